/* CPU oracle: ray casting + camera (rows S2, C3 of SURVEY.md §8(a)) -- TEST INFRASTRUCTURE ONLY.
 *
 * Restates [upstream] mj_ray as used by the <rangefinder> sensors the reference reads out at
 * `stretch_mujoco/mujoco_server_sensor_manager.py:77-83`, and the geometric part of
 * `mujoco.Renderer.render()` (`stretch_mujoco/mujoco_server_camera_manager.py:135-137`): a pinhole
 * camera looking down -z with +y up, planar depth along the optical axis, near/far =
 * vis.map.znear/zfar * stat.extent.  RGB uses a Blinn-Phong restatement of the fixed-function
 * lighting (headlight + scene lights, no shadows / reflections / MSAA): structural parity only.
 */
#include "ss_oracle.h"
#include "ss_oracle_internal.h"

#include <stdlib.h>

typedef struct { float lo[3], hi[3]; int left, right, first, count; } bvh_node;
typedef struct { bvh_node* nodes; int nnodes; int* tri; /* permuted triangle ids (mesh-local) */ } mesh_bvh;

struct om_rayaccel { mesh_bvh* mesh; };

static void tri_bounds(const om_model* m, int mid, int t, float* lo, float* hi, float* cen) {
  const int* f = m->rmesh_face + 3 * (m->rmesh_faceadr[mid] + t);
  const float* v = m->rmesh_vert + 3 * m->rmesh_vertadr[mid];
  for (int k = 0; k < 3; k++) {
    float a = v[3 * f[0] + k], b = v[3 * f[1] + k], c = v[3 * f[2] + k];
    lo[k] = fminf(a, fminf(b, c)); hi[k] = fmaxf(a, fmaxf(b, c)); cen[k] = (a + b + c) / 3.0f;
  }
}

typedef struct { const om_model* m; int mid; float* cen; mesh_bvh* B; } build_ctx;

static int build_node(build_ctx* c, int first, int count) {
  mesh_bvh* B = c->B;
  int id = B->nnodes++;
  bvh_node* n = &B->nodes[id];
  float clo[3] = {1e30f, 1e30f, 1e30f}, chi[3] = {-1e30f, -1e30f, -1e30f};
  for (int k = 0; k < 3; k++) { n->lo[k] = 1e30f; n->hi[k] = -1e30f; }
  for (int i = first; i < first + count; i++) {
    float lo[3], hi[3], cen[3];
    tri_bounds(c->m, c->mid, B->tri[i], lo, hi, cen);
    for (int k = 0; k < 3; k++) {
      n->lo[k] = fminf(n->lo[k], lo[k]); n->hi[k] = fmaxf(n->hi[k], hi[k]);
      clo[k] = fminf(clo[k], cen[k]); chi[k] = fmaxf(chi[k], cen[k]);
    }
  }
  n->first = first; n->count = count; n->left = n->right = -1;
  if (count <= 4) return id;
  int ax = 0;
  if (chi[1] - clo[1] > chi[ax] - clo[ax]) ax = 1;
  if (chi[2] - clo[2] > chi[ax] - clo[ax]) ax = 2;
  if (!(chi[ax] > clo[ax])) return id;
  /* quickselect the median centroid along ax */
  int lo = first, hi = first + count - 1, mid = first + count / 2;
  while (lo < hi) {
    float pivot = c->cen[3 * B->tri[(lo + hi) / 2] + ax];
    int i = lo, j = hi;
    while (i <= j) {
      while (c->cen[3 * B->tri[i] + ax] < pivot) i++;
      while (c->cen[3 * B->tri[j] + ax] > pivot) j--;
      if (i <= j) { int t = B->tri[i]; B->tri[i] = B->tri[j]; B->tri[j] = t; i++; j--; }
    }
    if (mid <= j) hi = j; else if (mid >= i) lo = i; else break;
  }
  int nl = mid - first;
  int l = build_node(c, first, nl);
  int r = build_node(c, mid, count - nl);
  B->nodes[id].left = l; B->nodes[id].right = r; B->nodes[id].count = 0;
  return id;
}

static struct om_rayaccel* get_accel(const om_model* m) {
  static struct om_rayaccel* cache_accel = NULL;
  static const om_model* cache_model = NULL;
  if (cache_model == m) return cache_accel;
  struct om_rayaccel* A = (struct om_rayaccel*)calloc(1, sizeof(*A));
  A->mesh = (mesh_bvh*)calloc(m->nmesh > 0 ? m->nmesh : 1, sizeof(mesh_bvh));
  for (int mid = 0; mid < m->nmesh; mid++) {
    int nf = m->rmesh_facenum[mid];
    if (nf <= 0 || m->rmesh_faceadr[mid] < 0) continue;
    mesh_bvh* B = &A->mesh[mid];
    B->nodes = (bvh_node*)malloc(sizeof(bvh_node) * (size_t)(2 * nf + 1));
    B->tri = (int*)malloc(sizeof(int) * (size_t)nf);
    float* cen = (float*)malloc(sizeof(float) * 3 * (size_t)nf);
    for (int t = 0; t < nf; t++) { float lo[3], hi[3]; B->tri[t] = t; tri_bounds(m, mid, t, lo, hi, cen + 3 * t); }
    build_ctx c = {m, mid, cen, B};
    build_node(&c, 0, nf);
    free(cen);
  }
  cache_accel = A; cache_model = m;
  return A;
}

/* ------------------------------------------------------------------------- ray / shape */

static double ray_quad(double a, double b, double c) {
  /* smallest non-negative root of a x^2 + 2 b x + c */
  double det = b * b - a * c;
  if (det < 0 || a < OM_MINVAL) return -1;
  det = sqrt(det);
  double x0 = (-b - det) / a, x1 = (-b + det) / a;
  if (x0 >= 0) return x0;
  if (x1 >= 0) return x1;
  return -1;
}

static double ray_tri(const float* v0, const float* v1, const float* v2, const double* o, const double* d, double* nrm) {
  double e1[3] = {v1[0] - v0[0], v1[1] - v0[1], v1[2] - v0[2]}, e2[3] = {v2[0] - v0[0], v2[1] - v0[1], v2[2] - v0[2]};
  double p[3], q[3], t[3] = {o[0] - v0[0], o[1] - v0[1], o[2] - v0[2]};
  v3cross(p, d, e2);
  double det = v3dot(e1, p);
  if (fabs(det) < 1e-30) return -1;
  double inv = 1.0 / det, u = v3dot(t, p) * inv;
  if (u < 0 || u > 1) return -1;
  v3cross(q, t, e1);
  double v = v3dot(d, q) * inv;
  if (v < 0 || u + v > 1) return -1;
  double x = v3dot(e2, q) * inv;
  if (x < 0) return -1;
  if (nrm) v3cross(nrm, e1, e2);
  return x;
}

static int ray_box_hit(const float* lo, const float* hi, const double* o, const double* inv, double tmax) {
  double t0 = 0, t1 = tmax;
  for (int k = 0; k < 3; k++) {
    double a = (lo[k] - o[k]) * inv[k], b = (hi[k] - o[k]) * inv[k];
    if (a > b) { double t = a; a = b; b = t; }
    if (a > t0) t0 = a;
    if (b < t1) t1 = b;
    if (t0 > t1) return 0;
  }
  return 1;
}

static double ray_mesh(const om_model* m, const struct om_rayaccel* A, int mid, const double* o, const double* d,
                       double tmin, double* nrm) {
  const mesh_bvh* B = &A->mesh[mid];
  if (!B->nodes) return -1;
  const float* V = m->rmesh_vert + 3 * m->rmesh_vertadr[mid];
  const int* F = m->rmesh_face + 3 * m->rmesh_faceadr[mid];
  double inv[3];
  for (int k = 0; k < 3; k++) inv[k] = 1.0 / (fabs(d[k]) > 1e-30 ? d[k] : (d[k] < 0 ? -1e-30 : 1e-30));
  double best = -1;
  int stack[64], sp = 0;
  stack[sp++] = 0;
  while (sp) {
    const bvh_node* n = &B->nodes[stack[--sp]];
    if (!ray_box_hit(n->lo, n->hi, o, inv, best >= 0 ? best : 1e30)) continue;
    if (n->left < 0) {
      for (int i = n->first; i < n->first + n->count; i++) {
        const int* f = F + 3 * B->tri[i];
        double tn[3] = {0, 0, 0};
        double x = ray_tri(V + 3 * f[0], V + 3 * f[1], V + 3 * f[2], o, d, tn);
        if (x >= tmin && (best < 0 || x < best)) { best = x; if (nrm) v3copy(nrm, tn); }
      }
    } else { stack[sp++] = n->left; stack[sp++] = n->right; }
  }
  return best;
}

/* nearest intersection of the local-frame ray with one geom; x >= tmin; nrm = local normal (unnormalised) */
static double ray_geom(const om_model* m, const struct om_rayaccel* A, int g, const double* o, const double* d,
                       double tmin, double* nrm) {
  const double* s = m->geom_size + 3 * g;
  double x = -1;
  switch (m->geom_type[g]) {
    case GEOM_PLANE:
      if (d[2] > -OM_MINVAL) return -1;
      x = -o[2] / d[2];
      if (x < tmin) return -1;
      {
        double px = o[0] + x * d[0], py = o[1] + x * d[1];
        if ((s[0] > 0 && fabs(px) > s[0]) || (s[1] > 0 && fabs(py) > s[1])) return -1;
      }
      if (nrm) { nrm[0] = 0; nrm[1] = 0; nrm[2] = 1; }
      return x;
    case GEOM_SPHERE: {
      double a = v3dot(d, d), b = v3dot(d, o), c = v3dot(o, o) - s[0] * s[0];
      x = ray_quad(a, b, c);
      if (x < tmin) {
        double det = b * b - a * c;
        if (det < 0) return -1;
        x = (-b + sqrt(det)) / a;
        if (x < tmin) return -1;
      }
      if (nrm) for (int k = 0; k < 3; k++) nrm[k] = o[k] + x * d[k];
      return x;
    }
    case GEOM_BOX: {
      double best = -1;
      for (int ax = 0; ax < 3; ax++) {
        if (fabs(d[ax]) < OM_MINVAL) continue;
        for (int sg = -1; sg <= 1; sg += 2) {
          double t = (sg * s[ax] - o[ax]) / d[ax];
          if (t < tmin) continue;
          int a1 = (ax + 1) % 3, a2 = (ax + 2) % 3;
          if (fabs(o[a1] + t * d[a1]) <= s[a1] && fabs(o[a2] + t * d[a2]) <= s[a2] && (best < 0 || t < best)) {
            best = t;
            if (nrm) { nrm[0] = nrm[1] = nrm[2] = 0; nrm[ax] = sg; }
          }
        }
      }
      return best;
    }
    case GEOM_CYLINDER: {
      double best = -1;
      double a = d[0] * d[0] + d[1] * d[1], b = d[0] * o[0] + d[1] * o[1], c = o[0] * o[0] + o[1] * o[1] - s[0] * s[0];
      double det = b * b - a * c;
      if (a > OM_MINVAL && det >= 0) {
        double sq = sqrt(det);
        for (int k = 0; k < 2; k++) {
          double t = (-b + (k ? sq : -sq)) / a;
          if (t >= tmin && fabs(o[2] + t * d[2]) <= s[1] && (best < 0 || t < best)) {
            best = t;
            if (nrm) { nrm[0] = o[0] + t * d[0]; nrm[1] = o[1] + t * d[1]; nrm[2] = 0; }
          }
        }
      }
      if (fabs(d[2]) > OM_MINVAL)
        for (int sg = -1; sg <= 1; sg += 2) {
          double t = (sg * s[1] - o[2]) / d[2];
          if (t < tmin) continue;
          double px = o[0] + t * d[0], py = o[1] + t * d[1];
          if (px * px + py * py <= s[0] * s[0] && (best < 0 || t < best)) {
            best = t;
            if (nrm) { nrm[0] = nrm[1] = 0; nrm[2] = sg; }
          }
        }
      return best;
    }
    case GEOM_MESH:
      return ray_mesh(m, A, m->geom_dataid[g], o, d, tmin, nrm);
    default:
      return -1;
  }
}

/* nearest hit over all ray-visible geoms. vec need not be unit length: distances are in units of |vec|. */
static double ray_scene(const om_model* m, const double* geom_xpos, const double* geom_xmat, const double* pnt,
                        const double* vec, int groupmask, int bodyexclude, double tmin, int* geomid, double* wnrm) {
  const struct om_rayaccel* A = get_accel(m);
  double best = -1;
  int bg = -1;
  double vv = v3dot(vec, vec);
  for (int k = 0; k < m->nraygeom; k++) {
    int g = m->raygeom_id[k];
    if (m->geom_bodyid[g] == bodyexclude) continue;
    if (groupmask && !((groupmask >> m->geom_group[g]) & 1)) continue;
    const double *gp = geom_xpos + 3 * g, *gm = geom_xmat + 9 * g;
    double dif[3], o[3], d[3], nrm[3];
    v3sub(dif, pnt, gp);
    if (m->geom_type[g] != GEOM_PLANE) {
      /* bounding-sphere rejection */
      double b = v3dot(vec, dif), c = v3dot(dif, dif) - m->geom_rbound[g] * m->geom_rbound[g];
      if (c > 0 && (b > 0 || b * b - vv * c < 0)) continue;
    }
    multmatvec3(o, gm, dif);
    multmatvec3(d, gm, vec);
    double x = ray_geom(m, A, g, o, d, tmin, wnrm ? nrm : NULL);
    if (x >= 0 && (best < 0 || x < best)) {
      best = x; bg = g;
      if (wnrm) mulmatvec3(wnrm, gm, nrm);
    }
  }
  if (geomid) *geomid = bg;
  return best;
}

double om_ray(const om_model* m, const double* xpos, const double* xmat, const double* geom_xpos,
              const double* geom_xmat, const double* pnt, const double* vec, int groupmask, int bodyexclude,
              int* geomid) {
  (void)xpos; (void)xmat;
  if (!m->has_ray) { if (geomid) *geomid = -1; return -1; }
  return ray_scene(m, geom_xpos, geom_xmat, pnt, vec, groupmask, bodyexclude, 0.0, geomid, NULL);
}

/* geom frames from body frames (xpos/xquat given by the caller, e.g. copied back from the GPU) */
static void geom_frames(const om_model* m, const double* xpos, const double* xquat, double* gx, double* gm) {
  for (int g = 0; g < m->ngeom; g++) {
    int b = m->geom_bodyid[g];
    double R[9], t[3], q[4];
    quat2mat(R, xquat + 4 * b);
    mulmatvec3(t, R, m->geom_pos + 3 * g); v3add(gx + 3 * g, xpos + 3 * b, t);
    quatmul(q, xquat + 4 * b, m->geom_quat + 4 * g); quatnormalize(q); quat2mat(gm + 9 * g, q);
  }
}

typedef struct {
  const double *xpos, *xquat, *origin, *dir;
  int nray, groupmask, bodyexclude;
  double* out_dist; int* out_geom;
} rays_ctx;

static void rays_env(const om_model* m, om_data* d, int e, void* vctx) {
  rays_ctx* c = (rays_ctx*)vctx;
  geom_frames(m, c->xpos + (size_t)e * m->nbody * 3, c->xquat + (size_t)e * m->nbody * 4, d->geom_xpos, d->geom_xmat);
  for (int r = 0; r < c->nray; r++) {
    size_t k = (size_t)e * c->nray + r;
    int g;
    c->out_dist[k] = ray_scene(m, d->geom_xpos, d->geom_xmat, c->origin + 3 * k, c->dir + 3 * k, c->groupmask,
                               c->bodyexclude, 0.0, &g, NULL);
    if (c->out_geom) c->out_geom[k] = g;
  }
}

int om_batch_rays(const om_model* m, int nenv, const double* xpos, const double* xquat, int nray,
                  const double* origin, const double* dir, int groupmask, int bodyexclude, double* out_dist,
                  int* out_geom, int nthreads) {
  if (!m->has_ray) return -1;
  get_accel(m); /* build once, before the worker threads start */
  rays_ctx c = {xpos, xquat, origin, dir, nray, groupmask, bodyexclude, out_dist, out_geom};
  om_parallel_for(m, nenv, nthreads, rays_env, &c);
  return 0;
}

typedef struct {
  const double *xpos, *xquat;
  int cam, W, H;
  double fovy;
  unsigned char* rgb; float* depth;
} render_ctx;

/* bilinear sample of 2-D texture t at (s, t) with GL_REPEAT wrapping; texel centres at (i + 0.5) / w, image row 0 on top
   (t = 1) [upstream: mjr uploads the PNG rows bottom-up] */
static void tex_sample(const om_model* m, int t, double s, double tt, double* rgb) {
  int w = m->tex_w[t], h = m->tex_h[t];
  const unsigned char* img = m->tex_rgb + m->tex_adr[t];
  double x = (s - floor(s)) * w - 0.5, y = (1.0 - (tt - floor(tt))) * h - 0.5;
  int x0 = (int)floor(x), y0 = (int)floor(y);
  double fx = x - x0, fy = y - y0;
  int xa = ((x0 % w) + w) % w, xb = (xa + 1) % w, ya = ((y0 % h) + h) % h, yb = (ya + 1) % h;
  for (int k = 0; k < 3; k++) {
    double c00 = img[3 * (ya * w + xa) + k], c10 = img[3 * (ya * w + xb) + k], c01 = img[3 * (yb * w + xa) + k], c11 = img[3 * (yb * w + xb) + k];
    rgb[k] = ((c00 * (1 - fx) + c10 * fx) * (1 - fy) + (c01 * (1 - fx) + c11 * fx) * fy) / 255.0;
  }
}

static void shade_pixel(const om_model* m, int g, const double* pos, const double* nrm_in, const double* eye,
                        const double* fwd, const double* xpos, const double* xquat, const double* gpos, const double* gmat,
                        unsigned char* out) {
  double sh[8];
  for (int k = 0; k < 8; k++) sh[k] = m->geom_shade[8 * g + k];
  if (m->geom_tex && m->geom_tex[4 * g] >= 0) {
    /* planar x-y projection of the geom frame; texrepeat per unit length (texuniform) or per geom extent */
    const double* tx = m->geom_tex + 4 * g;
    double d[3] = {pos[0] - gpos[0], pos[1] - gpos[1], pos[2] - gpos[2]};
    double lx = gmat[0] * d[0] + gmat[3] * d[1] + gmat[6] * d[2], ly = gmat[1] * d[0] + gmat[4] * d[1] + gmat[7] * d[2];
    double sx = m->geom_size[3 * g], sy = m->geom_size[3 * g + 1], s = 0, t = 0, c[3];
    if (tx[3] == 2) {
      /* the mesh's UV set: the triangle that holds the hit point (searched over the whole mesh -- textured meshes are
         stickers of < 100 triangles), barycentric interpolation of its three UV pairs */
      int mid = m->geom_dataid[g], nf = m->rmesh_facenum[mid];
      const float* V = m->rmesh_vert + 3 * m->rmesh_vertadr[mid];
      const int* F = m->rmesh_face + 3 * m->rmesh_faceadr[mid];
      const float* UV = m->rmesh_uv ? m->rmesh_uv + 6 * (size_t)m->rmesh_faceadr[mid] : NULL;
      double lz = gmat[2] * d[0] + gmat[5] * d[1] + gmat[8] * d[2], p[3] = {lx, ly, lz}, bestd = 1e30;
      for (int f = 0; f < nf && UV; f++) {
        const float *v0 = V + 3 * F[3 * f], *v1 = V + 3 * F[3 * f + 1], *v2 = V + 3 * F[3 * f + 2];
        double e1[3] = {v1[0] - v0[0], v1[1] - v0[1], v1[2] - v0[2]}, e2[3] = {v2[0] - v0[0], v2[1] - v0[1], v2[2] - v0[2]};
        double w[3] = {p[0] - v0[0], p[1] - v0[1], p[2] - v0[2]}, nn[3];
        double d00 = v3dot(e1, e1), d01 = v3dot(e1, e2), d11 = v3dot(e2, e2), d20 = v3dot(w, e1), d21 = v3dot(w, e2);
        double den = d00 * d11 - d01 * d01;
        if (fabs(den) < 1e-30) continue;
        double bu = (d11 * d20 - d01 * d21) / den, bv = (d00 * d21 - d01 * d20) / den;
        if (bu < -1e-6 || bv < -1e-6 || bu + bv > 1 + 1e-6) continue;
        v3cross(nn, e1, e2);
        double dist = fabs(v3dot(w, nn)) / sqrt(v3dot(nn, nn));
        if (dist < bestd) {
          bestd = dist;
          const float* uv = UV + 6 * f;
          s = (uv[0] + bu * (uv[2] - uv[0]) + bv * (uv[4] - uv[0])) * tx[1];
          t = (uv[1] + bu * (uv[3] - uv[1]) + bv * (uv[5] - uv[1])) * tx[2];
        }
      }
    }
    else if (tx[3] != 0 || sx <= 0 || sy <= 0) { s = lx * tx[1]; t = ly * tx[2]; }
    else { s = (lx / (2 * sx) + 0.5) * tx[1]; t = (ly / (2 * sy) + 0.5) * tx[2]; }
    tex_sample(m, (int)tx[0], s, t, c);
    for (int k = 0; k < 3; k++) sh[k] *= c[k];
  }
  double n[3] = {nrm_in[0], nrm_in[1], nrm_in[2]}, v[3], col[3];
  v3normalize(n);
  v3sub(v, eye, pos); v3normalize(v);
  if (v3dot(n, v) < 0) v3scl(n, n, -1); /* two-sided lighting */
  for (int k = 0; k < 3; k++) col[k] = sh[k] * sh[6];
  double shininess = fmax(sh[5] * 128.0, 1.0);
  for (int l = -1; l < m->nlight; l++) {
    double L[3], amb[3], dif[3], spc[3];
    if (l < 0) {
      if (!m->headlight_active) continue;
      v3scl(L, fwd, -1);
      for (int k = 0; k < 3; k++) { amb[k] = m->vis_headlight[k]; dif[k] = m->vis_headlight[3 + k]; spc[k] = m->vis_headlight[6 + k]; }
    } else {
      int b = m->light_bodyid[l];
      double R[9], t[3];
      quat2mat(R, xquat + 4 * b);
      if (m->light_directional[l]) { mulmatvec3(t, R, m->light_dir + 3 * l); v3scl(L, t, -1); }
      else { double lp[3]; mulmatvec3(t, R, m->light_pos + 3 * l); v3add(lp, xpos + 3 * b, t); v3sub(L, lp, pos); }
      v3normalize(L);
      for (int k = 0; k < 3; k++) { amb[k] = m->light_ambient[3 * l + k]; dif[k] = m->light_diffuse[3 * l + k]; spc[k] = m->light_specular[3 * l + k]; }
    }
    double nl = fmax(v3dot(n, L), 0.0), hs = 0;
    if (nl > 0) {
      double h[3];
      v3add(h, L, v); v3normalize(h);
      hs = pow(fmax(v3dot(n, h), 0.0), shininess);
    }
    for (int k = 0; k < 3; k++) col[k] += sh[k] * (amb[k] + dif[k] * nl) + sh[4] * spc[k] * hs;
  }
  for (int k = 0; k < 3; k++) out[k] = (unsigned char)(fmin(fmax(col[k], 0.0), 1.0) * 255.0 + 0.5);
}

static void render_env(const om_model* m, om_data* d, int e, void* vctx) {
  render_ctx* c = (render_ctx*)vctx;
  const double *xpos = c->xpos + (size_t)e * m->nbody * 3, *xquat = c->xquat + (size_t)e * m->nbody * 4;
  geom_frames(m, xpos, xquat, d->geom_xpos, d->geom_xmat);
  int cam = c->cam, b = m->cam_bodyid[cam], W = c->W, H = c->H;
  double Rb[9], Rc[9], q[4], t[3], eye[3];
  quat2mat(Rb, xquat + 4 * b);
  mulmatvec3(t, Rb, m->cam_pos + 3 * cam); v3add(eye, xpos + 3 * b, t);
  quatmul(q, xquat + 4 * b, m->cam_quat + 4 * cam); quatnormalize(q); quat2mat(Rc, q);
  double f = 0.5 * H / tan(c->fovy * M_PI / 360.0);
  double znear = m->vis_map[0] * m->extent, zfar = m->vis_map[1] * m->extent;
  double fwd[3] = {-Rc[2], -Rc[5], -Rc[8]};
  for (int v = 0; v < H; v++)
    for (int u = 0; u < W; u++) {
      double dl[3] = {(u + 0.5 - 0.5 * W) / f, -(v + 0.5 - 0.5 * H) / f, -1.0}, dw[3], nrm[3];
      mulmatvec3(dw, Rc, dl);
      int g;
      double x = ray_scene(m, d->geom_xpos, d->geom_xmat, eye, dw, 0x7, -1, znear, &g, nrm);
      size_t k = ((size_t)e * H + v) * W + u;
      if (x < 0 || x > zfar) { x = zfar; g = -1; }
      if (c->depth) c->depth[k] = (float)x;
      if (c->rgb) {
        unsigned char* px = c->rgb + 3 * k;
        if (g < 0) {
          double dn[3] = {dw[0], dw[1], dw[2]};
          v3normalize(dn);
          double tt = 0.5 * (1.0 + dn[2]);
          for (int a = 0; a < 3; a++) {
            double s1 = m->nsky >= 2 ? m->skybox_rgb[a] : 0.0, s2 = m->nsky >= 2 ? m->skybox_rgb[3 + a] : 0.0;
            px[a] = (unsigned char)(fmin(fmax(tt * s1 + (1 - tt) * s2, 0.0), 1.0) * 255.0 + 0.5);
          }
        } else {
          double pos[3];
          v3addscl(pos, eye, dw, x);
          shade_pixel(m, g, pos, nrm, eye, fwd, xpos, xquat, d->geom_xpos + 3 * g, d->geom_xmat + 9 * g, px);
        }
      }
    }
}

int om_batch_render(const om_model* m, int nenv, const double* xpos, const double* xquat, int cam_id, int W, int H,
                    double fovy_deg, unsigned char* rgb, float* depth, int nthreads) {
  if (!m->has_ray || cam_id < 0 || cam_id >= m->ncam) return -1;
  get_accel(m);
  render_ctx c = {xpos, xquat, cam_id, W, H, fovy_deg, rgb, depth};
  om_parallel_for(m, nenv, nthreads, render_env, &c);
  return 0;
}
