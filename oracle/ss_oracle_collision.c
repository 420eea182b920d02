/* CPU oracle: collision detection (row S1c of SURVEY.md §8(a)) -- TEST INFRASTRUCTURE ONLY.
 *
 * Candidate pairs come pre-filtered and pre-ordered from the model compiler
 * (stretch_mujoco_b200/compiler.py:_collision_pairs); this file does the run-time bounding
 * sphere test and the narrowphase.  Plane-vs-primitive routines restate MuJoCo's analytic
 * functions [upstream mjc_PlaneSphere/PlaneCylinder/PlaneBox]; every other convex pair runs
 * Minkowski Portal Refinement, the algorithm of libccd's ccdMPRPenetration that MuJoCo 3.2.6
 * calls for convex pairs [upstream mjc_Convex], including the `multiccd` perturbation contacts the
 * reference enables (`stretch_mujoco/models/stretch.xml:8`).  Plane-mesh adds up to three hull
 * neighbours of the support vertex [upstream mjc_PlaneConvex]; box-box and sphere-box are analytic
 * (own SAT + face clipping restatement; MuJoCo's mjc_BoxBox point ordering is not reproduced).
 */
#include "ss_oracle.h"
#include "ss_oracle_internal.h"

/* contact frame from a unit normal [upstream mju_makeFrame] */
static void make_frame(double* f) {
  double* x = f; double* y = f + 3; double* z = f + 6;
  y[0] = 0; y[1] = 1; y[2] = 0;
  if (x[1] > 0.5 || x[1] < -0.5) { y[1] = 0; y[2] = 1; }
  double d = v3dot(x, y);
  v3addscl(y, y, x, -d);
  v3normalize(y);
  v3cross(z, x, y);
}

static om_contact* add_contact(const om_model* m, om_data* d, int pair, double dist, const double* pos,
                               const double* normal) {
  if (d->ncon >= OM_MAXCON) { d->flags |= 2; return NULL; }
  om_contact* c = &d->contact[d->ncon++];
  c->dist = dist; v3copy(c->pos, pos); v3copy(c->frame, normal);
  make_frame(c->frame);
  c->geom1 = m->pair_geom1[pair]; c->geom2 = m->pair_geom2[pair];
  c->dim = m->pair_condim[pair];
  memcpy(c->friction, m->pair_friction + 5 * pair, 40);
  memcpy(c->solref, m->pair_solref + 2 * pair, 16);
  memcpy(c->solimp, m->pair_solimp + 5 * pair, 40);
  c->includemargin = m->pair_margin[pair] - m->pair_gap[pair];
  c->mu = 0; c->efc_address = -1;
  return c;
}

/* ------------------------------------------------------------------------- support mapping */

typedef struct {
  int type; const double *pos, *mat, *size; const double* verts; int nvert;
  const int *edgeadr, *edge;   /* hull adjacency (NULL when the blob carries none) */
} cvx;

static int support(const cvx* g, const double* dir, double* out) {
  double l[3], r[3] = {0, 0, 0};
  int index = -1;
  multmatvec3(l, g->mat, dir);
  switch (g->type) {
    case GEOM_SPHERE: {
      double n = v3norm(l);
      if (n > OM_MINVAL) v3scl(r, l, g->size[0] / n);
    } break;
    case GEOM_BOX:
      for (int i = 0; i < 3; i++) r[i] = l[i] > 0 ? g->size[i] : -g->size[i];
      break;
    case GEOM_CYLINDER: {
      double n = sqrt(l[0] * l[0] + l[1] * l[1]);
      if (n > OM_MINVAL) { r[0] = l[0] / n * g->size[0]; r[1] = l[1] / n * g->size[0]; }
      r[2] = l[2] > 0 ? g->size[1] : -g->size[1];
    } break;
    case GEOM_MESH: {
      double best = -1e300; int bi = 0;
      for (int i = 0; i < g->nvert; i++) {
        double s = v3dot(g->verts + 3 * i, l);
        if (s > best) { best = s; bi = i; }
      }
      v3copy(r, g->verts + 3 * bi);
      index = bi;
    } break;
    default: break;
  }
  mulmatvec3(out, g->mat, r);
  v3addto(out, g->pos);
  return index;
}

typedef struct { double v[3], v1[3], v2[3]; } spt;

/* support of the Minkowski difference A - B */
static void msupport(const cvx* a, const cvx* b, const double* dir, spt* s) {
  double nd[3] = {-dir[0], -dir[1], -dir[2]};
  support(a, dir, s->v1);
  support(b, nd, s->v2);
  v3sub(s->v, s->v1, s->v2);
}

#define MPR_TOL 1e-6
#define MPR_MAXIT 50
#define MPR_EPS 1e-14

static void portal_dir(const spt* p, double* dir) {
  double a[3], b[3];
  v3sub(a, p[2].v, p[1].v); v3sub(b, p[3].v, p[1].v);
  v3cross(dir, a, b); v3normalize(dir);
}

static void expand_portal(spt* p, const spt* v4) {
  double v4v0[3];
  v3cross(v4v0, v4->v, p[0].v);
  if (v3dot(p[1].v, v4v0) > 0) {
    if (v3dot(p[2].v, v4v0) > 0) p[1] = *v4; else p[3] = *v4;
  } else {
    if (v3dot(p[3].v, v4v0) > 0) p[2] = *v4; else p[1] = *v4;
  }
}

static int reach_tolerance(const spt* p, const spt* v4, const double* dir) {
  double dv1 = v3dot(p[1].v, dir), dv2 = v3dot(p[2].v, dir), dv3 = v3dot(p[3].v, dir), dv4 = v3dot(v4->v, dir);
  double m1 = dv4 - dv1, m2 = dv4 - dv2, m3 = dv4 - dv3;
  double mn = m1 < m2 ? m1 : m2; mn = mn < m3 ? mn : m3;
  return mn <= MPR_TOL;
}

/* squared distance from the origin to triangle (a,b,c); closest point in w */
static double origin_tri_dist2(const double* a, const double* b, const double* c, double* w) {
  double ab[3], ac[3], ap[3];
  v3sub(ab, b, a); v3sub(ac, c, a); v3scl(ap, a, -1);
  double d1 = v3dot(ab, ap), d2 = v3dot(ac, ap);
  if (d1 <= 0 && d2 <= 0) { v3copy(w, a); return v3dot(w, w); }
  double bp[3]; v3scl(bp, b, -1);
  double d3 = v3dot(ab, bp), d4 = v3dot(ac, bp);
  if (d3 >= 0 && d4 <= d3) { v3copy(w, b); return v3dot(w, w); }
  double vc = d1 * d4 - d3 * d2;
  if (vc <= 0 && d1 >= 0 && d3 <= 0) { double t = d1 / (d1 - d3); v3addscl(w, a, ab, t); return v3dot(w, w); }
  double cp[3]; v3scl(cp, c, -1);
  double d5 = v3dot(ab, cp), d6 = v3dot(ac, cp);
  if (d6 >= 0 && d5 <= d6) { v3copy(w, c); return v3dot(w, w); }
  double vb = d5 * d2 - d1 * d6;
  if (vb <= 0 && d2 >= 0 && d6 <= 0) { double t = d2 / (d2 - d6); v3addscl(w, a, ac, t); return v3dot(w, w); }
  double va = d3 * d6 - d5 * d4;
  if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) {
    double t = (d4 - d3) / ((d4 - d3) + (d5 - d6)), bc[3];
    v3sub(bc, c, b); v3addscl(w, b, bc, t); return v3dot(w, w);
  }
  double den = 1.0 / (va + vb + vc), v = vb * den, u = vc * den;
  v3addscl(w, a, ab, v); v3addscl(w, w, ac, u);
  return v3dot(w, w);
}

static void find_pos(const spt* p, double* pos) {
  double dir[3], b[4], t[3], sum;
  portal_dir(p, dir);
  v3cross(t, p[1].v, p[2].v); b[0] = v3dot(t, p[3].v);
  v3cross(t, p[3].v, p[2].v); b[1] = v3dot(t, p[0].v);
  v3cross(t, p[0].v, p[1].v); b[2] = v3dot(t, p[3].v);
  v3cross(t, p[2].v, p[1].v); b[3] = v3dot(t, p[0].v);
  sum = b[0] + b[1] + b[2] + b[3];
  if (sum <= 0) {
    b[0] = 0;
    v3cross(t, p[2].v, p[3].v); b[1] = v3dot(t, dir);
    v3cross(t, p[3].v, p[1].v); b[2] = v3dot(t, dir);
    v3cross(t, p[1].v, p[2].v); b[3] = v3dot(t, dir);
    sum = b[1] + b[2] + b[3];
  }
  double inv = 1.0 / sum, p1[3] = {0, 0, 0}, p2[3] = {0, 0, 0};
  for (int i = 0; i < 4; i++) { v3addscl(p1, p1, p[i].v1, b[i]); v3addscl(p2, p2, p[i].v2, b[i]); }
  for (int k = 0; k < 3; k++) pos[k] = 0.5 * (p1[k] + p2[k]) * inv;
}

/* Minkowski Portal Refinement penetration query. Returns 1 when the shapes intersect and fills
 * depth (>=0), dir (unit, from A towards B) and pos. */
static int mpr_penetration(const cvx* A, const cvx* B, double* depth, double* dir_out, double* pos) {
  spt p[4], v4;
  double dir[3], va[3], vb[3];
  /* --- discover portal --- */
  v3sub(p[0].v, A->pos, B->pos); v3copy(p[0].v1, A->pos); v3copy(p[0].v2, B->pos);
  if (fabs(p[0].v[0]) < MPR_EPS && fabs(p[0].v[1]) < MPR_EPS && fabs(p[0].v[2]) < MPR_EPS) p[0].v[0] = 1e-5;
  v3scl(dir, p[0].v, -1); v3normalize(dir);
  msupport(A, B, dir, &p[1]);
  if (v3dot(p[1].v, dir) <= 0) return 0;
  v3cross(dir, p[0].v, p[1].v);
  if (v3dot(dir, dir) < MPR_EPS * MPR_EPS) {
    /* origin lies on the ray v0 -> v1 */
    if (v3dot(p[1].v, p[1].v) < MPR_EPS * MPR_EPS) {
      *depth = 0; v3zero(dir_out);
      for (int k = 0; k < 3; k++) pos[k] = 0.5 * (p[1].v1[k] + p[1].v2[k]);
      return 1;
    }
    *depth = v3norm(p[1].v); v3copy(dir_out, p[1].v); v3normalize(dir_out);
    for (int k = 0; k < 3; k++) pos[k] = 0.5 * (p[1].v1[k] + p[1].v2[k]);
    return 1;
  }
  v3normalize(dir);
  msupport(A, B, dir, &p[2]);
  if (v3dot(p[2].v, dir) <= 0) return 0;
  v3sub(va, p[1].v, p[0].v); v3sub(vb, p[2].v, p[0].v);
  v3cross(dir, va, vb); v3normalize(dir);
  if (v3dot(dir, p[0].v) > 0) { spt t = p[1]; p[1] = p[2]; p[2] = t; v3scl(dir, dir, -1); }
  for (int guard = 0;; guard++) {
    if (guard > 100) return 0;
    msupport(A, B, dir, &p[3]);
    if (v3dot(p[3].v, dir) <= 0) return 0;
    int cont = 0;
    v3cross(va, p[1].v, p[3].v);
    if (v3dot(va, p[0].v) < -MPR_EPS) { p[2] = p[3]; cont = 1; }
    if (!cont) {
      v3cross(va, p[3].v, p[2].v);
      if (v3dot(va, p[0].v) < -MPR_EPS) { p[1] = p[3]; cont = 1; }
    }
    if (!cont) break;
    v3sub(va, p[1].v, p[0].v); v3sub(vb, p[2].v, p[0].v);
    v3cross(dir, va, vb); v3normalize(dir);
  }
  /* --- refine portal until the origin is inside --- */
  for (int it = 0;; it++) {
    portal_dir(p, dir);
    if (v3dot(dir, p[1].v) >= 0) break; /* portal encloses the origin */
    msupport(A, B, dir, &v4);
    if (v3dot(v4.v, dir) < 0 || reach_tolerance(p, &v4, dir) || it > MPR_MAXIT) return 0;
    expand_portal(p, &v4);
  }
  /* --- find penetration --- */
  for (int it = 0;; it++) {
    portal_dir(p, dir);
    msupport(A, B, dir, &v4);
    if (reach_tolerance(p, &v4, dir) || it > MPR_MAXIT) {
      double w[3];
      double d2 = origin_tri_dist2(p[1].v, p[2].v, p[3].v, w);
      *depth = sqrt(d2);
      if (*depth < MPR_EPS) v3zero(dir_out);
      else { v3copy(dir_out, w); v3normalize(dir_out); }
      find_pos(p, pos);
      return 1;
    }
    expand_portal(p, &v4);
  }
}

/* ------------------------------------------------------------------------- narrowphase */

static void cvx_of(const om_model* m, const om_data* d, int g, cvx* c) {
  c->type = m->geom_type[g]; c->pos = d->geom_xpos + 3 * g; c->mat = d->geom_xmat + 9 * g;
  c->size = m->geom_size + 3 * g; c->verts = NULL; c->nvert = 0; c->edgeadr = NULL; c->edge = NULL;
  if (c->type == GEOM_MESH) {
    int mid = m->geom_dataid[g];
    c->verts = m->hull_vert + 3 * m->mesh_hulladr[mid]; c->nvert = m->mesh_hullnum[mid];
    if (m->hull_edgeadr) { c->edgeadr = m->hull_edgeadr + m->mesh_hulladr[mid]; c->edge = m->hull_edge; }
  }
}

static void plane_sphere(const om_model* m, om_data* d, int pair, int g1, int g2, double margin) {
  const double *pp = d->geom_xpos + 3 * g1, *pm = d->geom_xmat + 9 * g1, *c = d->geom_xpos + 3 * g2;
  double n[3] = {pm[2], pm[5], pm[8]}, dif[3], r = m->geom_size[3 * g2];
  v3sub(dif, c, pp);
  double dist = v3dot(dif, n) - r;
  if (dist > margin) return;
  double pos[3];
  v3addscl(pos, c, n, -(r + 0.5 * dist));
  add_contact(m, d, pair, dist, pos, n);
}

static void plane_cylinder(const om_model* m, om_data* d, int pair, int g1, int g2, double margin) {
  const double *pp = d->geom_xpos + 3 * g1, *pm = d->geom_xmat + 9 * g1, *c = d->geom_xpos + 3 * g2,
               *cm = d->geom_xmat + 9 * g2;
  double n[3] = {pm[2], pm[5], pm[8]}, axis[3] = {cm[2], cm[5], cm[8]}, dif[3], vec[3], pos[3];
  double r = m->geom_size[3 * g2], h = m->geom_size[3 * g2 + 1];
  double prjaxis = v3dot(n, axis);
  if (prjaxis > 0) { v3scl(axis, axis, -1); prjaxis = -prjaxis; }
  v3sub(dif, c, pp);
  double dist0 = v3dot(dif, n);
  v3scl(vec, axis, prjaxis); v3sub(vec, vec, n);
  double len = v3norm(vec);
  if (len < 1e-12) { vec[0] = cm[0] * r; vec[1] = cm[3] * r; vec[2] = cm[6] * r; }
  else v3scl(vec, vec, r / len);
  double prjvec = v3dot(vec, n);
  v3scl(axis, axis, h); prjaxis *= h;
  double dist = dist0 + prjaxis + prjvec;
  if (dist > margin) return;
  for (int k = 0; k < 3; k++) pos[k] = c[k] + vec[k] + axis[k] - n[k] * dist * 0.5;
  add_contact(m, d, pair, dist, pos, n);
  dist = dist0 - prjaxis + prjvec;
  if (dist <= margin) {
    for (int k = 0; k < 3; k++) pos[k] = c[k] + vec[k] - axis[k] - n[k] * dist * 0.5;
    add_contact(m, d, pair, dist, pos, n);
  }
  double prjvec1 = -prjvec * 0.5;
  dist = dist0 + prjaxis + prjvec1;
  if (dist <= margin) {
    double vec1[3];
    v3cross(vec1, vec, axis); v3normalize(vec1); v3scl(vec1, vec1, r * sqrt(3.0) * 0.5);
    for (int k = 0; k < 3; k++) pos[k] = c[k] + vec1[k] + axis[k] - vec[k] * 0.5 - n[k] * dist * 0.5;
    add_contact(m, d, pair, dist, pos, n);
    for (int k = 0; k < 3; k++) pos[k] = c[k] - vec1[k] + axis[k] - vec[k] * 0.5 - n[k] * dist * 0.5;
    add_contact(m, d, pair, dist, pos, n);
  }
}

static void plane_box(const om_model* m, om_data* d, int pair, int g1, int g2, double margin) {
  const double *pp = d->geom_xpos + 3 * g1, *pm = d->geom_xmat + 9 * g1, *c = d->geom_xpos + 3 * g2,
               *bm = d->geom_xmat + 9 * g2, *sz = m->geom_size + 3 * g2;
  double n[3] = {pm[2], pm[5], pm[8]}, dif[3];
  v3sub(dif, c, pp);
  double dist = v3dot(dif, n);
  int cnt = 0;
  for (int i = 0; i < 8; i++) {
    double l[3] = {(i & 1) ? sz[0] : -sz[0], (i & 2) ? sz[1] : -sz[1], (i & 4) ? sz[2] : -sz[2]}, vec[3], pos[3];
    mulmatvec3(vec, bm, l);
    double ldist = v3dot(n, vec);
    if (dist + ldist > margin || ldist > 0) continue;
    double cd = dist + ldist;
    for (int k = 0; k < 3; k++) pos[k] = c[k] + vec[k] - n[k] * cd * 0.5;
    add_contact(m, d, pair, cd, pos, n);
    if (++cnt >= 4) return;
  }
}

/* [upstream mjc_PlaneConvex] support vertex, then up to three of its hull neighbours within the margin */
static void plane_mesh(const om_model* m, om_data* d, int pair, int g1, int g2, double margin) {
  const double *pp = d->geom_xpos + 3 * g1, *pm = d->geom_xmat + 9 * g1;
  double n[3] = {pm[2], pm[5], pm[8]}, nd[3] = {-pm[2], -pm[5], -pm[8]}, s[3], dif[3], pos[3];
  cvx g;
  cvx_of(m, d, g2, &g);
  int v0 = support(&g, nd, s);
  v3sub(dif, s, pp);
  double dist = v3dot(dif, n);
  if (dist > margin) return;
  v3addscl(pos, s, n, -0.5 * dist);
  add_contact(m, d, pair, dist, pos, n);
  if (!g.edgeadr || v0 < 0) return;
  int cnt = 1;
  for (int e = g.edgeadr[v0]; e < g.edgeadr[v0 + 1] && cnt < 4; e++) {
    double w[3];
    mulmatvec3(w, g.mat, g.verts + 3 * g.edge[e]);
    v3addto(w, g.pos);
    v3sub(dif, w, pp);
    dist = v3dot(dif, n);
    if (dist > margin) continue;
    v3addscl(pos, w, n, -0.5 * dist);
    add_contact(m, d, pair, dist, pos, n);
    cnt++;
  }
}

/* one MPR query -> contact candidate [upstream mjc_MPRIteration] */
static int mpr_contact(const cvx* A, const cvx* B, double margin, double* dist, double* dir, double* pos) {
  double depth;
  if (!mpr_penetration(A, B, &depth, dir, pos)) return 0;
  if (v3dot(dir, dir) < 0.5) return 0; /* touching contact without a direction */
  *dist = margin - depth;
  return 1;
}

/* rotate the frame (mat, pos) about `origin` by rot [upstream mju_rotateFrame in engine_collision_convex.c] */
static void rotate_frame(const double* origin, const double* rot, double* mat, double* pos) {
  double t[9], rel[3], v[3];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) t[3 * r + c] = rot[3 * r] * mat[c] + rot[3 * r + 1] * mat[3 + c] + rot[3 * r + 2] * mat[6 + c];
  memcpy(mat, t, sizeof(t));
  v3sub(rel, pos, origin);
  mulmatvec3(v, rot, rel);
  v3add(pos, origin, v);
}

/* [upstream mjc_Convex] MPR contact, then -- with the multiccd flag the reference sets
 * (stretch_mujoco/models/stretch.xml:8) -- four more queries with both geoms counter-rotated by
 * +-1e-3 rad about the two tangent axes of the first contact; a new contact is kept when its position
 * is farther than 1e-3 * min(rbound) from every contact found so far. */
static void convex_convex(const om_model* m, om_data* d, int pair, int g1, int g2, double margin) {
  cvx A, B;
  double dist, dir[3], pos[3];
  cvx_of(m, d, g1, &A); cvx_of(m, d, g2, &B);
  if (!mpr_contact(&A, &B, margin, &dist, dir, pos)) return;
  if (!add_contact(m, d, pair, dist, pos, dir)) return;
  int t1 = A.type, t2 = B.type;
  if (!m->multiccd || t1 == GEOM_SPHERE || t1 == GEOM_ELLIPSOID || t2 == GEOM_SPHERE || t2 == GEOM_ELLIPSOID) return;
  double found[5][3], frame[9];
  int nfound = 1;
  v3copy(found[0], pos);
  v3copy(frame, dir); make_frame(frame);
  double tol = 1e-3 * (m->geom_rbound[g1] < m->geom_rbound[g2] ? m->geom_rbound[g1] : m->geom_rbound[g2]);
  for (int axis_id = 0; axis_id < 2; axis_id++)
    for (int angle_id = 0; angle_id < 2; angle_id++) {
      const double* axis = frame + 3 + 3 * axis_id;
      double angle = angle_id ? 1e-3 : -1e-3, q[4], rot[9], irot[9];
      axisangle2quat(q, axis, angle);
      quat2mat(rot, q);
      for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) irot[3 * r + c] = rot[3 * c + r];
      double pA[3], mA[9], pB[3], mB[9];
      v3copy(pA, A.pos); memcpy(mA, A.mat, 72); v3copy(pB, B.pos); memcpy(mB, B.mat, 72);
      rotate_frame(found[0], rot, mA, pA);
      rotate_frame(found[0], irot, mB, pB);
      cvx A2 = A, B2 = B;
      A2.pos = pA; A2.mat = mA; B2.pos = pB; B2.mat = mB;
      double dist2, dir2[3], pos2[3];
      if (!mpr_contact(&A2, &B2, margin, &dist2, dir2, pos2)) continue;
      int distinct = 1;
      for (int k = 0; k < nfound; k++) {
        double df[3]; v3sub(df, found[k], pos2);
        if (v3norm(df) < tol) distinct = 0;
      }
      if (!distinct) continue;
      if (!add_contact(m, d, pair, dist2, pos2, dir2)) return;
      v3copy(found[nfound++], pos2);
    }
}

/* sphere-box [upstream mjc_SphereBox]: closest point of the box to the sphere centre */
static void sphere_box(const om_model* m, om_data* d, int pair, int g1, int g2, double margin) {
  const double *ps = d->geom_xpos + 3 * g1, *pb = d->geom_xpos + 3 * g2, *Rb = d->geom_xmat + 9 * g2, *sz = m->geom_size + 3 * g2;
  double r = m->geom_size[3 * g1], dif[3], c[3], q[3], dl[3];
  v3sub(dif, ps, pb);
  multmatvec3(c, Rb, dif);
  int inside = 1;
  for (int k = 0; k < 3; k++) {
    q[k] = c[k] < -sz[k] ? -sz[k] : (c[k] > sz[k] ? sz[k] : c[k]);
    if (q[k] != c[k]) inside = 0;
  }
  double nl[3] = {0, 0, 0}, dist;   /* nl: box -> sphere, box frame */
  if (!inside) {
    v3sub(dl, c, q);
    double len = v3norm(dl);
    dist = len - r;
    if (dist > margin) return;
    v3scl(nl, dl, 1.0 / len);
  } else {
    int best = 0; double bd = 1e300;
    for (int k = 0; k < 3; k++) { double t = sz[k] - fabs(c[k]); if (t < bd) { bd = t; best = k; } }
    nl[best] = c[best] >= 0 ? 1 : -1;
    q[best] = nl[best] * sz[best];
    dist = -bd - r;
  }
  double nw[3], qw[3], n[3], pos[3];
  mulmatvec3(nw, Rb, nl); mulmatvec3(qw, Rb, q); v3addto(qw, pb);
  v3scl(n, nw, -1);   /* geom1 (sphere) -> geom2 (box) */
  v3addscl(pos, qw, nw, 0.5 * dist);
  add_contact(m, d, pair, dist, pos, n);
}

/* clip a convex polygon (n <= 8 vertices, 2-D coordinates) against u <= lim (sign=+1) or u >= -lim (sign=-1) */
static int clip_axis(double (*poly)[2], int n, int axis, double sign, double lim) {
  double out[16][2];
  int no = 0;
  for (int i = 0; i < n; i++) {
    const double *a = poly[i], *b = poly[(i + 1) % n];
    double da = sign * a[axis] - lim, db = sign * b[axis] - lim;
    if (da <= 0) { out[no][0] = a[0]; out[no][1] = a[1]; no++; }
    if ((da < 0 && db > 0) || (da > 0 && db < 0)) {
      double t = da / (da - db);
      out[no][0] = a[0] + t * (b[0] - a[0]); out[no][1] = a[1] + t * (b[1] - a[1]); no++;
    }
  }
  if (no > 8) no = 8;
  memcpy(poly, out, sizeof(double) * 2 * no);
  return no;
}

/* box-box [upstream mjc_BoxBox]: separating-axis test over the 15 axes, then either the incident face
 * clipped against the reference face (up to 8 points) or one edge-edge point.  Restated as the
 * textbook algorithm; MuJoCo's own routine may order / merge the points differently. */
static void box_box(const om_model* m, om_data* d, int pair, int g1, int g2, double margin) {
  const double *p1 = d->geom_xpos + 3 * g1, *R1 = d->geom_xmat + 9 * g1, *s1 = m->geom_size + 3 * g1;
  const double *p2 = d->geom_xpos + 3 * g2, *R2 = d->geom_xmat + 9 * g2, *s2 = m->geom_size + 3 * g2;
  double a1[3][3], a2[3][3], pp[3];
  for (int i = 0; i < 3; i++) for (int k = 0; k < 3; k++) { a1[i][k] = R1[3 * k + i]; a2[i][k] = R2[3 * k + i]; }
  v3sub(pp, p2, p1);
  double bestf = -1e300, beste = -1e300, nf[3] = {0, 0, 0}, ne[3] = {0, 0, 0};
  int codef = -1, codee = -1;
  for (int w = 0; w < 2; w++)
    for (int i = 0; i < 3; i++) {
      const double* L = w ? a2[i] : a1[i];
      double t = v3dot(pp, L), ra = 0, rb = 0;
      for (int k = 0; k < 3; k++) { ra += s1[k] * fabs(v3dot(a1[k], L)); rb += s2[k] * fabs(v3dot(a2[k], L)); }
      double sep = fabs(t) - ra - rb;
      if (sep > margin) return;
      if (sep > bestf) { bestf = sep; codef = 3 * w + i; v3scl(nf, L, t >= 0 ? 1 : -1); }
    }
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double L[3];
      v3cross(L, a1[i], a2[j]);
      double len = v3norm(L);
      if (len < 1e-6) continue;
      v3scl(L, L, 1.0 / len);
      double t = v3dot(pp, L), ra = 0, rb = 0;
      for (int k = 0; k < 3; k++) { ra += s1[k] * fabs(v3dot(a1[k], L)); rb += s2[k] * fabs(v3dot(a2[k], L)); }
      double sep = fabs(t) - ra - rb;
      if (sep > margin) return;
      if (sep > beste) { beste = sep; codee = 3 * i + j; v3scl(ne, L, t >= 0 ? 1 : -1); }
    }
  if (codee >= 0 && beste > bestf + 0.05 * fabs(bestf) + 1e-9) {
    /* edge-edge: the closest points of the two supporting edges */
    int i = codee / 3, j = codee % 3;
    double c1[3], c2[3];
    v3copy(c1, p1); v3copy(c2, p2);
    for (int k = 0; k < 3; k++) {
      if (k != i) v3addscl(c1, c1, a1[k], (v3dot(ne, a1[k]) > 0 ? 1 : -1) * s1[k]);
      if (k != j) v3addscl(c2, c2, a2[k], (v3dot(ne, a2[k]) > 0 ? -1 : 1) * s2[k]);
    }
    double r[3], b = v3dot(a1[i], a2[j]), den = 1 - b * b;
    v3sub(r, c2, c1);
    double d1 = v3dot(r, a1[i]), d2 = v3dot(r, a2[j]);
    double s = den > 1e-12 ? (d1 - b * d2) / den : 0, t = den > 1e-12 ? (b * d1 - d2) / den : 0;
    double q1[3], q2[3], pos[3];
    v3addscl(q1, c1, a1[i], s); v3addscl(q2, c2, a2[j], t);
    for (int k = 0; k < 3; k++) pos[k] = 0.5 * (q1[k] + q2[k]);
    add_contact(m, d, pair, beste, pos, ne);
    return;
  }
  /* face contact: reference box owns the axis, incident face of the other box is clipped against it */
  int refis2 = codef >= 3, ax = codef % 3;
  const double *pr = refis2 ? p2 : p1, *sr = refis2 ? s2 : s1, *pc = refis2 ? p1 : p2, *sc = refis2 ? s1 : s2;
  double(*ar)[3] = refis2 ? a2 : a1;
  double(*ac)[3] = refis2 ? a1 : a2;
  double nr[3];
  v3scl(nr, nf, refis2 ? -1 : 1);   /* outward normal of the reference face, towards the incident box */
  int jc = 0; double bd = -1;
  for (int k = 0; k < 3; k++) { double t = fabs(v3dot(ac[k], nr)); if (t > bd) { bd = t; jc = k; } }
  double sgn = v3dot(ac[jc], nr) > 0 ? -1 : 1, fc[3];
  v3addscl(fc, pc, ac[jc], sgn * sc[jc]);
  int k1 = (jc + 1) % 3, k2 = (jc + 2) % 3, u = (ax + 1) % 3, v = (ax + 2) % 3;
  double poly[16][2], hgt[4], quad[4][3];
  for (int c = 0; c < 4; c++) {
    double e1 = (c == 0 || c == 3) ? -1 : 1, e2 = c < 2 ? -1 : 1;
    v3addscl(quad[c], fc, ac[k1], e1 * sc[k1]); v3addscl(quad[c], quad[c], ac[k2], e2 * sc[k2]);
    double rel[3]; v3sub(rel, quad[c], pr);
    poly[c][0] = v3dot(rel, ar[u]); poly[c][1] = v3dot(rel, ar[v]); hgt[c] = v3dot(rel, nr);
  }
  /* height over the reference plane is affine in the 2-D coordinates: h = h0 + gu*x + gv*y (least squares on the quad) */
  double e1[2] = {poly[1][0] - poly[0][0], poly[1][1] - poly[0][1]}, e2[2] = {poly[3][0] - poly[0][0], poly[3][1] - poly[0][1]};
  double dh1 = hgt[1] - hgt[0], dh2 = hgt[3] - hgt[0], det = e1[0] * e2[1] - e1[1] * e2[0];
  if (fabs(det) < 1e-14) return;
  double gu = (dh1 * e2[1] - dh2 * e1[1]) / det, gv = (e1[0] * dh2 - e2[0] * dh1) / det;
  double h0 = hgt[0] - gu * poly[0][0] - gv * poly[0][1];
  int n = 4;
  n = clip_axis(poly, n, 0, 1, sr[u]);
  if (n) n = clip_axis(poly, n, 0, -1, sr[u]);
  if (n) n = clip_axis(poly, n, 1, 1, sr[v]);
  if (n) n = clip_axis(poly, n, 1, -1, sr[v]);
  for (int c = 0; c < n; c++) {
    double h = h0 + gu * poly[c][0] + gv * poly[c][1], dist = h - sr[ax];
    if (dist > margin) continue;
    double pos[3];
    v3copy(pos, pr);
    v3addscl(pos, pos, ar[u], poly[c][0]); v3addscl(pos, pos, ar[v], poly[c][1]); v3addscl(pos, pos, nr, h - 0.5 * dist);
    if (!add_contact(m, d, pair, dist, pos, nf)) return;
  }
}

void om_collision(const om_model* m, om_data* d) {
  d->ncon = 0;
  for (int p = 0; p < m->npair; p++) {
    int g1 = m->pair_geom1[p], g2 = m->pair_geom2[p];
    double margin = m->pair_margin[p];
    int t1 = m->geom_type[g1], t2 = m->geom_type[g2];
    /* bounding-sphere / plane-sphere rejection [upstream mj_filterSphere] */
    if (t1 == GEOM_PLANE) {
      const double* pm = d->geom_xmat + 9 * g1;
      double n[3] = {pm[2], pm[5], pm[8]}, dif[3];
      v3sub(dif, d->geom_xpos + 3 * g2, d->geom_xpos + 3 * g1);
      if (v3dot(dif, n) > m->geom_rbound[g2] + margin) continue;
    } else {
      double dif[3], bound = m->geom_rbound[g1] + m->geom_rbound[g2] + margin;
      v3sub(dif, d->geom_xpos + 3 * g2, d->geom_xpos + 3 * g1);
      if (v3dot(dif, dif) > bound * bound) continue;
    }
    /* oriented-bounding-box rejection (6 face axes / lowest corner vs plane): conservative, so it
     * never changes the contact set; it only keeps far-apart meshes out of the MPR query */
    {
      const double *ab2 = m->geom_aabb + 6 * g2, *R2 = d->geom_xmat + 9 * g2;
      double c2w[3], dif[3];
      mulmatvec3(c2w, R2, ab2);
      v3sub(dif, d->geom_xpos + 3 * g2, d->geom_xpos + 3 * g1);
      if (t1 == GEOM_PLANE) {
        const double* pm = d->geom_xmat + 9 * g1;
        double n[3] = {pm[2], pm[5], pm[8]};
        double low = v3dot(dif, n) + v3dot(c2w, n);
        for (int k = 0; k < 3; k++) low -= fabs(R2[k] * n[0] + R2[3 + k] * n[1] + R2[6 + k] * n[2]) * ab2[3 + k];
        if (low > margin) continue;
      } else {
        const double *ab1 = m->geom_aabb + 6 * g1, *R1 = d->geom_xmat + 9 * g1;
        double c1w[3], t[3];
        int sep = 0;
        mulmatvec3(c1w, R1, ab1);
        for (int k = 0; k < 3; k++) t[k] = dif[k] + c2w[k] - c1w[k];
        for (int i = 0; i < 3 && !sep; i++) {
          double a1[3] = {R1[i], R1[3 + i], R1[6 + i]}, a2[3] = {R2[i], R2[3 + i], R2[6 + i]};
          double r1 = ab1[3 + i] + margin, r2 = ab2[3 + i] + margin;
          for (int k = 0; k < 3; k++) {
            r1 += fabs(R2[k] * a1[0] + R2[3 + k] * a1[1] + R2[6 + k] * a1[2]) * ab2[3 + k];
            r2 += fabs(R1[k] * a2[0] + R1[3 + k] * a2[1] + R1[6 + k] * a2[2]) * ab1[3 + k];
          }
          if (fabs(v3dot(t, a1)) > r1 || fabs(v3dot(t, a2)) > r2) sep = 1;
        }
        if (sep) continue;
      }
    }
    if (t1 == GEOM_PLANE) {
      if (t2 == GEOM_SPHERE) plane_sphere(m, d, p, g1, g2, margin);
      else if (t2 == GEOM_CYLINDER) plane_cylinder(m, d, p, g1, g2, margin);
      else if (t2 == GEOM_BOX) plane_box(m, d, p, g1, g2, margin);
      else if (t2 == GEOM_MESH) plane_mesh(m, d, p, g1, g2, margin);
    } else if (t1 == GEOM_BOX && t2 == GEOM_BOX) {
      box_box(m, d, p, g1, g2, margin);
    } else if (t1 == GEOM_SPHERE && t2 == GEOM_BOX) {
      sphere_box(m, d, p, g1, g2, margin);
    } else {
      convex_convex(m, d, p, g1, g2, margin);
    }
  }
}
