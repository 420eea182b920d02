// Ray casting kernels of libstretchsim: 2-D spinning lidar (rows S2/L1) and pinhole RGB+depth
// camera (rows C1-C4 of SURVEY.md §8(a)).
//
// Replaces the <rangefinder> evaluation inside mj_step that the reference reads out at
// stretch_mujoco/mujoco_server_sensor_manager.py:77-83 and mujoco.Renderer.update_scene()+render()
// at stretch_mujoco/mujoco_server_camera_manager.py:135-137.
//
// Data layout: triangle soups live once in HBM/L2 in mesh-local coordinates (shared by all envs)
// behind one BVH per mesh; per env only the 12-float world transform of each ray-visible geom is
// produced (ray_prepare_kernel) and staged in shared memory by the tracing blocks.  The camera
// kernel culls geoms against each 16x16 pixel tile's frustum before tracing and writes RGB/depth
// rows contiguously (the HBM-bound part: W*H*7 algorithmic bytes per env-frame).
#include "host.h"

#define RCHUNK 1024   // triangles per raster work chunk
#define MAXPRIM 64    // primitive (non-mesh) geoms per env that the raster path ray-casts per pixel

#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>

#define CUDA_OK(x)                                                                      \
  do {                                                                                  \
    cudaError_t e_ = (x);                                                               \
    if (e_ != cudaSuccess) return ss_fail("%s: %s", #x, cudaGetErrorString(e_));        \
  } while (0)

// ----------------------------------------------------------------------------- host: BVH build
namespace {
// BVH2 with the CHILDREN's boxes stored in the parent (one 64-byte node per visit, no re-test at pop):
//   n0 = (L.lo.xyz, L.hi.x)  n1 = (L.hi.yz, R.lo.xy)  n2 = (R.lo.z, R.hi.xyz)  n3 = (refL, refR, cntL, cntR) as ints
// ref >= 0: internal node index; ref < 0: leaf, first triangle = ~ref (global, leaf order), cnt triangles.
struct HostBVH {
  std::vector<float4> nodes;  // 4 per internal node
  std::vector<float4> tris;   // 3 per triangle (v0, e1, e2), leaf order
};
struct Box { float lo[3], hi[3]; };
static inline void box_init(Box& b) { for (int k = 0; k < 3; k++) { b.lo[k] = 1e30f; b.hi[k] = -1e30f; } }
static inline void box_add(Box& b, const Box& o) { for (int k = 0; k < 3; k++) { b.lo[k] = std::min(b.lo[k], o.lo[k]); b.hi[k] = std::max(b.hi[k], o.hi[k]); } }
static inline float box_area(const Box& b) {
  float d[3] = {b.hi[0] - b.lo[0], b.hi[1] - b.lo[1], b.hi[2] - b.lo[2]};
  if (d[0] < 0) return 0.f;
  return 2.f * (d[0] * d[1] + d[1] * d[2] + d[2] * d[0]);
}
static float as_float(int v) { float f; memcpy(&f, &v, 4); return f; }

struct Builder {
  const float* V; const int* F; std::vector<int> order; std::vector<float> cen; std::vector<Box> tb; HostBVH* out; int tri_base;
  // returns the child reference and its bounds / triangle count
  int build(int first, int count, Box& bounds, int& cnt) {
    Box cb; box_init(bounds); box_init(cb);
    for (int i = first; i < first + count; i++) {
      box_add(bounds, tb[order[i]]);
      for (int k = 0; k < 3; k++) { cb.lo[k] = std::min(cb.lo[k], cen[3 * order[i] + k]); cb.hi[k] = std::max(cb.hi[k], cen[3 * order[i] + k]); }
    }
    int ax = 0;
    if (cb.hi[1] - cb.lo[1] > cb.hi[ax] - cb.lo[ax]) ax = 1;
    if (cb.hi[2] - cb.lo[2] > cb.hi[ax] - cb.lo[ax]) ax = 2;
    if (count <= 4 || !(cb.hi[ax] > cb.lo[ax])) { cnt = count; return ~(tri_base + first); }
    // binned surface-area heuristic over the three axes (16 bins); median split as the fallback
    const int NB = 16;
    float best_cost = 1e30f; int best_ax = -1, best_bin = -1;
    for (int a = 0; a < 3; a++) {
      float ext = cb.hi[a] - cb.lo[a];
      if (!(ext > 0)) continue;
      Box bb[NB]; int bc[NB];
      for (int b = 0; b < NB; b++) { box_init(bb[b]); bc[b] = 0; }
      for (int i = first; i < first + count; i++) {
        int b = std::min(NB - 1, (int)(NB * (cen[3 * order[i] + a] - cb.lo[a]) / ext));
        box_add(bb[b], tb[order[i]]); bc[b]++;
      }
      float la[NB]; int lc[NB]; Box acc; box_init(acc); int n = 0;
      for (int b = 0; b < NB - 1; b++) { box_add(acc, bb[b]); n += bc[b]; la[b] = box_area(acc); lc[b] = n; }
      box_init(acc); n = 0;
      for (int b = NB - 1; b > 0; b--) {
        box_add(acc, bb[b]); n += bc[b];
        if (lc[b - 1] == 0 || n == 0) continue;
        float cost = la[b - 1] * lc[b - 1] + box_area(acc) * n;
        if (cost < best_cost) { best_cost = cost; best_ax = a; best_bin = b; }
      }
    }
    int mid;
    if (best_ax >= 0) {
      float ext = cb.hi[best_ax] - cb.lo[best_ax];
      auto it = std::partition(order.begin() + first, order.begin() + first + count, [&](int t) {
        return std::min(NB - 1, (int)(NB * (cen[3 * t + best_ax] - cb.lo[best_ax]) / ext)) < best_bin; });
      mid = (int)(it - order.begin());
    } else mid = first;
    if (mid <= first || mid >= first + count) {
      mid = first + count / 2;
      std::nth_element(order.begin() + first, order.begin() + mid, order.begin() + first + count,
                       [&](int a, int b) { return cen[3 * a + ax] < cen[3 * b + ax]; });
    }
    int id = (int)out->nodes.size() / 4;
    for (int k = 0; k < 4; k++) out->nodes.push_back(make_float4(0, 0, 0, 0));
    Box lb, rb; int lc2, rc2;
    int lref = build(first, mid - first, lb, lc2);
    int rref = build(mid, first + count - mid, rb, rc2);
    out->nodes[4 * id] = make_float4(lb.lo[0], lb.lo[1], lb.lo[2], lb.hi[0]);
    out->nodes[4 * id + 1] = make_float4(lb.hi[1], lb.hi[2], rb.lo[0], rb.lo[1]);
    out->nodes[4 * id + 2] = make_float4(rb.lo[2], rb.hi[0], rb.hi[1], rb.hi[2]);
    out->nodes[4 * id + 3] = make_float4(as_float(lref), as_float(rref), as_float(lc2), as_float(rc2));
    cnt = count;
    return id;
  }
};
}  // namespace

template <typename T>
static const T* upload(ss_model* M, const std::vector<T>& v) {
  void* p = nullptr;
  size_t n = std::max<size_t>(v.size(), 1) * sizeof(T);
  if (cudaMalloc(&p, n) != cudaSuccess) return nullptr;
  if (!v.empty()) cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
  M->dev_allocs.push_back(p);
  return (const T*)p;
}
static std::vector<float> f32(const ss_blob& b, const char* name) {
  const double* p = ss_blob_f64(&b, name);
  size_t n = ss_blob_count(&b, name);
  std::vector<float> v(n);
  for (size_t i = 0; i < n; i++) v[i] = (float)p[i];
  return v;
}
static std::vector<int> i32(const ss_blob& b, const char* name) {
  const int32_t* p = ss_blob_i32(&b, name);
  size_t n = ss_blob_count(&b, name);
  return std::vector<int>(p, p + n);
}

int ss_rays_model_init(ss_model* M) {
  RayModel& r = M->rm;
  memset(&r, 0, sizeof(r));
  const ss_blob& b = M->b;
  if (!ss_blob_find(&b, "raygeom_id")) return 0;  // physics-only blob
  std::vector<int> rg = i32(b, "raygeom_id"), gtype = i32(b, "geom_type"), gbody = i32(b, "geom_bodyid"),
                   gdata = i32(b, "geom_dataid"), ggroup = i32(b, "geom_group");
  std::vector<float> gpos = f32(b, "geom_pos"), gquat = f32(b, "geom_quat"), gsize = f32(b, "geom_size"),
                     grb = f32(b, "geom_rbound"), gshade = f32(b, "geom_shade");
  std::vector<int> t_type, t_body, t_mesh, t_group;
  std::vector<float> t_pos, t_quat, t_size, t_rb, t_shade;
  for (int g : rg) {
    t_type.push_back(gtype[g]); t_body.push_back(gbody[g]); t_mesh.push_back(gdata[g]); t_group.push_back(ggroup[g]);
    for (int k = 0; k < 3; k++) { t_pos.push_back(gpos[3 * g + k]); t_size.push_back(gsize[3 * g + k]); }
    for (int k = 0; k < 4; k++) t_quat.push_back(gquat[4 * g + k]);
    for (int k = 0; k < 8; k++) t_shade.push_back(gshade[8 * g + k]);
    t_rb.push_back(grb[g]);
  }
  // per-mesh BVH over the triangle soup
  std::vector<int> vadr = i32(b, "rmesh_vertadr"), fadr = i32(b, "rmesh_faceadr"), fnum = i32(b, "rmesh_facenum");
  const float* V = ss_blob_f32(&b, "rmesh_vert");
  const int32_t* F = ss_blob_i32(&b, "rmesh_face");
  HostBVH H;
  std::vector<float2> tri_uv;   // 3 per triangle (uv0, uv1 - uv0, uv2 - uv0), leaf order like H.tris
  const float* UVsrc = ss_blob_f32(&b, "rmesh_uv");
  std::vector<int> bvhadr(vadr.size(), 0), bvhcnt(vadr.size(), 0), tribase(vadr.size(), 0);
  std::vector<float4> mbox(2 * vadr.size(), make_float4(0, 0, 0, 0));
  for (size_t mid = 0; mid < vadr.size(); mid++) {
    if (fadr[mid] < 0 || fnum[mid] <= 0) continue;
    Builder B;
    B.V = V + 3 * (size_t)vadr[mid]; B.F = F + 3 * (size_t)fadr[mid]; B.out = &H; B.tri_base = (int)H.tris.size() / 3;
    int nf = fnum[mid];
    B.order.resize(nf); std::iota(B.order.begin(), B.order.end(), 0);
    B.cen.resize(3 * (size_t)nf); B.tb.resize(nf);
    for (int t = 0; t < nf; t++) {
      const int* f = B.F + 3 * t;
      for (int k = 0; k < 3; k++) {
        float a = B.V[3 * f[0] + k], b2 = B.V[3 * f[1] + k], c = B.V[3 * f[2] + k];
        B.cen[3 * t + k] = (a + b2 + c) / 3.0f;
        B.tb[t].lo[k] = std::min(a, std::min(b2, c)); B.tb[t].hi[k] = std::max(a, std::max(b2, c));
      }
    }
    Box bounds; int cnt = 0;
    bvhadr[mid] = B.build(0, nf, bounds, cnt);
    bvhcnt[mid] = cnt;
    tribase[mid] = B.tri_base;
    mbox[2 * mid] = make_float4(bounds.lo[0], bounds.lo[1], bounds.lo[2], 0);
    mbox[2 * mid + 1] = make_float4(bounds.hi[0], bounds.hi[1], bounds.hi[2], 0);
    for (int i = 0; i < nf; i++) {
      const int* f = B.F + 3 * B.order[i];
      const float *a = B.V + 3 * f[0], *c1 = B.V + 3 * f[1], *c2 = B.V + 3 * f[2];
      H.tris.push_back(make_float4(a[0], a[1], a[2], 0));
      H.tris.push_back(make_float4(c1[0] - a[0], c1[1] - a[1], c1[2] - a[2], 0));
      H.tris.push_back(make_float4(c2[0] - a[0], c2[1] - a[1], c2[2] - a[2], 0));
      if (UVsrc) {
        const float* uv = UVsrc + 6 * ((size_t)fadr[mid] + B.order[i]);
        tri_uv.push_back(make_float2(uv[0], uv[1])); tri_uv.push_back(make_float2(uv[2] - uv[0], uv[3] - uv[1]));
        tri_uv.push_back(make_float2(uv[4] - uv[0], uv[5] - uv[1]));
      }
    }
  }
  // per ray-geom record used by the tracing blocks: (type, mesh root ref, root triangle count, body | group << 16),
  // (rbound, size.xyz), local bounds lo / hi (for the mesh root test and the tile culling)
  std::vector<float4> rec(4 * rg.size());
  for (size_t k = 0; k < rg.size(); k++) {
    int type = t_type[k], mesh = t_mesh[k], ref = 0, cnt = 0;
    float lo[3], hi[3];
    for (int a = 0; a < 3; a++) { lo[a] = -t_size[3 * k + a]; hi[a] = t_size[3 * k + a]; }
    if (type == GEOM_SPHERE) for (int a = 0; a < 3; a++) { lo[a] = -t_size[3 * k]; hi[a] = t_size[3 * k]; }
    if (type == GEOM_CYLINDER) { lo[0] = lo[1] = -t_size[3 * k]; hi[0] = hi[1] = t_size[3 * k]; lo[2] = -t_size[3 * k + 1]; hi[2] = t_size[3 * k + 1]; }
    if (type == GEOM_MESH) {
      bool ok = mesh >= 0 && fnum[mesh] > 0;
      ref = ok ? bvhadr[mesh] : 0; cnt = ok ? bvhcnt[mesh] : -1;   // cnt < 0: nothing to hit
      for (int a = 0; a < 3; a++) { lo[a] = ok ? (&mbox[2 * mesh].x)[a] : 0.f; hi[a] = ok ? (&mbox[2 * mesh + 1].x)[a] : 0.f; }
    }
    rec[4 * k] = make_float4(as_float(type), as_float(ref), as_float(cnt), as_float(t_body[k] | (t_group[k] << 16)));
    rec[4 * k + 1] = make_float4(t_rb[k], t_size[3 * k], t_size[3 * k + 1], t_size[3 * k + 2]);
    rec[4 * k + 2] = make_float4(lo[0], lo[1], lo[2], 0);
    rec[4 * k + 3] = make_float4(hi[0], hi[1], hi[2], 0);
  }
  r.rg_rec = upload(M, rec);
  {
    std::vector<float4> rgtex(rg.size(), make_float4(-1.f, 1.f, 1.f, 0.f));
    std::vector<int4> tinfo;
    std::vector<unsigned char> texels;
    if (ss_blob_find(&b, "geom_tex") && ss_blob_count(&b, "tex_adr") > 0) {
      const double* gt = ss_blob_f64(&b, "geom_tex");
      const int32_t *ta = ss_blob_i32(&b, "tex_adr"), *tw = ss_blob_i32(&b, "tex_w"), *th = ss_blob_i32(&b, "tex_h");
      const unsigned char* tp = ss_blob_u8(&b, "tex_rgb");
      size_t nt = ss_blob_count(&b, "tex_adr"), nb = ss_blob_count(&b, "tex_rgb");
      if (gt && ta && tw && th && tp) {
        bool ok = true;
        for (size_t t = 0; t < nt; t++) {
          if (ta[t] < 0 || tw[t] <= 0 || th[t] <= 0 || (size_t)ta[t] + (size_t)tw[t] * th[t] * 3 > nb) ok = false;
          tinfo.push_back(make_int4(ta[t], tw[t], th[t], 0));
        }
        if (ok) {
          texels.assign(tp, tp + nb);
          for (size_t k = 0; k < rg.size(); k++) {
            const double* g4 = gt + 4 * (size_t)rg[k];
            if (g4[0] >= 0 && g4[0] < (double)nt) rgtex[k] = make_float4((float)g4[0], (float)g4[1], (float)g4[2], (float)g4[3]);
          }
        } else tinfo.clear();
      }
    }
    r.ntex = (int)tinfo.size();
    // the material record's (unused) alpha slot tells the shaders whether the geom is textured: no extra load for the rest
    for (size_t k = 0; k < rg.size(); k++) t_shade[8 * k + 3] = rgtex[k].x >= 0.f ? 1.f : 0.f;
    r.rg_tex = upload(M, rgtex); r.tex_info = upload(M, tinfo); r.tex_rgb = upload(M, texels);
    r.tri_uv = (UVsrc && tri_uv.size() == H.tris.size()) ? upload(M, tri_uv) : nullptr;
  }
  // raster work chunks of the camera-visible mesh geoms (groups 0..2)
  {
    std::vector<int4> chunks; std::vector<float4> cbox;
    for (size_t k = 0; k < rg.size(); k++) {
      int mesh = t_mesh[k];
      if (t_type[k] != GEOM_MESH || mesh < 0 || fnum[mesh] <= 0 || t_group[k] > 2) continue;
      for (int off = 0; off < fnum[mesh]; off += RCHUNK) {
        int first = tribase[mesh] + off, cnt = std::min(RCHUNK, fnum[mesh] - off);
        float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
        for (int t = first; t < first + cnt; t++) {
          const float4 &v0 = H.tris[3 * t], &e1 = H.tris[3 * t + 1], &e2 = H.tris[3 * t + 2];
          const float P[3][3] = {{v0.x, v0.y, v0.z}, {v0.x + e1.x, v0.y + e1.y, v0.z + e1.z}, {v0.x + e2.x, v0.y + e2.y, v0.z + e2.z}};
          for (int c = 0; c < 3; c++) for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], P[c][a]); hi[a] = std::max(hi[a], P[c][a]); }
        }
        chunks.push_back(make_int4((int)k, first, cnt, 0));
        cbox.push_back(make_float4(lo[0], lo[1], lo[2], 0)); cbox.push_back(make_float4(hi[0], hi[1], hi[2], 0));
      }
    }
    r.nchunk = (int)chunks.size();
    r.rchunk = upload(M, chunks); r.rchunk_box = upload(M, cbox);
    if (rg.size() >= 1024 || H.tris.size() / 3 >= (1u << 22)) r.nchunk = 0;   // key packing of the raster path: 10 bits geom, 22 bits triangle
  }
  r.present = 1;
  r.nraygeom = (int)rg.size(); r.nmesh = (int)vadr.size(); r.ngeom = M->dims.ngeom; r.nbody = M->dims.nbody;
  r.ncam = M->dims.ncam; r.nsite = M->dims.nsite;
  r.extent = (float)ss_blob_f64(&b, "stat_extent")[0];
  r.znear = (float)ss_blob_f64(&b, "vis_map")[0]; r.zfar = (float)ss_blob_f64(&b, "vis_map")[1];
  r.rg_geom = upload(M, rg); r.rg_type = upload(M, t_type); r.rg_body = upload(M, t_body); r.rg_mesh = upload(M, t_mesh);
  r.rg_group = upload(M, t_group); r.rg_pos = upload(M, t_pos); r.rg_quat = upload(M, t_quat); r.rg_size = upload(M, t_size);
  r.rg_rbound = upload(M, t_rb); r.rg_shade = upload(M, t_shade);
  r.rmesh_bvhadr = upload(M, bvhadr);
  r.tri = upload(M, H.tris); r.bvh = upload(M, H.nodes);
  r.cam_bodyid = upload(M, i32(b, "cam_bodyid")); r.cam_pos = upload(M, f32(b, "cam_pos")); r.cam_quat = upload(M, f32(b, "cam_quat"));
  M->cam_fovy_host = f32(b, "cam_fovy");
  r.cam_fovy = upload(M, M->cam_fovy_host);
  r.site_bodyid = upload(M, i32(b, "site_bodyid")); r.site_pos = upload(M, f32(b, "site_pos")); r.site_quat = upload(M, f32(b, "site_quat"));
  std::vector<int> stype = i32(b, "sensor_type"), sobj = i32(b, "sensor_objid"), sadr = i32(b, "sensor_adr"), rs, ra;
  std::vector<float> scut = f32(b, "sensor_cutoff"), rc;
  for (size_t s = 0; s < stype.size(); s++)
    if (stype[s] == SENS_RANGE) { rs.push_back(sobj[s]); ra.push_back(sadr[s]); rc.push_back(scut[s]); }
  r.nrange = (int)rs.size();
  r.range_site = upload(M, rs); r.range_adr = upload(M, ra); r.range_cutoff = upload(M, rc);
  std::vector<float> hl = f32(b, "vis_headlight"), sky = f32(b, "skybox_rgb");
  for (int k = 0; k < 9; k++) r.headlight[k] = hl[k];
  r.nsky = (int)sky.size() / 3;
  for (int k = 0; k < 6; k++) r.sky[k] = k < (int)sky.size() ? sky[k] : 0.f;
  r.headlight_active = ss_blob_i32(&b, "vis_headlight_active")[0];
  r.nlight = (int)ss_blob_count(&b, "light_bodyid");
  r.light_bodyid = upload(M, i32(b, "light_bodyid")); r.light_directional = upload(M, i32(b, "light_directional"));
  r.light_pos = upload(M, f32(b, "light_pos")); r.light_dir = upload(M, f32(b, "light_dir"));
  r.light_ambient = upload(M, f32(b, "light_ambient")); r.light_diffuse = upload(M, f32(b, "light_diffuse"));
  r.light_specular = upload(M, f32(b, "light_specular"));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return ss_fail("ray geometry upload failed: %s", cudaGetErrorString(e));
  return 0;
}

int ss_rays_set_fovy(ss_model* M, const double* fovy, size_t bytes) {
  if (!M->rm.present) return ss_fail("model has no ray geometry");
  if (bytes != M->cam_fovy_host.size() * sizeof(double)) return ss_fail("cam_fovy: expected %zu doubles", M->cam_fovy_host.size());
  for (size_t i = 0; i < M->cam_fovy_host.size(); i++) M->cam_fovy_host[i] = (float)fovy[i];
  CUDA_OK(cudaMemcpy((void*)M->rm.cam_fovy, M->cam_fovy_host.data(), M->cam_fovy_host.size() * sizeof(float), cudaMemcpyHostToDevice));
  return 0;
}

extern "C" int ss_model_num_rangefinders(const ss_model* M) { return M ? M->nrange : 0; }

// ----------------------------------------------------------------------------- device: tracing
__device__ __forceinline__ void q2m(float* R, const float* q) {
  float w = q[0], x = q[1], y = q[2], z = q[3];
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - w * z); R[2] = 2 * (x * z + w * y);
  R[3] = 2 * (x * y + w * z); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - w * x);
  R[6] = 2 * (x * z - w * y); R[7] = 2 * (y * z + w * x); R[8] = 1 - 2 * (x * x + y * y);
}
__device__ __forceinline__ void qmul(float* r, const float* a, const float* b) {
  float w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
        y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1], z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  float n = rsqrtf(w * w + x * x + y * y + z * z);
  r[0] = w * n; r[1] = x * n; r[2] = y * n; r[3] = z * n;
}

// world transform (pos[3], rot[9]) of every ray-visible geom of every env
__global__ void ray_prepare_kernel(RayModel r, int env_begin, int nenv, const float* __restrict__ xpos, const float* __restrict__ xquat,
                                   float* __restrict__ xf) {   // envs [env_begin, env_begin + nenv)
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nenv * r.nraygeom) return;
  i += env_begin * r.nraygeom;
  int e = i / r.nraygeom, k = i % r.nraygeom, b = r.rg_body[k];
  const float *bp = xpos + ((size_t)e * r.nbody + b) * 3, *bq = xquat + ((size_t)e * r.nbody + b) * 4;
  float Rb[9], q[4], gq[4] = {r.rg_quat[4 * k], r.rg_quat[4 * k + 1], r.rg_quat[4 * k + 2], r.rg_quat[4 * k + 3]};
  float bqq[4] = {bq[0], bq[1], bq[2], bq[3]};
  q2m(Rb, bqq);
  float lp[3] = {r.rg_pos[3 * k], r.rg_pos[3 * k + 1], r.rg_pos[3 * k + 2]};
  float* o = xf + (size_t)i * 12;
  o[0] = bp[0] + Rb[0] * lp[0] + Rb[1] * lp[1] + Rb[2] * lp[2];
  o[1] = bp[1] + Rb[3] * lp[0] + Rb[4] * lp[1] + Rb[5] * lp[2];
  o[2] = bp[2] + Rb[6] * lp[0] + Rb[7] * lp[1] + Rb[8] * lp[2];
  qmul(q, bqq, gq);
  q2m(o + 3, q);
}

struct Hit { float t; int k; float n[3]; int tri; };  // k = index into the ray-geom list; n = local-frame normal (unnormalised); tri = triangle (meshes)

// slab test of the ray (o, 1/d) against the box [lo, hi] on [0, tmax]; entry distance in *tnear
__device__ __forceinline__ bool slab(float lx, float ly, float lz, float hx, float hy, float hz, const float* o, const float* inv,
                                     float tmax, float* tnear) {
  float ax = (lx - o[0]) * inv[0], bx = (hx - o[0]) * inv[0];
  float ay = (ly - o[1]) * inv[1], by = (hy - o[1]) * inv[1];
  float az = (lz - o[2]) * inv[2], bz = (hz - o[2]) * inv[2];
  float t0 = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), 0.f));
  float t1 = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), tmax));
  *tnear = t0;
  return t0 <= t1;
}

// Nearest triangle hit below the node `ref` (>= 0 internal, < 0 leaf ~first with cnt triangles); the caller has
// already tested the mesh bounds.  One 64-byte node per visit (children boxes in the parent), nearer child
// first, the farther one is stacked with its entry distance and dropped at pop time once a closer hit exists.
__device__ float trace_mesh(const RayModel& r, int ref, int cnt, const float* o, const float* d, const float* inv, float tmin,
                            float tbest, float* nrm) {
  float best = tbest;  // only hits closer than tbest matter (<0: none yet)
  bool found = false;
  int stack[32], cstk[32], sp = 0;
  float tstk[32];
  int cur = ref, ccnt = cnt;
  while (true) {
    if (cur < 0) {
      int first = ~cur;
      for (int i = first; i < first + ccnt; i++) {
        float4 v0 = __ldg(r.tri + 3 * i), e1 = __ldg(r.tri + 3 * i + 1), e2 = __ldg(r.tri + 3 * i + 2);
        float p[3] = {d[1] * e2.z - d[2] * e2.y, d[2] * e2.x - d[0] * e2.z, d[0] * e2.y - d[1] * e2.x};
        float det = e1.x * p[0] + e1.y * p[1] + e1.z * p[2];
        if (fabsf(det) < 1e-30f) continue;
        float idet = 1.0f / det, t[3] = {o[0] - v0.x, o[1] - v0.y, o[2] - v0.z};
        float u = (t[0] * p[0] + t[1] * p[1] + t[2] * p[2]) * idet;
        if (u < 0 || u > 1) continue;
        float q[3] = {t[1] * e1.z - t[2] * e1.y, t[2] * e1.x - t[0] * e1.z, t[0] * e1.y - t[1] * e1.x};
        float v = (d[0] * q[0] + d[1] * q[1] + d[2] * q[2]) * idet;
        if (v < 0 || u + v > 1) continue;
        float x = (e2.x * q[0] + e2.y * q[1] + e2.z * q[2]) * idet;
        if (x >= tmin && (best < 0 || x < best)) {
          best = x; found = true;
          nrm[0] = e1.y * e2.z - e1.z * e2.y; nrm[1] = e1.z * e2.x - e1.x * e2.z; nrm[2] = e1.x * e2.y - e1.y * e2.x;
          nrm[3] = __int_as_float(i);
        }
      }
    } else {
      const float4* nd = r.bvh + 4 * (size_t)cur;
      float4 n0 = __ldg(nd), n1 = __ldg(nd + 1), n2 = __ldg(nd + 2), n3 = __ldg(nd + 3);
      float lim = best >= 0 ? best : 1e30f, tl, tr;
      bool hl = slab(n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, o, inv, lim, &tl);
      bool hr = slab(n1.z, n1.w, n2.x, n2.y, n2.z, n2.w, o, inv, lim, &tr);
      int rl = __float_as_int(n3.x), rr = __float_as_int(n3.y), cl = __float_as_int(n3.z), cr = __float_as_int(n3.w);
      if (hl && hr) {
        bool lfirst = tl <= tr;
        if (sp < 32) { stack[sp] = lfirst ? rr : rl; cstk[sp] = lfirst ? cr : cl; tstk[sp] = lfirst ? tr : tl; sp++; }
        cur = lfirst ? rl : rr; ccnt = lfirst ? cl : cr;
        continue;
      }
      if (hl) { cur = rl; ccnt = cl; continue; }
      if (hr) { cur = rr; ccnt = cr; continue; }
    }
    // pop the next stacked subtree that can still beat the best hit
    bool got = false;
    while (sp) {
      --sp;
      if (best >= 0 && tstk[sp] > best) continue;
      cur = stack[sp]; ccnt = cstk[sp]; got = true;
      break;
    }
    if (!got) break;
  }
  return found ? best : -1.f;
}

// per-geom record in shared memory: 4 float4 (see ss_rays_model_init)
#define REC_TYPE(rec) __float_as_int((rec)[0].x)
#define REC_REF(rec) __float_as_int((rec)[0].y)
#define REC_CNT(rec) __float_as_int((rec)[0].z)
#define REC_BODY(rec) (__float_as_int((rec)[0].w) & 0xffff)
#define REC_GROUP(rec) (__float_as_int((rec)[0].w) >> 16)

// nearest intersection with one geom in its local frame, x >= tmin and (tbest<0 or x<tbest)
__device__ float trace_geom(const RayModel& r, const float4* rec, const float* o, const float* d, float tmin, float tbest, float* nrm) {
  int type = REC_TYPE(rec);
  float s0 = rec[1].y, s1 = rec[1].z, s2 = rec[1].w;
  if (type == GEOM_MESH) {
    int cnt = REC_CNT(rec);
    if (cnt < 0) return -1.f;
    float inv[3], tn;
#pragma unroll
    for (int k = 0; k < 3; k++) inv[k] = 1.0f / (fabsf(d[k]) > 1e-20f ? d[k] : (d[k] < 0 ? -1e-20f : 1e-20f));
    if (!slab(rec[2].x, rec[2].y, rec[2].z, rec[3].x, rec[3].y, rec[3].z, o, inv, tbest >= 0 ? tbest : 1e30f, &tn)) return -1.f;
    return trace_mesh(r, REC_REF(rec), cnt, o, d, inv, tmin, tbest, nrm);
  }
  if (type == GEOM_PLANE) {
    if (d[2] > -1e-15f) return -1.f;
    float x = -o[2] / d[2];
    if (x < tmin) return -1.f;
    float px = o[0] + x * d[0], py = o[1] + x * d[1];
    if ((s0 > 0 && fabsf(px) > s0) || (s1 > 0 && fabsf(py) > s1)) return -1.f;
    nrm[0] = 0; nrm[1] = 0; nrm[2] = 1;
    return x;
  }
  if (type == GEOM_SPHERE) {
    float a = d[0] * d[0] + d[1] * d[1] + d[2] * d[2], b = d[0] * o[0] + d[1] * o[1] + d[2] * o[2];
    float c = o[0] * o[0] + o[1] * o[1] + o[2] * o[2] - s0 * s0, det = b * b - a * c;
    if (det < 0) return -1.f;
    float sq = sqrtf(det), x = (-b - sq) / a;
    if (x < tmin) x = (-b + sq) / a;
    if (x < tmin) return -1.f;
    nrm[0] = o[0] + x * d[0]; nrm[1] = o[1] + x * d[1]; nrm[2] = o[2] + x * d[2];
    return x;
  }
  if (type == GEOM_BOX) {
    float best = -1.f, sz[3] = {s0, s1, s2};
    for (int ax = 0; ax < 3; ax++) {
      if (fabsf(d[ax]) < 1e-15f) continue;
      int a1 = (ax + 1) % 3, a2 = (ax + 2) % 3;
      for (int sg = -1; sg <= 1; sg += 2) {
        float t = (sg * sz[ax] - o[ax]) / d[ax];
        if (t < tmin) continue;
        if (fabsf(o[a1] + t * d[a1]) <= sz[a1] && fabsf(o[a2] + t * d[a2]) <= sz[a2] && (best < 0 || t < best)) {
          best = t; nrm[0] = nrm[1] = nrm[2] = 0; nrm[ax] = (float)sg;
        }
      }
    }
    return best;
  }
  if (type == GEOM_CYLINDER) {
    float best = -1.f;
    float a = d[0] * d[0] + d[1] * d[1], b = d[0] * o[0] + d[1] * o[1], c = o[0] * o[0] + o[1] * o[1] - s0 * s0, det = b * b - a * c;
    if (a > 1e-15f && det >= 0) {
      float sq = sqrtf(det);
      for (int j = 0; j < 2; j++) {
        float t = (-b + (j ? sq : -sq)) / a;
        if (t >= tmin && fabsf(o[2] + t * d[2]) <= s1 && (best < 0 || t < best)) { best = t; nrm[0] = o[0] + t * d[0]; nrm[1] = o[1] + t * d[1]; nrm[2] = 0; }
      }
    }
    if (fabsf(d[2]) > 1e-15f)
      for (int sg = -1; sg <= 1; sg += 2) {
        float t = (sg * s1 - o[2]) / d[2];
        if (t < tmin) continue;
        float px = o[0] + t * d[0], py = o[1] + t * d[1];
        if (px * px + py * py <= s0 * s0 && (best < 0 || t < best)) { best = t; nrm[0] = nrm[1] = 0; nrm[2] = (float)sg; }
      }
    return best;
  }
  return -1.f;
}

// test one geom (index k into the ray-geom table) and keep the hit if it is nearer; xf / recs = this env's
// transforms and the geom records in shared memory.  FILTER: apply the group mask / body exclusion.
template <bool FILTER>
__device__ __forceinline__ void trace_one(const RayModel& r, const float* xf, const float4* recs, int k, const float* pnt, const float* vec,
                                          float vv, float tmin, int groupmask, int bodyexclude, Hit& h) {
  const float4* rec = recs + 4 * k;
  if (FILTER) {
    if (REC_BODY(rec) == bodyexclude) return;
    if (groupmask && !((groupmask >> REC_GROUP(rec)) & 1)) return;
  }
  const float* T = xf + 12 * k;
  float dif[3] = {pnt[0] - T[0], pnt[1] - T[1], pnt[2] - T[2]};
  if (REC_TYPE(rec) != GEOM_PLANE) {
    float rb = rec[1].x;
    float b = vec[0] * dif[0] + vec[1] * dif[1] + vec[2] * dif[2], c = dif[0] * dif[0] + dif[1] * dif[1] + dif[2] * dif[2] - rb * rb;
    if (c > 0 && (b > 0 || b * b - vv * c < 0)) return;
    if (h.t >= 0 && c > 0) {  // sphere entirely beyond the current best hit
      float tent = (-b - sqrtf(b * b - vv * c)) / vv;
      if (tent > h.t) return;
    }
  }
  const float* R = T + 3;
  float o[3] = {R[0] * dif[0] + R[3] * dif[1] + R[6] * dif[2], R[1] * dif[0] + R[4] * dif[1] + R[7] * dif[2], R[2] * dif[0] + R[5] * dif[1] + R[8] * dif[2]};
  float d[3] = {R[0] * vec[0] + R[3] * vec[1] + R[6] * vec[2], R[1] * vec[0] + R[4] * vec[1] + R[7] * vec[2], R[2] * vec[0] + R[5] * vec[1] + R[8] * vec[2]};
  float n[4];
  n[3] = __int_as_float(-1);
  float x = trace_geom(r, rec, o, d, tmin, h.t, n);
  if (x >= 0 && (h.t < 0 || x < h.t)) { h.t = x; h.k = k; h.n[0] = n[0]; h.n[1] = n[1]; h.n[2] = n[2]; h.tri = __float_as_int(n[3]); }
}
// nearest hit over all ray-visible geoms of the env (lidar / generic rays)
template <bool FILTER>
__device__ Hit trace_scene(const RayModel& r, const float* xf, const float4* recs, int n, const float* pnt, const float* vec, float tmin,
                           int groupmask, int bodyexclude) {
  Hit h; h.t = -1.f; h.k = -1; h.tri = -1; h.n[0] = h.n[1] = h.n[2] = 0;
  float vv = vec[0] * vec[0] + vec[1] * vec[1] + vec[2] * vec[2];
  for (int k = 0; k < n; k++) trace_one<FILTER>(r, xf, recs, k, pnt, vec, vv, tmin, groupmask, bodyexclude, h);
  return h;
}

// Does the geom (bounding sphere, then oriented bounding box) reach into the pyramid with apex at the eye and
// the four inward unit normals pn (camera frame: looks down -z, +y up)?  Conservative.
template <bool OBB>
__device__ __forceinline__ bool frustum_keeps(const float4* rec, const float* T, const float* eye, const float* cR, const float (*pn)[3],
                                              float znear) {
  const float* R = T + 3;
  float dw[3] = {T[0] - eye[0], T[1] - eye[1], T[2] - eye[2]};
  float c[3] = {cR[0] * dw[0] + cR[3] * dw[1] + cR[6] * dw[2], cR[1] * dw[0] + cR[4] * dw[1] + cR[7] * dw[2],
                cR[2] * dw[0] + cR[5] * dw[1] + cR[8] * dw[2]};
  float rb = rec[1].x;
  if (-c[2] < znear - rb) return false;
#pragma unroll
  for (int p = 0; p < 4; p++)
    if (pn[p][0] * c[0] + pn[p][1] * c[1] + pn[p][2] * c[2] < -rb) return false;
  if (!OBB) return true;
  // oriented box: centre and half extents in the geom frame, axes = columns of cR^T R
  float hc[3] = {0.5f * (rec[2].x + rec[3].x), 0.5f * (rec[2].y + rec[3].y), 0.5f * (rec[2].z + rec[3].z)};
  float hh[3] = {0.5f * (rec[3].x - rec[2].x), 0.5f * (rec[3].y - rec[2].y), 0.5f * (rec[3].z - rec[2].z)};
  float A[9];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) A[3 * i + j] = cR[i] * R[j] + cR[3 + i] * R[3 + j] + cR[6 + i] * R[6 + j];
  float bc[3];
#pragma unroll
  for (int i = 0; i < 3; i++) bc[i] = c[i] + A[3 * i] * hc[0] + A[3 * i + 1] * hc[1] + A[3 * i + 2] * hc[2];
#pragma unroll
  for (int p = 0; p < 4; p++) {
    float dist = pn[p][0] * bc[0] + pn[p][1] * bc[1] + pn[p][2] * bc[2], rad = 0.f;
#pragma unroll
    for (int j = 0; j < 3; j++) rad += fabsf(pn[p][0] * A[j] + pn[p][1] * A[3 + j] + pn[p][2] * A[6 + j]) * hh[j];
    if (dist < -rad * 1.0001f - 1e-6f) return false;
  }
  return true;
}
__device__ __forceinline__ void pyramid(float (*pn)[3], float x0, float x1, float y0, float y1) {   // y0 > y1
  float il = rsqrtf(1 + x0 * x0), ir = rsqrtf(1 + x1 * x1), it = rsqrtf(1 + y0 * y0), ib = rsqrtf(1 + y1 * y1);
  pn[0][0] = il; pn[0][1] = 0.f; pn[0][2] = x0 * il;
  pn[1][0] = -ir; pn[1][1] = 0.f; pn[1][2] = -x1 * ir;
  pn[2][0] = 0.f; pn[2][1] = -it; pn[2][2] = -y0 * it;
  pn[3][0] = 0.f; pn[3][1] = ib; pn[3][2] = y1 * ib;
}

// stage this env's geom transforms and the (env-independent) geom records in shared memory
__device__ __forceinline__ void stage_geoms(const RayModel& r, const float* xf, float4* srec, float* sxf, int tid, int nthr) {
  for (int i = tid; i < r.nraygeom * 4; i += nthr) srec[i] = __ldg(r.rg_rec + i);
  const float4* xf4 = reinterpret_cast<const float4*>(xf);
  float4* sxf4 = reinterpret_cast<float4*>(sxf);
  for (int i = tid; i < r.nraygeom * 3; i += nthr) sxf4[i] = xf4[i];
}
#define RAY_SMEM_BYTES(r) ((size_t)(r).nraygeom * (16 + 12) * sizeof(float))

// ----------------------------------------------------------------------------- lidar
// One block per env, one warp per bundle of 32 consecutive rangefinders.  The bundle's rays are enclosed in a
// cone (apex = the first ray's origin widened by the largest origin offset, axis = mean direction); geoms whose
// bounding sphere misses the cone are dropped once per bundle (one geom per lane, ballot compaction into the
// warp's list, geom-id order kept), then every lane traces its ray over the survivors.
#define LIDAR_WARPS 4
__global__ void __launch_bounds__(32 * LIDAR_WARPS) lidar_kernel(RayModel r, int nenv, int nsensordata, const float* __restrict__ xpos,
                                                                const float* __restrict__ xquat, const float* __restrict__ xf_all,
                                                                float* __restrict__ out, float* __restrict__ sensordata) {
  extern __shared__ float4 sm4[];
  float4* srec = sm4;
  float* sxf = reinterpret_cast<float*>(sm4 + 4 * r.nraygeom);
  int* wlist = reinterpret_cast<int*>(sxf + 12 * r.nraygeom) + (threadIdx.x >> 5) * r.nraygeom;   // [LIDAR_WARPS][nraygeom]
  const int e = blockIdx.x, lane = threadIdx.x & 31;
  stage_geoms(r, xf_all + (size_t)e * r.nraygeom * 12, srec, sxf, threadIdx.x, blockDim.x);
  __syncthreads();
  for (int base = (threadIdx.x >> 5) * 32; base < r.nrange; base += blockDim.x) {
    const int s = base + lane;
    const bool live = s < r.nrange;
    float p[3] = {0, 0, 0}, dir[3] = {0, 0, 0};
    int b = -1;
    if (live) {
      int site = r.range_site[s];
      b = r.site_bodyid[site];
      const float *bp = xpos + ((size_t)e * r.nbody + b) * 3, *bq = xquat + ((size_t)e * r.nbody + b) * 4;
      float Rb[9], q[4], bqq[4] = {bq[0], bq[1], bq[2], bq[3]}, R[9];
      float sq[4] = {r.site_quat[4 * site], r.site_quat[4 * site + 1], r.site_quat[4 * site + 2], r.site_quat[4 * site + 3]};
      float sp[3] = {r.site_pos[3 * site], r.site_pos[3 * site + 1], r.site_pos[3 * site + 2]};
      q2m(Rb, bqq);
      p[0] = bp[0] + Rb[0] * sp[0] + Rb[1] * sp[1] + Rb[2] * sp[2];
      p[1] = bp[1] + Rb[3] * sp[0] + Rb[4] * sp[1] + Rb[5] * sp[2];
      p[2] = bp[2] + Rb[6] * sp[0] + Rb[7] * sp[1] + Rb[8] * sp[2];
      qmul(q, bqq, sq);
      q2m(R, q);
      dir[0] = R[2]; dir[1] = R[5]; dir[2] = R[8];
    }
    // bundle cone: apex a (lane 0's origin), axis = normalised sum of the unit directions, cos of the half angle,
    // and the largest distance of an origin from the apex (0 for the Stretch lidar: all sites coincide)
    float ax[3], a0[3], cosa, spread;
    {
      float sx = dir[0], sy = dir[1], sz = dir[2];
      for (int o = 16; o > 0; o >>= 1) { sx += __shfl_xor_sync(0xffffffffu, sx, o); sy += __shfl_xor_sync(0xffffffffu, sy, o); sz += __shfl_xor_sync(0xffffffffu, sz, o); }
      float n = rsqrtf(fmaxf(sx * sx + sy * sy + sz * sz, 1e-20f));
      ax[0] = sx * n; ax[1] = sy * n; ax[2] = sz * n;
      for (int k = 0; k < 3; k++) a0[k] = __shfl_sync(0xffffffffu, p[k], 0);
      float c = live ? dir[0] * ax[0] + dir[1] * ax[1] + dir[2] * ax[2] : 1.f;
      float d2 = live ? (p[0] - a0[0]) * (p[0] - a0[0]) + (p[1] - a0[1]) * (p[1] - a0[1]) + (p[2] - a0[2]) * (p[2] - a0[2]) : 0.f;
      for (int o = 16; o > 0; o >>= 1) { c = fminf(c, __shfl_xor_sync(0xffffffffu, c, o)); d2 = fmaxf(d2, __shfl_xor_sync(0xffffffffu, d2, o)); }
      cosa = c; spread = sqrtf(d2);
    }
    const float sina = sqrtf(fmaxf(1.f - cosa * cosa, 0.f));
    const int bx = __shfl_sync(0xffffffffu, b, 0);
    const bool same_body = __all_sync(0xffffffffu, !live || b == bx);
    int nl = 0;
    for (int kb = 0; kb < r.nraygeom; kb += 32) {
      int k = kb + lane;
      bool keep = false;
      if (k < r.nraygeom) {
        const float4* rec = srec + 4 * k;
        keep = true;
        if (same_body && REC_BODY(rec) == bx) keep = false;            // the sensor's own body (mj_ray bodyexclude)
        else if (cosa > 0.f && REC_TYPE(rec) != GEOM_PLANE) {           // cone test only for bundles narrower than 180 degrees
          const float* T = sxf + 12 * k;
          float d[3] = {T[0] - a0[0], T[1] - a0[1], T[2] - a0[2]};
          float rb = rec[1].x + spread, dist2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
          if (dist2 > rb * rb) {
            float dist = sqrtf(dist2), sb = rb / dist, cb = sqrtf(1.f - sb * sb);
            float cang = (d[0] * ax[0] + d[1] * ax[1] + d[2] * ax[2]) / dist;   // cos of the angle centre-axis
            keep = cang >= cosa * cb - sina * sb - 1e-5f;                        // angle <= alpha + beta
          }
        }
      }
      unsigned mask = __ballot_sync(0xffffffffu, keep);
      if (keep) wlist[nl + __popc(mask & ((1u << lane) - 1))] = k;
      nl += __popc(mask);
    }
    __syncwarp();
    if (live) {
      Hit h; h.t = -1.f; h.k = -1; h.tri = -1; h.n[0] = h.n[1] = h.n[2] = 0;
      const float vv = dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2];
      for (int i = 0; i < nl; i++) trace_one<true>(r, sxf, srec, wlist[i], p, dir, vv, 0.f, 0, b, h);
      float dist = h.t;
      float cut = r.range_cutoff[s];
      if (dist >= 0 && cut > 0 && dist > cut) dist = cut;
      if (out) out[(size_t)e * r.nrange + s] = dist;
      if (sensordata) sensordata[(size_t)e * nsensordata + r.range_adr[s]] = dist;
    }
    __syncwarp();
  }
}

__global__ void rays_kernel(RayModel r, int nenv, int nray, const float* __restrict__ xf_all, const float* __restrict__ origin,
                            const float* __restrict__ dir, int groupmask, int bodyexclude, float* __restrict__ dist,
                            int32_t* __restrict__ geom) {
  extern __shared__ float4 sm4[];
  float4* srec = sm4;
  float* sxf = reinterpret_cast<float*>(sm4 + 4 * r.nraygeom);
  int e = blockIdx.x;
  stage_geoms(r, xf_all + (size_t)e * r.nraygeom * 12, srec, sxf, threadIdx.x, blockDim.x);
  __syncthreads();
  for (int s = threadIdx.x; s < nray; s += blockDim.x) {
    size_t k = (size_t)e * nray + s;
    float p[3] = {origin[3 * k], origin[3 * k + 1], origin[3 * k + 2]}, d[3] = {dir[3 * k], dir[3 * k + 1], dir[3 * k + 2]};
    Hit h = trace_scene<true>(r, sxf, srec, r.nraygeom, p, d, 0.f, groupmask, bodyexclude);
    dist[k] = h.t;
    if (geom) geom[k] = h.k >= 0 ? r.rg_geom[h.k] : -1;
  }
}

// ----------------------------------------------------------------------------- camera
#define TILE 16
#define TILES_PER_CTA 4   // horizontally adjacent tiles share one staging of the env's geoms
#define MAXLIGHT 8
// 2-D texture of ray-geom k at the world point pos (planar x-y projection of the geom frame, GL_REPEAT, bilinear, texel
// centres at (i + 0.5) / w, image row 0 on top); multiplies the material colour
__device__ __forceinline__ void texture_modulate(const RayModel& r, int k, int tri, const float* T, const float* pos, float* base) {
  const float4 tx = __ldg(r.rg_tex + k);
  if (tx.x < 0.f) return;
  const float* R = T + 3;
  const float d[3] = {pos[0] - T[0], pos[1] - T[1], pos[2] - T[2]};
  const float lx = R[0] * d[0] + R[3] * d[1] + R[6] * d[2], ly = R[1] * d[0] + R[4] * d[1] + R[7] * d[2];
  const float4 sz = __ldg(r.rg_rec + 4 * k + 1);   // (rbound, size.xyz)
  float s, t;
  if (tx.w == 2.f) {
    // the mesh's UV set: barycentric coordinates of the hit point in the winning triangle
    if (tri < 0 || !r.tri_uv) return;
    const float lz = R[2] * d[0] + R[5] * d[1] + R[8] * d[2];
    const float4 v0 = __ldg(r.tri + 3 * (size_t)tri), e1 = __ldg(r.tri + 3 * (size_t)tri + 1), e2 = __ldg(r.tri + 3 * (size_t)tri + 2);
    const float w[3] = {lx - v0.x, ly - v0.y, lz - v0.z};
    const float d00 = e1.x * e1.x + e1.y * e1.y + e1.z * e1.z, d01 = e1.x * e2.x + e1.y * e2.y + e1.z * e2.z, d11 = e2.x * e2.x + e2.y * e2.y + e2.z * e2.z;
    const float d20 = w[0] * e1.x + w[1] * e1.y + w[2] * e1.z, d21 = w[0] * e2.x + w[1] * e2.y + w[2] * e2.z;
    const float iden = 1.0f / (d00 * d11 - d01 * d01);
    const float bu = (d11 * d20 - d01 * d21) * iden, bv = (d00 * d21 - d01 * d20) * iden;
    const float2 uv0 = __ldg(r.tri_uv + 3 * (size_t)tri), du1 = __ldg(r.tri_uv + 3 * (size_t)tri + 1), du2 = __ldg(r.tri_uv + 3 * (size_t)tri + 2);
    s = (uv0.x + bu * du1.x + bv * du2.x) * tx.y; t = (uv0.y + bu * du1.y + bv * du2.y) * tx.z;
  }
  else if (tx.w != 0.f || sz.y <= 0.f || sz.z <= 0.f) { s = lx * tx.y; t = ly * tx.z; }
  else { s = (lx / (2.f * sz.y) + 0.5f) * tx.y; t = (ly / (2.f * sz.z) + 0.5f) * tx.z; }
  const int4 ti = __ldg(r.tex_info + (int)tx.x);
  const int w = ti.y, h = ti.z;
  const unsigned char* img = r.tex_rgb + ti.x;
  const float x = (s - floorf(s)) * w - 0.5f, y = (1.0f - (t - floorf(t))) * h - 0.5f;
  const int x0 = (int)floorf(x), y0 = (int)floorf(y);
  const float fx = x - x0, fy = y - y0;
  const int xa = ((x0 % w) + w) % w, xb = (xa + 1) % w, ya = ((y0 % h) + h) % h, yb = (ya + 1) % h;
#pragma unroll
  for (int a = 0; a < 3; a++) {
    const float c00 = img[3 * (ya * w + xa) + a], c10 = img[3 * (ya * w + xb) + a], c01 = img[3 * (yb * w + xa) + a], c11 = img[3 * (yb * w + xb) + a];
    base[a] *= ((c00 * (1 - fx) + c10 * fx) * (1 - fy) + (c01 * (1 - fx) + c11 * fx) * fy) * (1.0f / 255.0f);
  }
}

// Pixel epilogue shared by the ray-cast and the raster camera paths: far clip, depth limit, Blinn-Phong restatement
// of the fixed-function lighting, client-side post-processing of the reference fused in
// (status_stretch_camera.py:47-82).  xf = the env's geom transforms (12 floats per ray-geom, shared or global memory).
// post bits 1-2: 0 camera orientation, 1 = np.rot90(img, 1) (nav camera), 2 = np.rot90(img, -1) (d435i);
// bit 0: BGR channel order (cv2.COLOR_RGB2BGR).  Rotated images are [nenv, W, H(,3)].
__device__ __forceinline__ void shade_and_store(const RayModel& r, Hit h, const float* dw, const float* cam_eye, const float (*lvec)[4], const float (*lcol)[9],
                                                const float* xf, float zfar, int le, int u, int v, int W, int H,
                                                uint8_t* __restrict__ rgb, float* __restrict__ depth, float depth_limit, int post) {
  float x = h.t;
  if (x < 0 || x > zfar) { x = zfar; h.k = -1; }
  const int rot = (post >> 1) & 3;
  size_t pix = rot == 0 ? ((size_t)le * H + v) * W + u
             : rot == 1 ? ((size_t)le * W + (W - 1 - u)) * H + v
                        : ((size_t)le * W + u) * H + (H - 1 - v);
  if (depth) depth[pix] = (depth_limit > 0 && x > depth_limit) ? 0.f : x;   // utils.limit_depth_distance
  if (rgb) {
    float col[3];
    if (h.k < 0) {
      float inv = rsqrtf(dw[0] * dw[0] + dw[1] * dw[1] + dw[2] * dw[2]);
      float tt = 0.5f * (1.0f + dw[2] * inv);
      for (int a = 0; a < 3; a++) col[a] = r.nsky >= 2 ? tt * r.sky[a] + (1 - tt) * r.sky[3 + a] : 0.f;
    } else {
      const float* sh = r.rg_shade + 8 * h.k;
      const float* R = xf + 12 * h.k + 3;
      float n[3] = {R[0] * h.n[0] + R[1] * h.n[1] + R[2] * h.n[2], R[3] * h.n[0] + R[4] * h.n[1] + R[5] * h.n[2], R[6] * h.n[0] + R[7] * h.n[1] + R[8] * h.n[2]};
      float inv = rsqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
      n[0] *= inv; n[1] *= inv; n[2] *= inv;
      float pos[3] = {cam_eye[0] + x * dw[0], cam_eye[1] + x * dw[1], cam_eye[2] + x * dw[2]};
      float vw[3] = {-dw[0], -dw[1], -dw[2]};
      inv = rsqrtf(vw[0] * vw[0] + vw[1] * vw[1] + vw[2] * vw[2]);
      vw[0] *= inv; vw[1] *= inv; vw[2] *= inv;
      if (n[0] * vw[0] + n[1] * vw[1] + n[2] * vw[2] < 0) { n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2]; }
      float base[3] = {sh[0], sh[1], sh[2]};
      if (sh[3] > 0.f) texture_modulate(r, h.k, h.tri, xf + 12 * h.k, pos, base);
      for (int a = 0; a < 3; a++) col[a] = base[a] * sh[6];
      float shininess = fmaxf(sh[5] * 128.0f, 1.0f);
      int nl_ = min(r.nlight, MAXLIGHT - 1);
      for (int l = -1; l < nl_; l++) {
        float L[3], amb[3], dif[3], spc[3];
        const float* lv = lvec[l + 1];
        for (int a = 0; a < 3; a++) { amb[a] = lcol[l + 1][a]; dif[a] = lcol[l + 1][3 + a]; spc[a] = lcol[l + 1][6 + a]; }
        if (l < 0) {
          if (!r.headlight_active) continue;
          L[0] = lv[0]; L[1] = lv[1]; L[2] = lv[2];
        } else {
          if (lv[3] == 0.f) { L[0] = lv[0]; L[1] = lv[1]; L[2] = lv[2]; }
          else { L[0] = lv[0] - pos[0]; L[1] = lv[1] - pos[1]; L[2] = lv[2] - pos[2]; }
          float il = rsqrtf(L[0] * L[0] + L[1] * L[1] + L[2] * L[2]);
          L[0] *= il; L[1] *= il; L[2] *= il;
        }
        float nl = fmaxf(n[0] * L[0] + n[1] * L[1] + n[2] * L[2], 0.f), hs = 0.f;
        if (nl > 0) {
          float hv[3] = {L[0] + vw[0], L[1] + vw[1], L[2] + vw[2]};
          float ih = rsqrtf(hv[0] * hv[0] + hv[1] * hv[1] + hv[2] * hv[2]);
          hs = __powf(fmaxf((n[0] * hv[0] + n[1] * hv[1] + n[2] * hv[2]) * ih, 0.f), shininess);
        }
        for (int a = 0; a < 3; a++) col[a] += base[a] * (amb[a] + dif[a] * nl) + sh[4] * spc[a] * hs;
      }
    }
    uint8_t* px = rgb + 3 * pix;
    for (int a = 0; a < 3; a++) px[(post & 1) ? 2 - a : a] = (uint8_t)(fminf(fmaxf(col[a], 0.f), 1.f) * 255.0f + 0.5f);
  }
}

// camera pose of env e (thread-local): eye, rotation (columns = camera axes in the world)
__device__ __forceinline__ void camera_pose(const RayModel& r, int cam, const float* __restrict__ xpos, const float* __restrict__ xquat, int e,
                                            float* eye, float* cR) {
  int b = r.cam_bodyid[cam];
  const float *bp = xpos + ((size_t)e * r.nbody + b) * 3, *bq = xquat + ((size_t)e * r.nbody + b) * 4;
  float Rb[9], q[4], bqq[4] = {bq[0], bq[1], bq[2], bq[3]};
  float cq[4] = {r.cam_quat[4 * cam], r.cam_quat[4 * cam + 1], r.cam_quat[4 * cam + 2], r.cam_quat[4 * cam + 3]};
  float cp[3] = {r.cam_pos[3 * cam], r.cam_pos[3 * cam + 1], r.cam_pos[3 * cam + 2]};
  q2m(Rb, bqq);
  eye[0] = bp[0] + Rb[0] * cp[0] + Rb[1] * cp[1] + Rb[2] * cp[2];
  eye[1] = bp[1] + Rb[3] * cp[0] + Rb[4] * cp[1] + Rb[5] * cp[2];
  eye[2] = bp[2] + Rb[6] * cp[0] + Rb[7] * cp[1] + Rb[8] * cp[2];
  qmul(q, bqq, cq);
  q2m(cR, q);
}
// colours of light slot s (0 = headlight, s - 1 = scene light): ambient, diffuse, specular
__device__ __forceinline__ void light_colours(const RayModel& r, int s, float* c) {
  for (int a = 0; a < 3; a++) {
    c[a] = s == 0 ? r.headlight[a] : r.light_ambient[3 * (s - 1) + a];
    c[3 + a] = s == 0 ? r.headlight[3 + a] : r.light_diffuse[3 * (s - 1) + a];
    c[6 + a] = s == 0 ? r.headlight[6 + a] : r.light_specular[3 * (s - 1) + a];
  }
}
// world direction towards light l (w = 0) or its position (w = 1)
__device__ __forceinline__ void light_vector(const RayModel& r, int l, const float* __restrict__ xpos, const float* __restrict__ xquat, int e, float* L) {
  int b = r.light_bodyid[l];
  const float *bp = xpos + ((size_t)e * r.nbody + b) * 3, *bq = xquat + ((size_t)e * r.nbody + b) * 4;
  float Rb[9], bqq[4] = {bq[0], bq[1], bq[2], bq[3]};
  q2m(Rb, bqq);
  if (r.light_directional[l]) {
    const float* ld = r.light_dir + 3 * l;
    L[0] = -(Rb[0] * ld[0] + Rb[1] * ld[1] + Rb[2] * ld[2]); L[1] = -(Rb[3] * ld[0] + Rb[4] * ld[1] + Rb[5] * ld[2]);
    L[2] = -(Rb[6] * ld[0] + Rb[7] * ld[1] + Rb[8] * ld[2]); L[3] = 0.f;
  } else {
    const float* lp = r.light_pos + 3 * l;
    L[0] = bp[0] + Rb[0] * lp[0] + Rb[1] * lp[1] + Rb[2] * lp[2];
    L[1] = bp[1] + Rb[3] * lp[0] + Rb[4] * lp[1] + Rb[5] * lp[2];
    L[2] = bp[2] + Rb[6] * lp[0] + Rb[7] * lp[1] + Rb[8] * lp[2]; L[3] = 1.f;
  }
}

__global__ void __launch_bounds__(TILE * TILE) render_kernel(RayModel r, int env_begin, int cam, int W, int H, float fovy_deg,
                                                             const float* __restrict__ xpos, const float* __restrict__ xquat,
                                                             const float* __restrict__ xf_all, uint8_t* __restrict__ rgb,
                                                             float* __restrict__ depth, float depth_limit, int post) {
  extern __shared__ float4 sm4[];
  float4* srec = sm4;                                              // [nraygeom*4]
  float* sxf = reinterpret_cast<float*>(sm4 + 4 * r.nraygeom);     // [nraygeom*12]
  int* list = (int*)(sxf + r.nraygeom * 12);                       // [nraygeom]
  int* flag = list + r.nraygeom;                                   // [nraygeom]; doubles as the depth-sorted list
  float* skey = (float*)(flag + r.nraygeom);                       // [nraygeom] camera-space entry depth of the bounding sphere
  __shared__ int nlist;
  __shared__ float cam_eye[3], cam_R[9], focal;
  __shared__ float lvec[MAXLIGHT][4];   // world direction towards the light (w = 0) or its position (w = 1); slot 0 = headlight
  __shared__ float lcol[MAXLIGHT][9];
  int le = blockIdx.z, e = env_begin + le, tid = threadIdx.y * TILE + threadIdx.x;
  stage_geoms(r, xf_all + (size_t)e * r.nraygeom * 12, srec, sxf, tid, TILE * TILE);
  if (tid == 0) {
    camera_pose(r, cam, xpos, xquat, e, cam_eye, cam_R);
    focal = 0.5f * H / tanf(fovy_deg * 3.14159265358979f / 360.0f);
    lvec[0][0] = cam_R[2]; lvec[0][1] = cam_R[5]; lvec[0][2] = cam_R[8]; lvec[0][3] = 0.f;   // headlight: -forward = +z of the camera frame
  }
  if (rgb && tid >= 32 && tid < 32 + r.nlight && tid < 32 + MAXLIGHT - 1) light_vector(r, tid - 32, xpos, xquat, e, lvec[tid - 31]);
  if (rgb && tid >= 64 && tid < 64 + 1 + r.nlight && tid < 64 + MAXLIGHT) light_colours(r, tid - 64, lcol[tid - 64]);
  __syncthreads();
  float f = focal;
  float znear = r.znear * r.extent, zfar = r.zfar * r.extent;
  for (int tile = 0; tile < TILES_PER_CTA; tile++) {
  const int tx = blockIdx.x * TILES_PER_CTA + tile;
  if (tx * TILE >= W) break;
  // tile frustum culling: geoms whose bounding sphere, then oriented bounding box, misses the tile's pyramid are dropped
  const float ty0 = -((blockIdx.y * TILE) - 0.5f * H) / f;
  {
    float pn[4][3];
    pyramid(pn, (tx * TILE - 0.5f * W) / f, (fminf((tx + 1) * TILE, (float)W) - 0.5f * W) / f, ty0,
            -(fminf((blockIdx.y + 1) * TILE, (float)H) - 0.5f * H) / f);
    for (int k = tid; k < r.nraygeom; k += TILE * TILE) {
      const float4* rec = srec + 4 * k;
      bool keep = true;
      if (!((0x7 >> REC_GROUP(rec)) & 1)) keep = false;  // camera sees geom groups 0..2 (collision group 3 hidden)
      else if (REC_TYPE(rec) != GEOM_PLANE) keep = frustum_keeps<true>(rec, sxf + 12 * k, cam_eye, cam_R, pn, znear);
      float key = -1.f;   // planes first: cheap, and their hit bounds everything behind them
      if (keep && REC_TYPE(rec) != GEOM_PLANE) {
        const float* T = sxf + 12 * k;
        float dz = cam_R[2] * (T[0] - cam_eye[0]) + cam_R[5] * (T[1] - cam_eye[1]) + cam_R[8] * (T[2] - cam_eye[2]);
        key = fmaxf(-dz - rec[1].x, 0.f);
      }
      skey[k] = key;
      flag[k] = keep;
    }
  }
  __syncthreads();
  if (tid < 32) {  // ordered compaction by warp 0 (ballot + popc keeps the geom-id order deterministic)
    int n = 0;
    for (int base = 0; base < r.nraygeom; base += 32) {
      int k = base + tid;
      bool keep = k < r.nraygeom && flag[k];
      unsigned mask = __ballot_sync(0xffffffffu, keep);
      if (keep) list[n + __popc(mask & ((1u << tid) - 1))] = k;
      n += __popc(mask);
    }
    if (tid == 0) nlist = n;
  }
  __syncthreads();
  // front-to-back order (rank sort by entry depth, ties by geom id): near hits come first and prune the rest
  {
    const int nl = nlist;
    int mine = -1, rank = 0;
    if (tid < nl) {
      mine = list[tid];
      float km = skey[mine];
      for (int j = 0; j < nl; j++) { int o2 = list[j]; float kj = skey[o2]; rank += (kj < km) || (kj == km && o2 < mine); }
    }
    __syncthreads();
    if (mine >= 0) flag[rank] = mine;
    __syncthreads();
  }
  const int* slist = flag;
  int u = tx * TILE + threadIdx.x, v = blockIdx.y * TILE + threadIdx.y;
  const bool inb = u < W && v < H;
  float dl[3] = {(u + 0.5f - 0.5f * W) / f, -(v + 0.5f - 0.5f * H) / f, -1.0f};
  float dw[3] = {cam_R[0] * dl[0] + cam_R[1] * dl[1] + cam_R[2] * dl[2], cam_R[3] * dl[0] + cam_R[4] * dl[1] + cam_R[5] * dl[2],
                 cam_R[6] * dl[0] + cam_R[7] * dl[1] + cam_R[8] * dl[2]};
  Hit h; h.t = -1.f; h.k = -1; h.tri = -1; h.n[0] = h.n[1] = h.n[2] = 0;
  if (inb) {
    // (a second, per-warp culling level against 16 x 2 pixel strips was measured slower: 21.0 vs 19.7 ms)
    const float vv = dw[0] * dw[0] + dw[1] * dw[1] + dw[2] * dw[2];
    const int nl = nlist;
    for (int i = 0; i < nl; i++) trace_one<false>(r, sxf, srec, slist[i], cam_eye, dw, vv, znear, 0, -1, h);
  }
  if (inb) shade_and_store(r, h, dw, cam_eye, lvec, lcol, sxf, zfar, le, u, v, W, H, rgb, depth, depth_limit, post);
  __syncthreads();   // list / flag are rebuilt for the next tile
  }
}

// ----------------------------------------------------------------------------- raster camera path
// A primary-visibility ray caster pays ~1400 instructions per pixel on the robot's 394 k-triangle visual meshes
// (BVH traversal); most of those triangles are smaller than a pixel.  The raster path visits every TRIANGLE of the
// mesh chunks inside the image pyramid once instead: camera-space transform, pixel bounding box, the same
// Moeller-Trumbore test as the ray caster at the covered pixel centres, and a 64-bit atomicMin of
// (depth bits << 32 | geom << 22 | triangle) into a per-env depth/id buffer that stays L2-resident (envs are
// processed in sub-chunks).  The resolve kernel ray-casts the few analytic primitives per pixel, merges the buffer,
// shades and stores, and resets the buffer for the next call.  Pixel rays, near / far clipping, depth convention
// and shading are those of render_kernel (the ray-cast path, SS_RENDER=raycast); equal depths are decided by the
// key (lower geom / triangle wins), so the images do not depend on the order of the atomics.
#define ZEMPTY 0xffffffffffffffffull
#define RASTER_SMALL 24     // pixel boxes up to this area are walked by the owning thread, larger ones by the whole warp

// bounds (lo, hi in the geom frame, world = T[0..3) + R local) against the pyramid pn / the near plane; conservative
__device__ __forceinline__ bool box_in_pyramid(const float4 lo, const float4 hi, const float* T, const float* eye, const float* cR,
                                               const float (*pn)[3], float znear) {
  const float* R = T + 3;
  float hc[3] = {0.5f * (lo.x + hi.x), 0.5f * (lo.y + hi.y), 0.5f * (lo.z + hi.z)};
  float hh[3] = {0.5f * (hi.x - lo.x), 0.5f * (hi.y - lo.y), 0.5f * (hi.z - lo.z)};
  float dw[3] = {T[0] - eye[0], T[1] - eye[1], T[2] - eye[2]};
  float A[9], bc[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
#pragma unroll
    for (int j = 0; j < 3; j++) A[3 * i + j] = cR[i] * R[j] + cR[3 + i] * R[3 + j] + cR[6 + i] * R[6 + j];
    bc[i] = cR[i] * dw[0] + cR[3 + i] * dw[1] + cR[6 + i] * dw[2] + A[3 * i] * hc[0] + A[3 * i + 1] * hc[1] + A[3 * i + 2] * hc[2];
  }
  float radz = fabsf(A[6]) * hh[0] + fabsf(A[7]) * hh[1] + fabsf(A[8]) * hh[2];
  if (-bc[2] + radz < znear) return false;      // entirely nearer than the near plane / behind the camera
#pragma unroll
  for (int p = 0; p < 4; p++) {
    float dist = pn[p][0] * bc[0] + pn[p][1] * bc[1] + pn[p][2] * bc[2], rad = 0.f;
#pragma unroll
    for (int j = 0; j < 3; j++) rad += fabsf(pn[p][0] * A[j] + pn[p][1] * A[3 + j] + pn[p][2] * A[6 + j]) * hh[j];
    if (dist < -rad * 1.0001f - 1e-6f) return false;
  }
  return true;
}

// per env of the sub-chunk: camera pose, the primitives inside the image pyramid (geom-id order), chunk visibility
__global__ void __launch_bounds__(256) raster_setup_kernel(RayModel r, int env_begin, int cam, int W, int H, float fovy_deg,
                                                           const float* __restrict__ xpos, const float* __restrict__ xquat,
                                                           const float* __restrict__ xf_all, float* __restrict__ campose,
                                                           int* __restrict__ prim, unsigned char* __restrict__ cvis) {
  __shared__ float eye[3], cR[9], pn[4][3];
  __shared__ float focal;
  int le = blockIdx.x, e = env_begin + le, tid = threadIdx.x;
  if (tid == 0) {
    camera_pose(r, cam, xpos, xquat, e, eye, cR);
    focal = 0.5f * H / tanf(fovy_deg * 3.14159265358979f / 360.0f);
    pyramid(pn, -0.5f * W / focal, 0.5f * W / focal, 0.5f * H / focal, -0.5f * H / focal);
    float* cp = campose + 16 * le;
    for (int k = 0; k < 3; k++) cp[k] = eye[k];
    for (int k = 0; k < 9; k++) cp[3 + k] = cR[k];
    cp[12] = focal;
  }
  __syncthreads();
  const float* xf = xf_all + (size_t)e * r.nraygeom * 12;
  const float znear = r.znear * r.extent;
  for (int c = tid; c < r.nchunk; c += blockDim.x) {
    int4 ch = r.rchunk[c];
    cvis[(size_t)le * r.nchunk + c] = box_in_pyramid(r.rchunk_box[2 * c], r.rchunk_box[2 * c + 1], xf + 12 * ch.x, eye, cR, pn, znear);
  }
  if (tid < 32) {   // ordered compaction keeps the geom-id order (ties between primitives resolve as in the ray-cast path's id order)
    int n = 0;
    for (int base = 0; base < r.nraygeom; base += 32) {
      int k = base + tid;
      bool keep = false;
      if (k < r.nraygeom) {
        const float4* rec = r.rg_rec + 4 * k;
        int type = REC_TYPE(rec);
        if (type != GEOM_MESH && ((0x7 >> REC_GROUP(rec)) & 1))
          keep = type == GEOM_PLANE ? true : frustum_keeps<true>(rec, xf + 12 * k, eye, cR, pn, znear);
      }
      unsigned mask = __ballot_sync(0xffffffffu, keep);
      int slot = n + __popc(mask & ((1u << tid) - 1));
      if (keep && slot < MAXPRIM) prim[(size_t)le * (MAXPRIM + 1) + 1 + slot] = k;
      n += __popc(mask);
    }
    if (tid == 0) prim[(size_t)le * (MAXPRIM + 1)] = min(n, MAXPRIM);
  }
}

struct RTri { float a[3], e1[3], e2[3], q[3], c0; int u0, u1, v0, v1; unsigned id; };
struct __align__(16) RItem { RTri t; int le; int pad; };   // 80 bytes
#define RASTER_ITEM 1024   // pixels per queued item

__device__ __forceinline__ void raster_pixel(const RTri& t, int u, int v, int W, int H, float invf, float znear, float zfar,
                                             unsigned long long* __restrict__ zb) {
  const float EPS = 1e-6f;
  float dx = (u + 0.5f - 0.5f * W) * invf, dy = -(v + 0.5f - 0.5f * H) * invf;   // ray direction (dx, dy, -1), origin = the eye
  float px = dy * t.e2[2] + t.e2[1], py = -t.e2[0] - dx * t.e2[2], pz = dx * t.e2[1] - dy * t.e2[0];   // d x e2
  float det = t.e1[0] * px + t.e1[1] * py + t.e1[2] * pz;
  if (fabsf(det) < 1e-30f) return;
  float idet = 1.0f / det;
  float uu = -(t.a[0] * px + t.a[1] * py + t.a[2] * pz) * idet;
  if (uu < -EPS || uu > 1.f + EPS) return;
  float vv = (dx * t.q[0] + dy * t.q[1] - t.q[2]) * idet;
  if (vv < -EPS || uu + vv > 1.f + EPS) return;
  float x = t.c0 * idet;
  if (!(x >= znear && x <= zfar)) return;
  unsigned long long key = ((unsigned long long)__float_as_uint(x) << 32) | t.id;
  unsigned long long* z = zb + (size_t)v * W + u;
  if (key < *z) atomicMin(z, key);
}

__global__ void __launch_bounds__(128) raster_tri_kernel(RayModel r, int env_begin, int W, int H, const float* __restrict__ xf_all,
                                                         const float* __restrict__ campose, const unsigned char* __restrict__ cvis,
                                                         unsigned long long* __restrict__ zbuf, unsigned long long* __restrict__ stats,
                                                         RItem* __restrict__ queue, int* __restrict__ qcount, int qcap) {
  const int c = blockIdx.x, le = blockIdx.y;
  if (!cvis[(size_t)le * r.nchunk + c]) return;
  if (stats && threadIdx.x == 0) atomicAdd(stats + 0, 1ull);   // chunks inside the pyramid
  const int4 ch = r.rchunk[c];
  const float* T = xf_all + ((size_t)(env_begin + le) * r.nraygeom + ch.x) * 12;
  const float* cp = campose + 16 * le;
  // camera <- geom: x_cam = A x_local + tc
  float A[9], tc[3];
  {
    const float *eye = cp, *cR = cp + 3, *R = T + 3;
    float dw[3] = {T[0] - eye[0], T[1] - eye[1], T[2] - eye[2]};
#pragma unroll
    for (int i = 0; i < 3; i++) {
      tc[i] = cR[i] * dw[0] + cR[3 + i] * dw[1] + cR[6 + i] * dw[2];
#pragma unroll
      for (int j = 0; j < 3; j++) A[3 * i + j] = cR[i] * R[j] + cR[3 + i] * R[3 + j] + cR[6 + i] * R[6 + j];
    }
  }
  const float f = cp[12], invf = 1.0f / f, znear = r.znear * r.extent, zfar = r.zfar * r.extent;
  unsigned long long* zb = zbuf + (size_t)le * W * H;
  const int lane = threadIdx.x & 31;
  float4 nv0 = make_float4(0, 0, 0, 0), ne1 = nv0, ne2 = nv0;
  if ((int)threadIdx.x < ch.z) {
    const size_t ti = (size_t)ch.y + threadIdx.x;
    nv0 = __ldg(r.tri + 3 * ti); ne1 = __ldg(r.tri + 3 * ti + 1); ne2 = __ldg(r.tri + 3 * ti + 2);
  }
  for (int base = 0; base < ch.z; base += blockDim.x) {   // same trip count for every thread (warp-wide ballots below)
    const int i = base + threadIdx.x;
    const float4 v0 = nv0, e1 = ne1, e2 = ne2;
    if (i + (int)blockDim.x < ch.z) {   // the next triangle's loads fly while this one is rasterised
      const size_t ti = (size_t)ch.y + i + blockDim.x;
      nv0 = __ldg(r.tri + 3 * ti); ne1 = __ldg(r.tri + 3 * ti + 1); ne2 = __ldg(r.tri + 3 * ti + 2);
    }
    RTri t;
    bool large = false;
    t.u0 = 0; t.u1 = -1; t.v0 = 0; t.v1 = -1; t.id = 0; t.c0 = 0;
#pragma unroll
    for (int k = 0; k < 3; k++) { t.a[k] = t.e1[k] = t.e2[k] = t.q[k] = 0.f; }
    if (i < ch.z) {
      const int ti = ch.y + i;
#pragma unroll
      for (int k = 0; k < 3; k++) {
        t.a[k] = tc[k] + A[3 * k] * v0.x + A[3 * k + 1] * v0.y + A[3 * k + 2] * v0.z;
        t.e1[k] = A[3 * k] * e1.x + A[3 * k + 1] * e1.y + A[3 * k + 2] * e1.z;
        t.e2[k] = A[3 * k] * e2.x + A[3 * k + 1] * e2.y + A[3 * k + 2] * e2.z;
      }
      float P[3][3];
#pragma unroll
      for (int k = 0; k < 3; k++) { P[0][k] = t.a[k]; P[1][k] = t.a[k] + t.e1[k]; P[2][k] = t.a[k] + t.e2[k]; }
      const float d0 = -P[0][2], d1 = -P[1][2], d2 = -P[2][2];
      const float dmin = fminf(d0, fminf(d1, d2)), dmax = fmaxf(d0, fmaxf(d1, d2));
      if (dmax >= znear && dmin <= zfar) {
        float xmin = 1e30f, xmax = -1e30f, ymin = 1e30f, ymax = -1e30f;
        const float zc = znear * 0.999f;
#pragma unroll
        for (int k = 0; k < 3; k++) {
          const float* p = P[k];
          const float* qn = P[(k + 1) % 3];
          float dp = -p[2], dq = -qn[2];
          if (dp >= zc) {
            float sx = f * p[0] / dp, sy = -f * p[1] / dp;
            xmin = fminf(xmin, sx); xmax = fmaxf(xmax, sx); ymin = fminf(ymin, sy); ymax = fmaxf(ymax, sy);
          }
          if ((dp >= zc) != (dq >= zc)) {   // the edge crosses the near plane: its intersection bounds the visible part
            float sI = (zc - dp) / (dq - dp);
            float ix = p[0] + sI * (qn[0] - p[0]), iy = p[1] + sI * (qn[1] - p[1]);
            float sx = f * ix / zc, sy = -f * iy / zc;
            xmin = fminf(xmin, sx); xmax = fmaxf(xmax, sx); ymin = fminf(ymin, sy); ymax = fmaxf(ymax, sy);
          }
        }
        // pixel centre u sits at screen x = u + 0.5 - W/2
        xmin = fmaxf(xmin + 0.5f * W - 0.5f - 1e-3f, -1.f); xmax = fminf(xmax + 0.5f * W - 0.5f + 1e-3f, (float)W);
        ymin = fmaxf(ymin + 0.5f * H - 0.5f - 1e-3f, -1.f); ymax = fminf(ymax + 0.5f * H - 0.5f + 1e-3f, (float)H);
        t.u0 = max(0, (int)ceilf(xmin)); t.u1 = min(W - 1, (int)floorf(xmax));
        t.v0 = max(0, (int)ceilf(ymin)); t.v1 = min(H - 1, (int)floorf(ymax));
        if (t.u0 <= t.u1 && t.v0 <= t.v1) {
          // q = tvec x e1 with tvec = -a;  c0 = e2 . q
          t.q[0] = -(t.a[1] * t.e1[2] - t.a[2] * t.e1[1]); t.q[1] = -(t.a[2] * t.e1[0] - t.a[0] * t.e1[2]); t.q[2] = -(t.a[0] * t.e1[1] - t.a[1] * t.e1[0]);
          t.c0 = t.e2[0] * t.q[0] + t.e2[1] * t.q[1] + t.e2[2] * t.q[2];
          t.id = ((unsigned)ch.x << 22) | (unsigned)ti;
          const int area = (t.u1 - t.u0 + 1) * (t.v1 - t.v0 + 1);
          if (stats) {
            atomicAdd(stats + 1, 1ull); atomicAdd(stats + 2, (unsigned long long)area);
            if (area > RASTER_SMALL) { atomicAdd(stats + 3, 1ull); atomicAdd(stats + 4, (unsigned long long)area); }
            if (dmin < zc) { atomicAdd(stats + 5, 1ull); atomicAdd(stats + 6, (unsigned long long)area); }
            if (area > 4096) { atomicAdd(stats + 7, 1ull); atomicAdd(stats + 8, (unsigned long long)area); }
          }
          if (area <= RASTER_SMALL) {
            for (int v = t.v0; v <= t.v1; v++)
              for (int u = t.u0; u <= t.u1; u++) raster_pixel(t, u, v, W, H, invf, znear, zfar, zb);
          } else {
            // larger pixel boxes go to the queue of raster_large_kernel (one warp per item), cut into row bands of at
            // most RASTER_ITEM pixels so that a screen-filling triangle spreads over many warps
            const int bw = t.u1 - t.u0 + 1, rows = max(1, RASTER_ITEM / bw), nitem = (t.v1 - t.v0 + rows) / rows;
            const int q0 = atomicAdd(qcount, nitem);
            if (q0 + nitem <= qcap) {
              for (int k = 0; k < nitem; k++) {
                RItem it;
                it.t = t; it.le = le;
                it.t.v0 = t.v0 + k * rows; it.t.v1 = min(t.v1, it.t.v0 + rows - 1);
                queue[q0 + k] = it;
              }
            } else large = true;   // queue full: this warp walks the box itself (below)
          }
        }
      }
    }
    unsigned big = __ballot_sync(0xffffffffu, large);
    while (big) {
      const int src = __ffs(big) - 1;
      big &= big - 1;
      RTri b;
#pragma unroll
      for (int k = 0; k < 3; k++) {
        b.a[k] = __shfl_sync(0xffffffffu, t.a[k], src); b.e1[k] = __shfl_sync(0xffffffffu, t.e1[k], src);
        b.e2[k] = __shfl_sync(0xffffffffu, t.e2[k], src); b.q[k] = __shfl_sync(0xffffffffu, t.q[k], src);
      }
      b.c0 = __shfl_sync(0xffffffffu, t.c0, src); b.id = __shfl_sync(0xffffffffu, t.id, src);
      b.u0 = __shfl_sync(0xffffffffu, t.u0, src); b.u1 = __shfl_sync(0xffffffffu, t.u1, src);
      b.v0 = __shfl_sync(0xffffffffu, t.v0, src); b.v1 = __shfl_sync(0xffffffffu, t.v1, src);
      const int bw = b.u1 - b.u0 + 1, n = bw * (b.v1 - b.v0 + 1);
      for (int idx = lane; idx < n; idx += 32) raster_pixel(b, b.u0 + idx % bw, b.v0 + idx / bw, W, H, invf, znear, zfar, zb);
    }
  }
}

// one warp per queued (triangle, row band): lanes stride over the band's pixels
__global__ void __launch_bounds__(256) raster_large_kernel(RayModel r, int W, int H, const float* __restrict__ campose,
                                                           unsigned long long* __restrict__ zbuf, const RItem* __restrict__ queue,
                                                           const int* __restrict__ qcount, int qcap) {
  const int lane = threadIdx.x & 31, nwarp = gridDim.x * (blockDim.x >> 5);
  const int count = min(*qcount, qcap);
  const float znear = r.znear * r.extent, zfar = r.zfar * r.extent;
  for (int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < count; i += nwarp) {
    const int4* src = reinterpret_cast<const int4*>(queue + i);
    __align__(16) RItem it;
    int4* dst = reinterpret_cast<int4*>(&it);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(RItem) / 16); k++) dst[k] = __ldg(src + k);
    const float invf = 1.0f / campose[16 * it.le + 12];
    unsigned long long* zb = zbuf + (size_t)it.le * W * H;
    const RTri& b = it.t;
    // lanes stride over the band's pixels in row-major order; (u, v) advance without a division per pixel
    const int bw = b.u1 - b.u0 + 1;
    int u = b.u0 + lane % bw, v = b.v0 + lane / bw;
    const int du = 32 % bw, dv = 32 / bw;
    while (v <= b.v1) {
      raster_pixel(b, u, v, W, H, invf, znear, zfar, zb);
      u += du; v += dv;
      if (u > b.u1) { u -= bw; v++; }
    }
  }
}

// Pixel epilogue of the raster path: the arithmetic of shade_and_store with the per-frame constants hoisted (vector loads of
// the material and the geom rotation, light directions normalised once per CTA: lvec[.][3] < 0 marks a unit direction).
__device__ __forceinline__ void shade_fast(const RayModel& r, float x, int k, int tri, const float* hn, const float* dw, const float* cam_eye,
                                           const float (*lvec)[4], const float (*lcol)[9], int nslot, const float* xf, float zfar, size_t pix,
                                           uint8_t* __restrict__ rgb, float* __restrict__ depth, float depth_limit, int bgr) {
  if (x < 0 || x > zfar) { x = zfar; k = -1; }
  if (depth) depth[pix] = (depth_limit > 0 && x > depth_limit) ? 0.f : x;   // utils.limit_depth_distance
  if (!rgb) return;
  float col[3];
  const float idw = rsqrtf(dw[0] * dw[0] + dw[1] * dw[1] + dw[2] * dw[2]);
  if (k < 0) {
    float tt = 0.5f * (1.0f + dw[2] * idw);
#pragma unroll
    for (int a = 0; a < 3; a++) col[a] = r.nsky >= 2 ? tt * r.sky[a] + (1 - tt) * r.sky[3 + a] : 0.f;
  } else {
    const float4* sh4 = reinterpret_cast<const float4*>(r.rg_shade + 8 * k);
    const float4 s0 = __ldg(sh4), s1 = __ldg(sh4 + 1);          // rgb, (unused alpha slot = sh[3]); specular, shininess, emission, -
    const float4* x4 = reinterpret_cast<const float4*>(xf + 12 * k);
    const float4 t0 = x4[0], t1 = x4[1], t2 = x4[2];            // pos.xyz R00 | R01 R02 R10 R11 | R12 R20 R21 R22
    float n[3] = {t0.w * hn[0] + t1.x * hn[1] + t1.y * hn[2], t1.z * hn[0] + t1.w * hn[1] + t2.x * hn[2], t2.y * hn[0] + t2.z * hn[1] + t2.w * hn[2]};
    float inv = rsqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    n[0] *= inv; n[1] *= inv; n[2] *= inv;
    const float pos[3] = {cam_eye[0] + x * dw[0], cam_eye[1] + x * dw[1], cam_eye[2] + x * dw[2]};
    const float vw[3] = {-dw[0] * idw, -dw[1] * idw, -dw[2] * idw};
    if (n[0] * vw[0] + n[1] * vw[1] + n[2] * vw[2] < 0) { n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2]; }
    float base[3] = {s0.x, s0.y, s0.z};
    const float spec = s1.x, shininess = fmaxf(s1.y * 128.0f, 1.0f), emis = s1.z;
    if (s0.w > 0.f) texture_modulate(r, k, tri, xf + 12 * k, pos, base);
    // col = base * (emission + sum_l (ambient_l + diffuse_l nl_l)) + specular * sum_l specular_l hs_l
    float da[3] = {emis, emis, emis}, sa[3] = {0.f, 0.f, 0.f};
    for (int l = 0; l < nslot; l++) {
      const float* lv = lvec[l];
      const float* lc = lcol[l];
      float L[3] = {lv[0], lv[1], lv[2]};
      if (lv[3] > 0.f) {   // positional light
        L[0] -= pos[0]; L[1] -= pos[1]; L[2] -= pos[2];
        float il = rsqrtf(L[0] * L[0] + L[1] * L[1] + L[2] * L[2]);
        L[0] *= il; L[1] *= il; L[2] *= il;
      }
      float nl = fmaxf(n[0] * L[0] + n[1] * L[1] + n[2] * L[2], 0.f), hs = 0.f;
      if (nl > 0) {
        float hv[3] = {L[0] + vw[0], L[1] + vw[1], L[2] + vw[2]};
        float ih = rsqrtf(hv[0] * hv[0] + hv[1] * hv[1] + hv[2] * hv[2]);
        hs = __powf(fmaxf((n[0] * hv[0] + n[1] * hv[1] + n[2] * hv[2]) * ih, 0.f), shininess);
      }
#pragma unroll
      for (int a = 0; a < 3; a++) { da[a] += lc[a] + lc[3 + a] * nl; sa[a] = fmaf(lc[6 + a], hs, sa[a]); }
    }
#pragma unroll
    for (int a = 0; a < 3; a++) col[a] = base[a] * da[a] + spec * sa[a];
  }
  uint8_t* px = rgb + 3 * pix;
#pragma unroll
  for (int a = 0; a < 3; a++) px[bgr ? 2 - a : a] = (uint8_t)(fminf(fmaxf(col[a], 0.f), 1.f) * 255.0f + 0.5f);
}

#define MAXPLANE 4
#ifndef RESOLVE_CTAS
#define RESOLVE_CTAS 5   // CTAs of 256 threads per SM (48 registers): occupancy beats register-resident constants (640x480 frame of 4096 envs: 2 CTAs 37.0 ms, 3 30.4, 4 27.1, 5 25.7, 6 26.3)
#endif
struct SPlane { float n[3], c0, ax[3], ox, ay[3], oy, s0, s1; int k; };   // world normal, n . (eye - p); in-plane axes / offsets for sized planes

// per pixel: analytic primitives by ray casting (planes through a precomputed world-space form), merge with the depth/id
// buffer, shade, store, reset the buffer
__device__ __forceinline__ void plane_hit(const SPlane& P, const float* dw, float znear, Hit& h) {
  const float d2 = P.n[0] * dw[0] + P.n[1] * dw[1] + P.n[2] * dw[2];
  if (d2 > -1e-15f) return;
  const float x = -P.c0 / d2;
  if (x < znear || (h.t >= 0 && x >= h.t)) return;
  if (P.s0 > 0 || P.s1 > 0) {
    const float px = P.ox + x * (P.ax[0] * dw[0] + P.ax[1] * dw[1] + P.ax[2] * dw[2]);
    const float py = P.oy + x * (P.ay[0] * dw[0] + P.ay[1] * dw[1] + P.ay[2] * dw[2]);
    if ((P.s0 > 0 && fabsf(px) > P.s0) || (P.s1 > 0 && fabsf(py) > P.s1)) return;
  }
  h.t = x; h.k = P.k; h.tri = -1; h.n[0] = 0.f; h.n[1] = 0.f; h.n[2] = 1.f;
}
// the depth/id buffer's winner replaces the primitive hit when it is nearer
__device__ __forceinline__ void merge_mesh_hit(const RayModel& r, unsigned long long key, Hit& h) {
  const float x = __uint_as_float((unsigned)(key >> 32));
  if (h.t < 0 || x < h.t) {
    const unsigned id = (unsigned)key;
    const size_t ti = id & 0x3fffffu;
    const float4 e1 = __ldg(r.tri + 3 * ti + 1), e2 = __ldg(r.tri + 3 * ti + 2);
    h.t = x; h.k = (int)(id >> 22); h.tri = (int)ti;
    h.n[0] = e1.y * e2.z - e1.z * e2.y; h.n[1] = e1.z * e2.x - e1.x * e2.z; h.n[2] = e1.x * e2.y - e1.y * e2.x;
  }
}

template <int NSLOT>   // light slots known at launch (1..3: registers, unrolled), 0 = generic
__global__ void __launch_bounds__(256, RESOLVE_CTAS) raster_resolve_kernel(RayModel r, int env_begin, int out_begin, int cam, int W, int H,
                                                             const float* __restrict__ xpos, const float* __restrict__ xquat,
                                                             const float* __restrict__ xf_all, const float* __restrict__ campose,
                                                             const int* __restrict__ prim, unsigned long long* __restrict__ zbuf,
                                                             uint8_t* __restrict__ rgb, float* __restrict__ depth, float depth_limit, int post) {
  __shared__ float cam_eye[3], cam_R[9], focal;
  __shared__ float lvec[MAXLIGHT][4];
  __shared__ float lcol[MAXLIGHT][9];
  __shared__ int sprim[MAXPRIM], nprim, nplane, nslot;
  __shared__ SPlane planes[MAXPLANE];
  const int le = blockIdx.z, e = env_begin + le, tid = threadIdx.y * 32 + threadIdx.x;
  const float* xf = xf_all + (size_t)e * r.nraygeom * 12;
  if (tid < 13) {
    float x = campose[16 * le + tid];
    if (tid < 3) cam_eye[tid] = x; else if (tid < 12) cam_R[tid - 3] = x; else focal = x;
  }
  if (rgb && tid >= 14 && tid < 15 + r.nlight && tid < 14 + MAXLIGHT) light_colours(r, tid - 14, lcol[tid - 14]);
  if (rgb && tid >= 32 && tid < 32 + r.nlight && tid < 32 + MAXLIGHT - 1) {
    float* L = lvec[tid - 31];
    light_vector(r, tid - 32, xpos, xquat, e, L);
    if (L[3] == 0.f) {   // directional: normalise once
      float il = rsqrtf(L[0] * L[0] + L[1] * L[1] + L[2] * L[2]);
      L[0] *= il; L[1] *= il; L[2] *= il;
    }
  }
  __syncthreads();
  if (tid == 0) {
    // slot 0 = headlight along the camera's +z; inactive headlight: the scene lights move down one slot
    const int nl = min(r.nlight, MAXLIGHT - 1);
    if (r.headlight_active) { lvec[0][0] = cam_R[2]; lvec[0][1] = cam_R[5]; lvec[0][2] = cam_R[8]; lvec[0][3] = 0.f; nslot = nl + 1; }
    else {
      for (int l = 0; l < nl; l++) { for (int a = 0; a < 4; a++) lvec[l][a] = lvec[l + 1][a]; for (int a = 0; a < 9; a++) lcol[l][a] = lcol[l + 1][a]; }
      nslot = nl;
    }
    // planes leave the generic list: t = -n . (eye - p) / (n . d) per pixel
    const int n0 = prim[(size_t)le * (MAXPRIM + 1)];
    int np = 0, npl = 0;
    for (int i = 0; i < n0; i++) {
      const int k = prim[(size_t)le * (MAXPRIM + 1) + 1 + i];
      const float4* rec = r.rg_rec + 4 * k;
      if (REC_TYPE(rec) == GEOM_PLANE && npl < MAXPLANE) {
        const float *T = xf + 12 * k, *R = T + 3;
        SPlane& P = planes[npl++];
        const float dif[3] = {cam_eye[0] - T[0], cam_eye[1] - T[1], cam_eye[2] - T[2]};
        P.n[0] = R[2]; P.n[1] = R[5]; P.n[2] = R[8];
        P.ax[0] = R[0]; P.ax[1] = R[3]; P.ax[2] = R[6];
        P.ay[0] = R[1]; P.ay[1] = R[4]; P.ay[2] = R[7];
        P.c0 = P.n[0] * dif[0] + P.n[1] * dif[1] + P.n[2] * dif[2];
        P.ox = P.ax[0] * dif[0] + P.ax[1] * dif[1] + P.ax[2] * dif[2];
        P.oy = P.ay[0] * dif[0] + P.ay[1] * dif[1] + P.ay[2] * dif[2];
        P.s0 = rec[1].y; P.s1 = rec[1].z; P.k = k;
      } else sprim[np++] = k;
    }
    nprim = np; nplane = npl;
  }
  __syncthreads();
  const int u = blockIdx.x * 32 + threadIdx.x;
  if (u >= W) return;
  const float invf = 1.0f / focal, znear = r.znear * r.extent, zfar = r.zfar * r.extent;
  const int np = nprim, npl = nplane, rot = (post >> 1) & 3, bgr = post & 1, lo = out_begin + le;
  if constexpr (NSLOT == 0) {
    // generic path: any number of light slots, everything read from shared memory per pixel
    const int ns = nslot;
#pragma unroll 1
    for (int row = 0; row < 4; row++) {   // a 32 x 32 pixel tile per CTA: the prologue above is paid once per 1024 pixels
      const int v = blockIdx.y * 32 + row * 8 + threadIdx.y;
      if (v >= H) break;
      float dl[3] = {(u + 0.5f - 0.5f * W) * invf, -(v + 0.5f - 0.5f * H) * invf, -1.0f};   // same rays as raster_pixel
      float dw[3] = {cam_R[0] * dl[0] + cam_R[1] * dl[1] + cam_R[2] * dl[2], cam_R[3] * dl[0] + cam_R[4] * dl[1] + cam_R[5] * dl[2],
                     cam_R[6] * dl[0] + cam_R[7] * dl[1] + cam_R[8] * dl[2]};
      unsigned long long* z = zbuf + ((size_t)le * H + v) * W + u;
      const unsigned long long key = *z;
      Hit h; h.t = -1.f; h.k = -1; h.tri = -1; h.n[0] = h.n[1] = h.n[2] = 0;
      for (int i = 0; i < npl; i++) plane_hit(planes[i], dw, znear, h);
      if (np) {
        const float vv = dw[0] * dw[0] + dw[1] * dw[1] + dw[2] * dw[2];
        for (int i = 0; i < np; i++) trace_one<false>(r, xf, r.rg_rec, sprim[i], cam_eye, dw, vv, znear, 0, -1, h);
      }
      if (key != ZEMPTY) { *z = ZEMPTY; merge_mesh_hit(r, key, h); }
      const size_t pix = rot == 0 ? ((size_t)lo * H + v) * W + u
                       : rot == 1 ? ((size_t)lo * W + (W - 1 - u)) * H + v
                                  : ((size_t)lo * W + u) * H + (H - 1 - v);
      shade_fast(r, h.t, h.k, h.tri, h.n, dw, cam_eye, lvec, lcol, ns, xf, zfar, pix, rgb, depth, depth_limit, bgr);
    }
  } else {
    // NSLOT light slots known at launch: camera, first plane and lights live in registers for the four rows of the tile
    float cR[9], eye[3], Lv[NSLOT][4], Lc[NSLOT][9];
#pragma unroll
    for (int k = 0; k < 9; k++) cR[k] = cam_R[k];
#pragma unroll
    for (int k = 0; k < 3; k++) eye[k] = cam_eye[k];
#pragma unroll
    for (int l = 0; l < NSLOT; l++) {
#pragma unroll
      for (int k = 0; k < 4; k++) Lv[l][k] = lvec[l][k];
#pragma unroll
      for (int k = 0; k < 9; k++) Lc[l][k] = lcol[l][k];
    }
    const bool p0 = npl > 0 && !(planes[0].s0 > 0 || planes[0].s1 > 0);   // first plane unbounded (a floor): inline form
    const float pn[3] = {planes[0].n[0], planes[0].n[1], planes[0].n[2]}, pc0 = planes[0].c0;
    const int pk = planes[0].k;
    const float dx = (u + 0.5f - 0.5f * W) * invf;
#pragma unroll 1
    for (int row = 0; row < 4; row++) {
      const int v = blockIdx.y * 32 + row * 8 + threadIdx.y;
      if (v >= H) break;
      const float dy = -(v + 0.5f - 0.5f * H) * invf;
      const float dw[3] = {cR[0] * dx + cR[1] * dy - cR[2], cR[3] * dx + cR[4] * dy - cR[5], cR[6] * dx + cR[7] * dy - cR[8]};
      unsigned long long* z = zbuf + ((size_t)le * H + v) * W + u;
      const unsigned long long key = *z;
      Hit h; h.t = -1.f; h.k = -1; h.tri = -1; h.n[0] = h.n[1] = h.n[2] = 0;
      if (p0) {
        const float d2 = pn[0] * dw[0] + pn[1] * dw[1] + pn[2] * dw[2];
        if (d2 <= -1e-15f) {
          const float x = -pc0 / d2;
          if (x >= znear) { h.t = x; h.k = pk; h.tri = -1; h.n[2] = 1.f; }
        }
      }
      for (int i = p0 ? 1 : 0; i < npl; i++) plane_hit(planes[i], dw, znear, h);
      if (np) {
        const float vv = dw[0] * dw[0] + dw[1] * dw[1] + dw[2] * dw[2];
        for (int i = 0; i < np; i++) trace_one<false>(r, xf, r.rg_rec, sprim[i], eye, dw, vv, znear, 0, -1, h);
      }
      if (key != ZEMPTY) { *z = ZEMPTY; merge_mesh_hit(r, key, h); }
      const size_t pix = rot == 0 ? ((size_t)lo * H + v) * W + u
                       : rot == 1 ? ((size_t)lo * W + (W - 1 - u)) * H + v
                                  : ((size_t)lo * W + u) * H + (H - 1 - v);
      // ---- shade (the arithmetic of shade_fast with the light loop unrolled over registers)
      float x = h.t;
      int k = h.k;
      if (x < 0 || x > zfar) { x = zfar; k = -1; }
      if (depth) depth[pix] = (depth_limit > 0 && x > depth_limit) ? 0.f : x;
      if (!rgb) continue;
      float col[3];
      const float idw = rsqrtf(dw[0] * dw[0] + dw[1] * dw[1] + dw[2] * dw[2]);
      if (k < 0) {
        const float tt = 0.5f * (1.0f + dw[2] * idw);
#pragma unroll
        for (int a = 0; a < 3; a++) col[a] = r.nsky >= 2 ? tt * r.sky[a] + (1 - tt) * r.sky[3 + a] : 0.f;
      } else {
        const float4* sh4 = reinterpret_cast<const float4*>(r.rg_shade + 8 * k);
        const float4 s0 = __ldg(sh4), s1 = __ldg(sh4 + 1);
        const float4* x4 = reinterpret_cast<const float4*>(xf + 12 * k);
        const float4 t0 = x4[0], t1 = x4[1], t2 = x4[2];
        float n[3] = {t0.w * h.n[0] + t1.x * h.n[1] + t1.y * h.n[2], t1.z * h.n[0] + t1.w * h.n[1] + t2.x * h.n[2], t2.y * h.n[0] + t2.z * h.n[1] + t2.w * h.n[2]};
        const float inv = rsqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        const float vw[3] = {-dw[0] * idw, -dw[1] * idw, -dw[2] * idw};
        const float sgn = (n[0] * vw[0] + n[1] * vw[1] + n[2] * vw[2] < 0) ? -inv : inv;   // two-sided lighting
        n[0] *= sgn; n[1] *= sgn; n[2] *= sgn;
        const float pos[3] = {eye[0] + x * dw[0], eye[1] + x * dw[1], eye[2] + x * dw[2]};
        float base[3] = {s0.x, s0.y, s0.z};
        const float spec = s1.x, shininess = fmaxf(s1.y * 128.0f, 1.0f), emis = s1.z;
        if (s0.w > 0.f) texture_modulate(r, k, h.tri, xf + 12 * k, pos, base);
        float da[3] = {emis, emis, emis}, sa[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int l = 0; l < NSLOT; l++) {
          float L[3] = {Lv[l][0], Lv[l][1], Lv[l][2]};
          if (Lv[l][3] > 0.f) {   // positional light
            L[0] -= pos[0]; L[1] -= pos[1]; L[2] -= pos[2];
            const float il = rsqrtf(L[0] * L[0] + L[1] * L[1] + L[2] * L[2]);
            L[0] *= il; L[1] *= il; L[2] *= il;
          }
          const float nl = fmaxf(n[0] * L[0] + n[1] * L[1] + n[2] * L[2], 0.f);
          float hs = 0.f;
          if (nl > 0) {
            const float hv[3] = {L[0] + vw[0], L[1] + vw[1], L[2] + vw[2]};
            const float ih = rsqrtf(hv[0] * hv[0] + hv[1] * hv[1] + hv[2] * hv[2]);
            hs = __powf(fmaxf((n[0] * hv[0] + n[1] * hv[1] + n[2] * hv[2]) * ih, 0.f), shininess);
          }
#pragma unroll
          for (int a = 0; a < 3; a++) { da[a] += Lc[l][a] + Lc[l][3 + a] * nl; sa[a] = fmaf(Lc[l][6 + a], hs, sa[a]); }
        }
#pragma unroll
        for (int a = 0; a < 3; a++) col[a] = base[a] * da[a] + spec * sa[a];
      }
      uint8_t* px = rgb + 3 * pix;
#pragma unroll
      for (int a = 0; a < 3; a++) px[bgr ? 2 - a : a] = (uint8_t)(fminf(fmaxf(col[a], 0.f), 1.f) * 255.0f + 0.5f);
    }
  }
}

// ----------------------------------------------------------------------------- C ABI
static int prepare(ss_batch* B, cudaStream_t st, int env_begin = 0, int env_count = -1) {
  const RayModel& r = B->model->rm;
  if (!r.present) return ss_fail("the model was compiled without ray geometry (compile with with_render=True)");
  if (!B->bufs.xpos || !B->bufs.xquat) return ss_fail("ray casting needs the xpos/xquat buffers");
  cudaSetDevice(B->model->device);
  if (!B->ray_xf) {
    CUDA_OK(cudaMalloc((void**)&B->ray_xf, (size_t)B->nenv * r.nraygeom * 12 * sizeof(float)));
  }
  if (env_count < 0) env_count = B->nenv - env_begin;
  int n = env_count * r.nraygeom;
  ray_prepare_kernel<<<(n + 255) / 256, 256, 0, st>>>(r, env_begin, env_count, B->bufs.xpos, B->bufs.xquat, B->ray_xf);
  B->launches++;
  CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int ss_batch_lidar(ss_batch* B, float* out_dev, ss_stream s) {
  if (!B) return ss_fail("ss_batch_lidar: null batch");
  cudaStream_t st = (cudaStream_t)s;
  if (prepare(B, st) != 0) return -1;
  const RayModel& r = B->model->rm;
  if (r.nrange == 0) return ss_fail("model has no rangefinder sensors");
  size_t smem = RAY_SMEM_BYTES(r) + (size_t)LIDAR_WARPS * r.nraygeom * sizeof(int);
  cudaFuncSetAttribute(lidar_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  lidar_kernel<<<B->nenv, 32 * LIDAR_WARPS, smem, st>>>(r, B->nenv, B->dm.nsensordata, B->bufs.xpos, B->bufs.xquat, B->ray_xf, out_dev,
                                           out_dev ? nullptr : B->bufs.sensordata);
  B->launches++;
  CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int ss_batch_rays(ss_batch* B, int nray, const float* origin, const float* dir, int groupmask, int bodyexclude,
                             float* dist, int32_t* geom, ss_stream s) {
  if (!B || !origin || !dir || !dist || nray <= 0) return ss_fail("ss_batch_rays: bad argument");
  cudaStream_t st = (cudaStream_t)s;
  if (prepare(B, st) != 0) return -1;
  const RayModel& r = B->model->rm;
  size_t smem = RAY_SMEM_BYTES(r);
  cudaFuncSetAttribute(rays_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  rays_kernel<<<B->nenv, 128, smem, st>>>(r, B->nenv, nray, B->ray_xf, origin, dir, groupmask, bodyexclude, dist, geom);
  B->launches++;
  CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int ss_batch_render_post(ss_batch* B, int cam, int W, int H, float fovy, uint8_t* rgb, float* depth, float depth_limit,
                                    int env_begin, int env_count, int rot90, int bgr, ss_stream s);
extern "C" int ss_batch_render(ss_batch* B, int cam, int W, int H, float fovy, uint8_t* rgb, float* depth, float depth_limit,
                               int env_begin, int env_count, ss_stream s) {
  return ss_batch_render_post(B, cam, W, H, fovy, rgb, depth, depth_limit, env_begin, env_count, 0, 0, s);
}
extern "C" int ss_batch_render_post(ss_batch* B, int cam, int W, int H, float fovy, uint8_t* rgb, float* depth, float depth_limit,
                                    int env_begin, int env_count, int rot90, int bgr, ss_stream s) {
  if (!B || W <= 0 || H <= 0 || (!rgb && !depth)) return ss_fail("ss_batch_render: bad argument");
  if (rot90 < -1 || rot90 > 1) return ss_fail("ss_batch_render_post: rot90 must be -1, 0 or 1 (numpy.rot90 k)");
  int post = (bgr ? 1 : 0) | ((rot90 == 1 ? 1 : rot90 == -1 ? 2 : 0) << 1);
  const RayModel& r = B->model->rm;
  if (r.present && (cam < 0 || cam >= r.ncam)) return ss_fail("ss_batch_render: camera %d out of range", cam);
  if (env_begin < 0 || env_count <= 0 || env_begin + env_count > B->nenv) return ss_fail("ss_batch_render: env range out of bounds");
  cudaStream_t st = (cudaStream_t)s;
  if (prepare(B, st, env_begin, env_count) != 0) return -1;   // geom transforms of the rendered envs only
  if (fovy <= 0) fovy = B->model->cam_fovy_host[cam];
  if (B->render_mode == 1 && r.nchunk > 0) {
    // raster path: envs in sub-chunks sharing one depth/id buffer (8 B per pixel)
    const size_t npix = (size_t)W * H;
    // sub-chunks of up to 256 envs (measured at 640x480, 4096 envs: 20 envs 61.4 ms, 40 envs 49.0 ms, 128 envs 39.5 ms, 256 envs 37.6 ms:
    // amortising the launch tails beats keeping the depth/id buffer L2-resident); at most 1 GiB of buffer
    int nsub = (int)std::max<size_t>(1, std::min<size_t>(256, ((size_t)1 << 30) / (npix * 8)));
    if (B->raster_nsub > 0) nsub = B->raster_nsub;   // SS_RASTER_NSUB, read once at batch creation (tuning knob; results do not depend on it)
    nsub = std::min(nsub, env_count);
    if (B->zbuf_cap < npix * nsub) {
      if (B->zbuf) { CUDA_OK(cudaStreamSynchronize(st)); cudaFree(B->zbuf); B->zbuf = nullptr; }
      CUDA_OK(cudaMalloc((void**)&B->zbuf, npix * nsub * 8));
      CUDA_OK(cudaMemsetAsync(B->zbuf, 0xff, npix * nsub * 8, st));
      B->zbuf_cap = npix * nsub;
    }
    if (B->rs_nsub < nsub) {
      if (B->rs_cam) { CUDA_OK(cudaStreamSynchronize(st)); cudaFree(B->rs_cam); cudaFree(B->rs_prim); cudaFree(B->rs_cvis); }
      CUDA_OK(cudaMalloc((void**)&B->rs_cam, (size_t)nsub * 16 * sizeof(float)));
      CUDA_OK(cudaMalloc((void**)&B->rs_prim, (size_t)nsub * (MAXPRIM + 1) * sizeof(int)));
      CUDA_OK(cudaMalloc((void**)&B->rs_cvis, (size_t)nsub * r.nchunk));
      if (B->rs_queue) { cudaFree(B->rs_queue); cudaFree(B->rs_qcount); }
      B->rs_qcap = B->raster_qcap > 0 ? B->raster_qcap : nsub * 4096;   // a full queue sends the box back to its owner warp (raster_tri_kernel)
      CUDA_OK(cudaMalloc((void**)&B->rs_queue, (size_t)B->rs_qcap * sizeof(RItem)));
      CUDA_OK(cudaMalloc((void**)&B->rs_qcount, sizeof(int)));
      B->rs_nsub = nsub;
    }
    for (int off = 0; off < env_count; off += nsub) {
      const int n = std::min(nsub, env_count - off), e0 = env_begin + off;
      raster_setup_kernel<<<n, 256, 0, st>>>(r, e0, cam, W, H, fovy, B->bufs.xpos, B->bufs.xquat, B->ray_xf, B->rs_cam, B->rs_prim, B->rs_cvis);
      unsigned long long* stats = nullptr;
      if (B->raster_stats) { cudaMalloc((void**)&stats, 16 * 8); cudaMemsetAsync(stats, 0, 16 * 8, st); }   // SS_RASTER_STATS=1: work counters (diagnostic, synchronises)
      CUDA_OK(cudaMemsetAsync(B->rs_qcount, 0, sizeof(int), st));
      raster_tri_kernel<<<dim3(r.nchunk, n), 128, 0, st>>>(r, e0, W, H, B->ray_xf, B->rs_cam, B->rs_cvis, B->zbuf, stats,
                                                           (RItem*)B->rs_queue, B->rs_qcount, B->rs_qcap);
      raster_large_kernel<<<4 * 148, 256, 0, st>>>(r, W, H, B->rs_cam, B->zbuf, (const RItem*)B->rs_queue, B->rs_qcount, B->rs_qcap);
      if (stats) {
        unsigned long long h[16];
        cudaStreamSynchronize(st); cudaMemcpy(h, stats, sizeof(h), cudaMemcpyDeviceToHost); cudaFree(stats);
        fprintf(stderr, "[raster] %d envs, %d chunks of %d in view; per env: %.0f triangles with a pixel box, %.0f box pixels | large: %.0f tris, %.0f px | near-crossing: %.0f tris, %.0f px | > 4096 px: %.1f tris, %.0f px\n",
                n, (int)(h[0] / n), r.nchunk, (double)h[1] / n, (double)h[2] / n, (double)h[3] / n, (double)h[4] / n, (double)h[5] / n, (double)h[6] / n, (double)h[7] / n, (double)h[8] / n);
      }
      {
        const dim3 rg((W + 31) / 32, (H + 31) / 32, n), rb(32, 8);
        const int nslot = (r.headlight_active ? 1 : 0) + std::min(r.nlight, MAXLIGHT - 1);
#define SS_RESOLVE(NS) raster_resolve_kernel<NS><<<rg, rb, 0, st>>>(r, e0, off, cam, W, H, B->bufs.xpos, B->bufs.xquat, B->ray_xf, B->rs_cam, B->rs_prim, \
                                                                   B->zbuf, rgb, depth, depth_limit, post)
        if (!rgb || nslot < 1 || nslot > 3) SS_RESOLVE(0);
        else if (nslot == 1) SS_RESOLVE(1);
        else if (nslot == 2) SS_RESOLVE(2);
        else SS_RESOLVE(3);
#undef SS_RESOLVE
      }
      B->launches += 4;
    }
    CUDA_OK(cudaGetLastError());
    return 0;
  }
  size_t smem = RAY_SMEM_BYTES(r) + (size_t)r.nraygeom * 3 * sizeof(int);
  cudaFuncSetAttribute(render_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid((W + TILE * TILES_PER_CTA - 1) / (TILE * TILES_PER_CTA), (H + TILE - 1) / TILE, env_count), block(TILE, TILE);
  if (grid.z > 65535) return ss_fail("ss_batch_render: at most 65535 envs per call");
  render_kernel<<<grid, block, smem, st>>>(r, env_begin, cam, W, H, fovy, B->bufs.xpos, B->bufs.xquat, B->ray_xf, rgb, depth, depth_limit, post);
  B->launches++;
  CUDA_OK(cudaGetLastError());
  return 0;
}

// ----------------------------------------------------------------------------- depth colour map (F3)
// utils.get_depth_color_map (stretch_mujoco/utils.py:363-373), the `use_depth_color_map` option of
// StatusStretchCameras.get_camera_data (datamodels/status_stretch_camera.py:47-82): per image,
// v = uint8((1 - (d - min) / (max - min)) * 255), then cv2.COLORMAP_JET (BGR).  One CTA per image: block-wide
// min / max, then the table lookup.  The table is cv2.applyColorMap(arange(256), COLORMAP_JET).
__constant__ unsigned char c_jet_bgr[768] = {128,0,0,132,0,0,136,0,0,140,0,0,144,0,0,148,0,0,152,0,0,156,0,0,160,0,0,164,0,0,168,0,0,172,0,0,176,0,0,180,0,0,184,0,0,188,0,0,192,0,0,196,0,0,200,0,0,204,0,0,208,0,0,212,0,0,216,0,0,220,0,0,224,0,0,228,0,0,232,0,0,236,0,0,240,0,0,244,0,0,248,0,0,252,0,0,255,0,0,255,4,0,255,8,0,255,12,0,255,16,0,255,20,0,255,24,0,255,28,0,255,32,0,255,36,0,255,40,0,255,44,0,255,48,0,255,52,0,255,56,0,255,60,0,255,64,0,255,68,0,255,72,0,255,76,0,255,80,0,255,84,0,255,88,0,255,92,0,255,96,0,255,100,0,255,104,0,255,108,0,255,112,0,255,116,0,255,120,0,255,124,0,255,128,0,255,132,0,255,136,0,255,140,0,255,144,0,255,148,0,255,152,0,255,156,0,255,160,0,255,164,0,255,168,0,255,172,0,255,176,0,255,180,0,255,184,0,255,188,0,255,192,0,255,196,0,255,200,0,255,204,0,255,208,0,255,212,0,255,216,0,255,220,0,255,224,0,255,228,0,255,232,0,255,236,0,255,240,0,255,244,0,255,248,0,255,252,0,254,255,2,250,255,6,246,255,10,242,255,14,238,255,18,234,255,22,230,255,26,226,255,30,222,255,34,218,255,38,214,255,42,210,255,46,206,255,50,202,255,54,198,255,58,194,255,62,190,255,66,186,255,70,182,255,74,178,255,78,174,255,82,170,255,86,166,255,90,162,255,94,158,255,98,154,255,102,150,255,106,146,255,110,142,255,114,138,255,118,134,255,122,130,255,126,126,255,130,122,255,134,118,255,138,114,255,142,110,255,146,106,255,150,102,255,154,98,255,158,94,255,162,90,255,166,86,255,170,82,255,174,78,255,178,74,255,182,70,255,186,66,255,190,62,255,194,58,255,198,54,255,202,50,255,206,46,255,210,42,255,214,38,255,218,34,255,222,30,255,226,26,255,230,22,255,234,18,255,238,14,255,242,10,255,246,6,255,250,1,255,254,0,252,255,0,248,255,0,244,255,0,240,255,0,236,255,0,232,255,0,228,255,0,224,255,0,220,255,0,216,255,0,212,255,0,208,255,0,204,255,0,200,255,0,196,255,0,192,255,0,188,255,0,184,255,0,180,255,0,176,255,0,172,255,0,168,255,0,164,255,0,160,255,0,156,255,0,152,255,0,148,255,0,144,255,0,140,255,0,136,255,0,132,255,0,128,255,0,124,255,0,120,255,0,116,255,0,112,255,0,108,255,0,104,255,0,100,255,0,96,255,0,92,255,0,88,255,0,84,255,0,80,255,0,76,255,0,72,255,0,68,255,0,64,255,0,60,255,0,56,255,0,52,255,0,48,255,0,44,255,0,40,255,0,36,255,0,32,255,0,28,255,0,24,255,0,20,255,0,16,255,0,12,255,0,8,255,0,4,255,0,0,255,0,0,252,0,0,248,0,0,244,0,0,240,0,0,236,0,0,232,0,0,228,0,0,224,0,0,220,0,0,216,0,0,212,0,0,208,0,0,204,0,0,200,0,0,196,0,0,192,0,0,188,0,0,184,0,0,180,0,0,176,0,0,172,0,0,168,0,0,164,0,0,160,0,0,156,0,0,152,0,0,148,0,0,144,0,0,140,0,0,136,0,0,132,0,0,128};

__global__ void depth_colormap_kernel(const float* __restrict__ depth, int npix, unsigned char* __restrict__ out) {
  const float* d = depth + (size_t)blockIdx.x * npix;
  unsigned char* o = out + (size_t)blockIdx.x * npix * 3;
  float lo = 3.4e38f, hi = -3.4e38f;
  for (int i = threadIdx.x; i < npix; i += blockDim.x) { float v = d[i]; lo = fminf(lo, v); hi = fmaxf(hi, v); }
  for (int s = 16; s > 0; s >>= 1) { lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, s)); hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, s)); }
  __shared__ float slo[32], shi[32];
  if ((threadIdx.x & 31) == 0) { slo[threadIdx.x >> 5] = lo; shi[threadIdx.x >> 5] = hi; }
  __syncthreads();
  lo = slo[0]; hi = shi[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); w++) { lo = fminf(lo, slo[w]); hi = fmaxf(hi, shi[w]); }
  float inv = 1.0f / (hi - lo);   // a constant image divides by zero, as the reference does (NaN -> 0 after the uint8 cast)
  for (int i = threadIdx.x; i < npix; i += blockDim.x) {
    float t = (1.0f - (d[i] - lo) * inv) * 255.0f;
    int v = (t == t) ? (int)t : 0;
    v = v < 0 ? 0 : (v > 255 ? 255 : v);
    o[3 * i] = c_jet_bgr[3 * v]; o[3 * i + 1] = c_jet_bgr[3 * v + 1]; o[3 * i + 2] = c_jet_bgr[3 * v + 2];
  }
}

extern "C" int ss_depth_colormap(const float* depth_dev, int nimg, int npix, uint8_t* bgr_dev, ss_stream s) {
  if (!depth_dev || !bgr_dev || nimg <= 0 || npix <= 0) return ss_fail("ss_depth_colormap: bad argument");
  depth_colormap_kernel<<<nimg, 256, 0, (cudaStream_t)s>>>(depth_dev, npix, bgr_dev);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return ss_fail("ss_depth_colormap: %s", cudaGetErrorString(e));
  return 0;
}
