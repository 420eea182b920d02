// C ABI of libstretchsim (include/stretchsim.h): model upload, batch management, launches.
#include "../../include/stretchsim.h"
#include "../../include/ss_blob.h"
#include "model.cuh"
#include "batch.cuh"
#include "host.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

static thread_local std::string g_err;
int ss_fail(const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return -1;
}
extern "C" const char* ss_last_error(void) { return g_err.c_str(); }
extern "C" const char* ss_version(void) { return "stretchsim 0.1.0 (sm_100a)"; }

#define CUDA_OK(x)                                                                      \
  do {                                                                                  \
    cudaError_t e_ = (x);                                                               \
    if (e_ != cudaSuccess) return ss_fail("%s: %s", #x, cudaGetErrorString(e_));        \
  } while (0)

extern "C" __global__ void ss_smooth_kernel(DevModel m, StepArgs a);
extern "C" __global__ void ss_narrow_kernel(DevModel m, StepArgs a);
int ss_solve_tile(int nv);
cudaError_t ss_solve_set_smem(int tile, int bytes);
void ss_solve_launch(int tile, int grid, int block, size_t smem, cudaStream_t st, const DevModel& m, const StepArgs& a);
#define NP_SMEM 96   // floats of shared memory per warp of the narrowphase kernel (physics.cu)
#define NARROW_THREADS 256

// Env visiting order for the next physics launch: counting sort of the envs by the cost they reported in
// the LAST step of the previous launch (Newton iterations + narrowphase queries), heaviest first.  An env's
// cost is persistent over a few steps (step-to-step correlation 0.74) but not over a control period, so
// long rollouts are cut into short launches (ss_batch_step) and re-sorted in between.  The order only
// affects which warp simulates which env, never the results.
__global__ void schedule_kernel(int env0, int nenv, const int32_t* __restrict__ cost, int mode, int cost_scale, int32_t* __restrict__ order,
                                int32_t* __restrict__ work_counter) {   // envs [env0, env0 + nenv) -> order[0 .. nenv)
  cost += env0;
  __shared__ int hist[256], start[256];
  if (threadIdx.x < 4) work_counter[threadIdx.x] = 0;   // solve-kernel work counter, narrowphase item count / fetch counter, spare
  if (mode < 0) {  // SS_NOSORT=1: identity order (A/B knob)
    for (int e = threadIdx.x; e < nenv; e += blockDim.x) order[e] = env0 + e;
    return;
  }
  for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  for (int e = threadIdx.x; e < nenv; e += blockDim.x) atomicAdd(&hist[255 - min(255, cost_scale * cost[e])], 1);
  __syncthreads();
  if (threadIdx.x < 32) {   // exclusive scan of the 256 buckets by one warp (8 buckets per lane)
    int v[8], sum = 0;
    for (int k = 0; k < 8; k++) { v[k] = hist[threadIdx.x * 8 + k]; sum += v[k]; }
    int incl = sum;
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if ((int)threadIdx.x >= o) incl += t; }
    int acc = incl - sum;
    for (int k = 0; k < 8; k++) { start[threadIdx.x * 8 + k] = acc; acc += v[k]; }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < nenv; e += blockDim.x) order[atomicAdd(&start[255 - min(255, cost_scale * cost[e])], 1)] = env0 + e;
}

// ----------------------------------------------------------------------------- device upload helpers
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

// kernel: 1 = smooth kernel (persistent block + kinematics / dynamics scratch), 3 = solve kernel (persistent
// block + constraint rows + solver scratch).  The persistent block has the same offsets in both.
static int build_layout(DevModel& m, int kernel) {
  EnvLayout& o = m.L;
  memset(&o, 0, sizeof(o));
  int off = 0;
  auto take = [&](int n) { int r = off; off += (n + 3) & ~3; return r; };
  int nv = m.nv, nb = m.nbody;
  o.ldm = (nv + 3) & ~3;                           // float4 rows (register-tile Cholesky: one row per lane up to 32 dofs, two up to 64)
  if (nv > 32 && ((o.ldm >> 2) & 1) == 0) o.ldm += 4;   // odd number of 16-byte chunks per row: conflict-free row-per-lane float4 access
  o.ldj = (nv + 3) & ~3;   // 16-byte aligned rows (float4 operand loads in the Hessian build)
  // persistent block, part A (staged in shared memory by the solve kernel) ...
  o.qpos = take(m.nq); o.qvel = take(nv); o.warm = take(nv);
  o.cdof = take(nv * 6);
  o.M = take(nv * o.ldm);
  o.qfrc_smooth = take(nv); o.actforce = take(m.nu);
  o.pbA = off;
  // ... and part B (read from global memory where needed: narrowphase poses, contact Jacobians' reference points, IMU)
  o.ctrl = take(m.nu); o.actlen = take(m.nu); o.actvel = take(m.nu);
  o.xpos = take(nb * 3); o.xquat = take(nb * 4); o.cdofdot = take(nv * 6); o.cvel = take(nb * 6);
  o.gpos = take(m.ncgeom * 3);
  o.pb = off;
  if (kernel == 1) {
    o.xmat = take(nb * 9); o.cinert = take(nb * 10);
    o.cacc = take(nb * 10); o.cfrc = take(nb * 6);
    o.crb = o.cacc;                      // composite inertias (10/body) die before the RNE accelerations (6/body) are born
    // local offsets of the mass-matrix build (kinematics -> crb_mass_matrix) die before the RNE forces are born: same region
    if (nb * 3 + nv * 3 <= nb * 6) { o.dpos = o.cfrc; o.danchor = o.cfrc + nb * 3; }
    else { o.dpos = take(nb * 3); o.danchor = take(nv * 3); }
  } else {
    off = o.pbA;                         // part B stays in global memory: its offsets are only valid against the global block
    o.qacc = take(nv); o.qacc_smooth = take(nv); o.qfrc_con = take(nv);
    o.con = take(m.maxcon * CON_STRIDE);
    o.s_d1 = take(m.maxsimple); o.s_c1 = take(m.maxsimple); o.s_d2 = take(m.maxsimple); o.s_c2 = take(m.maxsimple);
    o.e_R = take(m.maxsimple); o.e_D = take(m.maxrow); o.e_aref = take(m.maxrow); o.e_floss = take(m.maxsimple); o.e_info = take(m.maxrow);   // R, friction loss: simple rows only
    o.J = take(std::max(m.maxjnz, nb * 6));
    o.cacc = o.J;                        // the IMU pass rebuilds body accelerations after the solve, when J is dead
    o.H = take(nv * o.ldm); o.tmpJ = take(3 * o.ldj);   // dense scratch rows of the Hessian build: two cone vectors + one expanded Jacobian row
    o.e_force = take(m.maxrow); o.e_jar = take(m.maxrow); o.e_jv = take(m.maxrow);
    o.v_Ma = take(nv); o.v_grad = take(nv); o.v_search = take(nv); o.v_mv = take(nv); o.v_tmp = take(nv);
  }
  o.total = off;
  return off;
}

// ----------------------------------------------------------------------------- model pack
struct PackBuilder {
  std::vector<uint32_t> w;
  int addi(const std::vector<int>& v) {
    int o = (int)w.size();
    for (int x : v) w.push_back((uint32_t)x);
    while (w.size() % 4) w.push_back(0);
    return o;
  }
  int addu(const std::vector<uint32_t>& v) {
    int o = (int)w.size();
    for (uint32_t x : v) w.push_back(x);
    while (w.size() % 4) w.push_back(0);
    return o;
  }
  int addf(const std::vector<float>& v) {
    int o = (int)w.size();
    for (float x : v) { uint32_t u; memcpy(&u, &x, 4); w.push_back(u); }
    while (w.size() % 4) w.push_back(0);
    return o;
  }
};

extern "C" int ss_model_load_blob(const void* blob, size_t nbytes, int device, ss_model** out) {
  if (!blob || !out) return ss_fail("ss_model_load_blob: null argument");
  ss_model* M = new ss_model();
  M->blob.assign((const unsigned char*)blob, (const unsigned char*)blob + nbytes);
  ss_blob b;
  if (ss_blob_open(&b, M->blob.data(), M->blob.size()) != 0) { delete M; return ss_fail("not a stretchsim model blob"); }
  M->b = b;
  const int32_t* sz = ss_blob_i32(&b, "sizes");
  if (!sz) { delete M; return ss_fail("model blob has no 'sizes' array"); }
  if (ss_blob_count(&b, "sizes") < 16) { delete M; return ss_fail("model blob: 'sizes' must hold 16 entries"); }
  {
    static const char* req_f64[] = {"opt_timestep", "opt_gravity", "opt_impratio", "opt_tolerance", "opt_ls_tolerance", "stat_meaninertia",
                                    "qpos0", "body_pos", "body_quat", "geom_size", "pair_margin", "hull_vert"};
    static const char* req_i32[] = {"opt_iterations", "opt_ls_iterations", "opt_cone", "body_parentid", "jnt_type", "dof_parentid",
                                    "geom_type", "pair_geom1", "pair_geom2", "mesh_hulladr", "mesh_hullnum"};
    for (const char* n : req_f64) if (!ss_blob_f64(&b, n)) { delete M; return ss_fail("model blob: missing or mistyped f64 array '%s'", n); }
    for (const char* n : req_i32) if (!ss_blob_i32(&b, n)) { delete M; return ss_fail("model blob: missing or mistyped i32 array '%s'", n); }
    if (ss_blob_count(&b, "opt_gravity") < 3) { delete M; return ss_fail("model blob: opt_gravity needs 3 entries"); }
  }
  M->device = device;
  if (cudaSetDevice(device) != cudaSuccess) { delete M; return ss_fail("cudaSetDevice(%d) failed", device); }
  DevModel& m = M->dm;
  memset(&m, 0, sizeof(m));
  m.nq = sz[0]; m.nv = sz[1]; m.nu = sz[2]; m.nbody = sz[3]; m.njnt = sz[4]; m.ngeom = sz[5]; m.nsite = sz[6];
  m.ncam = sz[7]; m.ntendon = sz[8]; m.neq = sz[9]; m.nsensor = sz[10]; m.nsensordata = sz[11]; m.nkey = sz[12];
  m.npair = sz[14]; m.nmesh = sz[15];
  M->dims = {sz[0], sz[1], sz[2], sz[3], sz[4], sz[5], sz[6], sz[7], sz[8], sz[9], sz[10], sz[11], sz[12], sz[13], sz[14], sz[15]};
  if (m.nv > SS_MAXNV || m.nbody > SS_MAXBODY) { delete M; return ss_fail("model too large: nv=%d nbody=%d", m.nv, m.nbody); }
  m.timestep = (float)ss_blob_f64(&b, "opt_timestep")[0];
  for (int k = 0; k < 3; k++) m.gravity[k] = (float)ss_blob_f64(&b, "opt_gravity")[k];
  m.impratio = (float)ss_blob_f64(&b, "opt_impratio")[0];
  m.tolerance = (float)ss_blob_f64(&b, "opt_tolerance")[0];
  m.ls_tolerance = (float)ss_blob_f64(&b, "opt_ls_tolerance")[0];
  m.meaninertia = (float)ss_blob_f64(&b, "stat_meaninertia")[0];
  m.iterations = ss_blob_i32(&b, "opt_iterations")[0];
  m.ls_iterations = ss_blob_i32(&b, "opt_ls_iterations")[0];
  if (ss_blob_i32(&b, "opt_cone")[0] != 1) { delete M; return ss_fail("only cone=elliptic is supported"); }

  int nb = m.nbody, nv = m.nv;
  std::vector<int> parent = i32(b, "body_parentid"), rootid = i32(b, "body_rootid"), dofnum = i32(b, "body_dofnum"),
                   dofadr = i32(b, "body_dofadr"), dof_parent = i32(b, "dof_parentid"), jnt_type = i32(b, "jnt_type"),
                   jnt_qposadr = i32(b, "jnt_qposadr"), dof_jnt = i32(b, "dof_jntid"), jnt_dofadr = i32(b, "jnt_dofadr");
  // tree levels, children, roots
  std::vector<int> depth(nb, 0), lvl_adr, lvl_body, child_adr(nb + 1, 0), child_list, root_list, rootidx(nb, 0);
  int nlevel = 0;
  for (int i = 1; i < nb; i++) { depth[i] = parent[i] == 0 ? 0 : depth[parent[i]] + 1; nlevel = std::max(nlevel, depth[i] + 1); }
  for (int l = 0; l < nlevel; l++) {
    lvl_adr.push_back((int)lvl_body.size());
    for (int i = 1; i < nb; i++) if (depth[i] == l) lvl_body.push_back(i);
  }
  lvl_adr.push_back((int)lvl_body.size());
  for (int i = 0; i < nb; i++) {
    child_adr[i] = (int)child_list.size();
    for (int c = 1; c < nb; c++) if (parent[c] == i && c != i) child_list.push_back(c);
  }
  child_adr[nb] = (int)child_list.size();
  for (int i = 1; i < nb; i++) if (parent[i] == 0) root_list.push_back(i);
  for (int i = 1; i < nb; i++) rootidx[i] = (int)(std::find(root_list.begin(), root_list.end(), rootid[i]) - root_list.begin());
  m.nlevel = nlevel; m.nroot = (int)root_list.size();
  // dof masks per body
  std::vector<uint32_t> dofmask(2 * nb, 0);
  std::vector<int> lastdof(nb, -1);
  for (int i = 1; i < nb; i++) lastdof[i] = dofnum[i] > 0 ? dofadr[i] + dofnum[i] - 1 : lastdof[parent[i]];
  for (int i = 1; i < nb; i++)
    for (int d = lastdof[i]; d >= 0; d = dof_parent[d]) dofmask[2 * i + (d >> 5)] |= 1u << (d & 31);
  std::vector<int> dof_qposadr(nv, 0);
  for (int d = 0; d < nv; d++) {
    int j = dof_jnt[d];
    dof_qposadr[d] = jnt_qposadr[j] + (jnt_type[j] >= JNT_SLIDE ? 0 : (d - jnt_dofadr[j]));
  }
  // friction-loss dofs, limited joints
  std::vector<float> floss = f32(b, "dof_frictionloss");
  std::vector<int> floss_list, limited_list, jnt_limited = i32(b, "jnt_limited");
  for (int d = 0; d < nv; d++) if (floss[d] > 0) floss_list.push_back(d);
  for (int j = 0; j < m.njnt; j++) if (jnt_limited[j] && jnt_type[j] >= JNT_SLIDE) limited_list.push_back(j);
  m.nfloss = (int)floss_list.size(); m.nlimited = (int)limited_list.size();
  // dense actuator moment (configuration independent for joint / fixed-tendon transmissions)
  std::vector<int> trntype = i32(b, "actuator_trntype"), trnid = i32(b, "actuator_trnid"), ten_adr = i32(b, "tendon_adr"),
                   ten_num = i32(b, "tendon_num"), wrap_obj = i32(b, "wrap_objid");
  std::vector<float> gear = f32(b, "actuator_gear"), wrap_prm = f32(b, "wrap_prm"), moment((size_t)m.nu * nv, 0.f);
  for (int a = 0; a < m.nu; a++) {
    if (trntype[a] == 0) moment[(size_t)a * nv + jnt_dofadr[trnid[a]]] = gear[a];
    else for (int w = ten_adr[trnid[a]]; w < ten_adr[trnid[a]] + ten_num[trnid[a]]; w++)
        moment[(size_t)a * nv + jnt_dofadr[wrap_obj[w]]] = gear[a] * wrap_prm[w];
  }
  // compact list of collision geoms
  std::vector<int> pg1 = i32(b, "pair_geom1"), pg2 = i32(b, "pair_geom2"), gtype = i32(b, "geom_type"),
                   gbody = i32(b, "geom_bodyid"), gdata = i32(b, "geom_dataid");
  std::vector<float> gsize = f32(b, "geom_size"), grb = f32(b, "geom_rbound"), gpos = f32(b, "geom_pos"), gquat = f32(b, "geom_quat");
  std::vector<int> cg_of(m.ngeom, -1), cg_geomid, cg_type, cg_body, cg_data, pc1, pc2;
  std::vector<float> cg_size, cg_rb, cg_pos, cg_quat, cg_aabb, gaabb = f32(b, "geom_aabb");
  auto cg = [&](int g) {
    if (cg_of[g] < 0) {
      cg_of[g] = (int)cg_geomid.size();
      cg_geomid.push_back(g); cg_type.push_back(gtype[g]); cg_body.push_back(gbody[g]); cg_data.push_back(gdata[g]);
      for (int k = 0; k < 3; k++) { cg_size.push_back(gsize[3 * g + k]); cg_pos.push_back(gpos[3 * g + k]); }
      for (int k = 0; k < 4; k++) cg_quat.push_back(gquat[4 * g + k]);
      for (int k = 0; k < 6; k++) cg_aabb.push_back(gaabb[6 * g + k]);
      cg_rb.push_back(grb[g]);
    }
    return cg_of[g];
  };
  for (int p = 0; p < m.npair; p++) { pc1.push_back(cg(pg1[p])); pc2.push_back(cg(pg2[p])); }
  m.ncgeom = (int)cg_geomid.size();
  std::vector<float> hv = f32(b, "hull_vert");
  std::vector<float4> hull4(hv.size() / 3);
  for (size_t i = 0; i < hull4.size(); i++) hull4[i] = make_float4(hv[3 * i], hv[3 * i + 1], hv[3 * i + 2], 0.f);
  std::vector<float> gravcomp = f32(b, "body_gravcomp");
  std::vector<int> stype = i32(b, "sensor_type");
  for (int s : stype) if (s == SENS_ACCEL) m.naccel++;
  M->nrange = 0;
  for (int s : stype) if (s == SENS_RANGE) M->nrange++;

  if (m.ncgeom > 65535) { delete M; return ss_fail("too many collision geoms"); }
  std::vector<int> pair_cg(m.npair);
  for (int p = 0; p < m.npair; p++) pair_cg[p] = pc1[p] | (pc2[p] << 16);
  std::vector<float> pmargin = f32(b, "pair_margin");
  m.max_margin = 0.f;
  for (float x : pmargin) m.max_margin = std::max(m.max_margin, x);
  PackBuilder P;
  PackOffsets& k = m.pk;
  // ---- part 1: tables the solve kernel needs (it stages only [0, nwords3) of the pack) ----
  k.body_parentid = P.addi(parent); k.body_rootidx = P.addi(rootidx);
  k.body_dofnum = P.addi(dofnum); k.body_dofadr = P.addi(dofadr);
  k.lvl_adr = P.addi(lvl_adr); k.lvl_body = P.addi(lvl_body);
  k.root_list = P.addi(root_list); k.body_dofmask = P.addu(dofmask);
  k.body_invweight0 = P.addf(f32(b, "body_invweight0"));
  k.jnt_type = P.addi(jnt_type); k.jnt_qposadr = P.addi(jnt_qposadr);
  k.jnt_dofadr = P.addi(jnt_dofadr); k.limited_list = P.addi(limited_list);
  k.jnt_range = P.addf(f32(b, "jnt_range")); k.jnt_margin = P.addf(f32(b, "jnt_margin")); k.jnt_solref = P.addf(f32(b, "jnt_solref"));
  k.jnt_solimp = P.addf(f32(b, "jnt_solimp"));
  M->qpos0_host = f32(b, "qpos0");
  k.qpos0 = P.addf(M->qpos0_host);
  k.dof_bodyid = P.addi(i32(b, "dof_bodyid")); k.floss_list = P.addi(floss_list);
  k.dof_damping = P.addf(f32(b, "dof_damping")); k.dof_frictionloss = P.addf(floss);
  k.dof_invweight0 = P.addf(f32(b, "dof_invweight0")); k.dof_solref = P.addf(f32(b, "dof_solref")); k.dof_solimp = P.addf(f32(b, "dof_solimp"));
  k.cg_geomid = P.addi(cg_geomid); k.cg_bodyid = P.addi(cg_body);
  // only the sites that sensors of the physics kernels use (IMU); lidar sites live in the ray model
  {
    std::vector<int> sobj = i32(b, "sensor_objid"), sadr = i32(b, "sensor_adr"), sbody = i32(b, "site_bodyid");
    std::vector<float> spos = f32(b, "site_pos"), squat = f32(b, "site_quat");
    std::vector<int> t2, o2, a2, sb2; std::vector<float> sp2, sq2;
    for (size_t s = 0; s < stype.size(); s++) {
      if (stype[s] == SENS_RANGE) continue;
      int site = sobj[s];
      t2.push_back(stype[s]); o2.push_back((int)sb2.size()); a2.push_back(sadr[s]);
      sb2.push_back(sbody[site]);
      for (int q = 0; q < 3; q++) sp2.push_back(spos[3 * site + q]);
      for (int q = 0; q < 4; q++) sq2.push_back(squat[4 * site + q]);
    }
    m.nsensor = (int)t2.size();
    k.sensor_type = P.addi(t2); k.sensor_objid = P.addi(o2); k.sensor_adr = P.addi(a2);
    k.site_bodyid = P.addi(sb2); k.site_pos = P.addf(sp2); k.site_quat = P.addf(sq2);
  }
  k.eq_obj1id = P.addi(i32(b, "eq_obj1id")); k.eq_obj2id = P.addi(i32(b, "eq_obj2id")); k.eq_active0 = P.addi(i32(b, "eq_active0"));
  k.eq_data = P.addf(f32(b, "eq_data")); k.eq_solref = P.addf(f32(b, "eq_solref")); k.eq_solimp = P.addf(f32(b, "eq_solimp"));
  k.actuator_forcelimited = P.addi(i32(b, "actuator_forcelimited"));
  k.actuator_biasprm = P.addf(f32(b, "actuator_biasprm")); k.actuator_forcerange = P.addf(f32(b, "actuator_forcerange"));
  k.act_moment = P.addf(moment);
  k.nwords3 = (int)P.w.size();
  // ---- part 2: kinematics, dynamics and collision tables (smooth and narrowphase kernels) ----
  k.body_jntnum = P.addi(i32(b, "body_jntnum")); k.body_jntadr = P.addi(i32(b, "body_jntadr"));
  k.child_adr = P.addi(child_adr); k.child_list = P.addi(child_list);
  k.body_pos = P.addf(f32(b, "body_pos")); k.body_quat = P.addf(f32(b, "body_quat")); k.body_ipos = P.addf(f32(b, "body_ipos"));
  k.body_iquat = P.addf(f32(b, "body_iquat")); k.body_mass = P.addf(f32(b, "body_mass")); k.body_inertia = P.addf(f32(b, "body_inertia"));
  k.body_gravcomp = P.addf(gravcomp);
  k.jnt_bodyid = P.addi(i32(b, "jnt_bodyid"));
  k.jnt_pos = P.addf(f32(b, "jnt_pos")); k.jnt_axis = P.addf(f32(b, "jnt_axis")); k.jnt_stiffness = P.addf(f32(b, "jnt_stiffness"));
  k.qpos_spring = P.addf(f32(b, "qpos_spring"));
  k.dof_jntid = P.addi(dof_jnt); k.dof_parentid = P.addi(dof_parent); k.dof_qposadr = P.addi(dof_qposadr);
  k.dof_armature = P.addf(f32(b, "dof_armature"));
  k.cg_type = P.addi(cg_type); k.cg_dataid = P.addi(cg_data);
  k.cg_size = P.addf(cg_size); k.cg_rbound = P.addf(cg_rb); k.cg_pos = P.addf(cg_pos); k.cg_quat = P.addf(cg_quat); k.cg_aabb = P.addf(cg_aabb);
  k.pair_cg = P.addi(pair_cg);
  k.mesh_hulladr = P.addi(i32(b, "mesh_hulladr")); k.mesh_hullnum = P.addi(i32(b, "mesh_hullnum"));
  k.actuator_ctrllimited = P.addi(i32(b, "actuator_ctrllimited"));
  k.actuator_gainprm = P.addf(f32(b, "actuator_gainprm")); k.actuator_ctrlrange = P.addf(f32(b, "actuator_ctrlrange"));
  k.nwords = (int)P.w.size();
  M->pack_host = P.w;
  m.pack = upload(M, P.w);
  m.pair_condim = upload(M, i32(b, "pair_condim"));
  m.pair_friction = upload(M, f32(b, "pair_friction")); m.pair_solref = upload(M, f32(b, "pair_solref"));
  m.pair_solimp = upload(M, f32(b, "pair_solimp")); m.pair_margin = upload(M, pmargin); m.pair_gap = upload(M, f32(b, "pair_gap"));
  m.hull_vert = upload(M, hull4);
  if (ss_blob_find(&b, "hull_edgeadr")) { m.hull_edgeadr = upload(M, i32(b, "hull_edgeadr")); m.hull_edge = upload(M, i32(b, "hull_edge")); }
  m.multiccd = ss_blob_find(&b, "opt_multiccd") ? ss_blob_i32(&b, "opt_multiccd")[0] : 0;
  m.key_qpos = upload(M, f32(b, "key_qpos")); m.key_ctrl = upload(M, f32(b, "key_ctrl"));
  M->qpos0_dev = upload(M, M->qpos0_host);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { ss_model_free(M); return ss_fail("model upload failed: %s", cudaGetErrorString(e)); }
  if (ss_rays_model_init(M) != 0) { ss_model_free(M); return -1; }
  *out = M;
  return 0;
}

extern "C" void ss_model_free(ss_model* M) {
  if (!M) return;
  cudaSetDevice(M->device);
  for (void* p : M->dev_allocs) cudaFree(p);
  delete M;
}

extern "C" int ss_model_dims(const ss_model* M, ss_dims* out) {
  if (!M || !out) return ss_fail("ss_model_dims: null argument");
  *out = M->dims;
  return 0;
}

extern "C" int ss_name2id(const ss_model* M, int objtype, const char* name) {
  if (!M || !name) return -1;
  return ss_blob_name2id(&M->b, (uint32_t)objtype, name);
}
extern "C" const char* ss_id2name(const ss_model* M, int objtype, int id) {
  if (!M) return nullptr;
  return ss_blob_name(&M->b, (uint32_t)objtype, id);
}

extern "C" long ss_model_get(const ss_model* M, const char* field, void* dst, size_t bytes) {
  if (!M || !field) return ss_fail("ss_model_get: null argument");
  if (strcmp(field, "qpos0") == 0 && dst) {  // reflects ss_model_set
    size_t n = M->qpos0_host.size() * sizeof(double);
    if (bytes < n) return ss_fail("ss_model_get(qpos0): buffer too small");
    for (size_t i = 0; i < M->qpos0_host.size(); i++) ((double*)dst)[i] = M->qpos0_host[i];
    return (long)n;
  }
  const ss_blob_entry* e = ss_blob_find(&M->b, field);
  if (!e) return ss_fail("ss_model_get: unknown field '%s'", field);
  if (!dst) return (long)e->nbytes;
  if (bytes < e->nbytes) return ss_fail("ss_model_get(%s): buffer too small (%zu < %zu)", field, bytes, (size_t)e->nbytes);
  memcpy(dst, M->b.base + e->offset, e->nbytes);
  return (long)e->nbytes;
}

extern "C" int ss_model_field_info(const ss_model* M, const char* field, int* dtype, int* ndim, uint64_t shape[4]) {
  if (!M || !field) return ss_fail("ss_model_field_info: null argument");
  const ss_blob_entry* e = ss_blob_find(&M->b, field);
  if (!e) return ss_fail("ss_model_field_info: unknown field '%s'", field);
  if (dtype) *dtype = (int)e->dtype;
  if (ndim) *ndim = (int)e->ndim;
  if (shape) for (int k = 0; k < 4; k++) shape[k] = e->shape[k];
  return 0;
}

extern "C" int ss_model_set(ss_model* M, const char* field, const void* src, size_t bytes) {
  if (!M || !field || !src) return ss_fail("ss_model_set: null argument");
  cudaSetDevice(M->device);
  if (strcmp(field, "qpos0") == 0) {
    if (bytes != M->qpos0_host.size() * sizeof(double)) return ss_fail("ss_model_set(qpos0): expected %zu doubles", M->qpos0_host.size());
    for (size_t i = 0; i < M->qpos0_host.size(); i++) M->qpos0_host[i] = (float)((const double*)src)[i];
    CUDA_OK(cudaMemcpy((void*)(M->dm.pack + M->dm.pk.qpos0), M->qpos0_host.data(), M->qpos0_host.size() * sizeof(float), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy((void*)M->qpos0_dev, M->qpos0_host.data(), M->qpos0_host.size() * sizeof(float), cudaMemcpyHostToDevice));
    return 0;
  }
  if (strcmp(field, "opt_iterations") == 0 && bytes == sizeof(int)) { M->dm.iterations = *(const int*)src; return 0; }
  if (strcmp(field, "opt_tolerance") == 0 && bytes == sizeof(double)) { M->dm.tolerance = (float)*(const double*)src; return 0; }
  if (strcmp(field, "cam_fovy") == 0) return ss_rays_set_fovy(M, (const double*)src, bytes);
  return ss_fail("ss_model_set: field '%s' is not settable", field);
}

// ----------------------------------------------------------------------------- batch
extern "C" int ss_batch_create(const ss_model* M, int nenv, int maxcon, int maxefc, const ss_buffers* bufs, ss_batch** out) {
  if (!M || !bufs || !out || nenv <= 0) return ss_fail("ss_batch_create: bad argument");
  if (!bufs->qpos || !bufs->qvel || !bufs->qacc_warmstart || !bufs->ctrl) return ss_fail("ss_batch_create: state buffers (qpos,qvel,qacc_warmstart,ctrl) are required");
  ss_batch* B = new ss_batch();
  B->model = M; B->nenv = nenv; B->bufs = *bufs; B->launches = 0;
  memset(&B->dbg, 0, sizeof(B->dbg));
  B->dm = M->dm;
  DevModel& m = B->dm;
  m.maxcon = maxcon > 0 ? maxcon : 32;
  int nsimple = m.neq + m.nfloss + 2 * m.nlimited;
  m.maxsimple = nsimple;
  m.maxcrow = maxefc > 0 ? std::max(maxefc - nsimple, 6) : 160;
  m.maxrow = m.maxsimple + m.maxcrow;
  // packed contact Jacobian: a row holds only the dofs between the two bodies' chains (wheel on floor: 7, wrist on base: ~10, free
  // object on table: 6), 14 floats per row of capacity cover the measured workloads (an overflow drops the contact and
  // sets env_flags bit 1, like the row capacity), but any single contact (6 rows of nv) must fit
  m.maxjnz = std::max(m.maxcrow * 14, 6 * ((m.nv + 3) & ~3));
  if (const char* e = getenv("SS_MAXJNZ")) m.maxjnz = std::max(m.maxjnz / 4, atoi(e));
  B->dm1 = m;
  B->smem_per_env = (size_t)build_layout(m, 3) * sizeof(float);
  B->smem_per_env1 = (size_t)build_layout(B->dm1, 1) * sizeof(float);
  B->pb_stride = m.L.pb;
  cudaSetDevice(M->device);
  int max_smem = 0, sms = 0;
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, M->device);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, M->device);
  size_t pack_bytes = (size_t)m.pk.nwords * 4 + 1024;  // + static shared (mbarrier, compiler scratch) slack
  size_t pack_bytes3 = (size_t)m.pk.nwords3 * 4 + 1024;
  int wpb = max_smem > (int)pack_bytes3 ? (int)((max_smem - pack_bytes3) / B->smem_per_env) : 0;
  int wpb1 = max_smem > (int)pack_bytes ? (int)((max_smem - pack_bytes) / B->smem_per_env1) : 0;
  if (wpb < 1 || wpb1 < 1) { delete B; return ss_fail("env working set (%zu B + %zu B model pack) exceeds shared memory (%d B)", B->smem_per_env, pack_bytes, max_smem); }
  wpb = std::min(wpb, 8);     // ss_solve_kernel: __launch_bounds__(256, 1), up to 255 registers per thread
  wpb1 = std::min(wpb1, 16);  // ss_smooth_kernel: __launch_bounds__(512, 1)
  // small batches (strong scaling: 512 envs per GPU) spread over all SMs instead of filling a few CTAs: a step's latency
  // is one warp's dependent chain whatever the CTA size, and fewer warps per CTA wait for fewer neighbours
  {
    const int per_sm = (nenv + sms - 1) / sms;
    wpb = std::max(std::min(wpb, 2), std::min(wpb, per_sm));
    wpb1 = std::max(std::min(wpb1, 2), std::min(wpb1, per_sm));
  }
  if (const char* e = getenv("SS_WPB")) wpb = std::max(1, std::min(wpb, atoi(e)));   // tuning knobs (A/B experiments; results never depend on them)
  if (const char* e = getenv("SS_WPB1")) wpb1 = std::max(1, std::min(wpb1, atoi(e)));
  B->sync_level = 9;   // 1: stage barriers in the smooth kernel, 8: lockstep Newton loop in the solve kernel (physics.cu)
  if (const char* e = getenv("SS_SYNC")) B->sync_level = atoi(e);
  B->group_warps = 0;
  B->pack_bytes = (size_t)m.pk.nwords * 4;
  B->warps_per_block = wpb; B->warps_per_block1 = wpb1;
  if (getenv("SS_VERBOSE"))
    fprintf(stderr, "[stretchsim] pack %zu B (solve kernel: %d B), smooth kernel %zu B/env x %d warps, solve kernel %zu B/env x %d warps (persistent block %d floats, part A %d), shared memory limit %d B\n",
            B->pack_bytes, m.pk.nwords3 * 4, B->smem_per_env1, wpb1, B->smem_per_env, wpb, m.L.pb, m.L.pbA, max_smem);
  B->grid = std::min((nenv + wpb - 1) / wpb, sms);
  B->grid1 = std::min((nenv + wpb1 - 1) / wpb1, sms);
  B->grid2 = 2 * sms;   // ss_narrow_kernel: __launch_bounds__(256, 2) (three CTAs of 80 registers per SM measured slower: 45.9 vs 43.4 ms per bench step)
  // pipeline scratch: persistent blocks, broadphase slots, narrowphase records, work-item queue
  B->maxslot = std::max(32, 3 * m.maxcon);   // broadphase candidates per env (global scratch only)
  if (const char* e = getenv("SS_MAXSLOT")) B->maxslot = std::max(8, atoi(e));
  size_t nslot = (size_t)nenv * B->maxslot;
  if (cudaMalloc((void**)&B->pb, sizeof(float) * (size_t)nenv * B->pb_stride) != cudaSuccess ||
      cudaMalloc((void**)&B->rec, sizeof(float) * nslot * NP_REC) != cudaSuccess ||
      cudaMalloc((void**)&B->npass, sizeof(int32_t) * nenv) != cudaSuccess || cudaMalloc((void**)&B->slot_pair, sizeof(int32_t) * nslot) != cudaSuccess ||
      cudaMalloc((void**)&B->items, sizeof(int32_t) * nslot) != cudaSuccess || cudaMemset(B->npass, 0, sizeof(int32_t) * nenv) != cudaSuccess) {
    ss_batch_free(B);
    return ss_fail("ss_batch_create: pipeline scratch: %s", cudaGetErrorString(cudaGetLastError()));
  }
  if (cudaMalloc((void**)&B->order, sizeof(int32_t) * nenv) != cudaSuccess || cudaMalloc((void**)&B->cost, sizeof(int32_t) * nenv) != cudaSuccess ||
      cudaMalloc((void**)&B->work_counter, sizeof(int32_t) * 4 * SS_MAXSETS) != cudaSuccess || cudaMemset(B->cost, 0, sizeof(int32_t) * nenv) != cudaSuccess) {
    ss_batch_free(B);
    return ss_fail("ss_batch_create: schedule buffers: %s", cudaGetErrorString(cudaGetLastError()));
  }
  if (const char* e = getenv("SS_RENDER")) B->render_mode = strcmp(e, "raycast") == 0 ? 0 : 1;   // A/B knob: the ray-cast camera path
  if (const char* e = getenv("SS_RASTER_NSUB")) B->raster_nsub = std::max(1, atoi(e));
  B->raster_stats = getenv("SS_RASTER_STATS") != nullptr;
  if (const char* e = getenv("SS_RASTER_QCAP")) B->raster_qcap = std::max(1, atoi(e));
  B->nosort = getenv("SS_NOSORT") != nullptr;
  // schedule key: bucket = min(255, cost_scale * cost), cost = cost_w x Newton iterations + narrowphase queries
  // (47.3 ms per 50 steps for 16 / x1 against 47.9 ms for 8 / x4)
  B->cost_w = 16; B->cost_scale = 1;
  if (const char* e = getenv("SS_COSTW")) B->cost_w = atoi(e);
  if (const char* e = getenv("SS_COSTSCALE")) B->cost_scale = atoi(e);
  B->nsets = nenv >= 16 * sms ? 3 : nenv >= 8 * sms ? 2 : 1;   // measured at 4096 envs (round 2, three-kernel pipeline): 1 set 45.9 ms, 2 sets 43.2 ms, 3 sets 41.6 ms, 4 sets 41.9 ms per 50 steps
  if (const char* e = getenv("SS_SETS")) B->nsets = std::max(1, std::min(SS_MAXSETS, atoi(e)));
  if (B->nsets > 1) {
    bool ok = cudaEventCreateWithFlags(&B->ev_fork, cudaEventDisableTiming) == cudaSuccess;
    for (int k = 0; k < B->nsets && ok; k++)
      ok = cudaStreamCreateWithFlags(&B->side[k], cudaStreamNonBlocking) == cudaSuccess &&
           cudaEventCreateWithFlags(&B->ev_join[k], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) { ss_batch_free(B); return ss_fail("ss_batch_create: side streams: %s", cudaGetErrorString(cudaGetLastError())); }
  }
  *out = B;
  return 0;
}

extern "C" void ss_batch_free(ss_batch* B) {
  if (!B) return;
  cudaSetDevice(B->model->device);
  if (B->ray_xf) cudaFree(B->ray_xf);
  if (B->zbuf) cudaFree(B->zbuf);
  if (B->rs_cam) cudaFree(B->rs_cam);
  if (B->rs_prim) cudaFree(B->rs_prim);
  if (B->rs_cvis) cudaFree(B->rs_cvis);
  if (B->rs_queue) cudaFree(B->rs_queue);
  if (B->rs_qcount) cudaFree(B->rs_qcount);
  if (B->pb) cudaFree(B->pb);
  if (B->rec) cudaFree(B->rec);
  if (B->npass) cudaFree(B->npass);
  if (B->slot_pair) cudaFree(B->slot_pair);
  if (B->items) cudaFree(B->items);
  if (B->order) cudaFree(B->order);
  if (B->cost) cudaFree(B->cost);
  if (B->work_counter) cudaFree(B->work_counter);
  if (B->ev_fork) cudaEventDestroy(B->ev_fork);
  for (int k = 0; k < SS_MAXSETS; k++) {
    if (B->side[k]) { cudaStreamSynchronize(B->side[k]); cudaStreamDestroy(B->side[k]); }
    if (B->ev_join[k]) cudaEventDestroy(B->ev_join[k]);
  }
  delete B;
}
extern "C" long ss_batch_launch_count(const ss_batch* B) { return B ? B->launches : 0; }

extern "C" int ss_batch_set_debug(ss_batch* B, const ss_debug_buffers* d) {
  if (!B) return ss_fail("ss_batch_set_debug: null batch");
  if (d) B->dbg = *d; else memset(&B->dbg, 0, sizeof(B->dbg));
  return 0;
}

static int launch_physics(ss_batch* B, int nsteps, int forward_only, ss_stream stream, cudaEvent_t* marks = nullptr) {
  StepArgs a;
  memset(&a, 0, sizeof(a));
  const ss_buffers& f = B->bufs;
  a.nenv = B->nenv; a.nsteps = nsteps; a.forward_only = forward_only; a.sync_level = B->sync_level; a.group_warps = B->group_warps; a.cost_w = B->cost_w;
  a.qpos = f.qpos; a.qvel = f.qvel; a.warm = f.qacc_warmstart; a.time = f.time; a.ctrl = f.ctrl;
  a.xpos = f.xpos; a.xquat = f.xquat; a.act_length = f.act_length; a.act_velocity = f.act_velocity;
  a.sensordata = f.sensordata; a.qacc = f.qacc; a.ncon = f.ncon; a.contact_geom = f.contact_geom;
  a.contact_dist = f.contact_dist; a.solver_iter = f.solver_iter; a.env_flags = f.env_flags;
  a.dbg_M = B->dbg.M; a.dbg_qacc_smooth = B->dbg.qacc_smooth; a.dbg_qfrc_smooth = B->dbg.qfrc_smooth;
  a.dbg_qfrc_constraint = B->dbg.qfrc_constraint; a.dbg_contact_pos = B->dbg.contact_pos;
  a.dbg_contact_normal = B->dbg.contact_normal; a.dbg_nefc = B->dbg.nefc;
  cudaSetDevice(B->model->device);
  B->dm.iterations = B->model->dm.iterations;  // runtime-settable solver options (ss_model_set)
  B->dm.tolerance = B->model->dm.tolerance;
  a.cost = B->cost;
  a.pb = B->pb; a.pb_stride = B->pb_stride; a.maxslot = B->maxslot; a.npass = B->npass; a.slot_pair = B->slot_pair; a.rec = B->rec;
  // dynamic shared memory is per-function state shared by all batches of the process: set it before every
  // launch chain (a batch created later with a smaller footprint must not lower it for this one)
  const size_t smem1 = B->pack_bytes + B->warps_per_block1 * B->smem_per_env1, smem3 = (size_t)B->dm.pk.nwords3 * 4 + B->warps_per_block * B->smem_per_env;
  const size_t smem2 = B->pack_bytes + (NARROW_THREADS / 32) * NP_SMEM * sizeof(float);
  CUDA_OK(cudaFuncSetAttribute(ss_smooth_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
  CUDA_OK(cudaFuncSetAttribute(ss_narrow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
  const int tile = ss_solve_tile(B->dm.nv);
  CUDA_OK(ss_solve_set_smem(tile, (int)smem3));
  // One mj_step = schedule_kernel (cost-sorted env order, counters) -> ss_smooth_kernel -> ss_narrow_kernel ->
  // ss_solve_kernel.  The env batch is cut into `nsets` contiguous sets whose launch chains run on library-owned
  // side streams, forked from and joined back into the caller's stream by events: one set's kernels fill the SMs
  // that the other set's tails leave idle.
  const int cost_scale = B->cost_scale;
  cudaStream_t user = (cudaStream_t)stream;
  int nsets = (forward_only || marks) ? 1 : B->nsets;   // single steps fork too: the sets' chains fill each other's kernel tails
  if (nsets > 1) {
    CUDA_OK(cudaEventRecord(B->ev_fork, user));
    for (int k = 0; k < nsets; k++) CUDA_OK(cudaStreamWaitEvent(B->side[k], B->ev_fork, 0));
  }
  a.nsteps = 1;
  for (int done = 0; done < nsteps; done++) {
    for (int k = 0; k < nsets; k++) {
      int e0 = (int)((long)B->nenv * k / nsets), e1 = (int)((long)B->nenv * (k + 1) / nsets);
      cudaStream_t st = nsets > 1 ? B->side[k] : user;
      a.nenv = e1 - e0; a.order = B->order + e0;
      a.work_counter = B->work_counter + 4 * k; a.item_count = a.work_counter + 1; a.item_next = a.work_counter + 2;
      a.items = B->items + (size_t)e0 * B->maxslot;
      schedule_kernel<<<1, 1024, 0, st>>>(e0, a.nenv, B->cost, B->nosort ? -1 : 0, cost_scale, B->order + e0, a.work_counter);
      int g1 = std::min((a.nenv + B->warps_per_block1 - 1) / B->warps_per_block1, B->grid1);
      int g3 = std::min((a.nenv + B->warps_per_block - 1) / B->warps_per_block, B->grid);
      if (marks) cudaEventRecord(marks[0], st);
      ss_smooth_kernel<<<g1, B->warps_per_block1 * 32, smem1, st>>>(B->dm1, a);
      if (marks) cudaEventRecord(marks[1], st);
      ss_narrow_kernel<<<B->grid2, NARROW_THREADS, smem2, st>>>(B->dm1, a);
      if (marks) cudaEventRecord(marks[2], st);
      ss_solve_launch(tile, g3, B->warps_per_block * 32, smem3, st, B->dm, a);
      if (marks) cudaEventRecord(marks[3], st);
      B->launches += 4;
    }
  }
  if (nsets > 1)
    for (int k = 0; k < nsets; k++) {
      CUDA_OK(cudaEventRecord(B->ev_join[k], B->side[k]));
      CUDA_OK(cudaStreamWaitEvent(user, B->ev_join[k], 0));
    }
  CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int ss_batch_step(ss_batch* B, int nsteps, ss_stream stream) {
  if (!B || nsteps <= 0) return ss_fail("ss_batch_step: bad argument");
  return launch_physics(B, nsteps, 0, stream);
}
// One mj_step with its three kernels serialised on the caller's stream (one env set) and CUDA events between them:
// per-kernel durations for the roofline line of bench.py.  The only entry point that synchronises (on its last event).
extern "C" int ss_batch_profile_step(ss_batch* B, float* ms3, ss_stream stream) {
  if (!B || !ms3) return ss_fail("ss_batch_profile_step: bad argument");
  cudaSetDevice(B->model->device);
  cudaEvent_t ev[4];
  for (int k = 0; k < 4; k++) CUDA_OK(cudaEventCreate(&ev[k]));
  int rc = launch_physics(B, 1, 0, stream, ev);
  if (rc == 0) {
    CUDA_OK(cudaEventSynchronize(ev[3]));
    for (int k = 0; k < 3; k++) CUDA_OK(cudaEventElapsedTime(&ms3[k], ev[k], ev[k + 1]));
  }
  for (int k = 0; k < 4; k++) cudaEventDestroy(ev[k]);
  return rc;
}
extern "C" int ss_batch_forward(ss_batch* B, ss_stream stream) {
  if (!B) return ss_fail("ss_batch_forward: null batch");
  return launch_physics(B, 1, 1, stream);
}
