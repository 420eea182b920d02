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

extern "C" __global__ void ss_physics_kernel(DevModel m, StepArgs a);

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

static int build_layout(DevModel& m) {
  EnvLayout& o = m.L;
  int off = 0;
  auto take = [&](int n) { int r = off; off += (n + 3) & ~3; return r; };
  int nv = m.nv, nb = m.nbody;
  o.ldm = nv | 1;
  o.ldj = nv | 1;
  o.qpos = take(m.nq); o.qvel = take(nv); o.ctrl = take(m.nu); o.warm = take(nv); o.qacc = take(nv);
  o.xpos = take(nb * 3); o.xquat = take(nb * 4); o.xmat = take(nb * 9); o.xipos = take(nb * 3); o.ximat = take(nb * 9);
  o.xanchor = take(m.njnt * 3); o.xaxis = take(m.njnt * 3);
  o.rootcom = take(m.nroot * 3); o.cinert = take(nb * 10); o.crb = take(nb * 10); o.cdof = take(nv * 6);
  o.cdofdot = take(nv * 6); o.cvel = take(nb * 6); o.cacc = take(nb * 6); o.cfrc = take(nb * 6);
  o.M = take(nv * o.ldm); o.H = take(nv * o.ldm);
  o.qfrc_smooth = take(nv); o.qacc_smooth = take(nv); o.qfrc_con = take(nv);
  o.actforce = take(m.nu); o.actlen = take(m.nu); o.actvel = take(m.nu);
  o.gpos = take(m.ncgeom * 3);
  o.con = take(m.maxcon * CON_STRIDE);
  o.J = take(m.maxcrow * o.ldj); o.tmpJ = take(6 * o.ldj);
  o.s_d1 = take(m.maxsimple); o.s_c1 = take(m.maxsimple); o.s_d2 = take(m.maxsimple); o.s_c2 = take(m.maxsimple);
  o.e_R = take(m.maxrow); o.e_D = take(m.maxrow); o.e_aref = take(m.maxrow); o.e_floss = take(m.maxrow);
  o.e_force = take(m.maxrow); o.e_jar = take(m.maxrow); o.e_jv = take(m.maxrow); o.e_type = take(m.maxrow);
  o.e_id = take(m.maxrow); o.e_state = take(m.maxrow);
  o.v_Ma = take(nv); o.v_grad = take(nv); o.v_search = take(nv); o.v_mv = take(nv); o.v_tmp = take(nv);
  o.total = off;
  return off;
}

extern "C" int ss_model_load_blob(const void* blob, size_t nbytes, int device, ss_model** out) {
  if (!blob || !out) return ss_fail("ss_model_load_blob: null argument");
  ss_model* M = new ss_model();
  M->blob.assign((const unsigned char*)blob, (const unsigned char*)blob + nbytes);
  ss_blob b;
  if (ss_blob_open(&b, M->blob.data(), M->blob.size()) != 0) { delete M; return ss_fail("not a stretchsim model blob"); }
  M->b = b;
  const int32_t* sz = ss_blob_i32(&b, "sizes");
  if (!sz) { delete M; return ss_fail("model blob has no 'sizes' array"); }
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
  std::vector<float> cg_size, cg_rb, cg_pos, cg_quat;
  auto cg = [&](int g) {
    if (cg_of[g] < 0) {
      cg_of[g] = (int)cg_geomid.size();
      cg_geomid.push_back(g); cg_type.push_back(gtype[g]); cg_body.push_back(gbody[g]); cg_data.push_back(gdata[g]);
      for (int k = 0; k < 3; k++) { cg_size.push_back(gsize[3 * g + k]); cg_pos.push_back(gpos[3 * g + k]); }
      for (int k = 0; k < 4; k++) cg_quat.push_back(gquat[4 * g + k]);
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

#define UP(field, vec) m.field = upload(M, vec)
  UP(body_parentid, parent); UP(body_rootidx, rootidx); UP(body_jntnum, i32(b, "body_jntnum")); UP(body_jntadr, i32(b, "body_jntadr"));
  UP(body_dofnum, dofnum); UP(body_dofadr, dofadr); UP(lvl_adr, lvl_adr); UP(lvl_body, lvl_body); UP(child_adr, child_adr);
  UP(child_list, child_list); UP(root_list, root_list); UP(body_dofmask, dofmask);
  UP(body_pos, f32(b, "body_pos")); UP(body_quat, f32(b, "body_quat")); UP(body_ipos, f32(b, "body_ipos"));
  UP(body_iquat, f32(b, "body_iquat")); UP(body_mass, f32(b, "body_mass")); UP(body_inertia, f32(b, "body_inertia"));
  UP(body_gravcomp, gravcomp); UP(body_invweight0, f32(b, "body_invweight0")); UP(body_subtreemass, f32(b, "body_subtreemass"));
  UP(jnt_type, jnt_type); UP(jnt_bodyid, i32(b, "jnt_bodyid")); UP(jnt_qposadr, jnt_qposadr); UP(jnt_dofadr, jnt_dofadr);
  UP(jnt_limited, jnt_limited); UP(limited_list, limited_list);
  UP(jnt_pos, f32(b, "jnt_pos")); UP(jnt_axis, f32(b, "jnt_axis")); UP(jnt_stiffness, f32(b, "jnt_stiffness"));
  UP(jnt_range, f32(b, "jnt_range")); UP(jnt_margin, f32(b, "jnt_margin")); UP(jnt_solref, f32(b, "jnt_solref"));
  UP(jnt_solimp, f32(b, "jnt_solimp")); UP(qpos_spring, f32(b, "qpos_spring"));
  M->qpos0_host = f32(b, "qpos0");
  UP(qpos0, M->qpos0_host);
  UP(dof_bodyid, i32(b, "dof_bodyid")); UP(dof_jntid, dof_jnt); UP(dof_parentid, dof_parent); UP(dof_qposadr, dof_qposadr);
  UP(floss_list, floss_list);
  UP(dof_armature, f32(b, "dof_armature")); UP(dof_damping, f32(b, "dof_damping")); UP(dof_frictionloss, floss);
  UP(dof_invweight0, f32(b, "dof_invweight0")); UP(dof_solref, f32(b, "dof_solref")); UP(dof_solimp, f32(b, "dof_solimp"));
  UP(cg_geomid, cg_geomid); UP(cg_type, cg_type); UP(cg_bodyid, cg_body); UP(cg_dataid, cg_data);
  UP(cg_size, cg_size); UP(cg_rbound, cg_rb); UP(cg_pos, cg_pos); UP(cg_quat, cg_quat);
  UP(pair_cg1, pc1); UP(pair_cg2, pc2); UP(pair_condim, i32(b, "pair_condim"));
  UP(pair_friction, f32(b, "pair_friction")); UP(pair_solref, f32(b, "pair_solref")); UP(pair_solimp, f32(b, "pair_solimp"));
  UP(pair_margin, f32(b, "pair_margin")); UP(pair_gap, f32(b, "pair_gap"));
  UP(mesh_hulladr, i32(b, "mesh_hulladr")); UP(mesh_hullnum, i32(b, "mesh_hullnum")); UP(hull_vert, hull4);
  UP(site_bodyid, i32(b, "site_bodyid")); UP(sensor_type, stype); UP(sensor_objid, i32(b, "sensor_objid"));
  UP(sensor_adr, i32(b, "sensor_adr")); UP(site_pos, f32(b, "site_pos")); UP(site_quat, f32(b, "site_quat"));
  UP(sensor_cutoff, f32(b, "sensor_cutoff"));
  UP(eq_obj1id, i32(b, "eq_obj1id")); UP(eq_obj2id, i32(b, "eq_obj2id")); UP(eq_active0, i32(b, "eq_active0"));
  UP(eq_data, f32(b, "eq_data")); UP(eq_solref, f32(b, "eq_solref")); UP(eq_solimp, f32(b, "eq_solimp"));
  UP(actuator_ctrllimited, i32(b, "actuator_ctrllimited")); UP(actuator_forcelimited, i32(b, "actuator_forcelimited"));
  UP(actuator_trntype, trntype); UP(actuator_trnid, trnid);
  UP(actuator_gainprm, f32(b, "actuator_gainprm")); UP(actuator_biasprm, f32(b, "actuator_biasprm"));
  UP(actuator_ctrlrange, f32(b, "actuator_ctrlrange")); UP(actuator_forcerange, f32(b, "actuator_forcerange"));
  UP(act_moment, moment); UP(actuator_gear, gear);
  UP(key_qpos, f32(b, "key_qpos")); UP(key_ctrl, f32(b, "key_ctrl"));
#undef UP
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

extern "C" int ss_model_set(ss_model* M, const char* field, const void* src, size_t bytes) {
  if (!M || !field || !src) return ss_fail("ss_model_set: null argument");
  cudaSetDevice(M->device);
  if (strcmp(field, "qpos0") == 0) {
    if (bytes != M->qpos0_host.size() * sizeof(double)) return ss_fail("ss_model_set(qpos0): expected %zu doubles", M->qpos0_host.size());
    for (size_t i = 0; i < M->qpos0_host.size(); i++) M->qpos0_host[i] = (float)((const double*)src)[i];
    CUDA_OK(cudaMemcpy((void*)M->dm.qpos0, M->qpos0_host.data(), M->qpos0_host.size() * sizeof(float), cudaMemcpyHostToDevice));
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
  m.maxcon = maxcon > 0 ? maxcon : 24;
  int nsimple = m.neq + m.nfloss + 2 * m.nlimited;
  m.maxsimple = nsimple;
  m.maxcrow = maxefc > 0 ? std::max(maxefc - nsimple, 6) : 96;
  m.maxrow = m.maxsimple + m.maxcrow;
  int floats = build_layout(m);
  B->smem_per_env = (size_t)floats * sizeof(float);
  cudaSetDevice(M->device);
  int max_smem = 0, sms = 0;
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, M->device);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, M->device);
  int wpb = (int)(max_smem / B->smem_per_env);
  if (wpb < 1) { delete B; return ss_fail("env working set (%zu B) exceeds shared memory (%d B)", B->smem_per_env, max_smem); }
  wpb = std::min(wpb, 8);
  B->warps_per_block = wpb;
  B->grid = std::min((nenv + wpb - 1) / wpb, sms);
  cudaError_t e = cudaFuncSetAttribute(ss_physics_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(wpb * B->smem_per_env));
  if (e != cudaSuccess) { delete B; return ss_fail("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); }
  *out = B;
  return 0;
}

extern "C" void ss_batch_free(ss_batch* B) {
  if (!B) return;
  if (B->ray_xf) { cudaSetDevice(B->model->device); cudaFree(B->ray_xf); }
  delete B;
}
extern "C" long ss_batch_launch_count(const ss_batch* B) { return B ? B->launches : 0; }

extern "C" int ss_batch_set_debug(ss_batch* B, const ss_debug_buffers* d) {
  if (!B) return ss_fail("ss_batch_set_debug: null batch");
  if (d) B->dbg = *d; else memset(&B->dbg, 0, sizeof(B->dbg));
  return 0;
}

static int launch_physics(ss_batch* B, int nsteps, int forward_only, ss_stream stream) {
  StepArgs a;
  memset(&a, 0, sizeof(a));
  const ss_buffers& f = B->bufs;
  a.nenv = B->nenv; a.nsteps = nsteps; a.forward_only = forward_only;
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
  size_t smem = B->warps_per_block * B->smem_per_env;
  ss_physics_kernel<<<B->grid, B->warps_per_block * 32, smem, (cudaStream_t)stream>>>(B->dm, a);
  B->launches++;
  CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int ss_batch_step(ss_batch* B, int nsteps, ss_stream stream) {
  if (!B || nsteps <= 0) return ss_fail("ss_batch_step: bad argument");
  return launch_physics(B, nsteps, 0, stream);
}
extern "C" int ss_batch_forward(ss_batch* B, ss_stream stream) {
  if (!B) return ss_fail("ss_batch_forward: null batch");
  return launch_physics(B, 1, 1, stream);
}
