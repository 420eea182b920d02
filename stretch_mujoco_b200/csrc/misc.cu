// Small per-env marshalling kernels of libstretchsim: reset, status (P1), commands (P2).
#include "host.h"
#include <cstring>

#define CUDA_OK(x)                                                                      \
  do {                                                                                  \
    cudaError_t e_ = (x);                                                               \
    if (e_ != cudaSuccess) return ss_fail("%s: %s", #x, cudaGetErrorString(e_));        \
  } while (0)

// constants of the reference's status/command maths (stretch_mujoco/config.py:1-11)
#define WHEEL_R 0.0508f
#define WHEEL_L 0.3153f
#define GRIP_REAL_MIN (-0.376f)
#define GRIP_REAL_MAX 0.56f
#define GRIP_SIM_MIN (-0.02f)
#define GRIP_SIM_MAX 0.04f
#define BASE_X_VEL 0.3f
#define BASE_R_VEL 1.0f

__device__ __forceinline__ float map_ranges(float v, float a0, float a1, float b0, float b1) {
  return (v - a0) * (b1 - b0) / (a1 - a0) + b0;  // utils.map_between_ranges (stretch_mujoco/utils.py:352-360)
}

__global__ void reset_kernel(int nenv, int nq, int nv, int nu, const float* qpos_src, const float* ctrl_src,
                             const int32_t* mask, float* qpos, float* qvel, float* warm, float* time, float* ctrl,
                             int32_t* flags) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nenv || (mask && !mask[e])) return;
  for (int i = 0; i < nq; i++) qpos[(size_t)e * nq + i] = qpos_src[i];
  for (int i = 0; i < nv; i++) { qvel[(size_t)e * nv + i] = 0; warm[(size_t)e * nv + i] = 0; }
  if (ctrl_src) for (int i = 0; i < nu; i++) ctrl[(size_t)e * nu + i] = ctrl_src[i];
  if (time) time[e] = 0;
  if (flags) flags[e] = 0;
}

extern "C" int ss_batch_reset(ss_batch* B, const int32_t* mask, int key_id, ss_stream stream) {
  if (!B) return ss_fail("ss_batch_reset: null batch");
  const DevModel& m = B->dm;
  if (key_id >= m.nkey) return ss_fail("ss_batch_reset: keyframe %d out of range", key_id);
  cudaSetDevice(B->model->device);
  const float* qsrc = key_id < 0 ? B->model->qpos0_dev : m.key_qpos + (size_t)key_id * m.nq;
  const float* csrc = key_id < 0 ? nullptr : m.key_ctrl + (size_t)key_id * m.nu;
  reset_kernel<<<(B->nenv + 127) / 128, 128, 0, (cudaStream_t)stream>>>(B->nenv, m.nq, m.nv, m.nu, qsrc, csrc, mask,
                                                                          B->bufs.qpos, B->bufs.qvel, B->bufs.qacc_warmstart,
                                                                          B->bufs.time, B->bufs.ctrl, B->bufs.env_flags);
  B->launches++;
  CUDA_OK(cudaGetLastError());
  return 0;
}


__global__ void status_kernel(int nenv, int nu, int nbody, StatusIds ids, const float* time, const float* len,
                              const float* vel, const float* xpos, const float* xquat, float* out) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nenv) return;
  float* s = out + (size_t)e * SS_STATUS_WIDTH;
  s[0] = time ? time[e] : 0.f;
  for (int k = 0; k < 8; k++) {
    int a = ids.act[k];
    float p = a >= 0 ? len[(size_t)e * nu + a] : 0.f, v = a >= 0 ? vel[(size_t)e * nu + a] : 0.f;
    if (k == 7) p = map_ranges(p, GRIP_SIM_MIN, GRIP_SIM_MAX, GRIP_REAL_MIN, GRIP_REAL_MAX);
    s[1 + 2 * k] = p; s[2 + 2 * k] = v;
  }
  int b = ids.base_body;
  const float* q = xquat + ((size_t)e * nbody + b) * 4;
  float r00 = 1 - 2 * (q[2] * q[2] + q[3] * q[3]), r10 = 2 * (q[1] * q[2] + q[0] * q[3]);
  s[17] = xpos[((size_t)e * nbody + b) * 3]; s[18] = xpos[((size_t)e * nbody + b) * 3 + 1]; s[19] = atan2f(r10, r00);
  float wl = ids.act[8] >= 0 ? vel[(size_t)e * nu + ids.act[8]] : 0.f, wr = ids.act[9] >= 0 ? vel[(size_t)e * nu + ids.act[9]] : 0.f;
  s[20] = WHEEL_R * (wl + wr) * 0.5f;           // utils.diff_drive_fwd_kinematics (stretch_mujoco/utils.py:94-114)
  s[21] = WHEEL_R * (wr - wl) / WHEEL_L;
  s[22] = 0; s[23] = 0;
}

// actuator / body ids of the status and command rows: looked up by name once per batch (the reference does the
// same named lookups on every call, stretch_mujoco/mujoco_server.py:475-504)
static int lookup_ids(ss_batch* B, StatusIds* ids) {
  if (!B->ids_valid) {
    static const char* names[10] = {"lift", "arm", "head_pan", "head_tilt", "wrist_yaw", "wrist_pitch", "wrist_roll",
                                    "gripper", "left_wheel_vel", "right_wheel_vel"};
    for (int k = 0; k < 10; k++) B->ids.act[k] = ss_name2id(B->model, SS_OBJ_ACTUATOR, names[k]);
    B->ids.base_body = ss_name2id(B->model, SS_OBJ_BODY, "base_link");
    if (B->ids.base_body < 0) return ss_fail("model has no body named base_link");
    B->ids_valid = true;
  }
  *ids = B->ids;
  return 0;
}

extern "C" int ss_batch_pull_status(ss_batch* B, float* status_dev, ss_stream stream) {
  if (!B || !status_dev) return ss_fail("ss_batch_pull_status: null argument");
  if (!B->bufs.act_length || !B->bufs.act_velocity || !B->bufs.xpos || !B->bufs.xquat)
    return ss_fail("ss_batch_pull_status needs act_length, act_velocity, xpos and xquat buffers");
  StatusIds ids;
  if (lookup_ids(B, &ids) != 0) return -1;
  cudaSetDevice(B->model->device);
  status_kernel<<<(B->nenv + 127) / 128, 128, 0, (cudaStream_t)stream>>>(B->nenv, B->dm.nu, B->dm.nbody, ids, B->bufs.time,
                                                                          B->bufs.act_length, B->bufs.act_velocity,
                                                                          B->bufs.xpos, B->bufs.xquat, status_dev);
  B->launches++;
  CUDA_OK(cudaGetLastError());
  return 0;
}

// base_state: [0] mode (0 none, 1 translate_by, 2 rotate_by, 3 velocity) [1..3] start x,y,theta [4] increment [5] v [6] omega
__global__ void command_kernel(int nenv, int nu, int nbody, int nkey, StatusIds ids, const float* key_ctrl,
                               const float* len, const float* xpos, const float* xquat, float* cmd, float* bs, float* ctrl) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nenv) return;
  float* c = cmd + (size_t)e * SS_CMD_WIDTH;
  float* u = ctrl + (size_t)e * nu;
  float* st = bs + (size_t)e * 8;
  int b = ids.base_body;
  const float* q = xquat + ((size_t)e * nbody + b) * 4;
  float px = xpos[((size_t)e * nbody + b) * 3], py = xpos[((size_t)e * nbody + b) * 3 + 1];
  float th = atan2f(2 * (q[1] * q[2] + q[0] * q[3]), 1 - 2 * (q[2] * q[2] + q[3] * q[3]));
  // move_by (stretch_mujoco/mujoco_server.py:532-549)
  for (int k = 0; k < 10; k++) {
    if (c[20 + k] == 0.f) continue;
    c[20 + k] = 0.f;
    float inc = c[30 + k];
    if (k >= 8) { st[0] = (k == 8) ? 1.f : 2.f; st[1] = px; st[2] = py; st[3] = th; st[4] = inc; continue; }
    int a = ids.act[k];
    if (a < 0) continue;
    float cur = len[(size_t)e * nu + a];
    if (k == 7) u[a] = map_ranges(map_ranges(cur, GRIP_SIM_MIN, GRIP_SIM_MAX, GRIP_REAL_MIN, GRIP_REAL_MAX) + inc, GRIP_REAL_MIN,
                                  GRIP_REAL_MAX, GRIP_SIM_MIN, GRIP_SIM_MAX);
    else u[a] = cur + inc;
  }
  // move_to (mujoco_server.py:552-564); base slots are rejected on the host like the reference does
  for (int k = 0; k < 8; k++) {
    if (c[k] == 0.f) continue;
    c[k] = 0.f;
    int a = ids.act[k];
    if (a < 0) continue;
    u[a] = (k == 7) ? map_ranges(c[10 + k], GRIP_REAL_MIN, GRIP_REAL_MAX, GRIP_SIM_MIN, GRIP_SIM_MAX) : c[10 + k];
  }
  // set_base_velocity (mujoco_server.py:567-569)
  if (c[40] != 0.f) { c[40] = 0.f; st[0] = 3.f; st[1] = px; st[2] = py; st[3] = th; st[5] = c[41]; st[6] = c[42]; }
  // keyframe (mujoco_server.py:572-574)
  int key = (int)c[43];
  if (key > 0 && key <= nkey) { for (int i = 0; i < nu; i++) u[i] = key_ctrl[(size_t)(key - 1) * nu + i]; }
  c[43] = 0.f;
  // BaseController.update (mujoco_server.py:111-165)
  int mode = (int)st[0];
  float v = 0, w = 0;
  bool set = false;
  if (mode == 1) {
    float dx = px - st[1], dy = py - st[2];
    if (!(sqrtf(dx * dx + dy * dy) <= fabsf(st[4]))) { st[0] = 0.f; set = true; }
    else { v = BASE_X_VEL * (st[4] > 0 ? 1.f : -1.f); set = true; }
  } else if (mode == 2) {
    if (!(fabsf(st[3] - th) <= fabsf(st[4]))) { st[0] = 0.f; set = true; }
    else { w = BASE_R_VEL * (st[4] > 0 ? 1.f : -1.f); set = true; }
  } else if (mode == 3) {
    v = st[5]; w = st[6]; set = true;
  }
  if (set && ids.act[8] >= 0 && ids.act[9] >= 0) {
    // utils.diff_drive_inv_kinematics (stretch_mujoco/utils.py:117-135)
    u[ids.act[8]] = (v - w * WHEEL_L * 0.5f) / WHEEL_R;
    u[ids.act[9]] = (v + w * WHEEL_L * 0.5f) / WHEEL_R;
  }
}

extern "C" int ss_batch_apply_commands(ss_batch* B, float* command_dev, float* base_state_dev, ss_stream stream) {
  if (!B || !command_dev || !base_state_dev) return ss_fail("ss_batch_apply_commands: null argument");
  if (!B->bufs.act_length || !B->bufs.xpos || !B->bufs.xquat)
    return ss_fail("ss_batch_apply_commands needs act_length, xpos and xquat buffers");
  StatusIds ids;
  if (lookup_ids(B, &ids) != 0) return -1;
  cudaSetDevice(B->model->device);
  command_kernel<<<(B->nenv + 127) / 128, 128, 0, (cudaStream_t)stream>>>(B->nenv, B->dm.nu, B->dm.nbody, B->dm.nkey, ids,
                                                                           B->dm.key_ctrl, B->bufs.act_length, B->bufs.xpos,
                                                                           B->bufs.xquat, command_dev, base_state_dev,
                                                                           B->bufs.ctrl);
  B->launches++;
  CUDA_OK(cudaGetLastError());
  return 0;
}

// The reference's control cadence: push_command + BaseController.update run after EVERY mj_step
// (stretch_mujoco/mujoco_server.py:378-379,450-463), so the stop test of a base move_by is evaluated per step.
extern "C" int ss_batch_step_controlled(ss_batch* B, int nsteps, float* command_dev, float* base_state_dev, ss_stream stream) {
  if (!B || nsteps <= 0) return ss_fail("ss_batch_step_controlled: bad argument");
  for (int k = 0; k < nsteps; k++) {
    if (ss_batch_apply_commands(B, command_dev, base_state_dev, stream) != 0) return -1;
    if (ss_batch_step(B, 1, stream) != 0) return -1;
  }
  return 0;
}
