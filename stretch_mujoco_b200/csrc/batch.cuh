// Per-launch arguments of the physics kernel (device pointers owned by the caller, see
// include/stretchsim.h ss_buffers / ss_debug_buffers).
#pragma once
#include <stdint.h>

struct StepArgs {
  int nenv, nsteps, forward_only, sync_level, group_warps, cost_w;
  float *qpos, *qvel, *warm, *time, *ctrl;
  float *xpos, *xquat, *act_length, *act_velocity, *sensordata, *qacc;
  int32_t *ncon, *contact_geom;
  float* contact_dist;
  int32_t *solver_iter, *env_flags;
  // debug taps
  float *dbg_M, *dbg_qacc_smooth, *dbg_qfrc_smooth, *dbg_qfrc_constraint, *dbg_contact_pos, *dbg_contact_normal;
  int32_t* dbg_nefc;
  // scheduling (library-owned): env visiting order (heaviest first), per-env cost of this launch, CTA work counter
  const int32_t* order;
  int32_t *cost, *work_counter;
  // pipeline scratch (library-owned)
  float* pb;               // [nenv, pb_stride] persistent block of every env (smooth kernel -> narrowphase / solve kernels)
  int pb_stride, maxslot;
  int32_t* npass;          // [nenv] candidate pairs that passed the broadphase
  int32_t* slot_pair;      // [nenv, maxslot] their pair ids, reference order
  float* rec;              // [nenv, maxslot, NP_REC] narrowphase records: count, then 8 floats per contact
  int32_t* items;          // work-item queue of the narrowphase kernel (env * maxslot + slot)
  int32_t *item_count, *item_next;
};
#define NP_REC 72
