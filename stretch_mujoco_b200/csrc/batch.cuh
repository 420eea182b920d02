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
};
