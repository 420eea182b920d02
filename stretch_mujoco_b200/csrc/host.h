// Host-side handle definitions shared by the translation units of libstretchsim.
#pragma once
#include <cstdarg>
#include <string>
#include <vector>

#include "../../include/ss_blob.h"
#include "../../include/stretchsim.h"
#include "model.cuh"
#include "rays.cuh"

struct ss_model {
  std::vector<unsigned char> blob;  // host copy of the compiled model (names, fp64 masters)
  ss_blob b;
  ss_dims dims;
  int device;
  int nrange;
  DevModel dm;
  RayModel rm;
  std::vector<float> qpos0_host, cam_fovy_host;
  std::vector<uint32_t> pack_host;
  const float* qpos0_dev = nullptr;
  std::vector<void*> dev_allocs;
};

struct StatusIds { int act[10]; int base_body; };  // lift, arm, head_pan, head_tilt, wrist_yaw, wrist_pitch, wrist_roll, gripper, left_wheel, right_wheel

#define SS_MAXSETS 4
struct ss_batch {
  const ss_model* model;
  int nenv;
  ss_buffers bufs;
  ss_debug_buffers dbg;
  DevModel dm;   // model + this batch's buffer sizes and the SOLVE kernel's shared-memory layout
  DevModel dm1;  // same with the SMOOTH kernel's layout (also used by the narrowphase kernel: persistent block only)
  size_t smem_per_env, smem_per_env1, pack_bytes;
  int warps_per_block, warps_per_block1, grid, grid1, grid2, sync_level, group_warps;
  int maxslot = 0, pb_stride = 0;
  float *pb = nullptr, *rec = nullptr;
  int32_t *npass = nullptr, *slot_pair = nullptr, *items = nullptr, *counters = nullptr;   // counters: 4 ints per env set
  long launches;
  bool nosort = false;
  int cost_w = 16, cost_scale = 1;
  int nsets = 1;               // env sets with their own launch chains on side streams (tail overlap)
  cudaStream_t side[SS_MAXSETS] = {};
  cudaEvent_t ev_fork = nullptr, ev_join[SS_MAXSETS] = {};
  int32_t *order = nullptr, *cost = nullptr, *work_counter = nullptr;  // cost-sorted env schedule (api.cu)
  StatusIds ids = {};
  bool ids_valid = false;
  float* ray_xf = nullptr;  // [nenv, nraygeom, 12] world transforms of ray-visible geoms (library-owned scratch)
  // raster camera path (library-owned scratch, sized for one sub-chunk of envs)
  int render_mode = 1;                       // 1 = rasterised meshes + analytic primitives, 0 = ray casting (SS_RENDER=raycast)
  unsigned long long* zbuf = nullptr;        // [nsub, H, W] depth | triangle keys
  size_t zbuf_cap = 0;                       // entries
  float* rs_cam = nullptr;                   // [nsub, 16] eye, rotation, focal
  int* rs_prim = nullptr;                    // [nsub, 1 + MAXPRIM] camera-visible primitive geoms inside the image pyramid
  unsigned char* rs_cvis = nullptr;          // [nsub, nchunk] chunk inside the image pyramid?
  void* rs_queue = nullptr;                  // [rs_qcap] (triangle, row band) items with pixel boxes too large for their owner thread
  int* rs_qcount = nullptr;
  int rs_qcap = 0;
  int rs_nsub = 0;
  int raster_nsub = 0;                       // SS_RASTER_NSUB: envs per raster sub-chunk (0 = automatic)
  int raster_qcap = 0;                       // SS_RASTER_QCAP: queue capacity override (test of the overflow path)
  bool raster_stats = false;                 // SS_RASTER_STATS: print work counters per sub-chunk (diagnostic)
};

int ss_fail(const char* fmt, ...);
int ss_rays_model_init(ss_model* M);
int ss_rays_set_fovy(ss_model* M, const double* fovy, size_t bytes);
