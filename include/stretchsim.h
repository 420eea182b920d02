/* libstretchsim -- C ABI of the B200-native batched Stretch simulation engine.
 *
 * This is the drop-in boundary of SURVEY.md §8(b)(ii): each entry point replaces one call
 * the reference makes into the un-vendored `mujoco` wheel.  Citations are relative to the
 * reference repository (hello-robot/stretch_mujoco).
 *
 * Conventions: plain pointers and sizes, no exceptions; every function returns 0 on success
 * and a negative code on error with a thread-local message in ss_last_error().  All per-env
 * device arrays are env-major `[nenv, n]`, fp32 (int32 for indices), owned by the CALLER
 * (torch allocates them and passes data_ptr()).  Work is ordered by the caller's stream: kernels run on it, or
 * (ss_batch_step on large batches) on library-owned side streams that are forked from and joined back into it with
 * events, so that everything enqueued on the caller's stream before / after a call happens before / after the
 * call's work.  No function synchronises the device or blocks the host.
 */
#ifndef STRETCHSIM_H
#define STRETCHSIM_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct ss_model ss_model;
typedef struct ss_batch ss_batch;
typedef void* ss_stream; /* cudaStream_t */

/* object types for ss_name2id / ss_id2name (replaces mujoco.mj_name2id / mj_id2name,
 * stretch_mujoco/mujoco_server.py:207,283) */
enum { SS_OBJ_BODY = 0, SS_OBJ_JOINT, SS_OBJ_GEOM, SS_OBJ_SITE, SS_OBJ_CAMERA, SS_OBJ_ACTUATOR, SS_OBJ_SENSOR,
       SS_OBJ_KEY, SS_OBJ_MESH, SS_OBJ_TENDON };

typedef struct {
  int nq, nv, nu, nbody, njnt, ngeom, nsite, ncam, ntendon, neq, nsensor, nsensordata, nkey, nM, npair, nmesh;
} ss_dims;

/* ---- model ------------------------------------------------------------------------------
 * Replaces MjModel.from_xml_path / from_xml_string (stretch_mujoco/mujoco_server.py:252,
 * stretch_mujoco/robocasa_gen.py:232).  The XML front-end is host Python
 * (stretch_mujoco_b200/mjcf.py + compiler.py); it hands the compiled model over as a blob. */
int ss_model_load_blob(const void* blob, size_t nbytes, int device, ss_model** out);
void ss_model_free(ss_model*);
int ss_model_dims(const ss_model*, ss_dims* out);
int ss_name2id(const ss_model*, int objtype, const char* name);
const char* ss_id2name(const ss_model*, int objtype, int id);
/* named model arrays as fp64/int32 host copies: "qpos0", "jnt_range", "key_ctrl",
 * "actuator_ctrlrange", "opt_timestep", ... (replaces direct mjModel attribute reads,
 * stretch_mujoco/mujoco_server.py:213-222,284,381,574).  Returns bytes written or <0. */
long ss_model_get(const ss_model*, const char* field, void* dst, size_t bytes);
/* element type (0 = f64, 1 = i32, 2 = f32, 3 = u8), rank and shape of a named model array */
int ss_model_field_info(const ss_model*, const char* field, int* dtype, int* ndim, uint64_t shape[4]);
/* runtime-mutable fields: "qpos0" (start pose, mujoco_server.py:219-225), "cam_fovy"
 * (mujoco_server_camera_manager.py:197-208), "opt_iterations", "opt_tolerance". */
int ss_model_set(ss_model*, const char* field, const void* src, size_t bytes);

/* ---- batch (replaces MjData, stretch_mujoco/mujoco_server.py:258) ------------------------ */
typedef struct {
  /* state, read and written by ss_batch_step */
  float* qpos;           /* [nenv, nq] */
  float* qvel;           /* [nenv, nv] */
  float* qacc_warmstart; /* [nenv, nv] */
  float* time;           /* [nenv] */
  /* input */
  float* ctrl;           /* [nenv, nu] */
  /* outputs of the last step of every ss_batch_step call (nullable) */
  float* xpos;           /* [nenv, nbody, 3] body frames: mjData.xpos  */
  float* xquat;          /* [nenv, nbody, 4]              mjData.xquat */
  float* act_length;     /* [nenv, nu]  mjData.actuator_length   (mujoco_server.py:475-499) */
  float* act_velocity;   /* [nenv, nu]  mjData.actuator_velocity (mujoco_server.py:476-504) */
  float* sensordata;     /* [nenv, nsensordata] gyro, accelerometer; rangefinders via ss_batch_lidar */
  float* qacc;           /* [nenv, nv] */
  int32_t* ncon;         /* [nenv] */
  int32_t* contact_geom; /* [nenv, maxcon, 2] (geom1, geom2) in reference order, -1 padded */
  float* contact_dist;   /* [nenv, maxcon] */
  int32_t* solver_iter;  /* [nenv] Newton iterations of the last step */
  int32_t* env_flags;    /* [nenv] bit0: env was reset by the bad-state guard (mj_checkPos/Vel/Acc),
                                   bit1: contact/constraint buffer overflow, sticky until reset */
} ss_buffers;

int ss_batch_create(const ss_model*, int nenv, int maxcon, int maxefc, const ss_buffers* bufs, ss_batch** out);
void ss_batch_free(ss_batch*);
/* qpos <- key_qpos[key] (or qpos0 when key<0), qvel = warmstart = time = 0, ctrl <- key_ctrl
 * for envs whose mask is non-zero (NULL mask = all). */
int ss_batch_reset(ss_batch*, const int32_t* env_mask_dev, int key_id, ss_stream);
/* nsteps x mj_step (stretch_mujoco/mujoco_server.py:378): S1..S10 of SURVEY.md §8(a) fused in one persistent
 * kernel per launch.  The call is cut into short launches whose env->warp assignment is re-sorted by the cost each
 * env reported in its last step (csrc/api.cu:launch_physics); results do not depend on that policy.  ctrl is
 * read at every step from the ctrl buffer as it is when the call's work executes. */
int ss_batch_step(ss_batch*, int nsteps, ss_stream);
/* forward dynamics only (mj_forward): fills the output buffers, leaves the state untouched */
int ss_batch_forward(ss_batch*, ss_stream);
/* measurement aid: one mj_step whose three kernels (smooth dynamics + broadphase, narrowphase, constraints + solver +
 * integration) run back to back on the caller's stream with CUDA events between them; ms3 (HOST) receives their
 * durations.  Advances the state like ss_batch_step(1); the only call that synchronises (on its own last event). */
int ss_batch_profile_step(ss_batch*, float* ms3, ss_stream);
/* number of kernels launched by this batch since creation (bench.py "gpu_launches") */
long ss_batch_launch_count(const ss_batch*);

/* ---- debug / parity taps (fp32 device arrays, nullable) ---------------------------------- */
typedef struct {
  float* M;               /* [nenv, nv, nv] */
  float* qacc_smooth;     /* [nenv, nv] */
  float* qfrc_smooth;     /* [nenv, nv] */
  float* qfrc_constraint; /* [nenv, nv] */
  int32_t* nefc;          /* [nenv] */
  float* contact_pos;     /* [nenv, maxcon, 3] */
  float* contact_normal;  /* [nenv, maxcon, 3] */
} ss_debug_buffers;
int ss_batch_set_debug(ss_batch*, const ss_debug_buffers*);

/* ---- status / command marshalling (rows P1, P2) ------------------------------------------
 * status[nenv, 24]: replaces MujocoServer.pull_status (stretch_mujoco/mujoco_server.py:465-515):
 *  0 time | 1..16 (pos,vel) of lift, arm, head_pan, head_tilt, wrist_yaw, wrist_pitch,
 *  wrist_roll, gripper (gripper pos mapped to the real range, config.py:4-5) |
 *  17 base.x 18 base.y 19 base.theta 20 base.x_vel 21 base.theta_vel | 22,23 reserved */
#define SS_STATUS_WIDTH 24
int ss_batch_pull_status(ss_batch*, float* status_dev, ss_stream);

/* command[nenv, SS_CMD_WIDTH] float: replaces MujocoServer.push_command + BaseController
 * (stretch_mujoco/mujoco_server.py:93-176,527-578).  Per env:
 *  [0..9]   move_to trigger mask per slot (non-zero = trigger), slots = lift, arm, head_pan,
 *           head_tilt, wrist_yaw, wrist_pitch, wrist_roll, gripper, base_translate, base_rotate
 *  [10..19] move_to positions   [20..29] move_by trigger  [30..39] move_by increments
 *  [40] base_velocity trigger [41] v_linear [42] omega   [43] keyframe id + 1 (0 = none)
 * Triggers are consumed (cleared) by the call, like the reference's edge-triggered flags.
 * base_state[nenv, 8] persists the BaseController between calls (mode, start x,y,theta,
 * target increment, v, omega). */
#define SS_CMD_WIDTH 44
int ss_batch_apply_commands(ss_batch*, float* command_dev, float* base_state_dev, ss_stream);

/* nsteps x { ss_batch_apply_commands; ss_batch_step(1) }: the reference's cadence, where push_command and
 * BaseController.update run after every mj_step (stretch_mujoco/mujoco_server.py:378-379,450-463). */
int ss_batch_step_controlled(ss_batch*, int nsteps, float* command_dev, float* base_state_dev, ss_stream);

/* ---- sensors ---------------------------------------------------------------------------- */
/* 2-D spinning lidar (row S2/L1): evaluates every <rangefinder> of the model for each env from
 * the body frames of the last step; out[nenv, nrange] in sensor order, -1 on miss, clamped to
 * cutoff (stretch.xml:278-280,540; mujoco_server_sensor_manager.py:77-83).  If out is NULL the
 * values are written into the sensordata buffer at the sensors' addresses. */
int ss_batch_lidar(ss_batch*, float* out_dev, ss_stream);
int ss_model_num_rangefinders(const ss_model*);
/* generic rays: origin/dir [nenv, nray, 3] world frame; dist_out [nenv, nray]; geom_out nullable */
int ss_batch_rays(ss_batch*, int nray, const float* origin_dev, const float* dir_dev, int groupmask, int bodyexclude,
                  float* dist_out, int32_t* geom_out, ss_stream);
/* pinhole camera (rows C1-C4): replaces mujoco.Renderer.update_scene + render
 * (stretch_mujoco/mujoco_server_camera_manager.py:108-137).  rgb [nenv,H,W,3] uint8 top-left
 * origin, depth [nenv,H,W] float metres along the optical axis; either may be NULL.
 * depth_limit > 0 applies utils.limit_depth_distance (stretch_mujoco/utils.py:87-91). */
int ss_batch_render(ss_batch*, int cam_id, int W, int H, float fovy_deg, uint8_t* rgb_dev, float* depth_dev,
                    float depth_limit, int env_begin, int env_count, ss_stream);
/* same, with the client-side post-processing of StatusStretchCameras.get_camera_data
 * (stretch_mujoco/datamodels/status_stretch_camera.py:47-82) fused into the kernel's epilogue:
 * rot90 = k of numpy.rot90 (-1 for the d435i cameras, +1 for the nav camera, 0 none; rotated outputs are
 * [nenv,W,H,3] / [nenv,W,H]), bgr != 0 = cv2.COLOR_RGB2BGR channel order. */
int ss_batch_render_post(ss_batch*, int cam_id, int W, int H, float fovy_deg, uint8_t* rgb_dev, float* depth_dev,
                         float depth_limit, int env_begin, int env_count, int rot90, int bgr, ss_stream);

/* utils.get_depth_color_map (stretch_mujoco/utils.py:363-373; `use_depth_color_map` of
 * StatusStretchCameras.get_camera_data): depth [nimg, npix] -> JET colour map, BGR uint8 [nimg, npix, 3],
 * normalised per image with its own min / max. */
int ss_depth_colormap(const float* depth_dev, int nimg, int npix, uint8_t* bgr_dev, ss_stream);

const char* ss_last_error(void);
const char* ss_version(void);
#ifdef __cplusplus
}
#endif
#endif
