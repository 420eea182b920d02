/* CPU oracle (fp64) -- TEST INFRASTRUCTURE ONLY.
 *
 * A from-scratch restatement of the per-step path the reference runs through the
 * un-vendored dependency mujoco==3.2.6 (`pyproject.toml:12`): the single call
 * `mujoco.mj_step(model, data)` at `stretch_mujoco/mujoco_server.py:378`, the rangefinder
 * sensors evaluated inside it (`stretch_mujoco/models/stretch.xml:278-280,540`) and the
 * camera render at `stretch_mujoco/mujoco_server_camera_manager.py:135-137`.
 *
 * PARITY UNPINNED: the reference's tests hold no golden vectors for this path and the
 * dependency cannot be imported here; the oracle is anchored on the
 * reference's printed outputs (tests/test_kat_reference.py) and on analytic checks.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.  The product path (stretch_mujoco_b200/csrc) never links it.
 */
#ifndef SS_ORACLE_H
#define SS_ORACLE_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct om_model om_model;

om_model* om_model_load(const void* blob, size_t nbytes);
void om_model_free(om_model*);
/* out[16]: nq,nv,nu,nbody,njnt,ngeom,nsite,ncam,ntendon,neq,nsensor,nsensordata,nkey,nM,npair,nmesh */
void om_model_sizes(const om_model*, int* out);
void om_set_options(om_model*, int max_iter, double tolerance, int enable_lidar);

/* Per-env outputs (any pointer may be NULL). All arrays env-major [nenv, n], fp64 / int32. */
typedef struct {
  double* xpos;            /* [nenv, nbody, 3] */
  double* xquat;           /* [nenv, nbody, 4] */
  double* act_length;      /* [nenv, nu] */
  double* act_velocity;    /* [nenv, nu] */
  double* sensordata;      /* [nenv, nsensordata] */
  double* qacc;            /* [nenv, nv] */
  double* qfrc_constraint; /* [nenv, nv] */
  int* ncon;               /* [nenv] */
  int* contact_geom;       /* [nenv, maxcon, 2] */
  double* contact_dist;    /* [nenv, maxcon] */
  double* contact_pos;     /* [nenv, maxcon, 3] */
  double* contact_frame;   /* [nenv, maxcon, 3] normal only */
  int maxcon;
  int* solver_iter;        /* [nenv] */
  int* nefc;               /* [nenv] */
  int* flags;              /* [nenv] bit0: reset by bad-state guard, bit1: contact overflow */
  double* M;               /* [nenv, nv, nv] dense mass matrix (debug) */
  double* qacc_smooth;     /* [nenv, nv] */
  double* qfrc_bias;       /* [nenv, nv] */
  double* qfrc_passive;    /* [nenv, nv] */
  double* qfrc_actuator;   /* [nenv, nv] */
  double* efc_J;           /* [nenv, maxefc_out, nv] (debug) */
  double* efc_aref;        /* [nenv, maxefc_out] */
  double* efc_D;           /* [nenv, maxefc_out] */
  double* efc_force;       /* [nenv, maxefc_out] */
  int maxefc_out;
} om_outputs;

/* Advance nsteps. State arrays are updated in place. ctrl is either [nenv,nu] (held) or, when
 * ctrl_per_step!=0, [nsteps,nenv,nu].  nthreads<=0 -> OpenMP default. */
int om_batch_step(const om_model*, int nenv, int nsteps, double* qpos, double* qvel, const double* ctrl,
                  int ctrl_per_step, double* qacc_warmstart, double* time, om_outputs* out, int nthreads);

/* forward only (no integration): fills outputs for the given state */
int om_batch_forward(const om_model*, int nenv, const double* qpos, const double* qvel, const double* ctrl,
                     const double* qacc_warmstart, om_outputs* out, int nthreads);

/* rays: nearest hit over all ray-visible geoms. geomgroup_mask: bit g set = group g visible
 * (0 = all groups); bodyexclude -1 = none.  out_dist -1 on miss; out_geom may be NULL. */
int om_batch_rays(const om_model*, int nenv, const double* xpos, const double* xquat, int nray,
                  const double* origin /*[nenv,nray,3]*/, const double* dir /*[nenv,nray,3]*/, int groupmask,
                  int bodyexclude, double* out_dist, int* out_geom, int nthreads);

/* camera: pinhole RGB (uint8 [nenv,H,W,3], top-left origin) + planar depth (float [nenv,H,W]) */
int om_batch_render(const om_model*, int nenv, const double* xpos, const double* xquat, int cam_id, int W, int H,
                    double fovy_deg, unsigned char* rgb, float* depth, int nthreads);

int om_max_threads(void);
#ifdef __cplusplus
}
#endif
#endif
