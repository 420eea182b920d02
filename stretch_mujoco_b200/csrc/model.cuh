// Device-side model description shared by the kernels of libstretchsim.
// Everything here is read-only and identical for all envs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define SS_MAXNV 64
#define SS_MAXBODY 64

enum { JNT_FREE = 0, JNT_BALL = 1, JNT_SLIDE = 2, JNT_HINGE = 3 };
enum { GEOM_PLANE = 0, GEOM_HFIELD, GEOM_SPHERE, GEOM_CAPSULE, GEOM_ELLIPSOID, GEOM_CYLINDER, GEOM_BOX, GEOM_MESH };
enum { SENS_GYRO = 0, SENS_ACCEL = 1, SENS_RANGE = 2 };
enum { CNSTR_EQUALITY = 0, CNSTR_FRICTION, CNSTR_LIMIT, CNSTR_CONTACT_FRICTIONLESS, CNSTR_CONTACT_ELLIPTIC };
enum { ST_SATISFIED = 0, ST_QUADRATIC, ST_LINEARNEG, ST_LINEARPOS, ST_CONE };

// Shared-memory layout of one env (offsets in floats), computed on the host.
struct EnvLayout {
  int qpos, qvel, ctrl, warm, qacc;
  int xpos, xquat, xmat, xipos, ximat, xanchor, xaxis;
  int rootcom, cinert, crb, cdof, cdofdot, cvel, cacc, cfrc;
  int M, H, ldm;                 // dense nv x nv matrices with (odd) row stride ldm
  int qfrc_smooth, qacc_smooth, qfrc_con, actforce, actlen, actvel;
  int gpos;                      // world centres of collision geoms [ncgeom*3]
  int con;                       // contacts [maxcon * CON_STRIDE]
  int J, ldj, tmpJ;              // dense Jacobian of CONTACT rows [maxcrow * ldj]; scratch [6 * ldj]
  int s_d1, s_c1, s_d2, s_c2;    // simple rows (equality, friction loss, limit): <=2 non-zeros each
  int e_R, e_D, e_aref, e_floss, e_force, e_jar, e_jv, e_type, e_id, e_state;  // all rows [maxrow]
  int v_Ma, v_grad, v_search, v_mv, v_tmp;
  int total;
};

#define CON_STRIDE 36
// contact record fields (float slots)
#define C_POS 0
#define C_FRAME 3
#define C_DIST 12
#define C_MU 13
#define C_FRICTION 14
#define C_SOLREF 19
#define C_SOLIMP 21
#define C_DIM 26
#define C_GEOM1 27
#define C_GEOM2 28
#define C_EFC 29
#define C_INCLMARGIN 30
#define C_BODY1 31
#define C_BODY2 32

struct DevModel {
  int nq, nv, nu, nbody, njnt, ngeom, nsite, ncam, ntendon, neq, nsensor, nsensordata, nkey, npair, nmesh;
  int nlevel, nroot, ncgeom, nfloss, nlimited, ngravcomp, naccel;
  int maxcon, maxcrow, maxsimple, maxrow;
  float timestep, gravity[3], impratio, tolerance, ls_tolerance, meaninertia;
  int iterations, ls_iterations;
  EnvLayout L;
  // bodies
  const int *body_parentid, *body_rootidx, *body_jntnum, *body_jntadr, *body_dofnum, *body_dofadr, *lvl_adr, *lvl_body,
      *child_adr, *child_list, *root_list;
  const uint32_t* body_dofmask;  // [nbody][2]
  const float *body_pos, *body_quat, *body_ipos, *body_iquat, *body_mass, *body_inertia, *body_gravcomp,
      *body_invweight0, *body_subtreemass;
  // joints / dofs
  const int *jnt_type, *jnt_bodyid, *jnt_qposadr, *jnt_dofadr, *jnt_limited, *limited_list;
  const float *jnt_pos, *jnt_axis, *jnt_stiffness, *jnt_range, *jnt_margin, *jnt_solref, *jnt_solimp, *qpos0,
      *qpos_spring;
  const int *dof_bodyid, *dof_jntid, *dof_parentid, *dof_qposadr, *floss_list;
  const float *dof_armature, *dof_damping, *dof_frictionloss, *dof_invweight0, *dof_solref, *dof_solimp;
  // collision geoms (compacted list of geoms that appear in candidate pairs)
  const int *cg_geomid, *cg_type, *cg_bodyid, *cg_dataid;
  const float *cg_size, *cg_rbound, *cg_pos, *cg_quat;
  const int *pair_cg1, *pair_cg2, *pair_condim;
  const float *pair_friction, *pair_solref, *pair_solimp, *pair_margin, *pair_gap;
  const int *mesh_hulladr, *mesh_hullnum;
  const float4* hull_vert;
  // sites / sensors
  const int *site_bodyid, *sensor_type, *sensor_objid, *sensor_adr;
  const float *site_pos, *site_quat, *sensor_cutoff;
  // tendons / equality / actuators
  const int *eq_obj1id, *eq_obj2id, *eq_active0;
  const float *eq_data, *eq_solref, *eq_solimp;
  const int *actuator_ctrllimited, *actuator_forcelimited, *actuator_trntype, *actuator_trnid;
  const float *actuator_gainprm, *actuator_biasprm, *actuator_ctrlrange, *actuator_forcerange, *act_moment,
      *actuator_gear;
  const float *key_qpos, *key_ctrl;
};
