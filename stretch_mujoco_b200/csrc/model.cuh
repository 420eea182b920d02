// Device-side model description shared by the kernels of libstretchsim.
// Everything here is read-only and identical for all envs.
//
// The small per-body / joint / dof / geom / pair tables are packed into ONE contiguous device
// buffer (the "model pack").  The physics kernel pulls the pack into shared memory once per CTA
// with a TMA bulk copy; device code addresses it through the PKF()/PKI() macros (word offsets in
// DevModel::pk).  Large or rarely used tables (hull vertices, per-pair contact parameters,
// keyframes) stay in global memory behind pointers.
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

// Shared-memory layout of one env (offsets in floats from the env's slice), computed on the host
// (api.cu:build_layout), one layout per kernel.  All layouts start with the same "persistent block"
// [0, pb): the fields that travel from the smooth kernel to the narrowphase and solve kernels through
// global memory (one contiguous float4 copy per env).
struct EnvLayout {
  // persistent block, part A [0, pbA): staged in shared memory by the solve kernel
  int qpos, qvel, warm;
  int cdof;
  int M, ldm;                                    // dense nv x nv joint-space inertia, row stride ldm
  int qfrc_smooth, actforce;
  int pbA;
  // part B [pbA, pb): the solve kernel reads it from the global block (Jacobian reference points, IMU), the
  // narrowphase kernel too (poses)
  int ctrl, actlen, actvel;
  int xpos, xquat, cdofdot, cvel;
  int gpos;                                      // world positions of the collision geoms
  int pb;                                        // size of the persistent block (multiple of 4)
  // smooth kernel only
  int xmat, cinert, crb, cfrc;
  int dpos, danchor;                             // body origin relative to the parent's origin; hinge anchors relative to the body origin
  int cacc;                                      // smooth kernel: RNE accelerations; solve kernel: rebuilt with qacc for the IMU
  // solve kernel only
  int qacc, qacc_smooth, qfrc_con;
  int con;                                       // contacts [maxcon * CON_STRIDE]
  int s_d1, s_c1, s_d2, s_c2;                    // simple rows (equality, friction loss, limit): <=2 non-zeros
  int e_R, e_D, e_aref, e_floss, e_info;         // all rows [maxrow]; e_info packs type | state<<4 | id<<8
  int J, ldj;                                    // packed Jacobian of the CONTACT rows [maxjnz] (C_JOFS / C_MASK*); ldj = nv rounded up to 4 (dense scratch rows)
  int H, tmpJ, e_force, e_jar, e_jv, v_Ma, v_grad, v_search, v_mv, v_tmp;
  int total;
};

#define CON_STRIDE 28
// contact record fields (float slots); friction/solref/solimp are looked up through the pair id
#define C_POS 0
#define C_FRAME 3
#define C_DIST 12
#define C_MU 13
#define C_DIM 14
#define C_PAIR 15
#define C_EFC 16
#define C_BODY1 17
#define C_BODY2 18
#define C_FRICTION 19   // friction[5]: tangent1, tangent2, torsional, rolling1, rolling2
// packed Jacobian rows of the contact: the dofs on exactly one of the two bodies' chains (bit i of the 64-bit mask), values
// in dof order, row stride C_JW (popcount rounded up to 4), first row at J + C_JOFS; C_CMASK: 4-column chunks with a dof
#define C_MASKLO 24
#define C_MASKHI 25
#define C_JOFS 26
#define C_CMASK 27

// word offsets into the model pack
struct PackOffsets {
  int body_parentid, body_rootidx, body_jntnum, body_jntadr, body_dofnum, body_dofadr, lvl_adr, lvl_body, child_adr,
      child_list, root_list, body_dofmask;
  int body_pos, body_quat, body_ipos, body_iquat, body_mass, body_inertia, body_gravcomp, body_invweight0;
  int jnt_type, jnt_bodyid, jnt_qposadr, jnt_dofadr, limited_list;
  int jnt_pos, jnt_axis, jnt_stiffness, jnt_range, jnt_margin, jnt_solref, jnt_solimp, qpos0, qpos_spring;
  int dof_bodyid, dof_jntid, dof_parentid, dof_qposadr, floss_list;
  int dof_armature, dof_damping, dof_frictionloss, dof_invweight0, dof_solref, dof_solimp;
  int cg_geomid, cg_type, cg_bodyid, cg_dataid, cg_size, cg_rbound, cg_pos, cg_quat, cg_aabb, pair_cg;
  int mesh_hulladr, mesh_hullnum;
  int site_bodyid, site_pos, site_quat, sensor_type, sensor_objid, sensor_adr;
  int eq_obj1id, eq_obj2id, eq_active0, eq_data, eq_solref, eq_solimp;
  int actuator_ctrllimited, actuator_forcelimited, actuator_gainprm, actuator_biasprm, actuator_ctrlrange,
      actuator_forcerange, act_moment;
  int nwords;   // pack size in 32-bit words (multiple of 4)
  int nwords3;  // leading part of the pack that the solve kernel stages (tables of the later part are not used there)
};

struct DevModel {
  int nq, nv, nu, nbody, njnt, ngeom, nsite, ncam, ntendon, neq, nsensor, nsensordata, nkey, npair, nmesh;
  int nlevel, nroot, ncgeom, nfloss, nlimited, naccel;
  int maxcon, maxcrow, maxsimple, maxrow, maxjnz;
  float timestep, gravity[3], impratio, tolerance, ls_tolerance, meaninertia, max_margin;
  int iterations, ls_iterations;
  EnvLayout L;
  PackOffsets pk;
  const uint32_t* pack;  // device copy of the model pack
  // global-memory tables
  const int* pair_condim;
  const float *pair_friction, *pair_solref, *pair_solimp, *pair_margin, *pair_gap;
  const float4* hull_vert;
  const int *hull_edgeadr, *hull_edge;   // hull vertex adjacency (nullptr when the blob carries none)
  int multiccd;                          // <flag multiccd> (stretch.xml:8)
  const float *key_qpos, *key_ctrl;
};
