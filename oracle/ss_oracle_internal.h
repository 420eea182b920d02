/* CPU oracle internals -- TEST INFRASTRUCTURE ONLY (see ss_oracle.h). */
#ifndef SS_ORACLE_INTERNAL_H
#define SS_ORACLE_INTERNAL_H
#include <math.h>
#include <string.h>

#define OM_MAXNV 64
#define OM_MAXCON 64
#define OM_MAXEFC 320
#define OM_MINVAL 1e-15
#define OM_MAXVAL 1e10
#define OM_MINIMP 0.0001
#define OM_MAXIMP 0.9999

enum { JNT_FREE = 0, JNT_BALL = 1, JNT_SLIDE = 2, JNT_HINGE = 3 };
enum { GEOM_PLANE = 0, GEOM_HFIELD, GEOM_SPHERE, GEOM_CAPSULE, GEOM_ELLIPSOID, GEOM_CYLINDER, GEOM_BOX, GEOM_MESH };
enum { SENS_GYRO = 0, SENS_ACCEL = 1, SENS_RANGE = 2 };
enum { CNSTR_EQUALITY = 0, CNSTR_FRICTION, CNSTR_LIMIT, CNSTR_CONTACT_FRICTIONLESS, CNSTR_CONTACT_ELLIPTIC };
enum { ST_SATISFIED = 0, ST_QUADRATIC, ST_LINEARNEG, ST_LINEARPOS, ST_CONE };

struct om_model {
  int nq, nv, nu, nbody, njnt, ngeom, nsite, ncam, ntendon, neq, nsensor, nsensordata, nkey, nM, npair, nmesh;
  double timestep, gravity[3], impratio, tolerance, ls_tolerance, meaninertia, extent;
  int iterations, ls_iterations, cone, enable_lidar;
  int *body_parentid, *body_rootid, *body_weldid, *body_jntnum, *body_jntadr, *body_dofnum, *body_dofadr, *body_lastdof;
  double *body_pos, *body_quat, *body_ipos, *body_iquat, *body_mass, *body_inertia, *body_gravcomp, *body_invweight0,
      *body_subtreemass;
  int *jnt_type, *jnt_bodyid, *jnt_qposadr, *jnt_dofadr, *jnt_limited;
  double *jnt_pos, *jnt_axis, *jnt_stiffness, *jnt_range, *jnt_margin, *jnt_solref, *jnt_solimp, *qpos0, *qpos_spring;
  int *dof_bodyid, *dof_jntid, *dof_parentid;
  double *dof_armature, *dof_damping, *dof_frictionloss, *dof_invweight0, *dof_solref, *dof_solimp;
  int *geom_type, *geom_bodyid, *geom_dataid, *geom_group, *geom_matid;
  double *geom_size, *geom_rbound, *geom_pos, *geom_quat, *geom_rgba, *geom_aabb;
  int* site_bodyid; double *site_pos, *site_quat;
  int* cam_bodyid; double *cam_pos, *cam_quat, *cam_fovy;
  int *tendon_adr, *tendon_num, *wrap_objid; double* wrap_prm;
  int *eq_obj1id, *eq_obj2id, *eq_active0; double *eq_data, *eq_solref, *eq_solimp;
  int *actuator_trntype, *actuator_trnid, *actuator_ctrllimited, *actuator_forcelimited;
  double *actuator_gear, *actuator_gainprm, *actuator_biasprm, *actuator_ctrlrange, *actuator_forcerange;
  int *sensor_type, *sensor_objid, *sensor_adr; double* sensor_cutoff;
  int *pair_geom1, *pair_geom2, *pair_condim;
  double *pair_friction, *pair_solref, *pair_solimp, *pair_margin, *pair_gap;
  int *mesh_hulladr, *mesh_hullnum; double* hull_vert;
  int *hull_edgeadr, *hull_edge, multiccd;   /* hull adjacency (NULL in blobs that predate it), <flag multiccd> */
  double *key_qpos, *key_ctrl;
  /* ray geometry */
  int has_ray, nraygeom, nsky, nlight, headlight_active;
  int *rmesh_vertadr, *rmesh_faceadr, *rmesh_facenum, *rmesh_bvhadr, *rmesh_face, *bvh_child, *raygeom_id;
  float *rmesh_vert, *bvh_aabb;
  double *geom_shade, *vis_headlight, *vis_map, *skybox_rgb;
  double* geom_tex;                 /* [ngeom,4] texture index (-1 none), texrepeat x/y, texuniform (NULL: blob without textures) */
  int *tex_adr, *tex_w, *tex_h, ntex;
  unsigned char* tex_rgb;
  float* rmesh_uv;                  /* [nface,6] UV pairs per triangle (NULL: none) */
  int *light_bodyid, *light_directional;
  double *light_pos, *light_dir, *light_ambient, *light_diffuse, *light_specular;
};

typedef struct {
  double dist, pos[3], frame[9], friction[5], solref[2], solimp[5], mu, includemargin;
  int dim, geom1, geom2, efc_address;
} om_contact;

typedef struct om_data {
  /* state (external storage) */
  double *qpos, *qvel, *qacc_warmstart;
  const double* ctrl;
  double time;
  int flags;
  /* position-dependent */
  double *xpos, *xquat, *xmat, *xipos, *ximat, *xanchor, *xaxis, *geom_xpos, *geom_xmat, *site_xpos, *site_xmat;
  double *subtree_com, *cinert, *crb, *cdof, *cdof_dot, *cvel, *cacc, *cfrc, *M, *L;
  double *qfrc_bias, *qfrc_passive, *qfrc_actuator, *qfrc_smooth, *qacc_smooth, *qacc, *qfrc_constraint;
  double *act_length, *act_velocity, *act_force, *act_moment, *ten_length, *sensordata;
  /* contacts */
  int ncon;
  om_contact contact[OM_MAXCON];
  /* constraints */
  int nefc, ne, nf, nl, nc, solver_iter;
  int efc_type[OM_MAXEFC], efc_id[OM_MAXEFC], efc_state[OM_MAXEFC];
  double efc_pos[OM_MAXEFC], efc_margin[OM_MAXEFC], efc_floss[OM_MAXEFC], efc_diagApprox[OM_MAXEFC], efc_R[OM_MAXEFC],
      efc_D[OM_MAXEFC], efc_vel[OM_MAXEFC], efc_aref[OM_MAXEFC], efc_force[OM_MAXEFC], efc_solref[2 * OM_MAXEFC],
      efc_solimp[5 * OM_MAXEFC];
  double *efc_J, *H, *Hc;
  double sol_jar[OM_MAXEFC], sol_jv[OM_MAXEFC], sol_Ma[OM_MAXNV], sol_grad[OM_MAXNV], sol_search[OM_MAXNV],
      sol_mv[OM_MAXNV], sol_tmp[OM_MAXNV];
  double jacbuf[12 * OM_MAXNV];
} om_data;

typedef struct om_model om_model;
om_data* om_data_new(const om_model* m);
void om_data_free(om_data* d);
void om_kinematics(const om_model* m, om_data* d);
void om_forward(const om_model* m, om_data* d);
void om_step1(const om_model* m, om_data* d);
void om_collision(const om_model* m, om_data* d);
void om_parallel_for(const om_model* m, int nenv, int nthreads, void (*fn)(const om_model*, struct om_data*, int, void*),
                     void* ctx);
int om_max_threads(void);
void om_jac(const om_model* m, const om_data* d, int body, const double* point, double* jacp, double* jacr);
double om_ray(const om_model* m, const double* xpos, const double* xmat, const double* geom_xpos,
              const double* geom_xmat, const double* pnt, const double* vec, int groupmask, int bodyexclude,
              int* geomid);

/* ------------------------------------------------------------------------- small math */
static inline void v3zero(double* a) { a[0] = a[1] = a[2] = 0; }
static inline void v3copy(double* a, const double* b) { a[0] = b[0]; a[1] = b[1]; a[2] = b[2]; }
static inline void v3add(double* r, const double* a, const double* b) { r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2]; }
static inline void v3addto(double* r, const double* a) { r[0] += a[0]; r[1] += a[1]; r[2] += a[2]; }
static inline void v3sub(double* r, const double* a, const double* b) { r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2]; }
static inline void v3scl(double* r, const double* a, double s) { r[0] = a[0] * s; r[1] = a[1] * s; r[2] = a[2] * s; }
static inline void v3addscl(double* r, const double* a, const double* b, double s) { r[0] = a[0] + b[0] * s; r[1] = a[1] + b[1] * s; r[2] = a[2] + b[2] * s; }
static inline double v3dot(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline double v3norm(const double* a) { return sqrt(v3dot(a, a)); }
static inline void v3cross(double* r, const double* a, const double* b) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline double v3normalize(double* a) {
  double n = v3norm(a);
  if (n < OM_MINVAL) { a[0] = 1; a[1] = a[2] = 0; return 0; }
  a[0] /= n; a[1] /= n; a[2] /= n;
  return n;
}
static inline double dot6(const double* a, const double* b) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3] + a[4] * b[4] + a[5] * b[5];
}
/* r = M v (M row-major 3x3) */
static inline void mulmatvec3(double* r, const double* M, const double* v) {
  double x = M[0] * v[0] + M[1] * v[1] + M[2] * v[2], y = M[3] * v[0] + M[4] * v[1] + M[5] * v[2],
         z = M[6] * v[0] + M[7] * v[1] + M[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
/* r = M^T v */
static inline void multmatvec3(double* r, const double* M, const double* v) {
  double x = M[0] * v[0] + M[3] * v[1] + M[6] * v[2], y = M[1] * v[0] + M[4] * v[1] + M[7] * v[2],
         z = M[2] * v[0] + M[5] * v[1] + M[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline void quatmul(double* r, const double* a, const double* b) {
  double w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
         y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1], z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
static inline void quatnormalize(double* q) {
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < OM_MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
  q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}
static inline void quat2mat(double* R, const double* q) {
  double w = q[0], x = q[1], y = q[2], z = q[3];
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - w * z); R[2] = 2 * (x * z + w * y);
  R[3] = 2 * (x * y + w * z); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - w * x);
  R[6] = 2 * (x * z - w * y); R[7] = 2 * (y * z + w * x); R[8] = 1 - 2 * (x * x + y * y);
}
static inline void rotvecquat(double* r, const double* v, const double* q) {
  double R[9];
  quat2mat(R, q);
  mulmatvec3(r, R, v);
}
static inline void axisangle2quat(double* q, const double* axis, double ang) {
  double s = sin(0.5 * ang);
  q[0] = cos(0.5 * ang); q[1] = axis[0] * s; q[2] = axis[1] * s; q[3] = axis[2] * s;
}
/* quat <- normalize(quat * exp(h*w/2)), w in the body frame [upstream mju_quatIntegrate] */
static inline void quat_integrate(double* q, const double* w, double h) {
  double ax[3] = {w[0], w[1], w[2]};
  double n = v3norm(ax);
  if (n < OM_MINVAL) { quatnormalize(q); return; }
  ax[0] /= n; ax[1] /= n; ax[2] /= n;
  double dq[4], r[4];
  axisangle2quat(dq, ax, h * n);
  quatmul(r, q, dq);
  quatnormalize(r);
  memcpy(q, r, 32);
}
/* spatial inertia (10-vector about a reference point) times motion vector -> force vector */
static inline void mulinertvec(double* r, const double* i, const double* v) {
  r[0] = i[0] * v[0] + i[3] * v[1] + i[4] * v[2] - i[8] * v[4] + i[7] * v[5];
  r[1] = i[3] * v[0] + i[1] * v[1] + i[5] * v[2] + i[8] * v[3] - i[6] * v[5];
  r[2] = i[4] * v[0] + i[5] * v[1] + i[2] * v[2] - i[7] * v[3] + i[6] * v[4];
  r[3] = i[8] * v[1] - i[7] * v[2] + i[9] * v[3];
  r[4] = i[6] * v[2] - i[8] * v[0] + i[9] * v[4];
  r[5] = i[7] * v[0] - i[6] * v[1] + i[9] * v[5];
}
/* motion x motion */
static inline void crossmotion(double* r, const double* vel, const double* v) {
  double a[3], b[3];
  v3cross(r, vel, v);
  v3cross(a, vel, v + 3);
  v3cross(b, vel + 3, v);
  r[3] = a[0] + b[0]; r[4] = a[1] + b[1]; r[5] = a[2] + b[2];
}
/* motion x* force */
static inline void crossforce(double* r, const double* vel, const double* f) {
  double a[3], b[3];
  v3cross(a, vel, f);
  v3cross(b, vel + 3, f + 3);
  r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2];
  v3cross(r + 3, vel, f + 3);
}
/* y = A x, A is [r][c] row-major */
static inline void mulmat(double* y, const double* A, const double* x, int r, int c) {
  for (int i = 0; i < r; i++) {
    double s = 0;
    const double* a = A + (size_t)i * c;
    for (int k = 0; k < c; k++) s += a[k] * x[k];
    y[i] = s;
  }
}
/* in-place dense Cholesky (lower triangle holds L) */
static inline void chol_factor(double* A, int n) {
  for (int j = 0; j < n; j++) {
    double s = A[j * n + j];
    for (int k = 0; k < j; k++) s -= A[j * n + k] * A[j * n + k];
    if (s < OM_MINVAL) s = OM_MINVAL;
    double l = sqrt(s);
    A[j * n + j] = l;
    for (int i = j + 1; i < n; i++) {
      double t = A[i * n + j];
      for (int k = 0; k < j; k++) t -= A[i * n + k] * A[j * n + k];
      A[i * n + j] = t / l;
    }
  }
}
static inline void chol_solve(const double* L, double* x, int n) {
  for (int i = 0; i < n; i++) {
    double s = x[i];
    for (int k = 0; k < i; k++) s -= L[i * n + k] * x[k];
    x[i] = s / L[i * n + i];
  }
  for (int i = n - 1; i >= 0; i--) {
    double s = x[i];
    for (int k = i + 1; k < n; k++) s -= L[k * n + i] * x[k];
    x[i] = s / L[i * n + i];
  }
}
#endif
