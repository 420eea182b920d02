/* CPU oracle (fp64) -- TEST INFRASTRUCTURE ONLY. See ss_oracle.h for scope and provenance.
 *
 * Stage names follow SURVEY.md §8(a) (S1..S10), which in turn follow the stage order of the
 * un-vendored mujoco==3.2.6 `mj_step` the reference calls at
 * `stretch_mujoco/mujoco_server.py:378`.  [upstream] marks a behaviour restated from memory.
 */
#include "ss_oracle.h"
#include "../include/ss_blob.h"
#include "ss_oracle_internal.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

/* ------------------------------------------------------------------------- model */

static double* dupd(const ss_blob* b, const char* n, size_t* cnt) {
  const double* p = ss_blob_f64(b, n);
  size_t c = ss_blob_count(b, n);
  if (cnt) *cnt = c;
  double* r = (double*)malloc((c ? c : 1) * sizeof(double));
  if (p && c) memcpy(r, p, c * sizeof(double));
  return r;
}
static int* dupi(const ss_blob* b, const char* n, size_t* cnt) {
  const int32_t* p = ss_blob_i32(b, n);
  size_t c = ss_blob_count(b, n);
  if (cnt) *cnt = c;
  int* r = (int*)malloc((c ? c : 1) * sizeof(int));
  if (p && c) memcpy(r, p, c * sizeof(int));
  return r;
}

om_model* om_model_load(const void* buf, size_t nbytes) {
  ss_blob b;
  if (ss_blob_open(&b, buf, nbytes) != 0) return NULL;
  const int32_t* sz = ss_blob_i32(&b, "sizes");
  if (!sz) return NULL;
  om_model* m = (om_model*)calloc(1, sizeof(om_model));
  m->nq = sz[0]; m->nv = sz[1]; m->nu = sz[2]; m->nbody = sz[3]; m->njnt = sz[4]; m->ngeom = sz[5];
  m->nsite = sz[6]; m->ncam = sz[7]; m->ntendon = sz[8]; m->neq = sz[9]; m->nsensor = sz[10];
  m->nsensordata = sz[11]; m->nkey = sz[12]; m->nM = sz[13]; m->npair = sz[14]; m->nmesh = sz[15];
  if (m->nv > OM_MAXNV) { free(m); return NULL; }
#define D(field, name) m->field = dupd(&b, name, NULL)
#define I(field, name) m->field = dupi(&b, name, NULL)
  m->timestep = ss_blob_f64(&b, "opt_timestep")[0];
  memcpy(m->gravity, ss_blob_f64(&b, "opt_gravity"), 24);
  m->impratio = ss_blob_f64(&b, "opt_impratio")[0];
  m->tolerance = ss_blob_f64(&b, "opt_tolerance")[0];
  m->ls_tolerance = ss_blob_f64(&b, "opt_ls_tolerance")[0];
  m->iterations = ss_blob_i32(&b, "opt_iterations")[0];
  m->ls_iterations = ss_blob_i32(&b, "opt_ls_iterations")[0];
  m->cone = ss_blob_i32(&b, "opt_cone")[0];
  m->meaninertia = ss_blob_f64(&b, "stat_meaninertia")[0];
  m->extent = ss_blob_f64(&b, "stat_extent")[0];
  I(body_parentid, "body_parentid"); I(body_rootid, "body_rootid"); I(body_weldid, "body_weldid");
  I(body_jntnum, "body_jntnum"); I(body_jntadr, "body_jntadr"); I(body_dofnum, "body_dofnum");
  I(body_dofadr, "body_dofadr"); D(body_pos, "body_pos"); D(body_quat, "body_quat"); D(body_ipos, "body_ipos");
  D(body_iquat, "body_iquat"); D(body_mass, "body_mass"); D(body_inertia, "body_inertia");
  D(body_gravcomp, "body_gravcomp"); D(body_invweight0, "body_invweight0"); D(body_subtreemass, "body_subtreemass");
  I(jnt_type, "jnt_type"); I(jnt_bodyid, "jnt_bodyid"); I(jnt_qposadr, "jnt_qposadr"); I(jnt_dofadr, "jnt_dofadr");
  D(jnt_pos, "jnt_pos"); D(jnt_axis, "jnt_axis"); D(jnt_stiffness, "jnt_stiffness"); D(jnt_range, "jnt_range");
  I(jnt_limited, "jnt_limited"); D(jnt_margin, "jnt_margin"); D(jnt_solref, "jnt_solref"); D(jnt_solimp, "jnt_solimp");
  D(qpos0, "qpos0"); D(qpos_spring, "qpos_spring");
  I(dof_bodyid, "dof_bodyid"); I(dof_jntid, "dof_jntid"); I(dof_parentid, "dof_parentid");
  D(dof_armature, "dof_armature"); D(dof_damping, "dof_damping"); D(dof_frictionloss, "dof_frictionloss");
  D(dof_invweight0, "dof_invweight0"); D(dof_solref, "dof_solref"); D(dof_solimp, "dof_solimp");
  I(geom_type, "geom_type"); I(geom_bodyid, "geom_bodyid"); I(geom_dataid, "geom_dataid"); I(geom_group, "geom_group");
  I(geom_matid, "geom_matid");
  D(geom_size, "geom_size"); D(geom_rbound, "geom_rbound"); D(geom_pos, "geom_pos"); D(geom_quat, "geom_quat");
  D(geom_rgba, "geom_rgba"); D(geom_aabb, "geom_aabb");
  I(site_bodyid, "site_bodyid"); D(site_pos, "site_pos"); D(site_quat, "site_quat");
  I(cam_bodyid, "cam_bodyid"); D(cam_pos, "cam_pos"); D(cam_quat, "cam_quat"); D(cam_fovy, "cam_fovy");
  I(tendon_adr, "tendon_adr"); I(tendon_num, "tendon_num"); I(wrap_objid, "wrap_objid"); D(wrap_prm, "wrap_prm");
  I(eq_obj1id, "eq_obj1id"); I(eq_obj2id, "eq_obj2id"); D(eq_data, "eq_data"); D(eq_solref, "eq_solref");
  D(eq_solimp, "eq_solimp"); I(eq_active0, "eq_active0");
  I(actuator_trntype, "actuator_trntype"); I(actuator_trnid, "actuator_trnid"); D(actuator_gear, "actuator_gear");
  D(actuator_gainprm, "actuator_gainprm"); D(actuator_biasprm, "actuator_biasprm");
  D(actuator_ctrlrange, "actuator_ctrlrange"); I(actuator_ctrllimited, "actuator_ctrllimited");
  D(actuator_forcerange, "actuator_forcerange"); I(actuator_forcelimited, "actuator_forcelimited");
  I(sensor_type, "sensor_type"); I(sensor_objid, "sensor_objid"); I(sensor_adr, "sensor_adr");
  D(sensor_cutoff, "sensor_cutoff");
  I(pair_geom1, "pair_geom1"); I(pair_geom2, "pair_geom2"); I(pair_condim, "pair_condim");
  D(pair_friction, "pair_friction"); D(pair_solref, "pair_solref"); D(pair_solimp, "pair_solimp");
  D(pair_margin, "pair_margin"); D(pair_gap, "pair_gap");
  I(mesh_hulladr, "mesh_hulladr"); I(mesh_hullnum, "mesh_hullnum"); D(hull_vert, "hull_vert");
  D(key_qpos, "key_qpos"); D(key_ctrl, "key_ctrl");
  if (ss_blob_find(&b, "hull_edgeadr")) { I(hull_edgeadr, "hull_edgeadr"); I(hull_edge, "hull_edge"); }
  m->multiccd = ss_blob_find(&b, "opt_multiccd") ? ss_blob_i32(&b, "opt_multiccd")[0] : 0;
  /* ray geometry (optional) */
  if (ss_blob_find(&b, "rmesh_vertadr")) {
    m->has_ray = 1;
    I(rmesh_vertadr, "rmesh_vertadr"); I(rmesh_faceadr, "rmesh_faceadr"); I(rmesh_facenum, "rmesh_facenum");
    I(rmesh_bvhadr, "rmesh_bvhadr");
    size_t c;
    const float* vf = ss_blob_f32(&b, "rmesh_vert"); c = ss_blob_count(&b, "rmesh_vert");
    m->rmesh_vert = (float*)malloc((c ? c : 1) * 4); if (c) memcpy(m->rmesh_vert, vf, c * 4);
    I(rmesh_face, "rmesh_face");
    const float* bf = ss_blob_f32(&b, "bvh_aabb"); c = ss_blob_count(&b, "bvh_aabb");
    m->bvh_aabb = (float*)malloc((c ? c : 1) * 4); if (c) memcpy(m->bvh_aabb, bf, c * 4);
    I(bvh_child, "bvh_child");
    I(raygeom_id, "raygeom_id"); m->nraygeom = (int)ss_blob_count(&b, "raygeom_id");
    D(geom_shade, "geom_shade");
    D(vis_headlight, "vis_headlight"); D(vis_map, "vis_map"); D(skybox_rgb, "skybox_rgb");
    m->nsky = (int)ss_blob_count(&b, "skybox_rgb") / 3;
    I(light_bodyid, "light_bodyid"); D(light_pos, "light_pos"); D(light_dir, "light_dir");
    I(light_directional, "light_directional"); D(light_ambient, "light_ambient"); D(light_diffuse, "light_diffuse");
    D(light_specular, "light_specular"); m->nlight = (int)ss_blob_count(&b, "light_bodyid");
    m->headlight_active = ss_blob_i32(&b, "vis_headlight_active")[0];
    if (ss_blob_find(&b, "geom_tex")) {
      D(geom_tex, "geom_tex"); I(tex_adr, "tex_adr"); I(tex_w, "tex_w"); I(tex_h, "tex_h");
      m->ntex = (int)ss_blob_count(&b, "tex_adr");
      size_t c2 = ss_blob_count(&b, "tex_rgb");
      const unsigned char* tp = ss_blob_u8(&b, "tex_rgb");
      m->tex_rgb = (unsigned char*)malloc(c2 ? c2 : 1); if (c2 && tp) memcpy(m->tex_rgb, tp, c2);
      const float* uf = ss_blob_f32(&b, "rmesh_uv"); size_t c3 = ss_blob_count(&b, "rmesh_uv");
      if (uf && c3) { m->rmesh_uv = (float*)malloc(c3 * 4); memcpy(m->rmesh_uv, uf, c3 * 4); }
    }
  }
#undef D
#undef I
  /* derived: last dof on the chain of every body */
  m->body_lastdof = (int*)malloc(sizeof(int) * m->nbody);
  for (int i = 0; i < m->nbody; i++) {
    if (i == 0) { m->body_lastdof[i] = -1; continue; }
    if (m->body_dofnum[i] > 0) m->body_lastdof[i] = m->body_dofadr[i] + m->body_dofnum[i] - 1;
    else m->body_lastdof[i] = m->body_lastdof[m->body_parentid[i]];
  }
  m->enable_lidar = 1;
  return m;
}

void om_model_free(om_model* m) { free(m); /* arrays leak by design: process-lifetime test helper */ }

void om_model_sizes(const om_model* m, int* o) {
  int v[16] = {m->nq, m->nv, m->nu, m->nbody, m->njnt, m->ngeom, m->nsite, m->ncam, m->ntendon, m->neq,
               m->nsensor, m->nsensordata, m->nkey, m->nM, m->npair, m->nmesh};
  memcpy(o, v, sizeof(v));
}

void om_set_options(om_model* m, int max_iter, double tol, int enable_lidar) {
  if (max_iter > 0) m->iterations = max_iter;
  if (tol > 0) m->tolerance = tol;
  m->enable_lidar = enable_lidar;
}

int om_max_threads(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}

/* minimal pthread parallel-for over envs (dynamic scheduling through an atomic counter) */
typedef struct {
  const om_model* m; int nenv; int next; pthread_mutex_t mu;
  void (*fn)(const om_model*, om_data*, int, void*); void* ctx;
} om_pool;
static void* pool_worker(void* arg) {
  om_pool* p = (om_pool*)arg;
  om_data* d = om_data_new(p->m);
  for (;;) {
    pthread_mutex_lock(&p->mu);
    int e = p->next++;
    pthread_mutex_unlock(&p->mu);
    if (e >= p->nenv) break;
    p->fn(p->m, d, e, p->ctx);
  }
  om_data_free(d);
  return NULL;
}
void om_parallel_for(const om_model* m, int nenv, int nthreads, void (*fn)(const om_model*, om_data*, int, void*),
                     void* ctx) {
  if (nthreads <= 0) nthreads = om_max_threads();
  if (nthreads > nenv) nthreads = nenv;
  if (nthreads < 1) nthreads = 1;
  om_pool p = {m, nenv, 0, PTHREAD_MUTEX_INITIALIZER, fn, ctx};
  if (nthreads == 1) { pool_worker(&p); return; }
  pthread_t th[256];
  if (nthreads > 256) nthreads = 256;
  for (int i = 0; i < nthreads; i++) pthread_create(&th[i], NULL, pool_worker, &p);
  for (int i = 0; i < nthreads; i++) pthread_join(th[i], NULL);
}

/* ------------------------------------------------------------------------- data */

om_data* om_data_new(const om_model* m) {
  om_data* d = (om_data*)calloc(1, sizeof(om_data));
  int nb = m->nbody, nv = m->nv;
#define AL(f, n) d->f = (double*)calloc((size_t)(n) > 0 ? (size_t)(n) : 1, sizeof(double))
  AL(xpos, nb * 3); AL(xquat, nb * 4); AL(xmat, nb * 9); AL(xipos, nb * 3); AL(ximat, nb * 9);
  AL(xanchor, m->njnt * 3); AL(xaxis, m->njnt * 3); AL(geom_xpos, m->ngeom * 3); AL(geom_xmat, m->ngeom * 9);
  AL(site_xpos, m->nsite * 3); AL(site_xmat, m->nsite * 9); AL(subtree_com, nb * 3); AL(cinert, nb * 10);
  AL(crb, nb * 10); AL(cdof, nv * 6); AL(cdof_dot, nv * 6); AL(cvel, nb * 6); AL(cacc, nb * 6); AL(cfrc, nb * 6);
  AL(M, nv * nv); AL(L, nv * nv); AL(qfrc_bias, nv); AL(qfrc_passive, nv); AL(qfrc_actuator, nv);
  AL(qfrc_smooth, nv); AL(qacc_smooth, nv); AL(qacc, nv); AL(qfrc_constraint, nv);
  AL(act_length, m->nu); AL(act_velocity, m->nu); AL(act_force, m->nu); AL(act_moment, m->nu * nv);
  AL(ten_length, m->ntendon); AL(sensordata, m->nsensordata);
  AL(efc_J, OM_MAXEFC * nv); AL(H, nv * nv); AL(Hc, nv * nv);
#undef AL
  return d;
}

void om_data_free(om_data* d) {
  double** p[] = {&d->xpos, &d->xquat, &d->xmat, &d->xipos, &d->ximat, &d->xanchor, &d->xaxis, &d->geom_xpos,
                  &d->geom_xmat, &d->site_xpos, &d->site_xmat, &d->subtree_com, &d->cinert, &d->crb, &d->cdof,
                  &d->cdof_dot, &d->cvel, &d->cacc, &d->cfrc, &d->M, &d->L, &d->qfrc_bias, &d->qfrc_passive,
                  &d->qfrc_actuator, &d->qfrc_smooth, &d->qacc_smooth, &d->qacc, &d->qfrc_constraint,
                  &d->act_length, &d->act_velocity, &d->act_force, &d->act_moment, &d->ten_length, &d->sensordata,
                  &d->efc_J, &d->H, &d->Hc};
  for (size_t i = 0; i < sizeof(p) / sizeof(p[0]); i++) free(*p[i]);
  free(d);
}

/* ------------------------------------------------------------------------- S1: position stage */

/* forward kinematics: body, inertial, geom and site frames [upstream mj_kinematics] */
void om_kinematics(const om_model* m, om_data* d) {
  v3zero(d->xpos); d->xquat[0] = 1; d->xquat[1] = d->xquat[2] = d->xquat[3] = 0;
  quat2mat(d->xmat, d->xquat);
  v3zero(d->xipos); quat2mat(d->ximat, d->xquat);
  for (int b = 1; b < m->nbody; b++) {
    int p = m->body_parentid[b], jn = m->body_jntnum[b], ja = m->body_jntadr[b];
    double pos[3], quat[4];
    if (jn == 1 && m->jnt_type[ja] == JNT_FREE) {
      const double* q = d->qpos + m->jnt_qposadr[ja];
      v3copy(pos, q); memcpy(quat, q + 3, 32); quatnormalize(quat);
      v3copy(d->xanchor + 3 * ja, pos);
      d->xaxis[3 * ja] = 0; d->xaxis[3 * ja + 1] = 0; d->xaxis[3 * ja + 2] = 1;
    } else {
      double t[3];
      mulmatvec3(t, d->xmat + 9 * p, m->body_pos + 3 * b);
      v3add(pos, d->xpos + 3 * p, t);
      quatmul(quat, d->xquat + 4 * p, m->body_quat + 4 * b);
      for (int j = ja; j < ja + jn; j++) {
        double anchor[3], axis[3], r[3];
        rotvecquat(r, m->jnt_pos + 3 * j, quat); v3add(anchor, pos, r);
        rotvecquat(axis, m->jnt_axis + 3 * j, quat);
        double dq = d->qpos[m->jnt_qposadr[j]] - m->qpos0[m->jnt_qposadr[j]];
        if (m->jnt_type[j] == JNT_SLIDE) {
          v3addscl(pos, pos, axis, dq);
        } else if (m->jnt_type[j] == JNT_HINGE) {
          double qr[4], qn[4];
          axisangle2quat(qr, m->jnt_axis + 3 * j, dq);
          quatmul(qn, quat, qr); memcpy(quat, qn, 32);
          rotvecquat(r, m->jnt_pos + 3 * j, quat); v3sub(pos, anchor, r);
        }
        v3copy(d->xanchor + 3 * j, anchor); v3copy(d->xaxis + 3 * j, axis);
      }
      quatnormalize(quat);
    }
    v3copy(d->xpos + 3 * b, pos); memcpy(d->xquat + 4 * b, quat, 32);
    quat2mat(d->xmat + 9 * b, quat);
    double t[3], qi[4];
    mulmatvec3(t, d->xmat + 9 * b, m->body_ipos + 3 * b); v3add(d->xipos + 3 * b, pos, t);
    quatmul(qi, quat, m->body_iquat + 4 * b); quat2mat(d->ximat + 9 * b, qi);
  }
  for (int g = 0; g < m->ngeom; g++) {
    int b = m->geom_bodyid[g];
    double t[3], q[4];
    mulmatvec3(t, d->xmat + 9 * b, m->geom_pos + 3 * g); v3add(d->geom_xpos + 3 * g, d->xpos + 3 * b, t);
    quatmul(q, d->xquat + 4 * b, m->geom_quat + 4 * g); quatnormalize(q); quat2mat(d->geom_xmat + 9 * g, q);
  }
  for (int s = 0; s < m->nsite; s++) {
    int b = m->site_bodyid[s];
    double t[3], q[4];
    mulmatvec3(t, d->xmat + 9 * b, m->site_pos + 3 * s); v3add(d->site_xpos + 3 * s, d->xpos + 3 * b, t);
    quatmul(q, d->xquat + 4 * b, m->site_quat + 4 * s); quatnormalize(q); quat2mat(d->site_xmat + 9 * s, q);
  }
}

/* subtree COM, body inertias and dof axes about the tree's COM [upstream mj_comPos] */
static void om_compos(const om_model* m, om_data* d) {
  int nb = m->nbody;
  for (int b = 0; b < nb; b++) v3scl(d->subtree_com + 3 * b, d->xipos + 3 * b, m->body_mass[b]);
  for (int b = nb - 1; b > 0; b--) v3addto(d->subtree_com + 3 * m->body_parentid[b], d->subtree_com + 3 * b);
  for (int b = 0; b < nb; b++) {
    if (m->body_subtreemass[b] < OM_MINVAL) v3copy(d->subtree_com + 3 * b, d->xipos + 3 * b);
    else v3scl(d->subtree_com + 3 * b, d->subtree_com + 3 * b, 1.0 / m->body_subtreemass[b]);
  }
  memset(d->cinert, 0, sizeof(double) * 10);
  for (int b = 1; b < nb; b++) {
    double off[3], *ci = d->cinert + 10 * b, mass = m->body_mass[b];
    const double *R = d->ximat + 9 * b, *I = m->body_inertia + 3 * b;
    v3sub(off, d->xipos + 3 * b, d->subtree_com + 3 * m->body_rootid[b]);
    /* R diag(I) R^T */
    double Iw[9];
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++)
        Iw[3 * r + c] = R[3 * r] * I[0] * R[3 * c] + R[3 * r + 1] * I[1] * R[3 * c + 1] + R[3 * r + 2] * I[2] * R[3 * c + 2];
    double o2 = v3dot(off, off);
    ci[0] = Iw[0] + mass * (o2 - off[0] * off[0]);
    ci[1] = Iw[4] + mass * (o2 - off[1] * off[1]);
    ci[2] = Iw[8] + mass * (o2 - off[2] * off[2]);
    ci[3] = Iw[1] - mass * off[0] * off[1];
    ci[4] = Iw[2] - mass * off[0] * off[2];
    ci[5] = Iw[5] - mass * off[1] * off[2];
    ci[6] = mass * off[0]; ci[7] = mass * off[1]; ci[8] = mass * off[2]; ci[9] = mass;
  }
  for (int j = 0; j < m->njnt; j++) {
    int b = m->jnt_bodyid[j], da = m->jnt_dofadr[j];
    double off[3];
    v3sub(off, d->subtree_com + 3 * m->body_rootid[b], d->xanchor + 3 * j);
    if (m->jnt_type[j] == JNT_FREE) {
      for (int k = 0; k < 3; k++) {
        double* c = d->cdof + 6 * (da + k);
        memset(c, 0, 48); c[3 + k] = 1;
        double ax[3] = {d->xmat[9 * b + k], d->xmat[9 * b + 3 + k], d->xmat[9 * b + 6 + k]};
        double* cr = d->cdof + 6 * (da + 3 + k);
        v3copy(cr, ax); v3cross(cr + 3, ax, off);
      }
    } else if (m->jnt_type[j] == JNT_SLIDE) {
      double* c = d->cdof + 6 * da;
      v3zero(c); v3copy(c + 3, d->xaxis + 3 * j);
    } else {
      double* c = d->cdof + 6 * da;
      v3copy(c, d->xaxis + 3 * j); v3cross(c + 3, d->xaxis + 3 * j, off);
    }
  }
}

/* fixed tendons and actuator transmission [upstream mj_tendon, mj_transmission] */
static void om_transmission(const om_model* m, om_data* d) {
  for (int t = 0; t < m->ntendon; t++) {
    double L = 0;
    for (int w = m->tendon_adr[t]; w < m->tendon_adr[t] + m->tendon_num[t]; w++)
      L += m->wrap_prm[w] * d->qpos[m->jnt_qposadr[m->wrap_objid[w]]];
    d->ten_length[t] = L;
  }
  memset(d->act_moment, 0, sizeof(double) * m->nu * m->nv);
  for (int a = 0; a < m->nu; a++) {
    double gear = m->actuator_gear[a];
    if (m->actuator_trntype[a] == 0) {
      int j = m->actuator_trnid[a];
      d->act_length[a] = gear * d->qpos[m->jnt_qposadr[j]];
      d->act_moment[a * m->nv + m->jnt_dofadr[j]] = gear;
    } else {
      int t = m->actuator_trnid[a];
      d->act_length[a] = gear * d->ten_length[t];
      for (int w = m->tendon_adr[t]; w < m->tendon_adr[t] + m->tendon_num[t]; w++)
        d->act_moment[a * m->nv + m->jnt_dofadr[m->wrap_objid[w]]] = gear * m->wrap_prm[w];
    }
  }
}

/* composite rigid body inertia -> dense joint-space M; Cholesky factor in L [upstream mj_crb, mj_factorM] */
static void om_crb(const om_model* m, om_data* d) {
  int nb = m->nbody, nv = m->nv;
  memcpy(d->crb, d->cinert, sizeof(double) * 10 * nb);
  for (int b = nb - 1; b > 0; b--) {
    int p = m->body_parentid[b];
    if (p > 0) for (int k = 0; k < 10; k++) d->crb[10 * p + k] += d->crb[10 * b + k];
  }
  memset(d->M, 0, sizeof(double) * nv * nv);
  for (int i = 0; i < nv; i++) {
    double buf[6];
    mulinertvec(buf, d->crb + 10 * m->dof_bodyid[i], d->cdof + 6 * i);
    for (int j = i; j >= 0; j = m->dof_parentid[j]) {
      double v = dot6(d->cdof + 6 * j, buf);
      d->M[i * nv + j] = v; d->M[j * nv + i] = v;
    }
    d->M[i * nv + i] += m->dof_armature[i];
  }
  memcpy(d->L, d->M, sizeof(double) * nv * nv);
  chol_factor(d->L, nv);
}

/* Jacobian of a world point attached to a body: jacp/jacr are [3][nv] (either may be NULL) */
void om_jac(const om_model* m, const om_data* d, int body, const double* point, double* jacp, double* jacr) {
  int nv = m->nv;
  if (jacp) memset(jacp, 0, sizeof(double) * 3 * nv);
  if (jacr) memset(jacr, 0, sizeof(double) * 3 * nv);
  if (body <= 0) return;
  double off[3];
  v3sub(off, point, d->subtree_com + 3 * m->body_rootid[body]);
  for (int i = m->body_lastdof[body]; i >= 0; i = m->dof_parentid[i]) {
    const double* c = d->cdof + 6 * i;
    if (jacr) { jacr[i] = c[0]; jacr[nv + i] = c[1]; jacr[2 * nv + i] = c[2]; }
    if (jacp) {
      double t[3];
      v3cross(t, c, off);
      jacp[i] = c[3] + t[0]; jacp[nv + i] = c[4] + t[1]; jacp[2 * nv + i] = c[5] + t[2];
    }
  }
}

/* ------------------------------------------------------------------------- S3: velocity stage */

static void om_comvel(const om_model* m, om_data* d) {
  memset(d->cvel, 0, 48);
  for (int b = 1; b < m->nbody; b++) {
    double cv[6];
    memcpy(cv, d->cvel + 6 * m->body_parentid[b], 48);
    int ja = m->body_jntadr[b];
    for (int j = ja; j < ja + m->body_jntnum[b]; j++) {
      int da = m->jnt_dofadr[j];
      if (m->jnt_type[j] == JNT_FREE) {
        memset(d->cdof_dot + 6 * da, 0, sizeof(double) * 18);
        for (int k = 0; k < 3; k++) for (int c = 0; c < 6; c++) cv[c] += d->cdof[6 * (da + k) + c] * d->qvel[da + k];
        for (int k = 3; k < 6; k++) crossmotion(d->cdof_dot + 6 * (da + k), cv, d->cdof + 6 * (da + k));
        for (int k = 3; k < 6; k++) for (int c = 0; c < 6; c++) cv[c] += d->cdof[6 * (da + k) + c] * d->qvel[da + k];
      } else {
        crossmotion(d->cdof_dot + 6 * da, cv, d->cdof + 6 * da);
        for (int c = 0; c < 6; c++) cv[c] += d->cdof[6 * da + c] * d->qvel[da];
      }
    }
    memcpy(d->cvel + 6 * b, cv, 48);
  }
}

/* springs, dampers, gravity compensation [upstream mj_passive] */
static void om_passive(const om_model* m, om_data* d) {
  int nv = m->nv;
  for (int i = 0; i < nv; i++) d->qfrc_passive[i] = -m->dof_damping[i] * d->qvel[i];
  for (int j = 0; j < m->njnt; j++) {
    if (m->jnt_stiffness[j] == 0) continue;
    if (m->jnt_type[j] == JNT_SLIDE || m->jnt_type[j] == JNT_HINGE) {
      int qa = m->jnt_qposadr[j];
      d->qfrc_passive[m->jnt_dofadr[j]] -= m->jnt_stiffness[j] * (d->qpos[qa] - m->qpos_spring[qa]);
    }
  }
  for (int b = 1; b < m->nbody; b++) {
    if (m->body_gravcomp[b] == 0) continue;
    double f[3], off[3];
    v3scl(f, m->gravity, -m->body_mass[b] * m->body_gravcomp[b]);
    v3sub(off, d->xipos + 3 * b, d->subtree_com + 3 * m->body_rootid[b]);
    for (int i = m->body_lastdof[b]; i >= 0; i = m->dof_parentid[i]) {
      const double* c = d->cdof + 6 * i;
      double t[3];
      v3cross(t, c, off);
      d->qfrc_passive[i] += (c[3] + t[0]) * f[0] + (c[4] + t[1]) * f[1] + (c[5] + t[2]) * f[2];
    }
  }
}

/* recursive Newton-Euler. flg_acc: include cdof*qacc [upstream mj_rne / mj_rnePostConstraint] */
static void om_rne(const om_model* m, om_data* d, int flg_acc, const double* qacc, double* result) {
  int nb = m->nbody;
  memset(d->cacc, 0, 48);
  for (int k = 0; k < 3; k++) d->cacc[3 + k] = -m->gravity[k];
  for (int b = 1; b < nb; b++) {
    double* ca = d->cacc + 6 * b;
    memcpy(ca, d->cacc + 6 * m->body_parentid[b], 48);
    int da = m->body_dofadr[b];
    for (int k = 0; k < m->body_dofnum[b]; k++) {
      for (int c = 0; c < 6; c++) ca[c] += d->cdof_dot[6 * (da + k) + c] * d->qvel[da + k];
      if (flg_acc) for (int c = 0; c < 6; c++) ca[c] += d->cdof[6 * (da + k) + c] * qacc[da + k];
    }
    double t1[6], t2[6], *cf = d->cfrc + 6 * b;
    mulinertvec(cf, d->cinert + 10 * b, ca);
    mulinertvec(t1, d->cinert + 10 * b, d->cvel + 6 * b);
    crossforce(t2, d->cvel + 6 * b, t1);
    for (int c = 0; c < 6; c++) cf[c] += t2[c];
  }
  memset(d->cfrc, 0, 48);
  if (!result) return;
  for (int b = nb - 1; b > 0; b--) {
    int p = m->body_parentid[b];
    if (p > 0) for (int c = 0; c < 6; c++) d->cfrc[6 * p + c] += d->cfrc[6 * b + c];
  }
  for (int i = 0; i < m->nv; i++) result[i] = dot6(d->cdof + 6 * i, d->cfrc + 6 * m->dof_bodyid[i]);
}

/* ------------------------------------------------------------------------- S5: actuation */

static void om_actuation(const om_model* m, om_data* d) {
  int nv = m->nv;
  memset(d->qfrc_actuator, 0, sizeof(double) * nv);
  for (int a = 0; a < m->nu; a++) {
    double vel = 0;
    for (int i = 0; i < nv; i++) vel += d->act_moment[a * nv + i] * d->qvel[i];
    d->act_velocity[a] = vel;
    double c = d->ctrl[a];
    if (m->actuator_ctrllimited[a]) c = fmin(fmax(c, m->actuator_ctrlrange[2 * a]), m->actuator_ctrlrange[2 * a + 1]);
    const double *g = m->actuator_gainprm + 3 * a, *bp = m->actuator_biasprm + 3 * a;
    double f = g[0] * c + bp[0] + bp[1] * d->act_length[a] + bp[2] * vel;
    if (m->actuator_forcelimited[a]) f = fmin(fmax(f, m->actuator_forcerange[2 * a]), m->actuator_forcerange[2 * a + 1]);
    d->act_force[a] = f;
    for (int i = 0; i < nv; i++) d->qfrc_actuator[i] += d->act_moment[a * nv + i] * f;
  }
}

/* ------------------------------------------------------------------------- S1m: constraints */

/* impedance d(r) and its derivative [upstream getimpedance] */
static double impedance(const double* solimp, double pos, double margin) {
  double dmin = fmin(fmax(solimp[0], OM_MINIMP), OM_MAXIMP), dmax = fmin(fmax(solimp[1], OM_MINIMP), OM_MAXIMP);
  double width = fmax(solimp[2], 0), mid = fmin(fmax(solimp[3], OM_MINIMP), OM_MAXIMP), power = fmax(solimp[4], 1);
  if (dmin == dmax || width <= OM_MINVAL) return 0.5 * (dmin + dmax);
  double x = fabs((pos - margin) / width);
  if (x >= 1) return dmax;
  if (x <= 0) return dmin;
  double y;
  if (power == 1) y = x;
  else if (x <= mid) y = pow(x, power) / pow(mid, power - 1);
  else y = 1 - pow(1 - x, power) / pow(1 - mid, power - 1);
  return dmin + y * (dmax - dmin);
}

static int add_row(const om_model* m, om_data* d, int type, int id, double pos, double margin, double floss,
                   double diag, const double* solref, const double* solimp) {
  if (d->nefc >= OM_MAXEFC) { d->flags |= 2; return -1; }
  int i = d->nefc++;
  d->efc_type[i] = type; d->efc_id[i] = id; d->efc_pos[i] = pos; d->efc_margin[i] = margin;
  d->efc_floss[i] = floss; d->efc_diagApprox[i] = diag;
  memcpy(d->efc_solref + 2 * i, solref, 16); memcpy(d->efc_solimp + 5 * i, solimp, 40);
  memset(d->efc_J + (size_t)i * m->nv, 0, sizeof(double) * m->nv);
  return i;
}

/* rows in the reference order: equality, friction loss, limits, contacts [upstream mj_makeConstraint] */
static void om_make_constraint(const om_model* m, om_data* d) {
  int nv = m->nv;
  d->nefc = 0;
  for (int e = 0; e < m->neq; e++) {
    if (!m->eq_active0[e]) continue;
    int j1 = m->eq_obj1id[e], j2 = m->eq_obj2id[e];
    const double* c = m->eq_data + 5 * e;
    int q1 = m->jnt_qposadr[j1], d1 = m->jnt_dofadr[j1];
    double pos, diag = m->dof_invweight0[d1];
    int row;
    if (j2 >= 0) {
      int q2 = m->jnt_qposadr[j2], d2 = m->jnt_dofadr[j2];
      double dif = d->qpos[q2] - m->qpos0[q2];
      double poly = c[0] + dif * (c[1] + dif * (c[2] + dif * (c[3] + dif * c[4])));
      double deriv = c[1] + dif * (2 * c[2] + dif * (3 * c[3] + dif * 4 * c[4]));
      pos = d->qpos[q1] - m->qpos0[q1] - poly;
      diag += m->dof_invweight0[d2];
      row = add_row(m, d, CNSTR_EQUALITY, e, pos, 0, 0, diag, m->eq_solref + 2 * e, m->eq_solimp + 5 * e);
      if (row < 0) return;
      d->efc_J[(size_t)row * nv + d1] = 1; d->efc_J[(size_t)row * nv + d2] = -deriv;
    } else {
      pos = d->qpos[q1] - m->qpos0[q1] - c[0];
      row = add_row(m, d, CNSTR_EQUALITY, e, pos, 0, 0, diag, m->eq_solref + 2 * e, m->eq_solimp + 5 * e);
      if (row < 0) return;
      d->efc_J[(size_t)row * nv + d1] = 1;
    }
  }
  d->ne = d->nefc;
  for (int i = 0; i < nv; i++) {
    if (m->dof_frictionloss[i] <= 0) continue;
    int row = add_row(m, d, CNSTR_FRICTION, i, 0, 0, m->dof_frictionloss[i], m->dof_invweight0[i],
                      m->dof_solref + 2 * i, m->dof_solimp + 5 * i);
    if (row < 0) return;
    d->efc_J[(size_t)row * nv + i] = 1;
  }
  d->nf = d->nefc - d->ne;
  for (int j = 0; j < m->njnt; j++) {
    if (!m->jnt_limited[j]) continue;
    int t = m->jnt_type[j];
    if (t != JNT_SLIDE && t != JNT_HINGE) continue;
    double q = d->qpos[m->jnt_qposadr[j]], margin = m->jnt_margin[j];
    int da = m->jnt_dofadr[j];
    for (int side = -1; side <= 1; side += 2) {
      double dist = side * (m->jnt_range[2 * j + (side + 1) / 2] - q);
      if (dist < margin) {
        int row = add_row(m, d, CNSTR_LIMIT, j, dist, margin, 0, m->dof_invweight0[da], m->jnt_solref + 2 * j,
                          m->jnt_solimp + 5 * j);
        if (row < 0) return;
        d->efc_J[(size_t)row * nv + da] = -side;
      }
    }
  }
  d->nl = d->nefc - d->ne - d->nf;
  double* jp1 = d->jacbuf; double* jp2 = jp1 + 3 * nv; double* jr1 = jp2 + 3 * nv; double* jr2 = jr1 + 3 * nv;
  for (int c = 0; c < d->ncon; c++) {
    om_contact* con = &d->contact[c];
    int b1 = m->geom_bodyid[con->geom1], b2 = m->geom_bodyid[con->geom2];
    int dim = con->dim;
    if (d->nefc + dim > OM_MAXEFC) { d->flags |= 2; d->ncon = c; break; }
    om_jac(m, d, b1, con->pos, jp1, jr1);
    om_jac(m, d, b2, con->pos, jp2, jr2);
    double tran = m->body_invweight0[2 * b1] + m->body_invweight0[2 * b2];
    double rot = m->body_invweight0[2 * b1 + 1] + m->body_invweight0[2 * b2 + 1];
    con->efc_address = d->nefc;
    for (int r = 0; r < dim; r++) {
      int type = dim == 1 ? CNSTR_CONTACT_FRICTIONLESS : CNSTR_CONTACT_ELLIPTIC;
      int row = add_row(m, d, type, c, r == 0 ? con->dist : 0, r == 0 ? con->includemargin : 0, 0,
                        r < 3 ? tran : rot, con->solref, con->solimp);
      double* J = d->efc_J + (size_t)row * nv;
      const double* ax = con->frame + 3 * (r < 3 ? r : r - 3);
      const double *ja = r < 3 ? jp1 : jr1, *jb = r < 3 ? jp2 : jr2;
      for (int i = 0; i < nv; i++)
        J[i] = ax[0] * (jb[i] - ja[i]) + ax[1] * (jb[nv + i] - ja[nv + i]) + ax[2] * (jb[2 * nv + i] - ja[2 * nv + i]);
    }
  }
  d->nc = d->nefc - d->ne - d->nf - d->nl;
}

/* R, D, reference acceleration [upstream mj_makeImpedance, mj_referenceConstraint] */
static void om_impedance_and_reference(const om_model* m, om_data* d) {
  int nv = m->nv;
  for (int i = 0; i < d->nefc; i++) {
    const double *solref = d->efc_solref + 2 * i, *solimp = d->efc_solimp + 5 * i;
    double imp = impedance(solimp, d->efc_pos[i], d->efc_margin[i]);
    d->efc_R[i] = fmax(OM_MINVAL, (1 - imp) / imp * d->efc_diagApprox[i]);
    double K, B, dmax = fmin(fmax(solimp[1], OM_MINIMP), OM_MAXIMP);
    if (solref[0] > 0) {
      double tc = fmax(solref[0], 2 * m->timestep), dr = solref[1];
      K = 1 / fmax(OM_MINVAL, dmax * dmax * tc * tc * dr * dr);
      B = 2 / fmax(OM_MINVAL, dmax * tc);
    } else {
      K = -solref[0] / fmax(OM_MINVAL, dmax * dmax);
      B = -solref[1] / fmax(OM_MINVAL, dmax);
    }
    double vel = 0;
    const double* J = d->efc_J + (size_t)i * nv;
    for (int k = 0; k < nv; k++) vel += J[k] * d->qvel[k];
    d->efc_vel[i] = vel;
    d->efc_aref[i] = -B * vel - K * imp * (d->efc_pos[i] - d->efc_margin[i]);
  }
  /* elliptic cones: friction regularisation from the normal row [upstream mj_makeImpedance] */
  for (int c = 0; c < d->ncon; c++) {
    om_contact* con = &d->contact[c];
    int i = con->efc_address, dim = con->dim;
    if (dim == 1) { con->mu = 0; continue; }
    const double* fr = con->friction;
    d->efc_R[i + 1] = d->efc_R[i] / fmax(OM_MINVAL, m->impratio);
    con->mu = fr[0] * sqrt(d->efc_R[i + 1] / d->efc_R[i]);
    for (int j = 2; j < dim; j++) d->efc_R[i + j] = d->efc_R[i + 1] * fr[0] * fr[0] / (fr[j - 1] * fr[j - 1]);
  }
  for (int i = 0; i < d->nefc; i++) d->efc_D[i] = 1 / d->efc_R[i];
}

/* ------------------------------------------------------------------------- S7: Newton solver */

typedef struct {
  double cost;
} sol_ctx;

/* forces, states and cost at jar [upstream PrimalUpdateConstraint]; optionally cone Hessians */
static double constraint_update(const om_model* m, om_data* d, const double* jar, double* force, int* state) {
  double cost = 0;
  int ne = d->ne, nf = d->nf, nefc = d->nefc;
  for (int i = 0; i < ne; i++) {
    force[i] = -d->efc_D[i] * jar[i]; state[i] = ST_QUADRATIC; cost += 0.5 * d->efc_D[i] * jar[i] * jar[i];
  }
  for (int i = ne; i < ne + nf; i++) {
    double f = d->efc_floss[i], R = d->efc_R[i], D = d->efc_D[i], x = jar[i];
    if (x <= -R * f) { force[i] = f; state[i] = ST_LINEARNEG; cost += -0.5 * R * f * f - f * x; }
    else if (x >= R * f) { force[i] = -f; state[i] = ST_LINEARPOS; cost += -0.5 * R * f * f + f * x; }
    else { force[i] = -D * x; state[i] = ST_QUADRATIC; cost += 0.5 * D * x * x; }
  }
  for (int i = ne + nf; i < nefc; i++) {
    if (d->efc_type[i] != CNSTR_CONTACT_ELLIPTIC) {
      if (jar[i] < 0) { force[i] = -d->efc_D[i] * jar[i]; state[i] = ST_QUADRATIC; cost += 0.5 * d->efc_D[i] * jar[i] * jar[i]; }
      else { force[i] = 0; state[i] = ST_SATISFIED; }
      continue;
    }
    om_contact* con = &d->contact[d->efc_id[i]];
    int dim = con->dim;
    double mu = con->mu, U[6], T2 = 0;
    U[0] = jar[i] * mu;
    for (int j = 1; j < dim; j++) { U[j] = jar[i + j] * con->friction[j - 1]; T2 += U[j] * U[j]; }
    double N = U[0], T = sqrt(T2);
    if (N >= mu * T || (T <= 0 && N >= 0)) {
      for (int j = 0; j < dim; j++) { force[i + j] = 0; state[i + j] = ST_SATISFIED; }
    } else if (mu * N + T <= 0 || (T <= 0 && N < 0)) {
      for (int j = 0; j < dim; j++) {
        force[i + j] = -d->efc_D[i + j] * jar[i + j]; state[i + j] = ST_QUADRATIC;
        cost += 0.5 * d->efc_D[i + j] * jar[i + j] * jar[i + j];
      }
    } else {
      double Dm = d->efc_D[i] / fmax(mu * mu * (1 + mu * mu), OM_MINVAL);
      double NT = N - mu * T;
      cost += 0.5 * Dm * NT * NT;
      force[i] = -Dm * NT * mu; state[i] = ST_CONE;
      for (int j = 1; j < dim; j++) { force[i + j] = -force[i] / T * U[j] * con->friction[j - 1]; state[i + j] = ST_CONE; }
    }
    i += dim - 1;
  }
  return cost;
}

/* one-dimensional cost along the search direction: value and first two derivatives */
static void ls_eval(const om_model* m, const om_data* d, const double* jar, const double* jv, double quad_g1,
                    double quad_g2, double a, double* p0, double* p1, double* p2) {
  (void)m;
  double c = a * quad_g1 + 0.5 * a * a * quad_g2, g = quad_g1 + a * quad_g2, h = quad_g2;
  int ne = d->ne, nf = d->nf, nefc = d->nefc;
  for (int i = 0; i < ne; i++) {
    double x = jar[i] + a * jv[i], D = d->efc_D[i];
    c += 0.5 * D * x * x; g += D * x * jv[i]; h += D * jv[i] * jv[i];
  }
  for (int i = ne; i < ne + nf; i++) {
    double f = d->efc_floss[i], R = d->efc_R[i], D = d->efc_D[i], x = jar[i] + a * jv[i];
    if (x <= -R * f) { c += -0.5 * R * f * f - f * x; g += -f * jv[i]; }
    else if (x >= R * f) { c += -0.5 * R * f * f + f * x; g += f * jv[i]; }
    else { c += 0.5 * D * x * x; g += D * x * jv[i]; h += D * jv[i] * jv[i]; }
  }
  for (int i = ne + nf; i < nefc; i++) {
    if (d->efc_type[i] != CNSTR_CONTACT_ELLIPTIC) {
      double x = jar[i] + a * jv[i], D = d->efc_D[i];
      if (x < 0) { c += 0.5 * D * x * x; g += D * x * jv[i]; h += D * jv[i] * jv[i]; }
      continue;
    }
    const om_contact* con = &d->contact[d->efc_id[i]];
    int dim = con->dim;
    double mu = con->mu;
    double N = (jar[i] + a * jv[i]) * mu, N1 = jv[i] * mu, TT = 0, UV = 0, VV = 0;
    for (int j = 1; j < dim; j++) {
      double fj = con->friction[j - 1], u = (jar[i + j] + a * jv[i + j]) * fj, v = jv[i + j] * fj;
      TT += u * u; UV += u * v; VV += v * v;
    }
    double T = sqrt(TT);
    if (N >= mu * T || (T <= 0 && N >= 0)) {
      /* satisfied */
    } else if (mu * N + T <= 0 || (T <= 0 && N < 0)) {
      for (int j = 0; j < dim; j++) {
        double x = jar[i + j] + a * jv[i + j], D = d->efc_D[i + j];
        c += 0.5 * D * x * x; g += D * x * jv[i + j]; h += D * jv[i + j] * jv[i + j];
      }
    } else {
      double Dm = d->efc_D[i] / fmax(mu * mu * (1 + mu * mu), OM_MINVAL);
      double NT = N - mu * T, T1 = UV / T, T2d = (VV - T1 * T1) / T;
      double NT1 = N1 - mu * T1;
      c += 0.5 * Dm * NT * NT; g += Dm * NT * NT1; h += Dm * (NT1 * NT1 - NT * mu * T2d);
    }
    i += dim - 1;
  }
  *p0 = c; *p1 = g; *p2 = h;
}

/* exact line search on the convex piecewise-smooth 1-D cost: safeguarded Newton on p'(a)=0.
   p' is monotone but jumps at the kinks of the constraint laws, so the minimiser can sit ON a kink where no
   point has |p'| < gtol: the bracket [lo, hi] is then shrunk (Newton step if it stays inside, else secant,
   bisection every other step so that the width at least halves) until it is numerically a point, and the
   left end (p' < 0, guaranteed decrease) is returned. */
static double line_search(const om_model* m, const om_data* d, const double* jar, const double* jv, double g1,
                          double g2, double gtol) {
  double lo = 0, hi = -1, a = 0, p0, p1, p2, dlo, dhi = 0, d0;
  ls_eval(m, d, jar, jv, g1, g2, 0, &p0, &p1, &p2);
  if (p1 >= 0 || p2 <= 0) return 0;
  dlo = d0 = p1;
  a = -p1 / p2;
  int niter = m->ls_iterations > 100 ? m->ls_iterations : 100;
  for (int it = 0; it < niter; it++) {
    ls_eval(m, d, jar, jv, g1, g2, a, &p0, &p1, &p2);
    if (fabs(p1) < gtol || fabs(p1) <= 1e-14 * fabs(d0)) return a;
    if (p1 < 0) { lo = a; dlo = p1; } else { hi = a; dhi = p1; }
    if (hi > 0 && hi - lo <= 4e-16 * hi) return lo > 0 ? lo : a;
    double an = (p2 > 0) ? a - p1 / p2 : -1;
    if (hi < 0) {
      if (!(an > lo)) an = 2 * a + 1e-12;
    } else if (!(an > lo && an < hi) || (it & 1)) {
      double sec = lo + (hi - lo) * (-dlo) / (dhi - dlo);   /* secant on the monotone derivative */
      an = ((it & 1) || !(sec > lo && sec < hi)) ? 0.5 * (lo + hi) : sec;
    }
    a = an;
  }
  return lo > 0 ? lo : a;
}

/* grad = M qacc - qfrc_smooth - J^T force; also refreshes qfrc_constraint. Returns |grad|. */
static double solver_gradient(const om_model* m, om_data* d, const double* Ma, const double* force, double* grad) {
  int nv = m->nv, nefc = d->nefc;
  double g2 = 0;
  for (int i = 0; i < nv; i++) {
    double s = 0;
    for (int r = 0; r < nefc; r++) s += d->efc_J[(size_t)r * nv + i] * force[r];
    d->qfrc_constraint[i] = s;
    grad[i] = Ma[i] - d->qfrc_smooth[i] - s;
    g2 += grad[i] * grad[i];
  }
  return sqrt(g2);
}

/* H = M + J^T D J over quadratic rows + exact elliptic-cone blocks; Cholesky-factored in place */
static void solver_hessian(const om_model* m, om_data* d, const double* jar, const int* state) {
  int nv = m->nv, nefc = d->nefc;
  memcpy(d->H, d->M, sizeof(double) * nv * nv);
  for (int r = 0; r < nefc; r++) {
    if (state[r] == ST_QUADRATIC) {
      const double* J = d->efc_J + (size_t)r * nv;
      double D = d->efc_D[r];
      for (int i = 0; i < nv; i++) {
        if (J[i] == 0) continue;
        double s = D * J[i];
        for (int k = 0; k <= i; k++) d->H[i * nv + k] += s * J[k];
      }
    } else if (state[r] == ST_CONE) {
      om_contact* con = &d->contact[d->efc_id[r]];
      int dim = con->dim;
      double mu = con->mu, U[6], sc[6], T2 = 0, Hc[36];
      sc[0] = mu; U[0] = jar[r] * mu;
      for (int j = 1; j < dim; j++) { sc[j] = con->friction[j - 1]; U[j] = jar[r + j] * sc[j]; T2 += U[j] * U[j]; }
      double N = U[0], T = sqrt(T2), Dm = d->efc_D[r] / fmax(mu * mu * (1 + mu * mu), OM_MINVAL);
      Hc[0] = Dm;
      for (int j = 1; j < dim; j++) Hc[j] = Hc[j * dim] = -Dm * mu * U[j] / T;
      for (int j = 1; j < dim; j++)
        for (int k = 1; k < dim; k++)
          Hc[j * dim + k] = Dm * (mu * N * U[j] * U[k] / (T * T * T) - (j == k ? mu * (N - mu * T) / T : 0));
      for (int j = 0; j < dim; j++) for (int k = 0; k < dim; k++) Hc[j * dim + k] *= sc[j] * sc[k];
      for (int j = 0; j < dim; j++) {
        double* tmp = d->sol_tmp; /* tmp = sum_k Hc[j,k] J_k */
        memset(tmp, 0, sizeof(double) * nv);
        for (int k = 0; k < dim; k++) {
          const double* Jk = d->efc_J + (size_t)(r + k) * nv;
          double h = Hc[j * dim + k];
          for (int i = 0; i < nv; i++) tmp[i] += h * Jk[i];
        }
        const double* Jj = d->efc_J + (size_t)(r + j) * nv;
        for (int i = 0; i < nv; i++) {
          if (Jj[i] == 0) continue;
          for (int k = 0; k <= i; k++) d->H[i * nv + k] += Jj[i] * tmp[k];
        }
      }
      r += dim - 1;
    }
  }
  for (int i = 0; i < nv; i++) for (int k = 0; k < i; k++) d->H[k * nv + i] = d->H[i * nv + k];
  chol_factor(d->H, nv);
}

/* Newton with exact Hessian incl. elliptic-cone blocks [upstream mj_solNewton] */
static void om_solve(const om_model* m, om_data* d) {
  int nv = m->nv, nefc = d->nefc;
  if (nefc == 0) {
    memcpy(d->qacc, d->qacc_smooth, sizeof(double) * nv);
    memset(d->qfrc_constraint, 0, sizeof(double) * nv);
    d->solver_iter = 0;
    return;
  }
  double *jar = d->sol_jar, *force = d->efc_force, *Ma = d->sol_Ma, *grad = d->sol_grad, *search = d->sol_search,
         *mv = d->sol_mv, *jv = d->sol_jv;
  int* state = d->efc_state;
  double scale = 1.0 / (m->meaninertia * (nv > 1 ? nv : 1));
  /* warm start: keep qacc_warmstart only if it beats qacc_smooth */
  {
    double cw, cs;
    memcpy(d->qacc, d->qacc_warmstart, sizeof(double) * nv);
    mulmat(jar, d->efc_J, d->qacc, nefc, nv);
    for (int i = 0; i < nefc; i++) jar[i] -= d->efc_aref[i];
    cw = constraint_update(m, d, jar, force, state);
    mulmat(Ma, d->M, d->qacc, nv, nv);
    for (int i = 0; i < nv; i++) cw += 0.5 * (Ma[i] - d->qfrc_smooth[i]) * (d->qacc[i] - d->qacc_smooth[i]);
    mulmat(jar, d->efc_J, d->qacc_smooth, nefc, nv);
    for (int i = 0; i < nefc; i++) jar[i] -= d->efc_aref[i];
    cs = constraint_update(m, d, jar, force, state);
    if (cw > cs) memcpy(d->qacc, d->qacc_smooth, sizeof(double) * nv);
  }
  mulmat(Ma, d->M, d->qacc, nv, nv);
  mulmat(jar, d->efc_J, d->qacc, nefc, nv);
  for (int i = 0; i < nefc; i++) jar[i] -= d->efc_aref[i];
  double cost = constraint_update(m, d, jar, force, state);
  for (int i = 0; i < nv; i++) cost += 0.5 * (Ma[i] - d->qfrc_smooth[i]) * (d->qacc[i] - d->qacc_smooth[i]);
  int iter = 0;
  double gnorm = solver_gradient(m, d, Ma, force, grad);
  while (iter < m->iterations) {
    solver_hessian(m, d, jar, state);
    for (int i = 0; i < nv; i++) search[i] = -grad[i];
    chol_solve(d->H, search, nv);
    /* line search */
    mulmat(mv, d->M, search, nv, nv);
    mulmat(jv, d->efc_J, search, nefc, nv);
    double g1 = 0, g2 = 0, snorm = 0;
    for (int i = 0; i < nv; i++) { g1 += search[i] * (Ma[i] - d->qfrc_smooth[i]); g2 += search[i] * mv[i]; snorm += search[i] * search[i]; }
    snorm = sqrt(snorm);
    if (snorm < OM_MINVAL) break;
    double gtol = m->tolerance * m->ls_tolerance * snorm / scale;
    double alpha = line_search(m, d, jar, jv, g1, g2, gtol);
    if (alpha == 0) break;
    for (int i = 0; i < nv; i++) { d->qacc[i] += alpha * search[i]; Ma[i] += alpha * mv[i]; }
    for (int i = 0; i < nefc; i++) jar[i] += alpha * jv[i];
    double oldcost = cost;
    cost = constraint_update(m, d, jar, force, state);
    for (int i = 0; i < nv; i++) cost += 0.5 * (Ma[i] - d->qfrc_smooth[i]) * (d->qacc[i] - d->qacc_smooth[i]);
    iter++;
    gnorm = solver_gradient(m, d, Ma, force, grad);
    if (scale * (oldcost - cost) < m->tolerance || scale * gnorm < m->tolerance) break;
  }
  d->solver_iter = iter;
}

/* ------------------------------------------------------------------------- sensors */

static void om_sensors(const om_model* m, om_data* d) {
  for (int s = 0; s < m->nsensor; s++) {
    int type = m->sensor_type[s], site = m->sensor_objid[s], adr = m->sensor_adr[s];
    int b = m->site_bodyid[site];
    const double *R = d->site_xmat + 9 * site, *p = d->site_xpos + 3 * site;
    if (type == SENS_GYRO) {
      multmatvec3(d->sensordata + adr, R, d->cvel + 6 * b);
    } else if (type == SENS_ACCEL) {
      /* site acceleration from cacc (computed with qacc) + Coriolis correction [upstream mj_objectAcceleration] */
      double dif[3], lin[3], t[3], vlin[3], wl[3], vl[3], al[3];
      v3sub(dif, p, d->subtree_com + 3 * m->body_rootid[b]);
      const double *ca = d->cacc + 6 * b, *cv = d->cvel + 6 * b;
      v3cross(t, ca, dif); v3add(lin, ca + 3, t);
      v3cross(t, cv, dif); v3add(vlin, cv + 3, t);
      multmatvec3(al, R, lin); multmatvec3(wl, R, cv); multmatvec3(vl, R, vlin);
      v3cross(t, wl, vl);
      v3add(d->sensordata + adr, al, t);
    } else if (type == SENS_RANGE) {
      if (!m->enable_lidar || !m->has_ray) { d->sensordata[adr] = -1; continue; }
      double dir[3] = {R[2], R[5], R[8]};
      int geom;
      double dist = om_ray(m, d->xpos, d->xmat, d->geom_xpos, d->geom_xmat, p, dir, 0, b, &geom);
      if (dist >= 0 && m->sensor_cutoff[s] > 0 && dist > m->sensor_cutoff[s]) dist = m->sensor_cutoff[s];
      d->sensordata[adr] = dist;
    }
  }
}

/* ------------------------------------------------------------------------- forward + integrate */

static int bad_state(const double* x, int n) {
  for (int i = 0; i < n; i++) if (!(fabs(x[i]) < OM_MAXVAL)) return 1;
  return 0;
}

static void reset_env(const om_model* m, om_data* d) {
  memcpy(d->qpos, m->qpos0, sizeof(double) * m->nq);
  memset(d->qvel, 0, sizeof(double) * m->nv);
  memset(d->qacc_warmstart, 0, sizeof(double) * m->nv);
  d->flags |= 1;
}

void om_forward(const om_model* m, om_data* d) {
  int nv = m->nv;
  om_kinematics(m, d);
  om_compos(m, d);
  om_transmission(m, d);
  om_crb(m, d);
  om_collision(m, d);
  om_make_constraint(m, d);
  om_comvel(m, d);
  om_passive(m, d);
  om_impedance_and_reference(m, d);
  om_rne(m, d, 0, NULL, d->qfrc_bias);
  om_actuation(m, d);
  for (int i = 0; i < nv; i++) {
    d->qfrc_smooth[i] = d->qfrc_passive[i] - d->qfrc_bias[i] + d->qfrc_actuator[i];
    d->qacc_smooth[i] = d->qfrc_smooth[i];
  }
  chol_solve(d->L, d->qacc_smooth, nv);
  om_solve(m, d);
  memcpy(d->qacc_warmstart, d->qacc, sizeof(double) * nv);
  /* sensors: gyro after velocity, accelerometer after constraint (needs cacc with qacc) */
  om_rne(m, d, 1, d->qacc, NULL);
  om_sensors(m, d);
}

/* implicit-in-velocity integration, "fast" variant: symmetric derivative of passive and
 * actuator forces only [upstream mj_implicit with mjINT_IMPLICITFAST] */
static void om_implicitfast(const om_model* m, om_data* d) {
  int nv = m->nv;
  double h = m->timestep;
  double* A = d->H;
  memcpy(A, d->M, sizeof(double) * nv * nv);
  for (int i = 0; i < nv; i++) A[i * nv + i] += h * m->dof_damping[i];
  for (int a = 0; a < m->nu; a++) {
    double b2 = m->actuator_biasprm[3 * a + 2];
    if (b2 == 0) continue;
    if (m->actuator_forcelimited[a] &&
        (d->act_force[a] <= m->actuator_forcerange[2 * a] || d->act_force[a] >= m->actuator_forcerange[2 * a + 1]))
      continue; /* clamped force has zero velocity derivative [upstream mjd_actuator_vel] */
    const double* mom = d->act_moment + a * nv;
    for (int i = 0; i < nv; i++) {
      if (mom[i] == 0) continue;
      for (int k = 0; k < nv; k++) A[i * nv + k] -= h * b2 * mom[i] * mom[k];
    }
  }
  chol_factor(A, nv);
  double* qacc = d->sol_tmp;
  for (int i = 0; i < nv; i++) qacc[i] = d->qfrc_smooth[i] + d->qfrc_constraint[i];
  chol_solve(A, qacc, nv);
  /* advance [upstream mj_advance] */
  for (int i = 0; i < nv; i++) d->qvel[i] += h * qacc[i];
  for (int j = 0; j < m->njnt; j++) {
    int qa = m->jnt_qposadr[j], da = m->jnt_dofadr[j];
    if (m->jnt_type[j] == JNT_FREE) {
      for (int k = 0; k < 3; k++) d->qpos[qa + k] += h * d->qvel[da + k];
      quat_integrate(d->qpos + qa + 3, d->qvel + da + 3, h);
    } else {
      d->qpos[qa] += h * d->qvel[da];
    }
  }
  d->time += h;
}

void om_step1(const om_model* m, om_data* d) {
  if (bad_state(d->qpos, m->nq) || bad_state(d->qvel, m->nv)) reset_env(m, d);
  om_forward(m, d);
  if (bad_state(d->qacc, m->nv)) { reset_env(m, d); om_forward(m, d); }
  om_implicitfast(m, d);
}

/* ------------------------------------------------------------------------- batch drivers */

static void copy_outputs(const om_model* m, const om_data* d, om_outputs* o, int e) {
  if (!o) return;
  int nv = m->nv, nb = m->nbody;
  if (o->xpos) memcpy(o->xpos + (size_t)e * nb * 3, d->xpos, sizeof(double) * nb * 3);
  if (o->xquat) memcpy(o->xquat + (size_t)e * nb * 4, d->xquat, sizeof(double) * nb * 4);
  if (o->act_length) memcpy(o->act_length + (size_t)e * m->nu, d->act_length, sizeof(double) * m->nu);
  if (o->act_velocity) memcpy(o->act_velocity + (size_t)e * m->nu, d->act_velocity, sizeof(double) * m->nu);
  if (o->sensordata) memcpy(o->sensordata + (size_t)e * m->nsensordata, d->sensordata, sizeof(double) * m->nsensordata);
  if (o->qacc) memcpy(o->qacc + (size_t)e * nv, d->qacc, sizeof(double) * nv);
  if (o->qfrc_constraint) memcpy(o->qfrc_constraint + (size_t)e * nv, d->qfrc_constraint, sizeof(double) * nv);
  if (o->qacc_smooth) memcpy(o->qacc_smooth + (size_t)e * nv, d->qacc_smooth, sizeof(double) * nv);
  if (o->qfrc_bias) memcpy(o->qfrc_bias + (size_t)e * nv, d->qfrc_bias, sizeof(double) * nv);
  if (o->qfrc_passive) memcpy(o->qfrc_passive + (size_t)e * nv, d->qfrc_passive, sizeof(double) * nv);
  if (o->qfrc_actuator) memcpy(o->qfrc_actuator + (size_t)e * nv, d->qfrc_actuator, sizeof(double) * nv);
  if (o->M) memcpy(o->M + (size_t)e * nv * nv, d->M, sizeof(double) * nv * nv);
  if (o->ncon) o->ncon[e] = d->ncon;
  if (o->nefc) o->nefc[e] = d->nefc;
  if (o->solver_iter) o->solver_iter[e] = d->solver_iter;
  if (o->flags) o->flags[e] = d->flags;
  for (int c = 0; c < o->maxcon; c++) {
    int live = c < d->ncon;
    size_t k = (size_t)e * o->maxcon + c;
    if (o->contact_geom) { o->contact_geom[2 * k] = live ? d->contact[c].geom1 : -1; o->contact_geom[2 * k + 1] = live ? d->contact[c].geom2 : -1; }
    if (o->contact_dist) o->contact_dist[k] = live ? d->contact[c].dist : 0;
    if (o->contact_pos) for (int a = 0; a < 3; a++) o->contact_pos[3 * k + a] = live ? d->contact[c].pos[a] : 0;
    if (o->contact_frame) for (int a = 0; a < 3; a++) o->contact_frame[3 * k + a] = live ? d->contact[c].frame[a] : 0;
  }
  for (int r = 0; r < o->maxefc_out; r++) {
    int live = r < d->nefc;
    size_t k = (size_t)e * o->maxefc_out + r;
    if (o->efc_J) for (int i = 0; i < nv; i++) o->efc_J[k * nv + i] = live ? d->efc_J[(size_t)r * nv + i] : 0;
    if (o->efc_aref) o->efc_aref[k] = live ? d->efc_aref[r] : 0;
    if (o->efc_D) o->efc_D[k] = live ? d->efc_D[r] : 0;
    if (o->efc_force) o->efc_force[k] = live ? d->efc_force[r] : 0;
  }
}

typedef struct {
  int nenv, nsteps, ctrl_per_step;
  double *qpos, *qvel, *warm, *time;
  const double* ctrl;
  om_outputs* out;
} step_ctx;

static void step_env(const om_model* m, om_data* d, int e, void* vctx) {
  step_ctx* c = (step_ctx*)vctx;
  d->qpos = c->qpos + (size_t)e * m->nq; d->qvel = c->qvel + (size_t)e * m->nv;
  d->qacc_warmstart = c->warm + (size_t)e * m->nv;
  d->time = c->time ? c->time[e] : 0; d->flags = 0;
  for (int s = 0; s < c->nsteps; s++) {
    d->ctrl = c->ctrl + (c->ctrl_per_step ? ((size_t)s * c->nenv + e) * m->nu : (size_t)e * m->nu);
    om_step1(m, d);
  }
  if (c->time) c->time[e] = d->time;
  copy_outputs(m, d, c->out, e);
}

int om_batch_step(const om_model* m, int nenv, int nsteps, double* qpos, double* qvel, const double* ctrl,
                  int ctrl_per_step, double* warm, double* time, om_outputs* out, int nthreads) {
  step_ctx c = {nenv, nsteps, ctrl_per_step, qpos, qvel, warm, time, ctrl, out};
  om_parallel_for(m, nenv, nthreads, step_env, &c);
  return 0;
}

typedef struct {
  const double *qpos, *qvel, *ctrl, *warm;
  om_outputs* out;
} fwd_ctx;

static void fwd_env(const om_model* m, om_data* d, int e, void* vctx) {
  fwd_ctx* c = (fwd_ctx*)vctx;
  double q[7 * OM_MAXNV];
  memcpy(q, c->qpos + (size_t)e * m->nq, sizeof(double) * m->nq);
  memcpy(q + m->nq, c->qvel + (size_t)e * m->nv, sizeof(double) * m->nv);
  if (c->warm) memcpy(q + m->nq + m->nv, c->warm + (size_t)e * m->nv, sizeof(double) * m->nv);
  else memset(q + m->nq + m->nv, 0, sizeof(double) * m->nv);
  d->qpos = q; d->qvel = q + m->nq; d->qacc_warmstart = q + m->nq + m->nv;
  d->ctrl = c->ctrl + (size_t)e * m->nu; d->flags = 0;
  om_forward(m, d);
  copy_outputs(m, d, c->out, e);
}

int om_batch_forward(const om_model* m, int nenv, const double* qpos, const double* qvel, const double* ctrl,
                     const double* warm, om_outputs* out, int nthreads) {
  fwd_ctx c = {qpos, qvel, ctrl, warm, out};
  om_parallel_for(m, nenv, nthreads, fwd_env, &c);
  return 0;
}
