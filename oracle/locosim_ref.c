/*
 * CPU ORACLE (test infrastructure, NOT the product) -- see locosim_ref.h for scope and provenance.
 *
 * Plain C, fp64, one environment per RefSim. Restates MuJoCo 2.3.7's mj_step for hinge/slide trees:
 *   mj_fwdPosition : kinematics, comPos, crb, factorM, collision, makeConstraint
 *   mj_fwdVelocity : comVel, passive, referenceConstraint, rne
 *   mj_fwdActuation, mj_fwdAcceleration, mj_fwdConstraint (Newton solver, elliptic + pyramidal cones)
 *   mj_Euler (implicit-in-damping) / mj_RungeKutta(4)
 * Section comments name the MuJoCo routine being restated; the reference call site is
 * /root/reference/loco_mujoco/environments/base.py:25,109-111 (inherited MultiMuJoCo.step -> mujoco.mj_step).
 */
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <math.h>
#include <sched.h>
#include <time.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#include "../include/locosim_modelpack.h"
#include "../include/locosim_task.h"
#include "locosim_ref.h"

#define MAXNV 40
#define MAXBODY 64
#define MAXGEOM 256
#define MAXCON 128
#define MAXEFC 640
#define MAXCONPAIR 8

#define mjMINVAL 1e-15
#define mjMINIMP 0.0001
#define mjMAXIMP 0.9999
#define mjPI 3.14159265358979323846

enum { CNSTR_FRICTION_DOF = 1, CNSTR_LIMIT_JOINT = 3, CNSTR_CONTACT_FRICTIONLESS = 5, CNSTR_CONTACT_PYRAMIDAL = 6,
       CNSTR_CONTACT_ELLIPTIC = 7 };
enum { ST_SATISFIED = 0, ST_QUADRATIC, ST_LINEARNEG, ST_LINEARPOS, ST_CONE };

typedef struct {
  int nb, nv, ng, nu, np, nm, integrator, cone, iterations;
  double timestep, gravity[3], impratio, tolerance, meaninertia;
  int* ibuf;
  double* rbuf;
#define X(name, cnt) const int* name;
  LOCOSIM_MP_INT_FIELDS(X)
#undef X
#define X(name, cnt) const double* name;
  LOCOSIM_MP_REAL_FIELDS(X)
#undef X
} Model;

typedef struct {
  double dist, pos[3], frame[9], includemargin, friction[5], solref[2], solimp[5], mu;
  int dim, geom1, geom2, efc_address, exclude;
} Contact;

struct RefSim {
  Model m;
  double qpos[MAXNV], qvel[MAXNV], qacc[MAXNV], qacc_warmstart[MAXNV], ctrl[MAXNV];
  double qfrc_passive[MAXNV], qfrc_bias[MAXNV], qfrc_actuator[MAXNV], qfrc_smooth[MAXNV], qacc_smooth[MAXNV],
      qfrc_constraint[MAXNV];
  double xpos[MAXBODY][3], xquat[MAXBODY][4], xmat[MAXBODY][9], xipos[MAXBODY][3], ximat[MAXBODY][9];
  double xanchor[MAXNV][3], xaxis[MAXNV][3];
  double gxpos[MAXGEOM][3], gxmat[MAXGEOM][9];
  double subtree_com[MAXBODY][3], cinert[MAXBODY][10], crb[MAXBODY][10];
  double cdof[MAXNV][6], cdof_dot[MAXNV][6], cvel[MAXBODY][6], cacc[MAXBODY][6], cfrc[MAXBODY][6];
  double M[MAXNV][MAXNV], LM[MAXNV][MAXNV];
  int ncon, nf, nl, nefc, solver_iter;
  Contact con[MAXCON];
  double J[MAXEFC][MAXNV];
  double efc_pos[MAXEFC], efc_margin[MAXEFC], efc_frictionloss[MAXEFC], efc_diagApprox[MAXEFC], efc_KBIP[MAXEFC][4],
      efc_D[MAXEFC], efc_R[MAXEFC], efc_vel[MAXEFC], efc_aref[MAXEFC], efc_force[MAXEFC], efc_b[MAXEFC];
  int efc_type[MAXEFC], efc_id[MAXEFC], efc_state[MAXEFC];
};

/* ---------------------------------------------------------------------------------------------------- */
/* small vector helpers                                                                                 */
/* ---------------------------------------------------------------------------------------------------- */
static inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void cross3(double* r, const double* a, const double* b) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline double norm3(const double* a) { return sqrt(dot3(a, a)); }
static inline double normalize3(double* a) {
  double n = norm3(a);
  if (n < mjMINVAL) { a[0] = 1; a[1] = 0; a[2] = 0; }
  else { a[0] /= n; a[1] /= n; a[2] /= n; }
  return n;
}
static inline void mulmatvec3(double* r, const double* m, const double* v) {
  double x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2], y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2],
         z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline void mulmatTvec3(double* r, const double* m, const double* v) {
  double x = m[0] * v[0] + m[3] * v[1] + m[6] * v[2], y = m[1] * v[0] + m[4] * v[1] + m[7] * v[2],
         z = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline void mulmat3(double* r, const double* a, const double* b) {
  double t[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) t[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
  memcpy(r, t, sizeof(t));
}
static inline void quat2mat(double* m, const double* q) {
  double q00 = q[0] * q[0], q01 = q[0] * q[1], q02 = q[0] * q[2], q03 = q[0] * q[3], q11 = q[1] * q[1],
         q12 = q[1] * q[2], q13 = q[1] * q[3], q22 = q[2] * q[2], q23 = q[2] * q[3], q33 = q[3] * q[3];
  m[0] = q00 + q11 - q22 - q33; m[4] = q00 - q11 + q22 - q33; m[8] = q00 - q11 - q22 + q33;
  m[1] = 2 * (q12 - q03); m[2] = 2 * (q13 + q02); m[3] = 2 * (q12 + q03);
  m[5] = 2 * (q23 - q01); m[6] = 2 * (q13 - q02); m[7] = 2 * (q23 + q01);
}
static inline void mulquat(double* r, const double* a, const double* b) {
  double t[4] = {a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                 a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1], a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]};
  memcpy(r, t, sizeof(t));
}
static inline void normquat(double* q) {
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < mjMINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; }
  else if (fabs(n - 1) > mjMINVAL) { q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n; }
}
static inline void axisangle2quat(double* q, const double* axis, double angle) {
  if (angle == 0) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
  double s = sin(angle * 0.5);
  q[0] = cos(angle * 0.5); q[1] = axis[0] * s; q[2] = axis[1] * s; q[3] = axis[2] * s;
}

/* spatial (com-based) helpers: vectors are [rot(3); lin(3)], inertia is 10 numbers (mju_inertCom layout) */
static void mulInertVec(double* res, const double* i, const double* v) {
  res[0] = i[0] * v[0] + i[3] * v[1] + i[4] * v[2] - i[8] * v[4] + i[7] * v[5];
  res[1] = i[3] * v[0] + i[1] * v[1] + i[5] * v[2] + i[8] * v[3] - i[6] * v[5];
  res[2] = i[4] * v[0] + i[5] * v[1] + i[2] * v[2] - i[7] * v[3] + i[6] * v[4];
  res[3] = i[8] * v[1] - i[7] * v[2] + i[9] * v[3];
  res[4] = i[6] * v[2] - i[8] * v[0] + i[9] * v[4];
  res[5] = i[7] * v[0] - i[6] * v[1] + i[9] * v[5];
}
static void crossMotion(double* res, const double* vel, const double* v) {
  double a[3], b[3];
  cross3(res, vel, v);
  cross3(a, vel, v + 3);
  cross3(b, vel + 3, v);
  res[3] = a[0] + b[0]; res[4] = a[1] + b[1]; res[5] = a[2] + b[2];
}
static void crossForce(double* res, const double* vel, const double* f) {
  double a[3], b[3];
  cross3(a, vel, f);
  cross3(b, vel + 3, f + 3);
  res[0] = a[0] + b[0]; res[1] = a[1] + b[1]; res[2] = a[2] + b[2];
  cross3(res + 3, vel, f + 3);
}

/* ---------------------------------------------------------------------------------------------------- */
/* model loading                                                                                        */
/* ---------------------------------------------------------------------------------------------------- */
static int model_load(Model* m, const int* ints, int n_ints, const double* reals, int n_reals) {
  if (n_ints < MPI_HEADER_LEN || ints[MPI_MAGIC] != LOCOSIM_MP_MAGIC || ints[MPI_VERSION] != LOCOSIM_MP_VERSION) return -1;
  m->ibuf = (int*)malloc(sizeof(int) * n_ints);
  m->rbuf = (double*)malloc(sizeof(double) * n_reals);
  memcpy(m->ibuf, ints, sizeof(int) * n_ints);
  memcpy(m->rbuf, reals, sizeof(double) * n_reals);
  int nb = m->nb = ints[MPI_NBODY], nv = m->nv = ints[MPI_NV], ng = m->ng = ints[MPI_NGEOM], nu = m->nu = ints[MPI_NU],
      np = m->np = ints[MPI_NPAIR], nm = m->nm = ints[MPI_NMESHVERT];
  m->integrator = ints[MPI_INTEGRATOR]; m->cone = ints[MPI_CONE]; m->iterations = ints[MPI_ITERATIONS];
  m->timestep = reals[MPR_TIMESTEP];
  m->gravity[0] = reals[MPR_GRAV_X]; m->gravity[1] = reals[MPR_GRAV_Y]; m->gravity[2] = reals[MPR_GRAV_Z];
  m->impratio = reals[MPR_IMPRATIO]; m->tolerance = reals[MPR_TOLERANCE]; m->meaninertia = reals[MPR_MEANINERTIA];
  if (nb > MAXBODY || nv > MAXNV || ng > MAXGEOM) return -2;
  const int* ip = m->ibuf + MPI_HEADER_LEN;
  const double* rp = m->rbuf + MPR_HEADER_LEN;
#define X(name, cnt) m->name = ip; ip += (cnt);
  LOCOSIM_MP_INT_FIELDS(X)
#undef X
#define X(name, cnt) m->name = rp; rp += (cnt);
  LOCOSIM_MP_REAL_FIELDS(X)
#undef X
  if (ip - m->ibuf != n_ints || rp - m->rbuf != n_reals) return -3;
  return 0;
}

RefSim* ref_create(const int* ints, int n_ints, const double* reals, int n_reals) {
  RefSim* s = (RefSim*)calloc(1, sizeof(RefSim));
  int rc = model_load(&s->m, ints, n_ints, reals, n_reals);
  if (rc) { fprintf(stderr, "locosim_ref: bad ModelPack (%d)\n", rc); free(s); return NULL; }
  for (int i = 0; i < s->m.nv; i++) s->qpos[i] = s->m.qpos0[i];
  return s;
}
void ref_destroy(RefSim* s) { if (s) { free(s->m.ibuf); free(s->m.rbuf); free(s); } }
int ref_nv(const RefSim* s) { return s->m.nv; }
int ref_nu(const RefSim* s) { return s->m.nu; }
void ref_reset(RefSim* s, const double* qpos, const double* qvel) {
  int nv = s->m.nv;
  memset(s->qacc_warmstart, 0, sizeof(s->qacc_warmstart));
  memset(s->qacc, 0, sizeof(s->qacc));
  memset(s->ctrl, 0, sizeof(s->ctrl));
  for (int i = 0; i < nv; i++) { s->qpos[i] = qpos ? qpos[i] : s->m.qpos0[i]; s->qvel[i] = qvel ? qvel[i] : 0; }
}
void ref_get_state(const RefSim* s, double* qpos, double* qvel) {
  memcpy(qpos, s->qpos, sizeof(double) * s->m.nv); memcpy(qvel, s->qvel, sizeof(double) * s->m.nv);
}
void ref_get_warmstart(const RefSim* s, double* w) { memcpy(w, s->qacc_warmstart, sizeof(double) * s->m.nv); }
void ref_set_warmstart(RefSim* s, const double* w) { memcpy(s->qacc_warmstart, w, sizeof(double) * s->m.nv); }
int ref_ncon(const RefSim* s) { return s->ncon; }
int ref_nefc(const RefSim* s) { return s->nefc; }
int ref_solver_iter(const RefSim* s) { return s->solver_iter; }
void ref_get_qacc(const RefSim* s, double* qacc) { memcpy(qacc, s->qacc, sizeof(double) * s->m.nv); }
void ref_get_M(const RefSim* s, double* M) {
  int nv = s->m.nv;
  for (int i = 0; i < nv; i++) for (int j = 0; j < nv; j++) M[i * nv + j] = s->M[i][j];
}
void ref_get_bias(const RefSim* s, double* b) { memcpy(b, s->qfrc_bias, sizeof(double) * s->m.nv); }
void ref_get_contact(const RefSim* s, int k, double* o) {
  const Contact* c = &s->con[k];
  o[0] = c->dist; o[1] = c->pos[0]; o[2] = c->pos[1]; o[3] = c->pos[2];
  o[4] = c->frame[0]; o[5] = c->frame[1]; o[6] = c->frame[2]; o[7] = c->geom1; o[8] = c->geom2; o[9] = c->dim;
  for (int j = 0; j < 6; j++) o[10 + j] = 0;
  if (c->efc_address >= 0) {
    if (c->dim == 1 || s->m.cone == 1) for (int j = 0; j < c->dim; j++) o[10 + j] = s->efc_force[c->efc_address + j];
    else {  /* mju_decodePyramid */
      const double* p = s->efc_force + c->efc_address;
      for (int j = 0; j < 2 * (c->dim - 1); j++) o[10] += p[j];
      for (int j = 0; j < c->dim - 1; j++) o[11 + j] = (p[2 * j] - p[2 * j + 1]) * c->friction[j];
    }
  }
}

/* ---------------------------------------------------------------------------------------------------- */
/* mj_kinematics                                                                                        */
/* ---------------------------------------------------------------------------------------------------- */
static void kinematics(RefSim* s) {
  const Model* m = &s->m;
  s->xpos[0][0] = s->xpos[0][1] = s->xpos[0][2] = 0;
  s->xquat[0][0] = 1; s->xquat[0][1] = s->xquat[0][2] = s->xquat[0][3] = 0;
  quat2mat(s->xmat[0], s->xquat[0]);
  memcpy(s->ximat[0], s->xmat[0], sizeof(double) * 9);
  s->xipos[0][0] = s->xipos[0][1] = s->xipos[0][2] = 0;
  for (int b = 1; b < m->nb; b++) {
    int p = m->body_parentid[b];
    double pos[3], quat[4], tmp[3];
    mulmatvec3(tmp, s->xmat[p], m->body_pos + 3 * b);
    for (int k = 0; k < 3; k++) pos[k] = s->xpos[p][k] + tmp[k];
    mulquat(quat, s->xquat[p], m->body_quat + 4 * b);
    for (int k = 0; k < m->body_jntnum[b]; k++) {
      int j = m->body_jntadr[b] + k;
      double mat[9];
      quat2mat(mat, quat);
      mulmatvec3(tmp, mat, m->jnt_pos + 3 * j);
      for (int c = 0; c < 3; c++) s->xanchor[j][c] = pos[c] + tmp[c];
      mulmatvec3(s->xaxis[j], mat, m->jnt_axis + 3 * j);
      double q = s->qpos[j] - m->qpos0[j];
      if (m->jnt_type[j] == LS_JNT_SLIDE) {
        for (int c = 0; c < 3; c++) pos[c] += s->xaxis[j][c] * q;
      } else {
        double ql[4];
        axisangle2quat(ql, m->jnt_axis + 3 * j, q);
        mulquat(quat, quat, ql);
        quat2mat(mat, quat);
        mulmatvec3(tmp, mat, m->jnt_pos + 3 * j);
        for (int c = 0; c < 3; c++) pos[c] = s->xanchor[j][c] - tmp[c];
      }
    }
    normquat(quat);
    memcpy(s->xpos[b], pos, sizeof(pos));
    memcpy(s->xquat[b], quat, sizeof(quat));
    quat2mat(s->xmat[b], quat);
    mulmatvec3(tmp, s->xmat[b], m->body_ipos + 3 * b);
    for (int c = 0; c < 3; c++) s->xipos[b][c] = pos[c] + tmp[c];
    double iq[4], im[9];
    mulquat(iq, quat, m->body_iquat + 4 * b);
    quat2mat(im, iq);
    memcpy(s->ximat[b], im, sizeof(im));
  }
  for (int g = 0; g < m->ng; g++) {
    int b = m->geom_bodyid[g];
    double tmp[3], gm[9];
    mulmatvec3(tmp, s->xmat[b], m->geom_pos + 3 * g);
    for (int c = 0; c < 3; c++) s->gxpos[g][c] = s->xpos[b][c] + tmp[c];
    double gq[4];
    mulquat(gq, s->xquat[b], m->geom_quat + 4 * g);
    quat2mat(gm, gq);
    memcpy(s->gxmat[g], gm, sizeof(gm));
  }
}

/* ---------------------------------------------------------------------------------------------------- */
/* mj_comPos : subtree CoM, com-based body inertias (cinert) and motion dofs (cdof)                      */
/* ---------------------------------------------------------------------------------------------------- */
static void comPos(RefSim* s) {
  const Model* m = &s->m;
  double mass_subtree[MAXBODY];
  for (int b = 0; b < m->nb; b++) {
    mass_subtree[b] = m->body_mass[b];
    for (int c = 0; c < 3; c++) s->subtree_com[b][c] = m->body_mass[b] * s->xipos[b][c];
  }
  for (int b = m->nb - 1; b > 0; b--) {
    int p = m->body_parentid[b];
    mass_subtree[p] += mass_subtree[b];
    for (int c = 0; c < 3; c++) s->subtree_com[p][c] += s->subtree_com[b][c];
  }
  for (int b = 0; b < m->nb; b++) {
    if (mass_subtree[b] < mjMINVAL) memcpy(s->subtree_com[b], s->xipos[b], sizeof(double) * 3);
    else for (int c = 0; c < 3; c++) s->subtree_com[b][c] /= mass_subtree[b];
  }
  for (int b = 1; b < m->nb; b++) {
    /* mju_inertCom: inertia about subtree_com[root] in world orientation */
    const double* com = s->subtree_com[m->body_rootid[b]];
    double dif[3] = {s->xipos[b][0] - com[0], s->xipos[b][1] - com[1], s->xipos[b][2] - com[2]};
    const double* R = s->ximat[b];
    const double* I = m->body_inertia + 3 * b;
    double mass = m->body_mass[b];
    double tmp[9], res[9];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) tmp[3 * r + c] = R[3 * r + c] * I[c];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++)
      res[3 * r + c] = tmp[3 * r] * R[3 * c] + tmp[3 * r + 1] * R[3 * c + 1] + tmp[3 * r + 2] * R[3 * c + 2];
    double* ci = s->cinert[b];
    ci[0] = res[0] + mass * (dif[1] * dif[1] + dif[2] * dif[2]);
    ci[1] = res[4] + mass * (dif[0] * dif[0] + dif[2] * dif[2]);
    ci[2] = res[8] + mass * (dif[0] * dif[0] + dif[1] * dif[1]);
    ci[3] = res[1] - mass * dif[0] * dif[1];
    ci[4] = res[2] - mass * dif[0] * dif[2];
    ci[5] = res[5] - mass * dif[1] * dif[2];
    ci[6] = mass * dif[0]; ci[7] = mass * dif[1]; ci[8] = mass * dif[2]; ci[9] = mass;
  }
  memset(s->cinert[0], 0, sizeof(double) * 10);
  for (int j = 0; j < m->nv; j++) {
    int b = m->jnt_bodyid[j];
    const double* com = s->subtree_com[m->body_rootid[b]];
    double* cd = s->cdof[j];
    if (m->jnt_type[j] == LS_JNT_SLIDE) {
      cd[0] = cd[1] = cd[2] = 0;
      memcpy(cd + 3, s->xaxis[j], sizeof(double) * 3);
    } else {
      double off[3] = {com[0] - s->xanchor[j][0], com[1] - s->xanchor[j][1], com[2] - s->xanchor[j][2]};
      memcpy(cd, s->xaxis[j], sizeof(double) * 3);
      cross3(cd + 3, s->xaxis[j], off);
    }
  }
}

/* ---------------------------------------------------------------------------------------------------- */
/* mj_crb + mj_factorM (dense storage; same numbers as MuJoCo's sparse qM / qLD up to rounding)          */
/* ---------------------------------------------------------------------------------------------------- */
static int chol_factor(double L[MAXNV][MAXNV], int n) {
  int rank = n;
  for (int j = 0; j < n; j++) {
    double t = L[j][j];
    for (int k = 0; k < j; k++) t -= L[j][k] * L[j][k];
    if (t < mjMINVAL) { t = mjMINVAL; rank--; }
    L[j][j] = sqrt(t);
    double inv = 1.0 / L[j][j];
    for (int i = j + 1; i < n; i++) {
      double v = L[i][j];
      for (int k = 0; k < j; k++) v -= L[i][k] * L[j][k];
      L[i][j] = v * inv;
    }
  }
  return rank;
}
static void chol_solve(double L[MAXNV][MAXNV], int n, double* x) {
  for (int i = 0; i < n; i++) {
    double v = x[i];
    for (int k = 0; k < i; k++) v -= L[i][k] * x[k];
    x[i] = v / L[i][i];
  }
  for (int i = n - 1; i >= 0; i--) {
    double v = x[i];
    for (int k = i + 1; k < n; k++) v -= L[k][i] * x[k];
    x[i] = v / L[i][i];
  }
}

static void crb(RefSim* s) {
  const Model* m = &s->m;
  int nv = m->nv;
  for (int b = 0; b < m->nb; b++) memcpy(s->crb[b], s->cinert[b], sizeof(double) * 10);
  for (int b = m->nb - 1; b > 0; b--) {
    int p = m->body_parentid[b];
    if (p > 0) for (int k = 0; k < 10; k++) s->crb[p][k] += s->crb[b][k];
  }
  for (int i = 0; i < nv; i++) for (int j = 0; j < nv; j++) s->M[i][j] = 0;
  for (int i = 0; i < nv; i++) {
    double buf[6];
    mulInertVec(buf, s->crb[m->jnt_bodyid[i]], s->cdof[i]);
    s->M[i][i] = m->dof_armature[i];
    for (int j = i; j >= 0; j = m->dof_parentid[j]) {
      double v = 0;
      for (int k = 0; k < 6; k++) v += s->cdof[j][k] * buf[k];
      s->M[i][j] += v;
      if (j != i) s->M[j][i] = s->M[i][j];
    }
  }
  for (int i = 0; i < nv; i++) for (int j = 0; j <= i; j++) s->LM[i][j] = s->M[i][j];
  chol_factor(s->LM, nv);
}

/* ---------------------------------------------------------------------------------------------------- */
/* collision: narrow-phase functions (engine_collision_primitive.c restated) + static pair list           */
/* ---------------------------------------------------------------------------------------------------- */
typedef struct { double dist, pos[3], frame[9]; } RawCon;

static int plane_sphere(RawCon* c, double margin, const double* pos1, const double* mat1, const double* pos2, double r) {
  double n[3] = {mat1[2], mat1[5], mat1[8]};
  double tmp[3] = {pos2[0] - pos1[0], pos2[1] - pos1[1], pos2[2] - pos1[2]};
  double cdist = dot3(tmp, n);
  if (cdist > margin + r) return 0;
  c->dist = cdist - r;
  memset(c->frame, 0, sizeof(c->frame));
  memcpy(c->frame, n, sizeof(n));
  for (int k = 0; k < 3; k++) c->pos[k] = pos2[k] + n[k] * (-c->dist * 0.5 - r);
  return 1;
}
static int plane_capsule(RawCon* c, double margin, const double* pos1, const double* mat1, const double* pos2,
                         const double* mat2, const double* size2) {
  double axis[3] = {mat2[2], mat2[5], mat2[8]};
  double seg[3] = {axis[0] * size2[1], axis[1] * size2[1], axis[2] * size2[1]};
  double p[3] = {pos2[0] + seg[0], pos2[1] + seg[1], pos2[2] + seg[2]};
  int n = plane_sphere(c, margin, pos1, mat1, p, size2[0]);
  if (n) memcpy(c->frame + 3, axis, sizeof(axis));
  p[0] = pos2[0] - seg[0]; p[1] = pos2[1] - seg[1]; p[2] = pos2[2] - seg[2];
  int n2 = plane_sphere(c + n, margin, pos1, mat1, p, size2[0]);
  if (n2) memcpy(c[n].frame + 3, axis, sizeof(axis));
  return n + n2;
}
static int plane_cylinder(RawCon* c, double margin, const double* pos1, const double* mat1, const double* pos2,
                          const double* mat2, const double* size2) {
  double normal[3] = {mat1[2], mat1[5], mat1[8]};
  double axis[3] = {mat2[2], mat2[5], mat2[8]};
  double d[3] = {pos2[0] - pos1[0], pos2[1] - pos1[1], pos2[2] - pos1[2]};
  double dist0 = dot3(d, normal);
  double prjaxis = dot3(normal, axis);
  if (prjaxis > 0) { axis[0] = -axis[0]; axis[1] = -axis[1]; axis[2] = -axis[2]; prjaxis = -prjaxis; }
  double vec[3] = {axis[0] * prjaxis - normal[0], axis[1] * prjaxis - normal[1], axis[2] * prjaxis - normal[2]};
  double len_sqr = dot3(vec, vec);
  if (len_sqr >= mjMINVAL) {
    double scl = size2[0] / sqrt(len_sqr);
    vec[0] *= scl; vec[1] *= scl; vec[2] *= scl;
  } else {
    vec[0] = mat2[0] * size2[0]; vec[1] = mat2[3] * size2[0]; vec[2] = mat2[6] * size2[0];
  }
  double prjvec = dot3(vec, normal);
  axis[0] *= size2[1]; axis[1] *= size2[1]; axis[2] *= size2[1];
  prjaxis *= size2[1];
  int cnt = 0;
  if (dist0 + prjaxis + prjvec <= margin) {
    c[cnt].dist = dist0 + prjaxis + prjvec;
    for (int k = 0; k < 3; k++) c[cnt].pos[k] = pos2[k] + vec[k] + axis[k] - normal[k] * c[cnt].dist * 0.5;
    memset(c[cnt].frame, 0, sizeof(c[cnt].frame)); memcpy(c[cnt].frame, normal, sizeof(normal));
    cnt++;
  } else return 0;
  if (dist0 - prjaxis + prjvec <= margin) {
    c[cnt].dist = dist0 - prjaxis + prjvec;
    for (int k = 0; k < 3; k++) c[cnt].pos[k] = pos2[k] + vec[k] - axis[k] - normal[k] * c[cnt].dist * 0.5;
    memset(c[cnt].frame, 0, sizeof(c[cnt].frame)); memcpy(c[cnt].frame, normal, sizeof(normal));
    cnt++;
  }
  double prjvec1 = -prjvec * 0.5;
  if (dist0 + prjaxis + prjvec1 <= margin) {
    double vec1[3];
    cross3(vec1, vec, axis);
    normalize3(vec1);
    double sc = size2[0] * sqrt(3.0) / 2;
    vec1[0] *= sc; vec1[1] *= sc; vec1[2] *= sc;
    for (int sgn = 1; sgn >= -1; sgn -= 2) {
      c[cnt].dist = dist0 + prjaxis + prjvec1;
      for (int k = 0; k < 3; k++)
        c[cnt].pos[k] = pos2[k] + sgn * vec1[k] + axis[k] - vec[k] * 0.5 - normal[k] * c[cnt].dist * 0.5;
      memset(c[cnt].frame, 0, sizeof(c[cnt].frame)); memcpy(c[cnt].frame, normal, sizeof(normal));
      cnt++;
    }
  }
  return cnt;
}
static int plane_box(RawCon* c, double margin, const double* pos1, const double* mat1, const double* pos2,
                     const double* mat2, const double* size2) {
  double norm[3] = {mat1[2], mat1[5], mat1[8]};
  double d[3] = {pos2[0] - pos1[0], pos2[1] - pos1[1], pos2[2] - pos1[2]};
  double dist = dot3(d, norm);
  int cnt = 0;
  for (int i = 0; i < 8; i++) {
    double vec[3] = {(i & 1 ? size2[0] : -size2[0]), (i & 2 ? size2[1] : -size2[1]), (i & 4 ? size2[2] : -size2[2])};
    double corner[3];
    mulmatvec3(corner, mat2, vec);
    double ldist = dot3(norm, corner);
    if (dist + ldist > margin || ldist > 0) continue;
    c[cnt].dist = dist + ldist;
    for (int k = 0; k < 3; k++) c[cnt].pos[k] = pos2[k] + corner[k] - norm[k] * c[cnt].dist * 0.5;
    memset(c[cnt].frame, 0, sizeof(c[cnt].frame)); memcpy(c[cnt].frame, norm, sizeof(norm));
    if (++cnt >= 4) break;
  }
  return cnt;
}
/* mjc_PlaneConvex, mesh branch: the support vertex (deepest) first, then the next-deepest vertices within the margin that are
 * at least 0.3 * rbound away from the FIRST contact (not from each other: two neighbouring sole vertices may both be taken),
 * 3 contacts at most. MuJoCo walks the hull graph of the mesh from the support vertex (engine_collision_convex.c, not under
 * /root/reference); this rule was inferred from the reference goldens UnitreeH1.walk / .carry, whose stance rows are reproduced
 * to 1e-6 by exactly the vertex triple it selects (brute-force search over all triples of the 155 sole vertices near the floor:
 * that triple 1.2e-7, the next best 1.3e-3), and by no rule that also keeps the later contacts apart (round 1: 1e-2). */
static int plane_mesh(RawCon* c, double margin, const double* pos1, const double* mat1, const double* pos2,
                      const double* mat2, const double* verts, int nvert, double rbound) {
  double norm[3] = {mat1[2], mat1[5], mat1[8]};
  double nl[3];
  mulmatTvec3(nl, mat2, norm);
  double d[3] = {pos2[0] - pos1[0], pos2[1] - pos1[1], pos2[2] - pos1[2]};
  double dist0 = dot3(d, norm);
  int cnt = 0;
  int taken[3];
  double mind2 = (0.3 * rbound) * (0.3 * rbound);
  for (int pass = 0; pass < 3; pass++) {
    int best = -1; double bd = 1e300;
    for (int i = 0; i < nvert; i++) {
      const double* v = verts + 3 * i;
      double dd = dist0 + dot3(nl, v);
      if (dd > margin || dd >= bd) continue;
      if (cnt > 0) {
        if (i == taken[0] || (cnt > 1 && i == taken[1])) continue;
        const double* w = verts + 3 * taken[0];
        double e[3] = {v[0] - w[0], v[1] - w[1], v[2] - w[2]};
        if (dot3(e, e) < mind2) continue;
      }
      best = i; bd = dd;
    }
    if (best < 0) break;
    double vg[3];
    mulmatvec3(vg, mat2, verts + 3 * best);
    c[cnt].dist = bd;
    for (int k = 0; k < 3; k++) c[cnt].pos[k] = pos2[k] + vg[k] - norm[k] * bd * 0.5;
    memset(c[cnt].frame, 0, sizeof(c[cnt].frame)); memcpy(c[cnt].frame, norm, sizeof(norm));
    taken[cnt++] = best;
  }
  return cnt;
}
static int sphere_sphere_raw(RawCon* c, double margin, const double* pos1, double r1, const double* pos2, double r2) {
  double dif[3] = {pos2[0] - pos1[0], pos2[1] - pos1[1], pos2[2] - pos1[2]};
  double cdist = norm3(dif);
  if (cdist > margin + r1 + r2) return 0;
  c->dist = cdist - r1 - r2;
  memset(c->frame, 0, sizeof(c->frame));
  if (cdist < mjMINVAL) { c->frame[0] = 1; }
  else { c->frame[0] = dif[0] / cdist; c->frame[1] = dif[1] / cdist; c->frame[2] = dif[2] / cdist; }
  for (int k = 0; k < 3; k++) c->pos[k] = pos1[k] + c->frame[k] * (r1 + 0.5 * c->dist);
  return 1;
}
static int sphere_capsule(RawCon* c, double margin, const double* pos1, double r1, const double* pos2, const double* mat2,
                          const double* size2) {
  double axis[3] = {mat2[2], mat2[5], mat2[8]};
  double vec[3] = {pos1[0] - pos2[0], pos1[1] - pos2[1], pos1[2] - pos2[2]};
  double x = dot3(axis, vec);
  if (x > size2[1]) x = size2[1];
  if (x < -size2[1]) x = -size2[1];
  double p[3] = {pos2[0] + axis[0] * x, pos2[1] + axis[1] * x, pos2[2] + axis[2] * x};
  return sphere_sphere_raw(c, margin, pos1, r1, p, size2[0]);
}
/* mjc_SphereBox (engine_collision_box.c, not under /root/reference): sphere g1 against box g2 -- the sphere centre is
 * clamped to the box in the box frame; outside: normal along (clamped - centre), dist = |centre - clamped| - r; centre
 * inside the box: exit through the nearest face. Geometrically unique results; no reference golden exercises it
 * (parity unpinned). The contact normal points from geom 1 (sphere) to geom 2 (box). */
static int sphere_box(RawCon* c, double margin, const double* pos1, double r, const double* pos2, const double* mat2,
                      const double* size2) {
  double tmp[3] = {pos1[0] - pos2[0], pos1[1] - pos2[1], pos1[2] - pos2[2]}, cen[3], cl[3], d[3], nb[3] = {0, 0, 0}, pl[3];
  mulmatTvec3(cen, mat2, tmp);
  for (int k = 0; k < 3; k++) { cl[k] = fmax(-size2[k], fmin(size2[k], cen[k])); d[k] = cen[k] - cl[k]; }
  double dist = norm3(d);
  if (dist - r > margin) return 0;
  if (dist <= mjMINVAL) {
    double closest = 2 * (size2[0] + size2[1] + size2[2]);
    int kk = 0;
    for (int i = 0; i < 6; i++) {
      double f = fabs((i % 2 ? 1.0 : -1.0) * size2[i / 2] - cen[i / 2]);
      if (closest > f) { closest = f; kk = i; }
    }
    nb[kk / 2] = kk % 2 ? 1.0 : -1.0;                                  /* outward normal of the nearest face */
    for (int k = 0; k < 3; k++) pl[k] = cen[k] + nb[k] * 0.5 * (closest - r);
    c->dist = -closest - r;
  } else {
    for (int k = 0; k < 3; k++) { nb[k] = d[k] / dist; pl[k] = cl[k] + nb[k] * 0.5 * (dist - r); }
    c->dist = dist - r;
  }
  double nw[3], pw[3];
  mulmatvec3(nw, mat2, nb);
  mulmatvec3(pw, mat2, pl);
  memset(c->frame, 0, sizeof(c->frame));
  for (int k = 0; k < 3; k++) { c->frame[k] = -nw[k]; c->pos[k] = pos2[k] + pw[k]; }
  return 1;
}
static int capsule_capsule(RawCon* c, double margin, const double* pos1, const double* mat1, const double* size1,
                           const double* pos2, const double* mat2, const double* size2) {
  double a1[3] = {mat1[2], mat1[5], mat1[8]}, a2[3] = {mat2[2], mat2[5], mat2[8]};
  double dif[3] = {pos1[0] - pos2[0], pos1[1] - pos2[1], pos1[2] - pos2[2]};
  double ma = dot3(a1, a1), mb = -dot3(a1, a2), mc = dot3(a2, a2), u = -dot3(a1, dif), v = dot3(a2, dif);
  double det = ma * mc - mb * mb;
  double len1 = size1[1], len2 = size2[1];
  if (fabs(det) >= mjMINVAL) {
    double x1 = (mc * u - mb * v) / det, x2 = (ma * v - mb * u) / det;
    if (x1 > len1) { x1 = len1; x2 = (v - mb * len1) / mc; }
    else if (x1 < -len1) { x1 = -len1; x2 = (v + mb * len1) / mc; }
    if (x2 > len2) {
      x2 = len2; x1 = (u - mb * len2) / ma;
      if (x1 > len1) x1 = len1; else if (x1 < -len1) x1 = -len1;
    } else if (x2 < -len2) {
      x2 = -len2; x1 = (u + mb * len2) / ma;
      if (x1 > len1) x1 = len1; else if (x1 < -len1) x1 = -len1;
    }
    double p1[3] = {pos1[0] + a1[0] * x1, pos1[1] + a1[1] * x1, pos1[2] + a1[2] * x1};
    double p2[3] = {pos2[0] + a2[0] * x2, pos2[1] + a2[1] * x2, pos2[2] + a2[2] * x2};
    return sphere_sphere_raw(c, margin, p1, size1[0], p2, size2[0]);
  }
  /* parallel axes: test both ends of capsule 1 against capsule 2 and vice versa, keep distinct contacts (<=2) */
  int n = 0;
  for (int e = -1; e <= 1 && n < 2; e += 2) {
    double p1[3] = {pos1[0] + a1[0] * len1 * e, pos1[1] + a1[1] * len1 * e, pos1[2] + a1[2] * len1 * e};
    double vec[3] = {p1[0] - pos2[0], p1[1] - pos2[1], p1[2] - pos2[2]};
    double x2 = dot3(a2, vec);
    if (x2 > len2 || x2 < -len2) continue;
    double p2[3] = {pos2[0] + a2[0] * x2, pos2[1] + a2[1] * x2, pos2[2] + a2[2] * x2};
    n += sphere_sphere_raw(c + n, margin, p1, size1[0], p2, size2[0]);
  }
  for (int e = -1; e <= 1 && n < 2; e += 2) {
    double p2[3] = {pos2[0] + a2[0] * len2 * e, pos2[1] + a2[1] * len2 * e, pos2[2] + a2[2] * len2 * e};
    double vec[3] = {p2[0] - pos1[0], p2[1] - pos1[1], p2[2] - pos1[2]};
    double x1 = dot3(a1, vec);
    if (x1 >= len1 || x1 <= -len1) continue;
    double p1[3] = {pos1[0] + a1[0] * x1, pos1[1] + a1[1] * x1, pos1[2] + a1[2] * x1};
    n += sphere_sphere_raw(c + n, margin, p1, size1[0], p2, size2[0]);
  }
  return n;
}


/* ---------------------------------------------------------------------------------------------------- */
/* mjc_Convex: general convex pair = libccd's Minkowski Portal Refinement (ccdMPRPenetration, libccd 2.1 as   */
/* vendored by MuJoCo 2.3.7; engine_collision_convex.c: mjccd_center / mjccd_support / mjc_MPRIteration).     */
/* Neither source is under /root/reference (third party): restated from the published algorithm (G. Snethen,  */
/* "XenoCollide", Game Programming Gems 7; libccd src/mpr.c, src/vec3.c), pinned by the tails of the           */
/* reference goldens that contain bone-bone contacts (tests/test_oracle_golden.py).                            */
/* Each object is inflated by margin/2 in the support direction; dist = margin - depth.                        */
/* ---------------------------------------------------------------------------------------------------- */
#define CCD_EPS 2.220446049250313e-16
typedef struct { double v[3], v1[3], v2[3]; } CcdSup;
typedef struct { const RefSim* s; int g; double margin; } CcdObj;
static inline int ccd_is_zero(double x) { return fabs(x) < CCD_EPS; }
static inline int ccd_eq(double a, double b) {
  double ab = fabs(a - b);
  if (ab < CCD_EPS) return 1;
  a = fabs(a); b = fabs(b);
  return b > a ? ab < CCD_EPS * b : ab < CCD_EPS * a;
}
static inline int ccd_vec_eq(const double* a, const double* b) { return ccd_eq(a[0], b[0]) && ccd_eq(a[1], b[1]) && ccd_eq(a[2], b[2]); }
static inline void ccd_normalize(double* d) { double inv = 1.0 / sqrt(dot3(d, d)); d[0] *= inv; d[1] *= inv; d[2] *= inv; }

/* mjccd_support: farthest point of geom g (inflated by margin) in world direction dir (unit) */
static void ccd_support_geom(const CcdObj* o, const double* dir, double* res) {
  const Model* m = &o->s->m;
  const int g = o->g;
  const double* mat = o->s->gxmat[g];
  const double* size = m->geom_size + 3 * g;
  double ld[3], r[3] = {0, 0, 0};
  mulmatTvec3(ld, mat, dir);
  switch (m->geom_type[g]) {
    case LS_GEOM_SPHERE: for (int k = 0; k < 3; k++) r[k] = ld[k] * size[0]; break;
    case LS_GEOM_CAPSULE:
      for (int k = 0; k < 3; k++) r[k] = ld[k] * size[0];
      r[2] += (ld[2] >= 0 ? 1.0 : -1.0) * size[1];
      break;
    case LS_GEOM_CYLINDER: {
      double t = sqrt(ld[0] * ld[0] + ld[1] * ld[1]);
      if (t > mjMINVAL) { r[0] = ld[0] / t * size[0]; r[1] = ld[1] / t * size[0]; }
      r[2] = (ld[2] >= 0 ? 1.0 : -1.0) * size[1];
      break;
    }
    case LS_GEOM_BOX: for (int k = 0; k < 3; k++) r[k] = (ld[k] >= 0 ? 1.0 : -1.0) * size[k]; break;
    case LS_GEOM_MESH: {
      const double* v = m->mesh_vert + 3 * m->geom_meshadr[g];
      int n = m->geom_meshnum[g], best = 0;
      double mx = -1e300;
      for (int i = 0; i < n; i++) { double d = dot3(ld, v + 3 * i); if (d > mx) { mx = d; best = i; } }
      memcpy(r, v + 3 * best, sizeof(r));
      break;
    }
    default: break;
  }
  mulmatvec3(res, mat, r);
  for (int k = 0; k < 3; k++) res[k] += o->s->gxpos[g][k] + dir[k] * o->margin;
}
static long g_convex_support;
static int g_ccd_trace = 0;
void ref_debug_trace(int on) { g_ccd_trace = on; }
static void ccd_support(const CcdObj* o1, const CcdObj* o2, const double* dir, CcdSup* sp) {
  g_convex_support++;
  double nd[3] = {-dir[0], -dir[1], -dir[2]};
  ccd_support_geom(o1, dir, sp->v1);
  ccd_support_geom(o2, nd, sp->v2);
  for (int k = 0; k < 3; k++) sp->v[k] = sp->v1[k] - sp->v2[k];
  if (g_ccd_trace) printf("    [f64] g%d-g%d dir %.6f %.6f %.6f -> v %.7f %.7f %.7f\n", o1->g, o2->g, dir[0], dir[1], dir[2], sp->v[0], sp->v[1], sp->v[2]);
}
static void ccd_portal_dir(const CcdSup* p, double* dir) {
  double a[3], b[3];
  for (int k = 0; k < 3; k++) { a[k] = p[2].v[k] - p[1].v[k]; b[k] = p[3].v[k] - p[1].v[k]; }
  cross3(dir, a, b);
  ccd_normalize(dir);
}
static int ccd_reach_tolerance(const CcdSup* p, const CcdSup* v4, const double* dir, double tol) {
  double dv1 = dot3(p[1].v, dir), dv2 = dot3(p[2].v, dir), dv3 = dot3(p[3].v, dir), dv4 = dot3(v4->v, dir);
  double d1 = dv4 - dv1, d2 = dv4 - dv2, d3 = dv4 - dv3;
  d1 = d1 < d2 ? d1 : d2;
  d1 = d1 < d3 ? d1 : d3;
  return ccd_eq(d1, tol) || d1 < tol;
}
static void ccd_expand_portal(CcdSup* p, const CcdSup* v4) {
  double v4v0[3];
  cross3(v4v0, v4->v, p[0].v);
  if (dot3(p[1].v, v4v0) > 0) {
    if (dot3(p[2].v, v4v0) > 0) p[1] = *v4; else p[3] = *v4;
  } else {
    if (dot3(p[3].v, v4v0) > 0) p[2] = *v4; else p[1] = *v4;
  }
}
static double ccd_point_seg_dist2(const double* P, const double* x0, const double* b, double* wit) {
  double d[3], a[3];
  for (int k = 0; k < 3; k++) { d[k] = b[k] - x0[k]; a[k] = x0[k] - P[k]; }
  double t = -dot3(a, d) / dot3(d, d);
  if (t < 0 || ccd_is_zero(t)) memcpy(wit, x0, 3 * sizeof(double));
  else if (t > 1 || ccd_eq(t, 1.0)) memcpy(wit, b, 3 * sizeof(double));
  else for (int k = 0; k < 3; k++) wit[k] = d[k] * t + x0[k];
  double e[3] = {wit[0] - P[0], wit[1] - P[1], wit[2] - P[2]};
  return dot3(e, e);
}
static double ccd_point_tri_dist2(const double* P, const double* x0, const double* B, const double* C, double* wit) {
  double d1[3], d2[3], a[3];
  for (int k = 0; k < 3; k++) { d1[k] = B[k] - x0[k]; d2[k] = C[k] - x0[k]; a[k] = x0[k] - P[k]; }
  double v = dot3(d1, d1), w = dot3(d2, d2), p = dot3(a, d1), q = dot3(a, d2), r = dot3(d1, d2);
  double d = w * v - r * r, s, t, dist;
  if (ccd_is_zero(d)) s = t = -1.0;
  else { s = (q * r - w * p) / d; t = (-s * r - q) / w; }
  if ((ccd_is_zero(s) || s > 0) && (ccd_eq(s, 1.0) || s < 1) && (ccd_is_zero(t) || t > 0) && (ccd_eq(t, 1.0) || t < 1) &&
      (ccd_eq(t + s, 1.0) || t + s < 1)) {
    for (int k = 0; k < 3; k++) wit[k] = x0[k] + d1[k] * s + d2[k] * t;
    double e[3] = {wit[0] - P[0], wit[1] - P[1], wit[2] - P[2]};
    dist = dot3(e, e);
  } else {
    double w2[3], dist2;
    dist = ccd_point_seg_dist2(P, x0, B, wit);
    dist2 = ccd_point_seg_dist2(P, x0, C, w2);
    if (dist2 < dist) { dist = dist2; memcpy(wit, w2, sizeof(w2)); }
    dist2 = ccd_point_seg_dist2(P, B, C, w2);
    if (dist2 < dist) { dist = dist2; memcpy(wit, w2, sizeof(w2)); }
  }
  return dist;
}
static void ccd_find_pos(const CcdSup* p, double* pos) {
  double dir[3], vec[3], b[4], sum;
  ccd_portal_dir(p, dir);
  cross3(vec, p[1].v, p[2].v); b[0] = dot3(vec, p[3].v);
  cross3(vec, p[3].v, p[2].v); b[1] = dot3(vec, p[0].v);
  cross3(vec, p[0].v, p[1].v); b[2] = dot3(vec, p[3].v);
  cross3(vec, p[2].v, p[1].v); b[3] = dot3(vec, p[0].v);
  sum = b[0] + b[1] + b[2] + b[3];
  if (ccd_is_zero(sum) || sum < 0) {
    b[0] = 0;
    cross3(vec, p[2].v, p[3].v); b[1] = dot3(vec, dir);
    cross3(vec, p[3].v, p[1].v); b[2] = dot3(vec, dir);
    cross3(vec, p[1].v, p[2].v); b[3] = dot3(vec, dir);
    sum = b[1] + b[2] + b[3];
  }
  double inv = 1.0 / sum, p1[3] = {0, 0, 0}, p2[3] = {0, 0, 0};
  for (int i = 0; i < 4; i++) for (int k = 0; k < 3; k++) { p1[k] += p[i].v1[k] * b[i]; p2[k] += p[i].v2[k] * b[i]; }
  for (int k = 0; k < 3; k++) pos[k] = (p1[k] * inv + p2[k] * inv) * 0.5;
}
/* returns 0 and (depth, dir, pos) if the inflated objects intersect, -1 otherwise */
static int ccd_mpr_penetration(const CcdObj* o1, const CcdObj* o2, double tol, int max_iter, double* depth, double* pdir,
                               double* pos) {
  static const double origin[3] = {0, 0, 0};
  CcdSup p[4], v4;
  double dir[3], va[3], vb[3], dot;
  /* ---- discoverPortal ---- */
  for (int k = 0; k < 3; k++) {
    p[0].v1[k] = o1->s->gxpos[o1->g][k]; p[0].v2[k] = o2->s->gxpos[o2->g][k]; p[0].v[k] = p[0].v1[k] - p[0].v2[k];
  }
  if (ccd_vec_eq(p[0].v, origin)) p[0].v[0] += CCD_EPS * 10.0;
  for (int k = 0; k < 3; k++) dir[k] = -p[0].v[k];
  ccd_normalize(dir);
  ccd_support(o1, o2, dir, &p[1]);
  dot = dot3(p[1].v, dir);
  if (ccd_is_zero(dot) || dot < 0) return -1;
  cross3(dir, p[0].v, p[1].v);
  if (ccd_is_zero(dot3(dir, dir))) {
    if (ccd_vec_eq(p[1].v, origin)) {           /* findPenetrTouch */
      *depth = 0; pdir[0] = pdir[1] = pdir[2] = 0;
      for (int k = 0; k < 3; k++) pos[k] = 0.5 * (p[1].v1[k] + p[1].v2[k]);
    } else {                                    /* findPenetrSegment */
      for (int k = 0; k < 3; k++) { pos[k] = 0.5 * (p[1].v1[k] + p[1].v2[k]); pdir[k] = p[1].v[k]; }
      *depth = sqrt(dot3(pdir, pdir));
      ccd_normalize(pdir);
    }
    return 0;
  }
  ccd_normalize(dir);
  ccd_support(o1, o2, dir, &p[2]);
  dot = dot3(p[2].v, dir);
  if (ccd_is_zero(dot) || dot < 0) return -1;
  for (int k = 0; k < 3; k++) { va[k] = p[1].v[k] - p[0].v[k]; vb[k] = p[2].v[k] - p[0].v[k]; }
  cross3(dir, va, vb);
  ccd_normalize(dir);
  if (dot3(dir, p[0].v) > 0) {
    CcdSup t = p[1]; p[1] = p[2]; p[2] = t;
    for (int k = 0; k < 3; k++) dir[k] = -dir[k];
  }
  for (;;) {
    ccd_support(o1, o2, dir, &p[3]);
    dot = dot3(p[3].v, dir);
    if (ccd_is_zero(dot) || dot < 0) return -1;
    int cont = 0;
    cross3(va, p[1].v, p[3].v);
    dot = dot3(va, p[0].v);
    if (dot < 0 && !ccd_is_zero(dot)) { p[2] = p[3]; cont = 1; }
    if (!cont) {
      cross3(va, p[3].v, p[2].v);
      dot = dot3(va, p[0].v);
      if (dot < 0 && !ccd_is_zero(dot)) { p[1] = p[3]; cont = 1; }
    }
    if (!cont) break;
    for (int k = 0; k < 3; k++) { va[k] = p[1].v[k] - p[0].v[k]; vb[k] = p[2].v[k] - p[0].v[k]; }
    cross3(dir, va, vb);
    ccd_normalize(dir);
  }
  /* ---- refinePortal ---- */
  for (;;) {
    ccd_portal_dir(p, dir);
    dot = dot3(dir, p[1].v);
    if (ccd_is_zero(dot) || dot > 0) break;                       /* portal encapsules the origin */
    ccd_support(o1, o2, dir, &v4);
    dot = dot3(v4.v, dir);
    if (!(ccd_is_zero(dot) || dot > 0) || ccd_reach_tolerance(p, &v4, dir, tol)) return -1;
    ccd_expand_portal(p, &v4);
  }
  /* ---- findPenetr ---- */
  for (int iterations = 0;; iterations++) {
    ccd_portal_dir(p, dir);
    ccd_support(o1, o2, dir, &v4);
    if (ccd_reach_tolerance(p, &v4, dir, tol) || iterations > max_iter) {
      *depth = sqrt(ccd_point_tri_dist2(origin, p[1].v, p[2].v, p[3].v, pdir));
      if (ccd_is_zero(*depth)) pdir[0] = pdir[1] = pdir[2] = 0;
      else ccd_normalize(pdir);
      ccd_find_pos(p, pos);
      return 0;
    }
    ccd_expand_portal(p, &v4);
  }
}
static long g_convex_calls = 0, g_convex_hits = 0;
long ref_debug_convex(int k) { return k == 0 ? g_convex_calls : (k == 1 ? g_convex_hits : g_convex_support); }
static int convex_pair(RawCon* c, const RefSim* s, int g1, int g2, double margin) {
  g_convex_calls++;
  CcdObj o1 = {s, g1, 0.5 * margin}, o2 = {s, g2, 0.5 * margin};
  static const double origin[3] = {0, 0, 0};
  double depth, dir[3], pos[3];
  if (ccd_mpr_penetration(&o1, &o2, 1e-6, 50, &depth, dir, pos) != 0) return 0;      /* opt.mpr_tolerance / mpr_iterations */
  if (ccd_vec_eq(dir, origin)) return 0;
  g_convex_hits++;
  c->dist = margin - depth;
  memset(c->frame, 0, sizeof(c->frame));
  memcpy(c->frame, dir, sizeof(dir));
  memcpy(c->pos, pos, sizeof(pos));
  return 1;
}

/* mjc_BoxBox, edge-edge branch. MuJoCo's dedicated box-box routine (engine_collision_box.c, not under /root/reference) is a
 * separating-axis test over the 15 axes of two boxes; when the axis of least overlap (largest separation) is the cross
 * product of an edge of each box, the contact is ONE point: the midpoint of the closest points of the two edges, normal along
 * the axis, dist = separation. That case is geometrically unique and restated here; it is pinned by the reference golden
 * HumanoidTorque4Ages.run.all (row 9: the infant's foot boxes pass each other edge to edge, 0.17 mm apart inside the 1 mm
 * margin; MPR gets dist and normal right to 1e-7 but its contact POSITION is off by up to 1.6 mm there).
 * Face / vertex cases (several contact points in MuJoCo) return -1: the caller falls back to the MPR stand-in. */
static int box_box_edge(RawCon* c, double margin, const double* p1, const double* m1, const double* s1, const double* p2,
                        const double* m2, const double* s2) {
  double A[3][3], B[3][3], d[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
  for (int i = 0; i < 3; i++) for (int k = 0; k < 3; k++) { A[i][k] = m1[3 * k + i]; B[i][k] = m2[3 * k + i]; }   /* box axes = columns */
  double best = -1e300; int bi = -1, bj = -1; double bn[3] = {0, 0, 0};
  /* face axes */
  for (int f = 0; f < 6; f++) {
    const double* ax = f < 3 ? A[f] : B[f - 3];
    double ra = 0, rb = 0;
    for (int i = 0; i < 3; i++) { ra += s1[i] * fabs(dot3(A[i], ax)); rb += s2[i] * fabs(dot3(B[i], ax)); }
    double sep = fabs(dot3(d, ax)) - ra - rb;
    if (sep > best) { best = sep; bi = -1; bj = f; }
  }
  /* edge x edge axes */
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
    double ax[3];
    cross3(ax, A[i], B[j]);
    double l = norm3(ax);
    if (l < 1e-6) continue;                                   /* parallel edges: covered by the face axes */
    for (int k = 0; k < 3; k++) ax[k] /= l;
    double ra = 0, rb = 0;
    for (int k = 0; k < 3; k++) { ra += s1[k] * fabs(dot3(A[k], ax)); rb += s2[k] * fabs(dot3(B[k], ax)); }
    double sd = dot3(d, ax), sep = fabs(sd) - ra - rb;
    if (sep > best + 1e-12) { best = sep; bi = i; bj = j; for (int k = 0; k < 3; k++) bn[k] = sd >= 0 ? ax[k] : -ax[k]; }
  }
  if (bi < 0) return -1;                                       /* a face axis separates best: not the edge-edge case */
  if (best > margin) return 0;
  /* the two edges: on box 1 the edge along A[bi] that is extreme along +n in the other two axes, on box 2 along B[bj], extreme along -n */
  double e1[3], e2[3];
  for (int k = 0; k < 3; k++) { e1[k] = p1[k]; e2[k] = p2[k]; }
  for (int i = 0; i < 3; i++) if (i != bi) { double sg = dot3(A[i], bn) >= 0 ? 1.0 : -1.0; for (int k = 0; k < 3; k++) e1[k] += sg * s1[i] * A[i][k]; }
  for (int j = 0; j < 3; j++) if (j != bj) { double sg = dot3(B[j], bn) >= 0 ? -1.0 : 1.0; for (int k = 0; k < 3; k++) e2[k] += sg * s2[j] * B[j][k]; }
  /* closest points of the lines e1 + t A[bi], e2 + u B[bj] */
  double r[3] = {e1[0] - e2[0], e1[1] - e2[1], e1[2] - e2[2]};
  double b = dot3(A[bi], B[bj]), cc = dot3(A[bi], r), f = dot3(B[bj], r), den = 1.0 - b * b;
  double t = (b * f - cc) / den, u = (f - b * cc) / den;
  if (fabs(t) > s1[bi] || fabs(u) > s2[bj]) return -1;        /* closest points off the edges: a vertex is involved */
  double c1[3], c2[3];
  for (int k = 0; k < 3; k++) { c1[k] = e1[k] + t * A[bi][k]; c2[k] = e2[k] + u * B[bj][k]; }
  c->dist = best;
  memset(c->frame, 0, sizeof(c->frame));
  for (int k = 0; k < 3; k++) { c->frame[k] = bn[k]; c->pos[k] = 0.5 * (c1[k] + c2[k]); }
  return 1;
}

/* test entry points of the pair routines (tests/test_collision_pairs.py): out = dist, pos[3], normal[3] */
int ref_debug_box_box_edge(double margin, const double* p1, const double* m1, const double* s1, const double* p2, const double* m2,
                           const double* s2, double* out) {
  RawCon c;
  int n = box_box_edge(&c, margin, p1, m1, s1, p2, m2, s2);
  if (n > 0) { out[0] = c.dist; for (int k = 0; k < 3; k++) { out[1 + k] = c.pos[k]; out[4 + k] = c.frame[k]; } }
  return n;
}
int ref_debug_sphere_box(double margin, const double* p1, double r, const double* p2, const double* m2, const double* s2, double* out) {
  RawCon c;
  int n = sphere_box(&c, margin, p1, r, p2, m2, s2);
  if (n > 0) { out[0] = c.dist; for (int k = 0; k < 3; k++) { out[1 + k] = c.pos[k]; out[4 + k] = c.frame[k]; } }
  return n;
}

static void make_frame(double* f) {
  /* mju_makeFrame: normal given in f[0:3]; tangent f[3:6] optional */
  normalize3(f);
  if (norm3(f + 3) < 0.5) {
    f[3] = f[4] = f[5] = 0;
    if (f[1] < 0.5 && f[1] > -0.5) f[4] = 1; else f[5] = 1;
  }
  double t = dot3(f, f + 3);
  for (int k = 0; k < 3; k++) f[3 + k] -= t * f[k];
  normalize3(f + 3);
  cross3(f + 6, f, f + 3);
}

static void contact_params(const Model* m, int g1, int g2, Contact* con) {
  /* mj_contactParam */
  int p1 = m->geom_priority[g1], p2 = m->geom_priority[g2];
  double gap = fmax(m->geom_gap[g1], m->geom_gap[g2]);
  double margin = fmax(m->geom_margin[g1], m->geom_margin[g2]);
  con->includemargin = margin - gap;
  double fri[3];
  if (p1 != p2) {
    int g = p1 > p2 ? g1 : g2;
    con->dim = m->geom_condim[g];
    memcpy(con->solref, m->geom_solref + 2 * g, sizeof(double) * 2);
    memcpy(con->solimp, m->geom_solimp + 5 * g, sizeof(double) * 5);
    memcpy(fri, m->geom_friction + 3 * g, sizeof(double) * 3);
  } else {
    con->dim = m->geom_condim[g1] > m->geom_condim[g2] ? m->geom_condim[g1] : m->geom_condim[g2];
    double s1 = m->geom_solmix[g1], s2 = m->geom_solmix[g2], mix;
    if (s1 >= mjMINVAL && s2 >= mjMINVAL) mix = s1 / (s1 + s2);
    else if (s1 < mjMINVAL && s2 < mjMINVAL) mix = 0.5;
    else if (s1 < mjMINVAL) mix = 0.0;
    else mix = 1.0;
    const double *r1 = m->geom_solref + 2 * g1, *r2 = m->geom_solref + 2 * g2;
    if (r1[0] > 0 && r2[0] > 0) for (int k = 0; k < 2; k++) con->solref[k] = mix * r1[k] + (1 - mix) * r2[k];
    else for (int k = 0; k < 2; k++) con->solref[k] = fmin(r1[k], r2[k]);
    for (int k = 0; k < 5; k++) con->solimp[k] = mix * m->geom_solimp[5 * g1 + k] + (1 - mix) * m->geom_solimp[5 * g2 + k];
    for (int k = 0; k < 3; k++) fri[k] = fmax(m->geom_friction[3 * g1 + k], m->geom_friction[3 * g2 + k]);
  }
  con->friction[0] = fri[0]; con->friction[1] = fri[0]; con->friction[2] = fri[1]; con->friction[3] = fri[2];
  con->friction[4] = fri[2];
}

static void collision(RefSim* s) {
  const Model* m = &s->m;
  s->ncon = 0;
  RawCon raw[MAXCONPAIR];
  for (int p = 0; p < m->np; p++) {
    int g1 = m->pair_geom[2 * p], g2 = m->pair_geom[2 * p + 1];
    int t1 = m->geom_type[g1], t2 = m->geom_type[g2];
    double margin = fmax(m->geom_margin[g1], m->geom_margin[g2]);
    const double *pos1 = s->gxpos[g1], *mat1 = s->gxmat[g1], *pos2 = s->gxpos[g2], *mat2 = s->gxmat[g2];
    const double *size1 = m->geom_size + 3 * g1, *size2 = m->geom_size + 3 * g2;
    /* bounding-sphere / plane filter (mj_collideGeoms). NOTE: the reference's golden rollouts
     * (tests/golden/UnitreeA1.simple.real.npy rows 1,2,4,9) show that MuJoCo 2.3.7 does NOT activate a
     * sphere-plane contact while 0 < dist < margin, although capsule-plane contacts in the same margin zone are
     * active: the mid-phase plane filter cuts at centre distance > rbound, without the margin. */
    if (t1 == LS_GEOM_PLANE) {
      double n[3] = {mat1[2], mat1[5], mat1[8]};
      double d[3] = {pos2[0] - pos1[0], pos2[1] - pos1[1], pos2[2] - pos1[2]};
      if (dot3(d, n) > m->geom_rbound[g2]) continue;
    } else {
      double d[3] = {pos2[0] - pos1[0], pos2[1] - pos1[1], pos2[2] - pos1[2]};
      double bound = m->geom_rbound[g1] + m->geom_rbound[g2];
      if (dot3(d, d) > bound * bound) continue;
    }
    int n = 0;
    if (t1 == LS_GEOM_PLANE) {
      if (t2 == LS_GEOM_SPHERE) n = plane_sphere(raw, margin, pos1, mat1, pos2, size2[0]);
      else if (t2 == LS_GEOM_CAPSULE) n = plane_capsule(raw, margin, pos1, mat1, pos2, mat2, size2);
      else if (t2 == LS_GEOM_CYLINDER) n = plane_cylinder(raw, margin, pos1, mat1, pos2, mat2, size2);
      else if (t2 == LS_GEOM_BOX) n = plane_box(raw, margin, pos1, mat1, pos2, mat2, size2);
      else if (t2 == LS_GEOM_MESH)
        n = plane_mesh(raw, margin, pos1, mat1, pos2, mat2, m->mesh_vert + 3 * m->geom_meshadr[g2], m->geom_meshnum[g2],
                       m->geom_rbound[g2]);
    } else if (t1 == LS_GEOM_SPHERE && t2 == LS_GEOM_SPHERE) {
      n = sphere_sphere_raw(raw, margin, pos1, size1[0], pos2, size2[0]);
    } else if (t1 == LS_GEOM_SPHERE && t2 == LS_GEOM_CAPSULE) {
      n = sphere_capsule(raw, margin, pos1, size1[0], pos2, mat2, size2);
    } else if (t1 == LS_GEOM_CAPSULE && t2 == LS_GEOM_CAPSULE) {
      n = capsule_capsule(raw, margin, pos1, mat1, size1, pos2, mat2, size2);
    } else if (t1 == LS_GEOM_SPHERE && t2 == LS_GEOM_BOX) {
      n = sphere_box(raw, margin, pos1, size1[0], pos2, mat2, size2);
    } else {
      /* every other pair: mjc_Convex (MPR), one contact -- MuJoCo 2.3.7's collision table routes sphere | capsule |
         cylinder | box | mesh against cylinder | mesh there (and ellipsoids, which no in-scope model has).
         DEVIATION: capsule-box and box-box have dedicated multi-contact routines in MuJoCo (mjc_CapsuleBox, mjc_BoxBox);
         only the edge-edge branch of mjc_BoxBox is restated (box_box_edge, golden-pinned); everything else of these two
         goes through MPR here: ONE contact at the deepest point instead of up to 2 / 8
         (mjcf.py counts them in Model.n_approx_pairs; A1 trunk vs legs, the humanoid's two foot boxes). */
      n = -1;
      if (t1 == LS_GEOM_BOX && t2 == LS_GEOM_BOX) n = box_box_edge(raw, margin, pos1, mat1, size1, pos2, mat2, size2);
      if (n < 0) n = convex_pair(raw, s, g1, g2, margin);
    }
    for (int k = 0; k < n && s->ncon < MAXCON; k++) {
      Contact* c = &s->con[s->ncon++];
      c->dist = raw[k].dist;
      memcpy(c->pos, raw[k].pos, sizeof(c->pos));
      memcpy(c->frame, raw[k].frame, sizeof(c->frame));
      make_frame(c->frame);
      c->geom1 = g1; c->geom2 = g2;
      contact_params(m, g1, g2, c);
      c->exclude = (c->dist >= c->includemargin);
      c->efc_address = -1;
      c->mu = 0;
    }
  }
}

/* ---------------------------------------------------------------------------------------------------- */
/* mj_makeConstraint: frictionloss, joint limits, contacts; then impedance/reference parameters          */
/* ---------------------------------------------------------------------------------------------------- */
static void jac_point(const RefSim* s, double* jacp, double* jacr, const double* point, int body) {
  /* mj_jac: 3 x nv each, row-major */
  const Model* m = &s->m;
  int nv = m->nv;
  for (int k = 0; k < 3 * nv; k++) { jacp[k] = 0; if (jacr) jacr[k] = 0; }
  if (body == 0) return;
  const double* com = s->subtree_com[m->body_rootid[body]];
  double off[3] = {point[0] - com[0], point[1] - com[1], point[2] - com[2]};
  for (int d = m->body_lastdof[body]; d >= 0; d = m->dof_parentid[d]) {
    const double* cd = s->cdof[d];
    double t[3];
    cross3(t, cd, off);
    for (int k = 0; k < 3; k++) {
      jacp[k * nv + d] = cd[3 + k] + t[k];
      if (jacr) jacr[k * nv + d] = cd[k];
    }
  }
}

static void get_impedance(const double* solimp, double pos, double margin, double* imp, double* impP) {
  if (solimp[0] == solimp[1] || solimp[2] <= mjMINVAL) { *imp = 0.5 * (solimp[0] + solimp[1]); *impP = 0; return; }
  double x = (pos - margin) / solimp[2];
  double sgn = 1;
  if (x < 0) { x = -x; sgn = -1; }
  if (x >= 1 || x <= 0) { *imp = (x >= 1 ? solimp[1] : solimp[0]); *impP = 0; return; }
  double y, yP;
  if (solimp[4] == 1) { y = x; yP = 1; }
  else if (x <= solimp[3]) {
    double a = 1 / pow(solimp[3], solimp[4] - 1);
    y = a * pow(x, solimp[4]);
    yP = solimp[4] * a * pow(x, solimp[4] - 1);
  } else {
    double b = 1 / pow(1 - solimp[3], solimp[4] - 1);
    y = 1 - b * pow(1 - x, solimp[4]);
    yP = solimp[4] * b * pow(1 - x, solimp[4] - 1);
  }
  *imp = solimp[0] + y * (solimp[1] - solimp[0]);
  *impP = yP * sgn * (solimp[1] - solimp[0]) / solimp[2];
}

static void make_constraint(RefSim* s) {
  const Model* m = &s->m;
  int nv = m->nv, n = 0;
  /* --- dof frictionloss (mj_instantiateFriction) --- */
  for (int i = 0; i < nv; i++) {
    if (m->dof_frictionloss[i] > 0) {
      memset(s->J[n], 0, sizeof(double) * nv);
      s->J[n][i] = 1;
      s->efc_pos[n] = 0; s->efc_margin[n] = 0; s->efc_frictionloss[n] = m->dof_frictionloss[i];
      s->efc_type[n] = CNSTR_FRICTION_DOF; s->efc_id[n] = i;
      s->efc_diagApprox[n] = m->dof_invweight0[i];
      n++;
    }
  }
  s->nf = n;
  /* --- joint limits (mj_instantiateLimit) --- */
  for (int i = 0; i < nv; i++) {
    if (!m->jnt_limited[i]) continue;
    double value = s->qpos[i], margin = m->jnt_margin[i];
    for (int side = -1; side <= 1; side += 2) {
      double dist = side * (m->jnt_range[2 * i + (side + 1) / 2] - value);
      if (dist < margin) {
        memset(s->J[n], 0, sizeof(double) * nv);
        s->J[n][i] = -side;
        s->efc_pos[n] = dist; s->efc_margin[n] = margin; s->efc_frictionloss[n] = 0;
        s->efc_type[n] = CNSTR_LIMIT_JOINT; s->efc_id[n] = i;
        s->efc_diagApprox[n] = m->dof_invweight0[i];
        n++;
      }
    }
  }
  s->nl = n - s->nf;
  /* --- contacts (mj_instantiateContact) --- */
  static __thread double jp1[3 * MAXNV], jr1[3 * MAXNV], jp2[3 * MAXNV], jr2[3 * MAXNV];
  for (int ci = 0; ci < s->ncon; ci++) {
    Contact* c = &s->con[ci];
    if (c->exclude) continue;
    int b1 = m->geom_bodyid[c->geom1], b2 = m->geom_bodyid[c->geom2];
    int dim = c->dim;
    int nrow = (dim == 1) ? 1 : (m->cone == 1 ? dim : 2 * (dim - 1));
    if (n + nrow > MAXEFC) break;
    jac_point(s, jp1, jr1, c->pos, b1);
    jac_point(s, jp2, jr2, c->pos, b2);
    /* jacdif in contact frame: rows 0..2 translational (normal, t1, t2), rows 3..5 rotational */
    double jd[6][MAXNV];
    for (int r = 0; r < 3; r++)
      for (int d = 0; d < nv; d++) {
        double vp = 0, vr = 0;
        for (int k = 0; k < 3; k++) {
          vp += c->frame[3 * r + k] * (jp2[k * nv + d] - jp1[k * nv + d]);
          vr += c->frame[3 * r + k] * (jr2[k * nv + d] - jr1[k * nv + d]);
        }
        jd[r][d] = vp; jd[3 + r][d] = vr;
      }
    double tran = m->geom_invweight0[2 * c->geom1] + m->geom_invweight0[2 * c->geom2];
    double rot = m->geom_invweight0[2 * c->geom1 + 1] + m->geom_invweight0[2 * c->geom2 + 1];
    c->efc_address = n;
    if (dim == 1) {
      memcpy(s->J[n], jd[0], sizeof(double) * nv);
      s->efc_pos[n] = c->dist; s->efc_margin[n] = c->includemargin; s->efc_frictionloss[n] = 0;
      s->efc_type[n] = CNSTR_CONTACT_FRICTIONLESS; s->efc_id[n] = ci; s->efc_diagApprox[n] = tran;
      n++;
    } else if (m->cone == 1) {
      for (int r = 0; r < dim; r++) {
        memcpy(s->J[n], jd[r], sizeof(double) * nv);
        s->efc_pos[n] = c->dist; s->efc_margin[n] = c->includemargin; s->efc_frictionloss[n] = 0;
        s->efc_type[n] = CNSTR_CONTACT_ELLIPTIC; s->efc_id[n] = ci;
        s->efc_diagApprox[n] = r < 3 ? tran : rot;
        n++;
      }
    } else {
      for (int r = 1; r < dim; r++) {
        double fr = c->friction[r - 1];
        for (int sgn = 1; sgn >= -1; sgn -= 2) {
          for (int d = 0; d < nv; d++) s->J[n][d] = jd[0][d] + sgn * fr * jd[r][d];
          s->efc_pos[n] = c->dist; s->efc_margin[n] = c->includemargin; s->efc_frictionloss[n] = 0;
          s->efc_type[n] = CNSTR_CONTACT_PYRAMIDAL; s->efc_id[n] = ci;
          s->efc_diagApprox[n] = tran + c->friction[0] * c->friction[0] * tran;
          n++;
        }
      }
    }
  }
  s->nefc = n;

  /* --- mj_makeImpedance --- */
  for (int i = 0; i < n; i++) {
    double solref[2], solimp[5];
    int tp = s->efc_type[i], id = s->efc_id[i];
    if (tp == CNSTR_FRICTION_DOF) { memcpy(solref, m->dof_solref + 2 * id, 16); memcpy(solimp, m->dof_solimp + 5 * id, 40); }
    else if (tp == CNSTR_LIMIT_JOINT) { memcpy(solref, m->jnt_solref + 2 * id, 16); memcpy(solimp, m->jnt_solimp + 5 * id, 40); }
    else { memcpy(solref, s->con[id].solref, 16); memcpy(solimp, s->con[id].solimp, 40); }
    if (solref[0] > 0) solref[0] = fmax(solref[0], 2 * m->timestep);   /* refsafe */
    solimp[0] = fmin(mjMAXIMP, fmax(mjMINIMP, solimp[0]));
    solimp[1] = fmin(mjMAXIMP, fmax(mjMINIMP, solimp[1]));
    solimp[2] = fmax(0, solimp[2]);
    solimp[3] = fmin(mjMAXIMP, fmax(mjMINIMP, solimp[3]));
    solimp[4] = fmax(1, solimp[4]);
    double imp, impP;
    get_impedance(solimp, s->efc_pos[i], s->efc_margin[i], &imp, &impP);
    s->efc_R[i] = fmax(mjMINVAL, (1 - imp) * s->efc_diagApprox[i] / imp);
    double dmax = solimp[1];
    if (solref[0] > 0) {
      s->efc_KBIP[i][0] = 1 / fmax(mjMINVAL, dmax * dmax * solref[0] * solref[0] * solref[1] * solref[1]);
      s->efc_KBIP[i][1] = 2 / fmax(mjMINVAL, dmax * solref[0]);
    } else {
      s->efc_KBIP[i][0] = -solref[0] / fmax(mjMINVAL, dmax * dmax);
      s->efc_KBIP[i][1] = -solref[1] / fmax(mjMINVAL, dmax);
    }
    s->efc_KBIP[i][2] = imp; s->efc_KBIP[i][3] = impP;
    int friction_row = (tp == CNSTR_FRICTION_DOF) ||
                       (tp == CNSTR_CONTACT_ELLIPTIC && i > s->con[id].efc_address);
    if (friction_row) s->efc_KBIP[i][0] = 0;
  }
  /* frictional contacts: adjust R in the friction dimensions, set regularised cone mu */
  for (int ci = 0; ci < s->ncon; ci++) {
    Contact* c = &s->con[ci];
    if (c->efc_address < 0 || c->dim == 1) continue;
    int i = c->efc_address, dim = c->dim;
    if (m->cone == 0) {
      c->mu = c->friction[0] / sqrt(fmax(mjMINVAL, m->impratio));
      double Rpy = 2 * c->mu * c->mu * s->efc_R[i];
      for (int j = 0; j < 2 * (dim - 1); j++) s->efc_R[i + j] = Rpy;
    } else {
      s->efc_R[i + 1] = s->efc_R[i] / fmax(mjMINVAL, m->impratio);
      c->mu = c->friction[0] * sqrt(s->efc_R[i + 1] / s->efc_R[i]);
      for (int j = 1; j < dim - 1; j++)
        s->efc_R[i + j + 1] = s->efc_R[i + 1] * c->friction[0] * c->friction[0] / (c->friction[j] * c->friction[j]);
    }
  }
  for (int i = 0; i < n; i++) s->efc_D[i] = 1 / s->efc_R[i];
}

/* ---------------------------------------------------------------------------------------------------- */
/* mj_fwdVelocity pieces                                                                                */
/* ---------------------------------------------------------------------------------------------------- */
static void comVel(RefSim* s) {
  const Model* m = &s->m;
  memset(s->cvel[0], 0, sizeof(double) * 6);
  for (int b = 1; b < m->nb; b++) {
    double cvel[6];
    memcpy(cvel, s->cvel[m->body_parentid[b]], sizeof(cvel));
    for (int k = 0; k < m->body_jntnum[b]; k++) {
      int j = m->body_jntadr[b] + k;
      crossMotion(s->cdof_dot[j], cvel, s->cdof[j]);
      for (int c = 0; c < 6; c++) cvel[c] += s->cdof[j][c] * s->qvel[j];
    }
    memcpy(s->cvel[b], cvel, sizeof(cvel));
  }
}
static void passive(RefSim* s) {
  const Model* m = &s->m;
  for (int i = 0; i < m->nv; i++)
    s->qfrc_passive[i] = -m->jnt_stiffness[i] * (s->qpos[i] - m->qpos_spring[i]) - m->dof_damping[i] * s->qvel[i];
}
static void reference_constraint(RefSim* s) {
  int nv = s->m.nv;
  for (int i = 0; i < s->nefc; i++) {
    double v = 0;
    for (int d = 0; d < nv; d++) v += s->J[i][d] * s->qvel[d];
    s->efc_vel[i] = v;
    s->efc_aref[i] = -s->efc_KBIP[i][1] * v - s->efc_KBIP[i][0] * s->efc_KBIP[i][2] * (s->efc_pos[i] - s->efc_margin[i]);
  }
}
static void rne_bias(RefSim* s) {
  const Model* m = &s->m;
  memset(s->cacc[0], 0, sizeof(double) * 6);
  for (int k = 0; k < 3; k++) s->cacc[0][3 + k] = -m->gravity[k];
  for (int b = 1; b < m->nb; b++) {
    double* a = s->cacc[b];
    memcpy(a, s->cacc[m->body_parentid[b]], sizeof(double) * 6);
    for (int k = 0; k < m->body_jntnum[b]; k++) {
      int j = m->body_jntadr[b] + k;
      for (int c = 0; c < 6; c++) a[c] += s->cdof_dot[j][c] * s->qvel[j];
    }
    double t1[6], t2[6], t3[6];
    mulInertVec(t1, s->cinert[b], a);
    mulInertVec(t2, s->cinert[b], s->cvel[b]);
    crossForce(t3, s->cvel[b], t2);
    for (int c = 0; c < 6; c++) s->cfrc[b][c] = t1[c] + t3[c];
  }
  memset(s->cfrc[0], 0, sizeof(double) * 6);
  for (int b = m->nb - 1; b > 0; b--) {
    int p = m->body_parentid[b];
    for (int c = 0; c < 6; c++) s->cfrc[p][c] += s->cfrc[b][c];
  }
  for (int j = 0; j < m->nv; j++) {
    double v = 0;
    const double* f = s->cfrc[m->jnt_bodyid[j]];
    for (int c = 0; c < 6; c++) v += s->cdof[j][c] * f[c];
    s->qfrc_bias[j] = v;
  }
}
static void actuation(RefSim* s) {
  const Model* m = &s->m;
  memset(s->qfrc_actuator, 0, sizeof(double) * m->nv);
  for (int i = 0; i < m->nu; i++) {
    double c = s->ctrl[i];
    if (m->actuator_ctrllimited[i]) c = fmin(m->actuator_ctrlrange[2 * i + 1], fmax(m->actuator_ctrlrange[2 * i], c));
    int d = m->actuator_dof[i];
    double f = m->actuator_gain[i] * c + m->actuator_bias[3 * i] + m->actuator_bias[3 * i + 1] * s->qpos[d] +
               m->actuator_bias[3 * i + 2] * s->qvel[d];
    if (m->actuator_forcelimited[i]) f = fmin(m->actuator_forcerange[2 * i + 1], fmax(m->actuator_forcerange[2 * i], f));
    s->qfrc_actuator[d] += m->actuator_gear[i] * f;
  }
}

/* ---------------------------------------------------------------------------------------------------- */
/* constraint solver: mj_constraintUpdate + mj_solNewton (engine_solver.c restated, dense)               */
/* ---------------------------------------------------------------------------------------------------- */
typedef struct {
  double Ma[MAXNV], Jaref[MAXEFC], grad[MAXNV], Mgrad[MAXNV], search[MAXNV], Mv[MAXNV], Jv[MAXEFC];
  double quad[MAXEFC][3];
  double cost, gauss;
  double H[MAXNV][MAXNV];
  double coneH[MAXCON][36];
} SolverCtx;

/* cost, forces and states for given jar = J*qacc - aref.  cone Hessians (jar-space, dim x dim) -> ctx->coneH */
static double constraint_update(RefSim* s, const double* jar, SolverCtx* ctx, int want_force) {
  const Model* m = &s->m;
  double cost = 0;
  int n = s->nefc;
  for (int i = 0; i < n; i++) {
    int tp = s->efc_type[i];
    double D = s->efc_D[i], R = s->efc_R[i];
    if (tp == CNSTR_FRICTION_DOF) {
      double f = s->efc_frictionloss[i];
      if (jar[i] <= -R * f) { s->efc_state[i] = ST_LINEARNEG; s->efc_force[i] = f; cost += -0.5 * R * f * f - f * jar[i]; }
      else if (jar[i] >= R * f) { s->efc_state[i] = ST_LINEARPOS; s->efc_force[i] = -f; cost += -0.5 * R * f * f + f * jar[i]; }
      else { s->efc_state[i] = ST_QUADRATIC; s->efc_force[i] = -D * jar[i]; cost += 0.5 * D * jar[i] * jar[i]; }
    } else if (tp == CNSTR_CONTACT_ELLIPTIC) {
      Contact* c = &s->con[s->efc_id[i]];
      int dim = c->dim;
      double mu = c->mu, U[6];
      U[0] = jar[i] * mu;
      double TT = 0;
      for (int j = 1; j < dim; j++) { U[j] = jar[i + j] * c->friction[j - 1]; TT += U[j] * U[j]; }
      double N = U[0], T = sqrt(TT);
      if (N >= mu * T || (T <= 0 && N >= 0)) {
        for (int j = 0; j < dim; j++) { s->efc_force[i + j] = 0; s->efc_state[i + j] = ST_SATISFIED; }
      } else if (mu * N + T <= 0 || (T <= 0 && N < 0)) {
        for (int j = 0; j < dim; j++) {
          s->efc_force[i + j] = -s->efc_D[i + j] * jar[i + j];
          s->efc_state[i + j] = ST_QUADRATIC;
          cost += 0.5 * s->efc_D[i + j] * jar[i + j] * jar[i + j];
        }
      } else {
        double Dm = s->efc_D[i] / fmax(mjMINVAL, mu * mu * (1 + mu * mu));
        double NmT = N - mu * T;
        cost += 0.5 * Dm * NmT * NmT;
        s->efc_force[i] = -Dm * NmT * mu;
        for (int j = 1; j < dim; j++) s->efc_force[i + j] = -s->efc_force[i] / T * U[j] * c->friction[j - 1];
        for (int j = 0; j < dim; j++) s->efc_state[i + j] = ST_CONE;
        if (ctx) {
          double* H = ctx->coneH[s->efc_id[i]];
          double scl[6];
          scl[0] = mu;
          for (int j = 1; j < dim; j++) scl[j] = c->friction[j - 1];
          for (int a = 0; a < dim; a++)
            for (int b = 0; b < dim; b++) {
              double h;
              if (a == 0 && b == 0) h = 1;
              else if (a == 0) h = -mu * U[b] / T;
              else if (b == 0) h = -mu * U[a] / T;
              else h = mu * N / (T * T * T) * U[a] * U[b] + (a == b ? (mu * mu - mu * N / T) : 0);
              H[a * dim + b] = Dm * h * scl[a] * scl[b];
            }
        }
      }
      i += dim - 1;
    } else {
      if (jar[i] < 0) { s->efc_state[i] = ST_QUADRATIC; s->efc_force[i] = -D * jar[i]; cost += 0.5 * D * jar[i] * jar[i]; }
      else { s->efc_state[i] = ST_SATISFIED; s->efc_force[i] = 0; }
    }
  }
  (void)m; (void)want_force;
  return cost;
}

static void mulJ(const RefSim* s, double* res, const double* v) {
  int nv = s->m.nv;
  for (int i = 0; i < s->nefc; i++) {
    double a = 0;
    for (int d = 0; d < nv; d++) a += s->J[i][d] * v[d];
    res[i] = a;
  }
}
static void mulJT(const RefSim* s, double* res, const double* f) {
  int nv = s->m.nv;
  for (int d = 0; d < nv; d++) res[d] = 0;
  for (int i = 0; i < s->nefc; i++) {
    if (f[i] == 0) continue;
    for (int d = 0; d < nv; d++) res[d] += s->J[i][d] * f[i];
  }
}
static void mulM(const RefSim* s, double* res, const double* v) {
  int nv = s->m.nv;
  for (int i = 0; i < nv; i++) {
    double a = 0;
    for (int j = 0; j < nv; j++) a += s->M[i][j] * v[j];
    res[i] = a;
  }
}

static void primal_update_constraint(RefSim* s, SolverCtx* ctx) {
  int nv = s->m.nv;
  ctx->cost = constraint_update(s, ctx->Jaref, ctx, 1);
  mulJT(s, s->qfrc_constraint, s->efc_force);
  double g = 0;
  for (int i = 0; i < nv; i++) g += 0.5 * (ctx->Ma[i] - s->qfrc_smooth[i]) * (s->qacc[i] - s->qacc_smooth[i]);
  ctx->gauss = g;
  ctx->cost += g;
}

static void make_hessian(RefSim* s, SolverCtx* ctx) {
  int nv = s->m.nv;
  for (int i = 0; i < nv; i++) for (int j = 0; j <= i; j++) ctx->H[i][j] = s->M[i][j];
  for (int r = 0; r < s->nefc; r++) {
    if (s->efc_state[r] == ST_QUADRATIC) {
      double D = s->efc_D[r];
      const double* Jr = s->J[r];
      for (int i = 0; i < nv; i++) {
        if (Jr[i] == 0) continue;
        double t = D * Jr[i];
        for (int j = 0; j <= i; j++) ctx->H[i][j] += t * Jr[j];
      }
    } else if (s->efc_state[r] == ST_CONE) {
      Contact* c = &s->con[s->efc_id[r]];
      int dim = c->dim;
      const double* Hc = ctx->coneH[s->efc_id[r]];
      for (int a = 0; a < dim; a++)
        for (int b = 0; b < dim; b++) {
          double h = Hc[a * dim + b];
          if (h == 0) continue;
          const double *Ja = s->J[r + a], *Jb = s->J[r + b];
          for (int i = 0; i < nv; i++) {
            if (Ja[i] == 0) continue;
            double t = h * Ja[i];
            for (int j = 0; j <= i; j++) ctx->H[i][j] += t * Jb[j];
          }
        }
      r += dim - 1;
    }
  }
  chol_factor(ctx->H, nv);
}

static void primal_update_gradient(RefSim* s, SolverCtx* ctx) {
  int nv = s->m.nv;
  for (int i = 0; i < nv; i++) ctx->grad[i] = ctx->Ma[i] - s->qfrc_smooth[i] - s->qfrc_constraint[i];
  memcpy(ctx->Mgrad, ctx->grad, sizeof(double) * nv);
  chol_solve(ctx->H, nv, ctx->Mgrad);
}

/* line-search objective: cost(alpha) with first and second derivative (PrimalEval) */
typedef struct { double alpha, cost, d1, d2; } LSPoint;

static void ls_eval(RefSim* s, SolverCtx* ctx, const double* quadGauss, LSPoint* p, double alpha) {
  double c = alpha * alpha * quadGauss[2] + alpha * quadGauss[1] + quadGauss[0];
  double d1 = 2 * alpha * quadGauss[2] + quadGauss[1];
  double d2 = 2 * quadGauss[2];
  int n = s->nefc;
  for (int i = 0; i < n; i++) {
    int tp = s->efc_type[i];
    double x = ctx->Jaref[i] + alpha * ctx->Jv[i];
    const double* q = ctx->quad[i];
    if (tp == CNSTR_FRICTION_DOF) {
      double f = s->efc_frictionloss[i], Rf = s->efc_R[i] * f;
      if (x <= -Rf) { c += f * (-0.5 * Rf - ctx->Jaref[i]) - alpha * f * ctx->Jv[i]; d1 += -f * ctx->Jv[i]; }
      else if (x >= Rf) { c += f * (-0.5 * Rf + ctx->Jaref[i]) + alpha * f * ctx->Jv[i]; d1 += f * ctx->Jv[i]; }
      else { c += alpha * alpha * q[2] + alpha * q[1] + q[0]; d1 += 2 * alpha * q[2] + q[1]; d2 += 2 * q[2]; }
    } else if (tp == CNSTR_CONTACT_ELLIPTIC) {
      Contact* con = &s->con[s->efc_id[i]];
      int dim = con->dim;
      double mu = con->mu;
      double U0 = ctx->Jaref[i] * mu, V0 = ctx->Jv[i] * mu, UU = 0, UV = 0, VV = 0;
      for (int j = 1; j < dim; j++) {
        double fr = con->friction[j - 1];
        double u = ctx->Jaref[i + j] * fr, v = ctx->Jv[i + j] * fr;
        UU += u * u; UV += u * v; VV += v * v;
      }
      double N = U0 + alpha * V0;
      double Tsqr = UU + alpha * (2 * UV + alpha * VV);
      if (Tsqr <= 0) {
        if (N < 0) {
          for (int j = 0; j < dim; j++) {
            const double* qq = ctx->quad[i + j];
            c += alpha * alpha * qq[2] + alpha * qq[1] + qq[0]; d1 += 2 * alpha * qq[2] + qq[1]; d2 += 2 * qq[2];
          }
        }
      } else {
        double T = sqrt(Tsqr);
        if (N >= mu * T) {
          /* top zone: nothing */
        } else if (mu * N + T <= 0) {
          for (int j = 0; j < dim; j++) {
            const double* qq = ctx->quad[i + j];
            c += alpha * alpha * qq[2] + alpha * qq[1] + qq[0]; d1 += 2 * alpha * qq[2] + qq[1]; d2 += 2 * qq[2];
          }
        } else {
          double Dm = s->efc_D[i] / fmax(mjMINVAL, mu * mu * (1 + mu * mu));
          double N1 = V0, T1 = (UV + alpha * VV) / T, T2 = VV / T - (UV + alpha * VV) * T1 / (T * T);
          double NmT = N - mu * T;
          c += 0.5 * Dm * NmT * NmT;
          d1 += Dm * NmT * (N1 - mu * T1);
          d2 += Dm * ((N1 - mu * T1) * (N1 - mu * T1) + NmT * (-mu * T2));
        }
      }
      i += dim - 1;
    } else {
      if (x < 0) { c += alpha * alpha * q[2] + alpha * q[1] + q[0]; d1 += 2 * alpha * q[2] + q[1]; d2 += 2 * q[2]; }
    }
  }
  if (d2 <= 0) d2 = mjMINVAL;
  p->alpha = alpha; p->cost = c; p->d1 = d1; p->d2 = d2;
}

/* exact line search (PrimalSearch): returns alpha minimising cost(qacc + alpha*search) to gradient tolerance */
static double primal_search(RefSim* s, SolverCtx* ctx, double scale) {
  const Model* m = &s->m;
  int nv = m->nv, n = s->nefc;
  double snorm = 0;
  for (int i = 0; i < nv; i++) snorm += ctx->search[i] * ctx->search[i];
  snorm = sqrt(snorm);
  if (snorm < mjMINVAL) return 0;
  const double ls_tolerance = 0.01;
  const int ls_iterations = 50;
  double gtol = m->tolerance * ls_tolerance * snorm / scale;
  mulM(s, ctx->Mv, ctx->search);
  mulJ(s, ctx->Jv, ctx->search);
  double quadGauss[3] = {ctx->gauss, 0, 0};
  for (int i = 0; i < nv; i++) {
    quadGauss[1] += ctx->search[i] * (ctx->Ma[i] - s->qfrc_smooth[i]);
    quadGauss[2] += 0.5 * ctx->search[i] * ctx->Mv[i];
  }
  for (int i = 0; i < n; i++) {
    double D = s->efc_D[i];
    ctx->quad[i][0] = 0.5 * D * ctx->Jaref[i] * ctx->Jaref[i];
    ctx->quad[i][1] = D * ctx->Jv[i] * ctx->Jaref[i];
    ctx->quad[i][2] = 0.5 * D * ctx->Jv[i] * ctx->Jv[i];
  }
  LSPoint p0, p1, p2, pmid;
  int iter = 0;
  ls_eval(s, ctx, quadGauss, &p0, 0);
  ls_eval(s, ctx, quadGauss, &p1, p0.alpha - p0.d1 / p0.d2);
  if (p0.cost < p1.cost) p1 = p0;
  if (fabs(p1.d1) < gtol) return p1.alpha;
  /* stage 1: Newton steps in one direction until the derivative changes sign (bracket) */
  int dir = p1.d1 < 0 ? 1 : -1;
  p2 = p1;
  while (p1.d1 * dir <= -gtol && iter < ls_iterations) {
    p2 = p1;
    ls_eval(s, ctx, quadGauss, &p1, p1.alpha - p1.d1 / p1.d2);
    iter++;
    if (fabs(p1.d1) < gtol) return p1.alpha;
  }
  if (iter >= ls_iterations) return p1.alpha;
  /* stage 2: bracketed: p2 has d1*dir < 0 , p1 has d1*dir > 0 ; safeguarded Newton / bisection */
  LSPoint lo = p2, hi = p1;   /* lo: derivative sign = -dir (descent in dir), hi: opposite */
  LSPoint best = (p1.cost < p2.cost) ? p1 : p2;
  while (iter < ls_iterations) {
    /* Newton candidates from both ends, and the midpoint */
    double cand[3] = {lo.alpha - lo.d1 / lo.d2, hi.alpha - hi.d1 / hi.d2, 0.5 * (lo.alpha + hi.alpha)};
    double amin = fmin(lo.alpha, hi.alpha), amax = fmax(lo.alpha, hi.alpha);
    int moved = 0;
    for (int k = 0; k < 3; k++) {
      double a = cand[k];
      if (!(a > amin && a < amax)) continue;
      ls_eval(s, ctx, quadGauss, &pmid, a);
      iter++;
      if (fabs(pmid.d1) < gtol) return pmid.alpha;
      if (pmid.cost < best.cost) best = pmid;
      if (pmid.d1 * dir < 0) lo = pmid; else hi = pmid;
      moved = 1;
      amin = fmin(lo.alpha, hi.alpha); amax = fmax(lo.alpha, hi.alpha);
    }
    if (!moved || amax - amin < 1e-16 * fmax(1.0, fabs(amax))) break;
  }
  return best.alpha;
}

static void solve_newton(RefSim* s) {
  const Model* m = &s->m;
  int nv = m->nv, n = s->nefc;
  static __thread SolverCtx ctx_;
  SolverCtx* ctx = &ctx_;
  double scale = 1.0 / (m->meaninertia * (nv > 1 ? nv : 1));
  mulM(s, ctx->Ma, s->qacc);
  mulJ(s, ctx->Jaref, s->qacc);
  for (int i = 0; i < n; i++) ctx->Jaref[i] -= s->efc_aref[i];
  primal_update_constraint(s, ctx);
  make_hessian(s, ctx);
  primal_update_gradient(s, ctx);
  for (int i = 0; i < nv; i++) ctx->search[i] = -ctx->Mgrad[i];
  int iter = 0;
  while (iter < m->iterations) {
    double alpha = primal_search(s, ctx, scale);
    if (alpha == 0) break;
    for (int i = 0; i < nv; i++) { s->qacc[i] += alpha * ctx->search[i]; ctx->Ma[i] += alpha * ctx->Mv[i]; }
    for (int i = 0; i < n; i++) ctx->Jaref[i] += alpha * ctx->Jv[i];
    double oldcost = ctx->cost;
    primal_update_constraint(s, ctx);
    make_hessian(s, ctx);
    primal_update_gradient(s, ctx);
    double improvement = scale * (oldcost - ctx->cost);
    double gn = 0;
    for (int i = 0; i < nv; i++) gn += ctx->grad[i] * ctx->grad[i];
    double gradient = scale * sqrt(gn);
    iter++;
    if (improvement < m->tolerance || gradient < m->tolerance) break;
    for (int i = 0; i < nv; i++) ctx->search[i] = -ctx->Mgrad[i];
  }
  s->solver_iter = iter;
}

static void fwd_constraint(RefSim* s) {
  const Model* m = &s->m;
  int nv = m->nv, n = s->nefc;
  if (n == 0) {
    memcpy(s->qacc, s->qacc_smooth, sizeof(double) * nv);
    memcpy(s->qacc_warmstart, s->qacc_smooth, sizeof(double) * nv);
    memset(s->qfrc_constraint, 0, sizeof(double) * nv);
    s->solver_iter = 0;
    return;
  }
  mulJ(s, s->efc_b, s->qacc_smooth);
  for (int i = 0; i < n; i++) s->efc_b[i] -= s->efc_aref[i];
  /* warmstart: pick the cheaper of qacc_warmstart and qacc_smooth */
  {
    static __thread double jar[MAXEFC];
    double Ma[MAXNV];
    memcpy(s->qacc, s->qacc_warmstart, sizeof(double) * nv);
    mulJ(s, jar, s->qacc);
    for (int i = 0; i < n; i++) jar[i] -= s->efc_aref[i];
    double cw = constraint_update(s, jar, NULL, 0);
    mulM(s, Ma, s->qacc);
    for (int i = 0; i < nv; i++) cw += 0.5 * (Ma[i] - s->qfrc_smooth[i]) * (s->qacc[i] - s->qacc_smooth[i]);
    double cs = constraint_update(s, s->efc_b, NULL, 0);
    if (cw > cs) memcpy(s->qacc, s->qacc_smooth, sizeof(double) * nv);
  }
  solve_newton(s);
  memcpy(s->qacc_warmstart, s->qacc, sizeof(double) * nv);
}

/* ---------------------------------------------------------------------------------------------------- */
/* mj_forward / integrators / mj_step                                                                   */
/* ---------------------------------------------------------------------------------------------------- */
static void forward(RefSim* s) {
  const Model* m = &s->m;
  int nv = m->nv;
  kinematics(s);
  comPos(s);
  crb(s);
  collision(s);
  make_constraint(s);
  comVel(s);
  passive(s);
  reference_constraint(s);
  rne_bias(s);
  actuation(s);
  for (int i = 0; i < nv; i++) {
    s->qfrc_smooth[i] = s->qfrc_passive[i] - s->qfrc_bias[i] + s->qfrc_actuator[i];
    s->qacc_smooth[i] = s->qfrc_smooth[i];
  }
  chol_solve(s->LM, nv, s->qacc_smooth);
  fwd_constraint(s);
}
void ref_forward(RefSim* s, const double* ctrl) {
  if (ctrl) memcpy(s->ctrl, ctrl, sizeof(double) * s->m.nu);
  forward(s);
}

static void euler(RefSim* s) {
  const Model* m = &s->m;
  int nv = m->nv;
  double h = m->timestep;
  int has_damping = 0;
  for (int i = 0; i < nv; i++) if (m->dof_damping[i] > 0) has_damping = 1;
  double qacc[MAXNV];
  if (!has_damping) memcpy(qacc, s->qacc, sizeof(double) * nv);
  else {
    static __thread double A[MAXNV][MAXNV];
    for (int i = 0; i < nv; i++) {
      for (int j = 0; j <= i; j++) A[i][j] = s->M[i][j];
      A[i][i] += h * m->dof_damping[i];
      qacc[i] = s->qfrc_smooth[i] + s->qfrc_constraint[i];
    }
    chol_factor(A, nv);
    chol_solve(A, nv, qacc);
  }
  for (int i = 0; i < nv; i++) s->qvel[i] += h * qacc[i];
  for (int i = 0; i < nv; i++) s->qpos[i] += h * s->qvel[i];
}

static void rk4(RefSim* s) {
  const Model* m = &s->m;
  int nv = m->nv;
  double h = m->timestep;
  static const double A[3][3] = {{0.5, 0, 0}, {0, 0.5, 0}, {0, 0, 1}};
  static const double B[4] = {1.0 / 6, 1.0 / 3, 1.0 / 3, 1.0 / 6};
  double X0q[MAXNV], X0v[MAXNV], Fv[4][MAXNV], Fa[4][MAXNV];
  memcpy(X0q, s->qpos, sizeof(double) * nv);
  memcpy(X0v, s->qvel, sizeof(double) * nv);
  memcpy(Fv[0], s->qvel, sizeof(double) * nv);
  memcpy(Fa[0], s->qacc, sizeof(double) * nv);
  for (int i = 1; i < 4; i++) {
    for (int d = 0; d < nv; d++) {
      double dq = 0, dv = 0;
      for (int j = 0; j < i; j++) { dq += A[i - 1][j] * Fv[j][d]; dv += A[i - 1][j] * Fa[j][d]; }
      s->qpos[d] = X0q[d] + h * dq;
      s->qvel[d] = X0v[d] + h * dv;
    }
    forward(s);
    memcpy(Fv[i], s->qvel, sizeof(double) * nv);
    memcpy(Fa[i], s->qacc, sizeof(double) * nv);
  }
  for (int d = 0; d < nv; d++) {
    double dq = 0, dv = 0;
    for (int j = 0; j < 4; j++) { dq += B[j] * Fv[j][d]; dv += B[j] * Fa[j][d]; }
    s->qvel[d] = X0v[d] + h * dv;
    s->qpos[d] = X0q[d] + h * dq;
  }
}

void ref_step(RefSim* s, const double* ctrl, int nsub) {
  if (ctrl) memcpy(s->ctrl, ctrl, sizeof(double) * s->m.nu);
  for (int k = 0; k < nsub; k++) {
    forward(s);
    if (s->m.integrator == 1) rk4(s); else euler(s);
  }
}

/* ---------------------------------------------------------------------------------------------------- */
/* Full LocoEnv.step() contract on top of RefSim (TaskSpec, include/locosim_task.h)                       */
/* ---------------------------------------------------------------------------------------------------- */
typedef struct {
  int obs_dim, n_done, reward_type, n_substeps, n_traj, traj_len, n_goal, recenter[2], ri[4], use_absorbing;
  int n_grf, n_grf_geom, rot[3];
  double rp[2], track[4];
  int *ibuf; double* rbuf;
  const int *obs_src_type, *obs_src_idx, *done_obs_idx, *act_idx, *grf_group;
  const double *act_mean, *act_delta, *done_lo, *done_hi, *table;
} Task;

static int task_load(Task* t, int nu, int nv, const int* ti, int nti, const double* tr, int ntr) {
  if (nti < TKI_HEADER_LEN || ti[TKI_MAGIC] != LOCOSIM_TASK_MAGIC || ti[TKI_VERSION] != LOCOSIM_TASK_VERSION) return -1;
  t->ibuf = (int*)malloc(sizeof(int) * nti); memcpy(t->ibuf, ti, sizeof(int) * nti);
  t->rbuf = (double*)malloc(sizeof(double) * ntr); memcpy(t->rbuf, tr, sizeof(double) * ntr);
  t->obs_dim = ti[TKI_OBS_DIM]; t->n_done = ti[TKI_N_DONE]; t->reward_type = ti[TKI_REWARD_TYPE];
  t->n_substeps = ti[TKI_N_SUBSTEPS]; t->n_traj = ti[TKI_N_TRAJ]; t->traj_len = ti[TKI_TRAJ_LEN]; t->n_goal = ti[TKI_N_GOAL];
  t->recenter[0] = ti[TKI_RECENTER0]; t->recenter[1] = ti[TKI_RECENTER1];
  for (int k = 0; k < 4; k++) t->ri[k] = ti[TKI_REWARD_I0 + k];
  t->use_absorbing = ti[TKI_USE_ABSORBING];
  t->n_grf = ti[TKI_N_GRF]; t->n_grf_geom = ti[TKI_N_GRF_GEOM];
  if (t->n_grf < 0 || t->n_grf > LS_MAX_GRF) return -3;
  t->rp[0] = tr[TKR_REWARD_P0]; t->rp[1] = tr[TKR_REWARD_P1];
  for (int k = 0; k < 3; k++) t->rot[k] = ti[TKI_ROT_Q + k];
  for (int k = 0; k < 4; k++) t->track[k] = tr[TKR_TRACK_WP + k];
  const int* ip = t->ibuf + TKI_HEADER_LEN;
  t->obs_src_type = ip; ip += t->obs_dim; t->obs_src_idx = ip; ip += t->obs_dim; t->done_obs_idx = ip; ip += t->n_done;
  t->act_idx = ip; ip += nu;
  t->grf_group = ip; ip += t->n_grf_geom;
  const double* rp = t->rbuf + TKR_HEADER_LEN;
  t->act_mean = rp; rp += nu; t->act_delta = rp; rp += nu; t->done_lo = rp; rp += t->n_done; t->done_hi = rp; rp += t->n_done;
  t->table = rp; rp += (long)t->n_traj * t->traj_len * (2 * nv + t->n_goal);
  if (ip - t->ibuf != nti || rp - t->rbuf != ntr) return -2;
  return 0;
}

struct RefEnv { RefSim* sim; Task task; double goal[8]; double obs[128]; double grf[3 * LS_MAX_GRF]; double user[LS_POOL_USER];
                int cursor; double rot_angle; };

RefEnv* refenv_create(const int* ints, int n_ints, const double* reals, int n_reals, const int* ti, int nti,
                      const double* tr, int ntr) {
  RefEnv* e = (RefEnv*)calloc(1, sizeof(RefEnv));
  e->sim = ref_create(ints, n_ints, reals, n_reals);
  if (!e->sim) { free(e); return NULL; }
  int rc = task_load(&e->task, e->sim->m.nu, e->sim->m.nv, ti, nti, tr, ntr);
  if (rc) { fprintf(stderr, "locosim_ref: bad TaskSpec (%d)\n", rc); ref_destroy(e->sim); free(e); return NULL; }
  return e;
}
void refenv_destroy(RefEnv* e) { if (e) { ref_destroy(e->sim); free(e->task.ibuf); free(e->task.rbuf); free(e); } }
int refenv_obs_dim(const RefEnv* e) { return e->task.obs_dim; }
void refenv_set_user(RefEnv* e, const double* user) { memcpy(e->user, user, sizeof(e->user)); }
void refenv_set_rotation(RefEnv* e, double angle) { e->rot_angle = angle; }
int refenv_cursor(const RefEnv* e) { return e->cursor; }
RefSim* refenv_sim(RefEnv* e) { return e->sim; }

static void build_obs(const RefEnv* e, double* obs) {
  const Task* t = &e->task;
  for (int k = 0; k < t->obs_dim; k++) {
    int idx = t->obs_src_idx[k];
    switch (t->obs_src_type[k]) {
      case LS_OBS_QPOS: obs[k] = e->sim->qpos[idx]; break;
      case LS_OBS_QVEL: obs[k] = e->sim->qvel[idx]; break;
      case LS_OBS_PARAM: obs[k] = e->user[idx]; break;   /* e.g. the carried weight's mass (base_robot_humanoid.py:119-123) */
      case LS_OBS_GRF: obs[k] = e->grf[idx] / (1000.0 * t->n_substeps); break;   /* mean_grf.mean / 1000 (base.py:596-599) */
      default: obs[k] = e->goal[idx]; break;
    }
  }
}
static int has_fallen(const Task* t, const double* obs) {
  int f = 0;
  for (int k = 0; k < t->n_done; k++) {
    double v = obs[t->done_obs_idx[k]];
    if (v < t->done_lo[k] || v > t->done_hi[k]) f = 1;
  }
  return f;
}
static double reward_fn(const Task* t, const double* prev) {
  switch (t->reward_type) {
    case LS_REWARD_TARGET_VELOCITY: { double d = prev[t->ri[0]] - t->rp[0]; return exp(-d * d); }
    case LS_REWARD_VELOCITY_VECTOR: {
      double g = prev[t->ri[3]];
      double dx = prev[t->ri[0]] - g * prev[t->ri[2]], dy = prev[t->ri[1]] - g * prev[t->ri[2] + 1];
      return exp(-5.0 * sqrt(dx * dx + dy * dy));
    }
    case LS_REWARD_POS: return prev[t->ri[0]];
    default: return 0;
  }
}
void refenv_reset_to(RefEnv* e, int traj_no, int step_no, double* obs) {
  const Task* t = &e->task;
  int nv = e->sim->m.nv, ncol = 2 * nv + t->n_goal;
  const double* row = t->table + ((long)traj_no * t->traj_len + step_no) * ncol;
  double qpos[MAXNV], qvel[MAXNV];
  memcpy(qpos, row, sizeof(double) * nv);
  memcpy(qvel, row + nv, sizeof(double) * nv);
  if (t->recenter[0] >= 0) qpos[t->recenter[0]] = 0;
  if (t->recenter[1] >= 0) qpos[t->recenter[1]] = 0;
  if (t->rot[0] >= 0) {
    /* setup_random_rot (unitreeA1.py:270-285 -> utils/math.py:5-31 rotate_obs): yaw + angle wrapped to [-pi, pi), root
       (vx, vy) rotated; the angle comes from the caller (refenv_set_rotation), 0 otherwise */
    double a = e->rot_angle, vx = qvel[t->rot[1]], vy = qvel[t->rot[2]];
    qpos[t->rot[0]] = fmod(fmod(qpos[t->rot[0]] + a + M_PI, 2 * M_PI) + 2 * M_PI, 2 * M_PI) - M_PI;
    qvel[t->rot[1]] = cos(a) * vx - sin(a) * vy;
    qvel[t->rot[2]] = sin(a) * vx + cos(a) * vy;
  }
  e->cursor = traj_no * t->traj_len + step_no;
  ref_reset(e->sim, qpos, qvel);
  for (int k = 0; k < t->n_goal; k++) e->goal[k] = row[2 * nv + k];
  memset(e->grf, 0, sizeof(e->grf));   /* RunningAveragedWindow reset with the episode */
  build_obs(e, e->obs);
  if (obs) memcpy(obs, e->obs, sizeof(double) * t->obs_dim);
}
void refenv_step(RefEnv* e, const double* action, double* obs, double* reward, int* absorbing) {
  const Task* t = &e->task;
  int nu = e->sim->m.nu;
  double ctrl[MAXNV], cur[128];
  for (int i = 0; i < nu; i++) ctrl[t->act_idx[i]] = action[i] * t->act_delta[i] + t->act_mean[i];
  if (t->n_grf == 0) ref_step(e->sim, ctrl, t->n_substeps);
  else {
    /* use_foot_forces (base.py:94-98,623-631): n_intermediate_steps x mj_step(1), after each one the contact-frame force
       of the FIRST floor contact of every foot group (mushroom _get_collision_force -> mj_contactForce); the window
       of the running mean is exactly the n_substeps samples of this control step */
    memset(e->grf, 0, sizeof(e->grf));
    for (int k = 0; k < t->n_substeps; k++) {
      ref_step(e->sim, ctrl, 1);
      for (int g = 0; g < t->n_grf; g++) {
        for (int ci = 0; ci < e->sim->ncon; ci++) {
          int a = t->grf_group[e->sim->con[ci].geom1], b = t->grf_group[e->sim->con[ci].geom2];
          if (!((a == LS_GRF_FLOOR && b == g) || (b == LS_GRF_FLOOR && a == g))) continue;
          double o[16];
          ref_get_contact(e->sim, ci, o);
          for (int c = 0; c < 3; c++) e->grf[3 * g + c] += o[10 + c];
          break;
        }
      }
    }
  }
  build_obs(e, cur);
  int ab = t->use_absorbing ? has_fallen(t, cur) : 0;
  if (t->reward_type == LS_REWARD_TRACKING) {
    /* this package's mocap-tracking reward (include/locosim_task.h): cursor + 1 (clamped), post-step observation
       against that table row */
    int nv = e->sim->m.nv, tj = e->cursor / t->traj_len, sp = e->cursor % t->traj_len;
    sp = sp + 1 < t->traj_len ? sp + 1 : t->traj_len - 1;
    e->cursor = tj * t->traj_len + sp;
    const double* ref = t->table + (long)e->cursor * (2 * nv + t->n_goal);
    double ep = 0, ev = 0;
    for (int k = 0; k < t->obs_dim; k++) {
      int idx = t->obs_src_idx[k];
      if (t->obs_src_type[k] == LS_OBS_QPOS) { double d = e->sim->qpos[idx] - ref[idx]; ep += d * d; }
      else if (t->obs_src_type[k] == LS_OBS_QVEL) { double d = e->sim->qvel[idx] - ref[nv + idx]; ev += d * d; }
    }
    if (reward) *reward = t->track[0] * exp(-t->track[1] * ep) + t->track[2] * exp(-t->track[3] * ev);
  } else if (reward) *reward = reward_fn(t, e->obs);
  if (absorbing) *absorbing = ab;
  memcpy(e->obs, cur, sizeof(double) * t->obs_dim);
  if (obs) memcpy(obs, cur, sizeof(double) * t->obs_dim);
}

/* ---- batched CPU baseline loop (pthreads over independent envs) ------------------------------------- */
typedef struct {
  const int* ints; int n_ints; const double* reals; int n_reals; const int* ti; int nti; const double* tr; int ntr;
  int env0, env1, n_steps; unsigned long long seed; double* obs_out; long steps, resets; double cpu_s;
} Worker;
static inline unsigned long long splitmix(unsigned long long* x) {
  unsigned long long z = (*x += 0x9E3779B97F4A7C15ULL);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; return z ^ (z >> 31);
}
static void* worker_main(void* arg) {
  Worker* w = (Worker*)arg;
  RefEnv* e = refenv_create(w->ints, w->n_ints, w->reals, w->n_reals, w->ti, w->nti, w->tr, w->ntr);
  int nu = e->sim->m.nu, D = e->task.obs_dim;
  double act[MAXNV], obs[128], r; int ab;
  for (int env = w->env0; env < w->env1; env++) {
    unsigned long long st = w->seed * 0x100000001B3ULL + env;
    refenv_reset_to(e, (int)(splitmix(&st) % e->task.n_traj), (int)(splitmix(&st) % e->task.traj_len), obs);
    for (int k = 0; k < w->n_steps; k++) {
      for (int i = 0; i < nu; i++) act[i] = (double)(splitmix(&st) >> 11) * (2.0 / 9007199254740992.0) - 1.0;
      refenv_step(e, act, obs, &r, &ab);
      w->steps++;
      int bad = 0;
      for (int i = 0; i < D; i++) if (!isfinite(obs[i])) bad = 1;
      if (ab || bad) {
        refenv_reset_to(e, (int)(splitmix(&st) % e->task.n_traj), (int)(splitmix(&st) % e->task.traj_len), obs);
        w->resets++;
      }
    }
    if (w->obs_out) memcpy(w->obs_out + (long)env * D, obs, sizeof(double) * D);
  }
  refenv_destroy(e);
  struct timespec ts;
  if (clock_gettime(CLOCK_THREAD_CPUTIME_ID, &ts) == 0) w->cpu_s = (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
  return NULL;
}
/* stats_out (may be NULL): [0] wall seconds of the threaded region, [1] sum of the workers' own CPU seconds
   (CLOCK_THREAD_CPUTIME_ID; [1]/[0] = cores the host actually granted), [2] slowest worker's CPU seconds,
   [3] number of distinct CPUs the workers were pinned to (0: not pinned). pin != 0: worker t is pinned to the t-th CPU
   (round robin) of the calling thread's affinity mask. */
long ref_rollout_ex(const int* ints, int n_ints, const double* reals, int n_reals, const int* ti, int nti, const double* tr,
                    int ntr, int n_envs, int n_steps, int nthreads, unsigned long long seed, double* obs_out,
                    long* n_resets_out, int pin, double* stats_out) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > n_envs) nthreads = n_envs;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
  Worker* ws = (Worker*)calloc(nthreads, sizeof(Worker));
  cpu_set_t allowed;
  int cpus[CPU_SETSIZE], ncpu = 0;
  CPU_ZERO(&allowed);
  if (pin && sched_getaffinity(0, sizeof(allowed), &allowed) == 0)
    for (int c = 0; c < CPU_SETSIZE; c++) if (CPU_ISSET(c, &allowed)) cpus[ncpu++] = c;
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int t = 0; t < nthreads; t++) {
    Worker w = {ints, n_ints, reals, n_reals, ti, nti, tr, ntr, (int)((long)n_envs * t / nthreads),
                (int)((long)n_envs * (t + 1) / nthreads), n_steps, seed, obs_out, 0, 0, 0.0};
    ws[t] = w;
    pthread_attr_t at;
    pthread_attr_init(&at);
    if (ncpu > 0) {
      cpu_set_t one;
      CPU_ZERO(&one);
      CPU_SET(cpus[t % ncpu], &one);
      pthread_attr_setaffinity_np(&at, sizeof(one), &one);
    }
    if (pthread_create(&th[t], &at, worker_main, &ws[t]) != 0) pthread_create(&th[t], NULL, worker_main, &ws[t]);
    pthread_attr_destroy(&at);
  }
  long total = 0, resets = 0;
  double cpu_sum = 0, cpu_max = 0;
  for (int t = 0; t < nthreads; t++) {
    pthread_join(th[t], NULL);
    total += ws[t].steps; resets += ws[t].resets; cpu_sum += ws[t].cpu_s;
    if (ws[t].cpu_s > cpu_max) cpu_max = ws[t].cpu_s;
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  if (n_resets_out) *n_resets_out = resets;
  if (stats_out) {
    stats_out[0] = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    stats_out[1] = cpu_sum; stats_out[2] = cpu_max; stats_out[3] = ncpu < nthreads ? ncpu : nthreads;
  }
  free(th); free(ws);
  return total;
}
long ref_rollout(const int* ints, int n_ints, const double* reals, int n_reals, const int* ti, int nti, const double* tr,
                 int ntr, int n_envs, int n_steps, int nthreads, unsigned long long seed, double* obs_out,
                 long* n_resets_out) {
  return ref_rollout_ex(ints, n_ints, reals, n_reals, ti, nti, tr, ntr, n_envs, n_steps, nthreads, seed, obs_out,
                        n_resets_out, 0, NULL);
}
