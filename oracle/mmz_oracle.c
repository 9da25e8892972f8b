/* mmz_oracle.c - CPU float64 restatement of the reference's MazeEnv.step path.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under mujoco-maze_b200/ may include, link
 * or call this file; it is used by tests/, __graft_entry__.smoke() and
 * bench.py's CPU-baseline legs as the checker / reported baseline.
 *
 * What it restates, one environment at a time, scalar code, double precision:
 *   MazeEnv.step / _get_obs            reference mujoco_maze/maze_env.py:448-481, 351-369
 *   PointEnv.step                      reference mujoco_maze/point.py:44-61
 *   AntEnv.step / SwimmerEnv.step      reference mujoco_maze/ant.py:56-73, swimmer.py:32-47
 *   CollisionDetector.detect           reference mujoco_maze/maze_env_utils.py:186-206
 *   MazeTask.reward / termination      reference mujoco_maze/maze_task.py:43-47, 77-81 + variants
 *   mj_step (RK4) / mj_forward         THIRD PARTY: MuJoCo 2.0 via mujoco-py 2.0.2.13
 *                                      (reference poetry.lock:145-146), not in the reference
 *                                      tree and not installable here. Restated from MuJoCo's
 *                                      published "Computation" chapter.
 *
 * PARITY STATUS: the Python half (clamp, reward, termination, obs assembly) is
 * pinned against outputs of the real reference code (tests/golden/). The
 * PHYSICS IS PARITY-UNPINNED: no MuJoCo binary is available to pin it, and the
 * reference's own tests assert only shapes and reward signs. It is validated by
 * closed-form and conservation checks instead (tests/test_oracle_physics.py).
 */
#define MMZ_REAL_IS_DOUBLE
#include "../include/mmz_model.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NB MMZ_MAXBODY
#define NJ MMZ_MAXJNT
#define ND MMZ_MAXDOF
#define NQ MMZ_MAXQ
#define NG MMZ_MAXGEOM
#define MAXCON 96
#define MAXEFC (4 * MAXCON + 2 * NJ)
#define MINVAL 1e-15
#define MAXVAL 1e10
#define PI 3.14159265358979323846

typedef struct {
  double dist, pos[3], frame[9]; /* frame rows: normal, tangent1, tangent2 */
  int body1, body2;              /* -1 = world */
  double mu, solref[2], solimp[5], margin, invw;
} contact_t;

typedef struct ora_env {
  mmz_model m;
  /* state */
  double qpos[NQ], qvel[ND], ctrl[MMZ_MAXACT];
  int t;
  /* kinematics */
  double xpos[NB][3], xquat[NB][4], xmat[NB][9], xipos[NB][3], ximat[NB][9];
  double xanchor[NJ][3], xaxis[NJ][3];
  double gpos[NG][3], gmat[NG][9];
  /* dynamics */
  double cdof[ND][6];
  double M[ND][ND], L[ND][ND];
  double qfrc_bias[ND], qfrc_passive[ND], qfrc_act[ND], qfrc_smooth[ND], qacc_smooth[ND], qacc[ND];
  /* constraints */
  int ncon, nefc, niter, overflow;
  contact_t con[MAXCON];
  double J[MAXEFC][ND], efc_pos[MAXEFC], efc_margin[MAXEFC], efc_D[MAXEFC], efc_aref[MAXEFC], efc_force[MAXEFC];
  /* options */
  int use_warmstart;
} ora_env;

/* ------------------------------------------------------------------ small math */
static void cross3(double* r, const double* a, const double* b) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
static double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static double norm3(const double* a) { return sqrt(dot3(a, a)); }
static void quat_mul(double* r, const double* a, const double* b) {
  double w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  double x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  double y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  double z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
static void quat_norm(double* q) {
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
  for (int i = 0; i < 4; i++) q[i] /= n;
}
static void quat2mat(double* R, const double* q) { /* row-major 3x3 */
  double w = q[0], x = q[1], y = q[2], z = q[3];
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - w * z); R[2] = 2 * (x * z + w * y);
  R[3] = 2 * (x * y + w * z); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - w * x);
  R[6] = 2 * (x * z - w * y); R[7] = 2 * (y * z + w * x); R[8] = 1 - 2 * (x * x + y * y);
}
static void axisangle2quat(double* q, const double* axis, double ang) {
  double s = sin(0.5 * ang);
  q[0] = cos(0.5 * ang); q[1] = s * axis[0]; q[2] = s * axis[1]; q[3] = s * axis[2];
}
static void mat_vec(double* r, const double* R, const double* v) { /* r = R v */
  double x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
  double y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
  double z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static void matT_vec(double* r, const double* R, const double* v) { /* r = R^T v */
  double x = R[0] * v[0] + R[3] * v[1] + R[6] * v[2];
  double y = R[1] * v[0] + R[4] * v[1] + R[7] * v[2];
  double z = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}

/* ------------------------------------------------------------------ kinematics (mj_kinematics [EXT]) */
static void kinematics(ora_env* e) {
  const mmz_model* m = &e->m;
  for (int b = 0; b < m->nbody; b++) {
    int p = m->body_parent[b];
    double pos[3], quat[4], R[9];
    if (p < 0) {
      for (int k = 0; k < 3; k++) pos[k] = m->body_pos[b][k];
      for (int k = 0; k < 4; k++) quat[k] = m->body_quat[b][k];
    } else {
      mat_vec(pos, e->xmat[p], m->body_pos[b]);
      for (int k = 0; k < 3; k++) pos[k] += e->xpos[p][k];
      quat_mul(quat, e->xquat[p], m->body_quat[b]);
    }
    for (int j = m->body_jntadr[b]; j < m->body_jntadr[b] + m->body_jntnum[b]; j++) {
      int qa = m->jnt_qadr[j];
      if (m->jnt_type[j] == MMZ_JNT_FREE) {
        quat_norm(e->qpos + qa + 3); /* MuJoCo normalises the stored quaternion in place */
        for (int k = 0; k < 3; k++) pos[k] = e->qpos[qa + k];
        for (int k = 0; k < 4; k++) quat[k] = e->qpos[qa + 3 + k];
        for (int k = 0; k < 3; k++) { e->xanchor[j][k] = pos[k]; e->xaxis[j][k] = (k == 2); }
        continue;
      }
      quat2mat(R, quat);
      mat_vec(e->xanchor[j], R, m->jnt_pos[j]);
      for (int k = 0; k < 3; k++) e->xanchor[j][k] += pos[k];
      mat_vec(e->xaxis[j], R, m->jnt_axis[j]);
      double dq = e->qpos[qa] - m->qpos0[qa];
      if (m->jnt_type[j] == MMZ_JNT_SLIDE) {
        for (int k = 0; k < 3; k++) pos[k] += e->xaxis[j][k] * dq;
      } else { /* hinge: rotate about the anchor */
        double qr[4], q2[4], off[3];
        axisangle2quat(qr, m->jnt_axis[j], dq);
        quat_mul(q2, quat, qr);
        memcpy(quat, q2, sizeof q2);
        quat2mat(R, quat);
        mat_vec(off, R, m->jnt_pos[j]);
        for (int k = 0; k < 3; k++) pos[k] = e->xanchor[j][k] - off[k];
      }
    }
    quat_norm(quat);
    memcpy(e->xpos[b], pos, sizeof pos);
    memcpy(e->xquat[b], quat, sizeof quat);
    quat2mat(e->xmat[b], quat);
    mat_vec(e->xipos[b], e->xmat[b], m->body_ipos[b]);
    for (int k = 0; k < 3; k++) e->xipos[b][k] += pos[k];
    double qi[4];
    quat_mul(qi, quat, m->body_iquat[b]);
    quat2mat(e->ximat[b], qi);
  }
  for (int g = 0; g < m->ngeom; g++) {
    int b = m->geom_body[g];
    double q[4];
    mat_vec(e->gpos[g], e->xmat[b], m->geom_pos[g]);
    for (int k = 0; k < 3; k++) e->gpos[g][k] += e->xpos[b][k];
    quat_mul(q, e->xquat[b], m->geom_quat[g]);
    quat2mat(e->gmat[g], q);
  }
}

/* ------------------------------------------------------------------ spatial algebra about the world origin
 * motion vectors [angular(3); linear velocity of the point at the origin(3)],
 * force vectors  [torque about the origin(3); force(3)], all in world axes. */
static void motion_axes(ora_env* e) {
  const mmz_model* m = &e->m;
  for (int j = 0; j < m->njnt; j++) {
    int d = m->jnt_dadr[j], b = m->jnt_body[j];
    double* c = e->cdof[d];
    if (m->jnt_type[j] == MMZ_JNT_FREE) {
      for (int k = 0; k < 3; k++) { /* translation along world axes */
        for (int i = 0; i < 6; i++) e->cdof[d + k][i] = 0;
        e->cdof[d + k][3 + k] = 1;
      }
      for (int k = 0; k < 3; k++) { /* rotation about the body axes through the body origin */
        double ax[3] = {e->xmat[b][k], e->xmat[b][3 + k], e->xmat[b][6 + k]};
        double* r = e->cdof[d + 3 + k];
        r[0] = ax[0]; r[1] = ax[1]; r[2] = ax[2];
        cross3(r + 3, e->xpos[b], ax);
      }
    } else if (m->jnt_type[j] == MMZ_JNT_SLIDE) {
      c[0] = c[1] = c[2] = 0;
      for (int k = 0; k < 3; k++) c[3 + k] = e->xaxis[j][k];
    } else {
      for (int k = 0; k < 3; k++) c[k] = e->xaxis[j][k];
      cross3(c + 3, e->xanchor[j], e->xaxis[j]);
    }
  }
}
/* spatial inertia about the origin: I[0..5] = rotational (xx,yy,zz,xy,xz,yz), I[6..8] = m*com, I[9] = m */
static void body_inertia_world(const ora_env* e, int b, double* I) {
  const mmz_model* m = &e->m;
  const double* R = e->ximat[b];
  const double* d = m->body_inertia[b];
  double mass = m->body_mass[b];
  const double* c = e->xipos[b];
  double Ic[6];
  Ic[0] = R[0] * R[0] * d[0] + R[1] * R[1] * d[1] + R[2] * R[2] * d[2];
  Ic[1] = R[3] * R[3] * d[0] + R[4] * R[4] * d[1] + R[5] * R[5] * d[2];
  Ic[2] = R[6] * R[6] * d[0] + R[7] * R[7] * d[1] + R[8] * R[8] * d[2];
  Ic[3] = R[0] * R[3] * d[0] + R[1] * R[4] * d[1] + R[2] * R[5] * d[2];
  Ic[4] = R[0] * R[6] * d[0] + R[1] * R[7] * d[1] + R[2] * R[8] * d[2];
  Ic[5] = R[3] * R[6] * d[0] + R[4] * R[7] * d[1] + R[5] * R[8] * d[2];
  double cc = dot3(c, c);
  I[0] = Ic[0] + mass * (cc - c[0] * c[0]);
  I[1] = Ic[1] + mass * (cc - c[1] * c[1]);
  I[2] = Ic[2] + mass * (cc - c[2] * c[2]);
  I[3] = Ic[3] - mass * c[0] * c[1];
  I[4] = Ic[4] - mass * c[0] * c[2];
  I[5] = Ic[5] - mass * c[1] * c[2];
  I[6] = mass * c[0]; I[7] = mass * c[1]; I[8] = mass * c[2];
  I[9] = mass;
}
static void inert_mul(double* f, const double* I, const double* v) { /* f = I v */
  const double *w = v, *l = v + 3, *h = I + 6;
  double hxl[3], hxw[3];
  cross3(hxl, h, l);
  cross3(hxw, h, w);
  f[0] = I[0] * w[0] + I[3] * w[1] + I[4] * w[2] + hxl[0];
  f[1] = I[3] * w[0] + I[1] * w[1] + I[5] * w[2] + hxl[1];
  f[2] = I[4] * w[0] + I[5] * w[1] + I[2] * w[2] + hxl[2];
  for (int k = 0; k < 3; k++) f[3 + k] = I[9] * l[k] - hxw[k];
}
static void cross_motion(double* r, const double* v, const double* s) { /* v x s */
  double a[3], b[3], c[3];
  cross3(a, v, s);
  cross3(b, v, s + 3);
  cross3(c, v + 3, s);
  for (int k = 0; k < 3; k++) { r[k] = a[k]; r[3 + k] = b[k] + c[k]; }
}
static void cross_force(double* r, const double* v, const double* f) { /* v x* f */
  double a[3], b[3], c[3];
  cross3(a, v, f);
  cross3(b, v + 3, f + 3);
  cross3(c, v, f + 3);
  for (int k = 0; k < 3; k++) { r[k] = a[k] + b[k]; r[3 + k] = c[k]; }
}
static double dot6(const double* a, const double* b) {
  double s = 0;
  for (int k = 0; k < 6; k++) s += a[k] * b[k];
  return s;
}

/* composite rigid body -> M (mj_crb [EXT]) */
static void mass_matrix(ora_env* e) {
  const mmz_model* m = &e->m;
  double Ic[NB][10];
  for (int b = 0; b < m->nbody; b++) body_inertia_world(e, b, Ic[b]);
  for (int b = m->nbody - 1; b >= 0; b--) {
    int p = m->body_parent[b];
    if (p >= 0) for (int k = 0; k < 10; k++) Ic[p][k] += Ic[b][k];
  }
  for (int i = 0; i < m->nv; i++) for (int j = 0; j < m->nv; j++) e->M[i][j] = 0;
  for (int i = 0; i < m->nv; i++) {
    double f[6];
    inert_mul(f, Ic[m->dof_body[i]], e->cdof[i]);
    for (int j = i; j >= 0; j = m->dof_parent[j]) {
      double v = dot6(e->cdof[j], f);
      e->M[i][j] = e->M[j][i] = v;
    }
    e->M[i][i] += m->dof_armature[i];
  }
}

/* bias forces c(q, qvel) incl. gravity (mj_rne [EXT]) */
static void bias_forces(ora_env* e, double vel[NB][6]) {
  const mmz_model* m = &e->m;
  double acc[NB][6], frc[NB][6];
  for (int b = 0; b < m->nbody; b++) {
    int p = m->body_parent[b];
    double v[6], a[6];
    if (p < 0) {
      for (int k = 0; k < 6; k++) v[k] = 0, a[k] = 0;
      for (int k = 0; k < 3; k++) a[3 + k] = -m->gravity[k]; /* gravity as base acceleration */
    } else {
      memcpy(v, vel[p], sizeof v);
      memcpy(a, acc[p], sizeof a);
    }
    for (int j = m->body_jntadr[b]; j < m->body_jntadr[b] + m->body_jntnum[b]; j++) {
      int d = m->jnt_dadr[j];
      if (m->jnt_type[j] == MMZ_JNT_FREE) {
        /* world-aligned translation axes are constant: no axis-derivative term */
        for (int k = 0; k < 3; k++) for (int i = 0; i < 6; i++) v[i] += e->cdof[d + k][i] * e->qvel[d + k];
        double sd[3][6];
        for (int k = 0; k < 3; k++) cross_motion(sd[k], v, e->cdof[d + 3 + k]);
        for (int k = 0; k < 3; k++)
          for (int i = 0; i < 6; i++) {
            a[i] += sd[k][i] * e->qvel[d + 3 + k];
            v[i] += e->cdof[d + 3 + k][i] * e->qvel[d + 3 + k];
          }
      } else {
        double sd[6];
        cross_motion(sd, v, e->cdof[d]);
        for (int i = 0; i < 6; i++) {
          a[i] += sd[i] * e->qvel[d];
          v[i] += e->cdof[d][i] * e->qvel[d];
        }
      }
    }
    memcpy(vel[b], v, sizeof v);
    memcpy(acc[b], a, sizeof a);
    double I[10], Ia[6], Iv[6], vxIv[6];
    body_inertia_world(e, b, I);
    inert_mul(Ia, I, a);
    inert_mul(Iv, I, v);
    cross_force(vxIv, v, Iv);
    for (int k = 0; k < 6; k++) frc[b][k] = Ia[k] + vxIv[k];
  }
  for (int b = m->nbody - 1; b >= 0; b--) {
    int p = m->body_parent[b];
    if (p >= 0) for (int k = 0; k < 6; k++) frc[p][k] += frc[b][k];
  }
  for (int d = 0; d < m->nv; d++) e->qfrc_bias[d] = dot6(e->cdof[d], frc[m->dof_body[d]]);
}

/* damping + fluid forces (mj_passive [EXT]) */
static void passive_forces(ora_env* e, double vel[NB][6]) {
  const mmz_model* m = &e->m;
  for (int d = 0; d < m->nv; d++) e->qfrc_passive[d] = -m->dof_damping[d] * e->qvel[d];
  if (m->density <= 0 && m->viscosity <= 0) return;
  for (int b = 0; b < m->nbody; b++) {
    double mass = m->body_mass[b];
    if (mass < MINVAL) continue;
    const double* I = m->body_inertia[b];
    double box[3] = {sqrt(fmax(MINVAL, I[1] + I[2] - I[0]) / mass * 6.0),
                     sqrt(fmax(MINVAL, I[0] + I[2] - I[1]) / mass * 6.0),
                     sqrt(fmax(MINVAL, I[0] + I[1] - I[2]) / mass * 6.0)};
    /* velocity of the body com in the inertial (principal) frame */
    double wxc[3], vc[3], lw[3], lv[3], lf[6];
    cross3(wxc, vel[b], e->xipos[b]);
    for (int k = 0; k < 3; k++) vc[k] = vel[b][3 + k] + wxc[k];
    matT_vec(lw, e->ximat[b], vel[b]);
    matT_vec(lv, e->ximat[b], vc);
    for (int k = 0; k < 6; k++) lf[k] = 0;
    if (m->viscosity > 0) {
      double diam = (box[0] + box[1] + box[2]) / 3.0;
      for (int k = 0; k < 3; k++) {
        lf[k] = -PI * diam * diam * diam * m->viscosity * lw[k];
        lf[3 + k] = -3.0 * PI * diam * m->viscosity * lv[k];
      }
    }
    if (m->density > 0) {
      lf[3] -= 0.5 * m->density * box[1] * box[2] * fabs(lv[0]) * lv[0];
      lf[4] -= 0.5 * m->density * box[0] * box[2] * fabs(lv[1]) * lv[1];
      lf[5] -= 0.5 * m->density * box[0] * box[1] * fabs(lv[2]) * lv[2];
      lf[0] -= m->density * box[0] * (pow(box[1], 4) + pow(box[2], 4)) * fabs(lw[0]) * lw[0] / 64.0;
      lf[1] -= m->density * box[1] * (pow(box[0], 4) + pow(box[2], 4)) * fabs(lw[1]) * lw[1] / 64.0;
      lf[2] -= m->density * box[2] * (pow(box[0], 4) + pow(box[1], 4)) * fabs(lw[2]) * lw[2] / 64.0;
    }
    double tq[3], fc[3], cxf[3], sf[6];
    mat_vec(tq, e->ximat[b], lf);
    mat_vec(fc, e->ximat[b], lf + 3);
    cross3(cxf, e->xipos[b], fc);
    for (int k = 0; k < 3; k++) { sf[k] = tq[k] + cxf[k]; sf[3 + k] = fc[k]; }
    for (int d = 0; d < m->nv; d++)
      if (m->body_dofmask[b] >> d & 1) e->qfrc_passive[d] += dot6(e->cdof[d], sf);
  }
}

/* ------------------------------------------------------------------ dense Cholesky helpers */
static int chol(int n, double A[ND][ND], double Lo[ND][ND]) {
  for (int j = 0; j < n; j++) {
    double s = A[j][j];
    for (int k = 0; k < j; k++) s -= Lo[j][k] * Lo[j][k];
    if (s < MINVAL) s = MINVAL;
    Lo[j][j] = sqrt(s);
    for (int i = j + 1; i < n; i++) {
      double t = A[i][j];
      for (int k = 0; k < j; k++) t -= Lo[i][k] * Lo[j][k];
      Lo[i][j] = t / Lo[j][j];
    }
  }
  return 0;
}
static void chol_solve(int n, double Lo[ND][ND], double* x) { /* in place */
  for (int i = 0; i < n; i++) {
    double s = x[i];
    for (int k = 0; k < i; k++) s -= Lo[i][k] * x[k];
    x[i] = s / Lo[i][i];
  }
  for (int i = n - 1; i >= 0; i--) {
    double s = x[i];
    for (int k = i + 1; k < n; k++) s -= Lo[k][i] * x[k];
    x[i] = s / Lo[i][i];
  }
}

/* ------------------------------------------------------------------ narrow phase
 * Convention (MuJoCo [EXT]): normal points from geom1 to geom2, dist < 0 when
 * penetrating, position midway between the two surfaces. */
static void make_frame(double* fr) { /* fr[0..2] normal given; fr[3..5] optional hint (mju_makeFrame [EXT]) */
  double n = norm3(fr);
  for (int k = 0; k < 3; k++) fr[k] /= n;
  if (norm3(fr + 3) < 0.5) {
    fr[3] = fr[4] = fr[5] = 0;
    if (fr[1] < 0.5 && fr[1] > -0.5) fr[4] = 1; else fr[5] = 1;
  }
  double d = dot3(fr, fr + 3);
  for (int k = 0; k < 3; k++) fr[3 + k] -= d * fr[k];
  n = norm3(fr + 3);
  for (int k = 0; k < 3; k++) fr[3 + k] /= n;
  cross3(fr + 6, fr, fr + 3);
}

typedef struct { double dist, pos[3], normal[3], hint[3]; } raw_contact;

/* sphere (centre c, radius r) against a box (centre bc, rotation bR, half extents h) */
static int sphere_box(const double* c, double r, const double* bc, const double* bR, const double* h,
                      double margin, raw_contact* out) {
  double rel[3], loc[3], cl[3], dl[3];
  for (int k = 0; k < 3; k++) rel[k] = c[k] - bc[k];
  matT_vec(loc, bR, rel);
  int inside = 1;
  for (int k = 0; k < 3; k++) {
    cl[k] = fmin(fmax(loc[k], -h[k]), h[k]);
    if (cl[k] != loc[k]) inside = 0;
    dl[k] = cl[k] - loc[k];
  }
  double nl[3] = {0, 0, 0}, pl[3], dist;
  if (inside) { /* centre inside the box: leave through the nearest face */
    double best = 2 * (h[0] + h[1] + h[2]);
    int bi = 0;
    for (int i = 0; i < 6; i++) {
      double cd = fabs((i % 2 ? 1.0 : -1.0) * h[i / 2] - loc[i / 2]);
      if (cd < best) { best = cd; bi = i; }
    }
    nl[bi / 2] = (bi % 2) ? -1.0 : 1.0;
    dist = -best - r;
    for (int k = 0; k < 3; k++) pl[k] = loc[k] + nl[k] * (r - best) * 0.5;
  } else {
    double d = norm3(dl);
    if (d - r > margin) return 0;
    for (int k = 0; k < 3; k++) nl[k] = dl[k] / d;
    dist = d - r;
    for (int k = 0; k < 3; k++) pl[k] = loc[k] + nl[k] * (r + dist * 0.5);
  }
  if (dist > margin) return 0;
  out->dist = dist;
  mat_vec(out->normal, bR, nl);
  mat_vec(out->pos, bR, pl);
  for (int k = 0; k < 3; k++) out->pos[k] += bc[k];
  out->hint[0] = out->hint[1] = out->hint[2] = 0;
  return 1;
}

/* squared distance from a point (box-local) to the box, and d/dt along direction b */
static double box_excess_deriv(const double* a, const double* b, double t, const double* h) {
  double g = 0;
  for (int k = 0; k < 3; k++) {
    double x = a[k] + b[k] * t;
    double ex = x > h[k] ? x - h[k] : (x < -h[k] ? x + h[k] : 0.0);
    g += ex * b[k];
  }
  return g;
}
/* capsule (segment p0-p1, radius r) against a box: 2 end contacts when both ends are within the
 * margin, otherwise 1 contact at the segment point nearest the box. */
static int capsule_box(const double* p0, const double* p1, double r, const double* bc, const double* bR,
                       const double* h, double margin, raw_contact* out) {
  raw_contact c0, c1;
  int n0 = sphere_box(p0, r, bc, bR, h, margin, &c0);
  int n1 = sphere_box(p1, r, bc, bR, h, margin, &c1);
  if (n0 && n1) { out[0] = c0; out[1] = c1; return 2; }
  /* nearest point: root of the piecewise-linear derivative of the squared distance */
  double a[3], b[3], rel[3];
  for (int k = 0; k < 3; k++) rel[k] = p0[k] - bc[k];
  matT_vec(a, bR, rel);
  for (int k = 0; k < 3; k++) rel[k] = p1[k] - p0[k];
  matT_vec(b, bR, rel);
  double cand[8];
  int nc = 0;
  cand[nc++] = 0; cand[nc++] = 1;
  for (int k = 0; k < 3; k++)
    if (fabs(b[k]) > MINVAL)
      for (int s = -1; s <= 1; s += 2) {
        double t = (s * h[k] - a[k]) / b[k];
        if (t > 0 && t < 1) cand[nc++] = t;
      }
  double tlo = 0, thi = 1, glo = box_excess_deriv(a, b, 0, h), ghi = box_excess_deriv(a, b, 1, h), ts;
  if (glo > 0) ts = 0;
  else if (ghi <= 0) ts = 1;
  else {
    for (int i = 2; i < nc; i++) {
      double g = box_excess_deriv(a, b, cand[i], h);
      if (g <= 0) { if (cand[i] > tlo) { tlo = cand[i]; glo = g; } }
      else if (cand[i] < thi) { thi = cand[i]; ghi = g; }
    }
    ts = (ghi - glo) > MINVAL ? tlo + (-glo) * (thi - tlo) / (ghi - glo) : tlo;
  }
  double ps[3];
  for (int k = 0; k < 3; k++) ps[k] = p0[k] + ts * (p1[k] - p0[k]);
  return sphere_box(ps, r, bc, bR, h, margin, out);
}

/* box against box: separating-axis search, then face clipping or an edge-edge point.
 * Normal points from box A to box B. Up to 8 contacts. */
/* sphere against sphere [EXT: mjc_SphereSphere]: normal from sphere 1 to sphere 2, contact midway between the surfaces */
static int sphere_sphere(const double* c1, double r1, const double* c2, double r2, double margin, raw_contact* out) {
  double d[3] = {c2[0] - c1[0], c2[1] - c1[1], c2[2] - c1[2]};
  double len = norm3(d), dist = len - r1 - r2;
  if (dist > margin) return 0;
  double n[3] = {0, 0, 1};
  if (len > MINVAL) for (int k = 0; k < 3; k++) n[k] = d[k] / len;
  out->dist = dist;
  for (int k = 0; k < 3; k++) { out->normal[k] = n[k]; out->pos[k] = c1[k] + n[k] * (r1 + 0.5 * dist); out->hint[k] = 0; }
  return 1;
}
/* sphere against capsule [EXT: mjc_SphereCapsule]: the sphere against the capsule-radius sphere at the segment point nearest to it */
static int sphere_capsule(const double* c, double rs, const double* p0, const double* p1, double rc, double margin, raw_contact* out) {
  double ab[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]}, ac[3] = {c[0] - p0[0], c[1] - p0[1], c[2] - p0[2]};
  double den = dot3(ab, ab), t = den > MINVAL ? dot3(ac, ab) / den : 0.0;
  t = t < 0 ? 0 : (t > 1 ? 1 : t);
  double q[3] = {p0[0] + t * ab[0], p0[1] + t * ab[1], p0[2] + t * ab[2]};
  return sphere_sphere(c, rs, q, rc, margin, out);
}

static int box_box(const double* ca, const double* Ra, const double* ha, const double* cb, const double* Rb,
                   const double* hb, double margin, raw_contact* out) {
  double A[3][3], B[3][3], d[3];
  for (int i = 0; i < 3; i++) for (int k = 0; k < 3; k++) { A[i][k] = Ra[3 * k + i]; B[i][k] = Rb[3 * k + i]; }
  for (int k = 0; k < 3; k++) d[k] = cb[k] - ca[k];
  /* least-penetration axis, faces and edge pairs tracked separately */
  double best_f = -1e300, best_e = -1e300, nf[3] = {0, 0, 0}, ne[3] = {0, 0, 0};
  int code_f = -1, code_e = -1;
  for (int code = 0; code < 15; code++) {
    double L[3];
    if (code < 3) memcpy(L, A[code], sizeof L);
    else if (code < 6) memcpy(L, B[code - 3], sizeof L);
    else {
      cross3(L, A[(code - 6) / 3], B[(code - 6) % 3]);
      double n = norm3(L);
      if (n < 1e-6) continue; /* parallel edges: covered by the face axes */
      for (int k = 0; k < 3; k++) L[k] /= n;
    }
    double ra = 0, rb = 0;
    for (int i = 0; i < 3; i++) { ra += ha[i] * fabs(dot3(L, A[i])); rb += hb[i] * fabs(dot3(L, B[i])); }
    double dl = dot3(L, d);
    double s = fabs(dl) - ra - rb;
    if (s > margin) return 0;
    double sg = dl < 0 ? -1.0 : 1.0;
    if (code < 6) {
      if (s > best_f) { best_f = s; code_f = code; for (int k = 0; k < 3; k++) nf[k] = sg * L[k]; }
    } else if (s > best_e) { best_e = s; code_e = code; for (int k = 0; k < 3; k++) ne[k] = sg * L[k]; }
  }
  /* an edge pair wins only when clearly better than the best face (keeps resting contacts on faces) */
  double best, bn[3];
  int bcode;
  if (code_e >= 0 && best_e > best_f + 1e-6 + 0.05 * fabs(best_f)) { best = best_e; bcode = code_e; memcpy(bn, ne, sizeof bn); }
  else { best = best_f; bcode = code_f; memcpy(bn, nf, sizeof bn); }
  if (bcode < 0) return 0;
  if (bcode >= 6) { /* edge-edge */
    int ia = (bcode - 6) / 3, ib = (bcode - 6) % 3;
    double pa[3], pb[3];
    memcpy(pa, ca, sizeof pa);
    memcpy(pb, cb, sizeof pb);
    for (int i = 0; i < 3; i++) {
      if (i != ia) { double sg = dot3(bn, A[i]) > 0 ? 1.0 : -1.0; for (int k = 0; k < 3; k++) pa[k] += sg * ha[i] * A[i][k]; }
      if (i != ib) { double sg = dot3(bn, B[i]) > 0 ? -1.0 : 1.0; for (int k = 0; k < 3; k++) pb[k] += sg * hb[i] * B[i][k]; }
    }
    /* closest points of the two edge lines */
    double w[3], ua = 0, ub = 0;
    for (int k = 0; k < 3; k++) w[k] = pb[k] - pa[k];
    double uaub = dot3(A[ia], B[ib]), q1 = dot3(A[ia], w), q2 = -dot3(B[ib], w), den = 1 - uaub * uaub;
    if (den > 1e-9) { ua = (q1 + uaub * q2) / den; ub = (uaub * q1 + q2) / den; }
    ua = fmin(fmax(ua, -ha[ia]), ha[ia]);
    ub = fmin(fmax(ub, -hb[ib]), hb[ib]);
    for (int k = 0; k < 3; k++) {
      double xa = pa[k] + ua * A[ia][k], xb = pb[k] + ub * B[ib][k];
      out->pos[k] = 0.5 * (xa + xb);
      out->normal[k] = bn[k];
      out->hint[k] = 0;
    }
    out->dist = best;
    return 1;
  }
  /* face contact: reference box owns the axis, incident box is the other */
  const double *cr, *hr, *ci, *hi;
  double (*Rr)[3], (*Ri)[3];
  double nr[3]; /* outward normal of the reference face */
  int ax, ref_is_a = bcode < 3;
  if (ref_is_a) { cr = ca; hr = ha; Rr = A; ci = cb; hi = hb; Ri = B; ax = bcode; memcpy(nr, bn, sizeof nr); }
  else { cr = cb; hr = hb; Rr = B; ci = ca; hi = ha; Ri = A; ax = bcode - 3; for (int k = 0; k < 3; k++) nr[k] = -bn[k]; }
  /* incident face: most anti-parallel to nr */
  int iax = 0;
  double mind = 1e300;
  for (int i = 0; i < 3; i++) {
    double v = fabs(dot3(nr, Ri[i]));
    if (-v < mind) { mind = -v; iax = i; }
  }
  double isg = dot3(nr, Ri[iax]) > 0 ? -1.0 : 1.0;
  int u = (iax + 1) % 3, v = (iax + 2) % 3;
  double poly[16][3], tmp[16][3];
  int np = 4;
  for (int c = 0; c < 4; c++) {
    double su = (c == 0 || c == 3) ? -1.0 : 1.0, sv = (c < 2) ? -1.0 : 1.0;
    for (int k = 0; k < 3; k++)
      poly[c][k] = ci[k] + isg * hi[iax] * Ri[iax][k] + su * hi[u] * Ri[u][k] + sv * hi[v] * Ri[v][k];
  }
  /* clip against the four side planes of the reference face */
  int ru = (ax + 1) % 3, rv = (ax + 2) % 3;
  for (int side = 0; side < 4; side++) {
    const double* axs = Rr[side < 2 ? ru : rv];
    double sg = (side % 2) ? -1.0 : 1.0, lim = hr[side < 2 ? ru : rv];
    int nn = 0;
    for (int i = 0; i < np; i++) {
      double *p = poly[i], *q = poly[(i + 1) % np];
      double rp[3], rq[3];
      for (int k = 0; k < 3; k++) { rp[k] = p[k] - cr[k]; rq[k] = q[k] - cr[k]; }
      double dp = sg * dot3(axs, rp) - lim, dq = sg * dot3(axs, rq) - lim;
      if (dp <= 0) { memcpy(tmp[nn++], p, sizeof(double) * 3); }
      if ((dp <= 0) != (dq <= 0)) {
        double t = dp / (dp - dq);
        for (int k = 0; k < 3; k++) tmp[nn][k] = p[k] + t * (q[k] - p[k]);
        nn++;
      }
    }
    np = nn;
    memcpy(poly, tmp, sizeof(double) * 3 * np);
    if (np == 0) return 0;
  }
  int n = 0;
  for (int i = 0; i < np && n < 8; i++) {
    double rp[3];
    for (int k = 0; k < 3; k++) rp[k] = poly[i][k] - cr[k];
    double depth = dot3(nr, rp) - hr[ax]; /* signed distance of the incident point to the reference face */
    if (depth >= margin) continue;
    for (int k = 0; k < 3; k++) {
      out[n].pos[k] = poly[i][k] - nr[k] * depth * 0.5;
      out[n].normal[k] = bn[k];
      out[n].hint[k] = 0;
    }
    out[n].dist = depth;
    n++;
  }
  return n;
}

/* ------------------------------------------------------------------ collision driver (mj_collision [EXT]) */
typedef struct { double margin, friction[3], solref[2], solimp[5]; } geom_par;
static void mix_par(const geom_par* a, const geom_par* b, contact_t* c) {
  c->margin = fmax(a->margin, b->margin);
  c->mu = fmax(a->friction[0], b->friction[0]);
  for (int k = 0; k < 2; k++) c->solref[k] = 0.5 * (a->solref[k] + b->solref[k]);
  for (int k = 0; k < 5; k++) c->solimp[k] = 0.5 * (a->solimp[k] + b->solimp[k]);
}
static void geom_params(const mmz_model* m, int g, geom_par* p) {
  p->margin = m->geom_margin[g];
  memcpy(p->friction, m->geom_friction[g], sizeof p->friction);
  memcpy(p->solref, m->geom_solref[g], sizeof p->solref);
  memcpy(p->solimp, m->geom_solimp[g], sizeof p->solimp);
}
static void add_contacts(ora_env* e, int n, const raw_contact* rc, int b1, int b2, double invw, const contact_t* proto) {
  for (int i = 0; i < n; i++) {
    if (e->ncon >= MAXCON) { e->overflow = 1; return; }
    contact_t* c = &e->con[e->ncon++];
    *c = *proto;
    c->dist = rc[i].dist;
    memcpy(c->pos, rc[i].pos, sizeof c->pos);
    memcpy(c->frame, rc[i].normal, sizeof(double) * 3);
    memcpy(c->frame + 3, rc[i].hint, sizeof(double) * 3);
    make_frame(c->frame);
    c->body1 = b1; c->body2 = b2; c->invw = invw;
  }
}
static void capsule_ends(const ora_env* e, int g, double* p0, double* p1) {
  const mmz_model* m = &e->m;
  double ax[3] = {e->gmat[g][2], e->gmat[g][5], e->gmat[g][8]};
  for (int k = 0; k < 3; k++) {
    p0[k] = e->gpos[g][k] + ax[k] * m->geom_size[g][1];
    p1[k] = e->gpos[g][k] - ax[k] * m->geom_size[g][1];
  }
}
static const double IDENT[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};

static void collide_static_box(ora_env* e, int g, const double* bc, const double* bh, const geom_par* wp) {
  const mmz_model* m = &e->m;
  geom_par gp;
  contact_t proto;
  raw_contact rc[8];
  geom_params(m, g, &gp);
  mix_par(&gp, wp, &proto);
  int n = 0, b = m->geom_body[g], t = m->geom_type[g];
  if (t == MMZ_GEOM_SPHERE) {
    n = sphere_box(e->gpos[g], m->geom_size[g][0], bc, IDENT, bh, proto.margin, rc);
    add_contacts(e, n, rc, b, -1, m->geom_invweight[g], &proto); /* geom1 = sphere (robot), geom2 = wall */
  } else if (t == MMZ_GEOM_CAPSULE) {
    double p0[3], p1[3];
    capsule_ends(e, g, p0, p1);
    n = capsule_box(p0, p1, m->geom_size[g][0], bc, IDENT, bh, proto.margin, rc);
    add_contacts(e, n, rc, b, -1, m->geom_invweight[g], &proto);
  } else if (t == MMZ_GEOM_BOX) { /* geom1 = wall (lower geom id in MuJoCo's ordering), geom2 = robot box */
    n = box_box(bc, IDENT, bh, e->gpos[g], e->gmat[g], m->geom_size[g], proto.margin, rc);
    add_contacts(e, n, rc, -1, b, m->geom_invweight[g], &proto);
  }
}

static void collision(ora_env* e) {
  const mmz_model* m = &e->m;
  e->ncon = 0;
  e->overflow = 0;
  if (!m->collision_on) return;
  geom_par fp, wp;
  fp.margin = m->floor_margin; memcpy(fp.friction, m->floor_friction, sizeof fp.friction);
  memcpy(fp.solref, m->floor_solref, sizeof fp.solref); memcpy(fp.solimp, m->floor_solimp, sizeof fp.solimp);
  wp.margin = m->wall_margin; memcpy(wp.friction, m->wall_friction, sizeof wp.friction);
  memcpy(wp.solref, m->wall_solref, sizeof wp.solref); memcpy(wp.solimp, m->wall_solimp, sizeof wp.solimp);
  for (int g = 0; g < m->ngeom; g++) {
    int b = m->geom_body[g], t = m->geom_type[g];
    /* world geoms all have contype = conaffinity = 1 (reference maze_env.py:149-150, assets floor) */
    if (!((m->geom_contype[g] & 1) || (m->geom_conaffinity[g] & 1))) continue;
    geom_par gp;
    contact_t proto;
    raw_contact rc[8];
    geom_params(m, g, &gp);
    /* --- floor plane (normal +z); geom1 = plane */
    if (m->has_floor) {
      mix_par(&gp, &fp, &proto);
      int n = 0;
      if (t == MMZ_GEOM_SPHERE) {
        double r = m->geom_size[g][0], dist = e->gpos[g][2] - m->floor_z - r;
        if (dist < proto.margin) {
          rc[0].dist = dist;
          for (int k = 0; k < 3; k++) { rc[0].pos[k] = e->gpos[g][k]; rc[0].normal[k] = (k == 2); rc[0].hint[k] = 0; }
          rc[0].pos[2] -= r + 0.5 * dist;
          n = 1;
        }
      } else if (t == MMZ_GEOM_CAPSULE) {
        double p[2][3], r = m->geom_size[g][0];
        capsule_ends(e, g, p[0], p[1]);
        for (int s = 0; s < 2; s++) {
          double dist = p[s][2] - m->floor_z - r;
          if (dist < proto.margin) {
            rc[n].dist = dist;
            for (int k = 0; k < 3; k++) { rc[n].pos[k] = p[s][k]; rc[n].normal[k] = (k == 2); }
            rc[n].pos[2] -= r + 0.5 * dist;
            /* MuJoCo aligns the first tangent with the capsule axis */
            rc[n].hint[0] = e->gmat[g][2]; rc[n].hint[1] = e->gmat[g][5]; rc[n].hint[2] = e->gmat[g][8];
            if (fabs(rc[n].hint[2]) > 0.999999) rc[n].hint[0] = rc[n].hint[1] = rc[n].hint[2] = 0; /* vertical capsule */
            n++;
          }
        }
      } else if (t == MMZ_GEOM_BOX) {
        for (int c = 0; c < 8 && n < 4; c++) {
          double loc[3] = {(c & 1 ? 1 : -1) * m->geom_size[g][0], (c & 2 ? 1 : -1) * m->geom_size[g][1],
                           (c & 4 ? 1 : -1) * m->geom_size[g][2]};
          double w[3];
          mat_vec(w, e->gmat[g], loc);
          for (int k = 0; k < 3; k++) w[k] += e->gpos[g][k];
          double dist = w[2] - m->floor_z;
          if (dist < proto.margin) {
            rc[n].dist = dist;
            for (int k = 0; k < 3; k++) { rc[n].pos[k] = w[k]; rc[n].normal[k] = (k == 2); rc[n].hint[k] = 0; }
            rc[n].pos[2] -= 0.5 * dist;
            n++;
          }
        }
      }
      add_contacts(e, n, rc, -1, b, m->geom_invweight[g], &proto);
    }
    /* --- wall / platform boxes of the maze grid */
    for (int i = 0; i < m->grid_h; i++)
      for (int j = 0; j < m->grid_w; j++) {
        int code = m->grid[i * m->grid_w + j];
        double bc[3] = {j * m->cell_size - m->origin[0], i * m->cell_size - m->origin[1], 0};
        if (code & MMZ_CELL_WALL) { bc[2] = m->wall_z; collide_static_box(e, g, bc, m->wall_half, &wp); }
        if (code & MMZ_CELL_PLATFORM) { bc[2] = m->plat_z; collide_static_box(e, g, bc, m->wall_half, &wp); }
      }
  }
  /* --- pairs of geoms on different moving bodies: anything against a box, sphere against sphere / capsule (object balls) */
  for (int g1 = 0; g1 < m->ngeom; g1++)
    for (int g2 = g1 + 1; g2 < m->ngeom; g2++) {
      int b1 = m->geom_body[g1], b2 = m->geom_body[g2];
      if (b1 == b2 || m->body_parent[b1] == b2 || m->body_parent[b2] == b1) continue;
      if (!((m->geom_contype[g1] & m->geom_conaffinity[g2]) || (m->geom_contype[g2] & m->geom_conaffinity[g1]))) continue;
      int a = g1, b = g2; /* order by type, like MuJoCo */
      if (m->geom_type[a] > m->geom_type[b]) { a = g2; b = g1; }
      if (m->geom_type[b] != MMZ_GEOM_BOX && m->geom_type[a] != MMZ_GEOM_SPHERE) continue; /* capsule-capsule: not in any asset */
      geom_par pa, pb;
      contact_t proto;
      raw_contact rc[8];
      geom_params(m, a, &pa);
      geom_params(m, b, &pb);
      mix_par(&pa, &pb, &proto);
      int n = 0;
      if (m->geom_type[a] == MMZ_GEOM_SPHERE && m->geom_type[b] == MMZ_GEOM_SPHERE)
        n = sphere_sphere(e->gpos[a], m->geom_size[a][0], e->gpos[b], m->geom_size[b][0], proto.margin, rc);
      else if (m->geom_type[a] == MMZ_GEOM_SPHERE && m->geom_type[b] == MMZ_GEOM_CAPSULE) {
        double p0[3], p1[3];
        capsule_ends(e, b, p0, p1);
        n = sphere_capsule(e->gpos[a], m->geom_size[a][0], p0, p1, m->geom_size[b][0], proto.margin, rc);
      } else if (m->geom_type[a] == MMZ_GEOM_SPHERE)
        n = sphere_box(e->gpos[a], m->geom_size[a][0], e->gpos[b], e->gmat[b], m->geom_size[b], proto.margin, rc);
      else if (m->geom_type[a] == MMZ_GEOM_CAPSULE) {
        double p0[3], p1[3];
        capsule_ends(e, a, p0, p1);
        n = capsule_box(p0, p1, m->geom_size[a][0], e->gpos[b], e->gmat[b], m->geom_size[b], proto.margin, rc);
      } else if (m->geom_type[a] == MMZ_GEOM_BOX)
        n = box_box(e->gpos[a], e->gmat[a], m->geom_size[a], e->gpos[b], e->gmat[b], m->geom_size[b], proto.margin, rc);
      add_contacts(e, n, rc, m->geom_body[a], m->geom_body[b], m->geom_invweight[a] + m->geom_invweight[b], &proto);
    }
}

/* ------------------------------------------------------------------ constraint rows (mj_makeConstraint [EXT]) */
static double impedance(const double* si, double r) {
  double d0 = fmin(fmax(si[0], 1e-4), 0.9999), d1 = fmin(fmax(si[1], 1e-4), 0.9999);
  double width = si[2], mid = si[3], power = si[4];
  if (d0 == d1 || width <= MINVAL) return 0.5 * (d0 + d1);
  double x = fabs(r) / width, y;
  if (x >= 1) return d1;
  if (x <= 0) return d0;
  if (power == 1) y = x;
  else if (x <= mid) y = pow(x, power) / pow(mid, power - 1);
  else y = 1 - pow(1 - x, power) / pow(1 - mid, power - 1);
  return d0 + y * (d1 - d0);
}
static void point_jac_row(const ora_env* e, int body, const double* p, const double* dir, double sign, double* row) {
  const mmz_model* m = &e->m;
  if (body < 0) return;
  for (int d = 0; d < m->nv; d++)
    if (m->body_dofmask[body] >> d & 1) {
      double wxp[3];
      cross3(wxp, e->cdof[d], p);
      double v[3] = {e->cdof[d][3] + wxp[0], e->cdof[d][4] + wxp[1], e->cdof[d][5] + wxp[2]};
      row[d] += sign * dot3(dir, v);
    }
}
static void finish_row(ora_env* e, int r, const double* solref, const double* solimp, double diag_approx) {
  const mmz_model* m = &e->m;
  double pos = e->efc_pos[r], margin = e->efc_margin[r];
  double tc = fmax(solref[0], 2 * m->timestep), dr = solref[1]; /* refsafe */
  double dmax = fmin(fmax(solimp[1], 1e-4), 0.9999);
  double k = 1.0 / fmax(MINVAL, dmax * dmax * tc * tc * dr * dr), bb = 2.0 / fmax(MINVAL, dmax * tc);
  double imp = impedance(solimp, pos - margin);
  double R = fmax(MINVAL, (1 - imp) * diag_approx / imp);
  double vel = 0;
  for (int d = 0; d < m->nv; d++) vel += e->J[r][d] * e->qvel[d];
  e->efc_D[r] = 1.0 / R;
  e->efc_aref[r] = -bb * vel - k * imp * (pos - margin);
}
static void make_constraints(ora_env* e) {
  const mmz_model* m = &e->m;
  int r = 0;
  /* joint limits */
  for (int j = 0; j < m->njnt; j++) {
    if (!m->jnt_limited[j]) continue;
    double q = e->qpos[m->jnt_qadr[j]];
    for (int side = 0; side < 2; side++) {
      double dist = side == 0 ? q - m->jnt_range[j][0] : m->jnt_range[j][1] - q;
      if (dist >= m->jnt_margin[j]) continue;
      for (int d = 0; d < m->nv; d++) e->J[r][d] = 0;
      e->J[r][m->jnt_dadr[j]] = side == 0 ? 1.0 : -1.0;
      e->efc_pos[r] = dist;
      e->efc_margin[r] = m->jnt_margin[j];
      finish_row(e, r, m->jnt_solref[j], m->jnt_solimp[j], m->dof_invweight0[m->jnt_dadr[j]]);
      r++;
    }
  }
  /* frictional contacts, pyramidal cone, condim 3: 4 rows each */
  for (int c = 0; c < e->ncon; c++) {
    const contact_t* cn = &e->con[c];
    double Jf[3][ND];
    for (int a = 0; a < 3; a++) {
      for (int d = 0; d < m->nv; d++) Jf[a][d] = 0;
      point_jac_row(e, cn->body2, cn->pos, cn->frame + 3 * a, 1.0, Jf[a]);
      point_jac_row(e, cn->body1, cn->pos, cn->frame + 3 * a, -1.0, Jf[a]);
    }
    int r0 = r;
    for (int k = 0; k < 4; k++) {
      const double* Jt = Jf[1 + k / 2];
      double sg = (k % 2) ? -1.0 : 1.0;
      for (int d = 0; d < m->nv; d++) e->J[r][d] = Jf[0][d] + sg * cn->mu * Jt[d];
      e->efc_pos[r] = cn->dist;
      e->efc_margin[r] = cn->margin;
      finish_row(e, r, cn->solref, cn->solimp, cn->invw * (1 + cn->mu * cn->mu));
      r++;
    }
    /* all edges of the pyramid share R = 2 mu^2 R_first */
    double Rpy = 2 * cn->mu * cn->mu / e->efc_D[r0];
    for (int k = 0; k < 4; k++) e->efc_D[r0 + k] = 1.0 / fmax(MINVAL, Rpy);
  }
  e->nefc = r;
}

/* ------------------------------------------------------------------ Newton solver on the primal problem
 *   min_a 1/2 (a - a0)^T M (a - a0) + sum_i 1/2 D_i min(0, J_i a - aref_i)^2          (mj_solNewton [EXT]) */
static double line_deriv(const ora_env* e, int n, const double* jar, const double* jv, double alpha, double g0, double h0,
                         double* hess) {
  double g = g0 + alpha * h0, h = h0;
  for (int i = 0; i < n; i++) {
    double x = jar[i] + alpha * jv[i];
    if (x < 0) { g += e->efc_D[i] * x * jv[i]; h += e->efc_D[i] * jv[i] * jv[i]; }
  }
  *hess = h;
  return g;
}
static void solve(ora_env* e) {
  const mmz_model* m = &e->m;
  int nv = m->nv, n = e->nefc;
  double a[ND];
  if (e->use_warmstart) memcpy(a, e->qacc, sizeof a); else memcpy(a, e->qacc_smooth, sizeof a);
  e->niter = 0;
  if (n == 0) { memcpy(e->qacc, e->qacc_smooth, sizeof a); return; }
  static __thread double jar[MAXEFC], jv[MAXEFC];
  for (int it = 0; it < 100; it++) {
    double grad[ND], H[ND][ND], Lh[ND][ND], Ma[ND];
    for (int i = 0; i < n; i++) {
      double s = -e->efc_aref[i];
      for (int d = 0; d < nv; d++) s += e->J[i][d] * a[d];
      jar[i] = s;
    }
    for (int d = 0; d < nv; d++) {
      double s = 0;
      for (int k = 0; k < nv; k++) s += e->M[d][k] * a[k];
      Ma[d] = s;
      grad[d] = s - e->qfrc_smooth[d];
      for (int k = 0; k < nv; k++) H[d][k] = e->M[d][k];
    }
    for (int i = 0; i < n; i++)
      if (jar[i] < 0) {
        double f = -e->efc_D[i] * jar[i];
        for (int d = 0; d < nv; d++) {
          grad[d] -= e->J[i][d] * f;
          double dj = e->efc_D[i] * e->J[i][d];
          if (dj != 0) for (int k = 0; k < nv; k++) H[d][k] += dj * e->J[i][k];
        }
      }
    /* converged when the gradient is at round-off level of the terms it is the (cancelling) sum of */
    double gn = 0, ref = 0;
    for (int d = 0; d < nv; d++) {
      double mag = fabs(Ma[d]) + fabs(e->qfrc_smooth[d]);
      for (int i = 0; i < n; i++)
        if (jar[i] < 0) mag += fabs(e->J[i][d] * e->efc_D[i] * jar[i]);
      gn += grad[d] * grad[d];
      ref += mag * mag;
    }
    if (sqrt(gn) <= 1e-14 * sqrt(ref) + 1e-300) break;
    double dir[ND];
    for (int d = 0; d < nv; d++) dir[d] = -grad[d];
    chol(nv, H, Lh);
    chol_solve(nv, Lh, dir);
    /* exact line search: root of the monotone piecewise-linear derivative */
    double g0 = 0, h0 = 0;
    for (int d = 0; d < nv; d++) {
      double md = 0;
      for (int k = 0; k < nv; k++) md += e->M[d][k] * dir[k];
      g0 += dir[d] * (Ma[d] - e->qfrc_smooth[d]);
      h0 += dir[d] * md;
    }
    for (int i = 0; i < n; i++) {
      double s = 0;
      for (int d = 0; d < nv; d++) s += e->J[i][d] * dir[d];
      jv[i] = s;
    }
    double lo = 0, hi = -1, alpha = 1, hess;
    for (int ls = 0; ls < 60; ls++) {
      double g = line_deriv(e, n, jar, jv, alpha, g0, h0, &hess);
      if (fabs(g) < 1e-15 * fmax(1.0, fabs(g0))) break;
      if (g < 0) lo = alpha; else hi = alpha;
      double next = alpha - g / hess;
      if (hi >= 0 && (next <= lo || next >= hi)) next = 0.5 * (lo + hi);
      if (next <= lo && hi < 0) next = 2 * alpha + 1e-12;
      if (next == alpha) break;
      alpha = next;
    }
    for (int d = 0; d < nv; d++) a[d] += alpha * dir[d];
    e->niter = it + 1;
  }
  memcpy(e->qacc, a, sizeof a);
  for (int i = 0; i < n; i++) {
    double s = -e->efc_aref[i];
    for (int d = 0; d < nv; d++) s += e->J[i][d] * a[d];
    e->efc_force[i] = s < 0 ? -e->efc_D[i] * s : 0;
  }
}

/* ------------------------------------------------------------------ mj_forward [EXT] */
static void forward(ora_env* e) {
  const mmz_model* m = &e->m;
  double vel[NB][6];
  kinematics(e);
  motion_axes(e);
  mass_matrix(e);
  collision(e);
  bias_forces(e, vel);
  passive_forces(e, vel);
  for (int d = 0; d < m->nv; d++) e->qfrc_act[d] = 0;
  for (int a = 0; a < m->nu; a++) {
    double c = e->ctrl[a];
    if (m->act_limited[a]) c = fmin(fmax(c, m->act_ctrlrange[a][0]), m->act_ctrlrange[a][1]);
    e->qfrc_act[m->act_dof[a]] += m->act_gear[a] * c;
  }
  for (int d = 0; d < m->nv; d++) {
    e->qfrc_smooth[d] = e->qfrc_passive[d] - e->qfrc_bias[d] + e->qfrc_act[d];
    e->qacc_smooth[d] = e->qfrc_smooth[d];
  }
  chol(m->nv, e->M, e->L);
  chol_solve(m->nv, e->L, e->qacc_smooth);
  make_constraints(e);
  solve(e);
}

/* position update on the configuration manifold (mj_integratePos [EXT]) */
static void integrate_pos(const mmz_model* m, double* qpos, const double* vel, double h) {
  for (int j = 0; j < m->njnt; j++) {
    int qa = m->jnt_qadr[j], d = m->jnt_dadr[j];
    if (m->jnt_type[j] == MMZ_JNT_FREE) {
      for (int k = 0; k < 3; k++) qpos[qa + k] += h * vel[d + k];
      double w[3] = {vel[d + 3], vel[d + 4], vel[d + 5]};
      double ang = h * norm3(w);
      quat_norm(qpos + qa + 3);
      if (ang > 0) {
        double n = norm3(w), ax[3] = {w[0] / n, w[1] / n, w[2] / n}, qr[4], q2[4];
        axisangle2quat(qr, ax, ang);
        quat_mul(q2, qpos + qa + 3, qr);
        memcpy(qpos + qa + 3, q2, sizeof q2);
      }
    } else {
      qpos[qa] += h * vel[d];
    }
  }
}

static int state_bad(const ora_env* e) {
  const mmz_model* m = &e->m;
  for (int i = 0; i < m->nq; i++) if (!(fabs(e->qpos[i]) < MAXVAL)) return 1;
  for (int i = 0; i < m->nv; i++) if (!(fabs(e->qvel[i]) < MAXVAL)) return 1;
  return 0;
}

/* mj_step with the RK4 integrator (mj_RungeKutta [EXT]); returns 1 if the state blew up */
static int mj_step(ora_env* e) {
  const mmz_model* m = &e->m;
  static const double A[3] = {0.5, 0.5, 1.0}, Bw[4] = {1.0 / 6, 1.0 / 3, 1.0 / 3, 1.0 / 6};
  int nq = m->nq, nv = m->nv;
  double h = m->timestep, q0[NQ], v0[ND], Xv[4][ND], F[4][ND];
  if (state_bad(e)) return 1;
  forward(e);
  for (int d = 0; d < nv; d++) if (!(fabs(e->qacc[d]) < MAXVAL)) return 1;
  memcpy(q0, e->qpos, sizeof(double) * nq);
  memcpy(v0, e->qvel, sizeof(double) * nv);
  memcpy(Xv[0], e->qvel, sizeof(double) * nv);
  memcpy(F[0], e->qacc, sizeof(double) * nv);
  for (int i = 1; i < 4; i++) {
    double dv[ND];
    for (int d = 0; d < nv; d++) dv[d] = A[i - 1] * Xv[i - 1][d];
    memcpy(e->qpos, q0, sizeof(double) * nq);
    integrate_pos(m, e->qpos, dv, h);
    for (int d = 0; d < nv; d++) e->qvel[d] = v0[d] + h * A[i - 1] * F[i - 1][d];
    memcpy(Xv[i], e->qvel, sizeof(double) * nv);
    forward(e);
    memcpy(F[i], e->qacc, sizeof(double) * nv);
  }
  double dv[ND], da[ND];
  for (int d = 0; d < nv; d++) {
    dv[d] = da[d] = 0;
    for (int i = 0; i < 4; i++) { dv[d] += Bw[i] * Xv[i][d]; da[d] += Bw[i] * F[i][d]; }
  }
  memcpy(e->qpos, q0, sizeof(double) * nq);
  integrate_pos(m, e->qpos, dv, h);
  for (int d = 0; d < nv; d++) e->qvel[d] = v0[d] + h * da[d];
  /* derived arrays (xpos, contacts, qacc) deliberately stay at the 4th-stage state: quirk Q15 */
  return state_bad(e);
}

/* ------------------------------------------------------------------ segment clamp (maze_env_utils.py:84-206) */
static double cross2(double ax, double ay, double bx, double by) { return ax * by - ay * bx; }
static int seg_detect(const mmz_model* m, const double* o, const double* n, double* point, double* refl) {
  double mvx = n[0] - o[0], mvy = n[1] - o[1];
  if (sqrt(mvx * mvx + mvy * mvy) <= 1e-8) return 0;
  int hit = 0;
  double bestd = 0;
  for (int s = 0; s < m->nseg; s++) {
    double x1 = m->seg[s][0], y1 = m->seg[s][1], x2 = m->seg[s][2], y2 = m->seg[s][3];
    double wx = x2 - x1, wy = y2 - y1;
    /* the move's end points straddle the wall line, and the wall's end points straddle the move line */
    double sa = cross2(wx, wy, o[0] - x1, o[1] - y1) * cross2(wx, wy, n[0] - x1, n[1] - y1);
    double sb = cross2(mvx, mvy, x1 - o[0], y1 - o[1]) * cross2(mvx, mvy, x2 - o[0], y2 - o[1]);
    if (!(sa <= 0.0 && sb <= 0.0)) continue;
    double den = cross2(wx, wy, mvx, mvy), num = cross2(wx, wy, x2 - o[0], y2 - o[1]);
    if (den == 0.0) continue; /* collinear: the reference raises ZeroDivisionError; treated as no hit */
    double px = o[0] + num / den * mvx, py = o[1] + num / den * mvy;
    double d = sqrt((px - o[0]) * (px - o[0]) + (py - o[1]) * (py - o[1]));
    if (!hit || d < bestd) {
      double tt = ((n[0] - x1) * wx + (n[1] - y1) * wy) / (wx * wx + wy * wy);
      double fx = x1 + tt * wx, fy = y1 + tt * wy;
      hit = 1; bestd = d;
      point[0] = px; point[1] = py;
      refl[0] = fx + (fx - n[0]); refl[1] = fy + (fy - n[1]);
    }
  }
  return hit;
}

/* ------------------------------------------------------------------ obs / reward / termination */
/* update_view of get_top_down_view (maze_env.py:268-322): a unit square around the source's continuous
 * (row, col) is distributed over the 3x3 neighbourhood of its integer cell; out-of-raster parts are dropped. */
static void view_splat(double view[5][5][3], double x, double y, int d, double robot_x, double robot_y, double s) {
  x -= robot_x; y -= robot_y;                                      /* :270-271 */
  const double rowc = 2 + (y + s / 2) / s, colc = 2 + (x + s / 2) / s; /* _xy_to_rowcol, :90-93 */
  const int row = (int)rowc, col = (int)colc;                      /* int(): toward zero, :277 */
  const double rf = rowc - floor(rowc), cf = colc - floor(colc);   /* Python `% 1` is never negative */
  const double wr[3] = {fmax(0.0, 0.5 - rf), fmin(1.0, rf + 0.5) - fmax(0.0, rf - 0.5), fmax(0.0, rf - 0.5)};
  const double wc[3] = {fmax(0.0, 0.5 - cf), fmin(1.0, cf + 0.5) - fmax(0.0, cf - 0.5), fmax(0.0, cf - 0.5)};
  for (int a = -1; a <= 1; a++)
    for (int b = -1; b <= 1; b++) {
      const int r = row + a, c = col + b;
      if (r >= 0 && r < 5 && c >= 0 && c < 5) view[r][c][d] += wr[a + 1] * wc[b + 1]; /* the nine cases :284-322 */
    }
}
/* get_top_down_view (maze_env.py:262-349): channel 0 BLOCK cells, 1 CHASM cells, 2 movable blocks, all relative
 * to the torso's data.xpos (view body 0); movable blocks are view bodies 1.. */
static void top_down_view(const ora_env* e, double* out) {
  const mmz_model* m = &e->m;
  double view[5][5][3];
  memset(view, 0, sizeof view);
  const int* vb = m->obj_body + m->nobj;
  const double rx = e->xpos[vb[0]][0], ry = e->xpos[vb[0]][1], s = m->cell_size;
  for (int i = 0; i < m->grid_h; i++)
    for (int j = 0; j < m->grid_w; j++) {
      const int cell = m->grid[i * m->grid_w + j];
      if (cell & MMZ_CELL_WALL) view_splat(view, j * s - m->origin[0], i * s - m->origin[1], 0, rx, ry, s);
      if (cell & MMZ_CELL_CHASM) view_splat(view, j * s - m->origin[0], i * s - m->origin[1], 1, rx, ry, s);
    }
  for (int k = 1; k < m->nviewb; k++) view_splat(view, e->xpos[vb[k]][0], e->xpos[vb[k]][1], 2, rx, ry, s);
  memcpy(out, view, sizeof view);
}
static void observe(const ora_env* e, double* obs) {
  const mmz_model* m = &e->m;
  int k = 0;
  for (int i = 0; i < 3 && i < m->n_agent_q; i++) obs[k++] = e->qpos[i];
  for (int o = 0; o < m->nobj; o++) for (int i = 0; i < 3; i++) obs[k++] = e->xpos[m->obj_body[o]][i];
  for (int i = 3; i < m->n_agent_q; i++) obs[k++] = e->qpos[i];
  for (int i = 0; i < m->n_agent_v; i++) obs[k++] = e->qvel[i];
  if (m->view_dim) { top_down_view(e, obs + k); k += MMZ_VIEW_DIM; } /* maze_env.py:353-354,369 */
  obs[k++] = e->t * 0.001;
}
static int first_goal(const mmz_model* m, const double* where) {
  for (int g = 0; g < m->ngoal; g++) {
    double s = 0;
    for (int i = 0; i < m->goal_dim[g]; i++) s += (where[i] - m->goal_pos[g][i]) * (where[i] - m->goal_pos[g][i]);
    if (sqrt(s) <= m->goal_thr[g]) return g;
  }
  return -1;
}
static double goal_dist(const mmz_model* m, const double* where) {
  double s = 0;
  for (int i = 0; i < m->goal_dim[0]; i++) s += (where[i] - m->goal_pos[0][i]) * (where[i] - m->goal_pos[0][i]);
  return sqrt(s);
}
static void task_rules(const mmz_model* m, const double* obs, double* reward, int* done) {
  int term = 0;
  if (m->term_rule == MMZ_TERM_AGENT) term = first_goal(m, obs) >= 0;
  else if (m->term_rule == MMZ_TERM_OBJECT) term = first_goal(m, obs + 3) >= 0;
  double r = 0;
  int g;
  switch (m->reward_rule) {
    case MMZ_REWARD_REACH: r = term ? 1.0 : m->penalty; break;
    case MMZ_REWARD_SCALED: g = first_goal(m, obs); r = g >= 0 ? m->goal_scale[g] : m->penalty; break;
    case MMZ_REWARD_SCALED_OBJECT: g = first_goal(m, obs + 3); r = g >= 0 ? m->goal_scale[g] : m->penalty; break;
    case MMZ_REWARD_DIST_OBJECT: r = -goal_dist(m, obs + 3) / m->task_scale; break;
    case MMZ_REWARD_DIST: r = -goal_dist(m, obs) / m->task_scale; break;
    default: r = 0;
  }
  *reward = r;
  *done = term;
}

/* ================================================================== exported API */
ora_env* ora_create(const void* blob, size_t bytes) {
  if (bytes != sizeof(mmz_model)) return NULL;
  const mmz_model* m = (const mmz_model*)blob;
  if (m->magic != MMZ_MAGIC || m->version != MMZ_VERSION || m->real_bytes != 8) return NULL;
  ora_env* e = (ora_env*)calloc(1, sizeof(ora_env));
  memcpy(&e->m, blob, sizeof(mmz_model));
  memcpy(e->qpos, m->qpos0, sizeof(double) * m->nq);
  kinematics(e);
  return e;
}
void ora_destroy(ora_env* e) { free(e); }
size_t ora_model_bytes(void) { return sizeof(mmz_model); }
void ora_set_warmstart(ora_env* e, int on) { e->use_warmstart = on; }

/* MujocoEnv.set_state -> mj_forward: refreshes the derived arrays */
void ora_set_state(ora_env* e, const double* qpos, const double* qvel, int t) {
  memcpy(e->qpos, qpos, sizeof(double) * e->m.nq);
  memcpy(e->qvel, qvel, sizeof(double) * e->m.nv);
  e->t = t;
  kinematics(e);
}
void ora_get_state(const ora_env* e, double* qpos, double* qvel, int* t) {
  memcpy(qpos, e->qpos, sizeof(double) * e->m.nq);
  memcpy(qvel, e->qvel, sizeof(double) * e->m.nv);
  *t = e->t;
}
void ora_observe(const ora_env* e, double* obs) { observe(e, obs); }

/* one mj_forward under `action`; state is not advanced */
void ora_forward(ora_env* e, const double* action) {
  for (int a = 0; a < e->m.nu; a++) e->ctrl[a] = e->m.step_kind == MMZ_STEP_TELEPORT ? 0.0 : action[a];
  forward(e);
}
/* which: 0 qacc, 1 qacc_smooth, 2 qfrc_bias, 3 qfrc_passive, 4 qfrc_smooth, 5 qfrc_act (nv each) */
void ora_get_vec(const ora_env* e, int which, double* out) {
  const double* src[] = {e->qacc, e->qacc_smooth, e->qfrc_bias, e->qfrc_passive, e->qfrc_smooth, e->qfrc_act};
  memcpy(out, src[which], sizeof(double) * e->m.nv);
}
void ora_get_M(const ora_env* e, double* out) {
  for (int i = 0; i < e->m.nv; i++) for (int j = 0; j < e->m.nv; j++) out[i * e->m.nv + j] = e->M[i][j];
}
void ora_get_counts(const ora_env* e, int* out) { out[0] = e->ncon; out[1] = e->nefc; out[2] = e->niter; out[3] = e->overflow; }
/* contact c -> dist, pos[3], frame[9], body1, body2, mu, margin (17 doubles) */
void ora_get_contact(const ora_env* e, int c, double* out) {
  const contact_t* k = &e->con[c];
  out[0] = k->dist;
  memcpy(out + 1, k->pos, sizeof(double) * 3);
  memcpy(out + 4, k->frame, sizeof(double) * 9);
  out[13] = k->body1; out[14] = k->body2; out[15] = k->mu; out[16] = k->margin;
}
void ora_get_efc(const ora_env* e, double* J, double* D, double* aref, double* force) {
  for (int i = 0; i < e->nefc; i++) {
    for (int d = 0; d < e->m.nv; d++) J[i * e->m.nv + d] = e->J[i][d];
    D[i] = e->efc_D[i]; aref[i] = e->efc_aref[i]; force[i] = e->efc_force[i];
  }
}
void ora_get_xpos(const ora_env* e, double* xpos, double* xquat) {
  for (int b = 0; b < e->m.nbody; b++) {
    memcpy(xpos + 3 * b, e->xpos[b], sizeof(double) * 3);
    memcpy(xquat + 4 * b, e->xquat[b], sizeof(double) * 4);
  }
}
/* raw mj_step with a fixed control (physics tests) */
int ora_mj_step(ora_env* e, const double* ctrl) {
  for (int a = 0; a < e->m.nu; a++) e->ctrl[a] = ctrl ? ctrl[a] : 0.0;
  return mj_step(e);
}
/* CollisionDetector.detect: returns hit, fills point[2], reflection[2] */
int ora_detect(const ora_env* e, const double* old_xy, const double* new_xy, double* point, double* refl) {
  return seg_detect(&e->m, old_xy, new_xy, point, refl);
}
void ora_task_rules(const ora_env* e, const double* obs, double* reward, int* done) { task_rules(&e->m, obs, reward, done); }

static void reinit(ora_env* e) {
  memcpy(e->qpos, e->m.qpos0, sizeof(double) * e->m.nq);
  memset(e->qvel, 0, sizeof e->qvel);
  kinematics(e);
}

/* MazeEnv.step (maze_env.py:448-481). info = {x, y, reward_forward, reward_ctrl}; returns done bits */
int ora_step(ora_env* e, const double* action, double* obs, double* reward, double* info) {
  const mmz_model* m = &e->m;
  int bad = 0;
  double inner = 0, fwd = 0, cc = 0;
  e->t += 1;
  if (m->step_kind == MMZ_STEP_TELEPORT) { /* PointEnv.step (point.py:44-61) */
    double old_xy[2] = {e->qpos[0], e->qpos[1]};
    e->qpos[2] += action[1];
    if (e->qpos[2] < -PI) e->qpos[2] += 2 * PI;
    else if (PI < e->qpos[2]) e->qpos[2] -= 2 * PI;
    double ori = e->qpos[2];
    e->qpos[0] += cos(ori) * action[0];
    e->qpos[1] += sin(ori) * action[0];
    for (int d = 0; d < m->nv; d++) e->qvel[d] = fmin(fmax(e->qvel[d], -m->vel_limit), m->vel_limit);
    for (int a = 0; a < m->nu; a++) e->ctrl[a] = 0;
    for (int k = 0; k < m->frame_skip && !bad; k++) bad = mj_step(e);
    if (!bad && m->manual_collision) { /* maze_env.py:450-464 */
      double new_xy[2] = {e->qpos[0], e->qpos[1]}, pt[2], rf[2];
      if (seg_detect(m, old_xy, new_xy, pt, rf)) {
        double pos[2] = {pt[0] + m->restitution * (rf[0] - pt[0]), pt[1] + m->restitution * (rf[1] - pt[1])}, p2[2], r2[2];
        if (seg_detect(m, old_xy, pos, p2, r2)) { pos[0] = old_xy[0]; pos[1] = old_xy[1]; }
        e->qpos[0] = pos[0]; e->qpos[1] = pos[1];
        kinematics(e); /* set_xy -> set_state -> mj_forward refreshes xpos */
      }
    }
  } else { /* AntEnv.step / SwimmerEnv.step (ant.py:61-73) */
    double before[2] = {e->qpos[0], e->qpos[1]};
    for (int a = 0; a < m->nu; a++) e->ctrl[a] = action[a];
    for (int k = 0; k < m->frame_skip && !bad; k++) bad = mj_step(e);
    double dt = m->timestep * m->frame_skip;
    double vx = (e->qpos[0] - before[0]) / dt, vy = (e->qpos[1] - before[1]) / dt;
    /* forward_reward_fn (ant.py:18-23): vnorm, vabs, or left to the host wrapper */
    fwd = m->forward_reward_kind == MMZ_FWD_VABS ? fabs(vx) + fabs(vy) : m->forward_reward_kind == MMZ_FWD_HOST ? 0.0 : sqrt(vx * vx + vy * vy);
    for (int a = 0; a < m->nu; a++) cc += action[a] * action[a];
    cc *= m->ctrl_cost_weight;
    inner = m->forward_reward_weight * fwd - cc;
  }
  int bits = 0;
  if (bad) { reinit(e); bits |= 4; inner = 0; fwd = 0; cc = 0; }
  observe(e, obs);
  double outer;
  int term;
  task_rules(m, obs, &outer, &term);
  *reward = m->inner_reward_scale * inner + outer;
  if (term) bits |= 1;
  if (m->max_episode_steps > 0 && e->t >= m->max_episode_steps) bits |= 1 | 2;
  info[0] = e->qpos[0]; info[1] = e->qpos[1]; info[2] = fwd; info[3] = -cc;
  return bits;
}

/* Bounded CPU baseline: advance n independent envs for `steps` steps with OpenMP threads.
 * actions [steps][n][nu]. Returns the number of env-steps done. */
long ora_rollout(const void* blob, size_t bytes, int n, const double* qpos, const double* qvel, const double* actions,
                 int steps, int nthreads, double* out_obs, double* out_reward) {
  long total = 0;
  const mmz_model* m = (const mmz_model*)blob;
  int nu = m->nu, nq = m->nq, nv = m->nv, od = m->obs_dim;
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads) reduction(+ : total) schedule(dynamic, 1)
#endif
  for (int i = 0; i < n; i++) {
    ora_env* e = ora_create(blob, bytes);
    double obs[64 + MMZ_VIEW_DIM], r = 0, info[4];
    ora_set_state(e, qpos + (size_t)i * nq, qvel + (size_t)i * nv, 0);
    for (int s = 0; s < steps; s++) {
      ora_step(e, actions + ((size_t)s * n + i) * nu, obs, &r, info);
      total++;
    }
    if (out_obs) memcpy(out_obs + (size_t)i * od, obs, sizeof(double) * od);
    if (out_reward) out_reward[i] = r;
    ora_destroy(e);
  }
  return total;
}

/* Persistent batch of independent environments (CPU baseline legs of bench.py, tests):
 * state survives between calls, one OpenMP thread per chunk of environments. */
typedef struct ora_batch {
  int n;
  ora_env** env;
} ora_batch;

ora_batch* ora_batch_create(const void* blob, size_t bytes, int n) {
  ora_batch* b = (ora_batch*)calloc(1, sizeof(ora_batch));
  b->n = n;
  b->env = (ora_env**)calloc((size_t)n, sizeof(ora_env*));
  for (int i = 0; i < n; i++) {
    b->env[i] = ora_create(blob, bytes);
    if (!b->env[i]) { for (int k = 0; k < i; k++) ora_destroy(b->env[k]); free(b->env); free(b); return NULL; }
    b->env[i]->use_warmstart = 1;
  }
  return b;
}
void ora_batch_destroy(ora_batch* b) {
  if (!b) return;
  for (int i = 0; i < b->n; i++) ora_destroy(b->env[i]);
  free(b->env);
  free(b);
}
/* qpos [n][nq], qvel [n][nv] */
void ora_batch_set_state(ora_batch* b, const double* qpos, const double* qvel) {
  for (int i = 0; i < b->n; i++)
    ora_set_state(b->env[i], qpos + (size_t)i * b->env[i]->m.nq, qvel + (size_t)i * b->env[i]->m.nv, 0);
}
/* one MazeEnv.step for every environment: actions [n][nu] -> obs [n][obs_dim], reward [n], done bits [n] */
long ora_batch_step(ora_batch* b, const double* actions, double* obs, double* reward, int* done, int nthreads) {
  const int n = b->n;
  if (n == 0) return 0;
  const int nu = b->env[0]->m.nu, od = b->env[0]->m.obs_dim;
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 4)
#endif
  for (int i = 0; i < n; i++) {
    double info[4];
    done[i] = ora_step(b->env[i], actions + (size_t)i * nu, obs + (size_t)i * od, reward + i, info);
  }
  return n;
}
