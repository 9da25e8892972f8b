// mmz_hkernel.cuh - the hybrid step kernel for the Ant family: two thread-to-data mappings in one block.
//
// A block owns 32 environments (one SM: 512 threads, the whole shared memory) and switches between two
// views of the same shared-memory workspace, separated by block barriers:
//
//   * TREE phases, "lane = environment" : lane e of every warp works on environment e and
//     the warps split the bodies of a tree level / geoms / dofs / contact slots between them. Kinematics,
//     motion axes, spatial inertias, RNE, the mass matrix, collision and the contact rows run this way:
//     all 32 lanes do useful work (the lanes <-> bodies mapping of mmz_dyn.cuh keeps 4..13 of 32 busy),
//     and a warp without an item costs no issue slots.
//   * SOLVER phase, "16 lanes = one environment, lane = degree of freedom" (mmz_dyn.cuh): warp w solves
//     environments w and w + 16. The Newton solve is register-resident (Hessian row per lane, shuffle
//     Gaussian elimination) and its iteration count is only shared by the 2 environments of a warp, not by
//     all 32 of the block.
//
// The workspace is laid out [slot][33]: element (slot, env) lives at slot * 33 + env. In the tree view
// (fixed slot per instruction, lanes = 32 environments) the bank is (slot + env) % 32: conflict-free, also
// for lane-varying slots. In the solver view a warp touches 16 consecutive slots for environments w and
// w + 16: banks slot + w + {0..15} and slot + w + 16 + {0..15}: conflict-free too.
#pragma once
#include "mmz_layout.h"
#include "mmz_narrow.cuh"

namespace mmz {

constexpr int TW = 16;  // warps per block
constexpr int TE = 32;  // environments per block
constexpr unsigned kAll = 0xffffffffu;
constexpr int kTMaxNewton = 24;
constexpr int kMaxCTasks = 160;  // collision items + 2 nv
// roles of a warp in the tree walk (TDerived::walk_kind)
enum { A_IDLE = 0, A_KIN = 1, A_PAIR_KIN = 2, A_PAIR_DYN = 3, A_ROOTDYN = 4, A_BOTH = 5 };
constexpr int kTMaxLineSearch = 24;


enum { TMODE_STEP = 0, TMODE_FORWARD = 1, TMODE_OBSERVE = 2, TMODE_RESET = 3, TMODE_REFRESH = 4 };
enum { T_DONE_BIT = 1, T_TRUNC_BIT = 2, T_UNSTABLE_BIT = 4 };
enum { T_FLAG_AUTO_RESET = 1 };

// per-environment integers
enum { TN_CON = 0, TN_OVERFLOW = 1, TN_ITER = 2, TN_NOTCONV = 3, TN_ITER_SUM = 4, TN_LS_SUM = 5, TN_CON_MAX = 6,
       TN_CAPPED = 7, TN_MOVED = 8, TN_BAD = 9, TN_LIM = 10, TN_CNT = 12 };

struct TDerived {               // appended to the model blob in device memory
  int32_t anc[MMZ_MAXBODY];     // bit a set in anc[b]: body a is b or an ancestor of b
  int32_t lvl_off[MMZ_MAXBODY + 1];  // bodies of level l: lvl_body[lvl_off[l] .. lvl_off[l+1])
  int32_t lvl_body[MMZ_MAXBODY];
  int32_t boxg[MMZ_MAXGEOM];    // geoms of type box on moving bodies, in geom order
  int32_t boxord[MMZ_MAXGEOM];  // geom -> its index in boxg (or -1)
  int32_t nboxg;
  int32_t nlev;
  int32_t pad[2];
  // bodies below the roots, grouped by the warp that walks them: the subtree of the k-th level-1 body belongs to
  // warp k % TW, parents first, so a chain needs no block barrier (lane e of one warp reads what it wrote)
  int32_t chain_off[TW + 1];
  int32_t chain_body[2 * MMZ_MAXBODY];  // (a chain appears twice when a pair of warps walks it)
  // role of warp w in the tree walk after the roots' kinematics (A_* below), its named barrier, and how many
  // threads meet at the barrier that publishes the roots' dynamic halves
  int32_t walk_kind[TW];
  int32_t walk_bar[TW];
  int32_t walk_root_count;
  int32_t sub_end[MMZ_MAXBODY];  // bodies are in depth-first order: the subtree of b is [b, sub_end[b])
  int32_t dof_act[MMZ_MAXDOF];   // actuators driving dof d: bit k set for actuator k
  int32_t dof_rel[MMZ_MAXDOF];   // bit k set: dof k is d, an ancestor or a descendant of d (the sparsity pattern of row d of M)
  // phase C (contact counting, mass-matrix rows, smooth forces): the tasks of warp w, longest-processing-time-first
  // over a cost model (a collision item is ~4 smooth-force tasks): c_item[c_off[w] .. c_off[w + 1])
  int32_t c_off[TW + 1];
  int32_t c_item[kMaxCTasks];
  // Fast kinematics of a body with ONE hinge joint below another moving body (the Ant's 8 leg links): its pose is two
  // 3x3 products, R = R_parent (R_body R_joint(q)) with the joint rotation in Rodrigues form I + sin K + (1 - cos) K^2,
  // instead of a chain of three quat2mat and two quaternion products. All constants below are in the parent's frame.
  int32_t kin_fast[MMZ_MAXBODY];   // 1: use the fast path (no child reads this body's quaternion)
  float kin_Rb[MMZ_MAXBODY][9];    // rotation of body_quat
  float kin_c[MMZ_MAXBODY][3];     // joint anchor: body_pos + R_body jnt_pos
  float kin_ax[MMZ_MAXBODY][3];    // joint axis: R_body jnt_axis
  float kin_K[MMZ_MAXBODY][9];     // cross-product matrix of jnt_axis (body frame), and its square
  float kin_K2[MMZ_MAXBODY][9];
  float iq_R[MMZ_MAXBODY][9];      // rotation of body_iquat (every body): world inertia = (R iq_R) diag (R iq_R)^T
  int32_t pad2[3];
  float ident[9];
  float padf[3];
};

struct TLayout {
  int nb, nj, nv, nq, nu, ng, nobj, obs_dim, nlev;
  int nlatch, obs_core;  // as in Layout (mmz_layout.h)
  int maxcon, cstride, ldm, nstate, nslots, model_bytes;
  int o_qpos, o_qvel, o_qacc, o_objpos;  // persisted rows, in this order
  int o_ctrl, o_q0, o_v0, o_accv, o_acca;
  int o_xpos, o_xquat, o_xmat, o_gpos, o_gax, o_gmat, o_cdof;
  int o_iw, o_ic, o_vel, o_acc, o_frc, o_fsub;
  int o_M, o_smooth, o_dir;
  int o_con, o_cnt, o_gcnt, o_obs, o_act;
  int o_lim;  // [nv][4] joint-limit rows of each dof: D lower, D upper, aref lower, aref upper (0 = no row)
  // solver v2 (models without box geoms): the stored contact Jacobian and the per-contact forces / Hessian weights, as
  // float4 arrays in natural [environment][contact] order. They OVERLAY the slots [o_nat, o_con) of the [slot][33]
  // workspace, whose arrays are all dead while the solver runs.
  int v2, o_nat, jes, fa_off;  // first slot (multiple of 4); float4s per environment of the Jacobian area; float4 offset of FA
  int topo;                    // 1 (solver v2): the dof tree is the Ant's (free root + 2-dof chains): sparse elimination order (elim_solve2).
                               // Solver v3, bits: 2 = dofs 0..5 are one free joint (every contact mask holds all six or none),
                               // 4 = the Ant's tree + the 2-dof chain (14, 15) of one movable block (elim_solve2<3>)
  int v3, njac, es;            // solver v3 (box instances): float4 Jacobian entries per environment; float4s per environment of its natural area
  int tail0;                   // solver v3: the last dof tree has exactly two dofs, tail0 and tail0 + 1 (a movable block's slides), else -1
};

struct TArgs {
  TLayout L;
  const void* model;
  float* state;        // [L.nstate][npad]
  int* counters;       // [2][npad]: t, number of resets
  int n, npad;
  const float* action; // [n][nu]
  float* obs;          // [n][obs_dim]
  float* reward;       // [n]
  uint8_t* done;       // [n]
  float* info;         // [n][4] or null
  float* qacc_out;     // TMODE_FORWARD: [n][nv]
  int* diag;           // [n][4], optional
  const uint8_t* mask; // TMODE_RESET
  unsigned long long seed;
  unsigned flags;
  int block0;          // first block of this launch (mmz_step_host pipelines block ranges)
  int env_offset;
  float tol;           // Newton convergence: |grad_d| <= tol * (magnitude of the terms grad_d is the sum of)
  ObsPeers peers;      // TMODE_STEP: fused observation gather (mmz_layout.h)
};


constexpr int HS = 33;  // row stride of the [slot][33] workspace

#ifdef MMZ_PHASE_TIMING  // development aid (tools/build_debug.py): cycles per phase of block 0, Newton iterations per solve
__device__ unsigned long long g_phase[16];
__device__ unsigned g_iter_hist[16];
__device__ unsigned long long g_wsolve[16], g_wwait[16];  // per warp of block 0: its own solve, its wait for the slowest
__device__ unsigned long long g_wC[16], g_wD[16];  // per warp of block 0: its own work in phases C and D (before the barrier)
__device__ unsigned g_launch;               // step launches so far
__device__ unsigned g_blk[8];               // per-block phase cycles of this launch (thread 0 of every block; printed for launch MMZ_PRINT_LAUNCH)
#define MMZ_TICK(i) do { if (threadIdx.x == 0) { const long long t_ = clock64(); if (blockIdx.x == 0) g_phase[i] += t_ - tick_; bphase[i] += (unsigned)(t_ - tick_); tick_ = t_; } } while (0)
__device__ unsigned long long g_sec[16];  // solver sections, summed over the warps of block 0 (slot 15: Newton iterations)
#define MMZ_STICK(i) do { if (blockIdx.x == 0 && e == 0) { const long long t_ = clock64(); atomicAdd(&g_sec[i], (unsigned long long)(t_ - stick_)); stick_ = t_; } } while (0)
#else
#define MMZ_TICK(i) do { } while (0)
#define MMZ_STICK(i) do { } while (0)
#endif

// Contact record of solver v2 (floats; mmz_layout.h C_* is the record of the first solver). The narrow phase leaves
// point, normal, first tangent (the second is their cross product), distance, the inverse weight, the two bodies and
// the two geoms (the mixed solref / solimp / margin / friction are re-derived from those); the constraint rows then
// replace bodies by signed dof masks and the temporaries by (mu, D, aref[4]); once the Jacobian is stored, point and
// frame are dead and J a - aref of the four pyramid rows takes their place.
enum { K_POS = 0, K_N = 3, K_T1 = 6, K_BODY1 = 9, K_BODY2 = 10, K_MPOS = 9, K_MNEG = 10, K_MU = 11, K_D = 12, K_AREF = 13,
       K_DIST = 13, K_INVW = 14, K_GEOM = 15, K_OTHER = 16, K_JAR = 0, K_STRIDE = 17,
       K3_JV = 4, K3_JOFF = 17, K3_STRIDE = 19 };  // solver v3: J dir per pyramid row of contacts 16.. (over the dead point / normal); ONE word: the contact's dof mask | offset of its Jacobian entries << 16
// (Without box geoms the narrow-phase result travels BY VALUE, in registers: a RawContact passed by reference lives in local
// memory, and with the whole shared memory in use the SM has no L1 left - every local access is an L2 round trip: AntUMaze
// 6.05 -> 5.96 ms. The box instances keep their contacts in local arrays anyway and are faster by reference - measured.)
static __device__ __noinline__ void write_contact_record2(float* c, const RawContact& rc, int b1, int b2, float iw, int g, int other);
static __device__ __noinline__ void write_contact_record2v(float* c, float dist, float px, float py, float pz, float nx, float ny, float nz,
                                                          float hx, float hy, float hz, int b1, int b2, float iw, int g, int other);


template <int NVP, int BOX>
struct HEnv {
  static constexpr bool V2 = BOX == 0;  // which solver (and contact record) the instance uses
  static constexpr bool V3 = BOX != 0;  // box instances: the stored Jacobian on each contact's own dofs (solve_g3)
  const mmz_model* m;
  const TDerived* dv;
  float* sm;
  float4* jg;               // solver v2: this lane's environment in the Jacobian area, [contact][NVP + 1] float4
  float4* fg;               // solver v2: its per-contact (force, Hessian weights) pairs, [contact][2] float4
  float* hg;                // solver v3: this lane's row of the contact part of the Hessian, accumulated in shared memory
#ifdef MMZ_PHASE_TIMING
  unsigned bphase[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int bmaxit = 0;
#endif
  int e, wid;               // tree view: lane = environment e; warp wid takes items wid, wid + 16, ...
  int genv, lane, gshift;   // solver view: environment wid + 16 * (laneid / 16), lane = dof
  float limD[2], limA[2];   // this lane's joint-limit rows (solver view)
  float tol;                // Newton convergence tolerance (relative to the magnitude of the cancelling terms)

#define S(i) sm[(i) * HS + e]
#define W_(i) sm[(i) * HS + genv]
  MMZ_DI int& I(int i) const { return reinterpret_cast<int*>(sm)[i * HS + e]; }
  MMZ_DI int& IW(int i) const { return reinterpret_cast<int*>(sm)[i * HS + genv]; }
  // spatial quantities are taken about a point near the robot (its first three coordinates), not the world
  // origin: translation invariant, and it avoids fp32 cancellation far from the maze origin
  MMZ_DI void ref(const TLayout& L, float* r) const {
    r[0] = S(L.o_qpos); r[1] = S(L.o_qpos + 1); r[2] = (m->jnt_type[0] == MMZ_JNT_FREE) ? S(L.o_qpos + 2) : 0.f;
  }

  // ------------------------------------------------------------------ phase A: one body, in two halves
  // kinematic half: mj_kinematics of the body and its motion axes (reads the parent's pose)
  MMZ_DI void body_kin(const TLayout& L, int b) {
    const int p = m->body_parent[b];
    float pos[3], quat[4], R[9], rf[3];
    ref(L, rf);
    if (dv->kin_fast[b] == 1) {  // one hinge below a moving body: matrices only (TDerived::kin_*)
      const int j = m->body_jntadr[b], qa = m->jnt_qadr[j], d = m->jnt_dadr[j];
      float sn, cs;
      __sincosf(S(L.o_qpos + qa) - m->qpos0[qa], &sn, &cs);
      const float c1 = 1.f - cs;
      float Rj[9], Rbj[9], Rp[9];
#pragma unroll
      for (int k = 0; k < 9; k++) Rj[k] = ((k == 0 || k == 4 || k == 8) ? 1.f : 0.f) + sn * dv->kin_K[b][k] + c1 * dv->kin_K2[b][k];
#pragma unroll
      for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++)
          Rbj[3 * r + c] = dv->kin_Rb[b][3 * r] * Rj[c] + dv->kin_Rb[b][3 * r + 1] * Rj[3 + c] + dv->kin_Rb[b][3 * r + 2] * Rj[6 + c];
      // everything above is independent of the parent
#pragma unroll
      for (int k = 0; k < 9; k++) Rp[k] = S(L.o_xmat + 9 * p + k);
#pragma unroll
      for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) R[3 * r + c] = Rp[3 * r] * Rbj[c] + Rp[3 * r + 1] * Rbj[3 + c] + Rp[3 * r + 2] * Rbj[6 + c];
      float an[3], ax[3], off[3], cc[6];
      mat_vec(an, Rp, dv->kin_c[b]);
      mat_vec(ax, Rp, dv->kin_ax[b]);
#pragma unroll
      for (int k = 0; k < 3; k++) an[k] += S(L.o_xpos + 3 * p + k);
      mat_vec(off, R, m->jnt_pos[j]);
      const float at[3] = {an[0] - rf[0], an[1] - rf[1], an[2] - rf[2]};
      cross3(cc + 3, at, ax);
      cc[0] = ax[0]; cc[1] = ax[1]; cc[2] = ax[2];
#pragma unroll
      for (int i = 0; i < 6; i++) S(L.o_cdof + 6 * d + i) = cc[i];
#pragma unroll
      for (int k = 0; k < 3; k++) S(L.o_xpos + 3 * b + k) = an[k] - off[k];
#pragma unroll
      for (int k = 0; k < 9; k++) S(L.o_xmat + 9 * b + k) = R[k];
      return;
    }
    int freed = -1;  // first dof of a free joint of this body
    if (BOX != 0 && dv->kin_fast[b] == 2) {  // (compiled into the box instances only: the Point always carries its arrow box)
      // The Point (point.xml:19-26): a root body on slide x, slide y and a hinge about z, all through its origin, no body
      // rotation. The generic joint loop below does three quat2mat, a quaternion product and a dozen dependent loads of
      // joint constants for what is a translation and one rotation about z - and this body is the whole tree walk of the
      // small robots: their step waits for this one warp.
      const int j = m->body_jntadr[b], qa = m->jnt_qadr[j], d = m->jnt_dadr[j];
      const float x = S(L.o_qpos + qa) - m->qpos0[qa], y = S(L.o_qpos + qa + 1) - m->qpos0[qa + 1];
      const float th = S(L.o_qpos + qa + 2) - m->qpos0[qa + 2];
      pos[0] = m->body_pos[b][0] + x; pos[1] = m->body_pos[b][1] + y; pos[2] = m->body_pos[b][2];
      const float zax[3] = {0.f, 0.f, 1.f};
      axisangle2quat(quat, zax, th);
      const float at[2] = {pos[0] - rf[0], pos[1] - rf[1]};
#pragma unroll
      for (int i = 0; i < 6; i++) {
        S(L.o_cdof + 6 * d + i) = i == 3 ? 1.f : 0.f;
        S(L.o_cdof + 6 * (d + 1) + i) = i == 4 ? 1.f : 0.f;
      }
      S(L.o_cdof + 6 * (d + 2) + 0) = 0.f; S(L.o_cdof + 6 * (d + 2) + 1) = 0.f; S(L.o_cdof + 6 * (d + 2) + 2) = 1.f;
      S(L.o_cdof + 6 * (d + 2) + 3) = at[1]; S(L.o_cdof + 6 * (d + 2) + 4) = -at[0]; S(L.o_cdof + 6 * (d + 2) + 5) = 0.f;
    } else {
    if (p < 0) {
#pragma unroll
      for (int k = 0; k < 3; k++) pos[k] = m->body_pos[b][k];
#pragma unroll
      for (int k = 0; k < 4; k++) quat[k] = m->body_quat[b][k];
    } else {
      float Rp[9], qp[4];
#pragma unroll
      for (int k = 0; k < 9; k++) Rp[k] = S(L.o_xmat + 9 * p + k);
#pragma unroll
      for (int k = 0; k < 4; k++) qp[k] = S(L.o_xquat + 4 * p + k);
      mat_vec(pos, Rp, m->body_pos[b]);
#pragma unroll
      for (int k = 0; k < 3; k++) pos[k] += S(L.o_xpos + 3 * p + k);
      quat_mul(quat, qp, m->body_quat[b]);
    }
    const int j0 = m->body_jntadr[b], j1 = j0 + m->body_jntnum[b];
#pragma unroll 1
    for (int j = j0; j < j1; j++) {
      const int qa = m->jnt_qadr[j], type = m->jnt_type[j], d = m->jnt_dadr[j];
      if (type == MMZ_JNT_FREE) {
        float q[4];
#pragma unroll
        for (int k = 0; k < 4; k++) q[k] = S(L.o_qpos + qa + 3 + k);
        quat_norm(q);  // MuJoCo normalises the stored quaternion in place
#pragma unroll
        for (int k = 0; k < 4; k++) { S(L.o_qpos + qa + 3 + k) = q[k]; quat[k] = q[k]; }
#pragma unroll
        for (int k = 0; k < 3; k++) pos[k] = S(L.o_qpos + qa + k);
        freed = d;
        continue;
      }
      // hinge / slide: axis and anchor in the frame reached so far; its motion axis about the reference point
      quat2mat(R, quat);
      float an[3], ax[3], c[6];
      mat_vec(an, R, m->jnt_pos[j]);
#pragma unroll
      for (int k = 0; k < 3; k++) an[k] += pos[k];
      mat_vec(ax, R, m->jnt_axis[j]);
      const float dq = S(L.o_qpos + qa) - m->qpos0[qa];
      if (type == MMZ_JNT_SLIDE) {
        c[0] = c[1] = c[2] = 0.f; c[3] = ax[0]; c[4] = ax[1]; c[5] = ax[2];
#pragma unroll
        for (int k = 0; k < 3; k++) pos[k] += ax[k] * dq;
      } else {  // hinge: rotate about the anchor
        float at[3] = {an[0] - rf[0], an[1] - rf[1], an[2] - rf[2]};
        cross3(c + 3, at, ax);
        c[0] = ax[0]; c[1] = ax[1]; c[2] = ax[2];
        float qr[4], q2[4], off[3];
        axisangle2quat(qr, m->jnt_axis[j], dq);
        quat_mul(q2, quat, qr);
#pragma unroll
        for (int k = 0; k < 4; k++) quat[k] = q2[k];
        quat2mat(R, quat);
        mat_vec(off, R, m->jnt_pos[j]);
#pragma unroll
        for (int k = 0; k < 3; k++) pos[k] = an[k] - off[k];
      }
#pragma unroll
      for (int i = 0; i < 6; i++) S(L.o_cdof + 6 * d + i) = c[i];
    }
    }
    quat_norm(quat);
    quat2mat(R, quat);
#pragma unroll
    for (int k = 0; k < 3; k++) S(L.o_xpos + 3 * b + k) = pos[k];
#pragma unroll
    for (int k = 0; k < 4; k++) S(L.o_xquat + 4 * b + k) = quat[k];
#pragma unroll
    for (int k = 0; k < 9; k++) S(L.o_xmat + 9 * b + k) = R[k];
    if (freed >= 0) {  // free joint: 3 world-aligned translations, then rotations about the body axes through its origin
      const float at[3] = {pos[0] - rf[0], pos[1] - rf[1], pos[2] - rf[2]};
#pragma unroll
      for (int k = 0; k < 3; k++) {
#pragma unroll
        for (int i = 0; i < 6; i++) S(L.o_cdof + 6 * (freed + k) + i) = (i == 3 + k) ? 1.f : 0.f;
        const float ax[3] = {R[k], R[3 + k], R[6 + k]};
        float lin[3];
        cross3(lin, at, ax);
#pragma unroll
        for (int i = 0; i < 3; i++) { S(L.o_cdof + 6 * (freed + 3 + k) + i) = ax[i]; S(L.o_cdof + 6 * (freed + 3 + k) + 3 + i) = lin[i]; }
      }
    }
  }
  // The forward pass of RNE in two parts. body_vel: velocity and bias acceleration of the body from its parent's - the only
  // part that runs DOWN a chain, a handful of multiply-adds per dof. body_frc: world spatial inertia and the inertial
  // force I a + v x* I v - by far the longer part, and independent from body to body once velocities exist.
  MMZ_DI void body_vel(const TLayout& L, int b) {
    const int p = m->body_parent[b];
    const int d0 = m->body_dofadr[b], d1 = d0 + m->body_dofnum[b];
    float v[6], a[6];
    if (p < 0) {
#pragma unroll
      for (int k = 0; k < 6; k++) { v[k] = 0.f; a[k] = 0.f; }
#pragma unroll
      for (int k = 0; k < 3; k++) a[3 + k] = -m->gravity[k];  // gravity as base acceleration
    } else {
#pragma unroll
      for (int k = 0; k < 6; k++) { v[k] = S(L.o_vel + 6 * p + k); a[k] = S(L.o_acc + 6 * p + k); }
    }
    float vf[6];
#pragma unroll
    for (int k = 0; k < 6; k++) vf[k] = v[k];
#pragma unroll 1
    for (int d = d0; d < d1; d++) {
      const int j = m->dof_jnt[d], kk = d - m->jnt_dadr[j];
      const float qv = S(L.o_qvel + d);
      const bool isfree = m->jnt_type[j] == MMZ_JNT_FREE;
      if (isfree && kk < 3) {  // world-aligned translation: no axis derivative
#pragma unroll
        for (int k = 0; k < 3; k++) v[3 + k] += (k == kk) ? qv : 0.f;
        if (kk == 2) {
#pragma unroll
          for (int k = 0; k < 6; k++) vf[k] = v[k];
        }
        continue;
      }
      float s[6], sd[6], vs[6];
#pragma unroll
      for (int k = 0; k < 6; k++) { s[k] = S(L.o_cdof + 6 * d + k); vs[k] = isfree ? vf[k] : v[k]; }
      cross_motion(sd, vs, s);
#pragma unroll
      for (int i = 0; i < 6; i++) { a[i] += sd[i] * qv; v[i] += s[i] * qv; }
    }
#pragma unroll
    for (int k = 0; k < 6; k++) { S(L.o_vel + 6 * b + k) = v[k]; S(L.o_acc + 6 * b + k) = a[k]; }
  }
  MMZ_DI void body_frc(const TLayout& L, int b) {
    float pos[3], R[9], rf[3];
    ref(L, rf);
#pragma unroll
    for (int k = 0; k < 3; k++) pos[k] = S(L.o_xpos + 3 * b + k);
#pragma unroll
    for (int k = 0; k < 9; k++) R[k] = S(L.o_xmat + 9 * b + k);
    // ---- world spatial inertia about the reference point
    float Iw[10];
    {
      float ip[3], Ri[9], c[3];
      mat_vec(ip, R, m->body_ipos[b]);
#pragma unroll
      for (int r = 0; r < 3; r++)  // world axes of the inertial frame: R times the (constant) rotation of body_iquat
#pragma unroll
        for (int cc = 0; cc < 3; cc++)
          Ri[3 * r + cc] = R[3 * r] * dv->iq_R[b][cc] + R[3 * r + 1] * dv->iq_R[b][3 + cc] + R[3 * r + 2] * dv->iq_R[b][6 + cc];
#pragma unroll
      for (int k = 0; k < 3; k++) c[k] = ip[k] + pos[k] - rf[k];
      const float* dg = m->body_inertia[b];
      const float mass = m->body_mass[b], cc = dot3(c, c);
      Iw[0] = Ri[0] * Ri[0] * dg[0] + Ri[1] * Ri[1] * dg[1] + Ri[2] * Ri[2] * dg[2] + mass * (cc - c[0] * c[0]);
      Iw[1] = Ri[3] * Ri[3] * dg[0] + Ri[4] * Ri[4] * dg[1] + Ri[5] * Ri[5] * dg[2] + mass * (cc - c[1] * c[1]);
      Iw[2] = Ri[6] * Ri[6] * dg[0] + Ri[7] * Ri[7] * dg[1] + Ri[8] * Ri[8] * dg[2] + mass * (cc - c[2] * c[2]);
      Iw[3] = Ri[0] * Ri[3] * dg[0] + Ri[1] * Ri[4] * dg[1] + Ri[2] * Ri[5] * dg[2] - mass * c[0] * c[1];
      Iw[4] = Ri[0] * Ri[6] * dg[0] + Ri[1] * Ri[7] * dg[1] + Ri[2] * Ri[8] * dg[2] - mass * c[0] * c[2];
      Iw[5] = Ri[3] * Ri[6] * dg[0] + Ri[4] * Ri[7] * dg[1] + Ri[5] * Ri[8] * dg[2] - mass * c[1] * c[2];
      Iw[6] = mass * c[0]; Iw[7] = mass * c[1]; Iw[8] = mass * c[2];
      Iw[9] = mass;
#pragma unroll
      for (int k = 0; k < 10; k++) S(L.o_iw + 10 * b + k) = Iw[k];
    }
    float v[6], a[6], Ia[6], Iv[6], vxIv[6];
#pragma unroll
    for (int k = 0; k < 6; k++) { v[k] = S(L.o_vel + 6 * b + k); a[k] = S(L.o_acc + 6 * b + k); }
    inert_mul(Ia, Iw, a);
    inert_mul(Iv, Iw, v);
    cross_force(vxIv, v, Iv);
#pragma unroll
    for (int k = 0; k < 6; k++) S(L.o_frc + 6 * b + k) = Ia[k] + vxIv[k];
  }

  // kinematics only (refresh of the derived arrays after a reset / set_state)
  MMZ_DI void kinematics_only(const TLayout& L) {
#pragma unroll 1
    for (int lvl = 0; lvl < L.nlev; lvl++) {
      for (int i = dv->lvl_off[lvl] + wid; i < dv->lvl_off[lvl + 1]; i += TW) body_kin(L, dv->lvl_body[i]);
      __syncthreads();
    }
  }

  // ------------------------------------------------------------------ phase B tasks
  MMZ_DI void geom_pose(const TLayout& L, int g) {  // centre and long (local z) axis: all spheres and capsules need
    const int b = m->geom_body[g];
    float Rb[9], p[3], Rg[9], az[3], ax[3];
#pragma unroll
    for (int k = 0; k < 9; k++) Rb[k] = S(L.o_xmat + 9 * b + k);
    mat_vec(p, Rb, m->geom_pos[g]);
    quat2mat(Rg, m->geom_quat[g]);
    az[0] = Rg[2]; az[1] = Rg[5]; az[2] = Rg[8];
    mat_vec(ax, Rb, az);
#pragma unroll
    for (int k = 0; k < 3; k++) { S(L.o_gpos + 3 * g + k) = p[k] + S(L.o_xpos + 3 * b + k); S(L.o_gax + 3 * g + k) = ax[k]; }
    if (BOX && dv->boxord[g] >= 0) {  // box geoms need their whole frame
      float R[9];
#pragma unroll
      for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) R[3 * r + c] = Rb[3 * r] * Rg[c] + Rb[3 * r + 1] * Rg[3 + c] + Rb[3 * r + 2] * Rg[6 + c];
#pragma unroll
      for (int k = 0; k < 9; k++) S(L.o_gmat + 9 * dv->boxord[g] + k) = R[k];
    }
  }
  // sum over the subtree rooted at b of n-float records
  MMZ_DI void subtree_sum(const TLayout& L, int b, int src, int dst, int n) {
    float acc[10];
    for (int k = 0; k < n; k++) acc[k] = S(src + n * b + k);
#pragma unroll 1
    for (int c = b + 1; c < dv->sub_end[b]; c++)
      for (int k = 0; k < n; k++) acc[k] += S(src + n * c + k);
    for (int k = 0; k < n; k++) S(dst + n * b + k) = acc[k];
  }

  // ------------------------------------------------------------------ phase C tasks
  // row i of the mass matrix (composite rigid body): M[i][j] for the ancestors j of dof i
  MMZ_DI void mass_row(const TLayout& L, int i) {
    float Ic[10], ci[6], f[6];
    const int b = m->dof_body[i];
#pragma unroll
    for (int k = 0; k < 10; k++) Ic[k] = S(L.o_ic + 10 * b + k);
#pragma unroll
    for (int k = 0; k < 6; k++) ci[k] = S(L.o_cdof + 6 * i + k);
    inert_mul(f, Ic, ci);
#pragma unroll 1
    for (int j = i; j >= 0; j = m->dof_parent[j]) {
      float cj[6];
#pragma unroll
      for (int k = 0; k < 6; k++) cj[k] = S(L.o_cdof + 6 * j + k);
      float v = dot6(cj, f);
      if (j == i) v += m->dof_armature[i];
      S(L.o_M + i * L.ldm + j) = v;
      S(L.o_M + j * L.ldm + i) = v;
    }
  }
  // qfrc_smooth[d] = passive - bias + actuation
  MMZ_DI void smooth_dof(const TLayout& L, int d) {
    float cd[6], fs[6];
    const int b = m->dof_body[d];
#pragma unroll
    for (int k = 0; k < 6; k++) { cd[k] = S(L.o_cdof + 6 * d + k); fs[k] = S(L.o_fsub + 6 * b + k); }
    const float bias = dot6(cd, fs);
    const float passive = -m->dof_damping[d] * S(L.o_qvel + d);
    float act = 0.f;
#pragma unroll 1
    for (int bits = dv->dof_act[d]; bits; bits &= bits - 1) {
      const int k = __ffs(bits) - 1;
      float c = S(L.o_ctrl + k);
      if (m->act_limited[k]) c = fminf(fmaxf(c, m->act_ctrlrange[k][0]), m->act_ctrlrange[k][1]);
      act += m->act_gear[k] * c;
    }
    S(L.o_smooth + d) = passive - bias + act;
  }

  // mixed contact parameters of geom g against `other` (-1 floor, -2 wall / platform box, >= 0 another geom)
  MMZ_DI void mix_params(int g, int other, float* par /* margin, mu, solref[2], solimp[5] */) const {
    float om, of;
    const float *osr, *osi;
    if (other == -1) { om = m->floor_margin; of = m->floor_friction[0]; osr = m->floor_solref; osi = m->floor_solimp; }
    else if (other == -2) { om = m->wall_margin; of = m->wall_friction[0]; osr = m->wall_solref; osi = m->wall_solimp; }
    else { om = m->geom_margin[other]; of = m->geom_friction[other][0]; osr = m->geom_solref[other]; osi = m->geom_solimp[other]; }
    par[0] = fmaxf(m->geom_margin[g], om);
    par[1] = fmaxf(m->geom_friction[g][0], of);
#pragma unroll
    for (int k = 0; k < 2; k++) par[2 + k] = 0.5f * (m->geom_solref[g][k] + osr[k]);
#pragma unroll
    for (int k = 0; k < 5; k++) par[4 + k] = 0.5f * (m->geom_solimp[g][k] + osi[k]);
  }
  MMZ_DI void cell_range(const float* c, const float* ext, int* i0, int* i1, int* j0, int* j1) const {
    const float s = m->cell_size, hs = m->wall_half[0];
    *j0 = max(0, (int)ceilf((c[0] - ext[0] + m->origin[0] - hs) / s));
    *j1 = min(m->grid_w - 1, (int)floorf((c[0] + ext[0] + m->origin[0] + hs) / s));
    *i0 = max(0, (int)ceilf((c[1] - ext[1] + m->origin[1] - hs) / s));
    *i1 = min(m->grid_h - 1, (int)floorf((c[1] + ext[1] + m->origin[1] + hs) / s));
  }
  MMZ_DI bool moving_pair_ok(int g1, int g2) const {
    const int b1 = m->geom_body[g1], b2 = m->geom_body[g2];
    if (b1 == b2 || m->body_parent[b1] == b2 || m->body_parent[b2] == b1) return false;
    return (m->geom_contype[g1] & m->geom_conaffinity[g2]) || (m->geom_contype[g2] & m->geom_conaffinity[g1]);
  }
  // stores a narrow-phase record into contact slot `slot` (layout K_* above); the normal points
  // from body b1 (geom1) to body b2 (geom2), -1 = world
  MMZ_DI void write_contact(const TLayout& L, int slot, const RawContact& rc, int b1, int b2, float iw, int g, int other) {
    float* c = sm + (L.o_con + slot * L.cstride) * HS + e;
    if constexpr (BOX != 0) write_contact_record2(c, rc, b1, b2, iw, g, other);
    else write_contact_record2v(c, rc.dist, rc.pos[0], rc.pos[1], rc.pos[2], rc.normal[0], rc.normal[1], rc.normal[2], rc.hint[0], rc.hint[1],
                                rc.hint[2], b1, b2, iw, g, other);
  }
  // number of collision items: one per geom, then (BOX only) BCAND candidate slots per box geom
  static constexpr int BCELLS = 9;                       // maze cells a box geom can reach (3 x 3)
  MMZ_DI int box_cands() const { return 1 + 2 * BCELLS + (dv->nboxg - 1); }  // floor, cells x {wall, platform}, other boxes
  MMZ_DI int n_items(const TLayout& L) const { return L.ng + (BOX ? dv->nboxg * box_cands() : 0); }
  MMZ_DI int item_base(const TLayout& L, int item) const {
    int base = 0;
#pragma unroll 1
    for (int k = 0; k < item; k++) base += I(L.o_gcnt + k) & 31;
    return base;
  }
  // A box candidate item (item >= ng): candidate c of box geom k - the floor (plane-box corners), a maze box or a later
  // box geom (box-box). The narrow phase is by far the longest task of the small robots' step, so the warp that ran it
  // in the counting pass KEEPS what it found (local memory) across the block barrier and writes it in the second pass
  // instead of running it again (forward(): `kept`).
  struct BoxFound { RawContact rc[8]; int n, b1, b2, other, g; float iw; };
  MMZ_DI void box_item_find(const TLayout& L, int item, BoxFound& f) {
    const int bc_n = box_cands(), kbox = (item - L.ng) / bc_n, cand = (item - L.ng) - kbox * bc_n;
    const int g = dv->boxg[kbox];
    int n = 0;
    const int body = m->geom_body[g];
    f.g = g; f.b1 = -1; f.b2 = body; f.other = -1; f.iw = m->geom_invweight[g];
    if (((m->geom_contype[g] | m->geom_conaffinity[g]) & 1) && m->collision_on) {
      const float invw = m->geom_invweight[g];
      float gp[3], gm[9], sz[3];
#pragma unroll
      for (int k = 0; k < 3; k++) { gp[k] = S(L.o_gpos + 3 * g + k); sz[k] = m->geom_size[g][k]; }
#pragma unroll
      for (int k = 0; k < 9; k++) gm[k] = S(L.o_gmat + 9 * kbox + k);
      RawContact* rc = f.rc;
      if (cand == 0) {  // corners below the plane, at most 4; geom1 = plane
        if (m->has_floor) {
          const float margin = fmaxf(m->geom_margin[g], m->floor_margin);
          for (int c = 0; c < 8 && n < 4; c++) {
            float loc[3] = {(c & 1 ? 1.f : -1.f) * sz[0], (c & 2 ? 1.f : -1.f) * sz[1], (c & 4 ? 1.f : -1.f) * sz[2]}, wp[3];
            mat_vec(wp, gm, loc);
            const float dist = wp[2] + gp[2] - m->floor_z;
            if (dist < margin) {
              rc[n].dist = dist;
              rc[n].pos[0] = wp[0] + gp[0]; rc[n].pos[1] = wp[1] + gp[1]; rc[n].pos[2] = wp[2] + gp[2] - 0.5f * dist;
              rc[n].normal[0] = 0.f; rc[n].normal[1] = 0.f; rc[n].normal[2] = 1.f;
              rc[n].hint[0] = rc[n].hint[1] = rc[n].hint[2] = 0.f;
              n++;
            }
          }
        }
      } else if (cand - 1 < 2 * BCELLS) {  // maze boxes; geom1 = wall (lower geom id), geom2 = this box
        const float wallmargin = fmaxf(m->geom_margin[g], m->wall_margin);
        float ext[3];
#pragma unroll
        for (int k = 0; k < 3; k++) ext[k] = fabsf(gm[3 * k]) * sz[0] + fabsf(gm[3 * k + 1]) * sz[1] + fabsf(gm[3 * k + 2]) * sz[2];
        ext[0] += wallmargin; ext[1] += wallmargin;
        int i0, i1, j0, j1;
        cell_range(gp, ext, &i0, &i1, &j0, &j1);
        const int nj = max(0, j1 - j0 + 1), ncell = nj * max(0, i1 - i0 + 1);
        const int ci = (cand - 1) >> 1, slot = (cand - 1) & 1;
        if (ci < ncell) {
          const int i = i0 + ci / nj, j = j0 + ci % nj;
          const int code = m->grid[i * m->grid_w + j];
          if (code & (slot == 0 ? MMZ_CELL_WALL : MMZ_CELL_PLATFORM)) {
            const float bc[3] = {j * m->cell_size - m->origin[0], i * m->cell_size - m->origin[1], slot == 0 ? m->wall_z : m->plat_z};
            f.other = -2;
            n = box_box(bc, dv->ident, m->wall_half, gp, gm, sz, wallmargin, rc);
          }
        }
        if (cand == 1 && ncell > BCELLS) I(L.o_cnt + TN_OVERFLOW) = 1;  // the box reaches more cells than slots
      } else {  // box against a later box geom on another moving body
        const int k2 = kbox + 1 + (cand - 1 - 2 * BCELLS);
        if (k2 < dv->nboxg) {
          const int g2 = dv->boxg[k2];
          if (moving_pair_ok(g, g2)) {
            float gp2[3], gm2[9];
#pragma unroll
            for (int k = 0; k < 3; k++) gp2[k] = S(L.o_gpos + 3 * g2 + k);
#pragma unroll
            for (int k = 0; k < 9; k++) gm2[k] = S(L.o_gmat + 9 * k2 + k);
            f.other = g2; f.b1 = body; f.b2 = m->geom_body[g2];
            f.iw = invw + m->geom_invweight[g2];
            n = box_box(gp, gm, sz, gp2, gm2, m->geom_size[g2], fmaxf(m->geom_margin[g], m->geom_margin[g2]), rc);
          }
        }
      }
    }
    f.n = n;
  }
  // (all contacts of one box item share their normal - the clipped face's, the floor's - hence their frame: made once)
  MMZ_DI void box_item_write(const TLayout& L, int item, int base, const BoxFound& f) {
    const int nw = __reduce_max_sync(kAll, f.n);
    if (nw == 0) return;
    float fr[9];
#pragma unroll
    for (int k = 0; k < 3; k++) { fr[k] = f.n > 0 ? f.rc[0].normal[k] : (k == 2 ? 1.f : 0.f); fr[3 + k] = f.n > 0 ? f.rc[0].hint[k] : 0.f; }
    make_frame(fr);
#pragma unroll 1
    for (int k = 0; k < nw; k++) {
      if (k < f.n && base + k < L.maxcon) {
        float* c = sm + (L.o_con + (base + k) * L.cstride) * HS + e;
        const RawContact& rc = f.rc[k];
#pragma unroll
        for (int i = 0; i < 3; i++) c[(K_POS + i) * HS] = rc.pos[i];
#pragma unroll
        for (int i = 0; i < 6; i++) c[(K_N + i) * HS] = fr[i];
        c[K_DIST * HS] = rc.dist;
        c[K_INVW * HS] = f.iw;
        c[K_BODY1 * HS] = __int_as_float(f.b1);
        c[K_BODY2 * HS] = __int_as_float(f.b2);
        c[K_GEOM * HS] = __int_as_float(f.g);
        c[K_OTHER * HS] = __int_as_float(f.other);
      }
    }
  }
  // Collision item `item`. pass 0 counts its contacts (into o_gcnt), pass 1 writes them at the slots following those
  // of the items before it: the contact order is the item order, independent of warp timing.
  //   item < ng: sphere / capsule geom against the floor plane, the maze boxes near it and the movable box geoms;
  //   item >= ng (BOX): candidate c of box geom k: the floor (plane-box corners), a maze box or a later box geom (box-box).
  // Pass 0 also leaves, above the count (5 bits), which tests produced a contact: bit 5 + end for the floor, bits
  // 7 + 2 cand (+ 1) for box candidate cand < 12, bit 31 = "a later candidate": pass 1 skips the whole item when no
  // environment of the warp found anything, and otherwise repeats only the tests that hit in some environment.
  MMZ_DI void collide_item(const TLayout& L, int item, int pass) {
    int n = 0;
    unsigned hits = 0, wanted = 0xffffffffu;
    if (pass == 1) {  // (box candidate items only carry the count)
      wanted = __reduce_or_sync(kAll, (unsigned)I(L.o_gcnt + item));
      if ((wanted & 31u) == 0) return;
      if (item >= L.ng) wanted = 0xffffffffu;
    }
    const int base = pass == 1 ? item_base(L, item) : 0;
    if (item < L.ng) {
      const int g = item, type = m->geom_type[g];
      const bool capsule = type == MMZ_GEOM_CAPSULE;
      const bool valid = (type == MMZ_GEOM_SPHERE || capsule) && ((m->geom_contype[g] | m->geom_conaffinity[g]) & 1) && m->collision_on;
      if (valid) {
        const float r = m->geom_size[g][0], gmarg = m->geom_margin[g], invw = m->geom_invweight[g];
        const int body = m->geom_body[g];
        float gp[3], axv[3], p0[3], p1[3], ext[3];
#pragma unroll
        for (int k = 0; k < 3; k++) { gp[k] = S(L.o_gpos + 3 * g + k); axv[k] = S(L.o_gax + 3 * g + k); }
        const float hl = capsule ? m->geom_size[g][1] : 0.f;
#pragma unroll
        for (int k = 0; k < 3; k++) { const float a = axv[k] * hl; p0[k] = gp[k] + a; p1[k] = gp[k] - a; ext[k] = fabsf(a); }
        // ---- floor plane (normal +z); geom1 = plane, geom2 = this geom
        if (m->has_floor) {
          const float margin = fmaxf(gmarg, m->floor_margin);
          const float d0 = p0[2] - m->floor_z - r, d1 = p1[2] - m->floor_z - r;
          float hint[3] = {0.f, 0.f, 0.f};
          if (capsule && fabsf(axv[2]) <= 0.999999f) { hint[0] = axv[0]; hint[1] = axv[1]; hint[2] = axv[2]; }  // first tangent along the axis
#pragma unroll 1
          for (int end = 0; end < 2; end++) {
            const float d = end == 0 ? d0 : d1;
            if ((end == 1 && !capsule) || !(d < margin)) continue;
            hits |= 32u << end;
            if (pass == 1 && base + n < L.maxcon) {
              RawContact rc;
#pragma unroll
              for (int k = 0; k < 3; k++) { rc.pos[k] = end == 0 ? p0[k] : p1[k]; rc.normal[k] = (k == 2) ? 1.f : 0.f; rc.hint[k] = hint[k]; }
              rc.dist = d; rc.pos[2] -= r + 0.5f * d;
              write_contact(L, base + n, rc, -1, body, invw, g, -1);
              if (kRowsInD) contact_rows2(L, base + n);
            }
            n++;
          }
        }
        // ---- box-shaped obstacles; geom1 = this geom, geom2 = box (the normal points from the geom into the box):
        // the maze boxes of the cells it can reach (wall, then platform), then the movable box geoms
        const float mg = r + fmaxf(gmarg, m->wall_margin);
        ext[0] += mg; ext[1] += mg;
        int i0 = 0, i1 = -1, j0 = 0, j1 = -1;
        if (wanted >> 7) cell_range(gp, ext, &i0, &i1, &j0, &j1);
        const int nslot = m->elevated ? 2 : 1, nj = max(0, j1 - j0 + 1), ncell = nj * max(0, i1 - i0 + 1);
        const int ncand = (wanted >> 7) ? nslot * ncell + (BOX ? dv->nboxg : 0) : 0;
#pragma unroll 1
        for (int cand = 0; cand < ncand; cand++) {
          const unsigned cbit = cand < 12 ? 128u << (2 * cand) : 0x80000000u;
          if (!(wanted & (cand < 12 ? 3u * cbit : cbit))) continue;  // (pass 1) no environment of the warp hit this one
          float bc[3], bR[9], bh[3], margin, iw = invw;
          int other = -2, b2 = -1;
          if (cand < nslot * ncell) {
            const int ci = cand / nslot, slot = cand - ci * nslot, i = i0 + ci / nj, j = j0 + ci % nj;
            const int code = m->grid[i * m->grid_w + j];
            if (!(code & (slot == 0 ? MMZ_CELL_WALL : MMZ_CELL_PLATFORM))) continue;
            bc[0] = j * m->cell_size - m->origin[0]; bc[1] = i * m->cell_size - m->origin[1]; bc[2] = slot == 0 ? m->wall_z : m->plat_z;
#pragma unroll
            for (int k = 0; k < 9; k++) bR[k] = dv->ident[k];
#pragma unroll
            for (int k = 0; k < 3; k++) bh[k] = m->wall_half[k];
            margin = fmaxf(gmarg, m->wall_margin);
          } else {
            const int kb = cand - nslot * ncell, gb = dv->boxg[kb];
            if (!moving_pair_ok(g, gb)) continue;
#pragma unroll
            for (int k = 0; k < 3; k++) { bc[k] = S(L.o_gpos + 3 * gb + k); bh[k] = m->geom_size[gb][k]; }
#pragma unroll
            for (int k = 0; k < 9; k++) bR[k] = S(L.o_gmat + 9 * kb + k);
            margin = fmaxf(gmarg, m->geom_margin[gb]);
            other = gb; b2 = m->geom_body[gb];
            iw = invw + m->geom_invweight[gb];
          }
          // capsule: both end caps when both are within the margin, otherwise the segment point nearest the box
          RawContact r0, r1;
          int n0 = sphere_box(p0, r, bc, bR, bh, margin, &r0), n1 = 0;
          if (capsule) {
            n1 = sphere_box(p1, r, bc, bR, bh, margin, &r1);
            if (!(n0 && n1)) {
              const float ts = capsule_nearest(p0, p1, bc, bR, bh);
              float pt[3];
#pragma unroll
              for (int k = 0; k < 3; k++) pt[k] = p0[k] + ts * (p1[k] - p0[k]);
              n0 = sphere_box(pt, r, bc, bR, bh, margin, &r0);
              n1 = 0;
            }
          }
          if (n0) { hits |= cbit; if (pass == 1 && base + n < L.maxcon) { write_contact(L, base + n, r0, body, b2, iw, g, other); if (kRowsInD) contact_rows2(L, base + n); } n++; }
          if (n1) { hits |= cand < 12 ? cbit << 1 : cbit; if (pass == 1 && base + n < L.maxcon) { write_contact(L, base + n, r1, body, b2, iw, g, other); if (kRowsInD) contact_rows2(L, base + n); } n++; }
        }
      }
    } else if (BOX) {
      BoxFound f;
      box_item_find(L, item, f);
      n = f.n;
      if (pass == 1) box_item_write(L, item, base, f);
    }
    if (pass == 0) I(L.o_gcnt + item) = min(n, 31) | (int)hits;
  }

  // ------------------------------------------------------------------ constraint rows (mj_makeConstraint)
  MMZ_DI static float impedance(const float* si, float r) {
    const float d0 = fminf(fmaxf(si[0], 1e-4f), 0.9999f), d1 = fminf(fmaxf(si[1], 1e-4f), 0.9999f);
    const float width = si[2], mid = si[3], power = si[4];
    if (d0 == d1 || width <= kMinVal) return 0.5f * (d0 + d1);
    const float x = fabsf(r) / width;
    float y;
    if (x >= 1.f) return d1;
    if (x <= 0.f) return d0;
    if (power == 1.f) y = x;
    else if (power == 2.f) y = x <= mid ? x * x / mid : 1.f - (1.f - x) * (1.f - x) / (1.f - mid);  // MuJoCo's default power
    else if (x <= mid) y = powf(x, power) / powf(mid, power - 1.f);
    else y = 1.f - powf(1.f - x, power) / powf(1.f - mid, power - 1.f);
    return d0 + y * (d1 - d0);
  }
  MMZ_DI void row_params(const float* solref, const float* solimp, float pos, float margin, float diag, float* D, float* kr,
                         float* bb) const {
    const float tc = fmaxf(solref[0], 2.f * m->timestep), dr = solref[1];  // refsafe
    const float dmax = fminf(fmaxf(solimp[1], 1e-4f), 0.9999f);
    const float k = 1.f / fmaxf(kMinVal, dmax * dmax * tc * tc * dr * dr);
    *bb = 2.f / fmaxf(kMinVal, dmax * tc);
    const float imp = impedance(solimp, pos - margin);
    const float R = fmaxf(kMinVal, (1.f - imp) * diag / imp);
    *D = 1.f / R;
    *kr = k * imp * (pos - margin);
  }
  // contact slot c (tree view): narrow-phase record -> D, aref[4], signed dof masks, point relative to the
  // reference. J qvel is the point velocity of body2 minus body1, from the body velocities RNE has computed.
  // the same for the record of solver v2
  // Without box geoms (an item has one or two contacts) the warp that wrote a record computes its rows right away: no
  // barrier and no separate phase for the rows (AntUMaze 5.955 -> 5.926 ms). With box geoms an item can have 8 contacts
  // and the rows are spread over the warps in their own phase E (merged: PointUMaze 0.164 -> 0.175 ms, AntPush 6.41 -> 6.69).
  static constexpr bool kRowsInD = BOX == 0;
  MMZ_DI void contact_rows2(const TLayout& L, int c) {
    const int o = L.o_con + c * L.cstride;
    float rf[3], cp[3], fr[9], par[9], v[3] = {0.f, 0.f, 0.f};
    ref(L, rf);
#pragma unroll
    for (int k = 0; k < 3; k++) cp[k] = S(o + K_POS + k) - rf[k];
#pragma unroll
    for (int k = 0; k < 6; k++) fr[k] = S(o + K_N + k);
    cross3(fr + 6, fr, fr + 3);
    const float dist = S(o + K_DIST), invw = S(o + K_INVW);
    mix_params(__float_as_int(S(o + K_GEOM)), __float_as_int(S(o + K_OTHER)), par);
    const float margin = par[0], mu = par[1];
    const int b1 = __float_as_int(S(o + K_BODY1)), b2 = __float_as_int(S(o + K_BODY2));
    const int mask1 = b1 >= 0 ? m->body_dofmask[b1] : 0, mask2 = b2 >= 0 ? m->body_dofmask[b2] : 0;
#pragma unroll 1
    for (int side = 0; side < 2; side++) {
      const int b = side == 0 ? b2 : b1;
      if (b < 0) continue;
      float bv[6], wxp[3];
#pragma unroll
      for (int k = 0; k < 6; k++) bv[k] = S(L.o_vel + 6 * b + k);
      cross3(wxp, bv, cp);
      const float sg = side == 0 ? 1.f : -1.f;
#pragma unroll
      for (int k = 0; k < 3; k++) v[k] += sg * (bv[3 + k] + wxp[k]);
    }
    const float jv0 = dot3(fr, v), jv1 = dot3(fr + 3, v), jv2 = dot3(fr + 6, v);
    float D, kr, bb;
    row_params(par + 2, par + 4, dist, margin, invw * (1.f + mu * mu), &D, &kr, &bb);
#pragma unroll
    for (int k = 0; k < 3; k++) S(o + K_POS + k) = cp[k];
    S(o + K_MPOS) = __int_as_float(mask2 & ~mask1);
    S(o + K_MNEG) = __int_as_float(mask1 & ~mask2);
    S(o + K_MU) = mu;
    S(o + K_D) = 1.f / fmaxf(kMinVal, 2.f * mu * mu / D);  // all edges of the pyramid share R = 2 mu^2 R_first
    S(o + K_AREF + 0) = -bb * (jv0 + mu * jv1) - kr;
    S(o + K_AREF + 1) = -bb * (jv0 - mu * jv1) - kr;
    S(o + K_AREF + 2) = -bb * (jv0 + mu * jv2) - kr;
    S(o + K_AREF + 3) = -bb * (jv0 - mu * jv2) - kr;
  }
  // named barriers among a subset of the block's warps (ids 1..15; 0 is __syncthreads)
  MMZ_DI static void named_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
  MMZ_DI static void named_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

  // ================================================================== solver view (16 lanes = one environment)
  MMZ_DI unsigned gballot(bool p) const { return (__ballot_sync(kAll, p) >> gshift) & 0xffffu; }
  MMZ_DI static float gsum16(float v) {
#pragma unroll
    for (int off = 8; off > 0; off >>= 1) v += __shfl_xor_sync(kAll, v, off);
    return v;
  }
  // Joint-limit rows of dof d, tree view (lane = environment): side 0 = lower (J = +1), side 1 = upper (J = -1). They
  // depend on the state only, so idle warps compute them for all 32 environments at once while the roots' kinematics
  // run; the solver view then just loads its lane's four numbers.
  MMZ_DI void limit_rows_t(const TLayout& L, int d) {
    const int jl = m->dof_jnt[d];
    const bool limited = m->jnt_limited[jl] && m->jnt_type[jl] >= MMZ_JNT_SLIDE;
    float D2[2] = {0.f, 0.f}, A2[2] = {0.f, 0.f};
    if (limited) {
      const float q = S(L.o_qpos + m->jnt_qadr[jl]), qv = S(L.o_qvel + d), margin = m->jnt_margin[jl];
#pragma unroll
      for (int s = 0; s < 2; s++) {
        const float pos = s == 0 ? q - m->jnt_range[jl][0] : m->jnt_range[jl][1] - q;
        if (pos < margin) {
          float D, kr, bb;
          row_params(m->jnt_solref[jl], m->jnt_solimp[jl], pos, margin, m->dof_invweight0[d], &D, &kr, &bb);
          D2[s] = D;
          A2[s] = -bb * (s == 0 ? 1.f : -1.f) * qv - kr;
        }
      }
    }
    S(L.o_lim + 4 * d) = D2[0]; S(L.o_lim + 4 * d + 1) = D2[1]; S(L.o_lim + 4 * d + 2) = A2[0]; S(L.o_lim + 4 * d + 3) = A2[1];
  }
  // ... and the solver view (lane = dof) keeps them in registers
  MMZ_DI void limit_rows_g(const TLayout& L) {
    const bool me = lane < L.nv;
    limD[0] = me ? W_(L.o_lim + 4 * lane) : 0.f; limD[1] = me ? W_(L.o_lim + 4 * lane + 1) : 0.f;
    limA[0] = me ? W_(L.o_lim + 4 * lane + 2) : 0.f; limA[1] = me ? W_(L.o_lim + 4 * lane + 3) : 0.f;
  }
  // The same elimination for solver v2. `dg` is added to this lane's diagonal element on the fly (the joint-limit rows;
  // 1 for the identity rows beyond the model), so the caller's row is a plain copy of the mass-matrix row. TOPO 1 =
  // the Ant's tree (a free root, dofs 0-5, and 2-dof chains (6,7), (8,9), ... hanging off it): the pivots run from the
  // leaves to the root, where elimination has no fill-in, and step j only touches the columns of the ancestors of j -
  // 67 instead of 91 column updates (each a shuffle and a multiply-add) for the Ant.
  template <int TOPO>
  MMZ_DI float elim_solve2(float (&h)[NVP], float rhs, float dg) const {
    float invd = 1.f;
    constexpr bool kSide = (TOPO == 1 && NVP == 14) || (TOPO == 3 && NVP == 16);
    constexpr int NP = TOPO == 3 ? 5 : 4;  // TOPO 3: the Ant's tree and a second tree, the 2-dof chain (14, 15) of a movable block
    if (kSide) {
      // The four ankle rows do not couple with each other, and once they are gone neither do the four hip rows: their
      // pivots run SIDE BY SIDE (a lane that is one of the four pivots has a zero multiplier for the other three, so the
      // rows the shuffles read are the ones a one-by-one elimination would read). The chain of dependent (shuffle,
      // reciprocal, multiply-add) steps is 1 + 1 + 6 long instead of 14.
#pragma unroll
      for (int ph = 0; ph < 2; ph++) {
        float f[NP], t[NP];
#pragma unroll
        for (int m = 0; m < NP; m++) {
          const int j = 7 + 2 * m - ph;
          const bool own = lane == j;
          float hd = h[j];
          asm("" : "+f"(hd));  // opaque: four selects on (lane == j) over one array otherwise become h[lane] in LOCAL memory
          const float hj = own ? hd + dg : hd;
          const float piv = fmaxf(__shfl_sync(kAll, hj, j, 16), kMinVal);
          float inv;
          asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(piv));
          f[m] = own ? 0.f : hd * inv;
          invd = own ? inv : invd;
          t[m] = __shfl_sync(kAll, rhs, j, 16);
        }
#pragma unroll
        for (int m = 0; m < NP; m++) rhs -= f[m] * t[m];
#pragma unroll
        for (int k = 0; k < 6; k++) {  // (the block's chain does not hang off the root: m < 4)
#pragma unroll
          for (int m = 0; m < 4; m++) t[m] = __shfl_sync(kAll, h[k], 7 + 2 * m - ph, 16);
#pragma unroll
          for (int m = 0; m < 4; m++) h[k] -= f[m] * t[m];
        }
        if (ph == 0) {
#pragma unroll
          for (int m = 0; m < NP; m++) h[6 + 2 * m] -= f[m] * __shfl_sync(kAll, h[6 + 2 * m], 7 + 2 * m, 16);
        }
      }
    }
#pragma unroll
    for (int jj = kSide ? NVP - 6 : 0; jj < NVP; jj++) {
      const int j = TOPO != 0 ? NVP - 1 - jj : jj;
      const bool own = lane == j;
      float hd = h[j];
      asm("" : "+f"(hd));
      const float hj = own ? hd + dg : hd;
      const float piv = fmaxf(__shfl_sync(kAll, hj, j, 16), kMinVal);
      float inv;
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(piv));
      const float rj = __shfl_sync(kAll, rhs, j, 16);
      const float f = own ? 0.f : hd * inv;
      invd = own ? inv : invd;
      rhs -= f * rj;
      if (TOPO != 0) {
#pragma unroll
        for (int k = 0; k < NVP; k++) {
          const bool anc = j < 6 ? k < j : (k < 6 || ((j & 1) && k == j - 1));
          if (anc) h[k] -= f * __shfl_sync(kAll, h[k], j, 16);
        }
      } else {
#pragma unroll
        for (int k = j + 1; k < NVP; k++) h[k] -= f * __shfl_sync(kAll, h[k], j, 16);
      }
    }
    return rhs * invd;
  }
  // ================================================================== solver v2 (models without box geoms)
  // The contact Jacobian is built ONCE per forward evaluation and kept in shared memory, one float4 per (contact, dof):
  // (J_n, mu J_t1, mu J_t2, |J_n| + mu (|J_t1| + |J_t2|)). With the friction coefficient folded into the tangential
  // columns the four pyramid rows of a contact are n + t1, n - t1, n + t2, n - t2. The passes over the CONTACTS then run
  // with lane = contact (all contacts of an environment at once, no cross-lane reductions: J x is a loop over the dofs,
  // the rows, their forces and Hessian weights stay in the lane), the passes over the DOFS (gradient J^T f, Hessian rows)
  // with lane = dof read those results back. 3.3 k -> 1.9 k warp instructions per Newton iteration of a standing Ant.
  MMZ_DI float4* jrow(int c) const { return jg + c * (NVP + 1); }

  MMZ_DI void build_jac(const TLayout& L, const float (&cd)[6], int ncon, int ncw) {
#pragma unroll 1
    for (int c = 0; c < ncw; c++) {
      const int cs = L.o_con + c * L.cstride;
      const int mp = __float_as_int(W_(cs + K_MPOS)), mn = __float_as_int(W_(cs + K_MNEG));
      const float s = (float)(mp >> lane & 1) - (float)(mn >> lane & 1);
      float p[3], fr[9], wxp[3];
#pragma unroll
      for (int k = 0; k < 3; k++) p[k] = W_(cs + K_POS + k);
#pragma unroll
      for (int k = 0; k < 6; k++) fr[k] = W_(cs + K_N + k);
      cross3(fr + 6, fr, fr + 3);
      cross3(wxp, cd, p);
      const float v[3] = {s * (cd[3] + wxp[0]), s * (cd[4] + wxp[1]), s * (cd[5] + wxp[2])};
      const float mu = W_(cs + K_MU);
      float jn = dot3(fr, v), jt1 = mu * dot3(fr + 3, v), jt2 = mu * dot3(fr + 6, v);
      if (c >= ncon) { jn = 0.f; jt1 = 0.f; jt2 = 0.f; }  // the slot holds nothing for this environment (the other one of the warp has more contacts)
      if (lane < NVP) jrow(c)[lane] = make_float4(jn, jt1, jt2, fabsf(jn) + fabsf(jt1) + fabsf(jt2));
    }
  }
  // lane = contact, x = qacc: J x - aref per pyramid row (kept in the record for the line search), and for the dof passes
  // the force of the active rows in (n, t1, t2) coordinates + the scale of its rounding error, and the Hessian weights
  MMZ_DI void contact_pass_grad(const TLayout& L, int ncon, int ncw) {
    const int c = lane;
    if (c < ncw) {
      float4 F = make_float4(0.f, 0.f, 0.f, 0.f), Wt = F;
      if (c < ncon) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, sa = 0.f;
        const float4* jr = jrow(c);
#pragma unroll
        for (int k = 0; k < NVP; k++) {
          const float4 j = jr[k];
          const float x = W_(L.o_qacc + k);
          s0 = fmaf(j.x, x, s0); s1 = fmaf(j.y, x, s1); s2 = fmaf(j.z, x, s2); sa = fmaf(j.w, fabsf(x), sa);
        }
        const int cs = L.o_con + c * L.cstride;
        const float r0 = W_(cs + K_AREF), r1 = W_(cs + K_AREF + 1), r2 = W_(cs + K_AREF + 2), r3 = W_(cs + K_AREF + 3);
        const float D = W_(cs + K_D);
        const float j0 = s0 + s1 - r0, j1 = s0 - s1 - r1, j2 = s0 + s2 - r2, j3 = s0 - s2 - r3;
        W_(cs + K_JAR) = j0; W_(cs + K_JAR + 1) = j1; W_(cs + K_JAR + 2) = j2; W_(cs + K_JAR + 3) = j3;
        const float a0 = j0 < 0.f, a1 = j1 < 0.f, a2 = j2 < 0.f, a3 = j3 < 0.f;
        const float f0 = a0 * D * j0, f1 = a1 * D * j1, f2 = a2 * D * j2, f3 = a3 * D * j3;
        const float bound = sa + fmaxf(fmaxf(fabsf(r0), fabsf(r1)), fmaxf(fabsf(r2), fabsf(r3)));
        F = make_float4(f0 + f1 + f2 + f3, f0 - f1, f2 - f3, D * bound * (a0 + a1 + a2 + a3));
        Wt = make_float4(D * (a0 - a1), D * (a2 - a3), D * (a0 + a1), D * (a2 + a3));
      }
      fg[2 * c] = F; fg[2 * c + 1] = Wt;
    }
  }
  // Newton solver (mj_solNewton), same algorithm and stopping rules as solve_g
  MMZ_DI void solve_g2(const TLayout& L, bool warmstart) {
    const int nv = L.nv;
#ifdef MMZ_PHASE_TIMING
    long long stick_ = clock64();
#endif
    const bool me = lane < nv;
    const unsigned limbits = gballot(limD[0] > 0.f || limD[1] > 0.f);
    float mrow[NVP];  // this lane's row of the mass matrix (zero outside its sparsity pattern and outside the model)
    const int rel = me ? dv->dof_rel[lane] : 0;
#pragma unroll
    for (int k = 0; k < NVP; k++) mrow[k] = (rel >> k & 1) ? W_(L.o_M + lane * L.ldm + k) : 0.f;
    const float sm_ = me ? W_(L.o_smooth + lane) : 0.f;
    float cd[6];
#pragma unroll
    for (int k = 0; k < 6; k++) cd[k] = me ? W_(L.o_cdof + 6 * lane + k) : 0.f;
    // The Jacobian area overlays the mass matrix, the motion axes and every other array the solver view has read by
    // now (all environments of the block, in the other layout): no warp may write it before all warps are here. The
    // same barrier ends phase D (contact records, their rows, the contact count).
    __syncthreads();
    const int ncon = IW(L.o_cnt + TN_CON);
    const int ncw = max(ncon, __shfl_xor_sync(kAll, ncon, 16));  // the two environments of the warp
    const bool constrained = ncon > 0 || limbits != 0;
    build_jac(L, cd, ncon, ncw);
    float al = (warmstart && me) ? W_(L.o_qacc + lane) : 0.f;
    if (!(fabsf(al) < kMaxVal)) al = 0.f;  // a blown-up environment restarts from zero
    if (me) W_(L.o_qacc + lane) = al;
    const int nlim = __popc(gballot(limD[0] > 0.f)) + __popc(gballot(limD[1] > 0.f));
    if (lane == 0) {
      IW(L.o_cnt + TN_ITER) = 0; IW(L.o_cnt + TN_LIM) = nlim;
      IW(L.o_cnt + TN_CON_MAX) = max(IW(L.o_cnt + TN_CON_MAX), ncon);
    }
    __syncwarp();
    bool done = false;
    float Ma = 0.f, Mabs = 0.f;  // (M qacc)[lane] and the magnitude of its terms, updated with every step
#pragma unroll
    for (int k = 0; k < NVP; k++) { const float t = mrow[k] * W_(L.o_qacc + k); Ma += t; Mabs += fabsf(t); }
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    MMZ_STICK(0);  // prologue: register loads, block barrier, Jacobian
#pragma unroll 1
    for (int it = 0; it < kTMaxNewton; it++) {
      float grad = Ma - sm_, dadd = 0.f;
      float mag = Mabs + fabsf(sm_);
      float ljar[2];
#pragma unroll
      for (int s = 0; s < 2; s++) {
        const float sign = s == 0 ? 1.f : -1.f;
        ljar[s] = sign * al - limA[s];
        if (limD[s] > 0.f && ljar[s] < 0.f) {
          grad += limD[s] * ljar[s] * sign;
          mag += limD[s] * (fabsf(al) + fabsf(limA[s]));
          dadd += limD[s];
        }
      }
      contact_pass_grad(L, ncon, ncw);
      __syncwarp();
      MMZ_STICK(1);  // contact pass (lane = contact)
#pragma unroll 1
      for (int c = 0; c < ncw; c += 2) {  // two contacts per trip: four loads in flight (same order of additions as one by one)
        const bool two = c + 1 < ncw;
        const float4 j = lane < NVP ? jrow(c)[lane] : zero4, F = fg[2 * c];
        const float4 j1 = (two && lane < NVP) ? jrow(c + 1)[lane] : zero4, F1 = two ? fg[2 * c + 2] : zero4;
        grad += j.x * F.x + j.y * F.y + j.z * F.z;
        mag = fmaf(j.w, F.w, mag);
        grad += j1.x * F1.x + j1.y * F1.y + j1.z * F1.z;
        mag = fmaf(j1.w, F1.w, mag);
      }
      if (gballot(me && fabsf(grad) > tol * mag + 1e-30f) == 0) done = true;
      MMZ_STICK(2);  // gradient (lane = dof)
      if (__all_sync(kAll, done)) break;
      float hrow[NVP];
#pragma unroll
      for (int k = 0; k < NVP; k++) hrow[k] = mrow[k];  // the diagonal terms of the limit rows join in the elimination
#pragma unroll 1
      for (int c = 0; c < ncw; c += 2) {  // two contacts per trip (loads of both in flight; same order of additions as one by one)
        const bool two = c + 1 < ncw;
        const float4 Wt = fg[2 * c + 1], Wb = two ? fg[2 * c + 3] : zero4;
        const float wnn = Wt.z + Wt.w, wnb = Wb.z + Wb.w;
        if (!__any_sync(kAll, wnn != 0.f || wnb != 0.f)) continue;  // no active row in these contacts, in either environment
        const float4 j = lane < NVP ? jrow(c)[lane] : zero4, jb = (two && lane < NVP) ? jrow(c + 1)[lane] : zero4;
        const float u0 = wnn * j.x + Wt.x * j.y + Wt.y * j.z, u1 = Wt.x * j.x + Wt.z * j.y, u2 = Wt.y * j.x + Wt.w * j.z;
        const float v0 = wnb * jb.x + Wb.x * jb.y + Wb.y * jb.z, v1 = Wb.x * jb.x + Wb.z * jb.y, v2 = Wb.y * jb.x + Wb.w * jb.z;
        const float4* jr = jrow(c);
        const float4* jr2 = two ? jrow(c + 1) : jr;  // without a second contact: finite numbers times v = 0
#pragma unroll
        for (int k = 0; k < NVP; k++) {
          const float4 jk = jr[k], jl = jr2[k];
          hrow[k] += u0 * jk.x + u1 * jk.y + u2 * jk.z;
          hrow[k] += v0 * jl.x + v1 * jl.y + v2 * jl.z;
        }
      }
      MMZ_STICK(3);  // Hessian rows
      const float dg = me ? dadd : 1.f, rhs0 = me ? -grad : 0.f;
      const float dr = L.topo == 1 ? elim_solve2<1>(hrow, rhs0, dg) : elim_solve2<0>(hrow, rhs0, dg);
      if (me && !done) W_(L.o_dir + lane) = dr;
      __syncwarp();
      MMZ_STICK(4);  // elimination
      float alpha = 1.f;
      int ls = 0;
      bool exact = false;
      float md = 0.f, mdabs = 0.f;  // (M dir)[lane] and the magnitude of its terms
#pragma unroll
      for (int k = 0; k < NVP; k++) { const float t = mrow[k] * W_(L.o_dir + k); md += t; mdabs += fabsf(t); }
      if (__any_sync(kAll, constrained && !done)) {
        // This lane's rows - the two joint-limit rows of its dof and (lane = contact) the four pyramid rows of its contact -
        // stay in registers for the whole search. While jar + alpha jv < 0 a row adds D (jar + alpha jv) jv = rb + alpha ra
        // to the derivative along the direction and D jv^2 = ra to the curvature. Absent rows: jar = 1, the rest 0.
        float rj[6], rv[6], ra[6], rb[6];
#pragma unroll
        for (int s = 0; s < 2; s++) {
          const float jv = s == 0 ? dr : -dr;
          const bool on = limD[s] > 0.f;
          rj[s] = on ? ljar[s] : 1.f; rv[s] = on ? jv : 0.f;
          ra[s] = on ? limD[s] * jv * jv : 0.f; rb[s] = on ? limD[s] * ljar[s] * jv : 0.f;
        }
#pragma unroll
        for (int i = 2; i < 6; i++) { rj[i] = 1.f; rv[i] = 0.f; ra[i] = 0.f; rb[i] = 0.f; }
        if (lane < ncon && !done) {
          float s0 = 0.f, s1 = 0.f, s2 = 0.f;
          const float4* jr = jrow(lane);
#pragma unroll
          for (int k = 0; k < NVP; k++) {
            const float4 j = jr[k];
            const float x = W_(L.o_dir + k);
            s0 = fmaf(j.x, x, s0); s1 = fmaf(j.y, x, s1); s2 = fmaf(j.z, x, s2);
          }
          rv[2] = s0 + s1; rv[3] = s0 - s1; rv[4] = s0 + s2; rv[5] = s0 - s2;
          const int cs = L.o_con + lane * L.cstride;
          const float cD = W_(cs + K_D);
#pragma unroll
          for (int i = 0; i < 4; i++) {
            rj[2 + i] = W_(cs + K_JAR + i);
            ra[2 + i] = cD * rv[2 + i] * rv[2 + i]; rb[2 + i] = cD * rj[2 + i] * rv[2 + i];
          }
        }
        MMZ_STICK(5);  // M dir, rows of the line search
        bool lsdone = done || !constrained;
        bool flipped = false, lsconv = false;  // did a row change sides between 0 and alpha; did the search converge
        {
          // The full Newton step crosses no row: the cost along the direction is ONE quadratic on [0, 1], whose minimiser
          // is the Newton step itself. No search (and no rounding noise that could make it look unconverged).
          bool fl = false;
#pragma unroll
          for (int r = 0; r < 6; r++) fl |= (rj[r] + rv[r] < 0.f) != (rj[r] < 0.f);
          flipped = gballot(fl) != 0;
          if (!flipped && !lsdone) { lsdone = true; lsconv = true; }
        }
        if (!__all_sync(kAll, lsdone)) {
          const bool searched = !lsdone;
          const float g0 = gsum16(me ? dr * (Ma - sm_) : 0.f), h0 = gsum16(me ? dr * md : 0.f);
          float lo = 0.f, hi = -1.f;
#pragma unroll 1
          for (int k = 0; k < kTMaxLineSearch; k++) {
            float g = 0.f, h = 0.f;
#pragma unroll
            for (int r = 0; r < 6; r++) {
              if (rj[r] + alpha * rv[r] < 0.f) { g += rb[r] + alpha * ra[r]; h += ra[r]; }
            }
            g = gsum16(g) + g0 + alpha * h0;
            h = gsum16(h) + h0;
            if (!lsdone) {
              if (fabsf(g) < MMZ_LS_TOL * fmaxf(1e-6f, fabsf(g0))) { lsdone = true; lsconv = true; }
              else {
                if (g < 0.f) lo = alpha; else hi = alpha;
                float next = alpha - g / h;
                if (hi >= 0.f && (next <= lo || next >= hi)) next = 0.5f * (lo + hi);
                if (next <= lo && hi < 0.f) next = 2.f * alpha + 1e-6f;
                if (next == alpha) lsdone = true;
                else { alpha = next; ls++; }
              }
            }
            if (__all_sync(kAll, lsdone)) break;
          }
          {  // rows are linear in alpha: compare the sides at the accepted step with those at 0
            bool fl = false;
#pragma unroll
            for (int r = 0; r < 6; r++) fl |= (rj[r] + alpha * rv[r] < 0.f) != (rj[r] < 0.f);
            const bool any = gballot(fl) != 0;  // (a vote: outside the per-environment condition)
            if (searched) flipped = any;
          }
        }
        // No row of this environment changed sides on [0, alpha] and alpha minimises the cost along the Newton direction
        // of exactly that active set: the new point is the solution, and the gradient pass that would confirm it is
        // skipped.
        exact = !flipped && lsconv && fabsf(alpha - 1.f) < 1e-3f;
        if (exact) alpha = 1.f;  // the minimiser of that quadratic is the Newton step itself
      }
      MMZ_STICK(6);  // line search
      bool moved = false;
      if (me && !done) {
        const float st = alpha * dr;
        moved = fabsf(st) > 2e-6f * fabsf(al) + 1e-6f;
        al += st;
        W_(L.o_qacc + lane) = al;
        Ma += alpha * md;
        Mabs += fabsf(alpha) * mdabs;  // triangle inequality: still an upper bound of the magnitude of the terms
      }
      if (lane == 0 && !done) {
        IW(L.o_cnt + TN_ITER) = it + 1; IW(L.o_cnt + TN_ITER_SUM) += 1; IW(L.o_cnt + TN_LS_SUM) += ls;
        if (it == kTMaxNewton - 1) IW(L.o_cnt + TN_CAPPED) += 1;
      }
      __syncwarp();
      const unsigned movedbits = gballot(moved);
      if (!constrained || movedbits == 0 || exact) done = true;
      MMZ_STICK(7);  // update
#ifdef MMZ_PHASE_TIMING
      if (blockIdx.x == 0 && e == 0) atomicAdd(&g_sec[15], 1ull);
#endif
      if (__all_sync(kAll, done)) break;
    }
    __syncwarp();
#ifdef MMZ_PHASE_TIMING
    if (lane == 0) atomicAdd(&g_iter_hist[min(IW(L.o_cnt + TN_ITER), 15)], 1u);
#endif
  }

  // ================================================================== solver v3 (the box instances)
  // Solver v2's scheme where contacts are many (a pushed block rests on four corners, touches walls with up to eight points)
  // and most of them move few dofs (the block's two slides; an ant leg's chain of eight): the Jacobian of contact c is
  // stored COMPRESSED, one float4 per dof of its mask, at entries [off_c, off_c + popc(mask_c)) of a per-environment pool
  // (off_c = exclusive scan over the contacts, lane = contact). Passes over the contacts (lane = contact, 16 per trip) walk
  // the bits of their mask; passes over the dofs find their entry at rank(lane in mask). The contact part of the Hessian
  // row is accumulated in shared memory (dynamic column index), over the dofs of the mask only, then added to the
  // mass-matrix row in registers. Line-search rows of the contacts stay in their records (any number of contacts).
  MMZ_DI int cmask(int cs) const { return __float_as_int(W_(cs + K_MPOS)) | __float_as_int(W_(cs + K_MNEG)); }
  // Returns whether some contact of this environment moves dofs of BOTH trees of an Ant-and-block model (dofs 0..13 and
  // 14, 15): only then does the Hessian couple them (elim_solve2<3> otherwise).
  // `nnb`: one past the last contact, in either environment of the warp, that moves anything but the two dofs of the
  // tail tree (the block's own contacts come last in the contact order; the Hessian pass handles them lane = contact).
  MMZ_DI bool build_jac3(const TLayout& L, const float (&cd)[6], int ncon, int ncw, int* nnb) {
    int total = 0, last = 0;
    bool cross = false;
    const int tailmask = L.tail0 >= 0 ? 3 << L.tail0 : -1;
#pragma unroll 1
    for (int t0 = 0; t0 < ncw; t0 += 16) {  // offsets: lane = contact
      const int c = t0 + lane, cs = L.o_con + c * L.cstride;
      const bool valid = c < ncon;
      const int cnt = valid ? __popc(cmask(cs)) : 0;
      int incl = cnt;
#pragma unroll
      for (int off = 1; off < 16; off <<= 1) { const int v = __shfl_up_sync(kAll, incl, off, 16); if (lane >= off) incl += v; }
      const int excl = incl - cnt + total;
      // K3_JOFF holds ONE word for the passes below: the contact's dof mask (16 bits) and, above it, the offset of its
      // entries in the pool; 0 for the slots beyond this environment's contacts (the other one of the warp has more)
      if (valid) {
        const bool full = excl + cnt > L.njac;
        if (full) {  // the pool is full: the contact is dropped (and the overflow reported)
          W_(cs + K_MPOS) = __int_as_float(0); W_(cs + K_MNEG) = __int_as_float(0);
          IW(L.o_cnt + TN_OVERFLOW) = 1;
        }
        const int mk = cmask(cs);
        cross |= !full && (mk & 0x3fff) && (mk & 0xc000);
        if (!full && mk != tailmask) last = c + 1;
        IW(cs + K3_JOFF) = full ? 0 : (mk | excl << 16);
      } else if (c < ncw) {
        IW(cs + K3_JOFF) = 0;
      }
      total += __shfl_sync(kAll, incl, 15, 16);
    }
    __syncwarp();
    const int nn = __reduce_max_sync(kAll, last);
    *nnb = nn;
#pragma unroll 1
    for (int c = 0; c < nn; c++) {  // entries: lane = dof
      const int cs = L.o_con + c * L.cstride;
      const int mp = __float_as_int(W_(cs + K_MPOS)), mn = __float_as_int(W_(cs + K_MNEG));
      const int mask = (mp | mn) == tailmask ? 0 : (mp | mn);  // (tail-only contacts: below)
      const float s = (float)(mp >> lane & 1) - (float)(mn >> lane & 1);
      float p[3], fr[9], wxp[3];
#pragma unroll
      for (int k = 0; k < 3; k++) p[k] = W_(cs + K_POS + k);
#pragma unroll
      for (int k = 0; k < 6; k++) fr[k] = W_(cs + K_N + k);
      cross3(fr + 6, fr, fr + 3);
      cross3(wxp, cd, p);
      const float v[3] = {s * (cd[3] + wxp[0]), s * (cd[4] + wxp[1]), s * (cd[5] + wxp[2])};
      const float mu = W_(cs + K_MU);
      const float jn = dot3(fr, v), jt1 = mu * dot3(fr + 3, v), jt2 = mu * dot3(fr + 6, v);
      if (c < ncon && (mask >> lane & 1))
        jg[(IW(cs + K3_JOFF) >> 16) + __popc(mask & ((1 << lane) - 1))] = make_float4(jn, jt1, jt2, fabsf(jn) + fabsf(jt1) + fabsf(jt2));
    }
    if (L.tail0 >= 0) {  // the two entries of every tail-only contact, lane = contact (the motion axes of the two dofs by shuffle)
      float ca[6], cb[6];
#pragma unroll
      for (int k = 0; k < 6; k++) { ca[k] = __shfl_sync(kAll, cd[k], L.tail0, 16); cb[k] = __shfl_sync(kAll, cd[k], L.tail0 + 1, 16); }
#pragma unroll 1
      for (int t0 = 0; t0 < ncw; t0 += 16) {
        const int c = t0 + lane, cs = L.o_con + c * L.cstride;
        if (c >= ncon) continue;
        const int mp = __float_as_int(W_(cs + K_MPOS)), mn = __float_as_int(W_(cs + K_MNEG));
        if ((mp | mn) != tailmask) continue;
        float p[3], fr[9];
#pragma unroll
        for (int k = 0; k < 3; k++) p[k] = W_(cs + K_POS + k);
#pragma unroll
        for (int k = 0; k < 6; k++) fr[k] = W_(cs + K_N + k);
        cross3(fr + 6, fr, fr + 3);
        const float mu = W_(cs + K_MU);
        float4* out = jg + (IW(cs + K3_JOFF) >> 16);
#pragma unroll
        for (int w = 0; w < 2; w++) {
          const float* cw = w == 0 ? ca : cb;
          const int d = L.tail0 + w;
          const float s = (float)(mp >> d & 1) - (float)(mn >> d & 1);
          float wxp[3];
          cross3(wxp, cw, p);
          const float v[3] = {s * (cw[3] + wxp[0]), s * (cw[4] + wxp[1]), s * (cw[5] + wxp[2])};
          const float jn = dot3(fr, v), jt1 = mu * dot3(fr + 3, v), jt2 = mu * dot3(fr + 6, v);
          out[w] = make_float4(jn, jt1, jt2, fabsf(jn) + fabsf(jt1) + fabsf(jt2));
        }
      }
    }
    return gballot(cross) != 0;
  }
  // lane = contact (16 per trip): J x over the dofs of the contact's mask. which 0: x = qacc -> J a - aref, force and
  // Hessian weights of the active rows; which 1: x = direction -> J dir per pyramid row into the record
  // (which 1: the rows of the first 16 contacts stay in the lane, `jvreg`, for the line search; later ones go to the records)
  MMZ_DI void contact_pass3(const TLayout& L, int xoff, int ncon, int ncw, int which, float* jvreg = nullptr) {
#pragma unroll 1
    for (int t0 = 0; t0 < ncw; t0 += 16) {
      const int c = t0 + lane;
      if (c >= ncw) continue;
      float4 F = make_float4(0.f, 0.f, 0.f, 0.f), Wt = F;
      if (c < ncon) {
        const int cs = L.o_con + c * L.cstride;
        const int mj = IW(cs + K3_JOFF);
        const float4* jr = jg + (mj >> 16);
        unsigned bits = mj & 0xffff;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, sa = 0.f;
        if (NVP >= 6 && (L.topo & 2) && (bits & 0x3fu) == 0x3fu) {  // the six dofs of the free root: entries 0..5, no bit walk
#pragma unroll
          for (int k = 0; k < 6; k++) {
            const float4 j = jr[k];
            const float x = W_(xoff + k);
            s0 = fmaf(j.x, x, s0); s1 = fmaf(j.y, x, s1); s2 = fmaf(j.z, x, s2); sa = fmaf(j.w, fabsf(x), sa);
          }
          jr += 6; bits &= ~0x3fu;
        }
#pragma unroll 1
        while (bits) {  // two entries per trip (their loads in flight together; same order of additions as one by one)
          const int k0 = __ffs(bits) - 1;
          bits &= bits - 1;
          const bool two = bits != 0;
          const int k1 = two ? __ffs(bits) - 1 : k0;
          bits &= bits - 1;
          const float4 j = jr[0], j1 = jr[1];  // (jr[1] may belong to the next contact: unused then)
          const float x = W_(xoff + k0), x1 = W_(xoff + k1);
          jr += 2;
          s0 = fmaf(j.x, x, s0); s1 = fmaf(j.y, x, s1); s2 = fmaf(j.z, x, s2); sa = fmaf(j.w, fabsf(x), sa);
          if (two) { s0 = fmaf(j1.x, x1, s0); s1 = fmaf(j1.y, x1, s1); s2 = fmaf(j1.z, x1, s2); sa = fmaf(j1.w, fabsf(x1), sa); }
        }
        if (which == 1) {
          if (t0 == 0) { jvreg[0] = s0 + s1; jvreg[1] = s0 - s1; jvreg[2] = s0 + s2; jvreg[3] = s0 - s2; }
          else { W_(cs + K3_JV) = s0 + s1; W_(cs + K3_JV + 1) = s0 - s1; W_(cs + K3_JV + 2) = s0 + s2; W_(cs + K3_JV + 3) = s0 - s2; }
          continue;
        }
        const float r0 = W_(cs + K_AREF), r1 = W_(cs + K_AREF + 1), r2 = W_(cs + K_AREF + 2), r3 = W_(cs + K_AREF + 3);
        const float D = W_(cs + K_D);
        const float j0 = s0 + s1 - r0, j1 = s0 - s1 - r1, j2 = s0 + s2 - r2, j3 = s0 - s2 - r3;
        W_(cs + K_JAR) = j0; W_(cs + K_JAR + 1) = j1; W_(cs + K_JAR + 2) = j2; W_(cs + K_JAR + 3) = j3;
        const float a0 = j0 < 0.f, a1 = j1 < 0.f, a2 = j2 < 0.f, a3 = j3 < 0.f;
        const float f0 = a0 * D * j0, f1 = a1 * D * j1, f2 = a2 * D * j2, f3 = a3 * D * j3;
        const float bound = sa + fmaxf(fmaxf(fabsf(r0), fabsf(r1)), fmaxf(fabsf(r2), fabsf(r3)));
        F = make_float4(f0 + f1 + f2 + f3, f0 - f1, f2 - f3, D * bound * (a0 + a1 + a2 + a3));
        Wt = make_float4(D * (a0 - a1), D * (a2 - a3), D * (a0 + a1), D * (a2 + a3));
      }
      if (which == 0) { fg[2 * c] = F; fg[2 * c + 1] = Wt; }
    }
  }
  // derivative and curvature of the cost along the direction from the rows of this lane: its two joint-limit rows
  // (registers) and the pyramid rows r = lane, lane + 16, ... of the contacts (records); `fl`: a row changed sides
  // The four pyramid rows of contact `lane` (the first 16 contacts) come from registers: (rj, rv, rD) = (J a - aref, J dir, D),
  // (1, 0, 0) for a lane without a contact.
  MMZ_DI void ls_rows3(const TLayout& L, int ncon, const float (&ljar)[2], float dr, float alpha, const float (&rj)[4],
                       const float (&rv)[4], float rD, float* g, float* h, bool* fl) const {
    float gg = 0.f, hh = 0.f;
    bool f = false;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const float x = rj[i] + alpha * rv[i];
      if (x < 0.f) { gg += rD * x * rv[i]; hh += rD * rv[i] * rv[i]; }
      f |= (x < 0.f) != (rj[i] < 0.f);
    }
#pragma unroll
    for (int s = 0; s < 2; s++) {
      const float jv = s == 0 ? dr : -dr, x = ljar[s] + alpha * jv;
      if (limD[s] > 0.f && x < 0.f) { gg += limD[s] * x * jv; hh += limD[s] * jv * jv; }
      f |= limD[s] > 0.f && (x < 0.f) != (ljar[s] < 0.f);
    }
#pragma unroll 1
    for (int r = 64 + lane; r < 4 * ncon; r += 16) {
      const int cs = L.o_con + (r >> 2) * L.cstride;
      const float jv = W_(cs + K3_JV + (r & 3)), jar = W_(cs + K_JAR + (r & 3)), x = jar + alpha * jv, D = W_(cs + K_D);
      if (x < 0.f) { gg += D * x * jv; hh += D * jv * jv; }
      f |= (x < 0.f) != (jar < 0.f);
    }
    *g = gg; *h = hh; *fl = f;
  }
  MMZ_DI void solve_g3(const TLayout& L, bool warmstart) {
    const int nv = L.nv;
    const bool me = lane < nv;
#ifdef MMZ_PHASE_TIMING
    long long stick_ = clock64();
#endif
    const unsigned limbits = gballot(limD[0] > 0.f || limD[1] > 0.f);
    float mrow[NVP];  // this lane's row of the mass matrix (zero outside its sparsity pattern and outside the model)
    const int rel = me ? dv->dof_rel[lane] : 0;
#pragma unroll
    for (int k = 0; k < NVP; k++) mrow[k] = (rel >> k & 1) ? W_(L.o_M + lane * L.ldm + k) : 0.f;
    const float sm_ = me ? W_(L.o_smooth + lane) : 0.f;
    int ncon, ncw, nnb = 0;
    bool coupled = true;
    {
      float cd[6];
#pragma unroll
      for (int k = 0; k < 6; k++) cd[k] = me ? W_(L.o_cdof + 6 * lane + k) : 0.f;
      // the natural area overlays the mass matrix, the motion axes and every other array the solver view has read by now
      __syncthreads();
      ncon = IW(L.o_cnt + TN_CON);
      ncw = max(ncon, __shfl_xor_sync(kAll, ncon, 16));  // the two environments of the warp
#pragma unroll
      for (int k = 0; k < NVP; k++) hg[k] = 0.f;
      coupled = build_jac3(L, cd, ncon, ncw, &nnb);
    }
    // the Ant's tree and the block's chain eliminate side by side unless a contact couples them in either environment
    const bool sparse = NVP == 16 && (L.topo & 4) && !__any_sync(kAll, coupled);
    const bool constrained = ncon > 0 || limbits != 0;
    float al = (warmstart && me) ? W_(L.o_qacc + lane) : 0.f;
    if (!(fabsf(al) < kMaxVal)) al = 0.f;  // a blown-up environment restarts from zero
    if (me) W_(L.o_qacc + lane) = al;
    const int nlim = __popc(gballot(limD[0] > 0.f)) + __popc(gballot(limD[1] > 0.f));
    if (lane == 0) {
      IW(L.o_cnt + TN_ITER) = 0; IW(L.o_cnt + TN_LIM) = nlim;
      IW(L.o_cnt + TN_CON_MAX) = max(IW(L.o_cnt + TN_CON_MAX), ncon);
    }
    __syncwarp();
    bool done = false;
    float Ma = 0.f, Mabs = 0.f;  // (M qacc)[lane] and the magnitude of its terms, updated with every step
#pragma unroll
    for (int k = 0; k < NVP; k++) { const float t = mrow[k] * W_(L.o_qacc + k); Ma += t; Mabs += fabsf(t); }
    const unsigned lt = (1u << lane) - 1u;
    MMZ_STICK(0);
#pragma unroll 1
    for (int it = 0; it < kTMaxNewton; it++) {
      float grad = Ma - sm_, dadd = 0.f;
      float mag = Mabs + fabsf(sm_);
      float ljar[2];
#pragma unroll
      for (int s = 0; s < 2; s++) {
        const float sign = s == 0 ? 1.f : -1.f;
        ljar[s] = sign * al - limA[s];
        if (limD[s] > 0.f && ljar[s] < 0.f) {
          grad += limD[s] * ljar[s] * sign;
          mag += limD[s] * (fabsf(al) + fabsf(limA[s]));
          dadd += limD[s];
        }
      }
      contact_pass3(L, L.o_qacc, ncon, ncw, 0);
      __syncwarp();
      // The contacts that move only the two dofs (a, b) of the tail tree - a block's four corners on the floor, its points
      // on the walls - with lane = CONTACT, all at once: each lane forms J^T f and the 2 x 2 block J^T W J of its contact,
      // seven sums over the lanes, and the lanes of a and b take them (the Hessian part is added to their rows below).
      const unsigned tailmask = L.tail0 >= 0 ? 3u << L.tail0 : 0u;
      float tH0 = 0.f, tH1 = 0.f;
      if (L.tail0 >= 0) {
        float ga = 0.f, gb = 0.f, ma = 0.f, mb = 0.f, haa = 0.f, hab = 0.f, hbb = 0.f;
        bool any = false;
#pragma unroll 1
        for (int t0 = 0; t0 < ncw; t0 += 16) {
          const int c = t0 + lane;
          if (c < ncon) {
            const unsigned mj = (unsigned)IW(L.o_con + c * L.cstride + K3_JOFF);
            if ((mj & 0xffffu) == tailmask) {
              const float4 F = fg[2 * c], Wt = fg[2 * c + 1], ja = jg[mj >> 16], jb = jg[(mj >> 16) + 1];
              const float wnn = Wt.z + Wt.w;
              const float a0 = wnn * ja.x + Wt.x * ja.y + Wt.y * ja.z, a1 = Wt.x * ja.x + Wt.z * ja.y, a2 = Wt.y * ja.x + Wt.w * ja.z;
              const float b0 = wnn * jb.x + Wt.x * jb.y + Wt.y * jb.z, b1 = Wt.x * jb.x + Wt.z * jb.y, b2 = Wt.y * jb.x + Wt.w * jb.z;
              haa += a0 * ja.x + a1 * ja.y + a2 * ja.z;
              hab += a0 * jb.x + a1 * jb.y + a2 * jb.z;
              hbb += b0 * jb.x + b1 * jb.y + b2 * jb.z;
              ga += ja.x * F.x + ja.y * F.y + ja.z * F.z; ma = fmaf(ja.w, F.w, ma);
              gb += jb.x * F.x + jb.y * F.y + jb.z * F.z; mb = fmaf(jb.w, F.w, mb);
              any = true;
            }
          }
        }
        if (__any_sync(kAll, any)) {
          ga = gsum16(ga); gb = gsum16(gb); ma = gsum16(ma); mb = gsum16(mb);
          haa = gsum16(haa); hab = gsum16(hab); hbb = gsum16(hbb);
          if (lane == L.tail0) { grad += ga; mag += ma; tH0 = haa; tH1 = hab; }
          if (lane == L.tail0 + 1) { grad += gb; mag += mb; tH0 = hab; tH1 = hbb; }
        }
      }
      MMZ_STICK(1);
#pragma unroll 1
      for (int c = 0; c < nnb; c += 2) {  // gradient J^T f: this lane's entry of every contact that moves its dof (two per trip)
        const int cs = L.o_con + c * L.cstride;
        const unsigned mj = (unsigned)IW(cs + K3_JOFF), mj1 = c + 1 < nnb ? (unsigned)IW(cs + L.cstride + K3_JOFF) : 0u;
        const bool in0 = (mj >> lane & 1) && (mj & 0xffffu) != tailmask, in1 = (mj1 >> lane & 1) && (mj1 & 0xffffu) != tailmask;
        float4 j = make_float4(0.f, 0.f, 0.f, 0.f), F = j, j1 = j, F1 = j;
        if (in0) { j = jg[(mj >> 16) + __popc(mj & lt)]; F = fg[2 * c]; }
        if (in1) { j1 = jg[(mj1 >> 16) + __popc(mj1 & lt)]; F1 = fg[2 * c + 2]; }
        if (in0) { grad += j.x * F.x + j.y * F.y + j.z * F.z; mag = fmaf(j.w, F.w, mag); }
        if (in1) { grad += j1.x * F1.x + j1.y * F1.y + j1.z * F1.z; mag = fmaf(j1.w, F1.w, mag); }
      }
      if (gballot(me && fabsf(grad) > tol * mag + 1e-30f) == 0) done = true;
      MMZ_STICK(2);
      if (__all_sync(kAll, done)) break;
      // Contact part of the Hessian row over the dofs of each contact's mask. The six dofs of a free root (when the model
      // has one, every mask holds all six or none: they are its entries 0..5) go to REGISTERS, whatever else the contact
      // moves - a leg's hip and ankle, a block's slides - to the row in shared memory (dynamic column), two per trip.
      constexpr int NR = NVP >= 6 ? 6 : 1;
      float hacc[NR];
#pragma unroll
      for (int k = 0; k < NR; k++) hacc[k] = 0.f;
      if (L.tail0 >= 0 && (lane == L.tail0 || lane == L.tail0 + 1)) { hg[L.tail0] += tH0; hg[L.tail0 + 1] += tH1; }  // (the tail pass above)
#pragma unroll 1
      for (int c = 0; c < nnb; c++) {
        const float4 Wt = fg[2 * c + 1];
        const unsigned mj = (unsigned)IW(L.o_con + c * L.cstride + K3_JOFF);
        const float wnn = Wt.z + Wt.w;
        unsigned bits = (wnn != 0.f && (mj & 0xffffu) != tailmask) ? mj & 0xffffu : 0u;
        if (!__any_sync(kAll, bits != 0u)) continue;  // no active row in this contact, in either environment
        if (bits >> lane & 1) {
          const float4* jr = jg + (mj >> 16);
          const float4 j = jr[__popc(bits & lt)];
          const float u0 = wnn * j.x + Wt.x * j.y + Wt.y * j.z, u1 = Wt.x * j.x + Wt.z * j.y, u2 = Wt.y * j.x + Wt.w * j.z;
          if (NVP >= 6 && (L.topo & 2) && (bits & 0x3fu) == 0x3fu) {
#pragma unroll
            for (int k = 0; k < NR; k++) {
              const float4 jk = jr[k];
              hacc[k] += u0 * jk.x + u1 * jk.y + u2 * jk.z;
            }
            jr += 6; bits &= ~0x3fu;
          }
#pragma unroll 1
          while (bits) {
            const int k0 = __ffs(bits) - 1;
            bits &= bits - 1;
            const bool two = bits != 0;
            const int k1 = two ? __ffs(bits) - 1 : k0;
            bits &= bits - 1;
            const float4 jk = jr[0], jl = jr[1];  // (jr[1] may belong to the next contact: unused then)
            const float h0 = hg[k0], h1 = hg[k1];
            jr += 2;
            hg[k0] = h0 + (u0 * jk.x + u1 * jk.y + u2 * jk.z);
            if (two) hg[k1] = h1 + (u0 * jl.x + u1 * jl.y + u2 * jl.z);
          }
        }
      }
      float hrow[NVP];
#pragma unroll
      for (int k = 0; k < NVP; k++) { hrow[k] = mrow[k] + (k < NR && NVP >= 6 ? hg[k] + hacc[k] : hg[k]); hg[k] = 0.f; }
      MMZ_STICK(3);
      const float dg = me ? dadd : 1.f, rhs0 = me ? -grad : 0.f;
      float dr;
      if constexpr (NVP == 16) dr = sparse ? elim_solve2<3>(hrow, rhs0, dg) : elim_solve2<0>(hrow, rhs0, dg);
      else dr = elim_solve2<0>(hrow, rhs0, dg);
      if (me && !done) W_(L.o_dir + lane) = dr;
      __syncwarp();
      MMZ_STICK(4);
      float alpha = 1.f;
      int ls = 0;
      bool exact = false;
      float md = 0.f, mdabs = 0.f;  // (M dir)[lane] and the magnitude of its terms
#pragma unroll
      for (int k = 0; k < NVP; k++) { const float t = mrow[k] * W_(L.o_dir + k); md += t; mdabs += fabsf(t); }
      if (__any_sync(kAll, constrained && !done)) {
        const int nls = done ? 0 : ncon;
        float rj[4] = {1.f, 1.f, 1.f, 1.f}, rv[4] = {0.f, 0.f, 0.f, 0.f}, rD = 0.f;
        contact_pass3(L, L.o_dir, nls, ncw, 1, rv);
        if (lane < nls) {
          const int cs = L.o_con + lane * L.cstride;
#pragma unroll
          for (int i = 0; i < 4; i++) rj[i] = W_(cs + K_JAR + i);
          rD = W_(cs + K_D);
        }
        __syncwarp();
        bool lsdone = done || !constrained;
        bool flipped = false, lsconv = false;  // did a row change sides between 0 and alpha; did the search converge
        float g, h;
        bool fl;
        ls_rows3(L, nls, ljar, dr, 1.f, rj, rv, rD, &g, &h, &fl);
        // the full Newton step crosses no row: it is the minimiser along the direction, no search needed
        flipped = gballot(fl) != 0;
        if (!flipped && !lsdone) { lsdone = true; lsconv = true; }
        MMZ_STICK(5);
        if (!__all_sync(kAll, lsdone)) {
          const bool searched = !lsdone;
          const float g0 = gsum16(me ? dr * (Ma - sm_) : 0.f), h0 = gsum16(me ? dr * md : 0.f);
          float lo = 0.f, hi = -1.f;
#pragma unroll 1
          for (int k = 0; k < kTMaxLineSearch; k++) {
            if (k > 0) ls_rows3(L, nls, ljar, dr, alpha, rj, rv, rD, &g, &h, &fl);
            g = gsum16(g) + g0 + alpha * h0;
            h = gsum16(h) + h0;
            if (!lsdone) {
              if (fabsf(g) < MMZ_LS_TOL * fmaxf(1e-6f, fabsf(g0))) { lsdone = true; lsconv = true; }
              else {
                if (g < 0.f) lo = alpha; else hi = alpha;
                float next = alpha - g / h;
                if (hi >= 0.f && (next <= lo || next >= hi)) next = 0.5f * (lo + hi);
                if (next <= lo && hi < 0.f) next = 2.f * alpha + 1e-6f;
                if (next == alpha) lsdone = true;
                else { alpha = next; ls++; }
              }
            }
            if (__all_sync(kAll, lsdone)) break;
          }
          ls_rows3(L, nls, ljar, dr, alpha, rj, rv, rD, &g, &h, &fl);  // sides at the accepted step against those at 0
          const bool any = gballot(fl) != 0;
          if (searched) flipped = any;
        }
        exact = !flipped && lsconv && fabsf(alpha - 1.f) < 1e-3f;
        if (exact) alpha = 1.f;  // the minimiser of that quadratic is the Newton step itself
      }
      MMZ_STICK(6);
      bool moved = false;
      if (me && !done) {
        const float st = alpha * dr;
        moved = fabsf(st) > 2e-6f * fabsf(al) + 1e-6f;
        al += st;
        W_(L.o_qacc + lane) = al;
        Ma += alpha * md;
        Mabs += fabsf(alpha) * mdabs;  // triangle inequality: still an upper bound of the magnitude of the terms
      }
      if (lane == 0 && !done) {
        IW(L.o_cnt + TN_ITER) = it + 1; IW(L.o_cnt + TN_ITER_SUM) += 1; IW(L.o_cnt + TN_LS_SUM) += ls;
        if (it == kTMaxNewton - 1) IW(L.o_cnt + TN_CAPPED) += 1;
      }
      __syncwarp();
      const unsigned movedbits = gballot(moved);
      if (!constrained || movedbits == 0 || exact) done = true;
      MMZ_STICK(7);
#ifdef MMZ_PHASE_TIMING
      if (blockIdx.x == 0 && e == 0) atomicAdd(&g_sec[15], 1ull);
#endif
      if (__all_sync(kAll, done)) break;
    }
    __syncwarp();
  }

  // RK4 bookkeeping of stage i for this warp's two environments (solver view, lane = dof), right after their solve:
  // accumulate the stage (B weights), then move to the state of the next stage, or to the final combination after
  // stage 3 (classic tableau, A = 1/2, 1/2, 1). Positions integrate on the configuration manifold (mj_integratePos):
  // free-joint quaternion <- q (x) exp(h w / 2). A non-finite acceleration counts as zero and marks the environment.
  MMZ_DI void rk_update_g(const TLayout& L, int i) {
    const float hstep = m->timestep;
    const float B = (i == 0 || i == 3) ? (1.f / 6.f) : (1.f / 3.f), A = (i == 0 || i == 1) ? 0.5f : 1.f;
    const bool me = lane < L.nv;
    float f = me ? W_(L.o_qacc + lane) : 0.f;
    const bool badacc = gballot(me && !(fabsf(f) < kMaxVal)) != 0;
    if (badacc) f = 0.f;
    const float v = me ? W_(L.o_qvel + lane) : 0.f;
    float av = 0.f, aa = 0.f;
    bool badstate = false;  // mj_checkPos / mj_checkVel of the state this step ends in (stage 3 only)
    if (me) {
      av = W_(L.o_accv + lane) + B * v; aa = W_(L.o_acca + lane) + B * f;
      W_(L.o_accv + lane) = av; W_(L.o_acca + lane) = aa;
      W_(L.o_dir + lane) = (i == 3) ? av : v;  // the velocity that moves the positions (o_dir is free after the solve)
      const float vnew = W_(L.o_v0 + lane) + hstep * A * ((i == 3) ? aa : f);
      W_(L.o_qvel + lane) = vnew;
      badstate = !(fabsf(vnew) < kMaxVal);
    }
    __syncwarp();
#pragma unroll 1
    for (int j = lane; j < L.nj; j += 16) {
      const int qa = m->jnt_qadr[j], d = m->jnt_dadr[j];
      if (m->jnt_type[j] == MMZ_JNT_FREE) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
          const float pk = W_(L.o_q0 + qa + k) + hstep * A * W_(L.o_dir + d + k);
          W_(L.o_qpos + qa + k) = pk;
          badstate |= !(fabsf(pk) < kMaxVal);
        }
        float wv[3] = {A * W_(L.o_dir + d + 3), A * W_(L.o_dir + d + 4), A * W_(L.o_dir + d + 5)};
        float q[4] = {W_(L.o_q0 + qa + 3), W_(L.o_q0 + qa + 4), W_(L.o_q0 + qa + 5), W_(L.o_q0 + qa + 6)};
        const float nw = norm3(wv), ang = hstep * nw;
        quat_norm(q);
        if (ang > 0.f) {
          const float inv = 1.f / nw;
          float ax[3] = {wv[0] * inv, wv[1] * inv, wv[2] * inv}, qr[4], q2[4];
          axisangle2quat(qr, ax, ang);
          quat_mul(q2, q, qr);
#pragma unroll
          for (int k = 0; k < 4; k++) q[k] = q2[k];
        }
#pragma unroll
        for (int k = 0; k < 4; k++) { W_(L.o_qpos + qa + 3 + k) = q[k]; badstate |= !(fabsf(q[k]) < kMaxVal); }
      } else {
        const float pk = W_(L.o_q0 + qa) + hstep * A * W_(L.o_dir + d);
        W_(L.o_qpos + qa) = pk;
        badstate |= !(fabsf(pk) < kMaxVal);
      }
    }
    // the flag of this mj_step: assigned by stage 0 (every warp read the previous step's flag long before), or-ed later
    const bool flag = badacc || (i == 3 && gballot(badstate) != 0);
    if (lane == 0 && (i == 0 || flag)) IW(L.o_cnt + TN_BAD) = flag ? 1 : 0;
    __syncwarp();
  }

  // ------------------------------------------------------------------ mj_forward
  MMZ_DI void forward(const TLayout& L, bool warmstart, int rk_stage = -1) {
#ifdef MMZ_PHASE_TIMING
    long long tick_ = clock64();
    if (threadIdx.x == 0) { bmaxit = 0; }
#endif
    // A: the kinematic trees. Pass 0: pose and motion axes of the roots (and, on the idle warps, the joint-limit rows).
    // Pass 1: the roots' warps compute the roots' velocities (published through a named barrier), then their inertial forces; the subtree of every level-1 body is walked by a PAIR of warps: one runs pose +
    // velocity down the chain - the only dependent work - and the inertial force of the LAST body; its partner follows one
    // body behind with the inertial forces of the others (named barrier per pair), while the roots' warps do the roots'
    // inertial forces. Critical path for the Ant: 2 short kinematic steps + 1 inertial force, instead of 3 full body passes.
    // (Models with more chains than warp pairs walk every chain in one warp: A_BOTH.)
    {
      const int kind1 = dv->walk_kind[wid], bar = dv->walk_bar[wid];
#pragma unroll 1
      for (int pass = 0; pass < 2; pass++) {
        const bool legacy = dv->walk_root_count == 0;
        const int kind = pass == 0 ? (legacy ? A_BOTH : A_KIN) : kind1;
        const int32_t* list = (pass == 0 || kind == A_ROOTDYN) ? dv->lvl_body : dv->chain_body;
        const bool roots = pass == 0 || kind == A_ROOTDYN;
        const int i0 = roots ? wid : dv->chain_off[wid], i1 = roots ? dv->lvl_off[1] : dv->chain_off[wid + 1];
        if (kind != A_IDLE) {
#pragma unroll 1
          for (int i = i0; i < i1; i += roots ? TW : 1) {
            const int b = list[i];
            const bool last = !roots && i == i1 - 1;
            if (kind == A_KIN || kind == A_PAIR_KIN || kind == A_BOTH) body_kin(L, b);
            // the roots' velocities are computed in pass 1, beside the first kinematic step of the chains that wait for them
            if (kind == A_ROOTDYN) { body_vel(L, b); if (i + TW >= i1) named_arrive(15, dv->walk_root_count); }
            if (kind == A_PAIR_KIN && i == i0) named_sync(15, dv->walk_root_count);
            if (kind == A_PAIR_KIN || kind == A_BOTH) body_vel(L, b);
            if ((kind == A_PAIR_KIN || kind == A_PAIR_DYN) && !last) named_sync(bar, 64);
            if (kind == A_ROOTDYN || kind == A_BOTH || (kind == A_PAIR_DYN && !last) || (kind == A_PAIR_KIN && last)) body_frc(L, b);
          }
        }
        if (pass == 0)  // in the shadow of the roots' kinematics (the last warps first: the roots are on the first ones)
          for (int d = TW - 1 - wid; d < L.nv; d += TW) limit_rows_t(L, d);
        __syncthreads();
        MMZ_TICK(pass);
      }
    }
    // B: geom poses, composite inertias, subtree forces
    if (wid == TW - 1) I(L.o_cnt + TN_OVERFLOW) = 0;
    {
      const int nt = L.ng + 2 * L.nb;
      for (int t = wid; t < nt; t += TW) {
        if (t < L.ng) geom_pose(L, t);
        else if (t < L.ng + L.nb) subtree_sum(L, t - L.ng, L.o_iw, L.o_ic, 10);
        else subtree_sum(L, t - L.ng - L.nb, L.o_frc, L.o_fsub, 6);
      }
    }
    __syncthreads();
    MMZ_TICK(2);
    // C: contact counting, mass matrix rows, smooth forces. Box candidate items keep their contacts for phase D.
    const int nit = n_items(L);
#ifdef MMZ_PHASE_TIMING
    const long long wc0_ = clock64();
#endif
    constexpr int KEEP = BOX ? 3 : 1;
    BoxFound kept[KEEP];
    int nkept = 0;
    {
#pragma unroll 1
      for (int i = dv->c_off[wid]; i < dv->c_off[wid + 1]; i++) {
        const int t = dv->c_item[i];
        if (t < nit) {
          if (BOX && t >= L.ng && nkept < KEEP) {
            box_item_find(L, t, kept[nkept]);
            I(L.o_gcnt + t) = kept[nkept].n;
            nkept++;
          } else {
            collide_item(L, t, 0);
          }
        }
        else if (t < nit + L.nv) mass_row(L, t - nit);
        else smooth_dof(L, t - nit - L.nv);
      }
    }
#ifdef MMZ_PHASE_TIMING
    if (blockIdx.x == 0 && e == 0) g_wC[wid] += clock64() - wc0_;
#endif
    __syncthreads();
    MMZ_TICK(3);
#ifdef MMZ_PHASE_TIMING
    const long long wd0_ = clock64();
#endif
    // D: contacts into their slots (over the slots of arrays that are dead by now, see the layout): every warp writes
    // the items it counted - the kept box items as they are, the others through a second narrow phase
    if (!BOX) {
      for (int t = wid; t < nit; t += TW) collide_item(L, t, 1);
    } else {
      int ik = 0;
#pragma unroll 1
      for (int i = dv->c_off[wid]; i < dv->c_off[wid + 1]; i++) {
        const int t = dv->c_item[i];
        if (t >= nit) continue;
        if (BOX && t >= L.ng && ik < KEEP) {
          if (__any_sync(kAll, kept[ik].n > 0)) box_item_write(L, t, item_base(L, t), kept[ik]);
          ik++;
        } else {
          collide_item(L, t, 1);
        }
      }
    }
#ifdef MMZ_PHASE_TIMING
    if (blockIdx.x == 0 && e == 0) g_wD[wid] += clock64() - wd0_;
#endif
    if (wid == TW - 1) {
      int n = 0;
#pragma unroll 1
      for (int k = 0; k < nit; k++) n += I(L.o_gcnt + k) & 31;
      if (n > L.maxcon) I(L.o_cnt + TN_OVERFLOW) = 1;
      I(L.o_cnt + TN_CON) = min(n, L.maxcon);
    }
    if (kRowsInD) {
      // No barrier here: the contact rows were computed by the warps that wrote the records. The solver view first loads
      // its registers - mass matrix, smooth forces, motion axes, all complete since the barrier that ended phase C - and
      // then passes ITS block barrier before it reads the contact count and the records or touches the Jacobian area.
      MMZ_TICK(4);
    } else {
      __syncthreads();
      MMZ_TICK(4);
      // E: contact rows (the solver's own block barrier orders them against their readers)
      const int ncmax = __reduce_max_sync(kAll, I(L.o_cnt + TN_CON));
      for (int c = wid; c < ncmax; c += TW)
        if (c < I(L.o_cnt + TN_CON)) contact_rows2(L, c);
      MMZ_TICK(5);
    }
    // solver view: warp w owns environments w and w + 16
    limit_rows_g(L);
    if (V2) solve_g2(L, warmstart);
    else solve_g3(L, warmstart);
    if (rk_stage >= 0) rk_update_g(L, rk_stage);  // in the shadow of the wait for the slowest solve of the block
#ifdef MMZ_PHASE_TIMING
    const long long ts1_ = clock64();
    __syncthreads();
    if (blockIdx.x == 0 && e == 0) { g_wsolve[wid] += ts1_ - tick_; g_wwait[wid] += clock64() - ts1_; }
    MMZ_TICK(6);
    if (threadIdx.x == 0) { int mx = 0; for (int k = 0; k < 32; k++) mx = max(mx, reinterpret_cast<int*>(sm)[(L.o_cnt + TN_ITER) * HS + k]); bphase[7] += mx; }
#else
    __syncthreads();
#endif
  }

  // ------------------------------------------------------------------ state checks
  MMZ_DI bool state_bad(const TLayout& L) const {  // every thread evaluates its environment (cheap)
    bool bad = false;
#pragma unroll 1
    for (int i = 0; i < L.nq; i++) bad |= !(fabsf(S(L.o_qpos + i)) < kMaxVal);
#pragma unroll 1
    for (int i = 0; i < L.nv; i++) bad |= !(fabsf(S(L.o_qvel + i)) < kMaxVal);
    return bad;
  }

  // ------------------------------------------------------------------ mj_step, RK4 (mj_RungeKutta)
  // `dead`: the environment already blew up in this env-step; it is parked at qpos0 and keeps running.
  // `bad`: the state is not finite (the caller's state_bad before the first step, then the flag the previous step's
  // stage updates left) or the environment already blew up in this env-step: it is parked at qpos0 and keeps running.
  // No thread may still be reading the state when this is entered (the caller's barrier / the post-solve barrier).
  MMZ_DI bool mj_step(const TLayout& L, bool bad) {
    const int nq = L.nq, nv = L.nv;
    for (int i = wid; i < nq; i += TW) {
      const float q = bad ? m->qpos0[i] : S(L.o_qpos + i);
      if (bad) S(L.o_qpos + i) = q;
      S(L.o_q0 + i) = q;
    }
    for (int d = wid; d < nv; d += TW) {
      const float v = bad ? 0.f : S(L.o_qvel + d);
      if (bad) { S(L.o_qvel + d) = 0.f; S(L.o_qacc + d) = 0.f; }
      S(L.o_v0 + d) = v; S(L.o_accv + d) = 0.f; S(L.o_acca + d) = 0.f;
    }
    __syncthreads();
#pragma unroll 1
    for (int i = 0; i < 4; i++) forward(L, true, i);  // each evaluation ends with its RK4 stage update (rk_update_g)
    // a non-finite acceleration in some stage, or a non-finite state after the last one (written by rk_update_g before
    // the barrier that ends the evaluation). Derived arrays (xpos, contacts) deliberately stay at the 4th-stage
    // state: SURVEY quirk Q15
    return bad || I(L.o_cnt + TN_BAD) != 0;
  }
#undef S
#undef W_
};

static __device__ __noinline__ void write_contact_record2(float* c, const RawContact& rc, int b1, int b2, float iw, int g, int other) {
  float fr[9];
#pragma unroll
  for (int k = 0; k < 3; k++) { fr[k] = rc.normal[k]; fr[3 + k] = rc.hint[k]; }
  make_frame(fr);
#pragma unroll
  for (int k = 0; k < 3; k++) c[(K_POS + k) * HS] = rc.pos[k];
#pragma unroll
  for (int k = 0; k < 6; k++) c[(K_N + k) * HS] = fr[k];
  c[K_DIST * HS] = rc.dist;
  c[K_INVW * HS] = iw;
  c[K_BODY1 * HS] = __int_as_float(b1);
  c[K_BODY2 * HS] = __int_as_float(b2);
  c[K_GEOM * HS] = __int_as_float(g);
  c[K_OTHER * HS] = __int_as_float(other);
}
static __device__ __noinline__ void write_contact_record2v(float* c, float dist, float px, float py, float pz, float nx, float ny, float nz,
                                                           float hx, float hy, float hz, int b1, int b2, float iw, int g, int other) {
  float fr[9] = {nx, ny, nz, hx, hy, hz, 0.f, 0.f, 0.f};
  make_frame(fr);
  c[(K_POS + 0) * HS] = px; c[(K_POS + 1) * HS] = py; c[(K_POS + 2) * HS] = pz;
#pragma unroll
  for (int k = 0; k < 6; k++) c[(K_N + k) * HS] = fr[k];
  c[K_DIST * HS] = dist;
  c[K_INVW * HS] = iw;
  c[K_BODY1 * HS] = __int_as_float(b1);
  c[K_BODY2 * HS] = __int_as_float(b2);
  c[K_GEOM * HS] = __int_as_float(g);
  c[K_OTHER * HS] = __int_as_float(other);
}

}  // namespace mmz
