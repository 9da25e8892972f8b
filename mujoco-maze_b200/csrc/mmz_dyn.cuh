// mmz_dyn.cuh - forward dynamics of one environment on a group of G lanes (device, fp32).
//
// This is the from-scratch replacement for what the reference reaches through
// MujocoEnv.do_simulation / set_state -> mj_step / mj_forward (third-party MuJoCo 2.0; call
// sites reference ant.py:63,95,108, point.py:57-59,80,89, swimmer.py:39,67,73):
//   kinematics -> composite-rigid-body mass matrix -> collision -> bias (RNE) + passive forces
//   -> actuation -> soft-constraint rows (joint limits, pyramidal frictional contacts)
//   -> Newton solve with exact line search -> qacc,
// restricted to the features the reference's assets use.
//
// Lane mapping inside a group: lanes <-> bodies for the tree passes, lanes <-> geoms for the
// collision candidates, lanes <-> degrees of freedom for M / H rows (held in registers during
// the factorisation), lanes <-> contacts / constraint rows in the solver. Arrays that other
// lanes read at data-dependent indices live in the group's shared-memory workspace (mmz_layout.h).
//
// Code-size discipline: the step is instruction-fetch sensitive (a few resident warps per SM,
// each at its own place in a long program), so every large routine has exactly ONE call site
// (forward, the Cholesky, sphere_box, ...) and loops with big bodies are kept rolled.
#pragma once
#include <cstdio>

#include "mmz_layout.h"
#include "mmz_narrow.cuh"

namespace mmz {

// Control flow is kept WARP-UNIFORM: the 32/G environments that share a warp run every loop to the
// warp-wide maximum trip count (the shorter ones predicated off), so all 32 lanes reach every shuffle,
// ballot and __syncwarp together and the full mask can be used. (With per-group masks the compiler
// must guard every shuffle with a divergence check and a ~10-instruction collective fallback, which
// doubled the code size of the step.)
constexpr unsigned kFull = 0xffffffffu;
#ifdef MMZ_DEBUG_UNIFORM
#define MMZ_CONV(id) do { unsigned am_ = __activemask(); if (am_ != kFull && (threadIdx.x & 15) == 0) printf("diverged at %d: active %08x block %d thread %d\n", id, am_, blockIdx.x, threadIdx.x); } while (0)
#else
#define MMZ_CONV(id) do {} while (0)
#endif

template <int G>
MMZ_DI float gsum(float v) {  // sum over the group
#pragma unroll
  for (int off = G / 2; off > 0; off >>= 1) v += __shfl_xor_sync(kFull, v, off);
  return v;
}
MMZ_DI int wmax(int v) {  // maximum over the whole warp (all groups)
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v = max(v, __shfl_xor_sync(kFull, v, off));
  return v;
}
// inclusive prefix sum over the group; *total = sum over the group's lanes
template <int G>
MMZ_DI int gscan(int v, int lane, int* total) {
#pragma unroll
  for (int off = 1; off < G; off <<= 1) {
    int t = __shfl_up_sync(kFull, v, off, G);
    if (lane >= off) v += t;
  }
  *total = __shfl_sync(kFull, v, G - 1, G);
  return v;
}

constexpr int kMaxNewton = 24;
// optional features a kernel instance is compiled with (the rest of the code is absent from it:
// the step is instruction-fetch sensitive, so a model only pays for what it uses)
enum { FEAT_BOX = 1,    // box geoms on moving bodies (Point's arrow, movable blocks): box-box, plane-box
       FEAT_FLUID = 2,  // fluid drag (Swimmer: density / viscosity)
       FEAT_PAIR = 4,   // sphere-sphere / sphere-capsule contacts between moving bodies (object balls)
       FEAT_ALL = 7 };
constexpr int kMaxLineSearch = 24;

template <int G, int NVP, int FEAT>
struct Env {
  const mmz_model* m;  // shared memory
  const Derived* dv;   // shared memory
  float* w;            // this environment's workspace (shared memory)
  int lane;            // 0..G-1
  int gshift;          // bit position of this group's lane 0 inside the warp
  unsigned bsync;      // block-barrier placement (bit 0: start of every forward, bit 1: before the solve, bit 2: per mj_step)

  MMZ_DI void sync() const { __syncwarp(); }
  // ballot restricted to this group's lanes (bit i = lane i of the group)
  MMZ_DI unsigned gballot(bool p) const {
    const unsigned b = __ballot_sync(kFull, p);
    return G == 32 ? b : ((b >> gshift) & ((1u << (G % 32)) - 1u));
  }
  MMZ_DI int* cnt(const Layout& L) const { return reinterpret_cast<int*>(w + L.o_cnt); }
  // All spatial quantities (cdof, inertias, wrenches, contact Jacobians) are taken about the
  // origin of body 0 instead of the world origin: the physics is translation invariant, and in
  // fp32 this avoids cancelling m*|c|^2 terms against each other far from the maze origin.
  MMZ_DI void rel(const Layout& L, float* r, const float* p) const {
    r[0] = p[0] - w[L.o_xpos]; r[1] = p[1] - w[L.o_xpos + 1]; r[2] = p[2] - w[L.o_xpos + 2];
  }

  // ---------------------------------------------------------------- kinematics (mj_kinematics)
  MMZ_DI void body_kin(const Layout& L, int b) {
    float* qpos = w + L.o_qpos;
    float* xposA = w + L.o_xpos;
    float* xquatA = w + L.o_xquat;
    float* xmatA = w + L.o_xmat;
    int p = m->body_parent[b];
    float pos[3], quat[4], R[9];
    if (p < 0) {
#pragma unroll
      for (int k = 0; k < 3; k++) pos[k] = m->body_pos[b][k];
#pragma unroll
      for (int k = 0; k < 4; k++) quat[k] = m->body_quat[b][k];
    } else {
      mat_vec(pos, xmatA + 9 * p, m->body_pos[b]);
#pragma unroll
      for (int k = 0; k < 3; k++) pos[k] += xposA[3 * p + k];
      quat_mul(quat, xquatA + 4 * p, m->body_quat[b]);
    }
    int j0 = m->body_jntadr[b], j1 = j0 + m->body_jntnum[b];
#pragma unroll 1
    for (int j = j0; j < j1; j++) {
      int qa = m->jnt_qadr[j];
      float* anchor = w + L.o_xanchor + 3 * j;
      float* axis = w + L.o_xaxis + 3 * j;
      int type = m->jnt_type[j];
      if (type == MMZ_JNT_FREE) {
        quat_norm(qpos + qa + 3);  // MuJoCo normalises the stored quaternion in place
#pragma unroll
        for (int k = 0; k < 3; k++) pos[k] = qpos[qa + k];
#pragma unroll
        for (int k = 0; k < 4; k++) quat[k] = qpos[qa + 3 + k];
#pragma unroll
        for (int k = 0; k < 3; k++) { anchor[k] = pos[k]; axis[k] = (k == 2) ? 1.f : 0.f; }
        continue;
      }
      quat2mat(R, quat);
      float an[3], ax[3];
      mat_vec(an, R, m->jnt_pos[j]);
#pragma unroll
      for (int k = 0; k < 3; k++) an[k] += pos[k];
      mat_vec(ax, R, m->jnt_axis[j]);
#pragma unroll
      for (int k = 0; k < 3; k++) { anchor[k] = an[k]; axis[k] = ax[k]; }
      float dq = qpos[qa] - m->qpos0[qa];
      if (type == MMZ_JNT_SLIDE) {
#pragma unroll
        for (int k = 0; k < 3; k++) pos[k] += ax[k] * dq;
      } else {  // hinge: rotate about the anchor
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
    }
    quat_norm(quat);
    quat2mat(R, quat);
#pragma unroll
    for (int k = 0; k < 3; k++) xposA[3 * b + k] = pos[k];
#pragma unroll
    for (int k = 0; k < 4; k++) xquatA[4 * b + k] = quat[k];
#pragma unroll
    for (int k = 0; k < 9; k++) xmatA[9 * b + k] = R[k];
    float ip[3], qi[4], Ri[9];
    mat_vec(ip, R, m->body_ipos[b]);
#pragma unroll
    for (int k = 0; k < 3; k++) w[L.o_xipos + 3 * b + k] = ip[k] + pos[k];
    quat_mul(qi, quat, m->body_iquat[b]);
    quat2mat(Ri, qi);
#pragma unroll
    for (int k = 0; k < 9; k++) w[L.o_ximat + 9 * b + k] = Ri[k];
  }

  MMZ_DI void kinematics(const Layout& L) {
    // one pass over (body chunk, level): a lane handles body `b` when the loop reaches its level
    const int nlev = dv->nlev;
#pragma unroll 1
    for (int it = 0; it < nlev * ((L.nb + G - 1) / G); it++) {
      int lvl = it % nlev, b = (it / nlev) * G + lane;
      if (b < L.nb && m->body_level[b] == lvl) body_kin(L, b);
      sync();
    }
#pragma unroll 1
    for (int g = lane; g < L.ng; g += G) {
      int b = m->geom_body[g];
      float p[3], q[4], R[9];
      mat_vec(p, w + L.o_xmat + 9 * b, m->geom_pos[g]);
#pragma unroll
      for (int k = 0; k < 3; k++) w[L.o_gpos + 3 * g + k] = p[k] + w[L.o_xpos + 3 * b + k];
      quat_mul(q, w + L.o_xquat + 4 * b, m->geom_quat[g]);
      quat2mat(R, q);
#pragma unroll
      for (int k = 0; k < 9; k++) w[L.o_gmat + 9 * g + k] = R[k];
    }
    sync();
  }

  // ---------------------------------------------------------------- joint motion axes (cdof)
  MMZ_DI void motion_axes(const Layout& L) {
    // lanes <-> dofs: dof d of joint j (free joint: 3 translations then 3 body-axis rotations)
#pragma unroll 1
    for (int d = lane; d < L.nv; d += G) {
      int j = m->dof_jnt[d], b = m->jnt_body[j], type = m->jnt_type[j], k = d - m->jnt_dadr[j];
      float c[6];
      if (type == MMZ_JNT_FREE && k < 3) {
#pragma unroll
        for (int i = 0; i < 6; i++) c[i] = (i == 3 + k) ? 1.f : 0.f;
      } else if (type == MMZ_JNT_SLIDE) {
        const float* axis = w + L.o_xaxis + 3 * j;
        c[0] = c[1] = c[2] = 0.f;
        c[3] = axis[0]; c[4] = axis[1]; c[5] = axis[2];
      } else {
        float ax[3], at[3];
        if (type == MMZ_JNT_FREE) {  // rotation about body axis k-3 through the body origin
          const float* xm = w + L.o_xmat + 9 * b + (k - 3);
          ax[0] = xm[0]; ax[1] = xm[3]; ax[2] = xm[6];
          rel(L, at, w + L.o_xpos + 3 * b);
        } else {
          const float* axis = w + L.o_xaxis + 3 * j;
          ax[0] = axis[0]; ax[1] = axis[1]; ax[2] = axis[2];
          rel(L, at, w + L.o_xanchor + 3 * j);
        }
        cross3(c + 3, at, ax);
        c[0] = ax[0]; c[1] = ax[1]; c[2] = ax[2];
      }
      float* out = w + L.o_cdof + 6 * d;
#pragma unroll
      for (int i = 0; i < 6; i++) out[i] = c[i];
    }
    sync();
  }

  MMZ_DI void body_inertia_world(const Layout& L, int b, float* I) const {
    const float* R = w + L.o_ximat + 9 * b;
    const float* d = m->body_inertia[b];
    float c[3];
    rel(L, c, w + L.o_xipos + 3 * b);
    float mass = m->body_mass[b];
    float cc = dot3(c, c);
    I[0] = R[0] * R[0] * d[0] + R[1] * R[1] * d[1] + R[2] * R[2] * d[2] + mass * (cc - c[0] * c[0]);
    I[1] = R[3] * R[3] * d[0] + R[4] * R[4] * d[1] + R[5] * R[5] * d[2] + mass * (cc - c[1] * c[1]);
    I[2] = R[6] * R[6] * d[0] + R[7] * R[7] * d[1] + R[8] * R[8] * d[2] + mass * (cc - c[2] * c[2]);
    I[3] = R[0] * R[3] * d[0] + R[1] * R[4] * d[1] + R[2] * R[5] * d[2] - mass * c[0] * c[1];
    I[4] = R[0] * R[6] * d[0] + R[1] * R[7] * d[1] + R[2] * R[8] * d[2] - mass * c[0] * c[2];
    I[5] = R[3] * R[6] * d[0] + R[4] * R[7] * d[1] + R[5] * R[8] * d[2] - mass * c[1] * c[2];
    I[6] = mass * c[0]; I[7] = mass * c[1]; I[8] = mass * c[2];
    I[9] = mass;
  }

  // ---------------------------------------------------------------- composite rigid body -> M (mj_crb)
  MMZ_DI void mass_matrix(const Layout& L) {
    float* Iw = w + L.o_iw;
    float* Ic = w + L.o_ic;
    float* M = w + L.o_M;
#pragma unroll 1
    for (int b = lane; b < L.nb; b += G) {
      float I[10];
      body_inertia_world(L, b, I);
#pragma unroll
      for (int k = 0; k < 10; k++) Iw[10 * b + k] = I[k];
    }
#pragma unroll 1
    for (int i = lane; i < L.nv * L.ldm; i += G) M[i] = 0.f;
    sync();
#pragma unroll 1
    for (int b = lane; b < L.nb; b += G) {  // composite inertia of the subtree rooted at b
      float I[10];
#pragma unroll
      for (int k = 0; k < 10; k++) I[k] = Iw[10 * b + k];
#pragma unroll 1
      for (int c = b + 1; c < L.nb; c++)
        if (dv->anc[c] >> b & 1) {
#pragma unroll
          for (int k = 0; k < 10; k++) I[k] += Iw[10 * c + k];
        }
#pragma unroll
      for (int k = 0; k < 10; k++) Ic[10 * b + k] = I[k];
    }
    sync();
    const float* cdof = w + L.o_cdof;
#pragma unroll 1
    for (int i = lane; i < L.nv; i += G) {
      float f[6];
      inert_mul(f, Ic + 10 * m->dof_body[i], cdof + 6 * i);
#pragma unroll 1
      for (int j = i; j >= 0; j = m->dof_parent[j]) {
        float v = dot6(cdof + 6 * j, f);
        if (j == i) v += m->dof_armature[i];
        M[i * L.ldm + j] = v;
        M[j * L.ldm + i] = v;
      }
    }
    sync();
  }

  // ---------------------------------------------------------------- bias forces c(q, qvel) (mj_rne)
  MMZ_DI void bias_forces(const Layout& L) {
    const float* cdof = w + L.o_cdof;
    const float* qvel = w + L.o_qvel;
    float* vel = w + L.o_vel;
    float* acc = w + L.o_acc;
    float* frc = w + L.o_frc;
    const int nlev = dv->nlev;
#pragma unroll 1
    for (int it = 0; it < nlev * ((L.nb + G - 1) / G); it++) {
      int lvl = it % nlev, b = (it / nlev) * G + lane;
      if (b < L.nb && m->body_level[b] == lvl) {
        int p = m->body_parent[b];
        float v[6], a[6];
        if (p < 0) {
#pragma unroll
          for (int k = 0; k < 6; k++) { v[k] = 0.f; a[k] = 0.f; }
#pragma unroll
          for (int k = 0; k < 3; k++) a[3 + k] = -m->gravity[k];  // gravity as base acceleration
        } else {
#pragma unroll
          for (int k = 0; k < 6; k++) { v[k] = vel[6 * p + k]; a[k] = acc[6 * p + k]; }
        }
        // dofs of this body in order. A free joint's three rotation axes all take their axis
        // derivative against the velocity after its translations and before its rotations.
        int d0 = m->body_dofadr[b], d1 = d0 + m->body_dofnum[b];
        float vf[6];
#pragma unroll
        for (int k = 0; k < 6; k++) vf[k] = v[k];
#pragma unroll 1
        for (int d = d0; d < d1; d++) {
          int j = m->dof_jnt[d], kk = d - m->jnt_dadr[j];
          const float* s = cdof + 6 * d;
          float qv = qvel[d];
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
          float sd[6], vs[6];
#pragma unroll
          for (int k = 0; k < 6; k++) vs[k] = isfree ? vf[k] : v[k];
          cross_motion(sd, vs, s);
#pragma unroll
          for (int i = 0; i < 6; i++) { a[i] += sd[i] * qv; v[i] += s[i] * qv; }
        }
#pragma unroll
        for (int k = 0; k < 6; k++) { vel[6 * b + k] = v[k]; acc[6 * b + k] = a[k]; }
        float Ia[6], Iv[6], vxIv[6];
        const float* I = w + L.o_iw + 10 * b;
        inert_mul(Ia, I, a);
        inert_mul(Iv, I, v);
        cross_force(vxIv, v, Iv);
#pragma unroll
        for (int k = 0; k < 6; k++) frc[6 * b + k] = Ia[k] + vxIv[k];
      }
      sync();
    }
    float* fsub = w + L.o_fsub;
#pragma unroll 1
    for (int b = lane; b < L.nb; b += G) {
      float f[6];
#pragma unroll
      for (int k = 0; k < 6; k++) f[k] = frc[6 * b + k];
#pragma unroll 1
      for (int c = b + 1; c < L.nb; c++)
        if (dv->anc[c] >> b & 1) {
#pragma unroll
          for (int k = 0; k < 6; k++) f[k] += frc[6 * c + k];
        }
#pragma unroll
      for (int k = 0; k < 6; k++) fsub[6 * b + k] = f[k];
    }
    sync();
  }

  // ---------------------------------------------------------------- fluid forces (mj_passive, Swimmer only)
  MMZ_DI void fluid_forces(const Layout& L) {
    float* sf = w + L.o_frc;  // reuse: per-body fluid wrench about the origin
#pragma unroll 1
    for (int b = lane; b < L.nb; b += G) {
      float mass = m->body_mass[b];
      float out[6] = {0, 0, 0, 0, 0, 0};
      if (mass >= kMinVal) {
        const float* I = m->body_inertia[b];
        float box[3] = {sqrtf(fmaxf(kMinVal, I[1] + I[2] - I[0]) / mass * 6.f),
                        sqrtf(fmaxf(kMinVal, I[0] + I[2] - I[1]) / mass * 6.f),
                        sqrtf(fmaxf(kMinVal, I[0] + I[1] - I[2]) / mass * 6.f)};
        const float* v = w + L.o_vel + 6 * b;
        float xi[3];
        rel(L, xi, w + L.o_xipos + 3 * b);
        const float* Ri = w + L.o_ximat + 9 * b;
        float wxc[3], vc[3], lw[3], lv[3], lf[6] = {0, 0, 0, 0, 0, 0};
        cross3(wxc, v, xi);
#pragma unroll
        for (int k = 0; k < 3; k++) vc[k] = v[3 + k] + wxc[k];
        matT_vec(lw, Ri, v);
        matT_vec(lv, Ri, vc);
        if (m->viscosity > 0.f) {
          float diam = (box[0] + box[1] + box[2]) * (1.f / 3.f);
#pragma unroll
          for (int k = 0; k < 3; k++) {
            lf[k] = -kPi * diam * diam * diam * m->viscosity * lw[k];
            lf[3 + k] = -3.f * kPi * diam * m->viscosity * lv[k];
          }
        }
        if (m->density > 0.f) {
          float b0 = box[0], b1 = box[1], b2 = box[2];
          float p0 = b0 * b0 * b0 * b0, p1 = b1 * b1 * b1 * b1, p2 = b2 * b2 * b2 * b2;
          lf[3] -= 0.5f * m->density * b1 * b2 * fabsf(lv[0]) * lv[0];
          lf[4] -= 0.5f * m->density * b0 * b2 * fabsf(lv[1]) * lv[1];
          lf[5] -= 0.5f * m->density * b0 * b1 * fabsf(lv[2]) * lv[2];
          lf[0] -= m->density * b0 * (p1 + p2) * fabsf(lw[0]) * lw[0] * (1.f / 64.f);
          lf[1] -= m->density * b1 * (p0 + p2) * fabsf(lw[1]) * lw[1] * (1.f / 64.f);
          lf[2] -= m->density * b2 * (p0 + p1) * fabsf(lw[2]) * lw[2] * (1.f / 64.f);
        }
        float tq[3], fc[3], cxf[3];
        mat_vec(tq, Ri, lf);
        mat_vec(fc, Ri, lf + 3);
        cross3(cxf, xi, fc);
#pragma unroll
        for (int k = 0; k < 3; k++) { out[k] = tq[k] + cxf[k]; out[3 + k] = fc[k]; }
      }
#pragma unroll
      for (int k = 0; k < 6; k++) sf[6 * b + k] = out[k];
    }
    sync();
  }

  // qfrc_smooth = passive (damping + fluid) - bias + actuation, dofs over lanes
  MMZ_DI void smooth_forces(const Layout& L) {
    const float* cdof = w + L.o_cdof;
    const float* qvel = w + L.o_qvel;
    const float* fsub = w + L.o_fsub;
    const bool fluid = (FEAT & FEAT_FLUID) && (m->density > 0.f || m->viscosity > 0.f);
    if constexpr ((FEAT & FEAT_FLUID) != 0) { if (fluid) fluid_forces(L); }  // wrenches land in frc, which the subtree sums have consumed
    const float* sf = w + L.o_frc;
#pragma unroll 1
    for (int d = lane; d < L.nv; d += G) {
      float b = dot6(cdof + 6 * d, fsub + 6 * m->dof_body[d]);
      float p = -m->dof_damping[d] * qvel[d];
      if (fluid) {
#pragma unroll 1
        for (int bb = 0; bb < L.nb; bb++)
          if (m->body_dofmask[bb] >> d & 1) p += dot6(cdof + 6 * d, sf + 6 * bb);
      }
      float act = 0.f;
#pragma unroll 1
      for (int k = 0; k < L.nu; k++)
        if (m->act_dof[k] == d) {
          float c = w[L.o_ctrl + k];
          if (m->act_limited[k]) c = fminf(fmaxf(c, m->act_ctrlrange[k][0]), m->act_ctrlrange[k][1]);
          act += m->act_gear[k] * c;
        }
      w[L.o_smooth + d] = p - b + act;
    }
    sync();
  }

  // ---------------------------------------------------------------- collision (mj_collision)
  // mixed contact parameters of geom g against `other` (-1 floor, -2 wall/platform box, >= 0 geom)
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
  MMZ_DI void write_contact(const Layout& L, int slot, const RawContact& rc, int b1, int b2, float invw, int g,
                            int other) {
    float* c = w + L.o_con + slot * L.cstride;
    float fr[9], par[9];
    mix_params(g, other, par);
#pragma unroll
    for (int k = 0; k < 3; k++) { fr[k] = rc.normal[k]; fr[3 + k] = rc.hint[k]; }
    make_frame(fr);
    c[C_DIST] = rc.dist;
#pragma unroll
    for (int k = 0; k < 3; k++) c[C_POS + k] = rc.pos[k];
#pragma unroll
    for (int k = 0; k < 9; k++) c[C_FRAME + k] = fr[k];
    c[C_BODY1] = __int_as_float(b1);
    c[C_BODY2] = __int_as_float(b2);
    c[C_MU] = par[1];
    c[C_MARGIN] = par[0];
#pragma unroll
    for (int k = 0; k < 7; k++) c[C_SOLREF + k] = par[2 + k];
    c[C_INVW] = invw;
  }
  // grid cells whose wall / platform box can touch an axis-aligned extent
  MMZ_DI void cell_range(const float* c, const float* ext, int* i0, int* i1, int* j0, int* j1) const {
    float s = m->cell_size, hs = m->wall_half[0];
    *j0 = max(0, (int)ceilf((c[0] - ext[0] + m->origin[0] - hs) / s));
    *j1 = min(m->grid_w - 1, (int)floorf((c[0] + ext[0] + m->origin[0] + hs) / s));
    *i0 = max(0, (int)ceilf((c[1] - ext[1] + m->origin[1] - hs) / s));
    *i1 = min(m->grid_h - 1, (int)floorf((c[1] + ext[1] + m->origin[1] + hs) / s));
  }
  MMZ_DI bool moving_pair_ok(int g1, int g2) const {
    int b1 = m->geom_body[g1], b2 = m->geom_body[g2];
    if (b1 == b2 || m->body_parent[b1] == b2 || m->body_parent[b2] == b1) return false;
    return (m->geom_contype[g1] & m->geom_conaffinity[g2]) || (m->geom_contype[g2] & m->geom_conaffinity[g1]);
  }

  MMZ_DI void collision(const Layout& L) {
    int* cn = cnt(L);
    int ncon = 0;
    bool overflow = false;
    if (!m->collision_on) {
      if (lane == 0) { cn[N_CON] = 0; cn[N_OVERFLOW] = 0; }
      sync();
      return;
    }
    const int nslot = m->elevated ? 2 : 1;  // wall box, platform box per cell
    // ---- spheres and capsules: one lane per geom, candidates walked in lock step
#pragma unroll 1
    for (int gbase = 0; gbase < L.ng; gbase += G) {
      int g = gbase + lane;
      int type = g < L.ng ? m->geom_type[g] : -1;
      const bool capsule = type == MMZ_GEOM_CAPSULE;
      bool valid = (type == MMZ_GEOM_SPHERE || capsule) && ((m->geom_contype[g] | m->geom_conaffinity[g]) & 1);
      float r = 0.f, p0[3] = {0, 0, 0}, p1[3] = {0, 0, 0}, gmarg = 0.f;
      int i0 = 0, j0 = 0, nj = 1, ncell = 0, body = -1;
      float invw = 0.f;
      if (valid) {
        r = m->geom_size[g][0];
        body = m->geom_body[g];
        invw = m->geom_invweight[g];
        gmarg = m->geom_margin[g];
        const float* gp = w + L.o_gpos + 3 * g;
        const float* gm = w + L.o_gmat + 9 * g;
        float hl = capsule ? m->geom_size[g][1] : 0.f, ext[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
          float a = gm[3 * k + 2] * hl;
          p0[k] = gp[k] + a; p1[k] = gp[k] - a;
          ext[k] = fabsf(a);
        }
        float mg = r + fmaxf(gmarg, m->wall_margin);
        ext[0] += mg; ext[1] += mg;
        int i1, j1;
        cell_range(gp, ext, &i0, &i1, &j0, &j1);
        nj = max(1, j1 - j0 + 1);
        ncell = max(0, j1 - j0 + 1) * max(0, i1 - i0 + 1);
      }
      const int maxcell = wmax(ncell);  // warp-uniform candidate count
      const int ncand = 1 + nslot * maxcell + ((FEAT & FEAT_BOX) ? dv->nboxg : 0);
#pragma unroll 1
      for (int cand = 0; cand < ncand; cand++) {
        RawContact r0, r1;
        int n = 0, b1 = -1, b2 = -1, other = -1;
        float iw = invw;
        if (valid) {
          bool box = false;
          float bc[3] = {0.f, 0.f, 0.f}, margin = 0.f;
          const float* bR = dv->ident;
          const float* bh = m->wall_half;
          if (cand == 0) {  // floor plane (normal +z); geom1 = plane
            if (m->has_floor) {
              margin = fmaxf(gmarg, m->floor_margin);
              b2 = body;
              float d0 = p0[2] - m->floor_z - r, d1 = p1[2] - m->floor_z - r;
              float hint[3] = {0.f, 0.f, 0.f};
              if (capsule) {  // first tangent along the capsule axis
                const float* gm = w + L.o_gmat + 9 * g;
                if (fabsf(gm[8]) <= 0.999999f) { hint[0] = gm[2]; hint[1] = gm[5]; hint[2] = gm[8]; }
              }
              const bool c0 = d0 < margin, c1 = capsule && d1 < margin;
              // r0 takes the first active end, r1 the second
              float da = c0 ? d0 : d1;
#pragma unroll
              for (int k = 0; k < 3; k++) {
                r0.pos[k] = c0 ? p0[k] : p1[k];
                r1.pos[k] = p1[k];
                r0.normal[k] = r1.normal[k] = (k == 2) ? 1.f : 0.f;
                r0.hint[k] = r1.hint[k] = hint[k];
              }
              r0.dist = da; r0.pos[2] -= r + 0.5f * da;
              r1.dist = d1; r1.pos[2] -= r + 0.5f * d1;
              n = (int)c0 + (int)c1;
            }
          } else if (cand - 1 < nslot * maxcell) {  // maze boxes; geom1 = robot geom, geom2 = box
            int ci = (cand - 1) / nslot, slot = (cand - 1) % nslot;
            if (ci < ncell) {
              int i = i0 + ci / nj, j = j0 + ci % nj;
              int code = m->grid[i * m->grid_w + j];
              if (code & (slot == 0 ? MMZ_CELL_WALL : MMZ_CELL_PLATFORM)) {
                bc[0] = j * m->cell_size - m->origin[0];
                bc[1] = i * m->cell_size - m->origin[1];
                bc[2] = (slot == 0) ? m->wall_z : m->plat_z;
                margin = fmaxf(gmarg, m->wall_margin);
                other = -2; b1 = body;
                box = true;
              }
            }
          } else if (FEAT & FEAT_BOX) {  // box geoms on other moving bodies; geom1 = this geom, geom2 = box
            int gb = dv->boxg[cand - 1 - nslot * maxcell];
            if (moving_pair_ok(g, gb)) {
              margin = fmaxf(gmarg, m->geom_margin[gb]);
              other = gb; b1 = body; b2 = m->geom_body[gb];
              iw = invw + m->geom_invweight[gb];
              const float* gc = w + L.o_gpos + 3 * gb;
              bc[0] = gc[0]; bc[1] = gc[1]; bc[2] = gc[2];
              bR = w + L.o_gmat + 9 * gb;
              bh = m->geom_size[gb];
              box = true;
            }
          }
          if (box) {
            // capsule: both end caps when both are within the margin, otherwise the segment point
            // nearest the box; sphere: the centre. One sphere_box call site for all three probes.
            int n0 = 0, n1 = 0;
#pragma unroll 1
            for (int pr = 0; pr < 3; pr++) {
              float pt[3];
              if (pr == 0) { pt[0] = p0[0]; pt[1] = p0[1]; pt[2] = p0[2]; }
              else if (!capsule) break;
              else if (pr == 1) { pt[0] = p1[0]; pt[1] = p1[1]; pt[2] = p1[2]; }
              else {
                if (n0 && n1) break;
                float ts = capsule_nearest(p0, p1, bc, bR, bh);
#pragma unroll
                for (int k = 0; k < 3; k++) pt[k] = p0[k] + ts * (p1[k] - p0[k]);
              }
              RawContact t;
              int nt = sphere_box(pt, r, bc, bR, bh, margin, &t);
              if (pr == 1) { r1 = t; n1 = nt; }
              else { r0 = t; n0 = nt; if (pr == 2) n1 = 0; }
            }
            n = (n0 && n1) ? 2 : n0;  // after the third probe n1 == 0 and n0 is the nearest-point result
          }
        }
        if (!__any_sync(kFull, n > 0)) continue;
        int total, incl = gscan<G>(n, lane, &total);  // keeps the contact order deterministic
        int base = ncon + incl - n;
        if (n > 0 && base < L.maxcon) write_contact(L, base, r0, b1, b2, iw, g, other);
        if (n > 1 && base + 1 < L.maxcon) write_contact(L, base + 1, r1, b1, b2, iw, g, other);
        ncon += total;
        if (ncon > L.maxcon) { ncon = L.maxcon; overflow = true; }
      }
    }
    // ---- spheres against spheres / capsules on other moving bodies (object balls): one lane per listed pair
    if constexpr ((FEAT & FEAT_PAIR) != 0) {
#pragma unroll 1
      for (int pbase = 0; pbase < dv->npair; pbase += G) {
        const int p = pbase + lane;
        RawContact rc;
        int n = 0, ga = 0, gb = 0;
        if (p < dv->npair) {
          ga = dv->pair_a[p]; gb = dv->pair_b[p];
          const float* c1 = w + L.o_gpos + 3 * ga;
          float q[3];
          if (m->geom_type[gb] == MMZ_GEOM_CAPSULE) {
            const float* gp = w + L.o_gpos + 3 * gb;
            const float* gm = w + L.o_gmat + 9 * gb;
            const float hl = m->geom_size[gb][1];
            const float p0[3] = {gp[0] + gm[2] * hl, gp[1] + gm[5] * hl, gp[2] + gm[8] * hl};
            const float p1[3] = {gp[0] - gm[2] * hl, gp[1] - gm[5] * hl, gp[2] - gm[8] * hl};
            segment_nearest(p0, p1, c1, q);
          } else {
#pragma unroll
            for (int k = 0; k < 3; k++) q[k] = w[L.o_gpos + 3 * gb + k];
          }
          n = sphere_sphere(c1, m->geom_size[ga][0], q, m->geom_size[gb][0], fmaxf(m->geom_margin[ga], m->geom_margin[gb]), &rc);
        }
        if (!__any_sync(kFull, n > 0)) continue;
        int total, incl = gscan<G>(n, lane, &total);
        const int base = ncon + incl - n;
        if (n > 0 && base < L.maxcon)
          write_contact(L, base, rc, m->geom_body[ga], m->geom_body[gb], m->geom_invweight[ga] + m->geom_invweight[gb], ga, gb);
        ncon += total;
        if (ncon > L.maxcon) { ncon = L.maxcon; overflow = true; }
      }
    }
    if (lane == 0) { cn[N_CON] = ncon; cn[N_OVERFLOW] = overflow ? 1 : 0; }
    sync();
    // ---- box geoms (Point's arrow, movable blocks): few, handled by one lane each in order
    if constexpr ((FEAT & FEAT_BOX) != 0) {
#pragma unroll 1
      for (int k = 0; k < dv->nboxg; k++) {
        box_geom_contacts(L, dv->boxg[k], k);
      }
    }
  }

  // Box geom g (Point's arrow, a movable block): its candidates - the floor, the maze boxes near it and the
  // box geoms after it - are split over the lanes of the group (box-box is long, serial code); a scan over
  // the lanes keeps the contact order (candidate-major) independent of the mapping. Warp-uniform call.
  // (Same run, 20 steps: PointUMaze 4096 envs 0.42 -> 0.35 ms, AntPush 32768 envs 29.4 -> 26.1 ms per step.)
  MMZ_DI void box_geom_contacts(const Layout& L, int g, int kbox) {
    int* cn = cnt(L);
    int ncon = cn[N_CON];
    bool overflow = false;
    const bool active = ((m->geom_contype[g] | m->geom_conaffinity[g]) & 1) != 0;
    const float* gp = w + L.o_gpos + 3 * g;
    const float* gm = w + L.o_gmat + 9 * g;
    const float* sz = m->geom_size[g];
    const int body = m->geom_body[g];
    const float invw = m->geom_invweight[g];
    float ext[3], wallmargin = fmaxf(m->geom_margin[g], m->wall_margin);
#pragma unroll
    for (int k = 0; k < 3; k++) ext[k] = fabsf(gm[3 * k]) * sz[0] + fabsf(gm[3 * k + 1]) * sz[1] + fabsf(gm[3 * k + 2]) * sz[2];
    ext[0] += wallmargin; ext[1] += wallmargin;
    int i0, i1, j0, j1;
    cell_range(gp, ext, &i0, &i1, &j0, &j1);
    const int nj = max(0, j1 - j0 + 1), ncell = nj * max(0, i1 - i0 + 1);
    // candidates: floor, then (cell, wall|platform) pairs, then later box geoms on other bodies
    const int ncand = active ? 1 + 2 * ncell + (dv->nboxg - kbox - 1) : 0;
    const int ncandw = wmax(ncand);
    constexpr int P = G;
#pragma unroll 1
    for (int cbase = 0; cbase < ncandw; cbase += P) {
      const int cand = cbase + lane;
      RawContact rc[8];
      int n = 0, b1 = -1, b2 = body, other = -1;
      float iw = invw;
      if (lane < P && cand < ncand) {
        if (cand == 0) {  // corners below the plane, at most 4; geom1 = plane
          if (m->has_floor) {
            float margin = fmaxf(m->geom_margin[g], m->floor_margin);
            for (int c = 0; c < 8 && n < 4; c++) {
              float loc[3] = {(c & 1 ? 1.f : -1.f) * sz[0], (c & 2 ? 1.f : -1.f) * sz[1], (c & 4 ? 1.f : -1.f) * sz[2]}, wp[3];
              mat_vec(wp, gm, loc);
              float dist = wp[2] + gp[2] - m->floor_z;
              if (dist < margin) {
                rc[n].dist = dist;
                rc[n].pos[0] = wp[0] + gp[0]; rc[n].pos[1] = wp[1] + gp[1]; rc[n].pos[2] = wp[2] + gp[2] - 0.5f * dist;
                rc[n].normal[0] = 0.f; rc[n].normal[1] = 0.f; rc[n].normal[2] = 1.f;
                rc[n].hint[0] = rc[n].hint[1] = rc[n].hint[2] = 0.f;
                n++;
              }
            }
          }
        } else if (cand - 1 < 2 * ncell) {  // maze boxes; geom1 = wall (lower geom id), geom2 = this box
          int ci = (cand - 1) >> 1, slot = (cand - 1) & 1;
          int i = i0 + ci / nj, j = j0 + ci % nj;
          int code = m->grid[i * m->grid_w + j];
          if (code & (slot == 0 ? MMZ_CELL_WALL : MMZ_CELL_PLATFORM)) {
            float bc[3] = {j * m->cell_size - m->origin[0], i * m->cell_size - m->origin[1], slot == 0 ? m->wall_z : m->plat_z};
            other = -2;
            n = box_box(bc, dv->ident, m->wall_half, gp, gm, sz, wallmargin, rc);
          }
        } else {  // box against box on different moving bodies
          int g2 = dv->boxg[kbox + 1 + (cand - 1 - 2 * ncell)];
          if (moving_pair_ok(g, g2)) {
            other = g2; b1 = body; b2 = m->geom_body[g2];
            iw = invw + m->geom_invweight[g2];
            n = box_box(gp, gm, sz, w + L.o_gpos + 3 * g2, w + L.o_gmat + 9 * g2, m->geom_size[g2],
                        fmaxf(m->geom_margin[g], m->geom_margin[g2]), rc);
          }
        }
      }
      if (!__any_sync(kFull, n > 0)) continue;
      int total, incl = gscan<G>(n, lane, &total);
      const int base = ncon + incl - n;
      for (int k = 0; k < n; k++) {
        if (base + k < L.maxcon) write_contact(L, base + k, rc[k], b1, b2, iw, g, other);
      }
      ncon += total;
      if (ncon > L.maxcon) { ncon = L.maxcon; overflow = true; }
    }
    sync();
    if (lane == 0) { cn[N_CON] = ncon; if (overflow) cn[N_OVERFLOW] = 1; }
    sync();
  }

  // ---------------------------------------------------------------- constraint rows (mj_makeConstraint)
  MMZ_DI static float impedance(const float* si, float r) {
    float d0 = fminf(fmaxf(si[0], 1e-4f), 0.9999f), d1 = fminf(fmaxf(si[1], 1e-4f), 0.9999f);
    float width = si[2], mid = si[3], power = si[4];
    if (d0 == d1 || width <= kMinVal) return 0.5f * (d0 + d1);
    float x = fabsf(r) / width, y;
    if (x >= 1.f) return d1;
    if (x <= 0.f) return d0;
    if (power == 1.f) y = x;
    else if (power == 2.f) y = x <= mid ? x * x / mid : 1.f - (1.f - x) * (1.f - x) / (1.f - mid);  // MuJoCo's default power
    else if (x <= mid) y = powf(x, power) / powf(mid, power - 1.f);
    else y = 1.f - powf(1.f - x, power) / powf(1.f - mid, power - 1.f);
    return d0 + y * (d1 - d0);
  }
  // (solref, solimp, violation, 1/weight) -> D and the two pieces of aref = -bb * vel - kr
  MMZ_DI void row_params(const float* solref, const float* solimp, float pos, float margin, float diag, float* D,
                         float* kr, float* bb) const {
    float tc = fmaxf(solref[0], 2.f * m->timestep), dr = solref[1];  // refsafe
    float dmax = fminf(fmaxf(solimp[1], 1e-4f), 0.9999f);
    float k = 1.f / fmaxf(kMinVal, dmax * dmax * tc * tc * dr * dr);
    *bb = 2.f / fmaxf(kMinVal, dmax * tc);
    float imp = impedance(solimp, pos - margin);
    float R = fmaxf(kMinVal, (1.f - imp) * diag / imp);
    *D = 1.f / R;
    *kr = k * imp * (pos - margin);
  }

  // Joint limits: dof d's own limit rows live in lane d's registers (a limit row is +-e_d, so
  // nothing about it needs to be shared): side 0 = lower (J = +1), side 1 = upper (J = -1).
  // limD[s] = 0 marks an inactive side. Contacts: one lane per contact turns the narrow-phase
  // record into D, aref[4] and the signed dof masks. J qvel comes from the body velocities the
  // bias pass has just computed (point velocity of body2 minus body1), not from a Jacobian.
  float limD[2], limA[2];

  MMZ_DI void make_constraints(const Layout& L) {
    int* cn = cnt(L);
    const float* qpos = w + L.o_qpos;
    const float* qvel = w + L.o_qvel;
    const int ncon = cn[N_CON];
    limD[0] = limD[1] = 0.f; limA[0] = limA[1] = 0.f;
    // one loop over "row sources" so that row_params / impedance have a single call site:
    // pass 0, 1 = the two limit sides of this lane's dof, passes >= 2 = contacts lane, lane + G, ...
    const int jl = lane < L.nv ? m->dof_jnt[lane] : 0;
    const bool limited = lane < L.nv && m->jnt_limited[jl] && m->jnt_type[jl] >= MMZ_JNT_SLIDE;
    const int npass = 2 + (wmax(ncon) + G - 1) / G;
#pragma unroll 1
    for (int pass = 0; pass < npass; pass++) {
      float pos = 0.f, margin = 0.f, diag = 0.f, mu = 0.f, jv0 = 0.f, jv1 = 0.f, jv2 = 0.f;
      const float *solref = m->wall_solref, *solimp = m->wall_solimp;
      float* cs = nullptr;
      bool on = false;
      if (pass < 2) {
        if (limited) {
          float q = qpos[m->jnt_qadr[jl]];
          pos = pass == 0 ? q - m->jnt_range[jl][0] : m->jnt_range[jl][1] - q;
          margin = m->jnt_margin[jl];
          on = pos < margin;
          diag = m->dof_invweight0[lane];
          solref = m->jnt_solref[jl]; solimp = m->jnt_solimp[jl];
        }
      } else {
        const int c = (pass - 2) * G + lane;
        if (c < ncon) {
          on = true;
          cs = w + L.o_con + c * L.cstride;
          float cp[3], v[3] = {0.f, 0.f, 0.f};
#pragma unroll
          for (int k = 0; k < 3; k++) { cp[k] = cs[C_POS + k] - w[L.o_xpos + k]; cs[C_POS + k] = cp[k]; }
          const int b1 = __float_as_int(cs[C_BODY1]), b2 = __float_as_int(cs[C_BODY2]);
          const int mask1 = b1 >= 0 ? m->body_dofmask[b1] : 0, mask2 = b2 >= 0 ? m->body_dofmask[b2] : 0;
          if (b2 >= 0) {
            const float* bv = w + L.o_vel + 6 * b2;
            float wxp[3];
            cross3(wxp, bv, cp);
#pragma unroll
            for (int k = 0; k < 3; k++) v[k] += bv[3 + k] + wxp[k];
          }
          if (b1 >= 0) {
            const float* bv = w + L.o_vel + 6 * b1;
            float wxp[3];
            cross3(wxp, bv, cp);
#pragma unroll
            for (int k = 0; k < 3; k++) v[k] -= bv[3 + k] + wxp[k];
          }
          jv0 = dot3(cs + C_FRAME, v); jv1 = dot3(cs + C_FRAME + 3, v); jv2 = dot3(cs + C_FRAME + 6, v);
          mu = cs[C_MU];
          pos = cs[C_DIST]; margin = cs[C_MARGIN]; diag = cs[C_INVW] * (1.f + mu * mu);
          solref = cs + C_SOLREF; solimp = cs + C_SOLIMP;
          cs[C_MPOS] = __int_as_float(mask2 & ~mask1);
          cs[C_MNEG] = __int_as_float(mask1 & ~mask2);
        }
      }
      if (on) {
        float D, kr, bb;
        row_params(solref, solimp, pos, margin, diag, &D, &kr, &bb);
        if (pass < 2) {
          const float sign = pass == 0 ? 1.f : -1.f;
          limD[pass] = D;
          limA[pass] = -bb * sign * qvel[lane] - kr;
        } else {
          // all edges of the pyramid share R = 2 mu^2 R_first
          cs[C_D] = 1.f / fmaxf(kMinVal, 2.f * mu * mu / D);
          cs[C_AREF + 0] = -bb * (jv0 + mu * jv1) - kr;
          cs[C_AREF + 1] = -bb * (jv0 - mu * jv1) - kr;
          cs[C_AREF + 2] = -bb * (jv0 + mu * jv2) - kr;
          cs[C_AREF + 3] = -bb * (jv0 - mu * jv2) - kr;
        }
      }
    }
    const int nlim = __popc(gballot(limD[0] > 0.f)) + __popc(gballot(limD[1] > 0.f));
    if (lane == 0) cn[N_LIM] = nlim;
    sync();
  }

  // ---------------------------------------------------------------- dense solve in registers
  // Lane i holds row i of the symmetric positive-definite H (full row, NVP registers) and element i
  // of the right-hand side. Gaussian elimination without pivoting, the pivot row broadcast by
  // shuffles, then back substitution on the frozen upper rows: no shared memory, no barriers.
  // Rows >= nv are identity padding. Returns x = H^-1 rhs (element `lane`).
  // Gauss-Jordan, as in the hybrid kernel (mmz_hkernel.cuh: elim_solve): step j clears column j in every other row,
  // no back substitution; rows / columns beyond nv are the identity with a zero right-hand side (padded by the caller).
  MMZ_DI float elim_solve(float (&h)[NVP], float rhs) const {
    float invd = 1.f;
#pragma unroll
    for (int j = 0; j < NVP; j++) {
      const float piv = fmaxf(__shfl_sync(kFull, h[j], j, G), kMinVal);
      float inv;
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(piv));
      const float rj = __shfl_sync(kFull, rhs, j, G);
      const bool own = lane == j;
      const float f = own ? 0.f : h[j] * inv;
      invd = own ? inv : invd;
      rhs -= f * rj;
#pragma unroll
      for (int k = j + 1; k < NVP; k++) h[k] -= f * __shfl_sync(kFull, h[k], j, G);
    }
    return rhs * invd;
  }

  // this lane's column of the contact-frame Jacobian of contact block cs: (J_n, J_t1, J_t2)[lane]
  MMZ_DI void contact_jac(const float* cs, const float (&cd)[6], float* jn, float* jt1, float* jt2) const {
    const int mp = __float_as_int(cs[C_MPOS]), mn = __float_as_int(cs[C_MNEG]);
    const float s = (float)(mp >> lane & 1) - (float)(mn >> lane & 1);
    float wxp[3];
    cross3(wxp, cd, cs + C_POS);
    const float v[3] = {s * (cd[3] + wxp[0]), s * (cd[4] + wxp[1]), s * (cd[5] + wxp[2])};
    *jn = dot3(cs + C_FRAME, v); *jt1 = dot3(cs + C_FRAME + 3, v); *jt2 = dot3(cs + C_FRAME + 6, v);
  }

  // ---------------------------------------------------------------- Newton solver (mj_solNewton)
  //   min_a 1/2 a^T M a - a^T qfrc_smooth + sum_i 1/2 D_i min(0, J_i a - aref_i)^2
  // Lanes <-> dofs throughout. Row products J x are lane-local products reduced over the group
  // (contacts) or purely lane-local (limits). which = 0: x = a, stores jar = J a - aref per edge
  // and returns the gradient / round-off contributions; which = 1: x = dir, stores jv = J dir.
  // Contact loops run to ncw, the warp-wide maximum contact count (slots >= ncon are ignored).
  MMZ_DI void contact_products(const Layout& L, const float (&cd)[6], float x, int ncon, int ncw, int which, float* grad,
                               float* mag) {
    float* con = w + L.o_con;
#pragma unroll 1
    for (int c = 0; c < ncw; c++) {
      float* cs = con + c * L.cstride;
      const bool valid = c < ncon;
      float jn, jt1, jt2;
      contact_jac(cs, cd, &jn, &jt1, &jt2);
      if (!valid) { jn = 0.f; jt1 = 0.f; jt2 = 0.f; }
      const float mu = valid ? cs[C_MU] : 0.f;
      float s0 = jn * x, s1 = jt1 * x, s2 = jt2 * x, sa = (fabsf(jn) + mu * (fabsf(jt1) + fabsf(jt2))) * fabsf(x);
#pragma unroll
      for (int off = G / 2; off > 0; off >>= 1) {
        s0 += __shfl_xor_sync(kFull, s0, off);
        s1 += __shfl_xor_sync(kFull, s1, off);
        s2 += __shfl_xor_sync(kFull, s2, off);
        if (which == 0) sa += __shfl_xor_sync(kFull, sa, off);
      }
      if (!valid) continue;  // no shuffles below
      const float e0 = s0 + mu * s1, e1 = s0 - mu * s1, e2 = s0 + mu * s2, e3 = s0 - mu * s2;
      if (which == 0) {
        const float r0 = cs[C_AREF], r1 = cs[C_AREF + 1], r2 = cs[C_AREF + 2], r3 = cs[C_AREF + 3];
        const float j0 = e0 - r0, j1 = e1 - r1, j2 = e2 - r2, j3 = e3 - r3;
        if (lane == 0) { cs[C_JAR] = j0; cs[C_JAR + 1] = j1; cs[C_JAR + 2] = j2; cs[C_JAR + 3] = j3; }
        const float a0 = j0 < 0.f, a1 = j1 < 0.f, a2 = j2 < 0.f, a3 = j3 < 0.f;
        const float D = cs[C_D];
        // -J^T f with f_k = -D jar_k on the active edges
        const float f0 = a0 * D * j0, f1 = a1 * D * j1, f2 = a2 * D * j2, f3 = a3 * D * j3;
        *grad += jn * (f0 + f1 + f2 + f3) + jt1 * (mu * (f0 - f1)) + jt2 * (mu * (f2 - f3));
        const float bound = sa + fmaxf(fmaxf(fabsf(r0), fabsf(r1)), fmaxf(fabsf(r2), fabsf(r3)));
        *mag += D * bound * ((a0 + a1 + a2 + a3) * fabsf(jn) + mu * ((a0 + a1) * fabsf(jt1) + (a2 + a3) * fabsf(jt2)));
      } else if (lane == 0) {
        cs[C_JV] = e0; cs[C_JV + 1] = e1; cs[C_JV + 2] = e2; cs[C_JV + 3] = e3;
      }
    }
  }

  MMZ_DI void solve(const Layout& L, bool warmstart) {
    int* cn = cnt(L);
    const int nv = L.nv, ncon = cn[N_CON];
    const int ncw = wmax(ncon);
    float* a = w + L.o_qacc;
    const float* M = w + L.o_M;
    float* dir = w + L.o_dir;
    float* con = w + L.o_con;
    const bool me = lane < nv;
    const unsigned limbits = gballot(limD[0] > 0.f || limD[1] > 0.f);  // (collectives are never short-circuited)
    const bool constrained = ncon > 0 || limbits != 0;
    float cd[6];
#pragma unroll
    for (int k = 0; k < 6; k++) cd[k] = me ? w[L.o_cdof + 6 * lane + k] : 0.f;
    const float sm = me ? w[L.o_smooth + lane] : 0.f;
    float al = (warmstart && me) ? a[lane] : 0.f;
    if (!warmstart && me) a[lane] = 0.f;
    if (lane == 0) { cn[N_ITER] = 0; cn[N_CON_MAX] = max(cn[N_CON_MAX], ncon); }
    sync();
    // `done`: this environment's solve has finished; it keeps executing, predicated off, until every
    // environment of the warp has (uniform control flow)
    bool done = false;
#pragma unroll 1
    for (int it = 0; it < kMaxNewton; it++) {
      // ---- gradient at a; `mag` bounds the round-off of the terms it is the (cancelling) sum of
      float Ma = 0.f, mag = 0.f;
      if (me) {
#pragma unroll 2
        for (int k = 0; k < nv; k++) { float t = M[lane * L.ldm + k] * a[k]; Ma += t; mag += fabsf(t); }
      }
      float grad = Ma - sm, dadd = 0.f;
      mag += fabsf(sm);
      float ljar[2];
#pragma unroll
      for (int s = 0; s < 2; s++) {
        const float sign = s == 0 ? 1.f : -1.f;
        ljar[s] = sign * al - limA[s];
        if (limD[s] > 0.f && ljar[s] < 0.f) {
          grad += limD[s] * ljar[s] * sign;  // = -J^T force
          mag += limD[s] * (fabsf(al) + fabsf(limA[s]));
          dadd += limD[s];
        }
      }
      MMZ_CONV(10);
      contact_products(L, cd, al, ncon, ncw, 0, &grad, &mag);
      MMZ_CONV(11);
      // converged when every dof's gradient is at the fp32 round-off level of the terms it is the
      // (cancelling) sum of. The test is per dof, not on the norm: a light body (movable block,
      // 2e-4 kg) next to a heavy one would otherwise be left with a large acceleration error.
      if (gballot(fabsf(grad) > 2e-6f * mag + 1e-30f) == 0) done = true;
      if (__all_sync(kFull, done)) break;
      sync();  // the jar values lane 0 stored are read by every lane below
      // ---- Hessian row of this lane's dof: M + sum over active rows of D J^T J
      float hrow[NVP];
#pragma unroll
      for (int k = 0; k < NVP; k++) {
        float mk = (me && k < nv) ? M[lane * L.ldm + k] : 0.f;
        hrow[k] = (k == lane) ? (me ? mk + dadd : 1.f) : mk;
      }
#pragma unroll 1
      for (int c = 0; c < ncw; c++) {
        const float* cs = con + c * L.cstride;
        const bool valid = c < ncon;
        const float a0 = valid && cs[C_JAR] < 0.f, a1 = valid && cs[C_JAR + 1] < 0.f, a2 = valid && cs[C_JAR + 2] < 0.f,
                    a3 = valid && cs[C_JAR + 3] < 0.f;
        const bool act = a0 + a1 + a2 + a3 != 0.f;
        if (!__any_sync(kFull, act)) continue;
        float jn, jt1, jt2;
        contact_jac(cs, cd, &jn, &jt1, &jt2);
        if (!act) { jn = 0.f; jt1 = 0.f; jt2 = 0.f; }
        const float D = act ? cs[C_D] : 0.f, mu = act ? cs[C_MU] : 0.f;
        const float wnn = D * (a0 + a1 + a2 + a3), wn1 = D * mu * (a0 - a1), wn2 = D * mu * (a2 - a3);
        const float w11 = D * mu * mu * (a0 + a1), w22 = D * mu * mu * (a2 + a3);
        const float u0 = wnn * jn + wn1 * jt1 + wn2 * jt2, u1 = wn1 * jn + w11 * jt1, u2 = wn2 * jn + w22 * jt2;
#pragma unroll
        for (int k = 0; k < NVP; k++)
          hrow[k] += u0 * __shfl_sync(kFull, jn, k, G) + u1 * __shfl_sync(kFull, jt1, k, G) + u2 * __shfl_sync(kFull, jt2, k, G);
      }
      MMZ_CONV(12);
      const float dr = elim_solve(hrow, me ? -grad : 0.f);
      MMZ_CONV(13);
      if (me && !done) dir[lane] = dr;
      sync();
      // ---- exact line search along dir: root of the monotone piecewise-linear derivative
      float alpha = 1.f;
      int ls = 0;
      bool exact = false;
      if (__any_sync(kFull, constrained && !done)) {
        float dummy0 = 0.f, dummy1 = 0.f;
        contact_products(L, cd, dr, done ? 0 : ncon, ncw, 1, &dummy0, &dummy1);
        sync();
        float md = 0.f;
        if (me) {
#pragma unroll 2
          for (int k = 0; k < nv; k++) md += M[lane * L.ldm + k] * dir[k];
        }
        const float g0 = gsum<G>(me ? dr * (Ma - sm) : 0.f), h0 = gsum<G>(me ? dr * md : 0.f);
        float lo = 0.f, hi = -1.f;
        bool lsdone = done || !constrained;
        bool flipped = true, lsconv = false;  // did a row change sides between 0 and alpha; did the search converge
#pragma unroll 1
        for (int k = 0; k < kMaxLineSearch; k++) {
          float g = 0.f, h = 0.f;
          bool fl = false;
#pragma unroll
          for (int s = 0; s < 2; s++) {  // this lane's limit rows: jv = +-dr
            const float jv = s == 0 ? dr : -dr, x = ljar[s] + alpha * jv;
            if (limD[s] > 0.f && x < 0.f) { g += limD[s] * x * jv; h += limD[s] * jv * jv; }
            fl |= limD[s] > 0.f && (x < 0.f) != (ljar[s] < 0.f);
          }
          if (!lsdone) {
#pragma unroll 1
            for (int r = lane; r < 4 * ncon; r += G) {
              const float* cs = con + (r >> 2) * L.cstride;
              const float jv = cs[C_JV + (r & 3)], jar = cs[C_JAR + (r & 3)], x = jar + alpha * jv, D = cs[C_D];
              if (x < 0.f) { g += D * x * jv; h += D * jv * jv; }
              fl |= (x < 0.f) != (jar < 0.f);
            }
            flipped = fl;
          }
          g = gsum<G>(g) + g0 + alpha * h0;
          h = gsum<G>(h) + h0;
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
          if (__all_sync(kFull, lsdone)) break;
        }
        // no row changed sides on [0, alpha] and alpha is the full Newton step: the new point is the solution of
        // this active set, the gradient pass that would confirm it is skipped (as in the hybrid kernel)
        exact = gballot(flipped) == 0 && lsconv && fabsf(alpha - 1.f) < 1e-3f;
        if (exact) alpha = 1.f;  // the minimiser of that quadratic is the Newton step itself
      }
      bool moved = false;
      if (me && !done) {
        const float st = alpha * dr;
        moved = fabsf(st) > 2e-6f * fabsf(al) + 1e-6f;
        al += st;
        a[lane] = al;
      }
      if (lane == 0 && !done) { cn[N_ITER] = it + 1; cn[N_ITER_SUM] += 1; cn[N_LS_SUM] += ls; if (it == kMaxNewton - 1) cn[N_CAPPED] += 1; }
      sync();
      // unconstrained: one Newton step on the quadratic is exact. Otherwise stop at the fp32 floor,
      // when the step no longer changes any component of the iterate.
      const unsigned movedbits = gballot(moved);
      if (!constrained || movedbits == 0 || exact) done = true;
      if (__all_sync(kFull, done)) break;
    }
    sync();
  }

  // ---------------------------------------------------------------- mj_forward
  MMZ_DI void forward(const Layout& L, bool warmstart) {
    // Every warp of the block starts each forward evaluation together: the step is instruction-fetch
    // bound (a ~100 KB loop body against a much smaller instruction cache), and warps that walk the
    // same code at the same time share the fetched lines. All blocks are full (npad is a multiple of
    // the block's environment count) and every warp runs the same number of evaluations.
    if (bsync & 1) __syncthreads();
    MMZ_CONV(1);
    kinematics(L);
    MMZ_CONV(2);
    motion_axes(L);
    mass_matrix(L);
    MMZ_CONV(3);
    collision(L);
    MMZ_CONV(4);
    bias_forces(L);
    smooth_forces(L);
    MMZ_CONV(5);
    make_constraints(L);
    if (bsync & 2) __syncthreads();
    MMZ_CONV(6);
    solve(L, warmstart);
    MMZ_CONV(7);
  }

  MMZ_DI bool state_bad(const Layout& L) const {
    bool bad = false;
    for (int i = lane; i < L.nq; i += G) bad |= !(fabsf(w[L.o_qpos + i]) < kMaxVal);
    for (int i = lane; i < L.nv; i += G) bad |= !(fabsf(w[L.o_qvel + i]) < kMaxVal);
    return gballot(bad) != 0;
  }

  // ---------------------------------------------------------------- mj_step, RK4 (mj_RungeKutta)
  // Classic tableau; stages 2-4 re-run the whole forward at x0 + h * A * k. Positions integrate on
  // the configuration manifold (mj_integratePos). Returns true if the state blew up (MuJoCo would
  // auto-reset). One forward() call site: pass i = 0..3 evaluates stage i, pass 4 only combines.
  MMZ_DI bool mj_step(const Layout& L, bool dead) {
    const float h = m->timestep;
    const int nq = L.nq, nv = L.nv;
    float *qpos = w + L.o_qpos, *qvel = w + L.o_qvel, *qacc = w + L.o_qacc;
    float *q0 = w + L.o_q0, *v0 = w + L.o_v0, *xv = w + L.o_xv, *fa = w + L.o_fa, *accv = w + L.o_accv, *acca = w + L.o_acca;
    // An environment whose state is not finite is parked at qpos0 with zero velocity and keeps
    // executing (its results are discarded by the caller): every warp-level primitive below needs
    // all environments of the warp to run the same control flow.
    if (bsync & 4) __syncthreads();
    const bool startbad = state_bad(L);
    bool bad = dead || startbad;
    if (bad) park(L);
    sync();
    for (int i = lane; i < nq; i += G) q0[i] = qpos[i];
    for (int d = lane; d < nv; d += G) { v0[d] = qvel[d]; accv[d] = 0.f; acca[d] = 0.f; }
    sync();
#pragma unroll 1
    for (int i = 0; i < 5; i++) {
      if (i > 0) {
        // state of stage i (or the final combination when i == 4)
        const float A = (i == 1 || i == 2) ? 0.5f : 1.f;
        const float* vsrc = (i == 4) ? accv : xv;
        const float* asrc = (i == 4) ? acca : fa;
#pragma unroll 1
        for (int j = lane; j < L.nj; j += G) {
          int qa = m->jnt_qadr[j], d = m->jnt_dadr[j];
          if (m->jnt_type[j] == MMZ_JNT_FREE) {
#pragma unroll
            for (int k = 0; k < 3; k++) qpos[qa + k] = q0[qa + k] + h * A * vsrc[d + k];
            float wv[3] = {A * vsrc[d + 3], A * vsrc[d + 4], A * vsrc[d + 5]};
            float q[4] = {q0[qa + 3], q0[qa + 4], q0[qa + 5], q0[qa + 6]};
            float nw = norm3(wv), ang = h * nw;
            quat_norm(q);
            if (ang > 0.f) {
              float inv = 1.f / nw, ax[3] = {wv[0] * inv, wv[1] * inv, wv[2] * inv}, qr[4], q2[4];
              axisangle2quat(qr, ax, ang);
              quat_mul(q2, q, qr);
#pragma unroll
              for (int k = 0; k < 4; k++) q[k] = q2[k];
            }
#pragma unroll
            for (int k = 0; k < 4; k++) qpos[qa + 3 + k] = q[k];
          } else {
            qpos[qa] = q0[qa] + h * A * vsrc[d];
          }
        }
        for (int d = lane; d < nv; d += G) qvel[d] = v0[d] + h * A * asrc[d];
        sync();
        if (i == 4) break;
      }
      forward(L, true);
      const float B = (i == 0 || i == 3) ? (1.f / 6.f) : (1.f / 3.f);
      bool badacc = false;
      for (int d = lane; d < nv; d += G) {
        float v = qvel[d], f = qacc[d];
        badacc |= !(fabsf(f) < kMaxVal);
        if (badacc) f = 0.f;
        xv[d] = v; fa[d] = f;
        accv[d] += B * v; acca[d] += B * f;
      }
      if (gballot(badacc)) bad = true;  // keeps integrating (finite garbage), discarded by the caller
      sync();
    }
    // derived arrays (xpos, contacts) deliberately stay at the 4th-stage state: SURVEY quirk Q15
    const bool endbad = state_bad(L);
    return bad || endbad;
  }

  // qpos0, zero velocity and acceleration (keeps a blown-up environment finite); the caller syncs
  MMZ_DI void park(const Layout& L) {
    for (int i = lane; i < L.nq; i += G) w[L.o_qpos + i] = m->qpos0[i];
    for (int d = lane; d < L.nv; d += G) { w[L.o_qvel + d] = 0.f; w[L.o_qacc + d] = 0.f; }
  }
};

}  // namespace mmz
