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
#include "mmz_layout.h"
#include "mmz_narrow.cuh"

namespace mmz {

template <int G>
MMZ_DI float gsum(float v, unsigned mask) {
#pragma unroll
  for (int off = G / 2; off > 0; off >>= 1) v += __shfl_xor_sync(mask, v, off);
  return v;
}
template <int G>
MMZ_DI int gmax(int v, unsigned mask) {
#pragma unroll
  for (int off = G / 2; off > 0; off >>= 1) v = max(v, __shfl_xor_sync(mask, v, off));
  return v;
}
// inclusive prefix sum over the group; *total = sum over all lanes
template <int G>
MMZ_DI int gscan(int v, int lane, unsigned mask, int* total) {
#pragma unroll
  for (int off = 1; off < G; off <<= 1) {
    int t = __shfl_up_sync(mask, v, off, G);
    if (lane >= off) v += t;
  }
  *total = __shfl_sync(mask, v, G - 1, G);
  return v;
}

constexpr int kMaxNewton = 24;
constexpr int kMaxLineSearch = 24;

template <int G, int NVP>
struct Env {
  const mmz_model* m;  // shared memory
  const Derived* dv;   // shared memory
  float* w;            // this environment's workspace (shared memory)
  int lane;            // 0..G-1
  unsigned gmask;      // lanes of this group inside the warp

  MMZ_DI void sync() const { __syncwarp(gmask); }
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
    const bool fluid = m->density > 0.f || m->viscosity > 0.f;
    if (fluid) fluid_forces(L);  // wrenches land in frc, which the subtree sums have consumed
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
      const int maxcell = gmax<G>(ncell, gmask);
      const int ncand = 1 + nslot * maxcell + dv->nboxg;
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
          } else {  // box geoms on other moving bodies; geom1 = this geom, geom2 = box
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
        unsigned any = __ballot_sync(gmask, n > 0);
        if (!any) continue;
        int total, incl = gscan<G>(n, lane, gmask, &total);  // keeps the contact order deterministic
        int base = ncon + incl - n;
        if (n > 0 && base < L.maxcon) write_contact(L, base, r0, b1, b2, iw, g, other);
        if (n > 1 && base + 1 < L.maxcon) write_contact(L, base + 1, r1, b1, b2, iw, g, other);
        ncon += total;
        if (ncon > L.maxcon) { ncon = L.maxcon; overflow = true; }
      }
    }
    if (lane == 0) { cn[N_CON] = ncon; cn[N_OVERFLOW] = overflow ? 1 : 0; }
    sync();
    // ---- box geoms (Point's arrow, movable blocks): few, handled by one lane each in order
#pragma unroll 1
    for (int k = 0; k < dv->nboxg; k++) {
      int g = dv->boxg[k];
      if (lane == 0 && ((m->geom_contype[g] | m->geom_conaffinity[g]) & 1)) box_geom_contacts(L, g, k);
      sync();
    }
  }

  __device__ __noinline__ void box_geom_contacts(const Layout& L, int g, int kbox) {
    int* cn = cnt(L);
    int ncon = cn[N_CON];
    const float* gp = w + L.o_gpos + 3 * g;
    const float* gm = w + L.o_gmat + 9 * g;
    const float* sz = m->geom_size[g];
    const int body = m->geom_body[g];
    const float invw = m->geom_invweight[g];
    RawContact rc[8];
    float ext[3], wallmargin = fmaxf(m->geom_margin[g], m->wall_margin);
#pragma unroll
    for (int k = 0; k < 3; k++) ext[k] = fabsf(gm[3 * k]) * sz[0] + fabsf(gm[3 * k + 1]) * sz[1] + fabsf(gm[3 * k + 2]) * sz[2];
    ext[0] += wallmargin; ext[1] += wallmargin;
    int i0, i1, j0, j1;
    cell_range(gp, ext, &i0, &i1, &j0, &j1);
    const int nj = max(0, j1 - j0 + 1), ncell = nj * max(0, i1 - i0 + 1);
    // candidates: floor, then (cell, wall|platform) pairs, then later box geoms on other bodies
    const int ncand = 1 + 2 * ncell + (dv->nboxg - kbox - 1);
    for (int cand = 0; cand < ncand; cand++) {
      int n = 0, b1 = -1, b2 = body, other = -1;
      float iw = invw;
      if (cand == 0) {  // corners below the plane, at most 4; geom1 = plane
        if (!m->has_floor) continue;
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
      } else if (cand - 1 < 2 * ncell) {  // maze boxes; geom1 = wall (lower geom id), geom2 = this box
        int ci = (cand - 1) >> 1, slot = (cand - 1) & 1;
        int i = i0 + ci / nj, j = j0 + ci % nj;
        int code = m->grid[i * m->grid_w + j];
        if (!(code & (slot == 0 ? MMZ_CELL_WALL : MMZ_CELL_PLATFORM))) continue;
        float bc[3] = {j * m->cell_size - m->origin[0], i * m->cell_size - m->origin[1], slot == 0 ? m->wall_z : m->plat_z};
        other = -2;
        n = box_box(bc, dv->ident, m->wall_half, gp, gm, sz, wallmargin, rc);
      } else {  // box against box on different moving bodies
        int g2 = dv->boxg[kbox + 1 + (cand - 1 - 2 * ncell)];
        if (!moving_pair_ok(g, g2)) continue;
        other = g2; b1 = body; b2 = m->geom_body[g2];
        iw = invw + m->geom_invweight[g2];
        n = box_box(gp, gm, sz, w + L.o_gpos + 3 * g2, w + L.o_gmat + 9 * g2, m->geom_size[g2],
                    fmaxf(m->geom_margin[g], m->geom_margin[g2]), rc);
      }
      for (int k = 0; k < n; k++) {
        if (ncon < L.maxcon) write_contact(L, ncon++, rc[k], b1, b2, iw, g, other);
        else cn[N_OVERFLOW] = 1;
      }
    }
    cn[N_CON] = ncon;
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

  MMZ_DI void make_constraints(const Layout& L) {
    int* cn = cnt(L);
    const float* qpos = w + L.o_qpos;
    const float* qvel = w + L.o_qvel;
    const float* cdof = w + L.o_cdof;
    const int ncon = cn[N_CON];
    // One pass over "row sources": indices [0, 2 nj) are (joint, side) limit candidates, the rest
    // are contacts, so row_params / impedance have a single call site. Limit rows are compacted
    // in (joint, lower-then-upper) order with a group scan.
    int nlim = 0;
    const int nsrc = 2 * L.nj + ncon;
#pragma unroll 1
    for (int base = 0; base < nsrc; base += G) {
      const int idx = base + lane;
      const bool is_lim = idx < 2 * L.nj, is_con = !is_lim && idx < nsrc;
      // ---- limit candidate
      int j = idx >> 1, side = idx & 1, n = 0, d = 0;
      float pos = 0.f, margin = 0.f, diag = 0.f, mu = 0.f;
      const float *solref = m->wall_solref, *solimp = m->wall_solimp;
      float* cs = nullptr;
      float jv0 = 0.f, jv1 = 0.f, jv2 = 0.f;
      if (is_lim && m->jnt_limited[j]) {
        float q = qpos[m->jnt_qadr[j]];
        pos = side == 0 ? q - m->jnt_range[j][0] : m->jnt_range[j][1] - q;
        margin = m->jnt_margin[j];
        n = pos < margin;
        d = m->jnt_dadr[j];
        diag = m->dof_invweight0[d];
        solref = m->jnt_solref[j]; solimp = m->jnt_solimp[j];
      }
      if (is_con) {  // frame Jacobian of the relative point velocity (body2 - body1)
        cs = w + L.o_con + (idx - 2 * L.nj) * L.cstride;
        float* J = cs + C_J;
        float cp[3], fr[9];
#pragma unroll
        for (int k = 0; k < 3; k++) cp[k] = cs[C_POS + k] - w[L.o_xpos + k];
#pragma unroll
        for (int k = 0; k < 9; k++) fr[k] = cs[C_FRAME + k];
        int b1 = __float_as_int(cs[C_BODY1]), b2 = __float_as_int(cs[C_BODY2]);
        int mask1 = b1 >= 0 ? m->body_dofmask[b1] : 0, mask2 = b2 >= 0 ? m->body_dofmask[b2] : 0;
#pragma unroll 1
        for (int dd = 0; dd < L.nv; dd++) {
          float s = (float)(mask2 >> dd & 1) - (float)(mask1 >> dd & 1);
          float jn = 0.f, jt1 = 0.f, jt2 = 0.f;
          if (s != 0.f) {
            const float* cd = cdof + 6 * dd;
            float wxp[3];
            cross3(wxp, cd, cp);
            float v[3] = {cd[3] + wxp[0], cd[4] + wxp[1], cd[5] + wxp[2]};
            jn = s * dot3(fr, v); jt1 = s * dot3(fr + 3, v); jt2 = s * dot3(fr + 6, v);
            float qv = qvel[dd];
            jv0 += jn * qv; jv1 += jt1 * qv; jv2 += jt2 * qv;
          }
          J[dd] = jn; J[L.nv + dd] = jt1; J[2 * L.nv + dd] = jt2;
        }
        mu = cs[C_MU];
        pos = cs[C_DIST]; margin = cs[C_MARGIN]; diag = cs[C_INVW] * (1.f + mu * mu);
        solref = cs + C_SOLREF; solimp = cs + C_SOLIMP;
      }
      int slot = 0;
      if (base < 2 * L.nj) {  // chunks that contain limit candidates
        if (__ballot_sync(gmask, n > 0)) {
          int total;
          slot = nlim + gscan<G>(n, lane, gmask, &total) - n;
          nlim = min(nlim + total, L.maxlim);
        }
      }
      if ((n && slot < L.maxlim) || is_con) {
        float D, kr, bb;
        row_params(solref, solimp, pos, margin, diag, &D, &kr, &bb);
        if (is_con) {
          // all edges of the pyramid share R = 2 mu^2 R_first
          cs[C_DIST] = 1.f / fmaxf(kMinVal, 2.f * mu * mu / D);
          cs[C_AREF + 0] = -bb * (jv0 + mu * jv1) - kr;
          cs[C_AREF + 1] = -bb * (jv0 - mu * jv1) - kr;
          cs[C_AREF + 2] = -bb * (jv0 + mu * jv2) - kr;
          cs[C_AREF + 3] = -bb * (jv0 - mu * jv2) - kr;
        } else {
          float sign = side == 0 ? 1.f : -1.f;
          float* r = w + L.o_lim + slot * R_STRIDE;
          r[R_DOF] = __int_as_float(d);
          r[R_SIGN] = sign;
          r[R_D] = D;
          r[R_AREF] = -bb * sign * qvel[d] - kr;
        }
      }
    }
    if (lane == 0) cn[N_LIM] = nlim;
    sync();
  }

  // ---------------------------------------------------------------- dense Cholesky in shared memory
  // Lane i holds row i of a symmetric positive-definite matrix in registers; the rows are stored to
  // shared memory and factored in place by a ROLLED right-looking loop (lanes <-> rows): on return L
  // (lower) is at Lsm[i*ldm + k], k <= i. Rolled on purpose: a fully unrolled register version is
  // ~1500 instructions of straight-line code that every Newton iteration has to fetch.
  MMZ_DI void chol_rows(const Layout& L, const float (&row)[NVP], float* Lsm) {
    const int nv = L.nv, ld = L.ldm;
    float* my = Lsm + lane * ld;
    if (lane < nv) {
#pragma unroll
      for (int k = 0; k < NVP; k++)
        if (k <= lane) my[k] = row[k];
    }
    sync();
#pragma unroll 1
    for (int j = 0; j < nv; j++) {
      const float piv = fmaxf(Lsm[j * ld + j], kMinVal);
      const float inv = rsqrtf(piv);
      float lij = 0.f;
      if (lane >= j && lane < nv) {
        lij = (lane == j) ? piv * inv : my[j] * inv;
        my[j] = lij;
      }
      sync();
      if (lane > j && lane < nv) {
#pragma unroll 2
        for (int k = j + 1; k <= lane; k++) my[k] -= lij * Lsm[k * ld + j];
      }
      sync();
    }
  }
  // x <- (L L^T)^-1 x for the vector held one element per lane
  MMZ_DI float chol_solve(const Layout& L, const float* Lsm, float x) {
    const int nv = L.nv;
    float invd = (lane < nv) ? 1.f / Lsm[lane * L.ldm + lane] : 0.f;
    float acc = x;
#pragma unroll 1
    for (int k = 0; k < nv; k++) {
      float yk = __shfl_sync(gmask, acc * invd, k, G);
      if (lane > k && lane < nv) acc -= Lsm[lane * L.ldm + k] * yk;
      if (lane == k) acc = yk;
    }
#pragma unroll 1
    for (int k = nv - 1; k >= 0; k--) {
      float xk = __shfl_sync(gmask, acc * invd, k, G);
      if (lane < k) acc -= Lsm[k * L.ldm + lane] * xk;
      if (lane == k) acc = xk;
    }
    return acc;
  }

  // ---------------------------------------------------------------- Newton solver (mj_solNewton)
  //   min_a 1/2 a^T M a - a^T qfrc_smooth + sum_i 1/2 D_i min(0, J_i a - aref_i)^2
  // Row products J x for the limit rows and the 4 pyramid edges of every contact. which = 0:
  // jar = J x - aref, with the magnitude of the cancelling terms (for the round-off floor of the
  // convergence test) parked in the jv slot; which = 1: jv = J x.
  MMZ_DI void row_products(const Layout& L, const float* x, int nlim, int ncon, int which) {
    float* lim = w + L.o_lim;
#pragma unroll 1
    for (int r = lane; r < nlim; r += G) {
      float* lr = lim + r * R_STRIDE;
      float jx = lr[R_SIGN] * x[__float_as_int(lr[R_DOF])];
      if (which == 0) { lr[R_JAR] = jx - lr[R_AREF]; lr[R_JV] = fabsf(jx) + fabsf(lr[R_AREF]); }
      else lr[R_JV] = jx;
    }
#pragma unroll 1
    for (int c = lane; c < ncon; c += G) {
      float* cs = w + L.o_con + c * L.cstride;
      const float* J = cs + C_J;
      const float mu = cs[C_MU];
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, sa = 0.f;
#pragma unroll 2
      for (int d = 0; d < L.nv; d++) {
        float xd = x[d], jn = J[d], j1 = J[L.nv + d], j2 = J[2 * L.nv + d];
        s0 += jn * xd; s1 += j1 * xd; s2 += j2 * xd;
        sa += (fabsf(jn) + mu * (fabsf(j1) + fabsf(j2))) * fabsf(xd);
      }
      float e0 = s0 + mu * s1, e1 = s0 - mu * s1, e2 = s0 + mu * s2, e3 = s0 - mu * s2;
      if (which == 0) {
        float a0 = cs[C_AREF], a1 = cs[C_AREF + 1], a2 = cs[C_AREF + 2], a3 = cs[C_AREF + 3];
        cs[C_JAR + 0] = e0 - a0; cs[C_JAR + 1] = e1 - a1; cs[C_JAR + 2] = e2 - a2; cs[C_JAR + 3] = e3 - a3;
        cs[C_JV] = sa + fmaxf(fmaxf(fabsf(a0), fabsf(a1)), fmaxf(fabsf(a2), fabsf(a3)));
      } else {
        cs[C_JV + 0] = e0; cs[C_JV + 1] = e1; cs[C_JV + 2] = e2; cs[C_JV + 3] = e3;
      }
    }
    sync();
  }

  MMZ_DI void solve(const Layout& L, bool warmstart) {
    int* cn = cnt(L);
    const int nv = L.nv, ncon = cn[N_CON], nlim = cn[N_LIM], nrow = nlim + 4 * ncon;
    float* a = w + L.o_qacc;
    const float* M = w + L.o_M;
    const float* smooth = w + L.o_smooth;
    float* H = w + L.o_H;
    float* dir = w + L.o_dir;
    float* lim = w + L.o_lim;
    float* con = w + L.o_con;
    const bool me = lane < nv;
    if (!warmstart && me) a[lane] = 0.f;
    if (lane == 0) { cn[N_ITER] = 0; cn[N_CON_MAX] = max(cn[N_CON_MAX], ncon); }
    sync();
    // pass 2k evaluates jar = J a - aref, pass 2k+1 (after the Newton direction is known) jv = J dir:
    // a single row_products call site serves both
    float hrow[NVP];
    float Ma = 0.f, grad = 0.f, sm = 0.f, dr = 0.f;
    bool stop = false;
#pragma unroll 1
    for (int pass = 0; pass < 2 * kMaxNewton && !stop; pass++) {
      const int it = pass >> 1;
      if (nrow) row_products(L, (pass & 1) ? dir : a, nlim, ncon, pass & 1);
      if (!(pass & 1)) {
        // gradient and Hessian row of this lane's dof; `mag` bounds the round-off of the gradient
        float mag = 0.f;
        Ma = 0.f;
#pragma unroll
        for (int k = 0; k < NVP; k++) {
          float mk = (me && k < nv) ? M[lane * L.ldm + k] : ((k == lane) ? 1.f : 0.f);
          hrow[k] = mk;
          if (k < nv) { float t = mk * a[k]; Ma += t; mag += fabsf(t); }
        }
        if (!me) { Ma = 0.f; mag = 0.f; }
        sm = me ? smooth[lane] : 0.f;
        grad = Ma - sm;
        mag += fabsf(sm);
        float dadd = 0.f;
#pragma unroll 1
        for (int r = 0; r < nlim; r++) {
          const float* lr = lim + r * R_STRIDE;
          float jar = lr[R_JAR];
          if (jar < 0.f && __float_as_int(lr[R_DOF]) == lane) {
            grad += lr[R_D] * jar * lr[R_SIGN];  // = -J^T force
            mag += lr[R_D] * lr[R_JV];
            dadd += lr[R_D];
          }
        }
#pragma unroll 1
        for (int c = 0; c < ncon; c++) {
          const float* cs = con + c * L.cstride;
          float j0 = cs[C_JAR], j1 = cs[C_JAR + 1], j2 = cs[C_JAR + 2], j3 = cs[C_JAR + 3];
          float a0 = j0 < 0.f, a1 = j1 < 0.f, a2 = j2 < 0.f, a3 = j3 < 0.f;
          if (a0 + a1 + a2 + a3 == 0.f) continue;
          float D = cs[C_DIST], mu = cs[C_MU];
          const float* J = cs + C_J;
          float jn = me ? J[lane] : 0.f, jt1 = me ? J[nv + lane] : 0.f, jt2 = me ? J[2 * nv + lane] : 0.f;
          // -J^T f with f_k = -D jar_k on active edges
          float f0 = a0 * D * j0, f1 = a1 * D * j1, f2 = a2 * D * j2, f3 = a3 * D * j3;
          grad += jn * (f0 + f1 + f2 + f3) + jt1 * (mu * (f0 - f1)) + jt2 * (mu * (f2 - f3));
          mag += D * cs[C_JV] * ((a0 + a1 + a2 + a3) * fabsf(jn) + mu * ((a0 + a1) * fabsf(jt1) + (a2 + a3) * fabsf(jt2)));
          // H += Jc^T W Jc, W from the active edges
          float wnn = D * (a0 + a1 + a2 + a3), wn1 = D * mu * (a0 - a1), wn2 = D * mu * (a2 - a3);
          float w11 = D * mu * mu * (a0 + a1), w22 = D * mu * mu * (a2 + a3);
          float u0 = wnn * jn + wn1 * jt1 + wn2 * jt2, u1 = wn1 * jn + w11 * jt1, u2 = wn2 * jn + w22 * jt2;
#pragma unroll
          for (int k = 0; k < NVP; k++)
            if (k < nv) hrow[k] += u0 * J[k] + u1 * J[nv + k] + u2 * J[2 * nv + k];
        }
#pragma unroll
        for (int k = 0; k < NVP; k++)
          if (k == lane) hrow[k] += dadd;
        // converged when every dof's gradient is at the fp32 round-off level of the terms it is the
        // (cancelling) sum of. The test is per dof, not on the norm: a light body (movable block,
        // 2e-4 kg) next to a heavy one would otherwise be left with a large acceleration error.
        if (__ballot_sync(gmask, fabsf(grad) > 2e-6f * mag + 1e-30f) == 0) break;
        chol_rows(L, hrow, H);
        dr = chol_solve(L, H, -grad);
        if (me) dir[lane] = dr;
        sync();
      } else {
        float alpha = 1.f;
        int ls = 0;
        if (nrow) {
          // exact line search along dir: root of the monotone piecewise-linear derivative
          float md = 0.f;
          if (me) {
#pragma unroll 4
            for (int k = 0; k < nv; k++) md += M[lane * L.ldm + k] * dir[k];
          }
          const float g0 = gsum<G>(me ? dr * (Ma - sm) : 0.f, gmask), h0 = gsum<G>(me ? dr * md : 0.f, gmask);
          float lo = 0.f, hi = -1.f;
#pragma unroll 1
          for (; ls < kMaxLineSearch; ls++) {
            float g = 0.f, h = 0.f;
#pragma unroll 1
            for (int r = lane; r < nrow; r += G) {
              float jar, jv, D;
              if (r < nlim) { const float* lr = lim + r * R_STRIDE; jar = lr[R_JAR]; jv = lr[R_JV]; D = lr[R_D]; }
              else { const float* cs = con + ((r - nlim) >> 2) * L.cstride; int e = (r - nlim) & 3; jar = cs[C_JAR + e]; jv = cs[C_JV + e]; D = cs[C_DIST]; }
              float x = jar + alpha * jv;
              if (x < 0.f) { g += D * x * jv; h += D * jv * jv; }
            }
            g = gsum<G>(g, gmask) + g0 + alpha * h0;
            h = gsum<G>(h, gmask) + h0;
            if (fabsf(g) < 1e-6f * fmaxf(1e-6f, fabsf(g0))) break;
            if (g < 0.f) lo = alpha; else hi = alpha;
            float next = alpha - g / h;
            if (hi >= 0.f && (next <= lo || next >= hi)) next = 0.5f * (lo + hi);
            if (next <= lo && hi < 0.f) next = 2.f * alpha + 1e-6f;
            if (next == alpha) break;
            alpha = next;
          }
        }
        bool moved = false;
        if (me) {
          float av = a[lane], st = alpha * dr;
          a[lane] = av + st;
          moved = fabsf(st) > 2e-6f * fabsf(av) + 1e-6f;
        }
        if (lane == 0) { cn[N_ITER] = it + 1; cn[N_ITER_SUM] += 1; cn[N_LS_SUM] += ls; if (it == kMaxNewton - 1) cn[N_CAPPED] += 1; }
        sync();
        // unconstrained: one Newton step on the quadratic is exact. Otherwise stop at the fp32 floor,
        // when the step no longer changes any component of the iterate.
        stop = nrow == 0 || __ballot_sync(gmask, moved) == 0;
      }
    }
    sync();
  }

  // ---------------------------------------------------------------- mj_forward
  MMZ_DI void forward(const Layout& L, bool warmstart) {
    kinematics(L);
    motion_axes(L);
    mass_matrix(L);
    collision(L);
    bias_forces(L);
    smooth_forces(L);
    make_constraints(L);
    solve(L, warmstart);
  }

  MMZ_DI bool state_bad(const Layout& L) const {
    bool bad = false;
    for (int i = lane; i < L.nq; i += G) bad |= !(fabsf(w[L.o_qpos + i]) < kMaxVal);
    for (int i = lane; i < L.nv; i += G) bad |= !(fabsf(w[L.o_qvel + i]) < kMaxVal);
    return __ballot_sync(gmask, bad) != 0;
  }

  // ---------------------------------------------------------------- mj_step, RK4 (mj_RungeKutta)
  // Classic tableau; stages 2-4 re-run the whole forward at x0 + h * A * k. Positions integrate on
  // the configuration manifold (mj_integratePos). Returns true if the state blew up (MuJoCo would
  // auto-reset). One forward() call site: pass i = 0..3 evaluates stage i, pass 4 only combines.
  MMZ_DI bool mj_step(const Layout& L) {
    const float h = m->timestep;
    const int nq = L.nq, nv = L.nv;
    float *qpos = w + L.o_qpos, *qvel = w + L.o_qvel, *qacc = w + L.o_qacc;
    float *q0 = w + L.o_q0, *v0 = w + L.o_v0, *xv = w + L.o_xv, *fa = w + L.o_fa, *accv = w + L.o_accv, *acca = w + L.o_acca;
    if (state_bad(L)) return true;
    for (int i = lane; i < nq; i += G) q0[i] = qpos[i];
    for (int d = lane; d < nv; d += G) { v0[d] = qvel[d]; accv[d] = 0.f; acca[d] = 0.f; }
    sync();
    bool bad = false;
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
        xv[d] = v; fa[d] = f;
        accv[d] += B * v; acca[d] += B * f;
      }
      sync();
      if (i == 0 && __ballot_sync(gmask, badacc)) { bad = true; break; }
    }
    // derived arrays (xpos, contacts) deliberately stay at the 4th-stage state: SURVEY quirk Q15
    return bad || state_bad(L);
  }
};

}  // namespace mmz
