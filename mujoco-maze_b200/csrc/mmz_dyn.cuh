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
// the factorisations), lanes <-> contacts / constraint rows in the solver. Arrays that other
// lanes read at data-dependent indices live in the group's shared-memory workspace (mmz_layout.h).
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
    for (int lvl = 0; lvl < dv->nlev; lvl++) {
      for (int b = lane; b < L.nb; b += G)
        if (m->body_level[b] == lvl) body_kin(L, b);
      sync();
    }
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
    for (int j = lane; j < L.nj; j += G) {
      int d = m->jnt_dadr[j], b = m->jnt_body[j], type = m->jnt_type[j];
      float* c = w + L.o_cdof + 6 * d;
      const float* axis = w + L.o_xaxis + 3 * j;
      if (type == MMZ_JNT_FREE) {
        const float* xm = w + L.o_xmat + 9 * b;
        float xp[3];
        rel(L, xp, w + L.o_xpos + 3 * b);
#pragma unroll
        for (int k = 0; k < 3; k++) {
#pragma unroll
          for (int i = 0; i < 6; i++) c[6 * k + i] = (i == 3 + k) ? 1.f : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 3; k++) {
          float ax[3] = {xm[k], xm[3 + k], xm[6 + k]}, lin[3];
          cross3(lin, xp, ax);
          float* r = c + 6 * (3 + k);
          r[0] = ax[0]; r[1] = ax[1]; r[2] = ax[2]; r[3] = lin[0]; r[4] = lin[1]; r[5] = lin[2];
        }
      } else if (type == MMZ_JNT_SLIDE) {
        c[0] = c[1] = c[2] = 0.f;
        c[3] = axis[0]; c[4] = axis[1]; c[5] = axis[2];
      } else {
        float lin[3], an[3];
        rel(L, an, w + L.o_xanchor + 3 * j);
        cross3(lin, an, axis);
        c[0] = axis[0]; c[1] = axis[1]; c[2] = axis[2]; c[3] = lin[0]; c[4] = lin[1]; c[5] = lin[2];
      }
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
    for (int b = lane; b < L.nb; b += G) {
      float I[10];
      body_inertia_world(L, b, I);
#pragma unroll
      for (int k = 0; k < 10; k++) Iw[10 * b + k] = I[k];
    }
    for (int i = lane; i < L.nv * L.ldm; i += G) M[i] = 0.f;
    sync();
    for (int b = lane; b < L.nb; b += G) {  // composite inertia of the subtree rooted at b
      float I[10];
#pragma unroll
      for (int k = 0; k < 10; k++) I[k] = Iw[10 * b + k];
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
    for (int i = lane; i < L.nv; i += G) {
      float f[6];
      inert_mul(f, Ic + 10 * m->dof_body[i], cdof + 6 * i);
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
    for (int lvl = 0; lvl < dv->nlev; lvl++) {
      for (int b = lane; b < L.nb; b += G) {
        if (m->body_level[b] != lvl) continue;
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
        int j0 = m->body_jntadr[b], j1 = j0 + m->body_jntnum[b];
        for (int j = j0; j < j1; j++) {
          int d = m->jnt_dadr[j];
          if (m->jnt_type[j] == MMZ_JNT_FREE) {
#pragma unroll
            for (int k = 0; k < 3; k++) v[3 + k] += qvel[d + k];  // world-aligned translation axes
            // all three axis derivatives are taken against the same v (after translation, before rotation)
            float sdk[3][6];
#pragma unroll
            for (int k = 0; k < 3; k++) cross_motion(sdk[k], v, cdof + 6 * (d + 3 + k));
#pragma unroll
            for (int k = 0; k < 3; k++) {
              float qv = qvel[d + 3 + k];
              const float* s = cdof + 6 * (d + 3 + k);
#pragma unroll
              for (int i = 0; i < 6; i++) { a[i] += sdk[k][i] * qv; v[i] += s[i] * qv; }
            }
          } else {
            float sd[6];
            const float* s = cdof + 6 * d;
            float qv = qvel[d];
            cross_motion(sd, v, s);
#pragma unroll
            for (int i = 0; i < 6; i++) { a[i] += sd[i] * qv; v[i] += s[i] * qv; }
          }
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
    for (int b = lane; b < L.nb; b += G) {
      float f[6];
#pragma unroll
      for (int k = 0; k < 6; k++) f[k] = frc[6 * b + k];
      for (int c = b + 1; c < L.nb; c++)
        if (dv->anc[c] >> b & 1) {
#pragma unroll
          for (int k = 0; k < 6; k++) f[k] += frc[6 * c + k];
        }
#pragma unroll
      for (int k = 0; k < 6; k++) fsub[6 * b + k] = f[k];
    }
    sync();
    for (int d = lane; d < L.nv; d += G) w[L.o_bias + d] = dot6(cdof + 6 * d, fsub + 6 * m->dof_body[d]);
    sync();
  }

  // ---------------------------------------------------------------- damping + fluid forces (mj_passive)
  MMZ_DI void passive_forces(const Layout& L) {
    const float* cdof = w + L.o_cdof;
    const float* qvel = w + L.o_qvel;
    const bool fluid = m->density > 0.f || m->viscosity > 0.f;
    float* sf = w + L.o_frc;  // reuse: per-body fluid wrench about the origin
    if (fluid) {
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
    for (int d = lane; d < L.nv; d += G) {
      float p = -m->dof_damping[d] * qvel[d];
      if (fluid)
        for (int b = 0; b < L.nb; b++)
          if (m->body_dofmask[b] >> d & 1) p += dot6(cdof + 6 * d, sf + 6 * b);
      w[L.o_passive + d] = p;
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
  MMZ_DI void write_contact(const Layout& L, int slot, const RawContact& rc, int b1, int b2, float invw,
                            const float* par) {
    float* c = w + L.o_con + slot * L.cstride;
    float fr[9];
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
  MMZ_DI void capsule_ends(const Layout& L, int g, float* p0, float* p1) const {
    const float* gm = w + L.o_gmat + 9 * g;
    const float* gp = w + L.o_gpos + 3 * g;
    float hl = m->geom_size[g][1];
#pragma unroll
    for (int k = 0; k < 3; k++) {
      p0[k] = gp[k] + gm[3 * k + 2] * hl;
      p1[k] = gp[k] - gm[3 * k + 2] * hl;
    }
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
    const float ident[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    const int nslot = m->elevated ? 2 : 1;  // wall box, platform box per cell
    // ---- spheres and capsules: one lane per geom, candidates walked in lock step
    for (int gbase = 0; gbase < L.ng; gbase += G) {
      int g = gbase + lane;
      int type = g < L.ng ? m->geom_type[g] : -1;
      bool valid = (type == MMZ_GEOM_SPHERE || type == MMZ_GEOM_CAPSULE) &&
                   ((m->geom_contype[g] | m->geom_conaffinity[g]) & 1);
      float r = 0.f, p0[3] = {0, 0, 0}, p1[3] = {0, 0, 0};
      int i0 = 0, i1 = -1, j0 = 0, j1 = -1, nj = 1, ncell = 0, body = -1;
      float invw = 0.f;
      if (valid) {
        r = m->geom_size[g][0];
        body = m->geom_body[g];
        invw = m->geom_invweight[g];
        const float* gp = w + L.o_gpos + 3 * g;
        float ext[3];
        if (type == MMZ_GEOM_CAPSULE) {
          capsule_ends(L, g, p0, p1);
#pragma unroll
          for (int k = 0; k < 3; k++) ext[k] = fabsf(p0[k] - gp[k]);
        } else {
#pragma unroll
          for (int k = 0; k < 3; k++) { p0[k] = gp[k]; ext[k] = 0.f; }
        }
        float mg = r + fmaxf(m->geom_margin[g], m->wall_margin);
        ext[0] += mg; ext[1] += mg;
        cell_range(gp, ext, &i0, &i1, &j0, &j1);
        nj = max(1, j1 - j0 + 1);
        ncell = max(0, j1 - j0 + 1) * max(0, i1 - i0 + 1);
      }
      int maxcell = gmax<G>(ncell, gmask);
      int ncand = 1 + nslot * maxcell + dv->nboxg;
      for (int cand = 0; cand < ncand; cand++) {
        RawContact rc[2];
        int n = 0, b1 = -1, b2 = -1, other = -1;
        float iw = invw;
        if (valid) {
          if (cand == 0) {  // floor plane (normal +z); geom1 = plane
            if (m->has_floor) {
              float margin = fmaxf(m->geom_margin[g], m->floor_margin);
              other = -1; b1 = -1; b2 = body;
              const int nend = (type == MMZ_GEOM_CAPSULE) ? 2 : 1;
              for (int s = 0; s < nend; s++) {
                const float* p = s ? p1 : p0;
                float dist = p[2] - m->floor_z - r;
                if (dist < margin) {
                  rc[n].dist = dist;
                  rc[n].pos[0] = p[0]; rc[n].pos[1] = p[1]; rc[n].pos[2] = p[2] - (r + 0.5f * dist);
                  rc[n].normal[0] = 0.f; rc[n].normal[1] = 0.f; rc[n].normal[2] = 1.f;
                  rc[n].hint[0] = rc[n].hint[1] = rc[n].hint[2] = 0.f;
                  if (type == MMZ_GEOM_CAPSULE) {  // first tangent along the capsule axis
                    const float* gm = w + L.o_gmat + 9 * g;
                    if (fabsf(gm[8]) <= 0.999999f) { rc[n].hint[0] = gm[2]; rc[n].hint[1] = gm[5]; rc[n].hint[2] = gm[8]; }
                  }
                  n++;
                }
              }
            }
          } else if (cand - 1 < nslot * maxcell) {  // maze boxes; geom1 = robot geom, geom2 = box
            int ci = (cand - 1) / nslot, slot = (cand - 1) % nslot;
            if (ci < ncell) {
              int i = i0 + ci / nj, j = j0 + ci % nj;
              int code = m->grid[i * m->grid_w + j];
              bool hit = (code & (slot == 0 ? MMZ_CELL_WALL : MMZ_CELL_PLATFORM)) != 0;
              if (hit) {
                float bc[3] = {j * m->cell_size - m->origin[0], i * m->cell_size - m->origin[1],
                               (slot == 0) ? m->wall_z : m->plat_z};
                float margin = fmaxf(m->geom_margin[g], m->wall_margin);
                other = -2; b1 = body; b2 = -1;
                if (type == MMZ_GEOM_SPHERE) n = sphere_box(p0, r, bc, ident, m->wall_half, margin, rc);
                else n = capsule_box(p0, p1, r, bc, ident, m->wall_half, margin, rc);
              }
            }
          } else {  // box geoms on other moving bodies; geom1 = this geom, geom2 = box
            int gb = dv->boxg[cand - 1 - nslot * maxcell];
            if (moving_pair_ok(g, gb)) {
              float margin = fmaxf(m->geom_margin[g], m->geom_margin[gb]);
              other = gb; b1 = body; b2 = m->geom_body[gb];
              iw = invw + m->geom_invweight[gb];
              const float* bc = w + L.o_gpos + 3 * gb;
              const float* bR = w + L.o_gmat + 9 * gb;
              if (type == MMZ_GEOM_SPHERE) n = sphere_box(p0, r, bc, bR, m->geom_size[gb], margin, rc);
              else n = capsule_box(p0, p1, r, bc, bR, m->geom_size[gb], margin, rc);
            }
          }
        }
        unsigned any = __ballot_sync(gmask, n > 0);
        if (!any) continue;
        int incl = n;  // inclusive scan over the group keeps the contact order deterministic
#pragma unroll
        for (int off = 1; off < G; off <<= 1) {
          int t = __shfl_up_sync(gmask, incl, off, G);
          if (lane >= off) incl += t;
        }
        int total = __shfl_sync(gmask, incl, G - 1, G);
        int base = ncon + incl - n;
        if (n > 0) {
          float par[9];
          mix_params(g, other, par);
          for (int k = 0; k < n; k++)
            if (base + k < L.maxcon) write_contact(L, base + k, rc[k], b1, b2, iw, par);
        }
        ncon += total;
        if (ncon > L.maxcon) { ncon = L.maxcon; overflow = true; }
      }
    }
    if (lane == 0) { cn[N_CON] = ncon; cn[N_OVERFLOW] = overflow ? 1 : 0; }
    sync();
    // ---- box geoms (Point's arrow, movable blocks): few, handled by one lane each in order
    for (int k = 0; k < dv->nboxg; k++) {
      int g = dv->boxg[k];
      if (lane == 0 && ((m->geom_contype[g] | m->geom_conaffinity[g]) & 1)) box_geom_contacts(L, g, k);
      sync();
    }
  }

  __device__ __noinline__ void box_geom_contacts(const Layout& L, int g, int kbox) {
    int* cn = cnt(L);
    int ncon = cn[N_CON];
    const float ident[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    const float* gp = w + L.o_gpos + 3 * g;
    const float* gm = w + L.o_gmat + 9 * g;
    const float* sz = m->geom_size[g];
    int body = m->geom_body[g];
    float invw = m->geom_invweight[g];
    RawContact rc[8];
    float par[9];
    if (m->has_floor) {  // corners below the plane, at most 4; geom1 = plane
      mix_params(g, -1, par);
      int n = 0;
      for (int c = 0; c < 8 && n < 4; c++) {
        float loc[3] = {(c & 1 ? 1.f : -1.f) * sz[0], (c & 2 ? 1.f : -1.f) * sz[1], (c & 4 ? 1.f : -1.f) * sz[2]}, wp[3];
        mat_vec(wp, gm, loc);
        float dist = wp[2] + gp[2] - m->floor_z;
        if (dist < par[0]) {
          rc[n].dist = dist;
          rc[n].pos[0] = wp[0] + gp[0]; rc[n].pos[1] = wp[1] + gp[1]; rc[n].pos[2] = wp[2] + gp[2] - 0.5f * dist;
          rc[n].normal[0] = 0.f; rc[n].normal[1] = 0.f; rc[n].normal[2] = 1.f;
          rc[n].hint[0] = rc[n].hint[1] = rc[n].hint[2] = 0.f;
          n++;
        }
      }
      for (int k = 0; k < n; k++) {
        if (ncon < L.maxcon) write_contact(L, ncon++, rc[k], -1, body, invw, par);
        else cn[N_OVERFLOW] = 1;
      }
    }
    {  // maze boxes; geom1 = wall (lower geom id), geom2 = this box
      float ext[3];
#pragma unroll
      for (int k = 0; k < 3; k++) ext[k] = fabsf(gm[3 * k]) * sz[0] + fabsf(gm[3 * k + 1]) * sz[1] + fabsf(gm[3 * k + 2]) * sz[2];
      mix_params(g, -2, par);
      ext[0] += par[0]; ext[1] += par[0];
      int i0, i1, j0, j1;
      cell_range(gp, ext, &i0, &i1, &j0, &j1);
      for (int i = i0; i <= i1; i++)
        for (int j = j0; j <= j1; j++) {
          int code = m->grid[i * m->grid_w + j];
          for (int slot = 0; slot < 2; slot++) {
            if (!(code & (slot == 0 ? MMZ_CELL_WALL : MMZ_CELL_PLATFORM))) continue;
            float bc[3] = {j * m->cell_size - m->origin[0], i * m->cell_size - m->origin[1],
                           slot == 0 ? m->wall_z : m->plat_z};
            int n = box_box(bc, ident, m->wall_half, gp, gm, sz, par[0], rc);
            for (int k = 0; k < n; k++) {
              if (ncon < L.maxcon) write_contact(L, ncon++, rc[k], -1, body, invw, par);
              else cn[N_OVERFLOW] = 1;
            }
          }
        }
    }
    for (int k2 = kbox + 1; k2 < dv->nboxg; k2++) {  // box against box on different moving bodies
      int g2 = dv->boxg[k2];
      if (!moving_pair_ok(g, g2)) continue;
      mix_params(g, g2, par);
      int n = box_box(gp, gm, sz, w + L.o_gpos + 3 * g2, w + L.o_gmat + 9 * g2, m->geom_size[g2], par[0], rc);
      for (int k = 0; k < n; k++) {
        if (ncon < L.maxcon) write_contact(L, ncon++, rc[k], body, m->geom_body[g2], invw + m->geom_invweight[g2], par);
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
  // (k, b, impedance) -> D and aref pieces for one row
  MMZ_DI void row_params(const float* solref, const float* solimp, float pos, float margin, float diag, float* D,
                         float* kimp_r, float* bb) const {
    float tc = fmaxf(solref[0], 2.f * m->timestep), dr = solref[1];  // refsafe
    float dmax = fminf(fmaxf(solimp[1], 1e-4f), 0.9999f);
    float k = 1.f / fmaxf(kMinVal, dmax * dmax * tc * tc * dr * dr);
    *bb = 2.f / fmaxf(kMinVal, dmax * tc);
    float imp = impedance(solimp, pos - margin);
    float R = fmaxf(kMinVal, (1.f - imp) * diag / imp);
    *D = 1.f / R;
    *kimp_r = k * imp * (pos - margin);
  }

  MMZ_DI void make_constraints(const Layout& L) {
    int* cn = cnt(L);
    const float* qpos = w + L.o_qpos;
    const float* qvel = w + L.o_qvel;
    const float* cdof = w + L.o_cdof;
    // ---- joint limits
    int nlim = 0;
    for (int jbase = 0; jbase < L.nj; jbase += G) {
      int j = jbase + lane;
      int n = 0;
      float dist[2] = {0.f, 0.f};
      if (j < L.nj && m->jnt_limited[j]) {
        float q = qpos[m->jnt_qadr[j]];
        dist[0] = q - m->jnt_range[j][0];
        dist[1] = m->jnt_range[j][1] - q;
        n = (dist[0] < m->jnt_margin[j]) + (dist[1] < m->jnt_margin[j]);
      }
      unsigned any = __ballot_sync(gmask, n > 0);
      if (!any) continue;
      int incl = n;
#pragma unroll
      for (int off = 1; off < G; off <<= 1) {
        int t = __shfl_up_sync(gmask, incl, off, G);
        if (lane >= off) incl += t;
      }
      int total = __shfl_sync(gmask, incl, G - 1, G);
      int slot = nlim + incl - n;
      if (n > 0) {
        int d = m->jnt_dadr[j];
        for (int side = 0; side < 2; side++) {
          if (!(dist[side] < m->jnt_margin[j]) || slot >= L.maxlim) continue;
          float sign = side == 0 ? 1.f : -1.f, D, kr, bb;
          row_params(m->jnt_solref[j], m->jnt_solimp[j], dist[side], m->jnt_margin[j], m->dof_invweight0[d], &D, &kr, &bb);
          float* r = w + L.o_lim + slot * R_STRIDE;
          r[R_DOF] = __int_as_float(d);
          r[R_SIGN] = sign;
          r[R_D] = D;
          r[R_AREF] = -bb * sign * qvel[d] - kr;
          slot++;
        }
      }
      nlim = min(nlim + total, L.maxlim);
    }
    // ---- frictional contacts, pyramidal cone, condim 3: frame Jacobian + 4 edge rows
    int ncon = cn[N_CON];
    for (int c = lane; c < ncon; c += G) {
      float* cs = w + L.o_con + c * L.cstride;
      float* J = cs + C_J;
      float pos[3], fr[9];
#pragma unroll
      for (int k = 0; k < 3; k++) pos[k] = cs[C_POS + k] - w[L.o_xpos + k];
#pragma unroll
      for (int k = 0; k < 9; k++) fr[k] = cs[C_FRAME + k];
      int b1 = __float_as_int(cs[C_BODY1]), b2 = __float_as_int(cs[C_BODY2]);
      int mask1 = b1 >= 0 ? m->body_dofmask[b1] : 0, mask2 = b2 >= 0 ? m->body_dofmask[b2] : 0;
      float jv[3] = {0.f, 0.f, 0.f};
      for (int d = 0; d < L.nv; d++) {
        float s = (float)(mask2 >> d & 1) - (float)(mask1 >> d & 1);
        float jn = 0.f, jt1 = 0.f, jt2 = 0.f;
        if (s != 0.f) {
          const float* cd = cdof + 6 * d;
          float wxp[3];
          cross3(wxp, cd, pos);
          float v[3] = {cd[3] + wxp[0], cd[4] + wxp[1], cd[5] + wxp[2]};
          jn = s * dot3(fr, v); jt1 = s * dot3(fr + 3, v); jt2 = s * dot3(fr + 6, v);
          float qv = qvel[d];
          jv[0] += jn * qv; jv[1] += jt1 * qv; jv[2] += jt2 * qv;
        }
        J[d] = jn; J[L.nv + d] = jt1; J[2 * L.nv + d] = jt2;
      }
      float mu = cs[C_MU], D, kr, bb;
      row_params(cs + C_SOLREF, cs + C_SOLIMP, cs[C_DIST], cs[C_MARGIN], cs[C_INVW] * (1.f + mu * mu), &D, &kr, &bb);
      // all edges of the pyramid share R = 2 mu^2 R_first
      cs[C_DIST] = 1.f / fmaxf(kMinVal, 2.f * mu * mu / D);
      cs[C_AREF + 0] = -bb * (jv[0] + mu * jv[1]) - kr;
      cs[C_AREF + 1] = -bb * (jv[0] - mu * jv[1]) - kr;
      cs[C_AREF + 2] = -bb * (jv[0] + mu * jv[2]) - kr;
      cs[C_AREF + 3] = -bb * (jv[0] - mu * jv[2]) - kr;
    }
    if (lane == 0) cn[N_LIM] = nlim;
    sync();
  }

  // ---------------------------------------------------------------- dense Cholesky on register rows
  // Lane i holds row i of a symmetric positive-definite matrix; on return L (lower) is in
  // shared memory at Lsm[i*ldm + k], k <= i.
  MMZ_DI void chol_rows(const Layout& L, float (&row)[NVP], float* Lsm) {
    float* col = w + L.o_col;  // 2 x NVP, double-buffered
    const int nv = L.nv;
#pragma unroll
    for (int j = 0; j < NVP; j++) {
      if (j < nv) {
        float piv = fmaxf(__shfl_sync(gmask, row[j], j, G), kMinVal);
        float inv = rsqrtf(piv);
        float lij = (lane == j) ? piv * inv : row[j] * inv;
        row[j] = lij;
        float* cb = col + (j & 1) * NVP;
        if (lane < nv) cb[lane] = lij;
        sync();
#pragma unroll
        for (int k = j + 1; k < NVP; k++)
          if (k <= lane && k < nv) row[k] -= lij * cb[k];
      }
    }
    if (lane < nv) {
#pragma unroll
      for (int k = 0; k < NVP; k++)
        if (k <= lane) Lsm[lane * L.ldm + k] = row[k];
    }
    sync();
  }
  // x <- (L L^T)^-1 x for the vector held one element per lane
  MMZ_DI float chol_solve(const Layout& L, const float* Lsm, float x) {
    const int nv = L.nv;
    float invd = (lane < nv) ? 1.f / Lsm[lane * L.ldm + lane] : 0.f;
    float acc = x;
    for (int k = 0; k < nv; k++) {
      float yk = __shfl_sync(gmask, acc * invd, k, G);
      if (lane > k && lane < nv) acc -= Lsm[lane * L.ldm + k] * yk;
      if (lane == k) acc = yk;
    }
    for (int k = nv - 1; k >= 0; k--) {
      float xk = __shfl_sync(gmask, acc * invd, k, G);
      if (lane < k) acc -= Lsm[k * L.ldm + lane] * xk;
      if (lane == k) acc = xk;
    }
    return acc;
  }

  // ---------------------------------------------------------------- Newton solver (mj_solNewton)
  //   min_a 1/2 (a - a0)^T M (a - a0) + sum_i 1/2 D_i min(0, J_i a - aref_i)^2
  MMZ_DI void contact_dots(const Layout& L, const float* x, int ncon, int dst) {
    for (int c = lane; c < ncon; c += G) {
      float* cs = w + L.o_con + c * L.cstride;
      const float* J = cs + C_J;
      float s0 = 0.f, s1 = 0.f, s2 = 0.f;
      for (int d = 0; d < L.nv; d++) {
        float xd = x[d];
        s0 += J[d] * xd; s1 += J[L.nv + d] * xd; s2 += J[2 * L.nv + d] * xd;
      }
      float mu = cs[C_MU];
      if (dst == C_JAR) {
        cs[C_JAR + 0] = s0 + mu * s1 - cs[C_AREF + 0];
        cs[C_JAR + 1] = s0 - mu * s1 - cs[C_AREF + 1];
        cs[C_JAR + 2] = s0 + mu * s2 - cs[C_AREF + 2];
        cs[C_JAR + 3] = s0 - mu * s2 - cs[C_AREF + 3];
      } else {
        cs[C_JV + 0] = s0 + mu * s1;
        cs[C_JV + 1] = s0 - mu * s1;
        cs[C_JV + 2] = s0 + mu * s2;
        cs[C_JV + 3] = s0 - mu * s2;
      }
    }
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
    if (!warmstart || nrow == 0) {
      for (int d = lane; d < nv; d += G) a[d] = w[L.o_qacc_smooth + d];
    }
    if (lane == 0) cn[N_ITER] = 0;
    sync();
    if (nrow == 0) return;
    const bool me = lane < nv;
    for (int it = 0; it < 30; it++) {
      // rows: jar = J a - aref
      for (int r = lane; r < nlim; r += G) {
        float* lr = lim + r * R_STRIDE;
        lr[R_JAR] = lr[R_SIGN] * a[__float_as_int(lr[R_DOF])] - lr[R_AREF];
      }
      contact_dots(L, a, ncon, C_JAR);
      sync();
      // gradient and Hessian row of this lane's dof
      float hrow[NVP];
      float Ma = 0.f, grad, mag;
#pragma unroll
      for (int k = 0; k < NVP; k++) {
        float mk = (me && k < nv) ? M[lane * L.ldm + k] : ((k == lane) ? 1.f : 0.f);
        hrow[k] = mk;
        if (k < nv) Ma += mk * a[k];
      }
      if (!me) Ma = 0.f;
      float sm = me ? smooth[lane] : 0.f;
      grad = Ma - sm;
      mag = fabsf(Ma) + fabsf(sm);
      float dadd = 0.f;
      for (int r = 0; r < nlim; r++) {
        const float* lr = lim + r * R_STRIDE;
        float jar = lr[R_JAR];
        if (jar < 0.f && __float_as_int(lr[R_DOF]) == lane) {
          float f = lr[R_D] * jar * lr[R_SIGN];  // = -J^T force
          grad += f;
          mag += fabsf(f);
          dadd += lr[R_D];
        }
      }
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
        float gn = f0 + f1 + f2 + f3, g1 = mu * (f0 - f1), g2 = mu * (f2 - f3);
        grad += jn * gn + jt1 * g1 + jt2 * g2;
        mag += fabsf(jn * gn) + fabsf(jt1 * g1) + fabsf(jt2 * g2);
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
      float gn2 = gsum<G>(grad * grad, gmask), ref2 = gsum<G>(mag * mag, gmask);
      if (gn2 <= 4e-12f * ref2 + 1e-30f) break;
      chol_rows(L, hrow, H);
      float dr = chol_solve(L, H, -grad);
      if (me) dir[lane] = dr;
      sync();
      // exact line search along dir: root of the monotone piecewise-linear derivative
      float md = 0.f;
      if (me)
        for (int k = 0; k < nv; k++) md += M[lane * L.ldm + k] * dir[k];
      float g0 = gsum<G>(me ? dr * (Ma - sm) : 0.f, gmask), h0 = gsum<G>(me ? dr * md : 0.f, gmask);
      for (int r = lane; r < nlim; r += G) {
        float* lr = lim + r * R_STRIDE;
        lr[R_JV] = lr[R_SIGN] * dir[__float_as_int(lr[R_DOF])];
      }
      contact_dots(L, dir, ncon, C_JV);
      sync();
      float lo = 0.f, hi = -1.f, alpha = 1.f;
      for (int ls = 0; ls < 24; ls++) {
        float g = 0.f, h = 0.f;
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
      float an = 0.f, sn = 0.f;
      if (me) {
        float av = a[lane], st = alpha * dr;
        a[lane] = av + st;
        an = av * av; sn = st * st;
      }
      if (lane == 0) cn[N_ITER] = it + 1;
      sync();
      // fp32 floor: stop when the step no longer changes the iterate
      an = gsum<G>(an, gmask); sn = gsum<G>(sn, gmask);
      if (sn <= 1e-11f * (an + 1e-6f)) break;
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
    passive_forces(L);
    const int nv = L.nv;
    const bool me = lane < nv;
    float fs = 0.f;
    if (me) {
      float act = 0.f;
      for (int k = 0; k < L.nu; k++)
        if (m->act_dof[k] == lane) {
          float c = w[L.o_ctrl + k];
          if (m->act_limited[k]) c = fminf(fmaxf(c, m->act_ctrlrange[k][0]), m->act_ctrlrange[k][1]);
          act += m->act_gear[k] * c;
        }
      fs = w[L.o_passive + lane] - w[L.o_bias + lane] + act;
      w[L.o_smooth + lane] = fs;
    }
    float hrow[NVP];
#pragma unroll
    for (int k = 0; k < NVP; k++) hrow[k] = (me && k < nv) ? w[L.o_M + lane * L.ldm + k] : ((k == lane) ? 1.f : 0.f);
    chol_rows(L, hrow, w + L.o_H);
    float qs = chol_solve(L, w + L.o_H, fs);
    if (me) w[L.o_qacc_smooth + lane] = qs;
    sync();
    make_constraints(L);
    solve(L, warmstart);
  }

  // position update on the configuration manifold (mj_integratePos)
  MMZ_DI void integrate_pos(const Layout& L, const float* q0, const float* vel, float scale, float h) {
    float* qpos = w + L.o_qpos;
    for (int j = lane; j < L.nj; j += G) {
      int qa = m->jnt_qadr[j], d = m->jnt_dadr[j];
      if (m->jnt_type[j] == MMZ_JNT_FREE) {
#pragma unroll
        for (int k = 0; k < 3; k++) qpos[qa + k] = q0[qa + k] + h * scale * vel[d + k];
        float wv[3] = {scale * vel[d + 3], scale * vel[d + 4], scale * vel[d + 5]};
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
        qpos[qa] = q0[qa] + h * scale * vel[d];
      }
    }
  }

  MMZ_DI bool state_bad(const Layout& L) const {
    bool bad = false;
    for (int i = lane; i < L.nq; i += G) bad |= !(fabsf(w[L.o_qpos + i]) < kMaxVal);
    for (int i = lane; i < L.nv; i += G) bad |= !(fabsf(w[L.o_qvel + i]) < kMaxVal);
    return __ballot_sync(gmask, bad) != 0;
  }

  // ---------------------------------------------------------------- mj_step, RK4 (mj_RungeKutta)
  // returns true if the state blew up (MuJoCo would auto-reset)
  MMZ_DI bool mj_step(const Layout& L) {
    const float h = m->timestep;
    const int nq = L.nq, nv = L.nv;
    float *qpos = w + L.o_qpos, *qvel = w + L.o_qvel, *qacc = w + L.o_qacc;
    float *q0 = w + L.o_q0, *v0 = w + L.o_v0, *xv = w + L.o_xv, *fa = w + L.o_fa, *accv = w + L.o_accv, *acca = w + L.o_acca;
    if (state_bad(L)) return true;
    forward(L, true);
    bool badacc = false;
    for (int d = lane; d < nv; d += G) badacc |= !(fabsf(qacc[d]) < kMaxVal);
    if (__ballot_sync(gmask, badacc)) return true;
    for (int i = lane; i < nq; i += G) q0[i] = qpos[i];
    for (int d = lane; d < nv; d += G) {
      float v = qvel[d], f = qacc[d];
      v0[d] = v; xv[d] = v; fa[d] = f;
      accv[d] = v * (1.f / 6.f); acca[d] = f * (1.f / 6.f);
    }
    sync();
#pragma unroll 1
    for (int i = 1; i < 4; i++) {
      const float A = (i == 3) ? 1.f : 0.5f, B = (i == 3) ? (1.f / 6.f) : (1.f / 3.f);
      integrate_pos(L, q0, xv, A, h);
      sync();
      for (int d = lane; d < nv; d += G) {
        float v = v0[d] + h * A * fa[d];
        qvel[d] = v; xv[d] = v;
        accv[d] += B * v;
      }
      sync();
      forward(L, true);
      for (int d = lane; d < nv; d += G) {
        float f = qacc[d];
        fa[d] = f;
        acca[d] += B * f;
      }
      sync();
    }
    integrate_pos(L, q0, accv, 1.f, h);
    for (int d = lane; d < nv; d += G) qvel[d] = v0[d] + h * acca[d];
    sync();
    // derived arrays (xpos, contacts) deliberately stay at the 4th-stage state: SURVEY quirk Q15
    return state_bad(L);
  }
};

}  // namespace mmz
