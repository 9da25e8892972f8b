// mmz_hstep.cuh - MazeEnv.step / reset / observe around the hybrid dynamics (mmz_hkernel.cuh).
//
// One launch of maze_hkernel<NVP, BOX, TMODE_STEP> is one MazeEnv.step (reference maze_env.py:448-481) of N lock-step
// environments for torque-driven agents (AntEnv.step ant.py:61-73, SwimmerEnv.step swimmer.py:37-47):
// frame_skip x mj_step, _get_obs (maze_env.py:351-369), MazeTask.reward / termination (maze_task.py),
// TimeLimit truncation (__init__.py:31) and the optional in-kernel auto-reset (reset_model, ant.py:84-96).
#pragma once
#include <cstdio>
#include "mmz_clamp.cuh"
#include "mmz_hkernel.cuh"

namespace mmz {

MMZ_DI unsigned h_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// persisted state row -> workspace slot
MMZ_DI int h_row_slot(const TLayout& L, int r) {
  if (r < L.nq) return L.o_qpos + r;
  r -= L.nq;
  if (r < L.nv) return L.o_qvel + r;
  r -= L.nv;
  if (r < L.nv) return L.o_qacc + r;
  return L.o_objpos + (r - L.nv);
}

// column of obs[N][obs_dim] for entry i of the assembled observation: the top-down view (written by the view kernel)
// sits between the state part and the trailing t * 0.001 (maze_env.py:368-369)
MMZ_DI int obs_column(const TLayout& L, int i) { return i == L.obs_core - 1 ? L.obs_dim - 1 : i; }

template <int NVP, int BOX>
struct HTask : HEnv<NVP, BOX> {
  using HEnv<NVP, BOX>::m; using HEnv<NVP, BOX>::sm; using HEnv<NVP, BOX>::e; using HEnv<NVP, BOX>::wid;
#define S(i) sm[(i) * HS + e]
  MMZ_DI int first_goal(int where) const {
    for (int g = 0; g < m->ngoal; g++) {
      float s = 0.f;
      for (int i = 0; i < m->goal_dim[g]; i++) { const float d = S(where + i) - m->goal_pos[g][i]; s += d * d; }
      if (sqrtf(s) <= m->goal_thr[g]) return g;
    }
    return -1;
  }
  MMZ_DI float goal_dist(int where) const {
    float s = 0.f;
    for (int i = 0; i < m->goal_dim[0]; i++) { const float d = S(where + i) - m->goal_pos[0][i]; s += d * d; }
    return sqrtf(s);
  }
  // MazeTask.reward / termination on the assembled observation (slots o_obs ...)
  MMZ_DI void task_rules(const TLayout& L, float* reward, bool* term) const {
    bool t = false;
    if (m->term_rule == MMZ_TERM_AGENT) t = first_goal(L.o_obs) >= 0;
    else if (m->term_rule == MMZ_TERM_OBJECT) t = first_goal(L.o_obs + 3) >= 0;
    float r = 0.f;
    int g;
    switch (m->reward_rule) {
      case MMZ_REWARD_REACH: r = t ? 1.f : m->penalty; break;
      case MMZ_REWARD_SCALED: g = first_goal(L.o_obs); r = g >= 0 ? m->goal_scale[g] : m->penalty; break;
      case MMZ_REWARD_SCALED_OBJECT: g = first_goal(L.o_obs + 3); r = g >= 0 ? m->goal_scale[g] : m->penalty; break;
      case MMZ_REWARD_DIST_OBJECT: r = -goal_dist(L.o_obs + 3) / m->task_scale; break;
      case MMZ_REWARD_DIST: r = -goal_dist(L.o_obs) / m->task_scale; break;
      default: r = 0.f;
    }
    *reward = r;
    *term = t;
  }
  // observed bodies: the reference reads data.xpos, which is only as fresh as the last kinematics pass
  MMZ_DI void latch_objpos(const TLayout& L, bool on) {
    if (on)
      for (int i = wid; i < 3 * L.nlatch; i += TW) S(L.o_objpos + i) = S(L.o_xpos + 3 * m->obj_body[i / 3] + i % 3);
  }
  // MazeEnv._get_obs (maze_env.py:351-369) into the o_obs slots
  MMZ_DI void assemble_obs(const TLayout& L, int t, bool on) {
    const int naq = m->n_agent_q, nav = m->n_agent_v, no = 3 * L.nobj;
    if (on) {
      for (int i = wid; i < L.obs_core; i += TW) {
        float v;
        if (i < 3 && i < naq) v = S(L.o_qpos + i);
        else if (i < 3 + no) v = S(L.o_objpos + i - 3);
        else if (i < naq + no) v = S(L.o_qpos + i - no);
        else if (i < naq + no + nav) v = S(L.o_qvel + i - naq - no);
        else v = t * 0.001f;
        S(L.o_obs + i) = v;
      }
    }
  }
  // reset_model (ant.py:84-96, swimmer.py:55-68, point.py:71-81): same distributions, Philox stream
  MMZ_DI void reset_state(const TLayout& L, unsigned long long seed, int genv, int nreset, bool noise, bool on) {
    const float amp = m->reset_noise;
    if (!on) return;
#pragma unroll 1
    for (int i = wid; i < L.nq + L.nv; i += TW) {
      const bool isq = i < L.nq;
      const int k = isq ? i : i - L.nq;
      float val = isq ? m->qpos0[k] : 0.f;
      if (noise && k < (isq ? m->n_agent_q : m->n_agent_v)) {
        uint32_t c[4] = {(uint32_t)genv, (uint32_t)nreset, (uint32_t)(isq ? k : 64 + k), 0u};
        philox4x32(c, (uint32_t)seed, (uint32_t)(seed >> 32));
        const float u = u01(c[0]);
        if (isq || m->reset_kind == MMZ_RESET_SWIMMER) val += amp * (2.f * u - 1.f);
        else if (m->reset_kind == MMZ_RESET_POINT) val += amp * u;
        else val += amp * sqrtf(-2.f * logf(1.f - u)) * cospif(2.f * u01(c[1]));  // Box-Muller
      }
      if (isq) S(L.o_qpos + k) = val;
      else { S(L.o_qvel + k) = val; S(L.o_qacc + k) = 0.f; }
    }
  }
#undef S
};

template <int NVP, int BOX, int MODE>
__global__ void __launch_bounds__(TW * 32, 1) maze_hkernel(const __grid_constant__ TArgs A) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar;
  const TLayout& L = A.L;
  const int tid = threadIdx.x;
  const int env0 = (A.block0 + blockIdx.x) * TE;
  float* ws = reinterpret_cast<float*>(smem + ((L.model_bytes + 127) & ~127));

  // ---- model constants: one bulk async copy global -> shared, completion on an mbarrier
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(h_smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(h_smem_u32(&bar)), "r"(L.model_bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(h_smem_u32(smem)), "l"(A.model), "r"(L.model_bytes), "r"(h_smem_u32(&bar)) : "memory");
  }
#ifdef MMZ_PHASE_TIMING
  if (MODE == TMODE_STEP && tid == 0 && blockIdx.x == 0) {  // totals of the launches before this one
    printf("phase cycles (block 0): rootkin %llu walk %llu B %llu C %llu D %llu E %llu solver %llu | iterations per solve:",
           g_phase[0], g_phase[1], g_phase[2], g_phase[3], g_phase[4], g_phase[5], g_phase[6]);
    for (int i = 0; i < 16; i++) printf(" %u", g_iter_hist[i]);
    printf("\n  per warp solve / wait (kcycles):");
    for (int i = 0; i < 16; i++) printf(" %llu/%llu", g_wsolve[i] / 1000, g_wwait[i] / 1000);
    printf("\n  solver sections (sum over the warps of block 0, kcycles): prologue %llu contact pass %llu gradient %llu hessian %llu elimination %llu rows %llu line search %llu update %llu | warp iterations %llu\n",
           g_sec[0] / 1000, g_sec[1] / 1000, g_sec[2] / 1000, g_sec[3] / 1000, g_sec[4] / 1000, g_sec[5] / 1000, g_sec[6] / 1000, g_sec[7] / 1000, g_sec[15]);
    printf("  per warp C / D own work (kcycles):");
    for (int i = 0; i < 16; i++) printf(" %llu/%llu", g_wC[i] / 1000, g_wD[i] / 1000);
    printf("\n");
    printf("  step (block 0, kcycles): prologue %llu mj_step %llu clamp / inner reward %llu reset + obs + rules %llu outputs %llu\n",
           g_phase[8] / 1000, g_phase[9] / 1000, g_phase[10] / 1000, g_phase[11] / 1000, g_phase[12] / 1000);
  }
#endif
#ifdef MMZ_PHASE_TIMING
  long long ktick_ = clock64();
#define MMZ_KTICK(i) do { if (tid == 0 && blockIdx.x == 0) { const long long t_ = clock64(); g_phase[i] += t_ - ktick_; ktick_ = t_; } } while (0)
#else
#define MMZ_KTICK(i) do { } while (0)
#endif
  HTask<NVP, BOX> T;
  T.m = reinterpret_cast<const mmz_model*>(smem);
  T.dv = reinterpret_cast<const TDerived*>(smem + ((sizeof(mmz_model) + 15) & ~15));
  T.sm = ws;
  T.e = tid % 32;
  T.wid = tid / 32;
  T.genv = T.wid + 16 * (T.e >> 4);
  {
    // solver v2: natural-order float4 areas over the workspace slots that are dead while the solver runs. The Jacobian
    // area of the upper 16 environments is skewed by one float4, so that the two environments of a warp read their
    // (broadcast) columns from different banks.
    float4* nat = reinterpret_cast<float4*>(ws + L.o_nat * HS);
    T.jg = nat + T.genv * L.jes + (T.genv >> 4);
    T.fg = nat + L.fa_off + T.genv * (2 * L.maxcon);
    T.hg = nullptr;
    if (BOX && L.v3) {  // solver v3: one block per environment - Jacobian pool, (force, weights) pairs, Hessian rows of the 16 lanes
      float4* base = nat + T.genv * L.es + (T.genv >> 4);
      T.jg = base;
      T.fg = base + L.njac;
      T.hg = reinterpret_cast<float*>(base + L.njac + 2 * L.maxcon) + (T.e & 15) * (NVP + 1);
    }
  }
  T.lane = T.e & 15;
  T.gshift = (T.e >> 4) * 16;
  T.tol = A.tol;
  const int e = T.e, wid = T.wid;
#define S(i) ws[(i) * HS + e]
  for (int i = L.nv + wid; i < NVP; i += TW) { S(L.o_qacc + i) = 0.f; S(L.o_dir + i) = 0.f; }  // padding the solver reads
  // ---- state tile: every row is 32 consecutive environments = one 128-byte line per warp load
  if (MODE != TMODE_RESET || A.mask != nullptr) {
    for (int r = wid; r < L.nstate; r += TW) S(h_row_slot(L, r)) = A.state[(size_t)r * A.npad + env0 + e];
  }
  __syncthreads();  // barrier initialised + state tile visible
  {
    unsigned ok = 0;
    while (!ok) {
      asm volatile(
          "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}\n"
          : "=r"(ok) : "r"(h_smem_u32(&bar)) : "memory");
    }
  }
  const int env = env0 + e;
  const bool real = env < A.n;  // [n, npad) are padding environments: they run, and write no outputs
  int t = A.counters[env], nreset = A.counters[A.npad + env];
  int* cn = reinterpret_cast<int*>(ws);

  if (MODE == TMODE_STEP) {
    const bool auto_reset = (A.flags & T_FLAG_AUTO_RESET) != 0;
    const bool teleport = T.m->step_kind == MMZ_STEP_TELEPORT;
    for (int a = wid; a < L.nu; a += TW) {
      const float v = real ? A.action[(size_t)env * L.nu + a] : 0.f;
      S(L.o_ctrl + a) = teleport ? 0.f : v;  // PointEnv.step never drives its motors (point.py:44-61, quirk Q16)
      S(L.o_act + a) = v;
    }
    if (wid == 0)
      for (int k = 0; k < 4; k++) cn[(L.o_cnt + TN_ITER_SUM + k) * HS + e] = 0;
    const float bx = S(L.o_qpos), by = S(L.o_qpos + 1);
    t += 1;
    if (teleport) {  // turn, move along the new heading, clip EVERY velocity (blocks and balls too)
      __syncthreads();
      if (wid == 0) {
        float ori = S(L.o_qpos + 2) + S(L.o_act + 1);
        if (ori < -kPi) ori += 2.f * kPi;
        else if (kPi < ori) ori -= 2.f * kPi;
        float sn, cs;
        sincosf(ori, &sn, &cs);
        const float a0 = S(L.o_act);
        S(L.o_qpos + 2) = ori;
        S(L.o_qpos) = bx + cs * a0;
        S(L.o_qpos + 1) = by + sn * a0;
      }
      for (int d = wid; d < L.nv; d += TW) S(L.o_qvel + d) = fminf(fmaxf(S(L.o_qvel + d), -T.m->vel_limit), T.m->vel_limit);
      __syncthreads();
    }
    bool bad = T.state_bad(L);  // mj_checkPos / mj_checkVel of the incoming state
    __syncthreads();
    MMZ_KTICK(8);  // prologue: model blob, state tile, actions, teleport
#pragma unroll 1
    for (int k = 0; k < T.m->frame_skip; k++) bad = T.mj_step(L, bad);
    MMZ_KTICK(9);  // mj_step x frame_skip
    float inner = 0.f, fwd = 0.f, cc = 0.f;
    bool moved = false;
    if (teleport) {
      if (T.m->manual_collision) {  // maze_env.py:450-464: a move through a wall segment bounces, or is undone
        if (wid == 0) {
          const float old[2] = {bx, by}, nw[2] = {S(L.o_qpos), S(L.o_qpos + 1)};
          float pos[2];
          const bool hit = clamp_move(T.m, old, nw, pos) && !bad;
          if (hit) { S(L.o_qpos) = pos[0]; S(L.o_qpos + 1) = pos[1]; }
          cn[(L.o_cnt + TN_MOVED) * HS + e] = hit ? 1 : 0;
        }
        __syncthreads();
        moved = cn[(L.o_cnt + TN_MOVED) * HS + e] != 0;  // set_xy -> set_state -> mj_forward refreshes xpos
      }
    } else if (!bad) {  // AntEnv.step / SwimmerEnv.step (ant.py:61-73, swimmer.py:37-47)
      const float dt = T.m->timestep * T.m->frame_skip;
      const float vx = (S(L.o_qpos) - bx) / dt, vy = (S(L.o_qpos + 1) - by) / dt;
      // forward_reward_fn (ant.py:18-23): vnorm, vabs, or left to the host wrapper
      const int fk = T.m->forward_reward_kind;
      fwd = fk == MMZ_FWD_VABS ? fabsf(vx) + fabsf(vy) : fk == MMZ_FWD_HOST ? 0.f : sqrtf(vx * vx + vy * vy);
      for (int a = 0; a < L.nu; a++) { const float v = S(L.o_act + a); cc += v * v; }
      cc *= T.m->ctrl_cost_weight;
      inner = T.m->forward_reward_weight * fwd - cc;
    }
    MMZ_KTICK(10);  // wall clamp / inner reward
    unsigned bits = bad ? T_UNSTABLE_BIT : 0;  // MuJoCo's mj_checkPos/Vel/Acc auto-reset: back to qpos0, zero velocity
    bool reset_now = bad, noise = false, refresh = bad || moved, live = true;
    float reward = 0.f, info0 = 0.f, info1 = 0.f;
#pragma unroll 1
    for (int pass = 0; pass < 2; pass++) {
      T.reset_state(L, A.seed, A.env_offset + env, nreset, noise, live && reset_now);
      T.latch_objpos(L, live && !refresh);  // stale derived arrays are the reference's behaviour (quirk Q15)
      __syncthreads();
      if (__any_sync(kAll, live && refresh)) {
        T.kinematics_only(L);
        T.latch_objpos(L, live && refresh);
        __syncthreads();
      }
      T.assemble_obs(L, t, live);
      __syncthreads();
      if (pass == 0) {
        float outer;
        bool term;
        T.task_rules(L, &outer, &term);
        reward = T.m->inner_reward_scale * inner + outer;
        if (term) bits |= T_DONE_BIT;
        // gym's TimeLimit: info['TimeLimit.truncated'] = not done, i.e. only when the task itself did not end the episode
        if (T.m->max_episode_steps > 0 && t >= T.m->max_episode_steps) bits |= T_DONE_BIT | (term ? 0 : T_TRUNC_BIT);
        info0 = S(L.o_qpos); info1 = S(L.o_qpos + 1);
        live = auto_reset && (bits & T_DONE_BIT);
        if (live) {  // the env that just ended starts its next episode inside this launch
          nreset += 1;
          t = 0;
          reset_now = true; noise = true; refresh = true;
        }
        if (!__any_sync(kAll, live)) break;
        __syncthreads();
      }
    }
    __syncthreads();
    MMZ_KTICK(11);  // reset / refresh / observation / task rules
    // ---- outputs: the block's observations are one contiguous chunk of obs[N][obs_dim]
    {
      const int nreal = min(TE, A.n - env0);
      for (int idx = tid; idx < nreal * L.obs_core; idx += TW * 32) {
        const int ee = idx / L.obs_core, i = idx - ee * L.obs_core;
        const float v = ws[(L.o_obs + i) * HS + ee];
        A.obs[(size_t)(env0 + ee) * L.obs_dim + obs_column(L, i)] = v;
        // fused observation gather: the same contiguous chunk again, into every rank's gathered tensor over NVLink
        if (A.peers.n) peer_store(A.peers, (size_t)(A.peers.row0 + env0 + ee) * L.obs_dim + obs_column(L, i), v);
      }
    }
    if (wid == 0) {
      if (real) {
        A.reward[env] = reward;
        A.done[env] = (uint8_t)bits;
        if (A.info) {
          float* o = A.info + (size_t)env * 4;
          if ((reinterpret_cast<uintptr_t>(A.info) & 15) == 0) *reinterpret_cast<float4*>(o) = make_float4(info0, info1, fwd, -cc);
          else { o[0] = info0; o[1] = info1; o[2] = fwd; o[3] = -cc; }
        }
        if (A.diag)
          for (int k = 0; k < 4; k++) A.diag[(size_t)env * 4 + k] = cn[(L.o_cnt + TN_ITER_SUM + k) * HS + e];
      }
      A.counters[env] = t;
      A.counters[A.npad + env] = nreset;
    }
    MMZ_KTICK(12);  // outputs
#ifdef MMZ_PHASE_TIMING
    if (tid == 0) {
      if (g_launch == 40)
        printf("blk %d: rootkin %u walk %u B %u C %u D %u E %u solver %u | sum of max Newton iterations %u\n", (int)blockIdx.x,
               T.bphase[0], T.bphase[1], T.bphase[2], T.bphase[3], T.bphase[4], T.bphase[5], T.bphase[6], T.bphase[7]);
      if (blockIdx.x == 0) atomicAdd(&g_launch, 1u);
    }
#endif
  } else if (MODE == TMODE_FORWARD) {
    for (int a = wid; a < L.nu; a += TW)  // (the Point's motors are never driven: point.py:44-61)
      S(L.o_ctrl + a) = (real && T.m->step_kind != MMZ_STEP_TELEPORT) ? A.action[(size_t)env * L.nu + a] : 0.f;
    __syncthreads();
    T.forward(L, false);
    if (real) {
      for (int d = wid; d < L.nv; d += TW) A.qacc_out[(size_t)env * L.nv + d] = S(L.o_qacc + d);
      if (A.diag && wid == 0) {
        const int nl = cn[(L.o_cnt + TN_LIM) * HS + e];
        A.diag[env * 4 + 0] = cn[(L.o_cnt + TN_CON) * HS + e];
        A.diag[env * 4 + 1] = nl + 4 * cn[(L.o_cnt + TN_CON) * HS + e];
        A.diag[env * 4 + 2] = cn[(L.o_cnt + TN_ITER) * HS + e];
        A.diag[env * 4 + 3] = cn[(L.o_cnt + TN_OVERFLOW) * HS + e];
      }
    }
  } else if (MODE == TMODE_OBSERVE) {
    T.assemble_obs(L, t, true);
    __syncthreads();
    const int nreal = min(TE, A.n - env0);
    for (int idx = tid; idx < nreal * L.obs_core; idx += TW * 32) {
      const int ee = idx / L.obs_core, i = idx - ee * L.obs_core;
      A.obs[(size_t)(env0 + ee) * L.obs_dim + obs_column(L, i)] = ws[(L.o_obs + i) * HS + ee];
    }
  } else if (MODE == TMODE_RESET) {
    const bool on = A.mask == nullptr || (real && A.mask[env]);
    if (on) { nreset += 1; t = 0; }
    T.reset_state(L, A.seed, A.env_offset + env, nreset, true, on);
    __syncthreads();
    // environments that are not reset keep their (possibly stale) observed-body positions
    T.kinematics_only(L);
    T.latch_objpos(L, on);
    __syncthreads();
    T.assemble_obs(L, 0, on);
    __syncthreads();
    if (on && wid == 0) { A.counters[env] = 0; A.counters[A.npad + env] = nreset; }
    if (A.obs && real && on)
      for (int i = wid; i < L.obs_core; i += TW) A.obs[(size_t)env * L.obs_dim + obs_column(L, i)] = S(L.o_obs + i);
  } else if (MODE == TMODE_REFRESH) {  // after set_state: mj_forward refreshes the derived arrays
    for (int d = wid; d < L.nv; d += TW) S(L.o_qacc + d) = 0.f;
    __syncthreads();
    T.kinematics_only(L);
    T.latch_objpos(L, true);
  }
  __syncthreads();
  if (MODE == TMODE_STEP || MODE == TMODE_RESET || MODE == TMODE_REFRESH) {
    for (int r = wid; r < L.nstate; r += TW) A.state[(size_t)r * A.npad + env0 + e] = S(h_row_slot(L, r));
  }
#undef S
}

typedef void (*hkernel_fn)(const TArgs);
hkernel_fn get_hkernel(int nvp, int mode);

}  // namespace mmz
