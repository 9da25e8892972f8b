// mmz_kernels.cuh - the fused MazeEnv.step kernel and its siblings (reset / observe / forward).
//
// One launch of maze_kernel<G, NVP, MODE_STEP> is one MazeEnv.step (reference
// maze_env.py:448-481) for N lock-step environments:
//   AgentModel.step  (point.py:44-61 teleport + 1 x mj_step | ant.py:61-73, swimmer.py:37-47
//                     frame_skip x mj_step with RK4)          -> Env::mj_step (mmz_dyn.cuh)
//   CollisionDetector.detect + bounce (maze_env.py:450-464, maze_env_utils.py:186-206)
//   MazeEnv._get_obs (maze_env.py:351-369), MazeTask.reward / termination (maze_task.py:77-81
//   and variants), TimeLimit truncation (__init__.py:31), optional in-kernel auto-reset.
//
// Block prologue: the model constants (mmz_model + Derived, ~8 KB) are staged into shared
// memory by ONE 1-D bulk asynchronous copy (TMA, cp.async.bulk ... mbarrier::complete_tx) issued
// by thread 0, while all threads load the block's tile of the structure-of-arrays state
// (rows x EPB consecutive environments: each row segment is contiguous in HBM).
#pragma once
#include "mmz_clamp.cuh"
#include "mmz_dyn.cuh"

namespace mmz {

enum { MODE_STEP = 0, MODE_FORWARD = 1, MODE_OBSERVE = 2, MODE_RESET = 3, MODE_REFRESH = 4 };
enum { DONE_BIT = 1, TRUNC_BIT = 2, UNSTABLE_BIT = 4 };
enum { FLAG_AUTO_RESET = 1 };

struct KArgs {
  Layout L;
  const void* model;   // device: mmz_model (float) + Derived, L.model_bytes
  float* state;        // [L.nstate][npad]
  int* counters;       // [2][npad]: t, number of resets
  int n, npad;
  const float* action; // [n][nu]
  float* obs;          // [n][obs_dim]
  float* reward;       // [n]
  uint8_t* done;       // [n]
  float* info;         // [n][4] or null
  float* qacc_out;     // MODE_FORWARD: [n][nv]
  int* diag;           // MODE_FORWARD: [n][4] {contacts, rows, iterations, overflow}; MODE_STEP (optional):
                       // [n][4] {Newton iterations, line-search iterations, max contacts, capped solves} of the step
  const uint8_t* mask; // MODE_RESET: [n] or null
  unsigned long long seed;
  unsigned flags;
  int env_offset;       // global index of env 0 (keeps the Philox streams independent of the sharding)
  int block0;           // first block of this launch: mmz_step_host runs the batch as a few pipelined block ranges
  ObsPeers peers;       // MODE_STEP: fused observation gather (mmz_layout.h)
};

MMZ_DI unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// state row -> workspace offset
MMZ_DI int row_offset(const Layout& L, int r) {
  if (r < L.nq) return L.o_qpos + r;
  r -= L.nq;
  if (r < L.nv) return L.o_qvel + r;
  r -= L.nv;
  if (r < L.nv) return L.o_qacc + r;
  return L.o_objpos + (r - L.nv);
}

template <int G, int NVP, int FEAT>
struct Task : Env<G, NVP, FEAT> {
  using E = Env<G, NVP, FEAT>;
  using E::dv; using E::lane; using E::m; using E::sync; using E::w;

  MMZ_DI int first_goal(const float* where) const {
    for (int g = 0; g < m->ngoal; g++) {
      float s = 0.f;
      for (int i = 0; i < m->goal_dim[g]; i++) { float d = where[i] - m->goal_pos[g][i]; s += d * d; }
      if (sqrtf(s) <= m->goal_thr[g]) return g;
    }
    return -1;
  }
  MMZ_DI float goal_dist(const float* where) const {
    float s = 0.f;
    for (int i = 0; i < m->goal_dim[0]; i++) { float d = where[i] - m->goal_pos[0][i]; s += d * d; }
    return sqrtf(s);
  }
  // MazeTask.reward / termination on the assembled observation
  MMZ_DI void task_rules(const float* obs, float* reward, bool* term) const {
    bool t = false;
    if (m->term_rule == MMZ_TERM_AGENT) t = first_goal(obs) >= 0;
    else if (m->term_rule == MMZ_TERM_OBJECT) t = first_goal(obs + 3) >= 0;
    float r = 0.f;
    int g;
    switch (m->reward_rule) {
      case MMZ_REWARD_REACH: r = t ? 1.f : m->penalty; break;
      case MMZ_REWARD_SCALED: g = first_goal(obs); r = g >= 0 ? m->goal_scale[g] : m->penalty; break;
      case MMZ_REWARD_SCALED_OBJECT: g = first_goal(obs + 3); r = g >= 0 ? m->goal_scale[g] : m->penalty; break;
      case MMZ_REWARD_DIST_OBJECT: r = -goal_dist(obs + 3) / m->task_scale; break;
      case MMZ_REWARD_DIST: r = -goal_dist(obs) / m->task_scale; break;
      default: r = 0.f;
    }
    *reward = r;
    *term = t;
  }

  // observed bodies: the reference reads data.xpos, which is only as fresh as the last kinematics pass.
  // `on` selects the environments that latch (warp-uniform call, predicated stores).
  MMZ_DI void latch_objpos(const Layout& L, bool on) {
    if (on)
      for (int i = lane; i < 3 * L.nlatch; i += G) w[L.o_objpos + i] = w[L.o_xpos + 3 * m->obj_body[i / 3] + i % 3];
    sync();
  }
  // MazeEnv._get_obs (maze_env.py:351-369) written straight to global memory, env-major
  MMZ_DI void write_obs(const Layout& L, float* obs_g, float* obs_s, int t, bool on) {
    const int naq = m->n_agent_q, nav = m->n_agent_v, no = 3 * L.nobj;
    if (on) {
      for (int i = lane; i < L.obs_core; i += G) {
        float v;
        if (i < 3 && i < naq) v = w[L.o_qpos + i];
        else if (i < 3 + no) v = w[L.o_objpos + i - 3];
        else if (i < naq + no) v = w[L.o_qpos + i - no];
        else if (i < naq + no + nav) v = w[L.o_qvel + i - naq - no];
        else v = t * 0.001f;
        obs_s[i] = v;
        if (obs_g) obs_g[i == L.obs_core - 1 ? L.obs_dim - 1 : i] = v;  // the view kernel fills the gap before t
      }
    }
    sync();
  }

  // reset_model (point.py:71-81, ant.py:84-96, swimmer.py:55-68): same distributions, Philox stream.
  // Only writes qpos / qvel / qacc; the caller syncs and refreshes the derived arrays.
  MMZ_DI void reset_state(const Layout& L, unsigned long long seed, int env, int nreset, bool noise) {
    const float amp = m->reset_noise;
#pragma unroll 1
    for (int i = lane; i < L.nq + L.nv; i += G) {
      const bool isq = i < L.nq;
      const int k = isq ? i : i - L.nq;
      float val = isq ? m->qpos0[k] : 0.f;
      if (noise && k < (isq ? m->n_agent_q : m->n_agent_v)) {
        uint32_t c[4] = {(uint32_t)env, (uint32_t)nreset, (uint32_t)(isq ? k : 64 + k), 0u};
        philox4x32(c, (uint32_t)seed, (uint32_t)(seed >> 32));
        float u = u01(c[0]);
        if (isq || m->reset_kind == MMZ_RESET_SWIMMER) val += amp * (2.f * u - 1.f);
        else if (m->reset_kind == MMZ_RESET_POINT) val += amp * u;
        else val += amp * sqrtf(-2.f * logf(1.f - u)) * cospif(2.f * u01(c[1]));  // Box-Muller
      }
      if (isq) w[L.o_qpos + k] = val;
      else { w[L.o_qvel + k] = val; w[L.o_qacc + k] = 0.f; }
    }
  }

  // MazeEnv.step for this environment (maze_env.py:448-481). Returns the done bits.
  // Control flow is warp-uniform (mmz_dyn.cuh): per-environment decisions are predicates.
  // Tail passes (one kinematics / observation call site): pass 0 closes the step itself (clamp or
  // blow-up refresh, obs, reward, done); pass 1 runs only when some environment of the warp ended
  // its episode with auto-reset on, and re-observes the fresh episode of those environments.
  // `real` is false for the padding environments of the last warp: they compute, but write nothing.
  MMZ_DI unsigned step(const Layout& L, const float* action, float* obs_g, float* obs_s, float* reward, float* info4,
                       int* t_io, int* nreset_io, bool auto_reset, unsigned long long seed, int genv, bool real) {
    float* qpos = w + L.o_qpos;
    float* qvel = w + L.o_qvel;
    const bool teleport = m->step_kind == MMZ_STEP_TELEPORT;
    bool bad = false;
    float inner = 0.f, fwd = 0.f, cc = 0.f;
    int t = *t_io + 1;
    const float before[2] = {qpos[0], qpos[1]};
    float act[MMZ_MAXACT];
#pragma unroll
    for (int a = 0; a < MMZ_MAXACT; a++) act[a] = (real && a < L.nu) ? action[a] : 0.f;
    sync();
    if (teleport) {  // PointEnv.step (point.py:44-61): turn, move, clip qvel; the motors are never driven
      if (lane == 0) {
        float ori = qpos[2] + act[1];
        if (ori < -kPi) ori += 2.f * kPi;
        else if (kPi < ori) ori -= 2.f * kPi;
        float sn, cs;
        sincosf(ori, &sn, &cs);
        qpos[2] = ori;
        qpos[0] += cs * act[0];
        qpos[1] += sn * act[0];
      }
      for (int d = lane; d < L.nv; d += G) qvel[d] = fminf(fmaxf(qvel[d], -m->vel_limit), m->vel_limit);
    }
#pragma unroll
    for (int a = 0; a < MMZ_MAXACT; a++)
      if (a < L.nu && lane == (a % G)) w[L.o_ctrl + a] = teleport ? 0.f : act[a];
    sync();
#pragma unroll 1
    for (int k = 0; k < m->frame_skip; k++) bad = E::mj_step(L, bad);
    bool refresh = bad;
    if (teleport) {
      if (m->manual_collision) {  // maze_env.py:450-464
        float nw[2] = {qpos[0], qpos[1]}, pos[2];
        sync();
        if (clamp_move(m, before, nw, pos) && !bad) {
          if (lane == 0) { qpos[0] = pos[0]; qpos[1] = pos[1]; }
          refresh = true;  // set_xy -> set_state -> mj_forward refreshes xpos
        }
      }
    } else if (!bad) {  // AntEnv.step / SwimmerEnv.step (ant.py:61-73, swimmer.py:37-47)
      float dt = m->timestep * m->frame_skip;
      float vx = (qpos[0] - before[0]) / dt, vy = (qpos[1] - before[1]) / dt;
      // forward_reward_fn (ant.py:18-23): vnorm, vabs, or left to the host wrapper
      const int fk = m->forward_reward_kind;
      fwd = fk == MMZ_FWD_VABS ? fabsf(vx) + fabsf(vy) : fk == MMZ_FWD_HOST ? 0.f : sqrtf(vx * vx + vy * vy);
#pragma unroll
      for (int a = 0; a < MMZ_MAXACT; a++) cc += act[a] * act[a];
      cc *= m->ctrl_cost_weight;
      inner = m->forward_reward_weight * fwd - cc;
    }
    unsigned bits = 0;
    if (bad) bits |= UNSTABLE_BIT;  // MuJoCo's mj_checkPos/Vel/Acc auto-reset: back to qpos0, zero velocity
    bool reset_now = bad, noise = false, live = true;
#pragma unroll 1
    for (int pass = 0; pass < 2; pass++) {
      sync();  // every lane has read the ending episode's qpos (info, task rules) before the reset overwrites it (racecheck)
      if (live && reset_now) reset_state(L, seed, genv, *nreset_io, noise);
      sync();
      latch_objpos(L, live && !refresh);  // stale derived arrays are the reference's behaviour (quirk Q15)
      if (__any_sync(kFull, live && refresh)) {
        E::kinematics(L);
        latch_objpos(L, live && refresh);
      }
      write_obs(L, (real && live && (pass == 1 || !auto_reset)) ? obs_g : nullptr, obs_s, t, live);
      if (pass == 0) {
        float outer;
        bool term;
        task_rules(obs_s, &outer, &term);
        *reward = m->inner_reward_scale * inner + outer;
        if (term) bits |= DONE_BIT;
        // gym's TimeLimit: info['TimeLimit.truncated'] = not done, i.e. only when the task itself did not end the episode
        if (m->max_episode_steps > 0 && t >= m->max_episode_steps) bits |= DONE_BIT | (term ? 0 : TRUNC_BIT);
        info4[0] = qpos[0]; info4[1] = qpos[1]; info4[2] = fwd; info4[3] = -cc;
        live = auto_reset && (bits & DONE_BIT);
        if (!live && auto_reset && real) {  // no reset: the observation just assembled is the one to return
          for (int i = lane; i < L.obs_core; i += G) obs_g[i == L.obs_core - 1 ? L.obs_dim - 1 : i] = obs_s[i];  // as write_obs
        }
        if (live) {  // the env that just ended starts its next episode inside this launch
          *nreset_io += 1;
          t = 0;
          reset_now = true; noise = true; refresh = true;
        }
        if (!__any_sync(kFull, live)) break;
      }
    }
    *t_io = t;
    return bits;
  }

};

// register budget: enough resident groups per SM to hide the latency of the long dependent chains
template <int G> struct LaunchCfg { static constexpr int kMaxThreads = 512, kMinBlocks = 1; };  // <= 128 registers
template <> struct LaunchCfg<32> { static constexpr int kMaxThreads = 256, kMinBlocks = 1; };

template <int G, int NVP, int FEAT, int MODE>
__global__ void __launch_bounds__(LaunchCfg<G>::kMaxThreads, LaunchCfg<G>::kMinBlocks)
maze_kernel(const __grid_constant__ KArgs A) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar;
  const Layout& L = A.L;
  const int tid = threadIdx.x, epb = blockDim.x / G;
  const int env0 = (A.block0 + blockIdx.x) * epb;
  float* wsbase = reinterpret_cast<float*>(smem + ((L.model_bytes + 127) & ~127));

  // ---- model constants: one bulk async copy global -> shared, completion on an mbarrier
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(L.model_bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem)), "l"(A.model), "r"(L.model_bytes), "r"(smem_u32(&bar)) : "memory");
  }
  // ---- state tile: rows x epb consecutive environments (contiguous row segments)
  if (MODE != MODE_RESET || A.mask != nullptr) {
    for (int idx = tid; idx < L.nstate * epb; idx += blockDim.x) {
      int r = idx / epb, e = idx - r * epb;
      if (env0 + e < A.npad) wsbase[e * L.stride + row_offset(L, r)] = A.state[(size_t)r * A.npad + env0 + e];
    }
  }
  __syncthreads();  // barrier initialised + state tile visible
  {
    unsigned done = 0;
    while (!done) {
      asm volatile(
          "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}\n"
          : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    }
  }

  Task<G, NVP, FEAT> T;
  T.m = reinterpret_cast<const mmz_model*>(smem);
  T.dv = reinterpret_cast<const Derived*>(smem + ((sizeof(mmz_model) + 15) & ~15));
  const int ge = tid / G;  // group (environment) inside the block
  T.lane = tid % G;
  T.gshift = (tid % 32) / G * G;
  T.bsync = A.flags >> 8;
  T.w = wsbase + ge * L.stride;
  const int env = env0 + ge;
  float* obs_s = T.w + L.o_obs;
  // npad is a multiple of 32 and a warp's first environment a multiple of 32/G: a warp is either
  // entirely inside [0, npad) or entirely outside, so this branch is warp-uniform. Environments in
  // [n, npad) are padding: they run like real ones (uniform control flow) and write no outputs.
  const bool real = env < A.n;

  if (env < A.npad) {
    int t = A.counters[env], nreset = A.counters[A.npad + env];
    if (MODE == MODE_STEP) {
      float reward, info4[4];
      if (T.lane < 4) T.cnt(L)[N_ITER_SUM + T.lane] = 0;
      T.sync();
      const float* act = A.action + (size_t)env * L.nu;
      unsigned bits = T.step(L, act, A.obs + (size_t)env * L.obs_dim, obs_s, &reward, info4, &t, &nreset,
                             (A.flags & FLAG_AUTO_RESET) != 0, A.seed, A.env_offset + env, real);
      if (T.lane == 0) {
        if (real) { A.reward[env] = reward; A.done[env] = (uint8_t)bits; }
        A.counters[env] = t;
        A.counters[A.npad + env] = nreset;
      }
      if (real && A.info && T.lane < 4) A.info[(size_t)env * 4 + T.lane] = info4[T.lane];
      if (real && A.diag && T.lane < 4) A.diag[(size_t)env * 4 + T.lane] = T.cnt(L)[N_ITER_SUM + T.lane];
      if (A.peers.n && real) {  // fused observation gather: obs_s holds the row that went to A.obs (write_obs)
        T.sync();
        for (int i = T.lane; i < L.obs_core; i += G)
          peer_store(A.peers, (size_t)(A.peers.row0 + env) * L.obs_dim + (i == L.obs_core - 1 ? L.obs_dim - 1 : i), obs_s[i]);
      }
    } else if (MODE == MODE_FORWARD) {
      for (int a = T.lane; a < L.nu; a += G)
        T.w[L.o_ctrl + a] = (T.m->step_kind == MMZ_STEP_TELEPORT || !real) ? 0.f : A.action[(size_t)env * L.nu + a];
      T.sync();
      T.forward(L, false);
      if (real) {
        for (int d = T.lane; d < L.nv; d += G) A.qacc_out[(size_t)env * L.nv + d] = T.w[L.o_qacc + d];
        if (A.diag && T.lane == 0) {
          int* cn = T.cnt(L);
          A.diag[env * 4 + 0] = cn[N_CON];
          A.diag[env * 4 + 1] = cn[N_LIM] + 4 * cn[N_CON];
          A.diag[env * 4 + 2] = cn[N_ITER];
          A.diag[env * 4 + 3] = cn[N_OVERFLOW];
        }
      }
    } else if (MODE == MODE_OBSERVE) {
      T.write_obs(L, real ? A.obs + (size_t)env * L.obs_dim : nullptr, obs_s, t, true);
    } else if (MODE == MODE_RESET) {
      const bool on = A.mask == nullptr || (real && A.mask[env]);
      if (on) {
        nreset += 1;
        t = 0;
        T.reset_state(L, A.seed, A.env_offset + env, nreset, true);
      }
      T.sync();
      // environments that are not reset keep their (possibly stale) derived arrays: they latched
      // their observed-body positions when those were computed, and the tile store writes them back
      T.kinematics(L);
      T.latch_objpos(L, on);
      if (on && T.lane == 0) { A.counters[env] = 0; A.counters[A.npad + env] = nreset; }
      T.write_obs(L, (real && on && A.obs) ? A.obs + (size_t)env * L.obs_dim : nullptr, obs_s, 0, on);
    } else if (MODE == MODE_REFRESH) {  // after set_state: mj_forward refreshes the derived arrays
      for (int d = T.lane; d < L.nv; d += G) T.w[L.o_qacc + d] = 0.f;
      T.sync();
      T.kinematics(L);
      T.latch_objpos(L, true);
    }
  }
  __syncthreads();
  if (MODE == MODE_STEP || MODE == MODE_RESET || MODE == MODE_REFRESH) {
    for (int idx = tid; idx < L.nstate * epb; idx += blockDim.x) {
      int r = idx / epb, e = idx - r * epb;
      if (env0 + e < A.npad) A.state[(size_t)r * A.npad + env0 + e] = wsbase[e * L.stride + row_offset(L, r)];
    }
  }
}

typedef void (*kernel_fn)(const KArgs);
// one translation unit per (G, NVP, FEAT) instance (mmz_inst.cu, compiled in parallel); returns
// nullptr for an instance that is not built (the caller falls back to FEAT_ALL)
kernel_fn get_kernel(int g, int nvp, int feat, int mode);

#ifdef MMZ_API_TU
// [n][k] env-major <-> [k][npad] rows (get_state / set_state)
__global__ void rows_to_env_major(const float* rows, float* out, int n, int npad, int k) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n * k) { int e = i / k, r = i - e * k; out[i] = rows[(size_t)r * npad + e]; }
}
__global__ void env_major_to_rows(const float* in, float* rows, int n, int npad, int k) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n * k) { int e = i / k, r = i - e * k; rows[(size_t)r * npad + e] = in[i]; }
}
__global__ void rows_copy(const float* in, int in_ld, float* out, int out_ld, int n, int k) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n * k) { int r = i / n, e = i - r * n; out[(size_t)r * out_ld + e] = in[(size_t)r * in_ld + e]; }
}

#endif  // MMZ_API_TU

}  // namespace mmz
