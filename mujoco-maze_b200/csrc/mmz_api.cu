// mmz_api.cu - C ABI of libmmz.so (include/mmz.h): handle management and kernel launches.
//
// Boundary notes: plain C signatures, caller-owned device pointers, every call only enqueues
// on the given stream (except mmz_step_host / mmz_destroy), errors are negative return codes
// with a thread-local message. No torch types, no exceptions across the boundary.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <algorithm>
#include <new>
#include <utility>
#include <vector>

#include "../../include/mmz.h"
#define MMZ_API_TU
#include "mmz_kernels.cuh"
#include "mmz_hstep.cuh"
#include "mmz_view.cuh"
#include "mmz_render.cuh"

using namespace mmz;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return code;
}
#define CUDA_TRY(expr)                                                                              \
  do {                                                                                              \
    cudaError_t e_ = (expr);                                                                        \
    if (e_ != cudaSuccess) return fail(MMZ_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e_)); \
  } while (0)

typedef mmz::kernel_fn kernel_fn;

}  // namespace

// kernel instances (one translation unit each, see build_native.py INSTANCES)
#define MMZ_INSTANCES(X) X(8, 4, 1) X(8, 4, 7) X(8, 8, 2) X(8, 8, 7) X(16, 14, 0) X(16, 16, 1) X(16, 16, 7) X(32, 20, 7)
namespace mmz {
hkernel_fn get_hkernel_14(int mode);
hkernel_fn get_hkernel_16(int mode);
hkernel_fn get_hkernel_4(int mode);
#define MMZ_DECL(g, nvp, feat) kernel_fn get_kernel_##g##_##nvp##_##feat(int mode);
MMZ_INSTANCES(MMZ_DECL)
#undef MMZ_DECL
kernel_fn get_kernel(int g, int nvp, int feat, int mode) {
#define MMZ_PICK(G_, NVP_, FEAT_) if (g == G_ && nvp == NVP_ && feat == FEAT_) return get_kernel_##G_##_##NVP_##_##FEAT_(mode);
  MMZ_INSTANCES(MMZ_PICK)
#undef MMZ_PICK
  return nullptr;
}
}  // namespace mmz

struct StepGraphKey {  // everything a recorded mmz_step_k graph bakes in
  int K;
  const float* actions; float* obs; float* reward; uint8_t* done; float* info; int32_t* diag;
  uint32_t flags; int env_offset; uint64_t seed; float* peer0; long long peer_row0;
};

struct mmz_env {
  int device = 0;
  int n = 0, npad = 0;
  unsigned flags = 0;
  int G = 32, NVP = 20;  // kernel instance
  int feat = 0;          // FEAT_* bits the instance was compiled with
  unsigned bsync = 1;    // block-barrier placement inside the step (mmz_dyn.cuh: Env::bsync)
  int tpb = 128;         // threads per block
  int smem_bytes = 0;
  int envs_per_sm = 0;
  int env_offset = 0;    // global index of env 0 (multi-GPU sharding keeps the Philox streams global)
  Layout L;
  mmz_model hm;          // host copy (float layout)
  void* d_model = nullptr;
  float* d_state = nullptr;
  int* d_counters = nullptr;
  // staging for mmz_step_host
  float *d_action = nullptr, *d_obs = nullptr, *d_reward = nullptr, *d_info = nullptr;
  uint8_t* d_done = nullptr;
  // mmz_step_host pipeline: block ranges of the batch on their own streams (copies of one range under the kernel of another)
  cudaStream_t hs[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t hev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  unsigned long long seed = 0;
  unsigned long long launches = 0;
  int32_t* d_step_diag = nullptr;  // caller-owned, optional
  kernel_fn fn[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  // hybrid kernel (mmz_hkernel.cuh), used when the model is eligible
  bool use_t = false;
  char kname[48] = "";
  float tol = 2e-6f;  // Newton convergence tolerance of the hybrid kernel (fp32 round-off floor)
  TLayout TL;
  ObsPeers peers = {};  // fused observation gather (mmz_set_obs_peers)
  cudaGraphExec_t kgraph = nullptr;  // mmz_step_k: the K launches of the last (K, buffers) combination
  StepGraphKey kgraph_key;
  uint64_t launches_per_kgraph = 0;
  mmz::hkernel_fn tfn[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
};

namespace {

int round_up(int x, int a) { return (x + a - 1) / a * a; }

void make_layout(const mmz_model& m, int G, int NVP, int maxcon, Layout* out) {
  Layout L;
  memset(&L, 0, sizeof L);
  L.nb = m.nbody; L.nj = m.njnt; L.nv = m.nv; L.nq = m.nq; L.nu = m.nu; L.ng = m.ngeom; L.nobj = m.nobj;
  L.obs_dim = m.obs_dim; L.nlatch = m.nobj + m.nviewb; L.obs_core = m.obs_dim - m.view_dim;
  L.ldm = m.nv | 1;
  L.maxcon = maxcon;
  L.cstride = C_STRIDE;
  L.nstate = m.nq + 2 * m.nv + 3 * L.nlatch;
  int o = 0;
  auto take = [&](int n) { int r = o; o += n; return r; };
  L.o_qpos = take(L.nq); L.o_qvel = take(L.nv); L.o_ctrl = take(L.nu > 0 ? L.nu : 1);
  L.o_q0 = take(L.nq); L.o_v0 = take(L.nv); L.o_xv = take(L.nv); L.o_fa = take(L.nv);
  L.o_accv = take(L.nv); L.o_acca = take(L.nv);
  L.o_xpos = take(3 * L.nb); L.o_xquat = take(4 * L.nb); L.o_xmat = take(9 * L.nb);
  L.o_xipos = take(3 * L.nb); L.o_ximat = take(9 * L.nb);
  L.o_xanchor = take(3 * L.nj); L.o_xaxis = take(3 * L.nj);
  L.o_gpos = take(3 * L.ng); L.o_gmat = take(9 * L.ng);
  L.o_cdof = take(6 * L.nv);
  L.o_iw = take(10 * L.nb);
  // composite inertias are dead once M is built; the RNE velocity / acceleration arrays reuse them
  L.o_ic = take(12 * L.nb); L.o_vel = L.o_ic; L.o_acc = L.o_ic + 6 * L.nb;
  L.o_frc = take(6 * L.nb); L.o_fsub = take(6 * L.nb);
  L.o_M = take(L.nv * L.ldm);
  L.o_smooth = take(L.nv); L.o_qacc = take(L.nv); L.o_dir = take(L.nv);
  L.o_con = take(L.maxcon * L.cstride);
  L.o_cnt = take(N_CNT);
  L.o_objpos = take(3 * L.nlatch > 0 ? 3 * L.nlatch : 1);
  L.o_obs = take(L.obs_core);
  // stride % 32 == G % 32 so that the groups of one warp fall on disjoint bank ranges
  int stride = round_up(o, 32) + (G % 32);
  if (stride - 32 >= o) stride -= 32;
  L.stride = stride;
  L.model_bytes = round_up(round_up((int)sizeof(mmz_model), 16) + (int)sizeof(Derived), 16);
  *out = L;
}

// sphere-sphere / sphere-capsule pairs between moving bodies that pass the contact filter, ordered like MuJoCo (lower
// geom type first, then lower id). Returns -1 for a pair the kernels cannot handle (capsule-capsule), else the count.
int list_pairs(const mmz_model& m, int* pa, int* pb, int cap) {
  int n = 0;
  if (!m.collision_on) return 0;  // option collision="predefined" without pairs (swimmer.xml, reacher.xml)
  for (int g1 = 0; g1 < m.ngeom; g1++)
    for (int g2 = g1 + 1; g2 < m.ngeom; g2++) {
      const int b1 = m.geom_body[g1], b2 = m.geom_body[g2];
      if (b1 == b2 || m.body_parent[b1] == b2 || m.body_parent[b2] == b1) continue;
      if (!((m.geom_contype[g1] & m.geom_conaffinity[g2]) || (m.geom_contype[g2] & m.geom_conaffinity[g1]))) continue;
      int a = g1, b = g2;
      if (m.geom_type[a] > m.geom_type[b]) { a = g2; b = g1; }
      if (m.geom_type[b] == MMZ_GEOM_BOX) continue;  // handled with the box geoms
      if (m.geom_type[a] != MMZ_GEOM_SPHERE) return -1;
      if (n < cap) { if (pa) pa[n] = a; if (pb) pb[n] = b; }
      n++;
    }
  return n;
}

void make_derived(const mmz_model& m, Derived* d) {
  memset(d, 0, sizeof *d);
  d->npair = std::min((int)MMZ_MAXPAIR, std::max(0, list_pairs(m, d->pair_a, d->pair_b, MMZ_MAXPAIR)));
  int nlev = 0;
  for (int b = 0; b < m.nbody; b++) {
    int mask = 0;
    for (int a = b; a >= 0; a = m.body_parent[a]) mask |= 1 << a;
    d->anc[b] = mask;
    if (m.body_level[b] + 1 > nlev) nlev = m.body_level[b] + 1;
  }
  d->nlev = nlev;
  d->ident[0] = d->ident[4] = d->ident[8] = 1.f;
  for (int g = 0; g < m.ngeom; g++)
    if (m.geom_type[g] == MMZ_GEOM_BOX) d->boxg[d->nboxg++] = g;
}

int validate(const mmz_model& m) {
  if (m.magic != MMZ_MAGIC) return fail(MMZ_ERR_MODEL, "bad magic 0x%x", m.magic);
  if (m.version != MMZ_VERSION) return fail(MMZ_ERR_MODEL, "blob version %d, library expects %d", m.version, MMZ_VERSION);
  if (m.real_bytes != 4) return fail(MMZ_ERR_MODEL, "blob must use the float layout (real_bytes=4), got %d", m.real_bytes);
  if (m.nbody < 1 || m.nbody > MMZ_MAXBODY || m.njnt < 1 || m.njnt > MMZ_MAXJNT || m.nv < 1 || m.nv > MMZ_MAXDOF ||
      m.nq < 1 || m.nq > MMZ_MAXQ || m.ngeom < 0 || m.ngeom > MMZ_MAXGEOM || m.nu < 0 || m.nu > MMZ_MAXACT ||
      m.ngoal < 0 || m.ngoal > MMZ_MAXGOAL || m.nseg < 0 || m.nseg > MMZ_MAXSEG || m.nobj < 0 || m.nobj > 4 || m.nviewb < 0 || m.nobj + m.nviewb > MMZ_MAXOBJ ||
      m.grid_h * m.grid_w > MMZ_MAXCELL || m.grid_h < 1 || m.grid_w < 1)
    return fail(MMZ_ERR_CAPACITY, "model dimensions exceed the compiled capacities");
  if ((m.view_dim != 0 && (m.view_dim != MMZ_VIEW_DIM || m.nviewb < 1)) ||
      m.obs_dim != m.n_agent_q + m.n_agent_v + 3 * m.nobj + m.view_dim + 1 || m.obs_dim - m.view_dim > 64)
    return fail(MMZ_ERR_MODEL, "inconsistent obs_dim %d", m.obs_dim);
  for (int b = 0; b < m.nbody; b++)
    if (m.body_parent[b] >= b) return fail(MMZ_ERR_MODEL, "bodies must be ordered parents first");
  for (int j = 0; j < m.njnt; j++)
    if (m.jnt_type[j] == MMZ_JNT_BALL) return fail(MMZ_ERR_MODEL, "ball joints are not supported");
  {
    const int np = list_pairs(m, nullptr, nullptr, 0);
    if (np < 0) return fail(MMZ_ERR_MODEL, "capsule-capsule contacts between moving bodies are not supported");
    if (np > MMZ_MAXPAIR) return fail(MMZ_ERR_CAPACITY, "%d moving geom pairs exceed the capacity of %d", np, (int)MMZ_MAXPAIR);
  }
  return MMZ_OK;
}

// MazeEnv.get_top_down_view for tasks with TOP_DOWN_VIEW (none in the reference's registry): a second small launch
// fills the 75 view columns of the observations the step / reset / observe kernel has just written.
int launch_view(mmz_env* h, int mode, const KArgs& A, cudaStream_t s) {
  if (!h->hm.view_dim || !A.obs || (mode != MODE_STEP && mode != MODE_RESET && mode != MODE_OBSERVE)) return MMZ_OK;
  ViewArgs V;
  V.model = (const mmz_model*)h->d_model; V.state = h->d_state; V.mask = mode == MODE_RESET ? A.mask : nullptr;
  V.obs = A.obs; V.n = h->n; V.npad = h->npad;
  const int threads = 256, total = h->n * MMZ_VIEW_DIM;
  maze_view_kernel<<<(total + threads - 1) / threads, threads, 0, s>>>(V);
  h->launches++;
  CUDA_TRY(cudaGetLastError());
  return MMZ_OK;
}

int launch(mmz_env* h, int mode, KArgs& A, cudaStream_t s, int block0 = 0, int nblocks = 0) {
  if (h->use_t) {
    TArgs T;
    memset(&T, 0, sizeof T);
    T.L = h->TL;
    T.model = h->d_model; T.state = h->d_state; T.counters = h->d_counters;
    T.n = h->n; T.npad = h->npad;
    T.action = A.action; T.obs = A.obs; T.reward = A.reward; T.done = A.done; T.info = A.info;
    T.qacc_out = A.qacc_out; T.diag = A.diag; T.mask = A.mask; T.seed = A.seed;
    T.flags = h->flags; T.env_offset = h->env_offset; T.tol = h->tol;
    T.block0 = block0;
    if (mode == MODE_STEP) T.peers = h->peers;
    h->tfn[mode]<<<nblocks > 0 ? nblocks : h->npad / TE, TW * 32, h->smem_bytes, s>>>(T);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return launch_view(h, mode, A, s);
  }
  A.L = h->L;
  A.model = h->d_model;
  A.state = h->d_state;
  A.counters = h->d_counters;
  A.n = h->n;
  A.npad = h->npad;
  A.flags = h->flags | (h->bsync << 8);
  A.env_offset = h->env_offset;
  int epb = h->tpb / h->G;
  int blocks = nblocks > 0 ? nblocks : (h->npad + epb - 1) / epb;  // padding environments run too (warp-uniform control flow)
  A.block0 = block0;
  if (mode == MODE_STEP) A.peers = h->peers; else memset(&A.peers, 0, sizeof A.peers);
  h->fn[mode]<<<blocks, h->tpb, h->smem_bytes, s>>>(A);
  h->launches++;
  CUDA_TRY(cudaGetLastError());
  return launch_view(h, mode, A, s);
}

int nbox_geoms(const mmz_model& m) {
  int n = 0;
  for (int g = 0; g < m.ngeom; g++) n += m.geom_type[g] == MMZ_GEOM_BOX;
  return n;
}

// The hybrid kernel (mmz_hkernel.cuh) handles torque-driven agents with at most 16 dofs whose moving geoms are
// spheres / capsules that only touch the world (floor, maze boxes): the Ant family. Returns false if the model is
// not eligible (the lanes-per-environment kernel of mmz_dyn.cuh serves everything).
bool configure_h(mmz_env* h, int* rc) {
  const mmz_model& m = h->hm;
  *rc = MMZ_OK;
  const char* kenv = getenv("MMZ_KERNEL");  // development aid: "groups" forces the lanes-per-environment kernel
  if (kenv && !strcmp(kenv, "groups")) return false;
  if (m.density > 0.f || m.viscosity > 0.f) return false;
  if (m.nv > 16) return false;
  int nbox = 0;
  for (int g = 0; g < m.ngeom; g++) {
    const int t = m.geom_type[g];
    if (t != MMZ_GEOM_SPHERE && t != MMZ_GEOM_CAPSULE && t != MMZ_GEOM_BOX) return false;
    nbox += t == MMZ_GEOM_BOX;
    for (int g2 = g + 1; g2 < m.ngeom; g2++) {  // a moving-moving pair that passes the contact filter must involve a box
      const int b1 = m.geom_body[g], b2 = m.geom_body[g2];
      if (b1 == b2 || m.body_parent[b1] == b2 || m.body_parent[b2] == b1) continue;
      const bool pair = (m.geom_contype[g] & m.geom_conaffinity[g2]) || (m.geom_contype[g2] & m.geom_conaffinity[g]);
      if (pair && t != MMZ_GEOM_BOX && m.geom_type[g2] != MMZ_GEOM_BOX) return false;
    }
  }
  // the subtree sums walk [b, sub_end[b]): the bodies must be in depth-first order (MJCF order is)
  for (int b = 0; b < m.nbody; b++) {
    bool in_run = true;
    for (int c = b + 1; c < m.nbody; c++) {
      bool desc = false;
      for (int a = c; a >= 0; a = m.body_parent[a]) desc |= a == b;
      if (desc && !in_run) return false;
      if (!desc) in_run = false;
    }
  }
  const bool box = nbox > 0;
  // instances built: <14, no boxes> (solver v2) and <16, boxes>. The second also takes the small robots with a box geom
  // (the Point and its arrow, 3 dofs, teleport step + manual wall clamp): a step of them is bound by the LATENCY of a few
  // long dependent chains, which the lane = environment tree phases run for 32 environments per instruction.
  if (!box && (m.nv > 14 || m.nv < 9)) return false;
  if (m.nv < 9) {
    // Measured on B200 (profiles/r2_bench.md): PointUMaze 4096 envs 0.272 -> 0.239 ms, PointPush 65536 envs 5.84 -> 3.74 ms,
    // but the single-body Point at 65536 envs 1.83 -> 2.38 ms (the lanes kernel keeps 64 of them per SM in flight).
    bool use = m.nbody > 1 || h->n <= 16384;
    if (const char* e = getenv("MMZ_POINT_HYBRID")) use = atoi(e) != 0;  // development aid
    if (!use) return false;
  }
  TLayout L;
  memset(&L, 0, sizeof L);
  L.tail0 = -1;
  L.nb = m.nbody; L.nj = m.njnt; L.nv = m.nv; L.nq = m.nq; L.nu = m.nu; L.ng = m.ngeom; L.nobj = m.nobj; L.obs_dim = m.obs_dim;
  L.nlatch = m.nobj + m.nviewb; L.obs_core = m.obs_dim - m.view_dim;
  int nlev = 0;
  for (int b = 0; b < m.nbody; b++)
    if (m.body_level[b] + 1 > nlev) nlev = m.body_level[b] + 1;
  L.nlev = nlev; L.ldm = m.nv + 1;
  L.cstride = C_STRIDE;
  L.nstate = m.nq + 2 * m.nv + 3 * L.nlatch;
  const int nitems = L.ng + nbox * (1 + 2 * 9 + (nbox - 1));  // HEnv::n_items
  if (nitems + 2 * m.nv > kMaxCTasks) return false;  // the phase C schedule of TDerived
  const int nvp = box ? (m.nv <= 4 ? 4 : 16) : 14;  // the solver reads qacc / dir up to the instance's padded nv (zero beyond nv)
  L.model_bytes = round_up(round_up((int)sizeof(mmz_model), 16) + (int)sizeof(TDerived), 16);
  int dev_smem = 0;
  if (cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device) != cudaSuccess) return false;
  int o = 0;
  auto take = [&](int n) { int r = o; o += n; return r; };
  L.o_cnt = take(TN_CNT);
  L.o_qpos = take(L.nq); L.o_qvel = take(L.nv); L.o_qacc = take(nvp); L.o_objpos = take(3 * L.nlatch > 0 ? 3 * L.nlatch : 1);
  L.o_ctrl = take(L.nu > 0 ? L.nu : 1); L.o_act = take(L.nu > 0 ? L.nu : 1);
  L.o_q0 = take(L.nq); L.o_v0 = take(L.nv); L.o_accv = take(L.nv); L.o_acca = take(L.nv);
  L.o_xpos = take(3 * L.nb);
  L.o_lim = take(4 * L.nv);
  if (!box) {
    // Solver v2. In front: what stays live while the solver iterates. Then everything that is dead by then - the
    // solver view loads the mass matrix, the smooth forces and the motion axes into registers and passes a block
    // barrier first - and the stored contact Jacobian and the per-contact forces / weights OVERLAY it (natural
    // [environment][contact] order, float4). The contact records come last.
    L.v2 = 1;
    L.cstride = K_STRIDE;
    {  // the Ant's dof tree: a free root and 2-dof chains hanging off its last dof
      bool ant = m.nv == 14 && m.jnt_type[0] == MMZ_JNT_FREE;
      for (int d = 0; d < m.nv && ant; d++) ant = m.dof_parent[d] == (d < 6 ? d - 1 : (d & 1) ? d - 1 : 5);
      L.topo = ant ? 1 : 0;
    }
    L.o_dir = take(nvp);
    L.o_nat = o = round_up(o, 4);  // 4 slots = 528 bytes: the float4 areas start 16-byte aligned
    L.o_cdof = take(6 * L.nv);
    L.o_M = take(L.nv * L.ldm);
    L.o_smooth = take(L.nv);
    L.o_obs = take(L.obs_core);
    L.o_gpos = take(3 * L.ng); L.o_gax = take(3 * L.ng); L.o_gmat = take(1);
    L.o_vel = take(6 * L.nb);
    L.o_gcnt = take(nitems);
    L.o_xquat = take(4 * L.nb); L.o_xmat = take(9 * L.nb); L.o_iw = take(10 * L.nb);
    L.o_ic = take(10 * L.nb); L.o_acc = take(6 * L.nb); L.o_frc = take(6 * L.nb); L.o_fsub = take(6 * L.nb);
    const int dead1 = o;
    const int static_smem = 256;  // mbarrier
    const int avail = (dev_smem - round_up(L.model_bytes, 128) - static_smem) / (HS * 4);  // slots per environment
    int maxcon = 16;  // one lane per contact in the contact passes
    for (; maxcon >= 10; maxcon--) {
      L.jes = maxcon * (nvp + 1);
      L.fa_off = TE * L.jes + 1;
      const int nat_bytes = (L.fa_off + TE * 2 * maxcon) * 16;
      L.o_con = std::max(dead1, L.o_nat + (nat_bytes + HS * 4 - 1) / (HS * 4));
      L.nslots = L.o_con + maxcon * L.cstride;
      if (L.nslots <= avail) break;
    }
    if (maxcon < 10) return false;  // does not fit: use the lanes-per-environment kernel
    L.maxcon = maxcon;
  } else {
    // Solver v3 (mmz_hkernel.cuh: solve_g3). In front: what stays live while the solver iterates. Then [A] the arrays the
    // solver view has loaded into registers by its block barrier, or that are dead by then (motion axes, mass matrix,
    // smooth forces, geom poses, body velocities, contact counts, rotation matrices): the natural-order area - per
    // environment a pool of `nj` Jacobian entries, (force, weights) pairs and the 16 Hessian rows - OVERLAYS them and
    // runs on into its own extension. Last [B] the arrays that are dead once mass matrix and smooth forces exist
    // (quaternions, inertias, bias accelerations and forces): the contact records overlay THOSE, as before.
    L.v3 = 1;
    L.cstride = K3_STRIDE;
    // dofs 0..5 are one free joint: every contact's dof mask holds all six of them or none (bit 2 of topo, solve_g3)
    if (m.nv >= 6 && m.jnt_type[0] == MMZ_JNT_FREE && m.jnt_dadr[0] == 0) L.topo |= 2;
    {  // the Ant's dof tree (a free root, four 2-dof chains) and a second tree, the 2-dof chain (14, 15) of one movable block: bit 4
      bool ab = m.nv == 16 && (L.topo & 2);
      for (int d = 0; d < m.nv && ab; d++)
        ab = m.dof_parent[d] == (d < 6 ? d - 1 : d == 14 ? -1 : (d & 1) ? d - 1 : 5);
      if (ab) L.topo |= 4;
    }
    {  // the last dof tree has exactly two dofs (a movable block on two slides): its own contacts are handled lane = contact
      int t0 = m.nv - 1;
      while (t0 > 0 && m.dof_parent[t0] >= 0) t0--;
      L.tail0 = (m.nv >= 3 && m.nv - t0 == 2 && t0 > 0) ? t0 : -1;
    }
    L.o_dir = take(nvp);
    L.o_nat = o = round_up(o, 4);
    L.o_cdof = take(6 * L.nv);
    L.o_M = take(L.nv * L.ldm);
    L.o_smooth = take(L.nv);
    L.o_obs = take(L.obs_core);
    L.o_gpos = take(3 * L.ng); L.o_gax = take(3 * L.ng); L.o_gmat = take(nbox > 0 ? 9 * nbox : 1);
    L.o_vel = take(6 * L.nb);
    L.o_gcnt = take(nitems);
    L.o_xmat = take(9 * L.nb);
    const int a_end = o;
    const int tail = 4 * L.nb + 10 * L.nb + 10 * L.nb + 6 * L.nb + 6 * L.nb + 6 * L.nb;  // xquat, iw, ic, acc, frc, fsub
    const int static_smem = 256;  // mbarrier
    const int avail = (dev_smem - round_up(L.model_bytes, 128) - static_smem) / (HS * 4);  // slots per environment
    const int maxcon = std::min(32, 16 + 8 * nbox);  // two trips of 16 lanes
    const int hq = (16 * (nvp + 1) + 3) / 4;        // float4s of the 16 Hessian rows
    // (a contact has at most nv entries: the Point's pool is 72 entries, not 192 - with ~150 KB of shared memory per block
    // instead of 227 the SM keeps ~90 KB of L1, which is where the LOCAL arrays of the box-box narrow phase live; with the
    // full carve-out they spill to L2 and the narrow phase, the critical path of the small robots' step, runs ~2x slower)
    int nj = m.nv <= 4 ? std::min(192, round_up(maxcon * m.nv, 8)) : 192;
    for (; nj >= 48; nj -= 8) {
      L.es = nj + 2 * maxcon + hq;
      const int nat_bytes = (TE * L.es + 1) * 16;
      const int tail0 = std::max(a_end, L.o_nat + (nat_bytes + HS * 4 - 1) / (HS * 4));
      L.nslots = tail0 + std::max(tail, maxcon * L.cstride);
      if (L.nslots <= avail) {
        o = tail0;
        break;
      }
    }
    if (nj < 48) return false;  // does not fit: use the lanes-per-environment kernel
    L.njac = nj;
    L.maxcon = maxcon;
    L.o_con = o;
    L.o_xquat = take(4 * L.nb); L.o_iw = take(10 * L.nb); L.o_ic = take(10 * L.nb);
    L.o_acc = take(6 * L.nb); L.o_frc = take(6 * L.nb); L.o_fsub = take(6 * L.nb);
  }
  h->TL = L;
  h->smem_bytes = round_up(L.model_bytes, 128) + L.nslots * HS * 4;
  for (int mode = 0; mode < 5; mode++) {
    h->tfn[mode] = !box ? mmz::get_hkernel_14(mode) : nvp == 4 ? mmz::get_hkernel_4(mode) : mmz::get_hkernel_16(mode);
    cudaError_t e = cudaFuncSetAttribute(h->tfn[mode], cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_bytes);
    if (e != cudaSuccess) { *rc = fail(MMZ_ERR_CUDA, "cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return false; }
  }
  h->use_t = true;
  h->G = 16; h->NVP = nvp; h->feat = box ? FEAT_BOX : 0; h->tpb = TW * 32; h->envs_per_sm = TE;
  h->L.stride = L.nslots; h->L.nstate = L.nstate; h->L.model_bytes = L.model_bytes;
  return true;
}

void make_tderived(const mmz_model& m, TDerived* d) {
  memset(d, 0, sizeof *d);
  int nlev = 0;
  for (int b = 0; b < m.nbody; b++) {
    int mask = 0;
    for (int a = b; a >= 0; a = m.body_parent[a]) mask |= 1 << a;
    d->anc[b] = mask;
    if (m.body_level[b] + 1 > nlev) nlev = m.body_level[b] + 1;
  }
  d->nlev = nlev;
  int k = 0;
  for (int l = 0; l < nlev; l++) {
    d->lvl_off[l] = k;
    for (int b = 0; b < m.nbody; b++)
      if (m.body_level[b] == l) d->lvl_body[k++] = b;
  }
  d->lvl_off[nlev] = k;
  {
    // Tree walk roles. With enough warps every chain (the subtree of a level-1 body) gets a PAIR of warps from the top
    // of the block (kinematic halves / dynamic halves, named barrier 1 + chain), and the roots' warps (0..nroots-1) do
    // the roots' dynamic halves after the block barrier; otherwise chain c is walked whole by warp c % TW.
    int nroots = 0, nl1 = 0;
    for (int b = 0; b < m.nbody; b++) { nroots += m.body_level[b] == 0; nl1 += m.body_level[b] == 1; }
    const bool pairs = nroots + 2 * nl1 <= TW && nl1 <= 14 && nl1 > 0;
    int chain[MMZ_MAXBODY], c1 = 0;
    for (int b = 0; b < m.nbody; b++) {  // parents come first in the blob
      chain[b] = -1;
      if (m.body_level[b] == 1) chain[b] = c1++;
      else if (m.body_level[b] > 1) chain[b] = chain[m.body_parent[b]];
    }
    auto chain_of_warp = [&](int w, int* kind, int* bar) {
      *kind = A_IDLE; *bar = 0;
      if (!pairs) { *kind = A_BOTH; return w; }  // warp w walks chains w, w + TW, ... whole
      const int k = TW - 1 - w;                   // warps TW-1, TW-2: chain 0; TW-3, TW-4: chain 1; ...
      if (k < 2 * nl1) { *kind = (k & 1) ? A_PAIR_DYN : A_PAIR_KIN; *bar = 1 + k / 2; return k / 2; }
      if (w < nroots) *kind = A_ROOTDYN;
      return -1;
    };
    int n = 0;
    for (int w = 0; w < TW; w++) {
      int kind, bar;
      const int c = chain_of_warp(w, &kind, &bar);
      d->walk_kind[w] = kind; d->walk_bar[w] = bar;
      d->chain_off[w] = n;
      for (int b = 0; b < m.nbody; b++)
        if (chain[b] >= 0 && (pairs ? chain[b] == c : chain[b] % TW == c)) d->chain_body[n++] = b;
    }
    d->chain_off[TW] = n;
    d->walk_root_count = pairs ? 32 * (std::min(nroots, (int)TW) + nl1) : 0;
  }
  for (int b = 0; b < m.nbody; b++) {
    int e = b + 1;
    while (e < m.nbody && (d->anc[e] >> b & 1)) e++;
    d->sub_end[b] = e;
  }
  for (int k = 0; k < m.nu; k++) d->dof_act[m.act_dof[k]] |= 1 << k;
  for (int i = 0; i < m.nv; i++)
    for (int j = i; j >= 0; j = m.dof_parent[j]) { d->dof_rel[i] |= 1 << j; d->dof_rel[j] |= 1 << i; }
  {
    auto q2m = [](const float* q, float* R) {
      double n = std::sqrt((double)q[0] * q[0] + (double)q[1] * q[1] + (double)q[2] * q[2] + (double)q[3] * q[3]);
      double w = q[0] / n, x = q[1] / n, y = q[2] / n, z = q[3] / n;
      const double M[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                           2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                           2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)};
      for (int k = 0; k < 9; k++) R[k] = (float)M[k];
    };
    for (int b = 0; b < m.nbody; b++) {
      q2m(m.body_iquat[b], d->iq_R[b]);
      bool fast = m.body_parent[b] >= 0 && m.body_jntnum[b] == 1 && m.jnt_type[m.body_jntadr[b]] == MMZ_JNT_HINGE;
      for (int c = 0; c < m.nbody; c++)  // a child on the generic path would need this body's quaternion
        if (m.body_parent[c] == b && !(m.body_jntnum[c] == 1 && m.jnt_type[m.body_jntadr[c]] == MMZ_JNT_HINGE)) fast = false;
      if (getenv("MMZ_NO_FAST_KIN")) fast = false;  // development aid
      d->kin_fast[b] = fast ? 1 : 0;
      {  // 2: a root body without rotation on slide x, slide y, hinge z through its origin (the Point)
        const int j = m.body_jntadr[b];
        auto is = [&](int jj, int type, int axis) {
          for (int k = 0; k < 3; k++)
            if (m.jnt_axis[jj][k] != (k == axis ? 1.f : 0.f) || m.jnt_pos[jj][k] != 0.f) return false;
          return m.jnt_type[jj] == type;
        };
        const bool planar = m.body_parent[b] < 0 && m.body_jntnum[b] == 3 && m.body_quat[b][0] == 1.f && m.body_quat[b][1] == 0.f &&
                            m.body_quat[b][2] == 0.f && m.body_quat[b][3] == 0.f && is(j, MMZ_JNT_SLIDE, 0) && is(j + 1, MMZ_JNT_SLIDE, 1) &&
                            is(j + 2, MMZ_JNT_HINGE, 2) && m.jnt_qadr[j + 1] == m.jnt_qadr[j] + 1 && m.jnt_qadr[j + 2] == m.jnt_qadr[j] + 2 &&
                            m.jnt_dadr[j + 1] == m.jnt_dadr[j] + 1 && m.jnt_dadr[j + 2] == m.jnt_dadr[j] + 2;
        if (planar && !getenv("MMZ_NO_FAST_KIN")) d->kin_fast[b] = 2;
      }
      if (!fast) continue;
      const int j = m.body_jntadr[b];
      float* Rb = d->kin_Rb[b];
      q2m(m.body_quat[b], Rb);
      double a[3] = {m.jnt_axis[j][0], m.jnt_axis[j][1], m.jnt_axis[j][2]};
      const double an = std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
      for (double& v : a) v /= an;
      const double K[9] = {0, -a[2], a[1], a[2], 0, -a[0], -a[1], a[0], 0};
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {
          d->kin_K[b][3 * r + c] = (float)K[3 * r + c];
          d->kin_K2[b][3 * r + c] = (float)(K[3 * r] * K[c] + K[3 * r + 1] * K[3 + c] + K[3 * r + 2] * K[6 + c]);
        }
      for (int r = 0; r < 3; r++) {
        d->kin_c[b][r] = m.body_pos[b][r] + Rb[3 * r] * m.jnt_pos[j][0] + Rb[3 * r + 1] * m.jnt_pos[j][1] + Rb[3 * r + 2] * m.jnt_pos[j][2];
        d->kin_ax[b][r] = (float)(Rb[3 * r] * a[0] + Rb[3 * r + 1] * a[1] + Rb[3 * r + 2] * a[2]);
      }
    }
  }
  d->ident[0] = d->ident[4] = d->ident[8] = 1.f;
  for (int g = 0; g < MMZ_MAXGEOM; g++) d->boxord[g] = -1;
  for (int g = 0; g < m.ngeom; g++)
    if (m.geom_type[g] == MMZ_GEOM_BOX) { d->boxord[g] = d->nboxg; d->boxg[d->nboxg++] = g; }
  {
    // Phase C schedule: tasks [0, nit) collision items, [nit, nit + nv) mass-matrix rows, then nv smooth forces;
    // longest first onto the least loaded warp. Costs in units of ~250 cycles of one warp (measured on the Ant).
    const int nit = m.ngeom + d->nboxg * (1 + 2 * 9 + (d->nboxg - 1)), nt = nit + 2 * m.nv;
    std::vector<std::pair<int, int>> tasks;  // (cost, task)
    for (int t = 0; t < nt; t++) {
      int cost;
      if (t < m.ngeom) cost = m.geom_type[t] == MMZ_GEOM_CAPSULE ? 10 : m.geom_type[t] == MMZ_GEOM_SPHERE ? 7 : 1;
      else if (t < nit) cost = ((t - m.ngeom) % (1 + 2 * 9 + (d->nboxg - 1)) == 0) ? 6 : 14;  // plane-box : box-box
      else if (t < nit + m.nv) { cost = 2; for (int j = t - nit; j >= 0; j = m.dof_parent[j]) cost++; }
      else cost = 2;
      tasks.push_back({cost, t});
    }
    std::stable_sort(tasks.begin(), tasks.end(), [](const std::pair<int, int>& a, const std::pair<int, int>& b) { return a.first > b.first; });
    std::vector<int> load(TW, 0);
    std::vector<std::vector<int>> mine(TW);
    for (auto& ct : tasks) {
      int w = 0;
      for (int k = 1; k < TW; k++)
        if (load[k] < load[w]) w = k;
      load[w] += ct.first;
      mine[w].push_back(ct.second);
    }
    int n = 0;
    for (int w = 0; w < TW; w++) {
      d->c_off[w] = n;
      for (int t : mine[w])
        if (n < kMaxCTasks) d->c_item[n++] = t;
    }
    d->c_off[TW] = n;
  }
}

int configure(mmz_env* h, int G, int NVP) {
  h->G = G;
  h->NVP = NVP;
  // contact capacity: generous for box geoms (up to 8 points per box pair and several pairs per block), 16 otherwise
  int nbox = 0;
  for (int g = 0; g < h->hm.ngeom; g++) nbox += h->hm.geom_type[g] == MMZ_GEOM_BOX;
  int maxcon = h->hm.collision_on ? (nbox ? std::min(40, 16 + 8 * nbox) : 16) : 1;  // 1 block: 24, 2: 32, 3+: 40
  if (h->hm.collision_on && list_pairs(h->hm, nullptr, nullptr, 0) > 0) maxcon = std::min(40, maxcon + 8);  // object balls
  if (h->hm.ngeom <= 2 && nbox) maxcon = 16;
  make_layout(h->hm, G, NVP, maxcon, &h->L);
  int dev_smem = 0, sms = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
  CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
  const int model_smem = round_up(h->L.model_bytes, 128);
  int best_tpb = 0, best_envs = 0, best_smem = 0;
  bool best_spreads = false;
  // features this model needs; an instance compiled without the unused ones is preferred
  int feat = 0;
  if (nbox) feat |= FEAT_BOX;
  if (h->hm.density > 0.f || h->hm.viscosity > 0.f) feat |= FEAT_FLUID;
  if (list_pairs(h->hm, nullptr, nullptr, 0) > 0) feat |= FEAT_PAIR;
  if (!mmz::get_kernel(G, NVP, feat, 0)) feat = FEAT_ALL;
  h->feat = feat;
  for (int mode = 0; mode < 5; mode++) {
    h->fn[mode] = mmz::get_kernel(G, NVP, feat, mode);
    if (!h->fn[mode]) return fail(MMZ_ERR_CAPACITY, "no kernel instance for G=%d NVP=%d", G, NVP);
  }
  // block size: any whole number of warps up to the launch bound; keep the one with most resident envs
  const int max_tpb = G == 32 ? 256 : 512;  // the kernels' launch bounds (mmz_kernels.cuh: LaunchCfg)
  int force_tpb = 0;                        // development aid: MMZ_TPB pins the block size
  if (const char* e = getenv("MMZ_TPB")) force_tpb = atoi(e);
  if (const char* e = getenv("MMZ_SYNC")) h->bsync = (unsigned)atoi(e) & 7u;
  for (int tpb = max_tpb; tpb >= 32; tpb -= 32) {
    if (tpb < G || (force_tpb && tpb != force_tpb)) continue;
    int smem = model_smem + (tpb / G) * h->L.stride * 4;
    if (smem > dev_smem) continue;
    for (int mode = 0; mode < 5; mode++)
      CUDA_TRY(cudaFuncSetAttribute(h->fn[mode], cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int nblk = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nblk, h->fn[MODE_STEP], tpb, smem));
    int envs = nblk * (tpb / G);
    // a batch too small to give every SM a block of this size is better served by smaller blocks
    const bool spreads = (long)(h->n + tpb / G - 1) / (tpb / G) >= sms || tpb <= 64;
    if (envs > best_envs && (spreads || !best_tpb)) { best_envs = envs; best_tpb = tpb; best_smem = smem; }
    else if (best_tpb && !best_spreads && spreads) { best_envs = envs; best_tpb = tpb; best_smem = smem; }
    if (best_tpb == tpb) best_spreads = spreads;
  }
  if (!best_tpb) return fail(MMZ_ERR_CAPACITY, "workspace of %d bytes per environment does not fit in shared memory", h->L.stride * 4);
  h->tpb = best_tpb;
  h->smem_bytes = best_smem;
  h->envs_per_sm = best_envs;
  for (int mode = 0; mode < 5; mode++)
    CUDA_TRY(cudaFuncSetAttribute(h->fn[mode], cudaFuncAttributeMaxDynamicSharedMemorySize, best_smem));
  return MMZ_OK;
}

}  // namespace

extern "C" {

int mmz_abi_version(void) { return 7; }
const char* mmz_last_error(void) { return g_err; }

int mmz_create(const void* model_blob, size_t bytes, int num_envs, int device, uint32_t flags, mmz_handle* out) {
  if (!model_blob || !out) return fail(MMZ_ERR_INVALID, "null argument");
  if (num_envs < 1) return fail(MMZ_ERR_INVALID, "num_envs must be >= 1");
  if (bytes != sizeof(mmz_model)) return fail(MMZ_ERR_MODEL, "blob is %zu bytes, expected %zu", bytes, sizeof(mmz_model));
  mmz_env* h = new (std::nothrow) mmz_env();
  if (!h) return fail(MMZ_ERR_INVALID, "out of host memory");
  memcpy(&h->hm, model_blob, sizeof(mmz_model));
  int rc = validate(h->hm);
  if (rc != MMZ_OK) { delete h; return rc; }
  h->device = device;
  h->n = num_envs;
  h->flags = flags;
  auto bail = [&](int code) { mmz_destroy(h); return code; };
  if (cudaSetDevice(device) != cudaSuccess) return bail(fail(MMZ_ERR_CUDA, "cudaSetDevice(%d) failed", device));
  const int nv = h->hm.nv;
  const char* force_g = getenv("MMZ_FORCE_G");  // development aid: "G,NVP" pins a lanes-per-environment instance
  int fg = 0, fnvp = 0;
  if (force_g && sscanf(force_g, "%d,%d", &fg, &fnvp) == 2 && fg > 0) rc = configure(h, fg, fnvp);
  else if (configure_h(h, &rc)) rc = MMZ_OK;
  else if (rc != MMZ_OK) return bail(rc);
  else if (nv <= 4 && h->hm.nbody <= 8 && h->hm.ngeom <= 8) rc = configure(h, 8, 4);
  else if (nv <= 8) rc = configure(h, 8, 8);
  else if (nv <= 14 && nbox_geoms(h->hm) == 0 && h->hm.density <= 0.f && h->hm.viscosity <= 0.f)
    rc = configure(h, 16, 14);  // the Ant family without movable blocks: rows of exactly 14 registers
  else if (nv <= 16) rc = configure(h, 16, 16);
  else rc = configure(h, 32, 20);
  if (rc != MMZ_OK) return bail(rc);
  {
    // whole blocks and whole warps only: padding environments run like real ones (uniform control
    // flow, block-wide barriers inside the step) and write no outputs
    const int epb = h->use_t ? TE : h->tpb / h->G;
    int unit = epb;
    while (unit % 32) unit += epb;  // lcm(epb, 32)
    h->npad = round_up(num_envs, unit);
  }
  // device copy of the constants: model + derived tables, padded to a multiple of 16 bytes
  {
    unsigned char* host = new (std::nothrow) unsigned char[h->L.model_bytes];
    if (!host) return bail(fail(MMZ_ERR_INVALID, "out of host memory"));
    memset(host, 0, h->L.model_bytes);
    memcpy(host, &h->hm, sizeof(mmz_model));
    if (h->use_t) {
      TDerived dv;
      make_tderived(h->hm, &dv);
      memcpy(host + round_up((int)sizeof(mmz_model), 16), &dv, sizeof dv);
    } else {
      Derived dv;
      make_derived(h->hm, &dv);
      memcpy(host + round_up((int)sizeof(mmz_model), 16), &dv, sizeof dv);
    }
    cudaError_t e = cudaMalloc(&h->d_model, h->L.model_bytes);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_model, host, h->L.model_bytes, cudaMemcpyHostToDevice);
    delete[] host;
    if (e != cudaSuccess) return bail(fail(MMZ_ERR_CUDA, "model upload failed: %s", cudaGetErrorString(e)));
  }
  size_t state_bytes = (size_t)h->L.nstate * h->npad * sizeof(float);
  cudaError_t e = cudaMalloc(&h->d_state, state_bytes);
  if (e == cudaSuccess) e = cudaMemset(h->d_state, 0, state_bytes);
  if (e == cudaSuccess) e = cudaMalloc(&h->d_counters, 2 * (size_t)h->npad * sizeof(int));
  if (e == cudaSuccess) e = cudaMemset(h->d_counters, 0, 2 * (size_t)h->npad * sizeof(int));
  if (e != cudaSuccess) return bail(fail(MMZ_ERR_CUDA, "state allocation failed: %s", cudaGetErrorString(e)));
  // start every environment at qpos0 with fresh derived arrays (a model before its first reset)
  KArgs A;
  memset(&A, 0, sizeof A);
  A.seed = 0;
  {
    mmz_model& m = h->hm;
    float* row = new (std::nothrow) float[h->npad];
    if (!row) return bail(fail(MMZ_ERR_INVALID, "out of host memory"));
    for (int i = 0; i < m.nq; i++) {
      for (int k = 0; k < h->npad; k++) row[k] = m.qpos0[i];
      e = cudaMemcpy(h->d_state + (size_t)i * h->npad, row, h->npad * sizeof(float), cudaMemcpyHostToDevice);
      if (e != cudaSuccess) break;
    }
    delete[] row;
    if (e != cudaSuccess) return bail(fail(MMZ_ERR_CUDA, "state upload failed: %s", cudaGetErrorString(e)));
  }
  rc = launch(h, MODE_REFRESH, A, 0);
  if (rc != MMZ_OK) return bail(rc);
  if ((e = cudaStreamSynchronize(0)) != cudaSuccess)
    return bail(fail(MMZ_ERR_CUDA, "initial refresh failed: %s", cudaGetErrorString(e)));
  *out = h;
  return MMZ_OK;
}

int mmz_dims(mmz_handle h, int* num_envs, int* nq, int* nv, int* nu, int* obs_dim) {
  if (!h) return fail(MMZ_ERR_INVALID, "null handle");
  if (num_envs) *num_envs = h->n;
  if (nq) *nq = h->hm.nq;
  if (nv) *nv = h->hm.nv;
  if (nu) *nu = h->hm.nu;
  if (obs_dim) *obs_dim = h->hm.obs_dim;
  return MMZ_OK;
}

int mmz_kernel_config(mmz_handle h, int* lanes_per_env, int* threads_per_block, int* smem_bytes, int* envs_per_sm,
                      int* floats_per_env) {
  if (!h) return fail(MMZ_ERR_INVALID, "null handle");
  if (lanes_per_env) *lanes_per_env = h->G;
  if (threads_per_block) *threads_per_block = h->tpb;
  if (smem_bytes) *smem_bytes = h->smem_bytes;
  if (envs_per_sm) *envs_per_sm = h->envs_per_sm;
  if (floats_per_env) *floats_per_env = h->L.stride;
  return MMZ_OK;
}

const char* mmz_kernel_name(mmz_handle h) {
  if (!h) return "";
  if (h->kname[0] == 0) {
    if (h->use_t) snprintf(h->kname, sizeof h->kname, "maze_hkernel<%d,%d>", h->NVP, h->feat);
    else snprintf(h->kname, sizeof h->kname, "maze_kernel<%d,%d,%d>", h->G, h->NVP, h->feat);
  }
  return h->kname;
}

int mmz_set_step_diag(mmz_handle h, int32_t* d_diag) {
  if (!h) return fail(MMZ_ERR_INVALID, "null handle");
  h->d_step_diag = d_diag;
  return MMZ_OK;
}

int mmz_set_obs_peers(mmz_handle h, float* const* d_peer_obs, int npeers, int64_t row_offset, int multicast) {
  if (!h) return fail(MMZ_ERR_INVALID, "null handle");
  if (npeers < 0 || npeers > MMZ_MAXPEERS || (npeers > 0 && !d_peer_obs) || row_offset < 0 || (multicast && npeers != 1))
    return fail(MMZ_ERR_INVALID, "mmz_set_obs_peers: 0..%d buffers (exactly one for a multicast address), row_offset >= 0", MMZ_MAXPEERS);
  if (npeers > 0 && h->hm.view_dim)
    return fail(MMZ_ERR_INVALID, "mmz_set_obs_peers: tasks with TOP_DOWN_VIEW fill part of the observation in a second kernel and cannot be gathered in the step kernel");
  memset(&h->peers, 0, sizeof h->peers);
  for (int k = 0; k < npeers; k++) {
    if (!d_peer_obs[k]) return fail(MMZ_ERR_INVALID, "mmz_set_obs_peers: buffer %d is null", k);
    h->peers.buf[k] = d_peer_obs[k];
  }
  h->peers.n = npeers; h->peers.row0 = row_offset; h->peers.multicast = multicast ? 1 : 0;
  return MMZ_OK;
}

int mmz_set_env_offset(mmz_handle h, int first_global_env) {
  if (!h) return fail(MMZ_ERR_INVALID, "null handle");
  h->env_offset = first_global_env;
  return MMZ_OK;
}

int mmz_reset(mmz_handle h, const uint8_t* d_mask, uint64_t seed, float* d_obs, void* stream) {
  if (!h) return fail(MMZ_ERR_INVALID, "null handle");
  CUDA_TRY(cudaSetDevice(h->device));
  KArgs A;
  memset(&A, 0, sizeof A);
  A.mask = d_mask;
  A.seed = seed;
  A.obs = d_obs;
  h->seed = seed;
  return launch(h, MODE_RESET, A, (cudaStream_t)stream);
}

int mmz_step(mmz_handle h, const float* d_action, float* d_obs, float* d_reward, uint8_t* d_done, float* d_info,
             void* stream) {
  if (!h) return fail(MMZ_ERR_INVALID, "null handle");
  if (!d_action || !d_obs || !d_reward || !d_done) return fail(MMZ_ERR_INVALID, "null device pointer");
  CUDA_TRY(cudaSetDevice(h->device));
  KArgs A;
  memset(&A, 0, sizeof A);
  A.action = d_action; A.obs = d_obs; A.reward = d_reward; A.done = d_done; A.info = d_info;
  A.seed = h->seed;
  A.diag = h->d_step_diag;
  return launch(h, MODE_STEP, A, (cudaStream_t)stream);
}

int mmz_step_k(mmz_handle h, int K, const float* d_actions, float* d_obs, float* d_reward, uint8_t* d_done, float* d_info,
               void* stream) {
  if (!h) return fail(MMZ_ERR_INVALID, "null handle");
  if (K < 1 || !d_actions || !d_obs || !d_reward || !d_done) return fail(MMZ_ERR_INVALID, "mmz_step_k: K >= 1 and non-null device pointers");
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t s = (cudaStream_t)stream;
  // The K launches are recorded once into a CUDA graph per (K, buffers, flags) and replayed: one driver call per
  // rollout segment instead of K (and K view launches for TOP_DOWN_VIEW tasks).
  StepGraphKey key;
  memset(&key, 0, sizeof key);  // (padding bytes take part in the comparison)
  key.K = K; key.actions = d_actions; key.obs = d_obs; key.reward = d_reward; key.done = d_done; key.info = d_info;
  key.diag = h->d_step_diag; key.flags = h->flags; key.env_offset = h->env_offset; key.seed = h->seed;
  key.peer0 = h->peers.n ? h->peers.buf[0] : nullptr; key.peer_row0 = h->peers.row0;
  if (!h->kgraph || memcmp(&key, &h->kgraph_key, sizeof key) != 0) {
    if (h->kgraph) { cudaGraphExecDestroy(h->kgraph); h->kgraph = nullptr; }
    cudaStream_t cs = nullptr;
    CUDA_TRY(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
    cudaGraph_t g = nullptr;
    int rc = MMZ_OK;
    const uint64_t launches0 = h->launches;
    cudaError_t e = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
    if (e == cudaSuccess) {
      const size_t n = h->n;
      for (int k = 0; k < K && rc == MMZ_OK; k++) {
        KArgs A;
        memset(&A, 0, sizeof A);
        A.action = d_actions + (size_t)k * n * h->hm.nu; A.obs = d_obs + (size_t)k * n * h->hm.obs_dim;
        A.reward = d_reward + (size_t)k * n; A.done = d_done + (size_t)k * n; A.info = d_info ? d_info + (size_t)k * n * 4 : nullptr;
        A.seed = h->seed;
        A.diag = h->d_step_diag;
        rc = launch(h, MODE_STEP, A, cs);
      }
      e = cudaStreamEndCapture(cs, &g);
    }
    h->launches_per_kgraph = h->launches - launches0;
    h->launches = launches0;  // recorded, not run
    if (e == cudaSuccess && rc == MMZ_OK) e = cudaGraphInstantiate(&h->kgraph, g, 0);
    if (g) cudaGraphDestroy(g);
    cudaStreamDestroy(cs);
    if (rc != MMZ_OK) return rc;
    if (e != cudaSuccess) { h->kgraph = nullptr; return fail(MMZ_ERR_CUDA, "mmz_step_k: graph capture failed: %s", cudaGetErrorString(e)); }
    h->kgraph_key = key;
  }
  CUDA_TRY(cudaGraphLaunch(h->kgraph, s));
  h->launches += h->launches_per_kgraph;
  return MMZ_OK;
}

int mmz_step_host(mmz_handle h, const float* h_action, float* h_obs, float* h_reward, uint8_t* h_done, float* h_info,
                  void* stream) {
  if (!h) return fail(MMZ_ERR_INVALID, "null handle");
  if (!h_action || !h_obs || !h_reward || !h_done) return fail(MMZ_ERR_INVALID, "null host pointer");
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t s = (cudaStream_t)stream;
  const size_t n = h->n;
  {
    // Pinned (page-locked) host buffers are mapped into the device's address space (unified addressing): the step kernel
    // reads the actions and writes observations / reward / done / info straight over PCIe - each block as it finishes,
    // under the physics of the blocks still running. No staging copies, no launch split: ONE launch, then the
    // synchronisation the host contract asks for. Pageable buffers (or MMZ_HOST_ZERO_COPY=0) take the staged path below.
    const char* zc = getenv("MMZ_HOST_ZERO_COPY");
    const bool zero_copy = !(zc && zc[0] == '0');
    auto mapped = [](const void* p, void** out) {
      cudaPointerAttributes at;
      if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
      if (at.type != cudaMemoryTypeHost || at.devicePointer == nullptr) return false;
      *out = at.devicePointer;
      return true;
    };
    void *da = nullptr, *dob = nullptr, *dr = nullptr, *dd = nullptr, *di = nullptr;
    if (zero_copy && mapped(h_action, &da) && mapped(h_obs, &dob) && mapped(h_reward, &dr) && mapped(h_done, &dd) &&
        (!h_info || mapped(h_info, &di))) {
      int rc = mmz_step(h, (const float*)da, (float*)dob, (float*)dr, (uint8_t*)dd, (float*)di, stream);
      if (rc != MMZ_OK) return rc;
      CUDA_TRY(cudaStreamSynchronize(s));
      return MMZ_OK;
    }
  }
  if (!h->d_done) {  // d_done is allocated last: a staging set that failed half-way is rebuilt
    cudaFree(h->d_action); cudaFree(h->d_obs); cudaFree(h->d_reward); cudaFree(h->d_info);
    h->d_action = h->d_obs = h->d_reward = h->d_info = nullptr;
    CUDA_TRY(cudaMalloc(&h->d_action, n * (h->hm.nu > 0 ? h->hm.nu : 1) * sizeof(float)));
    CUDA_TRY(cudaMalloc(&h->d_obs, n * h->hm.obs_dim * sizeof(float)));
    CUDA_TRY(cudaMalloc(&h->d_reward, n * sizeof(float)));
    CUDA_TRY(cudaMalloc(&h->d_info, n * 4 * sizeof(float)));
    CUDA_TRY(cudaMalloc(&h->d_done, n));
  }
  // Large batches run as up to 4 block ranges on internal streams: the action upload and the result download of one
  // range overlap the kernel of another. Ranges are whole blocks, every environment is stepped exactly once, and the
  // result is identical to the single launch (environments are independent).
  const int epb = h->use_t ? TE : h->tpb / h->G;
  const int nblk = (h->npad + epb - 1) / epb;
  const int nchunk = (h->hm.view_dim || nblk < 64) ? 1 : 4;
  if (nchunk > 1) {
    for (int c = 0; c < 4; c++)
      if (!h->hs[c]) CUDA_TRY(cudaStreamCreateWithFlags(&h->hs[c], cudaStreamNonBlocking));
    for (int c = 0; c < 5; c++)
      if (!h->hev[c]) CUDA_TRY(cudaEventCreateWithFlags(&h->hev[c], cudaEventDisableTiming));
    const size_t nu = h->hm.nu, od = h->hm.obs_dim;
    CUDA_TRY(cudaEventRecord(h->hev[4], s));  // work already queued on the caller's stream comes first
    for (int c = 0; c < nchunk; c++) {
      const int b0 = (int)((long long)nblk * c / nchunk), b1 = (int)((long long)nblk * (c + 1) / nchunk);
      const size_t e0 = (size_t)b0 * epb, e1 = std::min((size_t)b1 * epb, n);
      cudaStream_t cs = h->hs[c];
      CUDA_TRY(cudaStreamWaitEvent(cs, h->hev[4], 0));
      if (e1 > e0 && nu)
        CUDA_TRY(cudaMemcpyAsync(h->d_action + e0 * nu, h_action + e0 * nu, (e1 - e0) * nu * sizeof(float), cudaMemcpyHostToDevice, cs));
      KArgs A;
      memset(&A, 0, sizeof A);
      A.action = h->d_action; A.obs = h->d_obs; A.reward = h->d_reward; A.done = h->d_done; A.info = h_info ? h->d_info : nullptr;
      A.seed = h->seed;
      A.diag = h->d_step_diag;
      int rc = launch(h, MODE_STEP, A, cs, b0, b1 - b0);
      if (rc != MMZ_OK) return rc;
      if (e1 > e0) {
        CUDA_TRY(cudaMemcpyAsync(h_obs + e0 * od, h->d_obs + e0 * od, (e1 - e0) * od * sizeof(float), cudaMemcpyDeviceToHost, cs));
        CUDA_TRY(cudaMemcpyAsync(h_reward + e0, h->d_reward + e0, (e1 - e0) * sizeof(float), cudaMemcpyDeviceToHost, cs));
        CUDA_TRY(cudaMemcpyAsync(h_done + e0, h->d_done + e0, e1 - e0, cudaMemcpyDeviceToHost, cs));
        if (h_info) CUDA_TRY(cudaMemcpyAsync(h_info + e0 * 4, h->d_info + e0 * 4, (e1 - e0) * 4 * sizeof(float), cudaMemcpyDeviceToHost, cs));
      }
      CUDA_TRY(cudaEventRecord(h->hev[c], cs));
      CUDA_TRY(cudaStreamWaitEvent(s, h->hev[c], 0));
    }
    CUDA_TRY(cudaStreamSynchronize(s));
    return MMZ_OK;
  }
  CUDA_TRY(cudaMemcpyAsync(h->d_action, h_action, n * h->hm.nu * sizeof(float), cudaMemcpyHostToDevice, s));
  int rc = mmz_step(h, h->d_action, h->d_obs, h->d_reward, h->d_done, h_info ? h->d_info : nullptr, stream);
  if (rc != MMZ_OK) return rc;
  CUDA_TRY(cudaMemcpyAsync(h_obs, h->d_obs, n * h->hm.obs_dim * sizeof(float), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(h_reward, h->d_reward, n * sizeof(float), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(h_done, h->d_done, n, cudaMemcpyDeviceToHost, s));
  if (h_info) CUDA_TRY(cudaMemcpyAsync(h_info, h->d_info, n * 4 * sizeof(float), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  return MMZ_OK;
}

int mmz_observe(mmz_handle h, float* d_obs, void* stream) {
  if (!h || !d_obs) return fail(MMZ_ERR_INVALID, "null argument");
  CUDA_TRY(cudaSetDevice(h->device));
  KArgs A;
  memset(&A, 0, sizeof A);
  A.obs = d_obs;
  return launch(h, MODE_OBSERVE, A, (cudaStream_t)stream);
}

int mmz_render(mmz_handle h, int first_env, int count, int width, int height, uint8_t* d_rgb, void* stream) {
  if (!h || !d_rgb) return fail(MMZ_ERR_INVALID, "null argument");
  if (first_env < 0 || count < 1 || first_env + count > h->n) return fail(MMZ_ERR_INVALID, "environment range out of bounds");
  if (width < 1 || height < 1 || width > 4096 || height > 4096) return fail(MMZ_ERR_INVALID, "image size must be 1..4096");
  CUDA_TRY(cudaSetDevice(h->device));
  const mmz_model& m = h->hm;
  RenderArgs R;
  R.model = (const mmz_model*)h->d_model; R.state = h->d_state; R.rgb = d_rgb;
  R.npad = h->npad; R.first_env = first_env; R.count = count; R.width = width; R.height = height;
  R.x0 = -m.origin[0] - m.cell_size; R.x1 = (m.grid_w - 1) * m.cell_size - m.origin[0] + m.cell_size;
  R.y0 = -m.origin[1] - m.cell_size; R.y1 = (m.grid_h - 1) * m.cell_size - m.origin[1] + m.cell_size;
  maze_render_kernel<<<count, 256, 0, (cudaStream_t)stream>>>(R);
  h->launches++;
  CUDA_TRY(cudaGetLastError());
  return MMZ_OK;
}

int mmz_get_state(mmz_handle h, int layout, float* d_qpos, float* d_qvel, int32_t* d_t, void* stream) {
  if (!h) return fail(MMZ_ERR_INVALID, "null handle");
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t s = (cudaStream_t)stream;
  const int n = h->n, npad = h->npad, nq = h->hm.nq, nv = h->hm.nv;
  const float* rq = h->d_state;
  const float* rv = h->d_state + (size_t)nq * npad;
  if (layout == MMZ_LAYOUT_ENV_MAJOR) {
    if (d_qpos) { rows_to_env_major<<<(n * nq + 255) / 256, 256, 0, s>>>(rq, d_qpos, n, npad, nq); h->launches++; }
    if (d_qvel) { rows_to_env_major<<<(n * nv + 255) / 256, 256, 0, s>>>(rv, d_qvel, n, npad, nv); h->launches++; }
  } else if (layout == MMZ_LAYOUT_SOA) {
    if (d_qpos) { rows_copy<<<(n * nq + 255) / 256, 256, 0, s>>>(rq, npad, d_qpos, n, n, nq); h->launches++; }
    if (d_qvel) { rows_copy<<<(n * nv + 255) / 256, 256, 0, s>>>(rv, npad, d_qvel, n, n, nv); h->launches++; }
  } else {
    return fail(MMZ_ERR_INVALID, "unknown layout %d", layout);
  }
  if (d_t) CUDA_TRY(cudaMemcpyAsync(d_t, h->d_counters, n * sizeof(int), cudaMemcpyDeviceToDevice, s));
  CUDA_TRY(cudaGetLastError());
  return MMZ_OK;
}

int mmz_set_state(mmz_handle h, int layout, const float* d_qpos, const float* d_qvel, const int32_t* d_t,
                  void* stream) {
  if (!h) return fail(MMZ_ERR_INVALID, "null handle");
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t s = (cudaStream_t)stream;
  const int n = h->n, npad = h->npad, nq = h->hm.nq, nv = h->hm.nv;
  float* rq = h->d_state;
  float* rv = h->d_state + (size_t)nq * npad;
  if (layout == MMZ_LAYOUT_ENV_MAJOR) {
    if (d_qpos) { env_major_to_rows<<<(n * nq + 255) / 256, 256, 0, s>>>(d_qpos, rq, n, npad, nq); h->launches++; }
    if (d_qvel) { env_major_to_rows<<<(n * nv + 255) / 256, 256, 0, s>>>(d_qvel, rv, n, npad, nv); h->launches++; }
  } else if (layout == MMZ_LAYOUT_SOA) {
    if (d_qpos) { rows_copy<<<(n * nq + 255) / 256, 256, 0, s>>>(d_qpos, n, rq, npad, n, nq); h->launches++; }
    if (d_qvel) { rows_copy<<<(n * nv + 255) / 256, 256, 0, s>>>(d_qvel, n, rv, npad, n, nv); h->launches++; }
  } else {
    return fail(MMZ_ERR_INVALID, "unknown layout %d", layout);
  }
  if (d_t) CUDA_TRY(cudaMemcpyAsync(h->d_counters, d_t, n * sizeof(int), cudaMemcpyDeviceToDevice, s));
  CUDA_TRY(cudaGetLastError());
  KArgs A;  // MujocoEnv.set_state -> mj_forward: refresh the derived arrays
  memset(&A, 0, sizeof A);
  return launch(h, MODE_REFRESH, A, s);
}

int mmz_forward(mmz_handle h, const float* d_action, float* d_qacc, int32_t* d_diag, void* stream) {
  if (!h || !d_action || !d_qacc) return fail(MMZ_ERR_INVALID, "null argument");
  CUDA_TRY(cudaSetDevice(h->device));
  KArgs A;
  memset(&A, 0, sizeof A);
  A.action = d_action; A.qacc_out = d_qacc; A.diag = d_diag;
  return launch(h, MODE_FORWARD, A, (cudaStream_t)stream);
}

uint64_t mmz_launch_count(mmz_handle h) { return h ? h->launches : 0; }

void mmz_destroy(mmz_handle h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  cudaFree(h->d_model); cudaFree(h->d_state); cudaFree(h->d_counters);
  cudaFree(h->d_action); cudaFree(h->d_obs); cudaFree(h->d_reward); cudaFree(h->d_info); cudaFree(h->d_done);
  if (h->kgraph) cudaGraphExecDestroy(h->kgraph);
  for (cudaStream_t st : h->hs) if (st) cudaStreamDestroy(st);
  for (cudaEvent_t ev : h->hev) if (ev) cudaEventDestroy(ev);
  delete h;
}

}  // extern "C"
