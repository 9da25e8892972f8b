// mmz_layout.h - per-environment shared-memory workspace layout and persisted-state layout.
//
// One group of G lanes (G = 8, 16 or 32; 32 = one warp per environment) owns one
// environment. Its working set lives in shared memory at `stride` floats per environment;
// every array below is an offset (in floats) into that block. The layout is computed on the
// host from the model dimensions (mmz_api.cu: make_layout) and passed to the kernels by
// value (constant bank).
//
// Persisted state in HBM is structure-of-arrays, one row per scalar, N (padded) floats per
// row: rows [0,nq) qpos, [nq,nq+nv) qvel, [nq+nv,nq+2nv) qacc (solver warm start),
// then 3*nobj rows of observed-body positions (the reference's stale data.xpos),
// plus one int32 row for the episode step counter t.
#pragma once
#include <stdint.h>

#include "../../include/mmz_model.h"

namespace mmz {

// per-contact block (floats). The contact Jacobian is never stored: in the solver lane d re-derives
// its own column (J_n, J_t1, J_t2)[d] from the contact point, frame and its motion axis cdof[d].
// Slots 16..25 hold the narrow-phase parameters until make_constraints turns them into the row data.
enum {
  C_POS = 0,     // [3] contact point (world; relative to the origin of body 0 after make_constraints)
  C_FRAME = 3,   // [9] rows: normal, tangent1, tangent2
  C_BODY1 = 12,  // body ids (-1 world) | after make_constraints: dof masks with sign +1 / -1
  C_BODY2 = 13,
  C_MPOS = 12,
  C_MNEG = 13,
  C_MU = 14,
  C_D = 15,      // D shared by the 4 pyramid edges (after make_constraints)
  C_AREF = 16,   // [4]
  C_JAR = 20,    // [4] J a - aref per edge
  C_JV = 24,     // [4] J dir per edge
  // narrow-phase temporaries (dead once the rows are built)
  C_DIST = 16,
  C_MARGIN = 17,
  C_SOLREF = 18,  // [2]
  C_SOLIMP = 20,  // [5]
  C_INVW = 25,
  C_STRIDE = 29,  // odd
};
// per-env integer counters (stored in the float workspace)
enum { N_CON = 0, N_LIM = 1, N_ITER = 2, N_OVERFLOW = 3,
       N_ITER_SUM = 4, N_LS_SUM = 5, N_CON_MAX = 6, N_CAPPED = 7, N_CNT = 8 };  // 4..7: accumulated over one env-step

constexpr int MMZ_MAXPAIR = 32;
constexpr int MMZ_MAXPEERS = 8;

// Fused observation gather (mmz_set_obs_peers): the step kernels store every observation row they write a second time
// at row `row0 + env` of each peer buffer - [total_envs][obs_dim] float32 in the other ranks' (peer-mapped) and this
// rank's memory - or ONCE through an NVLS multicast address that fans out to all of them (`multicast`). The NVLink
// transfers overlap the physics of the blocks that are still running: no collective runs after the kernel.
struct ObsPeers {
  float* buf[MMZ_MAXPEERS];
  long long row0;
  int n;          // 0 = off
  int multicast;  // buf[0] is a multicast address: multimem.st
};
#ifdef __CUDACC__
__device__ __forceinline__ void peer_store(const ObsPeers& P, size_t idx, float v) {
  if (P.multicast) {
    asm volatile("multimem.st.global.f32 [%0], %1;" ::"l"(P.buf[0] + idx), "f"(v) : "memory");
  } else {
#pragma unroll 1
    for (int k = 0; k < P.n; k++) P.buf[k][idx] = v;
  }
}
#endif

struct Derived {            // appended to the model blob in device memory
  int32_t anc[MMZ_MAXBODY];  // bit a set in anc[b]: body a is b or an ancestor of b
  int32_t boxg[MMZ_MAXGEOM]; // geoms of type box on moving bodies
  int32_t nboxg;
  int32_t nlev;              // number of tree levels
  int32_t npair;             // sphere-sphere / sphere-capsule pairs between moving bodies that pass the contact filter
  int32_t pad[1];
  int32_t pair_a[MMZ_MAXPAIR];  // the sphere (geom1 of the pair)
  int32_t pair_b[MMZ_MAXPAIR];  // the other sphere or the capsule (geom2)
  float ident[9];            // identity rotation (static maze boxes)
  float padf[3];
};

struct Layout {
  int nb, nj, nv, nq, nu, ng, nobj, obs_dim;
  int nlatch;    // latched body positions: nobj observed + nviewb for the top-down view
  int obs_core;  // obs_dim without the top-down view: what the step kernel assembles (the view kernel fills the rest)
  int ldm;      // row stride of M and H (odd)
  int maxcon;   // contact capacity per environment
  int cstride;  // floats per contact block (odd)
  int nstate;   // persisted float rows
  int stride;   // floats per environment (stride % 32 == G % 32: groups of a warp hit distinct banks)
  int model_bytes;  // bytes of model + Derived in device memory (multiple of 16)
  int o_qpos, o_qvel, o_ctrl, o_q0, o_v0, o_xv, o_fa, o_accv, o_acca;
  int o_xpos, o_xquat, o_xmat, o_xipos, o_ximat, o_xanchor, o_xaxis, o_gpos, o_gmat, o_cdof;
  int o_iw, o_ic, o_vel, o_acc, o_frc, o_fsub;
  int o_M, o_smooth, o_qacc, o_dir;
  int o_con, o_cnt, o_objpos, o_obs;
};

}  // namespace mmz
