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

// per-contact block: scalars then the 3 x nv contact-frame Jacobian (normal, tangent1, tangent2)
enum {
  C_DIST = 0,   // narrow phase: dist | after row build: D (shared by the 4 pyramid edges)
  C_POS = 1,    // narrow phase: pos[3], frame[9] | after row build: aref[4] (1..4), jar[4] (5..8), jv[4] (9..12)
  C_FRAME = 4,
  C_AREF = 1,
  C_JAR = 5,
  C_JV = 9,
  C_BODY1 = 13,
  C_BODY2 = 14,
  C_MU = 15,
  C_MARGIN = 16,
  C_SOLREF = 17,
  C_SOLIMP = 19,
  C_INVW = 24,
  C_J = 25,
};
// per joint-limit row
enum { R_DOF = 0, R_SIGN = 1, R_D = 2, R_AREF = 3, R_JAR = 4, R_JV = 5, R_STRIDE = 7 };
// per-env integer counters (stored in the float workspace)
enum { N_CON = 0, N_LIM = 1, N_ITER = 2, N_OVERFLOW = 3,
       N_ITER_SUM = 4, N_LS_SUM = 5, N_CON_MAX = 6, N_CAPPED = 7, N_CNT = 8 };  // 4..7: accumulated over one env-step

struct Derived {            // appended to the model blob in device memory
  int32_t anc[MMZ_MAXBODY];  // bit a set in anc[b]: body a is b or an ancestor of b
  int32_t boxg[MMZ_MAXGEOM]; // geoms of type box on moving bodies
  int32_t nboxg;
  int32_t nlev;              // number of tree levels
  int32_t pad[2];
  float ident[9];            // identity rotation (static maze boxes)
  float padf[3];
};

struct Layout {
  int nb, nj, nv, nq, nu, ng, nobj, obs_dim;
  int ldm;      // row stride of M and H (odd)
  int maxcon;   // contact capacity per environment
  int maxlim;   // joint-limit row capacity
  int cstride;  // floats per contact block (odd)
  int nstate;   // persisted float rows
  int stride;   // floats per environment (stride % 32 == G % 32: groups of a warp hit distinct banks)
  int model_bytes;  // bytes of model + Derived in device memory (multiple of 16)
  int o_qpos, o_qvel, o_ctrl, o_q0, o_v0, o_xv, o_fa, o_accv, o_acca;
  int o_xpos, o_xquat, o_xmat, o_xipos, o_ximat, o_xanchor, o_xaxis, o_gpos, o_gmat, o_cdof;
  int o_iw, o_ic, o_vel, o_acc, o_frc, o_fsub;
  int o_M, o_H, o_bias, o_passive, o_smooth, o_qacc_smooth, o_qacc, o_grad, o_dir, o_tmp, o_col;
  int o_con, o_lim, o_cnt, o_objpos, o_obs;
};

}  // namespace mmz
