// mmz_view.cuh - MazeEnv.get_top_down_view (reference maze_env.py:262-349) for all environments.
//
// The reference splats every BLOCK cell (channel 0), every CHASM cell (channel 1) and every movable block
// (channel 2) into a 5x5 egocentric raster: a unit square around the continuous (row, col) of the source is
// distributed over the 3x3 neighbourhood of its integer cell by overlap area. The nine products it spells out
// (maze_env.py:284-322) are separable, so the kernel GATHERS instead: one thread per (environment, view entry)
// sums weight_row * weight_col over the sources. No task of the reference's registry turns the view on
// (MazeTask.TOP_DOWN_VIEW, maze_task.py:68), so this runs as a second small launch after the step kernel and
// only for models with view_dim != 0; the step kernel leaves the 75 columns before the trailing t untouched.
//
// Positions are the LATCHED body origins (rows nq + 2 nv + 3 k of the state: the reference reads data.xpos through
// get_body_com, which is only as fresh as the last kinematics pass - SURVEY Q15): latch nobj is the torso, the
// following ones are the movable blocks.
#pragma once
#include "mmz_layout.h"

namespace mmz {

// weight of view row/column `target` for a source at continuous row/column `pos` (maze_env.py:277-322)
__device__ __forceinline__ float view_axis_weight(float pos, int target) {
  const int cell = (int)pos;             // int(): truncation toward zero, also for negative positions (:277)
  const float f = pos - floorf(pos);     // Python's `% 1` (:277); the `< 0` branches at :278-281 never fire
  const int d = target - cell;
  if (d == 0) return fminf(1.f, f + 0.5f) - fmaxf(0.f, f - 0.5f);
  if (d == -1) return fmaxf(0.f, 0.5f - f);
  if (d == 1) return fmaxf(0.f, f - 0.5f);
  return 0.f;
}

struct ViewArgs {
  const mmz_model* model;
  const float* state;   // [nstate][npad]
  const uint8_t* mask;  // optional: only environments with a non-zero byte are written (masked reset)
  float* obs;           // [n][obs_dim]
  int n, npad;
};

__global__ void __launch_bounds__(256) maze_view_kernel(const __grid_constant__ ViewArgs A) {
  const mmz_model* __restrict__ m = A.model;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int env = idx / MMZ_VIEW_DIM, v = idx - env * MMZ_VIEW_DIM;
  if (env >= A.n) return;
  if (A.mask && !A.mask[env]) return;
  const int row = v / 15, col = (v / 3) % 5, ch = v % 3;  // view[row][col][channel] flattened (:354)
  const float s = m->cell_size;
  const int latch0 = m->nq + 2 * m->nv + 3 * m->nobj;
  const float rx = A.state[(size_t)latch0 * A.npad + env], ry = A.state[(size_t)(latch0 + 1) * A.npad + env];
  float acc = 0.f;
  if (ch < 2) {
    const int bit = ch == 0 ? MMZ_CELL_WALL : MMZ_CELL_CHASM;
    for (int i = 0; i < m->grid_h; i++) {
      const float y = (i * s - m->origin[1]) - ry;               // :333-334 then :270-271
      const float wr = view_axis_weight(2.f + (y + s / 2.f) / s, row);  // _xy_to_rowcol, :90-93
      if (wr == 0.f) continue;
      for (int j = 0; j < m->grid_w; j++) {
        if (!(m->grid[i * m->grid_w + j] & bit)) continue;
        const float x = (j * s - m->origin[0]) - rx;
        acc += wr * view_axis_weight(2.f + (x + s / 2.f) / s, col);
      }
    }
  } else {
    for (int k = 1; k < m->nviewb; k++) {  // movable blocks (:345-347)
      const int r0 = latch0 + 3 * k;
      const float x = A.state[(size_t)r0 * A.npad + env] - rx, y = A.state[(size_t)(r0 + 1) * A.npad + env] - ry;
      acc += view_axis_weight(2.f + (y + s / 2.f) / s, row) * view_axis_weight(2.f + (x + s / 2.f) / s, col);
    }
  }
  A.obs[(size_t)env * m->obs_dim + (m->obs_dim - 1 - MMZ_VIEW_DIM) + v] = acc;
}

}  // namespace mmz
