// mmz_clamp.cuh - the manual wall collision of the Point robot (reference maze_env.py:450-464 on top of
// CollisionDetector.detect, maze_env_utils.py:186-206), shared by both step kernels. Scalar code: every thread that calls
// it computes the result for its own environment.
#pragma once
#include "mmz_layout.h"
#include "mmz_math.cuh"

namespace mmz {

// CollisionDetector.detect: the wall segment hit first on the way o -> n; `point` the intersection, `refl` the mirror image
// of n on that wall
MMZ_DI bool seg_detect(const mmz_model* m, const float* o, const float* n, float* point, float* refl) {
  float mvx = n[0] - o[0], mvy = n[1] - o[1];
  if (sqrtf(mvx * mvx + mvy * mvy) <= 1e-8f) return false;
  bool hit = false;
  float bestd = 0.f;
  for (int s = 0; s < m->nseg; s++) {
    float x1 = m->seg[s][0], y1 = m->seg[s][1], x2 = m->seg[s][2], y2 = m->seg[s][3];
    float wx = x2 - x1, wy = y2 - y1;
    float sa = (wx * (o[1] - y1) - wy * (o[0] - x1)) * (wx * (n[1] - y1) - wy * (n[0] - x1));
    float sb = (mvx * (y1 - o[1]) - mvy * (x1 - o[0])) * (mvx * (y2 - o[1]) - mvy * (x2 - o[0]));
    if (!(sa <= 0.f && sb <= 0.f)) continue;
    float den = wx * mvy - wy * mvx, num = wx * (y2 - o[1]) - wy * (x2 - o[0]);
    if (den == 0.f) continue;  // collinear: the reference raises ZeroDivisionError; treated as no hit
    float tq = num / den, px = o[0] + tq * mvx, py = o[1] + tq * mvy;
    float d = sqrtf((px - o[0]) * (px - o[0]) + (py - o[1]) * (py - o[1]));
    if (!hit || d < bestd) {
      float tt = ((n[0] - x1) * wx + (n[1] - y1) * wy) / (wx * wx + wy * wy);
      float fx = x1 + tt * wx, fy = y1 + tt * wy;
      hit = true; bestd = d;
      point[0] = px; point[1] = py;
      refl[0] = fx + (fx - n[0]); refl[1] = fy + (fy - n[1]);
    }
  }
  return hit;
}

// detect + the bounce of maze_env.py:457-464: true if the move old -> nw crossed a wall segment; pos = bounce position,
// or `old` when the bounce crosses a wall again
MMZ_DI bool clamp_move(const mmz_model* m, const float* old, const float* nw, float* pos) {
  bool hit = false;
  float target[2] = {nw[0], nw[1]};
#pragma unroll 1
  for (int pass = 0; pass < 2; pass++) {
    float pt[2], rf[2];
    bool h = seg_detect(m, old, target, pt, rf);
    if (pass == 0) {
      if (!h) return false;
      hit = true;
      target[0] = pt[0] + m->restitution * (rf[0] - pt[0]);
      target[1] = pt[1] + m->restitution * (rf[1] - pt[1]);
    } else if (h) {
      target[0] = old[0]; target[1] = old[1];
    }
  }
  pos[0] = target[0]; pos[1] = target[1];
  return hit;
}

}  // namespace mmz
