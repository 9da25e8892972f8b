// mmz_math.cuh - small fp32 vector / quaternion / spatial-algebra helpers (device).
//
// Spatial convention: everything is expressed in world axes about one REFERENCE POINT near the robot (the
// callers subtract it: translation invariant, and it avoids fp32 cancellation far from the maze origin).
// Motion vectors are [angular(3); linear velocity of the point at the reference(3)], force vectors are
// [torque about the reference(3); force(3)]. A spatial inertia is 10 floats:
// I[0..5] rotational about the reference (xx,yy,zz,xy,xz,yz), I[6..8] = m*com, I[9] = m.
#pragma once
#include <cuda_runtime.h>

namespace mmz {

#define MMZ_DI __device__ __forceinline__
constexpr float kMinVal = 1e-15f;
constexpr float kMaxVal = 1e10f;
constexpr float kPi = 3.14159265358979323846f;
// The line search of the Newton solver stops at |slope| < MMZ_LS_TOL * |slope at 0|: MuJoCo's own default, ls_tolerance =
// 0.01. It only has to find a good point along the direction: the precision of the solution is set by the Newton stopping
// rules. At 1e-6 (fp32 round-off, round 1) a third of the searches bounced until the step stopped changing; 1e-3 made the
// Ant step 7 % faster, 1e-2 another 0.6 %, both with the Newton iteration counts (39.4 per Ant env-step) and every error
// column against the oracle unchanged (profiles/r2_parity.md).
#ifndef MMZ_LS_TOL
#define MMZ_LS_TOL 1e-2f
#endif

MMZ_DI void cross3(float* r, const float* a, const float* b) {
  float x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
MMZ_DI float dot3(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
MMZ_DI float norm3(const float* a) { return sqrtf(dot3(a, a)); }
MMZ_DI float dot6(const float* a, const float* b) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3] + a[4] * b[4] + a[5] * b[5];
}
MMZ_DI void quat_mul(float* r, const float* a, const float* b) {
  float w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  float x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  float y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  float z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
MMZ_DI void quat_norm(float* q) {
  float n = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < kMinVal) { q[0] = 1.f; q[1] = q[2] = q[3] = 0.f; return; }
  float inv = 1.f / n;
  q[0] *= inv; q[1] *= inv; q[2] *= inv; q[3] *= inv;
}
MMZ_DI void quat2mat(float* R, const float* q) {  // row-major 3x3
  float w = q[0], x = q[1], y = q[2], z = q[3];
  R[0] = 1.f - 2.f * (y * y + z * z); R[1] = 2.f * (x * y - w * z); R[2] = 2.f * (x * z + w * y);
  R[3] = 2.f * (x * y + w * z); R[4] = 1.f - 2.f * (x * x + z * z); R[5] = 2.f * (y * z - w * x);
  R[6] = 2.f * (x * z - w * y); R[7] = 2.f * (y * z + w * x); R[8] = 1.f - 2.f * (x * x + y * y);
}
MMZ_DI void axisangle2quat(float* q, const float* axis, float ang) {
  // Half joint angles and half integration angles are small (|x| < pi): the SFU sine / cosine (abs. error ~5e-7
  // there) is as good as fp32 needs and sits on the critical path of every tree walk; measured +1.5 % on the Ant
  // step with unchanged errors against the oracle (profiles/r1_parity.md).
  float s, c;
  __sincosf(0.5f * ang, &s, &c);
  q[0] = c; q[1] = s * axis[0]; q[2] = s * axis[1]; q[3] = s * axis[2];
}
MMZ_DI void mat_vec(float* r, const float* R, const float* v) {  // r = R v
  float x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
  float y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
  float z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
MMZ_DI void matT_vec(float* r, const float* R, const float* v) {  // r = R^T v
  float x = R[0] * v[0] + R[3] * v[1] + R[6] * v[2];
  float y = R[1] * v[0] + R[4] * v[1] + R[7] * v[2];
  float z = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}

// f = I v (spatial inertia times motion vector)
MMZ_DI void inert_mul(float* f, const float* I, const float* v) {
  const float *w = v, *l = v + 3, *h = I + 6;
  float hxl[3], hxw[3];
  cross3(hxl, h, l);
  cross3(hxw, h, w);
  f[0] = I[0] * w[0] + I[3] * w[1] + I[4] * w[2] + hxl[0];
  f[1] = I[3] * w[0] + I[1] * w[1] + I[5] * w[2] + hxl[1];
  f[2] = I[4] * w[0] + I[5] * w[1] + I[2] * w[2] + hxl[2];
  f[3] = I[9] * l[0] - hxw[0];
  f[4] = I[9] * l[1] - hxw[1];
  f[5] = I[9] * l[2] - hxw[2];
}
MMZ_DI void cross_motion(float* r, const float* v, const float* s) {  // v x s
  float a[3], b[3], c[3];
  cross3(a, v, s);
  cross3(b, v, s + 3);
  cross3(c, v + 3, s);
  r[0] = a[0]; r[1] = a[1]; r[2] = a[2];
  r[3] = b[0] + c[0]; r[4] = b[1] + c[1]; r[5] = b[2] + c[2];
}
MMZ_DI void cross_force(float* r, const float* v, const float* f) {  // v x* f
  float a[3], b[3], c[3];
  cross3(a, v, f);
  cross3(b, v + 3, f + 3);
  cross3(c, v, f + 3);
  r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2];
  r[3] = c[0]; r[4] = c[1]; r[5] = c[2];
}

// Philox4x32-10 counter-based generator (Salmon et al. 2011), used for reset noise.
MMZ_DI void philox4x32(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
MMZ_DI float u01(uint32_t x) { return (x >> 8) * (1.0f / 16777216.0f); }  // [0, 1)

}  // namespace mmz
