// mmz_narrow.cuh - narrow-phase contact generation (device, fp32, one thread per geom pair).
//
// Replaces the narrow phase of MuJoCo's mj_collision [third party; reached by the reference
// through mj_step, call sites ant.py:63, point.py:59] for the geom types the reference's
// assets use: plane / sphere / capsule / box (assets/point.xml:21-22, ant.xml:22-66, the
// maze boxes of maze_env.py:138-152 and movable blocks :563-660).
// Convention: the normal points from geom1 to geom2, dist < 0 when penetrating, the contact
// position is midway between the two surfaces.
#pragma once
#include "mmz_math.cuh"

namespace mmz {

struct RawContact {
  float dist, pos[3], normal[3], hint[3];
};

// sphere (centre c, radius r) against a box (centre bc, rotation bR row-major, half extents h)
MMZ_DI int sphere_box(const float* c, float r, const float* bc, const float* bR, const float* h, float margin,
                      RawContact* out) {
  float rel[3] = {c[0] - bc[0], c[1] - bc[1], c[2] - bc[2]}, loc[3], dl[3];
  matT_vec(loc, bR, rel);
  bool inside = true;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    float cl = fminf(fmaxf(loc[k], -h[k]), h[k]);
    if (cl != loc[k]) inside = false;
    dl[k] = cl - loc[k];
  }
  float nl[3] = {0.f, 0.f, 0.f}, pl[3], dist;
  if (inside) {  // centre inside the box: leave through the nearest face
    float best = 2.f * (h[0] + h[1] + h[2]);
    int bi = 0;
#pragma unroll
    for (int i = 0; i < 6; i++) {
      float cd = fabsf(((i & 1) ? 1.f : -1.f) * h[i >> 1] - loc[i >> 1]);
      if (cd < best) { best = cd; bi = i; }
    }
#pragma unroll
    for (int k = 0; k < 3; k++) nl[k] = (k == (bi >> 1)) ? ((bi & 1) ? -1.f : 1.f) : 0.f;
    dist = -best - r;
#pragma unroll
    for (int k = 0; k < 3; k++) pl[k] = loc[k] + nl[k] * (r - best) * 0.5f;
  } else {
    float d = norm3(dl);
    if (d - r > margin) return 0;
    float inv = 1.f / d;
#pragma unroll
    for (int k = 0; k < 3; k++) nl[k] = dl[k] * inv;
    dist = d - r;
#pragma unroll
    for (int k = 0; k < 3; k++) pl[k] = loc[k] + nl[k] * (r + dist * 0.5f);
  }
  if (dist > margin) return 0;
  out->dist = dist;
  mat_vec(out->normal, bR, nl);
  mat_vec(out->pos, bR, pl);
#pragma unroll
  for (int k = 0; k < 3; k++) { out->pos[k] += bc[k]; out->hint[k] = 0.f; }
  return 1;
}

// sphere against sphere: normal from sphere 1 to sphere 2, contact midway between the surfaces. The capsule case is
// the sphere against the capsule-radius sphere at the segment point nearest to it (segment_nearest).
MMZ_DI int sphere_sphere(const float* c1, float r1, const float* c2, float r2, float margin, RawContact* out) {
  const float d[3] = {c2[0] - c1[0], c2[1] - c1[1], c2[2] - c1[2]};
  const float len = norm3(d), dist = len - r1 - r2;
  if (dist > margin) return 0;
  float nrm[3] = {0.f, 0.f, 1.f};
  if (len > kMinVal) { const float inv = 1.f / len; nrm[0] = d[0] * inv; nrm[1] = d[1] * inv; nrm[2] = d[2] * inv; }
  out->dist = dist;
#pragma unroll
  for (int k = 0; k < 3; k++) { out->normal[k] = nrm[k]; out->pos[k] = c1[k] + nrm[k] * (r1 + 0.5f * dist); out->hint[k] = 0.f; }
  return 1;
}
MMZ_DI void segment_nearest(const float* p0, const float* p1, const float* c, float* q) {
  const float ab[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]}, ac[3] = {c[0] - p0[0], c[1] - p0[1], c[2] - p0[2]};
  const float den = dot3(ab, ab);
  float t = den > kMinVal ? dot3(ac, ab) / den : 0.f;
  t = fminf(fmaxf(t, 0.f), 1.f);
#pragma unroll
  for (int k = 0; k < 3; k++) q[k] = p0[k] + t * ab[k];
}

MMZ_DI float box_excess_deriv(const float* a, const float* b, float t, const float* h) {
  float g = 0.f;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    float x = a[k] + b[k] * t;
    float ex = x > h[k] ? x - h[k] : (x < -h[k] ? x + h[k] : 0.f);
    g += ex * b[k];
  }
  return g;
}

// Capsule (segment p0-p1) against a box: parameter t in [0,1] of the segment point nearest the box,
// the root of the monotone piecewise-linear derivative of the squared distance along the segment.
// The caller probes both end caps first (two contacts when both are within the margin) and falls
// back to a single sphere_box at this point (mmz_dyn.cuh: collision).
MMZ_DI float capsule_nearest(const float* p0, const float* p1, const float* bc, const float* bR, const float* h) {
  float a[3], b[3], rel[3];
#pragma unroll
  for (int k = 0; k < 3; k++) rel[k] = p0[k] - bc[k];
  matT_vec(a, bR, rel);
#pragma unroll
  for (int k = 0; k < 3; k++) rel[k] = p1[k] - p0[k];
  matT_vec(b, bR, rel);
  float glo = box_excess_deriv(a, b, 0.f, h), ghi = box_excess_deriv(a, b, 1.f, h);
  if (glo > 0.f) return 0.f;
  if (ghi <= 0.f) return 1.f;
  float tlo = 0.f, thi = 1.f;
#pragma unroll 1
  for (int c = 0; c < 6; c++) {
    const int k = c >> 1;
    const float bk = (k == 0) ? b[0] : (k == 1 ? b[1] : b[2]);
    const float ak = (k == 0) ? a[0] : (k == 1 ? a[1] : a[2]);
    const float hk = (k == 0) ? h[0] : (k == 1 ? h[1] : h[2]);
    if (fabsf(bk) > kMinVal) {
      float t = (((c & 1) ? hk : -hk) - ak) / bk;
      if (t > 0.f && t < 1.f) {
        float g = box_excess_deriv(a, b, t, h);
        if (g <= 0.f) { if (t > tlo) { tlo = t; glo = g; } }
        else if (t < thi) { thi = t; ghi = g; }
      }
    }
  }
  return (ghi - glo) > kMinVal ? tlo + (-glo) * (thi - tlo) / (ghi - glo) : tlo;
}

// box against box: separating-axis search, then face clipping or an edge-edge point.
// Normal from box A to box B; up to 8 contacts. This is the longest dependent chain of the small robots' step (the Point's
// arrow box point.xml:22 against the maze boxes, movable blocks), so everything but the clipped polygon lives in REGISTERS:
// the 15-axis search is fully unrolled (static row indices), rows picked at run time go through selects, the polygon
// ping-pongs between two local arrays (one load per vertex and side). Local memory is slow here: with ~200 KB of shared
// memory per block the SM has almost no L1 left and every local access is an L2 round trip.
MMZ_DI void bb_row(const float (&M)[9], int i, float* r) {
#pragma unroll
  for (int k = 0; k < 3; k++) r[k] = i == 0 ? M[k] : (i == 1 ? M[3 + k] : M[6 + k]);
}
MMZ_DI float bb_sel(const float* h, int i) { return i == 0 ? h[0] : (i == 1 ? h[1] : h[2]); }
// one Sutherland-Hodgman pass: keep sg * (axs . (x - cr)) <= lim. Emits p if inside, then the crossing of (p, next)
MMZ_DI int bb_clip(const float (*src)[3], int np, float (*dst)[3], const float* axs, float sg, float lim, const float* cr) {
  constexpr int CAP = 12;
  int nn = 0;
  float p[3] = {src[0][0], src[0][1], src[0][2]};
  float dp = sg * ((p[0] - cr[0]) * axs[0] + (p[1] - cr[1]) * axs[1] + (p[2] - cr[2]) * axs[2]) - lim;
#pragma unroll 1
  for (int i = 0; i < np; i++) {
    const int in = i + 1 == np ? 0 : i + 1;
    const float q[3] = {src[in][0], src[in][1], src[in][2]};
    const float dq = sg * ((q[0] - cr[0]) * axs[0] + (q[1] - cr[1]) * axs[1] + (q[2] - cr[2]) * axs[2]) - lim;
    if (dp <= 0.f && nn < CAP) { dst[nn][0] = p[0]; dst[nn][1] = p[1]; dst[nn][2] = p[2]; nn++; }
    if ((dp <= 0.f) != (dq <= 0.f) && nn < CAP) {
      const float t = dp / (dp - dq);
#pragma unroll
      for (int k = 0; k < 3; k++) dst[nn][k] = p[k] + t * (q[k] - p[k]);
      nn++;
    }
    p[0] = q[0]; p[1] = q[1]; p[2] = q[2];
    dp = dq;
  }
  return nn;
}
static __device__ __noinline__ int box_box(const float* ca, const float* Ra, const float* ha, const float* cb,
                                    const float* Rb, const float* hb, float margin, RawContact* out) {
  float A[9], B[9], d[3];  // row i = axis i of the box in world coordinates
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int k = 0; k < 3; k++) { A[3 * i + k] = Ra[3 * k + i]; B[3 * i + k] = Rb[3 * k + i]; }
#pragma unroll
  for (int k = 0; k < 3; k++) d[k] = cb[k] - ca[k];
  const float hav[3] = {ha[0], ha[1], ha[2]}, hbv[3] = {hb[0], hb[1], hb[2]};
  float best_f = -3.0e38f, best_e = -3.0e38f, nf[3] = {0, 0, 0}, ne[3] = {0, 0, 0};
  int code_f = -1, code_e = -1;
#pragma unroll
  for (int code = 0; code < 15; code++) {
    float L[3];
    if (code < 3) { L[0] = A[3 * code]; L[1] = A[3 * code + 1]; L[2] = A[3 * code + 2]; }
    else if (code < 6) { L[0] = B[3 * (code - 3)]; L[1] = B[3 * (code - 3) + 1]; L[2] = B[3 * (code - 3) + 2]; }
    else {
      cross3(L, A + 3 * ((code - 6) / 3), B + 3 * ((code - 6) % 3));
      float n = norm3(L);
      if (n < 1e-6f) continue;  // parallel edges: covered by the face axes
      float inv = 1.f / n;
      L[0] *= inv; L[1] *= inv; L[2] *= inv;
    }
    float ra = 0.f, rb = 0.f;
#pragma unroll
    for (int i = 0; i < 3; i++) { ra += hav[i] * fabsf(dot3(L, A + 3 * i)); rb += hbv[i] * fabsf(dot3(L, B + 3 * i)); }
    float dl = dot3(L, d);
    float s = fabsf(dl) - ra - rb;
    if (s > margin) return 0;
    float sg = dl < 0.f ? -1.f : 1.f;
    if (code < 6) {
      if (s > best_f) { best_f = s; code_f = code; nf[0] = sg * L[0]; nf[1] = sg * L[1]; nf[2] = sg * L[2]; }
    } else if (s > best_e) { best_e = s; code_e = code; ne[0] = sg * L[0]; ne[1] = sg * L[1]; ne[2] = sg * L[2]; }
  }
  float best, bn[3];
  int bcode;
  if (code_e >= 0 && best_e > best_f + 1e-6f + 0.05f * fabsf(best_f)) { best = best_e; bcode = code_e; bn[0] = ne[0]; bn[1] = ne[1]; bn[2] = ne[2]; }
  else { best = best_f; bcode = code_f; bn[0] = nf[0]; bn[1] = nf[1]; bn[2] = nf[2]; }
  if (bcode < 0) return 0;
  if (bcode >= 6) {  // edge-edge
    const int ia = (bcode - 6) / 3, ib = (bcode - 6) % 3;
    float pa[3] = {ca[0], ca[1], ca[2]}, pb[3] = {cb[0], cb[1], cb[2]};
#pragma unroll
    for (int i = 0; i < 3; i++) {
      if (i != ia) {
        float sg = dot3(bn, A + 3 * i) > 0.f ? 1.f : -1.f;
#pragma unroll
        for (int k = 0; k < 3; k++) pa[k] += sg * hav[i] * A[3 * i + k];
      }
      if (i != ib) {
        float sg = dot3(bn, B + 3 * i) > 0.f ? -1.f : 1.f;
#pragma unroll
        for (int k = 0; k < 3; k++) pb[k] += sg * hbv[i] * B[3 * i + k];
      }
    }
    float Aia[3], Bib[3];
    bb_row(A, ia, Aia);
    bb_row(B, ib, Bib);
    const float hia = bb_sel(hav, ia), hib = bb_sel(hbv, ib);
    float w[3] = {pb[0] - pa[0], pb[1] - pa[1], pb[2] - pa[2]}, ua = 0.f, ub = 0.f;
    float uaub = dot3(Aia, Bib), q1 = dot3(Aia, w), q2 = -dot3(Bib, w), den = 1.f - uaub * uaub;
    if (den > 1e-9f) { ua = (q1 + uaub * q2) / den; ub = (uaub * q1 + q2) / den; }
    ua = fminf(fmaxf(ua, -hia), hia);
    ub = fminf(fmaxf(ub, -hib), hib);
#pragma unroll
    for (int k = 0; k < 3; k++) {
      float xa = pa[k] + ua * Aia[k], xb = pb[k] + ub * Bib[k];
      out->pos[k] = 0.5f * (xa + xb);
      out->normal[k] = bn[k];
      out->hint[k] = 0.f;
    }
    out->dist = best;
    return 1;
  }
  // face contact: the reference box owns the axis, the incident box is the other
  const bool refa = bcode < 3;
  const int ax = refa ? bcode : bcode - 3;
  float Rr[9], Ri[9], cr[3], ci[3], hr[3], hi[3], nr[3];
#pragma unroll
  for (int k = 0; k < 9; k++) { Rr[k] = refa ? A[k] : B[k]; Ri[k] = refa ? B[k] : A[k]; }
#pragma unroll
  for (int k = 0; k < 3; k++) {
    cr[k] = refa ? ca[k] : cb[k]; ci[k] = refa ? cb[k] : ca[k];
    hr[k] = refa ? hav[k] : hbv[k]; hi[k] = refa ? hbv[k] : hav[k];
    nr[k] = refa ? bn[k] : -bn[k];
  }
  int iax = 0;
  float mind = 3.0e38f;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    float v = fabsf(dot3(nr, Ri + 3 * i));
    if (-v < mind) { mind = -v; iax = i; }
  }
  float Rix[3], Riu[3], Riv[3], Rru[3], Rrv[3];
  const int u = (iax + 1) % 3, v = (iax + 2) % 3, ru = (ax + 1) % 3, rv = (ax + 2) % 3;
  bb_row(Ri, iax, Rix); bb_row(Ri, u, Riu); bb_row(Ri, v, Riv);
  bb_row(Rr, ru, Rru); bb_row(Rr, rv, Rrv);
  const float hix = bb_sel(hi, iax), hiu = bb_sel(hi, u), hiv = bb_sel(hi, v);
  const float hru = bb_sel(hr, ru), hrv = bb_sel(hr, rv), hrx = bb_sel(hr, ax);
  const float isg = dot3(nr, Rix) > 0.f ? -1.f : 1.f;
  float poly[12][3], tmp[12][3];
#pragma unroll
  for (int c = 0; c < 4; c++) {
    const float su = (c == 0 || c == 3) ? -1.f : 1.f, sv = (c < 2) ? -1.f : 1.f;
#pragma unroll
    for (int k = 0; k < 3; k++) poly[c][k] = ci[k] + isg * hix * Rix[k] + su * hiu * Riu[k] + sv * hiv * Riv[k];
  }
  int np = bb_clip(poly, 4, tmp, Rru, 1.f, hru, cr);
  if (np == 0) return 0;
  np = bb_clip(tmp, np, poly, Rru, -1.f, hru, cr);
  if (np == 0) return 0;
  np = bb_clip(poly, np, tmp, Rrv, 1.f, hrv, cr);
  if (np == 0) return 0;
  np = bb_clip(tmp, np, poly, Rrv, -1.f, hrv, cr);
  if (np == 0) return 0;
  int n = 0;
#pragma unroll 1
  for (int i = 0; i < np && n < 8; i++) {
    const float pp[3] = {poly[i][0], poly[i][1], poly[i][2]};
    const float depth = (pp[0] - cr[0]) * nr[0] + (pp[1] - cr[1]) * nr[1] + (pp[2] - cr[2]) * nr[2] - hrx;
    if (depth >= margin) continue;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      out[n].pos[k] = pp[k] - nr[k] * depth * 0.5f;
      out[n].normal[k] = bn[k];
      out[n].hint[k] = 0.f;
    }
    out[n].dist = depth;
    n++;
  }
  return n;
}

// orthonormal contact frame from a normal and an optional tangent hint (rows: n, t1, t2)
MMZ_DI void make_frame(float* fr) {
  float inv = 1.f / norm3(fr);
  fr[0] *= inv; fr[1] *= inv; fr[2] *= inv;
  if (norm3(fr + 3) < 0.5f) {
    fr[3] = fr[4] = fr[5] = 0.f;
    if (fr[1] < 0.5f && fr[1] > -0.5f) fr[4] = 1.f; else fr[5] = 1.f;
  }
  float d = dot3(fr, fr + 3);
  fr[3] -= d * fr[0]; fr[4] -= d * fr[1]; fr[5] -= d * fr[2];
  inv = 1.f / norm3(fr + 3);
  fr[3] *= inv; fr[4] *= inv; fr[5] *= inv;
  cross3(fr + 6, fr, fr + 3);
}

}  // namespace mmz
