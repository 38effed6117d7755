// Small pieces shared by the two box kernels (fit.cu: at most 500 points per box, 64 threads; fit_all.cu:
// every masked pixel, 256 threads): float64 min / max with NumPy's NaN-skipping comparisons, the
// "first strict minimum" ordering of the reference's `if area < min_area` loop (src/util_3dbox.py:216),
// the gift-wrapping step of the convex hull, the rectangle area of a rotated footprint and the sweep angles.
#pragma once

#include <math_constants.h>

#include "common.cuh"

namespace la3d {

__device__ __forceinline__ double dmin(double a, double b) { return b < a ? b : a; }   // NaN in b is ignored
__device__ __forceinline__ double dmax(double a, double b) { return b > a ? b : a; }

// (area, index) pairs ordered like the reference's `if area < min_area` loop: the
// smallest area wins, the earliest index among equals; NaN and +inf never win.
struct Best {
  double area;
  int idx;   // -1 = nothing qualified yet
  __device__ void offer(double a, int i) {
    if (!(a < CUDART_INF) || i < 0) return;
    if (idx < 0 || a < area || (a == area && i < idx)) { area = a; idx = i; }
  }
};

// One step of the gift wrapping (counter-clockwise): the candidate q for the vertex after `cur`; a point p
// replaces q when it lies clockwise of cur->q, or on that ray and farther away (collinear points are skipped:
// a strict hull, like Qhull's vertex list).
struct Wrap {
  double cx, cz, qx, qz;
  int qi;
  __device__ void offer(double px, double pz, int pi) {
    if (pi < 0 || !(px == px) || (px == cx && pz == cz)) return;
    if (qi < 0) { qx = px; qz = pz; qi = pi; return; }
    const double cr = (qx - cx) * (pz - cz) - (qz - cz) * (px - cx);
    bool take = cr < 0.0;                       // p is clockwise of cur->q: q cannot be the next vertex
    if (cr == 0.0) {
      const double dq = (qx - cx) * (qx - cx) + (qz - cz) * (qz - cz);
      const double dp = (px - cx) * (px - cx) + (pz - cz) * (pz - cz);
      take = dp > dq || (dp == dq && pi < qi);
    }
    if (take) { qx = px; qz = pz; qi = pi; }
  }
};

// Bounding-rectangle area of a footprint given as an index list into x[] / z[] (a list that contains every point
// that can be extreme) after a rotation.  kind 0: the reference's hull-edge test rotates by +ang (rot_2d of
// util_3dbox.py:206-210, kept although the box is later built with rotate_y(yaw) = -yaw in XZ); kind 1: the sweep
// rotates like rotate_y(ang).  list == nullptr: points 0 .. n-1.
__device__ __forceinline__ double rect_area(const double* __restrict__ x, const double* __restrict__ z,
                                            const unsigned short* __restrict__ list, int n, double ang, int kind) {
  double s, c;
  sincos(ang, &s, &c);
  if (kind) s = -s;
  double mnx = CUDART_INF, mxx = -CUDART_INF, mnz = CUDART_INF, mxz = -CUDART_INF;
#pragma unroll 4
  for (int k = 0; k < n; ++k) {
    const int idx = list ? list[k] : k;
    const double px = x[idx], pz = z[idx];
    const double rx = c * px - s * pz, rz = s * px + c * pz;
    mnx = dmin(mnx, rx); mxx = dmax(mxx, rx); mnz = dmin(mnz, rz); mxz = dmax(mxz, rz);
  }
  return (mxx - mnx) * (mxz - mnz);
}

// float32 form of "strictly inside a convex polygon of n <= 8 vertices (counter-clockwise; consecutive duplicates
// allowed)", safe under rounding.  One thread per edge d < 8 writes pre[d] = {a, b, c}: a point p passes edge d when
// a * x' + b * z' + c > 0 with (x', z') = float(p - centre).  c carries a margin that covers the rounding of a, b, c,
// x', z' and of the two FMAs for every point within twice the polygon's radius of its centre, and points farther
// out fail some edge by much more than any rounding; so a point that passes ALL edges lies strictly inside the
// polygon (it cannot be extreme in any direction), while a point that fails is merely kept.  Degenerate edges
// (repeated vertex) and unused slots always pass.  The caller synchronises between writing the vertices / centre
// and calling this, and again before using pre[].
__device__ __forceinline__ void polygon_pretest_edge(const double* __restrict__ vx, const double* __restrict__ vz, int n,
                                                     double cx, double cz, int d, float (&out)[4]) {
  float a = 0.f, b = 0.f, c = 1.f;
  if (d < n) {
    const int e = (d + 1 == n) ? 0 : d + 1;
    const double ex = vx[e] - vx[d], ez = vz[e] - vz[d];
    if (ex != 0.0 || ez != 0.0) {
      double radius = 0.0;
      for (int i = 0; i < n; ++i) radius = fmax(radius, fmax(fabs(vx[i] - cx), fabs(vz[i] - cz)));
      const double ox = vx[d] - cx, oz = vz[d] - cz;
      a = (float)(-ez); b = (float)ex;
      const double c0 = -((double)a * ox + (double)b * oz);
      const double margin = ldexp((fabs((double)a) + fabs((double)b)) * 2.0 * radius, -19);
      c = (float)(c0 - margin);
      if (!(fabsf(a) < CUDART_INF_F) || !(fabsf(b) < CUDART_INF_F) || !(fabsf(c) < CUDART_INF_F)) { a = 0.f; b = 0.f; c = -1.f; }   // never passes
      else c = nextafterf(c, -CUDART_INF_F);                // the rounding of c itself
    }
  }
  out[0] = a; out[1] = b; out[2] = c; out[3] = 0.f;
}

__device__ __forceinline__ double sweep_angle(int c, int K) {
  return __ddiv_rn(__dmul_rn((double)c, CUDART_PIO2), (double)K);     // k * (pi/2) / K as NumPy evaluates it
}

}  // namespace la3d
