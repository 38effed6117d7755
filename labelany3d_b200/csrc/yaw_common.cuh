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
  for (int k = 0; k < n; ++k) {
    const int idx = list ? list[k] : k;
    const double px = x[idx], pz = z[idx];
    const double rx = c * px - s * pz, rz = s * px + c * pz;
    mnx = dmin(mnx, rx); mxx = dmax(mxx, rx); mnz = dmin(mnz, rz); mxz = dmax(mxz, rz);
  }
  return (mxx - mnx) * (mxz - mnz);
}

__device__ __forceinline__ double sweep_angle(int c, int K) {
  return __ddiv_rn(__dmul_rn((double)c, CUDART_PIO2), (double)K);     // k * (pi/2) / K as NumPy evaluates it
}

}  // namespace la3d
