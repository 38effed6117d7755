// Shared declarations of libla3d_sm100a (B200 / sm_100a only).
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "la3d.h"

namespace la3d {

// A mask plane is cut into chunks of 512 pixels = 16 bit-words = what one warp
// converts per step (32 lanes x 16 bytes).
constexpr int kChunkPx = 512;
constexpr int kChunkWords = 16;

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t err, const char* what);

#define LA3D_CUDA(expr)                                        \
  do {                                                         \
    cudaError_t err__ = (expr);                                \
    if (err__ != cudaSuccess) return cuda_fail(err__, #expr);  \
  } while (0)

#define LA3D_REQUIRE(cond, msg)          \
  do {                                   \
    if (!(cond)) {                       \
      set_error("%s: %s", __func__, msg); \
      return LA3D_EINVAL;                \
    }                                    \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---------------------------------------------------------------------------
// Camera maths shared by the lift and the fit kernels.
// ---------------------------------------------------------------------------

// Inverse of a 3x3 by LU with partial pivoting, forward substitution, and back
// substitution that multiplies by the reciprocal of the pivot - the operation
// order of LAPACK dgesv as OpenBLAS runs it, so that for pinhole intrinsics the
// result is bit-identical to np.linalg.inv (src/util.py:56 of the reference).
// No fused multiply-add anywhere.
__device__ inline void invert3x3(const double* __restrict__ Kin, double* __restrict__ X) {
  double A[3][3];
  int perm[3] = {0, 1, 2};
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) A[i][j] = Kin[i * 3 + j];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    int p = k;
    double best = fabs(A[k][k]);
#pragma unroll
    for (int i = k + 1; i < 3; ++i) {
      double a = fabs(A[i][k]);
      if (a > best) { best = a; p = i; }
    }
    if (p != k) {
#pragma unroll
      for (int j = 0; j < 3; ++j) { double tmp = A[k][j]; A[k][j] = A[p][j]; A[p][j] = tmp; }
      int tp = perm[k]; perm[k] = perm[p]; perm[p] = tp;
    }
    double r = __ddiv_rn(1.0, A[k][k]);
#pragma unroll
    for (int i = k + 1; i < 3; ++i) {
      A[i][k] = __dmul_rn(A[i][k], r);
#pragma unroll
      for (int j = k + 1; j < 3; ++j) A[i][j] = __dsub_rn(A[i][j], __dmul_rn(A[i][k], A[k][j]));
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    double b[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) b[i] = (perm[i] == c) ? 1.0 : 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < i; ++j) b[i] = __dsub_rn(b[i], __dmul_rn(A[i][j], b[j]));
#pragma unroll
    for (int i = 2; i >= 0; --i) {
#pragma unroll
      for (int j = i + 1; j < 3; ++j) b[i] = __dsub_rn(b[i], __dmul_rn(A[i][j], b[j]));
      b[i] = __dmul_rn(b[i], __ddiv_rn(1.0, A[i][i]));
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) X[i * 3 + c] = b[i];
  }
}

// One pixel of depth_to_points (src/util.py:72) in the reference's operation
// order: ((d*Kinv[i][0])*u + (d*Kinv[i][1])*v) + (d*Kinv[i][2])*1, float64, no FMA.
__device__ __forceinline__ void lift_pixel_exact(double d, double u, double v, const double* __restrict__ Kinv,
                                                 double& x, double& y, double& z) {
  x = __dadd_rn(__dadd_rn(__dmul_rn(__dmul_rn(d, Kinv[0]), u), __dmul_rn(__dmul_rn(d, Kinv[1]), v)),
                __dmul_rn(d, Kinv[2]));
  y = __dadd_rn(__dadd_rn(__dmul_rn(__dmul_rn(d, Kinv[3]), u), __dmul_rn(__dmul_rn(d, Kinv[4]), v)),
                __dmul_rn(d, Kinv[5]));
  z = __dadd_rn(__dadd_rn(__dmul_rn(__dmul_rn(d, Kinv[6]), u), __dmul_rn(__dmul_rn(d, Kinv[7]), v)),
                __dmul_rn(d, Kinv[8]));
}

// R @ p + t (src/util.py:74), summed left to right.
__device__ __forceinline__ void rigid_exact(const double* __restrict__ R, const double* __restrict__ t, double& x,
                                            double& y, double& z) {
  double a = x, b = y, c = z;
  if (R) {
    x = __dadd_rn(__dadd_rn(__dmul_rn(R[0], a), __dmul_rn(R[1], b)), __dmul_rn(R[2], c));
    y = __dadd_rn(__dadd_rn(__dmul_rn(R[3], a), __dmul_rn(R[4], b)), __dmul_rn(R[5], c));
    z = __dadd_rn(__dadd_rn(__dmul_rn(R[6], a), __dmul_rn(R[7], b)), __dmul_rn(R[8], c));
  }
  if (t) {
    x = __dadd_rn(x, t[0]);
    y = __dadd_rn(y, t[1]);
    z = __dadd_rn(z, t[2]);
  }
}

}  // namespace la3d
