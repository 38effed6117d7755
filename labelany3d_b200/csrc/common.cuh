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

constexpr int kMtN = 624;      // MT19937 state words

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

__device__ __forceinline__ bool aligned16_dev(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---------------------------------------------------------------------------
// Programmatic dependent launch.  A kernel launched with launch_pdl(..., pdl = true) may start
// while its predecessor in the stream is still running, once every CTA of the predecessor has
// called pdl_trigger() (or exited); pdl_wait() then blocks until the predecessor has completed and
// its writes are visible.  Without the launch attribute both are no-ops.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, bool pdl,
                              Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---------------------------------------------------------------------------
// mbarrier + 1-D TMA (cp.async.bulk) helpers: global -> shared bulk copies that signal an mbarrier.
// Source, destination and size must be multiples of 16 bytes.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(bar);
  uint32_t done = 0;
  for (uint32_t spins = 0; !done; ++spins) {
    if (spins > (1u << 24)) __trap();            // a lost arrival must not hang the GPU
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar))
               : "memory");
}

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

// Rg of util_3dbox.py:128-134: Rodrigues rotation taking (0,-1,0) to the ground
// normal, which is flipped first when dot((0,-1,0), g) <= 0.  0/0 -> NaN when the
// two are parallel, exactly like the reference.  g == nullptr: identity.
__device__ inline void ground_rotation(const double* g, double* Rg) {
  if (!g) {
#pragma unroll
    for (int i = 0; i < 9; ++i) Rg[i] = (i % 4 == 0) ? 1.0 : 0.0;
    return;
  }
  double g0 = g[0], g1 = g[1], g2 = g[2];
  const double dotp = 0.0 * g0 + (-1.0) * g1 + 0.0 * g2;
  if (dotp <= 0.0) { g0 = -g0; g1 = -g1; g2 = -g2; }
  const double a0 = 0.0, a1 = -1.0, a2 = 0.0;
  const double nb = sqrt(g0 * g0 + g1 * g1 + g2 * g2);
  double b0 = g0, b1 = g1, b2 = g2;
  if (nb != 0.0) { b0 = g0 / nb; b1 = g1 / nb; b2 = g2 / nb; }
  const double ax = a1 * b2 - a2 * b1, ay = a2 * b0 - a0 * b2, az = a0 * b1 - a1 * b0;
  const double cosang = a0 * b0 + a1 * b1 + a2 * b2;
  const double S[9] = {0.0, -az, ay, az, 0.0, -ax, -ay, ax, 0.0};
  const double nrm = sqrt(ax * ax + ay * ay + az * az);
  const double nn = nrm * nrm;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const double s2 = S[i * 3 + 0] * S[0 * 3 + j] + S[i * 3 + 1] * S[1 * 3 + j] + S[i * 3 + 2] * S[2 * 3 + j];
      Rg[i * 3 + j] = ((i == j ? 1.0 : 0.0) + S[i * 3 + j]) + s2 * (1.0 - cosang) / nn;
    }
}

// ---------------------------------------------------------------------------
// "prep" buffer of the scanned-mask path (la3d_fit_prepare): everything a batch
// needs that does not depend on the masks - per image the intrinsics and their
// inverse and the pre-generated MT19937 words (+ the state to continue from),
// per box the ground rotation.
// ---------------------------------------------------------------------------
struct PrepCamera {
  double K[9];
  double Kinv[9];
};
struct PrepView {
  PrepCamera* cams;   // [B]
  double* Rg;         // [B*I][9]
  uint32_t* state;    // [B][624] generator state after the pre-generated words
  uint32_t* words;    // [B][nblk*624] tempered outputs
  int nblk;
  size_t bytes;
};
PrepView prep_view(void* base, int B, int I, int nblk);
int prep_blocks(int I);
// images [b0, b0 + B) of the batch the pointers describe
int launch_sample(const uint32_t* chunk_counts, const PrepView& pv, int B, int I, int chunks, int32_t* counts,
                  int32_t* ranks, cudaStream_t s, bool pdl = false, int b0 = 0);

}  // namespace la3d
