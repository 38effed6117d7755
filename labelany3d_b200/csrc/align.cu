// Arithmetic of the depth stage's RANSAC scale alignment on sm_100a ("next" row f3, second half):
// align_depth, src/batch_scripts/depth.py:52-92 of the reference = scikit-learn's
// RANSACRegressor(LinearRegression(fit_intercept=False), min_samples=0.2) on (relative depth, metric depth)
// pairs of the valid pixels.  The random subsets are drawn on the host exactly as scikit-learn draws them
// (same generator, same calls); these kernels do what follows each draw:
//   subset_fit:  least-squares slope through the origin of the subset = sum(x y) / sum(x x)  (float64 sums);
//   classify:    residual |y - x * coef| in float32 like NumPy, inlier iff residual <= threshold, and over the
//                inliers the count and the float64 sums the score and the final refit need;
//   scale_fill:  the aligned map: coef * relative depth under the mask, 10000.0 elsewhere.
// One CTA of 1024 threads per call, strided loads, fixed reduction tree: the same bits on every run.  (The sizes
// are one image: at most a few hundred thousand pairs; a trial is two ~20 us launches.)
#include "common.cuh"

namespace la3d {
namespace {

constexpr int kThreads = 1024;
constexpr int kWarps = kThreads / 32;
constexpr unsigned kFull = 0xffffffffu;

template <int N>
__device__ __forceinline__ void block_sum(double (&v)[N], double* out) {
  __shared__ double red[kWarps][N];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < N; ++k)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(kFull, v[k], o);
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < N; ++k) red[warp][k] = v[k];
  __syncthreads();
  if (threadIdx.x < N) {
    double acc = 0.0;
    for (int w = 0; w < kWarps; ++w) acc += red[w][threadIdx.x];
    out[threadIdx.x] = acc;
  }
}

__global__ void __launch_bounds__(kThreads) subset_fit_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                              const int64_t* __restrict__ idx, long long m,
                                                              double* __restrict__ out) {
  double s[2] = {0.0, 0.0};
  for (long long k = threadIdx.x; k < m; k += kThreads) {
    const long long i = idx[k];
    const double xv = (double)__ldg(x + i), yv = (double)__ldg(y + i);
    s[0] += xv * xv;
    s[1] += xv * yv;
  }
  block_sum(s, out);           // out[0] = sum x^2, out[1] = sum x y
}

__global__ void __launch_bounds__(kThreads) classify_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                            long long n, float coef, float threshold,
                                                            double* __restrict__ out) {
  // out: n_inliers, sum y, sum y^2, sum (y - y_pred)^2, sum x^2, sum x y   (all over the inliers)
  double s[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  for (long long i = threadIdx.x; i < n; i += kThreads) {
    const float xv = __ldg(x + i), yv = __ldg(y + i);
    const float diff = __fsub_rn(yv, __fmul_rn(xv, coef));          // float32, no FMA: y - X @ coef as NumPy evaluates it
    if (fabsf(diff) <= threshold) {
      const double xd = (double)xv, yd = (double)yv, dd = (double)diff;
      s[0] += 1.0; s[1] += yd; s[2] += yd * yd; s[3] += dd * dd; s[4] += xd * xd; s[5] += xd * yd;
    }
  }
  block_sum(s, out);
}

__global__ void scale_fill_kernel(const float* __restrict__ rel, const uint8_t* __restrict__ mask, long long n, float coef,
                                  float fill, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float r = rel[i];
  const bool on = mask ? mask[i] != 0 : !isinf(r);
  out[i] = on ? __fmul_rn(r, coef) : fill;
}

}  // namespace
}  // namespace la3d

extern "C" int la3d_ransac_subset_fit(const float* x, const float* y, const int64_t* idx, long long m, double* sums,
                                      la3d_stream_t stream) {
  using namespace la3d;
  LA3D_REQUIRE(x && y && idx && sums, "null pointer");
  LA3D_REQUIRE(m > 0, "empty subset");
  subset_fit_kernel<<<1, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(x, y, idx, m, sums);
  LA3D_CUDA(cudaGetLastError());
  return LA3D_OK;
}

extern "C" int la3d_ransac_classify(const float* x, const float* y, long long n, float coef, float threshold,
                                    double* stats, la3d_stream_t stream) {
  using namespace la3d;
  LA3D_REQUIRE(x && y && stats, "null pointer");
  LA3D_REQUIRE(n > 0, "no samples");
  classify_kernel<<<1, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(x, y, n, coef, threshold, stats);
  LA3D_CUDA(cudaGetLastError());
  return LA3D_OK;
}

extern "C" int la3d_scale_fill(const float* rel, const uint8_t* mask, long long n, float coef, float fill, float* out,
                               la3d_stream_t stream) {
  using namespace la3d;
  LA3D_REQUIRE(rel && out, "null pointer");
  LA3D_REQUIRE(n > 0, "no pixels");
  scale_fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(rel, mask, n, coef, fill, out);
  LA3D_CUDA(cudaGetLastError());
  return LA3D_OK;
}
