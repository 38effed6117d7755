// Depth-scale alignment by masked median on sm_100a ("next" row f3 of the scope table): the step
// just before the box fit, align_to_depth_match (src/util.py:464-494 of the reference):
//     overlap = mask & render_mask                           (:473)
//     ratios  = depth_map[overlap] / depth_render[overlap]   (:480-485)
//     scale   = np.median(ratios)                            (:486)
// per (image, instance), straight from the bit planes of la3d_mask_scan: no compaction, no sort.
// One CTA per plane runs an exact radix select (4 passes of 8 bits over the order-preserving
// integer image of the float32 ratios); an even count averages the two middle values the way
// np.median does for float32 (float32 add, then halve); any NaN ratio makes the median NaN
// (np.median's NaN check).  Bit-exact with NumPy on float32 inputs.
#include <math_constants.h>

#include "common.cuh"

namespace la3d {
namespace {

constexpr int kThreads = 256;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ uint32_t order_key(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_value(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

struct Plane {
  const uint32_t* a;
  const uint32_t* b;
  const float* num;
  const float* den;
  int words;
};

// Histogram of the digit at `shift` over the ratios whose key matches `prefix` on the bits above it.
// kFirst: also counts the overlap pixels and the NaN ratios.
template <bool kFirst>
__device__ __forceinline__ void histogram(const Plane& pl, uint32_t prefix, int shift, int* hist, int* n_total,
                                          int* n_nan) {
  int cnt = 0, nan = 0;
  for (int w = threadIdx.x; w < pl.words; w += kThreads) {
    uint32_t m = __ldg(pl.a + w) & __ldg(pl.b + w);
    while (m) {
      const int bit = __ffs(m) - 1;
      m &= m - 1;
      const int p = (w << 5) + bit;
      const float r = __fdiv_rn(__ldg(pl.num + p), __ldg(pl.den + p));
      if (kFirst) ++cnt;
      if (r != r) { if (kFirst) ++nan; continue; }
      const uint32_t key = order_key(r);
      if (shift == 24 || (key >> (shift + 8)) == (prefix >> (shift + 8))) atomicAdd(&hist[(key >> shift) & 0xffu], 1);
    }
  }
  if (kFirst) {
    cnt = __reduce_add_sync(kFull, cnt);
    nan = __reduce_add_sync(kFull, nan);
    if ((threadIdx.x & 31) == 0) { atomicAdd(n_total, cnt); atomicAdd(n_nan, nan); }
  }
}

__global__ void __launch_bounds__(kThreads) ratio_median_kernel(const float* __restrict__ depth_map,
                                                                const float* __restrict__ depth_render,
                                                                const uint32_t* __restrict__ bits_a,
                                                                const uint32_t* __restrict__ bits_b, int group, int HW,
                                                                int words, int32_t* __restrict__ n_overlap,
                                                                float* __restrict__ scale) {
  __shared__ int hist[256];
  __shared__ int s_total, s_nan, s_digit, s_below;
  const int plane = blockIdx.x, tid = threadIdx.x;
  Plane pl{bits_a + (size_t)plane * words, bits_b + (size_t)plane * words, depth_map + (size_t)(plane / group) * HW,
           depth_render + (size_t)plane * HW, words};
  float picked[2] = {0.f, 0.f};
  int n = 0, n_nan = 0;
  // selection 0 finds sorted[n/2]; selection 1 (even n only) finds sorted[n/2 - 1]
  for (int sel = 0; sel < 2; ++sel) {
    uint32_t prefix = 0;
    int k = 0;
    for (int shift = 24; shift >= 0; shift -= 8) {
      hist[tid] = 0;
      if (tid == 0 && sel == 0 && shift == 24) { s_total = 0; s_nan = 0; }
      __syncthreads();
      if (sel == 0 && shift == 24) histogram<true>(pl, prefix, shift, hist, &s_total, &s_nan);
      else histogram<false>(pl, prefix, shift, hist, nullptr, nullptr);
      __syncthreads();
      if (sel == 0 && shift == 24) {
        n = s_total; n_nan = s_nan;
        k = n / 2;
      } else if (shift == 24) {
        k = n / 2 - 1;
      }
      if (n == 0 || n_nan > 0) break;            // uniform
      if (tid == 0) {
        int below = 0, d = 0;
        for (; d < 255; ++d) {
          if (below + hist[d] > k) break;
          below += hist[d];
        }
        s_digit = d; s_below = below;
      }
      __syncthreads();
      prefix |= (uint32_t)s_digit << shift;
      k -= s_below;
      __syncthreads();
    }
    if (n == 0 || n_nan > 0) break;
    picked[sel] = key_value(prefix);
    if (n & 1) break;
  }
  if (tid == 0) {
    n_overlap[plane] = n;
    float out;
    if (n == 0 || n_nan > 0) out = CUDART_NAN_F;
    else if (n & 1) out = picked[0];
    else out = __fdiv_rn(__fadd_rn(picked[1], picked[0]), 2.0f);
    scale[plane] = out;
  }
}

}  // namespace
}  // namespace la3d

extern "C" int la3d_masked_ratio_median(const float* depth_map, const float* depth_render, const uint32_t* mask_bits,
                                        const uint32_t* render_bits, int planes, int group, int H, int W,
                                        int32_t* n_overlap, float* scale, la3d_stream_t stream) {
  using namespace la3d;
  LA3D_REQUIRE(depth_map && depth_render && mask_bits && render_bits && n_overlap && scale, "null pointer");
  LA3D_REQUIRE(planes > 0 && group > 0 && H > 0 && W > 0, "non-positive shape");
  LA3D_REQUIRE((long long)H * W < (1ll << 30), "image too large");
  ratio_median_kernel<<<(unsigned)planes, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      depth_map, depth_render, mask_bits, render_bits, group, H * W, (int)la3d_words_per_plane(H, W), n_overlap, scale);
  LA3D_CUDA(cudaGetLastError());
  return LA3D_OK;
}
