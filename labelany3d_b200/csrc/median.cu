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
// The first pass walks the two bit planes once (16-byte loads), gathers the two depths of every
// overlap pixel and parks the ratio keys in shared memory (up to 8192 of them);
// the remaining passes then run on that list.  Larger overlaps re-walk the bit planes per pass.
#include <math_constants.h>

#include "common.cuh"

namespace la3d {
namespace {

constexpr int kThreads = 256;
constexpr int kCap = 8192;             // ratio keys kept in shared memory (32 KB)
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ uint32_t order_key(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_value(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// Warp-aggregated shared-memory counters: the ratios of one object sit in a narrow band, so most
// lanes hit the same histogram bin; lanes with equal targets elect one to add for all of them.
__device__ __forceinline__ void hist_add(int* hist, uint32_t digit) {
  const unsigned act = __activemask();
  const unsigned peers = __match_any_sync(act, digit);
  if ((int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&hist[digit], __popc(peers));
}
__device__ __forceinline__ int take_slot(int* counter) {
  const unsigned act = __activemask();
  const int lane = threadIdx.x & 31, leader = __ffs(act) - 1;
  int base = 0;
  if (lane == leader) base = atomicAdd(counter, __popc(act));
  base = __shfl_sync(act, base, leader);
  return base + __popc(act & ((1u << lane) - 1u));
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
                                          int* n_nan, uint32_t* keys, int* n_keys) {
  int cnt = 0, nan = 0;
  const uint4* a4 = reinterpret_cast<const uint4*>(pl.a);
  const uint4* b4 = reinterpret_cast<const uint4*>(pl.b);
  for (int w4 = threadIdx.x; w4 < pl.words / 4; w4 += kThreads) {      // words is a multiple of 16
    const uint4 x = __ldg(a4 + w4), y = __ldg(b4 + w4);
    const uint32_t mm[4] = {x.x & y.x, x.y & y.y, x.z & y.z, x.w & y.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      uint32_t m = mm[k];
      while (m) {
        const int bit = __ffs(m) - 1;
        m &= m - 1;
        const int p = (((w4 << 2) + k) << 5) + bit;
        const float r = __fdiv_rn(__ldg(pl.num + p), __ldg(pl.den + p));
        if (kFirst) ++cnt;
        if (r != r) { if (kFirst) ++nan; continue; }
        const uint32_t key = order_key(r);
        if (kFirst) {
          const int slot = take_slot(n_keys);
          if (slot < kCap) keys[slot] = key;
        }
        if (shift == 24 || (key >> (shift + 8)) == (prefix >> (shift + 8))) hist_add(hist, (key >> shift) & 0xffu);
      }
    }
  }
  if (kFirst) {
    cnt = __reduce_add_sync(kFull, cnt);
    nan = __reduce_add_sync(kFull, nan);
    if ((threadIdx.x & 31) == 0) { atomicAdd(n_total, cnt); atomicAdd(n_nan, nan); }
  }
}

// the same over the key list parked in shared memory by the first pass
__device__ __forceinline__ void histogram_list(const uint32_t* keys, int n, uint32_t prefix, int shift, int* hist) {
  for (int k = threadIdx.x; k < n; k += kThreads) {
    const uint32_t key = keys[k];
    if (shift == 24 || (key >> (shift + 8)) == (prefix >> (shift + 8))) hist_add(hist, (key >> shift) & 0xffu);
  }
}

__global__ void __launch_bounds__(kThreads) ratio_median_kernel(const float* __restrict__ depth_map,
                                                                const float* __restrict__ depth_render,
                                                                const uint32_t* __restrict__ bits_a,
                                                                const uint32_t* __restrict__ bits_b, int group, int HW,
                                                                int words, int32_t* __restrict__ n_overlap,
                                                                float* __restrict__ scale) {
  extern __shared__ uint32_t keys[];                       // [kCap]
  __shared__ int hist[256];
  __shared__ int s_total, s_nan, s_digit, s_below, s_keys;
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
      if (tid == 0 && sel == 0 && shift == 24) { s_total = 0; s_nan = 0; s_keys = 0; }
      __syncthreads();
      if (sel == 0 && shift == 24) histogram<true>(pl, prefix, shift, hist, &s_total, &s_nan, keys, &s_keys);
      else if (s_keys <= kCap) histogram_list(keys, s_keys, prefix, shift, hist);
      else histogram<false>(pl, prefix, shift, hist, nullptr, nullptr, nullptr, nullptr);
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
  LA3D_REQUIRE(aligned16(mask_bits) && aligned16(render_bits), "bit planes must be 16-byte aligned");
  LA3D_CUDA(cudaFuncSetAttribute(ratio_median_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kCap * 4));
  ratio_median_kernel<<<(unsigned)planes, kThreads, kCap * 4, static_cast<cudaStream_t>(stream)>>>(
      depth_map, depth_render, mask_bits, render_bits, group, H * W, (int)la3d_words_per_plane(H, W), n_overlap, scale);
  LA3D_CUDA(cudaGetLastError());
  return LA3D_OK;
}
