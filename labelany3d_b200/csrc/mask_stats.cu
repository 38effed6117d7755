// Mask-stack statistics on the bit planes of la3d_mask_scan (sm_100a).  "Next" row f1 of the scope
// table: the integer bookkeeping the reference does per instance mask before / around the box fit
//   - analyze_mask            src/util.py:291-326   area, pixels inside the four border bands
//   - get_maximum_height      src/util.py:328-335   last - first non-empty row + 1
//   - rows with any pixel     src/util.py:369-370   (`height = np.sum(rows)` of the RLE branch)
//   - filter_component_masks  src/model_wrappers.py:33-37   |mask & foreground| per mask
// Everything is integer work on 1/8 of the mask bytes (the scan already paid for the byte pass):
// results are exact.  One CTA per plane; threads stride over the plane's words (coalesced); a
// word is cut into its per-row segments, so any W works.
#include "common.cuh"

namespace la3d {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr unsigned kFull = 0xffffffffu;

struct Bands {
  int t0, t1, b0, b1;   // top / bottom bands: rows [t0,t1) and [b0,b1)
  int l0, l1, r0, r1;   // left / right bands: columns [l0,l1) and [r0,r1)
};

// bits of `word` (which holds columns [u, u+len) of one row in its bits [s, s+len)) whose column lies in [c0, c1)
__device__ __forceinline__ int pop_cols(uint32_t word, int s, int u, int len, int c0, int c1) {
  const int lo = max(c0, u), hi = min(c1, u + len);
  if (hi <= lo) return 0;
  const int n = hi - lo, sh = s + (lo - u);
  const uint32_t m = (n >= 32 ? 0xffffffffu : ((1u << n) - 1u)) << sh;
  return __popc(word & m);
}

__global__ void __launch_bounds__(kThreads) mask_stats_kernel(const uint32_t* __restrict__ bits, int words_per_plane,
                                                              int H, int W, Bands bd, int32_t* __restrict__ stats) {
  extern __shared__ uint32_t row_any[];            // [ceil(H/32)] one bit per row
  __shared__ int red[kWarps][5];
  const int plane = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row_words = (H + 31) >> 5;
  for (int k = tid; k < row_words; k += kThreads) row_any[k] = 0u;
  __syncthreads();
  const uint32_t* src = bits + (size_t)plane * words_per_plane;
  const int HW = H * W;
  int area = 0, top = 0, bottom = 0, left = 0, right = 0;
  for (int w = tid; w < words_per_plane; w += kThreads) {
    const uint32_t word = __ldg(src + w);
    if (word == 0u) continue;
    int p = w << 5;                                // first pixel of the word
    const int p_end = min(p + 32, HW);
    int r = p / W, u = p - r * W;
    int s = 0;
    while (p < p_end) {
      const int len = min(W - u, p_end - p);
      const uint32_t seg = (len >= 32 ? word : (word >> s) & ((1u << len) - 1u));
      if (seg) {
        const int c = __popc(seg);
        area += c;
        if (r >= bd.t0 && r < bd.t1) top += c;
        if (r >= bd.b0 && r < bd.b1) bottom += c;
        left += pop_cols(word, s, u, len, bd.l0, bd.l1);
        right += pop_cols(word, s, u, len, bd.r0, bd.r1);
        atomicOr(&row_any[r >> 5], 1u << (r & 31));
      }
      p += len; s += len; u = 0; ++r;
    }
  }
  int v[5] = {area, top, bottom, left, right};
#pragma unroll
  for (int k = 0; k < 5; ++k) v[k] = __reduce_add_sync(kFull, v[k]);
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < 5; ++k) red[warp][k] = v[k];
  __syncthreads();                                 // also publishes row_any
  if (warp == 0) {
    int first = 0x7fffffff, last = -1, rows = 0;
    for (int k = lane; k < row_words; k += 32) {
      const uint32_t m = row_any[k];
      if (m) {
        rows += __popc(m);
        first = min(first, (k << 5) + __ffs(m) - 1);
        last = max(last, (k << 5) + 31 - __clz(m));
      }
    }
    rows = __reduce_add_sync(kFull, rows);
    first = __reduce_min_sync(kFull, first);
    last = __reduce_max_sync(kFull, last);
    if (lane < 5) {
      int acc = 0;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) acc += red[w][lane];
      stats[(size_t)plane * 8 + lane] = acc;
    }
    if (lane == 5) stats[(size_t)plane * 8 + 5] = last < 0 ? -1 : first;
    if (lane == 6) stats[(size_t)plane * 8 + 6] = last;
    if (lane == 7) stats[(size_t)plane * 8 + 7] = rows;
  }
}

// W a multiple of 128: a row is a whole number of 16-byte pieces, so a thread's piece (128 pixels)
// lies in one row and the (row, piece) position advances without a division.
__global__ void __launch_bounds__(kThreads) mask_stats_rows_kernel(const uint32_t* __restrict__ bits,
                                                                   int words_per_plane, int H, int W, Bands bd,
                                                                   int32_t* __restrict__ stats) {
  extern __shared__ uint32_t row_any[];            // [ceil(H/32)] one bit per row
  __shared__ int red[kWarps][5];
  const int plane = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row_words = (H + 31) >> 5;
  for (int k = tid; k < row_words; k += kThreads) row_any[k] = 0u;
  __syncthreads();
  const uint4* src = reinterpret_cast<const uint4*>(bits + (size_t)plane * words_per_plane);
  const int q4 = W >> 7;                           // pieces per row
  const int total = H * q4;
  const int d_row = kThreads / q4, d_col = kThreads - d_row * q4;
  int row = tid / q4, col = tid - row * q4;
  int area = 0, top = 0, bottom = 0, left = 0, right = 0;
  for (int idx = tid; idx < total; idx += kThreads) {
    const uint4 v = __ldg(src + idx);
    if (v.x | v.y | v.z | v.w) {
      const int c = __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
      area += c;
      if (row >= bd.t0 && row < bd.t1) top += c;
      if (row >= bd.b0 && row < bd.b1) bottom += c;
      const int u = col << 7;                      // first column of the piece
      if (bd.l1 > u && bd.l0 < u + 128) {
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) left += pop_cols(w[k], 0, u + 32 * k, 32, bd.l0, bd.l1);
      }
      if (bd.r1 > u && bd.r0 < u + 128) {
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) right += pop_cols(w[k], 0, u + 32 * k, 32, bd.r0, bd.r1);
      }
      atomicOr(&row_any[row >> 5], 1u << (row & 31));
    }
    row += d_row; col += d_col;
    if (col >= q4) { col -= q4; ++row; }
  }
  int v5[5] = {area, top, bottom, left, right};
#pragma unroll
  for (int k = 0; k < 5; ++k) v5[k] = __reduce_add_sync(kFull, v5[k]);
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < 5; ++k) red[warp][k] = v5[k];
  __syncthreads();
  if (warp == 0) {
    int first = 0x7fffffff, last = -1, rows = 0;
    for (int k = lane; k < row_words; k += 32) {
      const uint32_t m = row_any[k];
      if (m) {
        rows += __popc(m);
        first = min(first, (k << 5) + __ffs(m) - 1);
        last = max(last, (k << 5) + 31 - __clz(m));
      }
    }
    rows = __reduce_add_sync(kFull, rows);
    first = __reduce_min_sync(kFull, first);
    last = __reduce_max_sync(kFull, last);
    if (lane < 5) {
      int acc = 0;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) acc += red[w][lane];
      stats[(size_t)plane * 8 + lane] = acc;
    }
    if (lane == 5) stats[(size_t)plane * 8 + 5] = last < 0 ? -1 : first;
    if (lane == 6) stats[(size_t)plane * 8 + 6] = last;
    if (lane == 7) stats[(size_t)plane * 8 + 7] = rows;
  }
}

// |a[p] & b[p / group]| per plane p: one warp per plane
__global__ void __launch_bounds__(kThreads) mask_overlap_kernel(const uint32_t* __restrict__ a,
                                                                const uint32_t* __restrict__ b, int planes, int group,
                                                                int words_per_plane, int32_t* __restrict__ inter) {
  const int plane = blockIdx.x * kWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (plane >= planes) return;
  const uint4* pa = reinterpret_cast<const uint4*>(a + (size_t)plane * words_per_plane);
  const uint4* pb = reinterpret_cast<const uint4*>(b + (size_t)(plane / group) * words_per_plane);
  int n = 0;
  for (int k = lane; k < words_per_plane / 4; k += 32) {     // words_per_plane is a multiple of 16
    const uint4 x = __ldg(pa + k), y = __ldg(pb + k);
    n += __popc(x.x & y.x) + __popc(x.y & y.y) + __popc(x.z & y.z) + __popc(x.w & y.w);
  }
  n = __reduce_add_sync(kFull, n);
  if (lane == 0) inter[plane] = n;
}

}  // namespace
}  // namespace la3d

extern "C" int la3d_mask_stats(const uint32_t* bits, int planes, int H, int W, const int* bands, int32_t* stats,
                               la3d_stream_t stream) {
  using namespace la3d;
  LA3D_REQUIRE(bits && bands && stats, "null pointer");
  LA3D_REQUIRE(planes > 0 && H > 0 && W > 0, "non-positive shape");
  LA3D_REQUIRE((long long)H * W < (1ll << 30), "image too large");
  const size_t dyn = (size_t)((H + 31) / 32) * 4;
  LA3D_REQUIRE(dyn <= 48 * 1024, "more than 393216 rows");
  Bands bd{bands[0], bands[1], bands[2], bands[3], bands[4], bands[5], bands[6], bands[7]};
  if (W % 128 == 0 && W / 128 <= kThreads && aligned16(bits))
    mask_stats_rows_kernel<<<(unsigned)planes, kThreads, dyn, static_cast<cudaStream_t>(stream)>>>(
        bits, (int)la3d_words_per_plane(H, W), H, W, bd, stats);
  else
    mask_stats_kernel<<<(unsigned)planes, kThreads, dyn, static_cast<cudaStream_t>(stream)>>>(
        bits, (int)la3d_words_per_plane(H, W), H, W, bd, stats);
  LA3D_CUDA(cudaGetLastError());
  return LA3D_OK;
}

extern "C" int la3d_mask_overlap(const uint32_t* bits, const uint32_t* other_bits, int planes, int group, int H, int W,
                                 int32_t* inter, la3d_stream_t stream) {
  using namespace la3d;
  LA3D_REQUIRE(bits && other_bits && inter, "null pointer");
  LA3D_REQUIRE(planes > 0 && group > 0 && H > 0 && W > 0, "non-positive shape");
  LA3D_REQUIRE(aligned16(bits) && aligned16(other_bits), "bit planes must be 16-byte aligned");
  mask_overlap_kernel<<<(unsigned)((planes + kWarps - 1) / kWarps), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      bits, other_bits, planes, group, (int)la3d_words_per_plane(H, W), inter);
  LA3D_CUDA(cudaGetLastError());
  return LA3D_OK;
}
