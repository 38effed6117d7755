// Mask-stack scan on sm_100a: byte masks -> bit masks + per-chunk popcounts.
//
// This is the HBM-bound pass of the box-fitting path: it reads every mask byte
// exactly once (I bytes per pixel) and writes 1/8 of that back as bits plus four
// bytes per 512 pixels (the set-pixel counts of the chunk's four 128-pixel quarters,
// one byte each - a quarter holds at most 128 = 0x80, so the bytes never carry).  Everything downstream (counts, the row-major rank
// select that stands in for NumPy's pts[mask], mask statistics) works on the
// bit planes.  Integer work: results are exact.
//
// A warp converts 4 consecutive 512-pixel chunks per step: 4 independent
// 16-byte loads per lane are in flight before the first is used.
//
// kPrep: the first `B` CTAs of the launch do not scan; they run the batch's
// mask-independent, latency-bound preparation (MT19937 words, cameras, ground rotations;
// prep.cuh), which so hides under the HBM-bound scan without a second stream.
#include "prep.cuh"

namespace la3d {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kUnroll = 4;                       // chunks per warp
constexpr int kTileChunks = kWarps * kUnroll;    // chunks per CTA

__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// 4 mask bytes -> 4 bits in the TOP nibble of the result (bit 28+k = byte k).
template <bool k01>
__device__ __forceinline__ uint32_t top_nibble(uint32_t w) {
  if (k01) {
    // bytes are 0/1: bits 0,8,16,24 -> 28,29,30,31 (partial products never collide)
    return w * 0x10204080u;
  } else {
    // bit 7 of each byte := byte != 0, then bits 7,15,23,31 -> 28,29,30,31
    uint32_t nz = (((w & 0x7f7f7f7fu) + 0x7f7f7f7fu) | w) & 0x80808080u;
    return nz * 0x00204081u;
  }
}

// 16 mask bytes -> 16 bits (bit k = byte k).
template <bool k01>
__device__ __forceinline__ uint32_t pack16(uint4 q) {
  uint32_t acc = 0;
  acc = __funnelshift_l(top_nibble<k01>(q.w), acc, 4);
  acc = __funnelshift_l(top_nibble<k01>(q.z), acc, 4);
  acc = __funnelshift_l(top_nibble<k01>(q.y), acc, 4);
  acc = __funnelshift_l(top_nibble<k01>(q.x), acc, 4);
  return acc;
}

template <bool k01, bool kVec, bool kPrep>
__global__ void __launch_bounds__(kThreads, kPrep ? 8 : 1)
    mask_scan_kernel(const uint8_t* __restrict__ masks, int HW, int chunks_per_plane, int tiles_per_plane,
                     uint32_t* __restrict__ bits, uint32_t* __restrict__ chunk_counts, PrepArgs pa) {
  int bid = blockIdx.x;
  if (kPrep) {
    if (bid < pa.B) { prep_body<kThreads>(pa, bid); return; }     // CTA-uniform
    bid -= pa.B;
  }
  const int plane = bid / tiles_per_plane;
  const int tile = bid - plane * tiles_per_plane;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = tile * kTileChunks + warp * kUnroll;
  const uint8_t* src = masks + (size_t)plane * HW;

  uint4 q[kUnroll];
#pragma unroll
  for (int j = 0; j < kUnroll; ++j) {
    const int px = (c0 + j) * kChunkPx + lane * 16;
    q[j] = make_uint4(0u, 0u, 0u, 0u);
    if (kVec) {
      if (c0 + j < chunks_per_plane && px < HW) q[j] = ld_stream(reinterpret_cast<const uint4*>(src + px));
    } else if (c0 + j < chunks_per_plane) {
      uint32_t w[4] = {0u, 0u, 0u, 0u};
      for (int k = 0; k < 16; ++k)
        if (px + k < HW) w[k >> 2] |= (uint32_t)__ldg(src + px + k) << (8 * (k & 3));
      q[j] = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }

  uint32_t* dst_bits = bits + (size_t)plane * chunks_per_plane * kChunkWords;
  uint32_t* dst_cnt = chunk_counts + (size_t)plane * chunks_per_plane;
#pragma unroll
  for (int j = 0; j < kUnroll; ++j) {
    const int c = c0 + j;
    if (c >= chunks_per_plane) break;                       // warp-uniform
    const uint32_t half = pack16<k01>(q[j]);
    const uint32_t other = __shfl_xor_sync(0xffffffffu, half, 1);
    // lanes 8q..8q+7 hold quarter q: one warp reduction yields all four byte counts
    const uint32_t quarters = __reduce_add_sync(0xffffffffu, (uint32_t)__popc(half) << (8 * (lane >> 3)));
    if ((lane & 1) == 0) dst_bits[c * kChunkWords + (lane >> 1)] = half | (other << 16);
    if (lane == 0) dst_cnt[c] = quarters;
  }
}

}  // namespace
}  // namespace la3d

extern "C" size_t la3d_chunks_per_plane(int H, int W) {
  return ((size_t)H * W + la3d::kChunkPx - 1) / la3d::kChunkPx;
}
extern "C" size_t la3d_words_per_plane(int H, int W) { return la3d_chunks_per_plane(H, W) * la3d::kChunkWords; }

namespace la3d {
// prep == nullptr: the plain scan.  Otherwise prep->B extra CTAs at the front of the grid prepare the batch.
int launch_mask_scan(const uint8_t* masks, int planes, int H, int W, int mask_is_01, uint32_t* bits,
                     uint32_t* chunk_counts, const PrepArgs* prep, cudaStream_t s) {
  LA3D_REQUIRE(masks && bits && chunk_counts, "null pointer");
  LA3D_REQUIRE(planes > 0 && H > 0 && W > 0, "non-positive shape");
  LA3D_REQUIRE((long long)H * W < (1ll << 30), "image too large");
  const int HW = H * W;
  const int chunks = (int)la3d_chunks_per_plane(H, W);
  const int tiles = (chunks + kTileChunks - 1) / kTileChunks;
  const long long ctas = (long long)tiles * planes + (prep ? prep->B : 0);
  LA3D_REQUIRE(ctas < (1ll << 31), "grid too large");
  const bool vec = (HW % 16 == 0) && aligned16(masks);
  dim3 grid((unsigned)ctas), block(kThreads);
  const PrepArgs pa = prep ? *prep : PrepArgs{};
#define LAUNCH(B01, VEC, PREP) \
  mask_scan_kernel<B01, VEC, PREP><<<grid, block, 0, s>>>(masks, HW, chunks, tiles, bits, chunk_counts, pa)
#define LAUNCH2(B01, VEC) do { if (prep) LAUNCH(B01, VEC, true); else LAUNCH(B01, VEC, false); } while (0)
  if (mask_is_01) { if (vec) LAUNCH2(true, true); else LAUNCH2(true, false); }
  else            { if (vec) LAUNCH2(false, true); else LAUNCH2(false, false); }
#undef LAUNCH2
#undef LAUNCH
  LA3D_CUDA(cudaGetLastError());
  return LA3D_OK;
}
}  // namespace la3d

extern "C" int la3d_mask_scan(const uint8_t* masks, int planes, int H, int W, int mask_is_01, uint32_t* bits,
                              uint32_t* chunk_counts, la3d_stream_t stream) {
  return la3d::launch_mask_scan(masks, planes, H, W, mask_is_01, bits, chunk_counts, nullptr,
                                static_cast<cudaStream_t>(stream));
}
