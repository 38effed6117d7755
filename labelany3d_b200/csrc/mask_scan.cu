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
// kPrep: the first ceil(B/8) CTAs of the launch do not scan; each of their warps runs one
// image's mask-independent, latency-bound preparation (MT19937 words, cameras, ground
// rotations; prep.cuh), which so hides under the HBM-bound scan without a second stream.
#include <cstdlib>

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

// kSparse (the fused step only, whose bit planes live in its own workspace and are read by nothing but the fit
// kernel's rank select): the bit words of a chunk without set pixels are not written - the select never lands in
// such a chunk (its count is 0), and with masks covering a few percent of the image most of the bit planes'
// bytes, 1/9 of the scan's DRAM traffic, are such zeros.
template <bool k01, bool kVec, bool kPrep, bool kSparse = false>
__global__ void __launch_bounds__(kThreads, kPrep ? 8 : 1)
    mask_scan_kernel(const uint8_t* __restrict__ masks, int HW, int chunks_per_plane, int tiles_per_plane,
                     uint32_t* __restrict__ bits, uint32_t* __restrict__ chunk_counts, PrepArgs pa) {
  pdl_trigger();          // the sampler (if launched as a programmatic dependent) may be scheduled early; it waits
  int bid = blockIdx.x;
  if (kPrep) {
    const int prep_ctas = (pa.B + kWarps - 1) / kWarps;           // one warp per image
    if (bid == 0) publish_epoch(pa.pub);
    if (bid < prep_ctas) { prep_body<kThreads>(pa, bid * kWarps); return; }     // CTA-uniform
    bid -= prep_ctas;
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
    if ((lane & 1) == 0 && (!kSparse || quarters != 0u)) dst_bits[c * kChunkWords + (lane >> 1)] = half | (other << 16);
    if (lane == 0) dst_cnt[c] = quarters;
  }
}

// ---------------------------------------------------------------------------------------------
// The "thin" form of the scan: a persistent kernel, a few CTAs per SM, whose memory-level
// parallelism comes from the TMA unit instead of from resident threads.  One elected thread per CTA
// streams the mask stack (a flat run of planes x H*W bytes when H*W is a multiple of 512) through a
// ring of 16 KB shared-memory stages with cp.async.bulk + mbarrier; eight converter warps turn
// each stage into bits and quarter counts, 32 bytes -> one finished bit word per lane.
// Measured on B200 (config 2, tools/scan_variants.py): 3 CTAs/SM x 2 stages 5.90 TB/s, 2 x 3 5.51,
// 1 x 4 3.2 TB/s (one CTA per SM is bound by the producer -> TMA -> mbarrier -> converter latency
// chain, not by bytes in flight); the one-tile-per-CTA kernel above reaches 5.75 in the same loop.
// It was built to run UNDER the sampler / fit of another batch (two streams): measured, both
// kernels slow down by what the other takes (profiles/r1_pipeline_timeline.txt) - the SMs are
// issue- and register-bound across scan + fit, so the step stays serial and this kernel is an
// alternative entry point (la3d_mask_scan_thin), not the default.
// ---------------------------------------------------------------------------------------------
constexpr int kStageBytes = 16384;                 // 32 chunks
constexpr int kStageChunks = kStageBytes / kChunkPx;
constexpr int kConvWarps = 8;
constexpr int kThinThreads = 32 + kConvWarps * 32;

template <bool k01, int kStages>
__global__ void __launch_bounds__(kThinThreads, 1)
    mask_scan_thin_kernel(const uint8_t* __restrict__ masks, long long total_bytes, uint32_t* __restrict__ bits,
                          uint32_t* __restrict__ chunk_counts) {
  extern __shared__ __align__(128) unsigned char ring[];
  __shared__ __align__(8) uint64_t full[kStages], empty[kStages];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long n_tiles = (total_bytes + kStageBytes - 1) / kStageBytes;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kConvWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        const int s = it % kStages;
        mbar_wait(&empty[s], ((it / kStages) & 1) ^ 1);        // a fresh barrier passes the first round
        const long long off = t * kStageBytes;
        const uint32_t bytes = (uint32_t)min((long long)kStageBytes, total_bytes - off);
        mbar_expect_tx(&full[s], bytes);
        tma_load_1d(ring + (size_t)s * kStageBytes, masks + off, bytes, &full[s]);
      }
    }
    return;
  }
  const int cw = warp - 1;
  int it = 0;
  for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
    const int s = it % kStages;
    const long long off = t * kStageBytes;
    const int n_chunks = (int)(min((long long)kStageBytes, total_bytes - off) / kChunkPx);
    mbar_wait(&full[s], (it / kStages) & 1);
    const unsigned char* stage = ring + (size_t)s * kStageBytes;
    // a lane converts 32 consecutive bytes into one finished bit word: a warp covers two chunks
    // (1024 pixels) per round and a stage takes two rounds per warp
    constexpr int kRounds = kStageChunks / 2 / kConvWarps;
    uint4 q[kRounds][2];
#pragma unroll
    for (int j = 0; j < kRounds; ++j) {
      const int pair = cw + j * kConvWarps;                    // chunks 2*pair, 2*pair+1 of the stage
      const bool live = 2 * pair + (lane >> 4) < n_chunks;
      const uint4* src = reinterpret_cast<const uint4*>(stage + pair * (2 * kChunkPx) + lane * 32);
      q[j][0] = live ? src[0] : make_uint4(0u, 0u, 0u, 0u);
      q[j][1] = live ? src[1] : make_uint4(0u, 0u, 0u, 0u);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);                    // the stage is in registers: hand it back
    const long long chunk0 = t * kStageChunks;
#pragma unroll
    for (int j = 0; j < kRounds; ++j) {
      const int pair = cw + j * kConvWarps;
      if (2 * pair >= n_chunks) break;                         // warp-uniform
      const bool live = 2 * pair + (lane >> 4) < n_chunks;
      const uint32_t word = pack16<k01>(q[j][0]) | (pack16<k01>(q[j][1]) << 16);
      // lanes 4q..4q+3 of a half-warp hold quarter q of its chunk (4 x 32 = 128 pixels: a byte never carries)
      const uint32_t cnt = (uint32_t)__popc(word) << (8 * ((lane >> 2) & 3));
      const uint32_t lo = __reduce_add_sync(0xffffffffu, lane < 16 ? cnt : 0u);
      const uint32_t hi = __reduce_add_sync(0xffffffffu, lane < 16 ? 0u : cnt);
      if (live) bits[(chunk0 + 2 * pair) * kChunkWords + lane] = word;
      if (lane == 0) chunk_counts[chunk0 + 2 * pair] = lo;
      if (lane == 16 && live) chunk_counts[chunk0 + 2 * pair + 1] = hi;
    }
  }
}

}  // namespace
}  // namespace la3d

extern "C" size_t la3d_chunks_per_plane(int H, int W) {
  return ((size_t)H * W + la3d::kChunkPx - 1) / la3d::kChunkPx;
}
extern "C" size_t la3d_words_per_plane(int H, int W) { return la3d_chunks_per_plane(H, W) * la3d::kChunkWords; }

namespace la3d {
static int g_scan_variant = -1;
int scan_variant() {
  if (g_scan_variant < 0) g_scan_variant = getenv("LA3D_SCAN_VARIANT") ? atoi(getenv("LA3D_SCAN_VARIANT")) : 0;
  return g_scan_variant;
}
void set_scan_variant(int v) { g_scan_variant = v < 0 ? 0 : v; }

// thin persistent form (see mask_scan_thin_kernel)
static int launch_thin(const uint8_t* masks, int planes, int HW, int mask_is_01, uint32_t* bits, uint32_t* chunk_counts,
                       int ctas_per_sm, int stages, cudaStream_t s) {
  int dev = 0, sms = 0;
  LA3D_CUDA(cudaGetDevice(&dev));
  LA3D_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const long long total = (long long)planes * HW;
  const long long n_tiles = (total + kStageBytes - 1) / kStageBytes;
  const unsigned g = (unsigned)min((long long)sms * (ctas_per_sm > 0 ? ctas_per_sm : 1), n_tiles);
#define THIN(B01, ST) do { \
    const int smem = ST * kStageBytes; \
    LA3D_CUDA(cudaFuncSetAttribute(mask_scan_thin_kernel<B01, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
    mask_scan_thin_kernel<B01, ST><<<g, kThinThreads, smem, s>>>(masks, total, bits, chunk_counts); } while (0)
  if (stages >= 6) { if (mask_is_01) THIN(true, 6); else THIN(false, 6); }
  else if (stages >= 4) { if (mask_is_01) THIN(true, 4); else THIN(false, 4); }
  else if (stages >= 3) { if (mask_is_01) THIN(true, 3); else THIN(false, 3); }
  else { if (mask_is_01) THIN(true, 2); else THIN(false, 2); }
#undef THIN
  LA3D_CUDA(cudaGetLastError());
  return LA3D_OK;
}

// prep == nullptr: the plain scan.  Otherwise prep->B extra CTAs at the front of the grid prepare the batch.
// sparse_bits (needs prep): see kSparse.
int launch_mask_scan(const uint8_t* masks, int planes, int H, int W, int mask_is_01, uint32_t* bits,
                     uint32_t* chunk_counts, const PrepArgs* prep, cudaStream_t s, bool sparse_bits) {
  LA3D_REQUIRE(masks && bits && chunk_counts, "null pointer");
  LA3D_REQUIRE(planes > 0 && H > 0 && W > 0, "non-positive shape");
  LA3D_REQUIRE((long long)H * W < (1ll << 30), "image too large");
  const int HW = H * W;
  const int chunks = (int)la3d_chunks_per_plane(H, W);
  const int tiles = (chunks + kTileChunks - 1) / kTileChunks;
  const long long ctas = (long long)tiles * planes + (prep ? (prep->B + kWarps - 1) / kWarps : 0);
  LA3D_REQUIRE(ctas < (1ll << 31), "grid too large");
  const bool vec = (HW % 16 == 0) && aligned16(masks);
  const int variant = scan_variant();
  if (!prep && variant > 0 && HW % kChunkPx == 0 && aligned16(masks)) {
    static const int per_sm = getenv("LA3D_SCAN_CTAS") ? atoi(getenv("LA3D_SCAN_CTAS")) : 3;
    return launch_thin(masks, planes, HW, mask_is_01, bits, chunk_counts, per_sm, variant, s);
  }
  dim3 grid((unsigned)ctas), block(kThreads);
  const PrepArgs pa = prep ? *prep : PrepArgs{};
  static const bool sparse_env = !(getenv("LA3D_SCAN_SPARSE") && atoi(getenv("LA3D_SCAN_SPARSE")) == 0);
  const bool sparse = sparse_bits && prep && sparse_env;
#define LAUNCH(B01, VEC, PREP) \
  mask_scan_kernel<B01, VEC, PREP><<<grid, block, 0, s>>>(masks, HW, chunks, tiles, bits, chunk_counts, pa)
#define LAUNCH_SPARSE(B01, VEC) \
  mask_scan_kernel<B01, VEC, true, true><<<grid, block, 0, s>>>(masks, HW, chunks, tiles, bits, chunk_counts, pa)
#define LAUNCH2(B01, VEC) do { if (prep && sparse) LAUNCH_SPARSE(B01, VEC); else if (prep) LAUNCH(B01, VEC, true); else LAUNCH(B01, VEC, false); } while (0)
  if (mask_is_01) { if (vec) LAUNCH2(true, true); else LAUNCH2(true, false); }
  else            { if (vec) LAUNCH2(false, true); else LAUNCH2(false, false); }
#undef LAUNCH2
#undef LAUNCH_SPARSE
#undef LAUNCH
  LA3D_CUDA(cudaGetLastError());
  return LA3D_OK;
}
}  // namespace la3d

extern "C" int la3d_mask_scan(const uint8_t* masks, int planes, int H, int W, int mask_is_01, uint32_t* bits,
                              uint32_t* chunk_counts, la3d_stream_t stream) {
  return la3d::launch_mask_scan(masks, planes, H, W, mask_is_01, bits, chunk_counts, nullptr,
                                static_cast<cudaStream_t>(stream), false);
}

extern "C" int la3d_mask_scan_thin(const uint8_t* masks, int planes, int H, int W, int mask_is_01, uint32_t* bits,
                                   uint32_t* chunk_counts, int ctas_per_sm, int stages, la3d_stream_t stream) {
  using namespace la3d;
  LA3D_REQUIRE(masks && bits && chunk_counts, "null pointer");
  LA3D_REQUIRE(planes > 0 && H > 0 && W > 0, "non-positive shape");
  LA3D_REQUIRE((long long)H * W < (1ll << 30), "image too large");
  LA3D_REQUIRE((H * W) % kChunkPx == 0 && aligned16(masks), "the thin scan needs H*W to be a multiple of 512 and a 16-byte aligned stack");
  LA3D_REQUIRE(ctas_per_sm >= 1 && ctas_per_sm <= 4 && stages >= 2, "1..4 CTAs per SM, at least 2 stages");
  return launch_thin(masks, planes, H * W, mask_is_01, bits, chunk_counts, ctas_per_sm, stages, static_cast<cudaStream_t>(stream));
}
