// COCO run-length masks -> bit planes + quarter counts on sm_100a.
//
// The reference builds its bool [I,H,W] mask stack on the host by decoding COCO / COCONUT
// run-length annotations one by one (mask_utils.decode, src/util.py:361-370) and only then
// hands it to the path.  Here the runs are the input: one CTA per plane turns them straight
// into what la3d_mask_scan would have produced from the decoded bytes (bit k of word j =
// row-major pixel 32 j + k; one count byte per 128-pixel quarter of a 512-pixel chunk), so the
// byte masks (I bytes per pixel, the dominant HBM stream of the scanned path and 8x the bits
// over PCIe) never exist.  Integer work: bit-exact.
//
// COCO runs are COLUMN-major (pixel (y,x) has run position x*H + y), the bit planes are
// row-major, so the kernel is a bit transposition:
//   1. inclusive prefix sums of the plane's run lengths -> run ends E[] (shared memory);
//   2. per column the run that holds its top pixel (one binary search per column, so every later
//      search stays inside one column's handful of runs), and the row / column range the set
//      pixels can lie in (most of a plane is empty: those rows are written as zeros directly);
//   3. a band of 32 rows is exactly W words of the plane's bit stream (32*W pixels) and belongs to one
//      warp: per strip of 32 columns lane l owns column 32 s + l, walks the (few) runs crossing its 32
//      pixels into a 32-bit column word, and the warp transposes the 32x32 bit tile with five shuffle
//      rounds (skipped when the tile is empty).  W % 128 == 0 (640, 1536, ...): rows are whole words and
//      quarters four strips of a row, so each lane stores its row's word straight from the register over
//      the zeroed band and accumulates its quarter's popcount (one red.global per non-empty quarter).
//      Any other W: the tiles go to a padded staging tile, the warp re-packs the staged rows with funnel
//      shifts into finished words, stores them coalesced and reduces their popcounts per quarter with a
//      segmented shuffle reduction.  No block barrier after the set-up.
//
// kPrep: like the mask scan, the first ceil(B/8) CTAs of the launch may instead run the
// mask-independent preparation of the batch (prep.cuh), so la3d_fit_boxes_rle stays at three launches.
#include <cstdlib>

#include "prep.cuh"

namespace la3d {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr unsigned kFull = 0xffffffffu;
constexpr int kMaxSmemRuns = 32768;            // 128 KB of run ends at most in shared memory

struct RleArgs {
  const uint32_t* counts;      // run lengths of all planes, back to back
  const long long* offsets;    // [planes+1] first run of every plane
  uint32_t* ends_ws;           // nullable [total runs]: run ends of planes that do not fit shared memory
  int H, W, HW, chunks, smem_runs, stage_words;   // stage_words: staging tile per warp (0 on the fast path)
  int fast;                    // W % 128 == 0 and 16-byte aligned bit planes: rows and quarters never straddle words
  int sparse;                  // bands without a set pixel are not written (fused step: the planes are read by the
                               // fit's rank select only, which never lands in a chunk whose count is 0)
  uint32_t* bits;
  uint32_t* chunk_counts;
  int32_t* status;
};

// a[l] bit j  ->  a[j] bit l  across the warp (recursive block swap, five rounds).
__device__ __forceinline__ uint32_t transpose32(uint32_t a, int lane) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    // bits j with (j & s) == 0
    const uint32_t low = s == 16 ? 0x0000ffffu : s == 8 ? 0x00ff00ffu : s == 4 ? 0x0f0f0f0fu : s == 2 ? 0x33333333u : 0x55555555u;
    const uint32_t other = __shfl_xor_sync(kFull, a, s);
    a = (lane & s) ? ((a & ~low) | ((other >> s) & low)) : ((a & low) | ((other << s) & ~low));
  }
  return a;
}

// number of run ends <= q, i.e. the index of the run that holds position q
__device__ __forceinline__ int run_of(const uint32_t* __restrict__ E, int m, uint32_t q) {
  int lo = 0, hi = m;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (E[mid] <= q) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ uint32_t bit_range(uint32_t a, uint32_t b) {       // bits [a, b), 0 <= a < b <= 32
  const uint32_t hi = b >= 32u ? 0xffffffffu : ((1u << b) - 1u);
  return hi & ~((1u << a) - 1u);
}

// number of run ends <= q among E[lo..hi), plus lo: run_of restricted to a column's runs
__device__ __forceinline__ int run_of_in(const uint32_t* __restrict__ E, int lo, int hi, uint32_t q) {
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (E[mid] <= q) lo = mid + 1; else hi = mid;
  }
  return lo;
}

template <bool kPrep, int kMinCtas>
__global__ void __launch_bounds__(kThreads, kMinCtas) rle_decode_kernel(RleArgs a, PrepArgs pa) {
  extern __shared__ __align__(16) uint32_t dyn[];
  __shared__ unsigned long long warp_tot[kWarps];
  __shared__ int s_ylo[kWarps], s_yhi[kWarps];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int plane = blockIdx.x;
  if (kPrep) {
    const int prep_ctas = (pa.B + kWarps - 1) / kWarps;           // one warp per image
    if (plane == 0) publish_epoch(pa.pub);
    if (plane < prep_ctas) { prep_body<kThreads, true>(pa, plane * kWarps, dyn); return; }   // CTA-uniform
    plane -= prep_ctas;
  }
  const int H = a.H, W = a.W, HW = a.HW;
  const int pitch = (W + 31) >> 5;               // 32-column strips per row
  const int P = pitch | 1;                       // staging pitch: odd, so a tile's 32 rows hit 32 banks
  uint32_t* stage = dyn + (size_t)warp * a.stage_words;                     // [kWarps][32][P]: every warp stages its own band
  int* col_run = reinterpret_cast<int*>(dyn + (size_t)kWarps * a.stage_words);   // [W+1] run holding the top pixel of a column
  uint32_t* ends_smem = dyn + (size_t)kWarps * a.stage_words + (W + 1);     // [smem_runs]

  const long long r0 = a.offsets[plane];
  const long long m_ll = a.offsets[plane + 1] - r0;
  int m = (int)min(m_ll, (long long)0x7fffffff);
  const uint32_t* cnt = a.counts + r0;
  uint32_t* E = ends_smem;
  int st = 0;
  if (m > a.smem_runs) {
    if (a.ends_ws) E = a.ends_ws + r0;
    else { st = 2; m = 0; }                      // more runs than the caller announced and no workspace: refuse
  }

  // ---- 1. run ends: inclusive prefix sums, clamped to HW+1 (anything beyond the image is equal) ----
  const int per = (m + kThreads - 1) / kThreads;
  const int j_lo = min(tid * per, m), j_hi = min(j_lo + per, m);
  unsigned long long run = 0;
  for (int j = j_lo; j < j_hi; ++j) run += cnt[j];
  unsigned long long incl = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long up = __shfl_up_sync(kFull, incl, o);
    if (lane >= o) incl += up;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  unsigned long long base = incl - run, total = 0;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) {
    if (w < warp) base += warp_tot[w];
    total += warp_tot[w];
  }
  for (int j = j_lo; j < j_hi; ++j) {
    base += cnt[j];
    E[j] = (uint32_t)min(base, (unsigned long long)HW + 1ull);
  }
  if (total > (unsigned long long)HW) st = 1;
  if (tid == 0) a.status[plane] = st;
  uint32_t* cc = a.chunk_counts + (size_t)plane * a.chunks;
  for (int c = tid; c < a.chunks; c += kThreads) cc[c] = 0u;
  __syncthreads();

  // ---- 2. where the set pixels can be: per column the run that holds its top pixel (all later searches
  // stay inside one column's runs), the column range, and the row range (a 1-run that crosses into the
  // next column touches the last and the first row)
  for (int x = tid; x < W; x += kThreads) col_run[x] = run_of(E, m, (uint32_t)x * (uint32_t)H);
  if (tid == 0) col_run[W] = m;
  int y_lo = H, y_hi = -1;
  for (int j = 1 + 2 * tid; j < m; j += 2 * kThreads) {
    const uint32_t s0 = E[j - 1], e0 = min(E[j], (uint32_t)HW);
    if (e0 > s0) {
      const uint32_t cs = s0 / (uint32_t)H, ce = (e0 - 1u) / (uint32_t)H;
      if (cs != ce) { y_lo = 0; y_hi = H - 1; }
      else { y_lo = min(y_lo, (int)(s0 - cs * (uint32_t)H)); y_hi = max(y_hi, (int)(e0 - 1u - cs * (uint32_t)H)); }
    }
  }
  y_lo = __reduce_min_sync(kFull, y_lo);
  y_hi = __reduce_max_sync(kFull, y_hi);
  if (lane == 0) { s_ylo[warp] = y_lo; s_yhi[warp] = y_hi; }
  __syncthreads();
#pragma unroll
  for (int w = 0; w < kWarps; ++w) { y_lo = min(y_lo, s_ylo[w]); y_hi = max(y_hi, s_yhi[w]); }
  int x_lo = W, x_hi = -1;
  if (m >= 2 && y_hi >= y_lo) {
    const uint32_t first = E[0];
    const uint32_t last = min((m & 1) ? E[m - 2] : E[m - 1], (uint32_t)HW);
    if (last > first) { x_lo = (int)(first / (uint32_t)H); x_hi = (int)((last - 1u) / (uint32_t)H); }
  }
  const int s_lo = x_lo >> 5, s_hi = x_hi >> 5;  // strips that can hold a set pixel (none if x_hi < x_lo)

  // ---- 3. bands of 32 rows = W words of the bit stream each; a warp owns whole bands ----
  const long long words_total = (long long)a.chunks * kChunkWords;
  uint32_t* out_bits = a.bits + (size_t)plane * words_total;
  const int n_bands = (int)((words_total + W - 1) / W);
  for (int band = warp; band < n_bands; band += kWarps) {
    const long long w_base = (long long)band * W;
    const int n_words = (int)min((long long)W, words_total - w_base);
    const long long y0_ll = (long long)band * 32;
    const int nrows = y0_ll >= H ? 0 : min(32, H - (int)y0_ll);
    const int y0 = (int)min(y0_ll, (long long)H);
    if (nrows == 0 || x_hi < x_lo || y0 > y_hi || y0 + nrows - 1 < y_lo) {   // nothing set in these rows
      if (a.sparse) {
        // only the words of chunks (16 words) this band shares with a neighbouring band: that one may hold pixels
        const long long hi = w_base + n_words;
        const long long head_end = min(hi, (w_base + 15) & ~15ll), tail_beg = max(head_end, hi & ~15ll);
        for (long long wi = w_base + lane; wi < head_end; wi += 32) out_bits[wi] = 0u;
        for (long long wi = tail_beg + lane; wi < hi; wi += 32) out_bits[wi] = 0u;
        continue;
      }
      if (a.fast) {
        uint4* z = reinterpret_cast<uint4*>(out_bits + w_base);
        for (int i = lane; i < (n_words >> 2); i += 32) z[i] = make_uint4(0u, 0u, 0u, 0u);
      } else {
        for (int wi = lane; wi < n_words; wi += 32) out_bits[w_base + wi] = 0u;
      }
      continue;
    }
    // one 32x32 tile: lane l owns column 32 s + l, walks the runs crossing its rows of the band into a
    // column word, and the warp transposes the tile (skipped when empty): lane = row, bit = column
    auto tile = [&](int s) -> uint32_t {
      const int x = 32 * s + lane;
      uint32_t word = 0;
      if (x < W) {
        const int k_top = col_run[x], k_next = col_run[x + 1];
        const uint32_t q0 = (uint32_t)x * (uint32_t)H + (uint32_t)y0, q1 = q0 + (uint32_t)nrows;
        int k = run_of_in(E, k_top, k_next, q0);
        uint32_t pos = q0;
        while (pos < q1 && k < m) {
          const uint32_t e = E[k];
          const uint32_t end = min(e, q1);
          if ((k & 1) && end > pos) word |= bit_range(pos - q0, end - q0);
          pos = end;
          if (e <= q1) ++k;
        }
      }
      if (__any_sync(kFull, word != 0u)) word = transpose32(word, lane);
      return word;
    };
    if (a.fast) {
      // W % 128 == 0: a row is `pitch` whole words and a quarter is four strips of one row.  Zero the band with
      // 16-byte stores, then every lane stores its row's word of each tile straight from the register and
      // keeps the popcount of the quarter it is in; no staging, no shuffles for the counts.
      uint4* z = reinterpret_cast<uint4*>(out_bits + w_base);
      for (int i = lane; i < (n_words >> 2); i += 32) z[i] = make_uint4(0u, 0u, 0u, 0u);
      __syncwarp();                                       // orders the zeros before the other lanes' words below
      uint32_t acc = 0;
      for (int s = s_lo; s <= s_hi; ++s) {
        const uint32_t word = tile(s);
        const long long wg = w_base + (long long)lane * pitch + s;
        if (word) out_bits[wg] = word;                    // rows past the image hold no bits and are never stored
        acc += __popc(word);
        if ((s & 3) == 3 || s == s_hi) {
          if (acc) atomicAdd(&cc[wg >> 4], acc << (8 * (int)((wg >> 2) & 3)));
          acc = 0;
        }
      }
      continue;
    }
    for (int s = s_lo; s <= s_hi; ++s) stage[lane * P + s] = tile(s);
    __syncwarp();
    for (int base_w = 0; base_w < n_words; base_w += 32) {                    // warp-uniform trip count
      const int wi = base_w + lane;
      const bool active = wi < n_words;
      uint32_t out = 0;
      if (active) {
        const uint32_t p = 32u * (uint32_t)wi;
        int r = (int)(p / (uint32_t)W);
        int x = (int)(p - (uint32_t)r * (uint32_t)W);
        int got = 0;
        while (got < 32 && r < 32) {
          const int n = min(32 - got, W - x);
          const int c = x >> 5;
          const uint32_t w0 = (c >= s_lo && c <= s_hi) ? stage[r * P + c] : 0u;
          const uint32_t w1 = (c + 1 >= s_lo && c + 1 <= s_hi) ? stage[r * P + c + 1] : 0u;
          uint32_t v = __funnelshift_r(w0, w1, x & 31);
          if (n < 32) v &= (1u << n) - 1u;
          out |= v << got;
          got += n; ++r; x = 0;
        }
        out_bits[w_base + wi] = out;
      }
      if (!__any_sync(kFull, out != 0u)) continue;                            // warp-uniform
      // quarter counts: lanes of one 4-word quarter are neighbours; sum them (segmented shuffle
      // reduction), one red per quarter
      const long long wg = w_base + wi;
      const uint32_t key = active ? (uint32_t)(wg >> 2) : (0x80000000u | (uint32_t)lane);
      uint32_t v = __popc(out);
      const uint32_t k1 = __shfl_down_sync(kFull, key, 1), v1 = __shfl_down_sync(kFull, v, 1);
      if (lane + 1 < 32 && k1 == key) v += v1;
      const uint32_t k2 = __shfl_down_sync(kFull, key, 2), v2 = __shfl_down_sync(kFull, v, 2);
      if (lane + 2 < 32 && k2 == key) v += v2;
      const uint32_t kp = __shfl_up_sync(kFull, key, 1);
      const bool head = lane == 0 || kp != key;
      if (active && head && v) atomicAdd(&cc[wg >> 4], v << (8 * (int)((wg >> 2) & 3)));
    }
    __syncwarp();
  }
}

}  // namespace

// prep == nullptr: the plain decode.  Otherwise ceil(prep->B / 8) extra CTAs at the front of the grid prepare the batch.
int launch_rle_decode(const uint32_t* counts, const int64_t* offsets, int planes, int H, int W, int max_runs,
                      uint32_t* ends_ws, uint32_t* bits, uint32_t* chunk_counts, int32_t* status, const PrepArgs* prep,
                      cudaStream_t s, bool sparse_bits) {
  LA3D_REQUIRE(counts && offsets && bits && chunk_counts && status, "null pointer");
  LA3D_REQUIRE(planes > 0 && H > 0 && W > 0, "non-positive shape");
  LA3D_REQUIRE((long long)H * W < (1ll << 30), "image too large");
  LA3D_REQUIRE(W <= 4096, "image wider than 4096 pixels");
  LA3D_REQUIRE(max_runs >= 0, "negative max_runs");
  static_assert(sizeof(long long) == sizeof(int64_t), "offsets are 64-bit");
  RleArgs a{};
  a.counts = counts; a.offsets = reinterpret_cast<const long long*>(offsets); a.ends_ws = ends_ws;
  a.H = H; a.W = W; a.HW = H * W; a.chunks = (int)la3d_chunks_per_plane(H, W);
  a.bits = bits; a.chunk_counts = chunk_counts; a.status = status;
  a.fast = (W % 128 == 0) && aligned16(bits);
  static const bool sparse_env = !(getenv("LA3D_SCAN_SPARSE") && atoi(getenv("LA3D_SCAN_SPARSE")) == 0);
  a.sparse = sparse_bits && prep && sparse_env;
  // shared memory: a staging tile per warp, the per-column run table, and the run ends if they fit beside them
  const int P = ((W + 31) >> 5) | 1;
  a.stage_words = a.fast ? 0 : 32 * P;
  const size_t fixed = ((size_t)kWarps * a.stage_words + (size_t)W + 1) * 4;
  const size_t budget = 200 * 1024;
  a.smem_runs = (max_runs <= kMaxSmemRuns && fixed + (size_t)max_runs * 4 <= budget) ? max_runs : 0;
  if (a.smem_runs == 0 && !ends_ws && max_runs > 0) {
    set_error("la3d_rle_decode: %d runs per plane do not fit shared memory; pass the ends_ws workspace", max_runs);
    return LA3D_EINVAL;
  }
  size_t smem = fixed + (size_t)a.smem_runs * 4;
  if (prep && smem < (size_t)kWarps * kMtN * 4) smem = (size_t)kWarps * kMtN * 4;      // the preparation CTAs' generator states
  const long long ctas = (long long)planes + (prep ? (prep->B + kWarps - 1) / kWarps : 0);
  LA3D_REQUIRE(ctas < (1ll << 31), "grid too large");
  const PrepArgs pa = prep ? *prep : PrepArgs{};
  // LA3D_RLE_VARIANT=1: compile-time bound of 6 resident CTAs per SM (40 registers, spills) instead of 5 (tuning knob)
  const char* env = getenv("LA3D_RLE_VARIANT");
  const int variant = env ? atoi(env) : 0;
  auto launch = [&](auto kernel, int slot) -> int {
    // the opt-in is per device and cheap: repeat it whenever a launch needs it (no process-wide cache)
    (void)slot;
    if (smem > 48 * 1024)
      LA3D_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kernel<<<(unsigned)ctas, kThreads, smem, s>>>(a, pa);
    return LA3D_OK;
  };
  int rc;
  if (variant == 1) rc = prep ? launch(rle_decode_kernel<true, 6>, 0) : launch(rle_decode_kernel<false, 6>, 1);
  else rc = prep ? launch(rle_decode_kernel<true, 5>, 2) : launch(rle_decode_kernel<false, 5>, 3);
  if (rc) return rc;
  LA3D_CUDA(cudaGetLastError());
  return LA3D_OK;
}

}  // namespace la3d

extern "C" int la3d_rle_decode(const uint32_t* counts, const int64_t* offsets, int planes, int H, int W, int max_runs,
                               uint32_t* ends_ws, uint32_t* bits, uint32_t* chunk_counts, int32_t* status,
                               la3d_stream_t stream) {
  return la3d::launch_rle_decode(counts, offsets, planes, H, W, max_runs, ends_ws, bits, chunk_counts, status, nullptr,
                                 static_cast<cudaStream_t>(stream));
}
