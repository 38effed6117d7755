// Per-image subsample ranks on sm_100a: NumPy's legacy RandomState.randint stream.
//
// Replaces `rand_ind = np.random.randint(0, N, 500)` (src/util_3dbox.py:123-125 of
// the reference) for a batch, in two kernels:
//
// prep (prep.cuh; one WARP per image; depends on nothing the mask scan produces, so
// la3d_fit_boxes runs it as a few extra CTAs inside the scan's launch, see mask_scan.cu):
//   - lane 0 seeds MT19937 the way np.random.seed(int) does (init_genrand, a serial 624-step
//     recurrence), lane 1 inverts the image's intrinsics, the other lanes build the ground
//     rotation of each instance (the scalar float64 work the fit kernel would otherwise repeat
//     per box);
//   - the warp then advances the generator 624 words at a time (the twist has dependency
//     distance 227, so a block is three data-parallel phases) and writes the TEMPERED words of
//     the first `nblk` blocks to global memory, followed by the raw state, from which the
//     consumer can continue should it ever run out of words.
//
// sample_kernel (one CTA of 256 threads per image, after the scan):
//   - totals the per-chunk quarter counts of the image's planes (N per instance);
//   - walks the instances in order over the pre-generated words, staged in shared
//     memory: an instance with N > 500 consumes draws until 500 of them pass NumPy's
//     masked-rejection test `(draw & mask) <= N-1`, mask = 2^k-1 >= N-1.  One pass tests
//     a window of 1024 words (4 consecutive words per thread, block-wide prefix sum),
//     stores the accepted values at their ranks and locates the draw that completed
//     the instance; the next instance starts at the word after it.
// Integer work: bit-exact with NumPy (tests/test_oracle_golden.py pins the
// restatement, tests/test_gpu_parity.py pins these kernels, including the
// out-of-words continuation).
#include "prep.cuh"

namespace la3d {
namespace {

constexpr int kPrepThreads = 128;
constexpr int kThreads = 256;              // sample_kernel
constexpr int kWarps = kThreads / 32;
constexpr int kPer = 4;                    // words tested per thread per pass
constexpr int kWindow = kThreads * kPer;   // words per pass
constexpr int kSegBlocksMax = 16;          // words staged in shared memory at a time: 16 x 624 x 4 B = 39 KB
constexpr unsigned kFull = 0xffffffffu;

int g_mt_blocks = 0;                       // 0 = automatic (la3d_set_mt_blocks)
int g_seg_blocks = 0;                      // 0 = automatic (la3d_set_sample_seg_blocks)
__device__ unsigned long long* g_sample_clocks = nullptr;   // debug: [images][4] globaltimer stamps (la3d_debug_sample_clocks)
__device__ __forceinline__ void sample_stamp(int b, int i) {
  if (g_sample_clocks && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    g_sample_clocks[(size_t)b * 4 + i] = t;
  }
}

__global__ void __launch_bounds__(kPrepThreads) prep_kernel(PrepArgs pa) {
  prep_body<kPrepThreads>(pa, blockIdx.x * (kPrepThreads / 32));
}

__global__ void __launch_bounds__(kThreads) sample_kernel(const uint32_t* __restrict__ chunk_counts, int I, int chunks,
                                                          PrepView pv, int seg_cap, int32_t* __restrict__ counts,
                                                          int32_t* __restrict__ ranks, int b0) {
  __shared__ uint32_t mt[kMtN];               // generator state, only if the pre-generated words run out
  __shared__ int wtot[kWarps], end_pos;
  extern __shared__ __align__(16) uint32_t dyn[];
  uint32_t* seg = dyn;                        // [seg_cap] staged words
  uint32_t* n_of = dyn + seg_cap;             // [I] set pixels per instance
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x + b0;
  const int nwords = pv.nblk * kMtN;          // pre-generated words of this image
  const uint32_t* words = pv.words + (size_t)b * nwords;

  // everything this kernel reads comes from the scan launch (counts, pre-generated words): wait for it,
  // then let the fit kernel start its own prologue (which reads only what the scan launch wrote)
  pdl_wait();
  pdl_trigger();
  sample_stamp(b, 0);
  // stage the first segment (one TMA bulk copy, signalled on an mbarrier) while the counts are totalled
  __shared__ __align__(8) uint64_t seg_bar;
  uint32_t seg_phase = 0;
  int seg_base = 0, seg_len = min(seg_cap, nwords);
  if (tid == 0) {
    mbar_init(&seg_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(&seg_bar, (uint32_t)seg_len * 4u);
    tma_load_1d(seg, words, (uint32_t)seg_len * 4u, &seg_bar);
  }
  for (int i = warp; i < I; i += kWarps) {
    const uint32_t* cc = chunk_counts + (size_t)(b * I + i) * chunks;
    uint32_t n = 0;
    if ((chunks & 3) == 0 && aligned16_dev(cc)) {   // 16-byte loads, all of a lane's in flight together
      const uint4* cc4 = reinterpret_cast<const uint4*>(cc);
#pragma unroll 8
      for (int c = lane; c < chunks / 4; c += 32) {
        const uint4 q = __ldg(cc4 + c);
        n = __dp4a(q.x, 0x01010101u, n); n = __dp4a(q.y, 0x01010101u, n);
        n = __dp4a(q.z, 0x01010101u, n); n = __dp4a(q.w, 0x01010101u, n);
      }
    } else {
#pragma unroll 4
      for (int c = lane; c < chunks; c += 32) n = __dp4a(__ldg(cc + c), 0x01010101u, n);   // sum of the 4 quarter bytes
    }
    n = __reduce_add_sync(kFull, n);
    if (lane == 0) { n_of[i] = n; counts[b * I + i] = (int32_t)n; }
  }
  __syncthreads();                              // n_of and the barrier's initialisation are visible
  sample_stamp(b, 1);
  mbar_wait(&seg_bar, seg_phase);
  seg_phase ^= 1u;
  sample_stamp(b, 2);

  // every variable below is uniform across the CTA
  int pos = 0;        // next unread word of the image's stream
  int got = 0;        // accepted draws of the current instance so far
  bool have_state = false;
  int i = 0;
  while (i < I && n_of[i] <= (uint32_t)LA3D_SUBSAMPLE) ++i;
  while (i < I) {
    if (pos >= seg_base + seg_len) {
      __syncthreads();                        // everyone is done with the old segment
      if (pos < nwords) {
        seg_base = pos & ~3;                    // 16-byte aligned source for the bulk copy
        seg_len = min(seg_cap, nwords - seg_base);
        if (tid == 0) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy accesses of seg come first
          mbar_expect_tx(&seg_bar, (uint32_t)seg_len * 4u);
          tma_load_1d(seg, words + seg_base, (uint32_t)seg_len * 4u, &seg_bar);
        }
        mbar_wait(&seg_bar, seg_phase);
        seg_phase ^= 1u;
      } else {
        seg_base = pos;
        // out of pre-generated words (pos == nwords + a multiple of 624): continue the generator
        if (!have_state) {
          for (int k = tid; k < kMtN; k += kThreads) mt[k] = __ldg(pv.state + (size_t)b * kMtN + k);
          have_state = true;
          __syncthreads();
        }
        mt_next_block<kThreads>(mt, seg);
        seg_len = kMtN;
      }
      __syncthreads();
    }
    const uint32_t top = n_of[i] - 1u;
    uint32_t mask = top;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    int32_t* dst = ranks + (size_t)(b * I + i) * LA3D_SUBSAMPLE;

    // thread t tests words [pos + 4t, pos + 4t + 4) of the window
    const int w_end = min(pos + kWindow, seg_base + seg_len);
    const int e0 = pos + tid * kPer;
    uint32_t val[kPer];
    int ok = 0;
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
      const bool in = e0 + j < w_end;
      val[j] = in ? (seg[e0 + j - seg_base] & mask) : 0xffffffffu;
      ok |= (int)(in && val[j] <= top) << j;
    }
    const int mine = __popc(ok);
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int up = __shfl_up_sync(kFull, incl, o);
      if (lane >= o) incl += up;
    }
    if (lane == 31) wtot[warp] = incl;
    if (tid == 0) end_pos = -1;
    __syncthreads();
    int before = got + incl - mine, total = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      if (w < warp) before += wtot[w];
      total += wtot[w];
    }
    // store accepted values at their ranks; the thread holding the 500th reports its draw
    int slot = before;
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
      if ((ok >> j) & 1) {
        if (slot < LA3D_SUBSAMPLE) dst[slot] = (int32_t)val[j];
        if (slot == LA3D_SUBSAMPLE - 1) end_pos = e0 + j;      // last draw this instance consumes
        ++slot;
      }
    }
    __syncthreads();
    if (end_pos >= 0) {
      pos = end_pos + 1;
      got = 0;
      ++i;
      while (i < I && n_of[i] <= (uint32_t)LA3D_SUBSAMPLE) ++i;
    } else {
      got += total;
      pos = w_end;
    }
    __syncthreads();      // end_pos / wtot are rewritten by the next pass
  }
  sample_stamp(b, 3);
}

int auto_blocks(int I) {
  if (g_mt_blocks > 0) return g_mt_blocks;
  // an instance needs 500 / p draws, p = N / 2^ceil(log2 N) in (0.5, 1]: at most ~1000 on average
  long long n = ((long long)I * 1024 + kMtN - 1) / kMtN + 1;
  return (int)(n > 64 ? 64 : n);
}

}  // namespace

PrepView prep_view(void* base, int B, int I, int nblk) {
  auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
  unsigned char* p = static_cast<unsigned char*>(base);
  PrepView v{};
  size_t off = 0;
  v.cams = reinterpret_cast<PrepCamera*>(p + off);       off = up(off + (size_t)B * sizeof(PrepCamera));
  v.Rg = reinterpret_cast<double*>(p + off);             off = up(off + (size_t)B * I * 9 * 8);
  v.state = reinterpret_cast<uint32_t*>(p + off);        off = up(off + (size_t)B * kMtN * 4);
  v.words = reinterpret_cast<uint32_t*>(p + off);        off = up(off + (size_t)B * nblk * kMtN * 4);
  v.nblk = nblk;
  v.bytes = off;
  return v;
}

int prep_blocks(int I) { return auto_blocks(I); }

// Blocks of 624 words a sampler CTA stages in shared memory at a time (it refills when they are used up).
static int seg_blocks(int B) {
  (void)B;
  if (g_seg_blocks > 0) return g_seg_blocks < kSegBlocksMax ? g_seg_blocks : kSegBlocksMax;
  return kSegBlocksMax;
}

int launch_sample(const uint32_t* chunk_counts, const PrepView& pv, int B, int I, int chunks, int32_t* counts,
                  int32_t* ranks, cudaStream_t s, bool pdl, int b0) {
  const int seg_cap = (pv.nblk < seg_blocks(B) ? pv.nblk : seg_blocks(B)) * kMtN;
  const size_t dyn = ((size_t)seg_cap + I) * 4;
  if (dyn > 48 * 1024)
    LA3D_CUDA(cudaFuncSetAttribute(sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
  LA3D_CUDA(launch_pdl(sample_kernel, dim3((unsigned)B), dim3(kThreads), dyn, s, pdl, chunk_counts, I, chunks, pv, seg_cap,
                       counts, ranks, b0));
  return LA3D_OK;
}

}  // namespace la3d

extern "C" void la3d_debug_sample_clocks(unsigned long long* clocks) {
  cudaMemcpyToSymbol(la3d::g_sample_clocks, &clocks, sizeof(clocks));
}

extern "C" void la3d_set_sample_seg_blocks(int n) { la3d::g_seg_blocks = n > 0 ? n : 0; }

extern "C" void la3d_set_mt_blocks(int n) { la3d::g_mt_blocks = n > 0 ? (n > 4096 ? 4096 : n) : 0; }

extern "C" size_t la3d_prep_bytes(int B, int I) {
  if (B <= 0 || I <= 0) return 0;
  return la3d::prep_view(nullptr, B, I, la3d::prep_blocks(I)).bytes;
}

extern "C" int la3d_fit_prepare(const double* K, const double* ground, int B, int I, uint32_t seed,
                                uint32_t image_offset, void* prep, size_t prep_bytes, la3d_stream_t stream) {
  using namespace la3d;
  LA3D_REQUIRE(K && prep, "null pointer");
  LA3D_REQUIRE(B > 0 && I > 0, "non-positive shape");
  LA3D_REQUIRE(I <= 8192, "at most 8192 instances per image");
  LA3D_REQUIRE((reinterpret_cast<uintptr_t>(prep) & 255u) == 0, "prep buffer must be 256-byte aligned");
  const PrepView pv = prep_view(prep, B, I, prep_blocks(I));
  if (prep_bytes < pv.bytes) {
    set_error("la3d_fit_prepare: buffer of %zu bytes, %zu needed", prep_bytes, pv.bytes);
    return LA3D_ENOMEM;
  }
  prep_kernel<<<(unsigned)((B + kPrepThreads / 32 - 1) / (kPrepThreads / 32)), kPrepThreads, 0, static_cast<cudaStream_t>(stream)>>>(PrepArgs{K, ground, B, I, seed + image_offset, pv});
  LA3D_CUDA(cudaGetLastError());
  return LA3D_OK;
}

extern "C" int la3d_sample_ranks(const uint32_t* chunk_counts, const void* prep, int B, int I, int H, int W,
                                 int32_t* counts, int32_t* ranks, la3d_stream_t stream) {
  using namespace la3d;
  LA3D_REQUIRE(chunk_counts && prep && counts && ranks, "null pointer");
  LA3D_REQUIRE(B > 0 && I > 0 && H > 0 && W > 0, "non-positive shape");
  LA3D_REQUIRE(I <= 8192, "at most 8192 instances per image");
  const PrepView pv = prep_view(const_cast<void*>(prep), B, I, prep_blocks(I));
  return launch_sample(chunk_counts, pv, B, I, (int)la3d_chunks_per_plane(H, W), counts, ranks,
                       static_cast<cudaStream_t>(stream));
}
