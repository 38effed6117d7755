// Per-image subsample ranks on sm_100a: NumPy's legacy RandomState.randint stream.
//
// Replaces `rand_ind = np.random.randint(0, N, 500)` (src/util_3dbox.py:123-125 of
// the reference) for a batch.  One CTA of 128 threads owns one image:
//   - thread 0 seeds MT19937 the way np.random.seed(int) does (init_genrand; an
//     inherently serial 624-step recurrence) while warps 1..3 total the per-chunk
//     quarter counts of the image's planes (N per instance);
//   - the generator is then advanced 624 words at a time by the whole CTA (the twist
//     has dependency distance 227, so it runs as three data-parallel phases) and the
//     tempered words are parked in shared memory;
//   - the instances are walked in order over that buffer; an instance with N > 500
//     consumes draws until 500 of them pass NumPy's masked-rejection test
//     `(draw & mask) <= N-1`, mask = 2^k-1 >= N-1.  Each pass tests the rest of the
//     buffer in parallel (warp ballots, cross-warp prefix), stores the accepted
//     values at their ranks and locates the draw that completed the instance.
// Integer work: bit-exact with NumPy (tests/test_oracle_golden.py pins the
// restatement, tests/test_gpu_parity.py pins this kernel).
#include "common.cuh"

namespace la3d {
namespace {

constexpr int kThreads = 128;
constexpr int kWarps = kThreads / 32;
constexpr int kMtN = 624, kMtM = 397;
constexpr int kSegMax = 5;                 // ballots per warp per pass: ceil(624 / 4 / 32)
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ uint32_t twist(uint32_t cur, uint32_t nxt) {
  uint32_t y = (cur & 0x80000000u) | (nxt & 0x7fffffffu);
  return (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}

__device__ __forceinline__ uint32_t temper(uint32_t y) {
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}

// Next 624 words.  Word kk needs old[kk], old[kk+1] and old[kk+397] (kk < 227) or NEW[kk-227]:
// three phases [0,227), [227,454), [454,623] each read only words no thread of the phase writes,
// except old[kk+1] at the seam, so every phase reads, synchronises, then writes.
__device__ __forceinline__ void mt_next_block(uint32_t* mt, uint32_t* out) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int phase = 0; phase < 3; ++phase) {
    const int lo = phase * (kMtN - kMtM), hi = min(lo + (kMtN - kMtM), kMtN - 1);
    uint32_t val[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int kk = lo + tid + j * kThreads;
      val[j] = 0;
      if (kk < hi) {
        const uint32_t far = (kk < kMtN - kMtM) ? mt[kk + kMtM] : mt[kk - (kMtN - kMtM)];
        val[j] = far ^ twist(mt[kk], mt[kk + 1]);
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int kk = lo + tid + j * kThreads;
      if (kk < hi) mt[kk] = val[j];
    }
    __syncthreads();
  }
  if (tid == 0) mt[kMtN - 1] = mt[kMtM - 1] ^ twist(mt[kMtN - 1], mt[0]);
  __syncthreads();
  for (int k = tid; k < kMtN; k += kThreads) out[k] = temper(mt[k]);
  __syncthreads();
}

// init_genrand for one image per warp (lane 0 runs the recurrence, the warp stores the state).
// Launched ahead of the mask scan so that the serial part of seeding is off the critical path.
__global__ void __launch_bounds__(kThreads) seed_kernel(int B, uint32_t seed0, uint32_t* __restrict__ states) {
  __shared__ uint32_t st[kWarps][kMtN];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * kWarps + warp;
  if (b >= B) return;
  if (lane == 0) {
    uint32_t s = seed0 + (uint32_t)b;
#pragma unroll 8
    for (int i = 0; i < kMtN; ++i) {
      st[warp][i] = s;
      s = 1812433253u * (s ^ (s >> 30)) + (uint32_t)i + 1u;
    }
  }
  __syncwarp();
  for (int k = lane; k < kMtN; k += 32) states[(size_t)b * kMtN + k] = st[warp][k];
}

__global__ void __launch_bounds__(kThreads) sample_kernel(const uint32_t* __restrict__ chunk_counts, int I, int chunks,
                                                          uint32_t seed0, const uint32_t* __restrict__ states,
                                                          int32_t* __restrict__ counts,
                                                          int32_t* __restrict__ ranks) {
  __shared__ uint32_t mt[kMtN], out[kMtN];
  __shared__ int wtot[kWarps], end_pos;
  extern __shared__ uint32_t n_of[];          // [I] set pixels per instance
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x;

  if (states) {
    for (int k = tid; k < kMtN; k += kThreads) mt[k] = __ldg(states + (size_t)b * kMtN + k);   // seeded earlier
  }
  if (warp == 0) {
    if (lane == 0 && !states) {
      uint32_t s = seed0 + (uint32_t)b;       // mod 2^32, as np.random.seed requires
#pragma unroll 8
      for (int i = 0; i < kMtN; ++i) {
        mt[i] = s;
        s = 1812433253u * (s ^ (s >> 30)) + (uint32_t)i + 1u;
      }
    }
  } else {
    for (int i = warp - 1; i < I; i += kWarps - 1) {
      const uint32_t* cc = chunk_counts + (size_t)(b * I + i) * chunks;
      uint32_t n = 0;
#pragma unroll 8
      for (int c = lane; c < chunks; c += 32) n = __dp4a(__ldg(cc + c), 0x01010101u, n);   // sum of the 4 quarter bytes
      n = __reduce_add_sync(kFull, n);
      if (lane == 0) { n_of[i] = n; counts[b * I + i] = (int32_t)n; }
    }
  }
  __syncthreads();

  int pos = kMtN;     // next unread word of `out`; every variable below is uniform across the CTA
  int got = 0;
  int i = 0;
  while (i < I && n_of[i] <= (uint32_t)LA3D_SUBSAMPLE) ++i;
  while (i < I) {
    if (pos == kMtN) { mt_next_block(mt, out); pos = 0; }
    const uint32_t top = n_of[i] - 1u;
    uint32_t mask = top;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    int32_t* dst = ranks + (size_t)(b * I + i) * LA3D_SUBSAMPLE;

    // warp w tests the contiguous segment [pos + w*seg, pos + (w+1)*seg) of the buffer
    const int seg = (((kMtN - pos) + kWarps - 1) / kWarps + 31) & ~31;
    const int w0 = pos + warp * seg;
    uint32_t val[kSegMax], bal[kSegMax];
    int mine = 0;
#pragma unroll
    for (int j = 0; j < kSegMax; ++j) {
      const int e = w0 + j * 32 + lane;
      const bool in = (j * 32 < seg) && (e < kMtN);
      val[j] = in ? (out[e] & mask) : 0xffffffffu;
      bal[j] = __ballot_sync(kFull, in && val[j] <= top);
      mine += __popc(bal[j]);
    }
    if (lane == 0) wtot[warp] = mine;
    if (tid == 0) end_pos = -1;
    __syncthreads();
    int before = got, total = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      if (w < warp) before += wtot[w];
      total += wtot[w];
    }
    // store accepted values at their ranks; the warp holding the 500th locates its draw
    int run = before;
#pragma unroll
    for (int j = 0; j < kSegMax; ++j) {
      const int slot = run + __popc(bal[j] & ((1u << lane) - 1u));
      if (((bal[j] >> lane) & 1u) && slot < LA3D_SUBSAMPLE) dst[slot] = (int32_t)val[j];
      const int cnt = __popc(bal[j]);
      if (run < LA3D_SUBSAMPLE && run + cnt >= LA3D_SUBSAMPLE && lane == 0)
        end_pos = w0 + j * 32 + (int)__fns(bal[j], 0, LA3D_SUBSAMPLE - run);   // last draw this instance consumes
      run += cnt;
    }
    __syncthreads();
    if (end_pos >= 0) {
      pos = end_pos + 1;
      got = 0;
      ++i;
      while (i < I && n_of[i] <= (uint32_t)LA3D_SUBSAMPLE) ++i;
    } else {
      got += total;
      pos = kMtN;
    }
    __syncthreads();      // end_pos / wtot are rewritten by the next pass
  }
}

}  // namespace
}  // namespace la3d

extern "C" int la3d_sample_ranks(const uint32_t* chunk_counts, int B, int I, int H, int W, uint32_t seed,
                                 uint32_t image_offset, int32_t* counts, int32_t* ranks, la3d_stream_t stream) {
  using namespace la3d;
  LA3D_REQUIRE(chunk_counts && counts && ranks, "null pointer");
  LA3D_REQUIRE(B > 0 && I > 0 && H > 0 && W > 0, "non-positive shape");
  LA3D_REQUIRE(I <= 8192, "at most 8192 instances per image");
  const int chunks = (int)la3d_chunks_per_plane(H, W);
  sample_kernel<<<(unsigned)B, kThreads, (size_t)I * 4, static_cast<cudaStream_t>(stream)>>>(
      chunk_counts, I, chunks, seed + image_offset, nullptr, counts, ranks);
  LA3D_CUDA(cudaGetLastError());
  return LA3D_OK;
}

namespace la3d {
// Internal (pipeline in api.cu): seeding split from sampling.
int seed_states(int B, uint32_t seed0, uint32_t* states, cudaStream_t s) {
  seed_kernel<<<(unsigned)((B + kWarps - 1) / kWarps), kThreads, 0, s>>>(B, seed0, states);
  LA3D_CUDA(cudaGetLastError());
  return LA3D_OK;
}
int sample_seeded(const uint32_t* chunk_counts, int B, int I, int chunks, const uint32_t* states, int32_t* counts,
                  int32_t* ranks, cudaStream_t s) {
  sample_kernel<<<(unsigned)B, kThreads, (size_t)I * 4, s>>>(chunk_counts, I, chunks, 0u, states, counts, ranks);
  LA3D_CUDA(cudaGetLastError());
  return LA3D_OK;
}
}  // namespace la3d
