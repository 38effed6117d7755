// Per-image subsample ranks on sm_100a: NumPy's legacy RandomState.randint stream.
//
// Replaces `rand_ind = np.random.randint(0, N, 500)` (src/util_3dbox.py:123-125 of
// the reference) for a batch.  One warp owns one image: it seeds MT19937 the way
// np.random.seed(int) does (init_genrand), then walks the image's instances in
// order; an instance with N > 500 set pixels consumes draws until 500 of them
// pass the masked-rejection test `(draw & mask) <= N-1`, mask = 2^k-1 >= N-1.
// Integer work: bit-exact with NumPy (tests/test_oracle_golden.py pins the
// restatement, tests/test_gpu_parity.py pins this kernel).
#include "common.cuh"

namespace la3d {
namespace {

constexpr int kWarpsPerCta = 4;
constexpr int kMtN = 624, kMtM = 397;

__device__ __forceinline__ uint32_t twist(uint32_t cur, uint32_t nxt) {
  uint32_t y = (cur & 0x80000000u) | (nxt & 0x7fffffffu);
  return (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}

// Regenerate all 624 words, 32 at a time.  Word kk needs old[kk], old[kk+1] and
// either old[kk+397] (kk < 227) or NEW[kk-227]; a batch of 32 never reaches a
// word that the same batch writes except through old[kk+1], hence the
// read / sync / write / sync pattern.
__device__ __forceinline__ void mt_regenerate(uint32_t* mt, int lane) {
  for (int base = 0; base < kMtN - 1; base += 32) {
    const int kk = base + lane;
    uint32_t val = 0;
    if (kk < kMtN - 1) {
      const uint32_t far = (kk < kMtN - kMtM) ? mt[kk + kMtM] : mt[kk - (kMtN - kMtM)];
      val = far ^ twist(mt[kk], mt[kk + 1]);
    }
    __syncwarp();
    if (kk < kMtN - 1) mt[kk] = val;
    __syncwarp();
  }
  if (lane == 0) mt[kMtN - 1] = mt[kMtM - 1] ^ twist(mt[kMtN - 1], mt[0]);
  __syncwarp();
}

__device__ __forceinline__ uint32_t temper(uint32_t y) {
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}

__global__ void __launch_bounds__(kWarpsPerCta * 32) sample_kernel(const uint16_t* __restrict__ chunk_counts, int B,
                                                                   int I, int chunks, uint32_t seed0,
                                                                   int32_t* __restrict__ counts,
                                                                   int32_t* __restrict__ ranks) {
  __shared__ uint32_t mt_all[kWarpsPerCta][kMtN];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * kWarpsPerCta + warp;
  if (b >= B) return;
  uint32_t* mt = mt_all[warp];

  if (lane == 0) {
    uint32_t s = seed0 + (uint32_t)b;   // mod 2^32, as np.random.seed requires
    for (int i = 0; i < kMtN; ++i) {
      mt[i] = s;
      s = 1812433253u * (s ^ (s >> 30)) + (uint32_t)i + 1u;
    }
  }
  __syncwarp();
  int pos = kMtN;

  for (int i = 0; i < I; ++i) {
    const int plane = b * I + i;
    const uint16_t* cc = chunk_counts + (size_t)plane * chunks;
    uint32_t n = 0;
    for (int c = lane; c < chunks; c += 32) n += cc[c];
    n = __reduce_add_sync(0xffffffffu, n);
    if (lane == 0) counts[plane] = (int32_t)n;
    if (n <= (uint32_t)LA3D_SUBSAMPLE) continue;

    const uint32_t top = n - 1u;
    uint32_t mask = top;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    int32_t* dst = ranks + (size_t)plane * LA3D_SUBSAMPLE;
    int got = 0;
    while (got < LA3D_SUBSAMPLE) {
      if (pos == kMtN) { mt_regenerate(mt, lane); pos = 0; }
      const int take = min(32, kMtN - pos);
      const uint32_t v = (lane < take) ? (temper(mt[pos + lane]) & mask) : 0xffffffffu;
      const bool ok = (lane < take) && (v <= top);
      const uint32_t bal = __ballot_sync(0xffffffffu, ok);
      const int slot = got + __popc(bal & ((1u << lane) - 1u));
      if (ok && slot < LA3D_SUBSAMPLE) dst[slot] = (int32_t)v;
      const int tot = __popc(bal);
      if (got + tot >= LA3D_SUBSAMPLE) {
        // the draw that produced the 500th accepted value is the last one consumed
        pos += (int)__fns(bal, 0, LA3D_SUBSAMPLE - got) + 1;
        got = LA3D_SUBSAMPLE;
      } else {
        pos += take;
        got += tot;
      }
    }
  }
}

}  // namespace
}  // namespace la3d

extern "C" int la3d_sample_ranks(const uint16_t* chunk_counts, int B, int I, int H, int W, uint32_t seed,
                                 uint32_t image_offset, int32_t* counts, int32_t* ranks, la3d_stream_t stream) {
  using namespace la3d;
  LA3D_REQUIRE(chunk_counts && counts && ranks, "null pointer");
  LA3D_REQUIRE(B > 0 && I > 0 && H > 0 && W > 0, "non-positive shape");
  const int chunks = (int)la3d_chunks_per_plane(H, W);
  dim3 grid((unsigned)((B + kWarpsPerCta - 1) / kWarpsPerCta)), block(kWarpsPerCta * 32);
  sample_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(chunk_counts, B, I, chunks,
                                                                       seed + image_offset, counts, ranks);
  LA3D_CUDA(cudaGetLastError());
  return LA3D_OK;
}
