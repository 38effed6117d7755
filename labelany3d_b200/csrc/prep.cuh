// MT19937 block generation and the mask-independent preparation of a batch (see sample.cu).
// Shared by the stand-alone prep kernel (sample.cu) and by the prep CTAs that ride in the mask
// scan's launch (mask_scan.cu).
#pragma once

#include "common.cuh"

namespace la3d {

constexpr int kMtM = 397;

// Multi-GPU: the launch that OPENS a step also publishes the previous step's epoch in every rank's flag row (sink.cuh):
// at a kernel boundary every record the previous fit kernel stored into peer memory has been performed, so one
// thread's release store does what a fence + counter in every fit CTA would.
struct PeerPublish {
  uint32_t* flags[LA3D_MAX_PEERS];
  int n, rank;            // n == 0: nothing to publish
  uint32_t epoch;
};

struct PrepArgs {
  const double* K;        // [B][9]
  const double* ground;   // [B*I][3] or null
  int B, I;
  uint32_t seed0;         // image b seeds with seed0 + b (mod 2^32)
  PrepView pv;
  PeerPublish pub;
};

// First preparation CTA of a launch: threads 0 .. n-1 publish (no-op without peers).
__device__ __forceinline__ void publish_epoch(const PeerPublish& pub) {
  if ((int)threadIdx.x < pub.n)
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(pub.flags[threadIdx.x] + pub.rank), "r"(pub.epoch) : "memory");
}

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
// except old[kk+1] at the seam, so every phase reads, synchronises, then writes.  The tempered
// words go to `out` (shared or global); the caller synchronises before reading them.
template <int kT>
__device__ __forceinline__ void mt_next_block(uint32_t* mt, uint32_t* out) {
  const int tid = threadIdx.x;
  constexpr int kSpan = kMtN - kMtM;                 // 227
  constexpr int kIter = (kSpan + kT - 1) / kT;
#pragma unroll
  for (int phase = 0; phase < 3; ++phase) {
    const int lo = phase * kSpan, hi = min(lo + kSpan, kMtN - 1);
    uint32_t val[kIter];
#pragma unroll
    for (int j = 0; j < kIter; ++j) {
      const int kk = lo + tid + j * kT;
      val[j] = 0;
      if (kk < hi) {
        const uint32_t far = (kk < kSpan) ? mt[kk + kMtM] : mt[kk - kSpan];
        val[j] = far ^ twist(mt[kk], mt[kk + 1]);
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kIter; ++j) {
      const int kk = lo + tid + j * kT;
      if (kk < hi) mt[kk] = val[j];
    }
    __syncthreads();
  }
  if (tid == 0) mt[kMtN - 1] = mt[kMtM - 1] ^ twist(mt[kMtN - 1], mt[0]);
  __syncthreads();
  for (int k = tid; k < kMtN; k += kT) out[k] = temper(mt[k]);
}

// The same for ONE WARP (no block barrier): 227 words per phase = 8 rounds of 32 lanes.
__device__ __forceinline__ void mt_next_block_warp(uint32_t* mt, uint32_t* out) {
  const int lane = threadIdx.x & 31;
  constexpr int kSpan = kMtN - kMtM;                 // 227
  constexpr int kIter = (kSpan + 31) / 32;
#pragma unroll
  for (int phase = 0; phase < 3; ++phase) {
    const int lo = phase * kSpan, hi = min(lo + kSpan, kMtN - 1);
    uint32_t val[kIter];
#pragma unroll
    for (int j = 0; j < kIter; ++j) {
      const int kk = lo + lane + j * 32;
      val[j] = 0;
      if (kk < hi) {
        const uint32_t far = (kk < kSpan) ? mt[kk + kMtM] : mt[kk - kSpan];
        val[j] = far ^ twist(mt[kk], mt[kk + 1]);
      }
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < kIter; ++j) {
      const int kk = lo + lane + j * 32;
      if (kk < hi) mt[kk] = val[j];
    }
    __syncwarp();
  }
  if (lane == 0) mt[kMtN - 1] = mt[kMtM - 1] ^ twist(mt[kMtN - 1], mt[0]);
  __syncwarp();
  for (int k = lane; k < kMtN; k += 32) out[k] = temper(mt[k]);
  __syncwarp();
}

// One WARP prepares one image (a CTA of kT threads prepares kT/32 consecutive images; `first` is the
// CTA's first image): lane 0 seeds MT19937 (init_genrand), lane 1 inverts the intrinsics, the other
// lanes build the ground rotations; then the warp generates the image's first pv.nblk blocks of
// tempered words and saves the state.  Latency-bound on purpose: these warps ride in the launch of
// the HBM-bound mask scan and should take as few of its slots as possible.
// kExtMt: the generator states live in caller-provided shared memory (`mt_ext`, kT/32 x 624 words) instead
// of a static array, for host kernels whose other CTAs should not carry those 20 KB.
template <int kT, bool kExtMt = false>
__device__ __forceinline__ void prep_body(const PrepArgs& pa, int first, uint32_t* mt_ext = nullptr) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = first + warp;
  if (b >= pa.B) return;                               // warp-uniform; no block barrier below
  uint32_t* mt;
  if constexpr (kExtMt) {
    mt = mt_ext + warp * kMtN;
  } else {
    __shared__ uint32_t mt_all[kT / 32][kMtN];
    mt = mt_all[warp];
  }
  const PrepView& pv = pa.pv;
  if (lane == 0) {
    uint32_t s = pa.seed0 + (uint32_t)b;       // mod 2^32, as np.random.seed requires
#pragma unroll 8
    for (int i = 0; i < kMtN; ++i) {
      mt[i] = s;
      s = 1812433253u * (s ^ (s >> 30)) + (uint32_t)i + 1u;
    }
  } else if (lane == 1) {
    PrepCamera* cam = pv.cams + b;
    double Km[9], Kinv[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) Km[i] = pa.K[(size_t)b * 9 + i];
    invert3x3(Km, Kinv);
#pragma unroll
    for (int i = 0; i < 9; ++i) { cam->K[i] = Km[i]; cam->Kinv[i] = Kinv[i]; }
  } else {
    for (int i = lane - 2; i < pa.I; i += 30) {
      double Rg[9];
      ground_rotation(pa.ground ? pa.ground + ((size_t)b * pa.I + i) * 3 : nullptr, Rg);
#pragma unroll
      for (int k = 0; k < 9; ++k) pv.Rg[((size_t)b * pa.I + i) * 9 + k] = Rg[k];
    }
  }
  __syncwarp();
  uint32_t* words = pv.words + (size_t)b * pv.nblk * kMtN;
  for (int blk = 0; blk < pv.nblk; ++blk) mt_next_block_warp(mt, words + (size_t)blk * kMtN);
  for (int k = lane; k < kMtN; k += 32) pv.state[(size_t)b * kMtN + k] = mt[k];
}

int launch_mask_scan(const uint8_t* masks, int planes, int H, int W, int mask_is_01, uint32_t* bits,
                     uint32_t* chunk_counts, const PrepArgs* prep, cudaStream_t s, bool sparse_bits = false);

int launch_rle_decode(const uint32_t* counts, const int64_t* offsets, int planes, int H, int W, int max_runs,
                      uint32_t* ends_ws, uint32_t* bits, uint32_t* chunk_counts, int32_t* status, const PrepArgs* prep,
                      cudaStream_t s, bool sparse_bits = false);

struct RecordSink;
int fit_scanned_sink(const float* depth, const void* prep, const uint32_t* bits, const uint32_t* chunk_counts,
                     const int32_t* ranks, int B, int I, int H, int W, int method, int yaw_steps,
                     const RecordSink& sink, cudaStream_t stream, bool pdl = false, int b0 = 0, int Bp = -1);
int fit_all_sink(const float* depth, const void* prep, const uint32_t* bits, int B, int I, int H, int W, int method,
                 int yaw_steps, const RecordSink& sink, cudaStream_t stream);

}  // namespace la3d
