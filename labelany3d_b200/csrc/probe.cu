// Measurement aid, not part of the box path: scattered 4-byte reads from a buffer (device memory, or pinned host
// memory mapped into the device's address space), `kInflight` independent loads per thread per round.  Gives the
// ceiling of the in-place depth gather when the depth map stays in host memory (the run-length end-to-end leg):
// tools/gather_ceiling.py, DESIGN.md section 6.
#include "common.cuh"

namespace la3d {

__device__ __forceinline__ uint32_t mix(uint32_t x) {          // integer hash (avalanche of a counter)
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

// share > 1: `share` neighbouring lanes read the same 128-byte line, `spread` elements apart (8: one 32-byte sector
// each, 1: the same sector), which is what address-sorted samples of one mask look like to the memory system.
template <int kInflight>
__global__ void __launch_bounds__(256) scatter_read_kernel(const float* __restrict__ src, uint32_t n, uint32_t window,
                                                           int rounds, int share, int spread, float* __restrict__ out) {
  const uint32_t t0 = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t t = t0 / (uint32_t)share, sub = (t0 % (uint32_t)share) * (uint32_t)spread;
  // every CTA draws inside its own window of the buffer (window == n: the whole buffer), like a box that
  // samples the pixels of one mask
  const uint32_t base = window < n ? mix(blockIdx.x) % (n - window + 1) : 0;
  float acc = 0.f;
  for (int r = 0; r < rounds; ++r) {
    float v[kInflight];
#pragma unroll
    for (int k = 0; k < kInflight; ++k)
      v[k] = __ldg(src + ((base + mix((t * (uint32_t)rounds + r) * kInflight + k + 0x9e3779b9u) % window) & (share > 1 ? ~31u : ~0u)) + sub);
#pragma unroll
    for (int k = 0; k < kInflight; ++k) acc += v[k];
  }
  out[t0] = acc;
}

}  // namespace la3d

extern "C" int la3d_debug_scatter_read(const float* src, size_t n, size_t window, int inflight, int rounds, int ctas,
                                       int share, int spread, float* out, la3d_stream_t stream) {
  using namespace la3d;
  LA3D_REQUIRE(src && out, "null pointer");
  LA3D_REQUIRE(n > 0 && n < (1ull << 32) && window > 0 && window <= n && rounds > 0 && ctas > 0, "bad sizes");
  LA3D_REQUIRE(share >= 1 && share <= 32 && spread >= 0 && (share - 1) * spread < 32 && n % 32 == 0, "bad sharing");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const uint32_t n32 = (uint32_t)n, w32 = (uint32_t)window;
  switch (inflight) {
    case 1: scatter_read_kernel<1><<<ctas, 256, 0, s>>>(src, n32, w32, rounds, share, spread, out); break;
    case 2: scatter_read_kernel<2><<<ctas, 256, 0, s>>>(src, n32, w32, rounds, share, spread, out); break;
    case 4: scatter_read_kernel<4><<<ctas, 256, 0, s>>>(src, n32, w32, rounds, share, spread, out); break;
    case 8: scatter_read_kernel<8><<<ctas, 256, 0, s>>>(src, n32, w32, rounds, share, spread, out); break;
    case 16: scatter_read_kernel<16><<<ctas, 256, 0, s>>>(src, n32, w32, rounds, share, spread, out); break;
    default: set_error("la3d_debug_scatter_read: inflight must be 1, 2, 4, 8 or 16"); return LA3D_EINVAL;
  }
  LA3D_CUDA(cudaGetLastError());
  return LA3D_OK;
}
