// 3D -> 2D pinhole projection of explicit points on sm_100a.  Replaces project_to_2d
// (src/util.py:227-229 = src/tools/combine_results.py:105-108 of the reference) for
// batches of points: [u, v] = (K p)[:2] / (K p)[2], float64, no behind-camera handling.
// (The box-fitting kernel projects its own 8 corners; this entry serves draw_cube and
// combine_results, which re-project corners read back from JSON.)
#include "common.cuh"

namespace la3d {
namespace {
__global__ void project_kernel(const double* __restrict__ pts, const double* __restrict__ K, int k_stride,
                               const int32_t* __restrict__ k_index, long long n, double* __restrict__ uv) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* Km = K + (size_t)(k_index ? k_index[i] : 0) * k_stride;
  const double x = pts[i * 3], y = pts[i * 3 + 1], z = pts[i * 3 + 2];
  const double h0 = Km[0] * x + Km[1] * y + Km[2] * z;
  const double h1 = Km[3] * x + Km[4] * y + Km[5] * z;
  const double h2 = Km[6] * x + Km[7] * y + Km[8] * z;
  uv[i * 2] = h0 / h2;
  uv[i * 2 + 1] = h1 / h2;
}
}  // namespace
}  // namespace la3d

extern "C" int la3d_project_points(const double* pts, const double* K, const int32_t* k_index, long long n, double* uv,
                                   la3d_stream_t stream) {
  using namespace la3d;
  LA3D_REQUIRE(pts && K && uv, "null pointer");
  LA3D_REQUIRE(n > 0, "non-positive point count");
  const int threads = 128;
  project_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, static_cast<cudaStream_t>(stream)>>>(
      pts, K, 9, k_index, n, uv);
  LA3D_CUDA(cudaGetLastError());
  return LA3D_OK;
}
