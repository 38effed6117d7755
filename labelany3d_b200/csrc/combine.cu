// Arithmetic of the combine stage on sm_100a ("next" row f2 of the scope table): the step right
// after the box fit, src/tools/combine_results.py of the reference.
//   - la3d_iou_matrix:        iou2D (:111-123) for every pair of two box lists (the cost matrix of
//                             hungarian_matching, :130-134), per scene, all scenes in one launch.
//                             Same operation order, no FMA: bit-identical to the Python floats.
//   - la3d_box2d_from_corners: bbox2D_proj / bbox2D_trunc (:234-252) of boxes given by their 8
//                             corners (as read back from 3dbbox.json), Python min()/max() semantics.
// The assignment itself (scipy.optimize.linear_sum_assignment, :137) stays on the host like the
// reference: it is a few boxes per scene.
#include "common.cuh"

namespace la3d {
namespace {

__device__ __forceinline__ double py_max(double a, double b) { return b > a ? b : a; }   // max(a, b) of Python
__device__ __forceinline__ double py_min(double a, double b) { return b < a ? b : a; }   // min(a, b) of Python

__global__ void iou_kernel(const double* __restrict__ boxes0, const int64_t* __restrict__ off0,
                           const double* __restrict__ boxes1, const int64_t* __restrict__ off1,
                           const int64_t* __restrict__ out_off, double* __restrict__ iou) {
  const int g = blockIdx.x;
  const long long a0 = off0[g], n0 = off0[g + 1] - a0;
  const long long a1 = off1[g], n1 = off1[g + 1] - a1;
  double* out = iou + out_off[g];
  for (long long e = threadIdx.x; e < n0 * n1; e += blockDim.x) {
    const long long i = e / n1, j = e - i * n1;
    const double* p = boxes0 + (a0 + i) * 4;
    const double* q = boxes1 + (a1 + j) * 4;
    const double x1 = py_max(p[0], q[0]), y1 = py_max(p[1], q[1]);
    const double x2 = py_min(p[2], q[2]), y2 = py_min(p[3], q[3]);
    const double inter = __dmul_rn(py_max(0.0, __dsub_rn(x2, x1)), py_max(0.0, __dsub_rn(y2, y1)));
    const double area1 = __dmul_rn(__dsub_rn(p[2], p[0]), __dsub_rn(p[3], p[1]));
    const double area2 = __dmul_rn(__dsub_rn(q[2], q[0]), __dsub_rn(q[3], q[1]));
    out[e] = __ddiv_rn(inter, __dadd_rn(__dsub_rn(__dadd_rn(area1, area2), inter), 1e-6));
  }
}

__global__ void box2d_kernel(const double* __restrict__ corners, const double* __restrict__ K,
                             const int32_t* __restrict__ k_index, const double* __restrict__ wh, int n,
                             double* __restrict__ proj, double* __restrict__ trunc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int m = k_index ? k_index[i] : 0;
  const double* Km = K + (size_t)m * 9;
  double mnu = 0, mnv = 0, mxu = 0, mxv = 0;
  for (int c = 0; c < 8; ++c) {
    const double* p = corners + ((size_t)i * 8 + c) * 3;
    const double h0 = Km[0] * p[0] + Km[1] * p[1] + Km[2] * p[2];
    const double h1 = Km[3] * p[0] + Km[4] * p[1] + Km[5] * p[2];
    const double h2 = Km[6] * p[0] + Km[7] * p[1] + Km[8] * p[2];
    const double u = h0 / h2, v = h1 / h2;
    if (c == 0) { mnu = mxu = u; mnv = mxv = v; }
    else {                                      // Python min()/max() over a generator: `if x < m` / `if x > m`
      if (u < mnu) mnu = u;
      if (v < mnv) mnv = v;
      if (u > mxu) mxu = u;
      if (v > mxv) mxv = v;
    }
  }
  proj[(size_t)i * 4] = mnu; proj[(size_t)i * 4 + 1] = mnv; proj[(size_t)i * 4 + 2] = mxu; proj[(size_t)i * 4 + 3] = mxv;
  const double W = wh[(size_t)m * 2], H = wh[(size_t)m * 2 + 1];
  trunc[(size_t)i * 4] = py_max(0.0, mnu);
  trunc[(size_t)i * 4 + 1] = py_max(0.0, mnv);
  trunc[(size_t)i * 4 + 2] = py_min(W, mxu);
  trunc[(size_t)i * 4 + 3] = py_min(H, mxv);
}

}  // namespace
}  // namespace la3d

extern "C" int la3d_iou_matrix(const double* boxes0, const int64_t* off0, const double* boxes1, const int64_t* off1,
                               const int64_t* out_off, int groups, double* iou, la3d_stream_t stream) {
  using namespace la3d;
  LA3D_REQUIRE(boxes0 && off0 && boxes1 && off1 && out_off && iou, "null pointer");
  LA3D_REQUIRE(groups > 0, "non-positive group count");
  iou_kernel<<<(unsigned)groups, 128, 0, static_cast<cudaStream_t>(stream)>>>(boxes0, off0, boxes1, off1, out_off, iou);
  LA3D_CUDA(cudaGetLastError());
  return LA3D_OK;
}

extern "C" int la3d_box2d_from_corners(const double* corners, const double* K, const int32_t* k_index, const double* wh,
                                       int n, double* proj, double* trunc, la3d_stream_t stream) {
  using namespace la3d;
  LA3D_REQUIRE(corners && K && wh && proj && trunc, "null pointer");
  LA3D_REQUIRE(n > 0, "non-positive box count");
  box2d_kernel<<<(unsigned)((n + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(corners, K, k_index, wh, n,
                                                                                        proj, trunc);
  LA3D_CUDA(cudaGetLastError());
  return LA3D_OK;
}
