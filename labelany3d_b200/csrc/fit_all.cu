// Oriented 3D box per instance from ALL masked pixels (no 500-point subsample) on sm_100a.
//
// The reference draws 500 of a mask's points at random before fitting (src/util_3dbox.py:123-125, an
// unseeded global-RNG draw).  This kernel is estimate_bbox (src/util_3dbox.py:106-178, method='pca') with
// that draw replaced by the identity: every set pixel of the plane takes part, so the result is
// deterministic and uses all the data.  It is the "reduction" form of the path: per instance the
// centroid / covariance sums and the extents are block-wide reductions over the masked pixels.
//
// One CTA of 256 threads per box, two sweeps over the plane's bit words (la3d_mask_scan /
// la3d_rle_decode layout); a warp takes 32 consecutive words and every non-empty one of them is handled by
// the whole warp, lane k taking bit k:
//   sweep 1: pixel -> depth -> exact float64 lift (src/util.py:72 operation order) -> p @ Rg -> NaN-row
//            filter -> n, sum x, sum z, sum xx, sum xz, sum zz, min / max y;  then the closed-form first
//            principal axis of scikit-learn's PCA(2) gives the yaw;
//   sweep 2: the same points rotated by that yaw -> min / max x, z;
//   tail:    float16-rounded corners, back-rotation with the reference's Rg / Rg^T convention, centre,
//            dimensions, R_cam, projected corners and their 2D bounds - the same arithmetic as the tail of
//            fit.cu's kernel.
// Every thread adds its points in a fixed order and the partial sums are combined in a fixed tree, so the
// record does not change from run to run (the two walk orders differ in the last bits of the sums only).
#include <math_constants.h>

#include <cstdlib>

#include "box_tail.cuh"
#include "common.cuh"
#include "prep.cuh"
#include "sink.cuh"

namespace la3d {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr unsigned kFull = 0xffffffffu;

struct AllArgs {
  const float* depth;
  const uint32_t* bits;
  const PrepCamera* cams;   // [images] intrinsics and their inverse (la3d_fit_prepare)
  const double* Rg_pre;     // [boxes][9] ground rotations (la3d_fit_prepare)
  int I, HW, W, words;      // words: bit words per plane (la3d_words_per_plane)
  RecordSink sink;          // sink.cuh
};

__device__ __forceinline__ double dmin(double a, double b) { return b < a ? b : a; }   // NaN in b is ignored
__device__ __forceinline__ double dmax(double a, double b) { return b > a ? b : a; }

struct Smem {
  double red[kWarps][8];
  int ired[kWarps][4];
  double Kinv[9], Kmat[9], Rg[9];
  double yaw, cos_yaw, sin_yaw;
  double rec[LA3D_REC];
};

// kind[k]: 0 sum, 1 min, 2 max.  Fixed combination order: xor tree inside a warp, then warps 0..7.
template <int N>
__device__ __forceinline__ void block_reduce(double (&v)[N], const int (&kind)[N], Smem& sm) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < N; ++k)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double other = __shfl_xor_sync(kFull, v[k], o);
      v[k] = kind[k] == 0 ? v[k] + other : kind[k] == 1 ? dmin(v[k], other) : dmax(v[k], other);
    }
  __syncthreads();
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < N; ++k) sm.red[warp][k] = v[k];
  __syncthreads();
#pragma unroll
  for (int k = 0; k < N; ++k) {
    double acc = sm.red[0][k];
#pragma unroll
    for (int w = 1; w < kWarps; ++w) {
      const double other = sm.red[w][k];
      acc = kind[k] == 0 ? acc + other : kind[k] == 1 ? dmin(acc, other) : dmax(acc, other);
    }
    v[k] = acc;
  }
}

__device__ __forceinline__ void block_sum_int(int (&v)[4], Smem& sm) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < 4; ++k) v[k] = __reduce_add_sync(kFull, v[k]);
  __syncthreads();
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < 4; ++k) sm.ired[warp][k] = v[k];
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    int acc = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) acc += sm.ired[w][k];
    v[k] = acc;
  }
}

// Calls f(rx, ry, rz) for every set pixel of the plane that this thread owns (words tid, tid+256, ...),
// in ascending pixel order: the ground-aligned point p @ Rg of the pixel's lifted depth.
template <typename F>
__device__ __forceinline__ void for_each_point(const AllArgs& a, const uint32_t* __restrict__ plane,
                                               const float* __restrict__ depth_img, const Smem& sm, F&& f) {
  const int used = (a.HW + 31) >> 5;
  for (int w = threadIdx.x; w < used; w += kThreads) {
    uint32_t word = __ldg(plane + w);
    if (!word) continue;
    const int p0 = w << 5;
    const int v0 = p0 / a.W, u0 = p0 - v0 * a.W;
    while (word) {
      const int k = __ffs((int)word) - 1;
      word &= word - 1u;
      const int p = p0 + k;
      if (p >= a.HW) break;                       // padding bits of the last word (always zero)
      int u = u0 + k, v = v0;
      while (u >= a.W) { u -= a.W; ++v; }         // a word may run over the end of a row
      const double d = (double)__ldg(depth_img + p);
      double X, Y, Z;
      lift_pixel_exact(d, (double)u, (double)v, sm.Kinv, X, Y, Z);
      // np.dot(in_pc, Rg): r_j = sum_i p_i Rg[i][j]  (inf * 0 -> NaN drops the row, as in NumPy)
      const double rx = X * sm.Rg[0] + Y * sm.Rg[3] + Z * sm.Rg[6];
      const double ry = X * sm.Rg[1] + Y * sm.Rg[4] + Z * sm.Rg[7];
      const double rz = X * sm.Rg[2] + Y * sm.Rg[5] + Z * sm.Rg[8];
      f(rx, ry, rz);
    }
  }
}

// The same walk, warp-cooperatively: a warp takes 32 consecutive words, and every non-empty one of them (found with
// a ballot) is handled by the whole warp, lane k taking bit k - so the lanes are busy on the dense words inside
// a mask instead of idling while the few lanes that own those words walk 32 bits each.
template <typename F>
__device__ __forceinline__ void for_each_point_coop(const AllArgs& a, const uint32_t* __restrict__ plane,
                                                    const float* __restrict__ depth_img, const Smem& sm, F&& f) {
  const int used = (a.HW + 31) >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int wbase = warp * 32; wbase < used; wbase += kThreads) {        // warp-uniform trip count
    const int w = wbase + lane;
    const uint32_t mine = (w < used) ? __ldg(plane + w) : 0u;
    unsigned nonempty = __ballot_sync(kFull, mine != 0u);
    while (nonempty) {                                                   // warp-uniform
      const int src = __ffs((int)nonempty) - 1;
      nonempty &= nonempty - 1u;
      const uint32_t word = __shfl_sync(kFull, mine, src);
      if (!((word >> lane) & 1u)) continue;
      const int p0 = (wbase + src) << 5;
      const int p = p0 + lane;
      if (p >= a.HW) continue;                                           // padding bits of the last word (always zero)
      int v = p0 / a.W;
      int u = p0 - v * a.W + lane;
      while (u >= a.W) { u -= a.W; ++v; }                                // a word may run over the end of a row
      const double d = (double)__ldg(depth_img + p);
      double X, Y, Z;
      lift_pixel_exact(d, (double)u, (double)v, sm.Kinv, X, Y, Z);
      const double rx = X * sm.Rg[0] + Y * sm.Rg[3] + Z * sm.Rg[6];
      const double ry = X * sm.Rg[1] + Y * sm.Rg[4] + Z * sm.Rg[7];
      const double rz = X * sm.Rg[2] + Y * sm.Rg[5] + Z * sm.Rg[8];
      f(rx, ry, rz);
    }
  }
}

template <bool kCoop>
__global__ void __launch_bounds__(kThreads) fit_all_kernel(AllArgs a) {
  __shared__ Smem sm;
  const int box = blockIdx.x;
  const int tid = threadIdx.x;
  const int img = box / a.I;
  if (tid < 9) sm.Kmat[tid] = __ldg(&a.cams[img].K[tid]);
  else if (tid < 18) sm.Kinv[tid - 9] = __ldg(&a.cams[img].Kinv[tid - 9]);
  else if (tid < 27) sm.Rg[tid - 18] = __ldg(a.Rg_pre + (size_t)box * 9 + (tid - 18));
  __syncthreads();
  const uint32_t* plane = a.bits + (size_t)box * a.words;
  const float* depth_img = a.depth + (size_t)img * a.HW;

  // ---- sweep 1: counts, NaN-row filter (util_3dbox.py:139-143), moments of the XZ footprint, y range ----
  double s[7] = {0.0, 0.0, 0.0, 0.0, 0.0, CUDART_INF, -CUDART_INF};      // sums x, z, xx, xz, zz; min y; max y
  int cnt[4] = {0, 0, 0, 0};                                             // valid, set pixels, inf in x/z, inf in y
  auto sweep = [&](auto&& f) {
    if (kCoop) for_each_point_coop(a, plane, depth_img, sm, f);
    else for_each_point(a, plane, depth_img, sm, f);
  };
  sweep([&](double rx, double ry, double rz) {
    ++cnt[1];
    if (isnan(rx) || isnan(ry) || isnan(rz)) return;
    ++cnt[0];
    cnt[2] |= (int)(isinf(rx) || isinf(rz));
    cnt[3] |= (int)isinf(ry);
    s[0] += rx; s[1] += rz; s[2] += rx * rx; s[3] += rx * rz; s[4] += rz * rz;
    s[5] = dmin(s[5], ry); s[6] = dmax(s[6], ry);
  });
  {
    const int kind[7] = {0, 0, 0, 0, 0, 1, 2};
    block_reduce(s, kind, sm);
    block_sum_int(cnt, sm);
  }
  const int n_valid = cnt[0], n_src = cnt[1], inf_xz = cnt[2], inf_y = cnt[3];
  int status = LA3D_ST_OK;
  if (n_valid == 0) status = LA3D_ST_NO_VALID;
  else if (inf_xz) status = LA3D_ST_NONFINITE;            // scikit-learn's input check raises
  else if (n_valid == 1) status = LA3D_ST_PCA_UNDEFINED;  // PCA(2) needs 2 samples
  if (status != LA3D_ST_OK) {                             // uniform across the CTA
    fill_failed_record(sm.rec, status, n_valid, n_src, kThreads);
    __syncthreads();
  } else {
  // ---- yaw: util_3dbox.py:181-186 with scikit-learn's arithmetic in closed form ----
  if (tid == 0) {
    const double n = (double)n_valid;
    const double mx = s[0] / n, mz = s[1] / n;
    const double ca = (s[2] - n * mx * mx) / (n - 1.0);
    const double cb = (s[3] - n * mx * mz) / (n - 1.0);
    const double cc = (s[4] - n * mz * mz) / (n - 1.0);
    const double theta = 0.5 * atan2(2.0 * cb, ca - cc);
    double vz, vx;
    sincos(theta, &vz, &vx);
    if (fabs(vx) >= fabs(vz)) { if (vx < 0.0) { vx = -vx; vz = -vz; } }
    else if (vz < 0.0) { vx = -vx; vz = -vz; }
    sm.cos_yaw = vx; sm.sin_yaw = vz;
    sm.yaw = atan2(vz, vx);
  }
  __syncthreads();
  const double yaw = sm.yaw, cy_ = sm.cos_yaw, sy_ = sm.sin_yaw;

  // ---- sweep 2: extents at that yaw: rotate_y(yaw) @ pc^T, per-axis min / max (:154-160) ----
  double e[4] = {CUDART_INF, CUDART_INF, -CUDART_INF, -CUDART_INF};      // min x, min z, max x, max z
  sweep([&](double px, double py, double pz) {
    if (isnan(px) || isnan(py) || isnan(pz)) return;
    const double rx = cy_ * px + sy_ * pz, rz = cy_ * pz - sy_ * px;
    e[0] = dmin(e[0], rx); e[2] = dmax(e[2], rx);
    e[1] = dmin(e[1], rz); e[3] = dmax(e[3], rz);
  });
  {
    const int kind[4] = {1, 1, 2, 2};
    block_reduce(e, kind, sm);
  }
  double ext[6] = {e[0], s[5], e[1], e[2], s[6], e[3]};
  if (inf_y) { ext[0] = ext[3] = ext[2] = ext[5] = CUDART_NAN; }        // 0 * inf in the x / z rows of the product
  const double dim[3] = {ext[3] - ext[0], ext[4] - ext[1], ext[5] - ext[2]};
  const double ctr[3] = {(ext[0] + ext[3]) / 2, (ext[1] + ext[4]) / 2, (ext[2] + ext[5]) / 2};

  // ---- tail (box_tail.cuh: util_3dbox.py:165-176, util.py:227-229) ----
  write_box_record(dim, ctr, yaw, cy_, sy_, sm.Rg, sm.Kmat, true, sm.rec, n_valid, n_src, tid);
  }
  sink_acquire(a.sink);
  sink_store(a.sink, (size_t)box, sm.rec, kThreads);
  sink_release(a.sink);
}

}  // namespace
}  // namespace la3d

namespace la3d {
int fit_all_sink(const float* depth, const void* prep, const uint32_t* bits, int B, int I, int H, int W, int method,
                 int yaw_steps, const RecordSink& sink, cudaStream_t stream) {
  LA3D_REQUIRE(depth && prep && bits, "null pointer");
  LA3D_REQUIRE(B > 0 && I > 0 && H > 0 && W > 0, "non-positive shape");
  LA3D_REQUIRE((long long)H * W < (1ll << 30), "image too large");
  LA3D_REQUIRE(I <= 8192, "at most 8192 instances per image");
  LA3D_REQUIRE(method == LA3D_METHOD_PCA, "the all-pixels fit supports method pca");
  (void)yaw_steps;
  const PrepView pv = prep_view(const_cast<void*>(prep), B, I, prep_blocks(I));
  AllArgs a{};
  a.depth = depth; a.bits = bits; a.cams = pv.cams; a.Rg_pre = pv.Rg;
  a.I = I; a.HW = H * W; a.W = W; a.words = (int)la3d_words_per_plane(H, W);
  a.sink = sink;
  // default: warp-cooperative walk of the set bits (lane = bit); LA3D_FITALL_VARIANT=0 selects the first form, one
  // thread per word (measured on B200, config 2: 0.86 ms vs 2.04 ms per step)
  const char* env = getenv("LA3D_FITALL_VARIANT");
  if (!env || atoi(env) != 0) fit_all_kernel<true><<<(unsigned)(B * I), kThreads, 0, stream>>>(a);
  else fit_all_kernel<false><<<(unsigned)(B * I), kThreads, 0, stream>>>(a);
  LA3D_CUDA(cudaGetLastError());
  return LA3D_OK;
}
}  // namespace la3d

extern "C" int la3d_fit_all_points(const float* depth, const void* prep, const uint32_t* bits, int B, int I, int H, int W,
                                   void* records, int rec_f64, la3d_stream_t stream) {
  using namespace la3d;
  LA3D_REQUIRE(records, "null pointer");
  return fit_all_sink(depth, prep, bits, B, I, H, W, LA3D_METHOD_PCA, 0, local_sink(records, rec_f64),
                      static_cast<cudaStream_t>(stream));
}
