// Oriented 3D box per instance from ALL masked pixels (no 500-point subsample) on sm_100a.
//
// The reference draws 500 of a mask's points at random before fitting (src/util_3dbox.py:123-125, an
// unseeded global-RNG draw).  This kernel is estimate_bbox (src/util_3dbox.py:106-178) with that draw
// replaced by the identity: every set pixel of the plane takes part, so the result is deterministic and
// uses all the data.  It is the "reduction" form of the path: per instance the centroid / covariance sums
// and the extents are block-wide reductions over the masked pixels.  All three yaw estimators of the
// sampled path are available (method pca | convex_hull | sweep, util_3dbox.py:146-151 + SURVEY.md 8 a7).
//
// One CTA of 256 threads per box, sweeps over the plane's bit words (la3d_mask_scan / la3d_rle_decode
// layout); a warp takes 32 consecutive words and every non-empty one of them is handled by the whole warp,
// lane k taking bit k.  Every sweep turns pixel -> depth -> exact float64 lift (src/util.py:72 operation
// order) -> p @ Rg -> NaN-row filter.
//   sweep 1: n, shifted sums of x, z, xx, xz, zz (shift = the plane's first point: the covariance is formed
//            from sums of small numbers instead of raw moments), min / max y; for hull / sweep also the 8
//            extreme points of the footprint (max / min of x, x+z, z, z-x with the pixel that attains them).
//   pca:     closed-form first principal axis of scikit-learn's PCA(2) gives the yaw;
//            sweep 2 reduces the extents at that yaw.
//   hull / sweep: the footprint's convex hull can only contain points that are NOT strictly inside a convex
//            polygon of footprint points.  The polygon starts as the octagon of the 8 extreme points; a
//            classification sweep keeps the points outside or on it in shared memory (at most kCap) and, for
//            every polygon edge, tracks the farthest point outside it (QuickHull's step).  If more than kCap
//            points survive, those farthest points join the polygon (8 -> 16 -> 32 -> 64 vertices) and the
//            sweep repeats; a float32 test against the octagon, shrunk by a rounding margin, dismisses the
//            ~90 % of the points that are well inside before any float64 edge test.  The survivors then feed
//            the same yaw code as the sampled path (block-wide gift wrapping + the reference's hull-edge
//            search with its +yaw rotation, or the uniform sweep) and the extents (the extremes of a rotated
//            footprint are hull vertices, so no further sweep over the plane is needed).
//   tail:    box_tail.cuh (float16-rounded corners, Rg / Rg^T convention, projection) -> sink.cuh.
// Every thread adds its points in a fixed order and partial results are combined in a fixed tree, so the
// record does not change from run to run.
#include <math_constants.h>

#include <cstdlib>

#include "box_tail.cuh"
#include "common.cuh"
#include "prep.cuh"
#include "sink.cuh"
#include "yaw_common.cuh"

namespace la3d {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr unsigned kFull = 0xffffffffu;
constexpr int kCap = 2048;       // hull / sweep: candidate points kept in shared memory
constexpr int kMaxPoly = 64;     // vertices of the filter polygon
constexpr int kTrack = 8;        // polygon edges whose farthest outside point one sweep tracks (registers)

struct AllArgs {
  const float* depth;
  const uint32_t* bits;
  const PrepCamera* cams;   // [images] intrinsics and their inverse (la3d_fit_prepare)
  const double* Rg_pre;     // [boxes][9] ground rotations (la3d_fit_prepare)
  int I, HW, W, words;      // words: bit words per plane (la3d_words_per_plane)
  int method, yaw_steps, n_areas;
  RecordSink sink;          // sink.cuh
};

struct Smem {
  double red[kWarps][8];
  int ired[kWarps][8];
  double Kinv[9], Kmat[9], Rg[9];
  double yaw, cos_yaw, sin_yaw;
  double shift_x, shift_z;                       // subtracted from x / z inside the moment sums
  double rec[LA3D_REC];
  // hull / sweep
  double polyx[kMaxPoly], polyz[kMaxPoly];       // filter polygon, counter-clockwise, distinct consecutive vertices
  double farv[kMaxPoly];                         // per edge: farthest point outside it (value = -cross, pixel)
  int farp[kMaxPoly];
  double newx[kMaxPoly], newz[kMaxPoly];         // those points lifted
  float pre[8][4];                               // float32 pre-test: a, b, c - margin per octagon edge
  double pre_cx, pre_cz;
  double wq[kWarps][2];                          // gift wrapping: per-warp candidates
  int wi[kWarps];
  int npoly, npre, n_store, hull_n, added;
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

// kind: 0 sum, 1 min
template <int N>
__device__ __forceinline__ void block_reduce_int(int (&v)[N], const int (&kind)[N], Smem& sm) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < N; ++k) v[k] = kind[k] == 0 ? __reduce_add_sync(kFull, v[k]) : __reduce_min_sync(kFull, v[k]);
  __syncthreads();
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < N; ++k) sm.ired[warp][k] = v[k];
  __syncthreads();
#pragma unroll
  for (int k = 0; k < N; ++k) {
    int acc = sm.ired[0][k];
#pragma unroll
    for (int w = 1; w < kWarps; ++w) acc = kind[k] == 0 ? acc + sm.ired[w][k] : min(acc, sm.ired[w][k]);
    v[k] = acc;
  }
}

// N (value, pixel) pairs: the largest value wins, the smallest pixel among equal values; pixel < 0 = no candidate.
__device__ __forceinline__ bool better(double v, int p, double ov, int op) {
  return op >= 0 && (p < 0 || ov > v || (ov == v && op < p));
}
template <int N>
__device__ __forceinline__ void block_argmax(double (&v)[N], int (&p)[N], Smem& sm) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < N; ++k)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(kFull, v[k], o);
      const int op = __shfl_xor_sync(kFull, p[k], o);
      if (better(v[k], p[k], ov, op)) { v[k] = ov; p[k] = op; }
    }
  __syncthreads();
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < N; ++k) { sm.red[warp][k] = v[k]; sm.ired[warp][k] = p[k]; }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < N; ++k) {
    double bv = sm.red[0][k];
    int bp = sm.ired[0][k];
#pragma unroll
    for (int w = 1; w < kWarps; ++w)
      if (better(bv, bp, sm.red[w][k], sm.ired[w][k])) { bv = sm.red[w][k]; bp = sm.ired[w][k]; }
    v[k] = bv; p[k] = bp;
  }
}

__device__ __forceinline__ int first_strict_min(const double* areas, int n, Smem& sm) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  Best b{CUDART_INF, -1};
  for (int i = threadIdx.x; i < n; i += kThreads) b.offer(areas[i], i);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double oa = __shfl_xor_sync(kFull, b.area, o);
    const int oi = __shfl_xor_sync(kFull, b.idx, o);
    b.offer(oa, oi);
  }
  __syncthreads();
  if (lane == 0) { sm.red[warp][0] = b.area; sm.ired[warp][0] = b.idx; }
  __syncthreads();
  Best r{CUDART_INF, -1};
#pragma unroll
  for (int w = 0; w < kWarps; ++w) r.offer(sm.red[w][0], sm.ired[w][0]);
  return r.idx;
}

// Append (x, z) to the survivor list (at most kCap kept; the counter keeps counting).  Called from divergent code:
// the lanes that are active together share one atomic.
__device__ __forceinline__ void keep_point(Smem& sm, double* __restrict__ sx, double* __restrict__ sz, double x, double z) {
  const unsigned m = __activemask();
  const int lane = threadIdx.x & 31, leader = __ffs((int)m) - 1;
  int base = 0;
  if (lane == leader) base = atomicAdd(&sm.n_store, __popc(m));
  base = __shfl_sync(m, base, leader);
  const int slot = base + __popc(m & ((1u << lane) - 1u));
  if (slot < kCap) { sx[slot] = x; sz[slot] = z; }
}

// The ground-aligned point p @ Rg of one pixel's lifted depth (np.dot(in_pc, Rg): inf * 0 -> NaN drops the
// row, as in NumPy).
__device__ __forceinline__ void aligned_point(const float* __restrict__ depth_img, const Smem& sm, int p, int u, int v,
                                              double& rx, double& ry, double& rz) {
  const double d = (double)__ldg(depth_img + p);
  double X, Y, Z;
  lift_pixel_exact(d, (double)u, (double)v, sm.Kinv, X, Y, Z);
  rx = X * sm.Rg[0] + Y * sm.Rg[3] + Z * sm.Rg[6];
  ry = X * sm.Rg[1] + Y * sm.Rg[4] + Z * sm.Rg[7];
  rz = X * sm.Rg[2] + Y * sm.Rg[5] + Z * sm.Rg[8];
}
__device__ __forceinline__ void aligned_pixel(const AllArgs& a, const float* __restrict__ depth_img, const Smem& sm, int p,
                                              double& rx, double& ry, double& rz) {
  const int v = p / a.W;
  aligned_point(depth_img, sm, p, p - v * a.W, v, rx, ry, rz);
}

// Calls f(p, rx, ry, rz) for every set pixel of the plane that this thread owns (words tid, tid+256, ...),
// in ascending pixel order.
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
      double rx, ry, rz;
      aligned_point(depth_img, sm, p, u, v, rx, ry, rz);
      f(p, rx, ry, rz);
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
      double rx, ry, rz;
      aligned_point(depth_img, sm, p, u, v, rx, ry, rz);
      f(p, rx, ry, rz);
    }
  }
}

// util_3dbox.py:181-186 with scikit-learn's arithmetic in closed form from the SHIFTED sums
// s = {sum (x-sx), sum (z-sz), sum (x-sx)^2, sum (x-sx)(z-sz), sum (z-sz)^2}: the covariance does not depend on the shift.
__device__ __forceinline__ void yaw_from_moments(const double (&s)[7], int n_valid, Smem& sm) {
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

// ---- hull / sweep: the filter polygon ---------------------------------------------------------------
// One thread: drop NaN entries, consecutive duplicates and a last vertex equal to the first from cx / cz [0..n).
__device__ __forceinline__ void set_polygon(Smem& sm, const double* cx, const double* cz, int n) {
  int m = 0;
  for (int i = 0; i < n; ++i) {
    if (!(cx[i] == cx[i])) continue;
    if (m > 0 && sm.polyx[m - 1] == cx[i] && sm.polyz[m - 1] == cz[i]) continue;
    if (m < kMaxPoly) { sm.polyx[m] = cx[i]; sm.polyz[m] = cz[i]; ++m; }
  }
  while (m > 1 && sm.polyx[m - 1] == sm.polyx[0] && sm.polyz[m - 1] == sm.polyz[0]) --m;
  sm.npoly = m;
}

// All threads (contains barriers): the float32 pre-test of the octagon.  A point whose a*x' + b*z' + c is positive
// for every edge (x', z' relative to the polygon's centre, rounded to float32; the margin is folded into c) lies
// strictly inside the octagon whatever the float32 rounding did, hence strictly inside every finer polygon grown
// from it.
__device__ __forceinline__ void build_pretest(Smem& sm) {
  const int n = sm.npoly, d = threadIdx.x;
  if (d == 0) {
    double cx = 0.0, cz = 0.0;
    for (int i = 0; i < n; ++i) { cx += sm.polyx[i]; cz += sm.polyz[i]; }
    sm.pre_cx = cx / n; sm.pre_cz = cz / n;
    sm.npre = (n >= 3 && n <= 8) ? n : 0;
  }
  __syncthreads();
  if (d < 8 && sm.npre) polygon_pretest_edge(sm.polyx, sm.polyz, n, sm.pre_cx, sm.pre_cz, d, sm.pre[d]);
  __syncthreads();
}

// Block-wide gift wrapping of the n candidate points in x[] / z[], counter-clockwise from the lexicographically
// smallest one.  Writes hull[] / sm.hull_n; hull_n = 0 when Qhull would have raised (fewer than 3 vertices).
__device__ void hull_wrap_block(Smem& sm, const double* __restrict__ x, const double* __restrict__ z, int n,
                                unsigned short* __restrict__ hull) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // start vertex: lexicographic minimum (x, then z, then index)
  double bx = CUDART_INF, bz = CUDART_INF;
  int bi = -1;
  auto lexmin = [&](double ox, double oz, int oi) {
    if (oi < 0) return;
    if (bi < 0 || ox < bx || (ox == bx && (oz < bz || (oz == bz && oi < bi)))) { bx = ox; bz = oz; bi = oi; }
  };
  for (int k = threadIdx.x; k < n; k += kThreads) lexmin(x[k], z[k], k);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    lexmin(__shfl_xor_sync(kFull, bx, o), __shfl_xor_sync(kFull, bz, o), __shfl_xor_sync(kFull, bi, o));
  __syncthreads();
  if (lane == 0) { sm.wq[warp][0] = bx; sm.wq[warp][1] = bz; sm.wi[warp] = bi; }
  __syncthreads();
  bi = -1;
  for (int w = 0; w < kWarps; ++w) lexmin(sm.wq[w][0], sm.wq[w][1], sm.wi[w]);
  const double sx0 = bx, sz0 = bz;

  Wrap wr{sx0, sz0, 0.0, 0.0, -1};
  int cur = bi, hn = 0;
  bool closed = false;
  for (int step = 0; step < kCap && cur >= 0; ++step) {          // uniform across the CTA
    if (threadIdx.x == 0) hull[hn] = (unsigned short)cur;
    ++hn;
    wr.qi = -1;
    for (int k = threadIdx.x; k < n; k += kThreads) wr.offer(x[k], z[k], k);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ox = __shfl_xor_sync(kFull, wr.qx, o), oz = __shfl_xor_sync(kFull, wr.qz, o);
      const int oi = __shfl_xor_sync(kFull, wr.qi, o);
      wr.offer(ox, oz, oi);
    }
    __syncthreads();
    if (lane == 0) { sm.wq[warp][0] = wr.qx; sm.wq[warp][1] = wr.qz; sm.wi[warp] = wr.qi; }
    __syncthreads();
    wr.qi = -1;
    for (int w = 0; w < kWarps; ++w) wr.offer(sm.wq[w][0], sm.wq[w][1], sm.wi[w]);
    if (wr.qi < 0) break;                                           // every point coincides
    if (wr.qx == sx0 && wr.qz == sz0) { closed = true; break; }     // wrapped around
    wr.cx = wr.qx; wr.cz = wr.qz; cur = wr.qi;
  }
  __syncthreads();
  if (threadIdx.x == 0) sm.hull_n = (closed && hn >= 3) ? hn : 0;
  __syncthreads();
}

// ---- the kernel ------------------------------------------------------------------------------------
// kSearch: method convex_hull / sweep (dynamic shared memory: survivors, areas, hull); otherwise pca.
template <bool kCoop, bool kSearch>
__global__ void __launch_bounds__(kThreads) fit_all_kernel(AllArgs a) {
  extern __shared__ __align__(16) unsigned char dyn_raw[];
  __shared__ Smem sm;
  const int box = blockIdx.x;
  const int tid = threadIdx.x;
  const int img = box / a.I;
  if (tid < 9) sm.Kmat[tid] = __ldg(&a.cams[img].K[tid]);
  else if (tid < 18) sm.Kinv[tid - 9] = __ldg(&a.cams[img].Kinv[tid - 9]);
  else if (tid < 27) sm.Rg[tid - 18] = __ldg(a.Rg_pre + (size_t)box * 9 + (tid - 18));
  const uint32_t* plane = a.bits + (size_t)box * a.words;
  const float* depth_img = a.depth + (size_t)img * a.HW;

  // the plane's first set pixel: its aligned point is the shift of the moment sums
  {
    const int used = (a.HW + 31) >> 5;
    int first[1] = {0x7fffffff};
    for (int w = tid; w < used; w += kThreads) {
      const uint32_t word = __ldg(plane + w);
      if (word) { first[0] = (w << 5) + __ffs((int)word) - 1; break; }
    }
    const int kind[1] = {1};
    block_reduce_int(first, kind, sm);                      // its barriers also publish Kinv / Kmat / Rg
    if (tid == 0) {
      double sx = 0.0, sy, sz = 0.0;
      if (first[0] < a.HW) {
        aligned_pixel(a, depth_img, sm, first[0], sx, sy, sz);
        if (!(fabs(sx) < CUDART_INF) || !(fabs(sz) < CUDART_INF)) { sx = 0.0; sz = 0.0; }
      }
      sm.shift_x = sx; sm.shift_z = sz;
      sm.n_store = 0;
    }
    __syncthreads();
  }
  const double shx = sm.shift_x, shz = sm.shift_z;

  auto sweep = [&](auto&& f) {
    if (kCoop) for_each_point_coop(a, plane, depth_img, sm, f);
    else for_each_point(a, plane, depth_img, sm, f);
  };

  // ---- sweep 1: counts, NaN-row filter (util_3dbox.py:139-143), moments of the XZ footprint, y range ----
  double s[7] = {0.0, 0.0, 0.0, 0.0, 0.0, CUDART_INF, -CUDART_INF};      // shifted sums x, z, xx, xz, zz; min y; max y
  int cnt[4] = {0, 0, 0, 0};                                             // valid, set pixels, inf in x/z, inf in y
  // hull / sweep: extreme points in the directions x, x+z, z, z-x (maxima 0..3, minima 4..7 stored negated)
  double ev[8];
  int ep[8];
#pragma unroll
  for (int d = 0; d < 8; ++d) { ev[d] = -CUDART_INF; ep[d] = -1; }
  double* sx = reinterpret_cast<double*>(dyn_raw);                       // [kCap] survivors (hull / sweep)
  double* sz = sx + kCap;
  sweep([&](int p, double rx, double ry, double rz) {
    ++cnt[1];
    if (isnan(rx) || isnan(ry) || isnan(rz)) return;
    ++cnt[0];
    cnt[2] |= (int)(isinf(rx) || isinf(rz));
    cnt[3] |= (int)isinf(ry);
    const double qx = rx - shx, qz = rz - shz;
    s[0] += qx; s[1] += qz; s[2] += qx * qx; s[3] += qx * qz; s[4] += qz * qz;
    s[5] = dmin(s[5], ry); s[6] = dmax(s[6], ry);
    if (kSearch) {
      const double f[4] = {rx, rx + rz, rz, rz - rx};
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        if (f[d] > ev[d]) { ev[d] = f[d]; ep[d] = p; }
        if (-f[d] > ev[d + 4]) { ev[d + 4] = -f[d]; ep[d + 4] = p; }
      }
      // a plane with at most kCap points needs no filter: keep them all (the order does not matter)
      if (*reinterpret_cast<volatile int*>(&sm.n_store) <= kCap) keep_point(sm, sx, sz, rx, rz);
    }
  });
  {
    const int kind[7] = {0, 0, 0, 0, 0, 1, 2};
    block_reduce(s, kind, sm);
    const int ikind[4] = {0, 0, 0, 0};
    block_reduce_int(cnt, ikind, sm);
  }
  const int n_valid = cnt[0], n_src = cnt[1], inf_xz = cnt[2], inf_y = cnt[3];
  int status = LA3D_ST_OK;
  const bool bad_method =
      a.method != LA3D_METHOD_PCA && a.method != LA3D_METHOD_CONVEX_HULL && a.method != LA3D_METHOD_SWEEP;
  // same order as the reference: the NaN filter raises first (:142-143), then the method check (:151)
  if (n_valid == 0) status = LA3D_ST_NO_VALID;
  else if (bad_method) status = LA3D_ST_BAD_METHOD;
  else if (inf_xz) status = LA3D_ST_NONFINITE;            // scikit-learn's input check raises
  else if (n_valid == 1 && a.method != LA3D_METHOD_SWEEP) status = LA3D_ST_PCA_UNDEFINED;  // PCA(2) needs 2 samples

  int n_cand = 0;                                          // survivors in sx / sz (hull / sweep)
  if (kSearch && status == LA3D_ST_OK) {
    n_cand = n_valid;
    if (n_valid > kCap) {
      // ---- the filter polygon: octagon of the extreme points, refined until at most kCap points survive ----
      block_argmax(ev, ep, sm);
      if (tid < 8) {
        double rx = CUDART_NAN, ry, rz = CUDART_NAN;
        if (ep[tid] >= 0) aligned_pixel(a, depth_img, sm, ep[tid], rx, ry, rz);
        sm.newx[tid] = rx; sm.newz[tid] = rz;
      }
      __syncthreads();
      // direction order 0, 45, ..., 315 degrees: max x, max x+z, max z, max z-x, min x, min x+z, min z, min z-x
      if (tid == 0) set_polygon(sm, sm.newx, sm.newz, 8);
      __syncthreads();
      build_pretest(sm);
      for (int level = 0;; ++level) {                      // uniform across the CTA
        const int np = sm.npoly;
        const int groups = np < 3 ? 1 : (np + kTrack - 1) / kTrack;
        int survivors = 0;
        for (int g = 0; g < groups; ++g) {
          // group 0 classifies, stores and tracks edges 0..7; further groups only track (needed if we refine)
          if (g > 0 && survivors <= kCap) break;
          double fv[kTrack];
          int fp[kTrack];
#pragma unroll
          for (int t = 0; t < kTrack; ++t) { fv[t] = 0.0; fp[t] = -1; }
          if (g == 0) {
            __syncthreads();
            if (tid == 0) sm.n_store = 0;
            __syncthreads();
          }
          int mine = 0;
          const int npre = sm.npre;
          const double pcx = sm.pre_cx, pcz = sm.pre_cz;
          sweep([&](int p, double rx, double ry, double rz) {
            if (isnan(rx) || isnan(ry) || isnan(rz)) return;
            if (npre) {
              const float fx = (float)(rx - pcx), fz = (float)(rz - pcz);
              bool in = true;
#pragma unroll
              for (int d = 0; d < 8; ++d) in = in && (fmaf(sm.pre[d][0], fx, fmaf(sm.pre[d][1], fz, sm.pre[d][2])) > 0.f);
              if (in) return;
            }
            bool outside = np < 3;                         // no polygon (collinear extremes): everything survives
            for (int d = 0; d < np && np >= 3; ++d) {
              const int e = (d + 1 == np) ? 0 : d + 1;
              const double ox = sm.polyx[d], oz = sm.polyz[d];
              const double ex = sm.polyx[e] - ox, ez = sm.polyz[e] - oz;
              const double cr = ex * (rz - oz) - ez * (rx - ox);
              if (!(cr > 0.0)) outside = true;
              const int t = d - g * kTrack;
              if (cr < 0.0 && t >= 0 && t < kTrack) {
#pragma unroll
                for (int u = 0; u < kTrack; ++u)
                  if (u == t && -cr > fv[u]) { fv[u] = -cr; fp[u] = p; }
              }
            }
            if (outside && g == 0) {
              ++mine;
              keep_point(sm, sx, sz, rx, rz);
            }
          });
          if (g == 0) {
            int c[1] = {mine};
            const int ikind[1] = {0};
            block_reduce_int(c, ikind, sm);
            survivors = c[0];
          }
          if (g > 0 || survivors > kCap) {
            block_argmax(fv, fp, sm);
            if (tid < kTrack && g * kTrack + tid < np) { sm.farv[g * kTrack + tid] = fv[tid]; sm.farp[g * kTrack + tid] = fp[tid]; }
            __syncthreads();
          }
        }
        n_cand = survivors;
        if (survivors <= kCap) break;
        // refine: the farthest point outside an edge joins the polygon after the edge's first vertex
        if (tid < np) {
          double rx = CUDART_NAN, ry, rz = CUDART_NAN;
          if (sm.farp[tid] >= 0) aligned_pixel(a, depth_img, sm, sm.farp[tid], rx, ry, rz);
          sm.newx[tid] = rx; sm.newz[tid] = rz;
        }
        __syncthreads();
        if (tid == 0) {
          double cx[2 * kMaxPoly], cz[2 * kMaxPoly];
          int m = 0, added = 0;
          for (int d = 0; d < np; ++d) {
            cx[m] = sm.polyx[d]; cz[m] = sm.polyz[d]; ++m;
            if (np >= 3 && sm.farp[d] >= 0 && np + added < kMaxPoly) { cx[m] = sm.newx[d]; cz[m] = sm.newz[d]; ++m; ++added; }
          }
          set_polygon(sm, cx, cz, m);
          sm.added = added;
        }
        __syncthreads();
        if (sm.added == 0 || level >= 8) { status = LA3D_ST_TOO_MANY; break; }   // > kCap points on the hull's rim
      }
    }
  }

  if (status != LA3D_ST_OK) {                             // uniform across the CTA
    __syncthreads();
    fill_failed_record(sm.rec, status, n_valid, n_src, kThreads);
    __syncthreads();
  } else {
  // ---- yaw -------------------------------------------------------------------------------------------
  bool have_trig = false, hull_fallback = false;
  if (!kSearch || a.method == LA3D_METHOD_PCA) {
    if (tid == 0) yaw_from_moments(s, n_valid, sm);
    have_trig = true;
  } else {
    double* areas = sz + kCap;                                         // [n_areas]
    unsigned short* hull = reinterpret_cast<unsigned short*>(areas + a.n_areas);   // [kCap]
    __syncthreads();                                                   // the survivors are in place
    if (a.method == LA3D_METHOD_CONVEX_HULL) {
      hull_wrap_block(sm, sx, sz, n_cand, hull);
      const int hn = sm.hull_n;
      if (hn == 0) {
        if (tid == 0) yaw_from_moments(s, n_valid, sm);                // Qhull would have raised: fall back to PCA
        have_trig = true;
        hull_fallback = true;
      } else {
        // util_3dbox.py:202-218: one thread per hull edge, first strict minimum of the area
        auto edge_angle = [&](int e) {
          const int i0 = hull[e], i1 = hull[(e + 1 == hn) ? 0 : e + 1];
          return atan2(sz[i1] - sz[i0], sx[i1] - sx[i0]);
        };
        for (int e = tid; e < hn; e += kThreads) areas[e] = rect_area(sx, sz, hull, hn, edge_angle(e), 0);
        __syncthreads();
        const int e = first_strict_min(areas, hn, sm);
        if (tid == 0) sm.yaw = e < 0 ? 0.0 : edge_angle(e);
      }
    } else {
      // uniform sweep: yaw_k = k*(pi/2)/K, first strict minimum of dx*dz
      const int K = a.yaw_steps;
      if (inf_y) {
        if (tid == 0) sm.yaw = 0.0;                                    // 0*inf = NaN in every area: nothing beats +inf
      } else {
        for (int c = tid; c < K; c += kThreads) areas[c] = rect_area(sx, sz, nullptr, n_cand, sweep_angle(c, K), 1);
        __syncthreads();
        const int c = first_strict_min(areas, K, sm);
        if (tid == 0) sm.yaw = c < 0 ? 0.0 : sweep_angle(c, K);
      }
    }
  }
  __syncthreads();
  const double yaw = sm.yaw;
  double sy_, cy_;
  if (have_trig) { sy_ = sm.sin_yaw; cy_ = sm.cos_yaw; }
  else sincos(yaw, &sy_, &cy_);

  // ---- extents at that yaw: rotate_y(yaw) @ pc^T, per-axis min / max (:154-160) ----
  double e[4] = {CUDART_INF, CUDART_INF, -CUDART_INF, -CUDART_INF};      // min x, min z, max x, max z
  auto extent = [&](double px, double pz) {
    const double rx = cy_ * px + sy_ * pz, rz = cy_ * pz - sy_ * px;
    e[0] = dmin(e[0], rx); e[2] = dmax(e[2], rx);
    e[1] = dmin(e[1], rz); e[3] = dmax(e[3], rz);
  };
  if (kSearch) {
    // every extreme of a rotated footprint is a hull vertex, and the hull's vertices are among the survivors
    for (int k = tid; k < n_cand; k += kThreads) extent(sx[k], sz[k]);
  } else {
    sweep([&](int, double px, double py, double pz) {
      if (isnan(px) || isnan(py) || isnan(pz)) return;
      extent(px, pz);
    });
  }
  {
    const int kind[4] = {1, 1, 2, 2};
    block_reduce(e, kind, sm);
  }
  double ext[6] = {e[0], s[5], e[1], e[2], s[6], e[3]};
  if (inf_y) { ext[0] = ext[3] = ext[2] = ext[5] = CUDART_NAN; }        // 0 * inf in the x / z rows of the product
  const double dim[3] = {ext[3] - ext[0], ext[4] - ext[1], ext[5] - ext[2]};
  const double ctr[3] = {(ext[0] + ext[3]) / 2, (ext[1] + ext[4]) / 2, (ext[2] + ext[5]) / 2};

  // ---- tail (box_tail.cuh: util_3dbox.py:165-176, util.py:227-229) ----
  write_box_record(dim, ctr, yaw, cy_, sy_, sm.Rg, sm.Kmat, true, sm.rec, n_valid, n_src, tid,
                   hull_fallback ? (double)LA3D_FLAG_HULL_FALLBACK : 0.0);
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
  LA3D_REQUIRE(method != LA3D_METHOD_SWEEP || (yaw_steps > 0 && yaw_steps <= 16384), "sweep needs 1 <= yaw_steps <= 16384");
  const PrepView pv = prep_view(const_cast<void*>(prep), B, I, prep_blocks(I));
  AllArgs a{};
  a.depth = depth; a.bits = bits; a.cams = pv.cams; a.Rg_pre = pv.Rg;
  a.I = I; a.HW = H * W; a.W = W; a.words = (int)la3d_words_per_plane(H, W);
  a.method = method; a.yaw_steps = yaw_steps;
  a.sink = sink;
  // an unknown method id takes the pca kernel, which reports LA3D_ST_BAD_METHOD per box like the sampled path
  const bool search = method == LA3D_METHOD_CONVEX_HULL || method == LA3D_METHOD_SWEEP;
  a.n_areas = !search ? 0 : method == LA3D_METHOD_SWEEP ? ((yaw_steps + 1) & ~1) : kCap;
  const size_t dyn = !search ? 0 : (size_t)kCap * 16 + (size_t)a.n_areas * 8 + (size_t)kCap * 2;
  const unsigned grid = (unsigned)(B * I);
  // default: warp-cooperative walk of the set bits (lane = bit); LA3D_FITALL_VARIANT=0 selects the first form, one
  // thread per word (measured on B200, config 2: 0.86 ms vs 2.04 ms per step)
  const char* env = getenv("LA3D_FITALL_VARIANT");
  const bool coop = !env || atoi(env) != 0;
  if (search) {
    LA3D_REQUIRE(dyn + sizeof(Smem) <= 220 * 1024, "yaw sweep too large for shared memory");
    if (coop) {
      LA3D_CUDA(cudaFuncSetAttribute(fit_all_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
      fit_all_kernel<true, true><<<grid, kThreads, dyn, stream>>>(a);
    } else {
      LA3D_CUDA(cudaFuncSetAttribute(fit_all_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
      fit_all_kernel<false, true><<<grid, kThreads, dyn, stream>>>(a);
    }
  } else if (coop) {
    fit_all_kernel<true, false><<<grid, kThreads, 0, stream>>>(a);
  } else {
    fit_all_kernel<false, false><<<grid, kThreads, 0, stream>>>(a);
  }
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

extern "C" int la3d_fit_all_points_to(const float* depth, const void* prep, const uint32_t* bits, int B, int I, int H,
                                      int W, int method, int yaw_steps, const la3d_sink* sink, la3d_stream_t stream) {
  using namespace la3d;
  RecordSink rs;
  if (int rc = sink_from_public(sink, &rs)) return rc;
  if (int rc = publish_previous_epoch(rs, static_cast<cudaStream_t>(stream))) return rc;
  return fit_all_sink(depth, prep, bits, B, I, H, W, method, yaw_steps, rs, static_cast<cudaStream_t>(stream));
}
