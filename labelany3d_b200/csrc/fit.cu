// Oriented 3D box per instance on sm_100a.  Replaces estimate_bbox and its yaw
// estimators (src/util_3dbox.py:106-224 of the reference) and the corner
// reprojection (src/util.py:227-229, src/tools/combine_results.py:238-246).
//
// One CTA of 64 threads fits one box from at most 500 points, all in float64:
//   1. point source
//      - scanned masks: rank r (the reference's random row of pts[mask]) -> binary
//        search over the per-chunk prefix sums -> quarter of the chunk (packed byte
//        counts) -> bit select in that quarter's 4 words -> pixel (v,u) -> depth gather
//        -> exact lift (src/util.py:72 operation order); a few samples of a thread (kGroup,
//        measured: 2 - 4) move through that chain together;
//      - explicit points (the reference's own call pattern, util_3dbox.py:269-278).
//   2. ground alignment p @ Rg, NaN-row filter (util_3dbox.py:128-143);
//   3. yaw: PCA closed form | convex-hull edge search | uniform sweep.  Hull and sweep
//      first discard every point strictly inside the octagon of the footprint's 8
//      extreme points (such a point is never extreme in any direction); the sweep
//      then evaluates its candidates over the survivors, the hull method gift-wraps
//      them (one warp) and visits the hull edges;
//   4. extents at that yaw, float16-rounded corners, back-rotation with the
//      reference's Rg / Rg^T convention, centre, dimensions, R_cam (:154-176);
//   5. projection of the 8 corners and their 2D bounds.
// No tensor cores: there is no dense contraction here.
#include <math_constants.h>

#include <cstdlib>

#include "box_tail.cuh"
#include "common.cuh"
#include "prep.cuh"
#include "sink.cuh"
#include "yaw_common.cuh"

namespace la3d {
namespace {

constexpr int kThreads = 64;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxPts = 512;   // >= LA3D_SUBSAMPLE
constexpr unsigned kFull = 0xffffffffu;

struct FitArgs {
  // scanned-mask source
  const float* depth;
  const uint32_t* bits;
  const uint32_t* chunk_counts;
  const int32_t* ranks;
  const PrepCamera* cams;   // [images] intrinsics and their inverse (la3d_fit_prepare)
  const double* Rg_pre;     // [boxes][9] ground rotations (la3d_fit_prepare)
  int I, HW, W, chunks;
  int search_top;           // largest power of two <= chunks - 1 (find_chunks)
  uint32_t w_magic;         // p / W for p < 2^30 as umulhi(p, w_magic) >> w_shift; 0: W == 1
  int w_shift;
  // explicit-point source
  const double* pts;
  const int64_t* offsets;
  const int32_t* sample_idx;
  // common
  const double* K;        // [images or boxes][9] intrinsics; may be null for explicit points
  const double* ground;   // [boxes][3] or null
  int method, yaw_steps, n_areas;
  long long* phase_clocks;  // debug (la3d_debug_fit_clocks): [boxes][8] clock64 stamps of thread 0, or null
  int box0;               // first box of this launch (a step may be cut into several launches over one set of buffers)
  RecordSink sink;        // one local buffer, or the gathered buffers of all ranks (peer memory); sink.cuh
};

// Static shared memory.  The y coordinate is not kept: only its minimum / maximum matter (they do
// not depend on the yaw) and those are reduced on the fly.
struct Smem {
  double x[kMaxPts], z[kMaxPts];               // ground-aligned footprint; x = NaN marks a dropped row
  double red[kWarps][8];
  double Kinv[9], Kmat[9], Rg[9];
  int oct_idx[8];                              // extreme points of the footprint (octagon, CCW): indices into x / z
  double octx[8], octz[8], oct_cx, oct_cz;     // their coordinates and centre
  float pre[8][4];                             // float32 inside test per octagon edge (yaw_common.cuh)
  double yaw, cos_yaw, sin_yaw;
  int ired[kWarps][4];
  int cand_n, hull_n;
  double* rec;                                 // [64] record under construction (aliases the prefix table)
  unsigned short* cand;                        // points that can be extreme in some direction
  unsigned short* hull;                        // hull vertices (indices into x/z), counter-clockwise
};

// Dynamic shared memory: [areas: n_areas doubles][prefix table (chunks+1 u32) / record (64 doubles)]
//                        [cand: 512 u16 (hull, sweep)][hull: 512 u16 (hull)]
__host__ __device__ inline size_t region_b_bytes(int chunks, bool scanned) {
  size_t pref = scanned ? ((size_t)chunks + 1) * 4 : 0;
  size_t need = pref > (size_t)LA3D_REC * 8 ? pref : (size_t)LA3D_REC * 8;
  return (need + 15) & ~(size_t)15;
}

// ---- block-wide reductions; the result is broadcast to every thread -------------
template <int N, typename Op>
__device__ __forceinline__ void block_reduce(double (&v)[N], Smem& sm, Op op) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < N; ++k)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] = op(v[k], __shfl_xor_sync(kFull, v[k], o), k);
  __syncthreads();
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < N; ++k) sm.red[warp][k] = v[k];
  __syncthreads();
#pragma unroll
  for (int k = 0; k < N; ++k) {
    double acc = sm.red[0][k];
#pragma unroll
    for (int w = 1; w < kWarps; ++w) acc = op(acc, sm.red[w][k], k);
    v[k] = acc;
  }
}

template <int N>
__device__ __forceinline__ void block_sum_int(int (&v)[N], Smem& sm) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < N; ++k) v[k] = __reduce_add_sync(kFull, v[k]);
  __syncthreads();
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < N; ++k) sm.ired[warp][k] = v[k];
  __syncthreads();
#pragma unroll
  for (int k = 0; k < N; ++k) {
    int acc = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) acc += sm.ired[w][k];
    v[k] = acc;
  }
}

struct OpAdd { __device__ double operator()(double a, double b, int) const { return a + b; } };
// entries 0..2 are minima, 3..5 maxima
struct OpMinMax { __device__ double operator()(double a, double b, int k) const { return k < 3 ? dmin(a, b) : dmax(a, b); } };
// entries 0..3 are maxima, 4..7 minima
struct OpMax4Min4 { __device__ double operator()(double a, double b, int k) const { return k < 4 ? dmax(a, b) : dmin(a, b); } };

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

// ---- yaw estimators -----------------------------------------------------------------

// util_3dbox.py:181-186 with scikit-learn's arithmetic in closed form:
// C = (X^T X - n mu mu^T)/(n-1); first eigenvector angle theta = atan2(2b, a-c)/2; the
// component of larger magnitude is made positive (svd_flip, v-based).
__device__ void yaw_pca(Smem& sm, int nsel, int n_valid) {
  double s[5] = {0, 0, 0, 0, 0};
  for (int k = threadIdx.x; k < nsel; k += kThreads) {
    const double px = sm.x[k], pz = sm.z[k];
    if (px == px) { s[0] += px; s[1] += pz; s[2] += px * px; s[3] += px * pz; s[4] += pz * pz; }
  }
  block_reduce(s, sm, OpAdd());
  if (threadIdx.x == 0) {
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
    // yaw = atan2(vz, vx) (:186); (vx, vz) already is (cos yaw, sin yaw), which is all the
    // extents need - the angle itself is only reported
    sm.cos_yaw = vx; sm.sin_yaw = vz;
    sm.yaw = atan2(vz, vx);
  }
}

// Octagon filter.  The 8 points that maximise x, x+z, z, z-x, -x, -x-z, -z, x-z (in that,
// counter-clockwise, order of direction) span a convex polygon of input points; a point
// strictly inside it can never attain the maximum of a linear functional over the cloud, so
// rectangle extents (sweep, hull-edge search) and the hull itself only need the others.
// Degenerate edges (repeated extreme points) are skipped, which only keeps more points.
__device__ void octagon_candidates(Smem& sm, int nsel) {
  const int lane = threadIdx.x & 31;
  double e[8] = {-CUDART_INF, -CUDART_INF, -CUDART_INF, -CUDART_INF, CUDART_INF, CUDART_INF, CUDART_INF, CUDART_INF};
  for (int k = threadIdx.x; k < nsel; k += kThreads) {
    const double px = sm.x[k], pz = sm.z[k];       // dropped rows are NaN and lose every comparison
    e[0] = dmax(e[0], px); e[1] = dmax(e[1], px + pz); e[2] = dmax(e[2], pz); e[3] = dmax(e[3], pz - px);
    e[4] = dmin(e[4], px); e[5] = dmin(e[5], px + pz); e[6] = dmin(e[6], pz); e[7] = dmin(e[7], pz - px);
  }
  block_reduce(e, sm, OpMax4Min4());
  if (threadIdx.x == 0) sm.cand_n = 0;
  if (threadIdx.x < 8) sm.oct_idx[threadIdx.x] = 0x7fffffff;
  __syncthreads();
  // any point attaining an extreme serves as that octagon vertex; among ties the smallest index is elected, so
  // that both coordinates come from ONE input point whatever the thread schedule
  for (int k = threadIdx.x; k < nsel; k += kThreads) {
    const double px = sm.x[k], pz = sm.z[k];
    const double f[8] = {px, px + pz, pz, pz - px, px, px + pz, pz, pz - px};
#pragma unroll
    for (int d = 0; d < 8; ++d)
      if (f[d] == e[d]) atomicMin(&sm.oct_idx[d], k);
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    const int k = min(sm.oct_idx[threadIdx.x], nsel - 1);        // always elected: e[d] is attained by a valid point
    sm.octx[threadIdx.x] = sm.x[k]; sm.octz[threadIdx.x] = sm.z[k];
    __syncwarp(0xffu);
    double cx = 0.0, cz = 0.0;                                    // every edge thread sums the centre itself (same order)
    for (int d = 0; d < 8; ++d) { cx += sm.octx[d]; cz += sm.octz[d]; }
    cx /= 8; cz /= 8;
    if (threadIdx.x == 0) { sm.oct_cx = cx; sm.oct_cz = cz; }
    polygon_pretest_edge(sm.octx, sm.octz, 8, cx, cz, threadIdx.x, sm.pre[threadIdx.x]);
  }
  __syncthreads();
  float pa[8], pb[8], pc[8];
#pragma unroll
  for (int d = 0; d < 8; ++d) { pa[d] = sm.pre[d][0]; pb[d] = sm.pre[d][1]; pc[d] = sm.pre[d][2]; }
  const double ccx = sm.oct_cx, ccz = sm.oct_cz;
  for (int k0 = 0; k0 < nsel; k0 += kThreads) {
    const int k = k0 + threadIdx.x;
    bool keep = false;
    if (k < nsel) {
      const double px = sm.x[k], pz = sm.z[k];
      if (px == px) {
        // float32 test with a rounding margin: passing every edge = strictly inside the octagon
        const float fx = (float)(px - ccx), fz = (float)(pz - ccz);
        bool inside = true;
#pragma unroll
        for (int d = 0; d < 8; ++d) inside = inside && (fmaf(pa[d], fx, fmaf(pb[d], fz, pc[d])) > 0.f);
        keep = !inside;
      }
    }
    const unsigned bal = __ballot_sync(kFull, keep);
    int base = 0;
    if (lane == 0 && bal) base = atomicAdd(&sm.cand_n, __popc(bal));
    base = __shfl_sync(kFull, base, 0);
    if (keep) sm.cand[base + __popc(bal & ((1u << lane) - 1u))] = (unsigned short)k;
  }
  __syncthreads();
}

// Gift wrapping of the valid XZ points by warp 0, counter-clockwise from the
// lexicographically smallest point; collinear points are skipped (strict hull, like Qhull's
// vertex list).  Writes sm.hull / sm.hull_n; hull_n = 0 when Qhull would have raised
// (fewer than 3 vertices: coincident or collinear points).
__device__ void hull_wrap(Smem& sm) {
  const int lane = threadIdx.x & 31;
  const int nc = sm.cand_n;
  // start vertex: lexicographic minimum (x, then z, then index)
  double bx = CUDART_INF, bz = CUDART_INF;
  int bi = -1;
  auto lexmin = [&](double ox, double oz, int oi) {
    if (oi < 0) return;
    if (bi < 0 || ox < bx || (ox == bx && (oz < bz || (oz == bz && oi < bi)))) { bx = ox; bz = oz; bi = oi; }
  };
  for (int c = lane; c < nc; c += 32) {
    const int k = sm.cand[c];
    lexmin(sm.x[k], sm.z[k], k);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    lexmin(__shfl_xor_sync(kFull, bx, o), __shfl_xor_sync(kFull, bz, o), __shfl_xor_sync(kFull, bi, o));
  const double sx0 = bx, sz0 = bz;

  Wrap wr{sx0, sz0, 0.0, 0.0, -1};
  int cur = bi, hn = 0;
  bool closed = false;
  for (int step = 0; step < kMaxPts && cur >= 0; ++step) {
    if (lane == 0) sm.hull[hn] = (unsigned short)cur;
    ++hn;
    wr.qi = -1;
    for (int c = lane; c < nc; c += 32) {
      const int k = sm.cand[c];
      wr.offer(sm.x[k], sm.z[k], k);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ox = __shfl_xor_sync(kFull, wr.qx, o), oz = __shfl_xor_sync(kFull, wr.qz, o);
      const int oi = __shfl_xor_sync(kFull, wr.qi, o);
      wr.offer(ox, oz, oi);
    }
    // Wrap::offer is not symmetric under rounding (on nearly collinear points two lanes of a butterfly pair can
    // each keep their own candidate), so the lanes may end with different winners: lane 0's is THE next vertex.
    // Without this the lanes would leave the loop at different steps and the next shuffle would never complete.
    wr.qx = __shfl_sync(kFull, wr.qx, 0); wr.qz = __shfl_sync(kFull, wr.qz, 0); wr.qi = __shfl_sync(kFull, wr.qi, 0);
    if (wr.qi < 0) break;                                           // every point coincides
    if (wr.qx == sx0 && wr.qz == sz0) { closed = true; break; }     // wrapped around
    wr.cx = wr.qx; wr.cz = wr.qz; cur = wr.qi;
  }
  if (lane == 0) sm.hull_n = (closed && hn >= 3) ? hn : 0;
}

__device__ __forceinline__ double edge_angle(const Smem& sm, int hn, int e) {
  const int i0 = sm.hull[e], i1 = sm.hull[(e + 1 == hn) ? 0 : e + 1];
  return atan2(sm.z[i1] - sm.z[i0], sm.x[i1] - sm.x[i0]);
}

// ---- gather: rank -> pixel -> depth -> camera point ---------------------------------------
// Branch-free search for the chunk holding rank r: the largest c with pref[c] <= r.  Uniform step count, so that
// the searches of a thread's samples interleave (independent shared-memory loads per step).  `top` is the largest
// power of two <= chunks - 1 (0 for one chunk): the first probe folds the non-power-of-two remainder, every later
// probe pos + half stays below chunks.
template <int G>
__device__ __forceinline__ void find_chunks(const uint32_t* pref, int chunks, int top, const uint32_t (&r)[G],
                                            int (&chunk)[G], uint32_t (&rem)[G]) {
  int pos[G];
  const int first = chunks - 1 - top;                           // >= 0; pref[first + top] is the last chunk's offset
#pragma unroll
  for (int j = 0; j < G; ++j) pos[j] = (top && pref[first + 1] <= r[j]) ? first + 1 : 0;   // invariant: pref[pos] <= r
  // after the first probe the answer lies in [pos, pos + top): halve
  for (int half = top >> 1; half > 0; half >>= 1) {
#pragma unroll
    for (int j = 0; j < G; ++j) {
      const int nxt = pos[j] + half;
      pos[j] = (pref[nxt] <= r[j]) ? nxt : pos[j];
    }
  }
#pragma unroll
  for (int j = 0; j < G; ++j) { chunk[j] = pos[j]; rem[j] = r[j] - pref[pos[j]]; }
}

// ---- the kernel ------------------------------------------------------------------------
// kGroup: samples per thread that move through the dependent loads of the gather together.  Measured on B200
// (tools/fit_bench.py; 36-step sweep / pca / 360-step sweep / hull, us): 2048 images: 8 in flight 406 / 255 / 829 / 996,
// 4: 390 / 233 / 796 / 924, 2: 381 / 248 / 759 / 874; 256 images: 8: 64 / 41 / 167 / 147, 4: 60 / 37 / 158 / 141,
// 2: 59 / 35 / 160 / 140.  The gather is bound by sector throughput, not by latency, and fewer live registers help
// everything after it.  launch_fit picks 2, or 4 for pca on launches of 8192 boxes or more; LA3D_FIT_GROUP overrides.
//
// kSortDepth (depth maps left in pinned HOST memory, all samples of the box in one pass: kGroup * kThreads >= 512):
// a scattered read over PCIe costs one 128-byte line request per distinct line per warp instruction, whatever the
// loads in flight (tools/gather_ceiling.py: 343 M requests/s, 2x / 4x the samples when 2 / 4 lanes share a line).  The
// box's pixel indices are therefore sorted (bitonic, in the footprint arrays, which are still free), consecutive
// sorted samples are read by consecutive lanes, and the values return to sample order through shared memory - the
// points, their order and every later operation are unchanged.
template <bool kScanned, int kGroup, bool kSortDepth = false>
__global__ void __launch_bounds__(kThreads, 14) fit_kernel(FitArgs a) {
  static_assert(!kSortDepth || (kScanned && kGroup * kThreads >= kMaxPts && kMaxPts == 512), "one pass over 512 samples");
  extern __shared__ __align__(16) unsigned char dyn_raw[];
  __shared__ Smem sm;
  double* areas = reinterpret_cast<double*>(dyn_raw);                               // [n_areas]
  unsigned char* region_b = dyn_raw + (size_t)a.n_areas * 8;
  uint32_t* pref = reinterpret_cast<uint32_t*>(region_b);                           // [chunks+1], scanned source
  if (threadIdx.x == 0) {
    sm.rec = reinterpret_cast<double*>(region_b);                                   // used after pref is dead
    sm.cand = reinterpret_cast<unsigned short*>(region_b + region_b_bytes(a.chunks, kScanned));
    sm.hull = sm.cand + kMaxPts;
  }

  const int box = blockIdx.x + a.box0;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  auto stamp = [&](int i) { if (a.phase_clocks && tid == 0) a.phase_clocks[(size_t)box * 8 + i] = clock64(); };
  stamp(0);
  const int img = kScanned ? box / a.I : box;

  if (kScanned) {
    // camera and ground rotation were prepared per image / per box (la3d_fit_prepare)
    if (tid < 9) sm.Kmat[tid] = __ldg(&a.cams[img].K[tid]);
    else if (tid < 18) sm.Kinv[tid - 9] = __ldg(&a.cams[img].Kinv[tid - 9]);
    else if (tid < 27) sm.Rg[tid - 18] = __ldg(a.Rg_pre + (size_t)box * 9 + (tid - 18));
  } else {
    // explicit points: one thread each, overlapped with the loads below
    if (tid == 0 && a.K) {
#pragma unroll
      for (int i = 0; i < 9; ++i) sm.Kmat[i] = a.K[(size_t)img * 9 + i];
    }
    if (tid == 32) ground_rotation(a.ground ? a.ground + (size_t)box * 3 : nullptr, sm.Rg);
  }

  long long n_src;
  const double* src_pts = nullptr;
  const uint32_t* cc = nullptr;
  if (kScanned) {
    // exclusive prefix over the chunk totals.  The totals come in with COALESCED loads (thread t takes chunks t, t+64,
    // ...) into the table itself; each thread then scans a contiguous run of the table in shared memory and writes
    // the exclusive prefix back in place.  (Reading the contiguous runs straight from global memory, the first form,
    // touched 32 sectors per warp load and read every word twice: at 1536 x 1536, 4608 chunks per plane, that was six
    // times the sector requests of the whole sample gather.  Keeping the quarter-count words in shared memory as
    // well measured 40 % slower: 2.4 KB more per CTA push the SM's carve-out to 228 KB and leave the L1 28 KB.)
    cc = a.chunk_counts + (size_t)box * a.chunks;
    for (int c = tid; c < a.chunks; c += kThreads) pref[c] = __dp4a(__ldg(cc + c), 0x01010101u, 0u);
    __syncthreads();
    const int per = (a.chunks + kThreads - 1) / kThreads;
    const int c_lo = min(tid * per, a.chunks), c_hi = min(c_lo + per, a.chunks);
    uint32_t run = 0;
    for (int c = c_lo; c < c_hi; ++c) run += pref[c];
    uint32_t incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t up = __shfl_up_sync(kFull, incl, o);
      if (lane >= o) incl += up;
    }
    if (lane == 31) sm.ired[warp][0] = (int)incl;
    __syncthreads();
    uint32_t base = incl - run;
    for (int w = 0; w < warp; ++w) base += (uint32_t)sm.ired[w][0];
    for (int c = c_lo; c < c_hi; ++c) { const uint32_t t = pref[c]; pref[c] = base; base += t; }
    if (tid == kThreads - 1) pref[a.chunks] = base;
    n_src = 0;
  } else {
    const long long o0 = a.offsets[box];
    n_src = a.offsets[box + 1] - o0;
    src_pts = a.pts + (size_t)o0 * 3;
  }
  __syncthreads();
  if (kScanned) {
    n_src = pref[a.chunks];                    // = counts[box]
    // up to here only outputs of the scan launch were read; the ranks come from the sampler, which may
    // still be running when this kernel was launched as its programmatic dependent
    pdl_wait();
  }

  stamp(1);                                    // prologue (cameras, chunk prefix) done
  const bool subsample = n_src > LA3D_SUBSAMPLE;
  int status = LA3D_ST_OK;
  if (!kScanned && subsample && !a.sample_idx) status = LA3D_ST_TOO_MANY;
  const bool bad_method =
      a.method != LA3D_METHOD_PCA && a.method != LA3D_METHOD_CONVEX_HULL && a.method != LA3D_METHOD_SWEEP;
  const int nsel = status ? 0 : (int)min((long long)LA3D_SUBSAMPLE, n_src);

  // ---- gather + lift + ground alignment + NaN-row filter ---------------------------
  // The kGroup samples of a thread advance through the dependent loads together.
  int n_valid = 0, inf_xz = 0, inf_y = 0;
  double y_lo = CUDART_INF, y_hi = -CUDART_INF;
  const int search_top = a.search_top;
  // (kSortDepth: every thread runs the one pass, the block barriers of the sort sit inside it)
  const int k_end = kSortDepth ? (nsel > 0 ? kMaxPts : 0) : nsel;
  for (int k0 = tid; k0 < k_end; k0 += kThreads * kGroup) {
    double X[kGroup], Y[kGroup], Z[kGroup];
    if (kScanned) {
      uint32_t r[kGroup];
#pragma unroll
      for (int j = 0; j < kGroup; ++j) {
        const int k = k0 + j * kThreads;
        r[j] = (k < nsel && subsample) ? (uint32_t)__ldg(a.ranks + (size_t)box * LA3D_SUBSAMPLE + k) : (uint32_t)min(k, nsel - 1);
      }
      int chunk[kGroup];
      uint32_t rem[kGroup];
      find_chunks<kGroup>(pref, a.chunks, search_top, r, chunk, rem);
      uint32_t qw[kGroup];
#pragma unroll
      for (int j = 0; j < kGroup; ++j) qw[j] = __ldg(cc + chunk[j]);
      uint4 w4[kGroup];
      int base_px[kGroup];
#pragma unroll
      for (int j = 0; j < kGroup; ++j) {
        // quarter of the chunk: byte q of qw = set pixels of its 128-pixel quarter q
        int q = 0;
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          const uint32_t cnt = (qw[j] >> (8 * t)) & 0xffu;
          const bool next = (q == t) && (rem[j] >= cnt);
          rem[j] -= next ? cnt : 0u;
          q += next ? 1 : 0;
        }
        base_px[j] = chunk[j] * kChunkPx + q * 128;
        w4[j] = __ldg(reinterpret_cast<const uint4*>(a.bits + ((size_t)box * a.chunks + chunk[j]) * kChunkWords) + q);
      }
      float d[kGroup];
      int p[kGroup];
#pragma unroll
      for (int j = 0; j < kGroup; ++j) {
        const uint32_t w[4] = {w4[j].x, w4[j].y, w4[j].z, w4[j].w};
        uint32_t word = w[0];
        int wi = 0;
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          const uint32_t pc = __popc(w[t]);
          const bool next = (wi == t) && (rem[j] >= pc);
          rem[j] -= next ? pc : 0u;
          wi += next ? 1 : 0;
          word = next ? w[t + 1] : word;
        }
        p[j] = min(base_px[j] + wi * 32 + (int)__fns(word, 0, (int)rem[j] + 1), a.HW - 1);
        if (!kSortDepth) d[j] = __ldg(a.depth + (size_t)img * a.HW + p[j]);
      }
      if (kSortDepth) {
        unsigned long long* keys = reinterpret_cast<unsigned long long*>(sm.x);   // [512] (pixel << 32 | sample)
        float* dsh = reinterpret_cast<float*>(sm.z);                               // [512] depth by sample
#pragma unroll
        for (int j = 0; j < kGroup; ++j) {
          const int k = k0 + j * kThreads;
          if (k < kMaxPts) keys[k] = k < nsel ? ((unsigned long long)(uint32_t)p[j] << 32) | (uint32_t)k : ~0ull;
        }
        __syncthreads();
        for (int size = 2; size <= kMaxPts; size <<= 1) {
          for (int stride = size >> 1; stride > 0; stride >>= 1) {
#pragma unroll
            for (int t = tid; t < kMaxPts / 2; t += kThreads) {
              const int lo = ((t & ~(stride - 1)) << 1) | (t & (stride - 1)), hi = lo | stride;
              const unsigned long long ka = keys[lo], kb = keys[hi];
              if ((ka > kb) == ((lo & size) == 0)) { keys[lo] = kb; keys[hi] = ka; }
            }
            __syncthreads();
          }
        }
        unsigned long long key[kMaxPts / kThreads];
#pragma unroll
        for (int j = 0; j < kMaxPts / kThreads; ++j) key[j] = keys[j * kThreads + tid];
        float got[kMaxPts / kThreads];
#pragma unroll
        for (int j = 0; j < kMaxPts / kThreads; ++j)
          got[j] = key[j] != ~0ull ? __ldg(a.depth + (size_t)img * a.HW + (uint32_t)(key[j] >> 32)) : 0.f;
#pragma unroll
        for (int j = 0; j < kMaxPts / kThreads; ++j)
          if (key[j] != ~0ull) dsh[(uint32_t)key[j] & (kMaxPts - 1)] = got[j];
        __syncthreads();
#pragma unroll
        for (int j = 0; j < kGroup; ++j) {
          const int k = k0 + j * kThreads;
          d[j] = k < nsel ? dsh[k] : 0.f;
        }
        __syncthreads();                       // dsh / keys alias the footprint arrays written below
      }
#pragma unroll
      for (int j = 0; j < kGroup; ++j) {
        const int v = a.w_magic ? (int)(__umulhi((uint32_t)p[j], a.w_magic) >> a.w_shift) : p[j], u = p[j] - v * a.W;
        lift_pixel_exact((double)d[j], (double)u, (double)v, sm.Kinv, X[j], Y[j], Z[j]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < kGroup; ++j) {
        const int k = k0 + j * kThreads;
        X[j] = Y[j] = Z[j] = 0.0;
        if (k < nsel) {
          const long long row = subsample ? (long long)a.sample_idx[(size_t)box * LA3D_SUBSAMPLE + k] : (long long)k;
          X[j] = src_pts[row * 3]; Y[j] = src_pts[row * 3 + 1]; Z[j] = src_pts[row * 3 + 2];
        }
      }
    }
#pragma unroll
    for (int j = 0; j < kGroup; ++j) {
      const int k = k0 + j * kThreads;
      if (k < nsel) {
        // np.dot(in_pc, Rg): r_j = sum_i p_i Rg[i][j]  (inf * 0 -> NaN drops the row, as in NumPy)
        double rx = X[j] * sm.Rg[0] + Y[j] * sm.Rg[3] + Z[j] * sm.Rg[6];
        const double ry = X[j] * sm.Rg[1] + Y[j] * sm.Rg[4] + Z[j] * sm.Rg[7];
        const double rz = X[j] * sm.Rg[2] + Y[j] * sm.Rg[5] + Z[j] * sm.Rg[8];
        if (isnan(rx) || isnan(ry) || isnan(rz)) {
          rx = CUDART_NAN;
        } else {
          ++n_valid;
          inf_xz |= (int)(isinf(rx) || isinf(rz));
          inf_y |= (int)isinf(ry);
          y_lo = dmin(y_lo, ry); y_hi = dmax(y_hi, ry);
        }
        sm.x[k] = rx; sm.z[k] = rz;
      }
    }
  }
  {
    int r[3] = {n_valid, inf_xz, inf_y};
    block_sum_int(r, sm);           // its barriers also publish sm.x / y / z
    n_valid = r[0]; inf_xz = r[1]; inf_y = r[2];
  }
  stamp(2);                                    // gather + lift + alignment done
  if (status == LA3D_ST_OK) {
    // same order as the reference: the NaN filter raises first (:142-143), then the method check (:151)
    if (n_valid == 0) status = LA3D_ST_NO_VALID;
    else if (bad_method) status = LA3D_ST_BAD_METHOD;
    else if (inf_xz) status = LA3D_ST_NONFINITE;      // scikit-learn's input check raises; Qhull fails first and falls back to it
    else if (n_valid == 1 && a.method != LA3D_METHOD_SWEEP) status = LA3D_ST_PCA_UNDEFINED;   // PCA(2) needs 2 samples
  }

  if (status != LA3D_ST_OK) {                          // uniform across the CTA
    fill_failed_record(sm.rec, status, n_valid, n_src, kThreads);
    __syncthreads();
  } else {
  // ---- yaw ---------------------------------------------------------------------------
  bool have_trig = false;          // yaw_pca leaves cos / sin of the yaw in shared memory
  bool hull_fallback = false;
  if (a.method == LA3D_METHOD_PCA) {
    yaw_pca(sm, nsel, n_valid);
    have_trig = true;
  } else {
    octagon_candidates(sm, nsel);
    const int nc = sm.cand_n;
    stamp(3);                                  // octagon filter done
    if (a.method == LA3D_METHOD_CONVEX_HULL) {
      if (warp == 0) hull_wrap(sm);
      __syncthreads();
      const int hn = sm.hull_n;
      if (hn == 0) {
        yaw_pca(sm, nsel, n_valid);                    // Qhull would have raised (QH6214 / QH6154): fall back
        have_trig = true;
        hull_fallback = true;                          // reported in the record (LA3D_FLAG_HULL_FALLBACK)
      } else {
        // util_3dbox.py:202-218: one thread per hull edge, first strict minimum of the area
        for (int e = tid; e < hn; e += kThreads) areas[e] = rect_area(sm.x, sm.z, sm.hull, hn, edge_angle(sm, hn, e), 0);
        __syncthreads();
        const int e = first_strict_min(areas, hn, sm);
        if (tid == 0) sm.yaw = e < 0 ? 0.0 : edge_angle(sm, hn, e);
      }
    } else {
      // uniform sweep: yaw_k = k*(pi/2)/K, first strict minimum of dx*dz
      const int K = a.yaw_steps;
      if (inf_y) {
        if (tid == 0) sm.yaw = 0.0;                    // 0*inf = NaN in every area: nothing beats +inf
      } else {
        for (int c = tid; c < K; c += kThreads) areas[c] = rect_area(sm.x, sm.z, sm.cand, nc, sweep_angle(c, K), 1);
        __syncthreads();
        const int c = first_strict_min(areas, K, sm);
        if (tid == 0) sm.yaw = c < 0 ? 0.0 : sweep_angle(c, K);
      }
    }
  }
  __syncthreads();
  stamp(4);                                    // yaw known
  const double yaw = sm.yaw;

  // ---- extents at that yaw: rotate_y(yaw) @ pc^T, per-axis min / max (:154-160) ------
  double sy_, cy_;
  if (have_trig) { sy_ = sm.sin_yaw; cy_ = sm.cos_yaw; }
  else sincos(yaw, &sy_, &cy_);
  double ext[6] = {CUDART_INF, y_lo, CUDART_INF, -CUDART_INF, y_hi, -CUDART_INF};
  for (int k = tid; k < nsel; k += kThreads) {
    const double px = sm.x[k], pz = sm.z[k];         // a dropped row is NaN in x and loses every comparison
    const double rx = cy_ * px + sy_ * pz, rz = cy_ * pz - sy_ * px;
    ext[0] = dmin(ext[0], rx); ext[3] = dmax(ext[3], rx);
    ext[2] = dmin(ext[2], rz); ext[5] = dmax(ext[5], rz);
  }
  block_reduce(ext, sm, OpMinMax());
  if (inf_y) { ext[0] = ext[3] = ext[2] = ext[5] = CUDART_NAN; }      // 0 * inf in the x / z rows of the product
  const double dim[3] = {ext[3] - ext[0], ext[4] - ext[1], ext[5] - ext[2]};
  const double ctr[3] = {(ext[0] + ext[3]) / 2, (ext[1] + ext[4]) / 2, (ext[2] + ext[5]) / 2};

  stamp(5);                                    // extents done
  write_box_record(dim, ctr, yaw, cy_, sy_, sm.Rg, sm.Kmat, kScanned || a.K != nullptr, sm.rec, n_valid, n_src, tid,
                   hull_fallback ? (double)LA3D_FLAG_HULL_FALLBACK : 0.0);
  }
  stamp(6);                                    // record built
  sink_acquire(a.sink);
  sink_store(a.sink, (size_t)box, sm.rec, kThreads);
  sink_release(a.sink);
  stamp(7);
}

long long* g_phase_clocks = nullptr;

int launch_fit(bool scanned, FitArgs& a, int nboxes, cudaStream_t s, bool pdl = false) {
  a.phase_clocks = scanned ? g_phase_clocks : nullptr;
  a.n_areas = a.method == LA3D_METHOD_SWEEP ? (a.yaw_steps > 0 ? a.yaw_steps : 1)
            : a.method == LA3D_METHOD_CONVEX_HULL ? kMaxPts : 0;
  a.n_areas = (a.n_areas + 1) & ~1;                       // keep what follows 16-byte aligned
  const size_t lists = a.method == LA3D_METHOD_PCA ? 0 : (size_t)kMaxPts * 2 * (a.method == LA3D_METHOD_SWEEP ? 1 : 2);
  const size_t dyn = (size_t)a.n_areas * 8 + region_b_bytes(a.chunks, scanned) + lists;
  if (dyn + sizeof(Smem) > 227 * 1024) {
    set_error("la3d fit: image or yaw sweep too large for shared memory (%zu bytes needed)", dyn + sizeof(Smem));
    return LA3D_EINVAL;
  }
  static const int group_env = getenv("LA3D_FIT_GROUP") ? atoi(getenv("LA3D_FIT_GROUP")) : 0;
  // measured choice (see the kernel): 2 samples in flight, except the pca method on many-wave launches; depth maps
  // left in pinned HOST memory (the end-to-end path: 32-byte reads over PCIe, microseconds each) want all 8 in flight
  bool host_depth = false;
  if (scanned && a.depth) {
    cudaPointerAttributes attr{};
    if (cudaPointerGetAttributes(&attr, a.depth) == cudaSuccess) host_depth = attr.type == cudaMemoryTypeHost;
    else cudaGetLastError();
  }
  const int group = group_env ? group_env : host_depth ? 8 : (a.method == LA3D_METHOD_PCA && nboxes >= 8192 ? 4 : 2);
  auto launch = [&](auto kernel) -> int {
    if (dyn > 16 * 1024) LA3D_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    LA3D_CUDA(launch_pdl(kernel, dim3((unsigned)nboxes), dim3(kThreads), dyn, s, pdl && scanned, a));
    return LA3D_OK;
  };
  int rc;
  static const bool sort_env = !(getenv("LA3D_FIT_SORT_DEPTH") && atoi(getenv("LA3D_FIT_SORT_DEPTH")) == 0);
  if (scanned && host_depth && group == 8 && sort_env) rc = launch(fit_kernel<true, 8, true>);
  else if (scanned) rc = group == 2 ? launch(fit_kernel<true, 2>) : group == 8 ? launch(fit_kernel<true, 8>) : launch(fit_kernel<true, 4>);
  else rc = launch(fit_kernel<false, 4>);
  if (rc) return rc;
  LA3D_CUDA(cudaGetLastError());
  return LA3D_OK;
}

}  // namespace
}  // namespace la3d

namespace la3d {
// The boxes of images [b0, b0 + Bp) of a batch of B; every pointer is the whole batch's.
int fit_scanned_sink(const float* depth, const void* prep, const uint32_t* bits, const uint32_t* chunk_counts,
                     const int32_t* ranks, int B, int I, int H, int W, int method, int yaw_steps,
                     const RecordSink& sink, cudaStream_t stream, bool pdl, int b0, int Bp) {
  LA3D_REQUIRE(depth && prep && bits && chunk_counts && ranks, "null pointer");
  LA3D_REQUIRE(B > 0 && I > 0 && H > 0 && W > 0, "non-positive shape");
  LA3D_REQUIRE((long long)H * W < (1ll << 30), "image too large");
  LA3D_REQUIRE(method != LA3D_METHOD_SWEEP || (yaw_steps > 0 && yaw_steps <= 16384), "sweep needs 1 <= yaw_steps <= 16384");
  if (Bp < 0) Bp = B - b0;
  LA3D_REQUIRE(b0 >= 0 && Bp > 0 && b0 + Bp <= B, "bad image range");
  const PrepView pv = prep_view(const_cast<void*>(prep), B, I, prep_blocks(I));
  FitArgs a{};
  a.depth = depth; a.bits = bits; a.chunk_counts = chunk_counts; a.ranks = ranks;
  a.cams = pv.cams; a.Rg_pre = pv.Rg;
  a.I = I; a.HW = H * W; a.W = W; a.chunks = (int)la3d_chunks_per_plane(H, W);
  a.search_top = 0;
  while (a.chunks - 1 >= 2 * (a.search_top ? a.search_top : 1) || (a.search_top == 0 && a.chunks > 1)) a.search_top = a.search_top ? a.search_top * 2 : 1;
  // exact division of any p < 2^30 by W: M = ceil(2^(30+s) / W), s = max(2, ceil(log2 W)), q = (p * M) >> (30 + s)
  if (W > 1) {
    int sh = 2;
    while ((1ll << sh) < W) ++sh;
    a.w_magic = (uint32_t)(((1ull << (30 + sh)) + (uint64_t)W - 1) / (uint64_t)W);
    a.w_shift = sh - 2;
  }
  a.method = method; a.yaw_steps = yaw_steps;
  a.box0 = b0 * I;
  a.sink = sink;
  return launch_fit(true, a, Bp * I, stream, pdl);
}
}  // namespace la3d

extern "C" void la3d_debug_fit_clocks(long long* clocks) { la3d::g_phase_clocks = clocks; }

extern "C" int la3d_fit_scanned(const float* depth, const void* prep, const uint32_t* bits,
                                const uint32_t* chunk_counts, const int32_t* ranks, int B, int I, int H, int W,
                                int method, int yaw_steps, void* records, int rec_f64, la3d_stream_t stream) {
  using namespace la3d;
  LA3D_REQUIRE(records, "null pointer");
  return fit_scanned_sink(depth, prep, bits, chunk_counts, ranks, B, I, H, W, method, yaw_steps,
                          local_sink(records, rec_f64), static_cast<cudaStream_t>(stream), false);
}

extern "C" int la3d_fit_scanned_to(const float* depth, const void* prep, const uint32_t* bits,
                                   const uint32_t* chunk_counts, const int32_t* ranks, int B, int I, int H, int W,
                                   int method, int yaw_steps, const la3d_sink* sink, la3d_stream_t stream) {
  using namespace la3d;
  RecordSink rs;
  if (int rc = sink_from_public(sink, &rs)) return rc;
  if (int rc = publish_previous_epoch(rs, static_cast<cudaStream_t>(stream))) return rc;
  return fit_scanned_sink(depth, prep, bits, chunk_counts, ranks, B, I, H, W, method, yaw_steps, rs,
                          static_cast<cudaStream_t>(stream), false);
}

extern "C" int la3d_fit_points(const double* pts, const int64_t* offsets, const int32_t* sample_idx, const double* K,
                               const double* ground, int nboxes, int method, int yaw_steps, void* records,
                               int rec_f64, la3d_stream_t stream) {
  using namespace la3d;
  LA3D_REQUIRE(pts && offsets && records, "null pointer");
  LA3D_REQUIRE(nboxes > 0, "non-positive box count");
  LA3D_REQUIRE(method != LA3D_METHOD_SWEEP || (yaw_steps > 0 && yaw_steps <= 16384), "sweep needs 1 <= yaw_steps <= 16384");
  FitArgs a{};
  a.pts = pts; a.offsets = offsets; a.sample_idx = sample_idx;
  a.K = K; a.ground = ground; a.method = method; a.yaw_steps = yaw_steps;
  a.sink = local_sink(records, rec_f64);
  return launch_fit(false, a, nboxes, static_cast<cudaStream_t>(stream));
}
