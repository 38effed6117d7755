// Oriented 3D box per instance on sm_100a.  Replaces estimate_bbox and its yaw
// estimators (src/util_3dbox.py:106-224 of the reference) and the corner
// reprojection (src/util.py:227-229, src/tools/combine_results.py:238-246).
//
// One CTA of 128 threads fits one box from at most 500 points, all in float64:
//   1. point source
//      - scanned masks: rank r (the reference's random row of pts[mask]) ->
//        binary search over the per-chunk prefix sums -> bit select in the
//        chunk's 16 words -> pixel (v,u) -> depth gather -> exact lift
//        (src/util.py:72 operation order);
//      - explicit points (the reference's own call pattern, util_3dbox.py:269-278).
//   2. ground alignment p @ Rg, NaN-row filter (util_3dbox.py:128-143);
//   3. yaw: PCA closed form | convex-hull edge search | uniform sweep;
//   4. extents at that yaw, float16-rounded corners, back-rotation with the
//      reference's Rg / Rg^T convention, centre, dimensions, R_cam (:154-176);
//   5. projection of the 8 corners and their 2D bounds.
// No tensor cores: there is no dense contraction here.
#include <math_constants.h>

#include "common.cuh"

namespace la3d {
namespace {

constexpr int kThreads = 128;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxPts = 512;   // >= LA3D_SUBSAMPLE
constexpr unsigned kFull = 0xffffffffu;

struct FitArgs {
  // scanned-mask source
  const float* depth;
  const uint32_t* bits;
  const uint16_t* chunk_counts;
  const int32_t* counts;
  const int32_t* ranks;
  int I, HW, W, chunks;
  // explicit-point source
  const double* pts;
  const int64_t* offsets;
  const int32_t* sample_idx;
  // common
  const double* K;        // [images or boxes][9] intrinsics; may be null for explicit points
  const double* ground;   // [boxes][3] or null
  int method, yaw_steps, n_areas;
  void* records;
  int rec_f64;
};

struct Smem {
  double x[kMaxPts], y[kMaxPts], z[kMaxPts];   // ground-aligned points; x = NaN marks a dropped row
  double hx[kMaxPts], hz[kMaxPts];             // hull vertices, counter-clockwise
  double red[kWarps][8];
  double Kinv[9], Kmat[9], Rg[9];
  double rec[LA3D_REC];
  int ired[kWarps][4];
};

// ---- block-wide reductions; the result is broadcast to every thread -------------
template <int N, typename Op>
__device__ __forceinline__ void block_reduce(double (&v)[N], Smem& sm, Op op) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < N; ++k)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] = op(v[k], __shfl_xor_sync(kFull, v[k], o));
  __syncthreads();
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < N; ++k) sm.red[warp][k] = v[k];
  __syncthreads();
#pragma unroll
  for (int k = 0; k < N; ++k) {
    double acc = sm.red[0][k];
#pragma unroll
    for (int w = 1; w < kWarps; ++w) acc = op(acc, sm.red[w][k]);
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

struct OpAdd { __device__ double operator()(double a, double b) const { return a + b; } };
struct OpMin { __device__ double operator()(double a, double b) const { return fmin(a, b); } };
struct OpMax { __device__ double operator()(double a, double b) const { return fmax(a, b); } };

// (area, index) pairs ordered like the reference's `if area < min_area` loop: the
// smallest area wins, the earliest index among equals; NaN and +inf never win.
struct Best {
  double area;
  int idx;   // -1 = nothing qualified yet
  __device__ void offer(double a, int i) {
    if (!(a < CUDART_INF)) return;
    if (idx < 0 || a < area || (a == area && i < idx)) { area = a; idx = i; }
  }
};

__device__ __forceinline__ int first_strict_min(const double* areas, int n, Smem& sm) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  Best b{CUDART_INF, -1};
  for (int i = threadIdx.x; i < n; i += kThreads) b.offer(areas[i], i);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double oa = __shfl_xor_sync(kFull, b.area, o);
    const int oi = __shfl_xor_sync(kFull, b.idx, o);
    if (oi >= 0) b.offer(oa, oi);
  }
  __syncthreads();
  if (lane == 0) { sm.red[warp][0] = b.area; sm.ired[warp][0] = b.idx; }
  __syncthreads();
  Best r{CUDART_INF, -1};
#pragma unroll
  for (int w = 0; w < kWarps; ++w)
    if (sm.ired[w][0] >= 0) r.offer(sm.red[w][0], sm.ired[w][0]);
  return r.idx;
}

// Rg of util_3dbox.py:128-134: Rodrigues rotation taking (0,-1,0) to the ground
// normal, which is flipped first when dot((0,-1,0), g) <= 0.  0/0 -> NaN when the
// two are parallel, exactly like the reference.
__device__ void ground_rotation(const double* g, double* Rg) {
  if (!g) {
#pragma unroll
    for (int i = 0; i < 9; ++i) Rg[i] = (i % 4 == 0) ? 1.0 : 0.0;
    return;
  }
  double g0 = g[0], g1 = g[1], g2 = g[2];
  const double dotp = 0.0 * g0 + (-1.0) * g1 + 0.0 * g2;
  if (dotp <= 0.0) { g0 = -g0; g1 = -g1; g2 = -g2; }
  const double a0 = 0.0, a1 = -1.0, a2 = 0.0;
  const double nb = sqrt(g0 * g0 + g1 * g1 + g2 * g2);
  double b0 = g0, b1 = g1, b2 = g2;
  if (nb != 0.0) { b0 = g0 / nb; b1 = g1 / nb; b2 = g2 / nb; }
  const double ax = a1 * b2 - a2 * b1, ay = a2 * b0 - a0 * b2, az = a0 * b1 - a1 * b0;
  const double cosang = a0 * b0 + a1 * b1 + a2 * b2;
  const double S[9] = {0.0, -az, ay, az, 0.0, -ax, -ay, ax, 0.0};
  const double nrm = sqrt(ax * ax + ay * ay + az * az);
  const double nn = nrm * nrm;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const double s2 = S[i * 3 + 0] * S[0 * 3 + j] + S[i * 3 + 1] * S[1 * 3 + j] + S[i * 3 + 2] * S[2 * 3 + j];
      Rg[i * 3 + j] = ((i == j ? 1.0 : 0.0) + S[i * 3 + j]) + s2 * (1.0 - cosang) / nn;
    }
}

// r-th (0-based) set pixel of a plane in row-major order.  pref = exclusive prefix
// sums of the chunk counts, pref[chunks] = N.
__device__ __forceinline__ int select_pixel(const uint32_t* __restrict__ plane_bits, const uint32_t* pref, int chunks,
                                            uint32_t r) {
  int lo = 0, hi = chunks;            // invariant: pref[lo] <= r < pref[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (pref[mid] <= r) lo = mid; else hi = mid;
  }
  uint32_t rem = r - pref[lo];
  const uint4* w4 = reinterpret_cast<const uint4*>(plane_bits + (size_t)lo * kChunkWords);
  uint32_t w[kChunkWords];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint4 t = __ldg(w4 + q);
    w[4 * q] = t.x; w[4 * q + 1] = t.y; w[4 * q + 2] = t.z; w[4 * q + 3] = t.w;
  }
  int word = 0;
  uint32_t sel = w[0];
#pragma unroll
  for (int q = 0; q < kChunkWords - 1; ++q) {
    const uint32_t pc = __popc(w[q]);
    if (word == q && rem >= pc) { rem -= pc; word = q + 1; sel = w[q + 1]; }
  }
  return lo * kChunkPx + word * 32 + (int)__fns(sel, 0, (int)rem + 1);
}

// ---- yaw estimators -----------------------------------------------------------------

// util_3dbox.py:181-186 with scikit-learn's arithmetic in closed form:
// C = (X^T X - n mu mu^T)/(n-1); first eigenvector angle theta = atan2(2b, a-c)/2; the
// component of larger magnitude is made positive (svd_flip, v-based).
__device__ double yaw_pca(Smem& sm, int nsel, int n_valid) {
  double s[5] = {0, 0, 0, 0, 0};
  for (int k = threadIdx.x; k < nsel; k += kThreads) {
    const double px = sm.x[k], pz = sm.z[k];
    if (px == px) { s[0] += px; s[1] += pz; s[2] += px * px; s[3] += px * pz; s[4] += pz * pz; }
  }
  block_reduce(s, sm, OpAdd());
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
  return atan2(vz, vx);
}

// New feature: yaw_k = k*(pi/2)/K, area of the XZ bounding
// rectangle after yaw_matrix(yaw_k), first strict minimum.  One warp per candidate.
__device__ double yaw_sweep(Smem& sm, int nsel, int K, double* areas, bool inf_y) {
  if (inf_y || K <= 0) return 0.0;     // 0*inf = NaN in every area: nothing beats +inf
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = warp; c < K; c += kWarps) {
    const double ang = __ddiv_rn(__dmul_rn((double)c, CUDART_PIO2), (double)K);
    double s, co;
    sincos(ang, &s, &co);
    double mnx = CUDART_INF, mxx = -CUDART_INF, mnz = CUDART_INF, mxz = -CUDART_INF;
    for (int k = lane; k < nsel; k += 32) {
      const double px = sm.x[k], pz = sm.z[k];      // dropped rows carry NaN and fall out of fmin/fmax
      const double rx = co * px + s * pz, rz = co * pz - s * px;
      mnx = fmin(mnx, rx); mxx = fmax(mxx, rx); mnz = fmin(mnz, rz); mxz = fmax(mxz, rz);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mnx = fmin(mnx, __shfl_xor_sync(kFull, mnx, o));
      mxx = fmax(mxx, __shfl_xor_sync(kFull, mxx, o));
      mnz = fmin(mnz, __shfl_xor_sync(kFull, mnz, o));
      mxz = fmax(mxz, __shfl_xor_sync(kFull, mxz, o));
    }
    if (lane == 0) areas[c] = (mxx - mnx) * (mxz - mnz);
  }
  __syncthreads();
  const int best = first_strict_min(areas, K, sm);
  return best < 0 ? 0.0 : __ddiv_rn(__dmul_rn((double)best, CUDART_PIO2), (double)K);
}

// Gift wrapping of the valid XZ points, counter-clockwise from the lexicographically
// smallest point; collinear points are skipped (strict hull, like Qhull's vertex list).
// Returns the vertex count, or 0 when Qhull would have raised (< 3 vertices).
struct Wrap {
  double cx, cz, qx, qz;
  __device__ bool none() const { return qx == cx && qz == cz; }
  __device__ void offer(double px, double pz) {
    if (!(px == px) || (px == cx && pz == cz)) return;
    if (none()) { qx = px; qz = pz; return; }
    const double cr = (qx - cx) * (pz - cz) - (qz - cz) * (px - cx);
    if (cr < 0.0) { qx = px; qz = pz; }          // p is clockwise of cur->q: q cannot be next
    else if (cr == 0.0) {
      const double dq = (qx - cx) * (qx - cx) + (qz - cz) * (qz - cz);
      const double dp = (px - cx) * (px - cx) + (pz - cz) * (pz - cz);
      if (dp > dq) { qx = px; qz = pz; }
    }
  }
};

__device__ int hull_wrap(Smem& sm, int nsel) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double bx = CUDART_INF, bz = CUDART_INF;
  auto lexmin = [&](double ox, double oz) { if (ox < bx || (ox == bx && oz < bz)) { bx = ox; bz = oz; } };
  for (int k = threadIdx.x; k < nsel; k += kThreads) lexmin(sm.x[k], sm.z[k]);   // NaN never compares less
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lexmin(__shfl_xor_sync(kFull, bx, o), __shfl_xor_sync(kFull, bz, o));
  __syncthreads();
  if (lane == 0) { sm.red[warp][0] = bx; sm.red[warp][1] = bz; }
  __syncthreads();
  bx = sm.red[0][0]; bz = sm.red[0][1];
#pragma unroll
  for (int w = 1; w < kWarps; ++w) lexmin(sm.red[w][0], sm.red[w][1]);
  const double sx0 = bx, sz0 = bz;

  Wrap wr{sx0, sz0, sx0, sz0};
  int hn = 0;
  bool closed = false;
  for (int step = 0; step < kMaxPts; ++step) {
    if (threadIdx.x == 0) { sm.hx[hn] = wr.cx; sm.hz[hn] = wr.cz; }
    ++hn;
    wr.qx = wr.cx; wr.qz = wr.cz;
    for (int k = threadIdx.x; k < nsel; k += kThreads) wr.offer(sm.x[k], sm.z[k]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ox = __shfl_xor_sync(kFull, wr.qx, o), oz = __shfl_xor_sync(kFull, wr.qz, o);
      wr.offer(ox, oz);
    }
    __syncthreads();
    if (lane == 0) { sm.red[warp][0] = wr.qx; sm.red[warp][1] = wr.qz; }
    __syncthreads();
    wr.qx = sm.red[0][0]; wr.qz = sm.red[0][1];
#pragma unroll
    for (int w = 1; w < kWarps; ++w) wr.offer(sm.red[w][0], sm.red[w][1]);
    if (wr.none()) break;                                          // every point coincides
    if (wr.qx == sx0 && wr.qz == sz0) { closed = true; break; }    // wrapped around
    wr.cx = wr.qx; wr.cz = wr.qz;
  }
  __syncthreads();
  return (closed && hn >= 3) ? hn : 0;
}

// util_3dbox.py:202-218 over the hull edges.  The footprint is rotated by +yaw here
// (rot_2d) although the box is later built with rotate_y(yaw) = -yaw in XZ: kept as is.
// The extremes of a linear functional over the cloud are attained at hull vertices,
// so only those are visited.
__device__ double yaw_hull_edges(Smem& sm, int hn, double* areas) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int e = warp; e < hn; e += kWarps) {
    const int e1 = (e + 1 == hn) ? 0 : e + 1;
    const double ang = atan2(sm.hz[e1] - sm.hz[e], sm.hx[e1] - sm.hx[e]);
    double s, c;
    sincos(ang, &s, &c);
    double mnx = CUDART_INF, mxx = -CUDART_INF, mnz = CUDART_INF, mxz = -CUDART_INF;
    for (int k = lane; k < hn; k += 32) {
      const double px = sm.hx[k], pz = sm.hz[k];
      const double rx = c * px - s * pz, rz = s * px + c * pz;
      mnx = fmin(mnx, rx); mxx = fmax(mxx, rx); mnz = fmin(mnz, rz); mxz = fmax(mxz, rz);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mnx = fmin(mnx, __shfl_xor_sync(kFull, mnx, o));
      mxx = fmax(mxx, __shfl_xor_sync(kFull, mxx, o));
      mnz = fmin(mnz, __shfl_xor_sync(kFull, mnz, o));
      mxz = fmax(mxz, __shfl_xor_sync(kFull, mxz, o));
    }
    if (lane == 0) areas[e] = (mxx - mnx) * (mxz - mnz);
  }
  __syncthreads();
  const int e = first_strict_min(areas, hn, sm);
  if (e < 0) return 0.0;
  const int e1 = (e + 1 == hn) ? 0 : e + 1;
  return atan2(sm.hz[e1] - sm.hz[e], sm.hx[e1] - sm.hx[e]);
}

// ---- the kernel ------------------------------------------------------------------------
template <bool kScanned>
__global__ void __launch_bounds__(kThreads) fit_kernel(FitArgs a) {
  extern __shared__ __align__(16) unsigned char dyn_raw[];
  __shared__ Smem sm;
  double* areas = reinterpret_cast<double*>(dyn_raw);                               // [n_areas]
  uint32_t* pref = reinterpret_cast<uint32_t*>(dyn_raw + (size_t)a.n_areas * 8);    // [chunks+1], scanned source

  const int box = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int img = kScanned ? box / a.I : box;

  // camera and ground rotation: one thread each, overlapped with the prefix build
  if (tid == 0 && a.K) {
#pragma unroll
    for (int i = 0; i < 9; ++i) sm.Kmat[i] = a.K[(size_t)img * 9 + i];
    if (kScanned) invert3x3(sm.Kmat, sm.Kinv);
  }
  if (tid == 32) ground_rotation(a.ground ? a.ground + (size_t)box * 3 : nullptr, sm.Rg);

  long long n_src;
  const double* src_pts = nullptr;
  if (kScanned) {
    n_src = a.counts[box];
    // exclusive prefix over the chunk counts; each thread owns a contiguous run
    const uint16_t* cc = a.chunk_counts + (size_t)box * a.chunks;
    const int per = (a.chunks + kThreads - 1) / kThreads;
    const int c_lo = min(tid * per, a.chunks), c_hi = min(c_lo + per, a.chunks);
    uint32_t run = 0;
    for (int c = c_lo; c < c_hi; ++c) run += cc[c];
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
    for (int c = c_lo; c < c_hi; ++c) { pref[c] = base; base += cc[c]; }
    if (tid == kThreads - 1) pref[a.chunks] = base;
  } else {
    const long long o0 = a.offsets[box];
    n_src = a.offsets[box + 1] - o0;
    src_pts = a.pts + (size_t)o0 * 3;
  }
  __syncthreads();

  const bool subsample = n_src > LA3D_SUBSAMPLE;
  int status = LA3D_ST_OK;
  if (!kScanned && subsample && !a.sample_idx) status = LA3D_ST_TOO_MANY;
  const bool bad_method =
      a.method != LA3D_METHOD_PCA && a.method != LA3D_METHOD_CONVEX_HULL && a.method != LA3D_METHOD_SWEEP;
  const int nsel = status ? 0 : (int)min((long long)LA3D_SUBSAMPLE, n_src);

  // ---- gather + lift + ground alignment + NaN-row filter ---------------------------
  int n_valid = 0, inf_xz = 0, inf_y = 0;
  for (int k = tid; k < nsel; k += kThreads) {
    double X, Y, Z;
    if (kScanned) {
      const uint32_t r = subsample ? (uint32_t)a.ranks[(size_t)box * LA3D_SUBSAMPLE + k] : (uint32_t)k;
      const int p = select_pixel(a.bits + (size_t)box * a.chunks * kChunkWords, pref, a.chunks, r);
      const float d = __ldg(a.depth + (size_t)img * a.HW + p);
      const int v = p / a.W, u = p - v * a.W;
      lift_pixel_exact((double)d, (double)u, (double)v, sm.Kinv, X, Y, Z);
    } else {
      const long long row = subsample ? (long long)a.sample_idx[(size_t)box * LA3D_SUBSAMPLE + k] : (long long)k;
      X = src_pts[row * 3]; Y = src_pts[row * 3 + 1]; Z = src_pts[row * 3 + 2];
    }
    // np.dot(in_pc, Rg): r_j = sum_i p_i Rg[i][j]  (inf * 0 -> NaN drops the row, as in NumPy)
    double rx = X * sm.Rg[0] + Y * sm.Rg[3] + Z * sm.Rg[6];
    const double ry = X * sm.Rg[1] + Y * sm.Rg[4] + Z * sm.Rg[7];
    const double rz = X * sm.Rg[2] + Y * sm.Rg[5] + Z * sm.Rg[8];
    if (isnan(rx) || isnan(ry) || isnan(rz)) {
      rx = CUDART_NAN;
    } else {
      ++n_valid;
      inf_xz |= (int)(isinf(rx) || isinf(rz));
      inf_y |= (int)isinf(ry);
    }
    sm.x[k] = rx; sm.y[k] = ry; sm.z[k] = rz;
  }
  {
    int r[3] = {n_valid, inf_xz, inf_y};
    block_sum_int(r, sm);           // its barriers also publish sm.x / y / z
    n_valid = r[0]; inf_xz = r[1]; inf_y = r[2];
  }
  if (status == LA3D_ST_OK) {
    // same order as the reference: the NaN filter raises first (:142-143), then the method check (:151)
    if (n_valid == 0) status = LA3D_ST_NO_VALID;
    else if (bad_method) status = LA3D_ST_BAD_METHOD;
    else if (inf_xz) status = LA3D_ST_NONFINITE;      // scikit-learn's input check raises; Qhull fails first and falls back to it
    else if (n_valid == 1 && a.method != LA3D_METHOD_SWEEP) status = LA3D_ST_PCA_UNDEFINED;   // PCA(2) needs 2 samples
  }

  if (status != LA3D_ST_OK) {                          // uniform across the CTA
    if (tid < LA3D_REC) {
      double val = CUDART_NAN;
      if (tid == LA3D_O_NVALID) val = (double)n_valid;
      if (tid == LA3D_O_STATUS) val = (double)status;
      if (tid == LA3D_O_NMASK) val = (double)n_src;
      if (tid == LA3D_O_PAD) val = 0.0;
      if (a.rec_f64) reinterpret_cast<double*>(a.records)[(size_t)box * LA3D_REC + tid] = val;
      else reinterpret_cast<float*>(a.records)[(size_t)box * LA3D_REC + tid] = (float)val;
    }
    return;
  }

  // ---- yaw ---------------------------------------------------------------------------
  double yaw;
  if (a.method == LA3D_METHOD_SWEEP) {
    yaw = yaw_sweep(sm, nsel, a.yaw_steps, areas, inf_y != 0);
  } else {
    int hn = 0;
    if (a.method == LA3D_METHOD_CONVEX_HULL) hn = hull_wrap(sm, nsel);
    yaw = hn ? yaw_hull_edges(sm, hn, areas) : yaw_pca(sm, nsel, n_valid);
  }

  // ---- extents at that yaw: rotate_y(yaw) @ pc^T, per-axis min / max (:154-160) ------
  double sy_, cy_;
  sincos(yaw, &sy_, &cy_);
  double lo[3] = {CUDART_INF, CUDART_INF, CUDART_INF}, hi[3] = {-CUDART_INF, -CUDART_INF, -CUDART_INF};
  for (int k = tid; k < nsel; k += kThreads) {
    const double px = sm.x[k], py = sm.y[k], pz = sm.z[k];
    if (px == px) {
      const double rx = cy_ * px + sy_ * pz, rz = cy_ * pz - sy_ * px;
      lo[0] = fmin(lo[0], rx); hi[0] = fmax(hi[0], rx);
      lo[1] = fmin(lo[1], py); hi[1] = fmax(hi[1], py);
      lo[2] = fmin(lo[2], rz); hi[2] = fmax(hi[2], rz);
    }
  }
  block_reduce(lo, sm, OpMin());
  block_reduce(hi, sm, OpMax());
  if (inf_y) { lo[0] = hi[0] = lo[2] = hi[2] = CUDART_NAN; }      // 0 * inf in the x / z rows of the product
  const double dim[3] = {hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]};
  const double ctr[3] = {(lo[0] + hi[0]) / 2, (lo[1] + hi[1]) / 2, (lo[2] + hi[2]) / 2};

  // rotate_y(-yaw) = [[c,0,-s],[0,1,0],[s,0,c]] with c = cos(yaw), s = sin(yaw)
  const double Ry[9] = {cy_, 0.0, -sy_, 0.0, 1.0, 0.0, sy_, 0.0, cy_};
  __syncthreads();
  if (tid < 8) {
    // convert_box_vertices(cx,cy,cz,dx,dy,dz,0).astype(float16)  (:71-103, :165)
    const double sgx = (tid == 1 || tid == 2 || tid == 5 || tid == 6) ? 1.0 : -1.0;
    const double sgy = (tid == 2 || tid == 3 || tid == 6 || tid == 7) ? 1.0 : -1.0;
    const double sgz = (tid >= 4) ? 1.0 : -1.0;
    const double lx = sgx * (dim[0] / 2), ly = sgy * (dim[1] / 2), lz = sgz * (dim[2] / 2);
    // local @ rot(0)^T with rot(0) = [[1,0,0],[0,1,0],[-0,0,1]]: kept explicit for inf/NaN parity
    double V[3];
    V[0] = (lx * 1.0 + ly * 0.0 + lz * 0.0) + ctr[0];
    V[1] = (lx * 0.0 + ly * 1.0 + lz * 0.0) + ctr[1];
    V[2] = (lx * -0.0 + ly * 0.0 + lz * 1.0) + ctr[2];
#pragma unroll
    for (int i = 0; i < 3; ++i) V[i] = (double)__half2float(__double2half(V[i]));
    // vertices = (rotate_y(-yaw) @ V^T)^T @ Rg^T  (:168-169)
    double v1[3], v2[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) v1[i] = Ry[i * 3] * V[0] + Ry[i * 3 + 1] * V[1] + Ry[i * 3 + 2] * V[2];
#pragma unroll
    for (int i = 0; i < 3; ++i) v2[i] = v1[0] * sm.Rg[i * 3] + v1[1] * sm.Rg[i * 3 + 1] + v1[2] * sm.Rg[i * 3 + 2];
#pragma unroll
    for (int i = 0; i < 3; ++i) sm.rec[LA3D_O_VERT + tid * 3 + i] = v2[i];
    // project_to_2d (util.py:227-229)
    double uu = CUDART_NAN, vv = CUDART_NAN;
    if (a.K) {
      const double h0 = sm.Kmat[0] * v2[0] + sm.Kmat[1] * v2[1] + sm.Kmat[2] * v2[2];
      const double h1 = sm.Kmat[3] * v2[0] + sm.Kmat[4] * v2[1] + sm.Kmat[5] * v2[2];
      const double h2 = sm.Kmat[6] * v2[0] + sm.Kmat[7] * v2[1] + sm.Kmat[8] * v2[2];
      uu = h0 / h2; vv = h1 / h2;
    }
    sm.rec[LA3D_O_UV + tid * 2] = uu;
    sm.rec[LA3D_O_UV + tid * 2 + 1] = vv;
  } else if (tid == 32) {
    // center_cam = Rg^T @ (rotate_y(-yaw) @ c)  (:172-173; Rg^T where the corners used Rg - kept)
    double w[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) w[i] = Ry[i * 3] * ctr[0] + Ry[i * 3 + 1] * ctr[1] + Ry[i * 3 + 2] * ctr[2];
#pragma unroll
    for (int i = 0; i < 3; ++i) sm.rec[LA3D_O_CENTER + i] = sm.Rg[i] * w[0] + sm.Rg[3 + i] * w[1] + sm.Rg[6 + i] * w[2];
    sm.rec[LA3D_O_DIM] = dim[2]; sm.rec[LA3D_O_DIM + 1] = dim[1]; sm.rec[LA3D_O_DIM + 2] = dim[0];
    sm.rec[LA3D_O_YAW] = yaw;
    sm.rec[LA3D_O_NVALID] = (double)n_valid;
    sm.rec[LA3D_O_STATUS] = (double)LA3D_ST_OK;
    sm.rec[LA3D_O_NMASK] = (double)n_src;
    sm.rec[LA3D_O_PAD] = 0.0;
  } else if (tid == 64) {
    // R_cam = Rg^T @ rotate_y(-yaw)  (:176)
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int k = 0; k < 3; ++k)
        sm.rec[LA3D_O_RCAM + i * 3 + k] = sm.Rg[i] * Ry[k] + sm.Rg[3 + i] * Ry[3 + k] + sm.Rg[6 + i] * Ry[6 + k];
  }
  __syncthreads();
  if (tid == 0) {
    // Python min()/max() over the 8 projections (combine_results.py:241-246): sequential
    // `if x < m` / `if x > m`, so a leading NaN sticks and a later NaN is skipped.
    double mnu = sm.rec[LA3D_O_UV], mnv = sm.rec[LA3D_O_UV + 1], mxu = mnu, mxv = mnv;
    for (int j = 1; j < 8; ++j) {
      const double uu = sm.rec[LA3D_O_UV + 2 * j], vv = sm.rec[LA3D_O_UV + 2 * j + 1];
      if (uu < mnu) mnu = uu;
      if (vv < mnv) mnv = vv;
      if (uu > mxu) mxu = uu;
      if (vv > mxv) mxv = vv;
    }
    sm.rec[LA3D_O_BOX2D] = mnu; sm.rec[LA3D_O_BOX2D + 1] = mnv;
    sm.rec[LA3D_O_BOX2D + 2] = mxu; sm.rec[LA3D_O_BOX2D + 3] = mxv;
  }
  __syncthreads();
  if (tid < LA3D_REC) {
    if (a.rec_f64) reinterpret_cast<double*>(a.records)[(size_t)box * LA3D_REC + tid] = sm.rec[tid];
    else reinterpret_cast<float*>(a.records)[(size_t)box * LA3D_REC + tid] = (float)sm.rec[tid];
  }
}

int launch_fit(bool scanned, FitArgs& a, int nboxes, cudaStream_t s) {
  a.n_areas = a.method == LA3D_METHOD_SWEEP ? (a.yaw_steps > 0 ? a.yaw_steps : 1) : kMaxPts;
  a.n_areas = (a.n_areas + 1) & ~1;                       // keep pref 16-byte aligned
  const size_t dyn = (size_t)a.n_areas * 8 + (scanned ? ((size_t)a.chunks + 1) * 4 : 0);
  if (dyn + sizeof(Smem) > 227 * 1024) {
    set_error("la3d fit: image or yaw sweep too large for shared memory (%zu bytes needed)", dyn + sizeof(Smem));
    return LA3D_EINVAL;
  }
  if (scanned) {
    if (dyn > 16 * 1024) LA3D_CUDA(cudaFuncSetAttribute(fit_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    fit_kernel<true><<<nboxes, kThreads, dyn, s>>>(a);
  } else {
    if (dyn > 16 * 1024) LA3D_CUDA(cudaFuncSetAttribute(fit_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    fit_kernel<false><<<nboxes, kThreads, dyn, s>>>(a);
  }
  LA3D_CUDA(cudaGetLastError());
  return LA3D_OK;
}

}  // namespace
}  // namespace la3d

extern "C" int la3d_fit_scanned(const float* depth, const double* K, const double* ground, const uint32_t* bits,
                                const uint16_t* chunk_counts, const int32_t* counts, const int32_t* ranks, int B,
                                int I, int H, int W, int method, int yaw_steps, void* records, int rec_f64,
                                la3d_stream_t stream) {
  using namespace la3d;
  LA3D_REQUIRE(depth && K && bits && chunk_counts && counts && ranks && records, "null pointer");
  LA3D_REQUIRE(B > 0 && I > 0 && H > 0 && W > 0, "non-positive shape");
  LA3D_REQUIRE((long long)H * W < (1ll << 30), "image too large");
  LA3D_REQUIRE(method != LA3D_METHOD_SWEEP || (yaw_steps > 0 && yaw_steps <= 16384), "sweep needs 1 <= yaw_steps <= 16384");
  FitArgs a{};
  a.depth = depth; a.bits = bits; a.chunk_counts = chunk_counts; a.counts = counts; a.ranks = ranks;
  a.I = I; a.HW = H * W; a.W = W; a.chunks = (int)la3d_chunks_per_plane(H, W);
  a.K = K; a.ground = ground; a.method = method; a.yaw_steps = yaw_steps; a.records = records; a.rec_f64 = rec_f64;
  return launch_fit(true, a, B * I, static_cast<cudaStream_t>(stream));
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
  a.K = K; a.ground = ground; a.method = method; a.yaw_steps = yaw_steps; a.records = records; a.rec_f64 = rec_f64;
  return launch_fit(false, a, nboxes, static_cast<cudaStream_t>(stream));
}
