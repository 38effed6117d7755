// Depth -> camera-space points on sm_100a.  Replaces depth_to_points
// (src/util.py:52-75 of the reference).  HBM-bound: 4 B read + 12 B (float) or
// 24 B (double) written per pixel.
//
// lift_bulk_kernel (the fast path).  Cameras (inverse intrinsics, optional rigid transform): for
// short launches every CTA prepares its image's camera itself while its depth loads are in
// flight; for long ones lift_prep_kernel prepares all of them once, one thread per image, into a
// stream-ordered scratch buffer (the threshold was measured, see the kernel).  One CTA of 128 threads
// owns a tile of 1536 (float) / 1024 (double) consecutive pixels of one image: each
// thread issues its 16-byte depth loads, reads its image's camera (uniform loads),
// converts, and writes its points into the CTA's shared-memory tile; one thread hands
// the finished tile (18 / 24 KB, contiguous in the output) to the TMA unit with a
// single cp.async.bulk shared -> global.  Measured on B200 (config 2 / config 4):
// float 6.15 / 6.75 TB/s, double 6.46 / 6.70 TB/s = 0.94-1.03 of the measured copy peak.
// lift_tile_kernel is the same without staging (direct 16-byte stores at a 48-byte
// stride: 5.2 TB/s); kept selectable (LA3D_LIFT_VARIANT=0) for comparison.
//
// lift_scalar_kernel (generic fallback: pixel counts not divisible by 4, unaligned
// buffers, no stream-ordered allocator): one pixel per thread, camera per CTA.
#include <cstdlib>

#include "common.cuh"

namespace la3d {
namespace {

constexpr int kThreads = 256;
constexpr int kStepPx = kThreads * 4;   // pixels per CTA step

struct Camera {
  double Kinv[9];   // exact path
  double M[9];      // fused path: R @ Kinv
  double t[3];
  double R[9];
};

__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(float* p, float a, float b, float c, float d) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void st_stream(double* p, double a, double b) {
  asm volatile("st.global.cs.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}

__device__ __forceinline__ void prepare_camera(Camera& cam, const double* Kb, int k_is_inverse, const double* R,
                                               const double* t) {
  if (k_is_inverse) {
#pragma unroll
    for (int i = 0; i < 9; ++i) cam.Kinv[i] = Kb[i];
  } else {
    invert3x3(Kb, cam.Kinv);
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) cam.t[i] = t ? t[i] : 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      cam.R[i * 3 + j] = R ? R[i * 3 + j] : (i == j ? 1.0 : 0.0);
      cam.M[i * 3 + j] = R ? (R[i * 3 + 0] * cam.Kinv[0 * 3 + j] + R[i * 3 + 1] * cam.Kinv[1 * 3 + j] +
                              R[i * 3 + 2] * cam.Kinv[2 * 3 + j])
                           : cam.Kinv[i * 3 + j];
    }
}

// One pixel, float32 result: a single fused float64 pass (ray = M.[u,v,1]; p = d*ray + t)
// rounded once.  A non-finite depth takes the reference's exact operation order instead so
// that inf / NaN propagate exactly as in NumPy.
__device__ __forceinline__ void pixel_f32(float d, double ud, double vd, double r0, double r1, double r2,
                                          const Camera& cam, const Camera* __restrict__ full, const double* Rp,
                                          const double* tp, float& x, float& y, float& z) {
  const double dd = (double)d;
  if (isfinite(d)) {
    x = (float)fma(dd, fma(cam.M[0], ud, r0), cam.t[0]);
    y = (float)fma(dd, fma(cam.M[3], ud, r1), cam.t[1]);
    z = (float)fma(dd, fma(cam.M[6], ud, r2), cam.t[2]);
  } else {
    double X, Y, Z;
    lift_pixel_exact(dd, ud, vd, full->Kinv, X, Y, Z);
    rigid_exact(Rp, tp, X, Y, Z);
    x = (float)X; y = (float)Y; z = (float)Z;
  }
}

// 4 consecutive pixels starting at (v,u) of one image -> 48 bytes (kSmem: into the CTA's staging tile)
template <bool kSmem>
__device__ __forceinline__ void quad_f32(const float4 dq, int u, int v, int W, const Camera& cam,
                                         const Camera* __restrict__ full, const double* Rp, const double* tp,
                                         float* __restrict__ dst) {
  float x0, y0, z0, x1, y1, z1, x2, y2, z2, x3, y3, z3;
  const double vd = (double)v, ud = (double)u;
  const double r0 = fma(cam.M[1], vd, cam.M[2]), r1 = fma(cam.M[4], vd, cam.M[5]), r2 = fma(cam.M[7], vd, cam.M[8]);
  if (u + 3 < W) {
    pixel_f32(dq.x, ud, vd, r0, r1, r2, cam, full, Rp, tp, x0, y0, z0);
    pixel_f32(dq.y, ud + 1.0, vd, r0, r1, r2, cam, full, Rp, tp, x1, y1, z1);
    pixel_f32(dq.z, ud + 2.0, vd, r0, r1, r2, cam, full, Rp, tp, x2, y2, z2);
    pixel_f32(dq.w, ud + 3.0, vd, r0, r1, r2, cam, full, Rp, tp, x3, y3, z3);
  } else {   // the quad wraps to the next row (W not a multiple of 4)
    int uu = u, vv = v;
    auto one = [&](float d, float& x, float& y, float& z) {
      const double vq = (double)vv;
      pixel_f32(d, (double)uu, vq, fma(cam.M[1], vq, cam.M[2]), fma(cam.M[4], vq, cam.M[5]),
                fma(cam.M[7], vq, cam.M[8]), cam, full, Rp, tp, x, y, z);
      if (++uu == W) { uu = 0; ++vv; }
    };
    one(dq.x, x0, y0, z0); one(dq.y, x1, y1, z1); one(dq.z, x2, y2, z2); one(dq.w, x3, y3, z3);
  }
  if (kSmem) {
    reinterpret_cast<float4*>(dst)[0] = make_float4(x0, y0, z0, x1);
    reinterpret_cast<float4*>(dst)[1] = make_float4(y1, z1, x2, y2);
    reinterpret_cast<float4*>(dst)[2] = make_float4(z2, x3, y3, z3);
  } else {
    st_stream(dst, x0, y0, z0, x1);
    st_stream(dst + 4, y1, z1, x2, y2);
    st_stream(dst + 8, z2, x3, y3, z3);
  }
}

// float64 result: the reference's operation order, bit for bit
template <bool kSmem>
__device__ __forceinline__ void quad_f64(const float4 dq, int u, int v, int W, const Camera& cam, const double* Rp,
                                         const double* tp, double* __restrict__ dst) {
  double x0, y0, z0, x1, y1, z1;
  int uu = u, vv = v;
  auto one = [&](float d, double& x, double& y, double& z) {
    lift_pixel_exact((double)d, (double)uu, (double)vv, cam.Kinv, x, y, z);
    rigid_exact(Rp, tp, x, y, z);
    if (++uu == W) { uu = 0; ++vv; }
  };
  auto put = [&](double* p, double a, double b) {
    if (kSmem) *reinterpret_cast<double2*>(p) = make_double2(a, b);
    else st_stream(p, a, b);
  };
  one(dq.x, x0, y0, z0); one(dq.y, x1, y1, z1);
  put(dst, x0, y0); put(dst + 2, z0, x1); put(dst + 4, y1, z1);
  one(dq.z, x0, y0, z0); one(dq.w, x1, y1, z1);
  put(dst + 6, x0, y0); put(dst + 8, z0, x1); put(dst + 10, y1, z1);
}

// Cameras of all images, prepared once per launch by lift_prep_kernel.
__global__ void lift_prep_kernel(const double* __restrict__ K, int k_stride, int k_is_inverse,
                                 const double* __restrict__ R, const double* __restrict__ t, int B,
                                 Camera* __restrict__ cams) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) prepare_camera(cams[b], K + (size_t)b * k_stride, k_is_inverse, R, t);
}

// One CTA = kQuads x 1024 consecutive pixels of ONE image.  No shared memory, no barrier:
// every thread issues its kQuads 16-byte depth loads, then reads its image's camera
// (uniform, L1/L2-resident) and converts quad by quad.
template <bool kF64, int kQuads, int kMinCtas>
__global__ void __launch_bounds__(kThreads, kMinCtas)
    lift_tile_kernel(const float* __restrict__ depth, const Camera* __restrict__ cams, int has_R, int has_t, int HW,
                     int W, int tiles_per_image, void* __restrict__ out_) {
  const int b = blockIdx.x / tiles_per_image;
  const int tile = blockIdx.x - b * tiles_per_image;
  int px = tile * (kQuads * kStepPx) + threadIdx.x * 4;
  const float* img = depth + (size_t)b * HW;
  float4 dq[kQuads];
#pragma unroll
  for (int q = 0; q < kQuads; ++q)
    dq[q] = (px + q * kStepPx < HW) ? ld_stream(reinterpret_cast<const float4*>(img + px + q * kStepPx))
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
  const Camera* gc = cams + b;
  Camera cam;                       // only the fields a path touches are actually loaded
  if (kF64) {
#pragma unroll
    for (int i = 0; i < 9; ++i) cam.Kinv[i] = __ldg(&gc->Kinv[i]);
  } else {
#pragma unroll
    for (int i = 0; i < 9; ++i) cam.M[i] = __ldg(&gc->M[i]);
#pragma unroll
    for (int i = 0; i < 3; ++i) cam.t[i] = __ldg(&gc->t[i]);
  }
  const double* Rp = has_R ? gc->R : nullptr;
  const double* tp = has_t ? gc->t : nullptr;
  int v = px / W, u = px - v * W;
  const int dv = kStepPx / W, du = kStepPx - dv * W;
#pragma unroll
  for (int q = 0; q < kQuads; ++q) {
    if (px < HW) {
      const size_t o = ((size_t)b * HW + px) * 3;
      if (kF64) quad_f64<false>(dq[q], u, v, W, cam, Rp, tp, reinterpret_cast<double*>(out_) + o);
      else quad_f32<false>(dq[q], u, v, W, cam, gc, Rp, tp, reinterpret_cast<float*>(out_) + o);
    }
    px += kStepPx;
    u += du; v += dv;
    if (u >= W) { u -= W; ++v; }
  }
}

// The bulk-store form of the tile kernel.  Direct stores of the tile kernel above write 16-byte
// pieces at a 48-byte (float) / 96-byte (double) stride per lane, i.e. half sectors per
// instruction.  Here the CTA assembles its tile of points in shared memory and ONE thread hands
// the whole contiguous tile (up to 48 KB) to the TMA unit (cp.async.bulk shared -> global), which
// writes full lines.  One CTA = kQuads x kT x 4 consecutive pixels of one image.
// kOwnCam: every CTA prepares its image's camera itself (one thread, while the depth loads are in flight) - no
// preparation launch and no scratch buffer, which wins for short launches (configs[1]: 206 vs 216 us per call);
// otherwise the cameras come from lift_prep_kernel's buffer, which wins when the launch is long enough to amortise
// the extra launch (configs[3]: 730 vs 763 us).
template <bool kF64, int kQuads, int kT, bool kOwnCam>
__global__ void __launch_bounds__(kT)
    lift_bulk_kernel(const float* __restrict__ depth, const Camera* __restrict__ cams, const double* __restrict__ K,
                     int k_stride, int k_is_inverse, const double* __restrict__ R, const double* __restrict__ t, int HW,
                     int W, int tiles_per_image, void* __restrict__ out_) {
  extern __shared__ __align__(128) unsigned char stage_raw[];
  __shared__ Camera scam;
  constexpr int kStep = kT * 4;
  constexpr int kTilePx = kQuads * kStep;
  const int b = blockIdx.x / tiles_per_image;
  const int tile = blockIdx.x - b * tiles_per_image;
  const int px0 = tile * kTilePx;
  int px = px0 + threadIdx.x * 4;
  const float* img = depth + (size_t)b * HW;
  float4 dq[kQuads];
#pragma unroll
  for (int q = 0; q < kQuads; ++q)
    dq[q] = (px + q * kStep < HW) ? ld_stream(reinterpret_cast<const float4*>(img + px + q * kStep))
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
  if (kOwnCam) {
    if (threadIdx.x == 0) prepare_camera(scam, K + (size_t)b * k_stride, k_is_inverse, R, t);
    __syncthreads();
  }
  const Camera* gc = kOwnCam ? &scam : cams + b;
  const int has_R = R != nullptr, has_t = t != nullptr;
  Camera cam;
  if (kF64) {
#pragma unroll
    for (int i = 0; i < 9; ++i) cam.Kinv[i] = kOwnCam ? gc->Kinv[i] : __ldg(&gc->Kinv[i]);
  } else {
#pragma unroll
    for (int i = 0; i < 9; ++i) cam.M[i] = kOwnCam ? gc->M[i] : __ldg(&gc->M[i]);
#pragma unroll
    for (int i = 0; i < 3; ++i) cam.t[i] = kOwnCam ? gc->t[i] : __ldg(&gc->t[i]);
  }
  const double* Rp = has_R ? gc->R : nullptr;
  const double* tp = has_t ? gc->t : nullptr;
  int v = px / W, u = px - v * W;
  const int dv = kStep / W, du = kStep - dv * W;
#pragma unroll
  for (int q = 0; q < kQuads; ++q) {
    if (px < HW) {
      const size_t o = (size_t)(px - px0) * 3;
      if (kF64) quad_f64<true>(dq[q], u, v, W, cam, Rp, tp, reinterpret_cast<double*>(stage_raw) + o);
      else quad_f32<true>(dq[q], u, v, W, cam, gc, Rp, tp, reinterpret_cast<float*>(stage_raw) + o);
    }
    px += kStep;
    u += du; v += dv;
    if (u >= W) { u -= W; ++v; }
  }
  // generic-proxy writes -> visible to the async proxy, then one bulk store of the whole tile
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    const int n = min(kTilePx, HW - px0);
    const uint32_t bytes = (uint32_t)n * 3u * (kF64 ? 8u : 4u);
    unsigned char* dst = static_cast<unsigned char*>(out_) + ((size_t)b * HW + px0) * 3 * (kF64 ? 8 : 4);
    const uint32_t src = (uint32_t)__cvta_generic_to_shared(stage_raw);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // the tile must outlive the read
  }
}

// Generic fallback: one pixel per thread; blockIdx.y = image.
template <bool kF64>
__global__ void __launch_bounds__(kThreads) lift_scalar_kernel(const float* __restrict__ depth,
                                                               const double* __restrict__ K, int k_stride,
                                                               int k_is_inverse, const double* __restrict__ R,
                                                               const double* __restrict__ t, int HW, int W,
                                                               void* __restrict__ out_) {
  __shared__ Camera cam;
  const int b = blockIdx.y;
  if (threadIdx.x == 0) prepare_camera(cam, K + (size_t)b * k_stride, k_is_inverse, R, t);
  __syncthreads();
  const double* Rp = R ? cam.R : nullptr;
  const double* tp = t ? cam.t : nullptr;
  for (int p = blockIdx.x * kThreads + threadIdx.x; p < HW; p += gridDim.x * kThreads) {
    const int v = p / W, u = p - v * W;
    const float d = __ldg(depth + (size_t)b * HW + p);
    const size_t o = ((size_t)b * HW + p) * 3;
    if (kF64) {
      double x, y, z;
      lift_pixel_exact((double)d, (double)u, (double)v, cam.Kinv, x, y, z);
      rigid_exact(Rp, tp, x, y, z);
      double* dst = reinterpret_cast<double*>(out_) + o;
      dst[0] = x; dst[1] = y; dst[2] = z;
    } else {
      const double vd = (double)v;
      float x, y, z;
      pixel_f32(d, (double)u, vd, fma(cam.M[1], vd, cam.M[2]), fma(cam.M[4], vd, cam.M[5]),
                fma(cam.M[7], vd, cam.M[8]), cam, &cam, Rp, tp, x, y, z);
      float* dst = reinterpret_cast<float*>(out_) + o;
      dst[0] = x; dst[1] = y; dst[2] = z;
    }
  }
}

template <bool kF64, int kQ, int kT>
int launch_bulk(const float* depth, const Camera* cams, const double* K, int k_stride, int k_is_inverse, const double* R,
                const double* t, int B, int HW, int W, void* out, cudaStream_t s) {
  constexpr int kTilePx = kQ * kT * 4;
  constexpr int kSmem = kTilePx * 3 * (kF64 ? 8 : 4);
  const int tiles = (HW + kTilePx - 1) / kTilePx;
  LA3D_REQUIRE((long long)tiles * B < (1ll << 31), "grid too large");
  static_assert(kSmem + sizeof(Camera) <= 48 * 1024, "tile + camera must fit the default shared memory limit");
  if (cams) lift_bulk_kernel<kF64, kQ, kT, false><<<(unsigned)(tiles * B), kT, kSmem, s>>>(depth, cams, K, k_stride, k_is_inverse, R, t, HW, W, tiles, out);
  else lift_bulk_kernel<kF64, kQ, kT, true><<<(unsigned)(tiles * B), kT, kSmem, s>>>(depth, nullptr, K, k_stride, k_is_inverse, R, t, HW, W, tiles, out);
  return LA3D_OK;
}

int lift_variant() {
  static const int v = getenv("LA3D_LIFT_VARIANT") ? atoi(getenv("LA3D_LIFT_VARIANT")) : 1;   // 0: direct stores (no staging)
  return v;
}

constexpr int kQuads = 4;      // 16-byte loads in flight per thread
constexpr int kMinCtas = 3;    // resident CTAs per SM the tile kernels are compiled for

}  // namespace
}  // namespace la3d

extern "C" int la3d_depth_lift(const float* depth, const double* K, int k_stride, int k_is_inverse, const double* R,
                               const double* t, int B, int H, int W, void* out, int out_f64, la3d_stream_t stream) {
  using namespace la3d;
  LA3D_REQUIRE(depth && K && out, "null pointer");
  LA3D_REQUIRE(B > 0 && H > 0 && W > 0, "non-positive shape");
  LA3D_REQUIRE(k_stride == 0 || k_stride == 9, "k_stride must be 0 (shared) or 9 (per image)");
  LA3D_REQUIRE((long long)H * W < (1ll << 30), "image too large");
  const int HW = H * W;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool vec = (HW % 4 == 0) && aligned16(depth) && aligned16(out);
  const int variant = lift_variant();
  // short launches: every CTA prepares its own camera; long ones: one preparation launch into a stream-ordered scratch
  const bool own_cam = variant != 0 && (long long)B * HW <= 150000000ll;
  Camera* cams = nullptr;
  if (vec && !own_cam && cudaMallocAsync(reinterpret_cast<void**>(&cams), sizeof(Camera) * (size_t)B, s) != cudaSuccess) {
    cudaGetLastError();   // no stream-ordered pool on this driver: the CTAs prepare their cameras themselves
    cams = nullptr;
  }
  if (vec && (cams || variant != 0)) {
    if (cams) lift_prep_kernel<<<(B + 127) / 128, 128, 0, s>>>(K, k_stride, k_is_inverse, R, t, B, cams);
    if (variant == 0) {
      // LA3D_LIFT_VARIANT=0: direct stores, one CTA per 4096-pixel tile of one image
      const int tiles = (HW + kQuads * kStepPx - 1) / (kQuads * kStepPx);
      LA3D_REQUIRE((long long)tiles * B < (1ll << 31), "grid too large");
      const unsigned grid = (unsigned)(tiles * B);
      if (out_f64)
        lift_tile_kernel<true, kQuads, kMinCtas><<<grid, kThreads, 0, s>>>(depth, cams, R != nullptr, t != nullptr, HW, W, tiles, out);
      else
        lift_tile_kernel<false, kQuads, kMinCtas><<<grid, kThreads, 0, s>>>(depth, cams, R != nullptr, t != nullptr, HW, W, tiles, out);
    } else {
      // tile sizes measured on B200 (tools/lift_variants.py): 1536-pixel tiles (18 KB) for float,
      // 1024-pixel tiles (24 KB) for double; variant 2 = 2048- / 1536-pixel tiles
      int rc = LA3D_OK;
      if (variant == 2) rc = out_f64 ? launch_bulk<true, 3, 128>(depth, cams, K, k_stride, k_is_inverse, R, t, B, HW, W, out, s)
                                     : launch_bulk<false, 4, 128>(depth, cams, K, k_stride, k_is_inverse, R, t, B, HW, W, out, s);
      else rc = out_f64 ? launch_bulk<true, 2, 128>(depth, cams, K, k_stride, k_is_inverse, R, t, B, HW, W, out, s)
                        : launch_bulk<false, 3, 128>(depth, cams, K, k_stride, k_is_inverse, R, t, B, HW, W, out, s);
      if (rc) return rc;
    }
    LA3D_CUDA(cudaGetLastError());
    if (cams) LA3D_CUDA(cudaFreeAsync(cams, s));
  } else {
    LA3D_REQUIRE(B <= 65535, "fallback path supports at most 65535 images per call");
    dim3 grid((unsigned)min((HW + kThreads - 1) / kThreads, 4096), (unsigned)B);
    if (out_f64) lift_scalar_kernel<true><<<grid, kThreads, 0, s>>>(depth, K, k_stride, k_is_inverse, R, t, HW, W, out);
    else lift_scalar_kernel<false><<<grid, kThreads, 0, s>>>(depth, K, k_stride, k_is_inverse, R, t, HW, W, out);
  }
  LA3D_CUDA(cudaGetLastError());
  return LA3D_OK;
}
