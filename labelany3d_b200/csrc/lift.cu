// Depth -> camera-space points on sm_100a.  Replaces depth_to_points
// (src/util.py:52-75 of the reference).  HBM-bound: 4 B read + 12 B (float) or
// 24 B (double) written per pixel.
//
// One CTA owns a tile of kTilePx consecutive pixels of ONE image, so the camera
// (inverse intrinsics, optional rigid transform) is prepared once per CTA in
// shared memory.  Each thread handles 4 consecutive pixels per step and issues
// all kUnroll 16-byte depth loads of its steps before the first use.
#include "common.cuh"

namespace la3d {
namespace {

constexpr int kThreads = 256;
constexpr int kUnroll = 4;
constexpr int kStepPx = kThreads * 4;          // pixels per CTA step
constexpr int kTilePx = kStepPx * kUnroll;     // pixels per CTA

struct Camera {
  double Kinv[9];   // exact path
  double M[9];      // fast path: R @ Kinv
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
__device__ __forceinline__ void st_stream(float4* p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_stream(double2* p, double2 v) {
  asm volatile("st.global.cs.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}

// kF64: the output dtype.  double -> reference operation order, bit for bit.
// float -> one fused pass in double (ray = M.[u,v,1]; p = d*ray + t), rounded
// once to float; pixels whose depth is not finite take the exact path so that
// inf/NaN propagate exactly as in NumPy.
template <bool kF64, bool kVec>
__global__ void __launch_bounds__(kThreads) lift_kernel(const float* __restrict__ depth,
                                                        const double* __restrict__ K, int k_stride,
                                                        int k_is_inverse, const double* __restrict__ R,
                                                        const double* __restrict__ t, int HW, int W,
                                                        int tiles_per_image, void* __restrict__ out_) {
  __shared__ Camera cam;
  const int b = blockIdx.x / tiles_per_image;
  const int tile = blockIdx.x - b * tiles_per_image;
  const int tile_px = tile * kTilePx;
  const float* img = depth + (size_t)b * HW;

  // Issue this thread's depth loads first; the camera set-up below overlaps them.
  float4 dv[kUnroll];
  if (kVec) {
#pragma unroll
    for (int j = 0; j < kUnroll; ++j) {
      int px = tile_px + j * kStepPx + threadIdx.x * 4;
      dv[j] = (px < HW) ? ld_stream(reinterpret_cast<const float4*>(img + px)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  } else {
#pragma unroll
    for (int j = 0; j < kUnroll; ++j) {
      int px = tile_px + j * kStepPx + threadIdx.x * 4;
      float e[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) e[q] = (px + q < HW) ? __ldg(img + px + q) : 0.f;
      dv[j] = make_float4(e[0], e[1], e[2], e[3]);
    }
  }

  if (threadIdx.x == 0) {
    const double* Kb = K + (size_t)b * k_stride;
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
  __syncthreads();
  const double* Rp = R ? cam.R : nullptr;
  const double* tp = t ? cam.t : nullptr;

#pragma unroll
  for (int j = 0; j < kUnroll; ++j) {
    const int px = tile_px + j * kStepPx + threadIdx.x * 4;
    if (px >= HW) continue;
    int v = px / W;
    int u = px - v * W;
    const float dd[4] = {dv[j].x, dv[j].y, dv[j].z, dv[j].w};
    double o[12];
    if (kF64) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        lift_pixel_exact((double)dd[q], (double)u, (double)v, cam.Kinv, o[q * 3], o[q * 3 + 1], o[q * 3 + 2]);
        rigid_exact(Rp, tp, o[q * 3], o[q * 3 + 1], o[q * 3 + 2]);
        if (++u == W) { u = 0; ++v; }
      }
    } else {
      double vd = (double)v, ud = (double)u;
      double r0 = fma(cam.M[1], vd, cam.M[2]), r1 = fma(cam.M[4], vd, cam.M[5]), r2 = fma(cam.M[7], vd, cam.M[8]);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        double d = (double)dd[q];
        if (isfinite(dd[q])) {
          o[q * 3 + 0] = fma(d, fma(cam.M[0], ud, r0), cam.t[0]);
          o[q * 3 + 1] = fma(d, fma(cam.M[3], ud, r1), cam.t[1]);
          o[q * 3 + 2] = fma(d, fma(cam.M[6], ud, r2), cam.t[2]);
        } else {
          lift_pixel_exact(d, ud, vd, cam.Kinv, o[q * 3], o[q * 3 + 1], o[q * 3 + 2]);
          rigid_exact(Rp, tp, o[q * 3], o[q * 3 + 1], o[q * 3 + 2]);
        }
        ud += 1.0;
        if (++u == W) {
          u = 0; ud = 0.0; ++v; vd += 1.0;
          r0 = fma(cam.M[1], vd, cam.M[2]); r1 = fma(cam.M[4], vd, cam.M[5]); r2 = fma(cam.M[7], vd, cam.M[8]);
        }
      }
    }
    const int n = min(4, HW - px);
    if (kF64) {
      double* dst = reinterpret_cast<double*>(out_) + ((size_t)b * HW + px) * 3;
      if (kVec) {
#pragma unroll
        for (int q = 0; q < 6; ++q) st_stream(reinterpret_cast<double2*>(dst) + q, make_double2(o[2 * q], o[2 * q + 1]));
      } else {
        for (int q = 0; q < 3 * n; ++q) dst[q] = o[q];
      }
    } else {
      float* dst = reinterpret_cast<float*>(out_) + ((size_t)b * HW + px) * 3;
      if (kVec) {
#pragma unroll
        for (int q = 0; q < 3; ++q)
          st_stream(reinterpret_cast<float4*>(dst) + q,
                    make_float4((float)o[4 * q], (float)o[4 * q + 1], (float)o[4 * q + 2], (float)o[4 * q + 3]));
      } else {
        for (int q = 0; q < 3 * n; ++q) dst[q] = (float)o[q];
      }
    }
  }
}

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
  const int tiles = (HW + kTilePx - 1) / kTilePx;
  LA3D_REQUIRE((long long)tiles * B < (1ll << 31), "grid too large");
  const bool vec = (HW % 4 == 0) && aligned16(depth) && aligned16(out);
  dim3 grid((unsigned)(tiles * B)), block(kThreads);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
#define LAUNCH(F64, VEC) \
  lift_kernel<F64, VEC><<<grid, block, 0, s>>>(depth, K, k_stride, k_is_inverse, R, t, HW, W, tiles, out)
  if (out_f64) { if (vec) LAUNCH(true, true); else LAUNCH(true, false); }
  else         { if (vec) LAUNCH(false, true); else LAUNCH(false, false); }
#undef LAUNCH
  LA3D_CUDA(cudaGetLastError());
  return LA3D_OK;
}
