// The tail every box kernel ends with: from the extents of the ground-aligned, yaw-rotated footprint to the
// packed record.  estimate_bbox, src/util_3dbox.py:160-176 of the reference (float16-rounded corners,
// back-rotation with the reference's Rg / Rg^T convention, centre, dimensions, R_cam), project_to_2d
// (src/util.py:227-229) on the 8 corners and their 2D bounds (src/tools/combine_results.py:241-246).
// Called by all threads of the CTA (at least 41); `rec` is a 64-double shared-memory record that is
// complete (and published by a block barrier) on return - the caller hands it to sink_store (sink.cuh).
// One copy for fit.cu and fit_all.cu: the records of both are asserted bit-identical to the round-1 build
// that carried these statements inline twice (tests/test_gpu_regress.py).
#pragma once

#include <math_constants.h>

#include "common.cuh"

namespace la3d {

__device__ __forceinline__ void write_box_record(const double (&dim)[3], const double (&ctr)[3], double yaw, double cy_,
                                                 double sy_, const double* __restrict__ Rg,
                                                 const double* __restrict__ Kmat, bool has_K, double* __restrict__ rec,
                                                 int n_valid, long long n_src, int tid, double flags = 0.0) {
  // rotate_y(-yaw) = [[c,0,-s],[0,1,0],[s,0,c]] with c = cos(yaw), s = sin(yaw)
  const double Ry[9] = {cy_, 0.0, -sy_, 0.0, 1.0, 0.0, sy_, 0.0, cy_};
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
    for (int i = 0; i < 3; ++i) v2[i] = v1[0] * Rg[i * 3] + v1[1] * Rg[i * 3 + 1] + v1[2] * Rg[i * 3 + 2];
#pragma unroll
    for (int i = 0; i < 3; ++i) rec[LA3D_O_VERT + tid * 3 + i] = v2[i];
    // project_to_2d (util.py:227-229)
    double uu = CUDART_NAN, vv = CUDART_NAN;
    if (has_K) {
      const double h0 = Kmat[0] * v2[0] + Kmat[1] * v2[1] + Kmat[2] * v2[2];
      const double h1 = Kmat[3] * v2[0] + Kmat[4] * v2[1] + Kmat[5] * v2[2];
      const double h2 = Kmat[6] * v2[0] + Kmat[7] * v2[1] + Kmat[8] * v2[2];
      uu = h0 / h2; vv = h1 / h2;
    }
    rec[LA3D_O_UV + tid * 2] = uu;
    rec[LA3D_O_UV + tid * 2 + 1] = vv;
  } else if (tid == 32) {
    // center_cam = Rg^T @ (rotate_y(-yaw) @ c)  (:172-173; Rg^T where the corners used Rg - kept)
    double w[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) w[i] = Ry[i * 3] * ctr[0] + Ry[i * 3 + 1] * ctr[1] + Ry[i * 3 + 2] * ctr[2];
#pragma unroll
    for (int i = 0; i < 3; ++i) rec[LA3D_O_CENTER + i] = Rg[i] * w[0] + Rg[3 + i] * w[1] + Rg[6 + i] * w[2];
    rec[LA3D_O_DIM] = dim[2]; rec[LA3D_O_DIM + 1] = dim[1]; rec[LA3D_O_DIM + 2] = dim[0];
    rec[LA3D_O_YAW] = yaw;
    rec[LA3D_O_NVALID] = (double)n_valid;
    rec[LA3D_O_STATUS] = (double)LA3D_ST_OK;
    rec[LA3D_O_NMASK] = (double)n_src;
    rec[LA3D_O_PAD] = flags;
  } else if (tid == 40) {
    // R_cam = Rg^T @ rotate_y(-yaw)  (:176)
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int k = 0; k < 3; ++k)
        rec[LA3D_O_RCAM + i * 3 + k] = Rg[i] * Ry[k] + Rg[3 + i] * Ry[3 + k] + Rg[6 + i] * Ry[6 + k];
  }
  __syncthreads();
  if (tid == 0) {
    // Python min()/max() over the 8 projections (combine_results.py:241-246): sequential
    // `if x < m` / `if x > m`, so a leading NaN sticks and a later NaN is skipped.
    double mnu = rec[LA3D_O_UV], mnv = rec[LA3D_O_UV + 1], mxu = mnu, mxv = mnv;
    for (int j = 1; j < 8; ++j) {
      const double uu = rec[LA3D_O_UV + 2 * j], vv = rec[LA3D_O_UV + 2 * j + 1];
      if (uu < mnu) mnu = uu;
      if (vv < mnv) mnv = vv;
      if (uu > mxu) mxu = uu;
      if (vv > mxv) mxv = vv;
    }
    rec[LA3D_O_BOX2D] = mnu; rec[LA3D_O_BOX2D + 1] = mnv;
    rec[LA3D_O_BOX2D + 2] = mxu; rec[LA3D_O_BOX2D + 3] = mxv;
  }
  __syncthreads();
}

}  // namespace la3d
