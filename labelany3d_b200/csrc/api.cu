// Library-level entry points of libla3d_sm100a: version, error text, and the
// one-call pipeline (mask scan with the preparation riding in its launch -> subsample ranks -> fit)
// with its workspace.
#include <cstdarg>
#include <cstdio>
#include "prep.cuh"

namespace la3d {
namespace {
thread_local char g_error[512] = "";
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t err, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)err, cudaGetErrorString(err), what);
  return LA3D_ECUDA;
}

struct Workspace {
  uint32_t* bits;
  uint32_t* chunk_counts;
  int32_t* counts;
  int32_t* ranks;
  void* prep;
  size_t prep_bytes;
  size_t bytes;
};

// All sub-buffers are 256-byte aligned relative to the workspace base.
static Workspace carve(void* base, int B, int I, int H, int W) {
  auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const size_t planes = (size_t)B * I;
  const size_t words = la3d_words_per_plane(H, W), chunks = la3d_chunks_per_plane(H, W);
  unsigned char* p = static_cast<unsigned char*>(base);
  Workspace w{};
  size_t off = 0;
  w.bits = reinterpret_cast<uint32_t*>(p + off);          off = up(off + planes * words * 4);
  w.chunk_counts = reinterpret_cast<uint32_t*>(p + off);  off = up(off + planes * chunks * 4);
  w.counts = reinterpret_cast<int32_t*>(p + off);         off = up(off + planes * 4);
  w.ranks = reinterpret_cast<int32_t*>(p + off);          off = up(off + planes * LA3D_SUBSAMPLE * 4);
  w.prep = p + off;  w.prep_bytes = la3d_prep_bytes(B, I); off = up(off + w.prep_bytes);
  w.bytes = off;
  return w;
}

}  // namespace la3d

extern "C" int la3d_version(void) { return LA3D_VERSION; }
extern "C" const char* la3d_last_error(void) { return la3d::g_error; }

extern "C" size_t la3d_fit_workspace_bytes(int B, int I, int H, int W) {
  if (B <= 0 || I <= 0 || H <= 0 || W <= 0) return 0;
  return la3d::carve(nullptr, B, I, H, W).bytes;
}

extern "C" int la3d_fit_boxes(const float* depth, const uint8_t* masks, const double* K, const double* ground, int B,
                              int I, int H, int W, int mask_is_01, int method, int yaw_steps, uint32_t seed,
                              uint32_t image_offset, void* workspace, size_t workspace_bytes, void* records,
                              int rec_f64, la3d_stream_t stream) {
  using namespace la3d;
  LA3D_REQUIRE(depth && masks && K && workspace && records, "null pointer");
  LA3D_REQUIRE(B > 0 && I > 0 && H > 0 && W > 0, "non-positive shape");
  LA3D_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, "workspace must be 256-byte aligned");
  const Workspace w = carve(workspace, B, I, H, W);
  if (workspace_bytes < w.bytes) {
    set_error("la3d_fit_boxes: workspace of %zu bytes, %zu needed", workspace_bytes, w.bytes);
    return LA3D_ENOMEM;
  }
  // one launch: the scan CTAs plus B CTAs that prepare the batch (MT19937 words, cameras, ground
  // rotations) under it
  LA3D_REQUIRE(I <= 8192, "at most 8192 instances per image");
  const PrepView pv = prep_view(w.prep, B, I, prep_blocks(I));
  const PrepArgs pa{K, ground, B, I, seed + image_offset, pv};
  int rc = launch_mask_scan(masks, B * I, H, W, mask_is_01, w.bits, w.chunk_counts, &pa, static_cast<cudaStream_t>(stream));
  if (rc) return rc;
  rc = la3d_sample_ranks(w.chunk_counts, w.prep, B, I, H, W, w.counts, w.ranks, stream);
  if (rc) return rc;
  return la3d_fit_scanned(depth, w.prep, w.bits, w.chunk_counts, w.ranks, B, I, H, W, method, yaw_steps, records,
                          rec_f64, stream);
}
