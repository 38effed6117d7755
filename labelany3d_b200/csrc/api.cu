// Library-level entry points of libla3d_sm100a: version, error text, and the
// one-call pipeline (mask scan with the preparation riding in its launch -> subsample ranks -> fit)
// with its workspace.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
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

// LA3D_PDL=1 launches the sampler and the fit as programmatic dependents (their launch latency and the
// fit's prologue overlap the kernel before them).  Off by default: measured on B200 it costs 17 us per
// step (204 vs 187 us) - the early-scheduled CTAs sit in griddepcontrol.wait on slots the scan's last
// waves would have used.
static bool pdl_enabled() {
  static const bool v = getenv("LA3D_PDL") && atoi(getenv("LA3D_PDL")) != 0;
  return v;
}

int cuda_fail(cudaError_t err, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)err, cudaGetErrorString(err), what);
  return LA3D_ECUDA;
}

struct PeerFlags {
  uint32_t* flags[LA3D_MAX_PEERS];
};

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

namespace la3d {
static int fit_boxes_multi(const float* depth, const uint8_t* masks, const double* K, const double* ground, int B, int I,
                           int H, int W, int mask_is_01, int method, int yaw_steps, uint32_t seed,
                           uint32_t image_offset, void* workspace, size_t workspace_bytes, void* const* records,
                           int n_out, int rec_f64, void* wait_before_fit, la3d_stream_t stream) {
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
  rc = launch_sample(w.chunk_counts, pv, B, I, (int)la3d_chunks_per_plane(H, W), w.counts, w.ranks,
                     static_cast<cudaStream_t>(stream), pdl_enabled());
  if (rc) return rc;
  // multi-GPU: the records go into peer buffers that may still be read from the step before last;
  // the caller's event (the peer barrier of the previous step) gates only the fit, so that barrier
  // and the skew between ranks hide under this step's scan and sampler
  if (wait_before_fit)
    LA3D_CUDA(cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), static_cast<cudaEvent_t>(wait_before_fit), 0));
  return fit_scanned_multi(depth, w.prep, w.bits, w.chunk_counts, w.ranks, B, I, H, W, method, yaw_steps, records, n_out,
                           rec_f64, static_cast<cudaStream_t>(stream), pdl_enabled() && !wait_before_fit);
}

// Cross-GPU barrier over peer memory: rank r stores `epoch` into slot r of every peer's flag array
// (release, system scope), then waits until every slot of its own array has reached `epoch`.
// Epochs only grow, so the flags never need a reset.  status[0] is set to 1 if a peer does not show up
// within ~2 s (the kernel returns instead of hanging the GPU).
__global__ void peer_barrier_kernel(PeerFlags pf, int rank, int world, uint32_t epoch, int* status) {
  const int p = threadIdx.x;
  if (p >= world) return;
  __threadfence_system();
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(pf.flags[p] + rank), "r"(epoch) : "memory");
  const uint32_t* mine = pf.flags[rank] + p;
  const long long t0 = clock64();
  for (;;) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
    if ((int32_t)(v - epoch) >= 0) break;
    if (clock64() - t0 > 4000000000ll) { if (status) *status = 1; break; }
    __nanosleep(64);
  }
}
}  // namespace la3d

extern "C" int la3d_fit_boxes(const float* depth, const uint8_t* masks, const double* K, const double* ground, int B,
                              int I, int H, int W, int mask_is_01, int method, int yaw_steps, uint32_t seed,
                              uint32_t image_offset, void* workspace, size_t workspace_bytes, void* records,
                              int rec_f64, la3d_stream_t stream) {
  return la3d::fit_boxes_multi(depth, masks, K, ground, B, I, H, W, mask_is_01, method, yaw_steps, seed, image_offset,
                               workspace, workspace_bytes, &records, 1, rec_f64, nullptr, stream);
}

// ---- the path from bit planes (no byte masks): annotations decoded on the device, or planes kept from an earlier scan
namespace la3d {
struct BitsWorkspace {
  int32_t* counts;
  int32_t* ranks;
  void* prep;
  size_t prep_bytes;
  size_t bytes;
};
static BitsWorkspace carve_bits(void* base, int B, int I) {
  auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const size_t planes = (size_t)B * I;
  unsigned char* p = static_cast<unsigned char*>(base);
  BitsWorkspace w{};
  size_t off = 0;
  w.counts = reinterpret_cast<int32_t*>(p + off);  off = up(off + planes * 4);
  w.ranks = reinterpret_cast<int32_t*>(p + off);   off = up(off + planes * LA3D_SUBSAMPLE * 4);
  w.prep = p + off;  w.prep_bytes = la3d_prep_bytes(B, I);  off = up(off + w.prep_bytes);
  w.bytes = off;
  return w;
}
}  // namespace la3d

extern "C" size_t la3d_fit_bits_workspace_bytes(int B, int I) {
  if (B <= 0 || I <= 0) return 0;
  return la3d::carve_bits(nullptr, B, I).bytes;
}

extern "C" int la3d_fit_boxes_bits(const float* depth, const uint32_t* bits, const uint32_t* chunk_counts, const double* K,
                                   const double* ground, int B, int I, int H, int W, int method, int yaw_steps,
                                   uint32_t seed, uint32_t image_offset, void* workspace, size_t workspace_bytes,
                                   void* records, int rec_f64, la3d_stream_t stream) {
  using namespace la3d;
  LA3D_REQUIRE(depth && bits && chunk_counts && K && workspace && records, "null pointer");
  LA3D_REQUIRE(B > 0 && I > 0 && H > 0 && W > 0, "non-positive shape");
  LA3D_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, "workspace must be 256-byte aligned");
  const BitsWorkspace w = carve_bits(workspace, B, I);
  if (workspace_bytes < w.bytes) {
    set_error("la3d_fit_boxes_bits: workspace of %zu bytes, %zu needed", workspace_bytes, w.bytes);
    return LA3D_ENOMEM;
  }
  int rc = la3d_fit_prepare(K, ground, B, I, seed, image_offset, w.prep, w.prep_bytes, stream);
  if (rc) return rc;
  rc = la3d_sample_ranks(chunk_counts, w.prep, B, I, H, W, w.counts, w.ranks, stream);
  if (rc) return rc;
  return la3d_fit_scanned(depth, w.prep, bits, chunk_counts, w.ranks, B, I, H, W, method, yaw_steps, records, rec_f64, stream);
}

extern "C" int la3d_fit_boxes_rle(const float* depth, const uint32_t* run_counts, const int64_t* run_offsets, int max_runs,
                                  uint32_t* ends_ws, const double* K, const double* ground, int B, int I, int H, int W,
                                  int method, int yaw_steps, uint32_t seed, uint32_t image_offset, void* workspace,
                                  size_t workspace_bytes, int32_t* rle_status, void* records, int rec_f64,
                                  la3d_stream_t stream) {
  using namespace la3d;
  LA3D_REQUIRE(depth && run_counts && run_offsets && K && workspace && rle_status && records, "null pointer");
  LA3D_REQUIRE(B > 0 && I > 0 && H > 0 && W > 0, "non-positive shape");
  LA3D_REQUIRE(I <= 8192, "at most 8192 instances per image");
  LA3D_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, "workspace must be 256-byte aligned");
  const Workspace w = carve(workspace, B, I, H, W);
  if (workspace_bytes < w.bytes) {
    set_error("la3d_fit_boxes_rle: workspace of %zu bytes, %zu needed", workspace_bytes, w.bytes);
    return LA3D_ENOMEM;
  }
  // one launch: a CTA per plane decodes its runs into bits + quarter counts, plus the CTAs that prepare the batch
  const PrepView pv = prep_view(w.prep, B, I, prep_blocks(I));
  const PrepArgs pa{K, ground, B, I, seed + image_offset, pv};
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int rc = launch_rle_decode(run_counts, run_offsets, B * I, H, W, max_runs, ends_ws, w.bits, w.chunk_counts, rle_status, &pa, s);
  if (rc) return rc;
  rc = launch_sample(w.chunk_counts, pv, B, I, (int)la3d_chunks_per_plane(H, W), w.counts, w.ranks, s, false);
  if (rc) return rc;
  void* rec = records;
  return fit_scanned_multi(depth, w.prep, w.bits, w.chunk_counts, w.ranks, B, I, H, W, method, yaw_steps, &rec, 1, rec_f64, s, false);
}

// ---- every masked pixel instead of the 500-point subsample: two launches (scan with the preparation CTAs, dense fit)
extern "C" int la3d_fit_boxes_all(const float* depth, const uint8_t* masks, const double* K, const double* ground, int B,
                                  int I, int H, int W, int mask_is_01, void* workspace, size_t workspace_bytes,
                                  void* records, int rec_f64, la3d_stream_t stream) {
  using namespace la3d;
  LA3D_REQUIRE(depth && masks && K && workspace && records, "null pointer");
  LA3D_REQUIRE(B > 0 && I > 0 && H > 0 && W > 0, "non-positive shape");
  LA3D_REQUIRE(I <= 8192, "at most 8192 instances per image");
  LA3D_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, "workspace must be 256-byte aligned");
  const Workspace w = carve(workspace, B, I, H, W);
  if (workspace_bytes < w.bytes) {
    set_error("la3d_fit_boxes_all: workspace of %zu bytes, %zu needed", workspace_bytes, w.bytes);
    return LA3D_ENOMEM;
  }
  const PrepView pv = prep_view(w.prep, B, I, prep_blocks(I));
  const PrepArgs pa{K, ground, B, I, 0u, pv};              // cameras and ground rotations; the random words go unused
  int rc = launch_mask_scan(masks, B * I, H, W, mask_is_01, w.bits, w.chunk_counts, &pa, static_cast<cudaStream_t>(stream));
  if (rc) return rc;
  return la3d_fit_all_points(depth, w.prep, w.bits, B, I, H, W, records, rec_f64, stream);
}

extern "C" int la3d_fit_boxes_p2p(const float* depth, const uint8_t* masks, const double* K, const double* ground, int B,
                                  int I, int H, int W, int mask_is_01, int method, int yaw_steps, uint32_t seed,
                                  uint32_t image_offset, void* workspace, size_t workspace_bytes,
                                  void* const* peer_records, int n_peers, int rec_f64, void* wait_before_fit,
                                  la3d_stream_t stream) {
  return la3d::fit_boxes_multi(depth, masks, K, ground, B, I, H, W, mask_is_01, method, yaw_steps, seed, image_offset,
                               workspace, workspace_bytes, peer_records, n_peers, rec_f64, wait_before_fit, stream);
}

extern "C" int la3d_peer_barrier(uint32_t* const* flags, int rank, int world, uint32_t epoch, int* status,
                                 la3d_stream_t stream) {
  using namespace la3d;
  LA3D_REQUIRE(flags, "null pointer");
  LA3D_REQUIRE(world >= 1 && world <= LA3D_MAX_PEERS && rank >= 0 && rank < world, "bad rank / world");
  PeerFlags pf{};
  for (int p = 0; p < world; ++p) {
    LA3D_REQUIRE(flags[p] != nullptr, "null flag array");
    pf.flags[p] = flags[p];
  }
  peer_barrier_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(pf, rank, world, epoch, status);
  LA3D_CUDA(cudaGetLastError());
  return LA3D_OK;
}
