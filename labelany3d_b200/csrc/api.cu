// Library-level entry points of libla3d_sm100a: version, error text, and the
// one-call pipeline (mask scan with the preparation riding in its launch -> subsample ranks -> fit)
// with its workspace.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include "prep.cuh"
#include "sink.cuh"

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
// ---- record sinks (sink.cuh) ---------------------------------------------------------------
// A peer that does not reach a flag within this time is fatal (sticky status word + trap).  Far above ordinary
// rank skew (first-call lazy initialisation, data loading, host preemption); LA3D_PEER_TIMEOUT_MS or
// la3d_set_peer_timeout_ms() change it.
static long long g_peer_timeout_ms = -1;
unsigned long long peer_timeout_ns() {
  if (g_peer_timeout_ms < 0) {
    const char* env = getenv("LA3D_PEER_TIMEOUT_MS");
    g_peer_timeout_ms = (env && atoll(env) > 0) ? atoll(env) : 120000;
  }
  return (unsigned long long)g_peer_timeout_ms * 1000000ull;
}

RecordSink local_sink(void* records, int rec_f64) {
  RecordSink s{};
  s.out[0] = records;
  s.n_out = 1;
  s.rec_f64 = rec_f64;
  return s;
}

int sink_from_public(const la3d_sink* pub, RecordSink* out) {
  LA3D_REQUIRE(pub && out, "null pointer (sink)");
  LA3D_REQUIRE(pub->n_out >= 1 && pub->n_out <= LA3D_MAX_PEERS, "between 1 and LA3D_MAX_PEERS destinations");
  RecordSink s{};
  for (int p = 0; p < pub->n_out; ++p) {
    LA3D_REQUIRE(pub->records[p] != nullptr, "null destination buffer");
    LA3D_REQUIRE((reinterpret_cast<uintptr_t>(pub->records[p]) & 15u) == 0, "destination buffers must be 16-byte aligned");
    s.out[p] = pub->records[p];
  }
  s.n_out = pub->n_out;
  s.rec_f64 = pub->rec_f64 ? 1 : 0;
  if (pub->flags[0]) {
    LA3D_REQUIRE(pub->counter != nullptr, "peer synchronisation needs the local counter word");
    LA3D_REQUIRE(pub->rank >= 0 && pub->rank < pub->n_out, "rank outside the destinations");
    LA3D_REQUIRE(pub->epoch != 0, "epochs start at 1");
    for (int p = 0; p < pub->n_out; ++p) {
      LA3D_REQUIRE(pub->flags[p] != nullptr, "null flag row");
      s.flags[p] = pub->flags[p];
    }
    s.counter = pub->counter;
    s.status = pub->status;
    s.epoch = pub->epoch;
    s.rank = pub->rank;
    s.timeout_ns = peer_timeout_ns();
  }
  *out = s;
  return LA3D_OK;
}

static int fit_boxes_sink(const float* depth, const uint8_t* masks, const double* K, const double* ground, int B, int I,
                          int H, int W, int mask_is_01, int method, int yaw_steps, uint32_t seed,
                          uint32_t image_offset, void* workspace, size_t workspace_bytes, const RecordSink& sink,
                          la3d_stream_t stream) {
  LA3D_REQUIRE(depth && masks && K && workspace, "null pointer");
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
  return fit_scanned_sink(depth, w.prep, w.bits, w.chunk_counts, w.ranks, B, I, H, W, method, yaw_steps, sink,
                          static_cast<cudaStream_t>(stream), pdl_enabled());
}

// Cross-GPU flag operations over peer memory (the fit kernels do both halves themselves, sink.cuh; these
// are the stand-alone forms).  signal: rank r stores `epoch` into slot r of every rank's flag row (release,
// system scope).  wait: until every slot of the own row has reached `epoch`.  Epochs only grow, so the flags
// never need a reset.  A peer that does not arrive within the timeout is fatal (wait_flag).
__global__ void peer_sync_kernel(RecordSink s, int do_signal, int do_wait) {
  const int p = threadIdx.x;
  if (p >= s.n_out) return;
  if (do_signal) {
    __threadfence_system();
    st_release_sys(s.flags[p] + s.rank, s.epoch);
  }
  if (do_wait) wait_flag(s.flags[s.rank] + p, s.epoch, s.status, s.timeout_ns);
}

static int peer_sync(uint32_t* const* flags, int rank, int world, uint32_t epoch, int32_t* status, int do_signal,
                     int do_wait, la3d_stream_t stream) {
  LA3D_REQUIRE(flags, "null pointer");
  LA3D_REQUIRE(world >= 1 && world <= LA3D_MAX_PEERS && rank >= 0 && rank < world, "bad rank / world");
  RecordSink s{};
  for (int p = 0; p < world; ++p) {
    LA3D_REQUIRE(flags[p] != nullptr || (!do_signal && p != rank), "null flag row");
    s.flags[p] = flags[p];
  }
  s.n_out = world; s.rank = rank; s.epoch = epoch; s.status = status; s.timeout_ns = peer_timeout_ns();
  peer_sync_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(s, do_signal, do_wait);
  LA3D_CUDA(cudaGetLastError());
  return LA3D_OK;
}
}  // namespace la3d

extern "C" void la3d_set_peer_timeout_ms(long long ms) { la3d::g_peer_timeout_ms = ms > 0 ? ms : 120000; }

extern "C" int la3d_fit_boxes(const float* depth, const uint8_t* masks, const double* K, const double* ground, int B,
                              int I, int H, int W, int mask_is_01, int method, int yaw_steps, uint32_t seed,
                              uint32_t image_offset, void* workspace, size_t workspace_bytes, void* records,
                              int rec_f64, la3d_stream_t stream) {
  using namespace la3d;
  LA3D_REQUIRE(records, "null pointer");
  return fit_boxes_sink(depth, masks, K, ground, B, I, H, W, mask_is_01, method, yaw_steps, seed, image_offset,
                        workspace, workspace_bytes, local_sink(records, rec_f64), stream);
}

extern "C" int la3d_fit_boxes_to(const float* depth, const uint8_t* masks, const double* K, const double* ground, int B,
                                 int I, int H, int W, int mask_is_01, int method, int yaw_steps, uint32_t seed,
                                 uint32_t image_offset, void* workspace, size_t workspace_bytes, const la3d_sink* sink,
                                 la3d_stream_t stream) {
  using namespace la3d;
  RecordSink rs;
  if (int rc = sink_from_public(sink, &rs)) return rc;
  return fit_boxes_sink(depth, masks, K, ground, B, I, H, W, mask_is_01, method, yaw_steps, seed, image_offset,
                        workspace, workspace_bytes, rs, stream);
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

namespace la3d {
static int fit_boxes_rle_sink(const float* depth, const uint32_t* run_counts, const int64_t* run_offsets, int max_runs,
                              uint32_t* ends_ws, const double* K, const double* ground, int B, int I, int H, int W,
                              int method, int yaw_steps, uint32_t seed, uint32_t image_offset, void* workspace,
                              size_t workspace_bytes, int32_t* rle_status, const RecordSink& sink, la3d_stream_t stream) {
  LA3D_REQUIRE(depth && run_counts && run_offsets && K && workspace && rle_status, "null pointer");
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
  return fit_scanned_sink(depth, w.prep, w.bits, w.chunk_counts, w.ranks, B, I, H, W, method, yaw_steps, sink, s, false);
}
}  // namespace la3d

extern "C" int la3d_fit_boxes_rle(const float* depth, const uint32_t* run_counts, const int64_t* run_offsets, int max_runs,
                                  uint32_t* ends_ws, const double* K, const double* ground, int B, int I, int H, int W,
                                  int method, int yaw_steps, uint32_t seed, uint32_t image_offset, void* workspace,
                                  size_t workspace_bytes, int32_t* rle_status, void* records, int rec_f64,
                                  la3d_stream_t stream) {
  using namespace la3d;
  LA3D_REQUIRE(records, "null pointer");
  return fit_boxes_rle_sink(depth, run_counts, run_offsets, max_runs, ends_ws, K, ground, B, I, H, W, method, yaw_steps,
                            seed, image_offset, workspace, workspace_bytes, rle_status, local_sink(records, rec_f64), stream);
}

extern "C" int la3d_fit_boxes_rle_to(const float* depth, const uint32_t* run_counts, const int64_t* run_offsets,
                                     int max_runs, uint32_t* ends_ws, const double* K, const double* ground, int B, int I,
                                     int H, int W, int method, int yaw_steps, uint32_t seed, uint32_t image_offset,
                                     void* workspace, size_t workspace_bytes, int32_t* rle_status, const la3d_sink* sink,
                                     la3d_stream_t stream) {
  using namespace la3d;
  RecordSink rs;
  if (int rc = sink_from_public(sink, &rs)) return rc;
  return fit_boxes_rle_sink(depth, run_counts, run_offsets, max_runs, ends_ws, K, ground, B, I, H, W, method, yaw_steps,
                            seed, image_offset, workspace, workspace_bytes, rle_status, rs, stream);
}

// ---- every masked pixel instead of the 500-point subsample: two launches (scan with the preparation CTAs, dense fit)
namespace la3d {
static int fit_boxes_all_sink(const float* depth, const uint8_t* masks, const double* K, const double* ground, int B,
                              int I, int H, int W, int mask_is_01, int method, int yaw_steps, void* workspace,
                              size_t workspace_bytes, const RecordSink& sink, la3d_stream_t stream) {
  LA3D_REQUIRE(depth && masks && K && workspace, "null pointer");
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
  return fit_all_sink(depth, w.prep, w.bits, B, I, H, W, method, yaw_steps, sink, static_cast<cudaStream_t>(stream));
}
}  // namespace la3d

extern "C" int la3d_fit_boxes_all(const float* depth, const uint8_t* masks, const double* K, const double* ground, int B,
                                  int I, int H, int W, int mask_is_01, void* workspace, size_t workspace_bytes,
                                  void* records, int rec_f64, la3d_stream_t stream) {
  using namespace la3d;
  LA3D_REQUIRE(records, "null pointer");
  return fit_boxes_all_sink(depth, masks, K, ground, B, I, H, W, mask_is_01, LA3D_METHOD_PCA, 0, workspace,
                            workspace_bytes, local_sink(records, rec_f64), stream);
}

extern "C" int la3d_fit_boxes_all_to(const float* depth, const uint8_t* masks, const double* K, const double* ground,
                                     int B, int I, int H, int W, int mask_is_01, int method, int yaw_steps,
                                     void* workspace, size_t workspace_bytes, const la3d_sink* sink, la3d_stream_t stream) {
  using namespace la3d;
  RecordSink rs;
  if (int rc = sink_from_public(sink, &rs)) return rc;
  return fit_boxes_all_sink(depth, masks, K, ground, B, I, H, W, mask_is_01, method, yaw_steps, workspace,
                            workspace_bytes, rs, stream);
}

extern "C" int la3d_peer_signal(uint32_t* const* flags, int rank, int world, uint32_t epoch, la3d_stream_t stream) {
  return la3d::peer_sync(flags, rank, world, epoch, nullptr, 1, 0, stream);
}
extern "C" int la3d_peer_wait(uint32_t* const* flags, int rank, int world, uint32_t epoch, int32_t* status,
                              la3d_stream_t stream) {
  return la3d::peer_sync(flags, rank, world, epoch, status, 0, 1, stream);
}
extern "C" int la3d_peer_barrier(uint32_t* const* flags, int rank, int world, uint32_t epoch, int32_t* status,
                                 la3d_stream_t stream) {
  return la3d::peer_sync(flags, rank, world, epoch, status, 1, 1, stream);
}
