// Library-level entry points of libla3d_sm100a: version, error text, and the
// one-call pipeline (mask scan with the preparation riding in its launch -> subsample ranks -> fit)
// with its workspace.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <mutex>
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
static int pdl_mode() {
  static const int v = getenv("LA3D_PDL") ? atoi(getenv("LA3D_PDL")) : 0;   // 1: sampler and fit, 2: the fit only
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
    LA3D_REQUIRE(pub->rank >= 0 && pub->rank < pub->n_out, "rank outside the destinations");
    LA3D_REQUIRE(pub->epoch != 0, "epochs start at 1");
    for (int p = 0; p < pub->n_out; ++p) {
      LA3D_REQUIRE(pub->flags[p] != nullptr, "null flag row");
      s.flags[p] = pub->flags[p];
    }
    s.status = pub->status;
    s.epoch = pub->epoch;
    s.rank = pub->rank;
    s.timeout_ns = peer_timeout_ns();
  }
  *out = s;
  return LA3D_OK;
}

// ---- the step as a pipeline over parts of the batch ------------------------------------------
// The mask scan (or the run-length decode) is the HBM-bound pass; the sampler and the fit that follow are
// latency- and issue-bound tails that need nothing from the NEXT images.  A batch of more than one part is
// therefore cut along the image axis: the scan of part p+1 runs on the caller's stream while the sampler and
// the fit of part p run on a high-priority side stream (two of them, alternating, so that the tails of two
// parts may also overlap each other).  The side streams fork from and join back into the caller's stream with
// events, so the call is still "asynchronous on `stream`" and can be captured into a CUDA graph.
// Measured on B200 (profiles/r2_f_pipe_sweep.json; 2048 images x 8 masks, 36-step sweep): unsplit 1350 us, parts of
// 512 / 256 / 128 / 64 images 1456 / 1502 / 1497 / 1806 us - the scan holds every thread slot and register of an SM
// (8 CTAs x 256 threads x 32 registers), so a resident fit CTA displaces scan CTAs and the loads of the latency-bound
// fit queue behind a saturated HBM; both kernels lose more than the overlap wins.  The split is therefore OFF by
// default: LA3D_PIPE_IMAGES = images per part (default 0 = never split) or la3d_set_pipeline_images().
constexpr int kMaxParts = 32;
struct Pipe {
  cudaStream_t side[2] = {nullptr, nullptr};
  cudaEvent_t scanned[kMaxParts] = {};
  cudaEvent_t done[2] = {nullptr, nullptr};
  bool ready = false;
  std::mutex mu;
};
static Pipe g_pipes[64];

static int pipe_images() {
  static const int v = getenv("LA3D_PIPE_IMAGES") ? atoi(getenv("LA3D_PIPE_IMAGES")) : 0;
  return v;
}
static int g_pipe_override = -1;                       // la3d_set_pipeline_images (tests, tuning)

static int pipe_get(Pipe** out) {
  int dev = 0;
  LA3D_CUDA(cudaGetDevice(&dev));
  LA3D_REQUIRE(dev >= 0 && dev < 64, "device index out of range");
  Pipe& p = g_pipes[dev];
  if (!p.ready) {
    int lo = 0, hi = 0;
    LA3D_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));     // hi = numerically lowest = greatest priority
    for (int k = 0; k < 2; ++k) {
      LA3D_CUDA(cudaStreamCreateWithPriority(&p.side[k], cudaStreamNonBlocking, hi));
      LA3D_CUDA(cudaEventCreateWithFlags(&p.done[k], cudaEventDisableTiming));
    }
    for (int k = 0; k < kMaxParts; ++k) LA3D_CUDA(cudaEventCreateWithFlags(&p.scanned[k], cudaEventDisableTiming));
    p.ready = true;
  }
  *out = &p;
  return LA3D_OK;
}

static PrepView prep_part(const PrepView& pv, int b0, int I) {
  PrepView v = pv;
  v.cams += b0;
  v.Rg += (size_t)b0 * I * 9;
  v.state += (size_t)b0 * kMtN;
  v.words += (size_t)b0 * pv.nblk * kMtN;
  return v;
}

static cudaEvent_t g_step_events[4] = {nullptr, nullptr, nullptr, nullptr};

// `produce(b0, Bp, prep_args)` launches the pass that turns images [b0, b0 + Bp) into bit planes + quarter counts
// (with the preparation CTAs of those images in its grid) on the caller's stream.
template <typename Produce>
static int run_step(Produce&& produce, const float* depth, const double* K, const double* ground, int B, int I, int H,
                    int W, int method, int yaw_steps, uint32_t seed0, const Workspace& w, RecordSink sink,
                    cudaStream_t stream) {
  const PrepView pv = prep_view(w.prep, B, I, prep_blocks(I));
  const int chunks = (int)la3d_chunks_per_plane(H, W);
  // the first launch of this step publishes the PREVIOUS step's epoch to the peers (sink.cuh)
  PeerPublish pub{};
  if (sink.flags[0] && sink.epoch > 1u) {
    for (int p = 0; p < sink.n_out; ++p) pub.flags[p] = sink.flags[p];
    pub.n = sink.n_out; pub.rank = sink.rank; pub.epoch = sink.epoch - 1u;
  }
  const int per = g_pipe_override >= 0 ? g_pipe_override : pipe_images();
  int parts = per > 0 ? (B + per - 1) / per : 1;
  if (parts > kMaxParts) parts = kMaxParts;
  if (parts <= 1) {
    const PrepArgs pa{K, ground, B, I, seed0, pv, pub};
    cudaEvent_t* ev = g_step_events[0] ? g_step_events : nullptr;       // la3d_debug_step_events
    if (ev) LA3D_CUDA(cudaEventRecord(ev[0], stream));
    int rc = produce(0, B, pa);
    if (rc) return rc;
    if (ev) LA3D_CUDA(cudaEventRecord(ev[1], stream));
    rc = launch_sample(w.chunk_counts, pv, B, I, chunks, w.counts, w.ranks, stream, pdl_mode() == 1);
    if (rc) return rc;
    if (ev) LA3D_CUDA(cudaEventRecord(ev[2], stream));
    rc = fit_scanned_sink(depth, w.prep, w.bits, w.chunk_counts, w.ranks, B, I, H, W, method, yaw_steps, sink, stream,
                          pdl_mode() != 0);
    if (rc) return rc;
    if (ev) LA3D_CUDA(cudaEventRecord(ev[3], stream));
    return LA3D_OK;
  }
  Pipe* pipe = nullptr;
  if (int rc = pipe_get(&pipe)) return rc;
  std::lock_guard<std::mutex> lock(pipe->mu);             // one enqueue sequence at a time per device
  const int step = (B + parts - 1) / parts;
  bool used[2] = {false, false};
  for (int p = 0, b0 = 0; b0 < B; ++p, b0 += step) {
    const int Bp = b0 + step <= B ? step : B - b0;
    const PrepArgs pa{K + (size_t)b0 * 9, ground ? ground + (size_t)b0 * I * 3 : nullptr, Bp, I, seed0 + (uint32_t)b0,
                      prep_part(pv, b0, I), p == 0 ? pub : PeerPublish{}};
    int rc = produce(b0, Bp, pa);
    if (rc) return rc;
    cudaStream_t side = pipe->side[p & 1];
    LA3D_CUDA(cudaEventRecord(pipe->scanned[p], stream));
    LA3D_CUDA(cudaStreamWaitEvent(side, pipe->scanned[p], 0));
    rc = launch_sample(w.chunk_counts, pv, Bp, I, chunks, w.counts, w.ranks, side, false, b0);
    if (rc) return rc;
    rc = fit_scanned_sink(depth, w.prep, w.bits, w.chunk_counts, w.ranks, B, I, H, W, method, yaw_steps, sink, side, false,
                          b0, Bp);
    if (rc) return rc;
    used[p & 1] = true;
  }
  for (int k = 0; k < 2; ++k) {
    if (!used[k]) continue;
    LA3D_CUDA(cudaEventRecord(pipe->done[k], pipe->side[k]));
    LA3D_CUDA(cudaStreamWaitEvent(stream, pipe->done[k], 0));
  }
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
  LA3D_REQUIRE(I <= 8192, "at most 8192 instances per image");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t HW = (size_t)H * W, words = la3d_words_per_plane(H, W), chunks = la3d_chunks_per_plane(H, W);
  // one launch per part: the scan CTAs plus the CTAs that prepare the part's images (MT19937 words, cameras,
  // ground rotations) under it
  auto produce = [&](int b0, int Bp, const PrepArgs& pa) {
    const size_t p0 = (size_t)b0 * I;
    // the workspace's bit planes are read by the fit's rank select only: chunks without set pixels are not stored
    return launch_mask_scan(masks + p0 * HW, Bp * I, H, W, mask_is_01, w.bits + p0 * words, w.chunk_counts + p0 * chunks, &pa, s, true);
  };
  return run_step(produce, depth, K, ground, B, I, H, W, method, yaw_steps, seed + image_offset, w, sink, s);
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
  if (do_wait) wait_flag<true>(s.flags[s.rank] + p, s.epoch, s.status, s.timeout_ns);
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

namespace la3d {
int publish_previous_epoch(const RecordSink& sink, cudaStream_t stream) {
  if (!sink.flags[0] || sink.epoch <= 1u) return LA3D_OK;
  RecordSink s = sink;
  s.epoch = sink.epoch - 1u;
  peer_sync_kernel<<<1, 32, 0, stream>>>(s, 1, 0);
  LA3D_CUDA(cudaGetLastError());
  return LA3D_OK;
}
}  // namespace la3d

extern "C" void la3d_debug_step_events(void* const* events) {
  for (int i = 0; i < 4; ++i) la3d::g_step_events[i] = events ? static_cast<cudaEvent_t>(events[i]) : nullptr;
}

extern "C" void la3d_set_pipeline_images(int images_per_part) { la3d::g_pipe_override = images_per_part; }
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
  // one launch per part: a CTA per plane decodes its runs into bits + quarter counts, plus the CTAs that prepare
  // the part's images
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t words = la3d_words_per_plane(H, W), chunks = la3d_chunks_per_plane(H, W);
  auto produce = [&](int b0, int Bp, const PrepArgs& pa) {
    const size_t p0 = (size_t)b0 * I;
    return launch_rle_decode(run_counts, run_offsets + p0, Bp * I, H, W, max_runs, ends_ws, w.bits + p0 * words,
                             w.chunk_counts + p0 * chunks, rle_status + p0, &pa, s, true);
  };
  return run_step(produce, depth, K, ground, B, I, H, W, method, yaw_steps, seed + image_offset, w, sink, s);
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
  PeerPublish pub{};
  if (sink.flags[0] && sink.epoch > 1u) {
    for (int p = 0; p < sink.n_out; ++p) pub.flags[p] = sink.flags[p];
    pub.n = sink.n_out; pub.rank = sink.rank; pub.epoch = sink.epoch - 1u;
  }
  const PrepArgs pa{K, ground, B, I, 0u, pv, pub};         // cameras and ground rotations; the random words go unused
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
