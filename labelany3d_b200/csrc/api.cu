// Library-level entry points of libla3d_sm100a: version, error text, and the
// one-call pipeline (mask scan -> subsample ranks -> fit) with its workspace.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "common.cuh"

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

int seed_states(int B, uint32_t seed0, uint32_t* states, cudaStream_t s);
int sample_seeded(const uint32_t* chunk_counts, int B, int I, int chunks, const uint32_t* states, int32_t* counts,
                  int32_t* ranks, cudaStream_t s);

struct Workspace {
  uint32_t* bits;
  uint32_t* chunk_counts;
  int32_t* counts;
  int32_t* ranks;
  uint32_t* mt_states;
  size_t bytes;
};

// The one-call pipeline cuts the batch into parts.  The scans of all parts run back to back on
// the caller's stream (they are the HBM-bound work); the sampler and the fit kernel of a part
// run on a high-priority side stream as soon as that part is scanned, so the latency-bound tail
// of part s hides under the scan of part s+1.  The MT19937 seeding of all images is started
// before the first scan.  Side streams and events are created once per device.
constexpr int kMaxParts = 8;
struct Pipeline {
  bool ready = false;
  cudaStream_t side[kMaxParts];
  cudaEvent_t fork, seeded, scanned[kMaxParts], done[kMaxParts];
  cudaEvent_t t0[kMaxParts], t1[kMaxParts];   // profiling only
  int last_parts = 0;
};
std::mutex g_mu;
Pipeline g_pipe[64];
bool g_profile = false;

int get_pipeline(Pipeline** out) {
  int dev = 0;
  LA3D_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) { set_error("device index %d out of range", dev); return LA3D_EINVAL; }
  Pipeline& p = g_pipe[dev];
  std::lock_guard<std::mutex> lock(g_mu);
  if (!p.ready) {
    int least = 0, greatest = 0;
    LA3D_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    for (int i = 0; i < kMaxParts; ++i) {
      LA3D_CUDA(cudaStreamCreateWithPriority(&p.side[i], cudaStreamNonBlocking, greatest));
      LA3D_CUDA(cudaEventCreateWithFlags(&p.scanned[i], cudaEventDisableTiming));
      LA3D_CUDA(cudaEventCreateWithFlags(&p.done[i], cudaEventDisableTiming));
      LA3D_CUDA(cudaEventCreate(&p.t0[i]));
      LA3D_CUDA(cudaEventCreate(&p.t1[i]));
    }
    LA3D_CUDA(cudaEventCreateWithFlags(&p.fork, cudaEventDisableTiming));
    LA3D_CUDA(cudaEventCreateWithFlags(&p.seeded, cudaEventDisableTiming));
    p.ready = true;
  }
  *out = &p;
  return LA3D_OK;
}

int choose_parts(int B) {
  static const int forced = getenv("LA3D_PARTS") ? atoi(getenv("LA3D_PARTS")) : 0;
  int parts = forced > 0 ? forced : 1;   // measured on B200: co-running the latency-bound kernels slows the scan more than it hides
  if (parts > kMaxParts) parts = kMaxParts;
  if (parts > B) parts = B;
  return parts < 1 ? 1 : parts;
}

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
  w.mt_states = reinterpret_cast<uint32_t*>(p + off);     off = up(off + (size_t)B * 624 * 4);
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
  const int parts = choose_parts(B);
  if (parts == 1) {
    int rc = la3d_mask_scan(masks, B * I, H, W, mask_is_01, w.bits, w.chunk_counts, stream);
    if (rc) return rc;
    rc = la3d_sample_ranks(w.chunk_counts, B, I, H, W, seed, image_offset, w.counts, w.ranks, stream);
    if (rc) return rc;
    return la3d_fit_scanned(depth, K, ground, w.bits, w.chunk_counts, w.counts, w.ranks, B, I, H, W, method,
                            yaw_steps, records, rec_f64, stream);
  }

  Pipeline* pp = nullptr;
  int rc = get_pipeline(&pp);
  if (rc) return rc;
  Pipeline& p = *pp;
  cudaStream_t main_s = static_cast<cudaStream_t>(stream);
  const size_t HW = (size_t)H * W, words = la3d_words_per_plane(H, W), chunks = la3d_chunks_per_plane(H, W);
  const size_t rec_bytes = rec_f64 ? 8 : 4;
  p.last_parts = parts;

  // side work is ordered after everything already queued on the caller's stream
  LA3D_CUDA(cudaEventRecord(p.fork, main_s));
  LA3D_CUDA(cudaStreamWaitEvent(p.side[0], p.fork, 0));
  rc = seed_states(B, seed + image_offset, w.mt_states, p.side[0]);
  if (rc) return rc;
  LA3D_CUDA(cudaEventRecord(p.seeded, p.side[0]));

  for (int s = 0; s < parts; ++s) {
    const int b0 = (int)((long long)B * s / parts), b1 = (int)((long long)B * (s + 1) / parts), nb = b1 - b0;
    const size_t pl0 = (size_t)b0 * I;
    if (g_profile) LA3D_CUDA(cudaEventRecord(p.t0[s], main_s));
    rc = la3d_mask_scan(masks + pl0 * HW, nb * I, H, W, mask_is_01, w.bits + pl0 * words, w.chunk_counts + pl0 * chunks,
                        stream);
    if (rc) return rc;
    if (g_profile) LA3D_CUDA(cudaEventRecord(p.t1[s], main_s));
    LA3D_CUDA(cudaEventRecord(p.scanned[s], main_s));
    LA3D_CUDA(cudaStreamWaitEvent(p.side[s], p.scanned[s], 0));
    if (s > 0) LA3D_CUDA(cudaStreamWaitEvent(p.side[s], p.seeded, 0));
    rc = sample_seeded(w.chunk_counts + pl0 * chunks, nb, I, (int)chunks, w.mt_states + (size_t)b0 * 624, w.counts + pl0,
                       w.ranks + pl0 * LA3D_SUBSAMPLE, p.side[s]);
    if (rc) return rc;
    rc = la3d_fit_scanned(depth + (size_t)b0 * HW, K + (size_t)b0 * 9, ground ? ground + pl0 * 3 : nullptr,
                          w.bits + pl0 * words, w.chunk_counts + pl0 * chunks, w.counts + pl0,
                          w.ranks + pl0 * LA3D_SUBSAMPLE, nb, I, H, W, method, yaw_steps,
                          static_cast<unsigned char*>(records) + pl0 * LA3D_REC * rec_bytes, rec_f64, p.side[s]);
    if (rc) return rc;
    LA3D_CUDA(cudaEventRecord(p.done[s], p.side[s]));
  }
  for (int s = 0; s < parts; ++s) LA3D_CUDA(cudaStreamWaitEvent(main_s, p.done[s], 0));
  return LA3D_OK;
}

extern "C" void la3d_set_profiling(int on) { la3d::g_profile = on != 0; }

extern "C" int la3d_last_scan_ms(float* ms, int max_parts) {
  using namespace la3d;
  Pipeline* pp = nullptr;
  int rc = get_pipeline(&pp);
  if (rc) return rc;
  LA3D_REQUIRE(ms && max_parts > 0, "bad argument");
  int n = pp->last_parts < max_parts ? pp->last_parts : max_parts;
  for (int s = 0; s < n; ++s) LA3D_CUDA(cudaEventElapsedTime(&ms[s], pp->t0[s], pp->t1[s]));
  return n;
}
