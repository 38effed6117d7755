// Where the records of a box kernel go: one local buffer, or this rank's slot in the gathered record
// buffer of EVERY rank of an NVLink / NVSwitch node (peer-mapped memory), with the cross-GPU
// synchronisation folded into the step's own launches:
//   acquire  before its first peer store a CTA of the box kernel makes sure every rank has published epoch - 1 in
//            this rank's flag row: the peers are then past the step that last READ the buffers about to be
//            overwritten (the gathered buffers are double-buffered by the caller).  A relaxed load: it only gates
//            later stores, and an acquire load would invalidate the SM's L1 under the other resident CTAs;
//   store    the 64 fields of a record, 16 bytes per store, to every destination;
//   release  NOT in the box kernel: the kernel boundary after it already guarantees that its stores have been
//            performed, so the epoch is published by the first CTA of the NEXT launch on the stream - the next
//            step's scan / decode (PeerPublish, prep.cuh) or the consumer's la3d_peer_barrier - with one release
//            store per rank.  (A fence + counter increment in every fit CTA, the first form, measured +6 us per step
//            on one GPU and its full-fence variant +30 us: MEMBAR.SYS / CCTL.IVALL once per box.)
// A consumer of the gathered records first runs la3d_peer_barrier: publish this rank's epoch, wait until every slot
// of ITS flag row has reached it.  No side stream, no event edge.
#pragma once

#include <math_constants.h>

#include "common.cuh"

namespace la3d {

struct RecordSink {
  void* out[LA3D_MAX_PEERS];          // record buffers: box j goes to out[p] + j * 64 elements, every p < n_out
  uint32_t* flags[LA3D_MAX_PEERS];    // flags[p] = rank p's flag row; flags[0] == nullptr: no peer synchronisation
  int32_t* status;                    // sticky error word (host-visible), set to 1 before a timeout trap
  unsigned long long timeout_ns;
  uint32_t epoch;
  int n_out, rank, rec_f64;
};

// host side (api.cu)
RecordSink local_sink(void* records, int rec_f64);
int sink_from_public(const la3d_sink* pub, RecordSink* out);
unsigned long long peer_timeout_ns();
// For entry points without an opening launch of their own (la3d_fit_scanned_to, la3d_fit_all_points_to): publish the
// previous step's epoch with a one-CTA launch before the box kernel.
int publish_previous_epoch(const RecordSink& sink, cudaStream_t stream);

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// no L1 invalidation behind it (an acquire load carries a CCTL.IVALL): for waits that only gate later STORES
__device__ __forceinline__ uint32_t ld_relaxed_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Spin until *flag has reached `epoch` (epochs only grow; compared modulo 2^32).  A peer that does not
// show up within the timeout is fatal: the sticky status word is set and the kernel traps, so the
// failure surfaces as a CUDA error on the host instead of as stale records.
// kAcquire: the wait precedes READS of what the peers wrote (la3d_peer_wait); otherwise it only gates later stores
// (the fit kernels' wait before they overwrite peer buffers) and a relaxed load does, which spares the SM's other
// CTAs the L1 invalidation an acquire load brings.
template <bool kAcquire>
__device__ __forceinline__ void wait_flag(const uint32_t* flag, uint32_t epoch, int32_t* status,
                                          unsigned long long timeout_ns) {
  auto peek = [&]() { return kAcquire ? ld_acquire_sys(flag) : ld_relaxed_sys(flag); };
  if ((int32_t)(peek() - epoch) >= 0) return;
  const unsigned long long t0 = global_ns();
  for (;;) {
    if ((int32_t)(peek() - epoch) >= 0) return;
    if (global_ns() - t0 > timeout_ns) {
      if (status) { *reinterpret_cast<volatile int32_t*>(status) = 1; __threadfence_system(); }
      __trap();
    }
    __nanosleep(100);
  }
}

// Called by every thread of the CTA before the first sink_store (contains a block barrier when the
// sink synchronises peers).
__device__ __forceinline__ void sink_acquire(const RecordSink& s) {
  if (!s.flags[0]) return;
  if (threadIdx.x < s.n_out) wait_flag<false>(s.flags[s.rank] + threadIdx.x, s.epoch - 1u, s.status, s.timeout_ns);
  __syncthreads();
}

// rec: the finished record, 64 doubles in shared memory.  Called by every thread of the CTA.
__device__ __forceinline__ void sink_store(const RecordSink& s, size_t box, const double* __restrict__ rec, int nthreads) {
  if (s.rec_f64) {
    constexpr int kPieces = LA3D_REC / 2;                       // 16 bytes = 2 doubles
    for (int job = threadIdx.x; job < s.n_out * kPieces; job += nthreads) {
      const int p = job / kPieces, c = job - p * kPieces;
      double2 v = make_double2(rec[2 * c], rec[2 * c + 1]);
      reinterpret_cast<double2*>(reinterpret_cast<double*>(s.out[p]) + box * LA3D_REC)[c] = v;
    }
  } else {
    constexpr int kPieces = LA3D_REC / 4;                       // 16 bytes = 4 floats
    for (int job = threadIdx.x; job < s.n_out * kPieces; job += nthreads) {
      const int p = job / kPieces, c = job - p * kPieces;
      float4 v = make_float4((float)rec[4 * c], (float)rec[4 * c + 1], (float)rec[4 * c + 2], (float)rec[4 * c + 3]);
      reinterpret_cast<float4*>(reinterpret_cast<float*>(s.out[p]) + box * LA3D_REC)[c] = v;
    }
  }
}

// The record of a box that could not be fitted (status != 0): NaN everywhere except the counters.
__device__ __forceinline__ void fill_failed_record(double* __restrict__ rec, int status, int n_valid, long long n_src,
                                                   int nthreads) {
  for (int f = threadIdx.x; f < LA3D_REC; f += nthreads) {
    double val = CUDART_NAN;
    if (f == LA3D_O_NVALID) val = (double)n_valid;
    if (f == LA3D_O_STATUS) val = (double)status;
    if (f == LA3D_O_NMASK) val = (double)n_src;
    if (f == LA3D_O_PAD) val = 0.0;
    rec[f] = val;
  }
}

// Called by every thread of the CTA after its last sink_store.  Nothing to do: the release happens at the next
// launch boundary (see the header comment).
__device__ __forceinline__ void sink_release(const RecordSink&) {}

}  // namespace la3d
