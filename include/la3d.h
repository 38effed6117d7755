/*
 * la3d.h - C ABI of libla3d_sm100a.so, the B200 (sm_100a) implementation of
 * LabelAny3D's per-object 3D box-fitting hot path.
 *
 * The reference is pure Python: its "plugin API" for this path is a handful of
 * module-level functions.  Each entry point below names
 * the reference function (path relative to the reference root, file:line) whose
 * arithmetic it replaces; labelany3d_b200/dropin/{util,util_3dbox,cam_utils}.py
 * re-create those functions, with the reference's signatures, on top of this
 * ABI through ctypes (INTEGRATION.md shows the binding).
 *
 * Conventions
 *   - every array argument is a DEVICE pointer unless the comment says "host";
 *   - every function is asynchronous on `stream` (a cudaStream_t passed as
 *     void*) and returns 0 on success or a negative LA3D_E* code; the text of
 *     the last failure on the calling thread is la3d_last_error();
 *   - shapes are row-major (C order); images are [H][W], pixel (v,u) has flat
 *     index p = v*W + u; a "plane" is one (image, instance) mask [H][W];
 *   - nothing here allocates device memory: callers pass workspaces.
 */
#ifndef LA3D_H_
#define LA3D_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LA3D_VERSION 200 /* 0.2.0 */

typedef void* la3d_stream_t; /* cudaStream_t */

/* error codes (return values) */
#define LA3D_OK 0
#define LA3D_EINVAL (-1) /* bad argument (null pointer, non-positive size, misaligned buffer) */
#define LA3D_ECUDA (-2)  /* CUDA runtime error, see la3d_last_error() */
#define LA3D_ENOMEM (-3) /* workspace too small */

/* yaw estimators */
#define LA3D_METHOD_PCA 0         /* src/util_3dbox.py:181-186 */
#define LA3D_METHOD_CONVEX_HULL 1 /* src/util_3dbox.py:189-224 */
#define LA3D_METHOD_SWEEP 2       /* new: uniform yaw sweep, SURVEY.md section 8 row a7 */

/* per-box status written into the record (a GPU batch cannot raise per box) */
#define LA3D_ST_OK 0
#define LA3D_ST_NO_VALID 1      /* ValueError("No valid points after removing NaN values"), util_3dbox.py:142-143 */
#define LA3D_ST_PCA_UNDEFINED 2 /* one point: scikit-learn PCA(2) refuses n_samples < 2 */
#define LA3D_ST_BAD_METHOD 3    /* ValueError("Unknown method ..."), util_3dbox.py:151 */
#define LA3D_ST_NONFINITE 4     /* +-inf left in the XZ footprint: scikit-learn's input check raises */
#define LA3D_ST_TOO_MANY 5      /* la3d_fit_points: more than 500 points and no sample_idx; all-pixels hull / sweep: more
                                   than 2048 hull candidates */

/* the reference keeps at most this many points per box (util_3dbox.py:123-125) */
#define LA3D_SUBSAMPLE 500

/* packed box record: 64 scalars (float or double), SURVEY.md section 8e */
#define LA3D_REC 64
#define LA3D_O_VERT 0    /* 8x3 corners, camera frame            util_3dbox.py:165-169 */
#define LA3D_O_CENTER 24 /* center_cam                           util_3dbox.py:172-173 */
#define LA3D_O_DIM 27    /* [dz, dy, dx]                         util_3dbox.py:175     */
#define LA3D_O_RCAM 30   /* R_cam row-major                      util_3dbox.py:176     */
#define LA3D_O_YAW 39
#define LA3D_O_NVALID 40 /* points that survived the NaN filter  util_3dbox.py:139-140 */
#define LA3D_O_STATUS 41
#define LA3D_O_UV 42     /* 8x2 projected corners                util.py:227-229       */
#define LA3D_O_BOX2D 58  /* [min u, min v, max u, max v]         tools/combine_results.py:241-246 */
#define LA3D_O_NMASK 62  /* pixels set in the mask / points given */
#define LA3D_O_PAD 63   /* flags: 0, or LA3D_FLAG_HULL_FALLBACK */
#define LA3D_FLAG_HULL_FALLBACK 1 /* method convex_hull: no hull (coincident / collinear points, Qhull raises in the
                                    reference, util_3dbox.py:222-224): the yaw is the PCA yaw */

int la3d_version(void);
const char* la3d_last_error(void);

/* ---------------------------------------------------------------------------
 * depth -> camera-space points.  Replaces depth_to_points, src/util.py:52-75.
 *   out[b][v][u][i] = sum_j R[i][j] * ( ((d*Kinv[i][0])*u + (d*Kinv[i][1])*v) + d*Kinv[i][2] ) + t[i]
 * evaluated in float64 in the reference's operation order; `out_f64` selects a
 * double (reference dtype) or float (rounded once from the double) output.
 *   depth  [B][H][W] float
 *   K      [B][9] (k_stride = 9) or one shared [9] (k_stride = 0), double.
 *          k_is_inverse = 0: intrinsics, inverted on the device by LU with partial
 *          pivoting (bit-identical to LAPACK for pinhole matrices);
 *          k_is_inverse = 1: the caller already inverted them (np.linalg.inv).
 *   R, t   nullable; [9] / [3] double shared by all images (the reference API
 *          takes one R, t)
 *   out    [B][H][W][3]
 * ------------------------------------------------------------------------- */
int la3d_depth_lift(const float* depth, const double* K, int k_stride, int k_is_inverse, const double* R,
                    const double* t, int B, int H, int W, void* out, int out_f64, la3d_stream_t stream);

/* ---------------------------------------------------------------------------
 * Mask-stack scan.  Replaces the NumPy boolean-gather bookkeeping of pts[mask]
 * (src/util.py:480-481 idiom; mask stack of src/util.py:382): one pass over the
 * [B][I][H][W] byte masks produces, per plane, a bit mask (bit k of word w is
 * pixel 32*w+k) and, for every 512-pixel chunk, the number of set pixels of its four
 * 128-pixel quarters packed as four bytes (byte q = quarter q, each 0..128), which
 * together locate the r-th set pixel in row-major order (= row r of pts[mask]).
 *   masks         [B*I][H*W] uint8; nonzero = set (mask_is_01 = 1 promises 0/1 bytes,
 *                 which is what numpy/torch bool arrays hold, and takes a shorter path)
 *   bits          [B*I][la3d_words_per_plane(H,W)] uint32
 *   chunk_counts  [B*I][la3d_chunks_per_plane(H,W)] uint32 (4 x uint8 quarter counts)
 * ------------------------------------------------------------------------- */
size_t la3d_chunks_per_plane(int H, int W);
size_t la3d_words_per_plane(int H, int W);
int la3d_mask_scan(const uint8_t* masks, int planes, int H, int W, int mask_is_01, uint32_t* bits,
                   uint32_t* chunk_counts, la3d_stream_t stream);

/* The same scan as a "thin" persistent kernel: ctas_per_sm CTAs per SM whose loads are issued by the
 * TMA unit (cp.async.bulk into a ring of `stages` 16 KB shared-memory stages, mbarrier-signalled),
 * so it reaches the same bandwidth with ~1/4 of the resident threads and leaves room on every SM for
 * the sampler / fit kernels of another batch to be resident beside it (DESIGN.md section 4 reports what
 * that overlap measured).
 * Needs H*W % 512 == 0 and a 16-byte aligned stack (LA3D_EINVAL otherwise: use la3d_mask_scan). */
int la3d_mask_scan_thin(const uint8_t* masks, int planes, int H, int W, int mask_is_01, uint32_t* bits,
                        uint32_t* chunk_counts, int ctas_per_sm, int stages, la3d_stream_t stream);

/* ---------------------------------------------------------------------------
 * Mask statistics on the bit planes of la3d_mask_scan ("next" row f1 of the scope table): the
 * integer bookkeeping of analyze_mask (src/util.py:291-326), get_maximum_height
 * (src/util.py:328-335), the row count of src/util.py:369-370 and the overlap counts of
 * filter_component_masks (src/model_wrappers.py:33-37).
 *   bands  HOST array of 8 ints: rows [t0,t1) = top band, rows [b0,b1) = bottom band, columns
 *          [l0,l1) = left band, columns [r0,r1) = right band (the caller resolves the reference's
 *          Python slices `[:b]` / `[-b:]` against H and W)
 *   stats  [planes][8] int32 (out): area, pixels in the top / bottom / left / right band, first
 *          non-empty row (-1 if none), last non-empty row (-1 if none), number of non-empty rows
 *   la3d_mask_overlap: inter[p] = |bits[p] & other_bits[p / group]| (group = instances per image
 *          when every image has one foreground plane)
 * ------------------------------------------------------------------------- */
int la3d_mask_stats(const uint32_t* bits, int planes, int H, int W, const int* bands, int32_t* stats,
                    la3d_stream_t stream);
int la3d_mask_overlap(const uint32_t* bits, const uint32_t* other_bits, int planes, int group, int H, int W,
                      int32_t* inter, la3d_stream_t stream);

/* ---------------------------------------------------------------------------
 * Mask-independent preparation of a batch for the scanned-mask path.  Everything the
 * sampler and the fit kernel need that does not depend on the masks, so that it can run
 * concurrently with the mask scan (la3d_fit_boxes folds it into the scan's launch):
 *   - per image b, NumPy's legacy MT19937 seeded with (seed + image_offset + b) mod 2^32
 *     (np.random.seed) and its first words pre-generated, plus the state to continue from;
 *   - per image, the intrinsics and their inverse (np.linalg.inv(K), src/util.py:56);
 *   - per box, the ground rotation Rg of src/util_3dbox.py:128-134.
 *   K [B][9] double intrinsics; ground nullable [B*I][3] double
 *   prep: la3d_prep_bytes(B, I) bytes, 256-byte aligned (opaque)
 * la3d_set_mt_blocks(n): tuning / test hook - pre-generate n blocks of 624 words per image
 * instead of the automatic ceil(I*1024/624)+1 (0 restores the automatic choice).  Results
 * never depend on it (the sampler continues the generator when the words run out); it must
 * not change between la3d_prep_bytes / la3d_fit_prepare and the calls that consume `prep`.
 * ------------------------------------------------------------------------- */
size_t la3d_prep_bytes(int B, int I);
int la3d_fit_prepare(const double* K, const double* ground, int B, int I, uint32_t seed, uint32_t image_offset,
                     void* prep, size_t prep_bytes, la3d_stream_t stream);
void la3d_set_mt_blocks(int n);
/* Tuning hook: blocks of 624 words a sampler CTA stages in shared memory at a time (1 .. 16; 0 = the library's
 * choice).  Fewer blocks = less shared memory per CTA = more resident CTAs, at the price of refills; results
 * never depend on it. */
void la3d_set_sample_seg_blocks(int n);

/* ---------------------------------------------------------------------------
 * Per-image subsample ranks.  Replaces `np.random.randint(0, N, 500)` of
 * src/util_3dbox.py:123-125 for a whole batch: the instances of image b draw from
 * that image's stream (see la3d_fit_prepare) in order, each only if its count N exceeds
 * 500 (masked rejection on 32-bit draws, exactly RandomState.randint).
 *   counts [B*I] int32  (out) set pixels per plane
 *   ranks  [B*I][500] int32 (out) row indices into pts[mask]; untouched when N <= 500
 * ------------------------------------------------------------------------- */
int la3d_sample_ranks(const uint32_t* chunk_counts, const void* prep, int B, int I, int H, int W, int32_t* counts,
                      int32_t* ranks, la3d_stream_t stream);

/* ---------------------------------------------------------------------------
 * Oriented box per (image, instance) from depth + scanned masks.  Replaces the
 * composition depth_to_points -> pts[mask] -> estimate_bbox -> project_to_2d
 * (src/util.py:52-75, src/util_3dbox.py:106-224, src/util.py:227-229,
 * src/tools/combine_results.py:238-246) without materialising the point cloud.
 *   depth [B][H][W] float; prep: output of la3d_fit_prepare (cameras, ground rotations)
 *   bits / chunk_counts / ranks: outputs of la3d_mask_scan and la3d_sample_ranks
 *   records [B*I][64] float (rec_f64 = 0) or double (rec_f64 = 1)
 * ------------------------------------------------------------------------- */
/* Debug aid: when set (device pointer to [boxes][8] int64, or NULL to switch off), thread 0 of every CTA of the
 * scanned-mask fit kernel stores clock64() at 8 phase boundaries (start, prologue, gather, octagon, yaw, extents,
 * record, stores): per-phase latency without a profiler (tools/fit_phases.py). */
void la3d_debug_fit_clocks(long long* clocks);
/* The same for the sampler: [images][4] globaltimer nanoseconds (start, counts totalled, words staged, ranks written). */
void la3d_debug_sample_clocks(unsigned long long* clocks);
/* Measurement aid: 4 cudaEvent_t (or NULL to switch off).  While set, the fused steps (la3d_fit_boxes, la3d_fit_boxes_rle
 * and their _to forms, unsplit) record them on the caller's stream before the first launch, after the scan / decode
 * launch, after the sampler and after the fit: the durations of the step's OWN three launches without a profiler
 * (bench.py's roofline).  Process-wide, not thread-safe. */
void la3d_debug_step_events(void* const* events);
/* Measurement aid (tools/gather_ceiling.py): ctas x 256 threads each issue `rounds` rounds of `inflight` (1, 2, 4, 8, 16)
 * independent scattered 4-byte reads of src[0 .. n), every CTA inside its own `window` elements (window == n: anywhere);
 * share > 1: that many neighbouring lanes read one 128-byte line, `spread` elements apart (address-sorted samples);
 * out [ctas * 256] float receives the sums.  src may be device memory or pinned host memory: the latter measures the
 * read-request rate the in-place depth gather of la3d_fit_boxes_rle can reach over PCIe.  Replaces nothing in the reference. */
int la3d_debug_scatter_read(const float* src, size_t n, size_t window, int inflight, int rounds, int ctas, int share,
                            int spread, float* out, la3d_stream_t stream);
int la3d_fit_scanned(const float* depth, const void* prep, const uint32_t* bits, const uint32_t* chunk_counts,
                     const int32_t* ranks, int B, int I, int H, int W, int method, int yaw_steps, void* records,
                     int rec_f64, la3d_stream_t stream);

/* The four calls above as one pipeline of three launches on `stream`: the preparation rides in the
 * mask scan's launch (B extra CTAs at the front of its grid, so the latency-bound seeding hides
 * under the HBM-bound scan), then la3d_sample_ranks and la3d_fit_scanned.  `workspace` needs
 * la3d_fit_workspace_bytes() bytes, 256-byte aligned.
 * `depth` (here, in la3d_fit_scanned and in la3d_fit_boxes_rle) may be device memory or page-locked host memory
 * mapped for the device: the fit reads only 500 values per box, so a host-resident map is read in place over
 * PCIe, the samples of a box sorted by address so that the lanes of a warp share 128-byte line requests (same
 * records; LA3D_FIT_SORT_DEPTH=0 switches the sort off). */
/* Optional pipeline over parts of a batch (LA3D_PIPE_IMAGES=n or la3d_set_pipeline_images(n): n images per part,
 * 0 = never split = the default, -1 = back to the environment's choice): the scan of part p+1 on `stream` overlaps
 * the sampler and fit of part p on internal high-priority streams that fork from and join back into `stream`
 * (events only; the call stays asynchronous and graph-capturable).  Results do not depend on the split.  Off by
 * default because it measured slower on B200. */
void la3d_set_pipeline_images(int images_per_part);
size_t la3d_fit_workspace_bytes(int B, int I, int H, int W);
int la3d_fit_boxes(const float* depth, const uint8_t* masks, const double* K, const double* ground, int B, int I,
                   int H, int W, int mask_is_01, int method, int yaw_steps, uint32_t seed, uint32_t image_offset,
                   void* workspace, size_t workspace_bytes, void* records, int rec_f64, la3d_stream_t stream);

/* ---------------------------------------------------------------------------
 * Mask stacks from COCO run-length annotations (input side of the path).  Replaces the host-side
 * `mask_utils.decode(rle)` per annotation of read_bounding_boxes_segmentations, src/util.py:361-370
 * (COCO RLE: column-major runs alternating 0, 1, 0, ... - the format the reference's own encoder
 * binary_mask_to_rle, src/download_coconut.py:167-175, writes): one CTA per plane turns the runs
 * straight into the bit plane and quarter counts la3d_mask_scan would have produced from the decoded
 * byte mask, which therefore never exists (neither in HBM nor on the bus).
 *   run_counts   run lengths of all planes back to back, uint32
 *   run_offsets  [planes+1] int64: plane p owns run_counts[run_offsets[p] .. run_offsets[p+1])
 *   max_runs     upper bound on the runs of one plane the CALLER knows (sizes the shared-memory copy of
 *                the run ends); 0 = unknown
 *   ends_ws      nullable [total runs] uint32 scratch, used by planes with more than max_runs runs
 *                (or by all when max_runs is 0 or above 32768)
 *   bits / chunk_counts: as la3d_mask_scan
 *   status       [planes] int32 (out): 0 ok; 1 the runs cover more than H*W pixels (pycocotools 2.0
 *                would write out of bounds; here the in-image part is decoded); 2 more runs than
 *                max_runs and no ends_ws (plane left empty)
 * ------------------------------------------------------------------------- */
int la3d_rle_decode(const uint32_t* run_counts, const int64_t* run_offsets, int planes, int H, int W, int max_runs,
                    uint32_t* ends_ws, uint32_t* bits, uint32_t* chunk_counts, int32_t* status, la3d_stream_t stream);

/* The box fit from bit planes that already exist (la3d_rle_decode, or an earlier la3d_mask_scan):
 * la3d_fit_prepare -> la3d_sample_ranks -> la3d_fit_scanned on `stream`, workspace of
 * la3d_fit_bits_workspace_bytes(B, I) bytes, 256-byte aligned. */
size_t la3d_fit_bits_workspace_bytes(int B, int I);
int la3d_fit_boxes_bits(const float* depth, const uint32_t* bits, const uint32_t* chunk_counts, const double* K,
                        const double* ground, int B, int I, int H, int W, int method, int yaw_steps, uint32_t seed,
                        uint32_t image_offset, void* workspace, size_t workspace_bytes, void* records, int rec_f64,
                        la3d_stream_t stream);

/* la3d_fit_boxes with run-length masks as the input: three launches (decode with the preparation riding
 * in its grid, subsample ranks, fit).  `workspace` as for la3d_fit_boxes (la3d_fit_workspace_bytes);
 * rle_status [B*I] as la3d_rle_decode's `status`. */
int la3d_fit_boxes_rle(const float* depth, const uint32_t* run_counts, const int64_t* run_offsets, int max_runs,
                       uint32_t* ends_ws, const double* K, const double* ground, int B, int I, int H, int W, int method,
                       int yaw_steps, uint32_t seed, uint32_t image_offset, void* workspace, size_t workspace_bytes,
                       int32_t* rle_status, void* records, int rec_f64, la3d_stream_t stream);

/* ---------------------------------------------------------------------------
 * The box from EVERY masked pixel: estimate_bbox (src/util_3dbox.py:106-178) with the random 500-point draw
 * of :123-125 replaced by the identity - deterministic, no generator involved.  One CTA per box sweeps the set
 * pixels of the plane (float64, fixed summation order):
 *   method pca          centroid / covariance sums -> closed form of scikit-learn's PCA(2) -> extents in a
 *                       second sweep;
 *   method convex_hull  the footprint's hull candidates are filtered out of all pixels (octagon of the extreme
 *   / sweep             points, refined QuickHull-style until at most 2048 points survive), then the same hull-edge
 *                       search (util_3dbox.py:189-224, + the reference's +yaw rotation) / uniform sweep
 *                        as the sampled path.  A footprint with more than 2048 points on or near
 *                       its hull after 64 polygon vertices gets status LA3D_ST_TOO_MANY.
 * Same record as la3d_fit_scanned (status codes included; LA3D_O_NVALID / LA3D_O_NMASK count all pixels).
 *   la3d_fit_all_points: bits from la3d_mask_scan / la3d_rle_decode, prep from la3d_fit_prepare (method pca;
 *                        la3d_fit_all_points_to takes the method and a sink)
 *   la3d_fit_boxes_all:  byte masks in; two launches (scan with the preparation riding in it, dense fit);
 *                        workspace as for la3d_fit_boxes (method pca; la3d_fit_boxes_all_to takes the method)
 * ------------------------------------------------------------------------- */
int la3d_fit_all_points(const float* depth, const void* prep, const uint32_t* bits, int B, int I, int H, int W,
                        void* records, int rec_f64, la3d_stream_t stream);
int la3d_fit_boxes_all(const float* depth, const uint8_t* masks, const double* K, const double* ground, int B, int I,
                       int H, int W, int mask_is_01, void* workspace, size_t workspace_bytes, void* records, int rec_f64,
                       la3d_stream_t stream);

/* ---------------------------------------------------------------------------
 * Record sinks: where the records of a fit go.  Every fit entry point above has a `_to` form that takes a
 * la3d_sink instead of (records, rec_f64):
 *   - n_out = 1, flags[0] = NULL: one local buffer (what the plain forms do);
 *   - the multi-GPU form (images sharded across the GPUs of one NVLink / NVSwitch node, the reference's
 *     --start_index/--end_index/--gpu_idx split of src/batch_scripts/whole.py:25-27,42 followed by the merge of
 *     the per-process results): records[p] = rank p's gathered record buffer + THIS rank's slot (include the
 *     local one), peer-mapped; the box kernel stores every record straight into all n_out buffers over NVLink,
 *     so the all-gather of the packed records costs no extra pass.  The cross-GPU
 *     synchronisation rides in the step's own launches:
 *       flags[p]  rank p's flag row, >= n_out uint32, peer-mapped, zero-initialised once
 *       counter   unused (reserved)
 *       status    nullable int32 in host-visible (pinned) or device memory, zero-initialised once
 *       epoch     the step number: 1 for the first call, growing by one per call, the same on every rank
 *       rank      this rank's index (its column in every flag row)
 *     Before its first store a CTA of the box kernel waits until every slot of flags[rank] has reached epoch-1 (the
 *     peers are past the step that last read the buffers about to be overwritten: alternate between TWO gathered
 *     buffers per rank, and consume a gathered buffer on the stream that issues the next call).  The box kernel does
 *     NOT publish its own epoch: the kernel boundary after it guarantees its stores have been performed, so the
 *     first CTA of the NEXT `_to` call on the stream publishes epoch-1 (one release store per rank), and a consumer of
 *     the gathered records of step `epoch` first runs la3d_peer_barrier(flags, rank, n_out, epoch) (publish + wait).
 *     A rank without images in a step calls la3d_peer_signal instead of a fit.
 *     A peer that does not arrive within the timeout (default 120 s, LA3D_PEER_TIMEOUT_MS or
 *     la3d_set_peer_timeout_ms) is fatal: *status is set to 1 and the kernel traps, so the host sees a CUDA
 *     error instead of stale records.
 * All destination buffers must be 16-byte aligned (records are stored 16 bytes at a time).
 * ------------------------------------------------------------------------- */
#define LA3D_MAX_PEERS 8
typedef struct la3d_sink {
  void* records[LA3D_MAX_PEERS];
  uint32_t* flags[LA3D_MAX_PEERS];
  uint32_t* counter;
  int32_t* status;
  uint32_t epoch;
  int32_t n_out;
  int32_t rank;
  int32_t rec_f64;
} la3d_sink;

int la3d_fit_scanned_to(const float* depth, const void* prep, const uint32_t* bits, const uint32_t* chunk_counts,
                        const int32_t* ranks, int B, int I, int H, int W, int method, int yaw_steps,
                        const la3d_sink* sink, la3d_stream_t stream);
int la3d_fit_boxes_to(const float* depth, const uint8_t* masks, const double* K, const double* ground, int B, int I,
                      int H, int W, int mask_is_01, int method, int yaw_steps, uint32_t seed, uint32_t image_offset,
                      void* workspace, size_t workspace_bytes, const la3d_sink* sink, la3d_stream_t stream);
int la3d_fit_boxes_rle_to(const float* depth, const uint32_t* run_counts, const int64_t* run_offsets, int max_runs,
                          uint32_t* ends_ws, const double* K, const double* ground, int B, int I, int H, int W,
                          int method, int yaw_steps, uint32_t seed, uint32_t image_offset, void* workspace,
                          size_t workspace_bytes, int32_t* rle_status, const la3d_sink* sink, la3d_stream_t stream);
int la3d_fit_all_points_to(const float* depth, const void* prep, const uint32_t* bits, int B, int I, int H, int W,
                           int method, int yaw_steps, const la3d_sink* sink, la3d_stream_t stream);
int la3d_fit_boxes_all_to(const float* depth, const uint8_t* masks, const double* K, const double* ground, int B, int I,
                          int H, int W, int mask_is_01, int method, int yaw_steps, void* workspace,
                          size_t workspace_bytes, const la3d_sink* sink, la3d_stream_t stream);

/* Stand-alone flag operations on `stream` (one tiny launch each).  flags: HOST array of `world` device
 * pointers (the flag rows); la3d_peer_wait only dereferences flags[rank]. */
int la3d_peer_signal(uint32_t* const* flags, int rank, int world, uint32_t epoch, la3d_stream_t stream);
int la3d_peer_wait(uint32_t* const* flags, int rank, int world, uint32_t epoch, int32_t* status, la3d_stream_t stream);
int la3d_peer_barrier(uint32_t* const* flags, int rank, int world, uint32_t epoch, int32_t* status,
                      la3d_stream_t stream);
void la3d_set_peer_timeout_ms(long long ms);

/* ---------------------------------------------------------------------------
 * Oriented box from explicit point sets.  Replaces estimate_bbox,
 * src/util_3dbox.py:106-178, for the way the reference itself calls it (500
 * points sampled from a mesh, src/util_3dbox.py:269-278).
 *   pts        [total][3] double, box j owns rows offsets[j] .. offsets[j+1]-1
 *   offsets    [nboxes+1] int64
 *   sample_idx nullable [nboxes][500] int32: rows (relative to offsets[j]) to keep
 *              when a box has more than 500 points (the caller draws them, e.g.
 *              from np.random so the global stream advances as in the reference)
 *   K          nullable [nboxes][9] double intrinsics for the 2D reprojection
 *   ground     nullable [nboxes][3] double
 * ------------------------------------------------------------------------- */
int la3d_fit_points(const double* pts, const int64_t* offsets, const int32_t* sample_idx, const double* K,
                    const double* ground, int nboxes, int method, int yaw_steps, void* records, int rec_f64,
                    la3d_stream_t stream);

/* ---------------------------------------------------------------------------
 * Pinhole projection of explicit points.  Replaces project_to_2d,
 * src/util.py:227-229 (= src/tools/combine_results.py:105-108): [u,v] = (K p)[:2] / (K p)[2].
 *   pts [n][3] double; K [m][9] double; k_index nullable [n] int32 (which K each point uses;
 *   null = all use K[0]); uv [n][2] double (out)
 * ------------------------------------------------------------------------- */
int la3d_project_points(const double* pts, const double* K, const int32_t* k_index, long long n, double* uv,
                        la3d_stream_t stream);

/* ---------------------------------------------------------------------------
 * Combine stage ("next" row f2): src/tools/combine_results.py of the reference.
 * la3d_iou_matrix replaces iou2D (:111-123) over all pairs of two box lists, i.e. the cost
 * matrix of hungarian_matching (:130-134), for `groups` scenes in one launch:
 *   boxes0 [total0][4], boxes1 [total1][4] double xyxy; off0 / off1 [groups+1] int64 row offsets
 *   iou    (out) concatenated row-major [n0_g][n1_g] blocks, block g at iou + out_off[g]
 * Same operation order as the Python floats, no FMA: bit-identical.  The assignment itself
 * (scipy linear_sum_assignment, :137) stays on the host.
 * la3d_box2d_from_corners replaces :234-252: corners [n][8][3] double, K [m][9], k_index nullable
 * [n] int32 (which K / image size a box uses), wh [m][2] = (W, H) double ->
 * proj [n][4] = [min u, min v, max u, max v], trunc [n][4] = [max(0,.), max(0,.), min(W,.), min(H,.)]
 * with Python's min()/max() NaN behaviour.
 * ------------------------------------------------------------------------- */
int la3d_iou_matrix(const double* boxes0, const int64_t* off0, const double* boxes1, const int64_t* off1,
                    const int64_t* out_off, int groups, double* iou, la3d_stream_t stream);
int la3d_box2d_from_corners(const double* corners, const double* K, const int32_t* k_index, const double* wh, int n,
                            double* proj, double* trunc, la3d_stream_t stream);

/* ---------------------------------------------------------------------------
 * Depth-scale alignment ("next" row f3): the arithmetic of align_to_depth_match,
 * src/util.py:473-486: per plane p, over the pixels of mask_bits[p] & render_bits[p],
 * scale[p] = np.median(depth_map[p / group] / depth_render[p]) in float32 (exact radix
 * select; an even count averages the two middle values like np.median; NaN if any ratio is
 * NaN) and n_overlap[p] = number of such pixels (0: the reference returns the identity
 * transform, scale[p] is NaN).
 *   depth_map [planes/group][H][W] float; depth_render [planes][H][W] float
 *   mask_bits / render_bits: bit planes of la3d_mask_scan
 * ------------------------------------------------------------------------- */
int la3d_masked_ratio_median(const float* depth_map, const float* depth_render, const uint32_t* mask_bits,
                             const uint32_t* render_bits, int planes, int group, int H, int W, int32_t* n_overlap,
                             float* scale, la3d_stream_t stream);

/* ---------------------------------------------------------------------------
 * Depth-stage scale alignment ("next" row f3, second half): the arithmetic of align_depth,
 * src/batch_scripts/depth.py:52-92 = RANSACRegressor(LinearRegression(fit_intercept=False), min_samples=0.2).fit
 * on x = relative_depth[valid], y = metric_depth[valid] (float32, row-major order of the valid pixels).  The
 * host draws each trial's subset exactly as scikit-learn does and drives the loop (dropin/depth_align.py):
 *   la3d_ransac_subset_fit  sums[0] = sum x^2, sums[1] = sum x y over x[idx[k]], y[idx[k]], k < m (float64):
 *                           the least-squares slope through the origin is sums[1] / sums[0]
 *   la3d_ransac_classify    residual |y - x*coef| in float32 (no FMA, as NumPy evaluates y - X @ coef), inlier iff
 *                           <= threshold; stats[0..5] = n_inliers, sum y, sum y^2, sum residual^2, sum x^2, sum x y
 *                           over the inliers (float64; score and final refit)
 *   la3d_scale_fill         out[i] = mask[i] (or, mask NULL, !isinf(rel[i])) ? rel[i]*coef : fill   (:82-90)
 * The reference's slope comes out of LAPACK's float32 least squares, which is not restated: parity with the reference
 * is statistical (same subsets, slopes within ~1e-6 relative), parity with the oracle's restatement is exact.
 * ------------------------------------------------------------------------- */
int la3d_ransac_subset_fit(const float* x, const float* y, const int64_t* idx, long long m, double* sums,
                           la3d_stream_t stream);
int la3d_ransac_classify(const float* x, const float* y, long long n, float coef, float threshold, double* stats,
                         la3d_stream_t stream);
int la3d_scale_fill(const float* rel, const uint8_t* mask, long long n, float coef, float fill, float* out,
                    la3d_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* LA3D_H_ */
