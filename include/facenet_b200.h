/* facenet_b200 -- C ABI of the B200-native embedding-evaluation hot path.
 *
 * The reference (sMedX/FaceNet) has no FFI layer: the path is pure Python/NumPy in
 * facenet/statistics.py.  Every entry point below therefore names the reference
 * FUNCTION it replaces (file:line into /root/reference); the Python binding a
 * maintainer would add is shown in INTEGRATION.md and shipped as
 * facenet_b200/statistics.py (ctypes).
 *
 * Conventions
 *  - plain C: pointers, sizes and DLPack tensor structs only (no torch / numpy types);
 *  - tensors are BORROWED `DLTensor*` (the struct inside a DLPack capsule): kDLCPU data is
 *    staged through a ring of pinned buffers (pageable memory) or copied in place (pinned memory),
 *    kDLCUDA data on the handle's device is used in place;
 *  - embeddings: float32, C-contiguous [N, D], D a multiple of 64, 64 <= D <= 4096;
 *    labels: int32 or int64 [N];
 *  - every function returns FNB_OK (0) or an FNB_ERR_* code; fnb_last_error() gives the text;
 *  - a handle is not thread-safe; use one handle per thread / per GPU.  ctypes releases the
 *    GIL around each call.
 *  - there is NO CPU fallback: without a CUDA device fnb_create fails.
 */
#ifndef FACENET_B200_H
#define FACENET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- DLPack (ABI-compatible restatement of dlpack.h v0.8 structs) ------------------- */
#ifndef DLPACK_DLPACK_H_
typedef enum { kDLCPU = 1, kDLCUDA = 2, kDLCUDAHost = 3, kDLCUDAManaged = 13 } DLDeviceType;
typedef struct { int32_t device_type; int32_t device_id; } DLDevice;
typedef enum { kDLInt = 0, kDLUInt = 1, kDLFloat = 2, kDLBfloat = 4 } DLDataTypeCode;
typedef struct { uint8_t code; uint8_t bits; uint16_t lanes; } DLDataType;
typedef struct {
    void* data;
    DLDevice device;
    int32_t ndim;
    DLDataType dtype;
    int64_t* shape;
    int64_t* strides;      /* NULL = compact row-major */
    uint64_t byte_offset;
} DLTensor;
#endif

typedef struct fnb_context* fnb_handle;

enum {
    FNB_OK = 0,
    FNB_ERR_INVALID = 1,         /* bad argument (shape, dtype, contiguity, NULL) */
    FNB_ERR_NOT_NORMALIZED = 2,  /* statistics.py:40-42 -> ValueError('embeddings must be normalized to 1 ...') */
    FNB_ERR_BAD_METRIC = 3,      /* statistics.py:55,260 -> ValueError('Undefined similarity metric ...') */
    FNB_ERR_CUDA = 4,
    FNB_ERR_UNSUPPORTED = 5
};

/* Arithmetic of the Gram contraction (all accumulate in fp32 in tensor memory):
 *   FP16X3  x = hi + lo with hi, lo fp16 (pre-scaled by 2^8): hi*hi + hi*lo + lo*hi, three
 *           kind::f16 MMAs per k-step.  The operands carry 22 bits; what limits the accuracy is the tensor core's
 *           truncating accumulation: a bias of up to -3.3e-6 |s| at D = 512 (growing ~ D^1.09), which is calibrated
 *           and corrected (csrc/fnb_bias.h) -- after the correction |dd| <~ 3e-6.  The DEFAULT of the drop-in classes;
 *   TF32X3  same split with tf32 operands, kind::tf32;
 *   TF32    single kind::tf32 pass on RN-rounded operands (|dd| ~ 3.5e-5 rms);
 *   BF16    single kind::f16 pass on bf16 operands;  FP16: single pass on fp16 operands;
 *   FP16F8  x = hi + lo, hi fp16 (pre-scaled by 2^12): hi*hi in kind::f16 plus the cross terms
 *           e4m3(x)*e4m3(lo) + e4m3(lo)*e4m3(x) in kind::f8f6f4 at twice the MMA rate -- two fp16-pass
 *           equivalents instead of three.  Error (after the bias correction) is statistical: sigma(dd) ~ 0.9e-6 for
 *           dense 512-d rows, growing with the peakedness of the rows and with |s| (1.7e-6 at |s| ~ 1); histogram
 *           launches therefore run every tile that can hold a same-identity pair in FP16X3 (strict_tiles) and report
 *           fnb_stats.error_bound; needs D % 128 == 0. */
/*   AUTO    (histogram entry points only) FP16F8 when its error model holds for the data, else FP16X3: the e4m3
 *           rounding errors of the 2 x D cross products are independent and each is bounded by 2^-16 |x_i y_i|, so
 *           sigma(ds) ~ 2^-16.5 sqrt(sum_i x_i^2 y_i^2) <= 2^-16.5 max_row ||x||_4^2.  The split kernel measures
 *           max_row sum x_i^4 / (sum x_i^2)^2 ("peakedness": 3/D for dense Gaussian-like rows, 1 for one-hot rows);
 *           FP16F8 is used when it is <= 1/128 and D % 128 == 0 (a-priori gate), and its result is KEPT only when the
 *           a-posteriori fnb_stats.error_bound -- evaluated on the histogram it produced -- is <= eps; otherwise the
 *           launch is repeated in FP16X3 (fnb_stats.fallback).  fnb_stats.mode_used tells. */
enum { FNB_MODE_FP16X3 = 0, FNB_MODE_TF32X3 = 1, FNB_MODE_TF32 = 2, FNB_MODE_BF16 = 3, FNB_MODE_FP16 = 4, FNB_MODE_FP16F8 = 5,
       FNB_MODE_AUTO = 6 };

typedef struct {
    int32_t mode;          /* FNB_MODE_*                                        default FP16X3 */
    int32_t metric;        /* 0: 2*(1-s)   1: arccos(s)   (statistics.py:48-55)  default 0     */
    float   atol;          /* normalisation tolerance (statistics.py:22,40)      default 1e-5  */
    float   eps;           /* |d - threshold| <= eps pairs are counted            default 1e-5  */
    int32_t rank, world;   /* this process computes 1/world of the tiles (see shard_*)   default 0,1 */
    int32_t cta_group;     /* 0 auto, 1: 128x128 tiles per CTA, 2: 256x256 per CTA pair           */
    int32_t region_rows;   /* rows per super-row of the tile order, 0 = auto (row panels sized to half of L2) */
    const float* cuts;     /* optional [T]: per threshold the smallest fp32 similarity whose distance is
                              < threshold (+inf if none); NULL = computed here with libm (metric 1:
                              acosf).  The Python host passes NumPy's so that results are bit-exact
                              with NumPy's arccos. */
    int32_t max_ctas;      /* 0 = all SMs (debug / profiling knob) */
    int32_t force_checked; /* != 0: every tile takes the fully checked epilogue path (exact cut comparison, exact
                              eps window, per-pair range check) -- the reference the arithmetic path is tested against */
    int32_t debug;         /* profiling knob, results are MEANINGLESS when != 0: bit 0 skips the epilogue work,
                              bit 1 loads operands for the first tile only (later MMAs re-use that shared memory) */
    int32_t cluster_pairs; /* histogram launches, cta_group 2 only: 2 = clusters of two CTA pairs that share (multicast) the
                              A operand, so it is read from L2 once per 256 x 512 super-tile; 4 = clusters of 2 x 2 pairs on
                              a 512 x 512 super-tile, A and B each read once per two tiles (fp16f8 / fp16x3 modes);
                              1 = one pair per cluster; 0 = auto (for launches long enough to be power-limited,
                              >= 5e10 pairs per rank) */
    int32_t normalize;     /* rows are normalised while the operands are prepared (the raw norms are kept for `theta`):
                              1: x / |x|                      (facenet/faceclass.py:63-64, np.linalg.norm)
                              2: x * rsqrt(max(sum x^2, 1e-10))  (tf.nn.l2_normalize(axis=1, epsilon=1e-10),
                                 facenet/models/inception_resnet_v1.py:491-492)                  default 0 */
    float   theta;         /* normalize == 1 only: the pair distance becomes 2 (1 - s) + theta (2 (|x|-|y|) / (|x|+|y|))^2
                              (FaceToFaceDistanceClassifier.distance, facenet/faceclass.py:71)   default 0 */
    int32_t raw_distance;  /* != 0: classifier distance (facenet/faceclass.py:106-116): no range check, no clamp */
    int32_t subset_rows;   /* fnb_region_histogram_bins: perm / cls describe a SUBSET of subset_rows rows of emb (perm[r] indexes emb);
                              0 = all rows.  Lets a k-fold validation keep the embeddings resident on the GPU (one tensor, 20
                              subsets) instead of copying every fold's rows (facenet/statistics.py:284-305) */
    int32_t shard_mod, shard_lo, shard_width;
    const int32_t* shard_slots;
                           /* multi-GPU work split by ROW BLOCK: this rank computes the tiles of the row blocks rb (per super-row
                              of the tile order) whose residue rb % shard_mod is one of its shard_width residues: the host array
                              shard_slots[shard_width] (ascending), or, when that is NULL, the range [shard_lo, shard_lo +
                              shard_width).  shard_mod == 0 (default): mod = world and the single residue `rank` (equal shares).
                              Unequal widths give faster GPUs more rows; the ranks' residue sets must partition [0, shard_mod).
                              The integer bins do not depend on the split. */
    int32_t panel_window;  /* histogram launches: the clusters of the persistent Gram kernel publish the column panel they work
                              on, and one that is more than panel_window panels ahead of the slowest waits (bounded: it gives
                              up after a few ms, so this can cost time but never hang).  Keeps the clusters inside the same
                              column panels, which keeps the super-row's row panels in L2 over long launches (1M rows: DRAM
                              reads 824 -> 193-208 GB per launch, kernel 834 -> 800 ms).  1..7 = window, -1 = off, 0 = auto (whole-set
                              launches -- fnb_pair_histogram[_bins] -- whose share on this rank is >= 5e10 pairs, the ones long
                              enough to drift: window 2; off otherwise and for keyed launches, whose regions can be one tile wide).  Timing only:
                              the integer bins do not depend on it. */
    int32_t strict_tiles;  /* FP16F8 histogram launches: tiles that can hold a same-identity pair (and ragged / diagonal tiles) run
                              the FP16X3 contraction inside the same launch, so the pairs behind TP / FN never depend on the e4m3
                              error model.  0 = on (default), -1 = off (measurement knob) */
    int32_t bias_correction; /* the calibrated correction of the tensor core's accumulation bias (fnb_stats.error_bound, DESIGN.md
                              section 2): 0 = on (default), -1 = off (measurement knob: scripts/probe_bias.py) */
    int32_t streamed;      /* fnb_pair_histogram[_bins] over HOST embeddings whose labels are non-decreasing (rows already in class
                              order -- what np.concatenate over the classes gives, facenet/facenet.py:184-201): the upload is cut
                              into column chunks of the pair matrix and launch k (pairs whose column lies in chunk k) runs while
                              chunk k + 1 is copied, so only the first chunk's copy is exposed.  0 = auto (inputs >= 64 MiB),
                              1 = always, -1 = off (one copy, one launch); 3 (fnb_pair_histogram_sharded, pinned host rows): one DMA of the
                              whole shard + class-order gather on the device instead of the chunk-wise gather by the copy threads.
                              The integer bins do not depend on it. */
    int32_t tile_queue;    /* histogram launches: the clusters of the persistent Gram kernel take their tiles from ONE queue (an atomic
                              counter; tiles leave in schedule order, so all clusters stay inside the same column panels -- no
                              progress window needed -- and a cluster that starts late or runs slower takes fewer tiles), and on one
                              GPU a second launch of plain CTA pairs drains the same queue on the 16 SMs that a grid of 4-CTA
                              clusters cannot use (148 of 148 SMs busy).  0 = on (default), 2 = queue without the second launch,
                              -1 = off: static interleaved schedule (+ panel_window).  An explicit panel_window > 0 also selects
                              the static schedule.  Timing only: the integer bins do not depend on it. */
} fnb_options;

typedef struct {
    uint64_t n_pairs;      /* pairs binned by this rank                                         */
    uint64_t eps_window;   /* of those, pairs within eps of some threshold: exact on checked tiles; interior tiles
                              count pairs within eps_counted (eps <= eps_counted < 2.5 eps), an upper bound     */
    float    smin, smax;   /* exact range of raw similarities over the tiles that took the checked path
                              (diagonal / edge / same-identity tiles; NaN if none)                */
    float    max_abs;      /* bound on |similarity| over the remaining (interior) tiles: the largest squared row
                              norm (Cauchy-Schwarz); above 1 + atol every tile takes the checked path           */
    float    kernel_ms;    /* device time of the Gram kernel (CUDA events)                      */
    float    prepare_ms;   /* device time of sort/split/convert                                 */
    uint64_t tiles;        /* tiles processed by this rank                                      */
    uint32_t kernel_launches;
    float    eps_counted;  /* distance half-width actually counted by interior tiles (see eps_window)           */
    uint32_t grid_ctas;    /* CTAs of the Gram launch (co-resident clusters x cluster size)                     */
    int32_t  mode_used;    /* FNB_MODE_* the contraction ran in (what AUTO resolved to)                         */
    float    peakedness;   /* max over rows of sum x^4 / (sum x^2)^2 of the prepared embeddings                  */
    int32_t  panel_window; /* cluster-progress window the histogram launch ran with (0 = off; fnb_options.panel_window)  */
    float    error_bound;  /* a-posteriori bound on |dd| of any binned pair under the error model of mode_used (calibrated spread of
                              the arithmetic x a union bound over the pairs actually found in each similarity bin + the residual
                              of the bias correction); the contract needs <= eps.  AUTO re-runs in FP16X3 when FP16F8 exceeds it */
    int32_t  fallback;     /* 1: AUTO ran FP16F8, its error_bound exceeded eps and the launch was repeated in FP16X3 */
    float    h2d_ms;       /* device time of the host -> device copy of a kDLCPU embedding tensor (0: device-resident input).
                              Pageable memory goes through a ring of pinned slots filled by a pool of host threads while the
                              previous slot is in flight (csrc/fnb_stage.cu); pinned memory is copied in place */
    uint64_t h2d_bytes;    /* bytes copied host -> device by the call */
    float    gather_ms;    /* fnb_pair_histogram_sharded: device time from the first to the last piece of the row exchange on the copy
                              stream (uploads of this rank's rows + broadcasts); all but the first chunk of it runs under the launches */
    int32_t  streamed_chunks; /* launches of a streamed pass (fnb_options.streamed; 0: one upload, one launch); kernel_ms is then
                              the sum of the launches' durations */
} fnb_stats;

/* One rectangle of the pair matrix (rows/cols index the PERMUTED embedding order).  tri != 0:
 * row range == col range and only pairs col > row are counted. */
typedef struct {
    int32_t row_begin, row_end, col_begin, col_end;
    int32_t tri;           /* 2: like 1, and the diagonal elements (row == col) are binned too, into slot key + 1 */
    int32_t key;           /* histogram slot the rectangle accumulates into */
} fnb_region;

int  fnb_version(void);
void fnb_default_options(fnb_options* o);
int  fnb_create(int device, fnb_handle* out);
void fnb_destroy(fnb_handle h);
const char* fnb_last_error(fnb_handle h);          /* h may be NULL: error of the last failed fnb_create */
int  fnb_device_info(fnb_handle h, int* sm_count, int* cc_major, int* cc_minor, uint64_t* total_mem);
/* Enqueue all work of this handle on `cuda_stream` (a cudaStream_t, e.g. the framework's current stream) so that
 * it orders with the caller's kernels, collectives and events.  NULL is the legacy default stream (stream 0, the
 * "current stream" of torch / TensorFlow until another one is selected); FNB_STREAM_OWN restores the handle's own
 * non-blocking stream (the state after fnb_create).  Tensors handed over as kDLCUDA must be ready on that stream. */
#define FNB_STREAM_OWN ((void*)(intptr_t)-1)
int  fnb_set_stream(fnb_handle h, void* cuda_stream);

/* Replaces pairwise_similarities(xa, xb=None, metric, atol)  (facenet/statistics.py:22-57).
 *   xb == NULL: out = float32 [n(n-1)/2], strict upper triangle in row-major triu_indices(n,1) order;
 *   else      : out = float32 [na, nb].
 * out may be kDLCPU or kDLCUDA.  range[0..1] (host) receive min/max raw similarity.
 * FNB_ERR_NOT_NORMALIZED when a similarity is outside [-(1+atol), 1+atol] (range still filled). */
int fnb_pairwise(fnb_handle h, const DLTensor* xa, const DLTensor* xb, const fnb_options* opt,
                 DLTensor* out, float* range);

/* Whole-set verification histogram: for every unordered pair {a,b}, a != b, of emb [N,D], bin the
 * similarity against the T thresholds.  This is the inner statement of ConfidenceMatrix
 * (count_nonzero(sims < threshold), facenet/statistics.py:130-131) applied to the whole set with
 * same/different identity decided by labels (statistics.py:68-79,124-126).
 *   bins_out: uint64 [2, T+1] (kDLCPU or kDLCUDA): row 0 all pairs, row 1 same-identity pairs;
 *             bin k = pairs with exactly k of the ascending similarity cuts <= s.
 * Integer results: identical for any (world) split after summing ranks' bins. */
int fnb_pair_histogram_bins(fnb_handle h, const DLTensor* emb, const DLTensor* labels,
                            const double* thresholds, int T, const fnb_options* opt,
                            DLTensor* bins_out, fnb_stats* stats);

/* ---- multi-GPU inside the library: one process per GPU, one handle per process, NCCL over NVLink / NVSwitch -----------------
 * (north_star: "each GPU takes row blocks of the pair matrix, the small embedding matrix is all-gathered over NVLink with NCCL, and
 * the per-threshold count histograms are all-reduced".)  NCCL is resolved at run time: the libnccl.so.2 already loaded into the
 * process (torch's), else the system's, else the path in FNB_NCCL_LIB.
 *   fnb_comm_unique_id: rank 0 fills 128 bytes (an ncclUniqueId) and ships them to the other processes by any means;
 *   fnb_comm_init:      every process, same id; collective (returns when the communicator is up);
 *   fnb_comm_destroy:   releases it (fnb_destroy does too);  fnb_comm_info: rank / world / NCCL version in use.
 * fnb_comm_init also maps every rank's tile-queue counters into every rank through CUDA IPC (peer access over NVLink): in a sharded
 * job a rank drains its own queue (the row blocks rb % world == rank) and then takes tiles from the other ranks' queues, so a GPU
 * that runs slower under the power limit is helped out instead of holding up the all-reduce (fnb_comm_shared_queue tells whether
 * the mapping worked; without it every rank computes exactly its own row blocks).
 * fnb_comm_last_error: text of a failed fnb_comm_unique_id (which has no handle). */
int fnb_comm_unique_id(void* id128);
int fnb_comm_init(fnb_handle h, const void* id128, int rank, int world);
int fnb_comm_destroy(fnb_handle h);
int fnb_comm_info(fnb_handle h, int* rank, int* world, int* nccl_version);
int fnb_comm_shared_queue(fnb_handle h);   /* 1: the ranks can take tiles from each other's queues (counters mapped through CUDA IPC) */
const char* fnb_comm_last_error(void);

/* fnb_pair_histogram_bins over a set that is spread over the ranks of the handle's communicator: every rank passes ITS rows
 * (emb_shard [n_r, D], labels_shard [n_r]; host or device; the shards may differ in size; the set is their concatenation in
 * rank order) and receives the bins of the WHOLE set (already summed: ncclAllReduce of the [2, T+1] integer bins), identical
 * on every rank and identical to the one-GPU result.  Collective: every rank must call it with the same thresholds / options.
 *   exchange: 8 N bytes of labels first (every rank sorts the classes itself), then the fp32 rows by ncclBroadcast from their
 *   owner straight into place.  When the concatenation is already in class order (labels non-decreasing across the ranks) the
 *   exchange is cut into column chunks of the pair matrix and chunk k + 1 travels on the copy stream while launch k runs
 *   (fnb_options.streamed; fnb_stats.streamed_chunks, gather_ms): only the first chunk's transfer is exposed.  Otherwise all
 *   rows are gathered first.  opt->rank / world are taken from the communicator; shard_* (unequal shares) are honoured.
 *   A similarity out of range on ANY rank fails the call on EVERY rank (FNB_ERR_NOT_NORMALIZED), and AUTO's choice between
 *   FP16F8 and the strict pass is agreed between the ranks. */
int fnb_pair_histogram_sharded(fnb_handle h, const DLTensor* emb_shard, const DLTensor* labels_shard,
                               const double* thresholds, int T, const fnb_options* opt,
                               DLTensor* bins_out, fnb_stats* stats);

/* Test hook, host only (no GPU needed): the column-chunk plan of a streamed pass (fnb_options.streamed, fnb_pair_histogram_sharded)
 * over n rows with super-rows of rr rows and chunk granule g -- regions [cap][6] = {chunk, row_begin, row_end, col_begin, col_end,
 * tri}; returns the number of regions (negative: cap too small).  Every pair (row < col) lies in exactly one region. */
int fnb_debug_chunk_plan(long long n, long long rr, long long g, int cap, int* regions, int* nchunks);

/* Host-side conversion of (summed) bins into per-threshold counts:
 *   same_lt[n] = #{same-identity pairs with d < thresholds[n]}, diff_lt[n] likewise (strict <,
 *   float64 compare against the fp32 distance, statistics.py:131). */
int fnb_counts_from_bins(const double* thresholds, int T, const fnb_options* opt,
                         const uint64_t* bins /* host [2][T+1] */,
                         uint64_t* same_lt, uint64_t* diff_lt, uint64_t* n_same, uint64_t* n_diff);

/* Convenience, one GPU: bins + counts. */
int fnb_pair_histogram(fnb_handle h, const DLTensor* emb, const DLTensor* labels,
                       const double* thresholds, int T, const fnb_options* opt,
                       uint64_t* same_lt, uint64_t* diff_lt, uint64_t* n_same, uint64_t* n_diff,
                       fnb_stats* stats);

/* Keyed histogram over caller-defined rectangles of the pair matrix -- the building block of the
 * class-balanced ConfidenceMatrix (facenet/statistics.py:111-138) and of k-fold validation
 * (statistics.py:277-311).  perm [n] (host): row r of the permuted order is emb[perm[r]];
 * cls [n] (host): class id of permuted row r, NON-DECREASING in r; n = opt->subset_rows, or all rows of emb when 0.
 *   bins_host: uint64 [nkeys][2][T+1]. */
int fnb_region_histogram_bins(fnb_handle h, const DLTensor* emb, const int64_t* perm, const int32_t* cls,
                              const fnb_region* regions, int nregions, int nkeys,
                              const double* thresholds, int T, const fnb_options* opt,
                              uint64_t* bins_host, fnb_stats* stats);

/* Class-balanced confidence matrix from keyed bins, on the device (prefix scan over bins, fp64):
 *   tp[n] = sum_key w_same[key] * #{same pairs of key with d < thr[n]},  fn[n] = total_same - tp[n],
 *   fp/tn likewise with w_diff (statistics.py:133-138), then accuracy argmax (first maximum,
 *   statistics.py:296) and the FAR threshold by piecewise-linear interpolation of thresholds over
 *   fp_rates at far_target (statistics.py:299-302; 0 when max(fp_rates) < far_target).
 * Uses the bins left on the device by the last fnb_region_histogram_bins call on this handle. */
int fnb_confidence_from_last_bins(fnb_handle h, int nkeys, const double* w_same, const double* w_diff,
                                  const double* thresholds, int T, const fnb_options* opt, double far_target,
                                  double* tp, double* tn, double* fp, double* fn,
                                  int32_t* argmax_accuracy, double* far_threshold);

/* Weighted binary cross entropy of the pair classifier over one P x K batch, fused behind the Gram product
 * (binary_cross_entropy_loss(model(embeddings_batch), options), facenet/apps/train_classifier.py:60-84,109-110, with the
 * logits alpha * (threshold - distance) of facenet/faceclass.py:23-27):
 *   batch float32 [B, D], rows grouped by class (row i belongs to group i / examples_per_class, facenet/facenet.py:108-113);
 *   label 1 iff same group, on the strict upper triangle; pos_weight = len(labels) / sum(labels) - 1;
 *   opt->normalize / opt->theta select FaceToFaceDistanceClassifier (normalize 1, theta) or
 *   FaceToFaceNormalizedEmbeddingsClassifier (normalize 0) distances.
 *   out [5] (host): mean loss, d loss / d alpha, d loss / d threshold, d loss / d theta (what TensorFlow's autodiff hands
 *   the optimiser, train_classifier.py:127), pos_weight. */
int fnb_pair_cross_entropy(fnb_handle h, const DLTensor* batch, int examples_per_class, float alpha, float threshold,
                           const fnb_options* opt, double* out, fnb_stats* stats);

/* binary_cross_entropy_loss(logits, options) for a materialised float32 [B, B] logits matrix (train_classifier.py:60-84). */
int fnb_logits_cross_entropy(fnb_handle h, const DLTensor* logits, int examples_per_class, double* loss);

/* Triplet mining on P x K batches (NOT in the reference fork -- semantics defined in oracle/mining_oracle.py; batch layout
 * facenet/facenet.py:89-123; hardest pairs = within-class argmax / cross-class argmin of the distance, the commented search of
 * facenet/statistics.py:357-387).  Distances are metric 0.  The hardest positive / negative of every anchor are folded into
 * the Gram epilogue as running packed-key arg-extrema; the semi-hard selection reads the B x B distance strip of the batch.
 *
 * fnb_mine_batched: emb float32 [S * B, D] = S batches of B rows each (S = nbatches), labels int32 / int64 [S * B]; every
 *   batch is mined on its own (all indices written are LOCAL to the batch, 0 .. B-1; -1 = empty set).
 *     hardest_pos, hardest_neg: int32 [S * B];  kmax >= max class size - 1, or 0 for hardest-only mining (the strip is then
 *     not materialised);  pos_index, semi_hard, eligible: int32 [S * B, kmax] (NULL when kmax == 0);
 *     status: optional int32 [4]: {largest number of positives seen if > kmax else 0, min s, max s (float bits), 0}.
 *   When EVERY output tensor is kDLCUDA the call only enqueues work on the handle's stream and returns (no host
 *   synchronisation; a training loop keeps the indices on the GPU): errors that depend on the data (embeddings not
 *   normalised, kmax too small) are then reported by fnb_mine_check, which synchronises.  With host outputs the call
 *   synchronises and returns those errors itself.
 * fnb_mine: one batch, host outputs (the synchronous form).
 * fnb_mine_select_kth: for queries (anchor a = global row of the last mining call, positive p local to a's batch, k >= 0):
 *   out = the k-th negative n (ascending index, local) with fp32(d(a, n) - d(a, p)) < alpha, or -1 -- the candidate list upstream
 *   davidsandberg/facenet select_triplets indexes with np.random.randint; with eligible[] as the range of the draw it replays
 *   that selection.  Needs the strips of a preceding call with kmax > 0 on this handle.  int32 tensors, host or device. */
int fnb_mine(fnb_handle h, const DLTensor* emb, const DLTensor* labels, float alpha, const fnb_options* opt,
             int32_t* hardest_pos, int32_t* hardest_neg,
             int kmax, int32_t* pos_index, int32_t* semi_hard, int32_t* eligible, fnb_stats* stats);
int fnb_mine_batched(fnb_handle h, const DLTensor* emb, const DLTensor* labels, int nbatches, float alpha,
                     const fnb_options* opt, int kmax, DLTensor* hardest_pos, DLTensor* hardest_neg,
                     DLTensor* pos_index, DLTensor* semi_hard, DLTensor* eligible, DLTensor* status, fnb_stats* stats);
int fnb_mine_check(fnb_handle h, int32_t* status /* [4], may be NULL */, fnb_stats* stats);
int fnb_mine_select_kth(fnb_handle h, const DLTensor* anchors, const DLTensor* positives, const DLTensor* kth,
                        float alpha, DLTensor* out);

/* False pairs at one threshold (the hardest-pair search of the commented-out FalseExamples class, facenet/statistics.py:334-387):
 * every same-identity pair with distance > threshold and every different-identity pair with distance < threshold, found by a
 * filter epilogue of the Gram kernel (strict FP16X3 arithmetic) and appended to a compact list -- the N x N matrix is never
 * materialised.  rows / cols: int32 [capacity] ORIGINAL row indices of the pair (row's class rank <= col's), dist: float32
 * [capacity], all host; *count = number of pairs found (may exceed capacity: only the first `capacity` appended are stored, in
 * no particular order -- call again with a larger capacity).  The reference's greedy per-class / per-class-pair top-k selection
 * runs on this list on the host (facenet_b200/statistics.py: FalseExamples). */
int fnb_false_pairs(fnb_handle h, const DLTensor* emb, const DLTensor* labels, double threshold, const fnb_options* opt,
                    long long capacity, int32_t* rows, int32_t* cols, float* dist, uint64_t* count, fnb_stats* stats);

#ifdef __cplusplus
}
#endif
#endif /* FACENET_B200_H */
