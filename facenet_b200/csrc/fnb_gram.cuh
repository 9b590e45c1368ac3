// facenet_b200 -- the Gram-matrix kernel: S = Xa * Xb^T on tcgen05 tensor cores with the
// consumer fused into the epilogue, so the N x N similarity matrix never reaches HBM.
//
// Replaces (reference, /root/reference/facenet/statistics.py):
//   :33,:36   xa @ xb.T                         -> TMA-fed tcgen05.mma, fp32 accumulators in TMEM
//   :40-46    range check + clamp               -> epilogue (max |s| / min,max; clamp)
//   :50,:53   2*(1-s) | arccos(s)               -> epilogue (PAIRWISE) or folded into similarity cuts (HIST)
//   :130-131  count_nonzero(sims < threshold)   -> epilogue per-thread histogram over similarity bins
//   :124-126  per-class-pair blocking           -> rows sorted by class, rectangles ("regions") with a key
//
// Structure (one CTA per SM, or one CTA PAIR per two SMs with cta_group::2; 640 threads per CTA):
//   warp 0      TMA producer  : 128-row x 128-byte operand boxes (SWIZZLE_128B) into a ring of 32 KB slots; boxes that two
//                               pairs of a cluster share are fetched once and multicast (kPairs = 2, 4)
//   warp 1      MMA issuer    : one elected lane issues tcgen05.mma (leader CTA only in pair mode)
//   warp 2      TMEM allocator
//   warps 4-19  epilogue      : tcgen05.ld the finished accumulator (double buffered in TMEM) and consume it while the
//                               next tile's MMAs run: HIST (per-threshold counts), PAIRWISE / ROWSTRIP (distances),
//                               BCE (pair-classifier cross entropy and its gradients)
// Tile = 128x128 (cta_group 1) or 256x256 per CTA pair (each CTA owns 128 rows x 256 columns).
#pragma once

#include "fnb_common.cuh"

namespace fnb {

constexpr int kMaxBins      = 128;                 // bins k in [0, T], T <= 127
constexpr int kEpiWarps     = 16;                  // four per TMEM lane quadrant: 4 warps per SM sub-partition hide LDS / LDTM latency
constexpr int kEpiThreads   = kEpiWarps * 32;
constexpr int kColSplit     = kEpiWarps / 4;       // accumulator columns are split four ways
constexpr int kFirstEpiWarp = 4;
constexpr int kGramThreads  = (kFirstEpiWarp + kEpiWarps) * 32;   // 640
constexpr int kRowsPerCta   = 128;                 // MMA M per CTA, and B rows loaded per CTA
constexpr int kBoxBytes     = kRowsPerCta * 128;   // one operand box: 128 rows x 128 B = 16 KB
constexpr int kSlotBytes    = 2 * kBoxBytes;       // {A part, B part}
constexpr int kMaxSlots     = 8;
constexpr int kHist8Row     = kEpiThreads;         // bytes per bin of the thread-private u8 counters

enum EpiKind : int { EPI_HIST = 0, EPI_PAIRWISE = 1, EPI_ROWSTRIP = 2, EPI_BCE = 3, EPI_FILTER = 4 };

struct RegionDev {
    int32_t row_begin, row_end, col_begin, col_end;
    int32_t tri, key;
    int32_t nrb, ncb;          // tile grid of the region (row blocks fastest)
    int32_t own_cnt;           // row blocks of the region that THIS rank owns (see ShardSpec)
    int32_t cp_begin;          // column panels of all earlier regions (a launch-wide column-panel counter: cp_begin + cb)
    long long tile_begin;      // first tile index of the region in this rank's own tile order (own_cnt * ncb tiles per region)
};

// Which row blocks of a region a rank owns: rb with rb % mod in the rank's set of `width` residues (a contiguous range or an
// explicit list).  The same subset in every column panel, so a rank re-reads only its own row panels from L2; width / mod
// is the rank's share of the work (equal shares: mod = world, the single residue `rank`; unequal ones follow measured
// per-GPU speed, with the residues of a rank spread evenly over [0, mod) so that the short rows at the bottom of a
// triangular region are shared out fairly -- facenet_b200/distributed.py).
struct ShardSpec {
    int mod, lo, width;
    const int32_t* slots;      // device array [width] of owned residues, ascending; NULL: the contiguous range [lo, lo + width)
    __device__ int block(int j) const {                        // j-th owned row block
        const int q = j / width, r = j - q * width;
        return q * mod + (slots ? __ldg(slots + r) : lo + r);
    }
};

struct GramParams {
    // schedule
    const RegionDev* regions;  // [nregions + 1], last entry is a sentinel with tile_begin = total_tiles
    int nregions;
    long long total_tiles;     // of this rank
    ShardSpec shard;
    int kblocks;               // D / (elements per 128 B)
    int num_slots;
    float acc_scale;           // similarity = accumulator * acc_scale
    int operand_fmt;           // kFmtF16 / kFmtBF16 / kFmtTF32 (must agree with the kTf32 template flag)
    int force_slow;            // take the fully-checked epilogue path for every tile
    int debug;                 // profiling knob (fnb_options.debug): bit 0 no epilogue work, bit 1 operands loaded for the first tile only
    const unsigned int* norm_max_ord;   // ordered-uint max squared row norm (written by the split kernel), may be NULL
    unsigned int norm_limit_ord;        // above this the interior tiles cannot be proven in range -> checked path
    // HIST epilogue
    const int32_t* row_cls;    // class id per (permuted) row, non-decreasing
    const int32_t* col_cls;
    const float* cuts;         // [kMaxBins] ascending similarity cuts, padded with +inf
    const float* wlo;          // [kMaxBins + 1] wlo[k]: lower edge of the eps window of cut k (+inf padded)
    const float* whi;          // [kMaxBins + 1] whi[k]: upper edge of the eps window of cut k-1 (whi[0] = -inf)
    int T;                     // number of cuts (bins 0..T)
    int T_fin;                 // number of finite cuts (the +inf ones sort last)
    int uniform;               // cuts are (numerically) an arithmetic progression -> arithmetic binning
    // arithmetic binning of the interior tiles (uniform != 0), see fast_bin():
    //   v  = sat(acc * f_s1 + f_b1)                      in [0, 1]  <->  (u + 0.5) / Q,  u = (s - cut_0) / h
    //   kw = v * f_k2 + 0.5 R  (R = 2^frac_bits)         = (u + 1) R  -> bin = kw >> frac_bits, clipped to T_fin
    float f_s1, f_b1, f_k2, f_magic_k, f_magic_n;
    unsigned int frac_bits, near_mask;
    // single-FMA form, used when no clipping is needed (noclip != 0):  word = bits(acc * f_g1 + f_g0) = magic + (u + 1) R;
    // near a threshold <=> ((word + near_half) & near_mask) == 0
    int noclip;
    float f_g1, f_g0;
    float f_g1x, f_g0x;        // the same map for the strict (fp16x3) tiles of an fp16f8 launch
    unsigned int near_half;
    int nb8;                   // bins of the u8 counters (T_fin + 3)
    unsigned long long* bins;  // [nkeys][2][bins_stride]  (0: all pairs, 1: same-identity pairs)
    int bins_stride;
    unsigned long long* counters;  // [0] eps-window pairs, [1] tiles processed
    unsigned int* range_ord;   // [0] min(s) [1] max(s) as ordered uints (checked tiles)
    // PAIRWISE / ROWSTRIP epilogue
    float* out;                // PAIRWISE: packed triangle or [na, nb]; ROWSTRIP: [n_rows, out_ld] distances
    long long out_ld;
    int tri_packed;
    int metric;
    int n_rows, n_cols;
    // classifier distances (facenet/faceclass.py): raw = no clamp; row_nrm / col_nrm = norms of the rows before they were
    // normalised -> d = 2 (1 - s) + theta (2 (|x| - |y|) / (|x| + |y|))^2  (faceclass.py:71)
    int raw;
    const float* row_nrm;
    const float* col_nrm;
    float theta;
    // BCE epilogue (train_classifier.py:60-84 fused behind the classifier, faceclass.py:23-27): logits alpha (threshold - d)
    // of the strict upper triangle, label = same group (row_cls == col_cls), weighted cross entropy and its derivatives
    float bce_alpha, bce_threshold, bce_pos_weight;
    double* bce_out;           // [4] sums over pairs: loss, dloss/dalpha, dloss/dthreshold, dloss/dtheta
    // ROWSTRIP epilogue, mining (fnb_mine_batched): running per-anchor arg-extrema folded into the Gram epilogue.  Every
    // region is one batch (rows == columns == the batch); mine_lab holds the label of every (global) row; each thread folds
    // its row's 64 columns into packed (distance, index) keys and merges them with one 64-bit atomic per row and tile:
    //   mine_pos_key[row] = max over same-label columns != row of (d << 32 | ~index)   -> hardest positive, ties -> lowest index
    //   mine_neg_key[row] = min over other-label columns of (d << 32 | index)          -> hardest negative
    // Column indices are local to the batch.  out == NULL: the distance strip is not materialised (hardest-only mining).
    const long long* mine_lab;
    unsigned long long* mine_pos_key;
    unsigned long long* mine_neg_key;
    // FILTER epilogue (false pairs at one threshold, the commented FalseExamples search of facenet/statistics.py:334-387): every
    // same-identity pair with d > threshold (a miss) and every different-identity pair with d < threshold (a false accept) is
    // appended to a compact list -- (original row, original row, distance) -- through one atomic counter; the N x N matrix is
    // never materialised.  Entries beyond filter_cap are counted but not stored.
    float filter_threshold;
    int* filter_rows;
    int* filter_cols;
    float* filter_dist;
    unsigned long long* filter_count;
    long long filter_cap;
    const long long* filter_perm;   // permuted row -> original row (class-sorted order of the launch)
    // Cluster-progress window (off unless sync_window > 0).  The clusters walk a static schedule, and over a long launch
    // they drift apart by many column panels: each then streams its own column panel AND evicts the row panels the
    // others still need (measured at 1M rows: 824 GB of DRAM reads per launch).  Every cluster publishes the column
    // panel it is entering and waits while it is more than sync_window panels ahead of the slowest one, so the clusters
    // share column panels inside L2 (1M rows: 193-208 GB, kernel 834 -> 800 ms, profiles/r01d_panel_window.md).
    // The wait is BOUNDED (kSyncSpinLimit polls, a few ms): a cluster that runs into the bound stops waiting for the rest
    // of the launch, so clusters that are not co-resident (SMs held by somebody else's kernel) cost time, never a hang.
    // Timing only: the integer bins do not depend on it.
    unsigned int* progress;    // [number of clusters], zeroed before the launch; 0xffffffff = finished
    int sync_window;
    // Accumulation-bias correction (fnb_bias.h).  The tensor core truncates (towards zero) when it aligns the products and
    // adds them into the fp32 accumulator, so a raw similarity comes out SMALLER in magnitude than the exact one by a
    // calibrated relative amount beta(|s|) (profiles/r02a_bias_vs_similarity.log: 1.5e-6 .. 3.3e-6 for fp16x3 at D = 512).
    // Wherever a similarity is evaluated exactly (checked HIST tiles, PAIRWISE / ROWSTRIP / BCE) it is first corrected,
    // s <- s (1 + beta(|s|)), beta piecewise linear over kBiasKnots knots in |s|; the arithmetic binning of interior tiles
    // folds the same correction into its linear map (host: CutTables::raw).  Two tables: [0] the launch's arithmetic,
    // [1] the fp16x3 arithmetic of strict tiles inside an fp16f8 launch.  NULL: no correction.
    const float* bias_beta;    // [2][kBiasStride]
    // fp16f8 launches: tiles that can hold a same-identity pair, and ragged / diagonal tiles, run the fp16x3 contraction
    // (hi*hi + hi*l16 + l16*hi, l16 = fp16 low part) instead of the e4m3 cross terms -- the pairs behind TP / FN get the
    // strict arithmetic, the e4m3 error model only has to hold for different-identity pairs.  1: whole-set launches (row_cls ==
    // col_cls over n_valid rows, strictness decided per global 512 x 512 block); 2: every tile (keyed launches).
    int strict_tiles;
    const unsigned int* strict_bits;   // strict_tiles == 1: bit (R / 512) * strict_nb + C / 512 of the block at (R, C)
    int strict_nb;
    // Tile queue (HIST launches; NULL: every cluster walks its static share, tile = cluster_id + i * num_clusters).  One 64-bit
    // counter, zeroed before the launch; a cluster takes its next super-tile with one atomicAdd (warp 3 of the cluster's first
    // CTA fetches a few tiles ahead and hands the index to every role of every CTA of the cluster through a small ring in
    // shared memory).  Tiles leave in schedule order -- row blocks fastest inside a column panel -- so at any time all clusters
    // work inside the same one or two column panels (no drift, no progress window), a cluster that starts late or runs on a
    // slower SM simply takes fewer tiles, and a SECOND launch of another cluster shape can drain the same queue (the 16 SMs a
    // grid of 4-CTA clusters cannot use run CTA pairs on it: fnb_gram.cu).  The integer bins do not depend on who took which tile.
    unsigned long long* tile_counter;
    // Sharded jobs whose ranks share their queues (fnb_comm_init mapped every rank's counters into every rank, CUDA IPC over
    // NVLink): steal_counter[v] is rank v's counter of THIS launch (steal_counter[steal_rank] == tile_counter).  All ranks walk
    // schedules of identical shape (every rank ceil(nrb / world) row blocks per region, rank v's j-th block = j * world + v; blocks
    // past the region's end are skipped), so a tile index taken from rank v's queue decodes with v in place of this rank.  A
    // cluster takes from its own queue while it has tiles -- local atomics, and its own row panels stay in its L2 -- then from the
    // other ranks' queues in cyclic order: whoever finishes early helps the GPUs that run slower under the power limit.
    unsigned long long* steal_counter[8];
    int steal_world, steal_rank;
};

constexpr int kStealShift = 27;                    // raw queue entry = tile index | owner rank << 27

constexpr int kSchedDepth = 4;                     // raw tile indices in flight per cluster (fetch warp -> decode warps)
constexpr int kTileRing = 8;                       // decoded tiles per CTA (the producer runs two tiles ahead of the epilogue, and every role looks one tile ahead)

constexpr int kBiasKnots  = 41;                    // |s| = 0, 0.025, ..., 1
constexpr int kBiasStride = 44;

// s (1 + beta(|s|)), beta linear between the knots of `tab` (shared memory)
__device__ __forceinline__ float bias_correct(const float* tab, float s) {
    const float a = fminf(fabsf(s) * (float)(kBiasKnots - 1), (float)(kBiasKnots - 1) - 0.001f);
    const int i = (int)a;
    const float f = a - (float)i;
    const float b = fmaf(f, tab[i + 1] - tab[i], tab[i]);
    return fmaf(s, b, s);
}

// faceclass.py:71 in float32, operation by operation (no contraction)
__device__ __forceinline__ float classifier_distance(float s, float nr, float nc, float theta) {
    const float g = __fdiv_rn(__fmul_rn(2.0f, __fsub_rn(nr, nc)), __fadd_rn(nr, nc));
    return __fadd_rn(__fmul_rn(2.0f, __fsub_rn(1.0f, s)), __fmul_rn(theta, __fmul_rn(g, g)));
}

// polls (about 1 us each: a 400 ns sleep plus one L2 round trip) before a cluster gives up the progress window; the waits of a
// healthy launch are tens of microseconds (one column panel of one cluster is ~30 us of work)
constexpr int kSyncSpinLimit = 4096;

struct TileInfo {
    int row0, col0, row_end, col_end, tri, key;
    int gcp;                   // launch-wide index of the tile's column panel (non-decreasing along a cluster's tile sequence)
    int cbeg;                  // first column of the tile's region (ROWSTRIP writes region-local column indices)
    int strictq;               // tile queue: the strict flag of the tile, decided by the CTA's decode warp (-1: static schedule)
};

constexpr int kTileWords = 12;                     // one decoded queue entry: the nine ints of TileInfo, padded to three 16-byte words

// One scheduler step hands a cluster a SUPER-TILE of (kPR * kTile) rows x (kPC * kTile) columns; the CTA pair at
// (pair_row, pair_col) of the cluster's kPR x kPC pair grid owns the tile at [row0 + pair_row * kTile, col0 + pair_col *
// kTile).  kPairs == 1: the super-tile is the tile; 2: 1 x 2 pairs (A shared); 4: 2 x 2 pairs (A and B shared).
template <int kPairs> struct Sched_PR { static constexpr int value = (kPairs == 4) ? 2 : 1; };
template <int kPairs> struct Sched_PC { static constexpr int value = (kPairs == 1) ? 1 : 2; };

// kSub > 1 (kPairs == 1 only): the scheduler step is kSub tiles wide and the pair works through them one after the other --
// a launch of plain CTA pairs walking the schedule of a launch of two-pair clusters (same super-tiles, same tile queue).
template <int kCtaGroup, int kPairs = 1, int kSub = 1>
struct TileScheduler {
    static constexpr int kTile = kRowsPerCta * kCtaGroup;
    static constexpr int kSuperRows = kTile * Sched_PR<kPairs>::value;
    static constexpr int kSuperCols = kTile * Sched_PC<kPairs>::value * kSub;
    const RegionDev* regions;
    long long pos, stride, total;
    int cur;
    ShardSpec shard;
    // position inside the current region, advanced incrementally (18 warps per CTA walk the schedule: no divisions per tile)
    RegionDev r;
    int cb, j;                 // column panel; index among this rank's row blocks of the region
    bool fresh;                // (cb, j) must be recomputed from pos (first call, or a new region)
    // tile queue (GramParams::tile_counter): the roles read DECODED tiles from a ring in shared memory that the CTA's decode
    // warp fills -- no division, no global load and no remote traffic on the roles' side
    const int* qtiles;         // [kTileRing][kTileWords]
    uint64_t* qfull;           // [kTileRing] entry written (one arrive: the decode warp)
    uint64_t* qempty;          // [kTileRing] entry read by every role of this CTA
    int qslot; uint32_t qphase;
    long long next_begin;      // locate(): first tile of the next region

    __device__ TileScheduler(const GramParams& p, int cluster_id, int num_clusters)
        : regions(p.regions), pos(cluster_id), stride(num_clusters), total(p.total_tiles), cur(0), shard(p.shard),
          cb(0), j(0), fresh(true), qtiles(nullptr), qfull(nullptr), qempty(nullptr), qslot(0), qphase(0), next_begin(0) {}

    __device__ void use_queue(const int* tiles_, uint64_t* full_, uint64_t* empty_) { qtiles = tiles_; qfull = full_; qempty = empty_; }

    __device__ void fill(TileInfo& t, int rb) const {
        t.row0 = r.row_begin + rb * kSuperRows;
        t.col0 = r.col_begin + cb * kSuperCols;
        t.row_end = r.row_end;
        t.col_end = r.col_end;
        t.tri = r.tri;
        t.key = r.key;
        t.gcp = r.cp_begin + cb;
        t.cbeg = r.col_begin;
        t.strictq = -1;
    }

    // decode warp: tile index (indices only grow) -> tile; false when the tile lies entirely on / below the diagonal
    __device__ bool locate(long long id, TileInfo& t, int owner = -1) {
        if (!fresh && id < r.tile_begin) { cur = 0; fresh = true; }     // another rank's queue (stealing): its position may lie behind ours
        if (fresh || id >= next_begin) {
            while (id >= regions[cur + 1].tile_begin) ++cur;
            r = regions[cur];                                   // one global load per REGION, not per tile
            next_begin = regions[cur + 1].tile_begin;
            fresh = false;
        }
        const int li = (int)(id - r.tile_begin);
        cb = li / r.own_cnt;
        j = li - cb * r.own_cnt;
        const int rb = (owner >= 0) ? j * shard.mod + owner
                     : (shard.width == 1 && shard.slots == nullptr) ? j * shard.mod + shard.lo : shard.block(j);
        if (rb >= r.nrb) return false;                          // padded share (ranks that share their queues): no such row block
        fill(t, rb);
        return !(r.tri && t.col0 + kSuperCols - 1 <= t.row0);
    }

    // every lane of the calling warp runs this (warp-uniform)
    __device__ bool next(TileInfo& t) {
        if (qtiles != nullptr) {
            mbar_wait(&qfull[qslot], qphase);
            const int4* e = reinterpret_cast<const int4*>(qtiles + qslot * kTileWords);
            const int4 a = e[0], b = e[1], c = e[2];
            __syncwarp();
            if ((threadIdx.x & 31) == 0) mbar_arrive(&qempty[qslot]);
            if (++qslot == kTileRing) { qslot = 0; qphase ^= 1u; }
            if (a.x < 0) return false;                          // the queue is empty
            t.row0 = a.x; t.col0 = a.y; t.row_end = a.z; t.col_end = a.w;
            t.tri = b.x; t.key = b.y; t.gcp = b.z; t.cbeg = b.w; t.strictq = c.x;
            return true;
        }
        while (pos < total) {
            if (fresh || pos >= regions[cur + 1].tile_begin) {
                while (pos >= regions[cur + 1].tile_begin) ++cur;
                r = regions[cur];
                const int li = (int)(pos - r.tile_begin);
                cb = li / r.own_cnt;
                j = li - cb * r.own_cnt;
                fresh = false;
            }
            const int rb = (shard.width == 1 && shard.slots == nullptr) ? j * shard.mod + shard.lo : shard.block(j);
            fill(t, rb);
            // advance by `stride` tiles inside the region (row blocks fastest)
            pos += stride;
            j += (int)stride;
            while (j >= r.own_cnt) { j -= r.own_cnt; ++cb; }
            if (rb >= r.nrb) continue;                                  // padded share (see locate)
            if (r.tri && t.col0 + kSuperCols - 1 <= t.row0) continue;   // entirely on/below the diagonal
            return true;
        }
        return false;
    }
};

struct __align__(16) GramSmemMisc {
    uint64_t full[kMaxSlots];
    uint64_t empty[kMaxSlots];
    uint64_t tfull[2];
    uint64_t tempty[2];
    uint32_t tmem_base;
    uint32_t pad[3];
    uint64_t ring_full[kSchedDepth];    // tile queue, raw indices: "written" (this CTA's copy, signalled by the cluster's fetch warp)
    uint64_t ring_empty[kSchedDepth];   // "read by the decode warp of every CTA" (the first CTA's copy is the one in use)
    uint64_t tile_full[kTileRing];      // decoded tiles of this CTA: "written" (its decode warp)
    uint64_t tile_empty[kTileRing];     // "read by every role of this CTA"
    int32_t ring[kSchedDepth];
    __align__(16) int32_t tiles[kTileRing][kTileWords];
    float cuts[kMaxBins];
    float wlo[kMaxBins + 4];
    float whi[kMaxBins + 4];
    uint32_t cta_all[kMaxBins + 4];     // CTA-level counters (all pairs / same-identity pairs) of the current key
    uint32_t cta_same[kMaxBins + 4];
    int32_t col_cls[2][256];
    float beta[2][kBiasStride];         // accumulation-bias tables (GramParams::bias_beta), zeros without correction
};

__device__ __forceinline__ int exact_bin(float s, const float* cuts_s) {
    // k = #{j : cuts[j] <= s}, cuts ascending, padded with +inf up to kMaxBins
    int k = 0;
#pragma unroll
    for (int step = kMaxBins / 2; step >= 1; step >>= 1) {
        if (cuts_s[k + step - 1] <= s) k += step;
    }
    return k;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// thread-private byte counters live in shared memory and are touched only through these (32-bit shared
// addresses, program order preserved among them by `volatile`)
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u8(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u8 [%0], %1;" :: "r"(addr), "r"(v));
}
__device__ __forceinline__ uint32_t mad_lo(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ float fma_sat(float a, float b, float c) {
    float d;
    asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

// ---------------------------------------------------------------------------------------
// the kernel

// kPairs == 4 (cta_group 2, HIST only): a cluster of 2 x 2 CTA pairs works on one 512 x 512 super-tile; every A box is shared
// by the two pairs of a pair-grid row and every B box by the two pairs of a pair-grid column, so each is read from L2 once
// and multicast (L2 -> SM reads drop to 1/2).  All four pairs run in lockstep.
// kPairs == 2 (cta_group 2, HIST only): a cluster of two CTA pairs works on one 256 x 512 super-tile; the pairs share
// the A rows, so every A box is read from L2 ONCE and multicast to the matching CTA of both pairs (each of the two
// CTAs issues one 64-row half of it).  L2 -> SM operand traffic drops to 3/4; the two pairs run in lockstep
// (a slot is refilled only after BOTH pairs' MMAs have consumed it).
// kSub == 2 (kPairs == 1, HIST only): a launch of plain CTA pairs that walks the 256 x 512 super-tiles of a two-pair-cluster
// launch, two tiles one after the other (GramParams::tile_counter).
template <int kCtaGroup, int kNumPass, bool kTf32, int kEpi, int kPairs = 1, int kSub = 1>
__global__ void __launch_bounds__(kGramThreads, 1)
gram_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
            const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
            const __grid_constant__ CUtensorMap tm_a_h8, const __grid_constant__ CUtensorMap tm_b_h8,
            const __grid_constant__ CUtensorMap tm_a_l16, const __grid_constant__ CUtensorMap tm_b_l16,
            const GramParams p)
{
    static_assert(kPairs == 1 || ((kPairs == 2 || kPairs == 4) && kCtaGroup == 2 && kEpi == EPI_HIST), "multicast clusters: CTA pairs, HIST only");
    static_assert(kSub == 1 || (kPairs == 1 && kEpi == EPI_HIST), "sub-tiles: plain pairs on the histogram schedule of a wider cluster");
    constexpr int kPR = Sched_PR<kPairs>::value;           // pair grid of the cluster: kPR x kPC pairs (1x1, 1x2, 2x2)
    constexpr int kPC = Sched_PC<kPairs>::value;
    constexpr int kTile   = kRowsPerCta * kCtaGroup;       // tile rows == tile cols
    constexpr int kUmmaN  = kTile;                         // accumulator columns per stage
    constexpr int kClusterCtas = kCtaGroup * kPairs;
    using Sched = TileScheduler<kCtaGroup, kPairs, kSub>;
    // kNumPass: 1 = one MMA per k-step; 3 = split operands x = hi + lo, hi*hi + hi*lo + lo*hi in the operand type;
    //           2 = hi*hi in fp16 plus the two cross terms in fp8 (e4m3) at twice the MMA rate (kSchemeF8)
    constexpr bool kF8    = (kNumPass == 2);
    constexpr int kParts  = (kNumPass == 3) ? 2 : 1;       // slots per k-block: {hi} or {hi, lo}
    constexpr int kColsPerWarp = kUmmaN / kColSplit;       // four epilogue warps per TMEM lane quadrant
    constexpr uint32_t kTmemCols = 2 * kUmmaN;             // double-buffered accumulator
    constexpr int kElemsPerBox = kTf32 ? 32 : 64;          // K elements per 128-byte row

    // 1024-byte alignment (SWIZZLE_128B atoms) by pointer arithmetic on the shared array itself, so the
    // compiler keeps every derived pointer in the shared address space (LDS/STS, not generic LD/ST)
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* slots = smem;
    GramSmemMisc* misc = reinterpret_cast<GramSmemMisc*>(smem + (size_t)p.num_slots * kSlotBytes);
    uint8_t* hist8 = reinterpret_cast<uint8_t*>(misc) + sizeof(GramSmemMisc);   // [nb8][kHist8Row] thread-private u8 counters

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t cluster_rank = (kCtaGroup == 2) ? cluster_ctarank() : 0u;
    const uint32_t cta_rank = cluster_rank & 1u;           // position inside the CTA pair: 0 = leader (issues the MMAs)
    const uint32_t pair_idx = cluster_rank >> 1;           // which pair of the cluster (0 when kPairs == 1)
    const uint32_t pair_row = pair_idx / kPC, pair_col = pair_idx % kPC;
    const uint32_t leader_rank = cluster_rank & ~1u;       // cluster rank of this pair's leader
    const bool is_leader = (cta_rank == 0);
    const int cluster_id = blockIdx.x / kClusterCtas;
    const int num_clusters = gridDim.x / kClusterCtas;
    const uint16_t pair_mask = (uint16_t)(3u << (pair_idx * 2));            // the two CTAs of this pair
    const uint16_t cluster_mask = (uint16_t)((1u << kClusterCtas) - 1u);    // every CTA of the cluster
    const int num_slots = p.num_slots;
    // fp16 vs bf16 is a runtime choice (same instruction kind, one descriptor field)
    const uint32_t idesc = make_idesc((uint32_t)p.operand_fmt, kTile, kUmmaN);
    const uint32_t idesc_f8 = make_idesc(kFmtE4M3, kTile, kUmmaN);

    // ---- one-time setup
    if (threadIdx.x == 0) {
        for (int i = 0; i < num_slots; ++i) { mbar_init(&misc->full[i], 1); mbar_init(&misc->empty[i], kPairs); }
        for (int i = 0; i < 2; ++i) { mbar_init(&misc->tfull[i], 1); mbar_init(&misc->tempty[i], kEpiWarps * kCtaGroup); }
        // tile queue: an entry is consumed by the producer warp and the epilogue warps of every CTA and by the MMA warp of every leader
        // tile queue: a raw index is read by the decode warp of every CTA; a decoded tile by the producer warp, the epilogue warps
        // and (leader CTAs) the MMA warp of this CTA
        for (int i = 0; i < kSchedDepth; ++i) { mbar_init(&misc->ring_full[i], 1); mbar_init(&misc->ring_empty[i], kClusterCtas); }
        for (int i = 0; i < kTileRing; ++i) { mbar_init(&misc->tile_full[i], 1); mbar_init(&misc->tile_empty[i], 1 + kEpiWarps + (is_leader ? 1 : 0)); }
        fence_mbar_init();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_a_hi); tma_prefetch_desc(&tm_b_hi);
        if (kNumPass != 1) { tma_prefetch_desc(&tm_a_lo); tma_prefetch_desc(&tm_b_lo); }
        if (kF8) { tma_prefetch_desc(&tm_a_h8); tma_prefetch_desc(&tm_b_h8); }
        if (kF8 && p.strict_tiles) { tma_prefetch_desc(&tm_a_l16); tma_prefetch_desc(&tm_b_l16); }
    }
    if (warp == 2) tmem_alloc<kCtaGroup>(&misc->tmem_base, kTmemCols);
    if (warp >= kFirstEpiWarp) {
        const int te = threadIdx.x - kFirstEpiWarp * 32;
        for (int i = te; i < 2 * kBiasStride; i += kEpiThreads) (&misc->beta[0][0])[i] = p.bias_beta ? p.bias_beta[i] : 0.f;
    }
    if (kEpi == EPI_HIST && warp >= kFirstEpiWarp) {
        const int te = threadIdx.x - kFirstEpiWarp * 32;
        for (int i = te; i < kMaxBins; i += kEpiThreads) misc->cuts[i] = p.cuts[i];
        for (int i = te; i < kMaxBins + 1; i += kEpiThreads) { misc->wlo[i] = p.wlo[i]; misc->whi[i] = p.whi[i]; }
        for (int i = te; i < kMaxBins + 4; i += kEpiThreads) { misc->cta_all[i] = 0; misc->cta_same[i] = 0; }
        uint32_t* h32 = reinterpret_cast<uint32_t*>(hist8);
        for (int i = te; i < p.nb8 * (kHist8Row / 4); i += kEpiThreads) h32[i] = 0;
    }
    __syncwarp();
    tc_fence_before();
    if (kCtaGroup == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = misc->tmem_base;

    // fp16f8 launches: does the cluster's super-tile run the strict fp16x3 contraction (GramParams::strict_tiles)?  Identical
    // in every role and every CTA of the cluster.  The rule is a function of the GLOBAL 512 x 512 block of the pair matrix the
    // super-tile lies in -- strict iff any tile of that block is ragged, touches the diagonal or can hold a same-identity pair
    // (class ranks of its corner rows) -- so it does not depend on the cluster shape, the super-row height or the rank split:
    // the integer bins of an fp16f8 launch stay identical across all of them (regions are 512-aligned, fnb_api.cu).
    auto strict_tile = [&](const TileInfo& t) -> bool {
        if (!kF8 || !p.strict_tiles) return false;
        if (t.strictq >= 0) return t.strictq != 0;               // decided by the decode warp (tile queue)
        if (p.strict_tiles == 2) return true;
        const unsigned int idx = (unsigned int)(t.row0 >> 9) * (unsigned int)p.strict_nb + (unsigned int)(t.col0 >> 9);
        return (__ldg(p.strict_bits + (idx >> 5)) >> (idx & 31u)) & 1u;          // one bit per 512 x 512 block (strict_blocks_kernel)
    };

    const bool use_queue = (kEpi == EPI_HIST) && (p.tile_counter != nullptr);
    // kSub > 1: sub-tile `sub` of the super-tile lies outside the region or on / below the diagonal (same test in every role)
    auto sub_null = [&](const TileInfo& t, int sub) -> bool {
        if (kSub == 1) return false;
        const int c0 = t.col0 + sub * kTile;
        return c0 >= t.col_end || (t.tri && c0 + kTile - 1 <= t.row0);
    };

    // =====================================================================================
    if (warp == 3) {
        // ------------------------------ tile queue: fetch + decode warp ------------------
        // First CTA of the cluster: one atomicAdd per super-tile, the raw index written into the ring of EVERY CTA of the cluster
        // (st.shared::cluster) and announced on their ring_full barriers (release.cluster); -1 ends the launch.  Every CTA: the
        // raw index is decoded here (region search, one division, the strict flag's global load) and published as a complete
        // TileInfo in the CTA's own ring, so the roles pay one local barrier wait and three shared-memory loads per tile.
        if (use_queue) {
            Sched loc(p, 0, 1);
            int rslot = 0; uint32_t rphase = 0;
            int qslot = 0; uint32_t qphase = 0;
            int victim = p.steal_rank, victim_tried = 0;             // (lane 0 of the fetch warp)
            for (;;) {
                if (cluster_rank == 0) {
                    mbar_wait_relaxed(&misc->ring_empty[rslot], rphase ^ 1u);
                    int got_id = 0;
                    if (lane == 0) {
                        got_id = -1;
                        if (p.steal_world > 1) {
                            // own queue first, then the other ranks' (system-scope atomics: those counters live on peer GPUs)
                            while (victim_tried < p.steal_world) {
                                const unsigned long long got = atomicAdd_system(p.steal_counter[victim], 1ull);
                                if (got < (unsigned long long)p.total_tiles) { got_id = (int)got | (victim << kStealShift); break; }
                                ++victim_tried;
                                victim = (victim + 1 == p.steal_world) ? 0 : victim + 1;
                            }
                        } else {
                            const unsigned long long got = atomicAdd(p.tile_counter, 1ull);
                            if (got < (unsigned long long)p.total_tiles) got_id = (int)got;
                        }
                    }
                    got_id = __shfl_sync(0xffffffffu, got_id, 0);
                    if (lane < kClusterCtas) {
                        st_cluster_s32(mapa_u32(smem_u32(&misc->ring[rslot]), (uint32_t)lane), got_id);
                        mbar_arrive_cluster_addr(mapa_u32(smem_u32(&misc->ring_full[rslot]), (uint32_t)lane));
                    }
                    __syncwarp();
                }
                // CTA-scope acquire: the index arrives by st.shared::cluster + an arrive with release.cluster and is read from
                // this CTA's own shared memory (no cache in between); a cluster-scope acquire would make ptxas invalidate the L1
                mbar_wait(&misc->ring_full[rslot], rphase);
                const int id = *reinterpret_cast<const volatile int*>(&misc->ring[rslot]);
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster_addr(mapa_u32(smem_u32(&misc->ring_empty[rslot]), 0u));
                if (++rslot == kSchedDepth) { rslot = 0; rphase ^= 1u; }
                TileInfo t;
                t.row0 = -1; t.col0 = 0; t.row_end = 0; t.col_end = 0; t.tri = 0; t.key = 0; t.gcp = 0; t.cbeg = 0; t.strictq = 0;
                if (id >= 0) {
                    const int owner = (p.steal_world > 1) ? (id >> kStealShift) : -1;
                    if (!loc.locate(id & ((1 << kStealShift) - 1), t, owner)) continue;   // below the diagonal / padding: nothing to hand out
                    t.strictq = -1;
                    t.strictq = strict_tile(t) ? 1 : 0;
                }
                mbar_wait_relaxed(&misc->tile_empty[qslot], qphase ^ 1u);
                if (lane == 0) {
                    int4* e = reinterpret_cast<int4*>(&misc->tiles[qslot][0]);
                    e[0] = make_int4(t.row0, t.col0, t.row_end, t.col_end);
                    e[1] = make_int4(t.tri, t.key, t.gcp, t.cbeg);
                    e[2] = make_int4(t.strictq, 0, 0, 0);
                    mbar_arrive(&misc->tile_full[qslot]);        // release (CTA scope): the stores above are visible to the waiters
                }
                __syncwarp();
                if (++qslot == kTileRing) { qslot = 0; qphase ^= 1u; }
                if (id < 0) break;
            }
        }
    } else if (warp == 0) {
        // ------------------------------ TMA producer ------------------------------------
        // The whole warp walks the schedule and polls the barriers (warp-uniform control flow, so the compiler keeps
        // addresses and descriptors in uniform registers); one elected lane issues the copies.
        Sched sched(p, cluster_id, num_clusters);
        if (use_queue) sched.use_queue(&misc->tiles[0][0], misc->tile_full, misc->tile_empty);
        TileInfo t;
        int slot = 0; uint32_t phase = 0;
        int ntile = -1;
        int last_gcp = -1;
        const bool sync_on = p.sync_window > 0 && cluster_rank == 0 && !use_queue;
        bool sync_wait = true;        // cleared for good when a wait runs into kSyncSpinLimit
        // the strict flag of tile i + 1 is fetched while tile i is processed: its global-load latency never sits in front of a tile
        TileInfo t_next;
        bool more = sched.next(t_next);
        bool strict_next = more && strict_tile(t_next);
        while (more) {
            t = t_next;
            const bool strict_now = strict_next;
            more = sched.next(t_next);
            strict_next = more && strict_tile(t_next);
            ++ntile;
            if (sync_on && t.gcp != last_gcp) {
                // entering a new column panel: publish it, then wait for the stragglers (the slowest cluster never waits)
                last_gcp = t.gcp;
                volatile unsigned int* prog = p.progress;
                if (lane == 0) prog[cluster_id] = (unsigned int)t.gcp + 1u;
                __syncwarp();
                int spins = 0;
                while (sync_wait) {
                    unsigned int mn = 0xffffffffu;
                    for (int i = lane; i < num_clusters; i += 32) mn = min(mn, prog[i]);
#pragma unroll
                    for (int o = 16; o; o >>= 1) mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
                    if (mn + (unsigned int)p.sync_window >= (unsigned int)t.gcp + 1u) break;
                    if (++spins >= kSyncSpinLimit) sync_wait = false;      // warp-uniform (mn is): give up waiting for good
                    __nanosleep(400);
                }
            }
            for (int sub = 0; sub < kSub; ++sub) {
            if (sub_null(t, sub)) continue;
            const int arow = t.row0 + (int)pair_row * kTile + (int)cta_rank * kRowsPerCta;
            const int brow = t.col0 + ((int)pair_col * kSub + sub) * kTile + (int)cta_rank * kRowsPerCta;
            auto load_slot = [&](const CUtensorMap* ma, const CUtensorMap* mb, int kcol) {
                mbar_wait_relaxed(&misc->empty[slot], phase ^ 1u);
                if (elect_one()) {
                    uint8_t* dst = slots + (size_t)slot * kSlotBytes;
                    if ((p.debug & 2) && ntile > 0) {
                        // profiling: only the first tile is really loaded, so that later MMAs run on representative data
                        if (is_leader) mbar_arrive(&misc->full[slot]);
                    } else if (kCtaGroup == 1) {
                        mbar_arrive_expect_tx(&misc->full[slot], kSlotBytes);
                        tma_load_2d(dst, ma, &misc->full[slot], kcol, arow);
                        tma_load_2d(dst + kBoxBytes, mb, &misc->full[slot], kcol, brow);
                    } else {
                        if (is_leader) mbar_arrive_expect_tx(&misc->full[slot], 2 * kSlotBytes);
                        const uint32_t bar = mapa_u32(smem_u32(&misc->full[slot]), leader_rank);
                        // An operand box that two pairs of the cluster need is read from L2 once: the maps carry 64-row
                        // boxes, each of the two CTAs that want the box (same position in their pair) fetches one half and
                        // multicasts it to both; the other half arrives from the partner.
                        if constexpr (kPC == 1) {
                            tma_load_2d_pair(dst, ma, bar, kcol, arow);
                        } else {
                            // A rows are shared along the pair-grid row: partners (pair_row, 0) and (pair_row, 1)
                            const uint16_t mc = (uint16_t)((1u << ((pair_row * kPC + 0) * 2 + cta_rank)) | (1u << ((pair_row * kPC + 1) * 2 + cta_rank)));
                            tma_load_2d_pair_mc(dst + pair_col * (kBoxBytes / 2), ma, bar, mc, kcol, arow + (int)pair_col * (kRowsPerCta / 2));
                        }
                        if constexpr (kPR == 1) {
                            tma_load_2d_pair(dst + kBoxBytes, mb, bar, kcol, brow);
                        } else {
                            // B rows (tile columns) are shared along the pair-grid column: partners (0, pair_col) and (1, pair_col)
                            const uint16_t mc = (uint16_t)((1u << ((0 * kPC + pair_col) * 2 + cta_rank)) | (1u << ((1 * kPC + pair_col) * 2 + cta_rank)));
                            tma_load_2d_pair_mc(dst + kBoxBytes + pair_row * (kBoxBytes / 2), mb, bar, mc, kcol, brow + (int)pair_row * (kRowsPerCta / 2));
                        }
                    }
                }
                __syncwarp();
                if (++slot == num_slots) { slot = 0; phase ^= 1u; }
            };
            if (kF8 && strict_now) {
                // strict tile of an fp16f8 launch: the fp16x3 slots {hi, hi}, {l16, l16} per k-block (same slot count)
                for (int kb = 0; kb < p.kblocks; ++kb) {
                    load_slot(&tm_a_hi, &tm_b_hi, kb * 64);
                    load_slot(&tm_a_l16, &tm_b_l16, kb * 64);
                }
            } else if constexpr (kF8) {
                // The small cross terms first -- {A e4m3(x), B e4m3(lo)} and {A e4m3(lo), B e4m3(x)} per 128 K-elements --
                // then the fp16 {hi, hi} slots.  The tensor core truncates when it adds into the fp32 accumulator, an error
                // proportional to the accumulator's magnitude per step; while the cross terms are summed the accumulator
                // is ~2^-11 of its final value, so only the hi*hi steps contribute.
                for (int ks = 0; ks < p.kblocks / 2; ++ks) {
                    load_slot(&tm_a_h8, &tm_b_lo, ks * 128);
                    load_slot(&tm_a_lo, &tm_b_h8, ks * 128);
                }
                for (int kb = 0; kb < p.kblocks; ++kb) load_slot(&tm_a_hi, &tm_b_hi, kb * 64);
            } else {
                for (int kb = 0; kb < p.kblocks; ++kb) {
                    load_slot(&tm_a_hi, &tm_b_hi, kb * kElemsPerBox);
                    if (kParts == 2) load_slot(&tm_a_lo, &tm_b_lo, kb * kElemsPerBox);
                }
            }
            }   // sub
        }
        if (sync_on && lane == 0) {
            volatile unsigned int* prog = p.progress;
            prog[cluster_id] = 0xffffffffu;              // finished: nobody waits for this cluster any more
        }
    } else if (warp == 1) {
        // ------------------------------ MMA issuer --------------------------------------
        // Leader CTA of the pair only.  Warp-uniform control flow; one elected lane issues every tcgen05.mma and
        // tcgen05.commit (the commits track the MMAs of the issuing thread: elect.sync picks the same lane each time).
        if (is_leader) {
            Sched sched(p, cluster_id, num_clusters);
            if (use_queue) sched.use_queue(&misc->tiles[0][0], misc->tile_full, misc->tile_empty);
            TileInfo t;
            int slot = 0; uint32_t phase = 0;
            uint32_t it = 0;
            TileInfo t_next;
            bool more = sched.next(t_next);
            bool strict_next = more && strict_tile(t_next);
            while (more) {
                t = t_next;
                const bool strict_now = strict_next;
                more = sched.next(t_next);
                strict_next = more && strict_tile(t_next);      // consumed one tile later: the load overlaps this tile's MMAs
                for (int sub = 0; sub < kSub; ++sub) {
                if (sub_null(t, sub)) continue;
                const uint32_t acc = it & 1u, acc_phase = (it >> 1) & 1u;
                mbar_wait(&misc->tempty[acc], acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * kUmmaN;
                // one k-block pass of the split contraction hi*hi + hi*lo + lo*hi over two slots {hi, hi}, {lo, lo}
                auto x3_tile = [&]() {
                    for (int kb = 0; kb < p.kblocks; ++kb) {
                        const int slot_hi = slot;
                        mbar_wait(&misc->full[slot_hi], phase);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(slots + (size_t)slot_hi * kSlotBytes);
                        const uint64_t a_hi = make_smem_desc(sa), b_hi = make_smem_desc(sa + kBoxBytes);
                        if (elect_one()) {
#pragma unroll
                            for (int k4 = 0; k4 < 4; ++k4)
                                umma<kCtaGroup, kTf32>(d_tmem, a_hi + 2 * k4, b_hi + 2 * k4, idesc, (kb | k4) != 0 ? 1u : 0u);
                        }
                        __syncwarp();
                        if (++slot == num_slots) { slot = 0; phase ^= 1u; }
                        const int slot_lo = slot;
                        mbar_wait(&misc->full[slot_lo], phase);
                        tc_fence_after();
                        if (elect_one()) {
                            const uint32_t sl = smem_u32(slots + (size_t)slot_lo * kSlotBytes);
                            const uint64_t a_lo = make_smem_desc(sl), b_lo = make_smem_desc(sl + kBoxBytes);
#pragma unroll
                            for (int k4 = 0; k4 < 4; ++k4)
                                umma<kCtaGroup, kTf32>(d_tmem, a_hi + 2 * k4, b_lo + 2 * k4, idesc, 1u);
#pragma unroll
                            for (int k4 = 0; k4 < 4; ++k4)
                                umma<kCtaGroup, kTf32>(d_tmem, a_lo + 2 * k4, b_hi + 2 * k4, idesc, 1u);
                            umma_commit<kCtaGroup>(&misc->empty[slot_hi], cluster_mask);
                            umma_commit<kCtaGroup>(&misc->empty[slot_lo], cluster_mask);
                        }
                        __syncwarp();
                        if (++slot == num_slots) { slot = 0; phase ^= 1u; }
                    }
                };
                if (kF8 && strict_now) {
                    x3_tile();
                } else if constexpr (kF8) {
                    // kblocks e4m3 slots (K = 128 each: the cross terms), then kblocks fp16 slots (K = 64 each: hi*hi)
                    for (int st = 0; st < p.kblocks; ++st) {
                        mbar_wait(&misc->full[slot], phase);
                        tc_fence_after();
                        if (elect_one()) {
                            const uint32_t sa = smem_u32(slots + (size_t)slot * kSlotBytes);
                            const uint64_t da = make_smem_desc(sa), db = make_smem_desc(sa + kBoxBytes);
#pragma unroll
                            for (int k4 = 0; k4 < 4; ++k4)
                                umma_f8<kCtaGroup>(d_tmem, da + 2 * k4, db + 2 * k4, idesc_f8, (st | k4) != 0 ? 1u : 0u);
                            umma_commit<kCtaGroup>(&misc->empty[slot], cluster_mask);
                        }
                        __syncwarp();
                        if (++slot == num_slots) { slot = 0; phase ^= 1u; }
                    }
                    for (int kb = 0; kb < p.kblocks; ++kb) {
                        mbar_wait(&misc->full[slot], phase);
                        tc_fence_after();
                        if (elect_one()) {
                            const uint32_t sa = smem_u32(slots + (size_t)slot * kSlotBytes);
                            const uint64_t da = make_smem_desc(sa), db = make_smem_desc(sa + kBoxBytes);
#pragma unroll
                            for (int k4 = 0; k4 < 4; ++k4)
                                umma<kCtaGroup, false>(d_tmem, da + 2 * k4, db + 2 * k4, idesc, 1u);
                            umma_commit<kCtaGroup>(&misc->empty[slot], cluster_mask);
                        }
                        __syncwarp();
                        if (++slot == num_slots) { slot = 0; phase ^= 1u; }
                    }
                } else if constexpr (kNumPass == 3) {
                    x3_tile();
                } else {
                    for (int kb = 0; kb < p.kblocks; ++kb) {
                        mbar_wait(&misc->full[slot], phase);
                        tc_fence_after();
                        if (elect_one()) {
                            const uint32_t sa = smem_u32(slots + (size_t)slot * kSlotBytes);
                            const uint64_t a_hi = make_smem_desc(sa), b_hi = make_smem_desc(sa + kBoxBytes);
#pragma unroll
                            for (int k4 = 0; k4 < 4; ++k4)
                                umma<kCtaGroup, kTf32>(d_tmem, a_hi + 2 * k4, b_hi + 2 * k4, idesc, (kb | k4) != 0 ? 1u : 0u);
                            umma_commit<kCtaGroup>(&misc->empty[slot], cluster_mask);
                        }
                        __syncwarp();
                        if (++slot == num_slots) { slot = 0; phase ^= 1u; }
                    }
                }
                if (elect_one()) umma_commit<kCtaGroup>(&misc->tfull[acc], pair_mask);
                __syncwarp();
                ++it;
                }   // sub
            }
        }
    } else if (warp >= kFirstEpiWarp) {
        // ------------------------------ epilogue ----------------------------------------
        const int e = warp - kFirstEpiWarp;          // 0..15
        const int q = e & 3;                         // TMEM lane quadrant (== warp % 4)
        const int colq = e >> 2;                     // column quarter
        const int te = e * 32 + lane;                // epilogue thread id 0..511
        const int row_in_tile = (int)cta_rank * kRowsPerCta + q * 32 + lane;
        const uint32_t tmem_lane = tmem_base + ((uint32_t)(q * 32) << 16);
        const float scale = p.acc_scale;

        Sched sched(p, cluster_id, num_clusters);
        if (use_queue) sched.use_queue(&misc->tiles[0][0], misc->tile_full, misc->tile_empty);
        TileInfo tq;
        uint32_t it = 0;

        float smin = INFINITY, smax = -INFINITY;
        double bce_acc[4] = {0.0, 0.0, 0.0, 0.0};
        uint32_t eps_cnt = 0, tiles_done = 0;
        int cur_key = -1;
        uint32_t fast_since_flush = 0, tiles_since_global = 0;
        // u8 counters hold <= 255: one fast tile adds at most kColsPerWarp (+1 for the pair rule) per counter
        constexpr uint32_t kFlushEvery = 254u / (uint32_t)kColsPerWarp;
        const bool all_slow = p.force_slow || !p.uniform || p.row_nrm != nullptr ||
                              (p.norm_max_ord != nullptr && __ldg(p.norm_max_ord) > p.norm_limit_ord);

        // thread-private byte counters: bin b of thread (colq, q, lane) lives at byte b * 512 + (colq * 32 + lane) * 4 + q,
        // i.e. a warp always touches 32 different banks.  hb folds in the exponent bits of the magic-number trick.
        const uint32_t fbits = p.frac_bits;
        const uint32_t hb = smem_u32(hist8) + (uint32_t)((colq * 32 + lane) * 4 + q) - ((0x4B400000u >> fbits) * (uint32_t)kHist8Row);

        // fold the u8 counters into the CTA-level u32 counters (every epilogue thread takes part)
        auto flush_u8 = [&]() {
            named_bar_sync(1, kEpiThreads);
            for (int b = e; b < p.nb8; b += kEpiWarps) {
                uint4* w = reinterpret_cast<uint4*>(hist8 + (size_t)b * kHist8Row) + lane;
                const uint4 v = *w;
                *w = make_uint4(0u, 0u, 0u, 0u);
                uint32_t sum = __dp4a(v.x, 0x01010101u, 0u);
                sum = __dp4a(v.y, 0x01010101u, sum);
                sum = __dp4a(v.z, 0x01010101u, sum);
                sum = __dp4a(v.w, 0x01010101u, sum);
                sum = __reduce_add_sync(0xffffffffu, sum);
                if (lane == 0 && sum) atomicAdd(&misc->cta_all[min(b, p.T_fin)], sum);
            }
            named_bar_sync(1, kEpiThreads);
            fast_since_flush = 0;
        };
        // CTA-level counters -> global 64-bit bins of `key`
        auto flush_cta = [&](int key) {
            if (fast_since_flush) flush_u8(); else named_bar_sync(1, kEpiThreads);
            if (key >= 0) {
                unsigned long long* dst_all = p.bins + ((size_t)key * 2 + 0) * p.bins_stride;
                unsigned long long* dst_same = p.bins + ((size_t)key * 2 + 1) * p.bins_stride;
                for (int b = te; b <= p.T; b += kEpiThreads) {
                    const uint32_t a = misc->cta_all[b], sm = misc->cta_same[b];
                    if (a) { atomicAdd(dst_all + b, (unsigned long long)a); misc->cta_all[b] = 0; }
                    if (sm) { atomicAdd(dst_same + b, (unsigned long long)sm); misc->cta_same[b] = 0; }
                }
            }
            named_bar_sync(1, kEpiThreads);
            tiles_since_global = 0;
        };

        TileInfo t_next;
        bool more = sched.next(t_next);
        bool strict_next = more && strict_tile(t_next);
        while (more) {
            tq = t_next;
            const bool strict = strict_next;
            more = sched.next(t_next);
            strict_next = more && strict_tile(t_next);
            for (int sub = 0; sub < kSub; ++sub) {
            if (sub_null(tq, sub)) continue;
            TileInfo t = tq;
            if constexpr (kSub > 1) t.col0 += sub * kTile;
            const uint32_t acc = it & 1u, acc_phase = (it >> 1) & 1u;
            if constexpr (kPairs > 1) { t.row0 += (int)pair_row * kTile; t.col0 += (int)pair_col * kTile; }   // this pair's tile of the super-tile
            // a pair whose tile lies outside the region / below the diagonal still runs the pipeline (lockstep) and drops the result
            const bool null_tile = (kPairs > 1) && (t.col0 >= t.col_end || t.row0 >= t.row_end || (t.tri && t.col0 + kTile - 1 <= t.row0));
            const int row = t.row0 + row_in_tile;
            const int colw = t.col0 + colq * kColsPerWarp;       // first column of this warp
            const uint32_t taddr0 = tmem_lane + acc * kUmmaN + colq * kColsPerWarp;

            if constexpr (kEpi == EPI_HIST) {
                if (t.key != cur_key || tiles_since_global >= 65536u) {
                    if (cur_key >= 0) flush_cta(cur_key);
                    cur_key = t.key;
                }
                ++tiles_since_global;
                // classify the tile (identical in every epilogue thread of the CTA pair)
                const bool edge = (t.row0 + kTile > t.row_end) || (t.col0 + kTile > t.col_end) ||
                                  (t.tri && t.col0 <= t.row0 + kTile - 1);
                const int rlast = min(t.row0 + kTile, t.row_end) - 1;
                const int clast = min(t.col0 + kTile, t.col_end) - 1;
                // (a pair tile outside the region is dropped below: its corner rows may lie past the class array)
                const bool lab = !null_tile && (__ldg(p.row_cls + t.row0) <= __ldg(p.col_cls + clast)) &&
                                 (__ldg(p.col_cls + t.col0) <= __ldg(p.row_cls + rlast));
                // a strict tile without ragged edges or same-identity pairs still bins arithmetically, with the map fitted to the
                // fp16x3 arithmetic it was computed in
                const bool slow = all_slow || edge || lab || (strict && !p.noclip);
                const float* beta = misc->beta[strict ? 1 : 0];

                if ((p.debug & 1) || null_tile) {
                    mbar_wait_relaxed(&misc->tfull[acc], acc_phase);
                    tc_fence_after();
                } else if (!slow) {
                    // ---------------- interior tile: every element valid, no same-identity pair -------------
                    if (fast_since_flush >= kFlushEvery) flush_u8();
                    ++fast_since_flush;
                    mbar_wait_relaxed(&misc->tfull[acc], acc_phase);
                    tc_fence_after();
                    const uint32_t nmask = p.near_mask;
                    // two elements per step: both counters are read before either is written, so equal addresses add 2
                    auto bump2 = [&](uint32_t k0, uint32_t k1) {     // k = bin index (+ the constant folded into hb)
                        const uint32_t a0 = mad_lo(k0, (uint32_t)kHist8Row, hb), a1 = mad_lo(k1, (uint32_t)kHist8Row, hb);
                        const uint32_t c0 = lds_u8(a0);
                        const uint32_t c1 = lds_u8(a1);
                        const uint32_t inc = (k0 == k1) ? 2u : 1u;
                        sts_u8(a0, c0 + inc);
                        sts_u8(a1, c1 + inc);
                    };
                    if (p.noclip) {
                        // the cuts span the whole similarity range (guard bins absorb |s| <= 1 + atol + mode error):
                        // one FMA takes the accumulator straight to the fixed-point bin word
                        const float g1 = strict ? p.f_g1x : p.f_g1, g0 = strict ? p.f_g0x : p.f_g0;
                        const uint32_t half = p.near_half;
#pragma unroll 1
                        for (int c = 0; c < kColsPerWarp / 32; ++c) {
                            uint32_t r[32];
                            tmem_ld32(taddr0 + c * 32, r);
                            tmem_ld_wait();
#pragma unroll
                            for (int j = 0; j < 32; j += 2) {
                                const uint32_t w0 = __float_as_uint(fmaf(__uint_as_float(r[j]), g1, g0));
                                const uint32_t w1 = __float_as_uint(fmaf(__uint_as_float(r[j + 1]), g1, g0));
                                eps_cnt += (((w0 + half) & nmask) == 0u) ? 1u : 0u;
                                eps_cnt += (((w1 + half) & nmask) == 0u) ? 1u : 0u;
                                bump2(w0 >> fbits, w1 >> fbits);
                            }
                        }
                    } else {
                        const float s1 = p.f_s1, b1 = p.f_b1, k2 = p.f_k2, mk = p.f_magic_k, mn = p.f_magic_n;
#pragma unroll 1
                        for (int c = 0; c < kColsPerWarp / 32; ++c) {
                            uint32_t r[32];
                            tmem_ld32(taddr0 + c * 32, r);
                            tmem_ld_wait();
#pragma unroll
                            for (int j = 0; j < 32; j += 2) {
                                const float v0 = fma_sat(__uint_as_float(r[j]), s1, b1);
                                const float v1 = fma_sat(__uint_as_float(r[j + 1]), s1, b1);
                                eps_cnt += ((__float_as_uint(fmaf(v0, k2, mn)) & nmask) == 0u) ? 1u : 0u;
                                eps_cnt += ((__float_as_uint(fmaf(v1, k2, mn)) & nmask) == 0u) ? 1u : 0u;
                                bump2(__float_as_uint(fmaf(v0, k2, mk)) >> fbits, __float_as_uint(fmaf(v1, k2, mk)) >> fbits);
                            }
                        }
                    }
                } else {
                    // ---------------- checked tile: edges, diagonal, same-identity pairs, general cuts --------
                    mbar_wait_relaxed(&misc->tfull[acc], acc_phase);
                    tc_fence_after();
                    // stage this tile's column classes (kTile <= 256 columns).  Safe after the tfull wait:
                    // every epilogue warp has released the tile that last used col_cls[acc].
                    if (te < kTile) {
                        const int c = t.col0 + te;
                        misc->col_cls[acc][te] = (c < t.col_end) ? __ldg(p.col_cls + c) : -2;
                    }
                    const int my_cls = (row < t.row_end) ? __ldg(p.row_cls + row) : -1;
                    const float my_nrm = (p.row_nrm != nullptr && row < t.row_end) ? __ldg(p.row_nrm + row) : 1.0f;
                    named_bar_sync(2, kEpiThreads);
                    const bool row_ok = row < t.row_end;
#pragma unroll 1
                    for (int c = 0; c < kColsPerWarp; c += 4) {
                        uint32_t r[4];
                        tmem_ld4(taddr0 + c, r);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int col = colw + c + j;
                            const bool on_diag = (t.tri == 2) && (col == row);
                            const bool ok = row_ok && (col < t.col_end) && (!t.tri || col > row || on_diag);
                            if (ok) {
                                float s = __uint_as_float(r[j]) * scale;
                                if (!on_diag) { smin = fminf(smin, s); smax = fmaxf(smax, s); }
                                s = bias_correct(beta, s);
                                if (p.row_nrm != nullptr)      // classifier distance with the norm term, back on the similarity axis
                                    s = __fsub_rn(1.0f, __fmul_rn(0.5f, classifier_distance(s, my_nrm, __ldg(p.col_nrm + col), p.theta)));
                                else if (!p.raw) s = fminf(fmaxf(s, -1.0f), 1.0f);
                                const int ke = exact_bin(s, misc->cuts);
                                if (on_diag) {
                                    // self pairs of the full n_i x n_i blocks of train_classifier.py:35: slot key + 1, N adds in total
                                    atomicAdd(p.bins + ((size_t)(t.key + 1) * 2) * p.bins_stride + ke, 1ull);
                                } else {
                                    if (s <= misc->whi[ke] || s >= misc->wlo[ke]) ++eps_cnt;
                                    atomicAdd(&misc->cta_all[ke], 1u);
                                    if (misc->col_cls[acc][col - t.col0] == my_cls) atomicAdd(&misc->cta_same[ke], 1u);
                                }
                            }
                        }
                    }
                }
            } else if constexpr (kEpi == EPI_FILTER) {
                // ------------------- false pairs at one threshold: compact candidate list ----------
                mbar_wait_relaxed(&misc->tfull[acc], acc_phase);
                tc_fence_after();
                const bool row_ok = row < t.row_end;
                const int my_cls = row_ok ? __ldg(p.row_cls + row) : -1;
                const float thr = p.filter_threshold;
#pragma unroll 1
                for (int c = 0; c < kColsPerWarp / 32; ++c) {
                    uint32_t r[32];
                    tmem_ld32(taddr0 + c * 32, r);
                    tmem_ld_wait();
                    if (row_ok) {
#pragma unroll 4
                        for (int j = 0; j < 32; ++j) {
                            const int col = colw + c * 32 + j;
                            if (col < t.col_end && (!t.tri || col > row)) {
                                float s = __uint_as_float(r[j]) * scale;
                                smin = fminf(smin, s); smax = fmaxf(smax, s);
                                s = bias_correct(misc->beta[0], s);
                                if (!p.raw) s = fminf(fmaxf(s, -1.0f), 1.0f);
                                const float d = (p.metric == 0) ? __fmul_rn(2.0f, __fsub_rn(1.0f, s)) : acosf(s);
                                const bool same = (__ldg(p.col_cls + col) == my_cls);
                                if (same ? (d > thr) : (d < thr)) {
                                    const unsigned long long idx = atomicAdd(p.filter_count, 1ull);
                                    if ((long long)idx < p.filter_cap) {
                                        p.filter_rows[idx] = (int)__ldg(p.filter_perm + row);
                                        p.filter_cols[idx] = (int)__ldg(p.filter_perm + col);
                                        p.filter_dist[idx] = d;
                                    }
                                }
                            }
                        }
                    }
                }
            } else if constexpr (kEpi == EPI_BCE) {
                // ------------------- weighted binary cross entropy over the triangle ----------
                mbar_wait_relaxed(&misc->tfull[acc], acc_phase);
                tc_fence_after();
                const bool row_ok = row < t.row_end;
                const float my_nrm = (p.row_nrm != nullptr && row_ok) ? __ldg(p.row_nrm + row) : 1.0f;
                const int my_cls = row_ok ? __ldg(p.row_cls + row) : -1;
                float t_loss = 0.f, t_da = 0.f, t_dt = 0.f, t_dth = 0.f;     // this tile's terms of this thread (<= 64)
#pragma unroll 1
                for (int c = 0; c < kColsPerWarp / 32; ++c) {
                    uint32_t r[32];
                    tmem_ld32(taddr0 + c * 32, r);
                    tmem_ld_wait();
                    if (row_ok) {
#pragma unroll 4
                        for (int j = 0; j < 32; ++j) {
                            const int col = colw + c * 32 + j;
                            if (col < t.col_end && col > row) {
                                const float sv = bias_correct(misc->beta[0], __uint_as_float(r[j]) * scale);
                                float d, g2 = 0.f;
                                if (p.row_nrm != nullptr) {
                                    const float nc = __ldg(p.col_nrm + col);
                                    const float g = __fdiv_rn(__fmul_rn(2.0f, __fsub_rn(my_nrm, nc)), __fadd_rn(my_nrm, nc));
                                    g2 = __fmul_rn(g, g);
                                    d = __fadd_rn(__fmul_rn(2.0f, __fsub_rn(1.0f, sv)), __fmul_rn(p.theta, g2));
                                } else {
                                    d = __fmul_rn(2.0f, __fsub_rn(1.0f, sv));
                                }
                                const float margin = __fsub_rn(p.bce_threshold, d);
                                const float x = __fmul_rn(p.bce_alpha, margin);           // the logit (faceclass.py:26)
                                const float z = (__ldg(p.col_cls + col) == my_cls) ? 1.0f : 0.0f;
                                const float lw = 1.0f + (p.bce_pos_weight - 1.0f) * z;
                                // tf.nn.weighted_cross_entropy_with_logits: (1 - z) x + lw (log1p(exp(-|x|)) + max(-x, 0))
                                const float e = expf(-fabsf(x));
                                t_loss += (1.0f - z) * x + lw * (log1pf(e) + fmaxf(-x, 0.0f));
                                const float sig_neg = (x >= 0.0f) ? e / (1.0f + e) : 1.0f / (1.0f + e);   // sigmoid(-x)
                                const float gr = (1.0f - z) - lw * sig_neg;                                  // d loss / d logit
                                t_da += gr * margin;
                                t_dt += gr * p.bce_alpha;
                                t_dth -= gr * p.bce_alpha * g2;
                            }
                        }
                    }
                }
                bce_acc[0] += (double)t_loss; bce_acc[1] += (double)t_da; bce_acc[2] += (double)t_dt; bce_acc[3] += (double)t_dth;
            } else {
                // ------------------- PAIRWISE / ROWSTRIP: materialise distances --------------
                mbar_wait_relaxed(&misc->tfull[acc], acc_phase);
                tc_fence_after();
                const bool row_ok = row < t.row_end;
                const float my_nrm = (p.row_nrm != nullptr && row_ok) ? __ldg(p.row_nrm + row) : 1.0f;
                const bool mining = (kEpi == EPI_ROWSTRIP) && (p.mine_lab != nullptr);
                const long long my_lab = (mining && row_ok) ? __ldg(p.mine_lab + row) : 0;
                unsigned long long best_pos = 0ull, best_neg = ~0ull;
                const int cbeg = (kEpi == EPI_ROWSTRIP) ? t.cbeg : 0;
                // Distances leave through a 32 x 32 transpose in shared memory (one padded tile per warp): the accumulator hands
                // every thread ONE ROW, so a direct store would scatter a warp's 32 words over 32 rows (32 sectors per store,
                // 8 x the bytes at L2: the 1800 x 1800 mining strips ran at 0.4 TB/s); after the transpose a warp writes 128
                // contiguous bytes of one row per store.
                float* tile_s = reinterpret_cast<float*>(hist8) + (size_t)e * (32 * 33);
                const int row_w0 = t.row0 + (int)cta_rank * kRowsPerCta + q * 32;       // first row of this warp
#pragma unroll 1
                for (int c = 0; c < kColsPerWarp / 32; ++c) {
                    uint32_t r[32];
                    tmem_ld32(taddr0 + c * 32, r);
                    tmem_ld_wait();
                    const int col_c0 = colw + c * 32;
                    const bool store = (kEpi == EPI_PAIRWISE) || (p.out != nullptr);
                    __syncwarp();                                    // the previous chunk's tile has been read by every lane
                    if (row_ok) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int col = col_c0 + j;
                            const bool ok = (col < t.col_end) && (!t.tri || col > row);
                            float d = 0.f;
                            if (ok) {
                                float s = __uint_as_float(r[j]) * scale;
                                if (kEpi == EPI_PAIRWISE || col != row) { smin = fminf(smin, s); smax = fmaxf(smax, s); }
                                s = bias_correct(misc->beta[0], s);
                                if (!p.raw) s = fminf(fmaxf(s, -1.0f), 1.0f);
                                if (p.row_nrm != nullptr) d = classifier_distance(s, my_nrm, __ldg(p.col_nrm + col), p.theta);
                                else if (p.metric == 0) d = __fmul_rn(2.0f, __fsub_rn(1.0f, s));
                                else d = acosf(s);
                                if (kEpi == EPI_ROWSTRIP && mining && col != row) {
                                    const unsigned long long dk = (unsigned long long)__float_as_uint(d) << 32;
                                    const unsigned int ci = (unsigned int)(col - cbeg);
                                    if (__ldg(p.mine_lab + col) == my_lab) best_pos = max(best_pos, dk | (0xFFFFFFFFu - ci));
                                    else best_neg = min(best_neg, dk | ci);
                                }
                            }
                            if (store) tile_s[lane * 33 + j] = d;
                        }
                    }
                    if (store) {
                        __syncwarp();
                        const int col = col_c0 + lane;
                        const int rows_here = min(32, t.row_end - row_w0);
                        if (col < t.col_end) {
#pragma unroll 4
                            for (int rr = 0; rr < rows_here; ++rr) {
                                const int rowg = row_w0 + rr;
                                if (!t.tri || col > rowg) {
                                    long long base;
                                    if (p.tri_packed) base = (long long)rowg * p.n_rows - ((long long)rowg * (rowg + 1)) / 2 - rowg - 1;
                                    else base = (long long)rowg * p.out_ld - cbeg;
                                    p.out[base + col] = tile_s[rr * 33 + lane];
                                }
                            }
                        }
                    }
                }
                if (kEpi == EPI_ROWSTRIP && mining && row_ok) {
                    if (best_pos != 0ull) atomicMax(p.mine_pos_key + row, best_pos);
                    if (best_neg != ~0ull) atomicMin(p.mine_neg_key + row, best_neg);
                }
            }

            // release this accumulator stage to the MMA issuer
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (kCtaGroup == 2) mbar_arrive_remote(&misc->tempty[acc], leader_rank);
                else mbar_arrive(&misc->tempty[acc]);
            }
            ++it; ++tiles_done;
            }   // sub
        }

        if constexpr (kEpi == EPI_HIST) {
            flush_cta(cur_key);
            eps_cnt = warp_sum(eps_cnt);
            if (lane == 0 && eps_cnt) atomicAdd(p.counters + 0, (unsigned long long)eps_cnt);
        }
        if constexpr (kEpi == EPI_BCE) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                double v = bce_acc[i];
#pragma unroll
                for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0 && v != 0.0) atomicAdd(p.bce_out + i, v);
            }
        }
        smin = warp_min(smin); smax = warp_max(smax);
        if (lane == 0) {
            if (smin <= smax) {
                atomicMin(p.range_ord + 0, float_to_ordered(smin));
                atomicMax(p.range_ord + 1, float_to_ordered(smax));
            }
            if (e == 0 && cluster_rank == 0) atomicAdd(p.counters + 1, (unsigned long long)tiles_done);
        }
    }

    // ---- teardown (reconverge each warp first: the barriers below are .aligned)
    __syncwarp();
    tc_fence_before();
    if (kCtaGroup == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 2) { tc_fence_after(); tmem_dealloc<kCtaGroup>(tmem_base, kTmemCols); }
}

}  // namespace fnb
