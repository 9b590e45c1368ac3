// facenet_b200 -- triplet mining on P x K batches: per-anchor hardest positive / hardest negative and, per (anchor, positive),
// the semi-hard negative and the size of the margin-eligible candidate set; S batches per launch, outputs left on the device.
//
// NOT IN THE REFERENCE FORK: semantics are the ones stated in
// oracle/mining_oracle.py.  Conventions taken from the reference: P x K batches
// (facenet/facenet.py:89-123), same identity <=> equal label (facenet/apps/train_classifier.py:62-73),
// distance = metric 0 of pairwise_similarities (facenet/statistics.py:33-50), hardest pairs = within-class
// argmax / cross-class argmin (commented search in facenet/statistics.py:357-387).
//
// Launches per call (any number of batches), all on the handle's stream, no host synchronisation when every output
// tensor lives on the device:
//   1. split_rows_kernel        operands of all S * B rows
//   2. mine_init_kernel         labels -> int64, key arrays reset
//   3. gram_kernel<ROWSTRIP>    one region per batch; the epilogue folds the hardest positive / negative of every anchor into
//                               packed (distance, index) keys (running arg-extrema, one 64-bit atomic per row and tile) and
//                               writes the B x B fp32 distance strip of every batch (skipped when kmax == 0)
//   4. mine_rows_kernel         one CTA per anchor: decodes the keys; ordered list of positives; one warp per positive for the
//                               semi-hard argmin and the eligible count over the strip row (kmax == 0: mine_decode_kernel)
// fnb_mine_select_kth answers "the k-th margin-eligible negative of (anchor, positive) in index order" from the strips of the
// last call -- what upstream davidsandberg/facenet's select_triplets draws with np.random.randint (not in the fork).
#include "fnb_host.h"

#include <cmath>
#include <cstring>

namespace fnb {

constexpr int kMineThreads = 256;

__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int o) {
    return __shfl_xor_sync(0xffffffffu, v, o);
}

// key for "min distance, then lowest index": distances are >= +0 so their bit patterns order like uints
__device__ __forceinline__ unsigned long long min_key(float d, int idx) {
    return ((unsigned long long)__float_as_uint(d) << 32) | (unsigned int)idx;
}
// key for "max distance, then lowest index" under a max-reduction
__device__ __forceinline__ unsigned long long max_key(float d, int idx) {
    return ((unsigned long long)__float_as_uint(d) << 32) | (0xFFFFFFFFu - (unsigned int)idx);
}

// labels (int32 / int64) -> int64, arg-extrema keys and status reset
template <typename L>
__global__ void mine_init_kernel(const L* __restrict__ labels, long long n, long long* __restrict__ lab64,
                                 unsigned long long* __restrict__ pos_key, unsigned long long* __restrict__ neg_key,
                                 int* __restrict__ status)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        lab64[i] = (long long)labels[i];
        pos_key[i] = 0ull;
        neg_key[i] = ~0ull;
    }
    if (blockIdx.x == 0 && threadIdx.x < 4) status[threadIdx.x] = 0;
}

__device__ __forceinline__ int decode_pos(unsigned long long k) { return k ? (int)(0xFFFFFFFFu - (unsigned int)(k & 0xFFFFFFFFull)) : -1; }
__device__ __forceinline__ int decode_neg(unsigned long long k) { return (k != ~0ull) ? (int)(k & 0xFFFFFFFFull) : -1; }

// status[1..2] = exact range of the raw similarities (float bits) over all off-diagonal pairs of the call
__device__ __forceinline__ void write_range_status(const unsigned int* range_ord, int* status) {
    status[1] = (int)__float_as_uint(ordered_to_float(range_ord[0]));
    status[2] = (int)__float_as_uint(ordered_to_float(range_ord[1]));
}

// hardest-only mining (kmax == 0): keys -> indices
__global__ void mine_decode_kernel(const unsigned long long* __restrict__ pos_key, const unsigned long long* __restrict__ neg_key,
                                   long long n, int* __restrict__ hardest_pos, int* __restrict__ hardest_neg,
                                   const unsigned int* __restrict__ range_ord, int* __restrict__ status)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        hardest_pos[i] = decode_pos(pos_key[i]);
        hardest_neg[i] = decode_neg(neg_key[i]);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) write_range_status(range_ord, status);
}

// One CTA per anchor (global row a of batch a / b; all indices written are local to the batch).
// The row of the distance strip is staged in shared memory with every non-negative column (self, positives) set to +inf -- the
// inner loop then needs no class test: fp32(inf - d(a,p)) < alpha is false.  Each warp takes kPosGroup positives at a time in
// registers and walks the negatives once for all of them (one shared-memory load per negative and kPosGroup x 6 instructions).
constexpr int kPosGroup = 5;

__global__ void __launch_bounds__(kMineThreads)
mine_rows_kernel(const float* __restrict__ dist, long long ld, const long long* __restrict__ labels, int b, float alpha, int kmax,
                 const unsigned long long* __restrict__ pos_key, const unsigned long long* __restrict__ neg_key,
                 int* __restrict__ hardest_pos, int* __restrict__ hardest_neg, int* __restrict__ pos_index,
                 int* __restrict__ semi_hard, int* __restrict__ eligible, const unsigned int* __restrict__ range_ord,
                 int* __restrict__ status)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* d_neg = reinterpret_cast<float*>(smem_raw);                        // [b] distance of a negative, +inf for the rest
    int* pos_list = reinterpret_cast<int*>(d_neg + b);                        // [kmax]
    float* pos_d = reinterpret_cast<float*>(pos_list + kmax);                 // [kmax]
    constexpr int kMaxIters = 8192 / 32 / (kMineThreads / 32);                // 32-column steps per warp at the largest batch
    __shared__ int warp_cnt[kMineThreads / 32];
    __shared__ unsigned int same_mask[(kMineThreads / 32) * kMaxIters];
    __shared__ int s_npos;

    const long long ag = blockIdx.x;                        // global anchor row
    const int a = (int)(ag % b);                            // index inside its batch
    const long long* lab = labels + (ag - a);               // labels of the batch
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long la = lab[a];
    const float* src = dist + ag * ld;

    // ---- stage the row, classify the columns, ordered list of positives.  Every warp owns a contiguous range of columns and
    //      keeps its loads in flight (no block barrier between them: a barrier per 256 columns made this latency-bound);
    //      the same-label ballots are kept, one block barrier exchanges the warps' totals, and the positives are then
    //      written in ascending order.
    constexpr int kWarps = kMineThreads / 32;
    const int span = (((b + kWarps - 1) / kWarps) + 31) & ~31;           // columns per warp, a multiple of 32
    const int w_begin = warp * span;
    const int iters = span / 32;                                          // <= kMaxIters (b <= 8192)
    int mine = 0;                                                         // positives in this warp's range
    for (int it0 = 0; it0 < iters; it0 += 8) {
        float dv[8];
        long long lv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int n = w_begin + (it0 + u) * 32 + lane;
            const bool in = (it0 + u < iters) && (n < b);
            dv[u] = in ? src[n] : 0.f;
            lv[u] = in ? lab[n] : 0;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int n = w_begin + (it0 + u) * 32 + lane;
            const bool in = (it0 + u < iters) && (n < b);
            const bool same = in && (lv[u] == la);
            if (in) d_neg[n] = dv[u];                                     // raw for now: the positives' distances are read back below
            const unsigned m = __ballot_sync(0xffffffffu, same);
            if (it0 + u < iters) {
                if (lane == 0) same_mask[warp * kMaxIters + it0 + u] = m;
                mine += __popc(m);
            }
        }
    }
    // the anchor itself carries the label too: it is not a positive
    const bool has_self = (a >= w_begin) && (a < w_begin + span);
    if (lane == 0) warp_cnt[warp] = mine - (has_self ? 1 : 0);
    __syncthreads();
    int before = 0, running = 0;
    for (int w = 0; w < kWarps; ++w) { if (w < warp) before += warp_cnt[w]; running += warp_cnt[w]; }
    for (int it = 0; it < iters; ++it) {
        const unsigned m = same_mask[warp * kMaxIters + it];
        const int n = w_begin + it * 32 + lane;
        if ((m >> lane) & 1u) {
            int j = before + __popc(m & ((1u << lane) - 1u));
            // the anchor's own bit, if it precedes this column in the warp's range, is not counted
            if (has_self && a < n) --j;
            if (n != a && j < kmax) { pos_list[j] = n; pos_d[j] = d_neg[n]; }
            d_neg[n] = INFINITY;                                          // self and positives never act as negatives
        }
        before += __popc(m);
    }
    __syncthreads();
    if (tid == 0) {
        s_npos = running;
        // the arg-extrema were folded in the Gram epilogue
        hardest_pos[ag] = decode_pos(pos_key[ag]);
        hardest_neg[ag] = decode_neg(neg_key[ag]);
        if (running > kmax) atomicMax(status, running);
        if (ag == 0) write_range_status(range_ord, status);
    }
    __syncthreads();
    const int npos = min(s_npos, kmax);

    // ---- groups of kPosGroup positives, one warp per group: eligible counts and semi-hard argmins over the negatives.
    // The hardest negative (d_min, n_min) of the anchor is already known from the Gram epilogue's running argmin: a positive
    // that is closer than EVERY negative (d(a,p) < d_min, the usual case) has all negatives behind it, so its semi-hard
    // negative is n_min itself (if it passes the margin test) and only the eligible count needs the scan -- one subtract, one
    // compare and one predicated add per (negative, positive).  Positives at or beyond d_min take the full scan.
    const unsigned long long nk = neg_key[ag];
    const bool have_neg = (nk != ~0ull);
    const float d_min = have_neg ? __uint_as_float((unsigned int)(nk >> 32)) : INFINITY;
    const int n_min = have_neg ? (int)(nk & 0xFFFFFFFFull) : -1;
    const int ngroups = (kmax + kPosGroup - 1) / kPosGroup;
    for (int g = warp; g < ngroups; g += kMineThreads / 32) {
        const int j0 = g * kPosGroup;
        float dp[kPosGroup];
        bool easy = true;
#pragma unroll
        for (int i = 0; i < kPosGroup; ++i) {
            // a slot without a positive compares against -inf: nothing is eligible (inf - (-inf) and d - (-inf) are +inf)
            dp[i] = (j0 + i < npos) ? pos_d[j0 + i] : -INFINITY;
            easy = easy && (dp[i] < d_min);
        }
        int cnt[kPosGroup], semi[kPosGroup];
        if (j0 >= npos) {
#pragma unroll
            for (int i = 0; i < kPosGroup; ++i) { cnt[i] = 0; semi[i] = -1; }
        } else if (easy) {
            float cf[kPosGroup];
#pragma unroll
            for (int i = 0; i < kPosGroup; ++i) cf[i] = 0.f;
            for (int n = lane; n < b; n += 32) {
                const float dn = d_neg[n];
#pragma unroll
                for (int i = 0; i < kPosGroup; ++i)
                    if (__fsub_rn(dn, dp[i]) < alpha) cf[i] += 1.0f;     // exact: counts stay far below 2^24
            }
#pragma unroll
            for (int i = 0; i < kPosGroup; ++i) {
                int c = (int)cf[i];
#pragma unroll
                for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
                cnt[i] = c;
                semi[i] = (have_neg && __fsub_rn(d_min, dp[i]) < alpha) ? n_min : -1;
            }
        } else {
            float best_d[kPosGroup];
            int best_i[kPosGroup];
#pragma unroll
            for (int i = 0; i < kPosGroup; ++i) { best_d[i] = INFINITY; best_i[i] = -1; cnt[i] = 0; }
            for (int n = lane; n < b; n += 32) {
                const float dn = d_neg[n];
#pragma unroll
                for (int i = 0; i < kPosGroup; ++i) {
                    const bool el = __fsub_rn(dn, dp[i]) < alpha;        // fp32 subtraction, like the oracle
                    cnt[i] += el ? 1 : 0;
                    if (el && dn > dp[i] && dn < best_d[i]) { best_d[i] = dn; best_i[i] = n; }     // ascending n: ties keep the lowest index
                }
            }
#pragma unroll
            for (int i = 0; i < kPosGroup; ++i) {
                unsigned long long key = (best_i[i] >= 0) ? min_key(best_d[i], best_i[i]) : ~0ull;
                int c = cnt[i];
#pragma unroll
                for (int o = 16; o; o >>= 1) {
                    c += __shfl_xor_sync(0xffffffffu, c, o);
                    key = min(key, shfl_xor_u64(key, o));
                }
                cnt[i] = c;
                semi[i] = (key != ~0ull) ? (int)(key & 0xFFFFFFFFull) : -1;
            }
        }
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < kPosGroup; ++i) {
                const int j = j0 + i;
                if (j < kmax) {
                    const long long o = ag * kmax + j;
                    pos_index[o] = (j < npos) ? pos_list[j] : -1;
                    semi_hard[o] = (j < npos) ? semi[i] : -1;
                    eligible[o] = (j < npos) ? cnt[i] : 0;
                }
            }
        }
    }
}

// One warp per query (a, p, k): the k-th (0-based, ascending column index) negative n of anchor a with
// fp32(d(a, n) - d(a, p)) < alpha; -1 when there are not that many (or the query is malformed).  a is a global row of the last
// call, p and the result are local to a's batch.
__global__ void __launch_bounds__(256)
mine_select_kth_kernel(const float* __restrict__ dist, long long ld, const long long* __restrict__ labels, int b, long long rows,
                       float alpha, const int* __restrict__ qa, const int* __restrict__ qp, const int* __restrict__ qk, long long m,
                       int* __restrict__ out)
{
    const long long q = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (q >= m) return;
    const long long ag = qa[q];
    const int p = qp[q], k = qk[q];
    int result = -1;
    if (ag >= 0 && ag < rows && p >= 0 && p < b && k >= 0) {
        const int a = (int)(ag % b);
        const long long* lab = labels + (ag - a);
        const long long la = lab[a];
        const float* src = dist + ag * ld;
        const float dp = src[p];
        int seen = 0;
        for (int c0 = 0; c0 < b && result < 0; c0 += 32) {
            const int n = c0 + lane;
            const bool el = (n < b) && (lab[n] != la) && (__fsub_rn(src[n], dp) < alpha);
            const unsigned mask = __ballot_sync(0xffffffffu, el);
            const int cnt = __popc(mask);
            if (k < seen + cnt) {
                // the (k - seen)-th set bit of mask
                unsigned mm = mask;
                for (int i = 0; i < k - seen; ++i) mm &= mm - 1;
                result = c0 + __ffs(mm) - 1;
            }
            seen += cnt;
        }
    }
    if (lane == 0) out[q] = result;
}

}  // namespace fnb

using namespace fnb;

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return h->fail(FNB_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); } while (0)

static int check_i32(fnb_context* h, const DLTensor* t, const char* name, long long rows, long long cols, DLView* v) {
    int rc = dl_view(h, t, name, 1, 2, v); if (rc) return rc;
    if (v->code != kDLInt || v->bits != 32) return h->fail(FNB_ERR_INVALID, "%s must be int32", name);
    if (v->rows * v->cols != rows * cols) return h->fail(FNB_ERR_INVALID, "%s has %lld elements, expected %lld", name, v->rows * v->cols, rows * cols);
    return FNB_OK;
}

// turns the status words of the last mining call into the error codes of the synchronous form
static int mine_status_to_error(fnb_context* h, const int* st, double atol, long long b, int kmax) {
    float smin, smax;
    memcpy(&smin, &st[1], 4); memcpy(&smax, &st[2], 4);
    const double lim = 1.0 + atol;
    if (b > 1 && ((double)smin < -lim || (double)smax > lim || smin != smin || smax != smax))
        return h->fail(FNB_ERR_NOT_NORMALIZED, "embeddings must be normalized to 1, range %.9g %.9g", smin, smax);
    if (kmax > 0 && st[0] > kmax)
        return h->fail(FNB_ERR_INVALID, "kmax = %d is smaller than the largest class size - 1 = %d", kmax, st[0]);
    return FNB_OK;
}

extern "C" int fnb_mine_batched(fnb_handle h, const DLTensor* emb, const DLTensor* labels, int nbatches, float alpha,
                                const fnb_options* opt_in, int kmax, DLTensor* hardest_pos, DLTensor* hardest_neg,
                                DLTensor* pos_index, DLTensor* semi_hard, DLTensor* eligible, DLTensor* status, fnb_stats* stats)
{
    if (!h) return FNB_ERR_INVALID;
    fnb_options opt; if (opt_in) opt = *opt_in; else fnb_default_options(&opt);
    if (!hardest_pos || !hardest_neg || kmax < 0 || (kmax > 0 && (!pos_index || !semi_hard || !eligible)))
        return h->fail(FNB_ERR_INVALID, "NULL output / negative kmax");
    if (nbatches < 1) return h->fail(FNB_ERR_INVALID, "nbatches must be >= 1");
    CK(cudaSetDevice(h->device));
    if (stats) memset(stats, 0, sizeof(*stats));
    GramOperands op;
    if (opt.mode == FNB_MODE_AUTO) opt.mode = FNB_MODE_FP16X3;
    if (mode_info(opt.mode, &op.num_pass, &op.tf32, &op.fmt, &op.elem_bytes, &op.prescale)) return h->fail(FNB_ERR_INVALID, "bad mode %d", opt.mode);
    DLView ve, vl, vo[5], vs;
    int rc = dl_view(h, emb, "embeddings", 2, 2, &ve); if (rc) return rc;
    if ((rc = dl_view(h, labels, "labels", 1, 1, &vl))) return rc;
    if (vl.code != kDLInt || (vl.bits != 32 && vl.bits != 64)) return h->fail(FNB_ERR_INVALID, "labels must be int32 or int64");
    if (vl.rows != ve.rows) return h->fail(FNB_ERR_INVALID, "len(labels) != embeddings.shape[0]");
    const long long rows = ve.rows;
    h->mine_rows = 0;                                    // the strips of an earlier call are gone from here on
    if (rows == 0) return FNB_OK;
    if ((rc = dl_check_embeddings(h, ve, "embeddings"))) return rc;
    if (rows % nbatches) return h->fail(FNB_ERR_INVALID, "%lld rows do not split into %d equal batches", rows, nbatches);
    const long long b = rows / nbatches;
    const int d = (int)ve.cols;
    if (b > 8192) return h->fail(FNB_ERR_UNSUPPORTED, "mining batches are limited to 8192 rows (got %lld)", b);
    DLTensor* outs[5] = {hardest_pos, hardest_neg, pos_index, semi_hard, eligible};
    const char* names[5] = {"hardest_pos", "hardest_neg", "pos_index", "semi_hard", "eligible"};
    const int nouts = kmax > 0 ? 5 : 2;
    bool all_dev = true;
    for (int i = 0; i < nouts; ++i) {
        if ((rc = check_i32(h, outs[i], names[i], rows, i < 2 ? 1 : kmax, &vo[i]))) return rc;
        all_dev = all_dev && vo[i].on_device;
    }
    if (status) {
        if ((rc = check_i32(h, status, "status", 4, 1, &vs))) return rc;
        all_dev = all_dev && vs.on_device;
    }
    const long long ld = ((b + 3) / 4) * 4;
    const size_t strip_bytes = kmax > 0 ? (size_t)rows * ld * 4 : 0;
    if (strip_bytes > ((size_t)16 << 30)) return h->fail(FNB_ERR_UNSUPPORTED, "distance strips of %d batches need %zu bytes (limit 16 GiB): use fewer batches per call", nbatches, strip_bytes);

    const void* de = nullptr; const void* dl = nullptr;
    CK(cudaEventRecord(h->ev[0], h->stream));
    if ((rc = dl_to_device(h, ve, (size_t)rows * d * 4, h->stage_a, &de))) return rc;
    if ((rc = dl_to_device(h, vl, (size_t)rows * (vl.bits / 8), h->stage_lab, &dl))) return rc;
    if ((rc = prepare_operand(h, opt.mode, (const float*)de, nullptr, rows, d, false, op))) return rc;
    if ((rc = self_b_maps(h, op, d))) return rc;

    // labels as int64, arg-extrema keys, status words
    CK(h->mine_lab.ensure((size_t)rows * 8));
    CK(h->mine_keys.ensure((size_t)rows * 16));
    CK(h->mine_status.ensure(64));
    long long* lab64 = h->mine_lab.as<long long>();
    unsigned long long* pos_key = h->mine_keys.as<unsigned long long>();
    unsigned long long* neg_key = pos_key + rows;
    int* st_dev = h->mine_status.as<int>();
    {
        const unsigned blocks = (unsigned)std::min<long long>((rows + 255) / 256, 148LL * 8);
        if (vl.bits == 64) mine_init_kernel<long long><<<blocks, 256, 0, h->stream>>>((const long long*)dl, rows, lab64, pos_key, neg_key, st_dev);
        else mine_init_kernel<int><<<blocks, 256, 0, h->stream>>>((const int*)dl, rows, lab64, pos_key, neg_key, st_dev);
        CK(cudaGetLastError());
    }

    // stage 1: one region per batch -- B x B distances (every ordered pair, diagonal included) + running arg-extrema
    // CTA pairs on 256 x 256 tiles for real batches: half the operand bytes per distance of 128 x 128 tiles (an 1800-row
    // batch in fp16x3 is L2 -> SM bandwidth bound: 112 MB of operand boxes per batch at 128, 64 MB at 256); small batches keep
    // the finer tiles
    const int cg = (b >= 768) ? 2 : 1;
    const int tile = kRowsPerCta * cg;
    std::vector<RegionDev> regs((size_t)nbatches);
    for (int sb = 0; sb < nbatches; ++sb) {
        RegionDev r = {};
        r.row_begin = r.col_begin = (int)(sb * b); r.row_end = r.col_end = (int)((sb + 1) * b);
        regs[sb] = r;
    }
    finish_regions(regs, tile);
    if ((rc = upload_regions(h, regs))) return rc;
    if ((rc = reset_scalars(h))) return rc;
    if (strip_bytes) CK(h->strip.ensure(strip_bytes));
    GramParams p = {};
    p.regions = h->regions.as<RegionDev>(); p.nregions = nbatches; p.total_tiles = regs.back().tile_begin;
    p.shard = ShardSpec{1, 0, 1, nullptr};
    p.kblocks = d / (128 / op.elem_bytes);
    p.acc_scale = gram_acc_scale(op);
    p.operand_fmt = op.fmt;
    DeviceScalars* sc = h->counters.as<DeviceScalars>();
    p.counters = sc->counters; p.range_ord = sc->range_ord;
    p.out = strip_bytes ? h->strip.as<float>() : nullptr; p.out_ld = ld; p.tri_packed = 0; p.metric = 0;
    p.n_rows = (int)rows; p.n_cols = (int)rows;
    p.mine_lab = lab64; p.mine_pos_key = pos_key; p.mine_neg_key = neg_key;
    if ((rc = upload_bias(h, op.mode, d, opt.bias_correction < 0, &p.bias_beta))) return rc;
    CK(cudaEventRecord(h->ev[1], h->stream));
    if ((rc = launch_gram(h, cg, EPI_ROWSTRIP, opt.max_ctas, op, p, 0))) return rc;

    // stage 2: per-anchor selection, into the caller's device tensors or a staging area
    const size_t n_out = (size_t)rows * 2 + (size_t)rows * kmax * 3;
    int* dst[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    if (all_dev) {
        for (int i = 0; i < nouts; ++i) dst[i] = (int*)vo[i].data;
    } else {
        CK(h->mine_out.ensure(n_out * 4));
        dst[0] = h->mine_out.as<int>(); dst[1] = dst[0] + rows; dst[2] = dst[1] + rows;
        dst[3] = dst[2] + (size_t)rows * kmax; dst[4] = dst[3] + (size_t)rows * kmax;
    }
    if (kmax > 0) {
        const size_t smem = (size_t)b * 4 + (size_t)kmax * 8;
        if (smem > 200 * 1024) return h->fail(FNB_ERR_UNSUPPORTED, "mining row does not fit shared memory");
        CK(cudaFuncSetAttribute(mine_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mine_rows_kernel<<<(unsigned)rows, kMineThreads, smem, h->stream>>>(h->strip.as<float>(), ld, lab64, (int)b, alpha, kmax, pos_key, neg_key,
                                                                           dst[0], dst[1], dst[2], dst[3], dst[4], sc->range_ord, st_dev);
    } else {
        const unsigned blocks = (unsigned)std::min<long long>((rows + 255) / 256, 148LL * 8);
        mine_decode_kernel<<<blocks, 256, 0, h->stream>>>(pos_key, neg_key, rows, dst[0], dst[1], sc->range_ord, st_dev);
    }
    CK(cudaGetLastError());
    CK(cudaEventRecord(h->ev[2], h->stream));
    h->mine_rows = rows; h->mine_b = b; h->mine_ld = ld; h->mine_kmax = kmax; h->mine_atol = opt.atol;
    h->mine_has_strip = strip_bytes != 0;
    if (stats) {
        stats->n_pairs = (uint64_t)nbatches * (uint64_t)b * (uint64_t)b;
        stats->kernel_launches = 4;                     // split_rows, mine_init, gram<ROWSTRIP>, mine_rows | mine_decode
        stats->grid_ctas = (uint32_t)h->last_grid;
        stats->mode_used = op.mode;
    }
    if (all_dev) {
        // stream-ordered: results, and the status words (kmax needed, similarity range), stay on the device; fnb_mine_check
        // reports errors and timings after the caller has synchronised or wants to
        if (status) CK(cudaMemcpyAsync(vs.data, st_dev, 16, cudaMemcpyDeviceToDevice, h->stream));
        return FNB_OK;
    }

    CK(h->pinned.ensure(n_out * 4 + 4096));
    int* host = reinterpret_cast<int*>((char*)h->pinned.p + 4096);
    CK(cudaMemcpyAsync(host, h->mine_out.p, n_out * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(h->pinned.p, st_dev, 16, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    int st[4];
    memcpy(st, h->pinned.p, 16);
    if (stats) {
        float ms = 0.f, pm = 0.f;
        cudaEventElapsedTime(&ms, h->ev[1], h->ev[2]);
        cudaEventElapsedTime(&pm, h->ev[0], h->ev[1]);
        stats->kernel_ms = ms; stats->prepare_ms = pm;
        memcpy(&stats->smin, &st[1], 4); memcpy(&stats->smax, &st[2], 4);
    }
    if (status) {
        if (vs.on_device) CK(cudaMemcpyAsync(vs.data, st_dev, 16, cudaMemcpyDeviceToDevice, h->stream));
        else memcpy(vs.data, st, 16);
    }
    if ((rc = mine_status_to_error(h, st, opt.atol, b, kmax))) return rc;
    const int* src[5] = {host, host + rows, host + 2 * rows, host + 2 * rows + (size_t)rows * kmax, host + 2 * rows + (size_t)rows * kmax * 2};
    for (int i = 0; i < nouts; ++i) {
        const size_t bytes = (size_t)rows * (i < 2 ? 1 : kmax) * 4;
        if (vo[i].on_device) CK(cudaMemcpyAsync(vo[i].data, dst[i], bytes, cudaMemcpyDeviceToDevice, h->stream));
        else memcpy(vo[i].data, src[i], bytes);
    }
    return FNB_OK;
}

extern "C" int fnb_mine_check(fnb_handle h, int32_t* status_out, fnb_stats* stats)
{
    if (!h) return FNB_ERR_INVALID;
    if (h->mine_rows == 0) return h->fail(FNB_ERR_INVALID, "fnb_mine_check: no mining call to check");
    CK(cudaSetDevice(h->device));
    CK(h->pinned.ensure(4096));
    CK(cudaMemcpyAsync(h->pinned.p, h->mine_status.p, 16, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    int st[4];
    memcpy(st, h->pinned.p, 16);
    if (status_out) memcpy(status_out, st, 16);
    if (stats) {
        float ms = 0.f, pm = 0.f;
        cudaEventElapsedTime(&ms, h->ev[1], h->ev[2]);
        cudaEventElapsedTime(&pm, h->ev[0], h->ev[1]);
        stats->kernel_ms = ms; stats->prepare_ms = pm;
        memcpy(&stats->smin, &st[1], 4); memcpy(&stats->smax, &st[2], 4);
    }
    return mine_status_to_error(h, st, h->mine_atol, h->mine_b, h->mine_kmax);
}

extern "C" int fnb_mine_select_kth(fnb_handle h, const DLTensor* anchors, const DLTensor* positives, const DLTensor* kth,
                                   float alpha, DLTensor* out)
{
    if (!h) return FNB_ERR_INVALID;
    if (h->mine_rows == 0 || !h->mine_has_strip)
        return h->fail(FNB_ERR_INVALID, "fnb_mine_select_kth needs the distance strips of a preceding fnb_mine / fnb_mine_batched call with kmax > 0");
    CK(cudaSetDevice(h->device));
    DLView va, vp, vk, vo;
    int rc = dl_view(h, anchors, "anchors", 1, 1, &va); if (rc) return rc;
    const long long m = va.rows;
    if (va.code != kDLInt || va.bits != 32) return h->fail(FNB_ERR_INVALID, "anchors must be int32");
    if ((rc = check_i32(h, positives, "positives", m, 1, &vp))) return rc;
    if ((rc = check_i32(h, kth, "kth", m, 1, &vk))) return rc;
    if ((rc = check_i32(h, out, "out", m, 1, &vo))) return rc;
    if (m == 0) return FNB_OK;
    const void *da, *dp, *dk;
    CK(h->select_io.ensure((size_t)m * 16));
    // host queries are staged into one device buffer (three int32 arrays + the result)
    int* io = h->select_io.as<int>();
    auto stage = [&](const DLView& v, int slot, const void** ptr) -> int {
        if (v.on_device) { *ptr = v.data; return FNB_OK; }
        CK(cudaMemcpyAsync(io + (size_t)slot * m, v.data, (size_t)m * 4, cudaMemcpyHostToDevice, h->stream));
        *ptr = io + (size_t)slot * m;
        return FNB_OK;
    };
    if ((rc = stage(va, 0, &da)) || (rc = stage(vp, 1, &dp)) || (rc = stage(vk, 2, &dk))) return rc;
    int* dout = vo.on_device ? (int*)vo.data : io + (size_t)3 * m;
    const unsigned blocks = (unsigned)((m * 32 + 255) / 256);
    mine_select_kth_kernel<<<blocks, 256, 0, h->stream>>>(h->strip.as<float>(), h->mine_ld, h->mine_lab.as<long long>(), (int)h->mine_b,
                                                          h->mine_rows, alpha, (const int*)da, (const int*)dp, (const int*)dk, m, dout);
    CK(cudaGetLastError());
    if (!vo.on_device) {
        CK(cudaMemcpyAsync(vo.data, dout, (size_t)m * 4, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    } else if (!va.on_device || !vp.on_device || !vk.on_device) {
        CK(cudaStreamSynchronize(h->stream));            // the staged queries were read from the caller's host arrays
    }
    return FNB_OK;
}

// one batch, host outputs: the synchronous form
extern "C" int fnb_mine(fnb_handle h, const DLTensor* emb, const DLTensor* labels, float alpha, const fnb_options* opt_in,
                        int32_t* hardest_pos, int32_t* hardest_neg, int kmax, int32_t* pos_index, int32_t* semi_hard,
                        int32_t* eligible, fnb_stats* stats)
{
    if (!h) return FNB_ERR_INVALID;
    if (!emb || emb->ndim != 2) return h->fail(FNB_ERR_INVALID, "embeddings: ndim not supported");
    if (!hardest_pos || !hardest_neg || kmax < 0 || (kmax > 0 && (!pos_index || !semi_hard || !eligible)))
        return h->fail(FNB_ERR_INVALID, "NULL output / negative kmax");
    const int64_t b = emb->shape[0];
    int64_t shape1[1] = {b}, shape2[2] = {b, kmax};
    auto host_i32 = [&](int32_t* ptr, bool two) {
        DLTensor t = {};
        t.data = ptr; t.device.device_type = kDLCPU; t.ndim = two ? 2 : 1; t.dtype.code = kDLInt; t.dtype.bits = 32; t.dtype.lanes = 1;
        t.shape = two ? shape2 : shape1;
        return t;
    };
    DLTensor t_hp = host_i32(hardest_pos, false), t_hn = host_i32(hardest_neg, false);
    DLTensor t_pi = host_i32(pos_index, true), t_sh = host_i32(semi_hard, true), t_el = host_i32(eligible, true);
    return fnb_mine_batched(h, emb, labels, 1, alpha, opt_in, kmax, &t_hp, &t_hn, kmax > 0 ? &t_pi : nullptr,
                            kmax > 0 ? &t_sh : nullptr, kmax > 0 ? &t_el : nullptr, nullptr, stats);
}

// False pairs at one threshold over a labelled set (FalseExamples, facenet/statistics.py:334-387): candidate list only; the greedy
// per-class selection of the reference is host work on the (short) list.
extern "C" int fnb_false_pairs(fnb_handle h, const DLTensor* emb, const DLTensor* labels, double threshold, const fnb_options* opt_in,
                               long long capacity, int32_t* rows, int32_t* cols, float* dist, uint64_t* count, fnb_stats* stats)
{
    if (!h) return FNB_ERR_INVALID;
    fnb_options opt; if (opt_in) opt = *opt_in; else fnb_default_options(&opt);
    if (opt.metric != 0 && opt.metric != 1) return h->fail(FNB_ERR_BAD_METRIC, "Undefined similarity metric %d", opt.metric);
    if (!count || capacity < 0 || (capacity > 0 && (!rows || !cols || !dist))) return h->fail(FNB_ERR_INVALID, "NULL output");
    CK(cudaSetDevice(h->device));
    if (stats) memset(stats, 0, sizeof(*stats));
    *count = 0;
    GramOperands op;
    opt.mode = FNB_MODE_FP16X3;                          // materialised decisions: the strict split
    DLView ve, vl;
    int rc = dl_view(h, emb, "embeddings", 2, 2, &ve); if (rc) return rc;
    if ((rc = dl_check_embeddings(h, ve, "embeddings"))) return rc;
    if ((rc = dl_view(h, labels, "labels", 1, 1, &vl))) return rc;
    if (vl.code != kDLInt || (vl.bits != 32 && vl.bits != 64)) return h->fail(FNB_ERR_INVALID, "labels must be int32 or int64");
    if (vl.rows != ve.rows) return h->fail(FNB_ERR_INVALID, "len(labels) != embeddings.shape[0]");
    const long long n = ve.rows;
    const int d = (int)ve.cols;
    if (n < 2) return FNB_OK;
    const void* de = nullptr; const void* dl = nullptr;
    CK(cudaEventRecord(h->ev[0], h->stream));
    h->last_h2d_bytes = 0; h->h2d_timed = false; h->h2d_timed_bytes = 0;
    if ((rc = dl_to_device(h, vl, (size_t)n * (vl.bits / 8), h->stage_lab, &dl))) return rc;
    if ((rc = sort_labels(h, dl, vl.bits, n))) return rc;
    if ((rc = dl_to_device(h, ve, (size_t)n * d * 4, h->stage_a, &de))) return rc;
    if ((rc = prepare_operand(h, opt.mode, (const float*)de, h->perm.as<long long>(), n, d, false, op, opt.normalize))) return rc;
    if ((rc = self_b_maps(h, op, d))) return rc;
    const int cg = 2, tile = kRowsPerCta * cg;
    std::vector<RegionDev> regs;
    {
        const long long rr = std::max<long long>(tile, std::min<long long>(32768, ((n / 6) / tile) * tile));
        for (long long r0 = 0; r0 < n; r0 += rr) {
            const long long r1 = std::min(n, r0 + rr);
            RegionDev a = {}; a.row_begin = a.col_begin = (int)r0; a.row_end = a.col_end = (int)r1; a.tri = 1; regs.push_back(a);
            if (r1 < n) { RegionDev b = {}; b.row_begin = (int)r0; b.row_end = (int)r1; b.col_begin = (int)r1; b.col_end = (int)n; regs.push_back(b); }
        }
    }
    finish_regions(regs, tile);
    if ((rc = upload_regions(h, regs))) return rc;
    if ((rc = reset_scalars(h))) return rc;
    const size_t cap = (size_t)capacity;
    CK(h->mine_out.ensure(cap * 12 + 64));
    CK(h->mine_status.ensure(64));
    CK(cudaMemsetAsync(h->mine_status.p, 0, 16, h->stream));
    GramParams p = {};
    p.regions = h->regions.as<RegionDev>(); p.nregions = (int)regs.size() - 1; p.total_tiles = regs.back().tile_begin;
    p.shard = ShardSpec{1, 0, 1, nullptr};
    p.kblocks = d / (128 / op.elem_bytes);
    p.acc_scale = gram_acc_scale(op);
    p.operand_fmt = op.fmt;
    DeviceScalars* sc = h->counters.as<DeviceScalars>();
    p.counters = sc->counters; p.range_ord = sc->range_ord;
    p.row_cls = h->cls.as<int32_t>(); p.col_cls = p.row_cls;
    p.metric = opt.metric; p.raw = opt.raw_distance;
    p.n_rows = (int)n; p.n_cols = (int)n;
    p.filter_threshold = (float)threshold;
    p.filter_rows = h->mine_out.as<int>(); p.filter_cols = p.filter_rows + cap; p.filter_dist = reinterpret_cast<float*>(p.filter_cols + cap);
    p.filter_count = h->mine_status.as<unsigned long long>();
    p.filter_cap = capacity; p.filter_perm = h->perm.as<long long>();
    if ((rc = upload_bias(h, op.mode, d, opt.bias_correction < 0, &p.bias_beta))) return rc;
    CK(cudaEventRecord(h->ev[1], h->stream));
    if ((rc = launch_gram(h, cg, EPI_FILTER, opt.max_ctas, op, p, 0))) return rc;
    CK(cudaEventRecord(h->ev[2], h->stream));
    CK(h->pinned.ensure(16384));
    DeviceScalars hs;
    CK(cudaMemcpyAsync(h->pinned.p, h->counters.p, sizeof(hs), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(h->pinned.as<char>() + 1024, h->mine_status.p, 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    memcpy(&hs, h->pinned.p, sizeof(hs));
    unsigned long long found = 0;
    memcpy(&found, h->pinned.as<char>() + 1024, 8);
    *count = found;
    const float smin = ordered_to_float(hs.range_ord[0]), smax = ordered_to_float(hs.range_ord[1]);
    if (stats) {
        float ms = 0.f, pm = 0.f;
        cudaEventElapsedTime(&ms, h->ev[1], h->ev[2]);
        cudaEventElapsedTime(&pm, h->ev[0], h->ev[1]);
        stats->kernel_ms = ms; stats->prepare_ms = pm; stats->smin = smin; stats->smax = smax;
        stats->n_pairs = (uint64_t)n * (uint64_t)(n - 1) / 2; stats->tiles = hs.counters[1];
        stats->kernel_launches = 4; stats->grid_ctas = (uint32_t)h->last_grid; stats->mode_used = op.mode;
    }
    const double lim = 1.0 + (double)opt.atol;
    if (!opt.raw_distance && ((double)smin < -lim || (double)smax > lim || smin != smin || smax != smax))
        return h->fail(FNB_ERR_NOT_NORMALIZED, "embeddings must be normalized to 1, range %.9g %.9g", smin, smax);
    const size_t take = (size_t)std::min<unsigned long long>(found, (unsigned long long)capacity);
    if (take) {
        // on the handle's stream (never the null stream): the list was written by the launch queued there
        CK(cudaMemcpyAsync(rows, p.filter_rows, take * 4, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(cols, p.filter_cols, take * 4, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(dist, p.filter_dist, take * 4, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    return FNB_OK;
}
