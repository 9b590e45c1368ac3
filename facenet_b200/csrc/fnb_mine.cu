// facenet_b200 -- triplet mining on one batch: per-anchor hardest positive / hardest negative and,
// per (anchor, positive), the semi-hard negative and the size of the margin-eligible candidate set.
//
// NOT IN THE REFERENCE FORK: semantics are the ones stated in
// oracle/mining_oracle.py.  Conventions taken from the reference: P x K batches
// (facenet/facenet.py:89-123), same identity <=> equal label (facenet/apps/train_classifier.py:62-73),
// distance = metric 0 of pairwise_similarities (facenet/statistics.py:33-50), hardest pairs = within-class
// argmax / cross-class argmin (commented search in facenet/statistics.py:357-387).
//
// Two stages on one stream:
//   1. Gram kernel with the ROWSTRIP epilogue: the B x B fp32 distance matrix (13 MB at B = 1800, stays in
//      the 126 MB L2) -- tensor cores, same arithmetic modes as the verification path;
//   2. mine_rows_kernel: one CTA per anchor row, row of distances + labels staged in shared memory,
//      ordered compaction of the positives, packed (distance, index) keys for the arg-reductions
//      (ties -> lowest index), one warp per positive for the semi-hard scan.
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

template <typename L>
__global__ void __launch_bounds__(kMineThreads)
mine_rows_kernel(const float* __restrict__ dist, long long ld, const L* __restrict__ labels, int b, float alpha, int kmax,
                 int* __restrict__ hardest_pos, int* __restrict__ hardest_neg, int* __restrict__ pos_index,
                 int* __restrict__ semi_hard, int* __restrict__ eligible, int* __restrict__ overflow)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* d_row = reinterpret_cast<float*>(smem_raw);                        // [b]
    unsigned char* is_pos = reinterpret_cast<unsigned char*>(d_row + b);      // [b] 1 = positive, 0 = negative, 2 = self
    int* pos_list = reinterpret_cast<int*>(is_pos + ((b + 15) / 16) * 16);    // [kmax]
    __shared__ unsigned long long red_a[kMineThreads / 32], red_b[kMineThreads / 32];
    __shared__ int warp_cnt[kMineThreads / 32];
    __shared__ int s_npos;

    const int a = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const L la = labels[a];
    const float* src = dist + (long long)a * ld;

    // ---- stage the row, classify the columns, arg-reductions, ordered list of positives
    unsigned long long best_pos = 0ull, best_neg = ~0ull;
    int running = 0;                                        // positives found in earlier chunks
    for (int c0 = 0; c0 < b; c0 += kMineThreads) {
        const int n = c0 + tid;
        int flag = 0;
        if (n < b) {
            const float d = src[n];
            d_row[n] = d;
            const bool same = (labels[n] == la);
            if (n == a) is_pos[n] = 2;
            else if (same) { is_pos[n] = 1; flag = 1; best_pos = max(best_pos, max_key(d, n)); }
            else { is_pos[n] = 0; best_neg = min(best_neg, min_key(d, n)); }
        }
        const unsigned m = __ballot_sync(0xffffffffu, flag);
        if (lane == 0) warp_cnt[warp] = __popc(m);
        __syncthreads();
        int before = running;
        for (int w = 0; w < warp; ++w) before += warp_cnt[w];
        int total = 0;
        for (int w = 0; w < kMineThreads / 32; ++w) total += warp_cnt[w];
        if (flag) {
            const int j = before + __popc(m & ((1u << lane) - 1u));
            if (j < kmax) pos_list[j] = n;
        }
        running += total;
        __syncthreads();
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        best_pos = max(best_pos, shfl_xor_u64(best_pos, o));
        best_neg = min(best_neg, shfl_xor_u64(best_neg, o));
    }
    if (lane == 0) { red_a[warp] = best_pos; red_b[warp] = best_neg; }
    if (tid == 0) s_npos = running;
    __syncthreads();
    const int npos = min(s_npos, kmax);
    if (tid == 0) {
        unsigned long long bp = 0ull, bn = ~0ull;
        for (int w = 0; w < kMineThreads / 32; ++w) { bp = max(bp, red_a[w]); bn = min(bn, red_b[w]); }
        hardest_pos[a] = (s_npos > 0) ? (int)(0xFFFFFFFFu - (unsigned int)(bp & 0xFFFFFFFFull)) : -1;
        hardest_neg[a] = (bn != ~0ull) ? (int)(bn & 0xFFFFFFFFull) : -1;
        if (s_npos > kmax) atomicMax(overflow, s_npos);
    }

    // ---- one warp per positive: eligible count and semi-hard argmin over the negatives
    for (int j = warp; j < kmax; j += kMineThreads / 32) {
        int p = -1, cnt = 0;
        unsigned long long best = ~0ull;
        if (j < npos) {
            p = pos_list[j];
            const float dp = d_row[p];
            for (int n = lane; n < b; n += 32) {
                if (is_pos[n] == 0) {
                    const float dn = d_row[n];
                    if (__fsub_rn(dn, dp) < alpha) {                     // fp32 subtraction, like the oracle
                        ++cnt;
                        if (dn > dp) best = min(best, min_key(dn, n));
                    }
                }
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
                best = min(best, shfl_xor_u64(best, o));
            }
        }
        if (lane == 0) {
            const long long o = (long long)a * kmax + j;
            pos_index[o] = p;
            semi_hard[o] = (best != ~0ull) ? (int)(best & 0xFFFFFFFFull) : -1;
            eligible[o] = cnt;
        }
    }
}

}  // namespace fnb

using namespace fnb;

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return h->fail(FNB_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); } while (0)

extern "C" int fnb_mine(fnb_handle h, const DLTensor* emb, const DLTensor* labels, float alpha, const fnb_options* opt_in,
                        int32_t* hardest_pos, int32_t* hardest_neg, int kmax, int32_t* pos_index, int32_t* semi_hard,
                        int32_t* eligible, fnb_stats* stats)
{
    if (!h) return FNB_ERR_INVALID;
    fnb_options opt; if (opt_in) opt = *opt_in; else fnb_default_options(&opt);
    if (!hardest_pos || !hardest_neg || kmax < 0 || (kmax > 0 && (!pos_index || !semi_hard || !eligible)))
        return h->fail(FNB_ERR_INVALID, "NULL output / negative kmax");
    CK(cudaSetDevice(h->device));
    if (stats) memset(stats, 0, sizeof(*stats));
    GramOperands op;
    if (opt.mode == FNB_MODE_AUTO) opt.mode = FNB_MODE_FP16X3;
    if (mode_info(opt.mode, &op.num_pass, &op.tf32, &op.fmt, &op.elem_bytes, &op.prescale)) return h->fail(FNB_ERR_INVALID, "bad mode %d", opt.mode);
    DLView ve, vl;
    int rc = dl_view(h, emb, "embeddings", 2, 2, &ve); if (rc) return rc;
    if ((rc = dl_check_embeddings(h, ve, "embeddings"))) return rc;
    if ((rc = dl_view(h, labels, "labels", 1, 1, &vl))) return rc;
    if (vl.code != kDLInt || (vl.bits != 32 && vl.bits != 64)) return h->fail(FNB_ERR_INVALID, "labels must be int32 or int64");
    if (vl.rows != ve.rows) return h->fail(FNB_ERR_INVALID, "len(labels) != embeddings.shape[0]");
    const long long b = ve.rows;
    const int d = (int)ve.cols;
    if (b == 0) return FNB_OK;
    if (b > 8192) return h->fail(FNB_ERR_UNSUPPORTED, "mining batches are limited to 8192 rows (got %lld)", b);

    const void* de = nullptr; const void* dl = nullptr;
    CK(cudaEventRecord(h->ev[0], h->stream));
    if ((rc = dl_to_device(h, ve, (size_t)b * d * 4, h->stage_a, &de))) return rc;
    if ((rc = dl_to_device(h, vl, (size_t)b * (vl.bits / 8), h->stage_lab, &dl))) return rc;
    if ((rc = prepare_operand(h, opt.mode, (const float*)de, nullptr, b, d, false, op))) return rc;
    if ((rc = self_b_maps(h, op, d))) return rc;

    // stage 1: B x B distances (every ordered pair, diagonal included)
    const int cg = 1;                                   // 128 x 128 tiles: 225 tiles at B = 1800 fill the 148 SMs better than 64 pair-tiles
    const int tile = kRowsPerCta * cg;
    std::vector<RegionDev> regs;
    RegionDev r = {}; r.row_end = (int)b; r.col_end = (int)b; regs.push_back(r);
    finish_regions(regs, tile);
    if ((rc = upload_regions(h, regs))) return rc;
    if ((rc = reset_scalars(h))) return rc;
    const long long ld = ((b + 3) / 4) * 4;
    CK(h->strip.ensure((size_t)b * ld * 4));
    GramParams p = {};
    p.regions = h->regions.as<RegionDev>(); p.nregions = 1; p.total_tiles = regs.back().tile_begin;
    p.shard = ShardSpec{1, 0, 1, nullptr};
    p.kblocks = d / (128 / op.elem_bytes);
    p.acc_scale = 1.0f / (op.prescale * op.prescale);
    p.operand_fmt = op.fmt;
    DeviceScalars* sc = h->counters.as<DeviceScalars>();
    p.counters = sc->counters; p.range_ord = sc->range_ord;
    p.out = h->strip.as<float>(); p.out_ld = ld; p.tri_packed = 0; p.metric = 0;
    p.n_rows = (int)b; p.n_cols = (int)b;
    CK(cudaEventRecord(h->ev[1], h->stream));
    if ((rc = launch_gram(h, cg, EPI_ROWSTRIP, opt.max_ctas, op, p, 0))) return rc;

    // stage 2: per-anchor selection
    const size_t n_out = (size_t)b * 2 + (size_t)b * kmax * 3 + 1;
    CK(h->mine_out.ensure(n_out * 4));
    int* o_hp = h->mine_out.as<int>();
    int* o_hn = o_hp + b;
    int* o_pi = o_hn + b;
    int* o_sh = o_pi + (size_t)b * kmax;
    int* o_el = o_sh + (size_t)b * kmax;
    int* o_ovf = o_el + (size_t)b * kmax;
    CK(cudaMemsetAsync(o_ovf, 0, 4, h->stream));
    const size_t smem = (size_t)b * 4 + ((b + 15) / 16) * 16 + (size_t)(kmax > 0 ? kmax : 1) * 4;
    if (smem > 200 * 1024) return h->fail(FNB_ERR_UNSUPPORTED, "mining row does not fit shared memory");
    if (vl.bits == 64) {
        auto kern = mine_rows_kernel<long long>;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)b, kMineThreads, smem, h->stream>>>(h->strip.as<float>(), ld, (const long long*)dl, (int)b, alpha, kmax,
                                                            o_hp, o_hn, o_pi, o_sh, o_el, o_ovf);
    } else {
        auto kern = mine_rows_kernel<int>;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)b, kMineThreads, smem, h->stream>>>(h->strip.as<float>(), ld, (const int*)dl, (int)b, alpha, kmax,
                                                            o_hp, o_hn, o_pi, o_sh, o_el, o_ovf);
    }
    CK(cudaGetLastError());
    CK(cudaEventRecord(h->ev[2], h->stream));

    CK(h->pinned.ensure(n_out * 4 + 4096));
    int* host = reinterpret_cast<int*>((char*)h->pinned.p + 4096);
    CK(cudaMemcpyAsync(host, h->mine_out.p, n_out * 4, cudaMemcpyDeviceToHost, h->stream));
    DeviceScalars hs;
    CK(cudaMemcpyAsync(h->pinned.p, h->counters.p, sizeof(hs), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    memcpy(&hs, h->pinned.p, sizeof(hs));
    const float smin = ordered_to_float(hs.range_ord[0]), smax = ordered_to_float(hs.range_ord[1]);
    if (stats) {
        float ms = 0.f, pm = 0.f;
        cudaEventElapsedTime(&ms, h->ev[1], h->ev[2]);
        cudaEventElapsedTime(&pm, h->ev[0], h->ev[1]);
        stats->kernel_ms = ms; stats->prepare_ms = pm;
        stats->smin = smin; stats->smax = smax;
        stats->tiles = hs.counters[1];
        stats->n_pairs = (uint64_t)b * (uint64_t)b;
        stats->kernel_launches = 3;                     // split_rows, gram<ROWSTRIP>, mine_rows
    }
    const double lim = 1.0 + (double)opt.atol;
    if (b > 1 && ((double)smin < -lim || (double)smax > lim || smin != smin || smax != smax))
        return h->fail(FNB_ERR_NOT_NORMALIZED, "embeddings must be normalized to 1, range %.9g %.9g", smin, smax);
    if (host[n_out - 1] > kmax)
        return h->fail(FNB_ERR_INVALID, "kmax = %d is smaller than the largest class size - 1 = %d", kmax, host[n_out - 1]);
    memcpy(hardest_pos, host, (size_t)b * 4);
    memcpy(hardest_neg, host + b, (size_t)b * 4);
    if (kmax > 0) {
        memcpy(pos_index, host + 2 * b, (size_t)b * kmax * 4);
        memcpy(semi_hard, host + 2 * b + (size_t)b * kmax, (size_t)b * kmax * 4);
        memcpy(eligible, host + 2 * b + (size_t)b * kmax * 2, (size_t)b * kmax * 4);
    }
    return FNB_OK;
}
