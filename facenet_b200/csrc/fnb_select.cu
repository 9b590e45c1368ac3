// facenet_b200 -- threshold selection on the device: suffix scan over the keyed histogram bins,
// class-balanced confidence matrix in fp64, accuracy argmax and FAR-threshold interpolation.
//
// Replaces (reference, /root/reference/facenet/statistics.py):
//   :130-138  count / weight accumulation of tp, fn, fp, tn per threshold   -> confidence_kernel
//   :140-175  accuracy / tn_rates / fp_rates                                 -> select_kernel
//   :296      thresholds[np.argmax(accuracy)] (first maximum)                -> select_kernel
//   :299-302  far_threshold = interp1d(fp_rates, thresholds, 'slinear')(far) -> select_kernel
// Tiny, latency-bound kernels (<= 127 thresholds x a few thousand keys); no tensor-core work.
#include "fnb_host.h"

#include <cmath>
#include <cstring>

namespace fnb {

constexpr int kScanStride = kMaxBins + 2;      // suffix[k], k in [0, T+1], suffix[T+1] = 0

// one warp per (key, row): suffix[k] = sum_{k' >= k} bins[k']
__global__ void bins_suffix_scan_kernel(const unsigned long long* __restrict__ bins, int rows, int stride, int T,
                                        unsigned long long* __restrict__ suffix)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= rows) return;
    const unsigned long long* src = bins + (size_t)warp * stride;
    unsigned long long* dst = suffix + (size_t)warp * kScanStride;
    unsigned long long carry = 0;
    if (lane == 0) dst[T + 1] = 0;
    // walk the bins from the top in chunks of 32; lane l of a chunk handles bin (top - l)
    for (int top = T; top >= 0; top -= 32) {
        const int k = top - lane;
        unsigned long long v = (k >= 0) ? src[k] : 0ull;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long u = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += u;
        }
        v += carry;
        if (k >= 0) dst[k] = v;
        carry = __shfl_sync(0xffffffffu, v, 31);
    }
}

// block n (one per threshold): rates[0..3][n] = tp, tn, fp, fn  (statistics.py:133-138 with the
// class-pair blocks grouped by key; fp64, fixed summation order: thread-strided partial sums, then a tree)
__global__ void confidence_kernel(const unsigned long long* __restrict__ suffix, int nkeys, int T,
                                  const int* __restrict__ pos, const double* __restrict__ w_same,
                                  const double* __restrict__ w_diff, double* __restrict__ rates)
{
    __shared__ double red[4][128];
    const int n = blockIdx.x;
    const int p = pos[n];
    double tp = 0, tn = 0, fp = 0, fn = 0;
    for (int key = threadIdx.x; key < nkeys; key += blockDim.x) {
        const unsigned long long* all = suffix + ((size_t)key * 2 + 0) * kScanStride;
        const unsigned long long* same = all + kScanStride;
        const unsigned long long all_lt = all[p], same_lt = same[p];
        const unsigned long long all_tot = all[0], same_tot = same[0];
        const double ws = w_same[key], wd = w_diff[key];
        tp = __dadd_rn(tp, __dmul_rn((double)same_lt, ws));
        fn = __dadd_rn(fn, __dmul_rn((double)(same_tot - same_lt), ws));
        fp = __dadd_rn(fp, __dmul_rn((double)(all_lt - same_lt), wd));
        tn = __dadd_rn(tn, __dmul_rn((double)((all_tot - same_tot) - (all_lt - same_lt)), wd));
    }
    red[0][threadIdx.x] = tp; red[1][threadIdx.x] = tn; red[2][threadIdx.x] = fp; red[3][threadIdx.x] = fn;
    __syncthreads();
    for (int s = 64; s >= 1; s >>= 1) {
        if ((int)threadIdx.x < s) {
#pragma unroll
            for (int q = 0; q < 4; ++q) red[q][threadIdx.x] = __dadd_rn(red[q][threadIdx.x], red[q][threadIdx.x + s]);
        }
        __syncthreads();
    }
    if (threadIdx.x < 4) rates[(size_t)threadIdx.x * T + n] = red[threadIdx.x][0];
}

// one block of kMaxBins threads: thread n forms accuracy[n] and fp_rate[n]; accuracy argmax (FIRST maximum; NaN counts as
// maximum like np.argmax) by a packed-key block reduction; the FAR threshold by thread 0 over the shared fp_rates:
//   j = last sample (in threshold order) with fp_rate <= target, then the k = 1 B-spline weights of scipy 1.4.1's
//   interp1d(kind='slinear') (statistics.py:299-302):  w = 1 / (x1 - x0);  thr[j] ((x1 - t) w) + thr[j+1] ((t - x0) w).
// (When the left sample belongs to a RUN of equal fp_rates the reference's unstable argsort decides which tied threshold is
// used: the Python host replays that case, facenet_b200/statistics.py:_slinear.)
// sel[0] = argmax index; far[0] = far threshold (0 if max(fp_rates) < far_target; NaN if outside the range)
__global__ void __launch_bounds__(kMaxBins)
select_kernel(const double* __restrict__ rates, int T, const double* __restrict__ thr, double far_target,
              int* __restrict__ sel, double* __restrict__ far)
{
    __shared__ double s_fpr[kMaxBins];
    __shared__ unsigned long long s_key[kMaxBins / 32];
    __shared__ double s_max[kMaxBins / 32];
    const int n = threadIdx.x;
    const double* tp = rates; const double* tn = rates + T; const double* fp = rates + 2 * T; const double* fn = rates + 3 * T;
    // key: larger accuracy wins, NaN beats everything, ties -> lowest index
    unsigned long long key = 0ull;
    double fpr = -INFINITY;
    if (n < T) {
        const double num = __dadd_rn(tp[n], tn[n]);
        const double den = __dadd_rn(__dadd_rn(__dadd_rn(tp[n], fp[n]), tn[n]), fn[n]);
        const double acc = __ddiv_rn(num, den);
        unsigned long long bits = (unsigned long long)__double_as_longlong(acc);
        bits = (bits & 0x8000000000000000ull) ? ~bits : (bits | 0x8000000000000000ull);      // order-preserving for doubles
        if (acc != acc) bits = ~0ull;
        // 57 bits of order + 7 bits of (127 - n) would lose accuracy bits: reduce (value, index) pairs instead
        key = bits;
        const double d = __dadd_rn(tn[n], fp[n]);
        const double tnr = d > 0 ? __ddiv_rn(tn[n], d) : 1.0;
        fpr = __dsub_rn(1.0, tnr);
        s_fpr[n] = fpr;
    }
    // block argmax of (key, lowest index) and max of fpr
    int idx = (n < T) ? n : 0x7fffffff;
    double mx = fpr;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const unsigned long long ok = __shfl_xor_sync(0xffffffffu, key, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        const double om = __shfl_xor_sync(0xffffffffu, mx, o);
        if (oi != 0x7fffffff && (idx == 0x7fffffff || ok > key || (ok == key && oi < idx))) { key = ok; idx = oi; }
        mx = fmax(mx, om);
    }
    __shared__ int s_idx[kMaxBins / 32];
    if ((n & 31) == 0) { s_key[n >> 5] = key; s_idx[n >> 5] = idx; s_max[n >> 5] = mx; }
    __syncthreads();
    if (n != 0) return;
    for (int w = 1; w < kMaxBins / 32; ++w) {
        if (s_idx[w] != 0x7fffffff && (idx == 0x7fffffff || s_key[w] > key || (s_key[w] == key && s_idx[w] < idx))) { key = s_key[w]; idx = s_idx[w]; }
        mx = fmax(mx, s_max[w]);
    }
    sel[0] = idx;
    double out = 0.0;
    if (mx >= far_target) {
        if (T < 2 || far_target < s_fpr[0] || far_target > s_fpr[T - 1]) {
            out = __longlong_as_double(0x7ff8000000000000LL);
        } else {
            int lo = 0, hi = T;                       // np.searchsorted(x, xq, side='right')
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (s_fpr[mid] <= far_target) lo = mid + 1; else hi = mid;
            }
            int j = lo - 1;
            if (j < 0) j = 0;
            if (j > T - 2) j = T - 2;
            const double x0 = s_fpr[j], x1 = s_fpr[j + 1];
            const double w = __ddiv_rn(1.0, __dsub_rn(x1, x0));
            out = __dadd_rn(__dmul_rn(thr[j], __dmul_rn(__dsub_rn(x1, far_target), w)), __dmul_rn(thr[j + 1], __dmul_rn(__dsub_rn(far_target, x0), w)));
        }
    }
    far[0] = out;
}

}  // namespace fnb

using namespace fnb;

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return h->fail(FNB_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); } while (0)

extern "C" int fnb_confidence_from_last_bins(fnb_handle h, int nkeys, const double* w_same, const double* w_diff,
                                             const double* thresholds, int T, const fnb_options* opt_in, double far_target,
                                             double* tp, double* tn, double* fp, double* fn,
                                             int32_t* argmax_accuracy, double* far_threshold)
{
    if (!h) return FNB_ERR_INVALID;
    fnb_options opt; if (opt_in) opt = *opt_in; else fnb_default_options(&opt);
    if (opt.metric != 0 && opt.metric != 1) return h->fail(FNB_ERR_BAD_METRIC, "Undefined similarity metric %d", opt.metric);
    if (!w_same || !w_diff || !thresholds || !tp || !tn || !fp || !fn) return h->fail(FNB_ERR_INVALID, "NULL argument");
    if (T < 1 || T >= kMaxBins) return h->fail(FNB_ERR_INVALID, "number of thresholds must be in [1, %d]", kMaxBins - 1);
    if (nkeys < 1 || nkeys != h->last_nkeys || T != h->last_T)
        return h->fail(FNB_ERR_INVALID, "no matching histogram on the device (have nkeys=%d T=%d, asked nkeys=%d T=%d)",
                       h->last_nkeys, h->last_T, nkeys, T);
    CK(cudaSetDevice(h->device));
    CutTables ct;
    if (build_cut_tables(thresholds, T, opt.metric, opt.eps, opt.cuts, &ct)) return h->fail(FNB_ERR_INVALID, "bad thresholds");

    // staging layout (doubles first, ints last): w_same[nkeys] w_diff[nkeys] thr[T] | pos[T]
    const size_t nd = (size_t)2 * nkeys + T;
    const size_t in_bytes = nd * 8 + (size_t)T * 4;
    const size_t out_doubles = (size_t)4 * T + 1;            // rates + far
    const size_t out_bytes = out_doubles * 8 + 8;             // + argmax
    CK(h->pinned.ensure(in_bytes + out_bytes + 64));
    CK(h->select_io.ensure(in_bytes + out_bytes + 64));
    CK(h->scan.ensure((size_t)nkeys * 2 * kScanStride * 8));
    double* hin = h->pinned.as<double>();
    memcpy(hin, w_same, (size_t)nkeys * 8);
    memcpy(hin + nkeys, w_diff, (size_t)nkeys * 8);
    memcpy(hin + 2 * nkeys, thresholds, (size_t)T * 8);
    memcpy(hin + nd, ct.pos, (size_t)T * 4);
    CK(cudaMemcpyAsync(h->select_io.p, hin, in_bytes, cudaMemcpyHostToDevice, h->stream));
    double* d_in = h->select_io.as<double>();
    const int* d_pos = reinterpret_cast<const int*>(d_in + nd);
    const size_t out_off = (in_bytes + 63) / 64 * 64;
    double* d_rates = reinterpret_cast<double*>((char*)h->select_io.p + out_off);
    double* d_far = d_rates + 4 * T;
    int* d_sel = reinterpret_cast<int*>(d_far + 1);

    const int rows = nkeys * 2;
    const int stride = kMaxBins + 1;                           // layout written by run_hist
    bins_suffix_scan_kernel<<<(rows * 32 + 255) / 256, 256, 0, h->stream>>>(h->bins.as<unsigned long long>(), rows, stride, T,
                                                                            h->scan.as<unsigned long long>());
    CK(cudaGetLastError());
    confidence_kernel<<<T, 128, 0, h->stream>>>(h->scan.as<unsigned long long>(), nkeys, T, d_pos, d_in, d_in + nkeys, d_rates);
    CK(cudaGetLastError());
    select_kernel<<<1, kMaxBins, 0, h->stream>>>(d_rates, T, d_in + 2 * nkeys, far_target, d_sel, d_far);
    CK(cudaGetLastError());
    char* hout = (char*)h->pinned.p + out_off;
    CK(cudaMemcpyAsync(hout, d_rates, out_bytes, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    const double* r = reinterpret_cast<const double*>(hout);
    memcpy(tp, r, (size_t)T * 8);
    memcpy(tn, r + T, (size_t)T * 8);
    memcpy(fp, r + 2 * T, (size_t)T * 8);
    memcpy(fn, r + 3 * T, (size_t)T * 8);
    if (far_threshold) *far_threshold = r[4 * T];
    if (argmax_accuracy) memcpy(argmax_accuracy, r + 4 * T + 1, 4);
    return FNB_OK;
}
