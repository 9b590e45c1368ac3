// facenet_b200 -- C-ABI entry points (include/facenet_b200.h): argument checking, DLPack
// ingestion, workspace management, similarity-cut tables, region schedules and kernel launches.
#include "fnb_host.h"
#include "fnb_bias.h"

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <limits>

using namespace fnb;

static thread_local std::string g_create_error;

int fnb_context::fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    err = buf;
    return code;
}

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return h->fail(FNB_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); } while (0)

// ---------------------------------------------------------------------------------------
// modes

namespace fnb {

int mode_info(int mode, int* num_pass, bool* tf32, int* fmt, int* elem_bytes, float* prescale) {
    switch (mode) {
        case FNB_MODE_FP16X3: *num_pass = 3; *tf32 = false; *fmt = kFmtF16;  *elem_bytes = 2; *prescale = 256.f; return 0;
        case FNB_MODE_TF32X3: *num_pass = 3; *tf32 = true;  *fmt = kFmtTF32; *elem_bytes = 4; *prescale = 1.f;   return 0;
        case FNB_MODE_TF32:   *num_pass = 1; *tf32 = true;  *fmt = kFmtTF32; *elem_bytes = 4; *prescale = 1.f;   return 0;
        case FNB_MODE_BF16:   *num_pass = 1; *tf32 = false; *fmt = kFmtBF16; *elem_bytes = 2; *prescale = 1.f;   return 0;
        case FNB_MODE_FP16:   *num_pass = 1; *tf32 = false; *fmt = kFmtF16;  *elem_bytes = 2; *prescale = 1.f;   return 0;
        case FNB_MODE_FP16F8: case FNB_MODE_AUTO:
                              *num_pass = 2; *tf32 = false; *fmt = kFmtF16;  *elem_bytes = 2; *prescale = 4096.f; return 0;
    }
    return -1;
}

// beta(|s|) knots of `mode` at dimension d (fnb_bias.h); zeros for the modes without a calibration
void bias_table(int mode, int d, float* knots /* [kBiasStride] */) {
    for (int i = 0; i < kBiasStride; ++i) knots[i] = 0.f;
    const float* src = nullptr;
    double f = std::pow((double)d / 512.0, 1.09);
    switch (mode) {
        case FNB_MODE_FP16X3: src = kBiasBetaX3; break;
        case FNB_MODE_TF32X3: src = kBiasBetaX3; f *= 1.50; break;     // same 96 steps, K = 8 per MMA: 4.57e-6 vs 3.05e-6 (profiles/r02a_bias_raw_tf32x3.log)
        case FNB_MODE_FP16F8: case FNB_MODE_AUTO: src = kBiasBetaF8; break;
        default: return;
    }
    for (int i = 0; i < kBiasKnots; ++i) knots[i] = (float)(src[i] * f);
    for (int i = kBiasKnots; i < kBiasStride; ++i) knots[i] = knots[kBiasKnots - 1];
}

static double beta_at(const float* knots, double a) {
    a = std::min(std::fabs(a), 1.0) * (kBiasKnots - 1);
    int i = std::min((int)a, kBiasKnots - 2);
    const double f = a - i;
    return knots[i] + f * (knots[i + 1] - knots[i]);
}

// Error model of one pair's similarity in `mode` AFTER the bias correction (profiles/r02a_bias_fine.log, D = 512, dense rows):
// spread of the truncating accumulation (fp16x3: 0.05e-6 .. 0.32e-6, ~ D^0.8) and, for fp16f8, of the 2 D e4m3 roundings of
// the cross terms (0.46e-6 .. 0.83e-6 for Gaussian rows, ~ sqrt(sum x_i^2 y_i^2) <= sqrt(peakedness)); the single-pass modes
// carry the rounding of the operands themselves.  Returns sigma(s).
double mode_sigma_s(int mode, int d, double abs_s, double peakedness) {
    const double trunc = (0.05e-6 + 0.27e-6 * abs_s) * std::pow((double)d / 512.0, 0.8);
    switch (mode) {
        case FNB_MODE_FP16X3: return trunc;
        case FNB_MODE_TF32X3: return 1.5 * trunc;
        case FNB_MODE_FP16F8: case FNB_MODE_AUTO: {
            const double pk = peakedness > 0 ? peakedness : 3.0 / d;
            const double e4m3 = (0.455e-6 + 0.375e-6 * abs_s) * std::sqrt(pk / (3.0 / 512.0));
            return std::sqrt(e4m3 * e4m3 + trunc * trunc / 3.0);
        }
        case FNB_MODE_BF16: return 1.5e-4;
        default: return 1.75e-5;                          // tf32 / fp16 single pass
    }
}

// ---------------------------------------------------------------------------------------
// similarity cuts.  dist32() restates statistics.py:45-53 on one fp32 value.

static inline float dist32(float s, int metric) {
    if (s < -1.f) s = -1.f;
    if (s > 1.f) s = 1.f;
    if (metric == 0) { volatile float one_minus = 1.0f - s; return 2.0f * one_minus; }
    return acosf(s);
}

// smallest fp32 similarity in [-1, 1] whose distance is < t (float64 compare, statistics.py:131); +inf if none
static float cut_for_threshold(double t, int metric) {
    auto pred = [&](float s) { return (double)dist32(s, metric) < t; };
    if (!pred(1.0f)) return std::numeric_limits<float>::infinity();
    if (pred(-1.0f)) return -1.0f;
    uint32_t lo = float_to_ordered(-1.0f), hi = float_to_ordered(1.0f);   // pred(lo) false, pred(hi) true
    while (hi - lo > 1) {
        const uint32_t mid = lo + (hi - lo) / 2;
        if (pred(ordered_to_float(mid))) hi = mid; else lo = mid;
    }
    return ordered_to_float(hi);
}

int build_cut_tables(const double* thresholds, int T, int metric, double eps, const float* cuts_override, CutTables* out,
                     const float* beta_knots, const float* beta_knots_strict) {
    if (T < 1 || T >= kMaxBins) return -1;
    CutTables& c = *out;
    c.T = T;
    std::vector<float> cut(T);
    for (int n = 0; n < T; ++n) cut[n] = cuts_override ? cuts_override[n] : cut_for_threshold(thresholds[n], metric);
    for (int n = 0; n < T; ++n) c.order[n] = n;
    std::stable_sort(c.order, c.order + T, [&](int a, int b) { return cut[a] < cut[b]; });
    const float inf = std::numeric_limits<float>::infinity();
    for (int j = 0; j < kMaxBins; ++j) c.cuts[j] = inf;
    for (int j = 0; j < kMaxBins + 4; ++j) { c.wlo[j] = inf; c.whi[j] = -inf; }
    c.T_fin = 0;
    for (int j = 0; j < T; ++j) {
        const int n = c.order[j];
        c.cuts[j] = cut[n];
        if (std::isfinite(cut[n])) c.T_fin = j + 1;
        const double t = thresholds[n];
        double lo_s, hi_s;
        if (metric == 0) { lo_s = 1.0 - (t + eps) / 2.0; hi_s = 1.0 - (t - eps) / 2.0; }
        else { lo_s = std::cos(std::min(t + eps, M_PI)); hi_s = std::cos(std::max(t - eps, 0.0)); if (t - eps > M_PI) hi_s = -2.0; if (t + eps < 0) lo_s = 2.0; }
        c.wlo[j] = (float)lo_s;
        c.whi[j + 1] = (float)hi_s;
    }
    for (int n = 0; n < T; ++n)
        c.pos[n] = (int)(std::upper_bound(c.cuts, c.cuts + T, cut[n]) - c.cuts);
    // arithmetic-progression fit of the finite cuts (true for np.linspace thresholds with metric 0) in RAW similarity: the
    // interior tiles bin the uncorrected accumulator, s_raw = s (1 - beta(|s|)) (fnb_bias.h), so the progression is fitted to
    // the raw positions of the cuts and the curvature of beta goes into `dev`, i.e. into the counted near-threshold window
    c.uniform = 0;
    if (c.T_fin >= 2) {
        std::vector<double> raw(c.T_fin);
        for (int j = 0; j < c.T_fin; ++j) {
            const double cj = (double)c.cuts[j];
            raw[j] = beta_knots ? cj * (1.0 - beta_at(beta_knots, cj)) : cj;
        }
        c.e0 = raw[0];
        c.h = (raw[c.T_fin - 1] - raw[0]) / (c.T_fin - 1);
        double dev = 0;
        for (int j = 0; j < c.T_fin; ++j) dev = std::max(dev, std::fabs(raw[j] - (c.e0 + j * c.h)));
        c.dev = dev;
        if (c.h > 1e-4 && dev / c.h < 2e-4) c.uniform = 1;
        // the same fit for the fp16x3 arithmetic of the strict tiles of an fp16f8 launch
        c.e0x = c.e0; c.hx = c.h; c.devx = c.dev;
        if (beta_knots_strict) {
            for (int j = 0; j < c.T_fin; ++j) { const double cj = (double)c.cuts[j]; raw[j] = cj * (1.0 - beta_at(beta_knots_strict, cj)); }
            c.e0x = raw[0];
            c.hx = (raw[c.T_fin - 1] - raw[0]) / (c.T_fin - 1);
            double devx = 0;
            for (int j = 0; j < c.T_fin; ++j) devx = std::max(devx, std::fabs(raw[j] - (c.e0x + j * c.hx)));
            c.devx = devx;
        }
    }
    return 0;
}

}  // namespace fnb

// ---------------------------------------------------------------------------------------
// tensors

int fnb::dl_view(fnb_context* h, const DLTensor* t, const char* name, int want_ndim_min, int want_ndim_max, DLView* v) {
    if (!t) return h->fail(FNB_ERR_INVALID, "%s: NULL tensor", name);
    if (t->ndim < want_ndim_min || t->ndim > want_ndim_max) return h->fail(FNB_ERR_INVALID, "%s: ndim %d not supported", name, t->ndim);
    if (t->dtype.lanes != 1) return h->fail(FNB_ERR_INVALID, "%s: vector dtypes not supported", name);
    long long expect = 1;
    for (int i = t->ndim - 1; i >= 0; --i) {
        if (t->strides && t->shape[i] > 1 && t->strides[i] != expect) return h->fail(FNB_ERR_INVALID, "%s: must be C-contiguous", name);
        expect *= t->shape[i];
    }
    v->rows = t->ndim >= 1 ? t->shape[0] : 1;
    v->cols = t->ndim >= 2 ? t->shape[1] : 1;
    v->bits = t->dtype.bits;
    v->code = t->dtype.code;
    v->data = (char*)t->data + t->byte_offset;
    if (t->device.device_type == kDLCUDA) {
        if (t->device.device_id != h->device) return h->fail(FNB_ERR_INVALID, "%s: tensor on cuda:%d, handle on cuda:%d", name, t->device.device_id, h->device);
        v->on_device = true;
    } else if (t->device.device_type == kDLCPU || t->device.device_type == kDLCUDAHost) {
        v->on_device = false;
    } else {
        return h->fail(FNB_ERR_INVALID, "%s: unsupported DLPack device type %d", name, t->device.device_type);
    }
    return FNB_OK;
}

// device pointer to the tensor's bytes (staged through `stage` when the tensor lives on the host)
int fnb::dl_to_device(fnb_context* h, const DLView& v, size_t bytes, DevBuf& stage, const void** out) {
    if (v.on_device || bytes == 0) { *out = v.data; return FNB_OK; }
    CK(stage.ensure(bytes));
    int rc = stage_to_device(h, stage.p, v.data, bytes);       // pinned ring + copy threads for pageable memory (fnb_stage.cu)
    if (rc) return rc;
    *out = stage.p;
    return FNB_OK;
}

constexpr int kFmtU8 = 100;          // byte arrays (e4m3 operands): 128 elements per 128-byte box row

static int make_tmap(fnb_context* h, CUtensorMap* m, void* base, int fmt, long long rows, int d, int box_rows = kRowsPerCta) {
    CUtensorMapDataType dt = fmt == kFmtTF32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                           : fmt == kFmtBF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                           : fmt == kFmtU8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    const int eb = fmt == kFmtTF32 ? 4 : fmt == kFmtU8 ? 1 : 2;
    cuuint64_t gdim[2] = {(cuuint64_t)d, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)d * eb};
    cuuint32_t box[2] = {(cuuint32_t)(128 / eb), (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = h->encode(m, dt, 2, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return h->fail(FNB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%lld d=%d", (int)r, rows, d);
    return FNB_OK;
}

// FNB_MODE_AUTO, a-priori gate: largest peakedness of any row FP16F8 is tried for.  Dense Gaussian-like rows have 3 / D on
// average (512-d: 0.0059, the largest of 1M rows ~0.010); 256-d rows (mean 0.0117, largest of a few thousand > 1/64; measured
// max |dd| 1.3e-5 in fp16f8, profiles/r02a_bias_d256.log) and sparse / heavy-tailed rows do not pass.  The a-posteriori
// error bound (error_certificate) decides whether the result of the fast pass is kept.
constexpr float kAutoPeakLimit = 1.0f / 64.0f;

static long long pad_rows(long long n) { return ((n + 255) / 256) * 256 + 256; }

// ---------------------------------------------------------------------------------------
// regions

// tile grid of every region: `tile` rows x (`tile` * pairs) columns per scheduler step (see TileScheduler)
int fnb::shard_from_options(fnb_context* h, const fnb_options& opt, ShardHost* out) {
    if (opt.world < 1 || opt.rank < 0 || opt.rank >= opt.world) return h->fail(FNB_ERR_INVALID, "bad rank/world %d/%d", opt.rank, opt.world);
    out->residues.clear();
    if (opt.shard_mod == 0) {
        out->spec = ShardSpec{opt.world, opt.rank, 1, nullptr};
        out->residues.push_back(opt.rank);
        return FNB_OK;
    }
    const int mod = opt.shard_mod, width = opt.shard_width;
    if (mod < 1 || width < 0 || width > mod) return h->fail(FNB_ERR_INVALID, "bad shard width %d of %d", width, mod);
    if (opt.shard_slots) {
        for (int i = 0; i < width; ++i) {
            const int r = opt.shard_slots[i];
            if (r < 0 || r >= mod || (i > 0 && r <= opt.shard_slots[i - 1])) return h->fail(FNB_ERR_INVALID, "shard_slots must be ascending residues in [0, %d)", mod);
            out->residues.push_back(r);
        }
        CK(h->shard_slots.ensure((size_t)std::max(width, 1) * 4));
        if (width) CK(cudaMemcpyAsync(h->shard_slots.p, out->residues.data(), (size_t)width * 4, cudaMemcpyHostToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));            // `residues` is the caller's stack object
        out->spec = ShardSpec{mod, 0, width, h->shard_slots.as<int32_t>()};
    } else {
        if (opt.shard_lo < 0 || opt.shard_lo + width > mod) return h->fail(FNB_ERR_INVALID, "bad shard range [%d, %d + %d) mod %d", opt.shard_lo, opt.shard_lo, width, mod);
        for (int i = 0; i < width; ++i) out->residues.push_back(opt.shard_lo + i);
        out->spec = ShardSpec{mod, opt.shard_lo, width, nullptr};
    }
    return FNB_OK;
}

void fnb::finish_regions(std::vector<RegionDev>& regs, int tile, int pairs, const ShardHost* shard, bool pad_equal) {
    const ShardHost whole;
    if (!shard) shard = &whole;
    int cp = 0;
    long long t = 0;
    // pair grid of a cluster: 1 x 1, 1 x 2 (pairs == 2) or 2 x 2 (pairs == 4) tiles per scheduler step
    const int super_rows = tile * (pairs == 4 ? 2 : 1);
    const int super_cols = tile * (pairs == 1 ? 1 : 2);
    for (auto& r : regs) {
        r.nrb = (r.row_end - r.row_begin + super_rows - 1) / super_rows;
        r.ncb = (r.col_end - r.col_begin + super_cols - 1) / super_cols;
        r.own_cnt = shard->spec.width > 0 ? shard->owned(r.nrb) : 0;
        // ranks that share their tile queues walk schedules of identical shape: everybody ceil(nrb / world) row blocks (the kernel skips
        // the blocks past the region's end)
        if (pad_equal && shard->spec.mod > 1) r.own_cnt = (r.nrb + shard->spec.mod - 1) / shard->spec.mod;
        r.cp_begin = cp;
        cp += r.ncb;
        r.tile_begin = t;
        t += (long long)r.own_cnt * r.ncb;
    }
    RegionDev sentinel = {};
    sentinel.tile_begin = t;
    regs.push_back(sentinel);
}

// strict upper triangle of an n x n pair matrix as super-rows of `rr` rows: one diagonal square
// (tri) plus one rectangle to its right per super-row -> L2-friendly tile order
static void triangle_regions(long long n, int rr, int key, std::vector<RegionDev>& regs) {
    for (long long r0 = 0; r0 < n; r0 += rr) {
        const long long r1 = std::min<long long>(n, r0 + rr);
        RegionDev d = {};
        d.row_begin = (int)r0; d.row_end = (int)r1; d.col_begin = (int)r0; d.col_end = (int)r1; d.tri = 1; d.key = key;
        regs.push_back(d);
        if (r1 < n) {
            RegionDev f = {};
            f.row_begin = (int)r0; f.row_end = (int)r1; f.col_begin = (int)r1; f.col_end = (int)n; f.tri = 0; f.key = key;
            regs.push_back(f);
        }
    }
}

static int pick_cta_group(const fnb_options* o) { return (o->cta_group == 1 || o->cta_group == 2) ? o->cta_group : 2; }

// CTA pairs per cluster of the histogram launches (multicast of the A operand): only with CTA pairs.
// Auto (cluster_pairs == 0): a launch long enough to run into the board's power limit (measured: from ~5e10 pairs per
// rank, profiles/r01c_tile_order_and_power.md) is bound by energy per pair, and reading the A operand from L2 once per
// two tiles saves more than the 16 SMs that cannot host a 4-CTA cluster cost; shorter launches run at full clock,
// are L2->SM bandwidth bound, and are faster on all 148 SMs.
static int pick_pairs(const fnb_options* o, int cta_group, long long n = 0) {
    if (cta_group != 2) return 1;
    if (o->cluster_pairs == 1 || o->cluster_pairs == 2 || o->cluster_pairs == 4) return o->cluster_pairs;
    const double local_pairs = 0.5 * (double)n * (double)(n - 1) / (double)std::max(1, o->world);
    return local_pairs >= 5.0e10 ? 2 : 1;
}

// Rows per super-row of the tile order (row blocks fastest inside a super-row).  The CTAs working at any time then
// share ONE column panel, and the super-row's row panels (rr x d x 4 bytes of split operands) are re-read from L2 for
// every column panel: auto sizes them to 64 MiB, half of the 126 MB L2 (32768 rows at d = 512; measured 2048 -> 32768:
// +13 % at 1M rows, HBM reads fall from 490 GB to 30 GB per pass).  Ranks interleave tiles (t % world), so each rank
// touches 1/world of a super-row's row panels: scale by world.  Small sets keep >= 6 super-rows for balance.
static int pick_region_rows(const fnb_options* o, int tile, long long n = 0, int d = 512, int clusters = 0, int row_block = 0) {
    long long rr = o->region_rows;
    const int world = std::max(1, o->world);
    // Tile queue (the default): a row block is read by whichever cluster asks for the tile, i.e. from both dies, and is then cached
    // in both L2 partitions -- the row panels get HALF the budget (32 MiB: 16,384 rows at d = 512; measured at 1M, queue + auxiliary
    // pairs: 8,192 ... 25,344 rows 653-666 G pairs/s, 33,792 rows 633-642, 50,688 rows 585; profiles/r02q_queue_region_rows.log),
    // and nothing ties row blocks to clusters, so the cluster alignment below does not apply.
    const bool queue = o->tile_queue >= 0 && o->panel_window <= 0;
    if (rr <= 0) {
        rr = ((queue ? 32ll : 64ll) << 20) / (4ll * std::max(d, 1)) * world;
        if (n > 0) rr = std::min(rr, std::max<long long>(n / 6, 8ll * tile));
        // row blocks per super-row divisible by world: rank r then owns the SAME row blocks (r, r + world, ...) in every
        // column panel, i.e. 1/world of the row panels; otherwise its rows drift from column to column and it ends up
        // streaming all of them (measured at 8 GPUs: 114 ms per rank instead of 1/8 of the single-GPU 817 ms)
        long long q = (long long)tile * (o->shard_mod > 0 ? o->shard_mod : world);
        // ... and this rank's row blocks per super-row a multiple of the number of clusters: every cluster then gets the same
        // number of tiles in every column panel -- no cluster waits at the progress window for one that has an extra tile, and
        // each cluster re-reads the same row panels (measured at 1M, 33 clusters of two pairs: 132 row blocks per super-row
        // 849-854 ms and 64-66 GB of DRAM reads, 96 row blocks 890-894 ms and 78-85 GB; profiles/r02b_rr_window_dram.log)
        if (clusters > 0 && row_block > 0 && !queue) {
            long long unit = (long long)clusters * row_block * (o->shard_mod > 0 ? o->shard_mod : world);
            while (unit % tile) unit *= 2;
            if (rr >= unit) q = unit;
            if (rr >= unit && rr % unit >= unit / 2 && (n <= 0 || rr + unit <= std::max<long long>(n / 3, unit))) rr += unit;   // round to nearest
        }
        if (rr >= q) rr = (rr / q) * q;
    }
    rr = std::max<long long>(tile, (rr / tile) * tile);
    return (int)std::min<long long>(rr, 1ll << 30);
}

// ---------------------------------------------------------------------------------------
// lifecycle

extern "C" int fnb_version(void) { return 100; }

extern "C" void fnb_default_options(fnb_options* o) {
    if (!o) return;
    memset(o, 0, sizeof(*o));
    o->mode = FNB_MODE_FP16X3;
    o->metric = 0;
    o->atol = 1.e-5f;
    o->eps = 1.e-5f;
    o->rank = 0;
    o->world = 1;
}

extern "C" int fnb_create(int device, fnb_handle* out) {
    if (!out) { g_create_error = "fnb_create: NULL out"; return FNB_ERR_INVALID; }
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_create_error = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                         " (facenet_b200 has no CPU fallback)";
        return FNB_ERR_CUDA;
    }
    if (device < 0 || device >= count) { g_create_error = "fnb_create: bad device index"; return FNB_ERR_INVALID; }
    cudaDeviceProp prop;
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
        g_create_error = std::string("cudaSetDevice/GetDeviceProperties: ") + cudaGetErrorString(e);
        return FNB_ERR_CUDA;
    }
    if (prop.major != 10) {
        char buf[160];
        snprintf(buf, sizeof(buf), "device %d is sm_%d%d; facenet_b200 kernels are built for sm_100a only", device, prop.major, prop.minor);
        g_create_error = buf;
        return FNB_ERR_UNSUPPORTED;
    }
    fnb_context* h = new fnb_context();
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    h->cc_major = prop.major; h->cc_minor = prop.minor;
    h->total_mem = prop.totalGlobalMem;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || !fn) { g_create_error = "cuTensorMapEncodeTiled entry point not found"; delete h; return FNB_ERR_CUDA; }
    h->encode = (PFN_tmapEncodeTiled)fn;
    if ((e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking)) != cudaSuccess) {
        g_create_error = std::string("cudaStreamCreate: ") + cudaGetErrorString(e); delete h; return FNB_ERR_CUDA;
    }
    h->stream = h->own_stream;
    for (int i = 0; i < 4; ++i) cudaEventCreate(&h->ev[i]);
    for (int i = 0; i < 3; ++i) cudaEventCreate(&h->copy_ev[i]);
    *out = h;
    return FNB_OK;
}

extern "C" void fnb_destroy(fnb_handle h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    if (h->aux_stream) { cudaStreamSynchronize(h->aux_stream); cudaStreamDestroy(h->aux_stream); }
    for (int i = 0; i < 2; ++i) if (h->aux_ev[i]) cudaEventDestroy(h->aux_ev[i]);
    h->tile_counter.release();
    DevBuf* bufs[] = {&h->stage_a, &h->stage_b, &h->stage_lab, &h->a_hi, &h->a_lo, &h->b_hi, &h->b_lo, &h->a_h8, &h->b_h8, &h->a_l16, &h->b_l16, &h->bias_tab, &h->strict_bits, &h->a_nrm, &h->b_nrm, &h->shard_slots, &h->progress, &h->perm, &h->cls,
                      &h->keys_in, &h->keys_out, &h->vals_in, &h->flags, &h->cub_tmp, &h->regions, &h->tables, &h->bins,
                      &h->counters, &h->out, &h->strip, &h->mine_out, &h->scan, &h->select_io, &h->mine_lab, &h->mine_keys, &h->mine_status};
    for (DevBuf* b : bufs) b->release();
    h->pinned.release();
    h->perm_host.release();
    comm_release(h);
    h->comm_buf.release(); h->lab_all.release(); h->local_perm.release();
    if (h->copier) { destroy_copier(h->copier); h->copier = nullptr; }
    if (h->copy_stream) { cudaStreamSynchronize(h->copy_stream); cudaStreamDestroy(h->copy_stream); }
    h->ring.release();
    for (int i = 0; i < 4; ++i) if (h->ring_ev[i]) cudaEventDestroy(h->ring_ev[i]);
    for (int i = 0; i < 3; ++i) if (h->copy_ev[i]) cudaEventDestroy(h->copy_ev[i]);
    for (int i = 0; i < 4; ++i) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
}

extern "C" int fnb_set_stream(fnb_handle h, void* cuda_stream) {
    if (!h) return FNB_ERR_INVALID;
    cudaSetDevice(h->device);
    // NULL is CUDA's legacy default stream (what torch reports as its current stream until another one is selected):
    // it must NOT fall back to the handle's own non-blocking stream, which would not be ordered after the caller's
    // kernels and collectives on stream 0
    cudaStream_t next = (cuda_stream == FNB_STREAM_OWN) ? h->own_stream : (cuda_stream ? (cudaStream_t)cuda_stream : cudaStreamLegacy);
    if (next != h->stream) {
        // no host synchronisation: the new stream is ordered after the handle's work on the old one (workspaces are shared)
        if (cudaEventRecord(h->ev[3], h->stream) == cudaSuccess) cudaStreamWaitEvent(next, h->ev[3], 0);
        else cudaGetLastError();                         // the old stream is gone (destroyed by its owner): nothing to order after
        h->stream = next;
    }
    return FNB_OK;
}

extern "C" const char* fnb_last_error(fnb_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

extern "C" int fnb_device_info(fnb_handle h, int* sm_count, int* cc_major, int* cc_minor, uint64_t* total_mem) {
    if (!h) return FNB_ERR_INVALID;
    if (sm_count) *sm_count = h->sm_count;
    if (cc_major) *cc_major = h->cc_major;
    if (cc_minor) *cc_minor = h->cc_minor;
    if (total_mem) *total_mem = h->total_mem;
    return FNB_OK;
}

// ---------------------------------------------------------------------------------------
// shared launch plumbing

// split/convert `x` ([n, d] fp32 on the device, optionally gathered through perm) into the operand arrays of one
// side of the Gram product and encode their TMA maps into `op`
int fnb::prepare_operand(fnb_context* h, int mode, const float* x, const long long* perm, long long n, int d,
                         bool side_b, GramOperands& op, int normalize, bool defer_split) {
    if (mode == FNB_MODE_AUTO) {
        // FP16F8 when the data satisfies its error model (see FNB_MODE_AUTO in the header), else FP16X3
        int rc = (d % 128 == 0) ? prepare_operand(h, FNB_MODE_FP16F8, x, perm, n, d, side_b, op, normalize) : FNB_OK;
        if (rc) return rc;
        bool f8_ok = (d % 128 == 0);
        if (f8_ok) {
            unsigned int pk = 0;
            CK(cudaMemcpyAsync(&pk, &h->counters.as<DeviceScalars>()->peak_max_ord, 4, cudaMemcpyDeviceToHost, h->stream));
            CK(cudaStreamSynchronize(h->stream));
            op.peakedness = pk ? ordered_to_float(pk) : 0.f;
            f8_ok = op.peakedness <= kAutoPeakLimit;            // false for NaN
        }
        if (f8_ok) return FNB_OK;
        return prepare_operand(h, FNB_MODE_FP16X3, x, perm, n, d, side_b, op, normalize);
    }
    op.mode = mode;
    if (mode_info(mode, &op.num_pass, &op.tf32, &op.fmt, &op.elem_bytes, &op.prescale)) return h->fail(FNB_ERR_INVALID, "bad mode %d", mode);
    DevBuf& hi = side_b ? h->b_hi : h->a_hi;
    DevBuf& lo = side_b ? h->b_lo : h->a_lo;
    DevBuf& h8 = side_b ? h->b_h8 : h->a_h8;
    CUtensorMap* m_hi = side_b ? &op.b_hi : &op.a_hi;
    CUtensorMap* m_lo = side_b ? &op.b_lo : &op.a_lo;
    CUtensorMap* m_h8 = side_b ? &op.b_h8 : &op.a_h8;
    const bool f8 = (op.num_pass == 2);
    const long long n_pad = pad_rows(n);
    const size_t bytes = (size_t)n_pad * d * op.elem_bytes;
    CK(hi.ensure(bytes));
    if (op.num_pass == 3) CK(lo.ensure(bytes));
    if (f8) { CK(lo.ensure((size_t)n_pad * d)); CK(h8.ensure((size_t)n_pad * d)); }
    DevBuf& l16 = side_b ? h->b_l16 : h->a_l16;
    const bool want_l16 = f8 && op.want_l16;
    if (want_l16) CK(l16.ensure(bytes));
    CK(h->counters.ensure(sizeof(DeviceScalars)));
    unsigned int* norm = &h->counters.as<DeviceScalars>()->norm_max_ord;    // [0] max squared norm, [1] max peakedness, [2] its sum over rows
    static_assert(offsetof(DeviceScalars, peak_max_ord) == offsetof(DeviceScalars, norm_max_ord) + 4, "layout");
    static_assert(offsetof(DeviceScalars, peak_sum) == offsetof(DeviceScalars, norm_max_ord) + 8, "layout");
    CK(cudaMemsetAsync(norm, 0, 12, h->stream));
    float* nrm_out = nullptr;
    if (normalize) {
        DevBuf& nb = side_b ? h->b_nrm : h->a_nrm;
        CK(nb.ensure((size_t)n_pad * 4));
        nrm_out = nb.as<float>();
    }
    (side_b ? op.b_nrm : op.a_nrm) = nrm_out;
    // defer_split (streamed launches): the arrays and maps are set up here, split_operand_rows fills them chunk by chunk
    if (!defer_split)
        CK(launch_split_rows(mode, x, perm, n, n_pad, d, hi.p, op.num_pass != 1 ? lo.p : nullptr, f8 ? h8.p : nullptr, norm, h->stream,
                             normalize, nrm_out, want_l16 ? l16.p : nullptr));
    // an operand that two pairs of a cluster share is fetched as two 64-row halves (A: pairs 2 and 4, B: pairs 4)
    const int box_rows = (side_b ? op.pairs == 4 : op.pairs > 1) ? kRowsPerCta / 2 : kRowsPerCta;
    if (!side_b) { op.a_rows_pad = n_pad; h->last_rows = n; }
    int rc = make_tmap(h, m_hi, hi.p, op.fmt, n_pad, d, box_rows);
    if (rc) return rc;
    if (op.num_pass == 3) rc = make_tmap(h, m_lo, lo.p, op.fmt, n_pad, d, box_rows);
    else if (f8) { rc = make_tmap(h, m_lo, lo.p, kFmtU8, n_pad, d, box_rows); if (!rc) rc = make_tmap(h, m_h8, h8.p, kFmtU8, n_pad, d, box_rows); }
    else *m_lo = *m_hi;
    if (!f8) *m_h8 = *m_hi;
    CUtensorMap* m_l16 = side_b ? &op.b_l16 : &op.a_l16;
    *m_l16 = *m_hi;
    if (!rc && want_l16) { rc = make_tmap(h, m_l16, l16.p, kFmtF16, n_pad, d, box_rows); if (!side_b) op.have_l16 = true; }
    return rc;
}

// rows [row_begin, row_end) of the A-side operand arrays that prepare_operand(defer_split) set up; row_end may reach the padded
// row count (the zero rows behind n).  The row-norm / peakedness maxima accumulate over the chunks.
int fnb::split_operand_rows(fnb_context* h, const GramOperands& op, const float* x, const long long* perm, long long n, int d,
                            int normalize, long long row_begin, long long row_end, cudaStream_t stream) {
    const bool f8 = (op.num_pass == 2);
    unsigned int* norm = &h->counters.as<DeviceScalars>()->norm_max_ord;
    CK(launch_split_rows(op.mode, x, perm, n, op.a_rows_pad, d, h->a_hi.p, op.num_pass != 1 ? h->a_lo.p : nullptr, f8 ? h->a_h8.p : nullptr,
                         norm, stream ? stream : h->stream, normalize, normalize ? h->a_nrm.as<float>() : nullptr, op.have_l16 ? h->a_l16.p : nullptr,
                         row_begin, row_end));
    return FNB_OK;
}

int fnb::self_b_maps(fnb_context* h, GramOperands& op, int d) {
    op.b_nrm = op.a_nrm;
    if (op.pairs == 1 || op.pairs == 4) { op.b_hi = op.a_hi; op.b_lo = op.a_lo; op.b_h8 = op.a_h8; op.b_l16 = op.a_l16; return FNB_OK; }
    // pairs == 2: the A maps carry half-height boxes, the B side is not shared: encode full-height boxes over the same arrays
    const bool f8 = (op.num_pass == 2);
    int rc = make_tmap(h, &op.b_hi, h->a_hi.p, op.fmt, op.a_rows_pad, d);
    if (rc) return rc;
    if (op.num_pass == 3) rc = make_tmap(h, &op.b_lo, h->a_lo.p, op.fmt, op.a_rows_pad, d);
    else if (f8) { rc = make_tmap(h, &op.b_lo, h->a_lo.p, kFmtU8, op.a_rows_pad, d); if (!rc) rc = make_tmap(h, &op.b_h8, h->a_h8.p, kFmtU8, op.a_rows_pad, d); }
    else op.b_lo = op.b_hi;
    if (!f8) op.b_h8 = op.b_hi;
    op.b_l16 = op.b_hi;
    if (!rc && op.have_l16) rc = make_tmap(h, &op.b_l16, h->a_l16.p, kFmtF16, op.a_rows_pad, d);
    return rc;
}

int fnb::dl_check_embeddings(fnb_context* h, const DLView& v, const char* name) {
    if (v.code != kDLFloat || v.bits != 32) return h->fail(FNB_ERR_INVALID, "%s: embeddings must be float32", name);
    if (v.cols < 64 || v.cols > 4096 || (v.cols % 64) != 0)
        return h->fail(FNB_ERR_INVALID, "%s: embedding dimension %lld not supported (multiple of 64 in [64, 4096])", name, v.cols);
    if (v.rows > (1LL << 30)) return h->fail(FNB_ERR_INVALID, "%s: too many rows", name);
    return FNB_OK;
}

int fnb::reset_scalars(fnb_context* h) {
    CK(h->counters.ensure(sizeof(DeviceScalars)));
    DeviceScalars init = {};
    init.counters[0] = init.counters[1] = 0;
    init.range_ord[0] = 0xFFFFFFFFu;                    // min
    init.range_ord[1] = 0;                              // max
    init.range_ord[2] = float_to_ordered(0.f);          // max |s|
    init.range_ord[3] = 0;
    CK(h->pinned.ensure(16384));                          // [0, 1 KB) scalars, [1 KB, 4 KB) tables, [4 KB, ..) bins of a whole-set launch
    memcpy(h->pinned.p, &init, sizeof(init));
    // the row-norm word behind the first 32 bytes is written by the split kernel and survives the reset
    CK(cudaMemcpyAsync(h->counters.p, h->pinned.p, offsetof(DeviceScalars, norm_max_ord), cudaMemcpyHostToDevice, h->stream));
    return FNB_OK;
}

// device copy of the accumulation-bias knots: [0] the launch's mode, [1] the fp16x3 contraction of strict tiles in an fp16f8
// launch (hi / l16 at the 2^12 pre-scale: the same 96 accumulation steps as FP16X3)
int fnb::upload_bias(fnb_context* h, int mode, int d, bool off, const float** dev) {
    if (off) { *dev = nullptr; return FNB_OK; }
    CK(h->bias_tab.ensure(2 * kBiasStride * 4));
    if (h->bias_mode != mode || h->bias_d != d) {
        float tab[2 * kBiasStride];
        bias_table(mode, d, tab);
        bias_table(FNB_MODE_FP16X3, d, tab + kBiasStride);
        // pageable source: the runtime stages the 352 bytes before the call returns
        CK(cudaMemcpyAsync(h->bias_tab.p, tab, sizeof(tab), cudaMemcpyHostToDevice, h->stream));
        h->bias_mode = mode; h->bias_d = d;
    }
    *dev = h->bias_tab.as<float>();
    return FNB_OK;
}

int fnb::upload_regions(fnb_context* h, const std::vector<RegionDev>& regs) {
    CK(h->regions.ensure(regs.size() * sizeof(RegionDev)));
    CK(cudaMemcpyAsync(h->regions.p, regs.data(), regs.size() * sizeof(RegionDev), cudaMemcpyHostToDevice, h->stream));
    return FNB_OK;
}

// ---------------------------------------------------------------------------------------
// fnb_pairwise

extern "C" int fnb_pairwise(fnb_handle h, const DLTensor* xa, const DLTensor* xb, const fnb_options* opt_in,
                            DLTensor* out, float* range)
{
    if (!h) return FNB_ERR_INVALID;
    fnb_options opt; if (opt_in) opt = *opt_in; else fnb_default_options(&opt);
    if (opt.metric != 0 && opt.metric != 1) return h->fail(FNB_ERR_BAD_METRIC, "Undefined similarity metric %d", opt.metric);
    CK(cudaSetDevice(h->device));
    GramOperands op;
    if (opt.mode == FNB_MODE_AUTO) opt.mode = FNB_MODE_FP16X3;       // materialised distances: always the fp32-equivalent split
    if (mode_info(opt.mode, &op.num_pass, &op.tf32, &op.fmt, &op.elem_bytes, &op.prescale)) return h->fail(FNB_ERR_INVALID, "bad mode %d", opt.mode);
    DLView va, vb, vo;
    int rc = dl_view(h, xa, "xa", 2, 2, &va); if (rc) return rc;
    if ((rc = dl_check_embeddings(h, va, "xa"))) return rc;
    const bool self = (xb == nullptr);
    if (!self) {
        if ((rc = dl_view(h, xb, "xb", 2, 2, &vb))) return rc;
        if ((rc = dl_check_embeddings(h, vb, "xb"))) return rc;
        if (vb.cols != va.cols) return h->fail(FNB_ERR_INVALID, "xa and xb have different dimensions");
    }
    if ((rc = dl_view(h, out, "out", 1, 2, &vo))) return rc;
    if (vo.code != kDLFloat || vo.bits != 32) return h->fail(FNB_ERR_INVALID, "out must be float32");
    const long long na = va.rows, nb = self ? va.rows : vb.rows;
    const int d = (int)va.cols;
    const long long out_elems = self ? na * (na - 1) / 2 : na * nb;
    const long long have = vo.rows * vo.cols;
    if (have != out_elems) return h->fail(FNB_ERR_INVALID, "out has %lld elements, expected %lld", have, out_elems);
    if (range) { range[0] = INFINITY; range[1] = -INFINITY; }
    if (out_elems == 0) return FNB_OK;                   // statistics.py:38: empty in -> empty out

    const void* da = nullptr; const void* db = nullptr;
    if ((rc = dl_to_device(h, va, (size_t)na * d * 4, h->stage_a, &da))) return rc;
    if (!self && (rc = dl_to_device(h, vb, (size_t)nb * d * 4, h->stage_b, &db))) return rc;
    if ((rc = prepare_operand(h, opt.mode, (const float*)da, nullptr, na, d, false, op, opt.normalize))) return rc;
    if (self) { if ((rc = self_b_maps(h, op, d))) return rc; }
    else if ((rc = prepare_operand(h, opt.mode, (const float*)db, nullptr, nb, d, true, op, opt.normalize))) return rc;

    const int cg = pick_cta_group(&opt);
    const int tile = kRowsPerCta * cg;
    std::vector<RegionDev> regs;
    if (self) triangle_regions(na, pick_region_rows(&opt, tile, na, d), 0, regs);
    else {
        RegionDev r = {}; r.row_end = (int)na; r.col_end = (int)nb; regs.push_back(r);
    }
    finish_regions(regs, tile);
    if ((rc = upload_regions(h, regs))) return rc;
    if ((rc = reset_scalars(h))) return rc;

    float* dout = nullptr;
    if (vo.on_device) dout = (float*)vo.data;
    else { CK(h->out.ensure((size_t)out_elems * 4)); dout = h->out.as<float>(); }

    GramParams p = {};
    p.regions = h->regions.as<RegionDev>(); p.nregions = (int)regs.size() - 1; p.total_tiles = regs.back().tile_begin;
    p.shard = ShardSpec{1, 0, 1, nullptr};
    p.kblocks = d / (128 / op.elem_bytes);
    p.acc_scale = 1.0f / (op.prescale * op.prescale);
    p.operand_fmt = op.fmt;
    DeviceScalars* sc = h->counters.as<DeviceScalars>();
    p.counters = sc->counters; p.range_ord = sc->range_ord;
    p.out = dout; p.out_ld = nb; p.tri_packed = self ? 1 : 0; p.metric = opt.metric;
    p.n_rows = (int)na; p.n_cols = (int)nb;
    p.raw = opt.raw_distance;
    if (opt.normalize == 1 && opt.theta != 0.f) { p.row_nrm = op.a_nrm; p.col_nrm = op.b_nrm; p.theta = opt.theta; }
    if ((rc = upload_bias(h, op.mode, d, opt.bias_correction < 0, &p.bias_beta))) return rc;
    if ((rc = launch_gram(h, cg, EPI_PAIRWISE, opt.max_ctas, op, p, 0))) return rc;

    DeviceScalars hs;
    CK(cudaMemcpyAsync(h->pinned.p, h->counters.p, sizeof(hs), cudaMemcpyDeviceToHost, h->stream));
    if (!vo.on_device) CK(cudaMemcpyAsync(vo.data, dout, (size_t)out_elems * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    memcpy(&hs, h->pinned.p, sizeof(hs));
    const float smin = ordered_to_float(hs.range_ord[0]), smax = ordered_to_float(hs.range_ord[1]);
    if (range) { range[0] = smin; range[1] = smax; }
    const double lim = 1.0 + (double)opt.atol;
    if (!opt.raw_distance && ((double)smin < -lim || (double)smax > lim || smin != smin || smax != smax))
        return h->fail(FNB_ERR_NOT_NORMALIZED, "embeddings must be normalized to 1, range %.9g %.9g", smin, smax);
    return FNB_OK;
}

// ---------------------------------------------------------------------------------------
// histograms

static double mode_slack(int mode);

struct HistLaunch {
    CutTables ct;
    int nkeys = 1;
    int stride = kMaxBins + 1;
    // the auto rule of fnb_options.panel_window applies (whole-set launches: every region is a super-row whose column panels
    // hold many tiles per cluster; keyed launches can have one-tile regions, where a progress window would serialise the clusters)
    bool auto_window = false;
    // Streamed pass: the launch is cut into one launch per column chunk of the pair matrix (chunks[k] = the regions whose
    // columns lie in chunk k, rows anywhere above); before_launch(k) queues whatever launch k waits for (the upload and the
    // split of the chunk's rows).  The bins, counters and range words accumulate over the launches.
    const std::vector<std::vector<RegionDev>>* chunks = nullptr;
    std::function<int(int)> before_launch;
    // sharded job (fnb_pair_histogram_sharded): SMs are kept free for the row exchange; steal_world > 1: the ranks can take tiles
    // from each other's queues (every rank's counters are mapped into every rank and zeroed by their owner before the label exchange)
    bool sharded = false;
    int steal_world = 0, steal_rank = 0;
    unsigned long long* steal_ctr[8] = {};          // rank v's counters for this pass (steal_ctr[steal_rank] = this rank's own)
};

// uploads tables, zeroes bins, launches the HIST kernel over `regs`; leaves bins on the device
static int run_hist(fnb_context* h, const fnb_options& opt, GramOperands& op, const std::vector<RegionDev>& regs, int cg,
                    int d, const int32_t* cls_dev, const double* thresholds, int T, HistLaunch& hl, int force_slow)
{
    float knots[kBiasStride], knots_x3[kBiasStride];
    bias_table(op.mode, d, knots);
    bias_table(FNB_MODE_FP16X3, d, knots_x3);
    const bool bias_off = opt.bias_correction < 0;
    if (build_cut_tables(thresholds, T, opt.metric, opt.eps, opt.cuts, &hl.ct, bias_off ? nullptr : knots, bias_off ? nullptr : knots_x3))
        return h->fail(FNB_ERR_INVALID, "number of thresholds must be in [1, %d]", kMaxBins - 1);
    int rc;
    std::vector<RegionDev> all_chunks;                   // streamed pass: every chunk's regions (each list ends with its sentinel)
    if (hl.chunks) for (const auto& c : *hl.chunks) all_chunks.insert(all_chunks.end(), c.begin(), c.end());
    if ((rc = upload_regions(h, hl.chunks ? all_chunks : regs))) return rc;
    if ((rc = reset_scalars(h))) return rc;
    const size_t tab_floats = kMaxBins + 2 * (kMaxBins + 4);
    CK(h->tables.ensure(tab_floats * 4));
    float* htab = h->pinned.as<float>() + 256;           // past the scalars staging area
    memcpy(htab, hl.ct.cuts, kMaxBins * 4);
    memcpy(htab + kMaxBins, hl.ct.wlo, (kMaxBins + 4) * 4);
    memcpy(htab + kMaxBins + kMaxBins + 4, hl.ct.whi, (kMaxBins + 4) * 4);
    CK(cudaMemcpyAsync(h->tables.p, htab, tab_floats * 4, cudaMemcpyHostToDevice, h->stream));
    const size_t bins_bytes = (size_t)hl.nkeys * 2 * hl.stride * 8;
    CK(h->bins.ensure(bins_bytes));
    CK(cudaMemsetAsync(h->bins.p, 0, bins_bytes, h->stream));

    GramParams p = {};
    p.regions = h->regions.as<RegionDev>();
    if (!hl.chunks) { p.nregions = (int)regs.size() - 1; p.total_tiles = regs.back().tile_begin; }     // streamed pass: set per launch below
    p.shard = h->last_shard;
    p.kblocks = d / (128 / op.elem_bytes);
    p.acc_scale = 1.0f / (op.prescale * op.prescale);
    p.operand_fmt = op.fmt;
    p.force_slow = force_slow || opt.force_checked;
    if ((rc = upload_bias(h, op.mode, d, bias_off, &p.bias_beta))) return rc;
    // whole-set launches decide per global 512 x 512 block (bins independent of the tiling); keyed launches have arbitrary
    // rectangles: every tile strict
    p.strict_tiles = (op.num_pass == 2 && op.have_l16 && opt.strict_tiles >= 0) ? (hl.auto_window ? 1 : 2) : 0;
    if (p.strict_tiles == 1) {
        const int n_rows = (int)h->last_rows, nb = (n_rows + 511) / 512;
        const size_t words = ((size_t)nb * nb + 31) / 32 + 1;
        CK(h->strict_bits.ensure(words * 4));
        CK(cudaMemsetAsync(h->strict_bits.p, 0, words * 4, h->stream));
        CK(launch_strict_blocks(cls_dev, n_rows, kRowsPerCta * cg, nb, h->strict_bits.as<unsigned int>(), h->stream));
        p.strict_bits = h->strict_bits.as<unsigned int>(); p.strict_nb = nb;
    }
    h->last_strict = p.strict_tiles;
    p.raw = opt.raw_distance;
    if (opt.normalize == 1 && opt.theta != 0.f) { p.row_nrm = op.a_nrm; p.col_nrm = op.b_nrm; p.theta = opt.theta; }
    p.debug = opt.debug & 3;
    // cluster-progress window (fnb_options.panel_window, GramParams::sync_window): auto = on for the launches long enough for the
    // clusters to drift apart (this rank's share of the pair matrix >= 5e10 pairs, the same launches pick_pairs calls long)
    auto own_pairs_of = [](const std::vector<RegionDev>& rv) {
        double own_pairs = 0.0;
        for (size_t i = 0; i + 1 < rv.size(); ++i) {
            const RegionDev& r = rv[i];
            const double area = (double)(r.row_end - r.row_begin) * (double)(r.col_end - r.col_begin) * (r.tri ? 0.5 : 1.0);
            if (r.nrb > 0) own_pairs += area * (double)r.own_cnt / (double)r.nrb;
        }
        return own_pairs;
    };
    // streamed pass: the job as a whole decides (its launches are as long, in sum, as the one launch they replace); launches
    // of less than 1e9 pairs have nothing to drift over
    double job_pairs = 0.0;
    if (hl.chunks) for (const auto& c : *hl.chunks) job_pairs += own_pairs_of(c);
    auto window_of = [&](const std::vector<RegionDev>& rv) {
        if (opt.panel_window > 0) return std::min(opt.panel_window, 7);
        if (opt.panel_window != 0 || !hl.auto_window) return 0;
        const double own = own_pairs_of(rv);
        return (own >= 5.0e10 || (job_pairs >= 5.0e10 && own >= 1.0e9)) ? 2 : 0;
    };
    const int nlaunch = hl.chunks ? (int)hl.chunks->size() : 1;
    p.sync_window = window_of(hl.chunks ? hl.chunks->back() : regs);
    h->last_window = p.sync_window;
    CK(h->progress.ensure((size_t)nlaunch * 1024 * 4));
    CK(cudaMemsetAsync(h->progress.p, 0, (size_t)nlaunch * 1024 * 4, h->stream));
    p.progress = h->progress.as<unsigned int>();
    // tile queue (fnb_options.tile_queue): on for every histogram launch; the auxiliary launch on the free SMs only for one-GPU
    // jobs (a sharded job's row exchange runs there) and never under the profiling knobs
    long long most_tiles = hl.chunks ? 0 : regs.back().tile_begin;
    if (hl.chunks) for (const auto& c : *hl.chunks) most_tiles = std::max(most_tiles, c.back().tile_begin);
    // (an explicit progress window asks for the static schedule; queue entries are 32-bit tile indices)
    const bool use_queue = opt.tile_queue >= 0 && opt.panel_window <= 0 && most_tiles < 0x7fffffffll;
    // (a sharded job keeps kAuxReserve of the free SMs for its row exchange -- gather kernels and NCCL's CTAs run there under the launches)
    const bool use_aux = use_queue && opt.tile_queue != 2 && opt.max_ctas == 0 && op.pairs == 2;
    // ... while an exchange is still to come: nothing travels under the LAST launch of a pass (or under the one launch of a pass
    // whose rows were gathered up front), which therefore takes every free SM
    int aux_reserve = (hl.sharded && hl.chunks && hl.chunks->size() > 1) ? kAuxReserveSms : 0;
    unsigned long long* counters = nullptr;
    if (use_queue) {
        // (a stolen tile travels as index | owner << kStealShift in one int: jobs with more tiles per rank than that keep to their own
        // queues -- the same decision on every rank, their schedules have identical shape)
        if (hl.steal_world > 1 && most_tiles < (1ll << kStealShift)) {
            counters = hl.steal_ctr[hl.steal_rank];          // zeroed by its owner at the start of the call
            p.steal_world = hl.steal_world; p.steal_rank = hl.steal_rank;
            for (int v = 0; v < hl.steal_world; ++v) p.steal_counter[v] = hl.steal_ctr[v];
        } else if (hl.steal_world > 1) {
            counters = hl.steal_ctr[hl.steal_rank];          // own queue only
        } else {
            CK(h->tile_counter.ensure((size_t)nlaunch * 8));
            CK(cudaMemsetAsync(h->tile_counter.p, 0, (size_t)nlaunch * 8, h->stream));
            counters = h->tile_counter.as<unsigned long long>();
        }
        p.tile_counter = counters;
        p.sync_window = 0;
        h->last_window = 0;
    }
    if (use_aux && !h->aux_stream) {
        CK(cudaStreamCreateWithFlags(&h->aux_stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&h->aux_ev[0], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->aux_ev[1], cudaEventDisableTiming));
    }
    auto hist_bytes_of = [](const GramParams& pl) { return (size_t)pl.nb8 * kHist8Row; };
    // main launch, then (behind it) the pairs on the SMs it leaves free; the handle's stream continues after both
    auto launch_both = [&](GramParams& pl) -> int {
        int r2;
        // the auxiliary launch starts BEHIND the main one (the event in front of it): the main grid takes its SMs first
        if (use_aux) CK(cudaEventRecord(h->aux_ev[0], h->stream));
        if ((r2 = launch_gram(h, cg, EPI_HIST, opt.max_ctas, op, pl, hist_bytes_of(pl)))) return r2;
        if (use_aux) {
            CK(cudaStreamWaitEvent(h->aux_stream, h->aux_ev[0], 0));
            const int grid_before = h->last_grid;
            if ((r2 = launch_gram_aux(h, op, pl, hist_bytes_of(pl), aux_reserve))) return r2;
            if (h->last_grid > grid_before) ++h->aux_launches;
            CK(cudaEventRecord(h->aux_ev[1], h->aux_stream));
            CK(cudaStreamWaitEvent(h->stream, h->aux_ev[1], 0));
        }
        return FNB_OK;
    };
    p.row_cls = cls_dev; p.col_cls = cls_dev;
    p.cuts = h->tables.as<float>(); p.wlo = p.cuts + kMaxBins; p.whi = p.wlo + kMaxBins + 4;
    p.T = T; p.T_fin = hl.ct.T_fin; p.uniform = hl.ct.uniform;
    p.nb8 = hl.ct.T_fin + 3;
    p.frac_bits = 12;                                    // keeps hb well-defined on the checked-only path
    if (hl.ct.uniform) {
        // Arithmetic binning of interior tiles (see GramParams): with u = (s - cut_0) / h and Q = T_fin + 2,
        //   v = sat((u + 0.5) / Q),  kw = v * Q * R + R/2 = (u + 1) R,  bin = kw >> F,  R = 2^F.
        // A pair is flagged "near a threshold" when kw is within `half` units of a multiple of R; `half` covers
        // the eps window, the deviation of the true cuts from the progression and the fp32 rounding (< 1.5 units).
        const double eps_s = (opt.metric == 0) ? opt.eps / 2.0 : opt.eps;
        const int Q = hl.ct.T_fin + 2;
        int F = 16;
        while (F > 4 && (double)(Q + 1) * (double)(1u << F) > 4194304.0) --F;
        const double R = (double)(1u << F);
        const double need = (eps_s / hl.ct.h + 2.0 * std::max(hl.ct.dev, hl.ct.devx) / hl.ct.h) * R + 1.5;
        int W = 1;
        while ((double)(1u << (W - 1)) < need && W < 30) ++W;
        if (W > F - 2) {
            p.uniform = 0;                               // thresholds too dense for the eps window: checked path
        } else {
            const double half = (double)(1u << (W - 1));
            p.f_s1 = (float)((double)p.acc_scale / (hl.ct.h * Q));
            p.f_b1 = (float)((-hl.ct.e0 / hl.ct.h + 0.5) / Q);
            p.f_k2 = (float)(Q * R);
            p.f_magic_k = (float)(12582912.0 + 0.5 * R);
            p.f_magic_n = (float)(12582912.0 + 0.5 * R + half);
            p.frac_bits = (unsigned)F;
            p.near_mask = ((1u << F) - 1u) & ~((1u << W) - 1u);
            // Without clipping the bin index is floor(u + 1): it must stay inside the counter rows [0, nb8) for every
            // similarity an interior tile can produce, |s| <= 1 + atol (row-norm bound) + the arithmetic error of the mode
            const double lim = 1.0 + (double)opt.atol + mode_slack(opt.mode) + 1.0e-4;
            const double idx_lo = (-lim - hl.ct.e0) / hl.ct.h + 1.0, idx_hi = (lim - hl.ct.e0) / hl.ct.h + 1.0;
            if (idx_lo > 0.02 && idx_hi < (double)p.nb8 - 0.02) {
                p.noclip = 1;
                p.f_g1 = (float)((double)p.acc_scale / hl.ct.h * R);
                p.f_g0 = (float)((-hl.ct.e0 / hl.ct.h + 1.0) * R + 12582912.0);
                p.f_g1x = (float)((double)p.acc_scale / hl.ct.hx * R);           // strict (fp16x3) tiles of an fp16f8 launch
                p.f_g0x = (float)((-hl.ct.e0x / hl.ct.hx + 1.0) * R + 12582912.0);
                p.near_half = (unsigned)half;
            }
            h->last_eps_counted = (opt.metric == 0 ? 2.0 : 1.0) * half / R * hl.ct.h;
        }
    }
    // interior tiles skip the per-pair range check when every row norm proves |s| <= 1 + atol (Cauchy-Schwarz);
    // the 3-pass modes reproduce fp32 to ~2e-7, single-pass modes get the checked path only on real violations
    p.norm_max_ord = &h->counters.as<DeviceScalars>()->norm_max_ord;
    p.norm_limit_ord = float_to_ordered((float)(1.0 + (double)opt.atol - 2.0e-6));
    p.bins = h->bins.as<unsigned long long>(); p.bins_stride = hl.stride;
    DeviceScalars* sc = h->counters.as<DeviceScalars>();
    p.counters = sc->counters; p.range_ord = sc->range_ord;
    p.metric = opt.metric;
    h->chunk_launches = 0;
    h->aux_launches = 0;
    if (!hl.chunks) {
        CK(cudaEventRecord(h->ev[1], h->stream));
        if ((rc = launch_both(p))) return rc;
        CK(cudaEventRecord(h->ev[2], h->stream));
    } else {
        while (h->chunk_ev.size() < 2 * (size_t)nlaunch) {
            cudaEvent_t e = nullptr;
            CK(cudaEventCreate(&e));
            h->chunk_ev.push_back(e);
        }
        size_t off = 0;
        for (int k = 0; k < nlaunch; ++k) {
            const std::vector<RegionDev>& rv = (*hl.chunks)[k];
            if (hl.before_launch && (rc = hl.before_launch(k))) return rc;
            p.regions = h->regions.as<RegionDev>() + off;
            p.nregions = (int)rv.size() - 1;
            p.total_tiles = rv.back().tile_begin;
            p.sync_window = use_queue ? 0 : window_of(rv);
            p.progress = h->progress.as<unsigned int>() + (size_t)k * 1024;
            if (use_queue) p.tile_counter = counters + k;
            if (use_queue && p.steal_world > 1) for (int v = 0; v < hl.steal_world; ++v) p.steal_counter[v] = hl.steal_ctr[v] + k;
            if (k + 1 == nlaunch) aux_reserve = 0;
            off += rv.size();
            CK(cudaEventRecord(h->chunk_ev[2 * k], h->stream));
            if (p.total_tiles > 0 && (rc = launch_both(p))) return rc;
            CK(cudaEventRecord(h->chunk_ev[2 * k + 1], h->stream));
        }
        h->chunk_launches = nlaunch;
    }
    h->last_nkeys = hl.nkeys; h->last_T = T;
    return FNB_OK;
}

static int finish_hist(fnb_context* h, const fnb_options& opt, fnb_stats* stats, float* smin_out, float* smax_out, bool* violated) {
    DeviceScalars hs;
    CK(cudaMemcpyAsync(h->pinned.p, h->counters.p, sizeof(hs), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    memcpy(&hs, h->pinned.p, sizeof(hs));
    // checked tiles track min/max exactly; interior tiles are covered by the row-norm bound (max squared norm)
    const bool have_checked = hs.range_ord[0] != 0xFFFFFFFFu;
    const float smin = have_checked ? ordered_to_float(hs.range_ord[0]) : NAN;
    const float smax = have_checked ? ordered_to_float(hs.range_ord[1]) : NAN;
    const unsigned int norm_ord = hs.norm_max_ord;
    const float amax = norm_ord ? ordered_to_float(norm_ord) : NAN;
    const double lim = 1.0 + (double)opt.atol + mode_slack(opt.mode);
    *violated = !opt.raw_distance && have_checked && !((double)smin >= -lim && (double)smax <= lim);
    *smin_out = smin;
    *smax_out = smax;
    h->last_peak = hs.peak_max_ord ? ordered_to_float(hs.peak_max_ord) : 0.f;
    h->last_peak_mean = h->last_rows > 0 ? hs.peak_sum / (float)h->last_rows : 0.f;
    if (stats) {
        float ms = 0.f, pm = 0.f;
        if (h->chunk_launches > 0) {
            // streamed pass: the launches' own durations (the waits for the chunks in between are not kernel time)
            for (int k = 0; k < h->chunk_launches; ++k) { float m1 = 0.f; cudaEventElapsedTime(&m1, h->chunk_ev[2 * k], h->chunk_ev[2 * k + 1]); ms += m1; }
            if (getenv("FNB_TRACE")) {
                // timeline of the launches on the handle's stream, ms since the start of the call: start, end of every launch
                fprintf(stderr, "[fnb trace] rank %d:", h->last_shard.lo);
                for (int k = 0; k < h->chunk_launches; ++k) {
                    float t0 = 0.f, t1 = 0.f;
                    cudaEventElapsedTime(&t0, h->ev[0], h->chunk_ev[2 * k]); cudaEventElapsedTime(&t1, h->ev[0], h->chunk_ev[2 * k + 1]);
                    fprintf(stderr, " [%.2f %.2f]", t0, t1);
                }
                fprintf(stderr, "\n[fnb trace] rank %d exchange (start, own rows in place, broadcast done):", h->last_shard.lo);
                for (size_t k = 0; 3 * k + 2 < h->xchg_ev.size() && (int)k < h->chunk_launches; ++k) {
                    float t0 = 0.f, t1 = 0.f, t2 = 0.f;
                    if (cudaEventElapsedTime(&t0, h->ev[0], h->xchg_ev[3 * k]) != cudaSuccess) { cudaGetLastError(); break; }
                    cudaEventElapsedTime(&t1, h->ev[0], h->xchg_ev[3 * k + 1]); cudaEventElapsedTime(&t2, h->ev[0], h->xchg_ev[3 * k + 2]);
                    fprintf(stderr, " [%.2f %.2f %.2f]", t0, t1, t2);
                }
                fprintf(stderr, "\n");
            }
            cudaEventElapsedTime(&pm, h->ev[0], h->chunk_ev[0]);
        } else {
            cudaEventElapsedTime(&ms, h->ev[1], h->ev[2]);
            cudaEventElapsedTime(&pm, h->ev[0], h->ev[1]);
        }
        stats->kernel_ms = ms;
        stats->prepare_ms = pm;
        stats->eps_window = hs.counters[0];
        stats->tiles = hs.counters[1];
        stats->smin = smin;
        stats->smax = smax;
        stats->max_abs = amax;
        stats->eps_counted = (float)h->last_eps_counted;
        stats->kernel_launches += (h->chunk_launches > 0 ? 2 * h->chunk_launches - 1 : 1) + h->aux_launches;     // streamed: a split + a Gram launch per chunk; + the launches of CTA pairs on the free SMs
        stats->grid_ctas = (uint32_t)h->last_grid;
        stats->mode_used = h->last_mode;
        stats->panel_window = h->last_window;
        stats->peakedness = hs.peak_max_ord ? ordered_to_float(hs.peak_max_ord) : 0.f;
        stats->h2d_bytes = h->last_h2d_bytes;
        if (h->h2d_timed) { float hm = 0.f; if (cudaEventElapsedTime(&hm, h->copy_ev[1], h->copy_ev[2]) == cudaSuccess) stats->h2d_ms = hm; }
    }
    return FNB_OK;
}

// absolute error of a similarity computed in a single low-precision pass: the range check must not reject
// normalised inputs because of the arithmetic mode (the 3-pass modes are fp32-equivalent: no slack)
static double mode_slack(int mode) {
    switch (mode) {
        case FNB_MODE_BF16: return 1.6e-2;
        case FNB_MODE_TF32: case FNB_MODE_FP16: return 2.0e-3;
        case FNB_MODE_FP16F8: return 2.0e-5;
        default: return 0.0;
    }
}

// the range check of finish_hist, on the device: *flag = 1 when this rank's checked tiles saw a similarity outside
// [-(1 + atol + slack), 1 + atol + slack] (NaN included).  Sharded jobs carry the flag through the all-reduce of the bins.
__global__ void range_violation_kernel(const unsigned int* __restrict__ range_ord, double lim, int raw, unsigned long long* flag) {
    const unsigned int mn = range_ord[0], mx = range_ord[1];
    const bool have = mn != 0xFFFFFFFFu;
    const double smin = (double)ordered_to_float(mn), smax = (double)ordered_to_float(mx);
    *flag = (!raw && have && !(smin >= -lim && smax <= lim)) ? 1ull : 0ull;
}

static int flag_range_violation(fnb_context* h, const fnb_options& opt, unsigned long long* flag) {
    const double lim = 1.0 + (double)opt.atol + mode_slack(opt.mode);
    range_violation_kernel<<<1, 1, 0, h->stream>>>(h->counters.as<DeviceScalars>()->range_ord, lim, opt.raw_distance, flag);
    CK(cudaGetLastError());
    return FNB_OK;
}

static uint64_t sum_bins(const uint64_t* b, int n) { uint64_t s = 0; for (int i = 0; i < n; ++i) s += b[i]; return s; }

// A-posteriori error bound of one histogram launch (fnb_stats.error_bound).  The histogram tells how many pairs were found in
// every similarity bin; for each non-empty bin the model error of the arithmetic that binned those pairs is
//     residual of the bias correction (35 % of beta |s|: the data dependence seen in profiles/r02a_bias_{relu,sparse,t3}.log)
//   + z sigma(|s|),  z = sqrt(2 ln(1000 (T + 1) M))  -- a union bound over the M pairs of the bin and all bins at 1e-3,
// in distance units (metric 0: dd = 2 ds; metric 1 keeps similarity units like the eps window).  Same-identity pairs of an
// fp16f8 launch with strict tiles ran in the fp16x3 contraction.  all / same: [T + 1] counts summed over keys.
static double error_certificate(const fnb_options& opt, int mode, bool strict_tiles, int d, double peak_mean, double peak_max,
                                long long n_rows, const CutTables& ct, const uint64_t* all, const uint64_t* same, int T)
{
    const double share = (double)std::max(1, opt.world);     // a rank sees 1 / world of every bin: bound the whole job
    float tab_mode[kBiasStride], tab_x3[kBiasStride];
    bias_table(mode, d, tab_mode);
    bias_table(FNB_MODE_FP16X3, d, tab_x3);
    const bool corrected = opt.bias_correction >= 0;
    double worst = 0.0;
    for (int k = 0; k <= T && k <= ct.T_fin; ++k) {
        const double lo = k > 0 ? (double)ct.cuts[k - 1] : -1.0, hi = k < ct.T_fin ? (double)ct.cuts[k] : 1.0;
        const double a = std::min(1.0, std::max(std::fabs(lo), std::fabs(hi)));
        const uint64_t m_same = same[k], m_diff = all[k] - same[k];
        for (int which = 0; which < 2; ++which) {
            const uint64_t m = which ? m_same : m_diff;
            if (!m) continue;
            const bool x3 = which && strict_tiles && mode == FNB_MODE_FP16F8;
            const int md = x3 ? FNB_MODE_FP16X3 : mode;
            const double beta = beta_at(x3 ? tab_x3 : tab_mode, a);
            const double resid = (corrected ? 0.35 : 1.0) * beta * a;
            // bulk of the bin: typical rows (mean peakedness); and the pairs of the peakiest row with typical rows
            // (sum x^2 y^2 <= sqrt(sum x^4 sum y^4)): at most n_rows of them fall into the bin
            const double z = std::sqrt(2.0 * std::log(1000.0 * (T + 1) * (double)m * share));
            const double m_row = std::min((double)m, (double)std::max<long long>(n_rows, 1));
            const double z_row = std::sqrt(2.0 * std::log(1000.0 * (T + 1) * m_row));
            const double pk_row = std::sqrt(std::max(peak_mean, 0.0) * std::max(peak_max, peak_mean));
            const double e_s = resid + std::max(z * mode_sigma_s(md, d, a, peak_mean), z_row * mode_sigma_s(md, d, a, pk_row));
            worst = std::max(worst, opt.metric == 0 ? 2.0 * e_s : e_s);
        }
    }
    return worst;
}

// Regions of one column chunk of a streamed pass: the pairs (row < col) with col in [c0, c1), per super-row [r0, r1) of the tile
// order -- columns inside the super-row's own span [a, b) see the rows [r0, a) above them (a rectangle) and each other (a
// triangle); columns to its right see all of its rows (a rectangle).  Over the chunks of a partition of [0, n) every pair is
// covered exactly once (tests/test_host_logic.py::test_chunk_plan_covers_every_pair_once through fnb_debug_chunk_plan).
static void chunk_regions(long long n, long long rr, long long c0, long long c1, std::vector<RegionDev>& out) {
    for (long long r0 = 0; r0 < c1; r0 += rr) {
        const long long r1 = std::min<long long>(n, r0 + rr);
        const long long a = std::max(c0, r0), b = std::min(c1, r1);
        RegionDev g = {};
        g.key = 0;
        if (a < b) {
            if (a > r0) { g.row_begin = (int)r0; g.row_end = (int)a; g.col_begin = (int)a; g.col_end = (int)b; g.tri = 0; out.push_back(g); }
            g.row_begin = (int)a; g.row_end = (int)b; g.col_begin = (int)a; g.col_end = (int)b; g.tri = 1; out.push_back(g);
        }
        const long long right = std::max(c0, r1);
        if (right < c1) { g.row_begin = (int)r0; g.row_end = (int)r1; g.col_begin = (int)right; g.col_end = (int)c1; g.tri = 0; out.push_back(g); }
    }
}

// chunk boundaries of a streamed pass: multiples of `g` rows growing with the work already queued (1, 1, 1, 1, 2, 3, 4, 6, 9 ...
// granules): launch k is then at least as long as the transfer of chunk k + 1
static std::vector<long long> chunk_schedule(long long n, long long g) {
    std::vector<long long> bounds = {0};
    for (long long pos = 0; pos < n;) {
        pos = std::min(n, pos + std::max(g, (pos / 2) / g * g));
        bounds.push_back(pos);
    }
    return bounds;
}

// Test hook (host only, no GPU): the chunk plan of a streamed pass over n rows with super-rows of rr rows and a granule of g rows.
// regions: up to cap entries {chunk, row_begin, row_end, col_begin, col_end, tri}; returns the number of regions (or -needed).
extern "C" int fnb_debug_chunk_plan(long long n, long long rr, long long g, int cap, int* regions /* [cap][6] */, int* nchunks) {
    if (n < 1 || rr < 1 || g < 1) return 0;
    const std::vector<long long> bounds = chunk_schedule(n, g);
    int count = 0;
    for (size_t k = 0; k + 1 < bounds.size(); ++k) {
        std::vector<RegionDev> regs;
        chunk_regions(n, rr, bounds[k], bounds[k + 1], regs);
        for (const RegionDev& r : regs) {
            if (count < cap && regions) {
                int* o = regions + (size_t)count * 6;
                o[0] = (int)k; o[1] = r.row_begin; o[2] = r.row_end; o[3] = r.col_begin; o[4] = r.col_end; o[5] = r.tri;
            }
            ++count;
        }
    }
    if (nchunks) *nchunks = (int)bounds.size() - 1;
    return count <= cap ? count : -count;
}

// ---------------------------------------------------------------------------------------
// whole-set histogram: one job object shared by fnb_pair_histogram_bins (one GPU, or a caller that shards by itself through
// fnb_options.rank / world) and fnb_pair_histogram_sharded (NCCL inside the library)

namespace {

struct WholeSetJob {
    fnb_context* h;
    fnb_options opt;
    fnb_stats* stats;
    const double* thresholds;
    int T;
    long long n = 0;
    int d = 0;
    int cg = 2, tile = 256;
    int requested = FNB_MODE_FP16X3;
    ShardHost shard;
    GramOperands op;
    const void* de = nullptr;               // fp32 rows on the device
    const long long* perm_dev = nullptr;    // class order of the rows at `de` (NULL: they are stored in class order)
    const long long* stream_perm = nullptr; // streamed pass: class position -> row of h->stage_a (NULL: the pieces are fed in class order)
    bool reduce = false;                    // all-reduce the bins (and the range-violation flag) over the handle's communicator
    bool sharded = false;                   // fnb_pair_histogram_sharded (row exchange under the launches)
    int steal_world = 0, steal_rank = 0;    // > 1: the ranks share their tile queues (counters mapped through CUDA IPC)
    unsigned long long* steal_ctr[8] = {};
    int pass_index = 0;                     // every pass of a call takes its own 64 counters
    void arm(HistLaunch& hl) {
        hl.sharded = sharded;
        hl.steal_world = steal_world; hl.steal_rank = steal_rank;
        const int base = 64 * std::min(pass_index++, 3);
        for (int v = 0; v < steal_world && v < 8; ++v) hl.steal_ctr[v] = steal_ctr[v] + base;
    }
    uint64_t host_bins[2 * (kMaxBins + 1)];
    size_t row_bytes = 0;

    long long super_rows() {
        const int cl_size = cg * op.pairs;
        const int clusters = h->hist_grid[op.pairs] > 0 ? h->hist_grid[op.pairs] / cl_size
                                                        : (op.pairs == 1 ? h->sm_count / cl_size : op.pairs == 2 ? h->sm_count / cl_size - 4 : 15);
        // 512-aligned regions (see strict_tile in the kernel), super-row height matched to the cluster count of the launch
        return pick_region_rows(&opt, 512, n, d, clusters, tile * (op.pairs == 4 ? 2 : 1));
    }

    // after the launches of a pass: (sharded) sum the bins over the ranks, then scalars + bins to the host
    int collect(HistLaunch& hl, double* bound, bool* violated, float* smin, float* smax) {
        if (reduce) {
            // slot [0][kMaxBins] of the bins (never a bin: T < kMaxBins) carries the ranks' range violations through the all-reduce
            int rc = flag_range_violation(h, opt, h->bins.as<unsigned long long>() + kMaxBins);
            if (rc) return rc;
            if ((rc = comm_all_reduce_u64(h, h->bins.p, 2 * (size_t)hl.stride, false, h->stream))) return rc;
        }
        CK(cudaMemcpyAsync(h->pinned.as<char>() + 4096, h->bins.p, 2 * (size_t)hl.stride * 8, cudaMemcpyDeviceToHost, h->stream));
        int rc = finish_hist(h, opt, stats, smin, smax, violated);          // synchronises
        if (rc) return rc;
        const uint64_t* pb = reinterpret_cast<const uint64_t*>(h->pinned.as<char>() + 4096);
        if (reduce && pb[kMaxBins] != 0) *violated = true;
        for (int r = 0; r < 2; ++r) memcpy(host_bins + (size_t)r * (T + 1), pb + (size_t)r * hl.stride, (size_t)(T + 1) * 8);
        if (!*violated) {
            fnb_options o = opt;
            if (reduce) o.world = 1;                     // the bins are the whole job's
            *bound = error_certificate(o, op.mode, h->last_strict != 0, d, h->last_peak_mean, h->last_peak, n, hl.ct, host_bins, host_bins + (T + 1), T);
        }
        return FNB_OK;
    }

    int not_normalized(HistLaunch& hl, const std::vector<RegionDev>& regs) {
        // re-run with every tile on the checked path to report the exact similarity range (of this rank's tiles)
        int rc;
        float smin, smax; bool violated = false;
        arm(hl);
        if ((rc = run_hist(h, opt, op, regs, cg, d, h->cls.as<int32_t>(), thresholds, T, hl, 1))) return rc;
        if ((rc = finish_hist(h, opt, stats, &smin, &smax, &violated))) return rc;
        if (smin == smin) return h->fail(FNB_ERR_NOT_NORMALIZED, "embeddings must be normalized to 1, range %.9g %.9g", smin, smax);
        return h->fail(FNB_ERR_NOT_NORMALIZED, "embeddings must be normalized to 1 (similarity out of range on another rank)");
    }

    // one pass in `mode` over the resident rows: operands, schedule, launch, range check, bins to the host, error certificate
    int pass(int mode, double* bound) {
        int rc;
        op = GramOperands();
        op.pairs = pick_pairs(&opt, cg, n);
        op.want_l16 = opt.strict_tiles >= 0;
        if ((rc = prepare_operand(h, mode, (const float*)de, perm_dev, n, d, false, op, opt.normalize))) return rc;
        if ((rc = self_b_maps(h, op, d))) return rc;
        opt.mode = op.mode;                              // AUTO resolved
        h->last_mode = op.mode; h->last_peak = op.peakedness;
        std::vector<RegionDev> regs;
        triangle_regions(n, (int)super_rows(), 0, regs);
        finish_regions(regs, tile, op.pairs, &shard, steal_world > 1);
        HistLaunch hl; hl.auto_window = true;
        arm(hl);
        if ((rc = run_hist(h, opt, op, regs, cg, d, h->cls.as<int32_t>(), thresholds, T, hl, 0))) return rc;
        float smin, smax; bool violated = false;
        if ((rc = collect(hl, bound, &violated, &smin, &smax))) return rc;
        if (violated) return not_normalized(hl, regs);
        return FNB_OK;
    }

    // Streamed pass: the rows arrive in class order, piece by piece (`feed(k, c0, c1)` makes rows [c0, c1) of h->stage_a ready on
    // the handle's stream -- an upload from the host, or a broadcast from the rank that holds them).  The pair matrix is cut into
    // column chunks at `bounds`; launch k covers the pairs (row < column, column in chunk k) and needs only the rows fed so far,
    // so the transfer of chunk k + 1 runs under launch k and only the first chunk's transfer is exposed.  The integer bins do
    // not depend on the chunking (512-aligned regions, same strict-tile rule).
    int pass_streamed(int mode, double* bound, const std::vector<long long>& bounds,
                      const std::function<int(int, long long, long long)>& feed) {
        int rc;
        op = GramOperands();
        op.pairs = pick_pairs(&opt, cg, n);
        op.want_l16 = opt.strict_tiles >= 0;
        if ((rc = prepare_operand(h, mode, h->stage_a.as<float>(), nullptr, n, d, false, op, opt.normalize, /*defer_split=*/true))) return rc;
        if ((rc = self_b_maps(h, op, d))) return rc;
        opt.mode = op.mode;
        h->last_mode = op.mode; h->last_peak = 0.f;
        const long long rr = super_rows();
        const int nchunks = (int)bounds.size() - 1;
        std::vector<std::vector<RegionDev>> chunks((size_t)nchunks);
        for (int k = 0; k < nchunks; ++k) {
            chunk_regions(n, rr, bounds[k], bounds[k + 1], chunks[k]);
            finish_regions(chunks[k], tile, op.pairs, &shard, steal_world > 1);
        }
        HistLaunch hl; hl.auto_window = true; hl.chunks = &chunks;
        arm(hl);
        hl.before_launch = [&](int k) -> int {
            const long long c0 = bounds[k], c1 = bounds[k + 1];
            int r = feed(k, c0, c1);
            if (r) return r;
            // Sharded jobs: the split follows the transfer on the COPY stream -- it runs on the SMs the launch in progress leaves
            // free (kAuxReserveSms) instead of sitting between two launches on the handle's stream.  One GPU: every SM is busy
            // under a launch, and a split parked on the copy stream would only hold back the next chunk's upload behind it.
            // (The last chunk also zeroes the padding rows.)
            cudaStream_t ss = (sharded && h->copy_stream) ? h->copy_stream : h->stream;
            if ((r = split_operand_rows(h, op, h->stage_a.as<float>(), stream_perm, n, d, opt.normalize, c0, k + 1 == nchunks ? op.a_rows_pad : c1, ss))) return r;
            if (ss != h->stream) {
                CK(cudaEventRecord(h->copy_ev[2], ss));
                CK(cudaStreamWaitEvent(h->stream, h->copy_ev[2], 0));
            }
            return FNB_OK;
        };
        std::vector<RegionDev> none;
        if ((rc = run_hist(h, opt, op, none, cg, d, h->cls.as<int32_t>(), thresholds, T, hl, 0))) return rc;
        de = h->stage_a.p; perm_dev = stream_perm;       // the rows are resident now: a later pass re-reads them from here
        const int launches = h->chunk_launches;
        float smin, smax; bool violated = false;
        if ((rc = collect(hl, bound, &violated, &smin, &smax))) return rc;
        if (stats) stats->streamed_chunks = launches;
        if (violated) {
            std::vector<RegionDev> regs;
            triangle_regions(n, (int)rr, 0, regs);
            finish_regions(regs, tile, op.pairs, &shard, steal_world > 1);
            HistLaunch whole; whole.auto_window = true;
            return not_normalized(whole, regs);
        }
        op.peakedness = h->last_peak;
        return FNB_OK;
    }

    // column-chunk boundaries of a streamed pass: multiples of a granule of about one super-row, growing with the work already
    // queued (1, 1, 1, 1, 2, 3, 4, 6, 9 ... granules): launch k is then at least as long as the transfer of chunk k + 1
    std::vector<long long> chunk_bounds() {
        op.pairs = pick_pairs(&opt, cg, n);
        long long g = super_rows();
        while (g > 49152) g /= 2;
        g = std::max<long long>(512, g / 512 * 512);
        std::vector<long long> bounds = chunk_schedule(n, g);
        return bounds;
    }

    // AUTO: FP16F8 when the a-priori gate and the a-posteriori certificate hold, else the strict pass over the resident rows.
    // `streamed_first` ran the first pass already (optimistically in FP16F8) and left the bound.
    int finish_auto(double bound, int* fallback) {
        bool again = requested == FNB_MODE_AUTO && op.mode == FNB_MODE_FP16F8 && (bound > (double)opt.eps || !(h->last_peak <= kAutoPeakLimit));
        if (reduce && requested == FNB_MODE_AUTO) {
            // the ranks must take the same branch (their certificates can differ in the last bits: float atomics): any rank's yes wins
            unsigned long long* flag = h->bins.as<unsigned long long>() + kMaxBins;
            const unsigned long long v = again ? 1ull : 0ull;
            CK(cudaMemcpyAsync(flag, &v, 8, cudaMemcpyHostToDevice, h->stream));
            int rc = comm_all_reduce_u64(h, flag, 1, true, h->stream);
            if (rc) return rc;
            unsigned long long got = 0;
            CK(cudaMemcpyAsync(&got, flag, 8, cudaMemcpyDeviceToHost, h->stream));
            CK(cudaStreamSynchronize(h->stream));
            again = got != 0;
        }
        if (again) {
            // the fast contraction cannot vouch for this data (e.g. many different-identity pairs at high similarity): strict pass
            int rc = pass(FNB_MODE_FP16X3, &bound);
            if (rc) return rc;
            if (stats) stats->kernel_launches += 1;      // the second split_rows
            *fallback = 1;
        }
        if (stats) {
            stats->kernel_launches += 3;      // labels_to_keys, boundary_flags, split_rows (+ the Gram kernel counted above)
            stats->n_pairs = sum_bins(host_bins, T + 1);
            stats->error_bound = (float)bound;
            stats->fallback = *fallback;
        }
        return FNB_OK;
    }

    int write_bins(const DLView& vb) {
        const size_t rb = (size_t)(T + 1) * 8;
        if (vb.on_device) {
            for (int r = 0; r < 2; ++r)
                CK(cudaMemcpyAsync((char*)vb.data + r * rb, h->bins.as<unsigned long long>() + (size_t)r * (kMaxBins + 1), rb, cudaMemcpyDeviceToDevice, h->stream));
        } else {
            memcpy(vb.data, host_bins, 2 * rb);
        }
        return FNB_OK;
    }
};

static bool labels_in_class_order(const DLView& vl, long long n) {
    if (vl.on_device) return false;
    if (vl.bits == 64) { const long long* l = (const long long*)vl.data; for (long long i = 1; i < n; ++i) if (l[i] < l[i - 1]) return false; }
    else { const int* l = (const int*)vl.data; for (long long i = 1; i < n; ++i) if (l[i] < l[i - 1]) return false; }
    return true;
}

static int check_hist_args(fnb_context* h, const fnb_options& opt, const DLTensor* emb, const DLTensor* labels, const double* thresholds, int T,
                           DLTensor* bins_out, DLView* ve, DLView* vl, DLView* vb)
{
    if (opt.metric != 0 && opt.metric != 1) return h->fail(FNB_ERR_BAD_METRIC, "Undefined similarity metric %d", opt.metric);
    if (!thresholds || T < 1 || T >= kMaxBins) return h->fail(FNB_ERR_INVALID, "number of thresholds must be in [1, %d]", kMaxBins - 1);
    int np; bool tf; int fmt, eb; float ps;
    if (mode_info(opt.mode, &np, &tf, &fmt, &eb, &ps)) return h->fail(FNB_ERR_INVALID, "bad mode %d", opt.mode);
    int rc = dl_view(h, emb, "embeddings", 2, 2, ve); if (rc) return rc;
    if ((rc = dl_check_embeddings(h, *ve, "embeddings"))) return rc;
    if ((rc = dl_view(h, labels, "labels", 1, 1, vl))) return rc;
    if (vl->code != kDLInt || (vl->bits != 32 && vl->bits != 64)) return h->fail(FNB_ERR_INVALID, "labels must be int32 or int64");
    if (vl->rows != ve->rows) return h->fail(FNB_ERR_INVALID, "len(labels) != embeddings.shape[0]");
    if ((rc = dl_view(h, bins_out, "bins_out", 2, 2, vb))) return rc;
    if (vb->bits != 64 || vb->rows != 2 || vb->cols != T + 1) return h->fail(FNB_ERR_INVALID, "bins_out must be a 64-bit integer [2, T+1] tensor");
    return FNB_OK;
}

}  // namespace

extern "C" int fnb_pair_histogram_bins(fnb_handle h, const DLTensor* emb, const DLTensor* labels,
                                       const double* thresholds, int T, const fnb_options* opt_in,
                                       DLTensor* bins_out, fnb_stats* stats)
{
    if (!h) return FNB_ERR_INVALID;
    WholeSetJob job;
    job.h = h; job.stats = stats; job.thresholds = thresholds; job.T = T;
    if (opt_in) job.opt = *opt_in; else fnb_default_options(&job.opt);
    fnb_options& opt = job.opt;
    if (opt.world < 1 || opt.rank < 0 || opt.rank >= opt.world) return h->fail(FNB_ERR_INVALID, "bad rank/world %d/%d", opt.rank, opt.world);
    CK(cudaSetDevice(h->device));
    if (stats) memset(stats, 0, sizeof(*stats));
    DLView ve, vl, vb;
    int rc = check_hist_args(h, opt, emb, labels, thresholds, T, bins_out, &ve, &vl, &vb);
    if (rc) return rc;
    const long long n = ve.rows;
    const int d = (int)ve.cols;
    job.n = n; job.d = d; job.row_bytes = (size_t)d * 4;

    if (n < 2) {
        CK(h->bins.ensure(2 * (kMaxBins + 1) * 8));
        CK(cudaMemsetAsync(h->bins.p, 0, 2 * (kMaxBins + 1) * 8, h->stream));
        memset(job.host_bins, 0, sizeof(job.host_bins));
        if ((rc = job.write_bins(vb))) return rc;
        CK(cudaStreamSynchronize(h->stream));
        return FNB_OK;
    }

    const void* dl = nullptr;
    CK(cudaEventRecord(h->ev[0], h->stream));
    h->last_h2d_bytes = 0; h->h2d_timed = false; h->h2d_timed_bytes = 0;
    // labels first: the class sort (np.unique ranks, statistics.py:68-79) runs on the stream while the embeddings are staged
    if ((rc = dl_to_device(h, vl, (size_t)n * (vl.bits / 8), h->stage_lab, &dl))) return rc;
    if ((rc = sort_labels(h, dl, vl.bits, n))) return rc;
    const bool host_emb = !ve.on_device;
    if (!host_emb) job.de = ve.data;                     // host rows: uploaded below, in one piece or chunk by chunk (streamed pass)
    job.perm_dev = h->perm.as<long long>();
    job.cg = pick_cta_group(&opt);
    job.tile = kRowsPerCta * job.cg;
    job.requested = opt.mode;
    if ((rc = shard_from_options(h, opt, &job.shard))) return rc;
    h->last_shard = job.shard.spec;

    double bound = 0.0;
    int fallback = 0;
    // Streamed pass (host embeddings, fnb_options.streamed): the upload is cut into column chunks of the pair matrix and the copy
    // of chunk k + 1 runs under launch k.  Rows out of class order: the host threads that fill the pinned ring GATHER them in
    // class order (they copy the bytes anyway), with the permutation the device sort produced -- the extra cost is its 8 n
    // byte copy back.
    const bool want_stream = host_emb && opt.streamed >= 0 && (opt.streamed > 0 || (size_t)n * d * 4 >= ((size_t)64 << 20));
    if (want_stream) {
        const long long* perm_h = nullptr;
        if (!labels_in_class_order(vl, n)) {
            CK(h->perm_host.ensure((size_t)n * 8));
            CK(cudaMemcpyAsync(h->perm_host.p, h->perm.p, (size_t)n * 8, cudaMemcpyDeviceToHost, h->stream));
            CK(cudaStreamSynchronize(h->stream));
            perm_h = h->perm_host.as<long long>();
        }
        CK(h->stage_a.ensure((size_t)n * d * 4));
        const char* src = (const char*)ve.data;
        const size_t rb = job.row_bytes;
        auto feed = [&](int k, long long c0, long long c1) -> int {
            return stage_chunk(h, h->stage_a.as<char>() + (size_t)c0 * rb, perm_h ? src : src + (size_t)c0 * rb, (size_t)(c1 - c0) * rb, k == 0,
                               perm_h ? perm_h + c0 : nullptr, rb);
        };
        // AUTO starts optimistically in FP16F8; the peakedness gate is evaluated when every chunk has been split (finish_auto)
        const int first = (job.requested == FNB_MODE_AUTO) ? (d % 128 == 0 ? FNB_MODE_FP16F8 : FNB_MODE_FP16X3) : job.requested;
        if ((rc = job.pass_streamed(first, &bound, job.chunk_bounds(), feed))) return rc;
    } else {
        if (host_emb && (rc = dl_to_device(h, ve, (size_t)n * d * 4, h->stage_a, &job.de))) return rc;
        if ((rc = job.pass(job.requested, &bound))) return rc;
    }
    // (a caller that shards by itself: every rank evaluates the same model on its share scaled to the whole; the shares are
    // interleaved row blocks, so the ranks agree except at the very edge of the bound)
    if ((rc = job.finish_auto(bound, &fallback))) return rc;
    // the bins are already on the host; hand them over (device output: one small copy on the stream)
    return job.write_bins(vb);
}


// cnt[k * world + r] = number of class positions j in chunk k (bounds[k] <= j < bounds[k + 1]) whose row perm[j] belongs to rank r
// (off[r] <= perm[j] < off[r + 1]); block-private counters in shared memory, one flush per block
__global__ void run_counts_kernel(const long long* __restrict__ perm, long long n, const long long* __restrict__ off, int world,
                                  const long long* __restrict__ bounds, int nchunks, unsigned int* __restrict__ cnt)
{
    extern __shared__ unsigned int s_cnt[];
    const int total = nchunks * world;
    for (int i = threadIdx.x; i < total; i += blockDim.x) s_cnt[i] = 0;
    __syncthreads();
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x) {
        const long long g = perm[j];
        int r = 0;
        while (r + 1 < world && g >= off[r + 1]) ++r;
        int k = 0;
        while (k + 1 < nchunks && j >= bounds[k + 1]) ++k;
        atomicAdd(&s_cnt[k * world + r], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < total; i += blockDim.x) if (s_cnt[i]) atomicAdd(cnt + i, s_cnt[i]);
}

// dst row i <- src row perm[i], i < rows (one warp per row; d is a multiple of 64 floats)
__global__ void __launch_bounds__(256)
gather_rows_kernel(const float* __restrict__ src, const long long* __restrict__ perm, long long rows, int d, float* __restrict__ dst)
{
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int vec = d >> 2;
    for (long long i = warp0; i < rows; i += nwarps) {
        const float4* s4 = reinterpret_cast<const float4*>(src + perm[i] * d);
        float4* d4 = reinterpret_cast<float4*>(dst + i * d);
        for (int c = lane; c < vec; c += 32) d4[c] = __ldg(s4 + c);
    }
}

extern "C" int fnb_pair_histogram_sharded(fnb_handle h, const DLTensor* emb_shard, const DLTensor* labels_shard,
                                          const double* thresholds, int T, const fnb_options* opt_in,
                                          DLTensor* bins_out, fnb_stats* stats)
{
    if (!h) return FNB_ERR_INVALID;
    if (!h->comm) return h->fail(FNB_ERR_INVALID, "fnb_pair_histogram_sharded needs a communicator (fnb_comm_init)");
    WholeSetJob job;
    job.h = h; job.stats = stats; job.thresholds = thresholds; job.T = T;
    if (opt_in) job.opt = *opt_in; else fnb_default_options(&job.opt);
    fnb_options& opt = job.opt;
    const int world = comm_world(h), rank = comm_rank(h);
    opt.rank = rank; opt.world = world;
    CK(cudaSetDevice(h->device));
    if (stats) memset(stats, 0, sizeof(*stats));
    DLView ve, vl, vb;
    int rc = check_hist_args(h, opt, emb_shard, labels_shard, thresholds, T, bins_out, &ve, &vl, &vb);
    if (rc) return rc;
    const long long n_local = ve.rows;
    const int d = (int)ve.cols;
    const size_t rb = (size_t)d * 4;
    job.d = d; job.row_bytes = rb; job.reduce = world > 1;
    CK(cudaEventRecord(h->ev[0], h->stream));
    h->last_h2d_bytes = 0; h->h2d_timed = false; h->h2d_timed_bytes = 0;
    float gather_ms = 0.f;

    // ---- 1. row counts of the ranks (shards may be ragged) and the dimension check
    std::vector<long long> off((size_t)world + 1, 0);
    {
        CK(h->comm_buf.ensure((size_t)(world + 1) * 16));
        long long mine[2] = {n_local, (long long)d};
        long long* dev = h->comm_buf.as<long long>();
        CK(cudaMemcpyAsync(dev + 2 * world, mine, 16, cudaMemcpyHostToDevice, h->stream));
        if ((rc = comm_all_gather(h, dev + 2 * world, dev, 16, h->stream))) return rc;
        std::vector<long long> all((size_t)2 * world);
        CK(cudaMemcpyAsync(all.data(), dev, (size_t)world * 16, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        for (int r = 0; r < world; ++r) {
            if (all[2 * r + 1] != d) return h->fail(FNB_ERR_INVALID, "rank %d has embedding dimension %lld, this rank %d", r, all[2 * r + 1], d);
            off[r + 1] = off[r] + all[2 * r];
        }
    }
    const long long n = off[world];
    job.n = n;
    if (n > (1LL << 30)) return h->fail(FNB_ERR_INVALID, "too many rows");
    if (n < 2) {
        CK(h->bins.ensure(2 * (kMaxBins + 1) * 8));
        CK(cudaMemsetAsync(h->bins.p, 0, 2 * (kMaxBins + 1) * 8, h->stream));
        memset(job.host_bins, 0, sizeof(job.host_bins));
        if ((rc = job.write_bins(vb))) return rc;
        CK(cudaStreamSynchronize(h->stream));
        return FNB_OK;
    }

    // ---- tile queues shared by the ranks (fnb_comm_init mapped every rank's counters into every rank): a rank drains its own queue,
    //      then helps the others (GramParams::steal_counter), so nobody waits in the all-reduce for the GPU that runs slowest
    //      under the power limit.  Every rank zeroes ITS counters here, in front of the label exchange -- no launch of any rank
    //      can start before every rank has contributed to that exchange.
    job.sharded = true;
    {
        int cap = 0;
        if (comm_shared_counters(h, rank, &cap) && world > 1 && world <= 8 && opt.tile_queue >= 0 && opt.shard_mod == 0 && opt.panel_window <= 0 && cap >= 256) {
            job.steal_world = world; job.steal_rank = rank;
            for (int v = 0; v < world; ++v) job.steal_ctr[v] = comm_shared_counters(h, v, nullptr);
            CK(cudaMemsetAsync(job.steal_ctr[rank], 0, (size_t)cap * 8, h->stream));
        }
    }

    // ---- 2. labels.  Every rank first puts ITS rows in class order (a local sort: the rows of a rank then form a sorted run),
    //         the sorted label runs are exchanged (8 n bytes), and every rank sorts the concatenation of the runs itself.  The global
    //         class order is then a MERGE of the ranks' runs: any range of it takes a CONTIGUOUS range of every run -- which is
    //         what lets the row exchange below be cut into column chunks whatever order the caller's rows came in.
    CK(h->lab_all.ensure((size_t)n * 8));
    long long* lab_all = h->lab_all.as<long long>();
    CK(h->local_perm.ensure((size_t)std::max<long long>(n_local, 1) * 8));
    if (n_local > 0) {
        const void* dl = nullptr;
        if ((rc = dl_to_device(h, vl, (size_t)n_local * (vl.bits / 8), h->stage_lab, &dl))) return rc;
        if ((rc = sort_labels(h, dl, vl.bits, n_local))) return rc;         // h->perm: local class order, h->keys_out: sorted labels
        CK(cudaMemcpyAsync(h->local_perm.p, h->perm.p, (size_t)n_local * 8, cudaMemcpyDeviceToDevice, h->stream));
        CK(cudaMemcpyAsync(lab_all + off[rank], h->keys_out.p, (size_t)n_local * 8, cudaMemcpyDeviceToDevice, h->stream));
    }
    if (world > 1) {
        if ((rc = comm_group_start(h))) return rc;
        for (int r = 0; r < world; ++r)
            if (off[r + 1] > off[r] && (rc = comm_broadcast(h, lab_all + off[r], lab_all + off[r], (size_t)(off[r + 1] - off[r]) * 8, r, h->stream))) return rc;
        if ((rc = comm_group_end(h))) return rc;
    }
    if ((rc = sort_labels(h, lab_all, 64, n))) return rc;                   // h->perm: class position -> row of the concatenated runs

    job.cg = pick_cta_group(&opt);
    job.tile = kRowsPerCta * job.cg;
    job.requested = opt.mode;
    if ((rc = shard_from_options(h, opt, &job.shard))) return rc;
    h->last_shard = job.shard.spec;

    // column chunks of the pair matrix (class positions) and, per chunk, how many rows of every rank's run it takes
    const bool want_stream = opt.streamed >= 0 && world > 1;
    std::vector<long long> bounds = want_stream ? job.chunk_bounds() : std::vector<long long>{0, n};
    const int nchunks = (int)bounds.size() - 1;
    std::vector<long long> run_pos((size_t)(nchunks + 1) * world, 0);      // run_pos[k * world + r]: rows of run r in front of chunk k
    {
        std::vector<long long> meta(off.begin(), off.end());
        meta.insert(meta.end(), bounds.begin(), bounds.end());
        const size_t cnt_bytes = (size_t)nchunks * world * 4;
        CK(h->comm_buf.ensure(meta.size() * 8 + cnt_bytes + 64));
        long long* meta_dev = h->comm_buf.as<long long>();
        unsigned int* cnt_dev = reinterpret_cast<unsigned int*>(meta_dev + meta.size());
        CK(cudaMemcpyAsync(meta_dev, meta.data(), meta.size() * 8, cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemsetAsync(cnt_dev, 0, cnt_bytes, h->stream));
        run_counts_kernel<<<(unsigned)std::min<long long>((n + 255) / 256, 148LL * 8), 256, cnt_bytes, h->stream>>>(
            h->perm.as<long long>(), n, meta_dev, world, meta_dev + world + 1, nchunks, cnt_dev);
        CK(cudaGetLastError());
        std::vector<unsigned int> cnt((size_t)nchunks * world);
        CK(cudaMemcpyAsync(cnt.data(), cnt_dev, cnt_bytes, cudaMemcpyDeviceToHost, h->stream));
        if (!ve.on_device && n_local > 0) {
            // host rows: the copy threads gather this rank's run while they fill the pinned ring (fnb_stage.cu)
            CK(h->perm_host.ensure((size_t)n_local * 8));
            CK(cudaMemcpyAsync(h->perm_host.p, h->local_perm.p, (size_t)n_local * 8, cudaMemcpyDeviceToHost, h->stream));
        }
        CK(cudaStreamSynchronize(h->stream));              // meta and cnt are local vectors: consumed here
        for (int k = 0; k < nchunks; ++k)
            for (int r = 0; r < world; ++r) run_pos[(size_t)(k + 1) * world + r] = run_pos[(size_t)k * world + r] + cnt[(size_t)k * world + r];
    }

    // ---- 3. the rows.  Every rank needs all of them (the columns of its tiles); they travel as fp32 (4 bytes per element; the
    //         split operands would be 6) on the copy stream, run by run: rank r's part of a chunk is gathered out of r's tensor in
    //         class order (device rows: a gather kernel; host rows: the copy threads + the pinned ring) straight into its place
    //         in the array of concatenated runs, and broadcast in place from there.
    CK(h->stage_a.ensure((size_t)n * rb));
    char* runs = h->stage_a.as<char>();
    const char* local = (const char*)ve.data;
    bool first_chunk = true;
    // Option (fnb_options.streamed == 3), PINNED host rows only: the whole shard is copied as it lies by the DMA engine alone (one
    // stream-ordered copy on the copy stream, no host thread touches the bytes) and the class-order gather runs on the device like
    // for device-resident rows.  Not the default: the whole shard then sits in front of the FIRST launch (2 GPUs, 1 GB per rank:
    // e2e 1,192 against 1,236 G pairs/s with the chunk-wise host gather); it is meant for hosts with very few cores per GPU.
    bool bulk_upload = false;
    const bool local_in_order = !ve.on_device && labels_in_class_order(vl, n_local);
    if (!ve.on_device && n_local > 0) {
        cudaPointerAttributes attr;
        const cudaError_t pe = cudaPointerGetAttributes(&attr, ve.data);
        if (pe != cudaSuccess) cudaGetLastError();
        bulk_upload = (pe == cudaSuccess) && attr.type == cudaMemoryTypeHost && opt.streamed == 3;
        if (bulk_upload) CK(h->stage_b.ensure((size_t)n_local * rb));
    }
    // every run's part of chunk k: this rank's part is gathered into place, then all parts are broadcast in place by their owners
    // (one NCCL group: the broadcasts of a chunk run side by side); the handle's stream waits for the chunk
    auto move_chunk = [&](int k) -> int {
        int r2;
        if (!h->copy_stream) CK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        if (first_chunk) {
            // the array may still be read by work queued earlier on the handle's stream; the local permutation is ready there too
            CK(cudaEventRecord(h->copy_ev[0], h->stream));
            CK(cudaStreamWaitEvent(h->copy_stream, h->copy_ev[0], 0));
            CK(cudaEventRecord(h->copy_ev[1], h->copy_stream));
            h->h2d_timed = true;
            first_chunk = false;
            if (bulk_upload) {
                CK(cudaMemcpyAsync(h->stage_b.p, ve.data, (size_t)n_local * rb, cudaMemcpyHostToDevice, h->copy_stream));
                h->last_h2d_bytes += (size_t)n_local * rb;
                local = h->stage_b.as<char>();
            }
        }
        const bool rows_on_device = ve.on_device || bulk_upload;
        const bool trace = getenv("FNB_TRACE") != nullptr;
        if (trace) {
            while (h->xchg_ev.size() < 3 * (size_t)(k + 1)) { cudaEvent_t e = nullptr; CK(cudaEventCreate(&e)); h->xchg_ev.push_back(e); }
            CK(cudaEventRecord(h->xchg_ev[3 * k], h->copy_stream));
        }
        const long long s0 = run_pos[(size_t)k * world + rank], s1 = run_pos[(size_t)(k + 1) * world + rank];
        char* own_dst = runs + (size_t)(off[rank] + s0) * rb;
        // host rows: the upload (copy engine, no SM) starts as soon as the copy stream gets to it
        if (s1 > s0 && !rows_on_device) {
            // rows already in class order (labels non-decreasing): the run IS the shard -- no gather; pinned memory then goes by DMA alone
            r2 = local_in_order ? stage_chunk(h, own_dst, local + (size_t)s0 * rb, (size_t)(s1 - s0) * rb, false)
                                : stage_chunk(h, own_dst, local, (size_t)(s1 - s0) * rb, false, h->perm_host.as<long long>() + s0, rb);
            if (r2) return r2;
        }
        if (k >= 1 && (size_t)(2 * k - 1) < h->chunk_ev.size()) {
            // The KERNELS of chunk k's exchange run while launch k - 1 runs -- and not earlier: they start when launch k - 1 is
            // about to start (the event in front of it).  A broadcast or gather kernel that holds SMs at the moment a launch starts
            // keeps some of its clusters from becoming resident (measured at 2 GPUs with the exchange running ahead of a static
            // schedule: one rank's launches 17 % longer).  Started behind the launch, the exchange only ever gets the SMs the
            // launch leaves free.
            CK(cudaStreamWaitEvent(h->copy_stream, h->chunk_ev[2 * (k - 1)], 0));
        }
        if (s1 > s0 && rows_on_device) {
            gather_rows_kernel<<<(unsigned)std::min<long long>((s1 - s0 + 7) / 8, 148LL * 8), 256, 0, h->copy_stream>>>(
                (const float*)local, h->local_perm.as<long long>() + s0, s1 - s0, d, (float*)own_dst);
            CK(cudaGetLastError());
        }
        if (trace) CK(cudaEventRecord(h->xchg_ev[3 * k + 1], h->copy_stream));
        if (world > 1) {
            if ((r2 = comm_group_start(h))) return r2;
            for (int r = 0; r < world; ++r) {
                const long long a = run_pos[(size_t)k * world + r], b = run_pos[(size_t)(k + 1) * world + r];
                char* dst = runs + (size_t)(off[r] + a) * rb;
                if (b > a && (r2 = comm_broadcast(h, dst, dst, (size_t)(b - a) * rb, r, h->copy_stream))) return r2;
            }
            if ((r2 = comm_group_end(h))) return r2;
        }
        if (trace) CK(cudaEventRecord(h->xchg_ev[3 * k + 2], h->copy_stream));
        CK(cudaEventRecord(h->copy_ev[2], h->copy_stream));
        CK(cudaStreamWaitEvent(h->stream, h->copy_ev[2], 0));
        return FNB_OK;
    };

    double bound = 0.0;
    int fallback = 0;
    job.stream_perm = h->perm.as<long long>();
    if (want_stream) {
        // chunk k + 1 travels on the copy stream while launch k runs
        auto feed = [&](int k, long long, long long) -> int { return move_chunk(k); };
        const int first = (job.requested == FNB_MODE_AUTO) ? (d % 128 == 0 ? FNB_MODE_FP16F8 : FNB_MODE_FP16X3) : job.requested;
        if ((rc = job.pass_streamed(first, &bound, bounds, feed))) return rc;
    } else {
        // streaming switched off (or one rank): everything first, then one launch
        if ((rc = move_chunk(0))) return rc;
        job.de = runs;
        job.perm_dev = h->perm.as<long long>();
        if ((rc = job.pass(job.requested, &bound))) return rc;
    }
    if ((rc = job.finish_auto(bound, &fallback))) return rc;
    if (stats) {
        if (cudaEventElapsedTime(&gather_ms, h->copy_ev[1], h->copy_ev[2]) == cudaSuccess) stats->gather_ms = gather_ms;
        stats->h2d_ms = 0.f;
    }
    return job.write_bins(vb);
}

extern "C" int fnb_counts_from_bins(const double* thresholds, int T, const fnb_options* opt_in, const uint64_t* bins,
                                    uint64_t* same_lt, uint64_t* diff_lt, uint64_t* n_same, uint64_t* n_diff)
{
    fnb_options opt; if (opt_in) opt = *opt_in; else fnb_default_options(&opt);
    if (!thresholds || !bins || T < 1 || T >= kMaxBins) return FNB_ERR_INVALID;
    if (opt.metric != 0 && opt.metric != 1) return FNB_ERR_BAD_METRIC;
    CutTables ct;
    if (build_cut_tables(thresholds, T, opt.metric, opt.eps, opt.cuts, &ct)) return FNB_ERR_INVALID;
    const uint64_t* all = bins; const uint64_t* same = bins + (T + 1);
    // suffix sums: pairs whose similarity is >= the pos-th smallest cut
    uint64_t suf_all[kMaxBins + 2], suf_same[kMaxBins + 2];
    suf_all[T + 1] = suf_same[T + 1] = 0;
    for (int k = T; k >= 0; --k) { suf_all[k] = suf_all[k + 1] + all[k]; suf_same[k] = suf_same[k + 1] + same[k]; }
    for (int n = 0; n < T; ++n) {
        const int pos = ct.pos[n];
        const uint64_t a = pos <= T ? suf_all[pos] : 0, s = pos <= T ? suf_same[pos] : 0;
        if (same_lt) same_lt[n] = s;
        if (diff_lt) diff_lt[n] = a - s;
    }
    if (n_same) *n_same = suf_same[0];
    if (n_diff) *n_diff = suf_all[0] - suf_same[0];
    return FNB_OK;
}

extern "C" int fnb_pair_histogram(fnb_handle h, const DLTensor* emb, const DLTensor* labels, const double* thresholds, int T,
                                  const fnb_options* opt, uint64_t* same_lt, uint64_t* diff_lt, uint64_t* n_same,
                                  uint64_t* n_diff, fnb_stats* stats)
{
    if (!h) return FNB_ERR_INVALID;
    if (T < 1 || T >= kMaxBins) return h->fail(FNB_ERR_INVALID, "number of thresholds must be in [1, %d]", kMaxBins - 1);
    std::vector<uint64_t> bins(2 * (size_t)(T + 1));
    int64_t shape[2] = {2, T + 1};
    DLTensor t = {};
    t.data = bins.data(); t.device.device_type = kDLCPU; t.ndim = 2; t.dtype.code = kDLUInt; t.dtype.bits = 64; t.dtype.lanes = 1;
    t.shape = shape;
    int rc = fnb_pair_histogram_bins(h, emb, labels, thresholds, T, opt, &t, stats);
    if (rc) return rc;
    rc = fnb_counts_from_bins(thresholds, T, opt, bins.data(), same_lt, diff_lt, n_same, n_diff);
    if (rc) return h->fail(rc, "fnb_counts_from_bins failed");
    return FNB_OK;
}

extern "C" int fnb_region_histogram_bins(fnb_handle h, const DLTensor* emb, const int64_t* perm, const int32_t* cls,
                                         const fnb_region* regions, int nregions, int nkeys,
                                         const double* thresholds, int T, const fnb_options* opt_in,
                                         uint64_t* bins_host, fnb_stats* stats)
{
    if (!h) return FNB_ERR_INVALID;
    fnb_options opt; if (opt_in) opt = *opt_in; else fnb_default_options(&opt);
    if (opt.metric != 0 && opt.metric != 1) return h->fail(FNB_ERR_BAD_METRIC, "Undefined similarity metric %d", opt.metric);
    if (!thresholds || T < 1 || T >= kMaxBins) return h->fail(FNB_ERR_INVALID, "number of thresholds must be in [1, %d]", kMaxBins - 1);
    if (!perm || !cls || !regions || nregions < 0 || nkeys < 1 || !bins_host) return h->fail(FNB_ERR_INVALID, "NULL / empty argument");
    if (opt.world < 1 || opt.rank < 0 || opt.rank >= opt.world) return h->fail(FNB_ERR_INVALID, "bad rank/world");
    CK(cudaSetDevice(h->device));
    if (stats) memset(stats, 0, sizeof(*stats));
    GramOperands op;
    if (mode_info(opt.mode, &op.num_pass, &op.tf32, &op.fmt, &op.elem_bytes, &op.prescale)) return h->fail(FNB_ERR_INVALID, "bad mode %d", opt.mode);
    DLView ve;
    int rc = dl_view(h, emb, "embeddings", 2, 2, &ve); if (rc) return rc;
    if ((rc = dl_check_embeddings(h, ve, "embeddings"))) return rc;
    const long long n_all = ve.rows;
    if (opt.subset_rows < 0 || opt.subset_rows > n_all) return h->fail(FNB_ERR_INVALID, "subset_rows %d out of range", opt.subset_rows);
    const long long n = opt.subset_rows > 0 ? opt.subset_rows : n_all;      // rows of the (permuted) subset
    for (long long i = 0; i < n; ++i)
        if (perm[i] < 0 || perm[i] >= n_all) return h->fail(FNB_ERR_INVALID, "perm[%lld] = %lld out of range", i, (long long)perm[i]);
    const int d = (int)ve.cols;
    const size_t out_bytes = (size_t)nkeys * 2 * (T + 1) * 8;
    memset(bins_host, 0, out_bytes);
    const int cg = pick_cta_group(&opt);
    const int tile = kRowsPerCta * cg;
    op.pairs = pick_pairs(&opt, cg);
    op.want_l16 = opt.strict_tiles >= 0;
    std::vector<RegionDev> regs;
    for (int i = 0; i < nregions; ++i) {
        const fnb_region& r = regions[i];
        if (r.row_begin < 0 || r.row_end > n || r.col_begin < 0 || r.col_end > n || r.key < 0 || r.key >= nkeys)
            return h->fail(FNB_ERR_INVALID, "region %d out of range", i);
        if (r.row_end <= r.row_begin || r.col_end <= r.col_begin) continue;
        if (r.tri && (r.row_begin != r.col_begin || r.row_end != r.col_end)) return h->fail(FNB_ERR_INVALID, "region %d: tri needs row range == col range", i);
        RegionDev rd = {};
        rd.row_begin = r.row_begin; rd.row_end = r.row_end; rd.col_begin = r.col_begin; rd.col_end = r.col_end;
        rd.tri = r.tri == 2 ? 2 : (r.tri ? 1 : 0); rd.key = r.key;
        if (rd.tri == 2 && r.key + 1 >= nkeys) return h->fail(FNB_ERR_INVALID, "region %d: tri == 2 bins the diagonal into slot key + 1", i);
        regs.push_back(rd);
    }
    ShardHost shard;
    if ((rc = shard_from_options(h, opt, &shard))) return rc;
    h->last_shard = shard.spec;
    finish_regions(regs, tile, op.pairs, &shard);
    if (regs.back().tile_begin == 0 || n < 1) { h->last_nkeys = 0; return FNB_OK; }

    const void* de = nullptr;
    CK(cudaEventRecord(h->ev[0], h->stream));
    h->last_h2d_bytes = 0; h->h2d_timed = false; h->h2d_timed_bytes = 0;
    if ((rc = dl_to_device(h, ve, (size_t)n_all * d * 4, h->stage_a, &de))) return rc;
    CK(h->perm.ensure(n * 8));
    CK(h->cls.ensure(n * 4));
    CK(cudaMemcpyAsync(h->perm.p, perm, n * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->cls.p, cls, n * 4, cudaMemcpyHostToDevice, h->stream));
    if ((rc = prepare_operand(h, opt.mode, (const float*)de, h->perm.as<long long>(), n, d, false, op, opt.normalize))) return rc;
    if ((rc = self_b_maps(h, op, d))) return rc;
    opt.mode = op.mode;                                  // AUTO resolved
    h->last_mode = op.mode; h->last_peak = op.peakedness;

    HistLaunch hl; hl.nkeys = nkeys;
    if ((rc = run_hist(h, opt, op, regs, cg, d, h->cls.as<int32_t>(), thresholds, T, hl, 0))) return rc;
    float smin, smax; bool violated = false;
    if ((rc = finish_hist(h, opt, stats, &smin, &smax, &violated))) return rc;
    if (violated) {
        if ((rc = run_hist(h, opt, op, regs, cg, d, h->cls.as<int32_t>(), thresholds, T, hl, 1))) return rc;
        if ((rc = finish_hist(h, opt, stats, &smin, &smax, &violated))) return rc;
        return h->fail(FNB_ERR_NOT_NORMALIZED, "embeddings must be normalized to 1, range %.9g %.9g", smin, smax);
    }
    // compact [nkeys][2][stride] -> [nkeys][2][T+1]
    CK(h->pinned.ensure((size_t)nkeys * 2 * hl.stride * 8 + 8192));
    uint64_t* stage = (uint64_t*)((char*)h->pinned.p + 8192);
    CK(cudaMemcpyAsync(stage, h->bins.p, (size_t)nkeys * 2 * hl.stride * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    uint64_t total = 0;
    uint64_t sums[2][kMaxBins + 1] = {};
    for (int k = 0; k < nkeys * 2; ++k) {
        memcpy(bins_host + (size_t)k * (T + 1), stage + (size_t)k * hl.stride, (size_t)(T + 1) * 8);
        if ((k & 1) == 0) total += sum_bins(stage + (size_t)k * hl.stride, T + 1);
        for (int b = 0; b <= T; ++b) sums[k & 1][b] += stage[(size_t)k * hl.stride + b];
    }
    if (stats) {
        stats->n_pairs = total;
        stats->error_bound = (float)error_certificate(opt, op.mode, h->last_strict != 0, d, h->last_peak_mean, h->last_peak, n, hl.ct, sums[0], sums[1], T);
    }
    return FNB_OK;
}

