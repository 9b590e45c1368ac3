// facenet_b200 -- weighted binary cross entropy of the pair classifier over one P x K batch
// (facenet/apps/train_classifier.py:60-84 behind facenet/faceclass.py:23-27).
//
//   fnb_pair_cross_entropy    embeddings -> loss and d loss / d(alpha, threshold, theta) in ONE Gram launch: the B x B logits
//                             never reach HBM (the reference gathers B(B-1)/2 logits with tf.gather_nd, :75)
//   fnb_logits_cross_entropy  the reference's own signature: a materialised [B, B] logits matrix -> loss
#include "fnb_host.h"

#include <math.h>
#include <string.h>

using namespace fnb;

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return h->fail(FNB_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); } while (0)

namespace fnb {

// strict upper triangle of a row-major [b, ld] logits matrix; label 1 iff i / k_per == j / k_per (train_classifier.py:66-73)
__global__ void __launch_bounds__(256)
logits_bce_kernel(const float* __restrict__ logits, long long ld, int b, int k_per, float pos_weight, double* __restrict__ out)
{
    double acc = 0.0;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int i = warp; i < b; i += nwarps) {
        const int gi = i / k_per;
        float part = 0.f;
        for (int j = i + 1 + lane; j < b; j += 32) {
            const float x = __ldg(logits + (long long)i * ld + j);
            const float z = (j / k_per == gi) ? 1.0f : 0.0f;
            const float lw = 1.0f + (pos_weight - 1.0f) * z;
            part += (1.0f - z) * x + lw * (log1pf(expf(-fabsf(x))) + fmaxf(-x, 0.0f));
        }
        acc += (double)part;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0 && acc != 0.0) atomicAdd(out, acc);
}

}  // namespace fnb

static int check_batch(fnb_context* h, long long b, int k_per, double* pos_weight, double* n_pairs) {
    if (k_per < 1 || b < 2 || b % k_per != 0)
        return h->fail(FNB_ERR_INVALID, "batch of %lld rows is not P x K with K = %d", b, k_per);
    const double total = 0.5 * (double)b * (double)(b - 1);
    const double pos = 0.5 * (double)b * (double)(k_per - 1);
    if (pos <= 0) return h->fail(FNB_ERR_INVALID, "K = 1: no same-class pair, pos_weight = len(labels) / sum(labels) - 1 is undefined");
    *pos_weight = total / pos - 1.0;                     // train_classifier.py:73
    *n_pairs = total;
    return FNB_OK;
}

extern "C" int fnb_pair_cross_entropy(fnb_handle h, const DLTensor* batch, int examples_per_class, float alpha, float threshold,
                                      const fnb_options* opt_in, double* out, fnb_stats* stats)
{
    if (!h) return FNB_ERR_INVALID;
    if (!out) return h->fail(FNB_ERR_INVALID, "NULL out");
    fnb_options opt; if (opt_in) opt = *opt_in; else fnb_default_options(&opt);
    CK(cudaSetDevice(h->device));
    if (stats) memset(stats, 0, sizeof(*stats));
    GramOperands op;
    opt.mode = FNB_MODE_FP16X3;
    if (mode_info(opt.mode, &op.num_pass, &op.tf32, &op.fmt, &op.elem_bytes, &op.prescale)) return h->fail(FNB_ERR_INVALID, "bad mode");
    DLView ve;
    int rc = dl_view(h, batch, "batch", 2, 2, &ve); if (rc) return rc;
    if ((rc = dl_check_embeddings(h, ve, "batch"))) return rc;
    const long long b = ve.rows;
    const int d = (int)ve.cols;
    double pos_weight = 0, n_pairs = 0;
    if ((rc = check_batch(h, b, examples_per_class, &pos_weight, &n_pairs))) return rc;

    const void* de = nullptr;
    CK(cudaEventRecord(h->ev[0], h->stream));
    if ((rc = dl_to_device(h, ve, (size_t)b * d * 4, h->stage_a, &de))) return rc;
    if ((rc = prepare_operand(h, opt.mode, (const float*)de, nullptr, b, d, false, op, opt.normalize))) return rc;
    if ((rc = self_b_maps(h, op, d))) return rc;
    // group of a row = row / K (rows grouped by class, facenet/facenet.py:108-113)
    CK(h->pinned.ensure((size_t)b * 4 + 8192));
    int32_t* grp = reinterpret_cast<int32_t*>((char*)h->pinned.p + 8192);
    for (long long i = 0; i < b; ++i) grp[i] = (int32_t)(i / examples_per_class);
    CK(h->cls.ensure((size_t)b * 4));
    CK(cudaMemcpyAsync(h->cls.p, grp, (size_t)b * 4, cudaMemcpyHostToDevice, h->stream));
    CK(h->scan.ensure(64));
    CK(cudaMemsetAsync(h->scan.p, 0, 32, h->stream));

    const int cg = 1;                                    // 128 x 128 tiles fill the SMs better at batch sizes of a few thousand
    const int tile = kRowsPerCta * cg;
    std::vector<RegionDev> regs;
    RegionDev r = {}; r.row_end = (int)b; r.col_end = (int)b; r.tri = 1; regs.push_back(r);
    finish_regions(regs, tile);
    if ((rc = upload_regions(h, regs))) return rc;
    if ((rc = reset_scalars(h))) return rc;
    GramParams p = {};
    p.regions = h->regions.as<RegionDev>(); p.nregions = 1; p.total_tiles = regs.back().tile_begin;
    p.shard = ShardSpec{1, 0, 1, nullptr};
    p.kblocks = d / (128 / op.elem_bytes);
    p.acc_scale = 1.0f / (op.prescale * op.prescale);
    p.operand_fmt = op.fmt;
    DeviceScalars* sc = h->counters.as<DeviceScalars>();
    p.counters = sc->counters; p.range_ord = sc->range_ord;
    p.row_cls = h->cls.as<int32_t>(); p.col_cls = p.row_cls;
    p.n_rows = (int)b; p.n_cols = (int)b;
    p.raw = 1;
    if (opt.normalize == 1 && opt.theta != 0.f) { p.row_nrm = op.a_nrm; p.col_nrm = op.b_nrm; p.theta = opt.theta; }
    p.bce_alpha = alpha; p.bce_threshold = threshold; p.bce_pos_weight = (float)pos_weight;
    p.bce_out = h->scan.as<double>();
    if ((rc = upload_bias(h, op.mode, d, opt.bias_correction < 0, &p.bias_beta))) return rc;
    CK(cudaEventRecord(h->ev[1], h->stream));
    if ((rc = launch_gram(h, cg, EPI_BCE, opt.max_ctas, op, p, 0))) return rc;
    CK(cudaEventRecord(h->ev[2], h->stream));
    double* host = reinterpret_cast<double*>(h->pinned.p);
    CK(cudaMemcpyAsync(host, h->scan.p, 32, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < 4; ++i) out[i] = host[i] / n_pairs;     // tf.reduce_mean (train_classifier.py:82)
    out[4] = pos_weight;
    if (stats) {
        float ms = 0.f, pm = 0.f;
        cudaEventElapsedTime(&ms, h->ev[1], h->ev[2]);
        cudaEventElapsedTime(&pm, h->ev[0], h->ev[1]);
        stats->kernel_ms = ms; stats->prepare_ms = pm;
        stats->n_pairs = (uint64_t)n_pairs;
        stats->kernel_launches = 2;                      // split_rows, gram<BCE>
        stats->grid_ctas = (uint32_t)h->last_grid;
        stats->mode_used = FNB_MODE_FP16X3;
    }
    return FNB_OK;
}

extern "C" int fnb_logits_cross_entropy(fnb_handle h, const DLTensor* logits, int examples_per_class, double* loss)
{
    if (!h) return FNB_ERR_INVALID;
    if (!loss) return h->fail(FNB_ERR_INVALID, "NULL loss");
    CK(cudaSetDevice(h->device));
    DLView v;
    int rc = dl_view(h, logits, "logits", 2, 2, &v); if (rc) return rc;
    if (v.code != kDLFloat || v.bits != 32 || v.rows != v.cols) return h->fail(FNB_ERR_INVALID, "logits must be a square float32 matrix");
    const long long b = v.rows;
    double pos_weight = 0, n_pairs = 0;
    if ((rc = check_batch(h, b, examples_per_class, &pos_weight, &n_pairs))) return rc;
    const void* dl = nullptr;
    if ((rc = dl_to_device(h, v, (size_t)b * b * 4, h->stage_a, &dl))) return rc;
    CK(h->scan.ensure(64));
    CK(cudaMemsetAsync(h->scan.p, 0, 8, h->stream));
    const unsigned blocks = (unsigned)std::min<long long>((b + 7) / 8, 148LL * 8);
    logits_bce_kernel<<<blocks, 256, 0, h->stream>>>((const float*)dl, b, (int)b, examples_per_class, (float)pos_weight, h->scan.as<double>());
    CK(cudaGetLastError());
    CK(h->pinned.ensure(4096));
    CK(cudaMemcpyAsync(h->pinned.p, h->scan.p, 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    *loss = *reinterpret_cast<double*>(h->pinned.p) / n_pairs;
    return FNB_OK;
}
