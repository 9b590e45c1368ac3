// facenet_b200 -- instantiations and launcher of the Gram kernel (see fnb_gram.cuh).
#include "fnb_host.h"

#include <algorithm>

namespace fnb {

static constexpr size_t kSmemLimit = 232448;   // 227 KB per CTA on sm_100

size_t gram_smem_bytes(int num_slots, size_t hist_bytes) {
    return 1024 + (size_t)num_slots * kSlotBytes + sizeof(GramSmemMisc) + hist_bytes;
}

int gram_pick_slots(size_t hist_bytes) {
    int s = kMaxSlots;
    while (s > 2 && gram_smem_bytes(s, hist_bytes) > kSmemLimit) --s;
    return s;
}

template <int kCtaGroup, int kNumPass, bool kTf32, int kEpi, int kPairs = 1>
static int launch_one(fnb_context* h, int max_ctas, const GramOperands& op, const GramParams& p, size_t smem)
{
    constexpr int kCluster = kCtaGroup * kPairs;
    auto kern = gram_kernel<kCtaGroup, kNumPass, kTf32, kEpi, kPairs, 1>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return h->fail(FNB_ERR_CUDA, "cudaFuncSetAttribute(smem=%zu): %s", smem, cudaGetErrorString(e));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(h->sm_count - h->sm_count % kCluster));
    cfg.blockDim = dim3(kGramThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = h->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kCluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    // persistent kernel, one CTA per SM: the grid is the number of CTAs that are co-resident.  Clusters larger than a
    // TPC cannot use every SM (GPC boundaries), so ask the driver how many clusters fit.
    int ctas = h->sm_count;
    if (kCluster > 2) {
        int nclusters = 0;
        e = cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg);
        if (e != cudaSuccess) return h->fail(FNB_ERR_CUDA, "cudaOccupancyMaxActiveClusters: %s", cudaGetErrorString(e));
        ctas = std::min(ctas, nclusters * kCluster);
    }
    if (max_ctas > 0 && max_ctas < ctas) ctas = max_ctas;
    ctas -= ctas % kCluster;
    if (ctas < kCluster) ctas = kCluster;
    cfg.gridDim = dim3((unsigned)ctas);
    h->last_grid = ctas;
    if (kEpi == EPI_HIST) h->hist_grid[kPairs] = ctas;
    e = cudaLaunchKernelEx(&cfg, kern, op.a_hi, op.a_lo, op.b_hi, op.b_lo, op.a_h8, op.b_h8, op.a_l16, op.b_l16, p);
    if (e != cudaSuccess) return h->fail(FNB_ERR_CUDA, "gram kernel launch: %s", cudaGetErrorString(e));
    return FNB_OK;
}

// HIST with two CTA pairs per cluster (A operand multicast)
static int launch_mode_pairs2(fnb_context* h, int max_ctas, const GramOperands& op, const GramParams& p, size_t smem)
{
    if (op.num_pass == 2 && !op.tf32) return launch_one<2, 2, false, EPI_HIST, 2>(h, max_ctas, op, p, smem);
    if (op.num_pass == 3 && !op.tf32) return launch_one<2, 3, false, EPI_HIST, 2>(h, max_ctas, op, p, smem);
    if (op.num_pass == 3 && op.tf32)  return launch_one<2, 3, true,  EPI_HIST, 2>(h, max_ctas, op, p, smem);
    if (op.num_pass == 1 && !op.tf32) return launch_one<2, 1, false, EPI_HIST, 2>(h, max_ctas, op, p, smem);
    if (op.num_pass == 1 && op.tf32)  return launch_one<2, 1, true,  EPI_HIST, 2>(h, max_ctas, op, p, smem);
    return h->fail(FNB_ERR_INVALID, "no kernel for num_pass=%d tf32=%d", op.num_pass, (int)op.tf32);
}

template <int kCtaGroup, int kEpi>
static int launch_mode(fnb_context* h, int max_ctas, const GramOperands& op, const GramParams& p, size_t smem)
{
    if (op.num_pass == 2 && !op.tf32) return launch_one<kCtaGroup, 2, false, kEpi>(h, max_ctas, op, p, smem);
    if (op.num_pass == 3 && !op.tf32) return launch_one<kCtaGroup, 3, false, kEpi>(h, max_ctas, op, p, smem);
    if (op.num_pass == 3 && op.tf32)  return launch_one<kCtaGroup, 3, true,  kEpi>(h, max_ctas, op, p, smem);
    if (op.num_pass == 1 && !op.tf32) return launch_one<kCtaGroup, 1, false, kEpi>(h, max_ctas, op, p, smem);
    if (op.num_pass == 1 && op.tf32)  return launch_one<kCtaGroup, 1, true, kEpi>(h, max_ctas, op, p, smem);
    return h->fail(FNB_ERR_INVALID, "no kernel for num_pass=%d tf32=%d", op.num_pass, (int)op.tf32);
}

// The SMs a grid of two-pair clusters leaves free (16 of 148: four-CTA clusters do not tile the GPCs) run plain CTA pairs on the
// SAME tile queue: a second launch on the handle's auxiliary stream, started behind the main launch (an event in front of
// it), that walks the main launch's 256 x 512 super-tiles two tiles at a time.  Self Gram products only (the full-height
// B-side maps serve as A maps too).  The caller orders the handle's stream after it (aux_done).
template <int kNumPass>
static int launch_aux_pairs(fnb_context* h, int ctas, const GramOperands& op, const GramParams& p, size_t smem)
{
    auto kern = gram_kernel<2, kNumPass, false, EPI_HIST, 1, 2>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return h->fail(FNB_ERR_CUDA, "cudaFuncSetAttribute(smem=%zu): %s", smem, cudaGetErrorString(e));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)ctas);
    cfg.blockDim = dim3(kGramThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = h->aux_stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, kern, op.b_hi, op.b_lo, op.b_hi, op.b_lo, op.b_h8, op.b_h8, op.b_l16, op.b_l16, p);
    if (e != cudaSuccess) return h->fail(FNB_ERR_CUDA, "gram kernel launch (auxiliary pairs): %s", cudaGetErrorString(e));
    return FNB_OK;
}

int launch_gram_aux(fnb_context* h, const GramOperands& op, const GramParams& p_main, size_t hist_bytes, int reserve_sms)
{
    const int ctas = std::max(0, h->sm_count - h->last_grid - reserve_sms) & ~1;
    if (ctas < 2 || op.pairs != 2 || op.tf32 || p_main.tile_counter == nullptr) return FNB_OK;     // (the tf32 kinds keep to the main grid)
    GramParams p = p_main;
    p.sync_window = 0;
    const size_t smem = gram_smem_bytes(p.num_slots, hist_bytes);
    if (!h->aux_stream) return h->fail(FNB_ERR_INVALID, "auxiliary stream missing");
    int rc = op.num_pass == 2 ? launch_aux_pairs<2>(h, ctas, op, p, smem)
           : op.num_pass == 3 ? launch_aux_pairs<3>(h, ctas, op, p, smem) : launch_aux_pairs<1>(h, ctas, op, p, smem);   // 1: fp16 / bf16 single pass
    if (rc) return rc;
    h->last_grid += ctas;
    return FNB_OK;
}

int launch_gram(fnb_context* h, int cta_group, int epi, int max_ctas, const GramOperands& op, GramParams& p, size_t hist_bytes)
{
    // distance epilogues stage a 32 x 33 float tile per epilogue warp (the transpose in front of the coalesced stores)
    if (epi == EPI_PAIRWISE || epi == EPI_ROWSTRIP) hist_bytes = (size_t)kEpiWarps * 32 * 33 * 4;
    p.num_slots = gram_pick_slots(hist_bytes);
    const size_t smem = gram_smem_bytes(p.num_slots, hist_bytes);
    if (smem > kSmemLimit) return h->fail(FNB_ERR_UNSUPPORTED, "shared memory budget exceeded (%zu bytes)", smem);
    if (op.num_pass == 2 && (p.kblocks & 1)) return h->fail(FNB_ERR_UNSUPPORTED, "fp16f8 mode needs an embedding dimension that is a multiple of 128");
    if (op.pairs == 2 || op.pairs == 4) {
        if (cta_group != 2 || epi != EPI_HIST) return h->fail(FNB_ERR_INVALID, "cluster pairs need cta_group 2 and the histogram epilogue");
        if (op.pairs == 2) return launch_mode_pairs2(h, max_ctas, op, p, smem);
        // 2 x 2 pair grids (8-CTA clusters): the two modes the headline workloads run in
        if (op.num_pass == 2 && !op.tf32) return launch_one<2, 2, false, EPI_HIST, 4>(h, max_ctas, op, p, smem);
        if (op.num_pass == 3 && !op.tf32) return launch_one<2, 3, false, EPI_HIST, 4>(h, max_ctas, op, p, smem);
        return h->fail(FNB_ERR_UNSUPPORTED, "cluster_pairs = 4 is built for the fp16f8 and fp16x3 modes");
    }
    if (epi == EPI_FILTER) {
        if (op.num_pass != 3 || op.tf32 || cta_group != 2) return h->fail(FNB_ERR_INVALID, "the false-pair filter runs in fp16x3 mode on CTA pairs");
        return launch_one<2, 3, false, EPI_FILTER>(h, max_ctas, op, p, smem);
    }
    if (epi == EPI_BCE) {
        // one batch of a few thousand rows: the strict fp32-equivalent split only (the loss feeds an optimiser)
        if (op.num_pass != 3 || op.tf32) return h->fail(FNB_ERR_INVALID, "the cross-entropy epilogue runs in fp16x3 mode");
        return cta_group == 2 ? launch_one<2, 3, false, EPI_BCE>(h, max_ctas, op, p, smem)
                              : launch_one<1, 3, false, EPI_BCE>(h, max_ctas, op, p, smem);
    }
    if (cta_group == 2) {
        if (epi == EPI_HIST)     return launch_mode<2, EPI_HIST>(h, max_ctas, op, p, smem);
        if (epi == EPI_PAIRWISE) return launch_mode<2, EPI_PAIRWISE>(h, max_ctas, op, p, smem);
        return launch_mode<2, EPI_ROWSTRIP>(h, max_ctas, op, p, smem);
    }
    if (epi == EPI_HIST)     return launch_mode<1, EPI_HIST>(h, max_ctas, op, p, smem);
    if (epi == EPI_PAIRWISE) return launch_mode<1, EPI_PAIRWISE>(h, max_ctas, op, p, smem);
    return launch_mode<1, EPI_ROWSTRIP>(h, max_ctas, op, p, smem);
}

}  // namespace fnb
