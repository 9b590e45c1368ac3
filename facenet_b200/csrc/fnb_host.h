// facenet_b200 -- host-side internals shared by the C-ABI translation units.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdarg.h>
#include <string>
#include <vector>

#include "../../include/facenet_b200.h"
#include "fnb_gram.cuh"

namespace fnb {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct HostBuf {   // pinned
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// similarity cuts derived from the distance thresholds (host)
struct CutTables {
    int T = 0, T_fin = 0;
    float cuts[kMaxBins];          // ascending, +inf padded
    float wlo[kMaxBins + 4];
    float whi[kMaxBins + 4];
    int order[kMaxBins];           // sorted position j -> threshold index
    int pos[kMaxBins];             // threshold n -> number of sorted cuts <= cut_n
    int uniform = 0;
    double e0 = 0, h = 0, dev = 0; // arithmetic-progression fit of the finite cuts (raw similarity of the launch's arithmetic)
    double e0x = 0, hx = 0, devx = 0;  // the same for the fp16x3 arithmetic of strict tiles
};

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace fnb

namespace fnb { struct HostCopier; void destroy_copier(HostCopier*); struct Comm; }

struct fnb_context {
    int device = 0;
    int sm_count = 0;
    int cc_major = 0, cc_minor = 0;
    size_t total_mem = 0;
    cudaStream_t stream = nullptr;       // stream in use
    cudaStream_t own_stream = nullptr;   // created by fnb_create
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    std::vector<cudaEvent_t> chunk_ev;   // streamed histogram launches: one (start, end) pair per launch
    std::vector<cudaEvent_t> xchg_ev;    // FNB_TRACE: three events per chunk of a sharded row exchange
    int aux_launches = 0;                // auxiliary Gram launches (launch_gram_aux) of the last pass
    int chunk_launches = 0;              // launches of the last streamed pass (0: one launch timed by ev[1] .. ev[2])
    fnb::PFN_tmapEncodeTiled encode = nullptr;
    std::string err;

    fnb::DevBuf stage_a, stage_b, stage_lab;          // H2D staging of kDLCPU inputs
    fnb::DevBuf a_hi, a_lo, b_hi, b_lo, a_h8, b_h8;   // split / converted operands
    fnb::DevBuf a_l16, b_l16;                         // fp16f8 mode: fp16 low parts (strict tiles)
    fnb::DevBuf a_nrm, b_nrm;                         // row norms before normalise-on-load (fnb_options.normalize)
    fnb::DevBuf shard_slots;                          // residues of fnb_options.shard_slots on the device
    fnb::DevBuf tile_counter;                         // one 64-bit tile-queue counter per launch of a pass (GramParams::tile_counter)
    cudaStream_t aux_stream = nullptr;                // second Gram launch on the SMs the main grid leaves free (launch_gram_aux)
    cudaEvent_t aux_ev[2] = {nullptr, nullptr};
    fnb::DevBuf progress;                             // per-cluster column-panel progress (GramParams::sync_window)
    fnb::DevBuf perm, cls, keys_in, keys_out, vals_in, flags, cub_tmp;
    fnb::DevBuf regions, tables, bins, counters, out, strip, mine_out, scan, select_io;
    fnb::DevBuf strict_bits;                          // one bit per 512 x 512 block of the pair matrix (fp16f8 strict tiles)
    fnb::DevBuf bias_tab;                             // [2][kBiasStride] accumulation-bias knots of the current (mode, d)
    int bias_mode = -1, bias_d = 0;
    fnb::DevBuf mine_lab, mine_keys, mine_status;     // mining: labels as int64, packed arg-extrema keys, status words
    long long mine_rows = 0, mine_b = 0, mine_ld = 0; // geometry of the last mining call (its strips stay in `strip`)
    int mine_kmax = 0; float mine_atol = 0.f; bool mine_has_strip = false;
    fnb::HostBuf pinned;
    fnb::HostBuf perm_host;              // class-order permutation on the host (streamed upload of rows that are not in class order)
    // pipelined staging of pageable host tensors (fnb_stage.cu)
    cudaStream_t copy_stream = nullptr;
    fnb::HostBuf ring;
    cudaEvent_t ring_ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t copy_ev[3] = {nullptr, nullptr, nullptr};
    fnb::HostCopier* copier = nullptr;
    fnb::Comm* comm = nullptr;           // NCCL communicator of a sharded job (fnb_comm_init, fnb_comm.cu)
    fnb::DevBuf comm_buf, lab_all, local_perm;   // sharded jobs: row counts of the ranks / chunk metadata, all labels as int64, class order of this rank's rows
    int ring_next = 0;                   // next ring slot of a chunked upload (stage_chunk)
    size_t last_h2d_bytes = 0;           // host -> device bytes of the call in progress
    size_t h2d_timed_bytes = 0;          // size of the copy those events bracket
    bool h2d_timed = false;              // copy_ev[1] .. copy_ev[2] bracket a staged copy of the call in progress
    int hist_grid[5] = {0, 0, 0, 0, 0};  // CTAs of the last histogram launch per cluster_pairs (co-resident clusters x cluster size)
    int last_nkeys = 0, last_T = 0, last_grid = 0, last_mode = 0, last_window = 0, last_strict = 0;
    float last_peak = 0.f, last_peak_mean = 0.f;
    long long last_rows = 0;             // rows of the last prepared A operand (mean peakedness = peak_sum / rows)
    fnb::ShardSpec last_shard = fnb::ShardSpec{1, 0, 1, nullptr};   // share of the launch being prepared
    double last_eps_counted = 0;         // distance half-width of the near-threshold window counted by interior tiles

    int fail(int code, const char* fmt, ...);
};

namespace fnb {

// tensors crossing the C ABI (fnb_api.cu)
struct DLView {
    void* data = nullptr;
    bool on_device = false;
    long long rows = 0, cols = 1;
    int bits = 0, code = 0;
};

// operand arrays of one Gram launch: TMA maps over the split / converted embeddings
struct GramOperands {
    CUtensorMap a_hi, a_lo, b_hi, b_lo;
    CUtensorMap a_h8, b_h8;        // fp16f8 mode only: e4m3(x) arrays (a_lo / b_lo then hold e4m3(lo))
    CUtensorMap a_l16, b_l16;      // fp16f8 mode only: fp16 low parts for the strict (fp16x3) tiles of a histogram launch
    bool have_l16 = false;         // a_l16 / b_l16 were prepared (prepare_operand with want_l16 in fp16f8 mode)
    bool want_l16 = false;         // set by the histogram entry points before prepare_operand
    int num_pass = 3; bool tf32 = false; int fmt = 0; int elem_bytes = 2; float prescale = 1.f;
    int mode = 0;                  // FNB_MODE_* the operands were prepared for (AUTO resolved)
    float peakedness = 0.f;        // max_row sum x^4 / (sum x^2)^2 (valid after prepare_operand in AUTO mode)
    int pairs = 1;                 // CTA pairs per cluster (2: the A maps carry 64-row boxes, see gram_kernel kPairs)
    long long a_rows_pad = 0;      // padded row count of the prepared A-side arrays
    const float* a_nrm = nullptr;  // norms of the rows before normalise-on-load (NULL without it)
    const float* b_nrm = nullptr;
};

struct DeviceScalars {      // layout of fnb_context::counters
    unsigned long long counters[2];
    unsigned int range_ord[4];
    unsigned int norm_max_ord;      // ordered-uint max squared row norm of the last prepared operand (not reset per launch)
    unsigned int peak_max_ord;      // ordered-uint max over rows of sum x^4 / (sum x^2)^2 (same lifetime)
    float peak_sum;                 // sum over rows of the same quantity (mean peakedness = peak_sum / rows)
    unsigned int pad[5];
};

// fnb_api.cu
int dl_view(fnb_context* h, const DLTensor* t, const char* name, int want_ndim_min, int want_ndim_max, DLView* v);
int dl_to_device(fnb_context* h, const DLView& v, size_t bytes, DevBuf& stage, const void** out);
int stage_to_device(fnb_context* h, void* dst, const void* src, size_t bytes);   // fnb_stage.cu
// one chunk of a streamed upload on the copy stream (the handle's stream waits for it; `first` orders the copy stream after
// the work already queued on the handle's stream and starts the h2d timing)
// perm != NULL: a gather -- dst row i <- src + perm[i] * row_bytes (src is then the BASE of the host array, dst the chunk's place)
int stage_chunk(fnb_context* h, void* dst, const void* src, size_t bytes, bool first, const long long* perm = nullptr, size_t row_bytes = 0);
int dl_check_embeddings(fnb_context* h, const DLView& v, const char* name);
int prepare_operand(fnb_context* h, int mode, const float* x, const long long* perm, long long n, int d,
                    bool side_b, GramOperands& op, int normalize = 0, bool defer_split = false);
int split_operand_rows(fnb_context* h, const GramOperands& op, const float* x, const long long* perm, long long n, int d,
                       int normalize, long long row_begin, long long row_end, cudaStream_t stream = nullptr);
// host view of a rank's share: the residues it owns (ascending) and the device spec
struct ShardHost {
    ShardSpec spec = ShardSpec{1, 0, 1, nullptr};
    std::vector<int> residues = {0};
    int owned(int nrb) const {                                 // number of owned row blocks among [0, nrb)
        int c = (nrb / spec.mod) * spec.width;
        const int rem = nrb % spec.mod;
        for (int r : residues) c += (r < rem) ? 1 : 0;
        return c;
    }
};
void finish_regions(std::vector<RegionDev>& regs, int tile, int pairs = 1, const ShardHost* shard = nullptr, bool pad_equal = false);
int shard_from_options(fnb_context* h, const fnb_options& opt, ShardHost* out);
int self_b_maps(fnb_context* h, GramOperands& op, int d);   // B side = the prepared A side (Gram of a set with itself)
int upload_regions(fnb_context* h, const std::vector<RegionDev>& regs);
int reset_scalars(fnb_context* h);
int mode_info(int mode, int* num_pass, bool* tf32, int* fmt, int* elem_bytes, float* prescale);
inline float gram_acc_scale(const GramOperands& op) { return 1.0f / (op.prescale * op.prescale); }
int build_cut_tables(const double* thresholds, int T, int metric, double eps, const float* cuts_override, CutTables* out,
                     const float* beta_knots = nullptr, const float* beta_knots_strict = nullptr);
void bias_table(int mode, int d, float* knots);
double mode_sigma_s(int mode, int d, double abs_s, double peakedness);
int upload_bias(fnb_context* h, int mode, int d, bool strict_x3, const float** dev);

// fnb_prepare.cu
cudaError_t launch_split_rows(int mode, const float* x, const long long* perm, long long n, long long n_pad, int d,
                              void* hi, void* lo, void* h8, unsigned int* norm_max_ord, cudaStream_t s,   // norm_max_ord[1] = peakedness
                              int normalize = 0, float* row_nrm = nullptr, void* l16 = nullptr,
                              long long row_begin = 0, long long row_end = -1);     // rows of [0, n_pad) to write (-1: n_pad)
cudaError_t launch_strict_blocks(const int32_t* cls, int n, int tile, int nb, unsigned int* bits, cudaStream_t s);
int sort_labels(fnb_context* h, const void* labels_dev, int label_bits, long long n);   // fills h->perm (i64), h->cls (i32)

// fnb_comm.cu (NCCL, resolved at run time)
int comm_world(const fnb_context* h);
int comm_rank(const fnb_context* h);
int comm_all_gather(fnb_context* h, const void* send, void* recv, size_t bytes_per_rank, cudaStream_t s);
int comm_broadcast(fnb_context* h, const void* send, void* recv, size_t bytes, int root, cudaStream_t s);
int comm_group_start(fnb_context* h);
int comm_group_end(fnb_context* h);
int comm_all_reduce_u64(fnb_context* h, void* buf, size_t count, bool max_op, cudaStream_t s);
void comm_release(fnb_context* h);
unsigned long long* comm_shared_counters(const fnb_context* h, int r, int* count);   // rank r's counters; NULL: the ranks do not share their queues

// fnb_gram.cu
int launch_gram(fnb_context* h, int cta_group, int epi, int max_ctas, const GramOperands& op, GramParams& p, size_t hist_bytes);
// on h->aux_stream; reserve_sms of the free SMs are left alone
int launch_gram_aux(fnb_context* h, const GramOperands& op, const GramParams& p_main, size_t hist_bytes, int reserve_sms = 0);
constexpr int kAuxReserveSms = 8;    // SMs a sharded job keeps free for its row exchange (= the CTA cap of its NCCL communicator)
size_t gram_smem_bytes(int num_slots, size_t hist_bytes);
int gram_pick_slots(size_t hist_bytes);

}  // namespace fnb
