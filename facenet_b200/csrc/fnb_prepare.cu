// facenet_b200 -- operand preparation: class ranking of labels (replaces np.unique(labels) /
// split_embeddings, facenet/statistics.py:68-79) and the row gather + precision split that turns
// fp32 embeddings into the tensor-core operand arrays.  HBM-bound streaming kernels.
#include "fnb_host.h"
#include <cub/cub.cuh>
#include <cuda_fp8.h>

namespace fnb {

__device__ __forceinline__ float tf32_rn(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

// One warp converts one row (lanes stride over 8-element groups); rows [n, n_pad) are written as zeros.
// The warp also forms the squared norm of the row; the maximum over rows (ordered-uint atomicMax) lets the
// Gram kernel skip the per-pair range check of statistics.py:40-42 on interior tiles (|s_ab| <= |a| |b|).
template <int kMode>
__global__ void __launch_bounds__(256)
split_rows_kernel(const float* __restrict__ x, const long long* __restrict__ perm, long long n, long long n_pad, int d,
                  void* __restrict__ hi, void* __restrict__ lo, void* __restrict__ h8, unsigned int* __restrict__ norm_max_ord,
                  int normalize, float* __restrict__ row_nrm, void* __restrict__ l16, long long row_begin, long long row_end)
{
    const int vec_per_row = d >> 3;
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    float norm_max = 0.f, peak_max = 0.f, peak_sum = 0.f;
    // rows [row_begin, row_end) of the padded operand arrays (a chunk of a streamed launch, or all of [0, n_pad))
    for (long long row = row_begin + warp0; row < row_end; row += nwarps) {
        const long long src = (row < n) ? (perm ? perm[row] : row) : 0;
        // normalise-on-load (fnb_options.normalize): a first pass over the row forms |x|; the second pass below re-reads
        // the row (L1/L2 hit) and scales every element before the split.
        //   1: x / |x|  (np.linalg.norm: sqrt of the fp32 sum of squares; faceclass.py:57-64)
        //   2: x * rsqrt(max(sum x^2, 1e-10))  (tf.nn.l2_normalize, inception_resnet_v1.py:491-492)
        float inv_or_nrm = 1.f;
        if (normalize) {
            float ss = 0.f;
            if (row < n)
                for (int c8 = lane; c8 < vec_per_row; c8 += 32) {
                    const float4* s4 = reinterpret_cast<const float4*>(x + src * d + (long long)c8 * 8);
                    const float4 p0 = __ldg(s4), p1 = __ldg(s4 + 1);
                    ss += p0.x * p0.x + p0.y * p0.y + p0.z * p0.z + p0.w * p0.w + p1.x * p1.x + p1.y * p1.y + p1.z * p1.z + p1.w * p1.w;
                }
#pragma unroll
            for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
            const float nr = __fsqrt_rn(ss);
            inv_or_nrm = (normalize == 1) ? nr : __frsqrt_rn(fmaxf(ss, 1.0e-10f));
            if (row_nrm && lane == 0) row_nrm[row] = (row < n) ? nr : 0.f;
        }
        float nrm = 0.f, x4 = 0.f;
        for (int c8 = lane; c8 < vec_per_row; c8 += 32) {
            float f[8];
            if (row < n) {
                const float4* s4 = reinterpret_cast<const float4*>(x + src * d + (long long)c8 * 8);
                const float4 p0 = __ldg(s4), p1 = __ldg(s4 + 1);
                f[0] = p0.x; f[1] = p0.y; f[2] = p0.z; f[3] = p0.w; f[4] = p1.x; f[5] = p1.y; f[6] = p1.z; f[7] = p1.w;
                if (normalize == 1) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) f[i] = __fdiv_rn(f[i], inv_or_nrm);
                } else if (normalize == 2) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) f[i] = __fmul_rn(f[i], inv_or_nrm);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] = 0.f;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) { const float sq = f[i] * f[i]; nrm += sq; x4 = fmaf(sq, sq, x4); }
            const long long o = row * d + (long long)c8 * 8;
            if (kMode == FNB_MODE_FP16X3 || kMode == FNB_MODE_FP16) {
                const float pre = (kMode == FNB_MODE_FP16X3) ? 256.0f : 1.0f;
                __half hh[8], ll[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float xs = f[i] * pre;
                    hh[i] = __float2half_rn(xs);
                    ll[i] = __float2half_rn(xs - __half2float(hh[i]));
                }
                *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(hi) + o) = *reinterpret_cast<uint4*>(hh);
                if (kMode == FNB_MODE_FP16X3)
                    *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(lo) + o) = *reinterpret_cast<uint4*>(ll);
            } else if (kMode == FNB_MODE_FP16F8) {
                // x * 2^12 = hi + lo, hi fp16; the cross terms hi*lo are formed from e4m3(x * 2^8) and e4m3(lo * 2^4):
                // both stay in the normal range of e4m3 for |x| >= 6e-5 and the product keeps the 2^24 scale of hi*hi
                // l16 (optional) = the same low part in fp16: the operand of the strict fp16x3 tiles of a histogram launch
                __half hh[8], ll[8];
                __align__(8) __nv_fp8_storage_t q8[8], l8[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float xs = f[i] * 4096.0f;
                    hh[i] = __float2half_rn(xs);
                    const float lo_f = xs - __half2float(hh[i]);
                    ll[i] = __float2half_rn(lo_f);
                    q8[i] = __nv_cvt_float_to_fp8(xs * 0.0625f, __NV_SATFINITE, __NV_E4M3);
                    l8[i] = __nv_cvt_float_to_fp8(lo_f * 16.0f, __NV_SATFINITE, __NV_E4M3);
                }
                *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(hi) + o) = *reinterpret_cast<uint4*>(hh);
                if (l16) *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(l16) + o) = *reinterpret_cast<uint4*>(ll);
                *reinterpret_cast<uint2*>(reinterpret_cast<uint8_t*>(h8) + o) = *reinterpret_cast<uint2*>(q8);
                *reinterpret_cast<uint2*>(reinterpret_cast<uint8_t*>(lo) + o) = *reinterpret_cast<uint2*>(l8);
            } else if (kMode == FNB_MODE_BF16) {
                __nv_bfloat16 hh[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) hh[i] = __float2bfloat16_rn(f[i]);
                *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(hi) + o) = *reinterpret_cast<uint4*>(hh);
            } else {
                float hh[8], ll[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) { hh[i] = tf32_rn(f[i]); ll[i] = tf32_rn(f[i] - hh[i]); }
                float4* dh = reinterpret_cast<float4*>(reinterpret_cast<float*>(hi) + o);
                dh[0] = make_float4(hh[0], hh[1], hh[2], hh[3]);
                dh[1] = make_float4(hh[4], hh[5], hh[6], hh[7]);
                if (kMode == FNB_MODE_TF32X3) {
                    float4* dl = reinterpret_cast<float4*>(reinterpret_cast<float*>(lo) + o);
                    dl[0] = make_float4(ll[0], ll[1], ll[2], ll[3]);
                    dl[1] = make_float4(ll[4], ll[5], ll[6], ll[7]);
                }
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) { nrm += __shfl_xor_sync(0xffffffffu, nrm, o); x4 += __shfl_xor_sync(0xffffffffu, x4, o); }
        // NaN rows must not hide: an unordered compare keeps them as "too large"
        norm_max = (nrm <= norm_max) ? norm_max : nrm;
        const float pk = (nrm > 0.f) ? x4 / (nrm * nrm) : 0.f;
        peak_max = (pk <= peak_max) ? peak_max : pk;
        if (row < n && pk == pk) peak_sum += pk;
    }
    if (norm_max_ord && lane == 0) {
        const float v = (norm_max == norm_max) ? norm_max : INFINITY;
        if (v > 0.f) atomicMax(norm_max_ord, float_to_ordered(v));
        const float w = (peak_max == peak_max) ? peak_max : INFINITY;
        if (w > 0.f) atomicMax(norm_max_ord + 1, float_to_ordered(w));
        if (peak_sum > 0.f) atomicAdd(reinterpret_cast<float*>(norm_max_ord + 2), peak_sum);      // DeviceScalars::peak_sum
    }
}

cudaError_t launch_split_rows(int mode, const float* x, const long long* perm, long long n, long long n_pad, int d,
                              void* hi, void* lo, void* h8, unsigned int* norm_max_ord, cudaStream_t s,
                              int normalize, float* row_nrm, void* l16, long long row_begin, long long row_end)
{
    if (row_end < 0) row_end = n_pad;
    if (row_end <= row_begin) return cudaSuccess;
    const int threads = 256;
    long long blocks = (row_end - row_begin + 7) / 8;    // 8 warps (rows) per block
    if (blocks > 148LL * 16) blocks = 148LL * 16;
    switch (mode) {
        case FNB_MODE_FP16X3: split_rows_kernel<FNB_MODE_FP16X3><<<(unsigned)blocks, threads, 0, s>>>(x, perm, n, n_pad, d, hi, lo, h8, norm_max_ord, normalize, row_nrm, l16, row_begin, row_end); break;
        case FNB_MODE_TF32X3: split_rows_kernel<FNB_MODE_TF32X3><<<(unsigned)blocks, threads, 0, s>>>(x, perm, n, n_pad, d, hi, lo, h8, norm_max_ord, normalize, row_nrm, l16, row_begin, row_end); break;
        case FNB_MODE_TF32:   split_rows_kernel<FNB_MODE_TF32><<<(unsigned)blocks, threads, 0, s>>>(x, perm, n, n_pad, d, hi, lo, h8, norm_max_ord, normalize, row_nrm, l16, row_begin, row_end); break;
        case FNB_MODE_BF16:   split_rows_kernel<FNB_MODE_BF16><<<(unsigned)blocks, threads, 0, s>>>(x, perm, n, n_pad, d, hi, lo, h8, norm_max_ord, normalize, row_nrm, l16, row_begin, row_end); break;
        case FNB_MODE_FP16:   split_rows_kernel<FNB_MODE_FP16><<<(unsigned)blocks, threads, 0, s>>>(x, perm, n, n_pad, d, hi, lo, h8, norm_max_ord, normalize, row_nrm, l16, row_begin, row_end); break;
        case FNB_MODE_FP16F8: split_rows_kernel<FNB_MODE_FP16F8><<<(unsigned)blocks, threads, 0, s>>>(x, perm, n, n_pad, d, hi, lo, h8, norm_max_ord, normalize, row_nrm, l16, row_begin, row_end); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

// One thread per 512 x 512 block of the (upper triangle of the) pair matrix of a class-sorted set of n rows: the block is strict
// iff any of its tile x tile tiles is ragged (crosses n), touches the diagonal, or can hold a same-identity pair (the class
// ranks of its corner rows overlap).  bits must be zeroed before the launch.
__global__ void strict_blocks_kernel(const int32_t* __restrict__ cls, int n, int tile, int nb, unsigned int* __restrict__ bits)
{
    const long long total = (long long)nb * nb;
    for (long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x; id < total; id += (long long)gridDim.x * blockDim.x) {
        const int br = (int)(id / nb), bc = (int)(id - (long long)br * nb);
        if (bc < br) continue;                                       // below the diagonal: never scheduled
        const int R = br << 9, C = bc << 9;
        bool strict = false;
        for (int i = 0; i < 512 / tile && !strict; ++i) {
            for (int j = 0; j < 512 / tile && !strict; ++j) {
                const int r0 = R + i * tile, c0 = C + j * tile;
                if (r0 >= n || c0 >= n || c0 + tile - 1 <= r0) continue;           // outside the set / below the diagonal: never binned
                if (r0 + tile > n || c0 + tile > n || c0 <= r0 + tile - 1) strict = true;
                else strict = (cls[r0] <= cls[c0 + tile - 1]) && (cls[c0] <= cls[r0 + tile - 1]);
            }
        }
        if (strict) atomicOr(bits + (id >> 5), 1u << (id & 31));
    }
}

cudaError_t launch_strict_blocks(const int32_t* cls, int n, int tile, int nb, unsigned int* bits, cudaStream_t s) {
    const long long total = (long long)nb * nb;
    const unsigned blocks = (unsigned)std::min<long long>((total + 255) / 256, 148LL * 16);
    strict_blocks_kernel<<<blocks, 256, 0, s>>>(cls, n, tile, nb, bits);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// labels -> (perm, cls): rows ordered by label value; cls = rank of the label value

template <typename L>
__global__ void labels_to_keys_kernel(const L* __restrict__ labels, long long n, long long* keys, long long* idx) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        keys[i] = (long long)labels[i];
        idx[i] = i;
    }
}

__global__ void boundary_flags_kernel(const long long* __restrict__ sorted_keys, long long n, int* flags) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        flags[i] = (i > 0 && sorted_keys[i] != sorted_keys[i - 1]) ? 1 : 0;
}

int sort_labels(fnb_context* h, const void* labels_dev, int label_bits, long long n)
{
#define CKS(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return h->fail(FNB_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); } while (0)
    CKS(h->keys_in.ensure(n * 8));
    CKS(h->keys_out.ensure(n * 8));
    CKS(h->vals_in.ensure(n * 8));
    CKS(h->perm.ensure(n * 8));
    CKS(h->flags.ensure(n * 4));
    CKS(h->cls.ensure(n * 4));
    const int threads = 256;
    unsigned blocks = (unsigned)std::min<long long>((n + threads - 1) / threads, 148LL * 16);
    if (label_bits == 64)
        labels_to_keys_kernel<long long><<<blocks, threads, 0, h->stream>>>((const long long*)labels_dev, n, h->keys_in.as<long long>(), h->vals_in.as<long long>());
    else
        labels_to_keys_kernel<int><<<blocks, threads, 0, h->stream>>>((const int*)labels_dev, n, h->keys_in.as<long long>(), h->vals_in.as<long long>());
    CKS(cudaGetLastError());
    size_t tmp1 = 0, tmp2 = 0;
    CKS(cub::DeviceRadixSort::SortPairs(nullptr, tmp1, h->keys_in.as<long long>(), h->keys_out.as<long long>(),
                                        h->vals_in.as<long long>(), h->perm.as<long long>(), (int)n, 0, 64, h->stream));
    CKS(cub::DeviceScan::InclusiveSum(nullptr, tmp2, h->flags.as<int>(), h->cls.as<int>(), (int)n, h->stream));
    CKS(h->cub_tmp.ensure(std::max(tmp1, tmp2)));
    size_t cap = h->cub_tmp.cap;
    CKS(cub::DeviceRadixSort::SortPairs(h->cub_tmp.p, cap, h->keys_in.as<long long>(), h->keys_out.as<long long>(),
                                        h->vals_in.as<long long>(), h->perm.as<long long>(), (int)n, 0, 64, h->stream));
    boundary_flags_kernel<<<blocks, threads, 0, h->stream>>>(h->keys_out.as<long long>(), n, h->flags.as<int>());
    CKS(cudaGetLastError());
    cap = h->cub_tmp.cap;
    CKS(cub::DeviceScan::InclusiveSum(h->cub_tmp.p, cap, h->flags.as<int>(), h->cls.as<int>(), (int)n, h->stream));
#undef CKS
    return FNB_OK;
}

}  // namespace fnb
