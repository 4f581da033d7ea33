// loss.cu -- the statistics of nnU-Net's DC_and_CE_loss for one deep-supervision scale in ONE pass over the logits, and
// its gradient in one more (reference training/loss/compound_losses.py DC_and_CE_loss, dice.py:58-112
// MemoryEfficientSoftDiceLoss with softmax, robust_ce_loss.py; nnUNetTrainer.py:833-863 calls it once per scale).
// Off SURVEY's named path but inside every timed train step: as torch ops it is ~25 launches per scale (fp32 copy of the
// logits, softmax, one-hot scatter, product, three reductions, log-softmax, nll, and their backward kernels; 0.9 ms of
// device time per step at config 3, tools/glue_sites.py).
//   forward : per pixel p = softmax(l);  stats[b][k] = (sum p_k [t = k], sum p_k, sum [t = k]),  ce = sum (-log p_t)
//   backward: dl_j = p_j (G_j - sum_k p_k G_k) + g_ce (p_j - [t = j]),  G_k = g_inter[b][k] [t = k] + g_pred[b][k]
// The dice formula itself (a few hundred numbers) stays in torch on the stats, so that batch-dice, the DDP gather of the
// statistics and the smoothing constants keep the reference's code path.  Logits are addressed through
// (batch, class, pixel) strides: NCHW and channels_last heads are both read in place; fp32 or bf16.
#include <cuda_bf16.h>

#include "common.cuh"

namespace mlagg {

template <typename T>
__device__ __forceinline__ float ls_ld(const T *p);
template <>
__device__ __forceinline__ float ls_ld<float>(const float *p) { return __ldg(p); }
template <>
__device__ __forceinline__ float ls_ld<__nv_bfloat16>(const __nv_bfloat16 *p) {
    return __uint_as_float((uint32_t)__ldg(reinterpret_cast<const unsigned short *>(p)) << 16);
}
template <typename T>
__device__ __forceinline__ void ls_st(T *p, float v);
template <>
__device__ __forceinline__ void ls_st<float>(float *p, float v) { *p = v; }
template <>
__device__ __forceinline__ void ls_st<__nv_bfloat16>(__nv_bfloat16 *p, float v) { *p = __float2bfloat16_rn(v); }

template <typename TT>
__device__ __forceinline__ int ls_label(const TT *p);
template <>
__device__ __forceinline__ int ls_label<float>(const float *p) { return __float2int_rn(__ldg(p)); }
template <>
__device__ __forceinline__ int ls_label<long long>(const long long *p) { return (int)__ldg(p); }

struct LossParams {
    const void *logits, *target;
    void *dlogits;
    float *stats, *ce;                   // stats (B, K, 3), ce (1)
    const float *g_stats, *g_ce;         // backward: (B, K, 3) (third component ignored), (1)
    long long sb, sc, sn;                // logits strides in elements: batch, class, pixel
    long long npix;                      // pixels per image
    int K, rows16;
    long long pix_per_block;
};

// the logits of one pixel.  VEC: bf16 rows of exactly 16 elements, 32-byte aligned (the segmentation heads padded to 16
// output channels, classes contiguous): two 16-byte loads instead of K strided 2-byte ones (each warp instruction of the
// scalar form touches 32 different sectors, 14 times over)
template <typename T, int KP, bool VEC>
__device__ __forceinline__ void ls_row(const T *lp, long long sc, int K, float (&p)[KP]) {
    if constexpr (VEC) {
        const uint4 a = __ldg(reinterpret_cast<const uint4 *>(lp)), b = __ldg(reinterpret_cast<const uint4 *>(lp) + 1);
        const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            p[2 * i] = 2 * i < K ? __uint_as_float(w[i] << 16) : -3.0e38f;
            p[2 * i + 1] = 2 * i + 1 < K ? __uint_as_float(w[i] & 0xffff0000u) : -3.0e38f;
        }
    } else {
#pragma unroll
        for (int k = 0; k < KP; ++k) p[k] = k < K ? ls_ld<T>(lp + k * sc) : -3.0e38f;
    }
}

// softmax of one pixel in registers; returns max and sum
template <typename T, int KP, bool VEC>
__device__ __forceinline__ void ls_softmax(const T *lp, long long sc, int K, float (&p)[KP], float &mx, float &sum) {
    ls_row<T, KP, VEC>(lp, sc, K, p);
    mx = -3.0e38f;
#pragma unroll
    for (int k = 0; k < KP; ++k) mx = fmaxf(mx, p[k]);
    sum = 0.f;
#pragma unroll
    for (int k = 0; k < KP; ++k) {
        p[k] = k < K ? __expf(p[k] - mx) : 0.f;
        sum += p[k];
    }
    const float inv = 1.f / sum;
#pragma unroll
    for (int k = 0; k < KP; ++k) p[k] *= inv;
}

template <typename T, typename TT, int KP, bool VEC>
__global__ void __launch_bounds__(256) dice_ce_fwd_kernel(const LossParams q) {
    __shared__ float red[8][3 * KP + 1];
    const int K = q.K, bi = blockIdx.y;
    const T *lb = static_cast<const T *>(q.logits) + bi * q.sb;
    const TT *tb = static_cast<const TT *>(q.target) + bi * q.npix;
    const long long n0 = blockIdx.x * q.pix_per_block, n1 = min(q.npix, n0 + q.pix_per_block);
    float inter[KP], pred[KP], cnt[KP], ce = 0.f;
#pragma unroll
    for (int k = 0; k < KP; ++k) inter[k] = pred[k] = cnt[k] = 0.f;
    for (long long n = n0 + threadIdx.x; n < n1; n += 256) {
        const T *lp = lb + n * q.sn;
        const int t = ls_label<TT>(tb + n);
        float p[KP], mx, sum;
        // the logit of the label, before the exponentials overwrite the registers
        const float lt = (t >= 0 && t < K) ? ls_ld<T>(lp + t * q.sc) : 0.f;
        ls_softmax<T, KP, VEC>(lp, q.sc, K, p, mx, sum);
        if (t >= 0 && t < K) ce += logf(sum) - (lt - mx);
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            pred[k] += p[k];
            if (k == t) inter[k] += p[k], cnt[k] += 1.f;
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < KP; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            inter[k] += __shfl_xor_sync(0xffffffffu, inter[k], o);
            pred[k] += __shfl_xor_sync(0xffffffffu, pred[k], o);
            cnt[k] += __shfl_xor_sync(0xffffffffu, cnt[k], o);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ce += __shfl_xor_sync(0xffffffffu, ce, o);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < KP; ++k) red[warp][3 * k] = inter[k], red[warp][3 * k + 1] = pred[k], red[warp][3 * k + 2] = cnt[k];
        red[warp][3 * KP] = ce;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * KP + 1; i += 256) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w][i];
        if (i == 3 * KP) atomicAdd(q.ce, s);
        else if (i / 3 < K) atomicAdd(q.stats + ((size_t)bi * K + i / 3) * 3 + i % 3, s);
    }
}

template <typename T, typename TT, int KP, bool VEC>
__global__ void __launch_bounds__(256) dice_ce_bwd_kernel(const LossParams q) {
    __shared__ float gi[KP], gp[KP];
    const int K = q.K, bi = blockIdx.y;
    if (threadIdx.x < KP) {
        const bool in = threadIdx.x < K;
        gi[threadIdx.x] = in ? q.g_stats[((size_t)bi * K + threadIdx.x) * 3] : 0.f;
        gp[threadIdx.x] = in ? q.g_stats[((size_t)bi * K + threadIdx.x) * 3 + 1] : 0.f;
    }
    __syncthreads();
    const float gce = __ldg(q.g_ce);
    const T *lb = static_cast<const T *>(q.logits) + bi * q.sb;
    T *db = static_cast<T *>(q.dlogits) + bi * q.sb;
    const TT *tb = static_cast<const TT *>(q.target) + bi * q.npix;
    const long long n0 = blockIdx.x * q.pix_per_block, n1 = min(q.npix, n0 + q.pix_per_block);
    for (long long n = n0 + threadIdx.x; n < n1; n += 256) {
        const T *lp = lb + n * q.sn;
        const int t = ls_label<TT>(tb + n);
        float p[KP], mx, sum;
        ls_softmax<T, KP, VEC>(lp, q.sc, K, p, mx, sum);
        float dot = 0.f;
#pragma unroll
        for (int k = 0; k < KP; ++k) dot = fmaf(p[k], gp[k] + (k == t ? gi[k] : 0.f), dot);
        const float cev = (t >= 0 && t < K) ? gce : 0.f;     // labels outside [0, K) contribute no cross-entropy term
        float v[KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            const float G = gp[k] + (k == t ? gi[k] : 0.f);
            v[k] = k < K ? p[k] * (G - dot) + cev * (p[k] - (k == t ? 1.f : 0.f)) : 0.f;
        }
        if constexpr (VEC) {            // the row's padding columns receive zeros
            uint32_t w[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
                w[i] = *reinterpret_cast<const uint32_t *>(&h2);
            }
            uint4 *dp = reinterpret_cast<uint4 *>(db + n * q.sn);
            dp[0] = make_uint4(w[0], w[1], w[2], w[3]);
            dp[1] = make_uint4(w[4], w[5], w[6], w[7]);
        } else {
#pragma unroll
            for (int k = 0; k < KP; ++k)
                if (k < K) ls_st<T>(db + n * q.sn + k * q.sc, v[k]);
        }
    }
}

template <typename T, typename TT>
static cudaError_t loss_launch(const LossParams &q, int Bn, bool bwd, cudaStream_t st) {
    long long chunks = (148LL * 8 + Bn - 1) / Bn;
    LossParams p = q;
    p.pix_per_block = (q.npix + chunks - 1) / chunks;
    if (p.pix_per_block < 256) p.pix_per_block = 256;
    const dim3 grid((unsigned)((q.npix + p.pix_per_block - 1) / p.pix_per_block), (unsigned)Bn);
#define MLAGG_LOSS(KP) \
    do { \
        if (bwd) dice_ce_bwd_kernel<T, TT, KP, false><<<grid, 256, 0, st>>>(p); \
        else dice_ce_fwd_kernel<T, TT, KP, false><<<grid, 256, 0, st>>>(p); \
    } while (0)
    if constexpr (sizeof(T) == 2) {
        // rows of 16 bf16 (padded heads), classes contiguous, everything 32-byte aligned: vector row access.  `rows16` is the
        // caller's statement that each pixel row owns all 16 elements (the gradient's padding columns are written).
        if (q.rows16 && q.sc == 1 && q.sn == 16 && q.K <= 16 && q.K > 8 && q.sb % 16 == 0 &&
            reinterpret_cast<uintptr_t>(q.logits) % 32 == 0 && (!bwd || reinterpret_cast<uintptr_t>(q.dlogits) % 32 == 0)) {
            if (bwd) dice_ce_bwd_kernel<T, TT, 16, true><<<grid, 256, 0, st>>>(p);
            else dice_ce_fwd_kernel<T, TT, 16, true><<<grid, 256, 0, st>>>(p);
            return cudaGetLastError();
        }
    }
    if (q.K <= 4) MLAGG_LOSS(4);
    else if (q.K <= 8) MLAGG_LOSS(8);
    else if (q.K <= 16) MLAGG_LOSS(16);
    else MLAGG_LOSS(32);
#undef MLAGG_LOSS
    return cudaGetLastError();
}

// dtype: 0 fp32, 1 bf16 logits; tdtype: 0 float labels, 1 int64 labels
cudaError_t dice_ce_dispatch(const void *logits, const void *target, void *dlogits, float *stats, float *ce,
                             const float *g_stats, const float *g_ce, long long sb, long long sc, long long sn,
                             long long npix, int K, int Bn, int dtype, int tdtype, bool bwd, cudaStream_t st) {
    LossParams q{};
    q.rows16 = (dtype & 2) ? 1 : 0;      // bit 1 of dtype: rows of 16 elements owned by the tensor (see loss_launch)
    dtype &= 1;
    q.logits = logits, q.target = target, q.dlogits = dlogits, q.stats = stats, q.ce = ce, q.g_stats = g_stats, q.g_ce = g_ce;
    q.sb = sb, q.sc = sc, q.sn = sn, q.npix = npix, q.K = K;
    if (dtype == 0) return tdtype == 0 ? loss_launch<float, float>(q, Bn, bwd, st) : loss_launch<float, long long>(q, Bn, bwd, st);
    return tdtype == 0 ? loss_launch<__nv_bfloat16, float>(q, Bn, bwd, st) : loss_launch<__nv_bfloat16, long long>(q, Bn, bwd, st);
}

}  // namespace mlagg
