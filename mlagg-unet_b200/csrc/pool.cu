// pool.cu -- adaptive average pooling of a tokens-major (B, H*W, C) map to (B, pH*pW, C) with the preceding GELU
// folded into the read: the pooled-token branch of AggregatedAttention (reference nnUNetTrainer_MLAgg_2D_dt_MS.py:720-723,
//   x_ = self.pool(self.act(self.sr(x)))  with nn.AdaptiveAvgPool2d((H/sr, W/sr)), nn.GELU()).
// torch's adaptive_average_pool_nhwc kernel needs 0.5 ms for the 24 MB stage-0 map (it parallelises over output
// pixels only); here a block owns one output token, its threads split the bin's pixels and the channel vectors, so
// the read is coalesced and HBM-bound (algorithmic bytes B*N*C*e).  Bin i covers rows floor(i*H/pH) .. ceil((i+1)*H/pH)
// (torch semantics; bins overlap when H % pH != 0, which the backward handles).
#include <cuda_bf16.h>

#include <algorithm>

#include "common.cuh"

namespace mlagg {

__device__ __forceinline__ void pl4_ld(const float *p, float (&v)[4]) {
    const float4 t = __ldg(reinterpret_cast<const float4 *>(p));
    v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
}
__device__ __forceinline__ void pl4_ld(const __nv_bfloat16 *p, float (&v)[4]) {
    const uint2 t = __ldg(reinterpret_cast<const uint2 *>(p));
    v[0] = __uint_as_float(t.x << 16), v[1] = __uint_as_float(t.x & 0xffff0000u);
    v[2] = __uint_as_float(t.y << 16), v[3] = __uint_as_float(t.y & 0xffff0000u);
}
__device__ __forceinline__ void pl4_st(float *p, const float (&v)[4]) {
    *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void pl4_st(__nv_bfloat16 *p, const float (&v)[4]) {
    const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 raw;
    raw.x = *reinterpret_cast<const uint32_t *>(&a);
    raw.y = *reinterpret_cast<const uint32_t *>(&b);
    *reinterpret_cast<uint2 *>(p) = raw;
}
// GELU of the pooled branch.  bf16 storage: the Abramowitz-Stegun erf of common.cuh (|error| <= 1.5e-7, far inside the
// output rounding).  fp32 storage is the 1e-4 parity path: libm's erff / expf -- the pooled tokens feed d lambda, a sum
// over every token that cancels to ~1e-5 of its terms, and a systematic 1e-7 in them showed up as 2e-2 there.
template <typename T>
__device__ __forceinline__ float pool_gelu(float x) {
    if (sizeof(T) == 4) return 0.5f * x * (1.f + erff(x * 0.70710678118654752f));
    return gelu_f(x);
}
template <typename T>
__device__ __forceinline__ float pool_dgelu(float x) {
    if (sizeof(T) == 4) return 0.5f * (1.f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * expf(-0.5f * x * x);
    return gelu_grad(x);
}
__device__ __forceinline__ int bin_lo(int i, int n, int pn) { return (int)(((long long)i * n) / pn); }
__device__ __forceinline__ int bin_hi(int i, int n, int pn) { return (int)((((long long)(i + 1)) * n + pn - 1) / pn); }

constexpr int kPoolThreads = 256;

template <typename T>
__global__ void __launch_bounds__(kPoolThreads) avgpool_fwd_kernel(const T *__restrict__ x, T *__restrict__ y, int H, int W,
                                                                   int C, int pH, int pW, int gelu) {
    extern __shared__ float red[];   // [ny][C]
    const int cv = C >> 2;
    const int p = blockIdx.x, bi = blockIdx.y;
    const int pi = p / pW, pj = p - pi * pW;
    const int r0 = bin_lo(pi, H, pH), r1 = bin_hi(pi, H, pH), c0 = bin_lo(pj, W, pW), c1 = bin_hi(pj, W, pW);
    const int bw = c1 - c0, npix = (r1 - r0) * bw;
    const int ny = max(1, kPoolThreads / cv);
    const int tx = threadIdx.x % cv, ty = threadIdx.x / cv;
    const T *xb = x + (size_t)bi * H * W * C;
    if (ty < ny) {
        for (int cb = tx; cb < cv; cb += (ny == 1 ? kPoolThreads : cv)) {   // ny == 1: cv may exceed the block
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            for (int k = ty; k < npix; k += ny) {
                const int rr = r0 + k / bw, cc = c0 + k % bw;
                float v[4];
                pl4_ld(xb + ((size_t)rr * W + cc) * C + 4 * cb, v);
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[i] += gelu ? pool_gelu<T>(v[i]) : v[i];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) red[ty * C + 4 * cb + i] = acc[i];
        }
    }
    __syncthreads();
    const float inv = 1.f / (float)npix;
    for (int cb = threadIdx.x; cb < cv; cb += kPoolThreads) {
        float o[4] = {0.f, 0.f, 0.f, 0.f};
        for (int j = 0; j < ny; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) o[i] += red[j * C + 4 * cb + i];
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] *= inv;
        pl4_st(y + ((size_t)bi * pH * pW + p) * C + 4 * cb, o);
    }
}

// dx[b, (r, c), :] = gelu'(x) * sum over the bins containing (r, c) of dy[b, bin, :] / |bin|
// A thread owns one (column, channel vector) and walks down the rows of its row chunk: the column's bins (at most two per
// dimension; exactly one when W % pW == 0) are found once, the row's bins once per row -- the first version redid two
// 64-bit divisions and up to nine bin tests for every 8-byte vector and ran at a tenth of the HBM roofline.
template <typename T>
__global__ void __launch_bounds__(256) avgpool_bwd_kernel(const T *__restrict__ x, const T *__restrict__ dy,
                                                          T *__restrict__ dx, int H, int W, int C, int pH, int pW,
                                                          int gelu, int rows_per_block) {
    const int cv = C >> 2;
    const int bi = blockIdx.z;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= W * cv) return;
    const int c = idx / cv, cb = idx - c * cv;
    int jb[2], nj = 0;
    float jinv[2];
    {
        const int j0 = (int)(((long long)c * pW) / W);
        for (int j = max(0, j0 - 1); j <= min(pW - 1, j0 + 1) && nj < 2; ++j) {
            const int c0 = bin_lo(j, W, pW), c1 = bin_hi(j, W, pW);
            if (c >= c0 && c < c1) jb[nj] = j, jinv[nj] = 1.f / (float)(c1 - c0), ++nj;
        }
    }
    const int r_begin = blockIdx.y * rows_per_block, r_end = min(H, r_begin + rows_per_block);
    const T *dyb = dy + (size_t)bi * pH * pW * C + 4 * cb;
    for (int r = r_begin; r < r_end; ++r) {
        float g[4] = {0.f, 0.f, 0.f, 0.f};
        const int i0 = (int)(((long long)r * pH) / H);
        for (int i = max(0, i0 - 1); i <= min(pH - 1, i0 + 1); ++i) {
            const int r0 = bin_lo(i, H, pH), r1 = bin_hi(i, H, pH);
            if (r < r0 || r >= r1) continue;
            const float iinv = 1.f / (float)(r1 - r0);
            for (int q = 0; q < nj; ++q) {
                float v[4];
                pl4_ld(dyb + ((size_t)i * pW + jb[q]) * C, v);
                const float inv = iinv * jinv[q];
#pragma unroll
                for (int k = 0; k < 4; ++k) g[k] = fmaf(v[k], inv, g[k]);
            }
        }
        const size_t off = (((size_t)bi * H + r) * W + c) * C + 4 * cb;
        if (gelu) {
            float v[4];
            pl4_ld(x + off, v);
#pragma unroll
            for (int k = 0; k < 4; ++k) g[k] *= pool_dgelu<T>(v[k]);
        }
        pl4_st(dx + off, g);
    }
}

cudaError_t avgpool_dispatch(const void *x, const void *dy, void *out, int Bn, int H, int W, int C, int pH, int pW,
                             int gelu, int dtype, bool bwd, cudaStream_t st) {
    const int cv = C / 4;
    if (!bwd) {
        const int ny = cv >= kPoolThreads ? 1 : kPoolThreads / cv;
        const size_t sm = (size_t)ny * C * sizeof(float);
        const dim3 grid(pH * pW, Bn);
        if (dtype == 0)
            avgpool_fwd_kernel<float><<<grid, kPoolThreads, sm, st>>>(static_cast<const float *>(x), static_cast<float *>(out), H, W, C, pH, pW, gelu);
        else
            avgpool_fwd_kernel<__nv_bfloat16><<<grid, kPoolThreads, sm, st>>>(static_cast<const __nv_bfloat16 *>(x), static_cast<__nv_bfloat16 *>(out), H, W, C, pH, pW, gelu);
    } else {
        // ~8 CTAs per SM in total, every thread walking >= 4 rows
        const int gx = (W * cv + 255) / 256;
        int chunks = std::max(1, (148 * 8) / std::max(1, gx * Bn));
        int rpb = std::max(4, (H + chunks - 1) / chunks);
        const dim3 grid(gx, (H + rpb - 1) / rpb, Bn);
        if (dtype == 0)
            avgpool_bwd_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float *>(x), static_cast<const float *>(dy), static_cast<float *>(out), H, W, C, pH, pW, gelu, rpb);
        else
            avgpool_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16 *>(x), static_cast<const __nv_bfloat16 *>(dy), static_cast<__nv_bfloat16 *>(out), H, W, C, pH, pW, gelu, rpb);
    }
    return cudaGetLastError();
}

}  // namespace mlagg
