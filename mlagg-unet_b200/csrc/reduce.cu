// reduce.cu -- column sums of tokens-major activations: out[c] += sum_m x[m][c].
// This is the bias gradient of every nn.Linear on the hot path (reference nnUNetTrainer_MLAgg_2D_dt_MS.py:849-850,
// :868, :180-186, :673-674 -- in_proj / act_proj / out_proj / Mlp.fc1,fc2 / q / kv; MambaSkip.py:567,570 -- ConvGLU):
// autograd's `grad_output.sum(0)` on a (tokens, C) bf16 matrix runs at ~190 GB/s in torch's generic reduce kernel; here
// it is one HBM-bound pass (algorithmic bytes M*C*e).  A warp reads 32 x VEC consecutive channels of a row (coalesced
// 128-bit fp32 / 64-bit bf16), the 8 warps of a block take different rows, partial sums meet in shared memory and leave
// as one fp32 atomic per column and block.
#include <cuda_bf16.h>

#include "common.cuh"

namespace mlagg {

template <typename T, int VEC>
__device__ __forceinline__ void cs_ld(const T *p, float (&v)[VEC]);
template <>
__device__ __forceinline__ void cs_ld<float, 4>(const float *p, float (&v)[4]) {
    const float4 t = __ldg(reinterpret_cast<const float4 *>(p));
    v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
}
template <>
__device__ __forceinline__ void cs_ld<float, 1>(const float *p, float (&v)[1]) { v[0] = __ldg(p); }
template <>
__device__ __forceinline__ void cs_ld<__nv_bfloat16, 4>(const __nv_bfloat16 *p, float (&v)[4]) {
    const uint2 t = __ldg(reinterpret_cast<const uint2 *>(p));
    v[0] = __uint_as_float(t.x << 16), v[1] = __uint_as_float(t.x & 0xffff0000u);
    v[2] = __uint_as_float(t.y << 16), v[3] = __uint_as_float(t.y & 0xffff0000u);
}
template <>
__device__ __forceinline__ void cs_ld<__nv_bfloat16, 1>(const __nv_bfloat16 *p, float (&v)[1]) {
    v[0] = __bfloat162float(*p);
}

constexpr int kCsWarps = 8;

template <typename T, int VEC>
__global__ void __launch_bounds__(32 * kCsWarps) colsum_kernel(const T *__restrict__ x, float *__restrict__ out,
                                                               long long M, int C, long long ld, int rows_per_block) {
    __shared__ float red[kCsWarps][32 * VEC + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = (blockIdx.x * 32 + lane) * VEC;
    const long long m0 = (long long)blockIdx.y * rows_per_block;
    const long long m1 = min(M, m0 + rows_per_block);
    float acc[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
    if (c0 < C) {
        long long m = m0 + warp;
        // two rows in flight per iteration
        for (; m + kCsWarps < m1; m += 2 * kCsWarps) {
            float a[VEC], b[VEC];
            cs_ld<T, VEC>(x + m * ld + c0, a);
            cs_ld<T, VEC>(x + (m + kCsWarps) * ld + c0, b);
#pragma unroll
            for (int i = 0; i < VEC; ++i) acc[i] += a[i] + b[i];
        }
        if (m < m1) {
            float a[VEC];
            cs_ld<T, VEC>(x + m * ld + c0, a);
#pragma unroll
            for (int i = 0; i < VEC; ++i) acc[i] += a[i];
        }
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) red[warp][lane * VEC + i] = acc[i];
    __syncthreads();
    for (int j = threadIdx.x; j < 32 * VEC; j += blockDim.x) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kCsWarps; ++w) s += red[w][j];
        const int c = blockIdx.x * 32 * VEC + j;
        if (c < C) atomicAdd(out + c, s);
    }
}

cudaError_t colsum_dispatch(const void *x, float *out, long long M, int C, long long ld, int dtype, cudaStream_t st) {
    const bool vec = (C % 4 == 0) && (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    const int VEC = vec ? 4 : 1;
    const int gx = (C + 32 * VEC - 1) / (32 * VEC);
    // ~8 blocks per SM in total; at least 64 rows per block so the atomics stay negligible
    long long by = (148LL * 8 + gx - 1) / gx;
    long long rpb = (M + by - 1) / by;
    if (rpb < 64) rpb = 64;
    by = (M + rpb - 1) / rpb;
    if (by > 65535) { by = 65535; rpb = (M + by - 1) / by; }
    const dim3 grid(gx, (unsigned)by);
    if (dtype == 0) {
        if (vec) colsum_kernel<float, 4><<<grid, 32 * kCsWarps, 0, st>>>(static_cast<const float *>(x), out, M, C, ld, (int)rpb);
        else colsum_kernel<float, 1><<<grid, 32 * kCsWarps, 0, st>>>(static_cast<const float *>(x), out, M, C, ld, (int)rpb);
    } else {
        if (vec) colsum_kernel<__nv_bfloat16, 4><<<grid, 32 * kCsWarps, 0, st>>>(static_cast<const __nv_bfloat16 *>(x), out, M, C, ld, (int)rpb);
        else colsum_kernel<__nv_bfloat16, 1><<<grid, 32 * kCsWarps, 0, st>>>(static_cast<const __nv_bfloat16 *>(x), out, M, C, ld, (int)rpb);
    }
    return cudaGetLastError();
}

}  // namespace mlagg
