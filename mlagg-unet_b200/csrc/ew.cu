// ew.cu -- the element-wise seams of the MLAgg block as single HBM-bound passes (128-bit fp32 / 64-bit bf16 accesses):
//   residual + stochastic depth : out = x + s[b] * y          (nnUNetTrainer_MLAgg_2D_dt_MS.py:907-908; timm DropPath's
//                                                               per-sample mask/keep as a (B) fp32 vector, or none)
//   SiLU gate                   : out = t * silu(z)            (:881 act_proj + SiLU, :907 `x * act_res`)
//   lambda of the differential attention: exp(<lq1, lk1>) - exp(<lq2, lk2>) + lambda_init   (:700-702, :745-747)
// torch ran these as 2-3 kernels each (a broadcast multiply falls off its vectorised path), plus 8 + 14 one-block kernels
// per attention module for the scalar lambda and its gradient.
#include <cuda_bf16.h>

#include <algorithm>
#include <type_traits>

#include "common.cuh"

namespace mlagg {

template <typename T>
__device__ __forceinline__ float4 ew_ld4(const T *p);
template <>
__device__ __forceinline__ float4 ew_ld4<float>(const float *p) {
    return __ldg(reinterpret_cast<const float4 *>(p));
}
template <>
__device__ __forceinline__ float4 ew_ld4<__nv_bfloat16>(const __nv_bfloat16 *p) {
    const uint2 t = __ldg(reinterpret_cast<const uint2 *>(p));
    return make_float4(__uint_as_float(t.x << 16), __uint_as_float(t.x & 0xffff0000u), __uint_as_float(t.y << 16),
                       __uint_as_float(t.y & 0xffff0000u));
}
template <typename T>
__device__ __forceinline__ void ew_st4(T *p, float4 v);
template <>
__device__ __forceinline__ void ew_st4<float>(float *p, float4 v) {
    *reinterpret_cast<float4 *>(p) = v;
}
template <>
__device__ __forceinline__ void ew_st4<__nv_bfloat16>(__nv_bfloat16 *p, float4 v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 raw;
    raw.x = *reinterpret_cast<uint32_t *>(&a);
    raw.y = *reinterpret_cast<uint32_t *>(&b);
    *reinterpret_cast<uint2 *>(p) = raw;
}

// out = (x ? x : 0) + s[b] * y ; n4 = elements / 4, per4 = elements per sample / 4
template <typename T>
__global__ void __launch_bounds__(256) residual_scale_kernel(const T *__restrict__ x, const T *__restrict__ y,
                                                             const float *__restrict__ s, T *__restrict__ out,
                                                             long long n4, long long per4) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float sc = s ? __ldg(s + i / per4) : 1.f;
        const float4 b = ew_ld4<T>(y + 4 * i);
        float4 a = x ? ew_ld4<T>(x + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
        a.x = fmaf(sc, b.x, a.x), a.y = fmaf(sc, b.y, a.y), a.z = fmaf(sc, b.z, a.z), a.w = fmaf(sc, b.w, a.w);
        ew_st4<T>(out + 4 * i, a);
    }
}

__device__ __forceinline__ float sigmoid_f(float z) { return rcp_approx(1.f + ex2_approx(-z * kLog2e)); }

template <typename T>
__global__ void __launch_bounds__(256) silu_gate_fwd_kernel(const T *__restrict__ t, const T *__restrict__ z,
                                                            T *__restrict__ out, long long n4) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 a = ew_ld4<T>(t + 4 * i), g = ew_ld4<T>(z + 4 * i);
        ew_st4<T>(out + 4 * i, make_float4(a.x * g.x * sigmoid_f(g.x), a.y * g.y * sigmoid_f(g.y),
                                           a.z * g.z * sigmoid_f(g.z), a.w * g.w * sigmoid_f(g.w)));
    }
}

// dt = g * silu(z);  dz = g * t * silu'(z),  silu'(z) = s (1 + z (1 - s))
template <typename T>
__global__ void __launch_bounds__(256) silu_gate_bwd_kernel(const T *__restrict__ t, const T *__restrict__ z,
                                                            const T *__restrict__ g, T *__restrict__ dt,
                                                            T *__restrict__ dz, long long n4) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 a = ew_ld4<T>(t + 4 * i), zz = ew_ld4<T>(z + 4 * i), gg = ew_ld4<T>(g + 4 * i);
        float4 o1, o2;
        float s;
        s = sigmoid_f(zz.x); o1.x = gg.x * zz.x * s; o2.x = gg.x * a.x * s * (1.f + zz.x * (1.f - s));
        s = sigmoid_f(zz.y); o1.y = gg.y * zz.y * s; o2.y = gg.y * a.y * s * (1.f + zz.y * (1.f - s));
        s = sigmoid_f(zz.z); o1.z = gg.z * zz.z * s; o2.z = gg.z * a.z * s * (1.f + zz.z * (1.f - s));
        s = sigmoid_f(zz.w); o1.w = gg.w * zz.w * s; o2.w = gg.w * a.w * s * (1.f + zz.w * (1.f - s));
        ew_st4<T>(dt + 4 * i, o1);
        ew_st4<T>(dz + 4 * i, o2);
    }
}

// out[0] = exp(<q1,k1>) - exp(<q2,k2>) + init, out[1] = exp(<q1,k1>), out[2] = exp(<q2,k2>)     (one warp)
__global__ void __launch_bounds__(32) diff_lambda_fwd_kernel(const float *__restrict__ q1, const float *__restrict__ k1,
                                                             const float *__restrict__ q2, const float *__restrict__ k2,
                                                             int n, float init, float *__restrict__ out) {
    float s1 = 0.f, s2 = 0.f;
    for (int i = threadIdx.x; i < n; i += 32) {
        s1 = fmaf(q1[i], k1[i], s1);
        s2 = fmaf(q2[i], k2[i], s2);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if (threadIdx.x == 0) {
        const float e1 = expf(s1), e2 = expf(s2);
        out[0] = e1 - e2 + init;
        out[1] = e1;
        out[2] = e2;
    }
}

// d(q1, k1, q2, k2) (4, n) = dlam * (e1 k1, e1 q1, -e2 k2, -e2 q2)
__global__ void __launch_bounds__(32) diff_lambda_bwd_kernel(const float *__restrict__ q1, const float *__restrict__ k1,
                                                             const float *__restrict__ q2, const float *__restrict__ k2,
                                                             const float *__restrict__ saved,
                                                             const float *__restrict__ dlam, int n,
                                                             float *__restrict__ grads) {
    const float g = dlam[0], e1 = saved[1] * g, e2 = -saved[2] * g;
    for (int i = threadIdx.x; i < n; i += 32) {
        grads[i] = e1 * k1[i];
        grads[n + i] = e1 * q1[i];
        grads[2 * n + i] = e2 * k2[i];
        grads[3 * n + i] = e2 * q2[i];
    }
}

__device__ __forceinline__ float4 ew_ld4_rw(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ float4 ew_ld4_rw(const __nv_bfloat16 *p) {
    const uint2 t = *reinterpret_cast<const uint2 *>(p);
    return make_float4(__uint_as_float(t.x << 16), __uint_as_float(t.x & 0xffff0000u), __uint_as_float(t.y << 16),
                       __uint_as_float(t.y & 0xffff0000u));
}

// y[pix][c] += b[c] in place on a channels_last / tokens-major map (C % 4 == 0)
template <typename T>
__global__ void __launch_bounds__(256) bias_add_cl_kernel(T *__restrict__ y, const float *__restrict__ b, long long n4,
                                                          int C4) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 bv = __ldg(reinterpret_cast<const float4 *>(b) + (i % C4));
        float4 a = ew_ld4_rw(y + 4 * i);      // plain (coherent) load: the same address is written below
        a.x += bv.x, a.y += bv.y, a.z += bv.z, a.w += bv.w;
        ew_st4<T>(y + 4 * i, a);
    }
}

// dst[b][r][0:cols] = src[b][r][0:cols] for row-strided views (ld*, bs* in elements): channel slices and stage segments
// of tokens-major activations packed / scattered without torch's generic strided-copy kernel (~1 TB/s on these shapes)
template <typename T, int V>
__global__ void __launch_bounds__(256) copy_rows_kernel(const T *__restrict__ src, long long ld_s, long long bs_s,
                                                        T *__restrict__ dst, long long ld_d, long long bs_d,
                                                        long long rows, int colsv) {
    const T *sb = src + (long long)blockIdx.y * bs_s;
    T *db = dst + (long long)blockIdx.y * bs_d;
    const long long n = rows * colsv;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / colsv;
        const int c = (int)(i % colsv) * V;
        if (V == 1) {
            db[r * ld_d + c] = sb[r * ld_s + c];
        } else if (sizeof(T) * V == 16) {
            *reinterpret_cast<uint4 *>(db + r * ld_d + c) = __ldg(reinterpret_cast<const uint4 *>(sb + r * ld_s + c));
        } else {
            *reinterpret_cast<uint2 *>(db + r * ld_d + c) = __ldg(reinterpret_cast<const uint2 *>(sb + r * ld_s + c));
        }
    }
}

// the same for <= 256 vectors per row without per-element index arithmetic: a thread owns one column vector and walks
// down its CTA's rows four at a time (all loads first); ADD: dst += src (gradient joins of channel slices)
template <typename T, int V, bool ADD>
__global__ void __launch_bounds__(256) rows_walk_kernel(const T *__restrict__ src, long long ld_s, long long bs_s,
                                                        T *__restrict__ dst, long long ld_d, long long bs_d,
                                                        long long rows, int colsv, long long rows_per_block) {
    using Vec = typename std::conditional<sizeof(T) * V == 16, uint4, uint2>::type;
    const int rpi = 256 / colsv;
    const int jc = threadIdx.x % colsv, jr = threadIdx.x / colsv;
    if (jr >= rpi) return;
    const T *sb = src + (long long)blockIdx.y * bs_s + jc * V;
    T *db = dst + (long long)blockIdx.y * bs_d + jc * V;
    const long long r0 = blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
    for (long long r = r0 + jr; r < r1; r += 4 * rpi) {
        Vec a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long rr = r + u * rpi;
            if (rr < r1) {
                a[u] = __ldg(reinterpret_cast<const Vec *>(sb + rr * ld_s));
                if (ADD) b[u] = *reinterpret_cast<const Vec *>(db + rr * ld_d);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long rr = r + u * rpi;
            if (rr < r1) {
                if (ADD) {
                    constexpr int NW = sizeof(Vec) / 4;
                    uint32_t *pa = reinterpret_cast<uint32_t *>(&a[u]);
                    const uint32_t *pb = reinterpret_cast<const uint32_t *>(&b[u]);
#pragma unroll
                    for (int k = 0; k < NW; ++k) {
                        if (sizeof(T) == 4) {
                            pa[k] = __float_as_uint(__uint_as_float(pa[k]) + __uint_as_float(pb[k]));
                        } else {
                            const float lo = __uint_as_float(pa[k] << 16) + __uint_as_float(pb[k] << 16);
                            const float hi = __uint_as_float(pa[k] & 0xffff0000u) + __uint_as_float(pb[k] & 0xffff0000u);
                            const __nv_bfloat162 o = __floats2bfloat162_rn(lo, hi);
                            pa[k] = *reinterpret_cast<const uint32_t *>(&o);
                        }
                    }
                }
                *reinterpret_cast<Vec *>(db + rr * ld_d) = a[u];
            }
        }
    }
}

static int ew_blocks(long long n4) {
    long long b = (n4 + 255) / 256;
    const long long cap = 148LL * 16;
    return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

cudaError_t residual_scale_dispatch(const void *x, const void *y, const float *s, void *out, long long n,
                                    long long per_sample, int dtype, cudaStream_t st) {
    const long long n4 = n / 4, per4 = per_sample / 4;
    if (dtype == 0)
        residual_scale_kernel<float><<<ew_blocks(n4), 256, 0, st>>>(static_cast<const float *>(x), static_cast<const float *>(y), s, static_cast<float *>(out), n4, per4);
    else
        residual_scale_kernel<__nv_bfloat16><<<ew_blocks(n4), 256, 0, st>>>(static_cast<const __nv_bfloat16 *>(x), static_cast<const __nv_bfloat16 *>(y), s, static_cast<__nv_bfloat16 *>(out), n4, per4);
    return cudaGetLastError();
}

cudaError_t bias_add_cl_dispatch(void *y, const float *b, long long n, int C, int dtype, cudaStream_t st) {
    const long long n4 = n / 4;
    if (dtype == 0) bias_add_cl_kernel<float><<<ew_blocks(n4), 256, 0, st>>>(static_cast<float *>(y), b, n4, C / 4);
    else bias_add_cl_kernel<__nv_bfloat16><<<ew_blocks(n4), 256, 0, st>>>(static_cast<__nv_bfloat16 *>(y), b, n4, C / 4);
    return cudaGetLastError();
}

cudaError_t copy_rows_dispatch(const void *src, long long ld_s, long long bs_s, void *dst, long long ld_d, long long bs_d,
                               int batch, long long rows, int cols, int dtype, bool add, cudaStream_t st) {
    const size_t es = dtype == 0 ? 4 : 2;
    // widest vector (in elements) that divides cols and every stride and keeps both bases aligned
    int V = (int)(16 / es);
    auto ok = [&](int v) {
        const size_t bytes = v * es;
        return cols % v == 0 && ld_s % v == 0 && ld_d % v == 0 && bs_s % v == 0 && bs_d % v == 0 &&
               reinterpret_cast<uintptr_t>(src) % bytes == 0 && reinterpret_cast<uintptr_t>(dst) % bytes == 0;
    };
    while (V > 1 && !ok(V)) V /= 2;
    if (V * es < 8) V = 1;                                   // 8- and 16-byte vectors, else scalar
    if (V * es >= 8 && cols / V <= 256) {
        const int rpi = 256 / (cols / V);
        long long chunks = std::max<long long>(1, (148LL * 8 + batch - 1) / batch);
        long long rpb = (rows + chunks - 1) / chunks;
        rpb = std::max<long long>(4LL * rpi, (rpb + 4LL * rpi - 1) / (4LL * rpi) * (4LL * rpi));
        const dim3 g2((unsigned)((rows + rpb - 1) / rpb), (unsigned)batch);
#define MLAGG_WALK(T_, V_) \
    do { \
        if (add) rows_walk_kernel<T_, V_, true><<<g2, 256, 0, st>>>(static_cast<const T_ *>(src), ld_s, bs_s, static_cast<T_ *>(dst), ld_d, bs_d, rows, cols / V_, rpb); \
        else rows_walk_kernel<T_, V_, false><<<g2, 256, 0, st>>>(static_cast<const T_ *>(src), ld_s, bs_s, static_cast<T_ *>(dst), ld_d, bs_d, rows, cols / V_, rpb); \
    } while (0)
        if (dtype == 0) {
            if (V == 4) MLAGG_WALK(float, 4);
            else MLAGG_WALK(float, 2);
        } else {
            if (V == 8) MLAGG_WALK(__nv_bfloat16, 8);
            else MLAGG_WALK(__nv_bfloat16, 4);
        }
#undef MLAGG_WALK
        return cudaGetLastError();
    }
    if (add) return cudaErrorNotSupported;
    const long long n = rows * (cols / V);
    long long bx = (n + 255) / 256;
    const long long cap = (148LL * 16 + batch - 1) / batch;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    const dim3 grid((unsigned)bx, (unsigned)batch);
#define MLAGG_COPY(T_, V_) \
    copy_rows_kernel<T_, V_><<<grid, 256, 0, st>>>(static_cast<const T_ *>(src), ld_s, bs_s, static_cast<T_ *>(dst), ld_d, bs_d, rows, cols / V_)
    if (dtype == 0) {
        if (V == 4) MLAGG_COPY(float, 4);
        else if (V == 2) MLAGG_COPY(float, 2);
        else MLAGG_COPY(float, 1);
    } else {
        if (V == 8) MLAGG_COPY(__nv_bfloat16, 8);
        else if (V == 4) MLAGG_COPY(__nv_bfloat16, 4);
        else MLAGG_COPY(__nv_bfloat16, 1);
    }
#undef MLAGG_COPY
    return cudaGetLastError();
}

cudaError_t silu_gate_dispatch(const void *t, const void *z, const void *g, void *o1, void *o2, long long n, int dtype,
                               bool bwd, cudaStream_t st) {
    const long long n4 = n / 4;
    if (dtype == 0) {
        auto *tt = static_cast<const float *>(t), *zz = static_cast<const float *>(z);
        if (!bwd) silu_gate_fwd_kernel<float><<<ew_blocks(n4), 256, 0, st>>>(tt, zz, static_cast<float *>(o1), n4);
        else silu_gate_bwd_kernel<float><<<ew_blocks(n4), 256, 0, st>>>(tt, zz, static_cast<const float *>(g), static_cast<float *>(o1), static_cast<float *>(o2), n4);
    } else {
        auto *tt = static_cast<const __nv_bfloat16 *>(t), *zz = static_cast<const __nv_bfloat16 *>(z);
        if (!bwd) silu_gate_fwd_kernel<__nv_bfloat16><<<ew_blocks(n4), 256, 0, st>>>(tt, zz, static_cast<__nv_bfloat16 *>(o1), n4);
        else silu_gate_bwd_kernel<__nv_bfloat16><<<ew_blocks(n4), 256, 0, st>>>(tt, zz, static_cast<const __nv_bfloat16 *>(g), static_cast<__nv_bfloat16 *>(o1), static_cast<__nv_bfloat16 *>(o2), n4);
    }
    return cudaGetLastError();
}

cudaError_t diff_lambda_dispatch(const float *q1, const float *k1, const float *q2, const float *k2, int n, float init,
                                 float *out, const float *dlam, float *grads, cudaStream_t st) {
    if (!grads) diff_lambda_fwd_kernel<<<1, 32, 0, st>>>(q1, k1, q2, k2, n, init, out);
    else diff_lambda_bwd_kernel<<<1, 32, 0, st>>>(q1, k1, q2, k2, out, dlam, n, grads);
    return cudaGetLastError();
}

}  // namespace mlagg
