// layernorm.cu -- LayerNorm over the channel dimension of tokens-major activations (rows = tokens), fp32 math with
// fp32 or bf16 input AND output.  Replaces nn.LayerNorm at reference nnUNetTrainer_MLAgg_2D_dt_MS.py:848,871,887,907
// (norm1 / norm2 of MLLABlock), :670,723 (pooled-token norm), MambaSkip.py:344,536 (out_norm), :686,690,741,742 (ln_1 /
// norm2 of VSS_Conv_Block).  Under the reference's autocast, layer_norm runs in fp32: the input is up-cast by a copy
// kernel, the fp32 result is written, and every consuming Linear down-casts it again with another copy.  Here one
// kernel reads the activation once in its storage type and writes the consumer's type.  HBM-bound:
// algorithmic bytes = M*C*(e_in + e_out) + 8*M (mean, rstd saved for the backward pass).
// One warp per row; a lane owns NV groups of 4 consecutive channels (128-bit fp32 / 64-bit bf16 accesses).
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace mlagg {

template <typename T>
__device__ __forceinline__ float4 ln_ld4(const T *p);
template <>
__device__ __forceinline__ float4 ln_ld4<float>(const float *p) { return *reinterpret_cast<const float4 *>(p); }
template <>
__device__ __forceinline__ float4 ln_ld4<__nv_bfloat16>(const __nv_bfloat16 *p) {
    const uint2 raw = *reinterpret_cast<const uint2 *>(p);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&raw.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&raw.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
template <typename T>
__device__ __forceinline__ void ln_st4(T *p, float4 v);
template <>
__device__ __forceinline__ void ln_st4<float>(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }
template <>
__device__ __forceinline__ void ln_st4<__nv_bfloat16>(__nv_bfloat16 *p, float4 v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 raw;
    raw.x = *reinterpret_cast<uint32_t *>(&a);
    raw.y = *reinterpret_cast<uint32_t *>(&b);
    *reinterpret_cast<uint2 *>(p) = raw;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <typename TI, typename TO, int NV>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const TI *__restrict__ x, const float *__restrict__ w,
                                                            const float *__restrict__ b, TO *__restrict__ y,
                                                            float *__restrict__ mean, float *__restrict__ rstd,
                                                            long long M, int C, float eps) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    float4 wv[NV], bv[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (lane + 32 * i) * 4;
        wv[i] = c < C ? __ldg(reinterpret_cast<const float4 *>(w + c)) : make_float4(0, 0, 0, 0);
        bv[i] = (c < C && b) ? __ldg(reinterpret_cast<const float4 *>(b + c)) : make_float4(0, 0, 0, 0);
    }
    const float invC = 1.f / C;
    for (long long row = warp0; row < M; row += nwarps) {
        float4 v[NV];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = (lane + 32 * i) * 4;
            v[i] = c < C ? ln_ld4<TI>(x + row * C + c) : make_float4(0, 0, 0, 0);
            s += v[i].x + v[i].y + v[i].z + v[i].w;
        }
        const float mu = warp_sum(s) * invC;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = (lane + 32 * i) * 4;
            if (c < C) {
                const float a0 = v[i].x - mu, a1 = v[i].y - mu, a2 = v[i].z - mu, a3 = v[i].w - mu;
                q += a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3;
            }
        }
        const float rs = rsqrtf(warp_sum(q) * invC + eps);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = (lane + 32 * i) * 4;
            if (c < C) {
                float4 o;
                o.x = (v[i].x - mu) * rs * wv[i].x + bv[i].x;
                o.y = (v[i].y - mu) * rs * wv[i].y + bv[i].y;
                o.z = (v[i].z - mu) * rs * wv[i].z + bv[i].z;
                o.w = (v[i].w - mu) * rs * wv[i].w + bv[i].w;
                ln_st4<TO>(y + row * C + c, o);
            }
        }
        if (lane == 0) {
            mean[row] = mu;
            rstd[row] = rs;
        }
    }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * w;  dw += dy * xhat, db += dy (per-lane partials over the
// rows this warp visits -> shared memory -> one atomic per channel per block)
template <typename TI, typename TO, int NV>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const TI *__restrict__ x, const float *__restrict__ w,
                                                            const float *__restrict__ mean,
                                                            const float *__restrict__ rstd, const TO *__restrict__ dy,
                                                            TI *__restrict__ dx, float *__restrict__ dw,
                                                            float *__restrict__ db, long long M, int C,
                                                            const TI *__restrict__ dres) {
    extern __shared__ float red[];  // [2][C]
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    float4 wv[NV], aw[NV], ab[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (lane + 32 * i) * 4;
        wv[i] = c < C ? __ldg(reinterpret_cast<const float4 *>(w + c)) : make_float4(0, 0, 0, 0);
        aw[i] = ab[i] = make_float4(0, 0, 0, 0);
    }
    const float invC = 1.f / C;
    for (long long row = warp0; row < M; row += nwarps) {
        const float mu = mean[row], rs = rstd[row];
        float4 xh[NV], g[NV];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = (lane + 32 * i) * 4;
            if (c < C) {
                const float4 xv = ln_ld4<TI>(x + row * C + c), gv = ln_ld4<TO>(dy + row * C + c);
                xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
                aw[i].x += gv.x * xh[i].x; aw[i].y += gv.y * xh[i].y; aw[i].z += gv.z * xh[i].z; aw[i].w += gv.w * xh[i].w;
                ab[i].x += gv.x; ab[i].y += gv.y; ab[i].z += gv.z; ab[i].w += gv.w;
                g[i] = make_float4(gv.x * wv[i].x, gv.y * wv[i].y, gv.z * wv[i].z, gv.w * wv[i].w);
                s1 += g[i].x + g[i].y + g[i].z + g[i].w;
                s2 += g[i].x * xh[i].x + g[i].y * xh[i].y + g[i].z * xh[i].z + g[i].w * xh[i].w;
            } else {
                xh[i] = g[i] = make_float4(0, 0, 0, 0);
            }
        }
        const float c1 = warp_sum(s1) * invC, c2 = warp_sum(s2) * invC;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = (lane + 32 * i) * 4;
            if (c < C) {
                float4 o;
                o.x = rs * (g[i].x - c1 - xh[i].x * c2);
                o.y = rs * (g[i].y - c1 - xh[i].y * c2);
                o.z = rs * (g[i].z - c1 - xh[i].z * c2);
                o.w = rs * (g[i].w - c1 - xh[i].w * c2);
                if (dres != nullptr) {   // gradient arriving at x along the residual path: dx = LN'(dy) + dres in one pass
                    const float4 rv = ln_ld4<TI>(dres + row * C + c);
                    o.x += rv.x; o.y += rv.y; o.z += rv.z; o.w += rv.w;
                }
                ln_st4<TI>(dx + row * C + c, o);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (lane + 32 * i) * 4;
        if (c < C) {
            atomicAdd(&red[c], aw[i].x); atomicAdd(&red[c + 1], aw[i].y); atomicAdd(&red[c + 2], aw[i].z); atomicAdd(&red[c + 3], aw[i].w);
            atomicAdd(&red[C + c], ab[i].x); atomicAdd(&red[C + c + 1], ab[i].y); atomicAdd(&red[C + c + 2], ab[i].z); atomicAdd(&red[C + c + 3], ab[i].w);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
        atomicAdd(dw + i, red[i]);
        if (db) atomicAdd(db + i, red[C + i]);
    }
}

// ---------------------------------------------------------------- sub-warp rows (C % V == 0, V = 8 bf16 / 4 with fp32)
// One warp per row leaves 20 of 32 lanes idle at C = 48 and keeps one 8-byte load per lane in flight: 30 - 36 % of the HBM
// roofline (tools/call_shapes.py).  Here a row belongs to a group of G = 8 | 16 | 32 lanes with one 16-byte vector per
// lane and pass, a warp works on U * 32 / G consecutive rows per iteration with every load issued before the first
// reduction, and the row reductions are G-lane butterflies.
template <typename T, int V>
struct LnVec;
template <>
struct LnVec<float, 4> {
    using Raw = float4;
    static __device__ __forceinline__ Raw ldr(const float *p) { return *reinterpret_cast<const float4 *>(p); }
    static __device__ __forceinline__ void un(const Raw &t, float (&v)[4]) { v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w; }
    static __device__ __forceinline__ void ld(const float *p, float (&v)[4]) { un(ldr(p), v); }
    static __device__ __forceinline__ void st(float *p, const float (&v)[4]) {
        *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <>
struct LnVec<__nv_bfloat16, 4> {
    using Raw = uint2;
    static __device__ __forceinline__ Raw ldr(const __nv_bfloat16 *p) { return *reinterpret_cast<const uint2 *>(p); }
    static __device__ __forceinline__ void un(const Raw &t, float (&v)[4]) {
        v[0] = __uint_as_float(t.x << 16), v[1] = __uint_as_float(t.x & 0xffff0000u);
        v[2] = __uint_as_float(t.y << 16), v[3] = __uint_as_float(t.y & 0xffff0000u);
    }
    static __device__ __forceinline__ void ld(const __nv_bfloat16 *p, float (&v)[4]) { un(ldr(p), v); }
    static __device__ __forceinline__ void st(__nv_bfloat16 *p, const float (&v)[4]) {
        const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
        uint2 raw;
        raw.x = *reinterpret_cast<const uint32_t *>(&a), raw.y = *reinterpret_cast<const uint32_t *>(&b);
        *reinterpret_cast<uint2 *>(p) = raw;
    }
};
template <>
struct LnVec<__nv_bfloat16, 8> {
    using Raw = uint4;
    static __device__ __forceinline__ Raw ldr(const __nv_bfloat16 *p) { return *reinterpret_cast<const uint4 *>(p); }
    static __device__ __forceinline__ void un(const Raw &t, float (&v)[8]) {
        const uint32_t r[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) v[2 * i] = __uint_as_float(r[i] << 16), v[2 * i + 1] = __uint_as_float(r[i] & 0xffff0000u);
    }
    static __device__ __forceinline__ void ld(const __nv_bfloat16 *p, float (&v)[8]) { un(ldr(p), v); }
    static __device__ __forceinline__ void st(__nv_bfloat16 *p, const float (&v)[8]) {
        uint32_t r[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const __nv_bfloat162 a = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
            r[i] = *reinterpret_cast<const uint32_t *>(&a);
        }
        *reinterpret_cast<uint4 *>(p) = make_uint4(r[0], r[1], r[2], r[3]);
    }
};
template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <typename TI, typename TO>
struct LnV {
    static constexpr int N = (sizeof(TI) == 2 && sizeof(TO) == 2) ? 8 : 4;
};

template <typename TI, typename TO, int G, int NV, int U>
__global__ void __launch_bounds__(256, NV == 1 ? 3 : 2) layernorm_fwd_rows_kernel(const TI *__restrict__ x, const float *__restrict__ w,
                                                                 const float *__restrict__ b, TO *__restrict__ y,
                                                                 float *__restrict__ mean, float *__restrict__ rstd,
                                                                 long long M, int C, float eps) {
    constexpr int V = LnV<TI, TO>::N, RPW = 32 / G;
    const int lane = threadIdx.x & 31, gl = lane % G, grp = lane / G;
    const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    float wv[NV][V], bv[NV][V];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (gl + G * i) * V;
#pragma unroll
        for (int k = 0; k < V; ++k) {
            wv[i][k] = c < C ? __ldg(w + c + k) : 0.f;
            bv[i][k] = (c < C && b) ? __ldg(b + c + k) : 0.f;
        }
    }
    const float invC = 1.f / C;
    for (long long r0 = warp * (RPW * U); r0 < M; r0 += nwarps * (RPW * U)) {
        typename LnVec<TI, V>::Raw raw[U][NV];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long row = r0 + u * RPW + grp;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int c = (gl + G * i) * V;
                if (row < M && c < C) raw[u][i] = LnVec<TI, V>::ldr(x + row * C + c);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long row = r0 + u * RPW + grp;
            float v[1][NV][V];
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                if (row < M && (gl + G * i) * V < C) {
                    LnVec<TI, V>::un(raw[u][i], v[0][i]);
                } else {
#pragma unroll
                    for (int k = 0; k < V; ++k) v[0][i][k] = 0.f;
                }
#pragma unroll
                for (int k = 0; k < V; ++k) s += v[0][i][k];
            }
            const float mu = group_sum<G>(s) * invC;
            float q = 0.f;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                if ((gl + G * i) * V < C) {
#pragma unroll
                    for (int k = 0; k < V; ++k) {
                        const float a = v[0][i][k] - mu;
                        q = fmaf(a, a, q);
                    }
                }
            }
            const float rs = rsqrtf(group_sum<G>(q) * invC + eps);
            if (row < M) {
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    const int c = (gl + G * i) * V;
                    if (c < C) {
                        float o[V];
#pragma unroll
                        for (int k = 0; k < V; ++k) o[k] = (v[0][i][k] - mu) * rs * wv[i][k] + bv[i][k];
                        LnVec<TO, V>::st(y + row * C + c, o);
                    }
                }
                if (gl == 0) mean[row] = mu, rstd[row] = rs;
            }
        }
    }
}

template <typename TI, typename TO, int G, int NV, int U>
__global__ void __launch_bounds__(256, NV <= 2 ? 2 : 1) layernorm_bwd_rows_kernel(const TI *__restrict__ x, const float *__restrict__ w,
                                                                 const float *__restrict__ mean,
                                                                 const float *__restrict__ rstd,
                                                                 const TO *__restrict__ dy, TI *__restrict__ dx,
                                                                 float *__restrict__ dw, float *__restrict__ db,
                                                                 long long M, int C, const TI *__restrict__ dres) {
    constexpr int V = LnV<TI, TO>::N, RPW = 32 / G;
    extern __shared__ float red[];  // [2][C]
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31, gl = lane % G, grp = lane / G;
    const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    float wv[NV][V], aw[NV][V], ab[NV][V];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (gl + G * i) * V;
#pragma unroll
        for (int k = 0; k < V; ++k) {
            wv[i][k] = c < C ? __ldg(w + c + k) : 0.f;
            aw[i][k] = ab[i][k] = 0.f;
        }
    }
    const float invC = 1.f / C;
    for (long long r0 = warp * (RPW * U); r0 < M; r0 += nwarps * (RPW * U)) {
        typename LnVec<TI, V>::Raw rx[U][NV], rr[U][NV];
        typename LnVec<TO, V>::Raw rg[U][NV];
        float mu[U], rs[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long row = r0 + u * RPW + grp;
            mu[u] = row < M ? mean[row] : 0.f;
            rs[u] = row < M ? rstd[row] : 0.f;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int c = (gl + G * i) * V;
                if (row < M && c < C) {
                    rx[u][i] = LnVec<TI, V>::ldr(x + row * C + c);
                    rg[u][i] = LnVec<TO, V>::ldr(dy + row * C + c);
                    if (dres != nullptr) rr[u][i] = LnVec<TI, V>::ldr(dres + row * C + c);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long row = r0 + u * RPW + grp;
            float xh[NV][V], g[NV][V];
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                if (row < M && (gl + G * i) * V < C) {
                    LnVec<TI, V>::un(rx[u][i], xh[i]);
                    LnVec<TO, V>::un(rg[u][i], g[i]);
#pragma unroll
                    for (int k = 0; k < V; ++k) {
                        const float h = (xh[i][k] - mu[u]) * rs[u];
                        xh[i][k] = h;
                        aw[i][k] = fmaf(g[i][k], h, aw[i][k]);
                        ab[i][k] += g[i][k];
                        const float gw = g[i][k] * wv[i][k];
                        g[i][k] = gw;
                        s1 += gw;
                        s2 = fmaf(gw, h, s2);
                    }
                }
            }
            const float c1 = group_sum<G>(s1) * invC, c2 = group_sum<G>(s2) * invC;
            if (row < M) {
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    const int c = (gl + G * i) * V;
                    if (c < C) {
                        float o[V];
#pragma unroll
                        for (int k = 0; k < V; ++k) o[k] = rs[u] * (g[i][k] - c1 - xh[i][k] * c2);
                        if (dres != nullptr) {   // gradient arriving along the residual path, added in the same pass
                            float rv[V];
                            LnVec<TI, V>::un(rr[u][i], rv);
#pragma unroll
                            for (int k = 0; k < V; ++k) o[k] += rv[k];
                        }
                        LnVec<TI, V>::st(dx + row * C + c, o);
                    }
                }
            }
        }
    }
    // groups of one warp hold partial sums of the same channels: butterflies across the groups, then one shared-memory
    // atomic per (warp, channel) instead of one per lane
#pragma unroll
    for (int i = 0; i < NV; ++i) {
#pragma unroll
        for (int k = 0; k < V; ++k) {
#pragma unroll
            for (int o = G; o < 32; o <<= 1) {
                aw[i][k] += __shfl_xor_sync(0xffffffffu, aw[i][k], o);
                ab[i][k] += __shfl_xor_sync(0xffffffffu, ab[i][k], o);
            }
        }
        const int c = (gl + G * i) * V;
        if (c < C && grp == 0) {
#pragma unroll
            for (int k = 0; k < V; ++k) {
                atomicAdd(&red[c + k], aw[i][k]);
                atomicAdd(&red[C + c + k], ab[i][k]);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
        atomicAdd(dw + i, red[i]);
        if (db) atomicAdd(db + i, red[C + i]);
    }
}

template <typename TI, typename TO, int G, int NV>
static cudaError_t ln_rows_launch(const void *x, const float *w, const float *b, void *y, float *mean, float *rstd,
                                  const void *dy, void *dx, float *dw, float *db, long long M, int C, float eps, bool bwd,
                                  cudaStream_t st, const void *dres) {
    constexpr int UF = NV == 1 ? 4 : NV == 2 ? 2 : 1, UB = NV == 1 ? 4 : 1;   // rows per group and iteration, forward / backward
    // persistent grid: exactly the CTAs that are resident at once (short-lived CTAs paid their start-up -- zeroing the
    // reduction buffer, the parameter loads -- and their tail of atomics four waves in a row)
    const long long rows_per_block = 8LL * (32 / G) * (bwd ? UB : UF);
    const long long need = (M + rows_per_block - 1) / rows_per_block;
    static int occ_f = 0, occ_b = 0;
    int &occ = bwd ? occ_b : occ_f;
    if (occ == 0) {
        cudaError_t e = bwd ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, layernorm_bwd_rows_kernel<TI, TO, G, NV, UB>, 256, 2 * C * sizeof(float))
                            : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, layernorm_fwd_rows_kernel<TI, TO, G, NV, UF>, 256, 0);
        if (e != cudaSuccess) return e;
        if (occ < 1) occ = 1;
    }
    const long long blocks = std::min<long long>(need, 148LL * occ);
    if (!bwd)
        layernorm_fwd_rows_kernel<TI, TO, G, NV, UF><<<(int)blocks, 256, 0, st>>>(static_cast<const TI *>(x), w, b, static_cast<TO *>(y), mean, rstd, M, C, eps);
    else
        layernorm_bwd_rows_kernel<TI, TO, G, NV, UB><<<(int)blocks, 256, 2 * C * sizeof(float), st>>>(
            static_cast<const TI *>(x), w, mean, rstd, static_cast<const TO *>(dy), static_cast<TI *>(dx), dw, db, M, C, static_cast<const TI *>(dres));
    return cudaGetLastError();
}
// false: shape not taken (C % V != 0 or more than 8 vectors per lane)
template <typename TI, typename TO>
static bool ln_rows(cudaError_t &e, const void *x, const float *w, const float *b, void *y, float *mean, float *rstd,
                    const void *dy, void *dx, float *dw, float *db, long long M, int C, float eps, bool bwd, cudaStream_t st,
                    const void *dres) {
    constexpr int V = LnV<TI, TO>::N;
    if (C % V != 0 || getenv("MLAGG_LAYERNORM_OLD")) return false;
    const int nvec = C / V;
#define LN_GO(G, NV) e = ln_rows_launch<TI, TO, G, NV>(x, w, b, y, mean, rstd, dy, dx, dw, db, M, C, eps, bwd, st, dres); return true
    if (nvec <= 8) { LN_GO(8, 1); }
    if (nvec <= 16) { LN_GO(16, 1); }
    if (nvec <= 32) { LN_GO(32, 1); }
    if (nvec <= 64) { LN_GO(32, 2); }
    if (nvec <= 96) { LN_GO(32, 3); }
    if (nvec <= 128) { LN_GO(32, 4); }
#undef LN_GO
    return false;
}

template <typename TI, typename TO, int NV>
static cudaError_t ln_launch(const void *x, const float *w, const float *b, void *y, float *mean, float *rstd,
                             const void *dy, void *dx, float *dw, float *db, long long M, int C, float eps, bool bwd,
                             cudaStream_t st, const void *dres) {
    long long blocks = (M + 7) / 8;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (!bwd) {
        layernorm_fwd_kernel<TI, TO, NV><<<(int)blocks, 256, 0, st>>>(static_cast<const TI *>(x), w, b,
                                                                      static_cast<TO *>(y), mean, rstd, M, C, eps);
    } else {
        layernorm_bwd_kernel<TI, TO, NV><<<(int)blocks, 256, 2 * C * sizeof(float), st>>>(
            static_cast<const TI *>(x), w, mean, rstd, static_cast<const TO *>(dy), static_cast<TI *>(dx), dw, db, M, C,
            static_cast<const TI *>(dres));
    }
    return cudaGetLastError();
}

template <typename TI, typename TO>
static cudaError_t ln_nv(int nv, const void *x, const float *w, const float *b, void *y, float *mean, float *rstd,
                         const void *dy, void *dx, float *dw, float *db, long long M, int C, float eps, bool bwd,
                         cudaStream_t st, const void *dres) {
    switch (nv) {
        case 1: return ln_launch<TI, TO, 1>(x, w, b, y, mean, rstd, dy, dx, dw, db, M, C, eps, bwd, st, dres);
        case 2: return ln_launch<TI, TO, 2>(x, w, b, y, mean, rstd, dy, dx, dw, db, M, C, eps, bwd, st, dres);
        case 3: return ln_launch<TI, TO, 3>(x, w, b, y, mean, rstd, dy, dx, dw, db, M, C, eps, bwd, st, dres);
        case 4: return ln_launch<TI, TO, 4>(x, w, b, y, mean, rstd, dy, dx, dw, db, M, C, eps, bwd, st, dres);
        case 5: case 6: return ln_launch<TI, TO, 6>(x, w, b, y, mean, rstd, dy, dx, dw, db, M, C, eps, bwd, st, dres);
        case 7: case 8: return ln_launch<TI, TO, 8>(x, w, b, y, mean, rstd, dy, dx, dw, db, M, C, eps, bwd, st, dres);
        default: return cudaErrorInvalidValue;
    }
}

// dt_in / dt_out: 0 = fp32, 1 = bf16
cudaError_t layernorm_dispatch(const void *x, const float *w, const float *b, void *y, float *mean, float *rstd,
                               const void *dy, void *dx, float *dw, float *db, long long M, int C, float eps,
                               int dt_in, int dt_out, bool bwd, cudaStream_t st, const void *dres) {
    const int nv = (C / 4 + 31) / 32;
    using bf = __nv_bfloat16;
    cudaError_t e = cudaSuccess;
    if (dt_in == 0 && dt_out == 0 && ln_rows<float, float>(e, x, w, b, y, mean, rstd, dy, dx, dw, db, M, C, eps, bwd, st, dres)) return e;
    if (dt_in == 0 && dt_out == 1 && ln_rows<float, bf>(e, x, w, b, y, mean, rstd, dy, dx, dw, db, M, C, eps, bwd, st, dres)) return e;
    if (dt_in == 1 && dt_out == 0 && ln_rows<bf, float>(e, x, w, b, y, mean, rstd, dy, dx, dw, db, M, C, eps, bwd, st, dres)) return e;
    if (dt_in == 1 && dt_out == 1 && ln_rows<bf, bf>(e, x, w, b, y, mean, rstd, dy, dx, dw, db, M, C, eps, bwd, st, dres)) return e;
    if (dt_in == 0 && dt_out == 0) return ln_nv<float, float>(nv, x, w, b, y, mean, rstd, dy, dx, dw, db, M, C, eps, bwd, st, dres);
    if (dt_in == 0 && dt_out == 1) return ln_nv<float, bf>(nv, x, w, b, y, mean, rstd, dy, dx, dw, db, M, C, eps, bwd, st, dres);
    if (dt_in == 1 && dt_out == 0) return ln_nv<bf, float>(nv, x, w, b, y, mean, rstd, dy, dx, dw, db, M, C, eps, bwd, st, dres);
    return ln_nv<bf, bf>(nv, x, w, b, y, mean, rstd, dy, dx, dw, db, M, C, eps, bwd, st, dres);
}

}  // namespace mlagg
