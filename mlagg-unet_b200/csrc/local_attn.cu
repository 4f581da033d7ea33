// local_attn.cu -- fused 3x3-window differential softmax attention + sub-LayerNorm (RMSNorm) of the MLAgg block.
// Replaces the op chain at reference nnUNetTrainer_MLAgg_2D_dt_MS.py:698-717 (2x nn.Unfold materialising 9 copies of
// k and v, a batched 1 x hd @ hd x 9 matmul, masked_fill, softmax, lambda-combine, 1 x 9 @ 9 x 2hd matmul, RMSNorm,
// x (1 - lambda_init)); math: SURVEY.md App. A.4.  Tokens-major I/O, nothing is materialised:
//   q (B, N, 2h, hd)            row stride ldq   (raw projection; `scale` = hd**-0.5 is applied here)
//   k (B, N, 2h, hd), v (B, N, h, 2hd)  row stride ldkv  (the two halves of the kv Linear output)
//   out (B, N, h, 2hd)          row stride ldo
// One thread owns one (token, head pair): q and the 2hd-wide output live in registers, the 9 neighbours of k and v
// are read with 128-bit loads (they are shared with the neighbouring tokens through L1/L2).
// Backward is two gather passes (no atomics on activations):
//   p1: recompute, RMSNorm / softmax backward -> dq, and per-token scratch dO (C), Abar (h*9), dlogit (2h*9);
//       d lambda and d subln_w are block-reduced then added atomically;
//   p2: dk[i], dv[i] gathered from the <= 9 tokens that have i in their window.
#include <cuda_bf16.h>

#include "common.cuh"

namespace mlagg {

template <typename T>
__device__ __forceinline__ float ldf(const T *p);
template <>
__device__ __forceinline__ float ldf<float>(const float *p) { return __ldg(p); }
template <>
__device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16 *p) {
    return __bfloat162float(__ldg(reinterpret_cast<const __nv_bfloat16 *>(p)));
}
template <typename T>
__device__ __forceinline__ void stf(T *p, float v);
template <>
__device__ __forceinline__ void stf<float>(float *p, float v) { *p = v; }
template <>
__device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16 *p, float v) { *p = __float2bfloat16_rn(v); }

// vector of V consecutive elements -> float[V]; V % 4 == 0 uses 128-bit (fp32) / 64-bit (bf16) loads
template <typename T, int V>
__device__ __forceinline__ void ldv(const T *p, float *dst) {
    if constexpr (V % 4 == 0) {
#pragma unroll
        for (int i = 0; i < V; i += 4) {
            if constexpr (sizeof(T) == 4) {
                const float4 t = __ldg(reinterpret_cast<const float4 *>(p + i));
                dst[i] = t.x; dst[i + 1] = t.y; dst[i + 2] = t.z; dst[i + 3] = t.w;
            } else {
                const uint2 raw = __ldg(reinterpret_cast<const uint2 *>(p + i));
                const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&raw.x));
                const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&raw.y));
                dst[i] = a.x; dst[i + 1] = a.y; dst[i + 2] = b.x; dst[i + 3] = b.y;
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < V; ++i) dst[i] = ldf<T>(p + i);
    }
}
template <typename T, int V>
__device__ __forceinline__ void stv(T *p, const float *src) {
    if constexpr (V % 4 == 0) {
#pragma unroll
        for (int i = 0; i < V; i += 4) {
            if constexpr (sizeof(T) == 4) {
                *reinterpret_cast<float4 *>(p + i) = make_float4(src[i], src[i + 1], src[i + 2], src[i + 3]);
            } else {
                __nv_bfloat162 a = __floats2bfloat162_rn(src[i], src[i + 1]), b = __floats2bfloat162_rn(src[i + 2], src[i + 3]);
                uint2 raw;
                raw.x = *reinterpret_cast<uint32_t *>(&a);
                raw.y = *reinterpret_cast<uint32_t *>(&b);
                *reinterpret_cast<uint2 *>(p + i) = raw;
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < V; ++i) stf<T>(p + i, src[i]);
    }
}

struct LocalAttnParams {
    const void *q, *k, *v, *dout;
    void *out, *dq, *dk, *dv;
    const float *subln_w;     // (2hd)
    float *d_subln_w;         // (2hd), accumulated
    float *d_lambda;          // (1), accumulated
    float *ws_dO, *ws_abar, *ws_dlog;   // scratch: (B,N,h,2hd), (B,N,h,9), (B,N,2h,9) fp32
    long long ldq, ldkv, ldo, lddo, lddq, lddkv;   // row strides in elements
    int Bn, H, W, h;
    const float *lamp;        // device scalar lambda_full
    float scale, eps, post;   // post = 1 - lambda_init
};

// shared forward core: logits -> A0, A1 -> Abar -> o (pre-norm); returns r = rsqrt(mean(o^2)+eps)
template <typename T, int HD>
__device__ __forceinline__ float local_core(const LocalAttnParams &p, long long tok, int hr, int wc, int m,
                                            const float (&qv)[2][HD], float (&A)[2][9], float (&abar)[9],
                                            float (&o)[2 * HD]) {
    const T *kb = static_cast<const T *>(p.k), *vb = static_cast<const T *>(p.v);
    float lg[2][9];
    bool ok[9];
#pragma unroll
    for (int pp = 0; pp < 9; ++pp) {
        const int rr = hr + pp / 3 - 1, cc = wc + pp % 3 - 1;
        ok[pp] = rr >= 0 && rr < p.H && cc >= 0 && cc < p.W;
        lg[0][pp] = lg[1][pp] = -INFINITY;
        if (ok[pp]) {
            const long long nt = tok + (long long)(pp / 3 - 1) * p.W + (pp % 3 - 1);
            float kv[2 * HD];
            ldv<T, 2 * HD>(kb + nt * p.ldkv + (long long)m * 2 * HD, kv);
            float d0 = 0.f, d1 = 0.f;
#pragma unroll
            for (int c = 0; c < HD; ++c) {
                d0 = fmaf(qv[0][c], kv[c], d0);
                d1 = fmaf(qv[1][c], kv[HD + c], d1);
            }
            lg[0][pp] = d0 * p.scale;
            lg[1][pp] = d1 * p.scale;
        }
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        float mx = -INFINITY;
#pragma unroll
        for (int pp = 0; pp < 9; ++pp) mx = fmaxf(mx, lg[j][pp]);
        float s = 0.f;
#pragma unroll
        for (int pp = 0; pp < 9; ++pp) {
            A[j][pp] = ok[pp] ? expf(lg[j][pp] - mx) : 0.f;
            s += A[j][pp];
        }
        const float inv = 1.f / s;
#pragma unroll
        for (int pp = 0; pp < 9; ++pp) A[j][pp] *= inv;
    }
    const float lam = __ldg(p.lamp);
#pragma unroll
    for (int c = 0; c < 2 * HD; ++c) o[c] = 0.f;
#pragma unroll
    for (int pp = 0; pp < 9; ++pp) {
        abar[pp] = A[0][pp] - lam * A[1][pp];
        if (ok[pp]) {
            const long long nt = tok + (long long)(pp / 3 - 1) * p.W + (pp % 3 - 1);
            float vv[2 * HD];
            ldv<T, 2 * HD>(vb + nt * p.ldkv + (long long)m * 2 * HD, vv);
#pragma unroll
            for (int c = 0; c < 2 * HD; ++c) o[c] = fmaf(abar[pp], vv[c], o[c]);
        }
    }
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < 2 * HD; ++c) ss = fmaf(o[c], o[c], ss);
    return 1.f / sqrtf(ss * (1.f / (2 * HD)) + p.eps);
}

template <typename T, int HD>
__global__ void __launch_bounds__(128) local_attn_fwd_kernel(const LocalAttnParams p) {
    const long long total = (long long)p.Bn * p.H * p.W * p.h;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int m = (int)(idx % p.h);
    const long long tok = idx / p.h;
    const int wc = (int)(tok % p.W), hr = (int)((tok / p.W) % p.H);
    float qv[2][HD];
    ldv<T, HD>(static_cast<const T *>(p.q) + tok * p.ldq + (long long)m * 2 * HD, qv[0]);
    ldv<T, HD>(static_cast<const T *>(p.q) + tok * p.ldq + (long long)m * 2 * HD + HD, qv[1]);
    float A[2][9], abar[9], o[2 * HD];
    const float r = local_core<T, HD>(p, tok, hr, wc, m, qv, A, abar, o);
#pragma unroll
    for (int c = 0; c < 2 * HD; ++c) o[c] = o[c] * r * __ldg(p.subln_w + c) * p.post;
    stv<T, 2 * HD>(static_cast<T *>(p.out) + tok * p.ldo + (long long)m * 2 * HD, o);
}

template <typename T, int HD>
__global__ void __launch_bounds__(128) local_attn_bwd_p1_kernel(const LocalAttnParams p) {
    __shared__ float red[2 * HD + 1];
    for (int i = threadIdx.x; i < 2 * HD + 1; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
    const long long total = (long long)p.Bn * p.H * p.W * p.h;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float dlam = 0.f;
    float dw[2 * HD];
#pragma unroll
    for (int c = 0; c < 2 * HD; ++c) dw[c] = 0.f;
    if (idx < total) {
        const int m = (int)(idx % p.h);
        const long long tok = idx / p.h;
        const int wc = (int)(tok % p.W), hr = (int)((tok / p.W) % p.H);
        float qv[2][HD];
        ldv<T, HD>(static_cast<const T *>(p.q) + tok * p.ldq + (long long)m * 2 * HD, qv[0]);
        ldv<T, HD>(static_cast<const T *>(p.q) + tok * p.ldq + (long long)m * 2 * HD + HD, qv[1]);
        float A[2][9], abar[9], o[2 * HD];
        const float r = local_core<T, HD>(p, tok, hr, wc, m, qv, A, abar, o);
        // RMSNorm backward: out = o * r * w * post
        float g[2 * HD];
        ldv<T, 2 * HD>(static_cast<const T *>(p.dout) + tok * p.lddo + (long long)m * 2 * HD, g);
        float dot = 0.f;
#pragma unroll
        for (int c = 0; c < 2 * HD; ++c) {
            dw[c] = g[c] * p.post * o[c] * r;
            g[c] *= p.post * __ldg(p.subln_w + c);
            dot = fmaf(g[c], o[c], dot);
        }
        const float k3 = r * r * r * dot * (1.f / (2 * HD));
        float dO[2 * HD];
#pragma unroll
        for (int c = 0; c < 2 * HD; ++c) dO[c] = r * g[c] - o[c] * k3;
        stv<float, 2 * HD>(p.ws_dO + (tok * p.h + m) * 2 * HD, dO);
        // d Abar_p = dO . v[nbr_p]
        const T *vb = static_cast<const T *>(p.v), *kb = static_cast<const T *>(p.k);
        float dab[9];
#pragma unroll
        for (int pp = 0; pp < 9; ++pp) {
            const int rr = hr + pp / 3 - 1, cc = wc + pp % 3 - 1;
            dab[pp] = 0.f;
            if (rr >= 0 && rr < p.H && cc >= 0 && cc < p.W) {
                const long long nt = tok + (long long)(pp / 3 - 1) * p.W + (pp % 3 - 1);
                float vv[2 * HD];
                ldv<T, 2 * HD>(vb + nt * p.ldkv + (long long)m * 2 * HD, vv);
                float d = 0.f;
#pragma unroll
                for (int c = 0; c < 2 * HD; ++c) d = fmaf(dO[c], vv[c], d);
                dab[pp] = d;
            }
        }
        // softmax backward of both maps; dA0 = dab, dA1 = -lam * dab
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int pp = 0; pp < 9; ++pp) {
            s0 = fmaf(A[0][pp], dab[pp], s0);
            s1 = fmaf(A[1][pp], dab[pp], s1);
        }
        dlam = -s1;
        float dl[2][9];
#pragma unroll
        for (int pp = 0; pp < 9; ++pp) {
            dl[0][pp] = A[0][pp] * (dab[pp] - s0);
            dl[1][pp] = -__ldg(p.lamp) * A[1][pp] * (dab[pp] - s1);
        }
        float *wa = p.ws_abar + (tok * p.h + m) * 9;
        float *wl = p.ws_dlog + (tok * p.h + m) * 18;
#pragma unroll
        for (int pp = 0; pp < 9; ++pp) {
            wa[pp] = abar[pp];
            wl[pp] = dl[0][pp] * p.scale;
            wl[9 + pp] = dl[1][pp] * p.scale;
        }
        // dq_j = scale * sum_p dlogit_jp k_j[nbr_p]
        float dq[2 * HD];
#pragma unroll
        for (int c = 0; c < 2 * HD; ++c) dq[c] = 0.f;
#pragma unroll
        for (int pp = 0; pp < 9; ++pp) {
            const int rr = hr + pp / 3 - 1, cc = wc + pp % 3 - 1;
            if (rr >= 0 && rr < p.H && cc >= 0 && cc < p.W) {
                const long long nt = tok + (long long)(pp / 3 - 1) * p.W + (pp % 3 - 1);
                float kv[2 * HD];
                ldv<T, 2 * HD>(kb + nt * p.ldkv + (long long)m * 2 * HD, kv);
                const float a0 = dl[0][pp] * p.scale, a1 = dl[1][pp] * p.scale;
#pragma unroll
                for (int c = 0; c < HD; ++c) {
                    dq[c] = fmaf(a0, kv[c], dq[c]);
                    dq[HD + c] = fmaf(a1, kv[HD + c], dq[HD + c]);
                }
            }
        }
        stv<T, 2 * HD>(static_cast<T *>(p.dq) + tok * p.lddq + (long long)m * 2 * HD, dq);
    }
    // block reduction of d lambda and d subln_w: warp shuffle, then shared atomics (<= 4 warps), then global atomics
#pragma unroll
    for (int o2 = 16; o2 > 0; o2 >>= 1) dlam += __shfl_xor_sync(0xffffffffu, dlam, o2);
#pragma unroll
    for (int c = 0; c < 2 * HD; ++c) {
#pragma unroll
        for (int o2 = 16; o2 > 0; o2 >>= 1) dw[c] += __shfl_xor_sync(0xffffffffu, dw[c], o2);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&red[2 * HD], dlam);
#pragma unroll
        for (int c = 0; c < 2 * HD; ++c) atomicAdd(&red[c], dw[c]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * HD + 1; i += blockDim.x) {
        if (i < 2 * HD) atomicAdd(p.d_subln_w + i, red[i]);
        else atomicAdd(p.d_lambda, red[i]);
    }
}

template <typename T, int HD>
__global__ void __launch_bounds__(128) local_attn_bwd_p2_kernel(const LocalAttnParams p) {
    const long long total = (long long)p.Bn * p.H * p.W * p.h;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int m = (int)(idx % p.h);
    const long long tok = idx / p.h;
    const int wc = (int)(tok % p.W), hr = (int)((tok / p.W) % p.H);
    float dk[2 * HD], dv[2 * HD];
#pragma unroll
    for (int c = 0; c < 2 * HD; ++c) dk[c] = dv[c] = 0.f;
    const T *qb = static_cast<const T *>(p.q);
#pragma unroll
    for (int pp = 0; pp < 9; ++pp) {
        // token n = tok - off_p has tok as its pp-th neighbour
        const int rr = hr - (pp / 3 - 1), cc = wc - (pp % 3 - 1);
        if (rr >= 0 && rr < p.H && cc >= 0 && cc < p.W) {
            const long long n = tok - (long long)(pp / 3 - 1) * p.W - (pp % 3 - 1);
            const float ab = __ldg(p.ws_abar + (n * p.h + m) * 9 + pp);
            const float l0 = __ldg(p.ws_dlog + (n * p.h + m) * 18 + pp);
            const float l1 = __ldg(p.ws_dlog + (n * p.h + m) * 18 + 9 + pp);
            float dO[2 * HD], qn[2 * HD];
            ldv<float, 2 * HD>(p.ws_dO + (n * p.h + m) * 2 * HD, dO);
            ldv<T, 2 * HD>(qb + n * p.ldq + (long long)m * 2 * HD, qn);
#pragma unroll
            for (int c = 0; c < 2 * HD; ++c) dv[c] = fmaf(ab, dO[c], dv[c]);
#pragma unroll
            for (int c = 0; c < HD; ++c) {
                dk[c] = fmaf(l0, qn[c], dk[c]);
                dk[HD + c] = fmaf(l1, qn[HD + c], dk[HD + c]);
            }
        }
    }
    stv<T, 2 * HD>(static_cast<T *>(p.dk) + tok * p.lddkv + (long long)m * 2 * HD, dk);
    stv<T, 2 * HD>(static_cast<T *>(p.dv) + tok * p.lddkv + (long long)m * 2 * HD, dv);
}

template <typename T, int HD>
static cudaError_t local_launch(const LocalAttnParams &p, int which, cudaStream_t st) {
    const long long total = (long long)p.Bn * p.H * p.W * p.h;
    const int blocks = (int)((total + 127) / 128);
    if (which == 0) local_attn_fwd_kernel<T, HD><<<blocks, 128, 0, st>>>(p);
    else if (which == 1) local_attn_bwd_p1_kernel<T, HD><<<blocks, 128, 0, st>>>(p);
    else local_attn_bwd_p2_kernel<T, HD><<<blocks, 128, 0, st>>>(p);
    return cudaGetLastError();
}

template <typename T>
static cudaError_t local_hd(const LocalAttnParams &p, int hd, int which, cudaStream_t st) {
    switch (hd) {
        case 2: return local_launch<T, 2>(p, which, st);
        case 4: return local_launch<T, 4>(p, which, st);
        case 8: return local_launch<T, 8>(p, which, st);
        case 16: return local_launch<T, 16>(p, which, st);
        case 24: return local_launch<T, 24>(p, which, st);
        case 32: return local_launch<T, 32>(p, which, st);
        default: return cudaErrorInvalidValue;
    }
}

bool local_attn_hd_supported(int hd) { return hd == 2 || hd == 4 || hd == 8 || hd == 16 || hd == 24 || hd == 32; }

cudaError_t local_attn_dispatch(const LocalAttnParams &p, int hd, int dtype, int which, cudaStream_t st) {
    return dtype == 0 ? local_hd<float>(p, hd, which, st) : local_hd<__nv_bfloat16>(p, hd, which, st);
}

}  // namespace mlagg
