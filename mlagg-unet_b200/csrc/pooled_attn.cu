// pooled_attn.cu -- fused differential softmax attention over the P pooled tokens + sub-LN of the MLAgg block.
// Replaces, in ONE pass over q, the four flash_attn_func calls + 2x cat + lambda-combine + RMSNorm + scale at
// reference nnUNetTrainer_MLAgg_2D_dt_MS.py:734-760, reproducing the shipped DOUBLE softmax scaling (q * hd**-0.5
// at :688 and flash-attn's own default softmax_scale = hd**-0.5): logits = (q . k) * scale * scale.
//   q    (B, N, h, 2, hd)   raw projection, row stride ldq
//   kp   (B, P, h, 2, hd), vp (B, P, h, 2hd): the halves of kv(norm(pooled)), common row stride ldkv
//   out  (B, N, h, 2hd)     row stride ldo;  lse (B, N, h, 2) fp32 saved for the backward pass
// Since attn1 - lam * attn2 = sum_p (A0_p - lam A1_p) v_p, both softmax maps feed one 2hd-wide accumulator pair.
// Block = 128 query tokens of one (batch, head pair); the pair's K and V (P x 2hd each) sit in shared memory as
// fp32 and every lane reads them as warp-wide broadcasts; q, the two online-softmax states and the output stay in
// registers.  Backward: a token-parallel kernel (dq, dO, D_j = dO.o_j, d lambda, d subln_w) and a pooled-token-
// parallel kernel (dK, dV: thread = pooled token p, loops over a 512-token slab staged in shared memory, one
// atomic per element per slab).
#include <cuda_bf16.h>

#include "common.cuh"

namespace mlagg {

template <typename T>
__device__ __forceinline__ float pl_ld(const T *p);
template <>
__device__ __forceinline__ float pl_ld<float>(const float *p) { return __ldg(p); }
template <>
__device__ __forceinline__ float pl_ld<__nv_bfloat16>(const __nv_bfloat16 *p) { return __bfloat162float(*p); }
template <typename T>
__device__ __forceinline__ void pl_st(T *p, float v);
template <>
__device__ __forceinline__ void pl_st<float>(float *p, float v) { *p = v; }
template <>
__device__ __forceinline__ void pl_st<__nv_bfloat16>(__nv_bfloat16 *p, float v) { *p = __float2bfloat16_rn(v); }

struct PooledAttnParams {
    const void *q, *kp, *vp, *dout;
    void *out, *dq;
    float *lse;                  // (B,N,h,2)
    float *dkp, *dvp;            // (B,P,h,2hd) fp32 each, row stride ldd (accumulated)
    const float *subln_w;
    float *d_subln_w, *d_lambda;
    float *ws_dO, *ws_D;         // (B,N,h,2hd), (B,N,h,2)
    const float *lamp;
    long long ldq, ldkv, ldo, lddo, lddq, ldd;
    int Bn, N, P, h;
    float scale2, eps, post;     // scale2 = scale * scale
};

constexpr int kPTok = 128;       // query tokens per block (fwd / token-parallel bwd)

template <typename T, int HD>
__device__ __forceinline__ void stage_kv(const PooledAttnParams &p, int b, int m, float *ks, float *vs) {
    const T *kb = static_cast<const T *>(p.kp) + (long long)b * p.P * p.ldkv + (long long)m * 2 * HD;
    const T *vb = static_cast<const T *>(p.vp) + (long long)b * p.P * p.ldkv + (long long)m * 2 * HD;
    for (int i = threadIdx.x; i < p.P * 2 * HD; i += blockDim.x) {
        const int pp = i / (2 * HD), c = i % (2 * HD);
        ks[i] = pl_ld<T>(kb + (long long)pp * p.ldkv + c);
        vs[i] = pl_ld<T>(vb + (long long)pp * p.ldkv + c);
    }
}

template <typename T, int HD>
__global__ void __launch_bounds__(kPTok) pooled_attn_fwd_kernel(const PooledAttnParams p) {
    extern __shared__ __align__(16) float smem[];
    float *ks = smem, *vs = smem + p.P * 2 * HD;
    const int b = blockIdx.z, m = blockIdx.y;
    stage_kv<T, HD>(p, b, m, ks, vs);
    __syncthreads();
    const int n = blockIdx.x * kPTok + threadIdx.x;
    if (n >= p.N) return;
    const long long tok = (long long)b * p.N + n;
    float q[2 * HD];
    const T *qp = static_cast<const T *>(p.q) + tok * p.ldq + (long long)m * 2 * HD;
    const float qs = p.scale2 * kLog2e;  // work in the exp2 domain
#pragma unroll
    for (int c = 0; c < 2 * HD; ++c) q[c] = pl_ld<T>(qp + c) * qs;
    float mx[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
    float o0[2 * HD], o1[2 * HD];
#pragma unroll
    for (int c = 0; c < 2 * HD; ++c) o0[c] = o1[c] = 0.f;
    for (int pp = 0; pp < p.P; ++pp) {
        const float4 *k4 = reinterpret_cast<const float4 *>(ks + pp * 2 * HD);
        float d0 = 0.f, d1 = 0.f;
        if constexpr (HD % 4 == 0) {
#pragma unroll
            for (int c = 0; c < HD / 4; ++c) {
                const float4 a = k4[c], bb = k4[HD / 4 + c];
                d0 = fmaf(q[4 * c], a.x, fmaf(q[4 * c + 1], a.y, fmaf(q[4 * c + 2], a.z, fmaf(q[4 * c + 3], a.w, d0))));
                d1 = fmaf(q[HD + 4 * c], bb.x, fmaf(q[HD + 4 * c + 1], bb.y, fmaf(q[HD + 4 * c + 2], bb.z, fmaf(q[HD + 4 * c + 3], bb.w, d1))));
            }
        } else {
#pragma unroll
            for (int c = 0; c < HD; ++c) {
                d0 = fmaf(q[c], ks[pp * 2 * HD + c], d0);
                d1 = fmaf(q[HD + c], ks[pp * 2 * HD + HD + c], d1);
            }
        }
        float w0, w1;
        if (d0 > mx[0]) {
            const float f = ex2_approx(mx[0] - d0);
            l[0] *= f;
#pragma unroll
            for (int c = 0; c < 2 * HD; ++c) o0[c] *= f;
            mx[0] = d0;
        }
        if (d1 > mx[1]) {
            const float f = ex2_approx(mx[1] - d1);
            l[1] *= f;
#pragma unroll
            for (int c = 0; c < 2 * HD; ++c) o1[c] *= f;
            mx[1] = d1;
        }
        w0 = ex2_approx(d0 - mx[0]);
        w1 = ex2_approx(d1 - mx[1]);
        l[0] += w0;
        l[1] += w1;
        const float *vrow = vs + pp * 2 * HD;
#pragma unroll
        for (int c = 0; c < 2 * HD; ++c) {
            const float vv = vrow[c];
            o0[c] = fmaf(w0, vv, o0[c]);
            o1[c] = fmaf(w1, vv, o1[c]);
        }
    }
    const float lam = __ldg(p.lamp);
    const float i0 = 1.f / l[0], i1 = lam / l[1];
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < 2 * HD; ++c) {
        o0[c] = o0[c] * i0 - o1[c] * i1;
        ss = fmaf(o0[c], o0[c], ss);
    }
    const float r = 1.f / sqrtf(ss * (1.f / (2 * HD)) + p.eps);
    T *op = static_cast<T *>(p.out) + tok * p.ldo + (long long)m * 2 * HD;
#pragma unroll
    for (int c = 0; c < 2 * HD; ++c) pl_st<T>(op + c, o0[c] * r * __ldg(p.subln_w + c) * p.post);
    if (p.lse) {
        // natural-log-domain log-sum-exp of the scaled logits, kept in the exp2 domain: lse2 = mx + log2(l)
        p.lse[(tok * p.h + m) * 2 + 0] = mx[0] + lg2_approx(l[0]);
        p.lse[(tok * p.h + m) * 2 + 1] = mx[1] + lg2_approx(l[1]);
    }
}

template <typename T, int HD>
__global__ void __launch_bounds__(kPTok) pooled_attn_bwd_q_kernel(const PooledAttnParams p) {
    extern __shared__ __align__(16) float smem[];
    __shared__ float red[2 * HD + 1];
    float *ks = smem, *vs = smem + p.P * 2 * HD;
    const int b = blockIdx.z, m = blockIdx.y;
    stage_kv<T, HD>(p, b, m, ks, vs);
    for (int i = threadIdx.x; i < 2 * HD + 1; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
    const int n = blockIdx.x * kPTok + threadIdx.x;
    float dlam = 0.f, dw[2 * HD];
#pragma unroll
    for (int c = 0; c < 2 * HD; ++c) dw[c] = 0.f;
    if (n < p.N) {
        const long long tok = (long long)b * p.N + n;
        float q[2 * HD];
        const T *qp = static_cast<const T *>(p.q) + tok * p.ldq + (long long)m * 2 * HD;
        const float qs = p.scale2 * kLog2e;
#pragma unroll
        for (int c = 0; c < 2 * HD; ++c) q[c] = pl_ld<T>(qp + c) * qs;
        const float lse0 = p.lse[(tok * p.h + m) * 2 + 0], lse1 = p.lse[(tok * p.h + m) * 2 + 1];
        const float lam = __ldg(p.lamp);
        float o0[2 * HD], o1[2 * HD];
#pragma unroll
        for (int c = 0; c < 2 * HD; ++c) o0[c] = o1[c] = 0.f;
        for (int pp = 0; pp < p.P; ++pp) {
            float d0 = 0.f, d1 = 0.f;
#pragma unroll
            for (int c = 0; c < HD; ++c) {
                d0 = fmaf(q[c], ks[pp * 2 * HD + c], d0);
                d1 = fmaf(q[HD + c], ks[pp * 2 * HD + HD + c], d1);
            }
            const float w0 = ex2_approx(d0 - lse0), w1 = ex2_approx(d1 - lse1);
#pragma unroll
            for (int c = 0; c < 2 * HD; ++c) {
                const float vv = vs[pp * 2 * HD + c];
                o0[c] = fmaf(w0, vv, o0[c]);
                o1[c] = fmaf(w1, vv, o1[c]);
            }
        }
        // o = o0 - lam o1; RMSNorm backward
        float g[2 * HD], ss = 0.f;
        const T *gp = static_cast<const T *>(p.dout) + tok * p.lddo + (long long)m * 2 * HD;
#pragma unroll
        for (int c = 0; c < 2 * HD; ++c) {
            const float oc = o0[c] - lam * o1[c];
            ss = fmaf(oc, oc, ss);
            g[c] = pl_ld<T>(gp + c);
        }
        const float r = 1.f / sqrtf(ss * (1.f / (2 * HD)) + p.eps);
        float dot = 0.f;
#pragma unroll
        for (int c = 0; c < 2 * HD; ++c) {
            const float oc = o0[c] - lam * o1[c];
            dw[c] = g[c] * p.post * oc * r;
            g[c] *= p.post * __ldg(p.subln_w + c);
            dot = fmaf(g[c], oc, dot);
        }
        const float k3 = r * r * r * dot * (1.f / (2 * HD));
        float D0 = 0.f, D1 = 0.f;
        float *wdO = p.ws_dO + (tok * p.h + m) * 2 * HD;
#pragma unroll
        for (int c = 0; c < 2 * HD; ++c) {
            const float oc = o0[c] - lam * o1[c];
            g[c] = r * g[c] - oc * k3;  // g now holds dO
            D0 = fmaf(g[c], o0[c], D0);
            D1 = fmaf(g[c], o1[c], D1);
            wdO[c] = g[c];
        }
        p.ws_D[(tok * p.h + m) * 2 + 0] = D0;
        p.ws_D[(tok * p.h + m) * 2 + 1] = D1;
        dlam = -D1;
        // second pass: dq_j = scale2 * sum_p dlogit_jp k_jp   (o0 / o1 registers are reused as the accumulator)
#pragma unroll
        for (int c = 0; c < 2 * HD; ++c) o0[c] = 0.f;
        for (int pp = 0; pp < p.P; ++pp) {
            float d0 = 0.f, d1 = 0.f, dab = 0.f;
#pragma unroll
            for (int c = 0; c < HD; ++c) {
                d0 = fmaf(q[c], ks[pp * 2 * HD + c], d0);
                d1 = fmaf(q[HD + c], ks[pp * 2 * HD + HD + c], d1);
            }
#pragma unroll
            for (int c = 0; c < 2 * HD; ++c) dab = fmaf(g[c], vs[pp * 2 * HD + c], dab);
            const float dl0 = ex2_approx(d0 - lse0) * (dab - D0);
            const float dl1 = -lam * ex2_approx(d1 - lse1) * (dab - D1);
#pragma unroll
            for (int c = 0; c < HD; ++c) {
                o0[c] = fmaf(dl0, ks[pp * 2 * HD + c], o0[c]);
                o0[HD + c] = fmaf(dl1, ks[pp * 2 * HD + HD + c], o0[HD + c]);
            }
        }
        T *dqp = static_cast<T *>(p.dq) + tok * p.lddq + (long long)m * 2 * HD;
#pragma unroll
        for (int c = 0; c < 2 * HD; ++c) pl_st<T>(dqp + c, o0[c] * p.scale2);
    }
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

// dK / dV: thread = pooled token; block = (slab of kSlab query tokens, head pair, batch)
constexpr int kSlab = 512, kSub = 32;

template <typename T, int HD>
__global__ void __launch_bounds__(256) pooled_attn_bwd_kv_kernel(const PooledAttnParams p) {
    __shared__ __align__(16) float sq[kSub][2 * HD];
    __shared__ __align__(16) float sdo[kSub][2 * HD];
    __shared__ float sl[kSub][4];  // lse0, lse1, D0, D1
    const int b = blockIdx.z, m = blockIdx.y;
    const int pp = threadIdx.x;
    const bool active = pp < p.P;
    float k[2 * HD], v[2 * HD], dk[2 * HD], dv[2 * HD];
    if (active) {
        const T *kb = static_cast<const T *>(p.kp) + ((long long)b * p.P + pp) * p.ldkv + (long long)m * 2 * HD;
        const T *vb = static_cast<const T *>(p.vp) + ((long long)b * p.P + pp) * p.ldkv + (long long)m * 2 * HD;
#pragma unroll
        for (int c = 0; c < 2 * HD; ++c) {
            k[c] = pl_ld<T>(kb + c);
            v[c] = pl_ld<T>(vb + c);
        }
    }
#pragma unroll
    for (int c = 0; c < 2 * HD; ++c) dk[c] = dv[c] = 0.f;
    const float lam = __ldg(p.lamp);
    const float qs = p.scale2 * kLog2e;
    const int n0 = blockIdx.x * kSlab, n1 = min(p.N, n0 + kSlab);
    for (int base = n0; base < n1; base += kSub) {
        const int cnt = min(kSub, n1 - base);
        __syncthreads();
        for (int i = threadIdx.x; i < cnt * 2 * HD; i += blockDim.x) {
            const int t = i / (2 * HD), c = i % (2 * HD);
            const long long tok = (long long)b * p.N + base + t;
            sq[t][c] = pl_ld<T>(static_cast<const T *>(p.q) + tok * p.ldq + (long long)m * 2 * HD + c);
            sdo[t][c] = p.ws_dO[(tok * p.h + m) * 2 * HD + c];
        }
        for (int i = threadIdx.x; i < cnt * 4; i += blockDim.x) {
            const int t = i / 4, w = i % 4;
            const long long tok = (long long)b * p.N + base + t;
            sl[t][w] = w < 2 ? p.lse[(tok * p.h + m) * 2 + w] : p.ws_D[(tok * p.h + m) * 2 + (w - 2)];
        }
        __syncthreads();
        if (active) {
            for (int t = 0; t < cnt; ++t) {
                float d0 = 0.f, d1 = 0.f, dab = 0.f;
#pragma unroll
                for (int c = 0; c < HD; ++c) {
                    d0 = fmaf(sq[t][c], k[c], d0);
                    d1 = fmaf(sq[t][HD + c], k[HD + c], d1);
                }
#pragma unroll
                for (int c = 0; c < 2 * HD; ++c) dab = fmaf(sdo[t][c], v[c], dab);
                const float a0 = ex2_approx(d0 * qs - sl[t][0]), a1 = ex2_approx(d1 * qs - sl[t][1]);
                const float ab = a0 - lam * a1;
                const float dl0 = a0 * (dab - sl[t][2]);
                const float dl1 = -lam * a1 * (dab - sl[t][3]);
#pragma unroll
                for (int c = 0; c < 2 * HD; ++c) dv[c] = fmaf(ab, sdo[t][c], dv[c]);
#pragma unroll
                for (int c = 0; c < HD; ++c) {
                    dk[c] = fmaf(dl0, sq[t][c], dk[c]);
                    dk[HD + c] = fmaf(dl1, sq[t][HD + c], dk[HD + c]);
                }
            }
        }
    }
    if (active) {
        float *dkb = p.dkp + ((long long)b * p.P + pp) * p.ldd + (long long)m * 2 * HD;
        float *dvb = p.dvp + ((long long)b * p.P + pp) * p.ldd + (long long)m * 2 * HD;
#pragma unroll
        for (int c = 0; c < 2 * HD; ++c) {
            atomicAdd(dkb + c, dk[c] * p.scale2);
            atomicAdd(dvb + c, dv[c]);
        }
    }
}

template <typename T, int HD>
static cudaError_t pooled_launch(const PooledAttnParams &p, int which, cudaStream_t st) {
    const size_t smem = (size_t)p.P * 4 * HD * sizeof(float);
    cudaError_t e;
    if (which == 0) {
        auto k = pooled_attn_fwd_kernel<T, HD>;
        if ((e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        k<<<dim3((p.N + kPTok - 1) / kPTok, p.h, p.Bn), kPTok, smem, st>>>(p);
    } else if (which == 1) {
        auto k = pooled_attn_bwd_q_kernel<T, HD>;
        if ((e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        k<<<dim3((p.N + kPTok - 1) / kPTok, p.h, p.Bn), kPTok, smem, st>>>(p);
    } else {
        const int threads = ((p.P + 31) / 32) * 32;
        pooled_attn_bwd_kv_kernel<T, HD><<<dim3((p.N + kSlab - 1) / kSlab, p.h, p.Bn), threads, 0, st>>>(p);
    }
    return cudaGetLastError();
}

template <typename T>
static cudaError_t pooled_hd(const PooledAttnParams &p, int hd, int which, cudaStream_t st) {
    switch (hd) {
        case 2: return pooled_launch<T, 2>(p, which, st);
        case 4: return pooled_launch<T, 4>(p, which, st);
        case 8: return pooled_launch<T, 8>(p, which, st);
        case 16: return pooled_launch<T, 16>(p, which, st);
        case 24: return pooled_launch<T, 24>(p, which, st);
        case 32: return pooled_launch<T, 32>(p, which, st);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t pooled_attn_dispatch(const PooledAttnParams &p, int hd, int dtype, int which, cudaStream_t st) {
    return dtype == 0 ? pooled_hd<float>(p, hd, which, st) : pooled_hd<__nv_bfloat16>(p, hd, which, st);
}

}  // namespace mlagg
