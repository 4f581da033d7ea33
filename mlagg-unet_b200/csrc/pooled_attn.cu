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

#include <cstdlib>
#include <type_traits>

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
    float *lse;                  // saved: (B,N,h,2) log-sum-exps, then (16 B aligned) (B,N,h,2,2hd) normalised O0 | O1
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

// All inner products are PACKED f32x2 (FFMA2: two fp32 FMAs per issue slot; exact fp32 arithmetic):
//   logits (d0, d1) += (q0[c], q1[c]) * (k0[c], k1[c])        K staged interleaved:  kI[p][c] = (K[p][c], K[p][hd + c])
//   (o0[c], o1[c])  += (w0, w1) * (v[c], v[c])
// and every shared-memory read is a warp-wide broadcast LDS.128.  ncu (profiles/pooled_fwd_r01_ncu_raw.txt): these
// kernels are bound by the shared-memory return path -- a broadcast LDS.128 still delivers 512 B of lane data, i.e. 4
// cycles of the 128 B/clk pipe, for 8..16 FMAs per lane -- at ~19 % of the FP32 peak; the way out is operand reuse in
// registers across tokens, i.e. tensor-core fragments.
__device__ __forceinline__ float2 p2(float a, float b) { return make_float2(a, b); }

__device__ __forceinline__ size_t pooled_osave_offset(const PooledAttnParams &p) {
    return (((size_t)p.Bn * p.N * p.h * 2) + 3) & ~(size_t)3;   // floats; keeps the O rows 16-byte aligned
}

template <typename T, int HD>
__device__ __forceinline__ void stage_k_interleaved(const PooledAttnParams &p, int b, int m, float2 *kI) {
    const T *kb = static_cast<const T *>(p.kp) + (long long)b * p.P * p.ldkv + (long long)m * 2 * HD;
    for (int i = threadIdx.x; i < p.P * HD; i += blockDim.x) {
        const int pp = i / HD, c = i % HD;
        kI[i] = p2(pl_ld<T>(kb + (long long)pp * p.ldkv + c), pl_ld<T>(kb + (long long)pp * p.ldkv + HD + c));
    }
}
template <typename T, int HD, bool DUP>
__device__ __forceinline__ void stage_v(const PooledAttnParams &p, int b, int m, float *vs) {
    const T *vb = static_cast<const T *>(p.vp) + (long long)b * p.P * p.ldkv + (long long)m * 2 * HD;
    for (int i = threadIdx.x; i < p.P * 2 * HD; i += blockDim.x) {
        const int pp = i / (2 * HD), c = i % (2 * HD);
        const float v = pl_ld<T>(vb + (long long)pp * p.ldkv + c);
        if (DUP) reinterpret_cast<float2 *>(vs)[i] = p2(v, v);
        else vs[i] = v;
    }
}
// (d0, d1) = sum_c q2[c] * kI[pp][c]
template <int HD>
__device__ __forceinline__ float2 logits2(const float2 (&q2)[HD], const float2 *krow) {
    float2 acc[4] = {p2(0.f, 0.f), p2(0.f, 0.f), p2(0.f, 0.f), p2(0.f, 0.f)};   // 4 independent chains
#pragma unroll
    for (int c = 0; c < HD; c += 2) {
        const float4 kk = *reinterpret_cast<const float4 *>(krow + c);
        acc[c & 3] = __ffma2_rn(q2[c], p2(kk.x, kk.y), acc[c & 3]);
        acc[(c & 3) + 1] = __ffma2_rn(q2[c + 1], p2(kk.z, kk.w), acc[(c & 3) + 1]);
    }
    return __fadd2_rn(__fadd2_rn(acc[0], acc[1]), __fadd2_rn(acc[2], acc[3]));
}

template <typename T, int HD>
__global__ void __launch_bounds__(kPTok) pooled_attn_fwd_kernel(const PooledAttnParams p) {
    extern __shared__ __align__(16) float smem[];
    float2 *kI = reinterpret_cast<float2 *>(smem);                 // [P][HD]
    float *vP = reinterpret_cast<float *>(kI + p.P * HD);          // [P][2HD]
    const int b = blockIdx.z, m = blockIdx.y;
    stage_k_interleaved<T, HD>(p, b, m, kI);
    stage_v<T, HD, false>(p, b, m, vP);
    __syncthreads();
    const int n = blockIdx.x * kPTok + threadIdx.x;
    if (n >= p.N) return;
    const long long tok = (long long)b * p.N + n;
    float2 q2[HD];
    const T *qp = static_cast<const T *>(p.q) + tok * p.ldq + (long long)m * 2 * HD;
    const float qs = p.scale2 * kLog2e;  // work in the exp2 domain
#pragma unroll
    for (int c = 0; c < HD; ++c) q2[c] = p2(pl_ld<T>(qp + c) * qs, pl_ld<T>(qp + HD + c) * qs);
    float2 mx = p2(-INFINITY, -INFINITY), l = p2(0.f, 0.f);
    float2 o[2 * HD];                                              // (o0[c], o1[c])
#pragma unroll
    for (int c = 0; c < 2 * HD; ++c) o[c] = p2(0.f, 0.f);
    float2 dnext = logits2<HD>(q2, kI);
    for (int pp = 0; pp < p.P; ++pp) {
        const float2 d = dnext;
        if (pp + 1 < p.P) dnext = logits2<HD>(q2, kI + (pp + 1) * HD);   // independent of the softmax / AV work below
        if (d.x > mx.x || d.y > mx.y) {
            const float2 nm = p2(fmaxf(mx.x, d.x), fmaxf(mx.y, d.y));
            const float2 f = p2(ex2_approx(mx.x - nm.x), ex2_approx(mx.y - nm.y));
            l = __fmul2_rn(l, f);
#pragma unroll
            for (int c = 0; c < 2 * HD; ++c) o[c] = __fmul2_rn(o[c], f);
            mx = nm;
        }
        const float2 w = p2(ex2_approx(d.x - mx.x), ex2_approx(d.y - mx.y));
        l = __fadd2_rn(l, w);
        // the shared-memory return path (128 B of lane data per cycle and SM), not the FMA pipe, bounds this loop: a
        // duplicated-pair layout of V would save the four register moves below but double the LDS traffic
        const float *vrow = vP + pp * 2 * HD;
#pragma unroll
        for (int c = 0; c < 2 * HD; c += 4) {
            const float4 vv = *reinterpret_cast<const float4 *>(vrow + c);
            o[c] = __ffma2_rn(w, p2(vv.x, vv.x), o[c]);
            o[c + 1] = __ffma2_rn(w, p2(vv.y, vv.y), o[c + 1]);
            o[c + 2] = __ffma2_rn(w, p2(vv.z, vv.z), o[c + 2]);
            o[c + 3] = __ffma2_rn(w, p2(vv.w, vv.w), o[c + 3]);
        }
    }
    const float lam = __ldg(p.lamp);
    const float2 inv = p2(1.f / l.x, 1.f / l.y);
    float ss = 0.f;
    float oc[2 * HD];
#pragma unroll
    for (int c = 0; c < 2 * HD; ++c) {
        o[c] = __fmul2_rn(o[c], inv);                              // normalised O0, O1
        oc[c] = o[c].x - lam * o[c].y;
        ss = fmaf(oc[c], oc[c], ss);
    }
    const float r = 1.f / sqrtf(ss * (1.f / (2 * HD)) + p.eps);
    T *op = static_cast<T *>(p.out) + tok * p.ldo + (long long)m * 2 * HD;
#pragma unroll
    for (int c = 0; c < 2 * HD; ++c) pl_st<T>(op + c, oc[c] * r * __ldg(p.subln_w + c) * p.post);
    if (p.lse) {
        // log-sum-exp of the scaled logits in the exp2 domain (lse2 = mx + log2 l), then O0 | O1 for the backward pass
        p.lse[(tok * p.h + m) * 2 + 0] = mx.x + lg2_approx(l.x);
        p.lse[(tok * p.h + m) * 2 + 1] = mx.y + lg2_approx(l.y);
        float4 *os = reinterpret_cast<float4 *>(p.lse + pooled_osave_offset(p) + (tok * p.h + m) * 4 * HD);
#pragma unroll
        for (int c = 0; c < 2 * HD; c += 4) {
            os[c / 4] = make_float4(o[c].x, o[c + 1].x, o[c + 2].x, o[c + 3].x);
            os[(2 * HD + c) / 4] = make_float4(o[c].y, o[c + 1].y, o[c + 2].y, o[c + 3].y);
        }
    }
}

template <typename T, int HD>
__global__ void __launch_bounds__(kPTok) pooled_attn_bwd_q_kernel(const PooledAttnParams p) {
    extern __shared__ __align__(16) float smem[];
    __shared__ float red[2 * HD + 1];
    float2 *kI = reinterpret_cast<float2 *>(smem);                 // [P][HD]
    float *vP = reinterpret_cast<float *>(kI + p.P * HD);          // [P][2HD]
    const int b = blockIdx.z, m = blockIdx.y;
    stage_k_interleaved<T, HD>(p, b, m, kI);
    stage_v<T, HD, false>(p, b, m, vP);
    for (int i = threadIdx.x; i < 2 * HD + 1; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
    const int n = blockIdx.x * kPTok + threadIdx.x;
    float dlam = 0.f, dw[2 * HD];
#pragma unroll
    for (int c = 0; c < 2 * HD; ++c) dw[c] = 0.f;
    if (n < p.N) {
        const long long tok = (long long)b * p.N + n;
        const float lam = __ldg(p.lamp);
        // ---- RMSNorm backward from the saved O0, O1:  o = O0 - lam O1
        float g[2 * HD];
        float D0 = 0.f, D1 = 0.f;
        {
            float o0[2 * HD], o1[2 * HD];
            const float4 *os = reinterpret_cast<const float4 *>(p.lse + pooled_osave_offset(p) + (tok * p.h + m) * 4 * HD);
#pragma unroll
            for (int c = 0; c < 2 * HD; c += 4) {
                const float4 a = os[c / 4], bb = os[(2 * HD + c) / 4];
                o0[c] = a.x, o0[c + 1] = a.y, o0[c + 2] = a.z, o0[c + 3] = a.w;
                o1[c] = bb.x, o1[c + 1] = bb.y, o1[c + 2] = bb.z, o1[c + 3] = bb.w;
            }
            float ss = 0.f;
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
            float *wdO = p.ws_dO + (tok * p.h + m) * 2 * HD;
#pragma unroll
            for (int c = 0; c < 2 * HD; ++c) {
                const float oc = o0[c] - lam * o1[c];
                g[c] = r * g[c] - oc * k3;  // g now holds dO
                D0 = fmaf(g[c], o0[c], D0);
                D1 = fmaf(g[c], o1[c], D1);
            }
#pragma unroll
            for (int c = 0; c < 2 * HD; c += 4)
                *reinterpret_cast<float4 *>(wdO + c) = make_float4(g[c], g[c + 1], g[c + 2], g[c + 3]);
            dlam = -D1;
        }
        // ---- dq_j = scale2 * sum_p dlogit_jp k_jp
        float2 q2[HD], dq2[HD];
        const T *qp = static_cast<const T *>(p.q) + tok * p.ldq + (long long)m * 2 * HD;
        const float qs = p.scale2 * kLog2e;
#pragma unroll
        for (int c = 0; c < HD; ++c) {
            q2[c] = p2(pl_ld<T>(qp + c) * qs, pl_ld<T>(qp + HD + c) * qs);
            dq2[c] = p2(0.f, 0.f);
        }
        const float lse0 = p.lse[(tok * p.h + m) * 2 + 0], lse1 = p.lse[(tok * p.h + m) * 2 + 1];
        auto dot_v = [&](int pp) {   // dO . v_p
            float2 da = p2(0.f, 0.f), db = p2(0.f, 0.f);
            const float *vrow = vP + pp * 2 * HD;
#pragma unroll
            for (int c = 0; c < 2 * HD; c += 4) {
                const float4 vv = *reinterpret_cast<const float4 *>(vrow + c);
                da = __ffma2_rn(p2(g[c], g[c + 1]), p2(vv.x, vv.y), da);
                db = __ffma2_rn(p2(g[c + 2], g[c + 3]), p2(vv.z, vv.w), db);
            }
            return (da.x + da.y) + (db.x + db.y);
        };
        // D_j = sum_p A_j[p] dab[p] from the same probabilities and dab values the loop below uses (see the tensor-core
        // kernel): dO . O_j is the same number mathematically, but where the softmax is nearly uniform dab - D_j cancels
        // to ~1e-3 of dab and the two evaluation orders differ by more than that leaves
        {
            float2 Dc = p2(0.f, 0.f);
            for (int pp = 0; pp < p.P; ++pp) {
                const float2 d = logits2<HD>(q2, kI + pp * HD);
                const float dab = dot_v(pp);
                Dc = __ffma2_rn(p2(ex2_approx(d.x - lse0), ex2_approx(d.y - lse1)), p2(dab, dab), Dc);
            }
            D0 = Dc.x, D1 = Dc.y;
            p.ws_D[(tok * p.h + m) * 2 + 0] = D0;
            p.ws_D[(tok * p.h + m) * 2 + 1] = D1;
        }
        float2 dnext = logits2<HD>(q2, kI);
        float dabnext = dot_v(0);
        for (int pp = 0; pp < p.P; ++pp) {
            const float2 d = dnext;
            const float dab = dabnext;
            if (pp + 1 < p.P) {          // next pooled token's reductions overlap this one's dq update
                dnext = logits2<HD>(q2, kI + (pp + 1) * HD);
                dabnext = dot_v(pp + 1);
            }
            const float2 dl = p2(ex2_approx(d.x - lse0) * (dab - D0), -lam * ex2_approx(d.y - lse1) * (dab - D1));
            const float2 *krow = kI + pp * HD;
#pragma unroll
            for (int c = 0; c < HD; c += 2) {
                const float4 kk = *reinterpret_cast<const float4 *>(krow + c);
                dq2[c] = __ffma2_rn(dl, p2(kk.x, kk.y), dq2[c]);
                dq2[c + 1] = __ffma2_rn(dl, p2(kk.z, kk.w), dq2[c + 1]);
            }
        }
        T *dqp = static_cast<T *>(p.dq) + tok * p.lddq + (long long)m * 2 * HD;
#pragma unroll
        for (int c = 0; c < HD; ++c) {
            pl_st<T>(dqp + c, dq2[c].x * p.scale2);
            pl_st<T>(dqp + HD + c, dq2[c].y * p.scale2);
        }
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
    __shared__ __align__(16) float2 sq[kSub][HD];       // (q0[c], q1[c]), already in the exp2 domain
    __shared__ __align__(16) float sdo[kSub][2 * HD];
    __shared__ float sl[kSub][4];                       // lse0, lse1, D0, D1
    const int b = blockIdx.z, m = blockIdx.y;
    const int pp = threadIdx.x;
    const bool active = pp < p.P;
    float2 k2[HD], v2[HD], dk2[HD], dv2[HD];            // k2 interleaved (k0[c], k1[c]); v2, dv2 = consecutive pairs
    if (active) {
        const T *kb = static_cast<const T *>(p.kp) + ((long long)b * p.P + pp) * p.ldkv + (long long)m * 2 * HD;
        const T *vb = static_cast<const T *>(p.vp) + ((long long)b * p.P + pp) * p.ldkv + (long long)m * 2 * HD;
#pragma unroll
        for (int c = 0; c < HD; ++c) {
            k2[c] = p2(pl_ld<T>(kb + c), pl_ld<T>(kb + HD + c));
            v2[c] = p2(pl_ld<T>(vb + 2 * c), pl_ld<T>(vb + 2 * c + 1));
        }
    } else {
#pragma unroll
        for (int c = 0; c < HD; ++c) k2[c] = v2[c] = p2(0.f, 0.f);
    }
#pragma unroll
    for (int c = 0; c < HD; ++c) dk2[c] = dv2[c] = p2(0.f, 0.f);
    const float lam = __ldg(p.lamp);
    const float qs = p.scale2 * kLog2e;
    const int n0 = blockIdx.x * kSlab, n1 = min(p.N, n0 + kSlab);
    for (int base = n0; base < n1; base += kSub) {
        const int cnt = min(kSub, n1 - base);
        __syncthreads();
        for (int i = threadIdx.x; i < cnt * HD; i += blockDim.x) {
            const int t = i / HD, c = i % HD;
            const long long tok = (long long)b * p.N + base + t;
            const T *qp = static_cast<const T *>(p.q) + tok * p.ldq + (long long)m * 2 * HD;
            sq[t][c] = p2(pl_ld<T>(qp + c) * qs, pl_ld<T>(qp + HD + c) * qs);
        }
        for (int i = threadIdx.x; i < cnt * 2 * HD; i += blockDim.x) {
            const int t = i / (2 * HD), c = i % (2 * HD);
            const long long tok = (long long)b * p.N + base + t;
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
                const float2 d = logits2<HD>(k2, &sq[t][0]);
                float2 da = p2(0.f, 0.f);
#pragma unroll
                for (int c = 0; c < HD; c += 2) {
                    const float4 gg = *reinterpret_cast<const float4 *>(&sdo[t][2 * c]);
                    da = __ffma2_rn(p2(gg.x, gg.y), v2[c], da);
                    da = __ffma2_rn(p2(gg.z, gg.w), v2[c + 1], da);
                }
                const float dab = da.x + da.y;
                const float a0 = ex2_approx(d.x - sl[t][0]), a1 = ex2_approx(d.y - sl[t][1]);
                const float ab = a0 - lam * a1;
                const float2 ab2 = p2(ab, ab);
                const float2 dl = p2(a0 * (dab - sl[t][2]), -lam * a1 * (dab - sl[t][3]));
#pragma unroll
                for (int c = 0; c < HD; c += 2) {
                    const float4 gg = *reinterpret_cast<const float4 *>(&sdo[t][2 * c]);
                    dv2[c] = __ffma2_rn(ab2, p2(gg.x, gg.y), dv2[c]);
                    dv2[c + 1] = __ffma2_rn(ab2, p2(gg.z, gg.w), dv2[c + 1]);
                    const float4 qq = *reinterpret_cast<const float4 *>(&sq[t][c]);
                    dk2[c] = __ffma2_rn(dl, p2(qq.x, qq.y), dk2[c]);
                    dk2[c + 1] = __ffma2_rn(dl, p2(qq.z, qq.w), dk2[c + 1]);
                }
            }
        }
    }
    if (active) {
        float *dkb = p.dkp + ((long long)b * p.P + pp) * p.ldd + (long long)m * 2 * HD;
        float *dvb = p.dvp + ((long long)b * p.P + pp) * p.ldd + (long long)m * 2 * HD;
        // sq carried scale2 * log2e: d logit / d k = q * scale2, and the exp2-domain factor cancels against ln 2
#pragma unroll
        for (int c = 0; c < HD; ++c) {
            atomicAdd(dkb + c, dk2[c].x * (1.f / kLog2e));
            atomicAdd(dkb + HD + c, dk2[c].y * (1.f / kLog2e));
            atomicAdd(dvb + 2 * c, dv2[c].x);
            atomicAdd(dvb + 2 * c + 1, dv2[c].y);
        }
    }
}

// Same computation with TWO threads per pooled token, each owning half of the channels (hd % 4 == 0): 96 instead of 192
// state registers per thread, so three blocks of 2P threads fit an SM instead of two of P -- the kernel is bound by the
// latency of its shared-memory operands, and 2.5x the resident warps hide more of it.  The two halves of a logit / of
// dO . v meet through one shuffle.
template <typename T, int HD>
__global__ void __launch_bounds__(256) pooled_attn_bwd_kv2_kernel(const PooledAttnParams p) {
    constexpr int HH = HD / 2;                          // float2 slots per thread and array
    __shared__ __align__(16) float2 sq[kSub][HD];
    __shared__ __align__(16) float sdo[kSub][2 * HD];
    __shared__ float sl[kSub][4];
    const int b = blockIdx.z, m = blockIdx.y;
    const int pp = threadIdx.x >> 1, half = threadIdx.x & 1;
    const bool active = pp < p.P;
    float2 k2[HH], v2[HH], dk2[HH], dv2[HH];
    if (active) {
        const T *kb = static_cast<const T *>(p.kp) + ((long long)b * p.P + pp) * p.ldkv + (long long)m * 2 * HD;
        const T *vb = static_cast<const T *>(p.vp) + ((long long)b * p.P + pp) * p.ldkv + (long long)m * 2 * HD;
#pragma unroll
        for (int c = 0; c < HH; ++c) {
            const int cc = half * HH + c;
            k2[c] = p2(pl_ld<T>(kb + cc), pl_ld<T>(kb + HD + cc));
            v2[c] = p2(pl_ld<T>(vb + 2 * cc), pl_ld<T>(vb + 2 * cc + 1));
        }
    } else {
#pragma unroll
        for (int c = 0; c < HH; ++c) k2[c] = v2[c] = p2(0.f, 0.f);
    }
#pragma unroll
    for (int c = 0; c < HH; ++c) dk2[c] = dv2[c] = p2(0.f, 0.f);
    const float lam = __ldg(p.lamp);
    const float qs = p.scale2 * kLog2e;
    const int n0 = blockIdx.x * kSlab, n1 = min(p.N, n0 + kSlab);
    for (int base = n0; base < n1; base += kSub) {
        const int cnt = min(kSub, n1 - base);
        __syncthreads();
        for (int i = threadIdx.x; i < cnt * HD; i += blockDim.x) {
            const int t = i / HD, c = i % HD;
            const long long tok = (long long)b * p.N + base + t;
            const T *qp = static_cast<const T *>(p.q) + tok * p.ldq + (long long)m * 2 * HD;
            sq[t][c] = p2(pl_ld<T>(qp + c) * qs, pl_ld<T>(qp + HD + c) * qs);
        }
        for (int i = threadIdx.x; i < cnt * 2 * HD; i += blockDim.x) {
            const int t = i / (2 * HD), c = i % (2 * HD);
            const long long tok = (long long)b * p.N + base + t;
            sdo[t][c] = p.ws_dO[(tok * p.h + m) * 2 * HD + c];
        }
        for (int i = threadIdx.x; i < cnt * 4; i += blockDim.x) {
            const int t = i / 4, w = i % 4;
            const long long tok = (long long)b * p.N + base + t;
            sl[t][w] = w < 2 ? p.lse[(tok * p.h + m) * 2 + w] : p.ws_D[(tok * p.h + m) * 2 + (w - 2)];
        }
        __syncthreads();
        for (int t = 0; t < cnt; ++t) {
            const float2 *qrow = &sq[t][half * HH];
            const float *grow = &sdo[t][2 * half * HH];
            float2 d = p2(0.f, 0.f), e = p2(0.f, 0.f), da = p2(0.f, 0.f), db = p2(0.f, 0.f);
#pragma unroll
            for (int c = 0; c < HH; c += 2) {
                const float4 qq = *reinterpret_cast<const float4 *>(qrow + c);
                d = __ffma2_rn(k2[c], p2(qq.x, qq.y), d);
                e = __ffma2_rn(k2[c + 1], p2(qq.z, qq.w), e);
                const float4 gg = *reinterpret_cast<const float4 *>(grow + 2 * c);
                da = __ffma2_rn(p2(gg.x, gg.y), v2[c], da);
                db = __ffma2_rn(p2(gg.z, gg.w), v2[c + 1], db);
            }
            d = __fadd2_rn(d, e);
            float dab = (da.x + da.y) + (db.x + db.y);
            d.x += __shfl_xor_sync(0xffffffffu, d.x, 1);
            d.y += __shfl_xor_sync(0xffffffffu, d.y, 1);
            dab += __shfl_xor_sync(0xffffffffu, dab, 1);
            const float a0 = ex2_approx(d.x - sl[t][0]), a1 = ex2_approx(d.y - sl[t][1]);
            const float ab = a0 - lam * a1;
            const float2 ab2 = p2(ab, ab);
            const float2 dl = p2(a0 * (dab - sl[t][2]), -lam * a1 * (dab - sl[t][3]));
#pragma unroll
            for (int c = 0; c < HH; c += 2) {
                const float4 gg = *reinterpret_cast<const float4 *>(grow + 2 * c);
                dv2[c] = __ffma2_rn(ab2, p2(gg.x, gg.y), dv2[c]);
                dv2[c + 1] = __ffma2_rn(ab2, p2(gg.z, gg.w), dv2[c + 1]);
                const float4 qq = *reinterpret_cast<const float4 *>(qrow + c);
                dk2[c] = __ffma2_rn(dl, p2(qq.x, qq.y), dk2[c]);
                dk2[c + 1] = __ffma2_rn(dl, p2(qq.z, qq.w), dk2[c + 1]);
            }
        }
    }
    if (active) {
        float *dkb = p.dkp + ((long long)b * p.P + pp) * p.ldd + (long long)m * 2 * HD;
        float *dvb = p.dvp + ((long long)b * p.P + pp) * p.ldd + (long long)m * 2 * HD;
#pragma unroll
        for (int c = 0; c < HH; ++c) {
            const int cc = half * HH + c;
            atomicAdd(dkb + cc, dk2[c].x * (1.f / kLog2e));
            atomicAdd(dkb + HD + cc, dk2[c].y * (1.f / kLog2e));
            atomicAdd(dvb + 2 * cc, dv2[c].x);
            atomicAdd(dvb + 2 * cc + 1, dv2[c].y);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Tensor-core forward (bf16 I/O, hd = 24, P <= 112): the two contractions of each map -- S_j = Q_j K_j^T (k = hd padded
// to 32) and O_j = softmax(S_j) V (k = P padded to 112) -- run as mma.sync.m16n8k16 bf16 with fp32 accumulators; the
// probabilities go from the accumulator layout straight into the A-operand layout (two adjacent n8 tiles = one k16
// step), so nothing but K and V^T ever sits in shared memory.  A warp owns 16 query tokens at a time; softmax, the
// lambda-combine and the RMSNorm reduce over the 4 lanes of a row with two shuffles.  Same saved tensors (log-sum-exps,
// normalised O0 | O1) as the FMA kernel, so the backward kernels are shared.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    const __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t *>(&t);
}

constexpr int kMmaNT = 14, kMmaPmax = 8 * kMmaNT;                // 112 pooled tokens at most; hd = 24 (shipped) or 32
constexpr int kMmaKS = 40;                                       // K row stride (bf16): <= 32 channels + zero pad, conflict-free
constexpr int kMmaVS = kMmaPmax + 8;                             // V^T row stride (bf16)
constexpr int kMmaTok = 256;                                     // query tokens per block (4 warps x 4 tiles of 16)

// K (both maps) / V of one (image, head pair) into the shared-memory layouts the mma kernels read, from 16-byte loads
// (one per 8 channels of a pooled token).  The first version moved one bf16 per loop iteration with three index
// divisions each: 37 % of the forward kernel's instructions and 40 % of its stall samples sat in front of the first
// barrier (ncu source page, stage-1 shape).  The caller zero-fills the arrays (padding rows / channels) and
// synchronises before this runs.  sK [2][kMmaPmax][kMmaKS]; optional sKt [2][HD][kMmaVS], sV [kMmaPmax][2 HD + 8],
// sVt [2 HD][kMmaVS].
template <int HD, int PM>
__device__ __forceinline__ void pooled_stage_kv(const __nv_bfloat16 *kb, const __nv_bfloat16 *vb, long long ldkv, int P,
                                                __nv_bfloat16 *sK, __nv_bfloat16 *sKt, __nv_bfloat16 *sV,
                                                __nv_bfloat16 *sVt) {
    constexpr int VPR = 2 * HD / 8, VR = 2 * HD + 8, VS = PM + 8;   // PM: pooled-token capacity of the shared-memory arrays
    for (int i = threadIdx.x; i < P * VPR; i += blockDim.x) {
        const int pp = i / VPR, v = i - pp * VPR;
        const uint4 kq = __ldg(reinterpret_cast<const uint4 *>(kb + (long long)pp * ldkv + v * 8));
        const uint4 vq = __ldg(reinterpret_cast<const uint4 *>(vb + (long long)pp * ldkv + v * 8));
        const uint32_t kw[4] = {kq.x, kq.y, kq.z, kq.w}, vw[4] = {vq.x, vq.y, vq.z, vq.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int c = v * 8 + 2 * e;
            const int j = c >= HD ? 1 : 0, d = c - j * HD;
            *reinterpret_cast<uint32_t *>(sK + ((size_t)j * PM + pp) * kMmaKS + d) = kw[e];
            if (sKt) {
                const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162 *>(&kw[e]);
                sKt[((size_t)j * HD + d) * VS + pp] = h2.x;
                sKt[((size_t)j * HD + d + 1) * VS + pp] = h2.y;
            }
            if (sVt) {
                const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162 *>(&vw[e]);
                sVt[(size_t)c * VS + pp] = h2.x;
                sVt[(size_t)(c + 1) * VS + pp] = h2.y;
            }
        }
        if (sV) *reinterpret_cast<uint4 *>(sV + (size_t)pp * VR + v * 8) = vq;
    }
}
__device__ __forceinline__ void smem_zero16(void *ptr, int bytes) {
    for (int i = threadIdx.x * 16; i < bytes; i += blockDim.x * 16) *reinterpret_cast<uint4 *>(static_cast<unsigned char *>(ptr) + i) = make_uint4(0, 0, 0, 0);
}

// NT n8-tiles of pooled tokens per chunk, NCH chunks: (14, 1) covers P <= 112 in one piece (the shipped P = 100); (16, 2)
// covers P <= 256 (config 5) with an online softmax across the two chunks -- the running maximum and sum of a row are
// rescaled once per chunk, so the S tile in registers stays NT x 4 floats.
template <int HD, int NT, int NCH>
__global__ void __launch_bounds__(128) pooled_attn_fwd_mma_kernel(const PooledAttnParams p) {
    constexpr int NC = 2 * HD / 8;      // n8 tiles over the 2hd output channels
    constexpr int PM = NT * 8 * NCH, VS = PM + 8;
    extern __shared__ __align__(16) unsigned char fw_smem[];
    auto sK = reinterpret_cast<__nv_bfloat16 (*)[PM][kMmaKS]>(fw_smem);                                   // K_j[p][d]
    auto sVt = reinterpret_cast<__nv_bfloat16 (*)[VS]>(fw_smem + sizeof(__nv_bfloat16) * 2 * PM * kMmaKS);   // V^T[c][p]
    const int b = blockIdx.z, m = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    {   // stage K (both maps, zero-padded to 32 channels / 112 rows) and V^T
        const __nv_bfloat16 *kb = static_cast<const __nv_bfloat16 *>(p.kp) + (long long)b * p.P * p.ldkv + (long long)m * 2 * HD;
        const __nv_bfloat16 *vb = static_cast<const __nv_bfloat16 *>(p.vp) + (long long)b * p.P * p.ldkv + (long long)m * 2 * HD;
        smem_zero16(fw_smem, (int)(sizeof(__nv_bfloat16) * (2 * PM * kMmaKS + 2 * HD * VS)));
        __syncthreads();
        pooled_stage_kv<HD, PM>(kb, vb, p.ldkv, p.P, &sK[0][0][0], nullptr, nullptr, &sVt[0][0]);
    }
    __syncthreads();
    const float qs = p.scale2 * kLog2e;
    const float lam = __ldg(p.lamp);
    float wv[NC][2];
#pragma unroll
    for (int nc = 0; nc < NC; ++nc) {
        wv[nc][0] = __ldg(p.subln_w + nc * 8 + 2 * t) * p.post;
        wv[nc][1] = __ldg(p.subln_w + nc * 8 + 2 * t + 1) * p.post;
    }
    for (int tile = warp; tile < kMmaTok / 16; tile += 4) {
        const int n0 = blockIdx.x * kMmaTok + tile * 16;
        if (n0 >= p.N) break;
        const int nr[2] = {n0 + g, n0 + g + 8};
        const bool ok[2] = {nr[0] < p.N, nr[1] < p.N};
        float O[2][NC][4];
        float lsev[2][2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            // ---- Q_j fragments (k = 32: channels 24..31 are zero)
            uint32_t qa[2][4];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const __nv_bfloat16 *qp = static_cast<const __nv_bfloat16 *>(p.q) + ((long long)b * p.N + nr[r]) * p.ldq +
                                          (long long)m * 2 * HD + j * HD;
                qa[0][r] = ok[r] ? *reinterpret_cast<const uint32_t *>(qp + 2 * t) : 0u;
                qa[0][2 + r] = ok[r] ? *reinterpret_cast<const uint32_t *>(qp + 8 + 2 * t) : 0u;
                qa[1][r] = ok[r] ? *reinterpret_cast<const uint32_t *>(qp + 16 + 2 * t) : 0u;
                qa[1][2 + r] = (HD > 24 && ok[r]) ? *reinterpret_cast<const uint32_t *>(qp + 24 + 2 * t) : 0u;
            }
            float mx0 = -INFINITY, mx1 = -INFINITY, l0 = 0.f, l1 = 0.f;
#pragma unroll
            for (int nc = 0; nc < NC; ++nc) O[j][nc][0] = O[j][nc][1] = O[j][nc][2] = O[j][nc][3] = 0.f;
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                const int pbase = ch * NT * 8;
                // ---- S_j = Q_j K_j^T (this chunk's pooled tokens)
                float S[NT][4];
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    S[nt][0] = S[nt][1] = S[nt][2] = S[nt][3] = 0.f;
                    const __nv_bfloat16 *kr = &sK[j][pbase + nt * 8 + g][2 * t];
                    mma_bf16_16816(S[nt], qa[0], *reinterpret_cast<const uint32_t *>(kr), *reinterpret_cast<const uint32_t *>(kr + 8));
                    mma_bf16_16816(S[nt], qa[1], *reinterpret_cast<const uint32_t *>(kr + 16), *reinterpret_cast<const uint32_t *>(kr + 24));
                }
                // ---- softmax over the P valid columns (exp2 domain); across chunks: running maximum, rescaled sum and O
                float c0m = -INFINITY, c1m = -INFINITY;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const int c0 = pbase + nt * 8 + 2 * t;
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        S[nt][e] = (c0 + (e & 1) < p.P) ? S[nt][e] * qs : -INFINITY;
                    }
                    c0m = fmaxf(c0m, fmaxf(S[nt][0], S[nt][1]));
                    c1m = fmaxf(c1m, fmaxf(S[nt][2], S[nt][3]));
                }
                c0m = fmaxf(c0m, __shfl_xor_sync(0xffffffffu, c0m, 1)); c0m = fmaxf(c0m, __shfl_xor_sync(0xffffffffu, c0m, 2));
                c1m = fmaxf(c1m, __shfl_xor_sync(0xffffffffu, c1m, 1)); c1m = fmaxf(c1m, __shfl_xor_sync(0xffffffffu, c1m, 2));
                if (NCH == 1) {
                    mx0 = c0m, mx1 = c1m;
                } else {
                    const float n0 = fmaxf(mx0, c0m), n1 = fmaxf(mx1, c1m);      // finite from the first chunk on (P >= 1)
                    const float s0 = ex2_approx(mx0 - n0), s1 = ex2_approx(mx1 - n1);
                    mx0 = n0, mx1 = n1;
                    l0 *= s0, l1 *= s1;
#pragma unroll
                    for (int nc = 0; nc < NC; ++nc) {
                        O[j][nc][0] *= s0; O[j][nc][1] *= s0;
                        O[j][nc][2] *= s1; O[j][nc][3] *= s1;
                    }
                }
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    S[nt][0] = ex2_approx(S[nt][0] - mx0); S[nt][1] = ex2_approx(S[nt][1] - mx0);
                    S[nt][2] = ex2_approx(S[nt][2] - mx1); S[nt][3] = ex2_approx(S[nt][3] - mx1);
                    l0 += S[nt][0] + S[nt][1];
                    l1 += S[nt][2] + S[nt][3];
                }
                // ---- O_j += P V  (probabilities: accumulator layout -> A operand)
#pragma unroll
                for (int kk = 0; kk < NT / 2; ++kk) {
                    uint32_t pa[4];
                    pa[0] = pack_bf16(S[2 * kk][0], S[2 * kk][1]);
                    pa[1] = pack_bf16(S[2 * kk][2], S[2 * kk][3]);
                    pa[2] = pack_bf16(S[2 * kk + 1][0], S[2 * kk + 1][1]);
                    pa[3] = pack_bf16(S[2 * kk + 1][2], S[2 * kk + 1][3]);
#pragma unroll
                    for (int nc = 0; nc < NC; ++nc) {
                        const __nv_bfloat16 *vr = &sVt[nc * 8 + g][pbase + kk * 16 + 2 * t];
                        mma_bf16_16816(O[j][nc], pa, *reinterpret_cast<const uint32_t *>(vr), *reinterpret_cast<const uint32_t *>(vr + 8));
                    }
                }
            }
            l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
            l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
            lsev[j][0] = mx0 + lg2_approx(l0);
            lsev[j][1] = mx1 + lg2_approx(l1);
            const float i0 = 1.f / l0, i1 = 1.f / l1;
#pragma unroll
            for (int nc = 0; nc < NC; ++nc) {
                O[j][nc][0] *= i0; O[j][nc][1] *= i0;
                O[j][nc][2] *= i1; O[j][nc][3] *= i1;
            }
        }
        // ---- o = O0 - lam O1, RMSNorm over the 48 channels of the row, scale, store
        float ss0 = 0.f, ss1 = 0.f;
        float oc[NC][4];
#pragma unroll
        for (int nc = 0; nc < NC; ++nc)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                oc[nc][e] = O[0][nc][e] - lam * O[1][nc][e];
                if (e < 2) ss0 = fmaf(oc[nc][e], oc[nc][e], ss0); else ss1 = fmaf(oc[nc][e], oc[nc][e], ss1);
            }
        ss0 += __shfl_xor_sync(0xffffffffu, ss0, 1); ss0 += __shfl_xor_sync(0xffffffffu, ss0, 2);
        ss1 += __shfl_xor_sync(0xffffffffu, ss1, 1); ss1 += __shfl_xor_sync(0xffffffffu, ss1, 2);
        const float rr[2] = {1.f / sqrtf(ss0 * (1.f / (2 * HD)) + p.eps), 1.f / sqrtf(ss1 * (1.f / (2 * HD)) + p.eps)};
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            if (!ok[r]) continue;
            const long long tok = (long long)b * p.N + nr[r];
            __nv_bfloat16 *op = static_cast<__nv_bfloat16 *>(p.out) + tok * p.ldo + (long long)m * 2 * HD;
#pragma unroll
            for (int nc = 0; nc < NC; ++nc)
                *reinterpret_cast<uint32_t *>(op + nc * 8 + 2 * t) =
                    pack_bf16(oc[nc][2 * r] * rr[r] * wv[nc][0], oc[nc][2 * r + 1] * rr[r] * wv[nc][1]);
            if (p.lse) {
                if (t == 0) {
                    p.lse[(tok * p.h + m) * 2 + 0] = lsev[0][r];
                    p.lse[(tok * p.h + m) * 2 + 1] = lsev[1][r];
                }
                float *os = p.lse + pooled_osave_offset(p) + (tok * p.h + m) * 4 * HD;
#pragma unroll
                for (int nc = 0; nc < NC; ++nc) {
                    *reinterpret_cast<float2 *>(os + nc * 8 + 2 * t) = make_float2(O[0][nc][2 * r], O[0][nc][2 * r + 1]);
                    *reinterpret_cast<float2 *>(os + 2 * HD + nc * 8 + 2 * t) = make_float2(O[1][nc][2 * r], O[1][nc][2 * r + 1]);
                }
            }
        }
    }
}

// Tensor-core backward, token-owner half: a warp owns 16 query tokens; RMSNorm backward on the saved O0 | O1 gives dO
// (accumulator layout), dab = dO V^T, S_j is recomputed, dS_j = A_j o (dab - D_j) and dq_j = dS_j K_j are mma products
// (B operands: V row-major, K_j row-major for S, K_j^T for dq).  Also writes dO / D_j for the dK, dV kernel and
// accumulates d lambda and d subln_w.

template <int HD, int NT, int NCH>
__global__ void __launch_bounds__(128) pooled_attn_bwd_q_mma_kernel(const PooledAttnParams p) {
    constexpr int NC = 2 * HD / 8, KC = 2 * HD / 16, ND = HD / 8, kMmaVR = 2 * HD + 8;
    constexpr int PM = NT * 8 * NCH, VS = PM + 8;       // pooled-token capacity: NCH chunks of NT n8-tiles (see the forward kernel)
    extern __shared__ __align__(16) unsigned char bq_smem[];
    auto sK = reinterpret_cast<__nv_bfloat16 (*)[PM][kMmaKS]>(bq_smem);                                  // K_j[p][d]
    auto sKt = reinterpret_cast<__nv_bfloat16 (*)[HD][VS]>(bq_smem + sizeof(__nv_bfloat16) * 2 * PM * kMmaKS);   // K_j^T[d][p]
    auto sV = reinterpret_cast<__nv_bfloat16 (*)[kMmaVR]>(bq_smem + sizeof(__nv_bfloat16) * (2 * PM * kMmaKS + 2 * HD * VS));   // V[p][c]
    __shared__ float red[2 * HD + 1];
    const int b = blockIdx.z, m = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    {
        const __nv_bfloat16 *kb = static_cast<const __nv_bfloat16 *>(p.kp) + (long long)b * p.P * p.ldkv + (long long)m * 2 * HD;
        const __nv_bfloat16 *vb = static_cast<const __nv_bfloat16 *>(p.vp) + (long long)b * p.P * p.ldkv + (long long)m * 2 * HD;
        smem_zero16(bq_smem, (int)(sizeof(__nv_bfloat16) * (2 * PM * kMmaKS + 2 * HD * VS + PM * kMmaVR)));
        __syncthreads();
        pooled_stage_kv<HD, PM>(kb, vb, p.ldkv, p.P, &sK[0][0][0], &sKt[0][0][0], &sV[0][0], nullptr);
        for (int i = threadIdx.x; i < 2 * HD + 1; i += blockDim.x) red[i] = 0.f;
    }
    __syncthreads();
    const float qs = p.scale2 * kLog2e;
    const float lam = __ldg(p.lamp);
    float wv[NC][2], dwacc[NC][2];
#pragma unroll
    for (int nc = 0; nc < NC; ++nc) {
        wv[nc][0] = __ldg(p.subln_w + nc * 8 + 2 * t);
        wv[nc][1] = __ldg(p.subln_w + nc * 8 + 2 * t + 1);
        dwacc[nc][0] = dwacc[nc][1] = 0.f;
    }
    float dlam = 0.f;
    for (int tile = warp; tile < kMmaTok / 16; tile += 4) {
        const int n0 = blockIdx.x * kMmaTok + tile * 16;
        if (n0 >= p.N) break;
        const int nr[2] = {n0 + g, n0 + g + 8};
        const bool ok[2] = {nr[0] < p.N, nr[1] < p.N};
        // ---- RMSNorm backward on the saved O0 | O1 (accumulator layout: row r, columns nc*8 + 2t, +1)
        float dO[NC][4];
        float lse0[2], lse1[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const long long tok = (long long)b * p.N + (ok[r] ? nr[r] : 0);
            const float *os = p.lse + pooled_osave_offset(p) + (tok * p.h + m) * 4 * HD;
            const __nv_bfloat16 *gp = static_cast<const __nv_bfloat16 *>(p.dout) + tok * p.lddo + (long long)m * 2 * HD;
            float o0[NC][2], o1[NC][2], oc[NC][2], gs[NC][2];
            float ss = 0.f;
#pragma unroll
            for (int nc = 0; nc < NC; ++nc) {
                const float2 a = ok[r] ? *reinterpret_cast<const float2 *>(os + nc * 8 + 2 * t) : make_float2(0.f, 0.f);
                const float2 c = ok[r] ? *reinterpret_cast<const float2 *>(os + 2 * HD + nc * 8 + 2 * t) : make_float2(0.f, 0.f);
                o0[nc][0] = a.x, o0[nc][1] = a.y, o1[nc][0] = c.x, o1[nc][1] = c.y;
                oc[nc][0] = a.x - lam * c.x, oc[nc][1] = a.y - lam * c.y;
                ss = fmaf(oc[nc][0], oc[nc][0], fmaf(oc[nc][1], oc[nc][1], ss));
                const uint32_t gg = ok[r] ? *reinterpret_cast<const uint32_t *>(gp + nc * 8 + 2 * t) : 0u;
                gs[nc][0] = __uint_as_float(gg << 16), gs[nc][1] = __uint_as_float(gg & 0xffff0000u);
            }
            ss += __shfl_xor_sync(0xffffffffu, ss, 1); ss += __shfl_xor_sync(0xffffffffu, ss, 2);
            const float rn = 1.f / sqrtf(ss * (1.f / (2 * HD)) + p.eps);
            float dot = 0.f;
#pragma unroll
            for (int nc = 0; nc < NC; ++nc)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    dwacc[nc][e] = fmaf(gs[nc][e] * p.post, oc[nc][e] * rn, dwacc[nc][e]);
                    gs[nc][e] *= p.post * wv[nc][e];
                    dot = fmaf(gs[nc][e], oc[nc][e], dot);
                }
            dot += __shfl_xor_sync(0xffffffffu, dot, 1); dot += __shfl_xor_sync(0xffffffffu, dot, 2);
            const float k3 = rn * rn * rn * dot * (1.f / (2 * HD));
            float d0 = 0.f, d1 = 0.f;
#pragma unroll
            for (int nc = 0; nc < NC; ++nc)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const float v = rn * gs[nc][e] - oc[nc][e] * k3;
                    dO[nc][2 * r + e] = v;
                    d0 = fmaf(v, o0[nc][e], d0);
                    d1 = fmaf(v, o1[nc][e], d1);
                }
            d0 += __shfl_xor_sync(0xffffffffu, d0, 1); d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
            d1 += __shfl_xor_sync(0xffffffffu, d1, 1); d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
            (void)d0;
            lse0[r] = ok[r] ? p.lse[(tok * p.h + m) * 2 + 0] : 0.f;
            lse1[r] = ok[r] ? p.lse[(tok * p.h + m) * 2 + 1] : 0.f;
            if (ok[r]) {
                if (t == 0) dlam -= d1;
                if (p.ws_dO) {
                    float *wdO = p.ws_dO + (tok * p.h + m) * 2 * HD;
#pragma unroll
                    for (int nc = 0; nc < NC; ++nc)
                        *reinterpret_cast<float2 *>(wdO + nc * 8 + 2 * t) = make_float2(dO[nc][2 * r], dO[nc][2 * r + 1]);
                }
            }
        }
        if constexpr (NCH == 1) {
        // ---- dab = dO V^T  (k = 48 channels = 3 steps; dO: accumulator layout -> A operand)
        float dab[NT][4];
        {
            uint32_t da[KC][4];
#pragma unroll
            for (int kc = 0; kc < KC; ++kc) {
                da[kc][0] = pack_bf16(dO[2 * kc][0], dO[2 * kc][1]);
                da[kc][1] = pack_bf16(dO[2 * kc][2], dO[2 * kc][3]);
                da[kc][2] = pack_bf16(dO[2 * kc + 1][0], dO[2 * kc + 1][1]);
                da[kc][3] = pack_bf16(dO[2 * kc + 1][2], dO[2 * kc + 1][3]);
            }
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                dab[nt][0] = dab[nt][1] = dab[nt][2] = dab[nt][3] = 0.f;
#pragma unroll
                for (int kc = 0; kc < KC; ++kc) {
                    const __nv_bfloat16 *vr = &sV[nt * 8 + g][kc * 16 + 2 * t];
                    mma_bf16_16816(dab[nt], da[kc], *reinterpret_cast<const uint32_t *>(vr), *reinterpret_cast<const uint32_t *>(vr + 8));
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            uint32_t qa[2][4];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const __nv_bfloat16 *qp = static_cast<const __nv_bfloat16 *>(p.q) + ((long long)b * p.N + (ok[r] ? nr[r] : 0)) * p.ldq +
                                          (long long)m * 2 * HD + j * HD;
                qa[0][r] = ok[r] ? *reinterpret_cast<const uint32_t *>(qp + 2 * t) : 0u;
                qa[0][2 + r] = ok[r] ? *reinterpret_cast<const uint32_t *>(qp + 8 + 2 * t) : 0u;
                qa[1][r] = ok[r] ? *reinterpret_cast<const uint32_t *>(qp + 16 + 2 * t) : 0u;
                qa[1][2 + r] = (HD > 24 && ok[r]) ? *reinterpret_cast<const uint32_t *>(qp + 24 + 2 * t) : 0u;
            }
            const float ls[2] = {j == 0 ? lse0[0] : lse1[0], j == 0 ? lse0[1] : lse1[1]};
            const float sgn = j == 0 ? 1.f : -lam;
            // D_j = sum_p A_j[p] dab[p], formed from the SAME probabilities and the SAME (bf16-operand) dab values that
            // dS_j = A_j (dab - D_j) uses below, so that sum_p dS_j[p] = 0 holds to rounding.  (Round 1 took D_j = dO . O_j
            // from the fp32 dO and the saved O_j: mathematically equal, but dab is computed from bf16-rounded dO, and
            // wherever the softmax is nearly uniform -- stage 0 / 1, logits damped by the double scale -- dab - D_j is a
            // difference of nearly equal numbers: dq came out at 3 - 8x its own magnitude in error against fp64,
            // tests/test_parity_shipped_gpu.py.)
            float Dj[2] = {0.f, 0.f};
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                float S[4] = {0.f, 0.f, 0.f, 0.f};
                const __nv_bfloat16 *kr = &sK[j][nt * 8 + g][2 * t];
                mma_bf16_16816(S, qa[0], *reinterpret_cast<const uint32_t *>(kr), *reinterpret_cast<const uint32_t *>(kr + 8));
                mma_bf16_16816(S, qa[1], *reinterpret_cast<const uint32_t *>(kr + 16), *reinterpret_cast<const uint32_t *>(kr + 24));
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int r = e >> 1;
                    const bool in = nt * 8 + 2 * t + (e & 1) < p.P;
                    const float a = in ? ex2_approx(S[e] * qs - ls[r]) : 0.f;
                    Dj[r] = fmaf(a, dab[nt][e], Dj[r]);
                }
            }
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                Dj[r] += __shfl_xor_sync(0xffffffffu, Dj[r], 1);
                Dj[r] += __shfl_xor_sync(0xffffffffu, Dj[r], 2);
                if (p.ws_D && ok[r] && t == 0) p.ws_D[(((long long)b * p.N + nr[r]) * p.h + m) * 2 + j] = Dj[r];
            }
            float dq[ND][4];
#pragma unroll
            for (int nd = 0; nd < ND; ++nd) dq[nd][0] = dq[nd][1] = dq[nd][2] = dq[nd][3] = 0.f;
#pragma unroll
            for (int kk = 0; kk < NT / 2; ++kk) {
                uint32_t sa[4];
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    const int nt = 2 * kk + hf;
                    float S[4] = {0.f, 0.f, 0.f, 0.f};
                    const __nv_bfloat16 *kr = &sK[j][nt * 8 + g][2 * t];
                    mma_bf16_16816(S, qa[0], *reinterpret_cast<const uint32_t *>(kr), *reinterpret_cast<const uint32_t *>(kr + 8));
                    mma_bf16_16816(S, qa[1], *reinterpret_cast<const uint32_t *>(kr + 16), *reinterpret_cast<const uint32_t *>(kr + 24));
                    float ds[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int r = e >> 1;
                        const bool in = nt * 8 + 2 * t + (e & 1) < p.P;
                        const float a = in ? ex2_approx(S[e] * qs - ls[r]) : 0.f;
                        ds[e] = sgn * a * (dab[nt][e] - Dj[r]);
                    }
                    sa[2 * hf] = pack_bf16(ds[0], ds[1]);
                    sa[2 * hf + 1] = pack_bf16(ds[2], ds[3]);
                }
#pragma unroll
                for (int nd = 0; nd < ND; ++nd) {
                    const __nv_bfloat16 *kt = &sKt[j][nd * 8 + g][kk * 16 + 2 * t];
                    mma_bf16_16816(dq[nd], sa, *reinterpret_cast<const uint32_t *>(kt), *reinterpret_cast<const uint32_t *>(kt + 8));
                }
            }
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                if (!ok[r]) continue;
                __nv_bfloat16 *dqp = static_cast<__nv_bfloat16 *>(p.dq) + ((long long)b * p.N + nr[r]) * p.lddq +
                                     (long long)m * 2 * HD + j * HD;
#pragma unroll
                for (int nd = 0; nd < ND; ++nd)
                    *reinterpret_cast<uint32_t *>(dqp + nd * 8 + 2 * t) =
                        pack_bf16(dq[nd][2 * r] * p.scale2, dq[nd][2 * r + 1] * p.scale2);
            }
        }
        } else
        {
            // ---- NCH > 1: two sweeps over the chunks.  Sweep 1 forms D_j = sum_p A_j[p] dab[p] over ALL pooled tokens, sweep 2
            // dS_j = A_j (dab - D_j) and dq_j += dS_j K_j; dab = dO V^T of a chunk is recomputed in each sweep (NT x 4 floats
            // in registers instead of NT x NCH x 4), from the same bf16 operands both times.
            uint32_t da[KC][4];
#pragma unroll
            for (int kc = 0; kc < KC; ++kc) {
                da[kc][0] = pack_bf16(dO[2 * kc][0], dO[2 * kc][1]);
                da[kc][1] = pack_bf16(dO[2 * kc][2], dO[2 * kc][3]);
                da[kc][2] = pack_bf16(dO[2 * kc + 1][0], dO[2 * kc + 1][1]);
                da[kc][3] = pack_bf16(dO[2 * kc + 1][2], dO[2 * kc + 1][3]);
            }
            uint32_t qa[2][2][4];
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const __nv_bfloat16 *qp = static_cast<const __nv_bfloat16 *>(p.q) + ((long long)b * p.N + (ok[r] ? nr[r] : 0)) * p.ldq +
                                              (long long)m * 2 * HD + j * HD;
                    qa[j][0][r] = ok[r] ? *reinterpret_cast<const uint32_t *>(qp + 2 * t) : 0u;
                    qa[j][0][2 + r] = ok[r] ? *reinterpret_cast<const uint32_t *>(qp + 8 + 2 * t) : 0u;
                    qa[j][1][r] = ok[r] ? *reinterpret_cast<const uint32_t *>(qp + 16 + 2 * t) : 0u;
                    qa[j][1][2 + r] = (HD > 24 && ok[r]) ? *reinterpret_cast<const uint32_t *>(qp + 24 + 2 * t) : 0u;
                }
            auto dab_chunk = [&](int pbase, float (&dab)[NT][4]) {
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    dab[nt][0] = dab[nt][1] = dab[nt][2] = dab[nt][3] = 0.f;
#pragma unroll
                    for (int kc = 0; kc < KC; ++kc) {
                        const __nv_bfloat16 *vr = &sV[pbase + nt * 8 + g][kc * 16 + 2 * t];
                        mma_bf16_16816(dab[nt], da[kc], *reinterpret_cast<const uint32_t *>(vr), *reinterpret_cast<const uint32_t *>(vr + 8));
                    }
                }
            };
            auto probs = [&](int j, int pbase, int nt, float (&a)[4]) {       // A_j of one n8 tile from the saved lse
                float S[4] = {0.f, 0.f, 0.f, 0.f};
                const __nv_bfloat16 *kr = &sK[j][pbase + nt * 8 + g][2 * t];
                mma_bf16_16816(S, qa[j][0], *reinterpret_cast<const uint32_t *>(kr), *reinterpret_cast<const uint32_t *>(kr + 8));
                mma_bf16_16816(S, qa[j][1], *reinterpret_cast<const uint32_t *>(kr + 16), *reinterpret_cast<const uint32_t *>(kr + 24));
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int r = e >> 1;
                    const bool in = pbase + nt * 8 + 2 * t + (e & 1) < p.P;
                    a[e] = in ? ex2_approx(S[e] * qs - (j == 0 ? lse0[r] : lse1[r])) : 0.f;
                }
            };
            float Dj[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
#pragma unroll 1
            for (int ch = 0; ch < NCH; ++ch) {
                float dab[NT][4];
                dab_chunk(ch * NT * 8, dab);
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) {
                        float a[4];
                        probs(j, ch * NT * 8, nt, a);
#pragma unroll
                        for (int e = 0; e < 4; ++e) Dj[j][e >> 1] = fmaf(a[e], dab[nt][e], Dj[j][e >> 1]);
                    }
            }
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    Dj[j][r] += __shfl_xor_sync(0xffffffffu, Dj[j][r], 1);
                    Dj[j][r] += __shfl_xor_sync(0xffffffffu, Dj[j][r], 2);
                    if (p.ws_D && ok[r] && t == 0) p.ws_D[(((long long)b * p.N + nr[r]) * p.h + m) * 2 + j] = Dj[j][r];
                }
            float dq[2][ND][4];
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int nd = 0; nd < ND; ++nd) dq[j][nd][0] = dq[j][nd][1] = dq[j][nd][2] = dq[j][nd][3] = 0.f;
#pragma unroll 1
            for (int ch = 0; ch < NCH; ++ch) {
                float dab[NT][4];
                dab_chunk(ch * NT * 8, dab);
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const float sgn = j == 0 ? 1.f : -lam;
#pragma unroll
                    for (int kk = 0; kk < NT / 2; ++kk) {
                        uint32_t sa[4];
#pragma unroll
                        for (int hf = 0; hf < 2; ++hf) {
                            const int nt = 2 * kk + hf;
                            float a[4], ds[4];
                            probs(j, ch * NT * 8, nt, a);
#pragma unroll
                            for (int e = 0; e < 4; ++e) ds[e] = sgn * a[e] * (dab[nt][e] - Dj[j][e >> 1]);
                            sa[2 * hf] = pack_bf16(ds[0], ds[1]);
                            sa[2 * hf + 1] = pack_bf16(ds[2], ds[3]);
                        }
#pragma unroll
                        for (int nd = 0; nd < ND; ++nd) {
                            const __nv_bfloat16 *kt = &sKt[j][nd * 8 + g][ch * NT * 8 + kk * 16 + 2 * t];
                            mma_bf16_16816(dq[j][nd], sa, *reinterpret_cast<const uint32_t *>(kt), *reinterpret_cast<const uint32_t *>(kt + 8));
                        }
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    if (!ok[r]) continue;
                    __nv_bfloat16 *dqp = static_cast<__nv_bfloat16 *>(p.dq) + ((long long)b * p.N + nr[r]) * p.lddq +
                                         (long long)m * 2 * HD + j * HD;
#pragma unroll
                    for (int nd = 0; nd < ND; ++nd)
                        *reinterpret_cast<uint32_t *>(dqp + nd * 8 + 2 * t) =
                            pack_bf16(dq[j][nd][2 * r] * p.scale2, dq[j][nd][2 * r + 1] * p.scale2);
                }
        }

    }
    // ---- d subln_w (column sums over the block's tokens) and d lambda
#pragma unroll
    for (int nc = 0; nc < NC; ++nc)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            float v = dwacc[nc][e];
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            if (g == 0) atomicAdd(&red[nc * 8 + 2 * t + e], v);
        }
#pragma unroll
    for (int o2 = 16; o2 > 0; o2 >>= 1) dlam += __shfl_xor_sync(0xffffffffu, dlam, o2);
    if (lane == 0) atomicAdd(&red[2 * HD], dlam);
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * HD + 1; i += blockDim.x) {
        if (i < 2 * HD) atomicAdd(p.d_subln_w + i, red[i]);
        else atomicAdd(p.d_lambda, red[i]);
    }
}

// Tensor-core backward, pooled-token-owner half (dK, dV): warp w owns the 16 pooled tokens [16w, 16w+16) with their K / V
// rows as A-operand registers and walks the block's query tokens 16 at a time: S_j^T = K_j Q_j^T and dab^T = V dO^T
// (B operands: Q, dO row-major in shared memory), probabilities and dS_j^T in the accumulator layout, then
// dV += Abar^T dO and dK_j += dS_j^T Q_j with the accumulator tiles reused as A operands (B operands: dO^T, Q_j^T,
// staged transposed).  One fp32 atomic per element and block at the end.
constexpr int kKvChunk = 64;                  // query tokens staged per iteration
constexpr int kKvSlab = 1024;                 // query tokens per block
constexpr int kKvTS = kKvChunk + 8;           // row stride of the transposed tiles (bf16), conflict-free

template <int HD>
__global__ void __launch_bounds__(224) pooled_attn_bwd_kv_mma_kernel(const PooledAttnParams p) {
    constexpr int NC = 2 * HD / 8, KC = 2 * HD / 16, ND = HD / 8, kMmaVR = 2 * HD + 8;
    __shared__ __align__(16) __nv_bfloat16 sQ[2][kKvChunk][kMmaKS];      // Q_j[tok][d] (zero-padded to 32 channels)
    __shared__ __align__(16) __nv_bfloat16 sQt[2][HD][kKvTS];            // Q_j^T[d][tok]
    __shared__ __align__(16) __nv_bfloat16 sG[kKvChunk][kMmaVR];         // dO[tok][c]
    __shared__ __align__(16) __nv_bfloat16 sGt[2 * HD][kKvTS];           // dO^T[c][tok]
    __shared__ float sl[kKvChunk][4];                                    // lse0, lse1, D0, D1
    // blockIdx.y = (chunk of 112 pooled tokens) * h + head pair: P > 112 (config 5: 256) takes several CTAs per slab
    const int b = blockIdx.z, m = blockIdx.y % p.h, p0 = (blockIdx.y / p.h) * kMmaPmax;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int prow[2] = {p0 + warp * 16 + g, p0 + warp * 16 + g + 8};
    const bool pok[2] = {prow[0] < p.P, prow[1] < p.P};
    const bool warp_on = p0 + warp * 16 < p.P;
    // ---- A operands: K_j rows (k = 32 channels, 24..31 zero) and V rows (k = 48 channels) of this warp's pooled tokens
    uint32_t ka[2][2][4], va[KC][4];
    {
        const __nv_bfloat16 *kb = static_cast<const __nv_bfloat16 *>(p.kp) + (long long)b * p.P * p.ldkv + (long long)m * 2 * HD;
        const __nv_bfloat16 *vb = static_cast<const __nv_bfloat16 *>(p.vp) + (long long)b * p.P * p.ldkv + (long long)m * 2 * HD;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const __nv_bfloat16 *kr = kb + (long long)(pok[r] ? prow[r] : 0) * p.ldkv;
            const __nv_bfloat16 *vr = vb + (long long)(pok[r] ? prow[r] : 0) * p.ldkv;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                ka[j][0][r] = pok[r] ? *reinterpret_cast<const uint32_t *>(kr + j * HD + 2 * t) : 0u;
                ka[j][0][2 + r] = pok[r] ? *reinterpret_cast<const uint32_t *>(kr + j * HD + 8 + 2 * t) : 0u;
                ka[j][1][r] = pok[r] ? *reinterpret_cast<const uint32_t *>(kr + j * HD + 16 + 2 * t) : 0u;
                ka[j][1][2 + r] = (HD > 24 && pok[r]) ? *reinterpret_cast<const uint32_t *>(kr + j * HD + 24 + 2 * t) : 0u;
            }
#pragma unroll
            for (int kc = 0; kc < KC; ++kc) {
                va[kc][r] = pok[r] ? *reinterpret_cast<const uint32_t *>(vr + kc * 16 + 2 * t) : 0u;
                va[kc][2 + r] = pok[r] ? *reinterpret_cast<const uint32_t *>(vr + kc * 16 + 8 + 2 * t) : 0u;
            }
        }
    }
    float dV[NC][4], dK[2][ND][4];
#pragma unroll
    for (int i = 0; i < NC; ++i) dV[i][0] = dV[i][1] = dV[i][2] = dV[i][3] = 0.f;
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int i = 0; i < ND; ++i) dK[j][i][0] = dK[j][i][1] = dK[j][i][2] = dK[j][i][3] = 0.f;
    const float qs = p.scale2 * kLog2e;
    const float lam = __ldg(p.lamp);
    const __nv_bfloat16 z = __float2bfloat16_rn(0.f);
    const int n0 = blockIdx.x * kKvSlab, n1 = min(p.N, n0 + kKvSlab);
    for (int base = n0; base < n1; base += kKvChunk) {
        __syncthreads();
        // ---- stage Q (both maps), dO, lse / D of the chunk; rows past the end are zero
        for (int i = threadIdx.x; i < kKvChunk * HD; i += blockDim.x) {         // pairs of channels
            const int tk = i / HD, c2 = (i % HD) * 2;                            // c2 in 0..46 over the 48 q channels
            const int n = base + tk;
            uint32_t v = 0u;
            if (n < n1) v = *reinterpret_cast<const uint32_t *>(static_cast<const __nv_bfloat16 *>(p.q) +
                                                              ((long long)b * p.N + n) * p.ldq + (long long)m * 2 * HD + c2);
            const int j = c2 / HD, d = c2 % HD;
            *reinterpret_cast<uint32_t *>(&sQ[j][tk][d]) = v;
            const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162 *>(&v);
            sQt[j][d][tk] = h2.x;
            sQt[j][d + 1][tk] = h2.y;
        }
        for (int i = threadIdx.x; i < 2 * kKvChunk * (kMmaKS - HD); i += blockDim.x) {   // zero pad channels 24..39
            const int j = i / (kKvChunk * (kMmaKS - HD)), tk = (i / (kMmaKS - HD)) % kKvChunk, d = HD + i % (kMmaKS - HD);
            sQ[j][tk][d] = z;
        }
        for (int i = threadIdx.x; i < kKvChunk * HD; i += blockDim.x) {         // pairs of dO channels (fp32 -> bf16)
            const int tk = i / HD, c2 = (i % HD) * 2;
            const int n = base + tk;
            float2 v = make_float2(0.f, 0.f);
            if (n < n1) v = *reinterpret_cast<const float2 *>(p.ws_dO + (((long long)b * p.N + n) * p.h + m) * 2 * HD + c2);
            const __nv_bfloat162 h2 = __floats2bfloat162_rn(v.x, v.y);
            *reinterpret_cast<__nv_bfloat162 *>(&sG[tk][c2]) = h2;
            sGt[c2][tk] = h2.x;
            sGt[c2 + 1][tk] = h2.y;
        }
        for (int i = threadIdx.x; i < kKvChunk * 4; i += blockDim.x) {
            const int tk = i / 4, w = i % 4;
            const int n = base + tk;
            const long long tm = ((long long)b * p.N + n) * p.h + m;
            sl[tk][w] = n < n1 ? (w < 2 ? p.lse[tm * 2 + w] : p.ws_D[tm * 2 + (w - 2)]) : (w < 2 ? INFINITY : 0.f);
        }
        __syncthreads();
        if (!warp_on) continue;
#pragma unroll 1
        for (int tk0 = 0; tk0 < kKvChunk; tk0 += 16) {
            if (base + tk0 >= n1) break;
            // ---- dab^T = V dO^T : rows = pooled tokens, columns = the 16 query tokens (two n8 tiles)
            float dab[2][4];
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                dab[nt][0] = dab[nt][1] = dab[nt][2] = dab[nt][3] = 0.f;
#pragma unroll
                for (int kc = 0; kc < KC; ++kc) {
                    const __nv_bfloat16 *gr = &sG[tk0 + nt * 8 + g][kc * 16 + 2 * t];
                    mma_bf16_16816(dab[nt], va[kc], *reinterpret_cast<const uint32_t *>(gr), *reinterpret_cast<const uint32_t *>(gr + 8));
                }
            }
            uint32_t abar[4] = {0u, 0u, 0u, 0u};
            float ab[2][4];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                uint32_t dsa[4];
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    float S[4] = {0.f, 0.f, 0.f, 0.f};
                    const __nv_bfloat16 *qr = &sQ[j][tk0 + nt * 8 + g][2 * t];
                    mma_bf16_16816(S, ka[j][0], *reinterpret_cast<const uint32_t *>(qr), *reinterpret_cast<const uint32_t *>(qr + 8));
                    mma_bf16_16816(S, ka[j][1], *reinterpret_cast<const uint32_t *>(qr + 16), *reinterpret_cast<const uint32_t *>(qr + 24));
                    float ds[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int tk = tk0 + nt * 8 + 2 * t + (e & 1);       // column = query token
                        const float a = pok[e >> 1] ? ex2_approx(S[e] * qs - sl[tk][j]) : 0.f;   // lse = +inf past the end -> 0
                        ds[e] = (j == 0 ? a : -lam * a) * (dab[nt][e] - sl[tk][2 + j]);
                        ab[nt][e] = j == 0 ? a : ab[nt][e] - lam * a;
                    }
                    dsa[2 * nt] = pack_bf16(ds[0], ds[1]);
                    dsa[2 * nt + 1] = pack_bf16(ds[2], ds[3]);
                }
                // dK_j += dS_j^T Q_j   (A = accumulator tiles of the two n8 token tiles, k = 16 tokens)
                const uint32_t a4[4] = {dsa[0], dsa[1], dsa[2], dsa[3]};
#pragma unroll
                for (int nd = 0; nd < ND; ++nd) {
                    const __nv_bfloat16 *qt = &sQt[j][nd * 8 + g][tk0 + 2 * t];
                    mma_bf16_16816(dK[j][nd], a4, *reinterpret_cast<const uint32_t *>(qt), *reinterpret_cast<const uint32_t *>(qt + 8));
                }
            }
            abar[0] = pack_bf16(ab[0][0], ab[0][1]);
            abar[1] = pack_bf16(ab[0][2], ab[0][3]);
            abar[2] = pack_bf16(ab[1][0], ab[1][1]);
            abar[3] = pack_bf16(ab[1][2], ab[1][3]);
#pragma unroll
            for (int nc = 0; nc < NC; ++nc) {
                const __nv_bfloat16 *gt = &sGt[nc * 8 + g][tk0 + 2 * t];
                mma_bf16_16816(dV[nc], abar, *reinterpret_cast<const uint32_t *>(gt), *reinterpret_cast<const uint32_t *>(gt + 8));
            }
        }
    }
    if (!warp_on) return;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        if (!pok[r]) continue;
        float *dkb = p.dkp + ((long long)b * p.P + prow[r]) * p.ldd + (long long)m * 2 * HD;
        float *dvb = p.dvp + ((long long)b * p.P + prow[r]) * p.ldd + (long long)m * 2 * HD;
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int nd = 0; nd < ND; ++nd) {
                atomicAdd(dkb + j * HD + nd * 8 + 2 * t, dK[j][nd][2 * r] * p.scale2);
                atomicAdd(dkb + j * HD + nd * 8 + 2 * t + 1, dK[j][nd][2 * r + 1] * p.scale2);
            }
#pragma unroll
        for (int nc = 0; nc < NC; ++nc) {
            atomicAdd(dvb + nc * 8 + 2 * t, dV[nc][2 * r]);
            atomicAdd(dvb + nc * 8 + 2 * t + 1, dV[nc][2 * r + 1]);
        }
    }
}

static bool pooled_use_mma() {
    const char *e = getenv("MLAGG_POOLED_MMA");
    return !(e && e[0] == '0');
}

// the vectorised K / V staging of the mma kernels: 16-byte aligned rows and head-pair offsets
static bool pooled_kv16(const PooledAttnParams &p) {
    return p.ldkv % 8 == 0 && reinterpret_cast<uintptr_t>(p.kp) % 16 == 0 && reinterpret_cast<uintptr_t>(p.vp) % 16 == 0;
}

template <typename T, int HD>
static cudaError_t pooled_launch(const PooledAttnParams &p, int which, cudaStream_t st) {
    const size_t smem = (size_t)p.P * HD * 16;   // kI (P x hd float2) + V (P x 2hd float)
    cudaError_t e;
    constexpr bool kMmaHd = HD == 24 || HD == 32;
    constexpr int MH = kMmaHd ? HD : 24;   // instantiate the tensor-core kernels only for the head sizes they support
    constexpr int kBigNT = 16, kBigNCH = 2;                 // P <= 256 in two chunks of 128
    const bool mma_ok = std::is_same<T, __nv_bfloat16>::value && kMmaHd && p.P <= kBigNT * 8 * kBigNCH && pooled_use_mma() &&
                        p.ldq % 2 == 0 && pooled_kv16(p);
    const bool big = p.P > kMmaPmax;
    auto fw_bytes = [](int pm) { return sizeof(__nv_bfloat16) * (size_t)(2 * pm * kMmaKS + 2 * MH * (pm + 8)); };
    auto bq_bytes = [](int pm) { return sizeof(__nv_bfloat16) * (size_t)(2 * pm * kMmaKS + 2 * MH * (pm + 8) + pm * (2 * MH + 8)); };
    auto go = [&](auto k, dim3 grid, int threads, size_t bytes) -> cudaError_t {
        cudaError_t e2 = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e2 != cudaSuccess) return e2;
        k<<<grid, threads, bytes, st>>>(p);
        return cudaSuccess;
    };
    const dim3 gq((p.N + kMmaTok - 1) / kMmaTok, p.h, p.Bn);
    if (which == 0 && mma_ok && p.ldo % 2 == 0) {
        if ((e = big ? go(pooled_attn_fwd_mma_kernel<MH, kBigNT, kBigNCH>, gq, 128, fw_bytes(kBigNT * 8 * kBigNCH))
                     : go(pooled_attn_fwd_mma_kernel<MH, kMmaNT, 1>, gq, 128, fw_bytes(kMmaPmax))) != cudaSuccess) return e;
    } else if (which == 0) {
        auto k = pooled_attn_fwd_kernel<T, HD>;
        if ((e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        k<<<dim3((p.N + kPTok - 1) / kPTok, p.h, p.Bn), kPTok, smem, st>>>(p);
    } else if (which == 1 && mma_ok && p.lddo % 2 == 0 && p.lddq % 2 == 0) {
        if ((e = big ? go(pooled_attn_bwd_q_mma_kernel<MH, kBigNT, kBigNCH>, gq, 128, bq_bytes(kBigNT * 8 * kBigNCH))
                     : go(pooled_attn_bwd_q_mma_kernel<MH, kMmaNT, 1>, gq, 128, bq_bytes(kMmaPmax))) != cudaSuccess) return e;
    } else if (which == 1) {
        auto k = pooled_attn_bwd_q_kernel<T, HD>;
        if ((e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        k<<<dim3((p.N + kPTok - 1) / kPTok, p.h, p.Bn), kPTok, smem, st>>>(p);
    } else if (mma_ok) {
        const int pchunks = (p.P + kMmaPmax - 1) / kMmaPmax;
        pooled_attn_bwd_kv_mma_kernel<MH><<<dim3((p.N + kKvSlab - 1) / kKvSlab, p.h * pchunks, p.Bn), 224, 0, st>>>(p);
    } else {
        if (HD % 4 == 0 && 2 * p.P <= 256) {
            const int threads = ((2 * p.P + 31) / 32) * 32;
            pooled_attn_bwd_kv2_kernel<T, (HD % 4 == 0 ? HD : 4)><<<dim3((p.N + kSlab - 1) / kSlab, p.h, p.Bn), threads, 0, st>>>(p);
        } else {
            const int threads = ((p.P + 31) / 32) * 32;
            pooled_attn_bwd_kv_kernel<T, HD><<<dim3((p.N + kSlab - 1) / kSlab, p.h, p.Bn), threads, 0, st>>>(p);
        }
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
