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
#include <stdlib.h>

#include <type_traits>

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

// =====================================================================================================================
// Tiled variants (round 2).  The thread-per-token kernels above read every neighbour row straight from global memory:
// a warp's 64-bit load touches 32 different 128-byte lines (32 tokens, rows 2C elements apart), the L1 pipe retires one
// line per cycle, and at stage 0 (256 000 tokens x 216 such loads) that alone is the kernel's run time -- 148 us forward,
// 320 us backward for 100 MB of traffic (profiles/trace_r02_mid_summary.txt).  Here a CTA owns an 8 x 16 tile of tokens of
// ONE (image, head pair): the k | v segments of the 10 x 18 halo, the q / dout rows of the tile and every per-token
// result are staged through shared memory with 16-byte chunks (a warp instruction touches ~6 lines), and the threads
// (still one per token, same arithmetic) read their 9 neighbours from shared memory; the per-token record stride is
// padded so that 16 (8) consecutive tokens hit distinct banks with 64-bit (128-bit) loads.
// =====================================================================================================================
constexpr int kTR = 8, kTC = 16, kHC = kTC + 2, kHalo = (kTR + 2) * kHC, kTile = kTR * kTC;

template <typename T, int V>
__device__ __forceinline__ void ldv_s(const unsigned char *bp, float *dst) {   // shared-memory twin of ldv (V % 4 == 0)
    const T *p = reinterpret_cast<const T *>(bp);
#pragma unroll
    for (int i = 0; i < V; i += 4) {
        if constexpr (sizeof(T) == 4) {
            const float4 t = *reinterpret_cast<const float4 *>(p + i);
            dst[i] = t.x; dst[i + 1] = t.y; dst[i + 2] = t.z; dst[i + 3] = t.w;
        } else {
            const uint2 raw = *reinterpret_cast<const uint2 *>(p + i);
            const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&raw.x));
            const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&raw.y));
            dst[i] = a.x; dst[i + 1] = a.y; dst[i + 2] = b.x; dst[i + 3] = b.y;
        }
    }
}
template <typename T, int V>
__device__ __forceinline__ void stv_s(unsigned char *bp, const float *src) {
    T *p = reinterpret_cast<T *>(bp);
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
}

template <typename T, int HD>
struct LT {
    static constexpr int SEG = 2 * HD;                               // elements of one k / v / q / out segment
    static constexpr int SEGB = SEG * (int)sizeof(T);                // its bytes (a multiple of 16 for the tiled path)
    static constexpr int PAD = sizeof(T) == 2 ? 8 : 16;
    static constexpr int CH = sizeof(T) == 2 ? 8 : 16;                // staging chunk of T-typed records
    static constexpr int KVS = 2 * SEGB + PAD;                       // halo record: [k | v]
    static constexpr int QS = SEGB + PAD;                            // tile record: q, later out / dq
    static constexpr int FS = SEG * 4 + 16;                          // fp32 record (dO)
};

// stage `seg_bytes` of every token of a (rows x cols) window whose top-left token is (r0, c0) (may lie outside the image:
// zero-filled) : smem record `rec` gets the bytes at dst_off.  src = first element of image b; tok_bytes = token stride.
// CH = chunk bytes: 16 when the record stride is a multiple of 16 (fp32 records), else 8 (bf16 records, stride = 8 mod 16)
// (rows, cols, segment bytes are compile-time: the index arithmetic of these loops -- divisions by run-time values in the
// first version -- cost as much as the attention arithmetic itself)
template <int CH, int SEG_BYTES, int ROWS, int COLS>
__device__ __forceinline__ void stage_window(unsigned char *smem, int rec_stride, int dst_off, const unsigned char *src,
                                             long long tok_bytes, int seg_off, int r0, int c0, int H, int W) {
    using V = typename std::conditional<CH == 16, uint4, uint2>::type;
    constexpr int nch = SEG_BYTES / CH, rows = ROWS, cols = COLS;
#pragma unroll 4
    for (int idx = threadIdx.x; idx < rows * cols * nch; idx += kTile) {
        const int ch = idx % nch, t = idx / nch;
        const int rr = r0 + t / cols, cc = c0 + t % cols;
        V v{};
        if (rr >= 0 && rr < H && cc >= 0 && cc < W)
            v = __ldg(reinterpret_cast<const V *>(src + ((long long)rr * W + cc) * tok_bytes + seg_off + CH * ch));
        *reinterpret_cast<V *>(smem + t * rec_stride + dst_off + CH * ch) = v;
    }
}
// the reverse: tile records -> global (only tokens inside the image)
template <int CH, int SEG_BYTES>
__device__ __forceinline__ void unstage_tile(const unsigned char *smem, int rec_stride, int src_off, unsigned char *dst,
                                             long long tok_bytes, int seg_off, int r0, int c0, int H, int W) {
    using V = typename std::conditional<CH == 16, uint4, uint2>::type;
    constexpr int nch = SEG_BYTES / CH;
#pragma unroll 4
    for (int idx = threadIdx.x; idx < kTile * nch; idx += kTile) {
        const int ch = idx % nch, t = idx / nch;
        const int rr = r0 + t / kTC, cc = c0 + t % kTC;
        if (rr < H && cc < W)
            *reinterpret_cast<V *>(dst + ((long long)rr * W + cc) * tok_bytes + seg_off + CH * ch) =
                *reinterpret_cast<const V *>(smem + t * rec_stride + src_off + CH * ch);
    }
}

// forward core on the staged halo: same arithmetic as local_core
template <typename T, int HD>
__device__ __forceinline__ float local_core_s(const LocalAttnParams &p, const unsigned char *kv, int tr, int tc,
                                              const bool (&ok)[9], const float (&qv)[2][HD], float (&A)[2][9],
                                              float (&abar)[9], float (&o)[2 * HD]) {
    using L = LT<T, HD>;
    float lg[2][9];
#pragma unroll
    for (int pp = 0; pp < 9; ++pp) {
        lg[0][pp] = lg[1][pp] = -INFINITY;
        if (ok[pp]) {
            float kk[2 * HD];
            ldv_s<T, 2 * HD>(kv + ((tr + pp / 3) * kHC + tc + pp % 3) * L::KVS, kk);
            float d0 = 0.f, d1 = 0.f;
#pragma unroll
            for (int c = 0; c < HD; ++c) {
                d0 = fmaf(qv[0][c], kk[c], d0);
                d1 = fmaf(qv[1][c], kk[HD + c], d1);
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
            float vv[2 * HD];
            ldv_s<T, 2 * HD>(kv + ((tr + pp / 3) * kHC + tc + pp % 3) * L::KVS + L::SEGB, vv);
#pragma unroll
            for (int c = 0; c < 2 * HD; ++c) o[c] = fmaf(abar[pp], vv[c], o[c]);
        }
    }
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < 2 * HD; ++c) ss = fmaf(o[c], o[c], ss);
    return 1.f / sqrtf(ss * (1.f / (2 * HD)) + p.eps);
}

struct TileIdx {
    int b, m, r0, c0, tr, tc, hr, wc;
    bool in;
    long long tok;        // global token index (b * N + hr * W + wc), valid when `in`
};
__device__ __forceinline__ TileIdx tile_index(const LocalAttnParams &p) {
    TileIdx t;
    const int tcols = (p.W + kTC - 1) / kTC;
    t.b = blockIdx.z; t.m = blockIdx.y;
    t.r0 = (blockIdx.x / tcols) * kTR; t.c0 = (blockIdx.x % tcols) * kTC;
    t.tr = threadIdx.x / kTC; t.tc = threadIdx.x % kTC;
    t.hr = t.r0 + t.tr; t.wc = t.c0 + t.tc;
    t.in = t.hr < p.H && t.wc < p.W;
    t.tok = ((long long)t.b * p.H + t.hr) * p.W + t.wc;
    return t;
}
__device__ __forceinline__ void window_mask(const LocalAttnParams &p, const TileIdx &t, bool (&ok)[9]) {
#pragma unroll
    for (int pp = 0; pp < 9; ++pp) {
        const int rr = t.hr + pp / 3 - 1, cc = t.wc + pp % 3 - 1;
        ok[pp] = t.in && rr >= 0 && rr < p.H && cc >= 0 && cc < p.W;
    }
}

template <typename T, int HD>
__global__ void __launch_bounds__(kTile) local_attn_fwd_tiled_kernel(const LocalAttnParams p) {
    using L = LT<T, HD>;
    extern __shared__ __align__(16) unsigned char lsm[];
    unsigned char *kv = lsm, *qs = lsm + kHalo * L::KVS;
    const TileIdx t = tile_index(p);
    const long long img = (long long)t.b * p.H * p.W;
    const unsigned char *kb = reinterpret_cast<const unsigned char *>(static_cast<const T *>(p.k) + img * p.ldkv);
    const unsigned char *vb = reinterpret_cast<const unsigned char *>(static_cast<const T *>(p.v) + img * p.ldkv);
    const unsigned char *qb = reinterpret_cast<const unsigned char *>(static_cast<const T *>(p.q) + img * p.ldq);
    const int so = t.m * L::SEGB;
    stage_window<L::CH, L::SEGB, kTR + 2, kHC>(kv, L::KVS, 0, kb, p.ldkv * (long long)sizeof(T), so, t.r0 - 1, t.c0 - 1, p.H, p.W);
    stage_window<L::CH, L::SEGB, kTR + 2, kHC>(kv, L::KVS, L::SEGB, vb, p.ldkv * (long long)sizeof(T), so, t.r0 - 1, t.c0 - 1, p.H, p.W);
    stage_window<L::CH, L::SEGB, kTR, kTC>(qs, L::QS, 0, qb, p.ldq * (long long)sizeof(T), so, t.r0, t.c0, p.H, p.W);
    __syncthreads();
    bool ok[9];
    window_mask(p, t, ok);
    float o[2 * HD];
    if (t.in) {
        float qv[2][HD];
        ldv_s<T, HD>(qs + threadIdx.x * L::QS, qv[0]);
        ldv_s<T, HD>(qs + threadIdx.x * L::QS + HD * sizeof(T), qv[1]);
        float A[2][9], abar[9];
        const float r = local_core_s<T, HD>(p, kv, t.tr, t.tc, ok, qv, A, abar, o);
#pragma unroll
        for (int c = 0; c < 2 * HD; ++c) o[c] = o[c] * r * __ldg(p.subln_w + c) * p.post;
        stv_s<T, 2 * HD>(qs + threadIdx.x * L::QS, o);          // each thread overwrites its OWN q record
    }
    __syncthreads();
    unstage_tile<L::CH, L::SEGB>(qs, L::QS, 0, reinterpret_cast<unsigned char *>(static_cast<T *>(p.out) + img * p.ldo), p.ldo * (long long)sizeof(T), so, t.r0, t.c0, p.H, p.W);
}

template <typename T, int HD>
__global__ void __launch_bounds__(kTile) local_attn_bwd_p1_tiled_kernel(const LocalAttnParams p) {
    using L = LT<T, HD>;
    extern __shared__ __align__(16) unsigned char lsm[];
    __shared__ float red[2 * HD + 1];
    unsigned char *kv = lsm, *qs = lsm + kHalo * L::KVS, *gs = qs + kTile * L::QS;   // q -> dq ; dout
    unsigned char *fs = gs + kTile * L::QS;                                          // fp32 staging: dO, then abar | dlog
    for (int i = threadIdx.x; i < 2 * HD + 1; i += blockDim.x) red[i] = 0.f;
    const TileIdx t = tile_index(p);
    const long long img = (long long)t.b * p.H * p.W;
    const int so = t.m * L::SEGB;
    stage_window<L::CH, L::SEGB, kTR + 2, kHC>(kv, L::KVS, 0, reinterpret_cast<const unsigned char *>(static_cast<const T *>(p.k) + img * p.ldkv), p.ldkv * (long long)sizeof(T), so, t.r0 - 1, t.c0 - 1, p.H, p.W);
    stage_window<L::CH, L::SEGB, kTR + 2, kHC>(kv, L::KVS, L::SEGB, reinterpret_cast<const unsigned char *>(static_cast<const T *>(p.v) + img * p.ldkv), p.ldkv * (long long)sizeof(T), so, t.r0 - 1, t.c0 - 1, p.H, p.W);
    stage_window<L::CH, L::SEGB, kTR, kTC>(qs, L::QS, 0, reinterpret_cast<const unsigned char *>(static_cast<const T *>(p.q) + img * p.ldq), p.ldq * (long long)sizeof(T), so, t.r0, t.c0, p.H, p.W);
    stage_window<L::CH, L::SEGB, kTR, kTC>(gs, L::QS, 0, reinterpret_cast<const unsigned char *>(static_cast<const T *>(p.dout) + img * p.lddo), p.lddo * (long long)sizeof(T), so, t.r0, t.c0, p.H, p.W);
    __syncthreads();
    bool ok[9];
    window_mask(p, t, ok);
    float dlam = 0.f;
    float dw[2 * HD];
#pragma unroll
    for (int c = 0; c < 2 * HD; ++c) dw[c] = 0.f;
    float wab[9], wdl[18];
#pragma unroll
    for (int pp = 0; pp < 9; ++pp) wab[pp] = wdl[pp] = wdl[9 + pp] = 0.f;
    if (t.in) {
        float qv[2][HD];
        ldv_s<T, HD>(qs + threadIdx.x * L::QS, qv[0]);
        ldv_s<T, HD>(qs + threadIdx.x * L::QS + HD * sizeof(T), qv[1]);
        float A[2][9], abar[9], o[2 * HD];
        const float r = local_core_s<T, HD>(p, kv, t.tr, t.tc, ok, qv, A, abar, o);
        float g[2 * HD];
        ldv_s<T, 2 * HD>(gs + threadIdx.x * L::QS, g);
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
        stv_s<float, 2 * HD>(fs + threadIdx.x * L::FS, dO);
        float dab[9];
#pragma unroll
        for (int pp = 0; pp < 9; ++pp) {
            dab[pp] = 0.f;
            if (ok[pp]) {
                float vv[2 * HD];
                ldv_s<T, 2 * HD>(kv + ((t.tr + pp / 3) * kHC + t.tc + pp % 3) * L::KVS + L::SEGB, vv);
                float d = 0.f;
#pragma unroll
                for (int c = 0; c < 2 * HD; ++c) d = fmaf(dO[c], vv[c], d);
                dab[pp] = d;
            }
        }
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int pp = 0; pp < 9; ++pp) {
            s0 = fmaf(A[0][pp], dab[pp], s0);
            s1 = fmaf(A[1][pp], dab[pp], s1);
        }
        dlam = -s1;
        const float lam = __ldg(p.lamp);
        float dq[2 * HD];
#pragma unroll
        for (int c = 0; c < 2 * HD; ++c) dq[c] = 0.f;
#pragma unroll
        for (int pp = 0; pp < 9; ++pp) {
            const float a0 = A[0][pp] * (dab[pp] - s0) * p.scale, a1 = -lam * A[1][pp] * (dab[pp] - s1) * p.scale;
            wab[pp] = abar[pp]; wdl[pp] = a0; wdl[9 + pp] = a1;
            if (ok[pp]) {
                float kk[2 * HD];
                ldv_s<T, 2 * HD>(kv + ((t.tr + pp / 3) * kHC + t.tc + pp % 3) * L::KVS, kk);
#pragma unroll
                for (int c = 0; c < HD; ++c) {
                    dq[c] = fmaf(a0, kk[c], dq[c]);
                    dq[HD + c] = fmaf(a1, kk[HD + c], dq[HD + c]);
                }
            }
        }
        stv_s<T, 2 * HD>(qs + threadIdx.x * L::QS, dq);          // own record: q -> dq
    }
    __syncthreads();
    // per-token results out: dq, dO (fp32), then abar | dlog through the same fp32 staging area
    unstage_tile<L::CH, L::SEGB>(qs, L::QS, 0, reinterpret_cast<unsigned char *>(static_cast<T *>(p.dq) + img * p.lddq), p.lddq * (long long)sizeof(T), so, t.r0, t.c0, p.H, p.W);
    unstage_tile<16, 2 * HD * 4>(fs, L::FS, 0, reinterpret_cast<unsigned char *>(p.ws_dO + img * p.h * 2 * HD), (long long)p.h * 2 * HD * 4, t.m * 2 * HD * 4, t.r0, t.c0, p.H, p.W);
    if (t.in) {                                                   // 27 floats per token: contiguous per token, direct
        float *wa = p.ws_abar + (t.tok * p.h + t.m) * 9;
        float *wl = p.ws_dlog + (t.tok * p.h + t.m) * 18;
#pragma unroll
        for (int pp = 0; pp < 9; ++pp) {
            wa[pp] = wab[pp];
            wl[pp] = wdl[pp];
            wl[9 + pp] = wdl[9 + pp];
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

template <typename T, int HD>
__global__ void __launch_bounds__(kTile) local_attn_bwd_p2_tiled_kernel(const LocalAttnParams p) {
    using L = LT<T, HD>;
    extern __shared__ __align__(16) unsigned char lsm[];
    constexpr int DOQ = 2 * HD * 4 + L::SEGB + 16;      // halo record: [dO fp32 | q], bank-spread
    unsigned char *dq = lsm;                                   // [kHalo][DOQ]
    float *sc = reinterpret_cast<float *>(lsm + kHalo * DOQ);  // [27 (+1)][kHalo]: abar 9 | dlog 18, token index fastest
    unsigned char *os = reinterpret_cast<unsigned char *>(sc + kHalo * 28);   // [kTile][2 * SEGB + PAD]: dk | dv
    const TileIdx t = tile_index(p);
    const long long img = (long long)t.b * p.H * p.W;
    const int so = t.m * L::SEGB;
    stage_window<16, 2 * HD * 4, kTR + 2, kHC>(dq, DOQ, 0, reinterpret_cast<const unsigned char *>(p.ws_dO + img * p.h * 2 * HD), (long long)p.h * 2 * HD * 4, t.m * 2 * HD * 4, t.r0 - 1, t.c0 - 1, p.H, p.W);
    stage_window<16, L::SEGB, kTR + 2, kHC>(dq, DOQ, 2 * HD * 4, reinterpret_cast<const unsigned char *>(static_cast<const T *>(p.q) + img * p.ldq), p.ldq * (long long)sizeof(T), so, t.r0 - 1, t.c0 - 1, p.H, p.W);
    for (int idx = threadIdx.x; idx < kHalo * 27; idx += blockDim.x) {
        const int j = idx % 27, ht = idx / 27;
        const int rr = t.r0 - 1 + ht / kHC, cc = t.c0 - 1 + ht % kHC;
        float v = 0.f;
        if (rr >= 0 && rr < p.H && cc >= 0 && cc < p.W) {
            const long long n = (img + (long long)rr * p.W + cc) * p.h + t.m;
            v = j < 9 ? __ldg(p.ws_abar + n * 9 + j) : __ldg(p.ws_dlog + n * 18 + (j - 9));
        }
        sc[j * kHalo + ht] = v;
    }
    __syncthreads();
    if (t.in) {
        float dk[2 * HD], dv[2 * HD];
#pragma unroll
        for (int c = 0; c < 2 * HD; ++c) dk[c] = dv[c] = 0.f;
#pragma unroll
        for (int pp = 0; pp < 9; ++pp) {
            // token n = tok - off_p has this token as its pp-th neighbour
            const int rr = t.hr - (pp / 3 - 1), cc = t.wc - (pp % 3 - 1);
            if (rr >= 0 && rr < p.H && cc >= 0 && cc < p.W) {
                const int ht = (t.tr + 1 - (pp / 3 - 1)) * kHC + (t.tc + 1 - (pp % 3 - 1));
                const float ab = sc[pp * kHalo + ht], l0 = sc[(9 + pp) * kHalo + ht], l1 = sc[(18 + pp) * kHalo + ht];
                float dO[2 * HD], qn[2 * HD];
                ldv_s<float, 2 * HD>(dq + ht * DOQ, dO);
                ldv_s<T, 2 * HD>(dq + ht * DOQ + 2 * HD * 4, qn);
#pragma unroll
                for (int c = 0; c < 2 * HD; ++c) dv[c] = fmaf(ab, dO[c], dv[c]);
#pragma unroll
                for (int c = 0; c < HD; ++c) {
                    dk[c] = fmaf(l0, qn[c], dk[c]);
                    dk[HD + c] = fmaf(l1, qn[HD + c], dk[HD + c]);
                }
            }
        }
        stv_s<T, 2 * HD>(os + threadIdx.x * L::KVS, dk);
        stv_s<T, 2 * HD>(os + threadIdx.x * L::KVS + L::SEGB, dv);
    }
    __syncthreads();
    unstage_tile<L::CH, L::SEGB>(os, L::KVS, 0, reinterpret_cast<unsigned char *>(static_cast<T *>(p.dk) + img * p.lddkv), p.lddkv * (long long)sizeof(T), so, t.r0, t.c0, p.H, p.W);
    unstage_tile<L::CH, L::SEGB>(os, L::KVS, L::SEGB, reinterpret_cast<unsigned char *>(static_cast<T *>(p.dv) + img * p.lddkv), p.lddkv * (long long)sizeof(T), so, t.r0, t.c0, p.H, p.W);
}

template <typename T, int HD>
static bool local_tiled_ok(const LocalAttnParams &p, int which) {
    using L = LT<T, HD>;
    if (L::SEGB % 16 != 0 || HD % 4 != 0 || getenv("MLAGG_LOCAL_UNTILED")) return false;
    // Measured on B200 (tools/local_attn_microbench.py, profiles/local_attn_microbench_r02.txt): the tiled FORWARD is 2x
    // faster on the large maps (160^2: 152 -> 76 us, 80^2: 78 -> 44 us) and slower on the small ones (tile quantisation:
    // a 20 x 20 map fills 52 % of its 8 x 16 tiles); the tiled BACKWARD passes execute 40 % more instructions (staging)
    // from a fully unrolled body that no longer fits the instruction cache (ncu: stall_no_instruction 2.0 per issue) and
    // are not faster anywhere -- they stay available behind MLAGG_LOCAL_TILED_BWD=1 for the next round's work.
    if (which == 0 ? (p.H < 64 || p.W < 64) : getenv("MLAGG_LOCAL_TILED_BWD") == nullptr) return false;
    auto a16 = [](const void *q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    const long long e = sizeof(T);
    bool ok = a16(p.q) && a16(p.k) && a16(p.v) && (p.ldq * e) % 16 == 0 && (p.ldkv * e) % 16 == 0 && p.Bn <= 65535 &&
              p.h <= 65535;
    if (which == 0) ok = ok && a16(p.out) && (p.ldo * e) % 16 == 0;
    if (which == 1) ok = ok && a16(p.dout) && a16(p.dq) && a16(p.ws_dO) && (p.lddo * e) % 16 == 0 && (p.lddq * e) % 16 == 0;
    if (which == 2) ok = ok && a16(p.dk) && a16(p.dv) && a16(p.ws_dO) && (p.lddkv * e) % 16 == 0;
    return ok;
}

template <typename T, int HD>
static cudaError_t local_launch_tiled(const LocalAttnParams &p, int which, cudaStream_t st) {
    using L = LT<T, HD>;
    const dim3 grid(((p.H + kTR - 1) / kTR) * ((p.W + kTC - 1) / kTC), p.h, p.Bn);
    cudaError_t e;
    if (which == 0) {
        const size_t sm = (size_t)kHalo * L::KVS + (size_t)kTile * L::QS;
        auto k = local_attn_fwd_tiled_kernel<T, HD>;
        if ((e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)) != cudaSuccess) return e;
        k<<<grid, kTile, sm, st>>>(p);
    } else if (which == 1) {
        const size_t sm = (size_t)kHalo * L::KVS + 2 * (size_t)kTile * L::QS + (size_t)kTile * L::FS;
        auto k = local_attn_bwd_p1_tiled_kernel<T, HD>;
        if ((e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)) != cudaSuccess) return e;
        k<<<grid, kTile, sm, st>>>(p);
    } else {
        constexpr int DOQ = 2 * HD * 4 + L::SEGB + 16;
        const size_t sm = (size_t)kHalo * DOQ + (size_t)kHalo * 28 * 4 + (size_t)kTile * L::KVS;
        auto k = local_attn_bwd_p2_tiled_kernel<T, HD>;
        if ((e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)) != cudaSuccess) return e;
        k<<<grid, kTile, sm, st>>>(p);
    }
    return cudaGetLastError();
}

template <typename T, int HD>
static cudaError_t local_launch(const LocalAttnParams &p, int which, cudaStream_t st) {
    if constexpr ((2 * HD * sizeof(T)) % 16 == 0 && HD % 4 == 0) {
        if (local_tiled_ok<T, HD>(p, which)) return local_launch_tiled<T, HD>(p, which, st);
    }
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
