// linattn.cu -- elu+1 linear attention core of MLLA (the ops BASELINE.json:north_star names by op).
// Replaces the op sequence at reference nnUNetTrainer_MLLA_UNet.py:234-246 (elu+1 on q and k, RoPE.forward :190-195 on
// both, z = 1/(q . mean_n k + 1e-6), kv = (k_rope^T n^-1/2)(v n^-1/2), out = q_rope kv z) with two kernels per direction:
//
//   forward   linattn_state_kernel   S[b,h] = (1/N) sum_n rope(phi k_n) (x) v_n,  kmean[b,h] = (1/N) sum_n phi k_n
//             linattn_apply_kernel   out_n  = z_n * rope(phi q_n) S,             z_n = 1 / (phi q_n . kmean + eps)
//   backward  linattn_bwd_q_kernel   dq, and dS = (1/N) sum_n rope(phi q_n) (x) z_n dO_n, dkmean (token-parallel + reduce)
//             linattn_bwd_kv_kernel  dk, dv from dS, dkmean (token-parallel)
//
// Layout: q, k, v, out are tokens-major (B, N, h, hd) views with arbitrary row strides (q and k are the two halves of
// the qk projection, consumed in place); S (B, h, hd, hd) and kmean (B, h, hd) are fp32.  phi, RoPE and every
// accumulation are fp32 (the reference forces fp32 in RoPE.forward); I/O is fp32 or bf16.
// The hd x hd state contraction is a rank-N update with tiny M = N = hd (32 in MLLA-UNet): a register-tiled FFMA
// outer-product loop fed from shared-memory tiles keeps it below the HBM time of reading k and v once for fp32 parity
// (1e-4) without split-precision tricks; tcgen05's minimum M = 64 tile would be half empty and bf16/tf32 operands
// would break the fp32 tolerance.
// RoPE angles come from a separable table the caller builds exactly like the reference does (:181-187):
//   rope_cs (H + W, C/4, 2) = [cos, sin] of row * theta_i (first H rows) and col * theta_i (next W rows).
#include <cuda_bf16.h>

#include "common.cuh"

namespace mlagg {

struct LinAttnParams {
    const void *q, *k, *v, *dout;
    void *out, *dq, *dk, *dv;
    float *S, *kmean;    // forward results, saved for backward
    float *dS, *dkm;     // backward scratch, zeroed by the entry point
    const float *rope;   // (H + W, C/4, 2)
    long long ldq, ldk, ldv, ldo, lddo, lddq, lddk, lddv;
    int Bn, H, W, h, chunk;
    float eps;
};

constexpr int kLTok = 128;  // tokens per tile == threads per block

// ---------------------------------------------------------------- row I/O: HD contiguous elements <-> fp32 registers
template <int HD>
__device__ __forceinline__ void la_load(const float *p, float (&r)[HD]) {
#pragma unroll
    for (int i = 0; i < HD / 4; ++i) {
        const float4 t = __ldg(reinterpret_cast<const float4 *>(p) + i);
        r[4 * i] = t.x, r[4 * i + 1] = t.y, r[4 * i + 2] = t.z, r[4 * i + 3] = t.w;
    }
}
template <int HD>
__device__ __forceinline__ void la_load(const __nv_bfloat16 *p, float (&r)[HD]) {
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
        const uint4 t = __ldg(reinterpret_cast<const uint4 *>(p) + i);
        const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            r[8 * i + 2 * j] = __uint_as_float(w[j] << 16);
            r[8 * i + 2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
        }
    }
}
template <int HD>
__device__ __forceinline__ void la_store(float *p, const float (&r)[HD]) {
#pragma unroll
    for (int i = 0; i < HD / 4; ++i)
        reinterpret_cast<float4 *>(p)[i] = make_float4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
}
template <int HD>
__device__ __forceinline__ void la_store(__nv_bfloat16 *p, const float (&r)[HD]) {
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const __nv_bfloat162 t = __floats2bfloat162_rn(r[8 * i + 2 * j], r[8 * i + 2 * j + 1]);
            w[j] = *reinterpret_cast<const uint32_t *>(&t);
        }
        reinterpret_cast<uint4 *>(p)[i] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}
template <int HD>
__device__ __forceinline__ void la_smem_row(float *row, const float (&r)[HD]) {
#pragma unroll
    for (int i = 0; i < HD / 4; ++i)
        reinterpret_cast<float4 *>(row)[i] = make_float4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
}

// phi(t) = elu(t) + 1 = t + 1 (t > 0) | exp(t) (t <= 0).  d phi / dt = 1 | phi, i.e. (phi > 1 ? 1 : phi).
__device__ __forceinline__ float la_phi(float t) { return t > 0.f ? t + 1.f : ex2_approx(t * kLog2e); }

// cos/sin of the HD/2 channel pairs of head hh at token (row, col)
template <int HD>
__device__ __forceinline__ void la_angles(const LinAttnParams &p, int hh, int n, float (&cs)[HD]) {
    const int quarter = p.h * HD / 4;
    const int row = n / p.W, col = n - row * p.W;
    const float2 *tr = reinterpret_cast<const float2 *>(p.rope) + (long long)row * quarter;
    const float2 *tc = reinterpret_cast<const float2 *>(p.rope) + (long long)(p.H + col) * quarter - quarter;
#pragma unroll
    for (int j = 0; j < HD / 2; ++j) {
        const int i = hh * (HD / 2) + j;
        const float2 t = __ldg(i < quarter ? tr + i : tc + i);
        cs[2 * j] = t.x, cs[2 * j + 1] = t.y;
    }
}
// (x0, x1) -> (c x0 - s x1, s x0 + c x1)
template <int HD>
__device__ __forceinline__ void la_rope(const float (&x)[HD], const float (&cs)[HD], float (&y)[HD]) {
#pragma unroll
    for (int j = 0; j < HD / 2; ++j) {
        const float c = cs[2 * j], s = cs[2 * j + 1], a = x[2 * j], b = x[2 * j + 1];
        y[2 * j] = c * a - s * b, y[2 * j + 1] = s * a + c * b;
    }
}
// transpose of the rotation: (g0, g1) -> (c g0 + s g1, -s g0 + c g1)
template <int HD>
__device__ __forceinline__ void la_rope_t(const float (&g)[HD], const float (&cs)[HD], float (&y)[HD]) {
#pragma unroll
    for (int j = 0; j < HD / 2; ++j) {
        const float c = cs[2 * j], s = cs[2 * j + 1], a = g[2 * j], b = g[2 * j + 1];
        y[2 * j] = c * a + s * b, y[2 * j + 1] = c * b - s * a;
    }
}

// ---------------------------------------------------------------- rank-128 update of an HD x HD accumulator
// Thread (g, di, ei) owns the RD x 4 block (RD*di.., 4*ei..) of group g's partial sum and walks the tile rows
// g, g + NG, ...; tiles are [kLTok][HD + 4] fp32 (the +4 keeps the row-per-thread float4 stores conflict-free).
template <int HD>
struct LaTile {
    static constexpr int LD = HD + 4;
    static constexpr int RD = HD == 64 ? 8 : 4;
    static constexpr int TPS = (HD / RD) * (HD / 4);  // threads per accumulator copy
    static constexpr int NG = kLTok / TPS;            // copies (token groups)
    static_assert(TPS <= kLTok && kLTok % TPS == 0, "tile shape");
};

template <int HD>
__device__ __forceinline__ void la_outer_accum(const float *sA, const float *sB, float (&acc)[LaTile<HD>::RD][4]) {
    using TL = LaTile<HD>;
    const int ts = threadIdx.x % TL::TPS, g = threadIdx.x / TL::TPS;
    const int di = ts / (HD / 4), ei = ts % (HD / 4);
#pragma unroll 4
    for (int r = g; r < kLTok; r += TL::NG) {
        const float4 b = *reinterpret_cast<const float4 *>(sB + r * TL::LD + 4 * ei);
#pragma unroll
        for (int x = 0; x < TL::RD / 4; ++x) {
            const float4 a = *reinterpret_cast<const float4 *>(sA + r * TL::LD + TL::RD * di + 4 * x);
            const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int y = 0; y < 4; ++y) {
                acc[4 * x + y][0] += av[y] * b.x;
                acc[4 * x + y][1] += av[y] * b.y;
                acc[4 * x + y][2] += av[y] * b.z;
                acc[4 * x + y][3] += av[y] * b.w;
            }
        }
    }
}
// partial column sums of a tile: thread owns column tid % HD over rows tid / HD, + kLTok / HD, ...
template <int HD>
__device__ __forceinline__ float la_colsum(const float *sW) {
    float s = 0.f;
    for (int r = threadIdx.x / HD; r < kLTok; r += kLTok / HD) s += sW[r * LaTile<HD>::LD + threadIdx.x % HD];
    return s;
}
// fold the per-thread partials through shared memory and add them, scaled, to the global accumulators
template <int HD>
__device__ __forceinline__ void la_flush(float *sRed, const float (&acc)[LaTile<HD>::RD][4], float colpart, float scale,
                                         float *gS, float *gvec) {
    using TL = LaTile<HD>;
    for (int i = threadIdx.x; i < HD * HD + HD; i += kLTok) sRed[i] = 0.f;
    __syncthreads();
    const int ts = threadIdx.x % TL::TPS;
    const int di = ts / (HD / 4), ei = ts % (HD / 4);
#pragma unroll
    for (int x = 0; x < TL::RD; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) atomicAdd(sRed + (TL::RD * di + x) * HD + 4 * ei + y, acc[x][y]);
    atomicAdd(sRed + HD * HD + threadIdx.x % HD, colpart);
    __syncthreads();
    for (int i = threadIdx.x; i < HD * HD; i += kLTok) atomicAdd(gS + i, sRed[i] * scale);
    if (threadIdx.x < HD) atomicAdd(gvec + threadIdx.x, sRed[HD * HD + threadIdx.x] * scale);
}

// ---------------------------------------------------------------- forward 1: state
template <typename T, int HD>
__global__ void __launch_bounds__(kLTok) linattn_state_kernel(const LinAttnParams p) {
    using TL = LaTile<HD>;
    extern __shared__ __align__(16) float smem[];
    float *sA = smem, *sB = sA + kLTok * TL::LD, *sW = sB + kLTok * TL::LD;
    const int b = blockIdx.z, hh = blockIdx.y, N = p.H * p.W;
    const int n0 = blockIdx.x * p.chunk, n1 = min(N, n0 + p.chunk);
    float acc[TL::RD][4] = {};
    float colpart = 0.f;
    for (int t0 = n0; t0 < n1; t0 += kLTok) {
        const int n = t0 + threadIdx.x;
        float x[HD], y[HD];
        if (n < n1) {
            const long long tok = (long long)b * N + n;
            la_load<HD>(static_cast<const T *>(p.k) + tok * p.ldk + hh * HD, x);
#pragma unroll
            for (int c = 0; c < HD; ++c) x[c] = la_phi(x[c]);
            la_smem_row<HD>(sW + threadIdx.x * TL::LD, x);
            float cs[HD];
            la_angles<HD>(p, hh, n, cs);
            la_rope<HD>(x, cs, y);
            la_smem_row<HD>(sA + threadIdx.x * TL::LD, y);
            la_load<HD>(static_cast<const T *>(p.v) + tok * p.ldv + hh * HD, x);
            la_smem_row<HD>(sB + threadIdx.x * TL::LD, x);
        } else {
#pragma unroll
            for (int c = 0; c < HD; ++c) x[c] = 0.f;
            la_smem_row<HD>(sW + threadIdx.x * TL::LD, x);
            la_smem_row<HD>(sA + threadIdx.x * TL::LD, x);
            la_smem_row<HD>(sB + threadIdx.x * TL::LD, x);
        }
        __syncthreads();
        la_outer_accum<HD>(sA, sB, acc);
        colpart += la_colsum<HD>(sW);
        __syncthreads();
    }
    const long long bh = (long long)b * p.h + hh;
    la_flush<HD>(smem, acc, colpart, 1.f / (float)N, p.S + bh * HD * HD, p.kmean + bh * HD);
}

// stage S (or dS) and kmean (or dkmean) of one (batch, head) in shared memory
template <int HD>
__device__ __forceinline__ void la_stage_state(const float *gS, const float *gvec, float *sS, float *sV) {
    for (int i = threadIdx.x; i < HD * HD; i += kLTok) sS[i] = gS[i];
    if (threadIdx.x < HD) sV[threadIdx.x] = gvec[threadIdx.x];
    __syncthreads();
}
// y[e] = sum_d x[d] M[d][e]      (M row-major HD x HD in shared memory, read as warp-wide broadcasts)
template <int HD>
__device__ __forceinline__ void la_vec_mat(const float (&x)[HD], const float *sM, float (&y)[HD]) {
#pragma unroll
    for (int e = 0; e < HD; ++e) y[e] = 0.f;
#pragma unroll
    for (int d = 0; d < HD; ++d) {
#pragma unroll
        for (int e = 0; e < HD / 4; ++e) {
            const float4 m = *reinterpret_cast<const float4 *>(sM + d * HD + 4 * e);
            y[4 * e] += x[d] * m.x, y[4 * e + 1] += x[d] * m.y, y[4 * e + 2] += x[d] * m.z, y[4 * e + 3] += x[d] * m.w;
        }
    }
}
// y[d] = sum_e M[d][e] x[e]
template <int HD>
__device__ __forceinline__ void la_mat_vec(const float *sM, const float (&x)[HD], float (&y)[HD]) {
#pragma unroll
    for (int d = 0; d < HD; ++d) {
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int e = 0; e < HD / 4; ++e) {
            const float4 m = *reinterpret_cast<const float4 *>(sM + d * HD + 4 * e);
            s0 += m.x * x[4 * e] + m.z * x[4 * e + 2];
            s1 += m.y * x[4 * e + 1] + m.w * x[4 * e + 3];
        }
        y[d] = s0 + s1;
    }
}

// ---------------------------------------------------------------- forward 2: apply
template <typename T, int HD>
__global__ void __launch_bounds__(kLTok) linattn_apply_kernel(const LinAttnParams p) {
    __shared__ __align__(16) float sS[HD * HD];
    __shared__ float sK[HD];
    const int b = blockIdx.z, hh = blockIdx.y, N = p.H * p.W;
    const long long bh = (long long)b * p.h + hh;
    la_stage_state<HD>(p.S + bh * HD * HD, p.kmean + bh * HD, sS, sK);
    const int n = blockIdx.x * kLTok + threadIdx.x;
    if (n >= N) return;
    const long long tok = (long long)b * N + n;
    float x[HD], qr[HD], o[HD];
    la_load<HD>(static_cast<const T *>(p.q) + tok * p.ldq + hh * HD, x);
    float den = p.eps;
#pragma unroll
    for (int c = 0; c < HD; ++c) {
        x[c] = la_phi(x[c]);
        den += x[c] * sK[c];
    }
    const float z = 1.f / den;
    {
        float cs[HD];
        la_angles<HD>(p, hh, n, cs);
        la_rope<HD>(x, cs, qr);
    }
    la_vec_mat<HD>(qr, sS, o);
#pragma unroll
    for (int c = 0; c < HD; ++c) o[c] *= z;
    la_store<HD>(static_cast<T *>(p.out) + tok * p.ldo + hh * HD, o);
}

// ---------------------------------------------------------------- backward 1: dq, dS, dkmean
template <typename T, int HD>
__global__ void __launch_bounds__(kLTok) linattn_bwd_q_kernel(const LinAttnParams p) {
    using TL = LaTile<HD>;
    extern __shared__ __align__(16) float smem[];
    float *sA = smem, *sB = sA + kLTok * TL::LD, *sW = sB + kLTok * TL::LD;
    float *sS = sW + kLTok * TL::LD, *sK = sS + HD * HD;
    const int b = blockIdx.z, hh = blockIdx.y, N = p.H * p.W;
    const long long bh = (long long)b * p.h + hh;
    const int n0 = blockIdx.x * p.chunk, n1 = min(N, n0 + p.chunk);
    la_stage_state<HD>(p.S + bh * HD * HD, p.kmean + bh * HD, sS, sK);
    float acc[TL::RD][4] = {};
    float colpart = 0.f;
    for (int t0 = n0; t0 < n1; t0 += kLTok) {
        const int n = t0 + threadIdx.x;
        if (n < n1) {
            const long long tok = (long long)b * N + n;
            float phi[HD], cs[HD], a[HD], g[HD];
            la_load<HD>(static_cast<const T *>(p.q) + tok * p.ldq + hh * HD, phi);
            float den = p.eps;
#pragma unroll
            for (int c = 0; c < HD; ++c) {
                phi[c] = la_phi(phi[c]);
                den += phi[c] * sK[c];
            }
            const float z = 1.f / den;
            la_angles<HD>(p, hh, n, cs);
            la_rope<HD>(phi, cs, a);                                   // a = q_rope
            la_smem_row<HD>(sA + threadIdx.x * TL::LD, a);
            la_vec_mat<HD>(a, sS, g);                                  // g = q_rope S  (out / z)
            la_load<HD>(static_cast<const T *>(p.dout) + tok * p.lddo + hh * HD, a);   // a = dO
            float dz = 0.f;
#pragma unroll
            for (int c = 0; c < HD; ++c) {
                dz += a[c] * g[c];
                a[c] *= z;                                             // a = dt = z dO
            }
            la_smem_row<HD>(sB + threadIdx.x * TL::LD, a);
            asm volatile("" ::: "memory");
            la_mat_vec<HD>(sS, a, g);                                  // g = d q_rope = S dt
            la_rope_t<HD>(g, cs, a);                                   // a = d phi (rotation part)
            const float cden = -z * z * dz;                            // d / d(phi q . kmean)
#pragma unroll
            for (int c = 0; c < HD; ++c) {
                a[c] = (a[c] + cden * sK[c]) * (phi[c] > 1.f ? 1.f : phi[c]);
                g[c] = cden * phi[c];
            }
            la_smem_row<HD>(sW + threadIdx.x * TL::LD, g);
            la_store<HD>(static_cast<T *>(p.dq) + tok * p.lddq + hh * HD, a);
        } else {
            float zr[HD];
#pragma unroll
            for (int c = 0; c < HD; ++c) zr[c] = 0.f;
            la_smem_row<HD>(sA + threadIdx.x * TL::LD, zr);
            la_smem_row<HD>(sB + threadIdx.x * TL::LD, zr);
            la_smem_row<HD>(sW + threadIdx.x * TL::LD, zr);
        }
        __syncthreads();
        la_outer_accum<HD>(sA, sB, acc);
        colpart += la_colsum<HD>(sW);
        __syncthreads();
    }
    la_flush<HD>(smem, acc, colpart, 1.f / (float)N, p.dS + bh * HD * HD, p.dkm + bh * HD);
}

// ---------------------------------------------------------------- backward 2: dk, dv
template <typename T, int HD>
__global__ void __launch_bounds__(kLTok) linattn_bwd_kv_kernel(const LinAttnParams p) {
    __shared__ __align__(16) float sS[HD * HD];
    __shared__ float sK[HD];
    const int b = blockIdx.z, hh = blockIdx.y, N = p.H * p.W;
    const long long bh = (long long)b * p.h + hh;
    la_stage_state<HD>(p.dS + bh * HD * HD, p.dkm + bh * HD, sS, sK);   // already scaled by 1/N
    const int n = blockIdx.x * kLTok + threadIdx.x;
    if (n >= N) return;
    const long long tok = (long long)b * N + n;
    float phi[HD], cs[HD], kr[HD], vv[HD], dv[HD];
    la_load<HD>(static_cast<const T *>(p.k) + tok * p.ldk + hh * HD, phi);
#pragma unroll
    for (int c = 0; c < HD; ++c) phi[c] = la_phi(phi[c]);
    la_angles<HD>(p, hh, n, cs);
    la_rope<HD>(phi, cs, kr);
    la_load<HD>(static_cast<const T *>(p.v) + tok * p.ldv + hh * HD, vv);
#pragma unroll
    for (int c = 0; c < HD; ++c) dv[c] = 0.f;
    // one pass over the rows of dS: dv += k_rope[d] * dS[d,:] and d k_rope[d] = dS[d,:] . v (kr[d] is overwritten)
#pragma unroll
    for (int d = 0; d < HD; ++d) {
        float s0 = 0.f, s1 = 0.f;
        const float kd = kr[d];
#pragma unroll
        for (int e = 0; e < HD / 4; ++e) {
            const float4 m = *reinterpret_cast<const float4 *>(sS + d * HD + 4 * e);
            dv[4 * e] += kd * m.x, dv[4 * e + 1] += kd * m.y, dv[4 * e + 2] += kd * m.z, dv[4 * e + 3] += kd * m.w;
            s0 += m.x * vv[4 * e] + m.z * vv[4 * e + 2];
            s1 += m.y * vv[4 * e + 1] + m.w * vv[4 * e + 3];
        }
        kr[d] = s0 + s1;
    }
    la_store<HD>(static_cast<T *>(p.dv) + tok * p.lddv + hh * HD, dv);
    la_rope_t<HD>(kr, cs, vv);
#pragma unroll
    for (int c = 0; c < HD; ++c) vv[c] = (vv[c] + sK[c]) * (phi[c] > 1.f ? 1.f : phi[c]);
    la_store<HD>(static_cast<T *>(p.dk) + tok * p.lddk + hh * HD, vv);
}

// ---------------------------------------------------------------- dispatch
bool linattn_hd_supported(int hd) { return hd == 8 || hd == 16 || hd == 32; }

int linattn_chunk(int Bn, int N, int h) {
    // ~4 reducing blocks per SM; a block walks whole 128-token tiles
    long long per = ((long long)Bn * N * h + 591) / 592;
    long long chunk = (per + kLTok - 1) / kLTok * kLTok;
    if (chunk < kLTok) chunk = kLTok;
    return (int)chunk;
}

template <typename T, int HD>
static cudaError_t linattn_launch(LinAttnParams p, int which, cudaStream_t st) {
    using TL = LaTile<HD>;
    const int N = p.H * p.W;
    p.chunk = linattn_chunk(p.Bn, N, p.h);
    const dim3 gtok((N + kLTok - 1) / kLTok, p.h, p.Bn), gred((N + p.chunk - 1) / p.chunk, p.h, p.Bn);
    const size_t tiles = (size_t)3 * kLTok * TL::LD * sizeof(float);
    if (which == 0) {
        auto k1 = linattn_state_kernel<T, HD>;
        cudaError_t e = cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tiles);
        if (e != cudaSuccess) return e;
        k1<<<gred, kLTok, tiles, st>>>(p);
        linattn_apply_kernel<T, HD><<<gtok, kLTok, 0, st>>>(p);
    } else {
        const size_t sm = tiles + (size_t)(HD * HD + HD) * sizeof(float);
        auto k1 = linattn_bwd_q_kernel<T, HD>;
        cudaError_t e = cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e != cudaSuccess) return e;
        k1<<<gred, kLTok, sm, st>>>(p);
        linattn_bwd_kv_kernel<T, HD><<<gtok, kLTok, 0, st>>>(p);
    }
    return cudaGetLastError();
}

template <typename T>
static cudaError_t linattn_by_hd(const LinAttnParams &p, int hd, int which, cudaStream_t st) {
    switch (hd) {
        case 8: return linattn_launch<T, 8>(p, which, st);
        case 16: return linattn_launch<T, 16>(p, which, st);
        case 32: return linattn_launch<T, 32>(p, which, st);
    }
    return cudaErrorInvalidValue;
}

cudaError_t linattn_dispatch(const LinAttnParams &p, int hd, int dtype, int which, cudaStream_t st) {
    return dtype == 0 ? linattn_by_hd<float>(p, hd, which, st) : linattn_by_hd<__nv_bfloat16>(p, hd, which, st);
}

}  // namespace mlagg
