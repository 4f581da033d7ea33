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

#include <cstdlib>

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

int linattn_chunk(int Bn, int N, int h);

// ================================================================= tensor-core variants (bf16 I/O, hd 16 / 32)
// The phi(K)^T V state contraction (k = tokens), the apply product (k = hd) and their backward counterparts as
// mma.sync.m16n8k16 bf16 with fp32 accumulators.  elu+1, RoPE, the normaliser and every reduction stay fp32; only the
// MMA operands are rounded to bf16 (the I/O type of this path).  RoPE rotates channel pairs (2i, 2i+1) -- exactly the
// register pairs of the A-operand / accumulator layouts -- so q, k rows are loaded, mapped and rotated directly in
// fragment layout without shared memory; contractions over tokens stage their operands transposed ([channel][token]).
__device__ __forceinline__ void la_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t la_pack(float lo, float hi) {
    const __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t *>(&t);
}
__device__ __forceinline__ float2 la_unpack(uint32_t v) {
    return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
}
__device__ __forceinline__ uint32_t la_lds32(const __nv_bfloat16 *p) { return *reinterpret_cast<const uint32_t *>(p); }

// 8x8 b16 tile held in fragment layout (lane (g, t): row g, columns 2t, 2t+1) -> its transpose in the same layout: turns a
// [token][channel] register tile into the [channel][token] operand a contraction over tokens needs, without shared memory
__device__ __forceinline__ uint32_t la_trans(uint32_t v) {
    uint32_t d;
    asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(v));
    return d;
}

constexpr int kLmTok = 256;      // tokens per block of the token-parallel mma kernels (4 warps x 4 tiles of 16)

// One token row pair set of a warp tile in fragment layout: rows (g, g+8), pairs pi = 0 .. HD/8-1 at columns 8*pi + 2t.
template <int HD>
struct LaFrag {
    float2 phi[2][HD / 8];   // elu+1 values
    float2 cs[2][HD / 8];    // (cos, sin) of the pair
};
template <int HD>
__device__ __forceinline__ void la_frag_load(const LinAttnParams &p, const __nv_bfloat16 *base, long long ld, int hh, int b,
                                             const int (&nr)[2], const bool (&ok)[2], int t, LaFrag<HD> &f) {
    const int N = p.H * p.W, quarter = p.h * HD / 4;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int n = ok[r] ? nr[r] : 0;
        const __nv_bfloat16 *row = base + ((long long)b * N + n) * ld + hh * HD;
        const int rr = n / p.W, cc = n - rr * p.W;
        const float2 *tr = reinterpret_cast<const float2 *>(p.rope) + (long long)rr * quarter;
        const float2 *tc = reinterpret_cast<const float2 *>(p.rope) + (long long)(p.H + cc) * quarter - quarter;
#pragma unroll
        for (int pi = 0; pi < HD / 8; ++pi) {
            const float2 x = ok[r] ? la_unpack(la_lds32(row + 8 * pi + 2 * t)) : make_float2(0.f, 0.f);
            f.phi[r][pi] = ok[r] ? make_float2(la_phi(x.x), la_phi(x.y)) : make_float2(0.f, 0.f);
            const int i = hh * (HD / 2) + 4 * pi + t;
            f.cs[r][pi] = __ldg(i < quarter ? tr + i : tc + i);
        }
    }
}
// rope(phi) packed as the A operands of the HD/16 k-steps
template <int HD>
__device__ __forceinline__ void la_frag_rope_a(const LaFrag<HD> &f, uint32_t (&a)[HD / 16][4]) {
#pragma unroll
    for (int ks = 0; ks < HD / 16; ++ks)
#pragma unroll
        for (int hf = 0; hf < 2; ++hf)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const float2 x = f.phi[r][2 * ks + hf], c = f.cs[r][2 * ks + hf];
                a[ks][2 * hf + r] = la_pack(c.x * x.x - c.y * x.y, c.y * x.x + c.x * x.y);
            }
}

// ---- forward 1: S = (1/N) sum rope(phi k) (x) v, kmean = (1/N) sum phi k        (k = tokens)
// A warp owns 16 tokens at a time: k and v rows are read in fragment layout, phi / RoPE happen in registers, and
// movmatrix turns the [token][channel] tiles into the [channel][token] A operand and the [token][channel] B operand --
// no shared-memory staging, no block-wide barriers inside the token loop.
template <int HD>
__global__ void __launch_bounds__(kLTok) linattn_state_mma_kernel(const LinAttnParams p) {
    __shared__ float red[HD * HD + HD];
    const int b = blockIdx.z, hh = blockIdx.y, N = p.H * p.W;
    const int n0 = blockIdx.x * p.chunk, n1 = min(N, n0 + p.chunk);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    float S[HD / 16][HD / 8][4], KM[HD / 16][4];
#pragma unroll
    for (int mt = 0; mt < HD / 16; ++mt) {
        KM[mt][0] = KM[mt][1] = KM[mt][2] = KM[mt][3] = 0.f;
#pragma unroll
        for (int nt = 0; nt < HD / 8; ++nt) S[mt][nt][0] = S[mt][nt][1] = S[mt][nt][2] = S[mt][nt][3] = 0.f;
    }
    const uint32_t ones = (g == 0) ? la_pack(1.f, 1.f) : 0u;     // B operand whose column 0 is all ones
    for (int nb = n0 + warp * 16; nb < n1; nb += 64) {
        const int nr[2] = {nb + g, nb + g + 8};
        const bool ok[2] = {nr[0] < n1, nr[1] < n1};
        LaFrag<HD> f;
        la_frag_load<HD>(p, static_cast<const __nv_bfloat16 *>(p.k), p.ldk, hh, b, nr, ok, t, f);
        uint32_t kt[2][HD / 8], pt[2][HD / 8], vt[2][HD / 8];     // transposed 8x8 tiles: [token half][channel group]
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int pi = 0; pi < HD / 8; ++pi) {
                const float2 x = f.phi[r][pi], c = f.cs[r][pi];
                kt[r][pi] = la_trans(la_pack(c.x * x.x - c.y * x.y, c.y * x.x + c.x * x.y));
                pt[r][pi] = la_trans(la_pack(x.x, x.y));
                const uint32_t vv = ok[r] ? la_lds32(static_cast<const __nv_bfloat16 *>(p.v) + ((long long)b * N + nr[r]) * p.ldv +
                                                     hh * HD + 8 * pi + 2 * t)
                                          : 0u;
                vt[r][pi] = la_trans(vv);
            }
#pragma unroll
        for (int mt = 0; mt < HD / 16; ++mt) {
            const uint32_t a[4] = {kt[0][2 * mt], kt[0][2 * mt + 1], kt[1][2 * mt], kt[1][2 * mt + 1]};
#pragma unroll
            for (int nt = 0; nt < HD / 8; ++nt) la_mma(S[mt][nt], a, vt[0][nt], vt[1][nt]);
            const uint32_t ap[4] = {pt[0][2 * mt], pt[0][2 * mt + 1], pt[1][2 * mt], pt[1][2 * mt + 1]};
            la_mma(KM[mt], ap, ones, ones);
        }
    }
    // fold the four warps' partial sums through shared memory, then one scaled atomic per element
    for (int i = threadIdx.x; i < HD * HD + HD; i += kLTok) red[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int mt = 0; mt < HD / 16; ++mt) {
#pragma unroll
        for (int nt = 0; nt < HD / 8; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e)
                atomicAdd(red + (mt * 16 + g + 8 * (e >> 1)) * HD + nt * 8 + 2 * t + (e & 1), S[mt][nt][e]);
        if (t == 0) {
            atomicAdd(red + HD * HD + mt * 16 + g, KM[mt][0]);
            atomicAdd(red + HD * HD + mt * 16 + g + 8, KM[mt][2]);
        }
    }
    __syncthreads();
    const long long bh = (long long)b * p.h + hh;
    const float scale = 1.f / (float)N;
    for (int i = threadIdx.x; i < HD * HD; i += kLTok) atomicAdd(p.S + bh * HD * HD + i, red[i] * scale);
    if (threadIdx.x < HD) atomicAdd(p.kmean + bh * HD + threadIdx.x, red[HD * HD + threadIdx.x] * scale);
}

// stage an HD x HD fp32 matrix as bf16, row-major (sM[d][e]) and / or transposed (sMt[e][d])
template <int HD>
__device__ __forceinline__ void la_stage_bf16(const float *gM, __nv_bfloat16 (*sM)[HD + 8], __nv_bfloat16 (*sMt)[HD + 8]) {
    for (int i = threadIdx.x; i < HD * HD; i += blockDim.x) {
        const __nv_bfloat16 v = __float2bfloat16_rn(gM[i]);
        if (sM) sM[i / HD][i % HD] = v;
        if (sMt) sMt[i % HD][i / HD] = v;
    }
}

// ---- forward 2: out = z * rope(phi q) S                                           (k = hd)
template <int HD>
__global__ void __launch_bounds__(128) linattn_apply_mma_kernel(const LinAttnParams p) {
    __shared__ __align__(16) __nv_bfloat16 sSt[HD][HD + 8];
    __shared__ float sKm[HD];
    const int b = blockIdx.z, hh = blockIdx.y, N = p.H * p.W;
    const long long bh = (long long)b * p.h + hh;
    la_stage_bf16<HD>(p.S + bh * HD * HD, nullptr, sSt);
    if (threadIdx.x < HD) sKm[threadIdx.x] = p.kmean[bh * HD + threadIdx.x];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    for (int tile = warp; tile < kLmTok / 16; tile += 4) {
        const int nb = blockIdx.x * kLmTok + tile * 16;
        if (nb >= N) break;
        const int nr[2] = {nb + g, nb + g + 8};
        const bool ok[2] = {nr[0] < N, nr[1] < N};
        LaFrag<HD> f;
        la_frag_load<HD>(p, static_cast<const __nv_bfloat16 *>(p.q), p.ldq, hh, b, nr, ok, t, f);
        float z[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            float den = 0.f;
#pragma unroll
            for (int pi = 0; pi < HD / 8; ++pi)
                den = fmaf(f.phi[r][pi].x, sKm[8 * pi + 2 * t], fmaf(f.phi[r][pi].y, sKm[8 * pi + 2 * t + 1], den));
            den += __shfl_xor_sync(0xffffffffu, den, 1);
            den += __shfl_xor_sync(0xffffffffu, den, 2);
            z[r] = 1.f / (den + p.eps);
        }
        uint32_t a[HD / 16][4];
        la_frag_rope_a<HD>(f, a);
#pragma unroll
        for (int nt = 0; nt < HD / 8; ++nt) {
            float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int ks = 0; ks < HD / 16; ++ks)
                la_mma(o, a[ks], la_lds32(&sSt[nt * 8 + g][16 * ks + 2 * t]), la_lds32(&sSt[nt * 8 + g][16 * ks + 8 + 2 * t]));
#pragma unroll
            for (int r = 0; r < 2; ++r)
                if (ok[r])
                    *reinterpret_cast<uint32_t *>(static_cast<__nv_bfloat16 *>(p.out) + ((long long)b * N + nr[r]) * p.ldo +
                                                  hh * HD + nt * 8 + 2 * t) = la_pack(o[2 * r] * z[r], o[2 * r + 1] * z[r]);
        }
    }
}

// ---- backward 1: dq; dS = (1/N) sum rope(phi q) (x) z dO, dkmean = (1/N) sum c phi q     (k = hd, then k = tokens)
template <int HD>
__global__ void __launch_bounds__(128) linattn_bwd_q_mma_kernel(const LinAttnParams p) {
    constexpr int GT = 64;                                      // tokens per iteration (4 warps x 16)
    __shared__ __align__(16) __nv_bfloat16 sS[HD][HD + 8], sSt[HD][HD + 8];
    __shared__ float sKm[HD], red[HD * HD + HD];
    const int b = blockIdx.z, hh = blockIdx.y, N = p.H * p.W;
    const long long bh = (long long)b * p.h + hh;
    const int n0 = blockIdx.x * p.chunk, n1 = min(N, n0 + p.chunk);
    la_stage_bf16<HD>(p.S + bh * HD * HD, sS, sSt);
    if (threadIdx.x < HD) sKm[threadIdx.x] = p.kmean[bh * HD + threadIdx.x];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    float dS[HD / 16][HD / 8][4], dKM[HD / 16][4];
#pragma unroll
    for (int mt = 0; mt < HD / 16; ++mt) {
        dKM[mt][0] = dKM[mt][1] = dKM[mt][2] = dKM[mt][3] = 0.f;
#pragma unroll
        for (int nt = 0; nt < HD / 8; ++nt) dS[mt][nt][0] = dS[mt][nt][1] = dS[mt][nt][2] = dS[mt][nt][3] = 0.f;
    }
    for (int t0 = n0; t0 < n1; t0 += GT) {
        const int nb = t0 + warp * 16;
        const int nr[2] = {nb + g, nb + g + 8};
        const bool ok[2] = {nr[0] < n1, nr[1] < n1};
        LaFrag<HD> f;
        la_frag_load<HD>(p, static_cast<const __nv_bfloat16 *>(p.q), p.ldq, hh, b, nr, ok, t, f);
        float z[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            float den = 0.f;
#pragma unroll
            for (int pi = 0; pi < HD / 8; ++pi)
                den = fmaf(f.phi[r][pi].x, sKm[8 * pi + 2 * t], fmaf(f.phi[r][pi].y, sKm[8 * pi + 2 * t + 1], den));
            den += __shfl_xor_sync(0xffffffffu, den, 1);
            den += __shfl_xor_sync(0xffffffffu, den, 2);
            z[r] = 1.f / (den + p.eps);
        }
        uint32_t a[HD / 16][4];
        la_frag_rope_a<HD>(f, a);
        // t = q_rope S (out / z);  dz = dO . t;  dt = z dO
        float dt[HD / 8][4];
        float dz[2] = {0.f, 0.f};
#pragma unroll
        for (int nt = 0; nt < HD / 8; ++nt) {
            float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int ks = 0; ks < HD / 16; ++ks)
                la_mma(o, a[ks], la_lds32(&sSt[nt * 8 + g][16 * ks + 2 * t]), la_lds32(&sSt[nt * 8 + g][16 * ks + 8 + 2 * t]));
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const float2 go = ok[r] ? la_unpack(la_lds32(static_cast<const __nv_bfloat16 *>(p.dout) +
                                                            ((long long)b * N + nr[r]) * p.lddo + hh * HD + nt * 8 + 2 * t))
                                        : make_float2(0.f, 0.f);
                dz[r] = fmaf(go.x, o[2 * r], fmaf(go.y, o[2 * r + 1], dz[r]));
                dt[nt][2 * r] = go.x * z[r];
                dt[nt][2 * r + 1] = go.y * z[r];
            }
        }
        float cden[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            dz[r] += __shfl_xor_sync(0xffffffffu, dz[r], 1);
            dz[r] += __shfl_xor_sync(0xffffffffu, dz[r], 2);
            cden[r] = -z[r] * z[r] * dz[r];
        }
        // d q_rope = dt S^T  (k = hd over e), rotate back, normaliser term, elu'
        {
            uint32_t da[HD / 16][4];
#pragma unroll
            for (int kk = 0; kk < HD / 16; ++kk) {
                da[kk][0] = la_pack(dt[2 * kk][0], dt[2 * kk][1]);
                da[kk][1] = la_pack(dt[2 * kk][2], dt[2 * kk][3]);
                da[kk][2] = la_pack(dt[2 * kk + 1][0], dt[2 * kk + 1][1]);
                da[kk][3] = la_pack(dt[2 * kk + 1][2], dt[2 * kk + 1][3]);
            }
#pragma unroll
            for (int nd = 0; nd < HD / 8; ++nd) {
                float dq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int kk = 0; kk < HD / 16; ++kk)
                    la_mma(dq, da[kk], la_lds32(&sS[nd * 8 + g][16 * kk + 2 * t]), la_lds32(&sS[nd * 8 + g][16 * kk + 8 + 2 * t]));
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    if (!ok[r]) continue;
                    const float2 c = f.cs[r][nd], ph = f.phi[r][nd];
                    const float g0 = dq[2 * r], g1 = dq[2 * r + 1];
                    float e0 = c.x * g0 + c.y * g1 + cden[r] * sKm[nd * 8 + 2 * t];
                    float e1 = c.x * g1 - c.y * g0 + cden[r] * sKm[nd * 8 + 2 * t + 1];
                    e0 *= ph.x > 1.f ? 1.f : ph.x;
                    e1 *= ph.y > 1.f ? 1.f : ph.y;
                    *reinterpret_cast<uint32_t *>(static_cast<__nv_bfloat16 *>(p.dq) + ((long long)b * N + nr[r]) * p.lddq +
                                                  hh * HD + nd * 8 + 2 * t) = la_pack(e0, e1);
                }
            }
        }
        {   // dS += q_rope^T dt, dkmean += (phi q)^T c over this warp's 16 tokens: movmatrix-transposed register tiles
            uint32_t dtt[2][HD / 8], ptt[2][HD / 8];
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int pi = 0; pi < HD / 8; ++pi) {
                    dtt[r][pi] = la_trans(la_pack(dt[pi][2 * r], dt[pi][2 * r + 1]));
                    ptt[r][pi] = la_trans(la_pack(f.phi[r][pi].x, f.phi[r][pi].y));
                }
            // B operand with c in column 0: lane (g = 0, t) needs c of tokens 2t, 2t+1 (rows g) and 8+2t, 9+2t (rows g+8)
            const float c00 = __shfl_sync(0xffffffffu, cden[0], 8 * t), c01 = __shfl_sync(0xffffffffu, cden[0], 8 * t + 4);
            const float c10 = __shfl_sync(0xffffffffu, cden[1], 8 * t), c11 = __shfl_sync(0xffffffffu, cden[1], 8 * t + 4);
            const uint32_t cb0 = (g == 0) ? la_pack(c00, c01) : 0u, cb1 = (g == 0) ? la_pack(c10, c11) : 0u;
#pragma unroll
            for (int mt = 0; mt < HD / 16; ++mt) {
                const uint32_t aq[4] = {la_trans(a[mt][0]), la_trans(a[mt][2]), la_trans(a[mt][1]), la_trans(a[mt][3])};
#pragma unroll
                for (int nt = 0; nt < HD / 8; ++nt) la_mma(dS[mt][nt], aq, dtt[0][nt], dtt[1][nt]);
                const uint32_t ap[4] = {ptt[0][2 * mt], ptt[0][2 * mt + 1], ptt[1][2 * mt], ptt[1][2 * mt + 1]};
                la_mma(dKM[mt], ap, cb0, cb1);
            }
        }
    }
    for (int i = threadIdx.x; i < HD * HD + HD; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int mt = 0; mt < HD / 16; ++mt) {
#pragma unroll
        for (int nt = 0; nt < HD / 8; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e)
                atomicAdd(red + (mt * 16 + g + 8 * (e >> 1)) * HD + nt * 8 + 2 * t + (e & 1), dS[mt][nt][e]);
        if (t == 0) {
            atomicAdd(red + HD * HD + mt * 16 + g, dKM[mt][0]);
            atomicAdd(red + HD * HD + mt * 16 + g + 8, dKM[mt][2]);
        }
    }
    __syncthreads();
    const float scale = 1.f / (float)N;
    for (int i = threadIdx.x; i < HD * HD; i += blockDim.x) atomicAdd(p.dS + bh * HD * HD + i, red[i] * scale);
    if (threadIdx.x < HD) atomicAdd(p.dkm + bh * HD + threadIdx.x, red[HD * HD + threadIdx.x] * scale);
}

// ---- backward 2: dv = rope(phi k) dS, dk from dS v                                (k = hd)
template <int HD>
__global__ void __launch_bounds__(128) linattn_bwd_kv_mma_kernel(const LinAttnParams p) {
    __shared__ __align__(16) __nv_bfloat16 sdS[HD][HD + 8], sdSt[HD][HD + 8];
    __shared__ float sdkm[HD];
    const int b = blockIdx.z, hh = blockIdx.y, N = p.H * p.W;
    const long long bh = (long long)b * p.h + hh;
    la_stage_bf16<HD>(p.dS + bh * HD * HD, sdS, sdSt);          // already scaled by 1/N
    if (threadIdx.x < HD) sdkm[threadIdx.x] = p.dkm[bh * HD + threadIdx.x];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    for (int tile = warp; tile < kLmTok / 16; tile += 4) {
        const int nb = blockIdx.x * kLmTok + tile * 16;
        if (nb >= N) break;
        const int nr[2] = {nb + g, nb + g + 8};
        const bool ok[2] = {nr[0] < N, nr[1] < N};
        LaFrag<HD> f;
        la_frag_load<HD>(p, static_cast<const __nv_bfloat16 *>(p.k), p.ldk, hh, b, nr, ok, t, f);
        uint32_t ak[HD / 16][4], av[HD / 16][4];
        la_frag_rope_a<HD>(f, ak);
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks)
#pragma unroll
            for (int hf = 0; hf < 2; ++hf)
#pragma unroll
                for (int r = 0; r < 2; ++r)
                    av[ks][2 * hf + r] = ok[r] ? la_lds32(static_cast<const __nv_bfloat16 *>(p.v) + ((long long)b * N + nr[r]) * p.ldv +
                                                         hh * HD + 16 * ks + 8 * hf + 2 * t)
                                               : 0u;
#pragma unroll
        for (int nt = 0; nt < HD / 8; ++nt) {
            float dv[4] = {0.f, 0.f, 0.f, 0.f}, dk[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int ks = 0; ks < HD / 16; ++ks) {
                la_mma(dv, ak[ks], la_lds32(&sdSt[nt * 8 + g][16 * ks + 2 * t]), la_lds32(&sdSt[nt * 8 + g][16 * ks + 8 + 2 * t]));
                la_mma(dk, av[ks], la_lds32(&sdS[nt * 8 + g][16 * ks + 2 * t]), la_lds32(&sdS[nt * 8 + g][16 * ks + 8 + 2 * t]));
            }
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                if (!ok[r]) continue;
                const long long tok = (long long)b * N + nr[r];
                *reinterpret_cast<uint32_t *>(static_cast<__nv_bfloat16 *>(p.dv) + tok * p.lddv + hh * HD + nt * 8 + 2 * t) =
                    la_pack(dv[2 * r], dv[2 * r + 1]);
                const float2 c = f.cs[r][nt], ph = f.phi[r][nt];
                const float g0 = dk[2 * r], g1 = dk[2 * r + 1];
                float e0 = c.x * g0 + c.y * g1 + sdkm[nt * 8 + 2 * t];
                float e1 = c.x * g1 - c.y * g0 + sdkm[nt * 8 + 2 * t + 1];
                e0 *= ph.x > 1.f ? 1.f : ph.x;
                e1 *= ph.y > 1.f ? 1.f : ph.y;
                *reinterpret_cast<uint32_t *>(static_cast<__nv_bfloat16 *>(p.dk) + tok * p.lddk + hh * HD + nt * 8 + 2 * t) =
                    la_pack(e0, e1);
            }
        }
    }
}

template <int HD>
static cudaError_t linattn_launch_mma(LinAttnParams p, int which, cudaStream_t st) {
    const int N = p.H * p.W;
    p.chunk = linattn_chunk(p.Bn, N, p.h);
    const dim3 gtok((N + kLmTok - 1) / kLmTok, p.h, p.Bn), gred((N + p.chunk - 1) / p.chunk, p.h, p.Bn);
    if (which == 0) {
        linattn_state_mma_kernel<HD><<<gred, kLTok, 0, st>>>(p);
        linattn_apply_mma_kernel<HD><<<gtok, 128, 0, st>>>(p);
    } else {
        linattn_bwd_q_mma_kernel<HD><<<gred, 128, 0, st>>>(p);
        linattn_bwd_kv_mma_kernel<HD><<<gtok, 128, 0, st>>>(p);
    }
    return cudaGetLastError();
}
static bool linattn_use_mma(const LinAttnParams &p, int which) {
    const char *e = getenv("MLAGG_LINATTN_MMA");
    if (e && e[0] == '0') return false;
    const bool ev = p.ldq % 2 == 0 && p.ldk % 2 == 0 && p.ldv % 2 == 0;
    return which == 0 ? (ev && p.ldo % 2 == 0) : (ev && p.lddo % 2 == 0 && p.lddq % 2 == 0 && p.lddk % 2 == 0 && p.lddv % 2 == 0);
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
    if (dtype == 1 && (hd == 16 || hd == 32) && linattn_use_mma(p, which))
        return hd == 16 ? linattn_launch_mma<16>(p, which, st) : linattn_launch_mma<32>(p, which, st);
    return dtype == 0 ? linattn_by_hd<float>(p, hd, which, st) : linattn_by_hd<__nv_bfloat16>(p, hd, which, st);
}

}  // namespace mlagg
