// dwconv.cu -- depthwise 3x3 convolution (pad 1, bias, optional fused SiLU) on TOKENS-MAJOR data, and the
// depthwise causal conv1d.  Replaces, without the NCHW<->tokens permute copies the reference needs,
//   nn.Conv2d(groups=C) at nnUNetTrainer_MLAgg_2D_dt_MS.py:851,890 (dwc), :680,782 (lepe);
//   MambaSkip.py:302-312,521-523 (SS2D_skip.conv2d[i] + SiLU), :545-556 (DWConv in ConvolutionalGLU);
//   nnUNetTrainer_MLLA_UNet.py:279,289 (cpe1/2), :215,248 (lepe);
//   causal_conv1d_fn of the un-vendored causal-conv1d package (north_star; not on the trainer's path, F3).
// HBM-bound stencils: every thread owns 4 consecutive channels of one pixel (128-bit fp32 / 64-bit bf16
// accesses, coalesced across the channel dimension); neighbour rows come from L1/L2.
#include <cuda_bf16.h>

#include <cstdlib>

#include "common.cuh"

namespace mlagg {

template <typename T>
__device__ __forceinline__ float4 ld4(const T *p);
template <>
__device__ __forceinline__ float4 ld4<float>(const float *p) {
    return __ldg(reinterpret_cast<const float4 *>(p));
}
template <>
__device__ __forceinline__ float4 ld4<__nv_bfloat16>(const __nv_bfloat16 *p) {
    const uint2 raw = __ldg(reinterpret_cast<const uint2 *>(p));
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162 *>(&raw.x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162 *>(&raw.y);
    const float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
}
template <typename T>
__device__ __forceinline__ void st4(T *p, float4 v);
template <>
__device__ __forceinline__ void st4<float>(float *p, float4 v) {
    *reinterpret_cast<float4 *>(p) = v;
}
template <>
__device__ __forceinline__ void st4<__nv_bfloat16>(__nv_bfloat16 *p, float4 v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 raw;
    raw.x = *reinterpret_cast<uint32_t *>(&a);
    raw.y = *reinterpret_cast<uint32_t *>(&b);
    *reinterpret_cast<uint2 *>(p) = raw;
}

__device__ __forceinline__ float silu_grad(float z) {  // d/dz [z * sigmoid(z)]
    const float s = rcp_approx(1.f + ex2_approx(-z * kLog2e));
    return s * (1.f + z * (1.f - s));
}

// weights as (C, 9) fp32 (the nn.Conv2d (C,1,3,3) tensor, contiguous); tap p = 3*(dr+1) + (dc+1).
// FLIP=false: y[pix] = b + sum_p w[p] x[pix + off_p]           (forward)
// FLIP=true : y[pix] =     sum_p w[p] x[pix - off_p]           (input gradient)
// A thread owns 4 consecutive channels and walks a strip of kStrip pixels along a row with the 3x3 window in
// registers: 3 new 64/128-bit loads, 36 FMAs and one store per pixel; the 36 weights are loaded once per strip (the first
// version re-read 9 neighbours and 36 weights for every pixel and ran at ~1/12 of the HBM roofline).
constexpr int kStrip = 16;

struct Row3 {
    const void *p[3];
};

// pixel (ld*) and image (bs*) strides in elements of the input, the output and the optional second input
// (forward: residual added to the result; weight-gradient pass: dy).  Contiguous tokens-major data: ld = C, bs = H*W*C.
struct DwLay {
    long long ldx, bsx, ldy, bsy, ldr, bsr;
    int rmul;              // forward: the second input MULTIPLIES the activated result (ConvolutionalGLU's a * v) instead of adding
    long long ldm, bsm;    // weight-gradient pass: strides of the multiplier v and of its gradient dv (same layout)
};

template <typename T>
__device__ __forceinline__ void dw_col(const Row3 &rows, int wc, int W, int C, float4 (&o)[3]) {
    const bool in = wc >= 0 && wc < W;
#pragma unroll
    for (int dr = 0; dr < 3; ++dr)
        o[dr] = (in && rows.p[dr]) ? ld4<T>(static_cast<const T *>(rows.p[dr]) + (long long)wc * C) : make_float4(0, 0, 0, 0);
}
__device__ __forceinline__ void dw_fma(float4 &acc, const float (&wr)[4][9], int p, const float4 &v) {
    acc.x = fmaf(wr[0][p], v.x, acc.x);
    acc.y = fmaf(wr[1][p], v.y, acc.y);
    acc.z = fmaf(wr[2][p], v.z, acc.z);
    acc.w = fmaf(wr[3][p], v.w, acc.w);
}
__device__ __forceinline__ float4 dw_window(const float4 &b, const float (&wr)[4][9], const float4 (&L)[3],
                                            const float4 (&M)[3], const float4 (&R)[3]) {
    float4 acc = b;
#pragma unroll
    for (int dr = 0; dr < 3; ++dr) {
        dw_fma(acc, wr, 3 * dr, L[dr]);
        dw_fma(acc, wr, 3 * dr + 1, M[dr]);
        dw_fma(acc, wr, 3 * dr + 2, R[dr]);
    }
    return acc;
}

template <typename T, bool FLIP, int ACT, bool RMUL>
__global__ void __launch_bounds__(256) dwconv3x3_kernel(const T *__restrict__ x, const float *__restrict__ w,
                                                        const float *__restrict__ bias, const T *__restrict__ res,
                                                        T *__restrict__ y, int Bn, int H, int W, int C, DwLay lay) {
    const int cv = C >> 2;
    const int SW = (W + kStrip - 1) / kStrip;
    const long long total = (long long)Bn * H * SW * cv;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c = (int)(idx % cv) << 2;
    long long s = idx / cv;
    const int w0 = (int)(s % SW) * kStrip;
    s /= SW;
    const int hr = (int)(s % H);
    const long long bimg = s / H;        // image index
    float wr[4][9];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int p = 0; p < 9; ++p) wr[k][p] = __ldg(w + (c + k) * 9 + (FLIP ? 8 - p : p));
    const float4 bv = (!FLIP && bias) ? __ldg(reinterpret_cast<const float4 *>(bias + c)) : make_float4(0, 0, 0, 0);
    Row3 rows;
#pragma unroll
    for (int dr = 0; dr < 3; ++dr) {
        const int rr = hr + dr - 1;
        rows.p[dr] = (rr >= 0 && rr < H) ? static_cast<const void *>(x + bimg * lay.bsx + ((long long)rr * W) * lay.ldx + c)
                                         : nullptr;
    }
    T *yr = y + bimg * lay.bsy + ((long long)hr * W) * lay.ldy + c;
    const T *rsr = res ? res + bimg * lay.bsr + ((long long)hr * W) * lay.ldr + c : nullptr;
    const int w1 = min(W, w0 + kStrip);
    const int ldx = (int)lay.ldx;
    float4 A[3], Bc[3], Cc[3];
    dw_col<T>(rows, w0 - 1, W, ldx, A);
    dw_col<T>(rows, w0, W, ldx, Bc);
    auto step = [&](int wc, const float4 (&L)[3], const float4 (&M)[3], float4 (&R)[3]) {
        dw_col<T>(rows, wc + 1, W, ldx, R);
        float4 acc = dw_window(bv, wr, L, M, R);
        if (ACT == 1) {
            acc.x = silu_f(acc.x); acc.y = silu_f(acc.y); acc.z = silu_f(acc.z); acc.w = silu_f(acc.w);
        }
        if (rsr) {
            const float4 r4 = ld4<T>(rsr + (long long)wc * lay.ldr);
            if (RMUL) {
                acc.x *= r4.x; acc.y *= r4.y; acc.z *= r4.z; acc.w *= r4.w;
            } else {
                acc.x += r4.x; acc.y += r4.y; acc.z += r4.z; acc.w += r4.w;
            }
        }
        st4<T>(yr + (long long)wc * lay.ldy, acc);
    };
    for (int wc = w0; wc < w1; wc += 3) {
        step(wc, A, Bc, Cc);
        if (wc + 1 < w1) step(wc + 1, Bc, Cc, A);
        if (wc + 2 < w1) step(wc + 2, Cc, A, Bc);
    }
}


// ---- C % 4 != 0 (narrow test configurations): one channel per thread, same arithmetic.
template <typename T>
__device__ __forceinline__ float ld1(const T *p);
template <>
__device__ __forceinline__ float ld1<float>(const float *p) { return __ldg(p); }
template <>
__device__ __forceinline__ float ld1<__nv_bfloat16>(const __nv_bfloat16 *p) { return __bfloat162float(*p); }
template <typename T>
__device__ __forceinline__ void st1(T *p, float v);
template <>
__device__ __forceinline__ void st1<float>(float *p, float v) { *p = v; }
template <>
__device__ __forceinline__ void st1<__nv_bfloat16>(__nv_bfloat16 *p, float v) { *p = __float2bfloat16_rn(v); }

// mode 0: y = act(conv(x)+b) [+ res]; mode 1 (FLIP): y = conv_flipped(x); mode 2: dz = dy * act'(conv(x)+b) and dw/db
// atomics (dy addressed through lay.ldr / lay.bsr)
template <typename T, int ACT>
__global__ void __launch_bounds__(256) dwconv3x3_scalar_kernel(const T *__restrict__ x, const float *__restrict__ w,
                                                               const float *__restrict__ bias,
                                                               const T *__restrict__ dy, T *__restrict__ y,
                                                               float *__restrict__ dw, float *__restrict__ db, int Bn,
                                                               int H, int W, int C, int mode, DwLay lay,
                                                               const T *__restrict__ mulv, T *__restrict__ dmul) {
    const long long total = (long long)Bn * H * W * C;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % C);
        const long long pix = idx / C;
        const int wc = (int)(pix % W), hr = (int)((pix / W) % H);
        const long long bimg = pix / ((long long)H * W);
        const long long lpix = (long long)hr * W + wc;
        const T *xb = x + bimg * lay.bsx + c;
        float acc = (mode != 1 && bias) ? __ldg(bias + c) : 0.f;
        float xn[9];
#pragma unroll
        for (int dr = -1; dr <= 1; ++dr)
#pragma unroll
            for (int dc = -1; dc <= 1; ++dc) {
                const int p = 3 * (dr + 1) + (dc + 1);
                const int rr = hr + dr, cc = wc + dc;
                const bool ok = rr >= 0 && rr < H && cc >= 0 && cc < W;
                xn[p] = ok ? ld1<T>(xb + ((long long)rr * W + cc) * lay.ldx) : 0.f;
                acc = fmaf(__ldg(w + c * 9 + (mode == 1 ? 8 - p : p)), xn[p], acc);
            }
        T *yp = y + bimg * lay.bsy + lpix * lay.ldy + c;
        if (mode == 2) {
            float g = ld1<T>(dy + bimg * lay.bsr + lpix * lay.ldr + c);
            if (mulv) {   // y = act(z) * v:  dv = g act(z),  g <- g v
                st1<T>(dmul + bimg * lay.bsm + lpix * lay.ldm + c, g * (ACT == 1 ? silu_f(acc) : acc));
                g *= ld1<T>(mulv + bimg * lay.bsm + lpix * lay.ldm + c);
            }
            if (ACT == 1) g *= silu_grad(acc);
            st1<T>(yp, g);
#pragma unroll
            for (int p = 0; p < 9; ++p) atomicAdd(dw + c * 9 + p, g * xn[p]);
            if (db) atomicAdd(db + c, g);
        } else {
            float o = (ACT == 1 && mode == 0) ? silu_f(acc) : acc;
            if (mode == 0 && dy) {                                                        // second input rides in `dy`
                const float r = ld1<T>(dy + bimg * lay.bsr + lpix * lay.ldr + c);
                o = lay.rmul ? o * r : o + r;
            }
            st1<T>(yp, o);
        }
    }
}

// Backward, pass 1: recompute z = conv(x) + b, dz = dy * act'(z); store dz; accumulate
// dw[c, p] += sum_pix dz[pix] x[pix + off_p] and db[c] += sum_pix dz[pix] (registers -> smem -> atomics).
// Block = (C/4 channel vectors) x PP strip lanes; lane pl walks `strips_per_lane` row strips of kStripW pixels with the
// 3x3 window of x in registers (same walk as the forward kernel).
constexpr int kStripW = 32;

template <typename T, int ACT, bool MUL>
__global__ void __launch_bounds__(256) dwconv3x3_bwd_w_kernel(const T *__restrict__ x, const float *__restrict__ w,
                                                              const float *__restrict__ bias,
                                                              const T *__restrict__ dy, T *__restrict__ dz,
                                                              float *__restrict__ dw, float *__restrict__ db,
                                                              int Bn, int H, int W, int C, int strips_per_lane,
                                                              DwLay lay, const T *__restrict__ mulv,
                                                              T *__restrict__ dmul) {
    extern __shared__ float red[];  // [PP][cv][40]    (x: lay.ldx/bsx, dy: lay.ldr/bsr, dz: lay.ldy/bsy)
    const int cv = C >> 2;
    const int PP = blockDim.x / cv;
    const int cvi = threadIdx.x % cv, pl = threadIdx.x / cv;
    const int c = cvi << 2;
    const int SW = (W + kStripW - 1) / kStripW;
    const long long nstrips = (long long)Bn * H * SW;
    float aw[4][9], ab[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        ab[k] = 0.f;
#pragma unroll
        for (int p = 0; p < 9; ++p) aw[k][p] = 0.f;
    }
    if (pl < PP) {
        float wr[4][9];
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int p = 0; p < 9; ++p) wr[k][p] = __ldg(w + (c + k) * 9 + p);
        const float4 bv = bias ? __ldg(reinterpret_cast<const float4 *>(bias + c)) : make_float4(0, 0, 0, 0);
        for (int j = 0; j < strips_per_lane; ++j) {
            long long s = ((long long)blockIdx.x * strips_per_lane + j) * PP + pl;
            if (s >= nstrips) break;
            const int w0 = (int)(s % SW) * kStripW;
            s /= SW;
            const int hr = (int)(s % H);
            const long long bimg = s / H;
            Row3 rows;
#pragma unroll
            for (int dr = 0; dr < 3; ++dr) {
                const int rr = hr + dr - 1;
                rows.p[dr] = (rr >= 0 && rr < H)
                                 ? static_cast<const void *>(x + bimg * lay.bsx + ((long long)rr * W) * lay.ldx + c)
                                 : nullptr;
            }
            const T *gyr = dy + bimg * lay.bsr + ((long long)hr * W) * lay.ldr + c;
            T *dzr = dz + bimg * lay.bsy + ((long long)hr * W) * lay.ldy + c;
            const T *mvr = MUL ? mulv + bimg * lay.bsm + ((long long)hr * W) * lay.ldm + c : nullptr;
            T *dmr = MUL ? dmul + bimg * lay.bsm + ((long long)hr * W) * lay.ldm + c : nullptr;
            const int w1 = min(W, w0 + kStripW);
            const int ldx = (int)lay.ldx;
            float4 A[3], Bc[3], Cc[3];
            dw_col<T>(rows, w0 - 1, W, ldx, A);
            dw_col<T>(rows, w0, W, ldx, Bc);
            auto step = [&](int wc, const float4 (&L)[3], const float4 (&M)[3], float4 (&R)[3]) {
                dw_col<T>(rows, wc + 1, W, ldx, R);
                float4 g = ld4<T>(gyr + (long long)wc * lay.ldr);
                if (ACT == 1 || MUL) {
                    const float4 z = dw_window(bv, wr, L, M, R);
                    if (MUL) {   // y = act(z) * v:  dv = g act(z),  g <- g v
                        const float4 v4 = ld4<T>(mvr + (long long)wc * lay.ldm);
                        float4 a4 = z;
                        if (ACT == 1) { a4.x = silu_f(z.x); a4.y = silu_f(z.y); a4.z = silu_f(z.z); a4.w = silu_f(z.w); }
                        st4<T>(dmr + (long long)wc * lay.ldm, make_float4(g.x * a4.x, g.y * a4.y, g.z * a4.z, g.w * a4.w));
                        g.x *= v4.x; g.y *= v4.y; g.z *= v4.z; g.w *= v4.w;
                    }
                    if (ACT == 1) {
                        g.x *= silu_grad(z.x); g.y *= silu_grad(z.y); g.z *= silu_grad(z.z); g.w *= silu_grad(z.w);
                    }
                }
                st4<T>(dzr + (long long)wc * lay.ldy, g);
                ab[0] += g.x; ab[1] += g.y; ab[2] += g.z; ab[3] += g.w;
#pragma unroll
                for (int dr = 0; dr < 3; ++dr) {
                    aw[0][3 * dr] = fmaf(g.x, L[dr].x, aw[0][3 * dr]);
                    aw[1][3 * dr] = fmaf(g.y, L[dr].y, aw[1][3 * dr]);
                    aw[2][3 * dr] = fmaf(g.z, L[dr].z, aw[2][3 * dr]);
                    aw[3][3 * dr] = fmaf(g.w, L[dr].w, aw[3][3 * dr]);
                    aw[0][3 * dr + 1] = fmaf(g.x, M[dr].x, aw[0][3 * dr + 1]);
                    aw[1][3 * dr + 1] = fmaf(g.y, M[dr].y, aw[1][3 * dr + 1]);
                    aw[2][3 * dr + 1] = fmaf(g.z, M[dr].z, aw[2][3 * dr + 1]);
                    aw[3][3 * dr + 1] = fmaf(g.w, M[dr].w, aw[3][3 * dr + 1]);
                    aw[0][3 * dr + 2] = fmaf(g.x, R[dr].x, aw[0][3 * dr + 2]);
                    aw[1][3 * dr + 2] = fmaf(g.y, R[dr].y, aw[1][3 * dr + 2]);
                    aw[2][3 * dr + 2] = fmaf(g.z, R[dr].z, aw[2][3 * dr + 2]);
                    aw[3][3 * dr + 2] = fmaf(g.w, R[dr].w, aw[3][3 * dr + 2]);
                }
            };
            for (int wc = w0; wc < w1; wc += 3) {
                step(wc, A, Bc, Cc);
                if (wc + 1 < w1) step(wc + 1, Bc, Cc, A);
                if (wc + 2 < w1) step(wc + 2, Cc, A, Bc);
            }
        }
        float *mine = red + ((size_t)pl * cv + cvi) * 40;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
            for (int p = 0; p < 9; ++p) mine[k * 10 + p] = aw[k][p];
            mine[k * 10 + 9] = ab[k];
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < cv * 40; i += blockDim.x) {
        float s = 0.f;
        for (int l = 0; l < PP; ++l) s += red[(size_t)l * cv * 40 + i];
        const int ch = (i / 40) * 4 + (i % 40) / 10, p = i % 10;
        if (p < 9) atomicAdd(dw + ch * 9 + p, s);
        else if (db) atomicAdd(db + ch, s);
    }
}

// ------------------------------------------------------------------ ring kernels (the production path)
// The strip kernels above take every operand straight from global memory, three 8-byte loads per pixel behind a
// rolling-window dependency: at 8 resident warps per SM (170 registers) they ran at 15 - 20 % of the HBM roofline
// (tools/call_shapes.py: 257 us for the 245 MB of a 10 x 160 x 160 x 96 backward).  Here a CTA owns (image, chunk of
// CC = 16 | 32 channels, band of rows) and walks DOWN the band RG rows at a time.  Rows of x (with one halo row above and
// below), of the second input (forward: residual / multiplier; weight pass: dy) and of the third (weight pass:
// ConvolutionalGLU's v) arrive in shared-memory rings through 16-byte cp.async (LDGSTS), issued one row group ahead of
// the arithmetic; every global byte is read once (+ 2 halo rows per band).  The arithmetic is the strip walk of the
// old kernels -- a thread owns 4 channels and SL consecutive pixels of one row, 3x3 window in registers -- fed from
// shared memory.  Zero columns on both sides of every ring row and zero rows outside the image replace all border
// predicates.  Results are identical to the strip kernels (same operation order per output).
constexpr int kRingThreads = 256;
constexpr int kRingSmemMax = 220 * 1024;

struct DwRing {
    const void *x, *a1, *a2;   // x: stencil input; a1: fwd second input / weight pass dy; a2: weight pass multiplier v
    void *o1, *o2;             // fwd y | flipped dx | weight pass dz;   o2: weight pass dv
    const float *w, *bias;
    float *dw, *db;
    int H, W;
    long long ldx, bsx, lda1, bsa1, lda2, bsa2, ldo1, bso1, ldo2, bso2;
    int nchunk, ncol, G, RG, NS, rmul, D;   // D: row groups in flight ahead of the arithmetic (1..4)
};

__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// a thread's four channels as two f32x2 pairs: all arithmetic below is FFMA2 / FMUL2 / FADD2 (two exact fp32 operations per
// issue slot -- the kernels are issue-bound, not HBM-bound, once the operands come from shared memory)
struct V4 {
    float2 lo, hi;
};
template <typename T>
__device__ __forceinline__ V4 lds4(const unsigned char *p);
template <>
__device__ __forceinline__ V4 lds4<float>(const unsigned char *p) {
    const float4 t = *reinterpret_cast<const float4 *>(p);
    return V4{make_float2(t.x, t.y), make_float2(t.z, t.w)};
}
template <>
__device__ __forceinline__ V4 lds4<__nv_bfloat16>(const unsigned char *p) {
    const uint2 raw = *reinterpret_cast<const uint2 *>(p);
    return V4{make_float2(__uint_as_float(raw.x << 16), __uint_as_float(raw.x & 0xffff0000u)),
              make_float2(__uint_as_float(raw.y << 16), __uint_as_float(raw.y & 0xffff0000u))};
}
__device__ __forceinline__ void stv4(float *p, const V4 &v) {
    *reinterpret_cast<float4 *>(p) = make_float4(v.lo.x, v.lo.y, v.hi.x, v.hi.y);
}
__device__ __forceinline__ void stv4(__nv_bfloat16 *p, const V4 &v) {
    const __nv_bfloat162 a = __floats2bfloat162_rn(v.lo.x, v.lo.y), b = __floats2bfloat162_rn(v.hi.x, v.hi.y);
    uint2 raw;
    raw.x = *reinterpret_cast<const uint32_t *>(&a);
    raw.y = *reinterpret_cast<const uint32_t *>(&b);
    *reinterpret_cast<uint2 *>(p) = raw;
}
__device__ __forceinline__ float2 silu2(float2 z) {          // == (silu_f(z.x), silu_f(z.y)) bit for bit
    const float2 t = __fmul2_rn(z, make_float2(-kLog2e, -kLog2e));
    const float2 d = __fadd2_rn(make_float2(ex2_approx(t.x), ex2_approx(t.y)), make_float2(1.f, 1.f));
    return __fmul2_rn(z, make_float2(rcp_approx(d.x), rcp_approx(d.y)));
}
__device__ __forceinline__ float2 silu_grad2(float2 z) {     // == (silu_grad(z.x), silu_grad(z.y))
    const float2 t = __fmul2_rn(z, make_float2(-kLog2e, -kLog2e));
    const float2 d = __fadd2_rn(make_float2(ex2_approx(t.x), ex2_approx(t.y)), make_float2(1.f, 1.f));
    const float2 sg = make_float2(rcp_approx(d.x), rcp_approx(d.y));
    const float2 om = __fadd2_rn(make_float2(1.f, 1.f), make_float2(-sg.x, -sg.y));
    return __fmul2_rn(sg, __ffma2_rn(z, om, make_float2(1.f, 1.f)));
}
// window sum in the tap order of dw_window (L, M, R per row), two channel pairs
__device__ __forceinline__ V4 dw_window2(const V4 &b, const float2 (&wl)[9], const float2 (&wh)[9], const V4 (&L)[3],
                                         const V4 (&M)[3], const V4 (&R)[3]) {
    V4 acc = b;
#pragma unroll
    for (int dr = 0; dr < 3; ++dr) {
        acc.lo = __ffma2_rn(wl[3 * dr], L[dr].lo, acc.lo);
        acc.hi = __ffma2_rn(wh[3 * dr], L[dr].hi, acc.hi);
        acc.lo = __ffma2_rn(wl[3 * dr + 1], M[dr].lo, acc.lo);
        acc.hi = __ffma2_rn(wh[3 * dr + 1], M[dr].hi, acc.hi);
        acc.lo = __ffma2_rn(wl[3 * dr + 2], R[dr].lo, acc.lo);
        acc.hi = __ffma2_rn(wh[3 * dr + 2], R[dr].hi, acc.hi);
    }
    return acc;
}

// MODE 0: y = act(conv(x) + b) [+ a1 | * a1];  MODE 1: y = conv_flipped(x);  MODE 2: dz = dy * act'(conv(x) + b) (dy = a1;
// with MUL: dv = dy * act(z) -> o2, dy <- dy * v first), dw / db accumulated in registers over the band -> shuffles ->
// shared memory -> one atomicAdd per (channel, tap) and CTA.
template <typename T, int SL, int CV, int MODE, int ACT, bool MUL>
__global__ void __launch_bounds__(kRingThreads) dw3x3_ring_kernel(const DwRing p) {
    constexpr int CC = CV * 4;
    constexpr int CCB = CC * (int)sizeof(T);     // bytes of one pixel's channel chunk
    constexpr int PIECES = CCB / 16;
    constexpr int VB = 4 * (int)sizeof(T);       // bytes of a thread's 4-channel vector
    constexpr int IT = kRingThreads / CV;        // (row, strip) items the CTA works on at once
    extern __shared__ __align__(16) unsigned char ring[];
    const int tid = threadIdx.x;
    const int NS = p.NS, RG = p.RG, W = p.W, H = p.H;
    const int col = blockIdx.x % p.ncol, bidx = blockIdx.x / p.ncol;
    const int nb = p.G / p.ncol + (col < p.G % p.ncol ? 1 : 0);
    const int h0 = (int)((long long)bidx * H / nb), h1 = (int)((long long)(bidx + 1) * H / nb);
    if (h0 >= h1) return;
    const int bimg = col / p.nchunk, c0 = (col % p.nchunk) * CC;
    const bool has1 = MODE == 2 || (MODE == 0 && p.a1 != nullptr);
    const bool rmul = p.rmul != 0;
    const int PW = NS * SL;
    const int RSX = (PW + 2) * CCB, RSA = PW * CCB;
    const int D = p.D;
    const int NRX = (D + 1) * RG + 2, NRA = (D + 1) * RG;
    unsigned char *xs = ring;
    unsigned char *a1s = xs + (size_t)NRX * RSX;
    unsigned char *a2s = a1s + (has1 ? (size_t)NRA * RSA : 0);
    const int ring_bytes = NRX * RSX + (has1 ? NRA * RSA : 0) + (MUL ? NRA * RSA : 0);
    for (int i = tid * 16; i < ring_bytes; i += kRingThreads * 16) *reinterpret_cast<uint4 *>(ring + i) = make_uint4(0, 0, 0, 0);
    __syncthreads();

    const T *xg = static_cast<const T *>(p.x) + bimg * p.bsx + c0;
    const T *g1 = has1 ? static_cast<const T *>(p.a1) + bimg * p.bsa1 + c0 : nullptr;
    const T *g2 = MUL ? static_cast<const T *>(p.a2) + bimg * p.bsa2 + c0 : nullptr;
    // rows [r0, r1) of x (zero rows outside the image) into ring slots sx.., rows [q0, q1) of the second / third input
    // into slots sa..; slots advance with wrap-around (no divisions in the loops)
    auto load = [&](int r0, int r1, int sx, int q0, int q1, int sa) {
        for (int row = r0; row < r1; ++row) {
            unsigned char *dst = xs + (size_t)sx * RSX + CCB;
            if (++sx == NRX) sx = 0;
            if (row < 0 || row >= H) {
                for (int i = tid; i < W * PIECES; i += kRingThreads) *reinterpret_cast<uint4 *>(dst + i * 16) = make_uint4(0, 0, 0, 0);
            } else {
                const T *src = xg + (long long)row * W * p.ldx;
                for (int i = tid; i < W * PIECES; i += kRingThreads)
                    cp_async16(dst + i * 16, src + (long long)(i / PIECES) * p.ldx + (i % PIECES) * (16 / (int)sizeof(T)));
            }
        }
        if (has1)
            for (int row = q0; row < q1; ++row) {
                unsigned char *dst = a1s + (size_t)sa * RSA;
                const T *src = g1 + (long long)row * W * p.lda1;
                for (int i = tid; i < W * PIECES; i += kRingThreads)
                    cp_async16(dst + i * 16, src + (long long)(i / PIECES) * p.lda1 + (i % PIECES) * (16 / (int)sizeof(T)));
                if (MUL) {
                    unsigned char *dst2 = a2s + (size_t)sa * RSA;
                    const T *src2 = g2 + (long long)row * W * p.lda2;
                    for (int i = tid; i < W * PIECES; i += kRingThreads)
                        cp_async16(dst2 + i * 16, src2 + (long long)(i / PIECES) * p.lda2 + (i % PIECES) * (16 / (int)sizeof(T)));
                }
                if (++sa == NRA) sa = 0;
            }
    };

    const int j = tid % CV;
    const int c = c0 + 4 * j;
    float2 wl[9], wh[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) {
        const int qq = MODE == 1 ? 8 - q : q;
        wl[q] = make_float2(__ldg(p.w + c * 9 + qq), __ldg(p.w + (c + 1) * 9 + qq));
        wh[q] = make_float2(__ldg(p.w + (c + 2) * 9 + qq), __ldg(p.w + (c + 3) * 9 + qq));
    }
    V4 bv{make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
    if (MODE != 1 && p.bias) {
        const float4 t = __ldg(reinterpret_cast<const float4 *>(p.bias + c));
        bv = V4{make_float2(t.x, t.y), make_float2(t.z, t.w)};
    }
    float2 awl[9], awh[9], abl = make_float2(0.f, 0.f), abh = make_float2(0.f, 0.f);
    if (MODE == 2) {
#pragma unroll
        for (int q = 0; q < 9; ++q) awl[q] = awh[q] = make_float2(0.f, 0.f);
    }
    T *o1 = static_cast<T *>(p.o1) + bimg * p.bso1 + c;
    T *o2 = MUL ? static_cast<T *>(p.o2) + bimg * p.bso2 + c : nullptr;
    // this thread's first item of a row group: (row r_first, strip s_first); later items advance by IT strips
    const int it0 = tid / CV;
    const int r_first = it0 / NS, s_first = it0 - r_first * NS;
    const int r_step = IT / NS, s_step = IT - r_step * NS;

    const int NG = (h1 - h0 + RG - 1) / RG;
    // row group k: second-input rows [s(k), e(k)), x rows (e(k-1), e(k)] (group 0 also brings the halo row h0 - 1 and row h0)
    int lx = 0, la = 0;                      // ring slots the next group is loaded into
    auto load_group = [&](int k) {
        const int sk = min(h1, h0 + k * RG), ek = min(h1, sk + RG);
        const int r0 = k == 0 ? h0 - 1 : sk + 1, r1 = (k == 0 || ek > sk) ? ek + 1 : sk + 1;
        load(r0, r1, lx, sk, ek, la);
        lx += r1 - r0;
        if (lx >= NRX) lx -= NRX;
        la += ek - sk;
        if (la >= NRA) la -= NRA;
        cp_async_commit();
    };
    for (int k = 0; k < D; ++k) load_group(k);
    int cx = 0, ca = 0;                      // ring slots of x row gs - 1 and of second-input row gs
    for (int g = 0; g < NG; ++g) {
        const int gs = h0 + g * RG, ge = min(h1, gs + RG);
        load_group(g + D);
        if (D == 1) cp_async_wait<1>();
        else if (D == 2) cp_async_wait<2>();
        else if (D == 3) cp_async_wait<3>();
        else cp_async_wait<4>();
        __syncthreads();
        for (int r = r_first, sidx = s_first; r < ge - gs;) {
            const int row = gs + r, w0 = sidx * SL;
            const unsigned char *xr[3];
#pragma unroll
            for (int dr = 0; dr < 3; ++dr) {
                int sl = cx + r + dr;
                if (sl >= NRX) sl -= NRX;
                xr[dr] = xs + (size_t)sl * RSX + w0 * CCB + j * VB;   // pixel w0 - 1
            }
            int sa = ca + r;
            if (sa >= NRA) sa -= NRA;
            const unsigned char *ar = a1s + (size_t)sa * RSA + w0 * CCB + j * VB;
            const unsigned char *mr = a2s + (size_t)sa * RSA + w0 * CCB + j * VB;
            const long long pix0 = (long long)row * W + w0;
            T *q1 = o1 + pix0 * p.ldo1;
            T *q2 = MUL ? o2 + pix0 * p.ldo2 : nullptr;
            const int nin = min(SL, W - w0);          // pixels of this strip inside the row
            V4 A[3], Bc[3], R[3];
#pragma unroll
            for (int dr = 0; dr < 3; ++dr) {
                A[dr] = lds4<T>(xr[dr]);
                Bc[dr] = lds4<T>(xr[dr] + CCB);
            }
#pragma unroll
            for (int i = 0; i < SL; ++i) {
#pragma unroll
                for (int dr = 0; dr < 3; ++dr) R[dr] = lds4<T>(xr[dr] + (i + 2) * CCB);
                if (MODE != 2) {
                    V4 acc = dw_window2(bv, wl, wh, A, Bc, R);
                    if (ACT == 1) acc.lo = silu2(acc.lo), acc.hi = silu2(acc.hi);
                    if (MODE == 0 && has1) {
                        const V4 r4 = lds4<T>(ar + i * CCB);
                        if (rmul) acc.lo = __fmul2_rn(acc.lo, r4.lo), acc.hi = __fmul2_rn(acc.hi, r4.hi);
                        else acc.lo = __fadd2_rn(acc.lo, r4.lo), acc.hi = __fadd2_rn(acc.hi, r4.hi);
                    }
                    if (i < nin) stv4(q1 + (long long)i * p.ldo1, acc);
                } else {
                    V4 gq = lds4<T>(ar + i * CCB);
                    if (ACT == 1 || MUL) {
                        const V4 z = dw_window2(bv, wl, wh, A, Bc, R);
                        if (MUL) {
                            const V4 v4 = lds4<T>(mr + i * CCB);
                            V4 a4 = z;
                            if (ACT == 1) a4.lo = silu2(z.lo), a4.hi = silu2(z.hi);
                            if (i < nin) stv4(q2 + (long long)i * p.ldo2, V4{__fmul2_rn(gq.lo, a4.lo), __fmul2_rn(gq.hi, a4.hi)});
                            gq.lo = __fmul2_rn(gq.lo, v4.lo), gq.hi = __fmul2_rn(gq.hi, v4.hi);
                        }
                        if (ACT == 1) gq.lo = __fmul2_rn(gq.lo, silu_grad2(z.lo)), gq.hi = __fmul2_rn(gq.hi, silu_grad2(z.hi));
                    }
                    if (i < nin) stv4(q1 + (long long)i * p.ldo1, gq);
                    abl = __fadd2_rn(abl, gq.lo), abh = __fadd2_rn(abh, gq.hi);
#pragma unroll
                    for (int dr = 0; dr < 3; ++dr) {
                        awl[3 * dr] = __ffma2_rn(gq.lo, A[dr].lo, awl[3 * dr]);
                        awh[3 * dr] = __ffma2_rn(gq.hi, A[dr].hi, awh[3 * dr]);
                        awl[3 * dr + 1] = __ffma2_rn(gq.lo, Bc[dr].lo, awl[3 * dr + 1]);
                        awh[3 * dr + 1] = __ffma2_rn(gq.hi, Bc[dr].hi, awh[3 * dr + 1]);
                        awl[3 * dr + 2] = __ffma2_rn(gq.lo, R[dr].lo, awl[3 * dr + 2]);
                        awh[3 * dr + 2] = __ffma2_rn(gq.hi, R[dr].hi, awh[3 * dr + 2]);
                    }
                }
#pragma unroll
                for (int dr = 0; dr < 3; ++dr) {
                    A[dr] = Bc[dr];
                    Bc[dr] = R[dr];
                }
            }
            r += r_step, sidx += s_step;
            if (sidx >= NS) sidx -= NS, ++r;
        }
        cx += ge - gs;
        if (cx >= NRX) cx -= NRX;
        ca += ge - gs;
        if (ca >= NRA) ca -= NRA;
        __syncthreads();
    }
    if (MODE == 2) {
        // threads with equal tid % CV hold partial sums of the same 4 channels: lanes first, then the 8 warps
        float *red = reinterpret_cast<float *>(ring);      // [8 warps][CV][40]; the rings are dead after the last barrier
        const int lane = tid & 31, warp = tid >> 5;
        float aw[4][10];
#pragma unroll
        for (int q = 0; q < 9; ++q) aw[0][q] = awl[q].x, aw[1][q] = awl[q].y, aw[2][q] = awh[q].x, aw[3][q] = awh[q].y;
        aw[0][9] = abl.x, aw[1][9] = abl.y, aw[2][9] = abh.x, aw[3][9] = abh.y;
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int q = 0; q < 10; ++q)
#pragma unroll
                for (int o = CV; o < 32; o <<= 1) aw[k][q] += __shfl_xor_sync(0xffffffffu, aw[k][q], o);
        if (lane < CV) {
            float *mine = red + (warp * CV + lane) * 40;
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int q = 0; q < 10; ++q) mine[k * 10 + q] = aw[k][q];
        }
        __syncthreads();
        for (int i = tid; i < CV * 40; i += kRingThreads) {
            float sum = 0.f;
#pragma unroll
            for (int wq = 0; wq < kRingThreads / 32; ++wq) sum += red[wq * CV * 40 + i];
            const int ch = c0 + (i / 40) * 4 + (i % 40) / 10, q = i % 10;
            if (q < 9) atomicAdd(p.dw + ch * 9 + q, sum);
            else if (p.db) atomicAdd(p.db + ch, sum);
        }
    }
}

// geometry of a ring launch; false when the shape / alignment / shared-memory need is outside what the ring kernels take
struct RingPlan {
    int SL, CV, RG, NS, G, nchunk, ncol, D;
    size_t smem;
};
static bool ring_aligned(const void *ptr, long long ld, long long bs, size_t esz) {
    return ptr == nullptr || ((reinterpret_cast<uintptr_t>(ptr) % 16 == 0) && (ld * (long long)esz) % 16 == 0 && (bs * (long long)esz) % 16 == 0);
}
static bool ring_plan(int Bn, int H, int W, int C, size_t esz, int nrings_a, bool two_per_sm, RingPlan &pl) {
    if (getenv("MLAGG_DWCONV_STRIP")) return false;
    if (C % 16 != 0 || H < 1 || W < 1) return false;
    for (int CV : {8, 4}) {
        if (C % (CV * 4) != 0) continue;
        double best = -1.0;
        RingPlan cand{};
        for (int SL : {8, 5}) {
            const int NS = (W + SL - 1) / SL, per_row = NS * CV;
            int RG = per_row >= kRingThreads ? 1 : kRingThreads / per_row;
            if (RG > H) RG = H;
            const int iters = (RG * per_row + kRingThreads - 1) / kRingThreads;
            const double eff = (double)RG * W * CV / ((double)iters * kRingThreads * SL);
            const int CCB = CV * 4 * (int)esz;
            // deepest prefetch (<= 4 row groups); the forward kernels (<= 128 registers) keep room for two CTAs per SM
            int D = 0;
            size_t smem = 0;
            for (int d = 4; d >= 1 && D == 0; --d) {
                const size_t need = (size_t)((d + 1) * RG + 2) * (NS * SL + 2) * CCB + (size_t)nrings_a * (d + 1) * RG * NS * SL * CCB;
                if (need <= (size_t)((two_per_sm && d > 2) ? 110 * 1024 : kRingSmemMax)) D = d, smem = need;
            }
            if (D == 0) continue;
            if (eff > best + 1e-9) {
                best = eff;
                cand = RingPlan{SL, CV, RG, NS, 0, C / (CV * 4), Bn * (C / (CV * 4)), D, smem};
            }
        }
        if (best < 0) continue;
        pl = cand;
        const long long rows_total = (long long)H * pl.ncol;
        int m = 2;
        while (m > 1 && rows_total / (148LL * m) < 2LL * pl.RG) --m;
        long long G = 148LL * m;
        if (G < pl.ncol) G = pl.ncol;
        const long long cap = (long long)pl.ncol * ((H + pl.RG - 1) / pl.RG);
        if (G > cap) G = cap;
        pl.G = (int)G;
        if (pl.smem < (size_t)(kRingThreads / 32) * pl.CV * 40 * sizeof(float)) pl.smem = (size_t)(kRingThreads / 32) * pl.CV * 40 * sizeof(float);
        return true;
    }
    return false;
}

template <typename T, int MODE, int ACT, bool MUL>
static cudaError_t ring_launch(const RingPlan &pl, const DwRing &rp, cudaStream_t st) {
    auto go = [&](auto k) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
        if (e != cudaSuccess) return e;
        k<<<pl.G, kRingThreads, pl.smem, st>>>(rp);
        return cudaGetLastError();
    };
    if (pl.SL == 8) return pl.CV == 8 ? go(dw3x3_ring_kernel<T, 8, 8, MODE, ACT, MUL>) : go(dw3x3_ring_kernel<T, 8, 4, MODE, ACT, MUL>);
    return pl.CV == 8 ? go(dw3x3_ring_kernel<T, 5, 8, MODE, ACT, MUL>) : go(dw3x3_ring_kernel<T, 5, 4, MODE, ACT, MUL>);
}

// ------------------------------------------------------------------ causal conv1d, x (B, C, L) row-major
// y[b,c,t] = bias[c] + sum_j w[c,j] x[b,c,t-(K-1)+j], optional SiLU.  One warp-row per (b,c); 4 steps per thread.
template <int ACT>
__global__ void __launch_bounds__(256) causal_conv1d_fwd_kernel(const float *__restrict__ x,
                                                                const float *__restrict__ w,
                                                                const float *__restrict__ bias,
                                                                float *__restrict__ y, int rows, int C, int L, int K) {
    const int row = blockIdx.y;
    const int c = row % C;
    float wk[4] = {0.f, 0.f, 0.f, 0.f};
    for (int j = 0; j < K; ++j) wk[4 - K + j] = __ldg(w + c * K + j);  // right-aligned: wk[3] multiplies x[t]
    const float bv = bias ? __ldg(bias + c) : 0.f;
    const float *xr = x + (size_t)row * L;
    float *yr = y + (size_t)row * L;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < L; t += gridDim.x * blockDim.x) {
        float acc = bv;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int tt = t - 3 + j;
            if (tt >= 0) acc = fmaf(wk[j], __ldg(xr + tt), acc);
        }
        yr[t] = ACT == 1 ? silu_f(acc) : acc;
    }
}

// backward: dz = dy * act'(z); dx[t] = sum_j w[j] dz[t + (K-1) - j]; dw[c,j] += sum_t dz[t] x[t-(K-1)+j]; db += sum dz
template <int ACT>
__global__ void __launch_bounds__(256) causal_conv1d_bwd_kernel(const float *__restrict__ x,
                                                                const float *__restrict__ w,
                                                                const float *__restrict__ bias,
                                                                const float *__restrict__ dy, float *__restrict__ dx,
                                                                float *__restrict__ dw, float *__restrict__ db,
                                                                int rows, int C, int L, int K) {
    __shared__ float red[8][5];
    const int row = blockIdx.y;
    const int c = row % C;
    float wk[4] = {0.f, 0.f, 0.f, 0.f};
    for (int j = 0; j < K; ++j) wk[4 - K + j] = __ldg(w + c * K + j);
    const float bv = bias ? __ldg(bias + c) : 0.f;
    const float *xr = x + (size_t)row * L;
    const float *gr = dy + (size_t)row * L;
    float *dxr = dx + (size_t)row * L;
    float aw[4] = {0.f, 0.f, 0.f, 0.f}, ab = 0.f;
    auto dz_at = [&](int t) -> float {  // dy[t] * act'(z[t]); 0 outside [0, L)
        if (t < 0 || t >= L) return 0.f;
        float g = __ldg(gr + t);
        if (ACT == 1) {
            float z = bv;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int tt = t - 3 + j;
                if (tt >= 0) z = fmaf(wk[j], __ldg(xr + tt), z);
            }
            g *= silu_grad(z);
        }
        return g;
    };
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < L; t += gridDim.x * blockDim.x) {
        const float g0 = dz_at(t);
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            // x[t] enters z[t + 3 - j] through wk[j]
            acc = fmaf(wk[j], j == 3 ? g0 : dz_at(t + 3 - j), acc);
            const int tt = t - 3 + j;
            if (tt >= 0) aw[j] = fmaf(g0, __ldg(xr + tt), aw[j]);
        }
        dxr[t] = acc;
        ab += g0;
    }
    // block reduction of the 5 partial sums
    float vals[5] = {aw[0], aw[1], aw[2], aw[3], ab};
#pragma unroll
    for (int k = 0; k < 5; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) vals[k] += __shfl_xor_sync(0xffffffffu, vals[k], o);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0)
        for (int k = 0; k < 5; ++k) red[warp][k] = vals[k];
    __syncthreads();
    if (threadIdx.x < 5) {
        float s = 0.f;
        for (int wdx = 0; wdx < (int)(blockDim.x >> 5); ++wdx) s += red[wdx][threadIdx.x];
        if (threadIdx.x < 4) {
            const int j = threadIdx.x - (4 - K);
            if (j >= 0) atomicAdd(dw + c * K + j, s);
        } else if (db) {
            atomicAdd(db + c, s);
        }
    }
}

// ------------------------------------------------------------------ host launchers
static DwLay dw_contig(int H, int W, int C) {
    const long long bs = (long long)H * W * C;
    return DwLay{C, bs, C, bs, C, bs, 0, C, bs};
}

template <typename T>
static cudaError_t dwconv_fwd_t(const void *x, const float *w, const float *b, const void *res, void *y, int Bn, int H,
                                int W, int C, int act, bool flip, DwLay lay, cudaStream_t st) {
    const T *xp = static_cast<const T *>(x), *rp = static_cast<const T *>(res);
    T *yp = static_cast<T *>(y);
    if (C % 4 != 0) {
        long long nb1 = ((long long)Bn * H * W * C + 255) / 256;
        if (nb1 > 148LL * 32) nb1 = 148LL * 32;
        if (act && !flip) dwconv3x3_scalar_kernel<T, 1><<<(int)nb1, 256, 0, st>>>(xp, w, b, rp, yp, nullptr, nullptr, Bn, H, W, C, 0, lay, nullptr, nullptr);
        else dwconv3x3_scalar_kernel<T, 0><<<(int)nb1, 256, 0, st>>>(xp, w, b, flip ? nullptr : rp, yp, nullptr, nullptr, Bn, H, W, C, flip ? 1 : 0, lay, nullptr, nullptr);
        return cudaGetLastError();
    }
    RingPlan pl;
    if (ring_plan(Bn, H, W, C, sizeof(T), (!flip && rp) ? 1 : 0, true, pl) && ring_aligned(x, lay.ldx, lay.bsx, sizeof(T)) &&
        ring_aligned(flip ? nullptr : res, lay.ldr, lay.bsr, sizeof(T)) && ring_aligned(y, lay.ldy, lay.bsy, sizeof(T))) {
        DwRing q{};
        q.x = x, q.a1 = flip ? nullptr : res, q.o1 = y, q.w = w, q.bias = flip ? nullptr : b;
        q.H = H, q.W = W, q.ldx = lay.ldx, q.bsx = lay.bsx, q.lda1 = lay.ldr, q.bsa1 = lay.bsr, q.ldo1 = lay.ldy, q.bso1 = lay.bsy;
        q.nchunk = pl.nchunk, q.ncol = pl.ncol, q.G = pl.G, q.RG = pl.RG, q.NS = pl.NS, q.rmul = lay.rmul, q.D = pl.D;
        if (flip) return ring_launch<T, 1, 0, false>(pl, q, st);
        return act ? ring_launch<T, 0, 1, false>(pl, q, st) : ring_launch<T, 0, 0, false>(pl, q, st);
    }
    const long long total = (long long)Bn * H * ((W + kStrip - 1) / kStrip) * (C / 4);   // one thread per strip
    const int blocks = (int)((total + 255) / 256);
    if (flip) dwconv3x3_kernel<T, true, 0, false><<<blocks, 256, 0, st>>>(xp, w, nullptr, nullptr, yp, Bn, H, W, C, lay);
    else if (lay.rmul && rp) {   // multiplicative second input: own instantiations, the others keep their code
        if (act) dwconv3x3_kernel<T, false, 1, true><<<blocks, 256, 0, st>>>(xp, w, b, rp, yp, Bn, H, W, C, lay);
        else dwconv3x3_kernel<T, false, 0, true><<<blocks, 256, 0, st>>>(xp, w, b, rp, yp, Bn, H, W, C, lay);
    } else if (act) dwconv3x3_kernel<T, false, 1, false><<<blocks, 256, 0, st>>>(xp, w, b, rp, yp, Bn, H, W, C, lay);
    else dwconv3x3_kernel<T, false, 0, false><<<blocks, 256, 0, st>>>(xp, w, b, rp, yp, Bn, H, W, C, lay);
    return cudaGetLastError();
}

// x (ldx, bsx), residual (ldr, bsr; nullable), y (ldy, bsy)
cudaError_t dwconv3x3_fwd_dispatch(const void *x, const float *w, const float *b, const void *res, void *y, int Bn, int H,
                                   int W, int C, long long ldx, long long bsx, long long ldr, long long bsr,
                                   long long ldy, long long bsy, int act, int res_mul, int dtype, cudaStream_t st) {
    const DwLay lay{ldx, bsx, ldy, bsy, ldr, bsr, res_mul ? 1 : 0, 0, 0};
    return dtype == 0 ? dwconv_fwd_t<float>(x, w, b, res, y, Bn, H, W, C, act, false, lay, st)
                      : dwconv_fwd_t<__nv_bfloat16>(x, w, b, res, y, Bn, H, W, C, act, false, lay, st);
}

template <typename T>
static cudaError_t dwconv_bwd_t(const void *x, const float *w, const float *b, const void *dy, void *dz, void *dx,
                                float *dw, float *db, int Bn, int H, int W, int C, int act, DwLay lx, long long lddx,
                                long long bsdx, const void *mulv, void *dmul, cudaStream_t st) {
    // pass 1: x (lx.ldx/bsx), dy (lx.ldr/bsr) -> dz contiguous workspace; pass 2: dz -> dx (lddx, bsdx)
    const DwLay c0 = dw_contig(H, W, C);
    const DwLay l1{lx.ldx, lx.bsx, c0.ldy, c0.bsy, lx.ldr, lx.bsr, 0, lx.ldm, lx.bsm};
    const DwLay l2{c0.ldx, c0.bsx, lddx, bsdx, c0.ldr, c0.bsr, 0, 0, 0};
    const T *mv = static_cast<const T *>(mulv);
    T *dm = static_cast<T *>(dmul);
    if (C % 4 != 0) {
        long long nb1 = ((long long)Bn * H * W * C + 255) / 256;
        if (nb1 > 148LL * 32) nb1 = 148LL * 32;
        const T *xq = static_cast<const T *>(x), *gq = static_cast<const T *>(dy);
        T *zq = static_cast<T *>(dz);
        if (act) dwconv3x3_scalar_kernel<T, 1><<<(int)nb1, 256, 0, st>>>(xq, w, b, gq, zq, dw, db, Bn, H, W, C, 2, l1, mv, dm);
        else dwconv3x3_scalar_kernel<T, 0><<<(int)nb1, 256, 0, st>>>(xq, w, b, gq, zq, dw, db, Bn, H, W, C, 2, l1, mv, dm);
        cudaError_t e1 = cudaGetLastError();
        if (e1 != cudaSuccess) return e1;
        return dwconv_fwd_t<T>(dz, w, nullptr, nullptr, dx, Bn, H, W, C, 0, true, l2, st);
    }
    RingPlan pl;
    if (ring_plan(Bn, H, W, C, sizeof(T), mulv ? 2 : 1, false, pl) && ring_aligned(x, lx.ldx, lx.bsx, sizeof(T)) &&
        ring_aligned(dy, lx.ldr, lx.bsr, sizeof(T)) && ring_aligned(mulv, lx.ldm, lx.bsm, sizeof(T)) && ring_aligned(dmul, lx.ldm, lx.bsm, sizeof(T)) &&
        ring_aligned(dz, c0.ldy, c0.bsy, sizeof(T))) {
        DwRing q{};
        q.x = x, q.a1 = dy, q.a2 = mulv, q.o1 = dz, q.o2 = dmul, q.w = w, q.bias = b, q.dw = dw, q.db = db;
        q.H = H, q.W = W, q.ldx = lx.ldx, q.bsx = lx.bsx, q.lda1 = lx.ldr, q.bsa1 = lx.bsr, q.lda2 = lx.ldm, q.bsa2 = lx.bsm;
        q.ldo1 = c0.ldy, q.bso1 = c0.bsy, q.ldo2 = lx.ldm, q.bso2 = lx.bsm;
        q.nchunk = pl.nchunk, q.ncol = pl.ncol, q.G = pl.G, q.RG = pl.RG, q.NS = pl.NS, q.D = pl.D;
        cudaError_t e0;
        if (mulv) e0 = act ? ring_launch<T, 2, 1, true>(pl, q, st) : ring_launch<T, 2, 0, true>(pl, q, st);
        else e0 = act ? ring_launch<T, 2, 1, false>(pl, q, st) : ring_launch<T, 2, 0, false>(pl, q, st);
        if (e0 != cudaSuccess) return e0;
        return dwconv_fwd_t<T>(dz, w, nullptr, nullptr, dx, Bn, H, W, C, 0, true, l2, st);
    }
    const int cv = C / 4;
    const int PP = max(1, 256 / cv);
    const int threads = cv * PP;  // <= 256 when cv <= 256
    const long long nstrips = (long long)Bn * H * ((W + kStripW - 1) / kStripW);
    int ppb = 1;   // strips per lane: keep >= ~2 blocks per SM, otherwise fewer atomics
    while (ppb < 4 && nstrips / ((long long)PP * ppb * 2) >= 148 * 2) ppb *= 2;
    const int blocks = (int)((nstrips + (long long)PP * ppb - 1) / ((long long)PP * ppb));
    const size_t smem = (size_t)PP * cv * 40 * sizeof(float);
    const T *xp = static_cast<const T *>(x), *gp = static_cast<const T *>(dy);
    T *zp = static_cast<T *>(dz);
    cudaError_t e;
    auto launch = [&](auto k) -> cudaError_t {
        cudaError_t e2 = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e2 != cudaSuccess) return e2;
        k<<<blocks, threads, smem, st>>>(xp, w, b, gp, zp, dw, db, Bn, H, W, C, ppb, l1, mv, dm);
        return cudaSuccess;
    };
    if (mv) e = act ? launch(dwconv3x3_bwd_w_kernel<T, 1, true>) : launch(dwconv3x3_bwd_w_kernel<T, 0, true>);
    else e = act ? launch(dwconv3x3_bwd_w_kernel<T, 1, false>) : launch(dwconv3x3_bwd_w_kernel<T, 0, false>);
    if (e != cudaSuccess) return e;
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    return dwconv_fwd_t<T>(dz, w, nullptr, nullptr, dx, Bn, H, W, C, 0, true, l2, st);
}

// x (ldx, bsx), dy (lddy, bsdy), dz: contiguous workspace, dx (lddx, bsdx)
cudaError_t dwconv3x3_bwd_dispatch(const void *x, const float *w, const float *b, const void *dy, void *dz, void *dx,
                                   float *dw, float *db, int Bn, int H, int W, int C, long long ldx, long long bsx,
                                   long long lddy, long long bsdy, long long lddx, long long bsdx, const void *mulv,
                                   void *dmul, long long ldm, long long bsm, int act, int dtype, cudaStream_t st) {
    const DwLay lx{ldx, bsx, 0, 0, lddy, bsdy, 0, ldm, bsm};
    return dtype == 0 ? dwconv_bwd_t<float>(x, w, b, dy, dz, dx, dw, db, Bn, H, W, C, act, lx, lddx, bsdx, mulv, dmul, st)
                      : dwconv_bwd_t<__nv_bfloat16>(x, w, b, dy, dz, dx, dw, db, Bn, H, W, C, act, lx, lddx, bsdx, mulv, dmul, st);
}

cudaError_t causal_conv1d_fwd_dispatch(const float *x, const float *w, const float *b, float *y, int rows, int C,
                                       int L, int K, int act, cudaStream_t st) {
    dim3 grid(min((L + 255) / 256, 64), rows);
    if (act) causal_conv1d_fwd_kernel<1><<<grid, 256, 0, st>>>(x, w, b, y, rows, C, L, K);
    else causal_conv1d_fwd_kernel<0><<<grid, 256, 0, st>>>(x, w, b, y, rows, C, L, K);
    return cudaGetLastError();
}

cudaError_t causal_conv1d_bwd_dispatch(const float *x, const float *w, const float *b, const float *dy, float *dx,
                                       float *dw, float *db, int rows, int C, int L, int K, int act,
                                       cudaStream_t st) {
    dim3 grid(min((L + 255) / 256, 64), rows);
    if (act) causal_conv1d_bwd_kernel<1><<<grid, 256, 0, st>>>(x, w, b, dy, dx, dw, db, rows, C, L, K);
    else causal_conv1d_bwd_kernel<0><<<grid, 256, 0, st>>>(x, w, b, dy, dx, dw, db, rows, C, L, K);
    return cudaGetLastError();
}

}  // namespace mlagg
