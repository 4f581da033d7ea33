// dwconv.cu -- depthwise 3x3 convolution (pad 1, bias, optional fused SiLU) on TOKENS-MAJOR data, and the
// depthwise causal conv1d.  Replaces, without the NCHW<->tokens permute copies the reference needs,
//   nn.Conv2d(groups=C) at nnUNetTrainer_MLAgg_2D_dt_MS.py:851,890 (dwc), :680,782 (lepe);
//   MambaSkip.py:302-312,521-523 (SS2D_skip.conv2d[i] + SiLU), :545-556 (DWConv in ConvolutionalGLU);
//   nnUNetTrainer_MLLA_UNet.py:279,289 (cpe1/2), :215,248 (lepe);
//   causal_conv1d_fn of the un-vendored causal-conv1d package (north_star; not on the trainer's path, F3).
// HBM-bound stencils: every thread owns 4 consecutive channels of one pixel (128-bit fp32 / 64-bit bf16
// accesses, coalesced across the channel dimension); neighbour rows come from L1/L2.
#include <cuda_bf16.h>

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
