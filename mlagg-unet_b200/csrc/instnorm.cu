// instnorm.cu -- per-(image, channel) normalisation over the pixels of a CHANNELS-LAST / tokens-major (B, N, C) map,
// optional affine, optional fused LeakyReLU.  One implementation serves nn.InstanceNorm2d (reference
// MambaSkip.py:714-716 conv branches of VSS_Conv_Block; monai UnetResBlock norm1..3 of encoder0 / decoder0,
// nnUNetTrainer_MLAgg_2D_dt_MS.py:1339-1357) and nn.GroupNorm(num_groups=C) (MedNeXtBlock / PatchExpand, :262,:497).
// These sit next to the hot path, not on it; they are here because torch's instance_norm / group_norm force an NCHW
// copy of every channels_last activation (and an fp32 round trip under autocast), which made them -- not the convs --
// the largest non-GEMM cost of the decoder.  Three HBM-bound passes forward (sums, finalise, apply), three backward.
//   stats (B, C, 2) fp32: (sum, sum of squares) -> finalised in place to (mean, rstd).
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace mlagg {

__device__ __forceinline__ void in_ld4(const float *p, float (&v)[4]) {
    const float4 t = __ldg(reinterpret_cast<const float4 *>(p));
    v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
}
__device__ __forceinline__ void in_ld4(const __nv_bfloat16 *p, float (&v)[4]) {
    const uint2 t = __ldg(reinterpret_cast<const uint2 *>(p));
    v[0] = __uint_as_float(t.x << 16), v[1] = __uint_as_float(t.x & 0xffff0000u);
    v[2] = __uint_as_float(t.y << 16), v[3] = __uint_as_float(t.y & 0xffff0000u);
}
__device__ __forceinline__ void in_st4(float *p, const float (&v)[4]) {
    *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void in_st4(__nv_bfloat16 *p, const float (&v)[4]) {
    const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 raw;
    raw.x = *reinterpret_cast<const uint32_t *>(&a);
    raw.y = *reinterpret_cast<const uint32_t *>(&b);
    *reinterpret_cast<uint2 *>(p) = raw;
}

constexpr int kInWarps = 8;

// activation fused behind the affine: 0 none, 1 LeakyReLU(slope), 2 SiLU
__device__ __forceinline__ float in_act(float z, int act, float slope) {
    if (act == 1) return z < 0.f ? z * slope : z;
    if (act == 2) return silu_f(z);
    return z;
}
__device__ __forceinline__ float in_dact(float z, int act, float slope) {
    if (act == 1) return z < 0.f ? slope : 1.f;
    if (act == 2) {
        const float sg = rcp_approx(1.f + ex2_approx(-z * kLog2e));
        return sg * (1.f + z * (1.f - sg));
    }
    return 1.f;
}

// kBwd == false: acc0 = sum x, acc1 = sum x^2.
// kBwd == true : acc0 = sum dz, acc1 = sum dz * xhat, dz = dy * act'(xhat * w + b).
template <typename T, bool kBwd>
__global__ void __launch_bounds__(32 * kInWarps) in_sums_kernel(const T *__restrict__ x, const T *__restrict__ dy,
                                                                const float *__restrict__ stats_in,
                                                                const float *__restrict__ w, const float *__restrict__ b,
                                                                float *__restrict__ out, int N, int C, int rows_per_block,
                                                                int act, float slope) {
    __shared__ float red[kInWarps][2][32 * 4 + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // narrow maps (C < 128, one block column): a warp covers rpw = 32 / (C/4) rows at once instead of idling the lanes
    // beyond C/4 (C = 48 at full resolution used 12 of 32 lanes: 1.4 TB/s)
    const int cv = C >> 2;
    const int rpw = (gridDim.x == 1 && cv < 32) ? 32 / cv : 1;
    const int sub = rpw > 1 ? lane / cv : 0;                       // row inside the warp's group of rpw rows
    const int c0 = rpw > 1 ? (lane % cv) * 4 : (blockIdx.x * 32 + lane) * 4;
    const bool active = rpw > 1 ? sub < rpw : c0 < C;
    const int bi = blockIdx.z;
    const int n0 = blockIdx.y * rows_per_block, n1 = min(N, n0 + rows_per_block);
    float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};
    if (active) {
        float mean[4], rstd[4], wv[4], bv[4];
        if (kBwd) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                mean[i] = stats_in[((size_t)bi * C + c0 + i) * 2];
                rstd[i] = stats_in[((size_t)bi * C + c0 + i) * 2 + 1];
                wv[i] = w ? w[c0 + i] : 1.f;
                bv[i] = b ? b[c0 + i] : 0.f;
            }
        }
        const T *xb = x + (size_t)bi * N * C + c0;
        const T *db = kBwd ? dy + (size_t)bi * N * C + c0 : nullptr;
        for (int n = n0 + warp * rpw + sub; n < n1; n += kInWarps * rpw) {
            float v[4];
            in_ld4(xb + (size_t)n * C, v);
            if (!kBwd) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    a0[i] += v[i];
                    a1[i] = fmaf(v[i], v[i], a1[i]);
                }
            } else {
                float g[4];
                in_ld4(db + (size_t)n * C, g);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float xh = (v[i] - mean[i]) * rstd[i];
                    const float dz = g[i] * in_dact(fmaf(xh, wv[i], bv[i]), act, slope);
                    a0[i] += dz;
                    a1[i] = fmaf(dz, xh, a1[i]);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        red[warp][0][lane * 4 + i] = a0[i];
        red[warp][1][lane * 4 + i] = a1[i];
    }
    __syncthreads();
    for (int j = threadIdx.x; j < 2 * 128; j += blockDim.x) {
        const int k = j >> 7, cc = j & 127;
        const int c = blockIdx.x * 128 + cc;
        if (c < C) {
            float s = 0.f;
            for (int r = 0; r < rpw; ++r)                          // lanes r * cv + c / 4 hold this channel's partials
#pragma unroll
                for (int ww = 0; ww < kInWarps; ++ww) s += red[ww][k][cc + r * cv * 4];
            atomicAdd(out + ((size_t)bi * C + c) * 2 + k, s);
        }
    }
}

// (sum, sumsq) -> (mean, rstd)
__global__ void in_finalize_kernel(float *stats, int BC, float invN, float eps) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= BC) return;
    const float m = stats[2 * i] * invN;
    const float var = fmaxf(stats[2 * i + 1] * invN - m * m, 0.f);
    stats[2 * i] = m;
    stats[2 * i + 1] = rsqrtf(var + eps);
}
// per-channel parameter gradients: dw[c] += sum_b S2[b][c], db[c] += sum_b S1[b][c]
__global__ void in_param_grad_kernel(const float *sums, float *dw, float *db, int Bn, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float s1 = 0.f, s2 = 0.f;
    for (int bi = 0; bi < Bn; ++bi) {
        s1 += sums[((size_t)bi * C + c) * 2];
        s2 += sums[((size_t)bi * C + c) * 2 + 1];
    }
    if (dw) atomicAdd(dw + c, s2);
    if (db) atomicAdd(db + c, s1);
}

// forward: y = lrelu(xhat * w + b).  backward: dx = rstd * w * (dz - S1/N - xhat * S2/N).
template <typename T, bool kBwd>
__global__ void __launch_bounds__(256) in_apply_kernel(const T *__restrict__ x, const T *__restrict__ dy,
                                                       const float *__restrict__ stats, const float *__restrict__ sums,
                                                       const float *__restrict__ w, const float *__restrict__ b,
                                                       T *__restrict__ out, int N, int C, int act, float slope,
                                                       float invN) {
    const int cv = C >> 2;
    const size_t total = (size_t)gridDim.y * N * cv;   // gridDim.y = batch
    (void)total;
    const int bi = blockIdx.y;
    const size_t per = (size_t)N * cv;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < per; idx += (size_t)gridDim.x * blockDim.x) {
        const int c0 = (int)(idx % cv) * 4;
        const size_t off = (size_t)bi * N * C + (idx / cv) * C + c0;
        float v[4], o[4];
        in_ld4(x + off, v);
        float g[4];
        if (kBwd) in_ld4(dy + off, g);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 st = __ldg(reinterpret_cast<const float2 *>(stats) + (size_t)bi * C + c0 + i);
            const float wv = w ? __ldg(w + c0 + i) : 1.f, bv = b ? __ldg(b + c0 + i) : 0.f;
            const float xh = (v[i] - st.x) * st.y;
            const float z = fmaf(xh, wv, bv);
            if (!kBwd) {
                o[i] = in_act(z, act, slope);
            } else {
                const float2 sm = __ldg(reinterpret_cast<const float2 *>(sums) + (size_t)bi * C + c0 + i);
                const float dz = g[i] * in_dact(z, act, slope);
                o[i] = st.y * wv * (dz - sm.x * invN - xh * sm.y * invN);
            }
        }
        in_st4(out + off, o);
    }
}

// ---------------------------------------------------------------- row-streaming kernels (C % (16 / sizeof(T)) == 0)
// The kernels above spend two 64-bit divisions and twelve scalar loads on every 8-byte vector and keep one load per
// thread in flight: 25 - 35 % of the HBM roofline at the full-resolution maps (tools/call_shapes.py).  Here a thread
// owns ONE 16-byte channel vector (8 bf16 / 4 fp32 channels, per-channel constants in registers) and walks down the
// rows of its CTA's row chunk, four rows per iteration with all loads issued before the arithmetic.
template <typename T>
struct InV {
    static constexpr int N = 16 / (int)sizeof(T);
};
template <typename T>
__device__ __forceinline__ uint4 in_ldraw(const T *p) {
    return __ldg(reinterpret_cast<const uint4 *>(p));
}
__device__ __forceinline__ void in_unpack(const uint4 &t, float (&v)[4]) {
    v[0] = __uint_as_float(t.x), v[1] = __uint_as_float(t.y), v[2] = __uint_as_float(t.z), v[3] = __uint_as_float(t.w);
}
__device__ __forceinline__ void in_unpack(const uint4 &t, float (&v)[8]) {
    const uint32_t r[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[2 * i] = __uint_as_float(r[i] << 16);
        v[2 * i + 1] = __uint_as_float(r[i] & 0xffff0000u);
    }
}
__device__ __forceinline__ void in_stv(float *p, const float (&v)[4]) { in_st4(p, v); }
__device__ __forceinline__ void in_stv(__nv_bfloat16 *p, const float (&v)[8]) {
    uint32_t r[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 a = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        r[i] = *reinterpret_cast<const uint32_t *>(&a);
    }
    *reinterpret_cast<uint4 *>(p) = make_uint4(r[0], r[1], r[2], r[3]);
}
constexpr int kInRows = 4;   // rows per thread and iteration

// RES (backward only): the forward was y = lrelu(xhat w + b + residual) -- the slope of the activation is read off the
// sign of the saved OUTPUT y (a LeakyReLU keeps the sign of its argument) instead of recomputing z without the residual
template <typename T, bool kBwd, bool RES = false>
__global__ void __launch_bounds__(256) in_sums_rows_kernel(const T *__restrict__ x, const T *__restrict__ dy,
                                                           const float *__restrict__ stats_in,
                                                           const float *__restrict__ w, const float *__restrict__ b,
                                                           float *__restrict__ out, int N, int C, int rows_per_block,
                                                           int act, float slope, const T *__restrict__ yout = nullptr) {
    constexpr int V = InV<T>::N;
    extern __shared__ float red[];                  // [rpi][cvn][2 V]
    const int cvn = C / V, rpi = 256 / cvn;
    const int jc = threadIdx.x % cvn, jr = threadIdx.x / cvn;
    const int bi = blockIdx.y;
    const int n0 = blockIdx.x * rows_per_block, n1 = min(N, n0 + rows_per_block);
    float a0[V], a1[V];
#pragma unroll
    for (int i = 0; i < V; ++i) a0[i] = a1[i] = 0.f;
    if (jr < rpi) {
        const int c0 = jc * V;
        float mean[V], rstd[V], wv[V], bv[V];
        if (kBwd) {
#pragma unroll
            for (int i = 0; i < V; ++i) {
                mean[i] = stats_in[((size_t)bi * C + c0 + i) * 2];
                rstd[i] = stats_in[((size_t)bi * C + c0 + i) * 2 + 1];
                wv[i] = w ? w[c0 + i] : 1.f;
                bv[i] = b ? b[c0 + i] : 0.f;
            }
        }
        const T *xb = x + (size_t)bi * N * C + c0;
        const T *db = kBwd ? dy + (size_t)bi * N * C + c0 : nullptr;
        const T *yb = RES ? yout + (size_t)bi * N * C + c0 : nullptr;
        for (int n = n0 + jr; n < n1; n += rpi * kInRows) {
            uint4 rv[kInRows], rg[kInRows], ry[kInRows];
#pragma unroll
            for (int u = 0; u < kInRows; ++u) {
                const int nn = n + u * rpi;
                if (nn < n1) {
                    rv[u] = in_ldraw(xb + (size_t)nn * C);
                    if (kBwd) rg[u] = in_ldraw(db + (size_t)nn * C);
                    if (RES) ry[u] = in_ldraw(yb + (size_t)nn * C);
                }
            }
#pragma unroll
            for (int u = 0; u < kInRows; ++u) {
                if (n + u * rpi < n1) {
                    float v[V], g[V], yv[V];
                    in_unpack(rv[u], v);
                    if (kBwd) in_unpack(rg[u], g);
                    if (RES) in_unpack(ry[u], yv);
#pragma unroll
                    for (int i = 0; i < V; ++i) {
                        if (!kBwd) {
                            a0[i] += v[i];
                            a1[i] = fmaf(v[i], v[i], a1[i]);
                        } else {
                            const float xh = (v[i] - mean[i]) * rstd[i];
                            const float dz = RES ? g[i] * (act == 1 && !(yv[i] > 0.f) ? slope : 1.f)
                                                 : g[i] * in_dact(fmaf(xh, wv[i], bv[i]), act, slope);
                            a0[i] += dz;
                            a1[i] = fmaf(dz, xh, a1[i]);
                        }
                    }
                }
            }
        }
        float *mine = red + (size_t)threadIdx.x * 2 * V;
#pragma unroll
        for (int i = 0; i < V; ++i) mine[i] = a0[i], mine[V + i] = a1[i];
    }
    __syncthreads();
    for (int o = threadIdx.x; o < 2 * C; o += blockDim.x) {
        const int c = o >> 1, k = o & 1;
        const int qc = c / V, qi = c - qc * V;
        float sum = 0.f;
        for (int r = 0; r < rpi; ++r) sum += red[((size_t)r * cvn + qc) * 2 * V + k * V + qi];
        atomicAdd(out + ((size_t)bi * C + c) * 2 + k, sum);
    }
}

// RES forward: y = act(xhat w + b + res).  RES backward: dz from the sign of the saved output `res` (= y), dx as always,
// and dz itself -- the gradient of the residual input -- stored to `dres`.
template <typename T, bool kBwd, bool RES = false>
__global__ void __launch_bounds__(256) in_apply_rows_kernel(const T *__restrict__ x, const T *__restrict__ dy,
                                                            const float *__restrict__ stats,
                                                            const float *__restrict__ sums, const float *__restrict__ w,
                                                            const float *__restrict__ b, T *__restrict__ out, int N,
                                                            int C, int rows_per_block, int act, float slope, float invN,
                                                            const T *__restrict__ res = nullptr,
                                                            T *__restrict__ dres = nullptr) {
    constexpr int V = InV<T>::N;
    const int cvn = C / V, rpi = 256 / cvn;
    const int jc = threadIdx.x % cvn, jr = threadIdx.x / cvn;
    if (jr >= rpi) return;
    const int bi = blockIdx.y;
    const int n0 = blockIdx.x * rows_per_block, n1 = min(N, n0 + rows_per_block);
    const int c0 = jc * V;
    float mean[V], rstd[V], wv[V], bv[V], k1[V], k2[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
        const float2 st = __ldg(reinterpret_cast<const float2 *>(stats) + (size_t)bi * C + c0 + i);
        mean[i] = st.x, rstd[i] = st.y;
        wv[i] = w ? __ldg(w + c0 + i) : 1.f;
        bv[i] = b ? __ldg(b + c0 + i) : 0.f;
        if (kBwd) {
            const float2 sm = __ldg(reinterpret_cast<const float2 *>(sums) + (size_t)bi * C + c0 + i);
            k1[i] = sm.x * invN, k2[i] = sm.y * invN;
        }
    }
    const T *xb = x + (size_t)bi * N * C + c0;
    const T *db = kBwd ? dy + (size_t)bi * N * C + c0 : nullptr;
    T *ob = out + (size_t)bi * N * C + c0;
    const T *rb = RES ? res + (size_t)bi * N * C + c0 : nullptr;
    T *drb = (RES && kBwd) ? dres + (size_t)bi * N * C + c0 : nullptr;
    for (int n = n0 + jr; n < n1; n += rpi * kInRows) {
        uint4 rv[kInRows], rg[kInRows], rr[kInRows];
#pragma unroll
        for (int u = 0; u < kInRows; ++u) {
            const int nn = n + u * rpi;
            if (nn < n1) {
                rv[u] = in_ldraw(xb + (size_t)nn * C);
                if (kBwd) rg[u] = in_ldraw(db + (size_t)nn * C);
                if (RES) rr[u] = in_ldraw(rb + (size_t)nn * C);
            }
        }
#pragma unroll
        for (int u = 0; u < kInRows; ++u) {
            const int nn = n + u * rpi;
            if (nn < n1) {
                float v[V], g[V], o[V], rvv[V], dzv[V];
                in_unpack(rv[u], v);
                if (kBwd) in_unpack(rg[u], g);
                if (RES) in_unpack(rr[u], rvv);
#pragma unroll
                for (int i = 0; i < V; ++i) {
                    const float xh = (v[i] - mean[i]) * rstd[i];
                    const float z = fmaf(xh, wv[i], bv[i]) + (RES && !kBwd ? rvv[i] : 0.f);
                    if (!kBwd) {
                        o[i] = in_act(z, act, slope);
                    } else {
                        const float dz = RES ? g[i] * (act == 1 && !(rvv[i] > 0.f) ? slope : 1.f) : g[i] * in_dact(z, act, slope);
                        dzv[i] = dz;
                        o[i] = rstd[i] * wv[i] * (dz - k1[i] - xh * k2[i]);
                    }
                }
                in_stv(ob + (size_t)nn * C, o);
                if (RES && kBwd) in_stv(drb + (size_t)nn * C, dzv);
            }
        }
    }
}

static void in_grid(int N, int C, int Bn, dim3 &gs, int &rpb) {
    const int gx = (C + 127) / 128;
    int by = (148 * 8 + gx * Bn - 1) / (gx * Bn);
    rpb = (N + by - 1) / by;
    if (rpb < 64) rpb = 64;
    by = (N + rpb - 1) / rpb;
    gs = dim3(gx, by, Bn);
}

template <typename T>
static cudaError_t instnorm_run(const T *x, const T *dy, const float *w, const float *b, T *out, float *stats,
                                float *sums, float *dw, float *db, int Bn, int N, int C, float eps, int act, float slope,
                                bool bwd, cudaStream_t st, const T *res = nullptr, T *dres = nullptr) {
    dim3 gs;
    int rpb;
    in_grid(N, C, Bn, gs, rpb);
    const size_t per = (size_t)N * (C / 4);
    const dim3 ga((unsigned)std::min<size_t>((per + 255) / 256, 148 * 16), Bn);
    const float invN = 1.f / (float)N;
    constexpr int V = InV<T>::N;
    if (C % V == 0 && C / V <= 256 && !getenv("MLAGG_INSTNORM_OLD")) {
        // ~8 CTAs per SM in total; a CTA's rows are a multiple of what its threads cover per iteration
        const int rpi = 256 / (C / V);
        int chunks = std::max(1, (148 * 8 + Bn - 1) / Bn);
        int rpb = (N + chunks - 1) / chunks;
        rpb = std::max(rpi * kInRows, (rpb + rpi * kInRows - 1) / (rpi * kInRows) * (rpi * kInRows));
        chunks = (N + rpb - 1) / rpb;
        const dim3 g2(chunks, Bn);
        const size_t smem = (size_t)256 * 2 * V * sizeof(float);
        if (res != nullptr) {      // residual form: forward res = the residual input; backward res = the saved output y
            if (!bwd) {
                in_sums_rows_kernel<T, false><<<g2, 256, smem, st>>>(x, nullptr, nullptr, nullptr, nullptr, stats, N, C, rpb, act, slope);
                in_finalize_kernel<<<(Bn * C + 255) / 256, 256, 0, st>>>(stats, Bn * C, invN, eps);
                in_apply_rows_kernel<T, false, true><<<g2, 256, 0, st>>>(x, nullptr, stats, nullptr, w, b, out, N, C, rpb, act, slope, invN, res, nullptr);
            } else {
                in_sums_rows_kernel<T, true, true><<<g2, 256, smem, st>>>(x, dy, stats, w, b, sums, N, C, rpb, act, slope, res);
                if (dw || db) in_param_grad_kernel<<<(C + 127) / 128, 128, 0, st>>>(sums, dw, db, Bn, C);
                in_apply_rows_kernel<T, true, true><<<g2, 256, 0, st>>>(x, dy, stats, sums, w, b, out, N, C, rpb, act, slope, invN, res, dres);
            }
            return cudaGetLastError();
        }
        if (!bwd) {
            in_sums_rows_kernel<T, false><<<g2, 256, smem, st>>>(x, nullptr, nullptr, nullptr, nullptr, stats, N, C, rpb, act, slope);
            in_finalize_kernel<<<(Bn * C + 255) / 256, 256, 0, st>>>(stats, Bn * C, invN, eps);
            in_apply_rows_kernel<T, false><<<g2, 256, 0, st>>>(x, nullptr, stats, nullptr, w, b, out, N, C, rpb, act, slope, invN);
        } else {
            in_sums_rows_kernel<T, true><<<g2, 256, smem, st>>>(x, dy, stats, w, b, sums, N, C, rpb, act, slope);
            if (dw || db) in_param_grad_kernel<<<(C + 127) / 128, 128, 0, st>>>(sums, dw, db, Bn, C);
            in_apply_rows_kernel<T, true><<<g2, 256, 0, st>>>(x, dy, stats, sums, w, b, out, N, C, rpb, act, slope, invN);
        }
        return cudaGetLastError();
    }
    if (res != nullptr) return cudaErrorNotSupported;      // the residual form exists for the row-streaming kernels only
    if (!bwd) {
        in_sums_kernel<T, false><<<gs, 32 * kInWarps, 0, st>>>(x, nullptr, nullptr, nullptr, nullptr, stats, N, C, rpb, act, slope);
        in_finalize_kernel<<<(Bn * C + 255) / 256, 256, 0, st>>>(stats, Bn * C, invN, eps);
        in_apply_kernel<T, false><<<ga, 256, 0, st>>>(x, nullptr, stats, nullptr, w, b, out, N, C, act, slope, invN);
    } else {
        in_sums_kernel<T, true><<<gs, 32 * kInWarps, 0, st>>>(x, dy, stats, w, b, sums, N, C, rpb, act, slope);
        if (dw || db) in_param_grad_kernel<<<(C + 127) / 128, 128, 0, st>>>(sums, dw, db, Bn, C);
        in_apply_kernel<T, true><<<ga, 256, 0, st>>>(x, dy, stats, sums, w, b, out, N, C, act, slope, invN);
    }
    return cudaGetLastError();
}

cudaError_t instnorm_dispatch(const void *x, const void *dy, const float *w, const float *b, void *out, float *stats,
                              float *sums, float *dw, float *db, int Bn, int N, int C, float eps, int act, float slope,
                              int dtype, bool bwd, cudaStream_t st, const void *res, void *dres) {
    if (dtype == 0)
        return instnorm_run<float>(static_cast<const float *>(x), static_cast<const float *>(dy), w, b,
                                   static_cast<float *>(out), stats, sums, dw, db, Bn, N, C, eps, act, slope, bwd, st,
                                   static_cast<const float *>(res), static_cast<float *>(dres));
    return instnorm_run<__nv_bfloat16>(static_cast<const __nv_bfloat16 *>(x), static_cast<const __nv_bfloat16 *>(dy), w, b,
                                       static_cast<__nv_bfloat16 *>(out), stats, sums, dw, db, Bn, N, C, eps, act, slope, bwd, st,
                                       static_cast<const __nv_bfloat16 *>(res), static_cast<__nv_bfloat16 *>(dres));
}

}  // namespace mlagg
