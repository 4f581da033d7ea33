// scan_bwd.cu -- selective scan (S6) backward for sm_100a.
// Replaces selective_scan_cuda.bwd (reference FFI shape: vmamba/csms6s.py:235-238); math: SURVEY.md App. A.2.
//
// Time tiles are visited in REVERSE order.  Inside a tile, 16-step chunks are visited in reverse; for each
// chunk the forward states are recomputed from the checkpoint the forward kernel saved (h_t and a_t live in
// registers: 128 of them), then the adjoint recurrence g_t = C_t dy_t + a_{t+1} g_{t+1} runs backwards.
// Reductions:
//   over the 16 states of a channel (du, ddelta)      -> 2-stage butterfly across the 4 lanes of the channel;
//   over the channels of a group (dB, dC)             -> 3-stage transpose-reduce across the 8 channel lanes of
//       the warp (state slots are XOR-permuted per lane so no selects are needed in the first two stages),
//       then across the W warps through per-warp shared-memory tiles, then ONE coalesced fp32 atomic per
//       (state, step) per CTA into global memory (only when the group spans several CTAs it is contended);
//   over batch and time (dA, dD, ddelta_bias)          -> registers, then one atomic per thread at the end.
#include "scan_common.cuh"

namespace mlagg {

constexpr int kAccStride = kTT + 1;  // odd: the 32 (array, state) rows a warp writes per step hit 32 banks

template <int W, int S, bool kBulk>
__global__ void __launch_bounds__(2 * W * 32, 1) scan_bwd_kernel(const ScanParams p) {
    constexpr int R = 8 * W;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *u_s = reinterpret_cast<float *>(smem_raw);  // [S][R][kRowF]  u      -> du
    float *dl_s = u_s + S * R * kRowF;                 // [S][R][kRowF]  delta  -> softplus -> ddelta
    float *dy_s = dl_s + S * R * kRowF;                // [S][R][kRowF]  dout
    float *sg_s = dy_s + S * R * kRowF;                // [S][R][kRowF]  d softplus / d x  (computed)
    float *B_s = sg_s + S * R * kRowF;                 // [S][kN][kRowF]
    float *C_s = B_s + S * kN * kRowF;                 // [S][kN][kRowF]
    float *acc_s = C_s + S * kN * kRowF;               // [W][2*kN][kAccStride]  per-warp dB | dC partials
    float *bias_s = acc_s + W * 2 * kN * kAccStride;   // [R]
    uint64_t *full = reinterpret_cast<uint64_t *>((reinterpret_cast<uintptr_t>(bias_s + R) + 7) & ~uintptr_t(7));
    uint64_t *empty = full + S;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.z, g = blockIdx.y;
    const int row0 = g * p.dpg + blockIdx.x * R;
    const int rows_valid = min(R, (g + 1) * p.dpg - row0);
    const int L = p.L;
    const int ntiles = (L + kTT - 1) / kTT;

    for (int i = threadIdx.x; i < S * (4 * R + 2 * kN) * kRowF; i += blockDim.x) u_s[i] = 0.f;
    for (int i = threadIdx.x; i < R; i += blockDim.x)
        bias_s[i] = (i < rows_valid && p.bias) ? p.bias[row0 + i] : 0.f;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], W);
            mbar_init(&empty[s], W);
        }
        mbar_fence_init();
    }
    fence_proxy_async();
    __syncthreads();

    if (warp >= W) {
        const int pw = warp - W;  // W producer warps: each issues (and accounts for) its share of the copies
        // ------------------------------------------------------------ producer warp (tiles in reverse)
        const size_t rowoff = ((size_t)b * p.dim + row0) * L;
        const float *srcs[3] = {p.u + rowoff, p.delta + rowoff, p.dout + rowoff};
        const float *Bb = p.B + ((size_t)b * p.G + g) * kN * (size_t)L;
        const float *Cb = p.C + ((size_t)b * p.G + g) * kN * (size_t)L;
        const int ncopies = 3 * rows_valid + 2 * kN;
        for (int k = 0; k < ntiles; ++k) {
            const int s = k % S;
            if (k >= S) mbar_wait(&empty[s], ((k / S) & 1) ^ 1);
            const int t0 = (ntiles - 1 - k) * kTT;
            const int nvalid = min(kTT, L - t0);
            float *own[3] = {u_s + s * R * kRowF, dl_s + s * R * kRowF, dy_s + s * R * kRowF};
            float *Bs = B_s + s * kN * kRowF, *Cs = C_s + s * kN * kRowF;
            const uint32_t bytes = nvalid * 4;
            if (kBulk) {
                int cnt = 0;
                for (int i = pw * 32; i < ncopies; i += 32 * W) cnt += min(32, ncopies - i);
                if (lane == 0) mbar_arrive_expect_tx(&full[s], bytes * cnt);
                __syncwarp();
            }
            for (int i = kBulk ? pw * 32 + lane : pw; i < ncopies; i += kBulk ? 32 * W : W) {
                const float *src;
                float *dst;
                if (i < 3 * rows_valid) {
                    const int which = i / rows_valid, rr = i % rows_valid;
                    src = srcs[which] + (size_t)rr * L;
                    dst = own[which] + rr * kRowF;
                } else if (i < 3 * rows_valid + kN) {
                    src = Bb + (size_t)(i - 3 * rows_valid) * L;
                    dst = Bs + (i - 3 * rows_valid) * kRowF;
                } else {
                    src = Cb + (size_t)(i - 3 * rows_valid - kN) * L;
                    dst = Cs + (i - 3 * rows_valid - kN) * kRowF;
                }
                if (kBulk) {
                    bulk_g2s(dst, src + t0, bytes, &full[s]);
                } else {
                    for (int t = lane; t < kTT; t += 32) dst[t] = t < nvalid ? __ldg(src + t0 + t) : 0.f;
                }
            }
            if (!kBulk) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[s]);
            }
        }
        return;
    }

    // ---------------------------------------------------------------- consumer warps
    const int r = lane >> 2, q = lane & 3;
    const int x = (r >> 1) & 3;  // slot permutation: slot s holds state n = q + 4 * (s ^ x)
    const int rl = warp * 8 + r;
    const bool valid = rl < rows_valid;
    const int d = row0 + rl;
    float A2[4], dAacc[4], gst[4], anext[4];
#pragma unroll
    for (int s4 = 0; s4 < 4; ++s4) {
        A2[s4] = valid ? p.A[(size_t)d * kN + q + 4 * (s4 ^ x)] * kLog2e : 0.f;
        dAacc[s4] = 0.f;
        gst[s4] = 0.f;
        anext[s4] = 0.f;
    }
    const float Dk = (valid && p.D) ? p.D[d] : 0.f;
    float dDacc = 0.f, dbacc = 0.f;
    const float *ck = p.ckpt_in + ((size_t)b * p.nchunks * p.dim + d) * kN + q * 4;
    const size_t ck_stride = (size_t)p.dim * kN;
    const size_t rowoff_w = ((size_t)b * p.dim + row0 + warp * 8) * L;
    float *acc_w = acc_s + warp * 2 * kN * kAccStride;
    const int acc_row = ((r & 1) * kN + q + 4 * x) * kAccStride;  // where this lane's reduced value goes

    // checkpoint prefetch: state BEFORE chunk gc is ckpt[gc - 1] (zero for gc == 0)
    auto load_ckpt = [&](int gc) -> float4 {
        if (gc <= 0 || !valid) return make_float4(0.f, 0.f, 0.f, 0.f);
        return *reinterpret_cast<const float4 *>(ck + (size_t)(gc - 1) * ck_stride);
    };
    float4 hnext = load_ckpt((L - 1) / kChunk);

    for (int k = 0; k < ntiles; ++k) {
        const int s = k % S;
        mbar_wait(&full[s], (k / S) & 1);
        const int t0 = (ntiles - 1 - k) * kTT;
        const int nvalid = min(kTT, L - t0);
        float *us = u_s + (s * R + warp * 8) * kRowF;
        float *dls = dl_s + (s * R + warp * 8) * kRowF;
        const float *dys = dy_s + (s * R + warp * 8) * kRowF;
        float *sgs = sg_s + (s * R + warp * 8) * kRowF;
        const float *Bq = B_s + (s * kN + q) * kRowF;
        const float *Cq = C_s + (s * kN + q) * kRowF;

        // delta <- softplus(delta + bias), sg <- sigmoid(delta + bias) for the warp's 8 x kTT tile
#pragma unroll
        for (int i = 0; i < (8 * kTT / 4) / 32; ++i) {
            const int idx = lane + 32 * i;
            const int rr = idx / (kTT / 4), c4 = idx % (kTT / 4);
            float4 *ptr = reinterpret_cast<float4 *>(dls + rr * kRowF + c4 * 4);
            float4 v = *ptr, sg = make_float4(1.f, 1.f, 1.f, 1.f);
            const float bb = bias_s[warp * 8 + rr];
            if (p.softplus) {
                v.x = softplus_fast(v.x + bb, &sg.x);
                v.y = softplus_fast(v.y + bb, &sg.y);
                v.z = softplus_fast(v.z + bb, &sg.z);
                v.w = softplus_fast(v.w + bb, &sg.w);
            } else {
                v.x += bb; v.y += bb; v.z += bb; v.w += bb;
            }
            *ptr = v;
            *reinterpret_cast<float4 *>(sgs + rr * kRowF + c4 * 4) = sg;
        }
        __syncwarp();

        const int nsub = (nvalid + kChunk - 1) / kChunk;
        for (int sc = nsub - 1; sc >= 0; --sc) {
            const int tb = sc * kChunk;          // first step of the chunk inside the tile
            const int ns = min(kChunk, nvalid - tb);
            const int gc = (t0 + tb) / kChunk;   // global chunk index
            const float4 h0v = hnext;
            hnext = load_ckpt(gc - 1);
            float hprev[4], h0[4];
            {
                const float hv[4] = {h0v.x, h0v.y, h0v.z, h0v.w};
#pragma unroll
                for (int s4 = 0; s4 < 4; ++s4) {  // checkpoint is stored in natural slot order
                    const int src = s4 ^ x;
                    hprev[s4] = src == 0 ? hv[0] : src == 1 ? hv[1] : src == 2 ? hv[2] : hv[3];
                    h0[s4] = hprev[s4];
                }
            }
            float hh[kChunk][4], aa[kChunk][4];
            // ---- recompute the forward states of this chunk
#pragma unroll
            for (int i4 = 0; i4 < kChunk; i4 += 4) {
                const float4 d4 = *reinterpret_cast<const float4 *>(dls + r * kRowF + tb + i4);
                const float4 u4 = *reinterpret_cast<const float4 *>(us + r * kRowF + tb + i4);
                float4 Bv[4];
#pragma unroll
                for (int s4 = 0; s4 < 4; ++s4)
                    Bv[s4] = *reinterpret_cast<const float4 *>(Bq + (s4 ^ x) * 4 * kRowF + tb + i4);
                const float dd[4] = {d4.x, d4.y, d4.z, d4.w};
                const float uu[4] = {u4.x, u4.y, u4.z, u4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const bool on = i4 + i < ns;
                    const float du = dd[i] * uu[i];
#pragma unroll
                    for (int s4 = 0; s4 < 4; ++s4) {
                        const float Bji = i == 0 ? Bv[s4].x : i == 1 ? Bv[s4].y : i == 2 ? Bv[s4].z : Bv[s4].w;
                        const float a = ex2_approx(dd[i] * A2[s4]);
                        const float hn = fmaf(a, hprev[s4], du * Bji);
                        aa[i4 + i][s4] = a;
                        hh[i4 + i][s4] = on ? hn : hprev[s4];
                        hprev[s4] = hh[i4 + i][s4];
                    }
                }
            }
            // ---- adjoint recurrence, last step of the chunk first
#pragma unroll
            for (int i4 = kChunk - 4; i4 >= 0; i4 -= 4) {
                const float4 d4 = *reinterpret_cast<const float4 *>(dls + r * kRowF + tb + i4);
                const float4 u4 = *reinterpret_cast<const float4 *>(us + r * kRowF + tb + i4);
                const float4 y4 = *reinterpret_cast<const float4 *>(dys + r * kRowF + tb + i4);
                const float4 g4 = *reinterpret_cast<const float4 *>(sgs + r * kRowF + tb + i4);
                float4 Bv[4], Cv[4];
#pragma unroll
                for (int s4 = 0; s4 < 4; ++s4) {
                    Bv[s4] = *reinterpret_cast<const float4 *>(Bq + (s4 ^ x) * 4 * kRowF + tb + i4);
                    Cv[s4] = *reinterpret_cast<const float4 *>(Cq + (s4 ^ x) * 4 * kRowF + tb + i4);
                }
                const float dd[4] = {d4.x, d4.y, d4.z, d4.w};
                const float uu[4] = {u4.x, u4.y, u4.z, u4.w};
                const float yy[4] = {y4.x, y4.y, y4.z, y4.w};
                const float ss[4] = {g4.x, g4.y, g4.z, g4.w};
                float du_out[4], dd_out[4];
#pragma unroll
                for (int i = 3; i >= 0; --i) {
                    const bool on = i4 + i < ns;  // warp-uniform
                    const float dy = on ? yy[i] : 0.f;
                    const float du = dd[i] * uu[i];
                    float s1 = 0.f, s2 = 0.f, P[4], Q[4];
#pragma unroll
                    for (int s4 = 0; s4 < 4; ++s4) {
                        const float Bji = i == 0 ? Bv[s4].x : i == 1 ? Bv[s4].y : i == 2 ? Bv[s4].z : Bv[s4].w;
                        const float Cji = i == 0 ? Cv[s4].x : i == 1 ? Cv[s4].y : i == 2 ? Cv[s4].z : Cv[s4].w;
                        const float gn = on ? fmaf(anext[s4], gst[s4], Cji * dy) : gst[s4];
                        const float ht = hh[i4 + i][s4];
                        Q[s4] = dy * ht;          // dC contribution
                        P[s4] = on ? gn * du : 0.f;  // dB contribution
                        s1 = fmaf(gn, Bji, s1);
                        const float hm1 = (i4 + i > 0) ? hh[(i4 + i > 0) ? i4 + i - 1 : 0][s4] : h0[s4];
                        const float tmp = gn * (aa[i4 + i][s4] * hm1);  // g * a_t * h_{t-1}
                        if (on) {
                            dAacc[s4] = fmaf(tmp, dd[i], dAacc[s4]);
                            s2 = fmaf(tmp, A2[s4], s2);
                            anext[s4] = aa[i4 + i][s4];
                        }
                        gst[s4] = gn;
                    }
                    if (!on) s1 = 0.f;
                    // sums over the 16 states: butterfly across the 4 lanes of the channel
                    s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
                    s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
                    s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
                    s2 += __shfl_xor_sync(0xffffffffu, s2, 2);
                    const float ddel = fmaf(s2, kLn2, uu[i] * s1);  // A = A2 * ln2
                    du_out[i] = fmaf(Dk, dy, dd[i] * s1);
                    dd_out[i] = ddel * ss[i];
                    // sums over the 8 channels of the warp (transpose-reduce, see header comment)
                    P[0] += __shfl_xor_sync(0xffffffffu, P[2], 16);
                    P[1] += __shfl_xor_sync(0xffffffffu, P[3], 16);
                    Q[0] += __shfl_xor_sync(0xffffffffu, Q[2], 16);
                    Q[1] += __shfl_xor_sync(0xffffffffu, Q[3], 16);
                    P[0] += __shfl_xor_sync(0xffffffffu, P[1], 8);
                    Q[0] += __shfl_xor_sync(0xffffffffu, Q[1], 8);
                    const bool oddr = (r & 1) != 0;
                    const float send = oddr ? P[0] : Q[0];
                    float keep = oddr ? Q[0] : P[0];
                    keep += __shfl_xor_sync(0xffffffffu, send, 4);
                    acc_w[acc_row + tb + i4 + i] = keep;
                    if (q == 0 && on) {
                        dDacc = fmaf(dy, uu[i], dDacc);
                        dbacc += dd_out[i];
                    }
                }
                if (q == 0) {
                    *reinterpret_cast<float4 *>(us + r * kRowF + tb + i4) =
                        make_float4(du_out[0], du_out[1], du_out[2], du_out[3]);
                    *reinterpret_cast<float4 *>(dls + r * kRowF + tb + i4) =
                        make_float4(dd_out[0], dd_out[1], dd_out[2], dd_out[3]);
                }
            }
        }

        // ---- tile epilogue: du / ddelta out, dB / dC reduced over the W warps and added to global
        if (kBulk) {
            fence_proxy_async();
            __syncwarp();
            if (lane < 8 && warp * 8 + lane < rows_valid) {
                bulk_s2g(p.du + rowoff_w + (size_t)lane * L + t0, us + lane * kRowF, nvalid * 4);
                bulk_s2g(p.ddelta + rowoff_w + (size_t)lane * L + t0, dls + lane * kRowF, nvalid * 4);
            }
            bulk_commit();
        } else {
            __syncwarp();
            for (int rr = 0; rr < 8 && warp * 8 + rr < rows_valid; ++rr)
                for (int t = lane; t < nvalid; t += 32) {
                    p.du[rowoff_w + (size_t)rr * L + t0 + t] = us[rr * kRowF + t];
                    p.ddelta[rowoff_w + (size_t)rr * L + t0 + t] = dls[rr * kRowF + t];
                }
        }
        named_bar_sync(1, W * 32);
        {
            float *dBg = p.dB + ((size_t)b * p.G + g) * kN * (size_t)L + t0;
            float *dCg = p.dC + ((size_t)b * p.G + g) * kN * (size_t)L + t0;
            const int tid = warp * 32 + lane;
            for (int idx = tid; idx < 2 * kN * kTT; idx += W * 32) {
                const int row = idx / kTT, t = idx % kTT;
                if (t < nvalid) {
                    float v = 0.f;
#pragma unroll
                    for (int w = 0; w < W; ++w) v += acc_s[(w * 2 * kN + row) * kAccStride + t];
                    float *dst = (row < kN ? dBg + (size_t)row * L : dCg + (size_t)(row - kN) * L) + t;
                    atomicAdd(dst, v);
                }
            }
        }
        named_bar_sync(1, W * 32);
        if (kBulk) {
            bulk_wait_read<1>();
            __syncwarp();
            if (lane == 0 && k >= 1) mbar_arrive(&empty[(k - 1) % S]);
        } else {
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
    }
    if (kBulk) bulk_wait<0>();

    // ---- parameter gradients: one atomic per thread
    if (valid) {
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) atomicAdd(p.dA + (size_t)d * kN + q + 4 * (s4 ^ x), dAacc[s4]);
        if (q == 0) {
            if (p.dD) atomicAdd(p.dD + d, dDacc);
            if (p.dbias) atomicAdd(p.dbias + d, dbacc);
        }
    }
}

template <int W, int S>
static size_t bwd_smem_bytes() {
    const size_t f = (size_t)S * (4 * 8 * W + 2 * kN) * kRowF + (size_t)W * 2 * kN * kAccStride + 8 * W;
    return f * 4 + 8 + 2 * S * 8;
}

template <int W, int S, bool kBulk>
static cudaError_t launch_bwd(const ScanParams &p, cudaStream_t st) {
    const size_t smem = bwd_smem_bytes<W, S>();
    auto kern = scan_bwd_kernel<W, S, kBulk>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid((p.dpg + 8 * W - 1) / (8 * W), p.G, p.batch);
    kern<<<grid, 2 * W * 32, smem, st>>>(p);
    return cudaGetLastError();
}

cudaError_t scan_bwd_dispatch(const ScanParams &p, bool bulk, int warps, cudaStream_t st) {
    if (warps >= 4) return bulk ? launch_bwd<4, 3, true>(p, st) : launch_bwd<4, 3, false>(p, st);
    if (warps >= 2) return bulk ? launch_bwd<2, 3, true>(p, st) : launch_bwd<2, 3, false>(p, st);
    return bulk ? launch_bwd<1, 3, true>(p, st) : launch_bwd<1, 3, false>(p, st);
}

}  // namespace mlagg
