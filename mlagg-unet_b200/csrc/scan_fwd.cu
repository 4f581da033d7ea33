// scan_fwd.cu -- selective scan (S6) forward for sm_100a.
// Replaces selective_scan_cuda.fwd as called at reference MambaSkip.py:445-451 (math: SURVEY.md App. A.1).
#include <type_traits>

#include "scan_common.cuh"

namespace mlagg {

template <int W, int S, bool kBulk>
__global__ void __launch_bounds__(2 * W * 32, 1) scan_fwd_kernel(const ScanParams p) {
    constexpr int R = 8 * W;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *u_s = reinterpret_cast<float *>(smem_raw);  // [S][R][kRowF]
    float *dl_s = u_s + S * R * kRowF;                 // [S][R][kRowF]   delta -> softplus(delta) -> y
    float *B_s = dl_s + S * R * kRowF;                 // [S][kN][kRowF]
    float *C_s = B_s + S * kN * kRowF;                 // [S][kN][kRowF]
    float *bias_s = C_s + S * kN * kRowF;              // [R]
    uint64_t *full = reinterpret_cast<uint64_t *>((reinterpret_cast<uintptr_t>(bias_s + R + 16) + 7) & ~uintptr_t(7));
    uint64_t *empty = full + S;                                 // [S]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.z, g = blockIdx.y;
    const int row0 = g * p.dpg + blockIdx.x * R;
    const int rows_valid = min(R, (g + 1) * p.dpg - row0);
    const int L = p.L;
    const int ntiles = (L + kTT - 1) / kTT;

    // Zero every tile once: rows past the end of the group and tile tails are never written by the loader.
    for (int i = threadIdx.x; i < S * (2 * R + 2 * kN) * kRowF; i += blockDim.x) u_s[i] = 0.f;
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
        // ------------------------------------------------------------ producer warp
        const float *ub = p.u + ((size_t)b * p.dim + row0) * L;
        const float *db = p.delta + ((size_t)b * p.dim + row0) * L;
        const float *Bb = p.B + ((size_t)b * p.G + g) * kN * (size_t)L;
        const float *Cb = p.C + ((size_t)b * p.G + g) * kN * (size_t)L;
        const int ncopies = 2 * rows_valid + 2 * kN;
        for (int c = 0; c < ntiles; ++c) {
            const int s = c % S;
            if (c >= S) mbar_wait(&empty[s], ((c / S) & 1) ^ 1);
            const int t0 = c * kTT;
            const int nvalid = min(kTT, L - t0);
            float *us = u_s + s * R * kRowF, *ds = dl_s + s * R * kRowF;
            float *Bs = B_s + s * kN * kRowF, *Cs = C_s + s * kN * kRowF;
            if (kBulk) {
                const uint32_t bytes = nvalid * 4;
                int cnt = 0;
                for (int i = pw * 32; i < ncopies; i += 32 * W) cnt += min(32, ncopies - i);
                if (lane == 0) mbar_arrive_expect_tx(&full[s], bytes * cnt);
                __syncwarp();
                for (int i = pw * 32 + lane; i < ncopies; i += 32 * W) {
                    const float *src;
                    float *dst;
                    if (i < rows_valid) {
                        src = ub + (size_t)i * L;
                        dst = us + i * kRowF;
                    } else if (i < 2 * rows_valid) {
                        src = db + (size_t)(i - rows_valid) * L;
                        dst = ds + (i - rows_valid) * kRowF;
                    } else if (i < 2 * rows_valid + kN) {
                        src = Bb + (size_t)(i - 2 * rows_valid) * L;
                        dst = Bs + (i - 2 * rows_valid) * kRowF;
                    } else {
                        src = Cb + (size_t)(i - 2 * rows_valid - kN) * L;
                        dst = Cs + (i - 2 * rows_valid - kN) * kRowF;
                    }
                    bulk_g2s(dst, src + t0, bytes, &full[s]);
                }
            } else {
                for (int i = pw; i < ncopies; i += W) {
                    const float *src;
                    float *dst;
                    if (i < rows_valid) {
                        src = ub + (size_t)i * L;
                        dst = us + i * kRowF;
                    } else if (i < 2 * rows_valid) {
                        src = db + (size_t)(i - rows_valid) * L;
                        dst = ds + (i - rows_valid) * kRowF;
                    } else if (i < 2 * rows_valid + kN) {
                        src = Bb + (size_t)(i - 2 * rows_valid) * L;
                        dst = Bs + (i - 2 * rows_valid) * kRowF;
                    } else {
                        src = Cb + (size_t)(i - 2 * rows_valid - kN) * L;
                        dst = Cs + (i - 2 * rows_valid - kN) * kRowF;
                    }
                    for (int t = lane; t < kTT; t += 32) dst[t] = t < nvalid ? __ldg(src + t0 + t) : 0.f;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[s]);
            }
        }
        return;
    }

    // ---------------------------------------------------------------- consumer warps
    const int r = lane >> 2, q = lane & 3;
    const int rl = warp * 8 + r;
    const bool valid = rl < rows_valid;
    const int d = row0 + rl;
    float A2[4], h[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        A2[j] = valid ? p.A[(size_t)d * kN + q + 4 * j] * kLog2e : 0.f;
        h[j] = 0.f;
    }
    const float Dk = (valid && p.D) ? p.D[d] : 0.f;
    float *ck = p.ckpt ? p.ckpt + ((size_t)b * p.nchunks * p.dim + d) * kN + q * 4 : nullptr;
    const size_t ck_stride = (size_t)p.dim * kN;
    float *outw = p.out + ((size_t)b * p.dim + row0 + warp * 8) * L;

    for (int c = 0; c < ntiles; ++c) {
        const int s = c % S;
        mbar_wait(&full[s], (c / S) & 1);
        const int t0 = c * kTT;
        const int nvalid = min(kTT, L - t0);
        float *us = u_s + (s * R + warp * 8) * kRowF;
        float *dls = dl_s + (s * R + warp * 8) * kRowF;
        const float *Bq = B_s + (s * kN + q) * kRowF;
        const float *Cq = C_s + (s * kN + q) * kRowF;

        // delta <- softplus(delta + bias) for this warp's 8 x kTT tile (one float4 per lane per pass)
#pragma unroll
        for (int i = 0; i < (8 * kTT / 4) / 32; ++i) {
            const int idx = lane + 32 * i;
            const int rr = idx / (kTT / 4), c4 = idx % (kTT / 4);
            float4 *ptr = reinterpret_cast<float4 *>(dls + rr * kRowF + c4 * 4);
            float4 v = *ptr;
            const float bb = bias_s[warp * 8 + rr];
            if (p.softplus) {
                v.x = softplus_fast(v.x + bb);
                v.y = softplus_fast(v.y + bb);
                v.z = softplus_fast(v.z + bb);
                v.w = softplus_fast(v.w + bb);
            } else {
                v.x += bb; v.y += bb; v.z += bb; v.w += bb;
            }
            *ptr = v;
        }
        __syncwarp();

        // Software pipeline, written out by hand because each SM sub-partition runs ONE warp of this kernel
        // (all latency hiding must come from ILP): while the FMA chain of 4-step group g runs, the B/C tiles
        // of group g+1 are already in registers, its 16 exponentials are in flight on the MUFU pipe, and the
        // delta/u of group g+2 are being fetched.  The cross-lane reduction + store of a 16-step block is
        // issued one block late so its shuffle latency hides behind the next block's MUFU stream.
        // Reads run up to 2 groups past the tile (row padding / neighbouring rows: finite garbage, never used).
        const float *dlr = dls + r * kRowF;
        const float *ur = us + r * kRowF;
        const bool hi = (q & 2) != 0, odd = (q & 1) != 0;
        auto reduce_store = [&](const float(&y)[16], int tb) {
            // transpose-reduce over the 4 lanes of the channel: lane q ends with steps tb+4q .. tb+4q+3
            float k8[8], z[4];
#pragma unroll
            for (int m = 0; m < 8; ++m)
                k8[m] = (hi ? y[8 + m] : y[m]) + __shfl_xor_sync(0xffffffffu, hi ? y[m] : y[8 + m], 2);
#pragma unroll
            for (int m = 0; m < 4; ++m)
                z[m] = (odd ? k8[4 + m] : k8[m]) + __shfl_xor_sync(0xffffffffu, odd ? k8[m] : k8[4 + m], 1);
            const float4 uq = *reinterpret_cast<const float4 *>(ur + tb + 4 * q);
            *reinterpret_cast<float4 *>(dls + r * kRowF + tb + 4 * q) =
                make_float4(fmaf(Dk, uq.x, z[0]), fmaf(Dk, uq.y, z[1]), fmaf(Dk, uq.z, z[2]), fmaf(Dk, uq.w, z[3]));
        };
        auto ld4 = [](const float *ptr) { return *reinterpret_cast<const float4 *>(ptr); };
        // pipeline registers
        float4 dc = ld4(dlr), uc = ld4(ur);          // delta / u of the group being consumed
        float4 dn = ld4(dlr + 4), un = ld4(ur + 4);  // ... of the next group
        float4 Bc[4], Cc[4];
        float ac[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            Bc[j] = ld4(Bq + j * 4 * kRowF);
            Cc[j] = ld4(Cq + j * 4 * kRowF);
        }
        {
            const float dd[4] = {dc.x, dc.y, dc.z, dc.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) ac[i][j] = ex2_approx(dd[i] * A2[j]);
        }
        float yprev[16];
        for (int tb = 0; tb < nvalid; tb += 16) {
            const int ns = nvalid - tb;  // >= 16 except in the last block of the sequence
            float y[16];
            auto block = [&](auto full_tag) {
                constexpr bool kFull = decltype(full_tag)::value;
#pragma unroll
                for (int i4 = 0; i4 < 16; i4 += 4) {
                    const int t = tb + i4;
                    // stage 1: operands of group g+1 (B, C) and g+2 (delta, u)
                    float4 Bn[4], Cn[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        Bn[j] = ld4(Bq + j * 4 * kRowF + t + 4);
                        Cn[j] = ld4(Cq + j * 4 * kRowF + t + 4);
                    }
                    const float4 dn2 = ld4(dlr + t + 8), un2 = ld4(ur + t + 8);
                    // stage 2: exponentials of group g+1
                    float an[4][4];
                    {
                        const float dd[4] = {dn.x, dn.y, dn.z, dn.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int j = 0; j < 4; ++j) an[i][j] = ex2_approx(dd[i] * A2[j]);
                    }
                    // stage 3: recurrence of group g
                    const float dd[4] = {dc.x, dc.y, dc.z, dc.w};
                    const float uu[4] = {uc.x, uc.y, uc.z, uc.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const bool on = kFull || (kBulk ? (i4 < ns) : (i4 + i < ns));
                        const float du = dd[i] * uu[i];
                        float acc = 0.f;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float Bji = i == 0 ? Bc[j].x : i == 1 ? Bc[j].y : i == 2 ? Bc[j].z : Bc[j].w;
                            const float Cji = i == 0 ? Cc[j].x : i == 1 ? Cc[j].y : i == 2 ? Cc[j].z : Cc[j].w;
                            const float hn = fmaf(ac[i][j], h[j], du * Bji);
                            h[j] = on ? hn : h[j];
                            acc = fmaf(Cji, h[j], acc);
                        }
                        y[i4 + i] = on ? acc : 0.f;
                    }
                    // rotate
                    dc = dn; uc = un; dn = dn2; un = un2;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        Bc[j] = Bn[j];
                        Cc[j] = Cn[j];
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) ac[i][j] = an[i][j];
                }
            };
            if (ns >= 16) block(std::true_type{}); else block(std::false_type{});
            if (ck != nullptr && valid && ns >= 16)
                *reinterpret_cast<float4 *>(ck + (size_t)((t0 + tb) / kChunk) * ck_stride) =
                    make_float4(h[0], h[1], h[2], h[3]);
            if (tb > 0) reduce_store(yprev, tb - 16);
#pragma unroll
            for (int i = 0; i < 16; ++i) yprev[i] = y[i];
        }
        reduce_store(yprev, ((nvalid - 1) / 16) * 16);

        if (kBulk) {
            fence_proxy_async();
            __syncwarp();
            if (lane < 8 && warp * 8 + lane < rows_valid)
                bulk_s2g(outw + (size_t)lane * L + t0, dls + lane * kRowF, nvalid * 4);
            bulk_commit();
            bulk_wait_read<1>();  // the store issued one tile ago has finished reading its stage
            __syncwarp();
            if (lane == 0 && c >= 1) mbar_arrive(&empty[(c - 1) % S]);
        } else {
            __syncwarp();
            for (int rr = 0; rr < 8 && warp * 8 + rr < rows_valid; ++rr)
                for (int t = lane; t < nvalid; t += 32) outw[(size_t)rr * L + t0 + t] = dls[rr * kRowF + t];
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
    }
    if (kBulk) bulk_wait<0>();
    if (p.last_state && valid) {
#pragma unroll
        for (int j = 0; j < 4; ++j) p.last_state[((size_t)b * p.dim + d) * kN + q + 4 * j] = h[j];
    }
}

template <int W, int S>
static size_t fwd_smem_bytes() {
    return (size_t)S * (2 * 8 * W + 2 * kN) * kRowF * 4 + 8 * W * 4 + 64 + 8 + 2 * S * 8;
}

template <int W, int S, bool kBulk>
static cudaError_t launch_fwd(const ScanParams &p, cudaStream_t st) {
    const size_t smem = fwd_smem_bytes<W, S>();
    auto kern = scan_fwd_kernel<W, S, kBulk>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid((p.dpg + 8 * W - 1) / (8 * W), p.G, p.batch);
    kern<<<grid, 2 * W * 32, smem, st>>>(p);
    return cudaGetLastError();
}

cudaError_t scan_fwd_dispatch(const ScanParams &p, bool bulk, int warps, cudaStream_t st) {
    if (warps >= 4) return bulk ? launch_fwd<4, 4, true>(p, st) : launch_fwd<4, 4, false>(p, st);
    if (warps >= 2) return bulk ? launch_fwd<2, 4, true>(p, st) : launch_fwd<2, 4, false>(p, st);
    return bulk ? launch_fwd<1, 4, true>(p, st) : launch_fwd<1, 4, false>(p, st);
}

}  // namespace mlagg
