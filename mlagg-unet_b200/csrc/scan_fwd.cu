// scan_fwd.cu -- selective scan (S6) forward for sm_100a.
// Replaces selective_scan_cuda.fwd as called at reference MambaSkip.py:445-451 (math: SURVEY.md App. A.1).
//
// CTA = 4 scan warps + 4 helper warps (one of each per SM sub-partition), 32 channels of one (batch, group).
//   helper warp h : streams its 8 channels' u / delta rows and the B / C rows {h, h+4, h+8, h+12} from HBM with
//                   coalesced loads that are issued one 64-step tile AHEAD and parked in registers; then runs
//                   softplus and writes the tile into shared memory in the layout the scan warps consume --
//                        pk[t][row]  = (delta, delta, delta*u, delta*u)      two ready-made f32x2 pairs
//                        bc[t][q]    = (B[q][t], B[q+4][t], B[q+8][t], B[q+12][t]),   bc[t][4+q] likewise for C
//                        yt[row][t]  = D[row] * u[row][t]
//                   and, one tile later, writes the finished yt rows back to HBM (coalesced).
//   scan warp w   : 8 channels x 4 lanes; lane (r, q) owns states q, q+4, q+8, q+12 of channel r and walks time
//                   in order with the state in registers.  Per step and lane: 3 LDS.128, 4 MUFU.EX2 and 8 packed
//                   FFMA2/FMUL2 -- the exponential (16/clk/SM) is the designed limiter, not issue slots or HBM.
//                   The loop is software-pipelined by hand (operands 3 steps ahead, exponentials 2 steps ahead)
//                   because only one scan warp runs per sub-partition: latency must be hidden by ILP.
// Hand-off: SP-stage ring in shared memory; mbarriers `ready` (4 helper arrivals) and `sdone` (4 scan arrivals).
// (Rounds 1's first versions staged raw tiles with 1-D bulk async copies; 96 x 256 B copies per tile serialised in
//  the issuing warp and bound the kernel at 1.8 ms -- see profiles/README.md.)
#include <type_traits>

#include "scan_common.cuh"

namespace mlagg {

// Operand tiles are TIME-major: one 128-byte line per step holds the eight 16-byte chunks a scan warp reads in that
// step (pk: one chunk per channel; bc: B chunks q = 0..3 then C chunks q = 0..3).  Measured on B200
// (tools/micro/smem_bench.cu): an LDS.128 whose 32 lanes fall into ONE line costs ~3.3 cycles of the shared-memory
// pipe however the lanes share chunks, while 8 (4) chunks in 8 (4) different lines cost 8 (5.1) -- the [row][t] layout
// of the previous version spent 73 of its ~80 cycles per step there.  Chunk c of step t sits at c ^ (t & 7), so the
// helper warps' stores (lane = step, fixed chunk) are conflict-free as well.
constexpr int kTS = kTT + 4;  // lines per tile: 64 steps + read-ahead padding of the 3-deep software pipeline

// W = scan warps (= helper warps) per CTA, 8 channels each.  W = 4 is the training shape (32 channels of one (batch,
// group) share the staged B / C tile).  W = 1 is the small-batch / inference shape (sliding-window tiles
// arrive at B = 1 .. 4): one scan warp + one helper warp per CTA, 4x the CTAs -- at B = 1 the W = 4 grid is 12 CTAs on
// 148 SMs -- and the 60 KiB of shared memory lets three of them share an SM.
template <int W_>
struct FwdCfg {
    static constexpr int W = W_, R = 8 * W_, SP = 3;
    static constexpr size_t bytes = (size_t)SP * W * kTS * 128 + (size_t)SP * kTS * 128 +
                                    (size_t)SP * R * kRowF * 4 + 2 * R * 4 + 64 + 2 * SP * 8 + 16;
};

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float ldg_stream(const float *p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

template <bool kFused, int kW>
__global__ void __launch_bounds__(64 * kW, 1) scan_fwd_kernel(const ScanParams p) {
    constexpr int W = FwdCfg<kW>::W, R = FwdCfg<kW>::R, SP = FwdCfg<kW>::SP;
    constexpr int NQ = 4 / W;      // state quads (B / C chunk pairs) staged by one helper warp
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4 *pk = reinterpret_cast<float4 *>(smem_raw);         // [SP][W][kTS][8 chunks]
    float4 *bc = pk + SP * W * kTS * 8;                        // [SP][kTS][8 chunks]
    float *yt = reinterpret_cast<float *>(bc + SP * kTS * 8);  // [SP][R][kRowF]
    float *bias_s = yt + SP * R * kRowF;                       // [R]
    float *D_s = bias_s + R;                                   // [R]
    uint64_t *ready = reinterpret_cast<uint64_t *>((reinterpret_cast<uintptr_t>(D_s + R + 16) + 7) & ~uintptr_t(7));
    uint64_t *sdone = ready + SP;                              // [SP]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.z, g = blockIdx.y;
    const int row0 = g * p.dpg + blockIdx.x * R;
    const int rows_valid = min(R, (g + 1) * p.dpg - row0);
    const int L = p.L;
    const int ntiles = (L + kTT - 1) / kTT;

    for (int i = threadIdx.x; i < (int)((reinterpret_cast<unsigned char *>(bias_s) - smem_raw) / 4); i += blockDim.x)
        reinterpret_cast<float *>(smem_raw)[i] = 0.f;
    for (int i = threadIdx.x; i < R; i += blockDim.x) {
        bias_s[i] = (i < rows_valid && p.bias) ? p.bias[row0 + i] : 0.f;
        D_s[i] = (i < rows_valid && p.D) ? p.D[row0 + i] : 0.f;
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < SP; ++s) {
            mbar_init(&ready[s], W);
            mbar_init(&sdone[s], W);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp >= W) {
        // =============================================================== helper warp
        const int h = warp - W;
        const int myrows = max(0, min(8, rows_valid - 8 * h));
        // operand sources: mamba interface or fused MSMM (see scan_common.cuh)
        constexpr bool fused = kFused;
        const int kdir = g;                                   // fused: direction = group
        const int C35 = p.Rk + 2 * kN;
        const int dloc0 = blockIdx.x * R + 8 * h;             // first channel of this warp inside the group
        const float *ub, *db = nullptr, *Bb, *Cb, *dtb = nullptr;
        if (fused) {
            ub = ((kdir & 1) ? p.xcol : p.xrow) + ((size_t)b * p.dpg + dloc0) * L;
            const float *xd = ((kdir & 1) ? p.xdbl_col : p.xdbl_row) + (((size_t)b * 2 + (kdir >> 1)) * C35) * (size_t)L;
            dtb = xd;
            Bb = xd + (size_t)p.Rk * L;
            Cb = xd + (size_t)(p.Rk + kN) * L;
        } else {
            ub = p.u + ((size_t)b * p.dim + row0 + 8 * h) * L;
            db = p.delta + ((size_t)b * p.dim + row0 + 8 * h) * L;
            Bb = p.B + (((size_t)b * p.G + g) * kN) * (size_t)L;   // quad q = rows q, q+4, q+8, q+12
            Cb = p.C + (((size_t)b * p.G + g) * kN) * (size_t)L;
        }
        const bool mirrored = fused && kdir >= 2;
        float *outb = p.out + ((size_t)b * p.dim + row0 + 8 * h) * L;
        float wdt[8][kMaxRk];                                 // fused: rows of W_dt for this warp's channels
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int rr = 0; rr < kMaxRk; ++rr)
                wdt[i][rr] = (fused && i < myrows && rr < p.Rk) ? p.Wdt[(size_t)(row0 + 8 * h + i) * p.Rk + rr] : 0.f;
        // the tile in flight: element i = 2 * row + half  ->  step lane + 32 * half.  Loaded values are NOT touched until the
        // next iteration (one tile of scan time later), so the HBM latency is never exposed.
        float ur[16], dr[16], Br[NQ][8], Cr[NQ][8], dtr[2][kMaxRk];   // helper h stages the quads h, h + W, ...

        auto fetch = [&](int c) {
            const int t0 = c * kTT;
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                const int t = t0 + lane + 32 * hf;
                const bool tin = t < L;
                const int tm = (mirrored && tin) ? mirror_pos(p, t) : t;
                if constexpr (fused) {
#pragma unroll
                    for (int rr = 0; rr < kMaxRk; ++rr)
                        dtr[hf][rr] = (tin && rr < p.Rk) ? ldg_stream(dtb + (size_t)rr * L + tm) : 0.f;
                }
#pragma unroll
                for (int rw = 0; rw < 8; ++rw) {
                    const int i = 2 * rw + hf;
                    const bool ok = rw < myrows && tin;
                    ur[i] = ok ? ldg_stream(ub + (size_t)rw * L + tm) : 0.f;
                    if constexpr (!fused) dr[i] = ok ? ldg_stream(db + (size_t)rw * L + tm) : 0.f;
                }
#pragma unroll
                for (int iq = 0; iq < NQ; ++iq)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int rowq = h + W * iq + 4 * j;
                        Br[iq][2 * j + hf] = tin ? ldg_stream(Bb + (size_t)rowq * L + tm) : 0.f;
                        Cr[iq][2 * j + hf] = tin ? ldg_stream(Cb + (size_t)rowq * L + tm) : 0.f;
                    }
            }
        };
        // per-row constants in registers (shared-memory reads here were dependent-load stalls on the helper's path)
        float bias_r[8], D_r[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            bias_r[i] = (i < myrows && p.bias) ? p.bias[row0 + 8 * h + i] : 0.f;
            D_r[i] = (i < myrows && p.D) ? p.D[row0 + 8 * h + i] : 0.f;
        }
        auto write_out = [&](int c) {  // yt rows of tile c -> global, coalesced; all loads first, then all stores
            const int sp = c % SP, t0 = c * kTT;
            const float *ys = yt + (sp * R + 8 * h) * kRowF;
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = ys[(i >> 1) * kRowF + lane + 32 * (i & 1)];
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                const int t = t0 + lane + 32 * hf;
                if (t < L) {
                    const int tm = mirrored ? mirror_pos(p, t) : t;
#pragma unroll
                    for (int rw = 0; rw < 8; ++rw)
                        if (rw < myrows) outb[(size_t)rw * L + tm] = v[2 * rw + hf];
                }
            }
        };

        // The helper runs up to SP - 1 tiles ahead of the scan warps: tile c is produced as soon as the scan warps
        // have released its ring slot (tile c - SP), whose finished y rows are written out first.
        fetch(0);
        for (int c = 0; c < ntiles; ++c) {
            const int sp = c % SP;
            if (c >= SP) {
                mbar_wait(&sdone[sp], ((c / SP) & 1) ^ 1);   // scan warps are done with tile c - SP
                write_out(c - SP);
            }
            float4 *pks = pk + (sp * W + h) * kTS * 8;             // feeds scan warp h (same 8 channels)
            float *ys = yt + (sp * R + 8 * h) * kRowF;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int rr = i >> 1, tl = lane + 32 * (i & 1);
                float draw;
                if constexpr (fused) {
                    draw = 0.f;
#pragma unroll
                    for (int k2 = 0; k2 < kMaxRk; ++k2) draw = fmaf(wdt[rr][k2], dtr[i & 1][k2], draw);
                } else {
                    draw = dr[i];
                }
                float dl = draw + bias_r[rr];
                if (p.softplus) dl = softplus_fast(dl);
                const float du = dl * ur[i];
                pks[tl * 8 + (rr ^ (tl & 7))] = make_float4(dl, dl, du, du);
                ys[rr * kRowF + tl] = D_r[rr] * ur[i];
            }
            {
                float4 *bcs = bc + sp * kTS * 8;
#pragma unroll
                for (int iq = 0; iq < NQ; ++iq)
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const int tl = lane + 32 * k, qd = h + W * iq;
                        bcs[tl * 8 + (qd ^ (tl & 7))] = make_float4(Br[iq][k], Br[iq][2 + k], Br[iq][4 + k], Br[iq][6 + k]);
                        bcs[tl * 8 + ((4 + qd) ^ (tl & 7))] = make_float4(Cr[iq][k], Cr[iq][2 + k], Cr[iq][4 + k], Cr[iq][6 + k]);
                    }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&ready[sp]);
            if (c + 1 < ntiles) fetch(c + 1);   // in flight while this warp waits for a free slot / writes y out
        }
        for (int c = max(0, ntiles - SP); c < ntiles; ++c) {
            mbar_wait(&sdone[c % SP], (c / SP) & 1);
            write_out(c);
        }
        return;
    }

    // =================================================================== scan warp
    const int r = lane >> 2, q = lane & 3;
    const int rl = warp * 8 + r;
    const bool valid = rl < rows_valid;
    const int d = row0 + rl;
    float2 A01, A23, h01 = f2(0.f, 0.f), h23 = f2(0.f, 0.f);
    A01.x = valid ? p.A[(size_t)d * kN + q] * kLog2e : 0.f;
    A01.y = valid ? p.A[(size_t)d * kN + q + 4] * kLog2e : 0.f;
    A23.x = valid ? p.A[(size_t)d * kN + q + 8] * kLog2e : 0.f;
    A23.y = valid ? p.A[(size_t)d * kN + q + 12] * kLog2e : 0.f;
    float *ck = p.ckpt ? p.ckpt + ((size_t)b * p.nchunks * p.dim + d) * kN + q * 4 : nullptr;
    const size_t ck_stride = (size_t)p.dim * kN;
    const bool hi = (q & 2) != 0, odd = (q & 1) != 0;

    for (int c = 0; c < ntiles; ++c) {
        const int sp = c % SP;
        mbar_wait(&ready[sp], (c / SP) & 1);
        const int t0 = c * kTT;
        const int nvalid = min(kTT, L - t0);
        const float4 *pkw = pk + (sp * W + warp) * kTS * 8;
        const float4 *bcw = bc + sp * kTS * 8;
        // operands of tile step t (t & 7 is a compile-time constant wherever this is used)
        auto ldP = [&](int t) { return pkw[t * 8 + (r ^ (t & 7))]; };
        auto ldB = [&](int t) { return bcw[t * 8 + (q ^ (t & 7))]; };
        auto ldC = [&](int t) { return bcw[t * 8 + ((4 + q) ^ (t & 7))]; };
        float *yr = yt + (sp * R + rl) * kRowF;

        auto reduce_store = [&](const float(&y)[16], int tb) {
            // transpose-reduce over the 4 lanes of the channel: lane q ends with steps tb+4q .. tb+4q+3
            float k8[8], z[4];
#pragma unroll
            for (int m = 0; m < 8; ++m)
                k8[m] = (hi ? y[8 + m] : y[m]) + __shfl_xor_sync(0xffffffffu, hi ? y[m] : y[8 + m], 2);
#pragma unroll
            for (int m = 0; m < 4; ++m)
                z[m] = (odd ? k8[4 + m] : k8[m]) + __shfl_xor_sync(0xffffffffu, odd ? k8[m] : k8[4 + m], 1);
            float4 *dst = reinterpret_cast<float4 *>(yr + tb + 4 * q);
            const float4 du = *dst;  // D * u, written by the helper
            *dst = make_float4(du.x + z[0], du.y + z[1], du.z + z[2], du.w + z[3]);
        };

        // 3-deep software pipeline over steps: operands of step t+3 loaded, exponentials of step t+2 issued, FMA chain
        // of step t executed.  Reads run up to 3 steps past the tile (row padding: zeros / stale, never used).
        float4 P0 = ldP(0), B0 = ldB(0), C0 = ldC(0);
        float4 P1 = ldP(1), B1 = ldB(1), C1 = ldC(1);
        float4 P2 = ldP(2), B2 = ldB(2), C2 = ldC(2);
        float2 ea01, ea23, eb01, eb23;   // exponentials of step t (ea) and t+1 (eb)
        {
            const float2 x01 = __fmul2_rn(f2(P0.x, P0.y), A01), x23 = __fmul2_rn(f2(P0.x, P0.y), A23);
            ea01 = f2(ex2_approx(x01.x), ex2_approx(x01.y)); ea23 = f2(ex2_approx(x23.x), ex2_approx(x23.y));
            const float2 y01 = __fmul2_rn(f2(P1.x, P1.y), A01), y23 = __fmul2_rn(f2(P1.x, P1.y), A23);
            eb01 = f2(ex2_approx(y01.x), ex2_approx(y01.y)); eb23 = f2(ex2_approx(y23.x), ex2_approx(y23.y));
        }
        float yprev[16];
        for (int tb = 0; tb < nvalid; tb += 16) {
            const int ns = nvalid - tb;
            float y[16];
            auto block = [&](auto full_tag) {
                constexpr bool kFull = decltype(full_tag)::value;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float4 P3 = ldP(tb + i + 3), B3 = ldB(tb + i + 3), C3 = ldC(tb + i + 3);
                    const float2 x01 = __fmul2_rn(f2(P2.x, P2.y), A01), x23 = __fmul2_rn(f2(P2.x, P2.y), A23);
                    const float2 ec01 = f2(ex2_approx(x01.x), ex2_approx(x01.y));
                    const float2 ec23 = f2(ex2_approx(x23.x), ex2_approx(x23.y));
                    const bool on = kFull || (i < ns);
                    const float2 bu01 = __fmul2_rn(f2(P0.z, P0.w), f2(B0.x, B0.y));
                    const float2 bu23 = __fmul2_rn(f2(P0.z, P0.w), f2(B0.z, B0.w));
                    const float2 n01 = __ffma2_rn(ea01, h01, bu01), n23 = __ffma2_rn(ea23, h23, bu23);
                    if (on) { h01 = n01; h23 = n23; }
                    float2 acc = __fmul2_rn(f2(C0.x, C0.y), h01);
                    acc = __ffma2_rn(f2(C0.z, C0.w), h23, acc);
                    y[i] = on ? acc.x + acc.y : 0.f;
                    P0 = P1; B0 = B1; C0 = C1; P1 = P2; B1 = B2; C1 = C2; P2 = P3; B2 = B3; C2 = C3;
                    ea01 = eb01; ea23 = eb23; eb01 = ec01; eb23 = ec23;
                }
            };
            if (ns >= 16) block(std::true_type{}); else block(std::false_type{});
            if (ck != nullptr && valid && ns >= 16)
                *reinterpret_cast<float4 *>(ck + (size_t)((t0 + tb) / kChunk) * ck_stride) =
                    make_float4(h01.x, h01.y, h23.x, h23.y);
            if (tb > 0) reduce_store(yprev, tb - 16);
#pragma unroll
            for (int i = 0; i < 16; ++i) yprev[i] = y[i];
        }
        reduce_store(yprev, ((nvalid - 1) / 16) * 16);
        __syncwarp();
        if (lane == 0) mbar_arrive(&sdone[sp]);
    }
    if (p.last_state && valid) {
        float *ls = p.last_state + ((size_t)b * p.dim + d) * kN + q;
        ls[0] = h01.x; ls[4] = h01.y; ls[8] = h23.x; ls[12] = h23.y;
    }
}

template <int kW>
static cudaError_t scan_fwd_launch(const ScanParams &p, cudaStream_t st) {
    const size_t smem = FwdCfg<kW>::bytes;
    auto kern = p.fused ? scan_fwd_kernel<true, kW> : scan_fwd_kernel<false, kW>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid((p.dpg + 8 * kW - 1) / (8 * kW), p.G, p.batch);
    kern<<<grid, 64 * kW, smem, st>>>(p);
    return cudaGetLastError();
}

// `warps`: 0 = choose (one scan warp per CTA when the 4-warp grid would leave more than half of the SMs empty), else 1 / 4
cudaError_t scan_fwd_dispatch(const ScanParams &p, bool bulk, int warps, cudaStream_t st) {
    (void)bulk;
    if (warps != 1 && warps != 4) {
        const long long ctas4 = (long long)((p.dpg + 31) / 32) * p.G * p.batch;
        warps = ctas4 < 74 ? 1 : 4;
    }
    return warps == 1 ? scan_fwd_launch<1>(p, st) : scan_fwd_launch<4>(p, st);
}

}  // namespace mlagg
