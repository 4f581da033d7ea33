// scan_common.cuh -- tiling constants and parameter block shared by the selective-scan kernels.
//
// Work decomposition (DESIGN.md "scan"): see the header comments of scan_fwd.cu / scan_bwd.cu.
//
// Two operand modes share the kernels (only the helper warps' address generation differs):
//   mamba interface : u, delta (B, D, L); B, C (B, G, N, L)        -- selective_scan_fn, reference MambaSkip.py:445-451
//   fused MSMM      : the 4-direction multi-scale cross-scan (MambaSkip.py:414-422), the dt projection (:434) and the
//                     fp32 casts (:437-443) are folded into the loads: direction k = group index reads channel d of
//                     x_row (k even) or x_col (k odd), both (B, Di, L), at position l for k < 2 and at the per-stage
//                     mirrored position for k >= 2; B / C / dt_rank rows come from xdbl_row / xdbl_col
//                     (B, 2, Rk + 2N, L) = W_x[k] applied to x in row- / column-major walk order; delta =
//                     W_dt[k*Di+d, :] . dts_r.  Outputs are stored at the un-mirrored position, so direction k's
//                     result is in plain row- (k even) or column-major (k odd) order.
#pragma once
#include "common.cuh"

namespace mlagg {

constexpr int kN = 16;              // d_state supported by the fast path
constexpr int kTT = 64;             // time steps per shared-memory tile (forward)
constexpr int kRowF = kTT + 4;      // floats per smem row (+16 B: bank spread, keeps 16 B alignment)
constexpr int kChunk = 16;          // checkpoint interval == MLAGG_SCAN_CHUNK
constexpr int kMaxStages = 8;       // fused mode: stages concatenated along L
constexpr int kMaxRk = 4;           // fused mode: dt_rank

struct ScanParams {
    const float *u, *delta, *A, *B, *C, *D, *bias;
    const float *dout;        // bwd
    const float *ckpt_in;     // bwd
    float *out, *ckpt, *last_state;                       // fwd outputs
    float *du, *ddelta, *dA, *dB, *dC, *dD, *dbias;       // bwd outputs
    int batch, dim, L, G, dpg, nchunks, softplus;
    // ---- fused MSMM mode (fused != 0): dim = 4 * Di, G = 4, dpg = Di
    int fused, Rk, nstage;
    int soff[kMaxStages + 1];                              // cumulative stage offsets along L
    const float *xrow, *xcol;                              // (B, Di, L)
    const float *xdbl_row, *xdbl_col;                      // (B, 2, Rk + 2N, L): directions {0,2} / {1,3}
    const float *Wdt;                                      // (4 * Di, Rk)
    float *dxdbl_row, *dxdbl_col, *dWdt;                   // bwd, accumulated into
    int dout_walks;                                        // bwd: dout is (B, 2, Di, L) in row / column walk order
};

// position read by scan position t of a mirrored direction (per-stage reversal), fused mode
__device__ __forceinline__ int mirror_pos(const ScanParams &p, int t) {
    int s = 0;
#pragma unroll
    for (int i = 1; i < kMaxStages; ++i) s += (i < p.nstage && t >= p.soff[i]) ? 1 : 0;
    return p.soff[s] + p.soff[s + 1] - 1 - t;
}

}  // namespace mlagg
