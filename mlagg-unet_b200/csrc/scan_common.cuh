// scan_common.cuh -- tiling constants and parameter block shared by the selective-scan kernels.
//
// Work decomposition (DESIGN.md "scan"):
//   CTA        = W consumer warps + 1 producer warp; handles R = 8*W channels of ONE (batch, group) so
//                that the B/C tiles (shared by every channel of a group) are loaded once per CTA.
//   warp       = 8 channels x 4 lanes; lane (r, q) owns the four states n = q, q+4, q+8, q+12 of channel r
//                (interleaved so that the four lanes of a channel read four different smem rows of B/C ->
//                conflict-free 128-bit loads) and walks the sequence in time order with the states in
//                registers: one MUFU.EX2 + 4 FMA-pipe ops per (channel, state, step), no scan tree.
//   time tile  = kTT steps staged in shared memory by 1-D bulk async copies (TMA, UBLKCP) signalled on
//                mbarriers; S-stage ring; results leave through bulk async stores.
//   checkpoint = the running state every kChunk steps (forward writes, backward reads) so the backward
//                pass can recompute states chunk by chunk with everything in registers.
#pragma once
#include "common.cuh"

namespace mlagg {

constexpr int kN = 16;              // d_state supported by the fast path
constexpr int kTT = 64;             // time steps per shared-memory tile
constexpr int kRowF = kTT + 4;      // floats per smem row (+16 B: bank spread, keeps 16 B alignment)
constexpr int kChunk = 16;          // checkpoint interval == MLAGG_SCAN_CHUNK

struct ScanParams {
    const float *u, *delta, *A, *B, *C, *D, *bias;
    const float *dout;        // bwd
    const float *ckpt_in;     // bwd
    float *out, *ckpt, *last_state;                       // fwd outputs
    float *du, *ddelta, *dA, *dB, *dC, *dD, *dbias;       // bwd outputs
    int batch, dim, L, G, dpg, nchunks, softplus;
};

}  // namespace mlagg
