// walk.cu -- the two layout changes around the MSMM selective scan as single passes:
//   pack   : tokens-major (B, L, C) activations  ->  channels-major fp32 planes (B, nc, L) in WALK order
//   unpack : channels-major fp32 planes (B, nc, L) in walk order (optionally the sum of two planes)  ->  tokens-major
// L is the stage-concatenated sequence (fine -> coarse); a walk is row-major (position p reads token p) or, per stage
// (H, W), column-major (position soff + q reads token soff + (q % H) * W + q / H) -- the index maps of the reference's
// cross-scan / cross-merge (variants/mamba/MambaSkip.py:414-422 and :454-471); the mirrored directions 2, 3 are handled
// inside the scan kernels.  These kernels replace, per train step, the transpose().contiguous(), per-stage
// reshape/transpose/cat, fp32 cast, direction-sum and slice-backward (zero-fill + copy + add) kernels torch ran for
// `SS2D_skip.forward_corev0` and its autograd graph: ~40 launches and ~5 ms -> 13 launches.
// HBM-bound tile transposes through shared memory (tile shapes below): token rows are read / written as 8- or 16-byte
// pieces per thread, planes as 64- or 128-byte row segments.
#include <cuda_bf16.h>

#include "common.cuh"

namespace mlagg {

constexpr int kWalkMaxStages = 8;
// 1-D tiles (row walk; the column walk uses the 2-D tiles further down): TC channels x TP positions, TC * TP = 2048.
// Shipped: (32, 64).  Measured at the config-3 shape (B = 10,
// L = 34 000, 96 channels, bf16 tokens -> fp32 planes), row / column walk: pack 39 / 150 us, unpack (two planes summed)
// 86 / 176 us.  The column walk is slower because consecutive positions are W tokens apart (every 64-byte token piece sits
// in a different DRAM page).  The (128, 16) shape -- a warp moves one whole token row -- was tried to fix that and is
// slower on BOTH walks (54 / 202 and 122 / 249 us): the fp32 plane side carries twice the bytes of the bf16 token side and
// its rows shrink to 64-byte segments.  The templates keep the TC parameter; the dispatchers use 32.

// shared-memory row of tile channel cl: identity for TC = 32 (conflict-free with 8 threads per position); for TC = 128 a
// warp writes the 4 x 32 channels of one position, so channel 4 cg + j goes to row 32 j + cg (odd row stride 17)
template <int TC>
__device__ __forceinline__ int walk_row(int cl) {
    return TC == 32 ? cl : ((cl & 3) * 32 + (cl >> 2));
}

struct WalkGeom {
    int nstage;
    int soff[kWalkMaxStages + 1];
    int H[kWalkMaxStages], W[kWalkMaxStages];
    int toff[kWalkMaxStages + 1];    // cumulative count of kColTI x kColTJ pixel tiles per stage (column walk)
};

// Column walk, 2-D tiles.  In column-major order position soff + j * H + i visits token soff + i * W + j, so a run of
// consecutive positions is a column of the image (tokens W apart: every token piece in another DRAM page -- the 1-D tiles
// above ran the column walk 4x slower than the row walk).  A tile of kColTI rows x kColTJ columns has both sides in runs:
// the token side reads / writes kColTJ consecutive tokens per image row, the plane side kColTI consecutive positions per
// image column.
constexpr int kColTI = 32, kColTJ = 8, kColTC = 32;
constexpr int kColCS = kColTJ * (kColTI + 1) + 1;   // floats per channel in shared memory: = 1 mod 8 keeps both phases conflict-free

struct ColTile {
    int soff, H, W, i0, j0;
};
__device__ __forceinline__ ColTile col_tile(const WalkGeom &g, int tile) {
    int s = 0;
#pragma unroll
    for (int i = 1; i < kWalkMaxStages; ++i) s += (i < g.nstage && tile >= g.toff[i]) ? 1 : 0;
    const int lt = tile - g.toff[s];
    const int ntj = (g.W[s] + kColTJ - 1) / kColTJ;
    return ColTile{g.soff[s], g.H[s], g.W[s], (lt / ntj) * kColTI, (lt % ntj) * kColTJ};
}

__device__ __forceinline__ int walk_token(const WalkGeom &g, int p, int col) {
    if (!col) return p;
    int s = 0;
#pragma unroll
    for (int i = 1; i < kWalkMaxStages; ++i) s += (i < g.nstage && p >= g.soff[i]) ? 1 : 0;
    const int q = p - g.soff[s], H = g.H[s];
    return g.soff[s] + (q % H) * g.W[s] + q / H;
}

template <typename T>
__device__ __forceinline__ float wk_ld(const T *p);
template <>
__device__ __forceinline__ float wk_ld<float>(const float *p) { return __ldg(p); }
template <>
__device__ __forceinline__ float wk_ld<__nv_bfloat16>(const __nv_bfloat16 *p) { return __bfloat162float(*p); }
template <typename T>
__device__ __forceinline__ void wk_st(T *p, float v);
template <>
__device__ __forceinline__ void wk_st<float>(float *p, float v) { *p = v; }
template <>
__device__ __forceinline__ void wk_st<__nv_bfloat16>(__nv_bfloat16 *p, float v) { *p = __float2bfloat16_rn(v); }

template <typename T>
__device__ __forceinline__ void wk_ld4(const T *p, float (&v)[4]);
template <>
__device__ __forceinline__ void wk_ld4<float>(const float *p, float (&v)[4]) {
    const float4 t = __ldg(reinterpret_cast<const float4 *>(p));
    v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
}
template <>
__device__ __forceinline__ void wk_ld4<__nv_bfloat16>(const __nv_bfloat16 *p, float (&v)[4]) {
    const uint2 t = __ldg(reinterpret_cast<const uint2 *>(p));
    v[0] = __uint_as_float(t.x << 16), v[1] = __uint_as_float(t.x & 0xffff0000u);
    v[2] = __uint_as_float(t.y << 16), v[3] = __uint_as_float(t.y & 0xffff0000u);
}
template <typename T>
__device__ __forceinline__ void wk_st4(T *p, const float (&v)[4]);
template <>
__device__ __forceinline__ void wk_st4<float>(float *p, const float (&v)[4]) {
    *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <>
__device__ __forceinline__ void wk_st4<__nv_bfloat16>(__nv_bfloat16 *p, const float (&v)[4]) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 raw;
    raw.x = *reinterpret_cast<uint32_t *>(&a);
    raw.y = *reinterpret_cast<uint32_t *>(&b);
    *reinterpret_cast<uint2 *>(p) = raw;
}

// src (B, L, >= c0 + nc) tokens-major, row stride ld_src, batch stride bs_src (elements)  ->  dst[b][c][p], c < nc, fp32,
// batch stride bs_dst.  VEC: 4-channel vector loads are legal (alignment checked by the dispatcher).
template <typename T, bool VEC, int TC>
__global__ void __launch_bounds__(256) walk_pack_kernel(const T *__restrict__ src, long long ld_src, long long bs_src,
                                                        int c0, int nc, float *__restrict__ dst, long long bs_dst, int L,
                                                        int col, WalkGeom g) {
    constexpr int TP = 2048 / TC, TPP = TC / 4, PPP = 256 / TPP;   // threads per position, positions per pass
    __shared__ float tile[TC][TP + 1];
    const int p0 = blockIdx.x * TP, cb = blockIdx.y * TC, b = blockIdx.z;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cg = tid % TPP, c = cb + 4 * cg;
    const T *sb = src + (long long)b * bs_src + c0 + c;
#pragma unroll
    for (int pass = 0; pass < TP / PPP; ++pass) {
        const int pos = pass * PPP + tid / TPP, p = p0 + pos;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (p < L && c < nc) {
            const T *r = sb + (long long)walk_token(g, p, col) * ld_src;
            if (VEC && c + 4 <= nc) {
                wk_ld4<T>(r, v);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (c + j < nc) v[j] = wk_ld<T>(r + j);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) tile[walk_row<TC>(4 * cg + j)][pos] = v[j];
    }
    __syncthreads();
    float *db = dst + (long long)b * bs_dst;
    if (TP >= 32) {
#pragma unroll
        for (int j = 0; j < TC / 8; ++j) {
            const int cl = (TC / 8) * warp + j, cc = cb + cl;
            if (cc < nc) {
#pragma unroll
                for (int half = 0; half < TP / 32; ++half) {
                    const int p = p0 + half * 32 + lane;
                    if (p < L) db[(long long)cc * L + p] = tile[walk_row<TC>(cl)][half * 32 + lane];
                }
            }
        }
    } else {   // TP == 16: a warp writes two channel rows of 16 positions (64 B each) per step
#pragma unroll
        for (int j = 0; j < TC / 16; ++j) {
            const int cl = (TC / 8) * warp + 2 * j + (lane >> 4), cc = cb + cl;
            const int p = p0 + (lane & 15);
            if (cc < nc && p < L) db[(long long)cc * L + p] = tile[walk_row<TC>(cl)][lane & 15];
        }
    }
}

// dst[b][token(p)][c0 + c] (+)= src0[b][c][p] (+ src1[b][c][p]) for c < nc, and 0 for nc <= c < nc_pad.
template <typename T, bool VEC, int TC>
__global__ void __launch_bounds__(256) walk_unpack_kernel(const float *__restrict__ src0, const float *__restrict__ src1,
                                                          long long bs_src, int nc, int nc_pad, T *__restrict__ dst,
                                                          long long ld_dst, long long bs_dst, int c0, int L, int col,
                                                          int accumulate, WalkGeom g) {
    constexpr int TP = 2048 / TC, TPP = TC / 4, PPP = 256 / TPP;
    __shared__ float tile[TC][TP + 1];
    const int p0 = blockIdx.x * TP, cb = blockIdx.y * TC, b = blockIdx.z;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *s0 = src0 + (long long)b * bs_src, *s1 = src1 ? src1 + (long long)b * bs_src : nullptr;
    if (TP >= 32) {
#pragma unroll
        for (int j = 0; j < TC / 8; ++j) {
            const int cl = (TC / 8) * warp + j, cc = cb + cl;
#pragma unroll
            for (int half = 0; half < TP / 32; ++half) {
                const int p = p0 + half * 32 + lane;
                float v = 0.f;
                if (cc < nc && p < L) {
                    v = __ldg(s0 + (long long)cc * L + p);
                    if (s1) v += __ldg(s1 + (long long)cc * L + p);
                }
                tile[walk_row<TC>(cl)][half * 32 + lane] = v;
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < TC / 16; ++j) {
            const int cl = (TC / 8) * warp + 2 * j + (lane >> 4), cc = cb + cl;
            const int p = p0 + (lane & 15);
            float v = 0.f;
            if (cc < nc && p < L) {
                v = __ldg(s0 + (long long)cc * L + p);
                if (s1) v += __ldg(s1 + (long long)cc * L + p);
            }
            tile[walk_row<TC>(cl)][lane & 15] = v;
        }
    }
    __syncthreads();
    const int cg = tid % TPP, c = cb + 4 * cg;
    T *dbase = dst + (long long)b * bs_dst + c0 + c;
#pragma unroll
    for (int pass = 0; pass < TP / PPP; ++pass) {
        const int pos = pass * PPP + tid / TPP, p = p0 + pos;
        if (p < L && c < nc_pad) {
            T *r = dbase + (long long)walk_token(g, p, col) * ld_dst;
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = tile[walk_row<TC>(4 * cg + j)][pos];
            if (VEC && c + 4 <= nc_pad) {
                if (accumulate) {
                    float o[4];
                    wk_ld4<T>(r, o);
#pragma unroll
                    for (int j = 0; j < 4; ++j) v[j] += o[j];
                }
                wk_st4<T>(r, v);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (c + j < nc_pad) wk_st<T>(r + j, accumulate ? v[j] + wk_ld<T>(r + j) : v[j]);
            }
        }
    }
}

template <typename T, bool VEC>
__global__ void __launch_bounds__(256) walk_pack_col2d_kernel(const T *__restrict__ src, long long ld_src, long long bs_src,
                                                              int c0, int nc, float *__restrict__ dst, long long bs_dst,
                                                              int L, WalkGeom g) {
    __shared__ float tile[kColTC * kColCS];
    const ColTile t = col_tile(g, blockIdx.x);
    const int cb = blockIdx.y * kColTC, b = blockIdx.z;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cg = tid & 7, c = cb + 4 * cg;
    const T *sb = src + (long long)b * bs_src + c0 + c;
#pragma unroll
    for (int pass = 0; pass < kColTI * kColTJ / 32; ++pass) {
        const int pp = pass * 32 + (tid >> 3), ii = pp / kColTJ, jj = pp % kColTJ;     // j fastest: consecutive tokens
        const int i = t.i0 + ii, j = t.j0 + jj;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (i < t.H && j < t.W && c < nc) {
            const T *r = sb + (long long)(t.soff + i * t.W + j) * ld_src;
            if (VEC && c + 4 <= nc) {
                wk_ld4<T>(r, v);
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (c + k < nc) v[k] = wk_ld<T>(r + k);
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) tile[(4 * cg + k) * kColCS + jj * (kColTI + 1) + ii] = v[k];
    }
    __syncthreads();
    float *db = dst + (long long)b * bs_dst;
    const int i = t.i0 + lane;
#pragma unroll
    for (int k = 0; k < kColTC / 8; ++k) {
        const int cl = (kColTC / 8) * warp + k, cc = cb + cl;
        if (cc < nc && i < t.H) {
#pragma unroll
            for (int jj = 0; jj < kColTJ; ++jj) {
                const int j = t.j0 + jj;
                if (j < t.W) db[(long long)cc * L + t.soff + j * t.H + i] = tile[cl * kColCS + jj * (kColTI + 1) + lane];
            }
        }
    }
}

template <typename T, bool VEC>
__global__ void __launch_bounds__(256) walk_unpack_col2d_kernel(const float *__restrict__ src0,
                                                                const float *__restrict__ src1, long long bs_src, int nc,
                                                                int nc_pad, T *__restrict__ dst, long long ld_dst,
                                                                long long bs_dst, int c0, int L, int accumulate,
                                                                WalkGeom g) {
    __shared__ float tile[kColTC * kColCS];
    const ColTile t = col_tile(g, blockIdx.x);
    const int cb = blockIdx.y * kColTC, b = blockIdx.z;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *s0 = src0 + (long long)b * bs_src, *s1 = src1 ? src1 + (long long)b * bs_src : nullptr;
    {
        const int i = t.i0 + lane;
#pragma unroll
        for (int k = 0; k < kColTC / 8; ++k) {
            const int cl = (kColTC / 8) * warp + k, cc = cb + cl;
#pragma unroll
            for (int jj = 0; jj < kColTJ; ++jj) {
                const int j = t.j0 + jj;
                float v = 0.f;
                if (cc < nc && i < t.H && j < t.W) {
                    const long long o = (long long)cc * L + t.soff + j * t.H + i;
                    v = __ldg(s0 + o);
                    if (s1) v += __ldg(s1 + o);
                }
                tile[cl * kColCS + jj * (kColTI + 1) + lane] = v;
            }
        }
    }
    __syncthreads();
    const int cg = tid & 7, c = cb + 4 * cg;
    T *dbase = dst + (long long)b * bs_dst + c0 + c;
#pragma unroll
    for (int pass = 0; pass < kColTI * kColTJ / 32; ++pass) {
        const int pp = pass * 32 + (tid >> 3), ii = pp / kColTJ, jj = pp % kColTJ;
        const int i = t.i0 + ii, j = t.j0 + jj;
        if (i < t.H && j < t.W && c < nc_pad) {
            T *r = dbase + (long long)(t.soff + i * t.W + j) * ld_dst;
            float v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = tile[(4 * cg + k) * kColCS + jj * (kColTI + 1) + ii];
            if (VEC && c + 4 <= nc_pad) {
                if (accumulate) {
                    float o[4];
                    wk_ld4<T>(r, o);
#pragma unroll
                    for (int k = 0; k < 4; ++k) v[k] += o[k];
                }
                wk_st4<T>(r, v);
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (c + k < nc_pad) wk_st<T>(r + k, accumulate ? v[k] + wk_ld<T>(r + k) : v[k]);
            }
        }
    }
}

static bool walk_geom(WalkGeom &g, int nstages, const int *Hs, const int *Ws, long long *L) {
    if (nstages <= 0 || nstages > kWalkMaxStages || !Hs || !Ws) return false;
    long long off = 0;
    g.nstage = nstages;
    for (int s = 0; s < kWalkMaxStages; ++s) {
        g.soff[s] = (int)off;
        g.H[s] = g.W[s] = 1;
        if (s < nstages) {
            if (Hs[s] <= 0 || Ws[s] <= 0) return false;
            g.H[s] = Hs[s], g.W[s] = Ws[s];
            off += (long long)Hs[s] * Ws[s];
        }
    }
    g.soff[kWalkMaxStages] = (int)off;
    for (int s = nstages; s < kWalkMaxStages; ++s) g.soff[s] = (int)off;
    long long tiles = 0;
    for (int s = 0; s <= kWalkMaxStages; ++s) {
        g.toff[s] = (int)tiles;
        if (s < nstages)
            tiles += (long long)((g.H[s] + kColTI - 1) / kColTI) * ((g.W[s] + kColTJ - 1) / kColTJ);
    }
    *L = off;
    return off > 0 && off <= 0x7fffffff && tiles <= 0x7fffffff;
}

// returns cudaErrorInvalidValue for shapes the caller should have rejected
cudaError_t walk_pack_dispatch(const void *src, int dtype, long long ld_src, long long bs_src, int c0, int nc, float *dst,
                               long long bs_dst, int batch, int nstages, const int *Hs, const int *Ws, int col,
                               cudaStream_t st) {
    WalkGeom g;
    long long L;
    if (!walk_geom(g, nstages, Hs, Ws, &L)) return cudaErrorInvalidValue;
    const int TC = 32, TP = 2048 / TC;
    const dim3 grid((unsigned)((L + TP - 1) / TP), (unsigned)((nc + TC - 1) / TC), (unsigned)batch);
    const size_t es = dtype == 0 ? 4 : 2;
    const bool vec = (c0 % 4 == 0) && (ld_src % 4 == 0) && (bs_src % 4 == 0) &&
                     ((reinterpret_cast<uintptr_t>(src) % (4 * es)) == 0);
    if (col) {
        const dim3 g2((unsigned)g.toff[kWalkMaxStages], (unsigned)((nc + kColTC - 1) / kColTC), (unsigned)batch);
        if (dtype == 0) {
            const float *sp = static_cast<const float *>(src);
            if (vec) walk_pack_col2d_kernel<float, true><<<g2, 256, 0, st>>>(sp, ld_src, bs_src, c0, nc, dst, bs_dst, (int)L, g);
            else walk_pack_col2d_kernel<float, false><<<g2, 256, 0, st>>>(sp, ld_src, bs_src, c0, nc, dst, bs_dst, (int)L, g);
        } else {
            const __nv_bfloat16 *sp = static_cast<const __nv_bfloat16 *>(src);
            if (vec) walk_pack_col2d_kernel<__nv_bfloat16, true><<<g2, 256, 0, st>>>(sp, ld_src, bs_src, c0, nc, dst, bs_dst, (int)L, g);
            else walk_pack_col2d_kernel<__nv_bfloat16, false><<<g2, 256, 0, st>>>(sp, ld_src, bs_src, c0, nc, dst, bs_dst, (int)L, g);
        }
        return cudaGetLastError();
    }
#define MLAGG_PACK(T_, V_, C_) \
    walk_pack_kernel<T_, V_, C_><<<grid, 256, 0, st>>>(static_cast<const T_ *>(src), ld_src, bs_src, c0, nc, dst, bs_dst, (int)L, col, g)
#define MLAGG_PACK_T(T_)                                  \
    do {                                                  \
        if (vec && TC == 128) MLAGG_PACK(T_, true, 128);  \
        else if (vec) MLAGG_PACK(T_, true, 32);           \
        else if (TC == 128) MLAGG_PACK(T_, false, 128);   \
        else MLAGG_PACK(T_, false, 32);                   \
    } while (0)
    if (dtype == 0) MLAGG_PACK_T(float);
    else MLAGG_PACK_T(__nv_bfloat16);
#undef MLAGG_PACK_T
#undef MLAGG_PACK
    return cudaGetLastError();
}

cudaError_t walk_unpack_dispatch(const float *src0, const float *src1, long long bs_src, int nc, int nc_pad, void *dst,
                                 int dtype, long long ld_dst, long long bs_dst, int c0, int batch, int nstages,
                                 const int *Hs, const int *Ws, int col, int accumulate, cudaStream_t st) {
    WalkGeom g;
    long long L;
    if (!walk_geom(g, nstages, Hs, Ws, &L)) return cudaErrorInvalidValue;
    const int TC = 32, TP = 2048 / TC;
    const dim3 grid((unsigned)((L + TP - 1) / TP), (unsigned)((nc_pad + TC - 1) / TC), (unsigned)batch);
    const size_t es = dtype == 0 ? 4 : 2;
    const bool vec = (c0 % 4 == 0) && (ld_dst % 4 == 0) && (bs_dst % 4 == 0) &&
                     ((reinterpret_cast<uintptr_t>(dst) % (4 * es)) == 0);
    if (col) {
        const dim3 g2((unsigned)g.toff[kWalkMaxStages], (unsigned)((nc_pad + kColTC - 1) / kColTC), (unsigned)batch);
        if (dtype == 0) {
            float *dp = static_cast<float *>(dst);
            if (vec) walk_unpack_col2d_kernel<float, true><<<g2, 256, 0, st>>>(src0, src1, bs_src, nc, nc_pad, dp, ld_dst, bs_dst, c0, (int)L, accumulate, g);
            else walk_unpack_col2d_kernel<float, false><<<g2, 256, 0, st>>>(src0, src1, bs_src, nc, nc_pad, dp, ld_dst, bs_dst, c0, (int)L, accumulate, g);
        } else {
            __nv_bfloat16 *dp = static_cast<__nv_bfloat16 *>(dst);
            if (vec) walk_unpack_col2d_kernel<__nv_bfloat16, true><<<g2, 256, 0, st>>>(src0, src1, bs_src, nc, nc_pad, dp, ld_dst, bs_dst, c0, (int)L, accumulate, g);
            else walk_unpack_col2d_kernel<__nv_bfloat16, false><<<g2, 256, 0, st>>>(src0, src1, bs_src, nc, nc_pad, dp, ld_dst, bs_dst, c0, (int)L, accumulate, g);
        }
        return cudaGetLastError();
    }
#define MLAGG_UNPACK(T_, V_, C_)                                                                                        \
    walk_unpack_kernel<T_, V_, C_><<<grid, 256, 0, st>>>(src0, src1, bs_src, nc, nc_pad, static_cast<T_ *>(dst), ld_dst, \
                                                         bs_dst, c0, (int)L, col, accumulate, g)
#define MLAGG_UNPACK_T(T_)                                  \
    do {                                                    \
        if (vec && TC == 128) MLAGG_UNPACK(T_, true, 128);  \
        else if (vec) MLAGG_UNPACK(T_, true, 32);           \
        else if (TC == 128) MLAGG_UNPACK(T_, false, 128);   \
        else MLAGG_UNPACK(T_, false, 32);                   \
    } while (0)
    if (dtype == 0) MLAGG_UNPACK_T(float);
    else MLAGG_UNPACK_T(__nv_bfloat16);
#undef MLAGG_UNPACK_T
#undef MLAGG_UNPACK
    return cudaGetLastError();
}

}  // namespace mlagg
