/*
 * mlagg_b200.h -- C ABI of libmlagg_b200.so (sm_100a kernels for the MLAgg + MSMM hot path).
 *
 * Conventions:
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless named host_*;
 *   - the CALLER owns every buffer, including workspaces; nothing is allocated, no global state is
 *     kept, no implicit synchronisation: work is enqueued on `stream` and the call returns;
 *   - return value 0 = enqueued; negative = MLAGG_ERR_* (nothing was enqueued);
 *   - all tensors are contiguous in the layout stated; "nullable" pointers may be NULL;
 *   - reentrant and thread-safe (autograd calls the *_bwd entry points from its own thread).
 *
 * Each entry point names the reference interface it replaces (paths relative to the reference repo,
 * mlagg/nnunetv2/training/nnUNetTrainer/...).
 */
#ifndef MLAGG_B200_H
#define MLAGG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void *mlagg_stream_t; /* cudaStream_t */

enum {
    MLAGG_OK = 0,
    MLAGG_ERR_BAD_SHAPE = -1,   /* non-positive size, dim % ngroups != 0, ... */
    MLAGG_ERR_UNSUPPORTED = -2, /* e.g. dstate not in {16}, unknown dtype code */
    MLAGG_ERR_NULL = -3,        /* required pointer is NULL */
    MLAGG_ERR_LAUNCH = -4,      /* cudaGetLastError() != cudaSuccess after the launch */
    MLAGG_ERR_ALIGN = -5        /* pointer not aligned to the element size */
};

/* element type codes for activations */
enum { MLAGG_F32 = 0, MLAGG_BF16 = 1 };

int mlagg_version(void);
const char *mlagg_error_string(int code);
/* text of the last CUDA error seen by the calling thread inside this library ("" if none) */
const char *mlagg_last_cuda_error(void);

/* --------------------------------------------------------------------------------------------
 * Selective scan (S6), forward.
 * Replaces selective_scan_cuda.fwd(u, delta, A, B, C, D, z, delta_bias, delta_softplus)
 *   -- FFI shape documented at variants/mamba/vmamba/csms6s.py:224, called through
 *      mamba_ssm's selective_scan_fn at variants/mamba/MambaSkip.py:445-451 (and :191-197).
 *   u, delta, out : (batch, dim, seqlen) fp32        A : (dim, dstate) fp32 (real, = -exp(A_log))
 *   B, C          : (batch, ngroups, dstate, seqlen) fp32; channel d uses group d / (dim/ngroups)
 *   D, delta_bias : (dim) fp32, nullable
 *   ckpt          : nullable.  When given, the running state is saved every MLAGG_SCAN_CHUNK steps for
 *                   the backward pass: mlagg_scan_ckpt_bytes(...) bytes.  Pass NULL for inference.
 *   last_state    : nullable, (batch, dim, dstate) fp32 -- h after the final step.
 * ------------------------------------------------------------------------------------------ */
#define MLAGG_SCAN_CHUNK 16
size_t mlagg_scan_ckpt_bytes(int batch, int dim, int seqlen, int dstate);

int mlagg_selective_scan_fwd(const float *u, const float *delta, const float *A, const float *B,
                             const float *C, const float *D, const float *delta_bias, float *out,
                             float *ckpt, float *last_state, int batch, int dim, int seqlen, int dstate,
                             int ngroups, int delta_softplus, mlagg_stream_t stream);

/* --------------------------------------------------------------------------------------------
 * Selective scan, backward.
 * Replaces selective_scan_cuda.bwd(u, delta, A, B, C, D, z, delta_bias, dout, x, out, dz,
 *                                  delta_softplus, recompute_out_z)  -- csms6s.py:235-238.
 *   dout, du, ddelta : (batch, dim, seqlen) fp32
 *   dA (dim, dstate), dD (dim, nullable iff D NULL), ddelta_bias (dim, nullable iff delta_bias NULL):
 *       fp32, ACCUMULATED INTO with atomics -- the caller zero-fills them (or passes running sums).
 *   dB, dC : (batch, ngroups, dstate, seqlen) fp32, ACCUMULATED INTO likewise (zero-fill first).
 *   ckpt   : the buffer the forward call filled.
 * ------------------------------------------------------------------------------------------ */
int mlagg_selective_scan_bwd(const float *u, const float *delta, const float *A, const float *B,
                             const float *C, const float *D, const float *delta_bias, const float *dout,
                             const float *ckpt, float *du, float *ddelta, float *dA, float *dB, float *dC,
                             float *dD, float *ddelta_bias, int batch, int dim, int seqlen, int dstate,
                             int ngroups, int delta_softplus, mlagg_stream_t stream);

/* --------------------------------------------------------------------------------------------
 * Depthwise 3x3 convolution, padding 1, on TOKENS-MAJOR activations x (batch, H, W, C)  [= (B, N, C)].
 * Replaces nn.Conv2d(C, C, 3, padding=1, groups=C) (+ the permute copies around it, + the SiLU after it) at
 *   nnUNetTrainer_MLAgg_2D_dt_MS.py:851,890 (dwc) and :680,782 (lepe); variants/mamba/MambaSkip.py:302-312,
 *   :521-523 (SS2D_skip.conv2d[i] + act) and :545-556 (DWConv); nnUNetTrainer_MLLA_UNet.py:279,289,215.
 *   x, y, dy, dx, dz_ws : (batch, H, W, C), element type `dtype` (MLAGG_F32 | MLAGG_BF16); C % 4 == 0 takes the
 *                         128-bit vector path, other C a scalar-channel path
 *   weight (C, 9) fp32 [the (C,1,3,3) parameter], bias (C) fp32 nullable
 *   act_silu != 0 : y = silu(conv(x) + bias)
 * Backward: dz_ws is a caller-owned scratch of x's size; dweight (C,9) / dbias (C) fp32 are ACCUMULATED
 * INTO (zero-fill first); dbias nullable.
 * ------------------------------------------------------------------------------------------ */
int mlagg_dwconv3x3_fwd(const void *x, const float *weight, const float *bias, void *y, int batch, int H, int W,
                        int C, int act_silu, int dtype, mlagg_stream_t stream);
int mlagg_dwconv3x3_bwd(const void *x, const float *weight, const float *bias, const void *dy, void *dz_ws,
                        void *dx, float *dweight, float *dbias, int batch, int H, int W, int C, int act_silu,
                        int dtype, mlagg_stream_t stream);
/* Strided variants: pixel stride ld* and image stride bs* in ELEMENTS for every activation operand (ld >= C; multiples of 4
 * when C % 4 == 0), so channel slices of wider activations (the v half of the kv projection for LePE, :680/:782; the first
 * half of ConvolutionalGLU's fc1 output, MambaSkip.py:567-575) and the per-stage segments of the stage-concatenated MSMM
 * sequence (MambaSkip.py:521-523) are read and written in place.  `residual` (nullable, same shape as y) is combined with
 * the result after the activation: residual_mul = 0: y = act(conv(x) + b) + residual -- the `attn_out + lepe(v)` of
 * :716/:759; residual_mul = 1: y = act(conv(x) + b) * residual -- the `act(dwconv(x)) * v` of ConvolutionalGLU,
 * MambaSkip.py:574.
 * The backward takes dy with its own strides, a CONTIGUOUS dz workspace, and stores dx with (lddx, bsdx).  With `mul`
 * (the forward's multiplicative `residual`, strides ldm / bsm; nullable) it also stores dmul = dy * act(conv(x) + b) with
 * the same strides and differentiates through the product -- so ConvolutionalGLU's gradient lands in the two halves of
 * ONE fc1-output gradient (dx -> first half, dmul -> second half). */
int mlagg_dwconv3x3_fwd_strided(const void *x, const float *weight, const float *bias, const void *residual, void *y,
                                int batch, int H, int W, int C, long long ldx, long long bsx, long long ldr,
                                long long bsr, long long ldy, long long bsy, int act_silu, int residual_mul, int dtype,
                                mlagg_stream_t stream);
int mlagg_dwconv3x3_bwd_strided(const void *x, const float *weight, const float *bias, const void *dy, void *dz_ws,
                                void *dx, float *dweight, float *dbias, int batch, int H, int W, int C, long long ldx,
                                long long bsx, long long lddy, long long bsdy, long long lddx, long long bsdx,
                                const void *mul, void *dmul, long long ldm, long long bsm, int act_silu, int dtype,
                                mlagg_stream_t stream);
int mlagg_causal_conv1d_fwd(const float *x, const float *weight, const float *bias, float *y, int batch, int C,
                            int L, int K, int act_silu, mlagg_stream_t stream);
int mlagg_causal_conv1d_bwd(const float *x, const float *weight, const float *bias, const float *dy, float *dx,
                            float *dweight, float *dbias, int batch, int C, int L, int K, int act_silu,
                            mlagg_stream_t stream);

/* --------------------------------------------------------------------------------------------
 * Local (3x3 window) differential softmax attention + sub-LN, tokens-major, fused.
 * Replaces the op chain of AggregatedAttention.forward, local branch, nnUNetTrainer_MLAgg_2D_dt_MS.py:698-717
 * (nn.Unfold x2, q@k, masked_fill, softmax, lambda-combine, attn@v, RMSNorm(subln), * (1 - lambda_init)).
 *   heads = h (the module's num_heads), head_dim = hd in {2,4,8,16,24,32}; C = 2*h*hd
 *   q   : (batch, H*W, 2h, hd)  RAW projection output; row stride ldq elements; `scale` (= hd**-0.5) applied inside
 *   k   : (batch, H*W, 2h, hd), v : (batch, H*W, h, 2hd); common row stride ldkv (the halves of the kv Linear)
 *   out : (batch, H*W, h, 2hd)  row stride ldo
 *   subln_w (2hd) fp32; lam = DEVICE pointer to the fp32 scalar lambda_full (no host sync); eps = 1e-5; post_scale = 1 - lambda_init
 *   dtype: MLAGG_F32 | MLAGG_BF16 for q, k, v, out, dout, dq, dk, dv.
 * Backward: dk / dv share row stride lddkv; d_subln_w (2hd) and d_lambda (1) fp32 are ACCUMULATED INTO;
 *   ws: caller-owned scratch of mlagg_local_diffattn_ws_bytes(...) bytes.
 * ------------------------------------------------------------------------------------------ */
size_t mlagg_local_diffattn_ws_bytes(int batch, int H, int W, int heads, int head_dim);
int mlagg_local_diffattn_fwd(const void *q, const void *k, const void *v, const float *subln_w, void *out,
                             int batch, int H, int W, int heads, int head_dim, long long ldq, long long ldkv,
                             long long ldo, float scale, const float *lam, float eps, float post_scale, int dtype,
                             mlagg_stream_t stream);
int mlagg_local_diffattn_bwd(const void *q, const void *k, const void *v, const float *subln_w, const void *dout,
                             void *dq, void *dk, void *dv, float *d_subln_w, float *d_lambda, void *ws, int batch,
                             int H, int W, int heads, int head_dim, long long ldq, long long ldkv, long long lddo,
                             long long lddq, long long lddkv, float scale, const float *lam, float eps, float post_scale,
                             int dtype, mlagg_stream_t stream);

/* --------------------------------------------------------------------------------------------
 * Pooled-token differential softmax attention + sub-LN, tokens-major, fused.
 * Replaces AggregatedAttention.forward, pooled branch, nnUNetTrainer_MLAgg_2D_dt_MS.py:734-760: the four
 * flash_attn_func(q_j, k_j, v_i) calls (:745-751), both torch.cat, the lambda-combine (:756), RMSNorm (:759) and
 * * (1 - lambda_init) (:760).  The shipped DOUBLE scaling is reproduced: logits = (q . k) * scale * scale where
 * `scale` = hd**-0.5 (once at :688, once more as flash-attn's default softmax_scale).
 *   q   : (batch, N, h, 2, hd) RAW projection, row stride ldq
 *   kp  : (batch, P, h, 2, hd), vp : (batch, P, h, 2hd); common row stride ldkv; P <= 256, P*hd*16 B of shared memory
 *   out : (batch, N, h, 2hd) row stride ldo;
 *   lse : nullable; mlagg_pooled_diffattn_saved_bytes(...) bytes, 16-byte aligned, saved for the backward pass:
 *         (batch, N, h, 2) log-sum-exps followed by the normalised per-map outputs O0 | O1 (batch, N, h, 2, 2hd) fp32
 * Backward: dkp / dvp are fp32 (batch, P, h, 2hd) with common row stride ldd, ACCUMULATED INTO (zero-fill);
 *   d_subln_w (2hd), d_lambda (1) ACCUMULATED INTO; ws: mlagg_pooled_diffattn_ws_bytes(...) bytes of scratch.
 * ------------------------------------------------------------------------------------------ */
size_t mlagg_pooled_diffattn_ws_bytes(int batch, int N, int heads, int head_dim);
size_t mlagg_pooled_diffattn_saved_bytes(int batch, int N, int heads, int head_dim);
int mlagg_pooled_diffattn_fwd(const void *q, const void *kp, const void *vp, const float *subln_w, void *out,
                              float *lse, int batch, int N, int P, int heads, int head_dim, long long ldq,
                              long long ldkv, long long ldo, float scale, const float *lam, float eps,
                              float post_scale, int dtype, mlagg_stream_t stream);
int mlagg_pooled_diffattn_bwd(const void *q, const void *kp, const void *vp, const float *subln_w,
                              const float *lse, const void *dout, void *dq, float *dkp, float *dvp,
                              float *d_subln_w, float *d_lambda, void *ws, int batch, int N, int P, int heads,
                              int head_dim, long long ldq, long long ldkv, long long lddo, long long lddq,
                              long long ldd, float scale, const float *lam, float eps, float post_scale,
                              int dtype, mlagg_stream_t stream);

/* --------------------------------------------------------------------------------------------
 * Fused MSMM scan: SS2D_skip.forward_corev0 (variants/mamba/MambaSkip.py:405-473) without materialising the four
 * direction-permuted copies of x (:414-422), the dt projection (:434), the fp32 casts (:437-443) or the
 * un-permutation (:454-471).  Same recurrence and kernels as mlagg_selective_scan_*; only operand addressing differs.
 *   xrow, xcol         : (batch, d_inner, L) fp32 -- the conv+SiLU output, channels-major, stages concatenated along L,
 *                        each stage in row-major (xrow) / column-major (xcol: W x H transposed image) order
 *   xdbl_row, xdbl_col : (batch, 2, dt_rank + 2*dstate, L) fp32 -- x_proj_weight[{0,2}] @ xrow and [{1,3}] @ xcol;
 *                        per direction: dt_rank rows of dts, dstate rows of B, dstate rows of C
 *   Wdt (4*d_inner, dt_rank), dt_bias (4*d_inner), A (4*d_inner, dstate) = -exp(A_logs), Ds (4*d_inner): fp32
 *   stage_lens : HOST array of nstages (<= 8) stage lengths H_s*W_s; L = their sum
 *   out  : (batch, 4, d_inner, L) fp32; direction k's result in row-major (k even) / column-major (k odd) order
 *          with the mirroring of directions 2, 3 already undone:  y = out0 + out2 + colmajor_to_rowmajor(out1 + out3)
 *   ckpt : as in mlagg_selective_scan_fwd (dim = 4*d_inner), nullable
 * Backward: dout like out; du (batch, 4, d_inner, L) plain stores in the same orders (dxrow = du0 + du2,
 *   dxcol = du1 + du3); dxdbl_row / dxdbl_col, dWdt, dA, dDs, ddt_bias fp32 ACCUMULATED INTO (zero-fill first).
 * dt_rank <= 4, dstate == 16.
 * ------------------------------------------------------------------------------------------ */
int mlagg_msmm_scan_fwd(const float *xrow, const float *xcol, const float *xdbl_row, const float *xdbl_col,
                        const float *Wdt, const float *dt_bias, const float *A, const float *Ds, float *out,
                        float *ckpt, int batch, int d_inner, int dstate, int dt_rank, int nstages,
                        const int *stage_lens, mlagg_stream_t stream);
int mlagg_msmm_scan_bwd(const float *xrow, const float *xcol, const float *xdbl_row, const float *xdbl_col,
                        const float *Wdt, const float *dt_bias, const float *A, const float *Ds, const float *dout,
                        const float *ckpt, float *du, float *dxdbl_row, float *dxdbl_col, float *dWdt,
                        float *ddt_bias, float *dA, float *dDs, int batch, int d_inner, int dstate, int dt_rank,
                        int nstages, const int *stage_lens, int dout_walks, mlagg_stream_t stream);
/* dout_walks = 0: dout is (batch, 4, d_inner, L), one plane per direction (the layout of `out`);
 * dout_walks = 1: dout is (batch, 2, d_inner, L) = the gradient of the MERGED output y in row-major (plane 0, read by
 *                 directions 0, 2) and column-major (plane 1, directions 1, 3) walk order -- what the cross-merge's
 *                 backward (MambaSkip.py:454-471) would broadcast into four planes. */

/* --------------------------------------------------------------------------------------------
 * Walk packing around the fused MSMM scan: the cross-scan / cross-merge index maps of MambaSkip.py:414-422 and :454-471
 * as one tile-transpose pass each, between the TOKENS-MAJOR activations the rest of the block uses and the
 * channels-major fp32 planes in walk order the scan kernels read and write.  The sequence is the concatenation of
 * nstages (<= 8) images of H_s x W_s tokens (HOST arrays Hs, Ws), L = sum H_s*W_s.  Walk position p visits token p
 * (col_walk = 0) or, inside stage s with local position q, token soff_s + (q % H_s) * W_s + q / H_s (col_walk = 1).
 *   mlagg_walk_pack  : dst[b][c][p] = (float) src[b][token(p)][c0 + c],  c < nc
 *       src (batch, L, .) of `dtype`, row stride ld_src and batch stride bs_src in ELEMENTS; dst fp32, plane stride L,
 *       batch stride bs_dst elements.
 *   mlagg_walk_unpack: dst[b][token(p)][c0 + c] (+)= src0[b][c][p] (+ src1[b][c][p]) for c < nc; columns
 *       nc <= c < nc_pad receive 0 (+ old value when accumulating).  src0 / src1 (nullable) fp32 planes with the same
 *       batch stride bs_src; dst (batch, L, .) of `dtype`, strides ld_dst / bs_dst; accumulate != 0 adds to dst.
 * Forward use: x, x_dbl -> xrow / xcol / xdbl_row / xdbl_col;  out -> y = out0 + out2 + colmajor_to_rowmajor(out1 + out3).
 * Backward use: dy -> dout (dout_walks = 1);  du, dxdbl_row / dxdbl_col -> dx, dx_dbl.
 * ------------------------------------------------------------------------------------------ */
int mlagg_walk_pack(const void *src, int dtype, long long ld_src, long long bs_src, int c0, int nc, float *dst,
                    long long bs_dst, int batch, int nstages, const int *host_Hs, const int *host_Ws, int col_walk,
                    mlagg_stream_t stream);
int mlagg_walk_unpack(const float *src0, const float *src1, long long bs_src, int nc, int nc_pad, void *dst, int dtype,
                      long long ld_dst, long long bs_dst, int c0, int batch, int nstages, const int *host_Hs,
                      const int *host_Ws, int col_walk, int accumulate, mlagg_stream_t stream);

/* --------------------------------------------------------------------------------------------
 * LayerNorm over the last dimension of tokens-major activations, fp32 math, fp32 or bf16 input and output.
 * Replaces nn.LayerNorm (+ the fp32 up-cast and the consumers' down-cast copies autocast inserts) at
 *   nnUNetTrainer_MLAgg_2D_dt_MS.py:848,871 (norm1, norm2), :670 (pooled-token norm);
 *   variants/mamba/MambaSkip.py:344 (out_norm), :686,690 (ln_1, norm2 of VSS_Conv_Block).
 *   x (M, C) of dt_in, y (M, C) of dt_out, weight / bias (C) fp32 (bias nullable), mean / rstd (M) fp32 saved for backward;
 *   C % 4 == 0, C <= 1024.
 * Backward: dy (M, C) of dt_out -> dx (M, C) of dt_in; dweight / dbias (C) fp32 ACCUMULATED INTO (zero-fill first).
 * ------------------------------------------------------------------------------------------ */
int mlagg_layernorm_fwd(const void *x, const float *weight, const float *bias, void *y, float *mean, float *rstd,
                        long long M, int C, float eps, int dt_in, int dt_out, mlagg_stream_t stream);
int mlagg_layernorm_bwd(const void *x, const float *weight, const float *mean, const float *rstd, const void *dy,
                        void *dx, float *dweight, float *dbias, long long M, int C, int dt_in, int dt_out,
                        mlagg_stream_t stream);
/* Same with the gradient that reaches x along the RESIDUAL path added in the same pass: dx = LN'(dy) + dres
 * (dres (M, C) of dt_in, nullable).  Every pre-norm residual branch of the path -- `x + f(norm(x))` at
 * nnUNetTrainer_MLAgg_2D_dt_MS.py:877-911 (twice per block) and variants/mamba/MambaSkip.py:738 -- otherwise costs one
 * full-tensor add in autograd's gradient accumulation. */
int mlagg_layernorm_bwd_res(const void *x, const float *weight, const float *mean, const float *rstd, const void *dy,
                            const void *dres, void *dx, float *dweight, float *dbias, long long M, int C, int dt_in,
                            int dt_out, mlagg_stream_t stream);

/* --------------------------------------------------------------------------------------------
 * elu+1 linear attention core of MLLA (the ops BASELINE.json:north_star names).
 * Replaces the op sequence at nnUNetTrainer_MLLA_UNet.py:234-246 (LinearAttention.forward: elu+1 on q and k,
 * RoPE.forward :190-195 on both, z = 1/(q . mean_n k + 1e-6), kv = (k_rope^T n^-1/2)(v n^-1/2), q_rope kv z);
 * the qk projection (:229) and the LePE conv (:250) stay outside (cuBLAS / mlagg_dwconv3x3_*).
 *   q, k, v, out : tokens-major (batch, N = H*W, heads, head_dim) of `dtype`, row strides ldq, ldk, ldv, ldo in
 *                  ELEMENTS (q and k are the halves of the qk projection, consumed in place); 16-byte aligned rows
 *   rope_cs      : (H + W, C/4, 2) fp32, C = heads*head_dim: [cos, sin] of row*theta_i for the H rows, then of
 *                  col*theta_i for the W columns, theta_i = base^(-i/(C/4)) (the reference's table, :181-187, separable)
 *   state        : mlagg_linattn_state_bytes(...) bytes; receives S (batch, heads, hd, hd) then kmean (batch, heads, hd),
 *                  fp32; saved for the backward pass
 *   head_dim in {8, 16, 32} (MLLA-UNet ships 32, :53-56)
 * Backward: dout like out (row stride lddo); dq, dk, dv plain stores with row strides lddq, lddk, lddv;
 *   ws: scratch of mlagg_linattn_state_bytes(...) bytes (dS, dkmean).
 * ------------------------------------------------------------------------------------------ */
size_t mlagg_linattn_state_bytes(int batch, int heads, int head_dim);
int mlagg_linattn_fwd(const void *q, const void *k, const void *v, const float *rope_cs, void *out, float *state,
                      int batch, int H, int W, int heads, int head_dim, long long ldq, long long ldk, long long ldv,
                      long long ldo, float eps, int dtype, mlagg_stream_t stream);
int mlagg_linattn_bwd(const void *q, const void *k, const void *v, const float *rope_cs, const float *state,
                      const void *dout, void *dq, void *dk, void *dv, float *ws, int batch, int H, int W, int heads,
                      int head_dim, long long ldq, long long ldk, long long ldv, long long lddo, long long lddq,
                      long long lddk, long long lddv, float eps, int dtype, mlagg_stream_t stream);

/* --------------------------------------------------------------------------------------------
 * Projection contractions on the 5th-generation tensor cores (csrc/gemm_tc.cu: TMA operand loads, tcgen05.mma with the
 * accumulator in tensor memory, tcgen05.ld epilogue).  Replace the cuBLAS GEMMs autograd runs for every nn.Linear /
 * 1x1 convolution of the named path and their two backward GEMMs:
 *   nnUNetTrainer_MLAgg_2D_dt_MS.py:673-674 (q, kv), :849-850 (in_proj, act_proj), :867 / :902 (out_proj), :176-192 (Mlp:
 *   fc1 -> GELU -> fc2), :660-661 (sr 1x1 conv); variants/mamba/MambaSkip.py:301 (in_proj), :345 (out_proj), :431 (x_proj as
 *   ONE tokens-major GEMM), :559-577 (ConvolutionalGLU fc1 / fc2).
 * All activations / weights are bf16, row-major with unit column stride and a row stride (ld*, in ELEMENTS, multiple of 8;
 * base pointers 16-byte aligned): channel slices of a wider activation are consumed and produced in place.  fp32
 * accumulation.  N % 8 == 0 and K % 8 == 0.  act: 0 none, 1 exact GELU, 2 SiLU.  out_dtype: MLAGG_BF16 / MLAGG_F32.
 *
 *   mlagg_linear_fwd        y[M,N] = act(x[M,K] . w[N,K]^T + bias[N]);   bias fp32, nullable;
 *                           pre (nullable, bf16 [M, ldpre]) additionally receives the PRE-activation for the backward.
 *   mlagg_linear_bwd_data   dx[M,K] = (dy[M,N] . w[N,K]) (*) act'(aux[M,K]);  aux nullable (then act is ignored): the
 *                           gradient through the activation in FRONT of this layer (Mlp: GELU between fc1 and fc2) is
 *                           applied in the epilogue.  w is read as stored (MN-major B operand, no transposed copy).
 *   mlagg_linear_bwd_weight dw[N,K] += dy[M,N]^T . x[M,K];  dw fp32, row stride lddw, ACCUMULATED INTO (zero-fill first);
 *                           the contraction over the M tokens is split across CTAs and reduced with fp32 vector atomics.
 *                           db (nullable, fp32 [N], ACCUMULATED INTO) += column sums of dy -- the bias gradient, taken from
 *                           the same operand tiles by one extra MMA against a tile of ones (replaces mlagg_colsum there).
 * ------------------------------------------------------------------------------------------ */
int mlagg_linear_fwd(const void *x, long long ldx, const void *w, long long ldw, const float *bias, void *y,
                     long long ldy, void *pre, long long ldpre, long long M, int N, int K, int act, int out_dtype,
                     mlagg_stream_t stream);
int mlagg_linear_bwd_data(const void *dy, long long lddy, const void *w, long long ldw, const void *aux, long long ldaux,
                          int act, void *dx, long long lddx, long long M, int N, int K, int out_dtype,
                          mlagg_stream_t stream);
int mlagg_linear_bwd_weight(const void *dy, long long lddy, const void *x, long long ldx, float *dw, long long lddw,
                            float *db, long long M, int N, int K, mlagg_stream_t stream);

/* --------------------------------------------------------------------------------------------
 * Column sums of a tokens-major matrix: out[c] += sum_m x[m * ld + c]  (out fp32, ACCUMULATED INTO: zero-fill first).
 * Replaces autograd's `grad_output.sum(0)` for the bias gradient of the nn.Linear layers of the hot path
 *   (nnUNetTrainer_MLAgg_2D_dt_MS.py:849-850, :868, :180-186, :673-674; variants/mamba/MambaSkip.py:567,570).
 *   x (M, C) of `dtype`, row stride ld >= C in ELEMENTS.
 * ------------------------------------------------------------------------------------------ */
int mlagg_colsum(const void *x, float *out, long long M, int C, long long ld, int dtype, mlagg_stream_t stream);

/* --------------------------------------------------------------------------------------------
 * Per-(image, channel) normalisation over the pixels of a channels-last / tokens-major (batch, N, C) map, fp32 math:
 *   y = act(w[c] * (x - mean[b,c]) * rstd[b,c] + b[c]),   act: 0 = identity, 1 = LeakyReLU(slope), 2 = SiLU.
 * Replaces nn.InstanceNorm2d (variants/mamba/MambaSkip.py:714-716; monai UnetResBlock norms of encoder0 / decoder0,
 *   nnUNetTrainer_MLAgg_2D_dt_MS.py:1339-1357) and nn.GroupNorm(num_groups=C) (:262, :497) without torch's NCHW copies.
 *   x, y (batch, N, C) of `dtype`; w, b (C) fp32, nullable (no affine); stats (batch, C, 2) fp32 receives (mean, rstd);
 *   C % 4 == 0, 16-byte aligned rows.
 * Backward: dy like y -> dx like x; sums (batch, C, 2) fp32 scratch; dw, db (C) fp32 nullable, ACCUMULATED INTO.
 * ------------------------------------------------------------------------------------------ */
int mlagg_instnorm_fwd(const void *x, const float *w, const float *b, void *y, float *stats, int batch, int N, int C,
                       float eps, int act, float slope, int dtype, mlagg_stream_t stream);
int mlagg_instnorm_bwd(const void *x, const float *w, const float *b, const float *stats, const void *dy, void *dx,
                       float *sums, float *dw, float *db, int batch, int N, int C, int act, float slope, int dtype,
                       mlagg_stream_t stream);
/* The residual form, y = act(norm(x) + residual) with act 0 | 1: the tail of monai's UnetResBlock,
 * `lrelu(norm2(conv2(.)) + residual)` (reference network: encoder0 / decoder stages, nnUNetTrainer_MLAgg_2D_dt_MS.py:1339-1357),
 * in the apply pass instead of an add and an activation kernel.  The backward takes the saved OUTPUT y (the LeakyReLU
 * slope is read off its sign), stores dx and d residual = dy * act'.  C % 8 == 0 (bf16) / % 4 (fp32), 16-byte aligned. */
int mlagg_instnorm_res_fwd(const void *x, const float *w, const float *b, const void *residual, void *y, float *stats,
                           int batch, int N, int C, float eps, int act, float slope, int dtype, mlagg_stream_t stream);
int mlagg_instnorm_res_bwd(const void *x, const float *w, const float *b, const float *stats, const void *y, const void *dy,
                           void *dx, void *dresidual, float *sums, float *dw, float *db, int batch, int N, int C, int act,
                           float slope, int dtype, mlagg_stream_t stream);

/* --------------------------------------------------------------------------------------------
 * Adaptive average pooling of a tokens-major map, optional exact GELU applied to x on the fly:
 *   y[b, (i, j), c] = mean over rows floor(i*H/pH) .. ceil((i+1)*H/pH) - 1 and the matching columns of act(x[b, (r, s), c]).
 * Replaces  self.pool(self.act(...))  of the pooled-token branch, nnUNetTrainer_MLAgg_2D_dt_MS.py:720-723
 *   (nn.AdaptiveAvgPool2d :668, nn.GELU :671) without the NCHW view or the materialised GELU output.
 *   x (batch, H*W, C), y (batch, pH*pW, C) of `dtype`; C % 4 == 0, C <= 4096.
 * Backward: dy like y -> dx like x (x is the forward input; only read when act_gelu != 0).
 * ------------------------------------------------------------------------------------------ */
int mlagg_avgpool_tokens_fwd(const void *x, void *y, int batch, int H, int W, int C, int pH, int pW, int act_gelu,
                             int dtype, mlagg_stream_t stream);
int mlagg_avgpool_tokens_bwd(const void *x, const void *dy, void *dx, int batch, int H, int W, int C, int pH, int pW,
                             int act_gelu, int dtype, mlagg_stream_t stream);

/* --------------------------------------------------------------------------------------------
 * Element-wise seams of the MLAgg block (contiguous tensors of `dtype`, n % 4 == 0, 16-byte (fp32) / 8-byte (bf16) aligned).
 *   mlagg_residual_scale : out = x + scale[b] * y -- `shortcut + drop_path(branch)`, nnUNetTrainer_MLAgg_2D_dt_MS.py:907-908
 *       (timm DropPath: scale[b] = Bernoulli(keep) / keep per sample, fp32 (batch); NULL = 1).  x NULL = 0, which is the
 *       backward of the branch input (dbranch = scale[b] * dout).  per_sample = n / batch, % 4 == 0.
 *   mlagg_silu_gate_fwd  : out = t * silu(z) -- `x * act_res` with act_res = SiLU(act_proj(.)), :881, :907.
 *   mlagg_silu_gate_bwd  : dt = dout * silu(z), dz = dout * t * silu'(z).
 *   mlagg_diff_lambda_fwd: out[0] = exp(<lq1, lk1>) - exp(<lq2, lk2>) + lambda_init (:700-702, :745-747), out[1], out[2] =
 *       the two exponentials (saved for the backward); vectors of n fp32.
 *   mlagg_diff_lambda_bwd: grads (4, n) fp32 = d out[0] / d (lq1, lk1, lq2, lk2) * dlam[0], plain stores.
 * ------------------------------------------------------------------------------------------ */
int mlagg_residual_scale(const void *x, const void *y, const float *scale, void *out, long long n, long long per_sample,
                         int dtype, mlagg_stream_t stream);
/* dst[b][r][0:cols] = src[b][r][0:cols], b < batch, r < rows, for row-strided views of tokens-major activations (row stride
 * ld*, batch stride bs*, in ELEMENTS; unit column stride): the channel split of the MSMM inputs (`x[:, :hidden]` /
 * `x[:, hidden:]`, MambaSkip.py:724-727), the per-stage split / concatenation of the token sequence (:728-746) and the
 * gradients of both, without intermediate zero-filled tensors. */
int mlagg_copy_rows(const void *src, long long ld_src, long long bs_src, void *dst, long long ld_dst, long long bs_dst,
                    int batch, long long rows, int cols, int dtype, mlagg_stream_t stream);
/* dst[b][r][0:cols] += src[b][r][0:cols], same addressing: the LePE gradient joins the v half of the kv gradient
 * (reference :716 / :759 `x + lepe(v)`, v = kv[..., C:]) without autograd's zero-filled kv-sized tensor and full-size add.
 * cols, strides and bases must allow 8-byte vectors, cols * elsize <= 4096; MLAGG_ERR_UNSUPPORTED otherwise. */
int mlagg_add_rows(const void *src, long long ld_src, long long bs_src, void *dst, long long ld_dst, long long bs_dst,
                   int batch, long long rows, int cols, int dtype, mlagg_stream_t stream);

/* ---- statistics of DC_and_CE_loss for one deep-supervision scale (reference training/loss/compound_losses.py
 * DC_and_CE_loss, dice.py:58-112 MemoryEfficientSoftDiceLoss with softmax, robust_ce_loss.py; called once per scale by
 * nnUNetTrainer.py:833-863).  logits (batch, K, npix) addressed through element strides (sb, sc, sn): NCHW heads have
 * sc = npix, sn = 1, channels_last heads sc = 1, sn = K; dtype fp32 / bf16.  target (batch, npix) class indices, float
 * (tdtype 0, rounded) or int64 (tdtype 1).  K <= 32.  dtype | 2: the caller states that every pixel row of logits (and of
 * dlogits) owns 16 contiguous elements (heads padded to 16 channels): bf16 rows are then read / written as two 16-byte
 * vectors, and the gradient's padding columns receive zeros.
 *   fwd: stats (batch, K, 3) fp32 += (sum_n p_k [t = k], sum_n p_k, sum_n [t = k]) with p = softmax over K;
 *        ce (1) += sum over every pixel of -log p_t.  Both accumulated into: zero-fill them.
 *   bwd: dlogits (strides of logits) = d loss / d logits given g_stats (batch, K, 3) (the count component is ignored) and
 *        g_ce (1) = d loss / d ce, softmax recomputed.
 * The dice formula itself stays host-side arithmetic on `stats` (batch-dice, the DDP gather, the smoothing terms). */
int mlagg_dice_ce_stats_fwd(const void *logits, const void *target, float *stats, float *ce, int batch, long long npix,
                            int K, long long sb, long long sc, long long sn, int dtype, int tdtype, mlagg_stream_t stream);
int mlagg_dice_ce_stats_bwd(const void *logits, const void *target, const float *g_stats, const float *g_ce, void *dlogits,
                            int batch, long long npix, int K, long long sb, long long sc, long long sn, int dtype, int tdtype,
                            mlagg_stream_t stream);
/* y[pix][c] += bias[c] IN PLACE on a channels_last / tokens-major (pixels, C) map of n elements, C % 4 == 0 -- the bias of
 * the conv stages' nn.Conv2d / nn.ConvTranspose2d (nnUNetTrainer_MLAgg_2D_dt_MS.py:230-366, MambaSkip.py:712-716), which torch
 * adds after the cuDNN call with an un-vectorised broadcast kernel and differentiates with its generic reduction; the
 * gradient here is mlagg_colsum. */
int mlagg_bias_add_cl(void *y, const float *bias, long long n, int C, int dtype, mlagg_stream_t stream);
int mlagg_silu_gate_fwd(const void *t, const void *z, void *out, long long n, int dtype, mlagg_stream_t stream);
int mlagg_silu_gate_bwd(const void *t, const void *z, const void *dout, void *dt, void *dz, long long n, int dtype,
                        mlagg_stream_t stream);
int mlagg_diff_lambda_fwd(const float *lq1, const float *lk1, const float *lq2, const float *lk2, int n,
                          float lambda_init, float *out, mlagg_stream_t stream);
int mlagg_diff_lambda_bwd(const float *lq1, const float *lk1, const float *lq2, const float *lk2, const float *saved,
                          const float *dlam, int n, float *grads, mlagg_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MLAGG_B200_H */
