// capi.cu -- extern "C" entry points of libmlagg_b200.so (declared in include/mlagg_b200.h).
// Argument validation only; all work is enqueued on the caller's stream by the *_dispatch functions.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include "../../include/mlagg_b200.h"
#include "scan_common.cuh"
#include "gemm_tc.cuh"

namespace mlagg {
cudaError_t scan_fwd_dispatch(const ScanParams &p, bool bulk, int warps, cudaStream_t st);
cudaError_t scan_bwd_dispatch(const ScanParams &p, bool bulk, int warps, cudaStream_t st);
cudaError_t residual_scale_dispatch(const void *x, const void *y, const float *s, void *out, long long n,
                                    long long per_sample, int dtype, cudaStream_t st);
cudaError_t dice_ce_dispatch(const void *logits, const void *target, void *dlogits, float *stats, float *ce,
                             const float *g_stats, const float *g_ce, long long sb, long long sc, long long sn,
                             long long npix, int K, int Bn, int dtype, int tdtype, bool bwd, cudaStream_t st);
cudaError_t copy_rows_dispatch(const void *src, long long ld_s, long long bs_s, void *dst, long long ld_d, long long bs_d,
                               int batch, long long rows, int cols, int dtype, bool add, cudaStream_t st);
cudaError_t bias_add_cl_dispatch(void *y, const float *b, long long n, int C, int dtype, cudaStream_t st);
cudaError_t silu_gate_dispatch(const void *t, const void *z, const void *g, void *o1, void *o2, long long n, int dtype,
                               bool bwd, cudaStream_t st);
cudaError_t diff_lambda_dispatch(const float *q1, const float *k1, const float *q2, const float *k2, int n, float init,
                                 float *out, const float *dlam, float *grads, cudaStream_t st);
cudaError_t walk_pack_dispatch(const void *src, int dtype, long long ld_src, long long bs_src, int c0, int nc, float *dst,
                               long long bs_dst, int batch, int nstages, const int *Hs, const int *Ws, int col,
                               cudaStream_t st);
cudaError_t walk_unpack_dispatch(const float *src0, const float *src1, long long bs_src, int nc, int nc_pad, void *dst,
                                 int dtype, long long ld_dst, long long bs_dst, int c0, int batch, int nstages,
                                 const int *Hs, const int *Ws, int col, int accumulate, cudaStream_t st);

cudaError_t dwconv3x3_fwd_dispatch(const void *x, const float *w, const float *b, const void *res, void *y, int Bn, int H,
                                   int W, int C, long long ldx, long long bsx, long long ldr, long long bsr,
                                   long long ldy, long long bsy, int act, int res_mul, int dtype, cudaStream_t st);
cudaError_t dwconv3x3_bwd_dispatch(const void *x, const float *w, const float *b, const void *dy, void *dz, void *dx,
                                   float *dw, float *db, int Bn, int H, int W, int C, long long ldx, long long bsx,
                                   long long lddy, long long bsdy, long long lddx, long long bsdx, const void *mulv,
                                   void *dmul, long long ldm, long long bsm, int act, int dtype, cudaStream_t st);
cudaError_t causal_conv1d_fwd_dispatch(const float *x, const float *w, const float *b, float *y, int rows, int C,
                                       int L, int K, int act, cudaStream_t st);
cudaError_t causal_conv1d_bwd_dispatch(const float *x, const float *w, const float *b, const float *dy, float *dx,
                                       float *dw, float *db, int rows, int C, int L, int K, int act,
                                       cudaStream_t st);

struct LocalAttnParams {
    const void *q, *k, *v, *dout;
    void *out, *dq, *dk, *dv;
    const float *subln_w;
    float *d_subln_w, *d_lambda, *ws_dO, *ws_abar, *ws_dlog;
    long long ldq, ldkv, ldo, lddo, lddq, lddkv;
    int Bn, H, W, h;
    const float *lamp;
    float scale, eps, post;
};
bool local_attn_hd_supported(int hd);
cudaError_t local_attn_dispatch(const LocalAttnParams &p, int hd, int dtype, int which, cudaStream_t st);

struct PooledAttnParams {
    const void *q, *kp, *vp, *dout;
    void *out, *dq;
    float *lse, *dkp, *dvp;
    const float *subln_w;
    float *d_subln_w, *d_lambda, *ws_dO, *ws_D;
    const float *lamp;
    long long ldq, ldkv, ldo, lddo, lddq, ldd;
    int Bn, N, P, h;
    float scale2, eps, post;
};
cudaError_t pooled_attn_dispatch(const PooledAttnParams &p, int hd, int dtype, int which, cudaStream_t st);

struct LinAttnParams {
    const void *q, *k, *v, *dout;
    void *out, *dq, *dk, *dv;
    float *S, *kmean, *dS, *dkm;
    const float *rope;
    long long ldq, ldk, ldv, ldo, lddo, lddq, lddk, lddv;
    int Bn, H, W, h, chunk;
    float eps;
};
bool linattn_hd_supported(int hd);
cudaError_t linattn_dispatch(const LinAttnParams &p, int hd, int dtype, int which, cudaStream_t st);

cudaError_t colsum_dispatch(const void *x, float *out, long long M, int C, long long ld, int dtype, cudaStream_t st);



cudaError_t instnorm_dispatch(const void *x, const void *dy, const float *w, const float *b, void *out, float *stats,
                              float *sums, float *dw, float *db, int Bn, int N, int C, float eps, int act, float slope,
                              int dtype, bool bwd, cudaStream_t st, const void *res = nullptr, void *dres = nullptr);

cudaError_t avgpool_dispatch(const void *x, const void *dy, void *out, int Bn, int H, int W, int C, int pH, int pW,
                             int gelu, int dtype, bool bwd, cudaStream_t st);

cudaError_t layernorm_dispatch(const void *x, const float *w, const float *b, void *y, float *mean, float *rstd,
                               const void *dy, void *dx, float *dw, float *db, long long M, int C, float eps,
                               int dt_in, int dt_out, bool bwd, cudaStream_t st, const void *dres = nullptr);

static thread_local char g_last_err[256] = "";

static int fail_cuda(cudaError_t e) {
    snprintf(g_last_err, sizeof(g_last_err), "%s: %s", cudaGetErrorName(e), cudaGetErrorString(e));
    return MLAGG_ERR_LAUNCH;
}
static bool aligned(const void *p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

// consumer warps per CTA: 8 channels per warp; keep CTAs inside one group and fill >= 148 SMs when possible
static int pick_warps(int batch, int G, int dpg) {
    (void)batch; (void)G; (void)dpg;
    const char *env = getenv("MLAGG_SCAN_WARPS");   // 1 or 4 forces the forward kernel's CTA shape; default: the dispatch
    return (env && *env) ? atoi(env) : 0;           // chooses by grid size (scan_fwd.cu)
}
}  // namespace mlagg

using namespace mlagg;

extern "C" int mlagg_version(void) { return 100; }

extern "C" const char *mlagg_error_string(int code) {
    switch (code) {
        case MLAGG_OK: return "ok";
        case MLAGG_ERR_BAD_SHAPE: return "bad shape";
        case MLAGG_ERR_UNSUPPORTED: return "unsupported configuration";
        case MLAGG_ERR_NULL: return "required pointer is NULL";
        case MLAGG_ERR_LAUNCH: return "CUDA launch failed";
        case MLAGG_ERR_ALIGN: return "pointer not aligned to element size";
        default: return "unknown error";
    }
}

extern "C" const char *mlagg_last_cuda_error(void) { return g_last_err; }

extern "C" size_t mlagg_scan_ckpt_bytes(int batch, int dim, int seqlen, int dstate) {
    if (batch <= 0 || dim <= 0 || seqlen <= 0 || dstate <= 0) return 0;
    const size_t nchunks = ((size_t)seqlen + MLAGG_SCAN_CHUNK - 1) / MLAGG_SCAN_CHUNK;
    return (size_t)batch * nchunks * dim * dstate * sizeof(float);
}

static int scan_check(const void *u, const void *delta, const void *A, const void *B, const void *C, int batch,
                      int dim, int seqlen, int dstate, int ngroups) {
    if (!u || !delta || !A || !B || !C) return MLAGG_ERR_NULL;
    if (batch <= 0 || dim <= 0 || seqlen <= 0 || ngroups <= 0 || dim % ngroups != 0) return MLAGG_ERR_BAD_SHAPE;
    if (batch > 65535 || ngroups > 65535) return MLAGG_ERR_BAD_SHAPE;
    if (dstate != kN) return MLAGG_ERR_UNSUPPORTED;
    if (!aligned(u, 4) || !aligned(delta, 4) || !aligned(A, 4) || !aligned(B, 4) || !aligned(C, 4))
        return MLAGG_ERR_ALIGN;
    return MLAGG_OK;
}

extern "C" int mlagg_selective_scan_fwd(const float *u, const float *delta, const float *A, const float *B,
                                        const float *C, const float *D, const float *delta_bias, float *out,
                                        float *ckpt, float *last_state, int batch, int dim, int seqlen,
                                        int dstate, int ngroups, int delta_softplus, mlagg_stream_t stream) {
    int rc = scan_check(u, delta, A, B, C, batch, dim, seqlen, dstate, ngroups);
    if (rc) return rc;
    if (!out) return MLAGG_ERR_NULL;
    ScanParams p;
    memset(&p, 0, sizeof(p));
    p.u = u; p.delta = delta; p.A = A; p.B = B; p.C = C; p.D = D; p.bias = delta_bias;
    p.out = out; p.ckpt = ckpt; p.last_state = last_state;
    p.batch = batch; p.dim = dim; p.L = seqlen; p.G = ngroups; p.dpg = dim / ngroups;
    p.nchunks = (seqlen + kChunk - 1) / kChunk; p.softplus = delta_softplus;
    const bool bulk = seqlen % 4 == 0 && aligned(u, 16) && aligned(delta, 16) && aligned(B, 16) &&
                      aligned(C, 16) && aligned(out, 16) && (!ckpt || aligned(ckpt, 16)) &&
                      !getenv("MLAGG_SCAN_NO_BULK");
    if (ckpt && !aligned(ckpt, 16)) return MLAGG_ERR_ALIGN;
    cudaError_t e = scan_fwd_dispatch(p, bulk, pick_warps(batch, ngroups, p.dpg), (cudaStream_t)stream);
    return e == cudaSuccess ? MLAGG_OK : fail_cuda(e);
}

extern "C" int mlagg_selective_scan_bwd(const float *u, const float *delta, const float *A, const float *B,
                                        const float *C, const float *D, const float *delta_bias,
                                        const float *dout, const float *ckpt, float *du, float *ddelta,
                                        float *dA, float *dB, float *dC, float *dD, float *ddelta_bias,
                                        int batch, int dim, int seqlen, int dstate, int ngroups,
                                        int delta_softplus, mlagg_stream_t stream) {
    int rc = scan_check(u, delta, A, B, C, batch, dim, seqlen, dstate, ngroups);
    if (rc) return rc;
    if (!dout || !ckpt || !du || !ddelta || !dA || !dB || !dC) return MLAGG_ERR_NULL;
    if ((D && !dD) || (delta_bias && !ddelta_bias)) return MLAGG_ERR_NULL;
    if (!aligned(ckpt, 16)) return MLAGG_ERR_ALIGN;
    ScanParams p;
    memset(&p, 0, sizeof(p));
    p.u = u; p.delta = delta; p.A = A; p.B = B; p.C = C; p.D = D; p.bias = delta_bias;
    p.dout = dout; p.ckpt_in = ckpt;
    p.du = du; p.ddelta = ddelta; p.dA = dA; p.dB = dB; p.dC = dC; p.dD = D ? dD : nullptr;
    p.dbias = delta_bias ? ddelta_bias : nullptr;
    p.batch = batch; p.dim = dim; p.L = seqlen; p.G = ngroups; p.dpg = dim / ngroups;
    p.nchunks = (seqlen + kChunk - 1) / kChunk; p.softplus = delta_softplus;
    const bool bulk = seqlen % 4 == 0 && aligned(u, 16) && aligned(delta, 16) && aligned(B, 16) &&
                      aligned(C, 16) && aligned(dout, 16) && aligned(du, 16) && aligned(ddelta, 16) &&
                      !getenv("MLAGG_SCAN_NO_BULK");
    cudaError_t e = scan_bwd_dispatch(p, bulk, pick_warps(batch, ngroups, p.dpg), (cudaStream_t)stream);
    return e == cudaSuccess ? MLAGG_OK : fail_cuda(e);
}

static int dwconv_check(const void *x, const float *w, const void *y, int batch, int H, int W, int C, int dtype) {
    if (!x || !w || !y) return MLAGG_ERR_NULL;
    if (batch <= 0 || H <= 0 || W <= 0 || C <= 0 || C > 1024) return MLAGG_ERR_BAD_SHAPE;
    if (dtype != MLAGG_F32 && dtype != MLAGG_BF16) return MLAGG_ERR_UNSUPPORTED;
    const size_t a = C % 4 != 0 ? (dtype == MLAGG_F32 ? 4 : 2) : (dtype == MLAGG_F32 ? 16 : 8);
    if (!aligned(x, a) || !aligned(y, a) || !aligned(w, 4)) return MLAGG_ERR_ALIGN;
    return MLAGG_OK;
}

static bool dw_strides_ok(int C, long long ld, long long bs) { return ld >= C && bs >= 0 && (C % 4 != 0 || (ld % 4 == 0 && bs % 4 == 0)); }

extern "C" int mlagg_dwconv3x3_fwd_strided(const void *x, const float *weight, const float *bias, const void *residual,
                                           void *y, int batch, int H, int W, int C, long long ldx, long long bsx,
                                           long long ldr, long long bsr, long long ldy, long long bsy, int act_silu,
                                           int residual_mul, int dtype, mlagg_stream_t stream) {
    int rc = dwconv_check(x, weight, y, batch, H, W, C, dtype);
    if (rc) return rc;
    if (!dw_strides_ok(C, ldx, bsx) || !dw_strides_ok(C, ldy, bsy) || (residual && !dw_strides_ok(C, ldr, bsr)))
        return MLAGG_ERR_BAD_SHAPE;
    const size_t a = C % 4 != 0 ? (dtype == MLAGG_F32 ? 4 : 2) : (dtype == MLAGG_F32 ? 16 : 8);
    if ((bias && !aligned(bias, C % 4 ? 4 : 16)) || (residual && !aligned(residual, a))) return MLAGG_ERR_ALIGN;
    cudaError_t e = dwconv3x3_fwd_dispatch(x, weight, bias, residual, y, batch, H, W, C, ldx, bsx, ldr, bsr, ldy, bsy,
                                           act_silu, residual_mul, dtype, (cudaStream_t)stream);
    return e == cudaSuccess ? MLAGG_OK : fail_cuda(e);
}

extern "C" int mlagg_dwconv3x3_fwd(const void *x, const float *weight, const float *bias, void *y, int batch, int H,
                                   int W, int C, int act_silu, int dtype, mlagg_stream_t stream) {
    const long long bs = (long long)H * W * C;
    return mlagg_dwconv3x3_fwd_strided(x, weight, bias, nullptr, y, batch, H, W, C, C, bs, C, bs, C, bs, act_silu, 0, dtype,
                                       stream);
}

extern "C" int mlagg_dwconv3x3_bwd_strided(const void *x, const float *weight, const float *bias, const void *dy,
                                           void *dz_ws, void *dx, float *dweight, float *dbias, int batch, int H, int W,
                                           int C, long long ldx, long long bsx, long long lddy, long long bsdy,
                                           long long lddx, long long bsdx, const void *mul, void *dmul, long long ldm,
                                           long long bsm, int act_silu, int dtype, mlagg_stream_t stream) {
    int rc = dwconv_check(x, weight, dx, batch, H, W, C, dtype);
    if (rc) return rc;
    if (!dy || !dz_ws || !dweight) return MLAGG_ERR_NULL;
    if (!dw_strides_ok(C, ldx, bsx) || !dw_strides_ok(C, lddy, bsdy) || !dw_strides_ok(C, lddx, bsdx))
        return MLAGG_ERR_BAD_SHAPE;
    const size_t a = C % 4 != 0 ? (dtype == MLAGG_F32 ? 4 : 2) : (dtype == MLAGG_F32 ? 16 : 8);
    if (!aligned(dy, a) || !aligned(dz_ws, a) || (bias && !aligned(bias, C % 4 ? 4 : 16))) return MLAGG_ERR_ALIGN;
    if (mul) {
        if (!dmul) return MLAGG_ERR_NULL;
        if (!dw_strides_ok(C, ldm, bsm)) return MLAGG_ERR_BAD_SHAPE;
        if (!aligned(mul, a) || !aligned(dmul, a)) return MLAGG_ERR_ALIGN;
    }
    cudaError_t e = dwconv3x3_bwd_dispatch(x, weight, bias, dy, dz_ws, dx, dweight, dbias, batch, H, W, C, ldx, bsx, lddy,
                                           bsdy, lddx, bsdx, mul, mul ? dmul : nullptr, ldm, bsm, act_silu, dtype,
                                           (cudaStream_t)stream);
    return e == cudaSuccess ? MLAGG_OK : fail_cuda(e);
}

extern "C" int mlagg_dwconv3x3_bwd(const void *x, const float *weight, const float *bias, const void *dy,
                                   void *dz_ws, void *dx, float *dweight, float *dbias, int batch, int H, int W,
                                   int C, int act_silu, int dtype, mlagg_stream_t stream) {
    const long long bs = (long long)H * W * C;
    return mlagg_dwconv3x3_bwd_strided(x, weight, bias, dy, dz_ws, dx, dweight, dbias, batch, H, W, C, C, bs, C, bs, C, bs,
                                       nullptr, nullptr, 0, 0, act_silu, dtype, stream);
}

extern "C" int mlagg_causal_conv1d_fwd(const float *x, const float *weight, const float *bias, float *y, int batch,
                                       int C, int L, int K, int act_silu, mlagg_stream_t stream) {
    if (!x || !weight || !y) return MLAGG_ERR_NULL;
    if (batch <= 0 || C <= 0 || L <= 0 || (long long)batch * C > 65535) return MLAGG_ERR_BAD_SHAPE;
    if (K < 1 || K > 4) return MLAGG_ERR_UNSUPPORTED;
    cudaError_t e = causal_conv1d_fwd_dispatch(x, weight, bias, y, batch * C, C, L, K, act_silu, (cudaStream_t)stream);
    return e == cudaSuccess ? MLAGG_OK : fail_cuda(e);
}

extern "C" int mlagg_causal_conv1d_bwd(const float *x, const float *weight, const float *bias, const float *dy,
                                       float *dx, float *dweight, float *dbias, int batch, int C, int L, int K,
                                       int act_silu, mlagg_stream_t stream) {
    if (!x || !weight || !dy || !dx || !dweight) return MLAGG_ERR_NULL;
    if (batch <= 0 || C <= 0 || L <= 0 || (long long)batch * C > 65535) return MLAGG_ERR_BAD_SHAPE;
    if (K < 1 || K > 4) return MLAGG_ERR_UNSUPPORTED;
    cudaError_t e = causal_conv1d_bwd_dispatch(x, weight, bias, dy, dx, dweight, dbias, batch * C, C, L, K, act_silu,
                                               (cudaStream_t)stream);
    return e == cudaSuccess ? MLAGG_OK : fail_cuda(e);
}

extern "C" size_t mlagg_local_diffattn_ws_bytes(int batch, int H, int W, int heads, int head_dim) {
    if (batch <= 0 || H <= 0 || W <= 0 || heads <= 0 || head_dim <= 0) return 0;
    return (size_t)batch * H * W * heads * (2 * head_dim + 9 + 18) * sizeof(float);
}

static int local_check(const void *q, const void *k, const void *v, const float *w, int batch, int H, int W, int heads,
                       int head_dim, int dtype) {
    if (!q || !k || !v || !w) return MLAGG_ERR_NULL;
    if (batch <= 0 || H <= 0 || W <= 0 || heads <= 0) return MLAGG_ERR_BAD_SHAPE;
    if (!local_attn_hd_supported(head_dim) || (dtype != MLAGG_F32 && dtype != MLAGG_BF16)) return MLAGG_ERR_UNSUPPORTED;
    const size_t a = dtype == MLAGG_F32 ? 16 : 8;
    if (!aligned(q, a) || !aligned(k, a) || !aligned(v, a)) return MLAGG_ERR_ALIGN;
    return MLAGG_OK;
}

extern "C" int mlagg_local_diffattn_fwd(const void *q, const void *k, const void *v, const float *subln_w, void *out,
                                        int batch, int H, int W, int heads, int head_dim, long long ldq,
                                        long long ldkv, long long ldo, float scale, const float *lam, float eps,
                                        float post_scale, int dtype, mlagg_stream_t stream) {
    int rc = local_check(q, k, v, subln_w, batch, H, W, heads, head_dim, dtype);
    if (rc) return rc;
    if (!out || !lam) return MLAGG_ERR_NULL;
    if (ldq % 4 || ldkv % 4 || ldo % 4) return MLAGG_ERR_ALIGN;
    LocalAttnParams p;
    memset(&p, 0, sizeof(p));
    p.q = q; p.k = k; p.v = v; p.out = out; p.subln_w = subln_w;
    p.ldq = ldq; p.ldkv = ldkv; p.ldo = ldo;
    p.Bn = batch; p.H = H; p.W = W; p.h = heads;
    p.scale = scale; p.lamp = lam; p.eps = eps; p.post = post_scale;
    cudaError_t e = local_attn_dispatch(p, head_dim, dtype, 0, (cudaStream_t)stream);
    return e == cudaSuccess ? MLAGG_OK : fail_cuda(e);
}

extern "C" int mlagg_local_diffattn_bwd(const void *q, const void *k, const void *v, const float *subln_w,
                                        const void *dout, void *dq, void *dk, void *dv, float *d_subln_w,
                                        float *d_lambda, void *ws, int batch, int H, int W, int heads, int head_dim,
                                        long long ldq, long long ldkv, long long lddo, long long lddq,
                                        long long lddkv, float scale, const float *lam, float eps, float post_scale,
                                        int dtype, mlagg_stream_t stream) {
    int rc = local_check(q, k, v, subln_w, batch, H, W, heads, head_dim, dtype);
    if (rc) return rc;
    if (!dout || !dq || !dk || !dv || !d_subln_w || !d_lambda || !ws || !lam) return MLAGG_ERR_NULL;
    if (ldq % 4 || ldkv % 4 || lddo % 4 || lddq % 4 || lddkv % 4 || !aligned(ws, 16)) return MLAGG_ERR_ALIGN;
    LocalAttnParams p;
    memset(&p, 0, sizeof(p));
    p.q = q; p.k = k; p.v = v; p.dout = dout; p.dq = dq; p.dk = dk; p.dv = dv; p.subln_w = subln_w;
    p.d_subln_w = d_subln_w; p.d_lambda = d_lambda;
    const size_t ntok = (size_t)batch * H * W * heads;
    p.ws_dO = static_cast<float *>(ws);
    p.ws_abar = p.ws_dO + ntok * 2 * head_dim;
    p.ws_dlog = p.ws_abar + ntok * 9;
    p.ldq = ldq; p.ldkv = ldkv; p.lddo = lddo; p.lddq = lddq; p.lddkv = lddkv;
    p.Bn = batch; p.H = H; p.W = W; p.h = heads;
    p.scale = scale; p.lamp = lam; p.eps = eps; p.post = post_scale;
    cudaError_t e = local_attn_dispatch(p, head_dim, dtype, 1, (cudaStream_t)stream);
    if (e == cudaSuccess) e = local_attn_dispatch(p, head_dim, dtype, 2, (cudaStream_t)stream);
    return e == cudaSuccess ? MLAGG_OK : fail_cuda(e);
}

extern "C" size_t mlagg_pooled_diffattn_ws_bytes(int batch, int N, int heads, int head_dim) {
    if (batch <= 0 || N <= 0 || heads <= 0 || head_dim <= 0) return 0;
    return (size_t)batch * N * heads * (2 * head_dim + 2) * sizeof(float);
}

extern "C" size_t mlagg_pooled_diffattn_saved_bytes(int batch, int N, int heads, int head_dim) {
    if (batch <= 0 || N <= 0 || heads <= 0 || head_dim <= 0) return 0;
    const size_t lse = (((size_t)batch * N * heads * 2) + 3) & ~(size_t)3;       // floats, O rows stay 16 B aligned
    return (lse + (size_t)batch * N * heads * 4 * head_dim) * sizeof(float);
}

static int pooled_check(const void *q, const void *kp, const void *vp, const float *w, const float *lam, int batch,
                        int N, int P, int heads, int head_dim, int dtype) {
    if (!q || !kp || !vp || !w || !lam) return MLAGG_ERR_NULL;
    if (batch <= 0 || N <= 0 || P <= 0 || heads <= 0 || batch > 65535 || heads > 65535) return MLAGG_ERR_BAD_SHAPE;
    if (!local_attn_hd_supported(head_dim) || (dtype != MLAGG_F32 && dtype != MLAGG_BF16)) return MLAGG_ERR_UNSUPPORTED;
    if (P > 256 || (size_t)P * head_dim * 16 > 200 * 1024) return MLAGG_ERR_UNSUPPORTED;
    return MLAGG_OK;
}

extern "C" int mlagg_pooled_diffattn_fwd(const void *q, const void *kp, const void *vp, const float *subln_w,
                                         void *out, float *lse, int batch, int N, int P, int heads, int head_dim,
                                         long long ldq, long long ldkv, long long ldo, float scale, const float *lam,
                                         float eps, float post_scale, int dtype, mlagg_stream_t stream) {
    int rc = pooled_check(q, kp, vp, subln_w, lam, batch, N, P, heads, head_dim, dtype);
    if (rc) return rc;
    if (!out) return MLAGG_ERR_NULL;
    PooledAttnParams p;
    memset(&p, 0, sizeof(p));
    p.q = q; p.kp = kp; p.vp = vp; p.out = out; p.lse = lse; p.subln_w = subln_w; p.lamp = lam;
    p.ldq = ldq; p.ldkv = ldkv; p.ldo = ldo;
    p.Bn = batch; p.N = N; p.P = P; p.h = heads;
    p.scale2 = scale * scale; p.eps = eps; p.post = post_scale;
    cudaError_t e = pooled_attn_dispatch(p, head_dim, dtype, 0, (cudaStream_t)stream);
    return e == cudaSuccess ? MLAGG_OK : fail_cuda(e);
}

extern "C" int mlagg_pooled_diffattn_bwd(const void *q, const void *kp, const void *vp, const float *subln_w,
                                         const float *lse, const void *dout, void *dq, float *dkp, float *dvp,
                                         float *d_subln_w, float *d_lambda, void *ws, int batch, int N, int P,
                                         int heads, int head_dim, long long ldq, long long ldkv, long long lddo,
                                         long long lddq, long long ldd, float scale, const float *lam, float eps,
                                         float post_scale, int dtype, mlagg_stream_t stream) {
    int rc = pooled_check(q, kp, vp, subln_w, lam, batch, N, P, heads, head_dim, dtype);
    if (rc) return rc;
    if (!lse || !dout || !dq || !dkp || !dvp || !d_subln_w || !d_lambda || !ws) return MLAGG_ERR_NULL;
    if (!aligned(lse, 16) || !aligned(ws, 16)) return MLAGG_ERR_ALIGN;
    PooledAttnParams p;
    memset(&p, 0, sizeof(p));
    p.q = q; p.kp = kp; p.vp = vp; p.dout = dout; p.dq = dq; p.lse = const_cast<float *>(lse);
    p.dkp = dkp; p.dvp = dvp; p.subln_w = subln_w; p.d_subln_w = d_subln_w; p.d_lambda = d_lambda; p.lamp = lam;
    p.ws_dO = static_cast<float *>(ws);
    p.ws_D = p.ws_dO + (size_t)batch * N * heads * 2 * head_dim;
    p.ldq = ldq; p.ldkv = ldkv; p.lddo = lddo; p.lddq = lddq; p.ldd = ldd;
    p.Bn = batch; p.N = N; p.P = P; p.h = heads;
    p.scale2 = scale * scale; p.eps = eps; p.post = post_scale;
    cudaError_t e = pooled_attn_dispatch(p, head_dim, dtype, 1, (cudaStream_t)stream);
    if (e == cudaSuccess) e = pooled_attn_dispatch(p, head_dim, dtype, 2, (cudaStream_t)stream);
    return e == cudaSuccess ? MLAGG_OK : fail_cuda(e);
}

static int msmm_fill(ScanParams &p, const float *xrow, const float *xcol, const float *xdbl_row, const float *xdbl_col,
                     const float *Wdt, const float *dt_bias, const float *A, const float *Ds, int batch, int d_inner,
                     int dstate, int dt_rank, int nstages, const int *stage_lens) {
    if (!xrow || !xcol || !xdbl_row || !xdbl_col || !Wdt || !A || !stage_lens) return MLAGG_ERR_NULL;
    if (batch <= 0 || batch > 65535 || d_inner <= 0 || nstages <= 0 || nstages > kMaxStages) return MLAGG_ERR_BAD_SHAPE;
    if (dstate != kN || dt_rank < 1 || dt_rank > kMaxRk) return MLAGG_ERR_UNSUPPORTED;
    memset(&p, 0, sizeof(p));
    long long L = 0;
    for (int s = 0; s < nstages; ++s) {
        if (stage_lens[s] <= 0) return MLAGG_ERR_BAD_SHAPE;
        p.soff[s] = (int)L;
        L += stage_lens[s];
    }
    for (int s = nstages; s <= kMaxStages; ++s) p.soff[s] = (int)L;
    if (L > 0x7fffffff) return MLAGG_ERR_BAD_SHAPE;
    p.fused = 1; p.Rk = dt_rank; p.nstage = nstages;
    p.xrow = xrow; p.xcol = xcol; p.xdbl_row = xdbl_row; p.xdbl_col = xdbl_col; p.Wdt = Wdt;
    p.A = A; p.D = Ds; p.bias = dt_bias;
    p.batch = batch; p.dim = 4 * d_inner; p.L = (int)L; p.G = 4; p.dpg = d_inner;
    p.nchunks = ((int)L + kChunk - 1) / kChunk; p.softplus = 1;
    return MLAGG_OK;
}

extern "C" int mlagg_msmm_scan_fwd(const float *xrow, const float *xcol, const float *xdbl_row, const float *xdbl_col,
                                   const float *Wdt, const float *dt_bias, const float *A, const float *Ds, float *out,
                                   float *ckpt, int batch, int d_inner, int dstate, int dt_rank, int nstages,
                                   const int *stage_lens, mlagg_stream_t stream) {
    ScanParams p;
    int rc = msmm_fill(p, xrow, xcol, xdbl_row, xdbl_col, Wdt, dt_bias, A, Ds, batch, d_inner, dstate, dt_rank,
                       nstages, stage_lens);
    if (rc) return rc;
    if (!out) return MLAGG_ERR_NULL;
    if (ckpt && !aligned(ckpt, 16)) return MLAGG_ERR_ALIGN;
    p.out = out; p.ckpt = ckpt;
    cudaError_t e = scan_fwd_dispatch(p, false, pick_warps(batch, 4, d_inner), (cudaStream_t)stream);
    return e == cudaSuccess ? MLAGG_OK : fail_cuda(e);
}

extern "C" int mlagg_msmm_scan_bwd(const float *xrow, const float *xcol, const float *xdbl_row, const float *xdbl_col,
                                   const float *Wdt, const float *dt_bias, const float *A, const float *Ds,
                                   const float *dout, const float *ckpt, float *du, float *dxdbl_row,
                                   float *dxdbl_col, float *dWdt, float *ddt_bias, float *dA, float *dDs, int batch,
                                   int d_inner, int dstate, int dt_rank, int nstages, const int *stage_lens,
                                   int dout_walks, mlagg_stream_t stream) {
    ScanParams p;
    int rc = msmm_fill(p, xrow, xcol, xdbl_row, xdbl_col, Wdt, dt_bias, A, Ds, batch, d_inner, dstate, dt_rank,
                       nstages, stage_lens);
    if (rc) return rc;
    p.dout_walks = dout_walks ? 1 : 0;
    if (!dout || !ckpt || !du || !dxdbl_row || !dxdbl_col || !dWdt || !dA) return MLAGG_ERR_NULL;
    if ((Ds && !dDs) || (dt_bias && !ddt_bias)) return MLAGG_ERR_NULL;
    if (!aligned(ckpt, 16)) return MLAGG_ERR_ALIGN;
    p.dout = dout; p.ckpt_in = ckpt; p.du = du; p.dxdbl_row = dxdbl_row; p.dxdbl_col = dxdbl_col; p.dWdt = dWdt;
    p.dA = dA; p.dD = Ds ? dDs : nullptr; p.dbias = dt_bias ? ddt_bias : nullptr;
    cudaError_t e = scan_bwd_dispatch(p, false, 4, (cudaStream_t)stream);
    return e == cudaSuccess ? MLAGG_OK : fail_cuda(e);
}

extern "C" int mlagg_walk_pack(const void *src, int dtype, long long ld_src, long long bs_src, int c0, int nc,
                               float *dst, long long bs_dst, int batch, int nstages, const int *host_Hs,
                               const int *host_Ws, int col_walk, mlagg_stream_t stream) {
    if (!src || !dst || !host_Hs || !host_Ws) return MLAGG_ERR_NULL;
    if (batch <= 0 || batch > 65535 || nc <= 0 || c0 < 0 || nstages <= 0 || nstages > 8 || ld_src < c0 + nc)
        return MLAGG_ERR_BAD_SHAPE;
    if (dtype != MLAGG_F32 && dtype != MLAGG_BF16) return MLAGG_ERR_UNSUPPORTED;
    if (!aligned(src, dtype == MLAGG_F32 ? 4 : 2) || !aligned(dst, 4)) return MLAGG_ERR_ALIGN;
    cudaError_t e = walk_pack_dispatch(src, dtype, ld_src, bs_src, c0, nc, dst, bs_dst, batch, nstages, host_Hs, host_Ws,
                                       col_walk ? 1 : 0, (cudaStream_t)stream);
    if (e == cudaErrorInvalidValue) return MLAGG_ERR_BAD_SHAPE;
    return e == cudaSuccess ? MLAGG_OK : fail_cuda(e);
}

extern "C" int mlagg_walk_unpack(const float *src0, const float *src1, long long bs_src, int nc, int nc_pad, void *dst,
                                 int dtype, long long ld_dst, long long bs_dst, int c0, int batch, int nstages,
                                 const int *host_Hs, const int *host_Ws, int col_walk, int accumulate,
                                 mlagg_stream_t stream) {
    if (!src0 || !dst || !host_Hs || !host_Ws) return MLAGG_ERR_NULL;
    if (batch <= 0 || batch > 65535 || nc <= 0 || nc_pad < nc || c0 < 0 || nstages <= 0 || nstages > 8 ||
        ld_dst < c0 + nc_pad)
        return MLAGG_ERR_BAD_SHAPE;
    if (dtype != MLAGG_F32 && dtype != MLAGG_BF16) return MLAGG_ERR_UNSUPPORTED;
    if (!aligned(dst, dtype == MLAGG_F32 ? 4 : 2) || !aligned(src0, 4) || (src1 && !aligned(src1, 4))) return MLAGG_ERR_ALIGN;
    cudaError_t e = walk_unpack_dispatch(src0, src1, bs_src, nc, nc_pad, dst, dtype, ld_dst, bs_dst, c0, batch, nstages,
                                         host_Hs, host_Ws, col_walk ? 1 : 0, accumulate ? 1 : 0, (cudaStream_t)stream);
    if (e == cudaErrorInvalidValue) return MLAGG_ERR_BAD_SHAPE;
    return e == cudaSuccess ? MLAGG_OK : fail_cuda(e);
}

static int ew_check(const void *a, const void *b, const void *c, long long n, int dtype) {
    if (!a || !b || !c) return MLAGG_ERR_NULL;
    if (n <= 0 || n % 4 != 0) return MLAGG_ERR_BAD_SHAPE;
    if (dtype != MLAGG_F32 && dtype != MLAGG_BF16) return MLAGG_ERR_UNSUPPORTED;
    const size_t al = dtype == MLAGG_F32 ? 16 : 8;
    if (!aligned(a, al) || !aligned(b, al) || !aligned(c, al)) return MLAGG_ERR_ALIGN;
    return MLAGG_OK;
}

extern "C" int mlagg_residual_scale(const void *x, const void *y, const float *scale, void *out, long long n,
                                    long long per_sample, int dtype, mlagg_stream_t stream) {
    int rc = ew_check(y, out, x ? x : y, n, dtype);
    if (rc) return rc;
    if (per_sample <= 0 || per_sample % 4 != 0 || n % per_sample != 0) return MLAGG_ERR_BAD_SHAPE;
    cudaError_t e = residual_scale_dispatch(x, y, scale, out, n, per_sample, dtype, (cudaStream_t)stream);
    return e == cudaSuccess ? MLAGG_OK : fail_cuda(e);
}

extern "C" int mlagg_copy_rows(const void *src, long long ld_src, long long bs_src, void *dst, long long ld_dst,
                               long long bs_dst, int batch, long long rows, int cols, int dtype, mlagg_stream_t stream) {
    if (!src || !dst) return MLAGG_ERR_NULL;
    if (batch <= 0 || batch > 65535 || rows <= 0 || cols <= 0 || ld_src < cols || ld_dst < cols || bs_src < 0 || bs_dst < 0)
        return MLAGG_ERR_BAD_SHAPE;
    if (dtype != MLAGG_F32 && dtype != MLAGG_BF16) return MLAGG_ERR_UNSUPPORTED;
    const size_t es = dtype == MLAGG_F32 ? 4 : 2;
    if (!aligned(src, es) || !aligned(dst, es)) return MLAGG_ERR_ALIGN;
    cudaError_t e = copy_rows_dispatch(src, ld_src, bs_src, dst, ld_dst, bs_dst, batch, rows, cols, dtype, false,
                                       (cudaStream_t)stream);
    return e == cudaSuccess ? MLAGG_OK : fail_cuda(e);
}

// ------------------------------------------------------------------------------------------------ dice + CE statistics
static int loss_check(const void *logits, const void *target, int batch, long long npix, int K, long long sb, long long sc,
                      long long sn, int dtype, int tdtype) {
    if (!logits || !target) return MLAGG_ERR_NULL;
    if (batch <= 0 || batch > 65535 || npix <= 0 || K < 1 || sb < 0 || sc < 1 || sn < 1) return MLAGG_ERR_BAD_SHAPE;
    if (K > 32) return MLAGG_ERR_UNSUPPORTED;
    if (((dtype & 1) != MLAGG_F32 && (dtype & 1) != MLAGG_BF16) || dtype < 0 || dtype > 3 || (tdtype != 0 && tdtype != 1)) return MLAGG_ERR_UNSUPPORTED;
    return MLAGG_OK;
}

extern "C" int mlagg_dice_ce_stats_fwd(const void *logits, const void *target, float *stats, float *ce, int batch,
                                       long long npix, int K, long long sb, long long sc, long long sn, int dtype,
                                       int tdtype, mlagg_stream_t stream) {
    int rc = loss_check(logits, target, batch, npix, K, sb, sc, sn, dtype, tdtype);
    if (rc) return rc;
    if (!stats || !ce) return MLAGG_ERR_NULL;
    cudaError_t e = dice_ce_dispatch(logits, target, nullptr, stats, ce, nullptr, nullptr, sb, sc, sn, npix, K, batch, dtype,
                                     tdtype, false, (cudaStream_t)stream);
    return e == cudaSuccess ? MLAGG_OK : fail_cuda(e);
}

extern "C" int mlagg_dice_ce_stats_bwd(const void *logits, const void *target, const float *g_stats, const float *g_ce,
                                       void *dlogits, int batch, long long npix, int K, long long sb, long long sc,
                                       long long sn, int dtype, int tdtype, mlagg_stream_t stream) {
    int rc = loss_check(logits, target, batch, npix, K, sb, sc, sn, dtype, tdtype);
    if (rc) return rc;
    if (!g_stats || !g_ce || !dlogits) return MLAGG_ERR_NULL;
    cudaError_t e = dice_ce_dispatch(logits, target, dlogits, nullptr, nullptr, g_stats, g_ce, sb, sc, sn, npix, K, batch, dtype,
                                     tdtype, true, (cudaStream_t)stream);
    return e == cudaSuccess ? MLAGG_OK : fail_cuda(e);
}

extern "C" int mlagg_add_rows(const void *src, long long ld_src, long long bs_src, void *dst, long long ld_dst,
                              long long bs_dst, int batch, long long rows, int cols, int dtype, mlagg_stream_t stream) {
    if (!src || !dst) return MLAGG_ERR_NULL;
    if (batch <= 0 || batch > 65535 || rows <= 0 || cols <= 0 || ld_src < cols || ld_dst < cols || bs_src < 0 || bs_dst < 0)
        return MLAGG_ERR_BAD_SHAPE;
    if (dtype != MLAGG_F32 && dtype != MLAGG_BF16) return MLAGG_ERR_UNSUPPORTED;
    const size_t es = dtype == MLAGG_F32 ? 4 : 2;
    // 8-byte vectors at least, at most 256 of them per row (the caller falls back to its own add otherwise)
    const int v = (int)(8 / es);
    if (cols % v || ld_src % v || ld_dst % v || bs_src % v || bs_dst % v || !aligned(src, 8) || !aligned(dst, 8)) return MLAGG_ERR_ALIGN;
    if (cols / v > 512) return MLAGG_ERR_UNSUPPORTED;
    cudaError_t e = copy_rows_dispatch(src, ld_src, bs_src, dst, ld_dst, bs_dst, batch, rows, cols, dtype, true,
                                       (cudaStream_t)stream);
    if (e == cudaErrorNotSupported) return MLAGG_ERR_UNSUPPORTED;
    return e == cudaSuccess ? MLAGG_OK : fail_cuda(e);
}

extern "C" int mlagg_bias_add_cl(void *y, const float *bias, long long n, int C, int dtype, mlagg_stream_t stream) {
    int rc = ew_check(y, y, y, n, dtype);
    if (rc) return rc;
    if (!bias) return MLAGG_ERR_NULL;
    if (C <= 0 || C % 4 != 0 || n % C != 0) return MLAGG_ERR_BAD_SHAPE;
    if (!aligned(bias, 16)) return MLAGG_ERR_ALIGN;
    cudaError_t e = bias_add_cl_dispatch(y, bias, n, C, dtype, (cudaStream_t)stream);
    return e == cudaSuccess ? MLAGG_OK : fail_cuda(e);
}

extern "C" int mlagg_silu_gate_fwd(const void *t, const void *z, void *out, long long n, int dtype,
                                   mlagg_stream_t stream) {
    int rc = ew_check(t, z, out, n, dtype);
    if (rc) return rc;
    cudaError_t e = silu_gate_dispatch(t, z, nullptr, out, nullptr, n, dtype, false, (cudaStream_t)stream);
    return e == cudaSuccess ? MLAGG_OK : fail_cuda(e);
}

extern "C" int mlagg_silu_gate_bwd(const void *t, const void *z, const void *dout, void *dt, void *dz, long long n,
                                   int dtype, mlagg_stream_t stream) {
    int rc = ew_check(t, z, dout, n, dtype);
    if (rc) return rc;
    rc = ew_check(dt, dz, dz, n, dtype);
    if (rc) return rc;
    cudaError_t e = silu_gate_dispatch(t, z, dout, dt, dz, n, dtype, true, (cudaStream_t)stream);
    return e == cudaSuccess ? MLAGG_OK : fail_cuda(e);
}

extern "C" int mlagg_diff_lambda_fwd(const float *lq1, const float *lk1, const float *lq2, const float *lk2, int n,
                                     float lambda_init, float *out, mlagg_stream_t stream) {
    if (!lq1 || !lk1 || !lq2 || !lk2 || !out) return MLAGG_ERR_NULL;
    if (n <= 0) return MLAGG_ERR_BAD_SHAPE;
    cudaError_t e = diff_lambda_dispatch(lq1, lk1, lq2, lk2, n, lambda_init, out, nullptr, nullptr, (cudaStream_t)stream);
    return e == cudaSuccess ? MLAGG_OK : fail_cuda(e);
}

extern "C" int mlagg_diff_lambda_bwd(const float *lq1, const float *lk1, const float *lq2, const float *lk2,
                                     const float *saved, const float *dlam, int n, float *grads,
                                     mlagg_stream_t stream) {
    if (!lq1 || !lk1 || !lq2 || !lk2 || !saved || !dlam || !grads) return MLAGG_ERR_NULL;
    if (n <= 0) return MLAGG_ERR_BAD_SHAPE;
    cudaError_t e = diff_lambda_dispatch(lq1, lk1, lq2, lk2, n, 0.f, const_cast<float *>(saved), dlam, grads,
                                         (cudaStream_t)stream);
    return e == cudaSuccess ? MLAGG_OK : fail_cuda(e);
}

static int ln_check(const void *x, const float *w, long long M, int C, int dt_in, int dt_out) {
    if (!x || !w) return MLAGG_ERR_NULL;
    if (M <= 0 || C <= 0 || C % 4 != 0 || C > 1024) return MLAGG_ERR_BAD_SHAPE;
    if ((dt_in != MLAGG_F32 && dt_in != MLAGG_BF16) || (dt_out != MLAGG_F32 && dt_out != MLAGG_BF16)) return MLAGG_ERR_UNSUPPORTED;
    if (!aligned(x, dt_in == MLAGG_F32 ? 16 : 8) || !aligned(w, 16)) return MLAGG_ERR_ALIGN;
    return MLAGG_OK;
}

extern "C" int mlagg_layernorm_fwd(const void *x, const float *weight, const float *bias, void *y, float *mean,
                                   float *rstd, long long M, int C, float eps, int dt_in, int dt_out,
                                   mlagg_stream_t stream) {
    int rc = ln_check(x, weight, M, C, dt_in, dt_out);
    if (rc) return rc;
    if (!y || !mean || !rstd) return MLAGG_ERR_NULL;
    if (!aligned(y, dt_out == MLAGG_F32 ? 16 : 8) || (bias && !aligned(bias, 16))) return MLAGG_ERR_ALIGN;
    cudaError_t e = layernorm_dispatch(x, weight, bias, y, mean, rstd, nullptr, nullptr, nullptr, nullptr, M, C, eps,
                                       dt_in, dt_out, false, (cudaStream_t)stream);
    return e == cudaSuccess ? MLAGG_OK : fail_cuda(e);
}

extern "C" int mlagg_layernorm_bwd_res(const void *x, const float *weight, const float *mean, const float *rstd,
                                       const void *dy, const void *dres, void *dx, float *dweight, float *dbias,
                                       long long M, int C, int dt_in, int dt_out, mlagg_stream_t stream) {
    int rc = ln_check(x, weight, M, C, dt_in, dt_out);
    if (rc) return rc;
    if (!mean || !rstd || !dy || !dx || !dweight) return MLAGG_ERR_NULL;
    if (!aligned(dy, dt_out == MLAGG_F32 ? 16 : 8) || !aligned(dx, dt_in == MLAGG_F32 ? 16 : 8)) return MLAGG_ERR_ALIGN;
    if (dres && !aligned(dres, dt_in == MLAGG_F32 ? 16 : 8)) return MLAGG_ERR_ALIGN;
    cudaError_t e = layernorm_dispatch(x, weight, nullptr, nullptr, const_cast<float *>(mean), const_cast<float *>(rstd),
                                       dy, dx, dweight, dbias, M, C, 0.f, dt_in, dt_out, true, (cudaStream_t)stream, dres);
    return e == cudaSuccess ? MLAGG_OK : fail_cuda(e);
}

extern "C" int mlagg_layernorm_bwd(const void *x, const float *weight, const float *mean, const float *rstd,
                                   const void *dy, void *dx, float *dweight, float *dbias, long long M, int C,
                                   int dt_in, int dt_out, mlagg_stream_t stream) {
    return mlagg_layernorm_bwd_res(x, weight, mean, rstd, dy, nullptr, dx, dweight, dbias, M, C, dt_in, dt_out, stream);
}

// ------------------------------------------------------------------------------------------------ linear attention
extern "C" size_t mlagg_linattn_state_bytes(int batch, int heads, int head_dim) {
    if (batch <= 0 || heads <= 0 || head_dim <= 0) return 0;
    return (size_t)batch * heads * (head_dim * head_dim + head_dim) * sizeof(float);
}

static int linattn_check(int batch, int H, int W, int heads, int head_dim, int dtype, const long long *lds, int nld,
                         const void *const *ptrs, int nptr) {
    for (int i = 0; i < nptr; ++i)
        if (!ptrs[i]) return MLAGG_ERR_NULL;
    if (batch <= 0 || H <= 0 || W <= 0 || heads <= 0 || batch > 65535 || heads > 65535) return MLAGG_ERR_BAD_SHAPE;
    if (!linattn_hd_supported(head_dim) || (dtype != MLAGG_F32 && dtype != MLAGG_BF16)) return MLAGG_ERR_UNSUPPORTED;
    if ((heads * head_dim) % 4 != 0) return MLAGG_ERR_BAD_SHAPE;
    const size_t es = dtype == MLAGG_F32 ? 4 : 2;
    for (int i = 0; i < nld; ++i)
        if (lds[i] < (long long)heads * head_dim || (lds[i] * es) % 16 != 0) return MLAGG_ERR_ALIGN;
    for (int i = 0; i < nptr; ++i)
        if (!aligned(ptrs[i], 16)) return MLAGG_ERR_ALIGN;
    return MLAGG_OK;
}

extern "C" int mlagg_linattn_fwd(const void *q, const void *k, const void *v, const float *rope_cs, void *out,
                                 float *state, int batch, int H, int W, int heads, int head_dim, long long ldq,
                                 long long ldk, long long ldv, long long ldo, float eps, int dtype,
                                 mlagg_stream_t stream) {
    const long long lds[] = {ldq, ldk, ldv, ldo};
    const void *ptrs[] = {q, k, v, rope_cs, out, state};
    int rc = linattn_check(batch, H, W, heads, head_dim, dtype, lds, 4, ptrs, 6);
    if (rc) return rc;
    LinAttnParams p;
    memset(&p, 0, sizeof(p));
    p.q = q; p.k = k; p.v = v; p.rope = rope_cs; p.out = out;
    p.S = state; p.kmean = state + (size_t)batch * heads * head_dim * head_dim;
    p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.ldo = ldo;
    p.Bn = batch; p.H = H; p.W = W; p.h = heads; p.eps = eps;
    cudaError_t e = cudaMemsetAsync(state, 0, mlagg_linattn_state_bytes(batch, heads, head_dim), (cudaStream_t)stream);
    if (e == cudaSuccess) e = linattn_dispatch(p, head_dim, dtype, 0, (cudaStream_t)stream);
    return e == cudaSuccess ? MLAGG_OK : fail_cuda(e);
}

extern "C" int mlagg_linattn_bwd(const void *q, const void *k, const void *v, const float *rope_cs, const float *state,
                                 const void *dout, void *dq, void *dk, void *dv, float *ws, int batch, int H, int W,
                                 int heads, int head_dim, long long ldq, long long ldk, long long ldv, long long lddo,
                                 long long lddq, long long lddk, long long lddv, float eps, int dtype,
                                 mlagg_stream_t stream) {
    const long long lds[] = {ldq, ldk, ldv, lddo, lddq, lddk, lddv};
    const void *ptrs[] = {q, k, v, rope_cs, state, dout, dq, dk, dv, ws};
    int rc = linattn_check(batch, H, W, heads, head_dim, dtype, lds, 7, ptrs, 10);
    if (rc) return rc;
    LinAttnParams p;
    memset(&p, 0, sizeof(p));
    p.q = q; p.k = k; p.v = v; p.rope = rope_cs; p.dout = dout; p.dq = dq; p.dk = dk; p.dv = dv;
    p.S = const_cast<float *>(state); p.kmean = p.S + (size_t)batch * heads * head_dim * head_dim;
    p.dS = ws; p.dkm = ws + (size_t)batch * heads * head_dim * head_dim;
    p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.lddo = lddo; p.lddq = lddq; p.lddk = lddk; p.lddv = lddv;
    p.Bn = batch; p.H = H; p.W = W; p.h = heads; p.eps = eps;
    cudaError_t e = cudaMemsetAsync(ws, 0, mlagg_linattn_state_bytes(batch, heads, head_dim), (cudaStream_t)stream);
    if (e == cudaSuccess) e = linattn_dispatch(p, head_dim, dtype, 1, (cudaStream_t)stream);
    return e == cudaSuccess ? MLAGG_OK : fail_cuda(e);
}

// ------------------------------------------------------------------------------------------------ tensor-core GEMMs
static int gemm_check(const void *a, long long lda, const void *b, long long ldb, const void *o, long long ldo, long long M,
                      int N, int K, int o_elt) {
    if (!a || !b || !o) return MLAGG_ERR_NULL;
    if (M <= 0 || M > 0x7fffffffLL || N <= 0 || K <= 0 || N % 8 != 0 || K % 8 != 0) return MLAGG_ERR_BAD_SHAPE;
    if (lda % 8 != 0 || ldb % 8 != 0 || ldo * o_elt % 16 != 0) return MLAGG_ERR_BAD_SHAPE;
    if (!aligned(a, 16) || !aligned(b, 16) || !aligned(o, 16)) return MLAGG_ERR_ALIGN;
    return MLAGG_OK;
}

extern "C" int mlagg_linear_fwd(const void *x, long long ldx, const void *w, long long ldw, const float *bias, void *y,
                                long long ldy, void *pre, long long ldpre, long long M, int N, int K, int act, int out_dtype,
                                mlagg_stream_t stream) {
    if (out_dtype != MLAGG_F32 && out_dtype != MLAGG_BF16) return MLAGG_ERR_UNSUPPORTED;
    int rc = gemm_check(x, ldx, w, ldw, y, ldy, M, N, K, out_dtype == MLAGG_F32 ? 4 : 2);
    if (rc) return rc;
    if (ldx < K || ldw < K || ldy < N || act < 0 || act > 2) return MLAGG_ERR_BAD_SHAPE;
    if (pre && (ldpre < N || ldpre % 8 != 0 || !aligned(pre, 16))) return MLAGG_ERR_ALIGN;
    GemmTcParams p{};
    p.out = y; p.ldo = ldy; p.pre = pre; p.ldpre = ldpre; p.bias = bias;
    p.M = (int)M; p.N = N; p.K = K; p.act = act; p.out_f32 = out_dtype == MLAGG_F32;
    cudaError_t e = gemm_tc_dispatch(x, ldx, 0, w, ldw, 0, p, (cudaStream_t)stream);
    return e == cudaSuccess ? MLAGG_OK : fail_cuda(e);
}

extern "C" int mlagg_linear_bwd_data(const void *dy, long long lddy, const void *w, long long ldw, const void *aux,
                                     long long ldaux, int act, void *dx, long long lddx, long long M, int N, int K,
                                     int out_dtype, mlagg_stream_t stream) {
    if (out_dtype != MLAGG_F32 && out_dtype != MLAGG_BF16) return MLAGG_ERR_UNSUPPORTED;
    int rc = gemm_check(dy, lddy, w, ldw, dx, lddx, M, N, K, out_dtype == MLAGG_F32 ? 4 : 2);
    if (rc) return rc;
    if (lddy < N || ldw < K || lddx < K || act < 0 || act > 2) return MLAGG_ERR_BAD_SHAPE;
    if (aux && (ldaux < K || ldaux % 8 != 0 || !aligned(aux, 16))) return MLAGG_ERR_ALIGN;
    GemmTcParams p{};                       // D[M, K_in] = dy[M, N_out] . w[N_out, K_in]: contraction over N_out
    p.out = dx; p.ldo = lddx; p.aux = aux; p.ldaux = ldaux;
    p.M = (int)M; p.N = K; p.K = N; p.act = aux ? act : 0; p.out_f32 = out_dtype == MLAGG_F32;
    cudaError_t e = gemm_tc_dispatch(dy, lddy, 0, w, ldw, 1, p, (cudaStream_t)stream);
    return e == cudaSuccess ? MLAGG_OK : fail_cuda(e);
}

extern "C" int mlagg_linear_bwd_weight(const void *dy, long long lddy, const void *x, long long ldx, float *dw,
                                       long long lddw, float *db, long long M, int N, int K, mlagg_stream_t stream) {
    int rc = gemm_check(dy, lddy, x, ldx, dw, lddw, M, N, K, 4);
    if (rc) return rc;
    if (lddy < N || ldx < K || lddw < K) return MLAGG_ERR_BAD_SHAPE;
    GemmTcParams p{};                       // D[N_out, K_in] += dy[M, N_out]^T . x[M, K_in]: contraction over the tokens
    p.out = dw; p.ldo = lddw; p.colsum = db;
    p.M = N; p.N = K; p.K = (int)M; p.reduce = 1;
    cudaError_t e = gemm_tc_dispatch(dy, lddy, 1, x, ldx, 1, p, (cudaStream_t)stream);
    return e == cudaSuccess ? MLAGG_OK : fail_cuda(e);
}

// ------------------------------------------------------------------------------------------------ column sums
extern "C" int mlagg_colsum(const void *x, float *out, long long M, int C, long long ld, int dtype,
                            mlagg_stream_t stream) {
    if (!x || !out) return MLAGG_ERR_NULL;
    if (M <= 0 || C <= 0 || ld < C) return MLAGG_ERR_BAD_SHAPE;
    if (dtype != MLAGG_F32 && dtype != MLAGG_BF16) return MLAGG_ERR_UNSUPPORTED;
    if (!aligned(x, dtype == MLAGG_F32 ? 4 : 2) || !aligned(out, 4)) return MLAGG_ERR_ALIGN;
    cudaError_t e = colsum_dispatch(x, out, M, C, ld, dtype, (cudaStream_t)stream);
    return e == cudaSuccess ? MLAGG_OK : fail_cuda(e);
}

// ------------------------------------------------------------------------------------------------ instance norm
static int instnorm_check(const void *x, const void *y, const float *stats, int batch, int N, int C, int dtype) {
    if (!x || !y || !stats) return MLAGG_ERR_NULL;
    if (batch <= 0 || batch > 65535 || N <= 0 || C <= 0 || C % 4 != 0) return MLAGG_ERR_BAD_SHAPE;
    if (dtype != MLAGG_F32 && dtype != MLAGG_BF16) return MLAGG_ERR_UNSUPPORTED;
    const size_t a = dtype == MLAGG_F32 ? 16 : 8;
    if (!aligned(x, a) || !aligned(y, a) || !aligned(stats, 8)) return MLAGG_ERR_ALIGN;
    return MLAGG_OK;
}

extern "C" int mlagg_instnorm_fwd(const void *x, const float *w, const float *b, void *y, float *stats, int batch,
                                  int N, int C, float eps, int act, float slope, int dtype, mlagg_stream_t stream) {
    int rc = instnorm_check(x, y, stats, batch, N, C, dtype);
    if (rc) return rc;
    if (act < 0 || act > 2) return MLAGG_ERR_UNSUPPORTED;
    cudaError_t e = cudaMemsetAsync(stats, 0, (size_t)batch * C * 2 * sizeof(float), (cudaStream_t)stream);
    if (e == cudaSuccess)
        e = instnorm_dispatch(x, nullptr, w, b, y, stats, nullptr, nullptr, nullptr, batch, N, C, eps, act, slope, dtype, false,
                              (cudaStream_t)stream);
    return e == cudaSuccess ? MLAGG_OK : fail_cuda(e);
}

extern "C" int mlagg_instnorm_bwd(const void *x, const float *w, const float *b, const float *stats, const void *dy,
                                  void *dx, float *sums, float *dw, float *db, int batch, int N, int C, int act,
                                  float slope, int dtype, mlagg_stream_t stream) {
    int rc = instnorm_check(x, dx, stats, batch, N, C, dtype);
    if (rc) return rc;
    if (!dy || !sums) return MLAGG_ERR_NULL;
    if (act < 0 || act > 2) return MLAGG_ERR_UNSUPPORTED;
    cudaError_t e = cudaMemsetAsync(sums, 0, (size_t)batch * C * 2 * sizeof(float), (cudaStream_t)stream);
    if (e == cudaSuccess)
        e = instnorm_dispatch(x, dy, w, b, dx, const_cast<float *>(stats), sums, dw, db, batch, N, C, 0.f, act, slope,
                              dtype, true, (cudaStream_t)stream);
    return e == cudaSuccess ? MLAGG_OK : fail_cuda(e);
}

// y = act(norm(x) + residual): the tail of monai's UnetResBlock (`lrelu(norm2(conv2(.)) + residual)`) in the apply pass
extern "C" int mlagg_instnorm_res_fwd(const void *x, const float *w, const float *b, const void *residual, void *y,
                                      float *stats, int batch, int N, int C, float eps, int act, float slope, int dtype,
                                      mlagg_stream_t stream) {
    int rc = instnorm_check(x, y, stats, batch, N, C, dtype);
    if (rc) return rc;
    if (!residual) return MLAGG_ERR_NULL;
    if (act < 0 || act > 1 || C % (dtype == MLAGG_F32 ? 4 : 8) != 0 || C / (dtype == MLAGG_F32 ? 4 : 8) > 256) return MLAGG_ERR_UNSUPPORTED;
    if (!aligned(x, 16) || !aligned(y, 16) || !aligned(residual, 16)) return MLAGG_ERR_ALIGN;
    cudaError_t e = cudaMemsetAsync(stats, 0, (size_t)batch * C * 2 * sizeof(float), (cudaStream_t)stream);
    if (e == cudaSuccess)
        e = instnorm_dispatch(x, nullptr, w, b, y, stats, nullptr, nullptr, nullptr, batch, N, C, eps, act, slope, dtype, false,
                              (cudaStream_t)stream, residual, nullptr);
    if (e == cudaErrorNotSupported) return MLAGG_ERR_UNSUPPORTED;
    return e == cudaSuccess ? MLAGG_OK : fail_cuda(e);
}

extern "C" int mlagg_instnorm_res_bwd(const void *x, const float *w, const float *b, const float *stats, const void *y,
                                      const void *dy, void *dx, void *dresidual, float *sums, float *dw, float *db,
                                      int batch, int N, int C, int act, float slope, int dtype, mlagg_stream_t stream) {
    int rc = instnorm_check(x, dx, stats, batch, N, C, dtype);
    if (rc) return rc;
    if (!dy || !sums || !y || !dresidual) return MLAGG_ERR_NULL;
    if (act < 0 || act > 1 || C % (dtype == MLAGG_F32 ? 4 : 8) != 0 || C / (dtype == MLAGG_F32 ? 4 : 8) > 256) return MLAGG_ERR_UNSUPPORTED;
    if (!aligned(x, 16) || !aligned(y, 16) || !aligned(dy, 16) || !aligned(dx, 16) || !aligned(dresidual, 16)) return MLAGG_ERR_ALIGN;
    cudaError_t e = cudaMemsetAsync(sums, 0, (size_t)batch * C * 2 * sizeof(float), (cudaStream_t)stream);
    if (e == cudaSuccess)
        e = instnorm_dispatch(x, dy, w, b, dx, const_cast<float *>(stats), sums, dw, db, batch, N, C, 0.f, act, slope,
                              dtype, true, (cudaStream_t)stream, y, dresidual);
    if (e == cudaErrorNotSupported) return MLAGG_ERR_UNSUPPORTED;
    return e == cudaSuccess ? MLAGG_OK : fail_cuda(e);
}

// ------------------------------------------------------------------------------------------------ adaptive avg pool
static int avgpool_check(const void *x, const void *y, int batch, int H, int W, int C, int pH, int pW, int dtype) {
    if (!x || !y) return MLAGG_ERR_NULL;
    if (batch <= 0 || batch > 65535 || H <= 0 || W <= 0 || pH <= 0 || pW <= 0 || pH > H || pW > W || C <= 0 || C % 4)
        return MLAGG_ERR_BAD_SHAPE;
    if (C > 4096 || (dtype != MLAGG_F32 && dtype != MLAGG_BF16)) return MLAGG_ERR_UNSUPPORTED;
    const size_t a = dtype == MLAGG_F32 ? 16 : 8;
    if (!aligned(x, a) || !aligned(y, a)) return MLAGG_ERR_ALIGN;
    return MLAGG_OK;
}
extern "C" int mlagg_avgpool_tokens_fwd(const void *x, void *y, int batch, int H, int W, int C, int pH, int pW,
                                        int act_gelu, int dtype, mlagg_stream_t stream) {
    int rc = avgpool_check(x, y, batch, H, W, C, pH, pW, dtype);
    if (rc) return rc;
    cudaError_t e = avgpool_dispatch(x, nullptr, y, batch, H, W, C, pH, pW, act_gelu, dtype, false, (cudaStream_t)stream);
    return e == cudaSuccess ? MLAGG_OK : fail_cuda(e);
}
extern "C" int mlagg_avgpool_tokens_bwd(const void *x, const void *dy, void *dx, int batch, int H, int W, int C, int pH,
                                        int pW, int act_gelu, int dtype, mlagg_stream_t stream) {
    int rc = avgpool_check(x, dx, batch, H, W, C, pH, pW, dtype);
    if (rc) return rc;
    if (!dy) return MLAGG_ERR_NULL;
    cudaError_t e = avgpool_dispatch(x, dy, dx, batch, H, W, C, pH, pW, act_gelu, dtype, true, (cudaStream_t)stream);
    return e == cudaSuccess ? MLAGG_OK : fail_cuda(e);
}
