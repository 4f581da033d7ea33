/*
 * oracle/scan_ref.c -- TEST INFRASTRUCTURE ONLY (parity oracle + CPU baseline).
 *
 * CPU restatement of the S6 selective scan as the reference parameterises it at
 *   mlagg/nnunetv2/training/nnUNetTrainer/variants/mamba/MambaSkip.py:437-451
 * (real A, time-varying B/C shared by groups of D/G channels, delta_bias added
 * before softplus, D skip, z=None).  The arithmetic itself lives in the
 * un-vendored, un-pinned dependency mamba-ssm (selective_scan_fn /
 * selective_scan_ref, README.md:49-50 of the reference); this file restates that
 * published recurrence and the analytic gradient of SURVEY.md App. A.1 / A.2.
 * PARITY UNPINNED against mamba-ssm itself (the package is absent everywhere);
 * pinned against the reference call site via tests/golden (see make_golden.py).
 *
 * Nothing in the product (mlagg-unet_b200/) may link, import or call this file.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs use it.
 *
 * Built twice by oracle/Makefile: REAL=float -> scan_ref_f32_*, REAL=double ->
 * scan_ref_f64_* (inputs are always fp32 buffers; REAL is the arithmetic type).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef REAL
#define REAL float
#endif
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(CAT(scan_ref_, SUFFIX), name)

static inline REAL r_exp(REAL x) { return sizeof(REAL) == 4 ? (REAL)expf((float)x) : (REAL)exp((double)x); }
static inline REAL r_log1p(REAL x) { return sizeof(REAL) == 4 ? (REAL)log1pf((float)x) : (REAL)log1p((double)x); }

/* torch.nn.functional.softplus(beta=1, threshold=20) */
static inline REAL softplus(REAL x) { return x > (REAL)20 ? x : r_log1p(r_exp(x)); }
static inline REAL sigmoid(REAL x) { return (REAL)1 / ((REAL)1 + r_exp(-x)); }

/*
 * Forward.  Shapes (all contiguous, fp32):
 *   u, delta, out : (Bn, D, L)      A : (D, N)     Bm, Cm : (Bn, G, N, L)
 *   Dskip, dbias  : (D) or NULL     last_state : (Bn, D, N) or NULL
 * group of channel d is d / (D / G).
 */
void FN(_fwd)(const float *u, const float *delta, const float *A, const float *Bm, const float *Cm,
              const float *Dskip, const float *dbias, int delta_softplus, int Bn, int D, int L, int N,
              int G, float *out, float *last_state)
{
    const int dpg = D / G;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < Bn; ++b) {
        for (int d = 0; d < D; ++d) {
            const int g = d / dpg;
            const float *ur = u + ((size_t)b * D + d) * L;
            const float *dr = delta + ((size_t)b * D + d) * L;
            const float *Br = Bm + ((size_t)b * G + g) * N * (size_t)L;
            const float *Cr = Cm + ((size_t)b * G + g) * N * (size_t)L;
            float *yr = out + ((size_t)b * D + d) * L;
            REAL h[256];
            for (int n = 0; n < N; ++n) h[n] = 0;
            const REAL bias = dbias ? (REAL)dbias[d] : (REAL)0;
            const REAL dsk = Dskip ? (REAL)Dskip[d] : (REAL)0;
            for (int t = 0; t < L; ++t) {
                REAL dt = (REAL)dr[t] + bias;
                if (delta_softplus) dt = softplus(dt);
                const REAL ut = (REAL)ur[t];
                REAL y = 0;
                for (int n = 0; n < N; ++n) {
                    const REAL a = r_exp(dt * (REAL)A[d * N + n]);
                    h[n] = a * h[n] + dt * (REAL)Br[(size_t)n * L + t] * ut;
                    y += (REAL)Cr[(size_t)n * L + t] * h[n];
                }
                yr[t] = (float)(y + dsk * ut);
            }
            if (last_state)
                for (int n = 0; n < N; ++n) last_state[((size_t)b * D + d) * N + n] = (float)h[n];
        }
    }
}

/*
 * Backward.  Gradient buffers are fp32 and are OVERWRITTEN:
 *   du, ddelta : (Bn, D, L)   dA : (D, N)   dBm, dCm : (Bn, G, N, L)
 *   dD, ddbias : (D) (may be NULL when the matching input is NULL)
 * Parallel over (b, g); the d loop inside a group is serial so dB/dC need no atomics;
 * dA/dD/ddbias are reduced over b through per-b scratch.
 */
void FN(_bwd)(const float *u, const float *delta, const float *A, const float *Bm, const float *Cm,
              const float *Dskip, const float *dbias, int delta_softplus, const float *dout, int Bn,
              int D, int L, int N, int G, float *du, float *ddelta, float *dA, float *dBm, float *dCm,
              float *dD, float *ddbias)
{
    const int dpg = D / G;
    double *pA = (double *)calloc((size_t)Bn * D * N, sizeof(double));
    double *pD = (double *)calloc((size_t)Bn * D, sizeof(double));
    double *pb = (double *)calloc((size_t)Bn * D, sizeof(double));
#pragma omp parallel
    {
        REAL *hs = (REAL *)malloc((size_t)(L + 1) * N * sizeof(REAL)); /* hs[(t+1)*N+n] = h_t */
        REAL *dts = (REAL *)malloc((size_t)L * sizeof(REAL));
        REAL *accB = (REAL *)malloc((size_t)N * L * sizeof(REAL));
        REAL *accC = (REAL *)malloc((size_t)N * L * sizeof(REAL));
#pragma omp for collapse(2) schedule(dynamic, 1)
        for (int b = 0; b < Bn; ++b) {
            for (int g = 0; g < G; ++g) {
                const float *Br = Bm + ((size_t)b * G + g) * N * (size_t)L;
                const float *Cr = Cm + ((size_t)b * G + g) * N * (size_t)L;
                memset(accB, 0, (size_t)N * L * sizeof(REAL));
                memset(accC, 0, (size_t)N * L * sizeof(REAL));
                for (int d = g * dpg; d < (g + 1) * dpg; ++d) {
                    const size_t row = ((size_t)b * D + d) * L;
                    const float *ur = u + row, *dr = delta + row, *gy = dout + row;
                    const REAL bias = dbias ? (REAL)dbias[d] : (REAL)0;
                    const REAL dsk = Dskip ? (REAL)Dskip[d] : (REAL)0;
                    for (int n = 0; n < N; ++n) hs[n] = 0;
                    for (int t = 0; t < L; ++t) {
                        REAL dt = (REAL)dr[t] + bias;
                        if (delta_softplus) dt = softplus(dt);
                        dts[t] = dt;
                        const REAL ut = (REAL)ur[t];
                        for (int n = 0; n < N; ++n) {
                            const REAL a = r_exp(dt * (REAL)A[d * N + n]);
                            hs[(size_t)(t + 1) * N + n] =
                                a * hs[(size_t)t * N + n] + dt * (REAL)Br[(size_t)n * L + t] * ut;
                        }
                    }
                    REAL gst[256];
                    REAL anext[256];
                    for (int n = 0; n < N; ++n) { gst[n] = 0; anext[n] = 0; }
                    double sD = 0, sb = 0;
                    for (int t = L - 1; t >= 0; --t) {
                        const REAL dy = (REAL)gy[t], ut = (REAL)ur[t], dt = dts[t];
                        REAL s1 = 0, s2 = 0;
                        for (int n = 0; n < N; ++n) {
                            const REAL An = (REAL)A[d * N + n];
                            const REAL a = r_exp(dt * An);
                            const REAL Bt = (REAL)Br[(size_t)n * L + t], Ct = (REAL)Cr[(size_t)n * L + t];
                            const REAL gn = Ct * dy + anext[n] * gst[n];
                            const REAL hprev = hs[(size_t)t * N + n];
                            accC[(size_t)n * L + t] += dy * hs[(size_t)(t + 1) * N + n];
                            accB[(size_t)n * L + t] += gn * dt * ut;
                            s1 += gn * Bt;
                            const REAL gha = gn * hprev * a;
                            s2 += gha * An;
                            pA[((size_t)b * D + d) * N + n] += (double)(gha * dt);
                            gst[n] = gn;
                            anext[n] = a;
                        }
                        du[row + t] = (float)(dsk * dy + dt * s1);
                        REAL ddt = s2 + ut * s1;
                        if (delta_softplus) {
                            const REAL x = (REAL)dr[t] + bias;
                            ddt = x > (REAL)20 ? ddt : ddt * sigmoid(x);
                        }
                        ddelta[row + t] = (float)ddt;
                        sD += (double)(dy * ut);
                        sb += (double)ddt;
                    }
                    pD[(size_t)b * D + d] = sD;
                    pb[(size_t)b * D + d] = sb;
                }
                float *dBr = dBm + ((size_t)b * G + g) * N * (size_t)L;
                float *dCr = dCm + ((size_t)b * G + g) * N * (size_t)L;
                for (size_t i = 0; i < (size_t)N * L; ++i) { dBr[i] = (float)accB[i]; dCr[i] = (float)accC[i]; }
            }
        }
        free(hs); free(dts); free(accB); free(accC);
    }
    for (int d = 0; d < D; ++d) {
        double sD = 0, sb = 0;
        for (int b = 0; b < Bn; ++b) { sD += pD[(size_t)b * D + d]; sb += pb[(size_t)b * D + d]; }
        if (dD) dD[d] = (float)sD;
        if (ddbias) ddbias[d] = (float)sb;
        for (int n = 0; n < N; ++n) {
            double s = 0;
            for (int b = 0; b < Bn; ++b) s += pA[((size_t)b * D + d) * N + n];
            dA[d * N + n] = (float)s;
        }
    }
    free(pA); free(pD); free(pb);
}

int FN(_threads)(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
