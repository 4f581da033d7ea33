// gemm_tc.cuh -- parameter block of the tensor-core GEMM kernels (csrc/gemm_tc.cu), shared with the C ABI (capi.cu).
#pragma once
#include <cuda_runtime.h>

namespace mlagg {

struct GemmTcParams {
    void *out;                  // [M][ldo]  bf16 / fp32 (store) or fp32 (reduce)
    void *pre;                  // optional second output: the pre-activation, bf16 [M][ldpre]
    const void *aux;            // optional bf16 [M][ldaux]: epilogue multiplies by act'(aux)
    const float *bias;          // optional [N]
    float *colsum;              // reduce mode, optional [M]: += sum over K of A[m][k] (the bias gradient), fp32 atomics
    long long ldo, ldpre, ldaux;
    int M, N, K;
    int BN;                     // UMMA N (multiple of 16, <= 256)
    int stages;
    int act;                    // 0 none, 1 GELU (erf), 2 SiLU
    int out_f32;                // store mode: 1 = fp32 output
    int reduce;                 // 1 = split-K: fp32 red.add into out
    int kblocks_per_split;
};

// a_mn / b_mn: 0 = memory [rows][K] (K-major), 1 = memory [K][rows] (MN-major)
cudaError_t gemm_tc_dispatch(const void *A, long long lda, int a_mn, const void *B, long long ldb, int b_mn, GemmTcParams p,
                             cudaStream_t st);

}  // namespace mlagg
