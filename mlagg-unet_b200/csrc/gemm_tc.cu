// gemm_tc.cu -- the projection contractions of the MLAgg block and the MSMM on the 5th-generation tensor cores.
//
// Replaces the cuBLAS GEMMs behind every nn.Linear / 1x1 conv of the named path (reference
// nnUNetTrainer_MLAgg_2D_dt_MS.py:673-674 q / kv, :849-850 in_proj / act_proj, :867 / :902 out_proj, :176-192 Mlp;
// MambaSkip.py:301 in_proj, :345 out_proj, :431 x_proj, :559-577 ConvolutionalGLU fc1 / fc2) and their two backward GEMMs.
//
//   D[M, N] (+)= A[M, K] . B[N, K]^T          bf16 operands, fp32 accumulation in TENSOR MEMORY
//
// One CTA = one 128 x BN output tile (BN <= 256), 6 warps with fixed roles:
//   warp 0     TMA producer: cp.async.bulk.tensor (SASS UTMALDG) of the A / B k-blocks (64 elements of K = one 128-byte
//              swizzle atom) into a ring of shared-memory stages, completion on `full` mbarriers; owns the TMEM allocation
//   warp 1     MMA issuer: ONE elected lane issues tcgen05.mma.cta_group::1.kind::f16 (SASS UTCHMMA), M = 128, N = BN,
//              K = 16 per instruction, operands straight from the swizzled stages through shared-memory descriptors;
//              tcgen05.commit releases a stage to the producer (`empty`) and finally hands the accumulator over (`accf`)
//   warps 2-5  epilogue: tcgen05.ld 32 lanes x 32 columns (SASS LDTM) -> bias / activation / activation-gradient product
//              -> bf16 or fp32 rows to HBM, or fp32 vector reductions (split-K weight gradients)
// Both operand majors are supported through the descriptors (no transposed copies anywhere):
//   K-major   memory [rows][K]  (activations x W^T: forward)          TMA box {64 k, rows}
//   MN-major  memory [K][rows]  (dY . W: data gradient, B operand;   TMA box {64 rows, 64 k} per 64-row slab
//                                dY^T . X: weight gradient, A and B)
// The weight-gradient GEMM contracts over the tokens: gridDim.z CTAs split K, each reduces its partial tile into the
// fp32 gradient with red.global.add.v4.f32.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "gemm_tc.cuh"

namespace mlagg {

constexpr int kBM = 128, kBK = 64;
constexpr int kABytes = kBM * kBK * 2;          // 16 KiB per stage
constexpr int kSlab = 64 * 128;                 // MN-major slab: 64 k-rows x 128 bytes


// ---------------------------------------------------------------- PTX wrappers (tcgen05 / TMA)
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tmap_prefetch(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t *slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor (sm_100 format, cute::UMMA::SmemDescriptor): start address >> 4 in [0,14), leading
// byte offset >> 4 in [16,30), stride byte offset >> 4 in [32,46), version 1 in [46,48), layout SWIZZLE_128B = 2 in
// [61,64).  K-major operand: rows of 128 bytes, 8-row groups SBO = 1024 bytes apart (LBO unused).  MN-major operand:
// k-rows of 128 bytes (64 MN elements), 8-k groups SBO = 1024 bytes apart, 64-element MN slabs LBO bytes apart.
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
           (1ull << 46) | (2ull << 61);
}

// gelu_parts / gelu_f / gelu_grad: common.cuh (Abramowitz-Stegun erf on one exponential shared with the density)
__device__ __forceinline__ float silu_grad(float x) {
    const float s = 1.f / (1.f + __expf(-x));
    return s * (1.f + x * (1.f - s));
}

template <bool kAmn, bool kBmn>
__global__ void __launch_bounds__(192) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                      const __grid_constant__ CUtensorMap tmB, const GemmTcParams p) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    const int BN = p.BN;
    const int nslabB = (BN + 63) / 64;
    const int bBytes = kBmn ? nslabB * kSlab : BN * 128;
    const int stageBytes = kABytes + bBytes;
    unsigned char *tail = smem + (size_t)p.stages * stageBytes;
    uint64_t *full = reinterpret_cast<uint64_t *>(tail);      // [stages]
    uint64_t *empty = full + p.stages;                        // [stages]
    uint64_t *accf = empty + p.stages;                        // [1]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(accf + 1);
    float *bias_s = reinterpret_cast<float *>(tmem_slot + 2); // [BN + 32]
    // reduce mode with `colsum`: a 16 x 64 K-major tile of bf16 ones (2 KiB, 1024-byte aligned).  One more MMA per k-step,
    // D2[128, 16] += A . ones^T, leaves sum_k A[m][k] -- the bias gradient of the layer whose weight gradient this CTA is
    // computing -- in every column of D2, so the column-sum kernel (145 launches per train step in round 1) disappears.
    unsigned char *ones_s = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(bias_s + BN + 32) + 1023) & ~uintptr_t(1023));
    const bool do_colsum = p.colsum != nullptr && blockIdx.y == 0;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * kBM, n0 = blockIdx.y * BN;
    const int kb_total = (p.K + kBK - 1) / kBK;
    const int kb0 = blockIdx.z * p.kblocks_per_split;
    const int kb1 = min(kb_total, kb0 + p.kblocks_per_split);
    const int nkb = kb1 - kb0;
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < BN + (p.colsum != nullptr ? 32 : 0)) tmem_cols <<= 1;
    if (p.colsum != nullptr) {
        for (int i = threadIdx.x; i < 512; i += blockDim.x) reinterpret_cast<uint32_t *>(ones_s)[i] = 0x3F803F80u;
        fence_proxy_async();
    }

    if (warp == 0) {
        if (lane == 0) {
            tmap_prefetch(&tmA);
            tmap_prefetch(&tmB);
            for (int s = 0; s < p.stages; ++s) {
                mbar_init(&full[s], 1);
                mbar_init(&empty[s], 1);
            }
            mbar_init(accf, 1);
            mbar_fence_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, tmem_cols);
    }
    for (int i = threadIdx.x; i < BN + 32; i += blockDim.x)      // + 32: the epilogue walks the columns in chunks of 32
        bias_s[i] = (p.bias != nullptr && n0 + i < p.N) ? p.bias[n0 + i] : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ============================================================ TMA producer
        if (lane == 0) {
            for (int i = 0; i < nkb; ++i) {
                const int s = i % p.stages;
                if (i >= p.stages) mbar_wait(&empty[s], ((i / p.stages) & 1) ^ 1);
                unsigned char *sa = smem + (size_t)s * stageBytes, *sb = sa + kABytes;
                const int k0 = (kb0 + i) * kBK;
                mbar_arrive_expect_tx(&full[s], (uint32_t)stageBytes);
                if (kAmn) {
                    tma_load_2d(sa, &tmA, m0, k0, &full[s]);
                    tma_load_2d(sa + kSlab, &tmA, m0 + 64, k0, &full[s]);
                } else {
                    tma_load_2d(sa, &tmA, k0, m0, &full[s]);
                }
                if (kBmn) {
                    for (int j = 0; j < nslabB; ++j) tma_load_2d(sb + j * kSlab, &tmB, n0 + 64 * j, k0, &full[s]);
                } else {
                    tma_load_2d(sb, &tmB, k0, n0, &full[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ============================================================ MMA issuer
        // instruction descriptor (cute::UMMA::InstrDescriptor): D fp32 [4,6) = 1, A / B bf16 [7,10) = [10,13) = 1, A / B
        // major bits 15 / 16 (1 = MN-major), N >> 3 in [17,23), M >> 4 in [24,29)
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((kAmn ? 1u : 0u) << 15) | ((kBmn ? 1u : 0u) << 16) |
                               ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
        for (int i = 0; i < nkb; ++i) {
            const int s = i % p.stages;
            mbar_wait(&full[s], (i / p.stages) & 1);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t sa = smem_u32(smem + (size_t)s * stageBytes), sb = sa + kABytes;
                const int kleft = p.K - (kb0 + i) * kBK;
                const int ksteps = (min(kleft, kBK) + 15) >> 4;
                for (int k = 0; k < ksteps; ++k) {
                    // K-major: 16 elements of K = 32 bytes inside the 128-byte swizzle row; MN-major: 16 k-rows = 2 atoms
                    const uint64_t ad = kAmn ? smem_desc(sa + k * 2048, kSlab, 1024) : smem_desc(sa + k * 32, 16, 1024);
                    const uint64_t bd = kBmn ? smem_desc(sb + k * 2048, kSlab, 1024) : smem_desc(sb + k * 32, 16, 1024);
                    umma_bf16(tmem_base, ad, bd, idesc, (i | k) != 0 ? 1u : 0u);
                    if (do_colsum) {
                        const uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((kAmn ? 1u : 0u) << 15) | (2u << 17) |
                                                ((uint32_t)(kBM >> 4) << 24);      // N = 16, B K-major
                        umma_bf16(tmem_base + (uint32_t)BN, ad, smem_desc(smem_u32(ones_s), 16, 1024), idesc1,
                                  (i | k) != 0 ? 1u : 0u);
                    }
                }
                umma_commit(&empty[s]);                 // stage free once these MMAs have read it
                if (i == nkb - 1) umma_commit(accf);    // accumulator complete
            }
            __syncwarp();
        }
    } else {
        // ============================================================ epilogue (TMEM lanes 32 * (warp % 4) ...)
        const int quad = warp & 3;
        const int row = m0 + 32 * quad + lane;
        const bool rok = row < p.M;
        if (nkb > 0) {
            mbar_wait(accf, 0);
            tc_fence_after();
        }
        if (do_colsum && nkb > 0) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(32 * quad) << 16) + (uint32_t)BN, v);
            if (rok) atomicAdd(p.colsum + row, __uint_as_float(v[0]));
        }
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t v[32];
            if (nkb > 0) {
                tmem_ld32(tmem_base + ((uint32_t)(32 * quad) << 16) + (uint32_t)c0, v);
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0u;
            }
            const int nbase = n0 + c0;
            if (p.reduce) {
                if (rok) {
                    float *o = reinterpret_cast<float *>(p.out) + (size_t)row * p.ldo + nbase;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        if (nbase + j < p.N)     // N is a multiple of 4 (checked on the host)
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + j),
                                         "f"(__uint_as_float(v[j])), "f"(__uint_as_float(v[j + 1])),
                                         "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3]))
                                         : "memory");
                    }
                }
                continue;
            }
            float x[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]) + bias_s[c0 + j];
            if (!rok) continue;
            if (p.pre != nullptr) {
                __nv_bfloat16 *o = reinterpret_cast<__nv_bfloat16 *>(p.pre) + (size_t)row * p.ldpre + nbase;
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    if (nbase + j < p.N) {
                        uint4 w;
                        __nv_bfloat162 t;
                        t = __floats2bfloat162_rn(x[j], x[j + 1]);     w.x = *reinterpret_cast<uint32_t *>(&t);
                        t = __floats2bfloat162_rn(x[j + 2], x[j + 3]); w.y = *reinterpret_cast<uint32_t *>(&t);
                        t = __floats2bfloat162_rn(x[j + 4], x[j + 5]); w.z = *reinterpret_cast<uint32_t *>(&t);
                        t = __floats2bfloat162_rn(x[j + 6], x[j + 7]); w.w = *reinterpret_cast<uint32_t *>(&t);
                        *reinterpret_cast<uint4 *>(o + j) = w;
                    }
                }
            }
            if (p.aux != nullptr) {      // gradient through the activation of the layer below: x *= act'(aux)
                const __nv_bfloat16 *a = reinterpret_cast<const __nv_bfloat16 *>(p.aux) + (size_t)row * p.ldaux + nbase;
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    if (nbase + j < p.N) {
                        const uint4 w = *reinterpret_cast<const uint4 *>(a + j);
                        const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&ww[e]));
                            if (p.act == 1) {
                                x[j + 2 * e] *= gelu_grad(f.x);
                                x[j + 2 * e + 1] *= gelu_grad(f.y);
                            } else if (p.act == 2) {
                                x[j + 2 * e] *= silu_grad(f.x);
                                x[j + 2 * e + 1] *= silu_grad(f.y);
                            }
                        }
                    }
                }
            } else if (p.act == 1) {
#pragma unroll
                for (int j = 0; j < 32; ++j) x[j] = gelu_f(x[j]);
            } else if (p.act == 2) {
#pragma unroll
                for (int j = 0; j < 32; ++j) x[j] = silu_f(x[j]);
            }
            if (p.out_f32) {
                float *o = reinterpret_cast<float *>(p.out) + (size_t)row * p.ldo + nbase;
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    if (nbase + j < p.N) *reinterpret_cast<float4 *>(o + j) = make_float4(x[j], x[j + 1], x[j + 2], x[j + 3]);
            } else {
                __nv_bfloat16 *o = reinterpret_cast<__nv_bfloat16 *>(p.out) + (size_t)row * p.ldo + nbase;
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    if (nbase + j < p.N) {
                        uint4 w;
                        __nv_bfloat162 t;
                        t = __floats2bfloat162_rn(x[j], x[j + 1]);     w.x = *reinterpret_cast<uint32_t *>(&t);
                        t = __floats2bfloat162_rn(x[j + 2], x[j + 3]); w.y = *reinterpret_cast<uint32_t *>(&t);
                        t = __floats2bfloat162_rn(x[j + 4], x[j + 5]); w.z = *reinterpret_cast<uint32_t *>(&t);
                        t = __floats2bfloat162_rn(x[j + 6], x[j + 7]); w.w = *reinterpret_cast<uint32_t *>(&t);
                        *reinterpret_cast<uint4 *>(o + j) = w;
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}

// ---------------------------------------------------------------- persistent store-mode kernel (forward, data gradient)
// One CTA per SM walks the output tiles of ONE column block (n fixed, m strided) when the whole B operand of that block
// fits in shared memory: W is then loaded ONCE per CTA ("resident") and only activation tiles stream through the TMA ring;
// otherwise (m, n) tiles are walked in row-major order and B streams with A.  Two TMEM accumulators alternate, so the
// epilogue of tile i (tcgen05.ld -> bias / activation -> bf16 -> 128-byte-swizzled staging slab -> TMA store, SASS
// UTMASTG) overlaps the loads and MMAs of tile i + 1.  A is K-major (activations / output gradients); B is K-major
// (forward: W[N][K]) or MN-major (data gradient: W[K][N] as stored).
struct GemmStoreCfg {
    int nt, mt, ctas_per_n, resident, bBytes, kb_total, stages;
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, const void *src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&t);
}

template <bool kBmn, int kAct, bool kPre, bool kAux>
__global__ void __launch_bounds__(320, 1) gemm_tc_store_kernel(const __grid_constant__ CUtensorMap tmA,
                                                               const __grid_constant__ CUtensorMap tmB,
                                                               const __grid_constant__ CUtensorMap tmO,
                                                               const __grid_constant__ CUtensorMap tmP,
                                                               const GemmTcParams p, const GemmStoreCfg c) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    const int BN = p.BN;
    const int nslabB = (BN + 63) / 64;
    const int stageBytes = kABytes + (c.resident ? 0 : c.bBytes);
    unsigned char *ring = smem;
    unsigned char *resB = ring + (size_t)c.stages * stageBytes;                       // [kb_total][bBytes] when resident
    unsigned char *stag = resB + (c.resident ? (size_t)c.kb_total * c.bBytes : 0);    // [2][128 rows][128 B]
    unsigned char *tail = stag + 4 * kABytes;                                         // 8 warps x 2 x (32 rows x 128 B)
    uint64_t *full = reinterpret_cast<uint64_t *>(tail);   // [stages]
    uint64_t *empty = full + c.stages;                     // [stages]
    uint64_t *accf = empty + c.stages;                     // [2] accumulator complete
    uint64_t *acce = accf + 2;                             // [2] accumulator drained
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acce + 2);
    float *bias_s = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(tmem_slot + 2) + 15) & ~uintptr_t(15));   // [BN + 64]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // tile walk: resident -> n fixed per CTA, m = mfirst + i * mstep; streaming -> t = cta + i * grid, (m, n) = (t / nt, t % nt)
    const int nfix = c.resident ? (int)(blockIdx.x % c.nt) : 0;
    const int mfirst = c.resident ? (int)(blockIdx.x / c.nt) : 0;
    const int ntiles = c.resident ? (c.mt > mfirst ? (c.mt - mfirst + c.ctas_per_n - 1) / c.ctas_per_n : 0)
                                  : ((c.mt * c.nt > (int)blockIdx.x) ? (c.mt * c.nt - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0);
    auto tile_mn = [&](int i, int &m0, int &n0) {
        if (c.resident) {
            m0 = (mfirst + i * c.ctas_per_n) * kBM;
            n0 = nfix * BN;
        } else {
            const int t = blockIdx.x + i * gridDim.x;
            m0 = (t / c.nt) * kBM;
            n0 = (t % c.nt) * BN;
        }
    };
    uint32_t acc_cols = 32;
    while ((int)acc_cols < BN) acc_cols <<= 1;
    const uint32_t tmem_cols = 2 * acc_cols;

    if (warp == 0) {
        if (lane == 0) {
            tmap_prefetch(&tmA);
            tmap_prefetch(&tmB);
            tmap_prefetch(&tmO);
            for (int s = 0; s < c.stages; ++s) {
                mbar_init(&full[s], 1);
                mbar_init(&empty[s], 1);
            }
            for (int b = 0; b < 2; ++b) {
                mbar_init(&accf[b], 1);
                mbar_init(&acce[b], 8);
            }
            mbar_fence_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, tmem_cols);
    }
    if (c.resident) {
        for (int i = threadIdx.x; i < BN + 64; i += blockDim.x)
            bias_s[i] = (p.bias != nullptr && nfix * BN + i < p.N) ? p.bias[nfix * BN + i] : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ============================================================ TMA producer
        if (lane == 0) {
            int kc = 0;
            for (int i = 0; i < ntiles; ++i) {
                int m0, n0;
                tile_mn(i, m0, n0);
                const bool loadB = !c.resident || i == 0;
                for (int kb = 0; kb < c.kb_total; ++kb, ++kc) {
                    const int s = kc % c.stages;
                    if (kc >= c.stages) mbar_wait(&empty[s], ((kc / c.stages) & 1) ^ 1);
                    unsigned char *sa = ring + (size_t)s * stageBytes;
                    unsigned char *sb = c.resident ? resB + (size_t)kb * c.bBytes : sa + kABytes;
                    mbar_arrive_expect_tx(&full[s], (uint32_t)(kABytes + (loadB ? c.bBytes : 0)));
                    tma_load_2d(sa, &tmA, kb * kBK, m0, &full[s]);
                    if (loadB) {
                        if (kBmn) {
                            for (int j = 0; j < nslabB; ++j) tma_load_2d(sb + j * kSlab, &tmB, n0 + 64 * j, kb * kBK, &full[s]);
                        } else {
                            tma_load_2d(sb, &tmB, kb * kBK, n0, &full[s]);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ============================================================ MMA issuer
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((kBmn ? 1u : 0u) << 16) | ((uint32_t)(BN >> 3) << 17) |
                               ((uint32_t)(kBM >> 4) << 24);
        int kc = 0;
        for (int i = 0; i < ntiles; ++i) {
            const int buf = i & 1;
            if (i >= 2) mbar_wait(&acce[buf], ((i >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t dtm = tmem_base + buf * acc_cols;
            for (int kb = 0; kb < c.kb_total; ++kb, ++kc) {
                const int s = kc % c.stages;
                mbar_wait(&full[s], (kc / c.stages) & 1);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t sa = smem_u32(ring + (size_t)s * stageBytes);
                    const uint32_t sb = c.resident ? smem_u32(resB + (size_t)kb * c.bBytes) : sa + kABytes;
                    const int ksteps = (min(p.K - kb * kBK, kBK) + 15) >> 4;
                    for (int k = 0; k < ksteps; ++k) {
                        const uint64_t ad = smem_desc(sa + k * 32, 16, 1024);
                        const uint64_t bd = kBmn ? smem_desc(sb + k * 2048, kSlab, 1024) : smem_desc(sb + k * 32, 16, 1024);
                        umma_bf16(dtm, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(&empty[s]);
                    if (kb == c.kb_total - 1) umma_commit(&accf[buf]);
                }
                __syncwarp();
            }
        }
    } else {
        // ============================================================ epilogue: 8 warps, no block-level barriers
        // warp w reads TMEM lanes 32 * (w % 4) ... (its 32 rows of the tile); warps 2-5 take the even 64-column slabs,
        // warps 6-9 the odd ones.  Each warp owns two 4 KiB staging buffers and issues its OWN TMA store (box 64 x 32),
        // so slabs only need __syncwarp: wait until the store that used this buffer two slabs ago has been read, write
        // the rows (16-byte chunk ch of row r at ch ^ (r & 7): the store's 128-byte swizzle, conflict-free across the
        // 32 rows of a warp), fence to the async proxy, lane 0 stores.  The epilogue math is selected at COMPILE time
        // (kAct / kPre / kAux): with run-time flags inside the 64-element unrolled body the first version executed
        // ~660 instructions per warp and tile and paced the whole kernel (ncu: profiles/gemm_store_r02_ncu.txt).
        const int quad = warp & 3;
        const int grp = (warp - 2) >> 2;
        const int rloc = 32 * quad + lane;                     // row inside the tile = TMEM lane
        unsigned char *mystag = stag + (size_t)(warp - 2) * 2 * 4096;
        int slabc = 0;
        for (int i = 0; i < ntiles; ++i) {
            int m0, n0;
            tile_mn(i, m0, n0);
            const int buf = i & 1;
            if (!c.resident) {                                  // bias of this tile's column block
                named_bar_sync(3, 256);
                for (int j = threadIdx.x - 64; j < BN + 64; j += 256)
                    bias_s[j] = (p.bias != nullptr && n0 + j < p.N) ? p.bias[n0 + j] : 0.f;
                named_bar_sync(3, 256);
            }
            mbar_wait(&accf[buf], (i >> 1) & 1);
            tc_fence_after();
            const uint32_t trow = tmem_base + buf * acc_cols + ((uint32_t)(32 * quad) << 16);
            const bool rok = m0 + rloc < p.M;
#pragma unroll
            for (int pass = 0; pass < (kPre ? 2 : 1); ++pass) {
                const bool to_pre = kPre && pass == 0;          // pass 0 of 2: the pre-activation goes to `pre`
                for (int c0 = 64 * grp; c0 < BN; c0 += 128, ++slabc) {
                    unsigned char *sbuf = mystag + (slabc & 1) * 4096;
                    unsigned char *sg = sbuf + lane * 128;
                    const bool two = c0 + 32 < BN;              // (uniform) second 32-column half inside the accumulator
                    uint32_t v[2][32];
                    tmem_ld32_nowait(trow + (uint32_t)c0, v[0]);
                    if (two) tmem_ld32_nowait(trow + (uint32_t)(c0 + 32), v[1]);
                    if (lane == 0) bulk_wait_read<1>();
                    tmem_ld_wait();
                    __syncwarp();
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        if (hf == 1 && !two) break;
                        const int cc = c0 + 32 * hf;
                        float x[32];
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 bb = *reinterpret_cast<const float4 *>(bias_s + cc + j);
                            x[j] = __uint_as_float(v[hf][j]) + bb.x;
                            x[j + 1] = __uint_as_float(v[hf][j + 1]) + bb.y;
                            x[j + 2] = __uint_as_float(v[hf][j + 2]) + bb.z;
                            x[j + 3] = __uint_as_float(v[hf][j + 3]) + bb.w;
                        }
                        if (!to_pre) {
                            if (kAux) {
                                if (rok) {
                                    const __nv_bfloat16 *a = reinterpret_cast<const __nv_bfloat16 *>(p.aux) +
                                                             (size_t)(m0 + rloc) * p.ldaux + n0 + cc;
#pragma unroll
                                    for (int j = 0; j < 32; j += 8) {
                                        if (n0 + cc + j < p.N) {
                                            const uint4 w = *reinterpret_cast<const uint4 *>(a + j);
                                            const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                                            for (int e = 0; e < 4; ++e) {
                                                const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&ww[e]));
                                                x[j + 2 * e] *= kAct == 1 ? gelu_grad(f.x) : silu_grad(f.x);
                                                x[j + 2 * e + 1] *= kAct == 1 ? gelu_grad(f.y) : silu_grad(f.y);
                                            }
                                        }
                                    }
                                }
                            } else if (kAct == 1) {
#pragma unroll
                                for (int j = 0; j < 32; ++j) x[j] = gelu_f(x[j]);
                            } else if (kAct == 2) {
#pragma unroll
                                for (int j = 0; j < 32; ++j) x[j] = silu_f(x[j]);
                            }
                        }
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            const int ch = (4 * hf + (j >> 3)) ^ (lane & 7);
                            *reinterpret_cast<uint4 *>(sg + ch * 16) = make_uint4(pack2(x[j], x[j + 1]), pack2(x[j + 2], x[j + 3]),
                                                                                  pack2(x[j + 4], x[j + 5]), pack2(x[j + 6], x[j + 7]));
                        }
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(to_pre ? &tmP : &tmO, sbuf, n0 + c0, m0 + 32 * quad);
                        bulk_commit();
                    }
                }
            }
            // every TMEM read of this tile by this warp is complete (tcgen05.wait::ld): hand the accumulator back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acce[buf]);
        }
        if (lane == 0) bulk_wait_read<0>();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// bf16 matrix [outer][inner] with row stride ld (elements): box {box_inner, box_outer}, 128-byte swizzle, zero fill
static bool make_map(CUtensorMap *m, const void *base, long long inner, long long outer, long long ld, int box_inner,
                     int box_outer) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
    cuuint32_t es[2] = {1, 1};
    const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS && getenv("MLAGG_DEBUG"))
        fprintf(stderr, "mlagg: cuTensorMapEncodeTiled -> %d (base %p inner %lld outer %lld ld %lld box %d x %d)\n", (int)r,
                base, inner, outer, ld, box_inner, box_outer);
    return r == CUDA_SUCCESS;
}

static int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

static bool make_map_out(CUtensorMap *m, const void *base, long long inner, long long outer, long long ld) {
    return make_map(m, base, inner, outer, ld, 64, 32);      // one epilogue warp's store: 64 columns x its 32 rows
}

// persistent store-mode launch (bf16 output); returns cudaErrorInvalidConfiguration when the shape does not fit its plan
static cudaError_t gemm_tc_store_launch(const void *A, long long lda, const void *B, long long ldb, int b_mn, GemmTcParams p,
                                        cudaStream_t st) {
    GemmStoreCfg c{};
    int BN = p.N <= 256 ? (p.N + 15) / 16 * 16 : 256;
    if (p.N > 256) {
        const int nt = (p.N + 255) / 256;
        BN = ((p.N + nt - 1) / nt + 63) / 64 * 64;       // several column blocks: whole 64-column store slabs
    }
    p.BN = BN;
    c.nt = (p.N + BN - 1) / BN;
    c.mt = (p.M + kBM - 1) / kBM;
    c.kb_total = (p.K + kBK - 1) / kBK;
    c.bBytes = b_mn ? ((BN + 63) / 64) * kSlab : BN * 128;
    const int sms = sm_count();
    const long long resBytes = (long long)c.kb_total * c.bBytes;
    c.resident = resBytes <= 96 * 1024 && c.nt <= sms;
    const int fixed = 4 * kABytes + 2048 + (BN + 64) * 4 + 1024;
    const int budget = 227 * 1024 - fixed - (c.resident ? (int)resBytes : 0);
    const int stageBytes = kABytes + (c.resident ? 0 : c.bBytes);
    c.stages = min(8, budget / stageBytes);
    if (c.stages < 2) return cudaErrorInvalidConfiguration;
    int grid;
    if (c.resident) {
        c.ctas_per_n = max(1, min(sms / c.nt, c.mt));
        grid = c.ctas_per_n * c.nt;
    } else {
        c.ctas_per_n = 0;
        grid = min(sms, c.mt * c.nt);
    }
    const size_t smem = (size_t)c.stages * stageBytes + (c.resident ? (size_t)resBytes : 0) + fixed;
    CUtensorMap tmA, tmB, tmO, tmP;
    auto encode = [&]() {
        bool ok = make_map(&tmA, A, p.K, p.M, lda, 64, kBM);
        ok = ok && (b_mn ? make_map(&tmB, B, p.N, p.K, ldb, 64, 64) : make_map(&tmB, B, p.K, p.N, ldb, 64, BN));
        ok = ok && make_map_out(&tmO, p.out, p.N, p.M, p.ldo);
        ok = ok && (p.pre ? make_map_out(&tmP, p.pre, p.N, p.M, p.ldpre) : make_map_out(&tmP, p.out, p.N, p.M, p.ldo));
        return ok;
    };
    if (!encode()) {
        cudaFree(nullptr);
        if (!encode()) return cudaErrorNotSupported;
    }
    using KernT = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const GemmTcParams,
                           const GemmStoreCfg);
    KernT kern = nullptr;
    const int act = p.act;
#define MLAGG_PICK(BMN, ACT, PRE, AUX) \
    if (b_mn == BMN && act == ACT && (p.pre != nullptr) == PRE && (p.aux != nullptr) == AUX) kern = gemm_tc_store_kernel<BMN, ACT, PRE, AUX>;
    MLAGG_PICK(false, 0, false, false) MLAGG_PICK(true, 0, false, false)
    MLAGG_PICK(false, 1, false, false) MLAGG_PICK(false, 2, false, false)
    MLAGG_PICK(false, 1, true, false) MLAGG_PICK(false, 2, true, false)
    MLAGG_PICK(true, 1, false, true) MLAGG_PICK(true, 2, false, true)
    MLAGG_PICK(false, 0, true, false)
#undef MLAGG_PICK
    if (kern == nullptr) return cudaErrorInvalidConfiguration;       // combination not instantiated: generic kernel
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, 320, smem, st>>>(tmA, tmB, tmO, tmP, p, c);
    return cudaGetLastError();
}

// a_mn / b_mn: 0 = memory [rows][K] (K-major), 1 = memory [K][rows] (MN-major).  Returns cudaErrorNotSupported when the
// driver entry point for tensor maps is unavailable.
cudaError_t gemm_tc_dispatch(const void *A, long long lda, int a_mn, const void *B, long long ldb, int b_mn, GemmTcParams p,
                             cudaStream_t st) {
    if (!p.reduce && !p.out_f32 && !a_mn && !getenv("MLAGG_GEMM_V1")) {
        cudaError_t e = gemm_tc_store_launch(A, lda, B, ldb, b_mn, p, st);
        if (e != cudaErrorInvalidConfiguration) return e;
    }
    int BN = p.N <= 256 ? (p.N + 15) / 16 * 16 : 256;
    if (p.N > 256) {   // balance the column tiles: 384 -> 2 x 192, 768 -> 3 x 256, 1536 -> 6 x 256
        const int nt = (p.N + 255) / 256;
        BN = ((p.N + nt - 1) / nt + 15) / 16 * 16;
    }
    p.BN = BN;
    const int nslabB = (BN + 63) / 64;
    const int bBytes = b_mn ? nslabB * kSlab : BN * 128;
    const int stageBytes = kABytes + bBytes;
    const int kb_total = (p.K + kBK - 1) / kBK;
    const int mt = (p.M + kBM - 1) / kBM, nt = (p.N + BN - 1) / BN;
    int splits = 1;
    if (p.reduce) {
        // about one CTA per SM, at least 8 k-blocks each; every split pays an fp32 atomic reduction of its whole tile, so
        // the total number of reduced elements (splits x M x N) is capped as well -- at 768 x 1536 outputs five splits
        // spent more time in red.global.add than in the MMAs
        const long long cap = max(1LL, 3000000LL / ((long long)p.M * p.N));
        splits = (int)max(1LL, min((long long)min((sm_count() + mt * nt - 1) / (mt * nt), (kb_total + 7) / 8), cap));
    }
    p.kblocks_per_split = (kb_total + splits - 1) / splits;
    splits = (kb_total + p.kblocks_per_split - 1) / p.kblocks_per_split;
    const int budget = p.reduce ? 200 * 1024 : 100 * 1024;     // store mode: two CTAs per SM overlap epilogue and loads
    p.stages = max(2, min(min(8, budget / stageBytes), max(2, p.kblocks_per_split)));
    const size_t smem = (size_t)p.stages * stageBytes + (2 * p.stages + 1) * 8 + 8 + (size_t)(BN + 32) * 4 + 1024 +
                        (p.colsum ? 3072 : 0);

    CUtensorMap tmA, tmB;
    auto encode = [&]() {
        bool ok = a_mn ? make_map(&tmA, A, p.M, p.K, lda, 64, 64) : make_map(&tmA, A, p.K, p.M, lda, 64, kBM);
        return ok && (b_mn ? make_map(&tmB, B, p.N, p.K, ldb, 64, 64) : make_map(&tmB, B, p.K, p.N, ldb, 64, BN));
    };
    if (!encode()) {
        // cuTensorMapEncodeTiled is a driver-API call: on a thread that has not touched the runtime yet (autograd's
        // backward thread) no context is current -- bind the primary context of the current device and try once more
        cudaFree(nullptr);
        if (!encode()) return cudaErrorNotSupported;
    }

    auto kern = a_mn ? (b_mn ? gemm_tc_kernel<true, true> : gemm_tc_kernel<true, false>)
                     : (b_mn ? gemm_tc_kernel<false, true> : gemm_tc_kernel<false, false>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid(mt, nt, splits);
    kern<<<grid, 192, smem, st>>>(tmA, tmB, p);
    return cudaGetLastError();
}

}  // namespace mlagg
