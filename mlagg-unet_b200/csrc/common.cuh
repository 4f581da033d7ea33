// common.cuh -- PTX helpers shared by the sm_100a kernels (mbarrier, bulk async copies, fast math).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mlagg {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug traps (visible as a launch failure) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins) {
        if (spins > (1u << 26)) __trap();
    }
}

// ---------------------------------------------------------------- bulk async copies (TMA, 1-D)
// global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).  16-byte aligned, size % 16 == 0.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global, tracked by the issuing thread's bulk async-group.
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
                 "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// make generic-proxy writes to shared memory visible to the async proxy (before bulk_s2g)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------- named barriers
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------- fast math (MUFU)
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// softplus(x) = x > 20 ? x : log1p(exp(x))  (torch semantics), plus d softplus/dx = sigmoid(x) (1 above 20).
// log1p(e) is a 4-term series for e < 1/32 (rel. error < e^4/5 < 2e-7) and ln(1+e) through MUFU.LG2 above.
__device__ __forceinline__ float softplus_fast(float x, float *dsp = nullptr) {
    // branch-free: both log1p forms are evaluated and selected, so independent elements interleave freely
    const float e = ex2_approx(fminf(x, 20.f) * kLog2e);
    const float series = e * (1.f - e * (0.5f - e * (0.33333334f - e * 0.25f)));
    const float viaLog = lg2_approx(1.f + e) * kLn2;
    const float sp = e < 0.03125f ? series : viaLog;
    if (dsp) *dsp = x > 20.f ? 1.f : e * rcp_approx(1.f + e);
    return x > 20.f ? x : sp;
}

__device__ __forceinline__ float silu_f(float x) { return x * rcp_approx(1.f + ex2_approx(-x * kLog2e)); }

// exact-GELU pieces (GEMM epilogues, pooled-branch GELU).  erf through Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, far inside the bf16
// output rounding) on ONE exponential shared with the density: libm's erff + expf cost ~40 instructions per element, which
// made the 8 epilogue warps -- not HBM -- the limiter of the fused fc2-data-gradient GEMM (87 us at stage 0).
__device__ __forceinline__ void gelu_parts(float x, float &cdf, float &pdf_x) {
    const float e = ex2_approx(-0.72134752044f * x * x);              // exp(-x^2 / 2)
    const float z = fabsf(x) * 0.70710678118654752f;
    const float t = rcp_approx(fmaf(0.3275911f, z, 1.f));
    const float poly = t * fmaf(t, fmaf(t, fmaf(t, fmaf(t, 1.061405429f, -1.453152027f), 1.421413741f), -0.284496736f), 0.254829592f);
    const float erfa = fmaf(-poly, e, 1.f);                            // erf(|x| / sqrt 2)
    cdf = 0.5f * (1.f + copysignf(erfa, x));
    pdf_x = x * e * 0.3989422804014327f;
}
__device__ __forceinline__ float gelu_f(float x) {
    float cdf, px;
    gelu_parts(x, cdf, px);
    return x * cdf;
}
__device__ __forceinline__ float gelu_grad(float x) {
    float cdf, px;
    gelu_parts(x, cdf, px);
    return cdf + px;
}

}  // namespace mlagg
