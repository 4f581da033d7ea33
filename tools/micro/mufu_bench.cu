// Microbenchmark: issue cost of MUFU.EX2 (ex2.approx.ftz.f32) per warp instruction and SM sub-partition, against FFMA.
// One block, W warps (W = 4: one warp per sub-partition; 8: two).  8 independent chains per thread, so the loop is
// throughput-bound, not latency-bound.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float *out, long long *cyc, int iters) {
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = -0.001f * (threadIdx.x + i + 1);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
            else if (MODE == 1) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(x[i]));
            else asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
        }
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE>
void run(const char *name, int warps, float *out, long long *cyc) {
    const int iters = 4000;
    k<MODE><<<1, 32 * warps>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    k<MODE><<<1, 32 * warps>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double per = (double)c / (iters * 8);          // cycles per warp instruction as seen by one warp
    printf("%-10s %d warps/SM : %.2f cycles per instruction per warp -> %.2f lanes/clk/SM\n", name, warps, per,
           32.0 * warps / per);
}
int main() {
    float *out; long long *cyc;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
    for (int w : {1, 4, 8, 16}) run<0>("MUFU.EX2", w, out, cyc);
    for (int w : {1, 4, 8, 16}) run<2>("MUFU.RCP", w, out, cyc);
    for (int w : {1, 4, 8, 16}) run<1>("FFMA", w, out, cyc);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
