// Microbenchmark: cost of LDS.32/64/128 under the broadcast patterns the scan kernels use (4 warps, one per SMSP).
#include <cstdio>
#include <cuda_runtime.h>
template <int VEC, int PAT>
__global__ void k(float *out, long long *cyc, int iters) {
    __shared__ __align__(16) float sm[8192];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i * 0.001f;
    __syncthreads();
    const int lane = threadIdx.x & 31, r = lane >> 2, q = lane & 3;
    int idx;  // element index (floats), multiple of VEC
    if (PAT == 0) idx = 0;                       // uniform
    else if (PAT == 1) idx = r * 68 * 4;         // by row r: 8 distinct, 4 contiguous lanes share (pk pattern, 16B*68 stride)
    else if (PAT == 2) idx = q * 68 * 4;         // by q: 4 distinct, interleaved lanes (BT pattern)
    else if (PAT == 3) idx = lane * VEC;         // all distinct, contiguous
    else if (PAT == 4) idx = r * 68;             // by row, 4B-row-stride layout (v1 delta pattern)
    else idx = (lane >> 3) * 68 * 4;             // by quarter-warp: 4 distinct, 8 contiguous lanes share
    float acc = 0.f;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const int off = ((it + u) & 15) * VEC;  // vary within the padding-free range, same pattern
            const unsigned a = (unsigned)__cvta_generic_to_shared(&sm[idx + off]);
            if (VEC == 4) { float x, y, z, w; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(a)); acc += x + w; }
            else if (VEC == 2) { float x, y; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(x), "=f"(y) : "r"(a)); acc += x + y; }
            else { float x; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(a)); acc += x; }
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int VEC, int PAT>
void run(const char *name, float *out, long long *cyc) {
    const int iters = 2000;
    k<VEC, PAT><<<1, 128>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    k<VEC, PAT><<<1, 128>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-34s VEC=%d : %.2f cycles per LDS per warp (4 warps concurrently) -> %.2f smem-cycles per LDS\n", name, VEC,
           (double)c / (iters * 16), (double)c / (iters * 16) / 4.0);
}
int main() {
    float *out; long long *cyc;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
    run<4, 0>("uniform", out, cyc); run<4, 1>("by row (4 contiguous lanes share)", out, cyc);
    run<4, 2>("by q (interleaved share)", out, cyc); run<4, 3>("all distinct contiguous", out, cyc);
    run<4, 5>("by quarter-warp", out, cyc);
    run<2, 0>("uniform", out, cyc); run<2, 1>("by row", out, cyc); run<2, 2>("by q", out, cyc); run<2, 3>("all distinct", out, cyc);
    run<1, 0>("uniform", out, cyc); run<1, 4>("by row", out, cyc); run<1, 2>("by q", out, cyc); run<1, 3>("all distinct", out, cyc);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
