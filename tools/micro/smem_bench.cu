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
    else if (PAT == 5) idx = (lane >> 3) * 68 * 4;  // by quarter-warp: 4 distinct, 8 contiguous lanes share
    else if (PAT == 6) idx = (lane >> 1) * VEC;  // 16 distinct contiguous, lane pairs share
    else if (PAT == 7) idx = (lane & 3) * VEC;   // 4 distinct contiguous, interleaved share
    else if (PAT == 8) idx = (lane >> 2) * VEC;  // 8 distinct contiguous, 4 contiguous lanes share
    else if (PAT == 9) idx = (lane & 7) * VEC;   // 8 distinct contiguous; every quarter-warp reads the same 8
    else if (PAT == 10) idx = (lane >> 3) * VEC; // 4 distinct contiguous, 8 contiguous lanes share
    else if (PAT == 11) idx = (lane >> 4) * VEC; // 2 distinct, half-warps share
    else idx = (lane & 15) * VEC;                // 16 distinct contiguous; both half-warps read the same 16
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
    run<4, 6>("16 distinct contig, pairs share", out, cyc); run<4, 7>("4 distinct contig, interleaved", out, cyc);
    run<4, 8>("8 distinct contig, 4 lanes share", out, cyc); run<4, 9>("8 distinct contig, per quarter", out, cyc);
    run<4, 10>("4 distinct contig, 8 lanes share", out, cyc); run<4, 11>("2 distinct, half-warps", out, cyc);
    run<4, 12>("16 distinct contig, per half", out, cyc);
    run<2, 6>("16 distinct contig, pairs share", out, cyc); run<2, 9>("8 distinct contig, per quarter", out, cyc);
    run<2, 10>("4 distinct contig, 8 lanes share", out, cyc); run<2, 12>("16 distinct contig, per half", out, cyc);
    run<2, 8>("8 distinct contig, 4 lanes share", out, cyc);
    run<2, 0>("uniform", out, cyc); run<2, 1>("by row", out, cyc); run<2, 2>("by q", out, cyc); run<2, 3>("all distinct", out, cyc);
    run<1, 0>("uniform", out, cyc); run<1, 4>("by row", out, cyc); run<1, 2>("by q", out, cyc); run<1, 3>("all distinct", out, cyc);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
