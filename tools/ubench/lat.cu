// Latency / throughput probes for the fp64 pipe on this GPU (study tool).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dep_chain(double* out, double a, double b, int n, long long* cyc) {
    double x = a + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) x = fma(x, b, a);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
template <int ILP>
__global__ void ilp_chain(double* out, double a, double b, int n, long long* cyc) {
    double x[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) x[j] = a + threadIdx.x + j;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
            for (int j = 0; j < ILP; ++j) x[j] = fma(x[j], b, a);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
    double s = 0;
#pragma unroll
    for (int j = 0; j < ILP; ++j) s += x[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    double* out; long long* cyc; long long h;
    cudaMalloc(&out, 8 * 1024 * 1024); cudaMalloc(&cyc, 8);
    const int n = 1000;
    for (int warps : {1, 2, 4, 8, 16, 32}) {
        dep_chain<<<1, 32 * warps>>>(out, 1.0, 0.999, n, cyc);
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("dependent DFMA chain, %2d warps/SM: %.2f cycles per DFMA (per warp)\n", warps, (double)h / (16.0 * n));
    }
#define RUN(ILP, W) { ilp_chain<ILP><<<1, 32 * W>>>(out, 1.0, 0.999, n, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
    printf("ILP %d, %2d warps/SM: %.2f cycles per DFMA issue slot -> %.1f DFMA lanes/clk/SM\n", ILP, W, (double)h / (8.0 * n * ILP), 32.0 * W * 8.0 * n * ILP / (double)h); }
    RUN(2, 4) RUN(4, 4) RUN(8, 4) RUN(2, 8) RUN(4, 8) RUN(8, 8) RUN(4, 16) RUN(8, 16) RUN(4, 32)
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
