// Micro-benchmark: fp64 dependent-issue latency and per-SM throughput on B200 (one CTA).
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void chain_kernel(double *out, long long *cyc, int n, double a, double b) {
    double x[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x + c;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) x[c] = __fma_rn(x[c], a, b);
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void add_chain_kernel(double *out, long long *cyc, int n, double b) {
    double x = threadIdx.x;
    const long long t0 = clock64();
    for (int i = 0; i < n; ++i) x = __dadd_rn(x, b);
    const long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void lds_chain_kernel(int *out, long long *cyc, int n) {
    __shared__ int idx[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) idx[i] = (i * 17 + 5) & 1023;
    __syncthreads();
    int p = threadIdx.x;
    const long long t0 = clock64();
    for (int i = 0; i < n; ++i) p = idx[p];
    const long long t1 = clock64();
    out[threadIdx.x] = p;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
    double *out;
    long long *cyc, h;
    cudaMalloc(&out, 1 << 20);
    cudaMalloc(&cyc, 1024);
    const int n = 4096;
    for (int warps = 1; warps <= 16; warps *= 2) {
        chain_kernel<1><<<1, 32 * warps>>>(out, cyc, n, 1.0000001, 1e-9);
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("DFMA chains=1 warps=%2d : %.2f cycles per dependent op (per warp)\n", warps, (double)h / n);
    }
    for (int warps = 4; warps <= 16; warps *= 2) {
        chain_kernel<4><<<1, 32 * warps>>>(out, cyc, n, 1.0000001, 1e-9);
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("DFMA chains=4 warps=%2d : %.2f cycles per 4 ops -> %.1f DFMA lanes/clk/SM\n", warps, (double)h / n, 4.0 * 32 * warps * n / h);
        chain_kernel<8><<<1, 32 * warps>>>(out, cyc, n, 1.0000001, 1e-9);
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("DFMA chains=8 warps=%2d : %.2f cycles per 8 ops -> %.1f DFMA lanes/clk/SM\n", warps, (double)h / n, 8.0 * 32 * warps * n / h);
    }
    add_chain_kernel<<<1, 32>>>(out, cyc, n, 1e-9);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DADD dependent latency: %.2f cycles\n", (double)h / n);
    lds_chain_kernel<<<1, 32>>>((int *)out, cyc, n);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("LDS dependent latency: %.2f cycles\n", (double)h / n);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
