// Micro-benchmark: per-SM store rate for the model's output pattern (12 planes, 90-wide rows, one strip per CTA)
//   mode 0: every thread stores consecutive cells of the strip (full rows, land and ocean together)
//   mode 1: ocean cells and land cells stored by separate instructions (compacted lists), as ensemble kernel v3
// usage: store_pattern mask.bin   (8100 bytes, 1 = land)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

template <int KIND> __device__ __forceinline__ void st(double *p, double v) {
    if (KIND == 0) __stcs(p, v);
    else if (KIND == 1) *p = v;
    else if (KIND == 2) __stcg(p, v);
    else __stwt(p, v);
}
template <int KIND>
__global__ void __launch_bounds__(512, 1)
pattern_kernel(double *out, size_t member_stride, int days, int mode, const int *list, int n_ocean, int n_cells, int row0) {
    const int tid = threadIdx.x;
    double *base = out + (size_t)blockIdx.x * member_stride;   // this CTA's "member": 12 planes x days x 8100
    const size_t plane_stride = (size_t)days * 8100;
    for (int d = 0; d < days; ++d) {
        double *slot = base + (size_t)d * 8100 + (size_t)row0 * 90;
        if (mode == 0) {
            for (int c = tid; c < n_cells; c += 512)
#pragma unroll
                for (int v = 0; v < 12; ++v) st<KIND>(slot + (size_t)v * plane_stride + c, (double)v);
        } else {
            for (int i = tid; i < n_ocean; i += 512) {
                const int c = list[i];
#pragma unroll
                for (int v = 0; v < 12; ++v) st<KIND>(slot + (size_t)v * plane_stride + c, (double)v);
            }
            for (int i = n_ocean + tid; i < n_cells; i += 512) {
                const int c = list[i];
#pragma unroll
                for (int v = 0; v < 12; ++v) st<KIND>(slot + (size_t)v * plane_stride + c, (double)v);
            }
        }
    }
}

int main(int argc, char **argv) {
    std::vector<unsigned char> mask(8100, 0);
    if (argc > 1) { FILE *f = fopen(argv[1], "rb"); if (f) { if (fread(mask.data(), 1, 8100, f) != 8100) return 2; fclose(f); } }
    const int row0 = 22, rows = 23, n_cells = rows * 90;
    std::vector<int> list;
    for (int c = 0; c < n_cells; ++c) if (!mask[row0 * 90 + c]) list.push_back(c);
    const int n_ocean = (int)list.size();
    for (int c = 0; c < n_cells; ++c) if (mask[row0 * 90 + c]) list.push_back(c);
    int *dlist;
    cudaMalloc(&dlist, list.size() * 4);
    cudaMemcpy(dlist, list.data(), list.size() * 4, cudaMemcpyHostToDevice);
    const int days = 64;
    const size_t member_stride = (size_t)12 * days * 8100;
    double *buf;
    if (cudaMalloc(&buf, member_stride * 148 * 8) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    printf("strip rows %d..%d: %d cells, %d ocean\n", row0, row0 + rows, n_cells, n_ocean);
    const int ctas[] = {1, 33, 132, 148};
    for (int kind = 0; kind < 4; ++kind)
    for (int mode = 0; mode < 2; ++mode)
        for (int ci = 1; ci < 4; ci += 1) {
            const int n = ctas[ci];
            float ms = 0;
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(e0);
                if (kind == 0) pattern_kernel<0><<<n, 512>>>(buf, member_stride, days, mode, dlist, n_ocean, n_cells, row0);
                if (kind == 1) pattern_kernel<1><<<n, 512>>>(buf, member_stride, days, mode, dlist, n_ocean, n_cells, row0);
                if (kind == 2) pattern_kernel<2><<<n, 512>>>(buf, member_stride, days, mode, dlist, n_ocean, n_cells, row0);
                if (kind == 3) pattern_kernel<3><<<n, 512>>>(buf, member_stride, days, mode, dlist, n_ocean, n_cells, row0);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                cudaEventElapsedTime(&ms, e0, e1);
            }
            const double gb = (double)n * days * n_cells * 12 * 8 / 1e9;
            printf("kind %d (0 cs,1 default,2 cg,3 wt) mode %d ctas %3d: %7.1f GB/s total, %5.1f GB/s per SM, %.2f us per strip-day\n", kind, mode, n, gb / (ms * 1e-3),
                   gb / (ms * 1e-3) / n, ms * 1e3 / days);
        }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
