// Micro-benchmark: how fast can ONE SM push stores to HBM, and how does it add up over SMs?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_bw store_bw.cu && ./store_bw
#include <cstdio>
#include <cuda_runtime.h>

template <int VEC, bool CS>
__global__ void __launch_bounds__(512, 1) store_kernel(double *out, size_t per_cta, int iters) {
    double *base = out + (size_t)blockIdx.x * per_cta;
    const double v = (double)threadIdx.x;
    for (int it = 0; it < iters; ++it) {
        double *p = base + ((size_t)it * blockDim.x * 8 * VEC) % per_cta;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (VEC == 1) {
                double *q = p + (size_t)u * blockDim.x + threadIdx.x;
                if (CS) __stcs(q, v); else *q = v;
            } else {
                double2 *q = reinterpret_cast<double2 *>(p) + (size_t)u * blockDim.x + threadIdx.x;
                if (CS) __stcs(q, make_double2(v, v)); else *q = make_double2(v, v);
            }
        }
    }
}

// 12 separate streams per CTA, 720-byte rows, like the model's output planes
__global__ void __launch_bounds__(512, 1) plane_store_kernel(double *out, size_t plane_stride, int days) {
    const int tid = threadIdx.x;
    for (int d = 0; d < days; ++d) {
        for (int c = tid; c < 2070; c += 512) {
#pragma unroll
            for (int v = 0; v < 12; ++v) __stcs(out + ((size_t)(blockIdx.x * 12 + v)) * plane_stride + (size_t)d * 8100 + c, (double)v);
        }
    }
}

int main() {
    double *buf;
    const size_t bytes = 24ull << 30;
    cudaMalloc(&buf, bytes);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int ctas[] = {1, 8, 33, 66, 132, 148};
    for (int which = 0; which < 4; ++which)
        for (int ci = 0; ci < 6; ++ci) {
            const int n = ctas[ci];
            const size_t per_cta = (bytes / 8 / 148) & ~(size_t)4095;
            const int iters = 4096;
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                if (which == 0) store_kernel<1, false><<<n, 512>>>(buf, per_cta, iters);
                if (which == 1) store_kernel<1, true><<<n, 512>>>(buf, per_cta, iters);
                if (which == 2) store_kernel<2, false><<<n, 512>>>(buf, per_cta, iters / 2);
                if (which == 3) store_kernel<2, true><<<n, 512>>>(buf, per_cta, iters / 2);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
            }
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            const double gb = (double)n * iters * 512 * 8 * 8 / 1e9;
            printf("%-12s ctas %3d  %8.1f GB/s total  %7.1f GB/s per SM\n",
                   which == 0 ? "st.64" : which == 1 ? "st.cs.64" : which == 2 ? "st.128" : "st.cs.128", n, gb / (ms * 1e-3), gb / (ms * 1e-3) / n);
        }
    for (int ci = 2; ci < 6; ++ci) {
        const int n = ctas[ci];
        const int days = 259 * 4;
        const size_t plane_stride = (size_t)days * 8100;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            plane_store_kernel<<<n, 512>>>(buf, plane_stride, days);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
        }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double gb = (double)n * days * 2070 * 12 * 8 / 1e9;
        printf("12-plane rows ctas %3d  %8.1f GB/s total  %7.1f GB/s per SM  (%.2f us per day)\n", n, gb / (ms * 1e-3),
               gb / (ms * 1e-3) / n, ms * 1e3 / days);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
