// Micro-benchmark 2: best achievable way to write 12 planes x (23 rows x 90) doubles per CTA per "day".
//   mode 0: cell-major  (thread owns cells, 12 planes per cell)       -- v1-style full rows
//   mode 2: plane-major (whole strip of plane 0, then plane 1, ...)   -- same bytes, fewer concurrent streams
//   mode 3: TMA bulk stores from shared memory, one 16.5 KB copy per plane, issued by one thread
//   mode 4: TMA bulk stores, 4 x 4 KB-ish chunks per plane, issued by 4 threads
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void bulk_store(double *dst, const double *src_smem, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
                 "r"((unsigned)__cvta_generic_to_shared(src_smem)), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(512, 1)
pattern_kernel(double *out, size_t member_stride, int days, int mode, int n_cells, int row0, int time_major) {
    extern __shared__ __align__(128) double stage[];   // [2][n_cells] for the TMA modes
    const int tid = threadIdx.x;
    // member-major: [member][var][day][8100]; time-major: [var][day][member][8100]
    const size_t nm = gridDim.x;
    double *base = time_major ? out + (size_t)blockIdx.x * 8100 : out + (size_t)blockIdx.x * member_stride;
    const size_t plane_stride = time_major ? (size_t)days * nm * 8100 : (size_t)days * 8100;
    const size_t day_stride = time_major ? nm * 8100 : 8100;
    for (int d = 0; d < days; ++d) {
        double *slot = base + (size_t)d * day_stride + (size_t)row0 * 90;
        if (mode == 0) {
            for (int c = tid; c < n_cells; c += 512)
#pragma unroll
                for (int v = 0; v < 12; ++v) __stcs(slot + (size_t)v * plane_stride + c, (double)v);
        } else if (mode == 2) {
            for (int v = 0; v < 12; ++v)
                for (int c = tid; c < n_cells; c += 512) __stcs(slot + (size_t)v * plane_stride + c, (double)v);
        } else {
            for (int v = 0; v < 12; ++v) {
                double *buf = stage + (size_t)(v & 1) * n_cells;
                if (v >= 2) {   // the copy that used this buffer two planes ago must have finished reading it
                    if (tid < 4) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    __syncthreads();
                }
                for (int c = tid; c < n_cells; c += 512) buf[c] = (double)(v + d);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncthreads();
                if (mode == 3) {
                    if (tid == 0) {
                        bulk_store(slot + (size_t)v * plane_stride, buf, n_cells * 8);
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                } else if (tid < 4) {
                    const int chunk = ((n_cells / 4) + 1) & ~1;
                    const int c0 = tid * chunk, c1 = min(n_cells, c0 + chunk);
                    bulk_store(slot + (size_t)v * plane_stride + c0, buf + c0, (c1 - c0) * 8);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
            if (tid < 4) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncthreads();
        }
    }
    if (mode >= 3 && tid < 4) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
    const int row0 = 22, rows = 23, n_cells = rows * 90;
    const int days = 64;
    const size_t member_stride = (size_t)12 * days * 8100;
    double *buf;
    if (cudaMalloc(&buf, member_stride * 148 * 8) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const size_t smem = (size_t)2 * n_cells * 8;
    cudaFuncSetAttribute(pattern_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int ctas[] = {1, 33, 132, 148};
    const int modes[] = {0, 2, 3, 4};
    for (int tm = 0; tm < 2; ++tm)
    for (int mi = 0; mi < 3; ++mi)
        for (int ci = 2; ci < 4; ++ci) {
            const int n = ctas[ci], mode = modes[mi];
            float ms = 0;
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(e0);
                pattern_kernel<<<n, 512, smem>>>(buf, member_stride, days, mode, n_cells, row0, tm);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                cudaEventElapsedTime(&ms, e0, e1);
            }
            const double gb = (double)n * days * n_cells * 12 * 8 / 1e9;
            printf("%s mode %d ctas %3d: %7.1f GB/s total, %5.1f GB/s per SM, %.2f us per strip-day  [%s]\n", tm ? "time-major  " : "member-major", mode, n,
                   gb / (ms * 1e-3), gb / (ms * 1e-3) / n, ms * 1e3 / days, cudaGetErrorString(cudaGetLastError()));
        }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
