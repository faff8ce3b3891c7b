// Micro-benchmark 3: the general day kernel's output pattern -- every CTA writes one TW x TH tile of each of 12 planes
// of an (ny x nx) grid with odd nx (rows are only 8-byte aligned) -- with
//   mode 0: one 8-byte store per thread per plane (what day_step_kernel does)
//   mode 1: tile staged in shared memory, one cp.async.bulk per tile row per plane (16-byte aligned part; the odd
//           first/last element of a misaligned row goes out as a scalar store)
// usage: tile_store [nx] ; prints GB/s for 32x16 and 64x8 tiles
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void bulk_store(double *dst, const double *src_smem, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
                 "r"((unsigned)__cvta_generic_to_shared(src_smem)), "r"(bytes) : "memory");
}

template <int TW, int TH>
__global__ void __launch_bounds__(512, 2) tile_kernel(double *out, int ny, int nx, int mode) {
    extern __shared__ __align__(16) double stage_raw[];
    double (*stage)[TH][TW + 2] = reinterpret_cast<double (*)[TH][TW + 2]>(stage_raw);
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const size_t plane = (size_t)ny * nx;
    const int tx = tid % TW, ty = tid / TW;           // 512 threads = TW*TH cells
    const int gx = x0 + tx, gy = y0 + ty;
    if (mode == 2) {      // sheared tile: every row's segment shifted left by 0..3 cells onto a 32-byte boundary (of even planes)
        const int d = (int)(((size_t)gy * nx) & 3);
        const int sx = gx - d;
        if (sx >= 0 && sx < nx && gy < ny)
#pragma unroll
            for (int v = 0; v < 12; ++v) out[v * plane + (size_t)gy * nx + sx] = (double)(v + tid);
        return;
    }
    if (mode == 0) {
        if (gx < nx && gy < ny)
#pragma unroll
            for (int v = 0; v < 12; ++v) out[v * plane + (size_t)gy * nx + gx] = (double)(v + tid);
        return;
    }
    // stage with the row's parity so that the 16-byte aligned global part is 16-byte aligned in shared memory too
    if (gy < ny) {
#pragma unroll
        for (int v = 0; v < 12; ++v) stage[v][ty][tx + (int)((v * plane + (size_t)gy * nx + x0) & 1)] = (double)(v + tid);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const int w = min(TW, nx - x0);
    for (int op = tid; op < 12 * TH; op += 512) {     // one (plane, row) per thread
        const int v = op / TH, r = op % TH;
        const int y = y0 + r;
        if (y >= ny) continue;
        const size_t e = (size_t)y * nx + x0;
        const int par = (int)((v * plane + e) & 1);
        double *dst = out + v * plane + e;
        const double *src = &stage[v][r][par];
        int first = 0, count = w;
        if (par) { dst[0] = src[0]; first = 1; count -= 1; }
        if (count & 1) { dst[first + count - 1] = src[first + count - 1]; count -= 1; }
        if (count > 0) bulk_store(dst + first, src + first, count * 8);
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

template <int TW, int TH>
void run(double *buf, int ny, int nx) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    dim3 grid((nx + TW - 1) / TW, (ny + TH - 1) / TH);
    cudaFuncSetAttribute(tile_kernel<TW, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 12 * TH * (TW + 2) * 8);
    for (int mode = 0; mode < 3; ++mode) {
        float best = 1e30f;
        for (int rep = 0; rep < 6; ++rep) {
            float ms;
            cudaEventRecord(e0);
            for (int k = 0; k < 10; ++k) tile_kernel<TW, TH><<<grid, 512, 12 * TH * (TW + 2) * 8>>>(buf + (size_t)(k & 1) * 12 * ny * nx, ny, nx, mode);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep && ms < best) best = ms;
        }
        const double gb = 10.0 * 12 * (double)ny * nx * 8 / 1e9;
        printf("tile %3dx%-2d nx=%d mode %d: %7.1f GB/s (%.1f us per 12-plane pass) [%s]\n", TW, TH, nx, mode, gb / (best * 1e-3),
               best * 1e3 / 10, cudaGetErrorString(cudaGetLastError()));
    }
}

int main(int argc, char **argv) {
    const int nx = argc > 1 ? atoi(argv[1]) : 1785, ny = nx;
    double *buf;
    if (cudaMalloc(&buf, (size_t)2 * 12 * ny * nx * 8) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    run<32, 16>(buf, ny, nx);
    run<64, 8>(buf, ny, nx);
    run<128, 4>(buf, ny, nx);
    // check mode 1 wrote what mode 0 writes (last launch was mode 1 into buffer 1; redo mode 0 into buffer 0 and compare)
    tile_kernel<64, 8><<<dim3((nx + 63) / 64, (ny + 7) / 8), 512, 12 * 8 * 66 * 8>>>(buf, ny, nx, 0);
    tile_kernel<64, 8><<<dim3((nx + 63) / 64, (ny + 7) / 8), 512, 12 * 8 * 66 * 8>>>(buf + (size_t)12 * ny * nx, ny, nx, 1);
    cudaDeviceSynchronize();
    size_t n = (size_t)12 * ny * nx;
    double *h = (double *)malloc(2 * n * 8);
    cudaMemcpy(h, buf, 2 * n * 8, cudaMemcpyDeviceToHost);
    size_t bad = 0;
    for (size_t i = 0; i < n; ++i) bad += h[i] != h[n + i];
    printf("mismatches between mode 0 and mode 1: %zu   %s\n", bad, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
