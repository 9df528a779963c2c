// microbench.cu — calibration of the epilogue building blocks on B200 (not part of the product):
// TMEM read throughput (tcgen05.ld 32x32b.x32) and integer pipe throughput (IMAD, VIMNMX, VIMNMX3).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__device__ __forceinline__ void tmem_ld32(uint32_t taddr) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}

__global__ void k_tmem(int iters, long long* cycles) {
    __shared__ uint32_t tptr;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tptr)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tptr + ((uint32_t)((warp & 3) * 32) << 16);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < 8; ++c) tmem_ld32(base + ((warp >> 2) & 1) * 256 + c * 32);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tptr), "r"(512u) : "memory");
}

template <int MODE>
__global__ void k_alu(int iters, int* sink, long long* cycles) {
    int a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 7 + i; b[i] = threadIdx.x * 13 + i * 5; }
    int x = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { a[i] = a[i] * -512 + b[i]; }                                   // IMAD
            if (MODE == 1) { a[i] = min(a[i], b[i] ^ it); }                                  // LOP + VIMNMX
            if (MODE == 2) { a[i] = min(min(a[i], b[i]), x + it); }                          // VIMNMX3 (+IADD)
            if (MODE == 3) { int k = b[i] - 512 * (x + it); int t = max(a[i], k); a[i] = min(a[i], k); b[i] = min(b[i], t); }  // full top-2 step
        }
    }
    long long t1 = clock64();
    int s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i] + b[i];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
    long long* d_cyc; int* d_sink;
    cudaMalloc(&d_cyc, 148 * 8); cudaMalloc(&d_sink, 148 * 1024 * 4);
    long long h[148];
    for (int warps : {4, 8}) {
        const int iters = 2000;
        k_tmem<<<148, warps * 32>>>(iters, d_cyc);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, d_cyc, 148 * 8, cudaMemcpyDeviceToHost);
        double bytes = (double)iters * 8 * 32 * 32 * 4 * warps;   // per SM
        printf("tmem_ld warps=%d: %s  %.1f B/clk/SM  (%.1f elements/clk/SM)\n", warps, cudaGetErrorString(e), bytes / h[0], bytes / 4 / h[0]);
    }
    for (int mode = 0; mode < 4; ++mode)
        for (int warps : {4, 8, 16}) {
            const int iters = 4000;
            if (mode == 0) k_alu<0><<<148, warps * 32>>>(iters, d_sink, d_cyc);
            if (mode == 1) k_alu<1><<<148, warps * 32>>>(iters, d_sink, d_cyc);
            if (mode == 2) k_alu<2><<<148, warps * 32>>>(iters, d_sink, d_cyc);
            if (mode == 3) k_alu<3><<<148, warps * 32>>>(iters, d_sink, d_cyc);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(h, d_cyc, 148 * 8, cudaMemcpyDeviceToHost);
            double ops = (double)iters * 8 * warps * 32;
            const char* names[] = {"IMAD", "LOP+VIMNMX", "VIMNMX3+IADD", "top2-step(4 ops)"};
            printf("%s warps=%d: %s  %.1f lane-steps/clk/SM\n", names[mode], warps, cudaGetErrorString(e), ops / h[0]);
        }
    return 0;
}
