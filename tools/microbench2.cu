// microbench2.cu — B200 calibration of the three engines the value-only kernel shares per SM (not part of the product):
//   (1) tcgen05.mma.kind::i8 issue rate for M128 x N{64,128,256} x K32 with both operands in shared memory,
//   (2) tcgen05.ld 32x32b.x32 throughput with 4 / 8 / 16 reading warps,
//   (3) both at once (TMEM port contention), optionally with bulk-copy traffic into shared memory (the TMA producer).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../sfm-mvs-pipeline_b200/csrc -o microbench2 microbench2.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

#include "common.cuh"
using namespace sfm;

struct Params {
    int n;            // MMA N (64 | 128 | 256)
    int ksteps;       // MMAs per tile (K = 32 B each)
    int mma_tiles;    // tiles the MMA warp issues (0 = no MMA)
    int ld_warps;     // reading warps (0, 4, 8, 16)
    int ld_iters;     // 8-chunk read rounds per reading warp
    int copy_kb;      // KB per bulk copy into smem by the producer warp (0 = none)
    int copy_iters;
    int a_stride;     // 0: every MMA re-reads the same A K-slab; 1: walk the four K slabs (like the kernel)
};

__global__ void __launch_bounds__(640, 1) k_mix(Params p, const uint8_t* __restrict__ gsrc, long long* cyc, uint32_t* sink) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* bp = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t offA = 0, offB = 16384, offC = 16384 + 32768, offBar = offC + 65536, offT = offBar + 64;
    const uint32_t bar_mma = base + offBar, bar_cp = base + offBar + 8;
    volatile uint32_t* tptr = reinterpret_cast<volatile uint32_t*>(bp + offT);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(bp)[i] = i * 2654435761u;
    if (warp == 1 && lane == 0) { mbar_init(bar_mma, 1); mbar_init(bar_cp, 1); fence_barrier_init(); }
    if (warp == 2) { tmem_alloc(base + offT, 512); tmem_relinquish(); }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tptr;
    long long t0 = clock64(), t1 = t0;
    if (warp == 1) {
        if (lane == 0 && p.mma_tiles > 0) {
            const uint32_t idesc = umma_idesc_u8(128, p.n);
            const uint64_t ad = umma_desc_sw128(base + offA), bd = umma_desc_sw128(base + offB);
            const int stages = 512 / p.n;
            uint32_t ph = 0;
            for (int t = 0; t < p.mma_tiles; ++t) {
                const uint32_t d = tmem + (t % stages) * p.n;
                for (int k = 0; k < p.ksteps; ++k) {
                    const int ks = p.a_stride ? (k & 3) : 0;
                    umma_i8(d, ad + 2 * ks, bd + 2 * ks, idesc, k > 0);
                }
                if ((t & 63) == 63 || t == p.mma_tiles - 1) {          // bound the queue: wait every 64 tiles
                    umma_commit(bar_mma);
                    mbar_wait(bar_mma, ph);
                    ph ^= 1;
                }
            }
            t1 = clock64();
            cyc[blockIdx.x * 4 + 0] = t1 - t0;
        }
    } else if (warp == 0) {
        if (lane == 0 && p.copy_kb > 0) {
            uint32_t ph = 0;
            for (int i = 0; i < p.copy_iters; ++i) {
                mbar_arrive_expect_tx(bar_cp, p.copy_kb * 1024);
                for (int c = 0; c < p.copy_kb; c += 16)
                    bulk_load_1d(base + offC + c * 1024, gsrc + (static_cast<size_t>(blockIdx.x) * 64 + (c + i * 16) % 64) * 1024,
                                 16384, bar_cp);
                mbar_wait(bar_cp, ph);
                ph ^= 1;
            }
            cyc[blockIdx.x * 4 + 2] = clock64() - t0;
        }
    } else if (warp >= 4 && warp < 4 + p.ld_warps) {
        const uint32_t taddr = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + (((warp - 4) >> 2) & 1) * 256;
        uint32_t acc = 0;
        for (int it = 0; it < p.ld_iters; ++it) {
            uint32_t v[2][32];
            tmem_ld_32x32(taddr, v[0]);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                uint32_t (&cur)[32] = v[c & 1];
                asm volatile("tcgen05.wait::ld.sync.aligned;"
                             : "+r"(cur[0]), "+r"(cur[1]), "+r"(cur[2]), "+r"(cur[3]), "+r"(cur[4]), "+r"(cur[5]),
                               "+r"(cur[6]), "+r"(cur[7]), "+r"(cur[8]), "+r"(cur[9]), "+r"(cur[10]), "+r"(cur[11]),
                               "+r"(cur[12]), "+r"(cur[13]), "+r"(cur[14]), "+r"(cur[15]), "+r"(cur[16]),
                               "+r"(cur[17]), "+r"(cur[18]), "+r"(cur[19]), "+r"(cur[20]), "+r"(cur[21]),
                               "+r"(cur[22]), "+r"(cur[23]), "+r"(cur[24]), "+r"(cur[25]), "+r"(cur[26]),
                               "+r"(cur[27]), "+r"(cur[28]), "+r"(cur[29]), "+r"(cur[30]), "+r"(cur[31])
                             :: "memory");
                if (c + 1 < 8) tmem_ld_32x32(taddr + (c + 1) * 32, v[(c + 1) & 1]);
#pragma unroll
                for (int j = 0; j < 32; ++j) acc ^= cur[j];
            }
        }
        if (warp == 4 && lane == 0) cyc[blockIdx.x * 4 + 1] = clock64() - t0;
        if (acc == 0x12345678u) sink[threadIdx.x] = acc;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem, 512);
}

static void run(const char* name, Params p, const uint8_t* gsrc, long long* d_cyc, uint32_t* d_sink) {
    const int grid = 148, smem = 16384 + 32768 + 65536 + 1024 + 2048;
    cudaFuncSetAttribute(k_mix, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaMemset(d_cyc, 0, grid * 4 * 8);
    k_mix<<<grid, 640, smem>>>(p, gsrc, d_cyc, d_sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148 * 4];
    cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double m = 0, l = 0, c = 0;
    for (int i = 0; i < grid; ++i) { m += h[i * 4]; l += h[i * 4 + 1]; c += h[i * 4 + 2]; }
    m /= grid; l /= grid; c /= grid;
    printf("%-44s %s |", name, cudaGetErrorString(e));
    if (p.mma_tiles) printf(" mma: %.1f cyc/MMA (N=%d, %.0f MAC/clk/SM)", m / (double(p.mma_tiles) * p.ksteps), p.n,
                            128.0 * p.n * 32 * p.mma_tiles * p.ksteps / m);
    if (p.ld_warps) printf(" | ld: %.1f B/clk/SM (%d warps, %.0f cyc per 32x32 load)", double(p.ld_warps) * p.ld_iters * 8 * 4096 / l,
                           p.ld_warps, l / (p.ld_iters * 8.0));
    if (p.copy_kb) printf(" | copy: %.1f B/clk/SM", double(p.copy_kb) * 1024 * p.copy_iters / c);
    printf("\n");
}

// Tight issue loop: compile-time N / K-steps / accumulator pattern, nothing but tcgen05.mma in the loop.
//   kIndep = 1: one accumulator per tile (dependent chain of kK MMAs);  2: two tiles interleaved (A k0, B k0, A k1, ...)
template <int kN, int kK, int kIndep>
__global__ void __launch_bounds__(640, 1) k_tight(int tiles, long long* cyc, int ld_warps, int ld_iters, int alu, uint32_t* sink, int commits, int ext) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* bp = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t offA = 0, offB = 16384, offBar = 16384 + 32768, offT = offBar + 64;
    const uint32_t bar_mma = base + offBar, bar_c1 = base + offBar + 8, bar_c2 = base + offBar + 16;
    volatile uint32_t* tptr = reinterpret_cast<volatile uint32_t*>(bp + offT);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(bp)[i] = i * 2654435761u;
    if (warp == 1 && lane == 0) { mbar_init(bar_mma, 1); mbar_init(bar_c1, 1); mbar_init(bar_c2, 1); fence_barrier_init(); }
    if (warp == 2) { tmem_alloc(base + offT, 512); tmem_relinquish(); }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tptr;
    if (warp == 1 && lane == 0) {
        constexpr uint32_t idesc = umma_idesc_u8(128, kN);
        const uint64_t ad = umma_desc_sw128(base + offA), bd = umma_desc_sw128(base + offB);
        long long t0 = clock64();
        for (int t = 0; t < tiles; t += kIndep) {
            const uint32_t d0 = tmem + ((t & 2) ? 2 * kN % 512 : 0);
#pragma unroll
            for (int k = 0; k < kK; ++k) {
#pragma unroll
                for (int j = 0; j < kIndep; ++j)
                    umma_i8(d0 + j * (kN % 256), ad + 2 * (k & 3), bd + 2 * (k & 3), idesc, k > 0 ? 1u : 0u);
            }
            if (ext) umma_i8(d0, umma_desc_sw32(base + offA), umma_desc_sw32(base + offB), umma_idesc_u8s8(128, kN), 1u);
            if (commits > 0) umma_commit(bar_c1);
            if (commits > 1) umma_commit(bar_c2);
        }
        umma_commit(bar_mma);
        mbar_wait(bar_mma, 0);
        cyc[blockIdx.x] = clock64() - t0;
    } else if (warp >= 4 && warp < 4 + ld_warps) {
        // reading warps: group g = (warp-4)>>2 reads column half g of BOTH stages alternately (like the epilogue)
        const uint32_t taddr = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + (((warp - 4) >> 2) & 1) * 128;
        uint32_t acc = 0;
        long long t0 = clock64();
        for (int it = 0; it < ld_iters; ++it) {
            uint32_t v[2][32];
            const uint32_t ta = taddr + (it & 1) * 256;
            tmem_ld_32x32(ta, v[0]);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t (&cur)[32] = v[c & 1];
                asm volatile("tcgen05.wait::ld.sync.aligned;"
                             : "+r"(cur[0]), "+r"(cur[1]), "+r"(cur[2]), "+r"(cur[3]), "+r"(cur[4]), "+r"(cur[5]),
                               "+r"(cur[6]), "+r"(cur[7]), "+r"(cur[8]), "+r"(cur[9]), "+r"(cur[10]), "+r"(cur[11]),
                               "+r"(cur[12]), "+r"(cur[13]), "+r"(cur[14]), "+r"(cur[15]), "+r"(cur[16]),
                               "+r"(cur[17]), "+r"(cur[18]), "+r"(cur[19]), "+r"(cur[20]), "+r"(cur[21]),
                               "+r"(cur[22]), "+r"(cur[23]), "+r"(cur[24]), "+r"(cur[25]), "+r"(cur[26]),
                               "+r"(cur[27]), "+r"(cur[28]), "+r"(cur[29]), "+r"(cur[30]), "+r"(cur[31])
                             :: "memory");
                if (c + 1 < 4) tmem_ld_32x32(ta + (c + 1) * 32, v[(c + 1) & 1]);
                if (alu) {
                    int32_t a[11];
#pragma unroll
                    for (int i = 0; i < 10; ++i) a[i] = __vimax3_s32((int)cur[3 * i], (int)cur[3 * i + 1], (int)cur[3 * i + 2]);
                    a[10] = max((int)cur[30], (int)cur[31]);
                    const int32_t b0 = __vimax3_s32(a[0], a[1], a[2]), b1 = __vimax3_s32(a[3], a[4], a[5]);
                    const int32_t b2 = __vimax3_s32(a[6], a[7], a[8]), b3 = max(a[9], a[10]);
                    acc = max((int)acc, max(__vimax3_s32(b0, b1, b2), b3));
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc ^= cur[j];
                }
            }
        }
        if (lane == 0 && warp == 4) cyc[148 + blockIdx.x] = clock64() - t0;
        if (acc == 0x12345678u) sink[threadIdx.x] = acc;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem, 512);
}

template <int kN, int kK, int kIndep>
static void run_tight(long long* d_cyc, int ld_warps = 0, int alu = 0, uint32_t* d_sink = nullptr, int commits = 0, int ext = 0) {
    const int grid = 148, smem = 16384 + 32768 + 1024 + 2048, tiles = 8192;
    cudaFuncSetAttribute(k_tight<kN, kK, kIndep>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int ld_iters = ld_warps ? tiles * 2 : 0;   // 4 chunks x half the columns per iteration = one tile per 2 iterations per group
    k_tight<kN, kK, kIndep><<<grid, 128 + 32 * ld_warps, smem>>>(tiles, d_cyc, ld_warps, ld_iters, alu, d_sink, commits, ext);
    if (commits || ext) printf("[commits/tile=%d ext-mma=%d] ", commits, ext);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[296];
    cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double m = 0, l = 0;
    for (int i = 0; i < grid; ++i) { m += h[i]; l += h[148 + i]; }
    m /= grid; l /= grid;
    printf("tight N=%3d k=%d accumulators=%d: %s  %.1f cyc/MMA, %.1f cyc/tile, %.0f MAC/clk/SM", kN, kK, kIndep,
           cudaGetErrorString(e), m / (double(tiles) * kK), m / tiles, 128.0 * kN * 32 * tiles * kK / m);
    if (ld_warps) printf(" | + %d ld warps (alu=%d): %.1f B/clk/SM, %.0f cyc per 32x32 load", ld_warps, alu,
                         double(ld_warps) * ld_iters * 4 * 4096 / l, l / (ld_iters * 4.0));
    printf("\n");
}

int main() {
    uint8_t* gsrc; long long* d_cyc; uint32_t* d_sink;
    cudaMalloc(&gsrc, size_t(148) * 64 * 1024 + 65536);
    cudaMemset(gsrc, 1, size_t(148) * 64 * 1024 + 65536);
    cudaMalloc(&d_cyc, 148 * 4 * 8);
    cudaMemset(d_cyc, 0, 148 * 4 * 8);
    cudaMalloc(&d_sink, 4096);
    for (int cm : {0, 1, 2}) run_tight<256, 5, 1>(d_cyc, 0, 0, d_sink, cm, 0);
    for (int cm : {0, 2}) run_tight<256, 4, 1>(d_cyc, 0, 0, d_sink, cm, 1);
    run_tight<256, 4, 1>(d_cyc, 8, 1, d_sink, 2, 1);
    for (int cm : {0, 2}) run_tight<128, 4, 1>(d_cyc, 0, 0, d_sink, cm, 1);
    for (int w : {4, 8, 16}) for (int alu : {0, 1}) run_tight<256, 5, 1>(d_cyc, w, alu, d_sink);
    for (int w : {8}) for (int alu : {0, 1}) run_tight<256, 4, 1>(d_cyc, w, alu, d_sink);
    run_tight<64, 4, 1>(d_cyc); run_tight<64, 4, 2>(d_cyc); run_tight<64, 16, 1>(d_cyc);
    run_tight<128, 4, 1>(d_cyc); run_tight<128, 5, 1>(d_cyc); run_tight<128, 4, 2>(d_cyc); run_tight<128, 16, 1>(d_cyc);
    run_tight<256, 1, 1>(d_cyc); run_tight<256, 2, 1>(d_cyc); run_tight<256, 4, 1>(d_cyc); run_tight<256, 5, 1>(d_cyc);
    run_tight<256, 8, 1>(d_cyc); run_tight<256, 16, 1>(d_cyc); run_tight<256, 4, 2>(d_cyc); run_tight<256, 5, 2>(d_cyc);
    for (int n : {64, 128, 256}) {
        char nm[96];
        snprintf(nm, sizeof nm, "mma only N=%d k=4 same-A-slab", n);
        run(nm, Params{n, 4, 4096, 0, 0, 0, 0, 0}, gsrc, d_cyc, d_sink);
        snprintf(nm, sizeof nm, "mma only N=%d k=4 walking K slabs", n);
        run(nm, Params{n, 4, 4096, 0, 0, 0, 0, 1}, gsrc, d_cyc, d_sink);
        snprintf(nm, sizeof nm, "mma only N=%d k=5", n);
        run(nm, Params{n, 5, 4096, 0, 0, 0, 0, 1}, gsrc, d_cyc, d_sink);
    }
    for (int w : {4, 8, 16}) {
        char nm[96];
        snprintf(nm, sizeof nm, "ld only warps=%d", w);
        run(nm, Params{256, 4, 0, w, 4096, 0, 0, 0}, gsrc, d_cyc, d_sink);
    }
    run("copy only 32 KB", Params{256, 4, 0, 0, 0, 32, 4096, 0}, gsrc, d_cyc, d_sink);
    run("copy only 64 KB", Params{256, 4, 0, 0, 0, 64, 2048, 0}, gsrc, d_cyc, d_sink);
    // concurrent: sized so that both sides run for about the same time
    for (int n : {128, 256}) {
        for (int w : {4, 8}) {
            char nm[96];
            snprintf(nm, sizeof nm, "mma N=%d k=5 + ld warps=%d", n, w);
            run(nm, Params{n, 5, 4096 * 256 / n, w, w == 4 ? 5000 : 2500, 0, 0, 1}, gsrc, d_cyc, d_sink);
        }
        char nm[96];
        snprintf(nm, sizeof nm, "mma N=%d k=5 + copy 32KB", n);
        run(nm, Params{n, 5, 4096 * 256 / n, 0, 0, 32, 6000, 1}, gsrc, d_cyc, d_sink);
        snprintf(nm, sizeof nm, "mma N=%d k=5 + ld warps=8 + copy 32KB", n);
        run(nm, Params{n, 5, 4096 * 256 / n, 8, 2500, 32, 6000, 1}, gsrc, d_cyc, d_sink);
    }
    return 0;
}
