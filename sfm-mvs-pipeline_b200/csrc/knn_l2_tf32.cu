// knn_l2_tf32.cu — non-integer float descriptors (128 x CV_32F) on the tensor cores: 3xTF32 candidate search.
//
// north_star (2): "Otherwise a tf32/3xTF32 path re-ranks the top candidates in fp32, with a stated tolerance of
// identical match indices except on distance ties within 1e-5 relative."
//
// Every bank row x is split at upload into hi = tf32(x) (low 13 mantissa bits cleared) and lo = x - hi (exact), so
//     a.b ~= hi_a.hi_b + hi_a.lo_b + lo_a.hi_b          (three tcgen05.mma.kind::tf32 per K step, fp32 accumulate)
// is accurate to ~2^-20 |a||b|.  One more K step (A = rows {1,1,1,0,...}, B = {-g_hi,-g_mid,-g_lo,0,...}, g = |b|^2/2
// split into three tf32-exact pieces) makes the accumulator  D~ = a.b - |b|^2/2  up to that error: larger D~ = closer.
// The epilogue (thread = TMEM lane = query row) keeps, per query row, the top-4 32-column chunks by chunk maximum of
// D~ plus the 5th-best value.  refine_f32_kernel (post.cu) recomputes the rows of those chunks exactly in fp32
// (sum (a-b)^2, the formulation cv::batchDistance uses), takes the exact top-2 and CERTIFIES it: every row outside
// the chunks has D~ <= v5, i.e. d^2 >= |a|^2 - 2 (v5 + eps); if the exact second neighbour is below that bound the
// answer is final, otherwise the row is brute-forced.  So the tensor cores only ever select candidates; every reported
// distance is an fp32 computation.
//
// Tiles: M = 128 query rows (A: hi + lo, 128 KB, resident per unit), N = 128 train rows per tile, K streamed in four
// 128-byte slabs (32 floats) per operand half: a B stage = one slab of hi + lo (32 KB), two stages.  49 MMAs per tile.
#include <cuda.h>

#include "common.cuh"
#include "kernels.h"

namespace sfm {

namespace tf {
constexpr int BM = 128, BN = 128, kSlabs = 4;
constexpr int kBStages = 2, kAccStages = 2;
constexpr int kSlabBytesA = BM * 128, kSlabBytesB = BN * 128;     // one 128-byte-wide slab of hi or lo
constexpr int kABytes = 2 * kSlabs * kSlabBytesA;                 // hi[4] | lo[4]
constexpr int kBStageBytes = 2 * kSlabBytesB;                     // hi slab | lo slab
constexpr int kEBytes = BN * 32;                                  // per tile: 8 floats per train row
constexpr int kAExtBytes = BM * 32;
constexpr int kThreads = 256;                                     // warps 0-3 control, 4-7 epilogue
constexpr int kEpiWarp0 = 4;
constexpr int offA = 0;
constexpr int offB = offA + kABytes;
constexpr int offE = offB + kBStages * kBStageBytes;              // 2 slots (tile parity)
constexpr int offAExt = offE + 2 * kEBytes;
constexpr int offBar = offAExt + kAExtBytes;
constexpr int kNumBars = 2 * kBStages + 2 + 2 * kAccStages;       // b_full/empty, a_full/empty, acc_full/empty
constexpr int offTmemPtr = offBar + kNumBars * 8;
constexpr int kSmemBytes = offTmemPtr + 16 + 1024;
// kind::tf32: D = f32 (1 << 4), A = B = tf32 (2 << 7, 2 << 10), K-major, M x N
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(BN >> 3) << 17) |
                            (static_cast<uint32_t>(BM >> 4) << 24);
}  // namespace tf

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

struct UnitInfoF { PairDesc pd; int rb; int n_tiles; };
__device__ __forceinline__ UnitInfoF decode_unit_f(const PairDesc* __restrict__ pairs, const int64_t* __restrict__ unit_prefix,
                                                   int n_pairs, int64_t unit) {
    UnitInfoF u;
    const int p = find_segment(unit_prefix, n_pairs, unit);
    u.pd = pairs[p];
    u.rb = static_cast<int>(unit - unit_prefix[p]);
    u.n_tiles = (u.pd.nt + tf::BN - 1) / tf::BN;
    return u;
}

__global__ void __launch_bounds__(tf::kThreads, 1)
knn2_l2_f32_tc3_kernel(const __grid_constant__ CUtensorMap tmap_hi_a, const __grid_constant__ CUtensorMap tmap_lo_a,
                       const __grid_constant__ CUtensorMap tmap_hi_b, const __grid_constant__ CUtensorMap tmap_lo_b,
                       const __grid_constant__ CUtensorMap tmap_e, const PairDesc* __restrict__ pairs,
                       const int64_t* __restrict__ unit_prefix, int n_pairs, int64_t n_units, Top2* __restrict__ out,
                       float* __restrict__ aux) {
    using namespace tf;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bar0 = base + offBar;
    auto b_full = [&](int i) { return bar0 + 8u * i; };
    auto b_empty = [&](int i) { return bar0 + 8u * (kBStages + i); };
    const uint32_t a_full = bar0 + 8u * (2 * kBStages), a_empty = bar0 + 8u * (2 * kBStages + 1);
    auto acc_full = [&](int i) { return bar0 + 8u * (2 * kBStages + 2 + i); };
    auto acc_empty = [&](int i) { return bar0 + 8u * (2 * kBStages + 2 + kAccStages + i); };
    volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(base_ptr + offTmemPtr);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_hi_a); prefetch_tmap(&tmap_lo_a); prefetch_tmap(&tmap_hi_b); prefetch_tmap(&tmap_lo_b);
        prefetch_tmap(&tmap_e);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kBStages; ++i) { mbar_init(b_full(i), 1); mbar_init(b_empty(i), 1); }
        mbar_init(a_full, 1); mbar_init(a_empty, 1);
        for (int i = 0; i < kAccStages; ++i) { mbar_init(acc_full(i), 1); mbar_init(acc_empty(i), 128); }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(base + offTmemPtr, 256);
        tmem_relinquish();
    }
    // constant A operand of the norm K step: rows {1,1,1,0, 1,1,1,0} (both 16-byte halves equal -> swizzle-invariant;
    // the second half meets zeros in B)
    for (int i = threadIdx.x; i < kAExtBytes / 4; i += kThreads)
        reinterpret_cast<float*>(base_ptr + offAExt)[i] = (i & 3) < 3 ? 1.0f : 0.0f;
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        // ================================================================ TMA producer
        uint32_t slab_iter = 0, tile_iter = 0, unit_iter = 0;
        for (int64_t unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
            const UnitInfoF u = decode_unit_f(pairs, unit_prefix, n_pairs, unit);
            if (u.n_tiles == 0) continue;
            mbar_wait(a_empty, (unit_iter & 1) ^ 1);
            if (elect_one()) {
                mbar_arrive_expect_tx(a_full, kABytes);
                const int row = u.pd.q_row0 + u.rb * BM;
#pragma unroll
                for (int sl = 0; sl < kSlabs; ++sl) {
                    tma_load_2d(base + offA + sl * kSlabBytesA, &tmap_hi_a, sl * 128, row, a_full);
                    tma_load_2d(base + offA + (kSlabs + sl) * kSlabBytesA, &tmap_lo_a, sl * 128, row, a_full);
                }
            }
            for (int t = 0; t < u.n_tiles; ++t, ++tile_iter) {
                const int row = u.pd.t_row0 + t * BN;
                for (int sl = 0; sl < kSlabs; ++sl, ++slab_iter) {
                    const int st = slab_iter % kBStages;
                    mbar_wait(b_empty(st), ((slab_iter / kBStages) & 1) ^ 1);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(b_full(st), kBStageBytes + (sl == 0 ? kEBytes : 0));
                        tma_load_2d(base + offB + st * kBStageBytes, &tmap_hi_b, sl * 128, row, b_full(st));
                        tma_load_2d(base + offB + st * kBStageBytes + kSlabBytesB, &tmap_lo_b, sl * 128, row, b_full(st));
                        // the tile's norm rows ride with slab 0; slot = tile parity (free again: the MMA that read it
                        // two tiles ago retired before the stage we are refilling was released)
                        if (sl == 0) tma_load_2d(base + offE + (tile_iter & 1) * kEBytes, &tmap_e, 0, row, b_full(st));
                    }
                    __syncwarp();
                }
            }
            ++unit_iter;
        }
    } else if (warp == 1) {
        // ================================================================ MMA issuer
        uint32_t slab_iter = 0, tile_iter = 0, unit_iter = 0;
        const uint64_t aext = umma_desc_sw32(base + offAExt);
        for (int64_t unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
            const UnitInfoF u = decode_unit_f(pairs, unit_prefix, n_pairs, unit);
            if (u.n_tiles == 0) continue;
            mbar_wait(a_full, unit_iter & 1);
            for (int t = 0; t < u.n_tiles; ++t, ++tile_iter) {
                const int acc = tile_iter % kAccStages;
                mbar_wait(acc_empty(acc), ((tile_iter / kAccStages) & 1) ^ 1);
                const uint32_t d = tmem_base + acc * BN;
                for (int sl = 0; sl < kSlabs; ++sl, ++slab_iter) {
                    const int st = slab_iter % kBStages;
                    mbar_wait(b_full(st), (slab_iter / kBStages) & 1);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t a_hi = umma_desc_sw128(base + offA + sl * kSlabBytesA);
                        const uint64_t a_lo = umma_desc_sw128(base + offA + (kSlabs + sl) * kSlabBytesA);
                        const uint64_t b_hi = umma_desc_sw128(base + offB + st * kBStageBytes);
                        const uint64_t b_lo = umma_desc_sw128(base + offB + st * kBStageBytes + kSlabBytesB);
                        if (sl == 0)                                    // D = -|b|^2/2 (three exact tf32 pieces)
                            umma_tf32(d, aext, umma_desc_sw32(base + offE + (tile_iter & 1) * kEBytes), kIdesc, 0);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {                   // K = 8 floats (32 B) per tcgen05.mma.kind::tf32
                            umma_tf32(d, a_lo + 2 * k, b_hi + 2 * k, kIdesc, 1);      // small terms first
                            umma_tf32(d, a_hi + 2 * k, b_lo + 2 * k, kIdesc, 1);
                            umma_tf32(d, a_hi + 2 * k, b_hi + 2 * k, kIdesc, 1);
                        }
                        umma_commit(b_empty(st));
                        if (sl == kSlabs - 1) {
                            umma_commit(acc_full(acc));
                            if (t == u.n_tiles - 1) umma_commit(a_empty);
                        }
                    }
                    __syncwarp();
                }
            }
            ++unit_iter;
        }
    } else if (warp >= kEpiWarp0) {
        // ================================================================ epilogue: top-4 chunks (+ 5th value) per query row
        const int quarter = warp & 3;
        const int row_in_unit = quarter * 32 + lane;
        uint32_t tile_iter = 0;
        const float ninf = __int_as_float(0xff800000);
        for (int64_t unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
            const UnitInfoF u = decode_unit_f(pairs, unit_prefix, n_pairs, unit);
            float v0 = ninf, v1 = ninf, v2 = ninf, v3 = ninf, v4 = ninf;
            int c0 = 0xFFFF, c1 = 0xFFFF, c2 = 0xFFFF, c3 = 0xFFFF;
            for (int t = 0; t < u.n_tiles; ++t, ++tile_iter) {
                const int acc = tile_iter % kAccStages;
                mbar_wait(acc_full(acc), (tile_iter / kAccStages) & 1);
                tc_fence_after();
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;
#pragma unroll 1
                for (int c = 0; c < BN / 32; ++c) {
                    uint32_t r[32];
                    tmem_ld_32x32(taddr + c * 32, r);
                    asm volatile("tcgen05.wait::ld.sync.aligned;"
                                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]),
                                   "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]),
                                   "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]),
                                   "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]),
                                   "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                                 :: "memory");
                    float x = __uint_as_float(r[0]);
#pragma unroll
                    for (int j = 1; j < 32; ++j) x = fmaxf(x, __uint_as_float(r[j]));
                    const int cx = t * (BN / 32) + c;
                    if (x > v4) {                                        // strict '>': equal maxima keep the earlier chunk
                        if (x > v0)      { v4 = v3; v3 = v2; c3 = c2; v2 = v1; c2 = c1; v1 = v0; c1 = c0; v0 = x; c0 = cx; }
                        else if (x > v1) { v4 = v3; v3 = v2; c3 = c2; v2 = v1; c2 = c1; v1 = x; c1 = cx; }
                        else if (x > v2) { v4 = v3; v3 = v2; c3 = c2; v2 = x; c2 = cx; }
                        else if (x > v3) { v4 = v3; v3 = x; c3 = cx; }
                        else             { v4 = x; }
                    }
                }
                tc_fence_before();
                mbar_arrive(acc_empty(acc));
            }
            const int row = u.rb * BM + row_in_unit;
            if (row < u.pd.nq) {
                Top2 o;
                o.i0 = c0 | (c1 << 16);
                o.i1 = c2 | (c3 << 16);
                o.d0 = v0;
                o.d1 = v1;
                out[u.pd.out_row0 + row] = o;
                aux[u.pd.out_row0 + row] = v4;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 256);
}

cudaError_t launch_knn2_l2_f32_tc3(const void* tmaps /* 5 CUtensorMap: hi_a, lo_a, hi_b, lo_b, ext */, const PairDesc* pairs,
                                   const int64_t* unit_prefix, int n_pairs, int64_t n_units, Top2* out, float* aux,
                                   int sm_count, cudaStream_t s) {
    if (n_units == 0) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(knn2_l2_f32_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tf::kSmemBytes);
    if (e != cudaSuccess) return e;
    const CUtensorMap* t = static_cast<const CUtensorMap*>(tmaps);
    const int grid = static_cast<int>(n_units < sm_count ? n_units : sm_count);
    knn2_l2_f32_tc3_kernel<<<grid, tf::kThreads, tf::kSmemBytes, s>>>(t[0], t[1], t[2], t[3], t[4], pairs, unit_prefix, n_pairs,
                                                                      n_units, out, aux);
    return cudaGetLastError();
}

}  // namespace sfm
