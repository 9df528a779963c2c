// knn_l2_tc.cu — SIFT (128-byte, u8-valued) squared-L2 knn(k=2) on the 5th-gen tensor cores.
//
// Replaces cv::batchDistance(NORM_L2, K=2) under BFMatcher::knnMatch for every pair the strategies issue
// (UnorderedFeatureMatchingStrategy.cpp:51, VideoFeatureMatchingStrategy.cpp:62, GridFeatureMatchingStrategy.cpp:105).
//
// dist^2(a,b) = |a|^2 + |b|^2 - 2 a.b ; a.b is a dense u8 x u8 -> s32 contraction (K = 128) computed with
// tcgen05.mma.kind::i8, operands staged in shared memory by TMA (128-byte swizzle, one swizzle atom per
// descriptor row), accumulators in TMEM.  The per-row top-2 never leaves the SM: epilogue warps read the
// accumulators with tcgen05.ld (one TMEM lane = one query row = one thread) and keep a running top-2 of
//     key = (|b_j|^2 - 2 a.b_j) * 256 + (j mod 256)          = ckey_j - 512 * acc
// a single IMAD per element from a per-train-row constant ckey_j that TMA-bulk-copies in with the tile.
// Keys are unique inside a tile, so "ties -> lowest trainIdx"  is free; across tiles the
// running state is a 64-bit (value, global column) key.  All arithmetic is integer: results are bit-exact.
// To keep the epilogue at ~0.6 ALU op per element the kernel tracks the top-2 of 32-column CHUNK MINIMA
// (VIMNMX3 trees): rank 1 is exact, rank 2 is the best element outside rank 1's chunk.  The true second
// neighbour is min(that, second-best inside rank 1's chunk); refine_second_kernel recomputes those 31
// distances with __dp4a, and only for rows whose provisional ratio test passes (a fraction of a percent).
//
// Persistent kernel, one CTA per SM, warp-specialised:
//   warp 0      TMA producer   (A tile once per unit, B tile + ckeys per train tile, 4-stage ring)
//   warps 1, 3  MMA issuers    (4 x tcgen05.mma M128 N256 K32 per K slab and tile; warp 1 issues the even tiles into
//                               TMEM accumulator stage 0, warp 3 the odd tiles into stage 1 — one thread cannot issue
//                               fast enough, see profiles/r1_tcv_issue_analysis.txt)
//   warp 2      TMEM allocator
//   warps 4-7   epilogue group 0 (columns   0..127 of every tile) } thread = query row, TMEM lane quarter = warp % 4
//   warps 8-11  epilogue group 1 (columns 128..255 of every tile) }
// unit = (pair, block of 128 query rows); units are dealt round-robin to CTAs so that concurrently running
// CTAs stream the same train image out of L2.
#include <cuda.h>

#include "common.cuh"
#include "kernels.h"

namespace sfm {

namespace tc {
constexpr int BM = 128, BN = 256;
constexpr int kAStages = 2, kAccStages = 2, kCkSlots = 8;
constexpr int kCkBytes = BN * 4;
constexpr int kThreads = 384;
constexpr int kEpiWarp0 = 4;
constexpr uint32_t kIdesc = umma_idesc_u8(BM, BN);
constexpr int64_t kEmptyKey = INT64_MAX;
// kSlabs = number of 128-byte K slabs of a descriptor row: 1 = SIFT (128 x u8), 2 = ORB bits expanded to 256 x u8
// (Hamming(a,b) = |a| + |b| - 2 a.b is the same contraction).  Shared memory carve-up from a 1024-byte aligned base.
template <int kSlabs>
struct Cfg {
    static constexpr int KB = 128 * kSlabs;
    static constexpr int kBStages = kSlabs == 1 ? 4 : 2;
    static constexpr int kABytes = BM * KB, kBBytes = BN * KB;
    static constexpr int offB = 0;
    static constexpr int offA = offB + kBStages * kBBytes;
    static constexpr int offCk = offA + kAStages * kABytes;
    static constexpr int offMerge = offCk + kCkSlots * kCkBytes;            // [128][2] int64
    static constexpr int offBar = offMerge + BM * 2 * 8;
    static constexpr int kNumBars = 2 * kBStages + 2 * kAStages + 2 * kAccStages + kCkSlots;
    static constexpr int offTmemPtr = offBar + kNumBars * 8;
    static constexpr int kSmemBytes = offTmemPtr + 16 + 1024;               // + alignment slack
};
}  // namespace tc

struct UnitInfo { PairDesc pd; int rb; int n_tiles; };

__device__ __forceinline__ UnitInfo decode_unit(const PairDesc* __restrict__ pairs, const int64_t* __restrict__ unit_prefix,
                                                int n_pairs, int64_t unit) {
    UnitInfo u;
    const int p = find_segment(unit_prefix, n_pairs, unit);
    u.pd = pairs[p];
    u.rb = static_cast<int>(unit - unit_prefix[p]);
    u.n_tiles = (u.pd.nt + tc::BN - 1) / tc::BN;
    return u;
}

// (m1, m2) <- two smallest of {m1, m2, k}, m1 <= m2
__device__ __forceinline__ void top2_key(int32_t k, int32_t& m1, int32_t& m2) {
    m2 = min(m2, max(m1, k));
    m1 = min(m1, k);
}
__device__ __forceinline__ void top2_key64(int64_t k, int64_t& m1, int64_t& m2) {
    m2 = min(m2, max(m1, k));
    m1 = min(m1, k);
}

template <int kSlabs>
__global__ void __launch_bounds__(tc::kThreads, 1)
knn2_l2_u8_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                     const int32_t* __restrict__ ckey, const int32_t* __restrict__ norm2,
                     const PairDesc* __restrict__ pairs, const int64_t* __restrict__ unit_prefix, int n_pairs,
                     int64_t n_units, Top2* __restrict__ out) {
    using namespace tc;
    using C = Cfg<kSlabs>;
    constexpr int kBStages = C::kBStages, kABytes = C::kABytes, kBBytes = C::kBBytes;
    constexpr int offA = C::offA, offB = C::offB, offCk = C::offCk, offMerge = C::offMerge, offBar = C::offBar;
    constexpr int offTmemPtr = C::offTmemPtr;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bar0 = base + offBar;
    auto b_full = [&](int i) { return bar0 + 8u * i; };
    auto b_empty = [&](int i) { return bar0 + 8u * (kBStages + i); };
    auto a_full = [&](int i) { return bar0 + 8u * (2 * kBStages + i); };
    auto a_empty = [&](int i) { return bar0 + 8u * (2 * kBStages + kAStages + i); };
    auto acc_full = [&](int i) { return bar0 + 8u * (2 * kBStages + 2 * kAStages + i); };
    auto acc_empty = [&](int i) { return bar0 + 8u * (2 * kBStages + 2 * kAStages + kAccStages + i); };
    auto ck_full = [&](int i) { return bar0 + 8u * (2 * kBStages + 2 * kAStages + 2 * kAccStages + i); };
    volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(base_ptr + offTmemPtr);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_a);
        prefetch_tmap(&tmap_b);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kBStages; ++i) { mbar_init(b_full(i), 1); mbar_init(b_empty(i), 1); }
        for (int i = 0; i < kAStages; ++i) { mbar_init(a_full(i), 1); mbar_init(a_empty(i), 2); }    // both MMA warps release A
        for (int i = 0; i < kAccStages; ++i) { mbar_init(acc_full(i), 1); mbar_init(acc_empty(i), 8); }  // one arrival per epilogue warp
        for (int i = 0; i < kCkSlots; ++i) mbar_init(ck_full(i), 1);
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(base + offTmemPtr, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        // ================================================================ TMA producer
        // the whole warp walks the schedule (keeps the warp converged for the final barrier); lane 0 issues
        uint32_t tile_iter = 0, unit_iter = 0;
        for (int64_t unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
            const UnitInfo u = decode_unit(pairs, unit_prefix, n_pairs, unit);
            if (u.n_tiles == 0) continue;
            const int as = unit_iter % kAStages;
            mbar_wait(a_empty(as), ((unit_iter / kAStages) & 1) ^ 1);
            if (elect_one()) {
                mbar_arrive_expect_tx(a_full(as), kABytes);
#pragma unroll
                for (int sl = 0; sl < kSlabs; ++sl)      // one 128-byte-wide box per K slab
                    tma_load_2d(base + offA + as * kABytes + sl * (BM * 128), &tmap_a, sl * 128, u.pd.q_row0 + u.rb * BM,
                                a_full(as));
            }
            for (int t = 0; t < u.n_tiles; ++t, ++tile_iter) {
                const int st = tile_iter % kBStages;
                mbar_wait(b_empty(st), ((tile_iter / kBStages) & 1) ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(b_full(st), kBBytes);
#pragma unroll
                    for (int sl = 0; sl < kSlabs; ++sl)
                        tma_load_2d(base + offB + st * kBBytes + sl * (BN * 128), &tmap_b, sl * 128, u.pd.t_row0 + t * BN,
                                    b_full(st));
                    // ckey ring: 8 slots, the producer is never more than kBStages + kAccStages = 6 tiles ahead
                    // of the epilogue, so a slot is free again before it is reloaded
                    const int cs = tile_iter % kCkSlots;
                    mbar_arrive_expect_tx(ck_full(cs), kCkBytes);
                    bulk_load_1d(base + offCk + cs * kCkBytes, ckey + u.pd.t_row0 + t * BN, kCkBytes, ck_full(cs));
                }
                __syncwarp();
            }
            ++unit_iter;
        }
    } else if (warp == 1 || warp == 3) {
        // ================================================================ MMA issuers (lane 0 issues)
        const uint32_t my_par = static_cast<uint32_t>(warp >> 1);       // warp 1: even tiles, warp 3: odd tiles
        uint32_t tile0 = 0, unit_iter = 0;                              // tile0: running tile number at the start of the unit
        for (int64_t unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
            const UnitInfo u = decode_unit(pairs, unit_prefix, n_pairs, unit);
            if (u.n_tiles == 0) continue;
            const int as = unit_iter % kAStages;
            mbar_wait(a_full(as), (unit_iter / kAStages) & 1);
            const uint32_t a_smem = base + offA + as * kABytes;
            const int first = static_cast<int>((my_par - tile0) & 1u);  // my first tile of this unit
            const int last = first < u.n_tiles ? first + 2 * ((u.n_tiles - 1 - first) / 2) : -1;
            if (last < 0 && lane == 0) mbar_arrive(a_empty(as));        // no tile of this unit is mine
            for (int t = first; t < u.n_tiles; t += 2) {
                const uint32_t tile_iter = tile0 + t;
                const int acc = tile_iter % kAccStages;
                const int st = tile_iter % kBStages;
                {
                    const uint32_t par_b = (tile_iter / kBStages) & 1, par_acc = ((tile_iter / kAccStages) & 1) ^ 1;
                    bool ok_b = mbar_try_wait(b_full(st), par_b);
                    bool ok_acc = mbar_try_wait(acc_empty(acc), par_acc);
                    while (!ok_b) ok_b = mbar_try_wait(b_full(st), par_b);
                    while (!ok_acc) ok_acc = mbar_try_wait(acc_empty(acc), par_acc);
                }
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t b_smem = base + offB + st * kBBytes;
                    const uint32_t d = tmem_base + acc * BN;
#pragma unroll
                    for (int sl = 0; sl < kSlabs; ++sl) {
                        const uint64_t adesc = umma_desc_sw128(a_smem + sl * (BM * 128));
                        const uint64_t bdesc = umma_desc_sw128(b_smem + sl * (BN * 128));
#pragma unroll
                        for (int k = 0; k < 4; ++k)        // K = 32 bytes per tcgen05.mma.kind::i8; +32 B = +2 encoded
                            umma_i8(d, adesc + 2 * k, bdesc + 2 * k, kIdesc, (sl | k) > 0);
                    }
                    umma_commit(b_empty(st));              // smem stage free once these MMAs retire
                    umma_commit(acc_full(acc));            // accumulator ready for the epilogue
                    if (t == last) umma_commit(a_empty(as));
                }
                __syncwarp();
            }
            tile0 += u.n_tiles;
            ++unit_iter;
        }
    } else if (warp >= kEpiWarp0) {
        // ================================================================ epilogue: running top-2 per query row
        const int group = (warp - kEpiWarp0) >> 2;         // column half of every tile
        const int quarter = warp & 3;                      // TMEM lanes 32*quarter .. +31
        const int row_in_unit = quarter * 32 + lane;
        int64_t* merge = reinterpret_cast<int64_t*>(base_ptr + offMerge);
        uint32_t tile_iter = 0;
        for (int64_t unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
            const UnitInfo u = decode_unit(pairs, unit_prefix, n_pairs, unit);
            int64_t r1 = kEmptyKey, r2 = kEmptyKey;
            for (int t = 0; t < u.n_tiles; ++t, ++tile_iter) {
                const int acc = tile_iter % kAccStages;
                const int cs = tile_iter % kCkSlots;
                mbar_wait(ck_full(cs), (tile_iter / kCkSlots) & 1);
                mbar_wait(acc_full(acc), (tile_iter / kAccStages) & 1);
                tc_fence_after();
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN + group * (BN / 2);
                const int4* ck4 = reinterpret_cast<const int4*>(base_ptr + offCk + cs * kCkBytes) + group * (BN / 8);
                // Per 32-column chunk: 32 IMAD build the keys, a VIMNMX3 tree (16 ops) takes the chunk minimum,
                // 3 more ops insert it into the running top-2 OF CHUNK MINIMA.  m1 is therefore the exact best
                // element; m2 is the best element outside m1's chunk, an upper bound of the true second
                // neighbour that refine_second_kernel (post.cu) tightens for the rows that need it.
                int32_t m1 = INT32_MAX, m2 = INT32_MAX;
                uint32_t v[2][32];
                tmem_ld_32x32(taddr, v[0]);
#pragma unroll
                for (int c = 0; c < BN / 64; ++c) {
                    uint32_t (&cur)[32] = v[c & 1];
                    // make the loaded registers depend on the wait
                    asm volatile("tcgen05.wait::ld.sync.aligned;"
                                 : "+r"(cur[0]), "+r"(cur[1]), "+r"(cur[2]), "+r"(cur[3]), "+r"(cur[4]), "+r"(cur[5]),
                                   "+r"(cur[6]), "+r"(cur[7]), "+r"(cur[8]), "+r"(cur[9]), "+r"(cur[10]), "+r"(cur[11]),
                                   "+r"(cur[12]), "+r"(cur[13]), "+r"(cur[14]), "+r"(cur[15]), "+r"(cur[16]),
                                   "+r"(cur[17]), "+r"(cur[18]), "+r"(cur[19]), "+r"(cur[20]), "+r"(cur[21]),
                                   "+r"(cur[22]), "+r"(cur[23]), "+r"(cur[24]), "+r"(cur[25]), "+r"(cur[26]),
                                   "+r"(cur[27]), "+r"(cur[28]), "+r"(cur[29]), "+r"(cur[30]), "+r"(cur[31])
                                 :: "memory");
                    if (c + 1 < BN / 64) tmem_ld_32x32(taddr + (c + 1) * 32, v[(c + 1) & 1]);
                    else {
                        // the last chunk is in registers (tcgen05.wait::ld is warp-wide): hand the stage back now
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(acc_empty(acc));
                    }
                    int32_t k[32];
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const int4 ck = ck4[c * 8 + (j >> 2)];           // smem broadcast
                        k[j] = ck.x - 512 * static_cast<int32_t>(cur[j]);
                        k[j + 1] = ck.y - 512 * static_cast<int32_t>(cur[j + 1]);
                        k[j + 2] = ck.z - 512 * static_cast<int32_t>(cur[j + 2]);
                        k[j + 3] = ck.w - 512 * static_cast<int32_t>(cur[j + 3]);
                    }
                    // balanced min3 tree: 32 -> 11 -> 4 -> 2 -> 1
                    int32_t a[11];
#pragma unroll
                    for (int i = 0; i < 10; ++i) a[i] = __vimin3_s32(k[3 * i], k[3 * i + 1], k[3 * i + 2]);
                    a[10] = min(k[30], k[31]);
                    const int32_t b0 = __vimin3_s32(a[0], a[1], a[2]), b1 = __vimin3_s32(a[3], a[4], a[5]);
                    const int32_t b2 = __vimin3_s32(a[6], a[7], a[8]), b3 = min(a[9], a[10]);
                    const int32_t cm = min(__vimin3_s32(b0, b1, b2), b3);
                    top2_key(cm, m1, m2);
                }
                const int32_t t1 = m1, t2 = m2;
                const int64_t colbase = static_cast<int64_t>(t) * BN;
                const int64_t k1 = static_cast<int64_t>(t1 >> 8) * (1ll << 32) + (colbase + (t1 & 255));
                const int64_t k2 = static_cast<int64_t>(t2 >> 8) * (1ll << 32) + (colbase + (t2 & 255));
                top2_key64(k1, r1, r2);
                top2_key64(k2, r1, r2);
            }
            // ---- unit end: combine the two groups, write the raw knn result of this query row
            if (group == 1) { merge[row_in_unit * 2] = r1; merge[row_in_unit * 2 + 1] = r2; }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (group == 0) {
                top2_key64(merge[row_in_unit * 2], r1, r2);
                top2_key64(merge[row_in_unit * 2 + 1], r1, r2);
                const int row = u.rb * BM + row_in_unit;
                if (row < u.pd.nq) {
                    const int32_t na = norm2[u.pd.q_row0 + row];
                    const int32_t v1 = static_cast<int32_t>(r1 >> 32), v2 = static_cast<int32_t>(r2 >> 32);
                    const bool has1 = r1 != kEmptyKey && v1 < (kSentinelKey >> 8);
                    const bool has2 = r2 != kEmptyKey && v2 < (kSentinelKey >> 8);
                    Top2 o;
                    o.i0 = has1 ? static_cast<int32_t>(r1 & 0xFFFFFFFF) : -1;
                    o.i1 = has2 ? static_cast<int32_t>(r2 & 0xFFFFFFFF) : -1;
                    o.d0 = has1 ? static_cast<float>(na + v1) : __int_as_float(0x7f800000);
                    o.d1 = has2 ? static_cast<float>(na + v2) : __int_as_float(0x7f800000);
                    out[u.pd.out_row0 + row] = o;
                }
            }
            asm volatile("bar.sync 2, 256;" ::: "memory");
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

template <int kSlabs>
static cudaError_t launch_tc(const CUtensorMap& ta, const CUtensorMap& tb, const int32_t* ckey, const int32_t* norm2,
                             const PairDesc* pairs, const int64_t* unit_prefix, int n_pairs, int64_t n_units, Top2* out,
                             int grid, cudaStream_t s) {
    // per launch: the attribute is per device, and one process may drive several GPUs
    cudaError_t e = cudaFuncSetAttribute(knn2_l2_u8_tc_kernel<kSlabs>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             tc::Cfg<kSlabs>::kSmemBytes);
    if (e != cudaSuccess) return e;
    knn2_l2_u8_tc_kernel<kSlabs><<<grid, tc::kThreads, tc::Cfg<kSlabs>::kSmemBytes, s>>>(ta, tb, ckey, norm2, pairs, unit_prefix,
                                                                                        n_pairs, n_units, out);
    return cudaGetLastError();
}

// slabs: 1 = 128-byte rows (SIFT), 2 = 256-byte rows (ORB bits expanded to bytes)
cudaError_t launch_knn2_l2_u8_tc(const void* tmap_a_host, const void* tmap_b_host, const int32_t* ckey,
                                 const int32_t* norm2, const PairDesc* pairs, const int64_t* unit_prefix,
                                 int n_pairs, int64_t n_units, Top2* out, int sm_count, int slabs, cudaStream_t s) {
    if (n_units == 0) return cudaSuccess;
    const CUtensorMap* ta = static_cast<const CUtensorMap*>(tmap_a_host);
    const CUtensorMap* tb = static_cast<const CUtensorMap*>(tmap_b_host);
    const int grid = static_cast<int>(n_units < sm_count ? n_units : sm_count);
    if (slabs == 2) return launch_tc<2>(*ta, *tb, ckey, norm2, pairs, unit_prefix, n_pairs, n_units, out, grid, s);
    return launch_tc<1>(*ta, *tb, ckey, norm2, pairs, unit_prefix, n_pairs, n_units, out, grid, s);
}

}  // namespace sfm
