// knn_l2_tcv.cu — "value-only" variant of the tcgen05 SIFT kernel: the throughput path of sfm_match_pairs.
//
// Same contraction, pipeline and warp roles as knn_l2_tc.cu (TMA -> smem ring -> tcgen05.mma.kind::i8 -> TMEM ->
// tcgen05.ld epilogue), but the per-train-row constant is folded into the MMA instead of the epilogue: every bank
// row carries 32 signed norm digits e (see common.cuh) and a fifth K-step (A = constant weight rows, u8;
// B = digits, s8) adds -(|b|^2 >> 1), so the accumulator holds
//        D = a.b - (|b_j|^2 >> 1)          and        |a - b_j|^2 = |a|^2 - 2 D + (|b_j|^2 & 1).
// A larger D is a closer row, strictly (the parity bit only orders rows with equal D).  The epilogue therefore
// needs no per-column constant at all: per 32-column chunk a VIMNMX3 tree takes the maximum of the raw accumulators
// (1/2 ALU op per element, no IMAD, no shared-memory loads) and the running state is the top-4 of CHUNK maxima,
// kept as (D, chunk).  The two nearest rows always lie in chunks whose maximum is >= the second-best chunk maximum,
// so the output per query row is: the two best chunks and their D values, a third chunk if it ties the second, and
// an "ambiguous" flag if a fourth ties too (probability ~1e-7 per row: brute force in the refine pass).
// refine_value_kernel (post.cu) turns this into the exact cv::batchDistance answer: rows whose ratio test cannot
// pass even with the bounds  d0^2 >= |a|^2 - 2 D1,  d1^2 <= |a|^2 - 2 D2 + 1  are rejected outright (~99.7 % of C3's
// rows); for the rest the two chunks (64 train rows) are recomputed exactly with __dp4a, ambiguous rows by brute
// force over the whole train image.  Everything is integer arithmetic -> bit-exact (tests/test_gpu_parity.py).
//
// Precondition (checked on the host): every |b|^2 of the bank <= kExtMaxNorm2.  Otherwise knn_l2_tc.cu is used.
#include <cuda.h>

#include <climits>

#include "common.cuh"
#include "kernels.h"
#include "refine_dot.cuh"

namespace sfm {

namespace tcv {
constexpr int BM = 128, KB = 128;
constexpr int kAStages = 2;
constexpr int kABytes = BM * KB, kAExtBytes = BM * kExtBytes;
constexpr int kEpiWarp0 = 4;
constexpr int32_t kValueBias = kExtPadValue;                      // D + bias >= 0, < 2^22
constexpr int kSeqBits = 9;                                       // chunks one epilogue warp visits per unit <= 512
constexpr int64_t kEmpty = INT64_MIN;

// Tile width BN = train rows per MMA tile; the 512 TMEM columns hold 512 / BN accumulator stages.  BN = 256 is the one
// in use: tools/microbench2.cu measured tcgen05.mma.kind::i8 M128 x N256 x K32 at exactly 128 cycles (8192 MAC/clk/SM)
// and N128 at 64, but every MMA costs its ISSUING thread 50-100 cycles (descriptor moves to uniform registers, the elect
// loop, the instruction), so 128-row tiles (twice the instructions per train row) were 1.6x slower (profiles/
// r1_tcv_issue_analysis.txt).
template <int BN_>
struct Cfg {
    static constexpr int BN = BN_;
    static constexpr int kAccStages = 512 / BN;
    static constexpr int kBStages = 1024 / BN;                     // 128 KB of train rows in flight either way
    static constexpr int kBBytes = BN * KB, kEBytes = BN * kExtBytes;
    static constexpr int offB = 0;
    static constexpr int offE = offB + kBStages * kBBytes;         // digit tiles, one per B stage
    static constexpr int offA = offE + kBStages * kEBytes;
    static constexpr int offAExt = offA + kAStages * kABytes;      // constant weight rows
    static constexpr int offMerge = offAExt + kAExtBytes;          // [3 groups][128][5] int64
    static constexpr int offBar = offMerge + 3 * BM * 5 * 8;
    static constexpr int kNumBars = 2 * kBStages + 2 * kAStages + 2 * kAccStages + 2 * 4;   // + unit hand-off to the finishing warps (kFinDepth each way)
    static constexpr int offTmemPtr = offBar + kNumBars * 8;
    static constexpr int kSmemBytes = offTmemPtr + 16 + 1024;
    static constexpr uint32_t kIdesc = umma_idesc_u8(BM, BN);
    static constexpr uint32_t kIdescExt = umma_idesc_u8s8(BM, BN);
};
// In-kernel unit end and re-rank (kRefine): the epilogue warps only stream accumulators.  At the end of a unit each of them
// leaves its five running keys per query row in shared memory (a ring of kFinDepth buffers) and goes on with the next
// unit; the four FINISHING warps merge the two column halves, apply the fused ratio bound and re-rank the few surviving rows
// exactly (refine_dot_row) while the epilogue is already in the next unit.  Measured before this split: ~2460 cycles per
// unit with both epilogue groups parked at two bar.sync around the unit end of group 0 (C3: 32 tiles per unit -> 77 of
// 620 cycles per tile; C5's 64-tile units ran 6.6 % faster per tile for that reason alone).
constexpr int kFinBufWords = 2 * 5 * BM;       // [column half][key][row] uint32 per buffer
struct __align__(16) SurvRec { Top2 t; int32_t v5, na, row, pad; };         // a row that survived the ratio bound (row: inside the query image)
struct SurvList { SurvRec rec[BM]; int count; };
constexpr int kFinDepth = 4;                   // buffers (units) the finishing warps may lag behind the epilogue; in the unused digit-tile area
}  // namespace tcv

struct UnitInfoV { PairDesc pd; int rb; int n_tiles; };

// uniform_units > 0: every pair of the launch has that many units (all images of the same size: the BASELINE configurations), so the
// pair of a unit is a division instead of a binary search over the prefix array — twelve dependent loads at the start of every
// unit in every warp of the CTA otherwise
template <int BN>
__device__ __forceinline__ UnitInfoV decode_unit_v(const PairDesc* __restrict__ pairs, const int64_t* __restrict__ unit_prefix,
                                                   int n_pairs, int64_t unit, int uniform_units, int& p_hint) {
    UnitInfoV u;
    if (uniform_units > 0) {
        const int p = static_cast<int>(unit / uniform_units);
        u.pd = pairs[p];
        u.rb = static_cast<int>(unit - static_cast<int64_t>(p) * uniform_units);
        u.n_tiles = (u.pd.nt + BN - 1) / BN;
        return u;
    }
    // ragged scene: a CTA visits its units in ascending order, so the pair of the next unit lies at or shortly after the pair of
    // the previous one: gallop from there (1, 2, 4, ... pairs ahead), then bisect the bracket — ~2 log2(distance) dependent loads
    // instead of log2(n_pairs)
    int lo = p_hint, step = 1;
    int hi = lo + 1;
    while (hi < n_pairs && __ldg(unit_prefix + hi) <= unit) { lo = hi; hi = min(n_pairs, hi + step); step <<= 1; }
    // invariant: prefix[lo] <= unit < prefix[hi]  (prefix[n_pairs] = total units > unit)
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(unit_prefix + mid) <= unit) lo = mid; else hi = mid;
    }
    const int p = lo;
    p_hint = p;
    u.pd = pairs[p];
    u.rb = static_cast<int>(unit - unit_prefix[p]);
    u.n_tiles = (u.pd.nt + BN - 1) / BN;
    return u;
}

// running top-3 (descending) insert
__device__ __forceinline__ void top3_max(int32_t k, int32_t& m1, int32_t& m2, int32_t& m3) {
    const int32_t t = min(m1, k);
    m1 = max(m1, k);
    const int32_t u = min(m2, t);
    m2 = max(m2, t);
    m3 = max(m3, u);
}
// running top-4 (descending) insert, 7 ops
__device__ __forceinline__ void top4_max(int32_t k, int32_t& m1, int32_t& m2, int32_t& m3, int32_t& m4) {
    const int32_t t = min(m1, k);
    m1 = max(m1, k);
    const int32_t u = min(m2, t);
    m2 = max(m2, t);
    const int32_t w = min(m3, u);
    m3 = max(m3, u);
    m4 = max(m4, w);
}
// running top-5 (descending) insert, 9 ops
__device__ __forceinline__ void top5_max(int32_t k, int32_t& m1, int32_t& m2, int32_t& m3, int32_t& m4, int32_t& m5) {
    const int32_t t = min(m1, k);
    m1 = max(m1, k);
    const int32_t u = min(m2, t);
    m2 = max(m2, t);
    const int32_t w = min(m3, u);
    m3 = max(m3, u);
    const int32_t x = min(m4, w);
    m4 = max(m4, w);
    m5 = max(m5, x);
}
__device__ __forceinline__ void top5_maxu(uint32_t k, uint32_t& m1, uint32_t& m2, uint32_t& m3, uint32_t& m4, uint32_t& m5) {
    const uint32_t t = min(m1, k);
    m1 = max(m1, k);
    const uint32_t u = min(m2, t);
    m2 = max(m2, t);
    const uint32_t w = min(m3, u);
    m3 = max(m3, u);
    const uint32_t x = min(m4, w);
    m4 = max(m4, w);
    m5 = max(m5, x);
}
__device__ __forceinline__ void top5_max64(int64_t k, int64_t& m1, int64_t& m2, int64_t& m3, int64_t& m4, int64_t& m5) {
    const int64_t t = min(m1, k);
    m1 = max(m1, k);
    const int64_t u = min(m2, t);
    m2 = max(m2, t);
    const int64_t w = min(m3, u);
    m3 = max(m3, u);
    const int64_t x = min(m4, w);
    m4 = max(m4, w);
    m5 = max(m5, x);
}
__device__ __forceinline__ void top4_max64(int64_t k, int64_t& m1, int64_t& m2, int64_t& m3, int64_t& m4) {
    const int64_t t = min(m1, k);
    m1 = max(m1, k);
    const int64_t u = min(m2, t);
    m2 = max(m2, t);
    const int64_t w = min(m3, u);
    m3 = max(m3, u);
    m4 = max(m4, w);
}
__device__ __forceinline__ void top3_max64(int64_t k, int64_t& m1, int64_t& m2, int64_t& m3) {
    const int64_t t = min(m1, k);
    m1 = max(m1, k);
    const int64_t u = min(m2, t);
    m2 = max(m2, t);
    m3 = max(m3, u);
}

// kParity x kHalves epilogue groups of 4 warps: a group owns the tiles of one parity (kParity = 2) or every tile
// (kParity = 1), and one of kHalves column ranges of the tile.  With two accumulator stages the MMA of tile t+2 waits for
// the complete epilogue of tile t, so the latency of ONE tile's epilogue has to stay below the MMA time of a tile:
// splitting the columns of every tile over the groups (1 x 2) halves that latency, alternating tiles (2 x 1) does not.
//
// kNorm = true : the accumulator carries the norm term (fifth K-step), output = best chunks by D = ab - (|b|^2 >> 1).
// kNorm = false: "norm-less" variant for banks whose |b|^2 vary little (SIFT: ~1 %): four K-steps only (-20 % tensor
//                work and energy, no digit tiles), the epilogue ranks chunks by the raw dot product a.b and reports the
//                four best chunks, V1, V2 and the fifth-best chunk maximum V5; refine_dot_kernel (post.cu) turns that
//                into the exact answer with the train image's norm range: d^2 >= |a|^2 + min|b|^2 - 2 V for every row
//                whose chunk maximum is V (see there).  Same exactness guarantee, everything integer.
// kCC = columns (train rows) per chunk: 32 or 64.  The epilogue is ALU-bound (ncu: ALU pipe 67 %, 0.85 instructions per
// accumulator element with 32-column chunks); the per-chunk bookkeeping (key, top-4/5 insert) halves with 64-column chunks,
// the refine pass then recomputes 64 train rows per candidate chunk.
template <int kParity, int kHalves, int kBN, bool kNorm, int kCC, bool kRefine = false>
__global__ void __launch_bounds__(128 + 128 * kParity * kHalves + (kRefine ? 128 : 0), 1)
knn2_l2_u8_tcv_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                      const __grid_constant__ CUtensorMap tmap_e, const PairDesc* __restrict__ pairs,
                      const int64_t* __restrict__ unit_prefix, int n_pairs, int64_t n_units, Top2* __restrict__ out,
                      int32_t* __restrict__ aux, int n_issuers, const TcvFuse fz) {
    using namespace tcv;
    using C = Cfg<kBN>;
    constexpr int BN = C::BN, kAccStages = C::kAccStages, kBStages = C::kBStages, kBBytes = C::kBBytes, kEBytes = C::kEBytes;
    constexpr int offB = C::offB, offE = C::offE, offA = C::offA, offAExt = C::offAExt, offMerge = C::offMerge;
    constexpr int offBar = C::offBar, offTmemPtr = C::offTmemPtr;
    constexpr uint32_t kIdesc = C::kIdesc, kIdescExt = C::kIdescExt;
    constexpr int kGroups = kParity * kHalves;
    constexpr int kThreads = 128 + 128 * kGroups + (kRefine ? 128 : 0);
    constexpr bool kKey32 = !kNorm && kParity == 1;                 // 32-bit keys with global chunk numbers (see the epilogue)
    constexpr int kGchBits = 11;
    constexpr uint32_t kGchMask = (1u << kGchBits) - 1;
    static_assert(!kRefine || (!kNorm && kGroups == 2), "refine warps: norm-less variant with 8 epilogue warps (512 threads x 128 registers)");
    static_assert(!kRefine || kFinDepth * kFinBufWords * 4 + static_cast<int>(sizeof(SurvList)) <= C::kBStages * C::kEBytes,
                  "hand-off buffers and the survivor list fit the unused digit-tile area");
    constexpr int kLoadsPerVisit = BN / 32 / kHalves;               // 32-column tcgen05.ld one warp issues per tile
    constexpr int kChunksPerVisit = BN / kCC / kHalves;             // chunks one warp reads per tile
    static_assert(kCC == 32 || kCC == 64 || kCC == 128, "chunk = one, two or four 32-column loads");
    static_assert(kCC != 128 || (!kNorm && kParity == 1), "128-row chunks: norm-less variant with global chunk keys only");
    constexpr int kLoadsPerChunk = kCC / 32;
    static_assert(kChunksPerVisit >= 1, "a warp reads at least one whole chunk per tile");
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
    auto fin_full = [&](int i) { return bar0 + 8u * (2 * kBStages + 2 * kAStages + 2 * kAccStages + i); };
    auto fin_empty = [&](int i) { return bar0 + 8u * (2 * kBStages + 2 * kAStages + 2 * kAccStages + kFinDepth + i); };
    volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(base_ptr + offTmemPtr);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_a);
        prefetch_tmap(&tmap_b);
        prefetch_tmap(&tmap_e);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kBStages; ++i) { mbar_init(b_full(i), 1); mbar_init(b_empty(i), 1); }
        for (int i = 0; i < kAStages; ++i) { mbar_init(a_full(i), 1); mbar_init(a_empty(i), n_issuers); }   // every MMA warp releases A
        for (int i = 0; i < kAccStages; ++i) { mbar_init(acc_full(i), 1); mbar_init(acc_empty(i), 4 * kHalves); }   // one arrival per reading warp
        for (int i = 0; i < kFinDepth; ++i) { mbar_init(fin_full(i), 4 * kGroups); mbar_init(fin_empty(i), 4); }   // per epilogue warp / per finishing warp
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(base + offTmemPtr, 512);
        tmem_relinquish();
    }
    // constant weight operand: 128 identical rows {255 x12, 1 x4, 255 x12, 1 x4}; both 16-byte halves of a row are
    // equal, so the 32-byte swizzle (which only swaps halves) leaves the image unchanged
    for (int i = threadIdx.x; i < kAExtBytes / 4; i += kThreads) {
        const int word = i & 3;                                     // word inside a 16-byte half
        reinterpret_cast<uint32_t*>(base_ptr + offAExt)[i] = word < 3 ? 0xFFFFFFFFu : 0x01010101u;
    }
    fence_proxy_async();                                            // generic-proxy writes -> visible to the MMA
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        // ================================================================ TMA producer
        uint32_t tile_iter = 0, unit_iter = 0;
        int p_hint = 0;
        // the descriptor of the NEXT unit is requested at the start of the current one: its global loads finish under the tile loop
        UnitInfoV nxt = decode_unit_v<BN>(pairs, unit_prefix, n_pairs, blockIdx.x, fz.uniform_units, p_hint);
        for (int64_t unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
            const UnitInfoV u = nxt;
            if (unit + gridDim.x < n_units) nxt = decode_unit_v<BN>(pairs, unit_prefix, n_pairs, unit + gridDim.x, fz.uniform_units, p_hint);
            if (u.n_tiles == 0) continue;
            const int as = unit_iter % kAStages;
            mbar_wait(a_empty(as), ((unit_iter / kAStages) & 1) ^ 1);
            if (elect_one()) {
                mbar_arrive_expect_tx(a_full(as), kABytes);
                tma_load_2d(base + offA + as * kABytes, &tmap_a, 0, u.pd.q_row0 + u.rb * BM, a_full(as));
            }
            for (int t = 0; t < u.n_tiles; ++t, ++tile_iter) {
                const int st = tile_iter % kBStages;
                mbar_wait(b_empty(st), ((tile_iter / kBStages) & 1) ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(b_full(st), kBBytes + (kNorm ? kEBytes : 0));
                    tma_load_2d(base + offB + st * kBBytes, &tmap_b, 0, u.pd.t_row0 + t * BN, b_full(st));
                    if (kNorm) tma_load_2d(base + offE + st * kEBytes, &tmap_e, 0, u.pd.t_row0 + t * BN, b_full(st));
                }
                __syncwarp();
            }
            ++unit_iter;
        }
    } else if (warp == 1 || (warp == 3 && n_issuers == 2)) {
        // ================================================================ MMA issuers (one elected lane issues)
        // The issue side is a serial chain per tile: two mbarrier polls (~60 cycles each even when satisfied), 5 tcgen05.mma
        // and 2-3 tcgen05.commit at >= 48 cycles per instruction.  Guarded by `lane == 0`, ptxas wrapped every one of them
        // in an elect / R2UR.BROADCAST / BRA.U.ANY loop (~75 cycles each; measured 770 cycles of issue work per tile against
        // 640 cycles of tensor-pipe work); guarded by elect.sync they are straight-line UTCIMMA / UTCBAR.  Two warps issue
        // (n_issuers = 2): warp 1 the even tiles (accumulator stage 0), warp 3 the odd tiles (stage 1), so the polls of
        // one tile overlap the issue of the other.  Each warp's tcgen05.commit covers its own MMAs, hence every issuing
        // warp arrives on a_empty after its last tile of a unit.
        const uint32_t my_par = static_cast<uint32_t>(warp >> 1);
        uint32_t tile0 = 0, unit_iter = 0;                          // tile0: running tile number at the start of the unit
        const uint64_t aext = umma_desc_sw32(base + offAExt);
        int p_hint = 0;
        // the descriptor of the NEXT unit is requested at the start of the current one: its global loads finish under the tile loop
        UnitInfoV nxt = decode_unit_v<BN>(pairs, unit_prefix, n_pairs, blockIdx.x, fz.uniform_units, p_hint);
        for (int64_t unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
            const UnitInfoV u = nxt;
            if (unit + gridDim.x < n_units) nxt = decode_unit_v<BN>(pairs, unit_prefix, n_pairs, unit + gridDim.x, fz.uniform_units, p_hint);
            if (u.n_tiles == 0) continue;
            const int as = unit_iter % kAStages;
            mbar_wait(a_full(as), (unit_iter / kAStages) & 1);
            const uint64_t adesc = umma_desc_sw128(base + offA + as * kABytes);
            const int first = n_issuers == 2 ? static_cast<int>((my_par - tile0) & 1u) : 0;   // my first tile of this unit
            const int last = first < u.n_tiles ? first + n_issuers * ((u.n_tiles - 1 - first) / n_issuers) : -1;
            if (last < 0 && lane == 0) mbar_arrive(a_empty(as));                       // no tile of this unit is mine
            for (int t = first; t < u.n_tiles; t += n_issuers) {
                const uint32_t tile_iter = tile0 + t;
                const int acc = tile_iter % kAccStages;
                const int st = tile_iter % kBStages;
                {
                    // both polls in flight at once (a satisfied try_wait still costs ~60 cycles); the train tile has
                    // normally landed long before the epilogue hands the accumulator stage back
                    const uint32_t par_b = (tile_iter / kBStages) & 1, par_acc = ((tile_iter / kAccStages) & 1) ^ 1;
                    bool ok_b = mbar_try_wait(b_full(st), par_b);
                    bool ok_acc = mbar_try_wait(acc_empty(acc), par_acc);
                    while (!ok_b) ok_b = mbar_try_wait(b_full(st), par_b);
                    while (!ok_acc) ok_acc = mbar_try_wait(acc_empty(acc), par_acc);
                }
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t bdesc = umma_desc_sw128(base + offB + st * kBBytes);
                    const uint64_t edesc = umma_desc_sw32(base + offE + st * kEBytes);
                    const uint32_t d = tmem_base + acc * BN;
#pragma unroll
                    for (int k = 0; k < KB / 32; ++k)
                        umma_i8(d, adesc + 2 * k, bdesc + 2 * k, kIdesc, k > 0);
                    if (kNorm) umma_i8(d, aext, edesc, kIdescExt, 1);          // += -(|b|^2 >> 1)
                    // two commits (B stage -> producer, accumulator -> epilogue): one shared mbarrier for both waiters
                    // measured 4 % slower
                    umma_commit(b_empty(st));
                    umma_commit(acc_full(acc));
                    if (t == last) umma_commit(a_empty(as));
                }
                __syncwarp();
            }
            tile0 += u.n_tiles;
            ++unit_iter;
        }
    } else if (warp >= kEpiWarp0 && warp < kEpiWarp0 + 4 * kGroups) {
        // ================================================================ epilogue: top-4 of chunk maxima per query row
        // 16 warps = 4 groups; group g works on the tiles of parity (g & 1) and on the column half (g >> 1), so four
        // tcgen05.ld are in flight per SM sub-partition (the loads are latency- not bandwidth-limited).
        // Running state for the WHOLE unit in 32-bit keys:  key = (D + bias) << 9 | (511 - seq), seq = running number
        // of the chunk in this warp's visiting order (ascending train rows) -> equal D keeps the earlier chunk.
        const int group = (warp - kEpiWarp0) >> 2;
        const int parity = group % kParity, half = group / kParity;
        const int quarter = warp & 3;
        const int row_in_unit = quarter * 32 + lane;
        int64_t* merge = reinterpret_cast<int64_t*>(base_ptr + offMerge);
        uint32_t tile_iter = 0, unit_no = 0;
        int p_hint = 0;
        // the descriptor of the NEXT unit is requested at the start of the current one: its global loads finish under the tile loop
        UnitInfoV nxt = decode_unit_v<BN>(pairs, unit_prefix, n_pairs, blockIdx.x, fz.uniform_units, p_hint);
        for (int64_t unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
            const UnitInfoV u = nxt;
            if (unit + gridDim.x < n_units) nxt = decode_unit_v<BN>(pairs, unit_prefix, n_pairs, unit + gridDim.x, fz.uniform_units, p_hint);
            int32_t m1 = -1, m2 = -1, m3 = -1, m4 = -1, m5 = -1;        // m5: fifth-best chunk key (kNorm = false only)
            // kKey32 (norm-less variant, every group sees every tile): the running keys carry the GLOBAL chunk number,
            //     key = a.b << 11 | (2047 - chunk),   a.b < 2^21 (|b|^2 <= kExtMaxNorm2), chunk < 2048,
            // so the groups' top-5 lists merge at unit end with 32-bit min / max only (the 64-bit merge of the other variants costs
            // ~6 % of the epilogue's instructions); 0 = no chunk
            uint32_t n1 = 0, n2 = 0, n3 = 0, n4 = 0, n5 = 0;
            uint32_t gbase = kGchMask - static_cast<uint32_t>(half * kChunksPerVisit);      // - (BN / kCC) per tile
            const int t_first = kParity == 2 ? static_cast<int>((parity - tile_iter) & 1u) : 0;   // first tile of this unit we own
            int seq = 0;
            // what the ratio bound at unit end needs from global memory is requested NOW (group 0 owns the unit end): |a|^2 of the
            // row and this lane's share of the train image's |b|^2 block ranges; the loads complete under the tile loop
            int na_pre = 0, mn_pre = INT_MAX, mx_pre = 0;
            if (!kRefine && group == 0) {
                const int prow = u.rb * BM + row_in_unit;
                if (prow < u.pd.nq) na_pre = __ldg(fz.norm2 + u.pd.q_row0 + prow);
                if (!kNorm) {
                    const int b0 = u.pd.t_row0 / kRowAlign, nblk = (u.pd.nt + kRowAlign - 1) / kRowAlign;
                    for (int b = lane; b < nblk; b += 32) { mn_pre = min(mn_pre, __ldg(fz.blk_min + b0 + b)); mx_pre = max(mx_pre, __ldg(fz.blk_max + b0 + b)); }
                }
            }
            for (int t = 0; t < u.n_tiles; ++t, ++tile_iter) {
                if (kParity == 2 && (tile_iter & 1) != static_cast<uint32_t>(parity)) continue;
                const int acc = tile_iter % kAccStages;
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN + half * (BN / kHalves);
                mbar_wait(acc_full(acc), (tile_iter / kAccStages) & 1);
                tc_fence_after();
                uint32_t v[2][32];
                tmem_ld_32x32(taddr, v[0]);
                int32_t cm_even = 0;
#pragma unroll
                for (int c = 0; c < kLoadsPerVisit; ++c) {
                    uint32_t (&cur)[32] = v[c & 1];
                    asm volatile("tcgen05.wait::ld.sync.aligned;"
                                 : "+r"(cur[0]), "+r"(cur[1]), "+r"(cur[2]), "+r"(cur[3]), "+r"(cur[4]), "+r"(cur[5]),
                                   "+r"(cur[6]), "+r"(cur[7]), "+r"(cur[8]), "+r"(cur[9]), "+r"(cur[10]), "+r"(cur[11]),
                                   "+r"(cur[12]), "+r"(cur[13]), "+r"(cur[14]), "+r"(cur[15]), "+r"(cur[16]),
                                   "+r"(cur[17]), "+r"(cur[18]), "+r"(cur[19]), "+r"(cur[20]), "+r"(cur[21]),
                                   "+r"(cur[22]), "+r"(cur[23]), "+r"(cur[24]), "+r"(cur[25]), "+r"(cur[26]),
                                   "+r"(cur[27]), "+r"(cur[28]), "+r"(cur[29]), "+r"(cur[30]), "+r"(cur[31])
                                 :: "memory");
                    if (c + 1 < kLoadsPerVisit) tmem_ld_32x32(taddr + (c + 1) * 32, v[(c + 1) & 1]);
                    else {
                        // last chunk is in registers (tcgen05.wait::ld is warp-wide): hand the stage back now, ONE arrival
                        // per warp (128 per-thread arrivals on one mbarrier serialise in shared memory)
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(acc_empty(acc));
                    }
                    // balanced max3 tree over the raw accumulators: 32 -> 11 -> 4 -> 1
                    int32_t a[11];
#pragma unroll
                    for (int i = 0; i < 10; ++i)
                        a[i] = __vimax3_s32(static_cast<int32_t>(cur[3 * i]), static_cast<int32_t>(cur[3 * i + 1]),
                                            static_cast<int32_t>(cur[3 * i + 2]));
                    a[10] = max(static_cast<int32_t>(cur[30]), static_cast<int32_t>(cur[31]));
                    const int32_t b0 = __vimax3_s32(a[0], a[1], a[2]), b1 = __vimax3_s32(a[3], a[4], a[5]);
                    const int32_t b2 = __vimax3_s32(a[6], a[7], a[8]), b3 = max(a[9], a[10]);
                    int32_t cmax = max(__vimax3_s32(b0, b1, b2), b3);
                    if (kLoadsPerChunk > 1) {                                          // a chunk spans several 32-column loads
                        if (c % kLoadsPerChunk != 0) cmax = max(cmax, cm_even);
                        if (c % kLoadsPerChunk != kLoadsPerChunk - 1) { cm_even = cmax; continue; }
                    }
                    const int cs = seq + c / kLoadsPerChunk;
                    if (kNorm) top4_max(cmax * (1 << kSeqBits) + (kValueBias * (1 << kSeqBits) + 511 - cs), m1, m2, m3, m4);
                    else if (kKey32) top5_maxu(static_cast<uint32_t>(cmax) * (1u << kGchBits) + (gbase - static_cast<uint32_t>(c / kLoadsPerChunk)), n1, n2, n3, n4, n5);
                    else top5_max(cmax * (1 << kSeqBits) + (511 - cs), m1, m2, m3, m4, m5);   // 0 <= a.b < 2^22
                }
                seq += kChunksPerVisit;
                gbase -= BN / kCC;
            }
            if constexpr (kRefine) {
                // ---- unit end, handed over: five keys per row into the buffer of this unit's parity, one arrival per warp
                const uint32_t fpar = unit_no % kFinDepth;
                mbar_wait(fin_empty(fpar), ((unit_no / kFinDepth) & 1u) ^ 1u);
                uint32_t* fin = reinterpret_cast<uint32_t*>(base_ptr + offE) + fpar * kFinBufWords + group * 5 * BM + row_in_unit;
                fin[0 * BM] = n1; fin[1 * BM] = n2; fin[2 * BM] = n3; fin[3 * BM] = n4; fin[4 * BM] = n5;
                __syncwarp();
                if (lane == 0) mbar_arrive(fin_full(fpar));         // release: the finishing warps acquire on the barrier
                ++unit_no;
                continue;
            }
            // ---- unit end: to (D + bias, -global chunk) 64-bit keys, merge the groups, write the candidates
            constexpr int kKeys = kNorm ? 4 : 5;                    // the fifth key only carries a value
            int64_t r[5];
            uint32_t r32[5] = {n1, n2, n3, n4, n5};
            if (!kKey32) {
                const int32_t mk[5] = {m1, m2, m3, m4, m5};
#pragma unroll
                for (int i = 0; i < kKeys; ++i) {
                    if (mk[i] < 0) { r[i] = kEmpty; continue; }
                    const int sq = 511 - (mk[i] & 511);
                    const int gch = (t_first + kParity * (sq / kChunksPerVisit)) * (BN / kCC) + half * kChunksPerVisit + (sq % kChunksPerVisit);
                    r[i] = static_cast<int64_t>(mk[i] >> kSeqBits) * (1ll << 32) + (0x7FFFFFFF - gch);
                }
            }
            uint32_t* merge32 = reinterpret_cast<uint32_t*>(merge);           // kKey32: [groups - 1][5][128] (conflict-free columns)
            if (group != 0) {
#pragma unroll
                for (int i = 0; i < kKeys; ++i) {
                    if (kKey32) merge32[((group - 1) * 5 + i) * BM + row_in_unit] = r32[i];
                    else merge[((group - 1) * BM + row_in_unit) * 5 + i] = r[i];
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(128 * kGroups) : "memory");
            if (group == 0) {
                for (int g = 0; g < kGroups - 1; ++g)
#pragma unroll
                    for (int i = 0; i < kKeys; ++i) {
                        if (kKey32) { top5_maxu(merge32[(g * 5 + i) * BM + row_in_unit], r32[0], r32[1], r32[2], r32[3], r32[4]); continue; }
                        const int64_t k = merge[(g * BM + row_in_unit) * 5 + i];
                        if (kNorm) top4_max64(k, r[0], r[1], r[2], r[3]);
                        else top5_max64(k, r[0], r[1], r[2], r[3], r[4]);
                    }
                // ---- fused ratio-test bound (north_star: "fused epilogue ... ratio test"): a row can only pass
                //     sqrtf(d0^2) < ratio * sqrtf(d1^2)   if it passes with the smallest possible d0^2 and the largest possible
                // d1^2 the chunk maxima allow.  kNorm:  d0^2 >= |a|^2 - 2 D1,  d1^2 <= |a|^2 - 2 D2 + 1.  Norm-less:
                // d0^2 >= |a|^2 + N- - 2 V1,  d1^2 <= |a|^2 + N+ - 2 V2  with [N-, N+] the |b|^2 range of the train image
                // (reduced here from the per-256-row-block ranges).  ~99.5 % of C3's rows stop here and write NOTHING;
                // survivors write their candidate record and append their staging row to the need list (one atomic per warp)
                // for the exact re-rank (post.cu: refine_dot_rows_kernel / refine_value_rows_kernel).
                const int row = u.rb * BM + row_in_unit;
                const bool in_range = row < u.pd.nq && u.pd.nt > 0;
                int32_t vv[5], ch[5];
                bool has[5];
#pragma unroll
                for (int i = 0; i < kKeys; ++i) {
                    if (kKey32) {
                        vv[i] = static_cast<int32_t>(r32[i] >> kGchBits);
                        ch[i] = static_cast<int32_t>(kGchMask - (r32[i] & kGchMask));
                        has[i] = vv[i] > 0;
                        continue;
                    }
                    vv[i] = static_cast<int32_t>(r[i] >> 32);
                    ch[i] = 0x7FFFFFFF - static_cast<int32_t>(r[i] & 0xFFFFFFFF);
                    // value part 0: kNorm: D == -bias, a chunk of padding rows only; norm-less: a.b == 0, padding rows or
                    // rows orthogonal to the query (the refine pass treats everything outside the candidates as a.b <= V5)
                    has[i] = r[i] != kEmpty && vv[i] > 0;
                }
                int nbmin = 0, nbmax = 0;
                if (!kNorm) {
                    int mn = mn_pre, mx = mx_pre;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) { mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
                    nbmin = mn; nbmax = mx;
                }
                bool need = false;
                Top2 o;
                o.i0 = -1; o.i1 = -1; o.d0 = 0.f; o.d1 = 0.f;
                if (in_range) {
                    const int na = na_pre;
                    if (kNorm) {
                        o.i0 = has[0] ? ch[0] : -1;                                          // chunk of the best D
                        o.i1 = has[1] ? ch[1] : -1;                                          // second chunk
                        if (has[1] && has[2] && vv[2] == vv[1]) {
                            o.i0 |= (ch[2] + 1) << 16;                                       // a third chunk ties the second
                            if (has[3] && vv[3] == vv[1]) o.i1 |= 0x40000000;                // and a fourth: ambiguous
                        }
                        const int D1 = vv[0] - kValueBias, D2 = vv[1] - kValueBias;
                        o.d0 = __int_as_float(D1);
                        o.d1 = __int_as_float(D2);
                        if (o.i0 >= 0) {
                            if (o.i1 < 0 || (o.i1 & 0x40000000)) need = true;
                            else {
                                const float lo0 = __fsqrt_rn(static_cast<float>(max(0, na - 2 * D1)));      // parity can make it -1
                                const float hi1 = __fsqrt_rn(static_cast<float>(na - 2 * D2 + 1));
                                need = static_cast<double>(lo0) < static_cast<double>(hi1) * fz.ratio;
                            }
                        }
                    } else {
                        // four candidate chunks (0xFFFF = none; only chunks with a positive maximum, i.e. with a real row),
                        // V1, V2 and V5 >= 0 = an upper bound of a.b for every train row outside those chunks
                        o.i0 = (has[0] ? ch[0] : 0xFFFF) | (has[1] ? ch[1] : 0xFFFF) << 16;
                        o.i1 = (has[2] ? ch[2] : 0xFFFF) | (has[3] ? ch[3] : 0xFFFF) << 16;
                        const int V1 = has[0] ? vv[0] : -1, V2 = has[1] ? vv[1] : -1;
                        o.d0 = __int_as_float(V1);
                        o.d1 = __int_as_float(V2);
                        if (V2 <= 0) need = true;                                            // fewer than two chunks with a real maximum
                        else {
                            const float lo0 = __fsqrt_rn(static_cast<float>(max(0, na + nbmin - 2 * V1)));
                            const float hi1 = __fsqrt_rn(static_cast<float>(max(0, na + nbmax - 2 * V2)));
                            need = static_cast<double>(lo0) < static_cast<double>(hi1) * fz.ratio;
                        }
                    }
                }
                const unsigned bal = __ballot_sync(0xffffffffu, need);
                if (bal) {
                    const int rank = __popc(bal & ((1u << lane) - 1));
                    const int64_t srow = u.pd.out_row0 + row;
                    {
                        int base = 0;
                        if (lane == 0) base = atomicAdd(fz.need_count, __popc(bal));
                        base = __shfl_sync(0xffffffffu, base, 0);
                        if (need) {
                            out[srow] = o;
                            if (!kNorm) aux[srow] = has[4] ? vv[4] : 0;
                            fz.need_list[base + rank] = static_cast<int32_t>(srow);
                        }
                    }
                }
            }
            asm volatile("bar.sync 2, %0;" ::"n"(128 * kGroups) : "memory");
        }
    } else if (kRefine && warp >= kEpiWarp0 + 4 * kGroups) {
        // ================================================================ finishing warps: unit end + exact re-rank
        // Warp w evaluates query rows 32 w .. 32 w + 31 of every unit.  Per unit: merge the two column halves' top-5 lists (32-bit
        // keys with global chunk numbers: min / max only), give the buffer back, then the fused ratio-test bound (north_star:
        // "fused epilogue ... ratio test"): a row can only pass  sqrtf(d0^2) < ratio * sqrtf(d1^2)  if it passes with the
        // smallest d0^2 and the largest d1^2 the chunk maxima allow,  d0^2 >= |a|^2 + N- - 2 V1,  d1^2 <= |a|^2 + N+ - 2 V2  with
        // [N-, N+] the |b|^2 range of the train image.  ~99.5 % of C3's rows stop here and write NOTHING.  Survivors are re-ranked
        // by the four warps together (refine_dot_row: __dp4a on L2-resident bank rows), a unit's survivors dealt round-robin.
        RefineCtx rc;
        rc.bank = fz.bank; rc.norm2 = fz.norm2; rc.top2 = out; rc.stats = fz.stats; rc.bf_list = fz.bf_list; rc.bf_count = fz.bf_count;
        rc.chunk_rows = kCC; rc.all_rows = 0; rc.ratio = fz.ratio;
        const int row_in_unit = (warp & 3) * 32 + lane;
        const uint32_t* fin0 = reinterpret_cast<const uint32_t*>(base_ptr + offE);
        SurvList* surv = reinterpret_cast<SurvList*>(base_ptr + offE + kFinDepth * kFinBufWords * 4);
        if (warp == kEpiWarp0 + 4 * kGroups && lane == 0) surv->count = 0;
        asm volatile("bar.sync 3, 128;" ::: "memory");
        unsigned long long n_rows_done = 0;
        uint32_t unit_no = 0;
        int p_hint = 0, range_of = -1, nbmin = 0, nbmax = 0;
        for (int64_t unit = blockIdx.x; unit < n_units; unit += gridDim.x, ++unit_no) {
            const UnitInfoV u = decode_unit_v<BN>(pairs, unit_prefix, n_pairs, unit, fz.uniform_units, p_hint);
            const int row = u.rb * BM + row_in_unit;
            const bool in_range = row < u.pd.nq && u.pd.nt > 0;
            const int na = in_range ? __ldg(fz.norm2 + u.pd.q_row0 + row) : 0;
            if (range_of != u.pd.t_row0) {                          // |b|^2 range of the train image, from the per-block ranges
                int mn = INT_MAX, mx = 0;
                const int b0 = u.pd.t_row0 / kRowAlign, nblk = (u.pd.nt + kRowAlign - 1) / kRowAlign;
                for (int b = lane; b < nblk; b += 32) { mn = min(mn, __ldg(fz.blk_min + b0 + b)); mx = max(mx, __ldg(fz.blk_max + b0 + b)); }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
                nbmin = mn; nbmax = mx; range_of = u.pd.t_row0;
            }
            const uint32_t fpar = unit_no % kFinDepth;
            mbar_wait(fin_full(fpar), (unit_no / kFinDepth) & 1u);
            const uint32_t* fin = fin0 + fpar * kFinBufWords + row_in_unit;
            uint32_t r0 = fin[0 * BM], r1 = fin[1 * BM], r2 = fin[2 * BM], r3 = fin[3 * BM], r4 = fin[4 * BM];   // half 0: descending
#pragma unroll
            for (int i = 0; i < 5; ++i) top5_maxu(fin[(5 + i) * BM], r0, r1, r2, r3, r4);
            __syncwarp();
            if (lane == 0) mbar_arrive(fin_empty(fpar));
            // key = a.b << 11 | (2047 - chunk); value 0 = no chunk with a real row (padding, or orthogonal to the query: the re-rank
            // treats everything outside the candidates as a.b <= V5)
            const uint32_t rk[5] = {r0, r1, r2, r3, r4};
            int32_t vv[5], ch[5];
#pragma unroll
            for (int i = 0; i < 5; ++i) { vv[i] = static_cast<int32_t>(rk[i] >> kGchBits); ch[i] = static_cast<int32_t>(kGchMask - (rk[i] & kGchMask)); }
            Top2 o;
            o.i0 = (vv[0] > 0 ? ch[0] : 0xFFFF) | (vv[1] > 0 ? ch[1] : 0xFFFF) << 16;       // four candidate chunks, 0xFFFF = none
            o.i1 = (vv[2] > 0 ? ch[2] : 0xFFFF) | (vv[3] > 0 ? ch[3] : 0xFFFF) << 16;
            const int V1 = vv[0] > 0 ? vv[0] : -1, V2 = vv[1] > 0 ? vv[1] : -1, V5 = vv[4] > 0 ? vv[4] : 0;
            o.d0 = __int_as_float(V1);
            o.d1 = __int_as_float(V2);
            bool need = false;
            if (in_range) {
                if (V2 <= 0) need = true;                           // fewer than two chunks with a real maximum
                else {
                    const float lo0 = __fsqrt_rn(static_cast<float>(max(0, na + nbmin - 2 * V1)));
                    const float hi1 = __fsqrt_rn(static_cast<float>(max(0, na + nbmax - 2 * V2)));
                    need = static_cast<double>(lo0) < static_cast<double>(hi1) * fz.ratio;
                }
            }
            // survivors of the whole unit into one shared list, then dealt round-robin: a unit's survivors cluster in few warps' rows
            const unsigned bal = __ballot_sync(0xffffffffu, need);
            if (bal) {
                int base = 0;
                if (lane == 0) base = atomicAdd(&surv->count, __popc(bal));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (need) {
                    SurvRec& q = surv->rec[base + __popc(bal & ((1u << lane) - 1))];
                    q.t = o; q.v5 = V5; q.na = na; q.row = row;
                }
            }
            asm volatile("bar.sync 3, 128;" ::: "memory");
            const int n_surv = *reinterpret_cast<volatile int*>(&surv->count);
            // "behind": the epilogue has filled every other buffer of the ring; one more unit and it would wait for these warps.
            // Sparse lists (fz.refine == 1): the rest of this warp's share goes to the need list and the post pass, which has the
            // whole GPU.  Dense lists (fz.refine == 2): the epilogue waits — a re-rank under the MMA costs less than one in the post pass.
            const uint32_t next_par = (unit_no + kFinDepth - 1) % kFinDepth, next_phase = ((unit_no + kFinDepth - 1) / kFinDepth) & 1u;
            int it = warp & 3, n_mine = 0;
            for (; it < n_surv; it += 4) {
                if (fz.refine == 1) {
                    // ONE lane reads the barrier: the warp must take this branch as a whole (refine_dot_row shuffles), and two
                    // lanes testing an mbarrier a few cycles apart may see different phases
                    const int behind = lane == 0 ? static_cast<int>(mbar_test(fin_full(next_par), next_phase)) : 0;
                    if (__shfl_sync(0xffffffffu, behind, 0)) break;
                } else if (fz.refine == 3 && n_mine >= 1) break;       // test mode: one row per warp and unit here, the rest diverted
                const SurvRec q = surv->rec[it];
                const int64_t srow = u.pd.out_row0 + q.row;
                refine_dot_row(rc, srow, lane, q.t, q.v5, q.na, u.pd.q_row0 + q.row, u.pd.t_row0, u.pd.nt, nbmin, nbmax);
                ++n_mine;
            }
            if (n_mine) {                                            // rows this warp re-ranked: one append for all of them
                int base = 0;
                if (lane == 0) base = atomicAdd(fz.done_count, n_mine);
                base = __shfl_sync(0xffffffffu, base, 0);
                if (lane < n_mine) fz.done_list[base + lane] = static_cast<int32_t>(u.pd.out_row0 + surv->rec[(warp & 3) + 4 * lane].row);
                n_rows_done += n_mine;
            }
            if (it < n_surv) {
                // candidate record + need list, one atomic per warp
                const int n_left = (n_surv - it + 3) / 4;
                int base = 0;
                if (lane == 0) base = atomicAdd(fz.need_count, n_left);
                base = __shfl_sync(0xffffffffu, base, 0);
                if (lane < n_left) {
                    const SurvRec q = surv->rec[it + 4 * lane];
                    const int64_t srow = u.pd.out_row0 + q.row;
                    out[srow] = q.t;
                    aux[srow] = q.v5;
                    fz.need_list[base + lane] = static_cast<int32_t>(srow);
                }
            }
            asm volatile("bar.sync 3, 128;" ::: "memory");           // every warp is done with the list
            if (warp == kEpiWarp0 + 4 * kGroups && lane == 0) surv->count = 0;
        }
        if (lane == 0 && n_rows_done && fz.stats) atomicAdd(fz.stats, n_rows_done);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

template <int kParity, int kHalves, int kBN, bool kNorm, int kCC, bool kRefine = false>
static cudaError_t launch_tcv(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& te, const PairDesc* pairs,
                              const int64_t* unit_prefix, int n_pairs, int64_t n_units, Top2* out, int32_t* aux, int grid,
                              int issuers, const TcvFuse& fz, cudaStream_t s) {
    // per launch: the attribute is per device, and one process may drive several GPUs
    cudaError_t e = cudaFuncSetAttribute(knn2_l2_u8_tcv_kernel<kParity, kHalves, kBN, kNorm, kCC, kRefine>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, tcv::Cfg<kBN>::kSmemBytes);
    if (e != cudaSuccess) return e;
    knn2_l2_u8_tcv_kernel<kParity, kHalves, kBN, kNorm, kCC, kRefine><<<grid, 128 + 128 * kParity * kHalves + (kRefine ? 128 : 0), tcv::Cfg<kBN>::kSmemBytes, s>>>(
        ta, tb, te, pairs, unit_prefix, n_pairs, n_units, out, aux, issuers, fz);
    return cudaGetLastError();
}

// layout: epilogue organisation = 10 * kParity + kHalves: 12 (8 warps, column halves of every tile; fastest), 14 (16
// warps, column quarters of every tile), 21 (8 warps, alternate tiles).  The 32-bit running keys number at most 512 chunks per epilogue warp and unit: train
// images up to 32768 rows with two groups (12, 21), 65536 with four (14).  tile_rows: 128 (four TMEM accumulator stages; tmap_b /
// tmap_e must have 128-row boxes) or 256 (two stages, 256-row boxes).
cudaError_t launch_knn2_l2_u8_tcv(const void* tmap_a_host, const void* tmap_b_host, const void* tmap_e_host,
                                  const PairDesc* pairs, const int64_t* unit_prefix, int n_pairs, int64_t n_units,
                                  Top2* out, int32_t* aux /* non-null selects the norm-less variant */, int sm_count, int layout,
                                  int tile_rows, int issuers, int chunk_rows, const TcvFuse& fz, cudaStream_t s) {
    if (n_units == 0) return cudaSuccess;
    const CUtensorMap* ta = static_cast<const CUtensorMap*>(tmap_a_host);
    const CUtensorMap* tb = static_cast<const CUtensorMap*>(tmap_b_host);
    const CUtensorMap* te = static_cast<const CUtensorMap*>(tmap_e_host);
    const int grid = static_cast<int>(n_units < sm_count ? n_units : sm_count);
    if (fz.refine && aux && layout == 12 && tile_rows == 256) {
        // norm-less variant with in-kernel re-rank: 8 epilogue warps + 4 refine warps
        if (chunk_rows == 128) return launch_tcv<1, 2, 256, false, 128, true>(*ta, *tb, *te, pairs, unit_prefix, n_pairs, n_units, out, aux, grid, issuers, fz, s);
        if (chunk_rows == 64) return launch_tcv<1, 2, 256, false, 64, true>(*ta, *tb, *te, pairs, unit_prefix, n_pairs, n_units, out, aux, grid, issuers, fz, s);
        return launch_tcv<1, 2, 256, false, 32, true>(*ta, *tb, *te, pairs, unit_prefix, n_pairs, n_units, out, aux, grid, issuers, fz, s);
    }
#define SFM_TCV_CASE(P, H, T)                                                                                          \
    if (layout == 10 * P + H && tile_rows == T) {                                                                      \
        if (chunk_rows == 64)                                                                                          \
            return aux ? launch_tcv<P, H, T, false, 64>(*ta, *tb, *te, pairs, unit_prefix, n_pairs, n_units, out, aux, grid, issuers, fz, s) \
                       : launch_tcv<P, H, T, true, 64>(*ta, *tb, *te, pairs, unit_prefix, n_pairs, n_units, out, aux, grid, issuers, fz, s); \
        return aux ? launch_tcv<P, H, T, false, 32>(*ta, *tb, *te, pairs, unit_prefix, n_pairs, n_units, out, aux, grid, issuers, fz, s) \
                   : launch_tcv<P, H, T, true, 32>(*ta, *tb, *te, pairs, unit_prefix, n_pairs, n_units, out, aux, grid, issuers, fz, s); \
    }
    SFM_TCV_CASE(1, 2, 256) SFM_TCV_CASE(1, 4, 256) SFM_TCV_CASE(2, 1, 256)
#undef SFM_TCV_CASE
    return cudaErrorInvalidValue;
}

}  // namespace sfm
