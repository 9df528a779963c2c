// post.cu — the small kernels around the knn kernels:
//   pack_f32_to_u8   CV_32F integer-valued SIFT rows -> u8 bank rows (+ integer-valued check)
//   norms_ckeys      |b|^2 per bank row and the packed per-column key constant of the tcgen05 epilogue
//   filter_*         Lowe ratio test in double (UnorderedFeatureMatchingStrategy.cpp:55-65), match() k=1 mode
//                    (…:66-72), cross-check (cv::BFMatcher crossCheck semantics), distinct (SfM.cpp:547-564)
//   scan_offsets     per-pair counts, min-match-count drop (SfM.cpp:566-570), output offsets
//   compact          ordered DMatch lists (ascending queryIdx, pair order)
#include <climits>

#include "common.cuh"
#include "kernels.h"
#include "refine_dot.cuh"

namespace sfm {

constexpr int kNormL2 = 4;

// ------------------------------------------------------------------------------------------------ upload helpers
__global__ void pack_f32_to_u8_kernel(const float* __restrict__ src, size_t stride, int n_rows, int cols,
                                      const int32_t* __restrict__ valid_in_block, uint8_t* __restrict__ dst,
                                      int* __restrict__ bad) {
    const int vec_per_row = cols >> 2;
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= static_cast<int64_t>(n_rows) * vec_per_row) return;
    const int r = static_cast<int>(i / vec_per_row), v = static_cast<int>(i - static_cast<int64_t>(r) * vec_per_row);
    uint32_t w = 0;
    if ((r & (kRowAlign - 1)) < valid_in_block[r / kRowAlign]) {       // padding rows become zero descriptors
        const float4 f = *reinterpret_cast<const float4*>(src + r * stride + 4 * v);
        const float e[4] = {f.x, f.y, f.z, f.w};
        bool ok = true;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float x = e[k];
            ok = ok && (x >= 0.f) && (x <= 255.f) && (x == rintf(x));   // NaN fails x >= 0
            w |= (static_cast<uint32_t>(ok ? static_cast<int>(x) : 0) & 0xFFu) << (8 * k);
        }
        if (!ok) atomicOr(bad, 1);
    }
    *reinterpret_cast<uint32_t*>(dst + static_cast<size_t>(r) * cols + 4 * v) = w;
}

cudaError_t launch_pack_f32_to_u8(const float* src, size_t src_stride_elems, int n_rows, int cols,
                                  const int32_t* valid_in_block, uint8_t* dst, int* not_integer_flag, cudaStream_t s) {
    const int64_t n = static_cast<int64_t>(n_rows) * (cols >> 2);
    if (n == 0) return cudaSuccess;
    pack_f32_to_u8_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(src, src_stride_elems, n_rows, cols,
                                                                                  valid_in_block, dst, not_integer_flag);
    return cudaGetLastError();
}

// zero the padding rows of a bank (rows past each image's last descriptor, up to the 256-row boundary)
__global__ void zero_padding_kernel(uint8_t* __restrict__ bank, int row_vecs /*16-byte units per row*/, int64_t padded_rows,
                                    const int32_t* __restrict__ valid_in_block) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= padded_rows * row_vecs) return;
    const int64_t r = i / row_vecs;
    if ((r & (kRowAlign - 1)) >= valid_in_block[r / kRowAlign]) reinterpret_cast<uint4*>(bank)[i] = make_uint4(0, 0, 0, 0);
}

cudaError_t launch_zero_padding(void* bank, int row_bytes, int64_t padded_rows, const int32_t* valid_in_block, cudaStream_t s) {
    const int64_t n = padded_rows * (row_bytes / 16);
    if (n == 0) return cudaSuccess;
    zero_padding_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(static_cast<uint8_t*>(bank), row_bytes / 16,
                                                                                padded_rows, valid_in_block);
    return cudaGetLastError();
}

// one warp per 128-byte row: |b|^2, the packed column key of the IMAD epilogue, and the norm digits of the
// value-only epilogue (see common.cuh)
constexpr int kNormRowsPerWarp = 8;              // 8 warps x 8 rows = 64 rows per CTA: one global atomic pair per CTA
__global__ void __launch_bounds__(256) norms_ckeys_kernel(const uint8_t* __restrict__ bank, int64_t padded_rows,
                                   const int32_t* __restrict__ valid_in_block, int32_t* __restrict__ norm2,
                                   int32_t* __restrict__ ckey, int8_t* __restrict__ ext, int* __restrict__ max_norm2,
                                   int32_t* __restrict__ blk_min, int32_t* __restrict__ blk_max) {
    __shared__ int s_max, s_min;
    if (threadIdx.x == 0) { s_max = 0; s_min = INT_MAX; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t row_base = (static_cast<int64_t>(blockIdx.x) * 8 + warp) * kNormRowsPerWarp;
    int w_max = 0, w_min = INT_MAX;                  // over this warp's valid rows (lane 0)
    for (int r = 0; r < kNormRowsPerWarp; ++r) {
        const int64_t row = row_base + r;
        if (row >= padded_rows) break;
        const uint32_t w = *reinterpret_cast<const uint32_t*>(bank + row * 128 + 4 * lane);
        uint32_t s = __dp4a(w, w, 0u);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const int col = static_cast<int>(row & (kRowAlign - 1));
        const bool valid = col < valid_in_block[row / kRowAlign];
        if (lane == 0) {
            norm2[row] = static_cast<int32_t>(s);
            // key = (|b|^2 - 2ab) * 256 + col  ==  ckey - 512 * ab
            ckey[row] = valid ? static_cast<int32_t>((s << 8) | static_cast<uint32_t>(col)) : (kSentinelKey | col);
            if (valid) { w_max = max(w_max, static_cast<int>(s)); w_min = min(w_min, static_cast<int>(s)); }
        }
        // digit of lane p: weight 255 at positions with (p & 15) < 12, weight 1 otherwise
        int digit = 128;
        if (valid && s <= static_cast<uint32_t>(kExtMaxNorm2)) {
            const int g = static_cast<int>(s >> 1), q = g / 255, rr = g - q * 255;
            const int sub = lane & 15, half = lane >> 4;
            if (sub < 12) digit = min(128, max(0, q - 128 * (half * 12 + sub)));
            else          digit = min(128, max(0, rr - 128 * (half * 4 + sub - 12)));
        }
        ext[row * kExtBytes + lane] = static_cast<int8_t>(-digit);
    }
    // |b|^2 ranges: the 8 rows of a warp lie in one 256-row block (row_base is a multiple of 8), so one atomic pair per
    // warp on the block range (bounds of the norm-less path, refine_dot_kernel) and one per CTA on the bank range --
    // a per-row atomic on ONE address cost 1.7 ms on the C3 bank
    if (lane == 0 && w_min != INT_MAX) {
        atomicMin(blk_min + row_base / kRowAlign, w_min);
        atomicMax(blk_max + row_base / kRowAlign, w_max);
        atomicMax(&s_max, w_max);
        atomicMin(&s_min, w_min);
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_min != INT_MAX) {
        atomicMax(max_norm2, s_max);
        atomicMin(max_norm2 + 1, s_min);
    }
}

cudaError_t launch_norms_ckeys(const uint8_t* bank, int64_t padded_rows, const int32_t* valid_in_block,
                               int32_t* norm2, int32_t* ckey, int8_t* ext, int* max_norm2, int32_t* blk_min, int32_t* blk_max,
                               cudaStream_t s) {
    if (padded_rows == 0) return cudaSuccess;
    const int64_t rows_per_cta = 8 * kNormRowsPerWarp;
    norms_ckeys_kernel<<<static_cast<unsigned>((padded_rows + rows_per_cta - 1) / rows_per_cta), 256, 0, s>>>(
        bank, padded_rows, valid_in_block, norm2, ckey, ext, max_norm2, blk_min, blk_max);
    return cudaGetLastError();
}

// ORB: expand every descriptor bit to one byte (256-byte rows for the tcgen05 kernel: Hamming(a,b) = |a| + |b| - 2 a.b
// over 0/1 bytes), popcount as the "squared norm", packed column key.  One warp per 32-byte row.
__global__ void expand_bits_kernel(const uint8_t* __restrict__ bank32, int64_t padded_rows,
                                   const int32_t* __restrict__ valid_in_block, uint8_t* __restrict__ bits,
                                   int32_t* __restrict__ norm2, int32_t* __restrict__ ckey) {
    const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= padded_rows) return;
    const uint32_t byte = bank32[row * 32 + lane];
    uint2 o;
    o.x = (byte & 1u) | ((byte & 2u) << 7) | ((byte & 4u) << 14) | ((byte & 8u) << 21);
    o.y = ((byte >> 4) & 1u) | (((byte >> 4) & 2u) << 7) | (((byte >> 4) & 4u) << 14) | (((byte >> 4) & 8u) << 21);
    *reinterpret_cast<uint2*>(bits + row * 256 + lane * 8) = o;
    uint32_t s = __popc(byte);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) {
        const int col = static_cast<int>(row & (kRowAlign - 1));
        const bool valid = col < valid_in_block[row / kRowAlign];
        norm2[row] = static_cast<int32_t>(s);
        ckey[row] = valid ? static_cast<int32_t>((s << 8) | static_cast<uint32_t>(col)) : (kSentinelKey | col);
    }
}

cudaError_t launch_expand_bits(const uint8_t* bank32, int64_t padded_rows, const int32_t* valid_in_block, uint8_t* bits,
                               int32_t* norm2, int32_t* ckey, cudaStream_t s) {
    if (padded_rows == 0) return cudaSuccess;
    const int64_t threads = padded_rows * 32;
    expand_bits_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, s>>>(bank32, padded_rows, valid_in_block, bits,
                                                                                     norm2, ckey);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ filters
// pair of a staging row: 256-row staging blocks never straddle pairs, so a per-block table replaces the binary search
__device__ __forceinline__ int pair_of_row(const int32_t* __restrict__ blk_pair, const int64_t* __restrict__ out_prefix,
                                           int n_pairs, int64_t srow) {
    return blk_pair ? __ldg(blk_pair + (srow >> 8)) : find_segment(out_prefix, n_pairs, srow);
}

__global__ void block_pairs_kernel(const int64_t* __restrict__ out_prefix, int n_pairs, int64_t n_blocks,
                                   int32_t* __restrict__ blk_pair) {
    const int64_t b = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (b < n_blocks) blk_pair[b] = find_segment(out_prefix, n_pairs, b * 256);
}

cudaError_t launch_block_pairs(const int64_t* out_prefix, int n_pairs, int64_t n_blocks, int32_t* blk_pair, cudaStream_t s) {
    if (n_blocks == 0) return cudaSuccess;
    block_pairs_kernel<<<static_cast<unsigned>((n_blocks + 255) / 256), 256, 0, s>>>(out_prefix, n_pairs, n_blocks, blk_pair);
    return cudaGetLastError();
}

struct RowCtx { int p; int row; bool valid; PairDesc pd; Top2 t; };

__device__ __forceinline__ RowCtx load_row(const FilterArgs& a, int64_t srow) {
    RowCtx c;
    c.p = pair_of_row(a.blk_pair, a.out_prefix, a.n_pairs, srow);
    c.pd = a.pairs[c.p];
    c.row = static_cast<int>(srow - a.out_prefix[c.p]);
    c.valid = c.row < c.pd.nq;
    if (c.valid) c.t = a.top2[srow];
    return c;
}

__device__ __forceinline__ float final_distance(int norm, float d) {
    // L2: DMatch.distance = float32(sqrt(float32(sum))) with IEEE rounding
    return norm == kNormL2 ? __fsqrt_rn(d) : d;
}

// keep-decision before the distinct filter
__device__ __forceinline__ bool keep_basic(const FilterArgs& a, const RowCtx& c, float& dist0) {
    if (!c.valid || c.t.i0 < 0) return false;
    dist0 = final_distance(a.fp.norm, c.t.d0);
    if (a.fp.cross_check) {
        // mutual nearest neighbour: NN(train row NN(q)) == q
        const Top2 r = a.rev[a.t_prefix[c.p] + c.t.i0];
        return r.i0 == c.row;
    }
    if (a.fp.k < 2 || c.t.i1 < 0) return true;       // match() semantics / single neighbour kept (:62-64)
    const float dist1 = final_distance(a.fp.norm, c.t.d1);
    // float < float * double(0.7): evaluated in double (UnorderedFeatureMatchingStrategy.cpp:55-59)
    return static_cast<double>(dist0) < static_cast<double>(dist1) * a.fp.ratio;
}

__device__ __forceinline__ bool keep_final(const FilterArgs& a, const RowCtx& c, float& dist0) {
    bool k = keep_basic(a, c, dist0);
    if (k && a.fp.distinct) k = a.train_cnt[a.t_prefix[c.p] + c.t.i0] == 1;
    return k;
}

// Sparse form (value-only tcgen05 path): the knn kernel's fused ratio bound leaves only a short list of staging rows
// (need_list) with a Top2 record at all; after their exact re-rank mark_keep_kernel evaluates the keep decision for
// those rows only and sets one bit per kept row.  The count / compact passes then read 1 bit per staging row instead of
// a 16-byte record, and the record itself only for kept rows.
__device__ __forceinline__ bool keep_bit(const uint32_t* __restrict__ bits, int64_t srow) {
    return (__ldg(bits + (srow >> 5)) >> (srow & 31)) & 1u;
}

__global__ void __launch_bounds__(256) mark_keep_kernel(FilterArgs a) {
    // the rows that carry a final Top2 record sit in up to three lists: re-ranked by the post pass (need_list), re-ranked by
    // the refine warps inside the knn kernel (done_list), finished by brute force (bf_list; may repeat rows of the other two)
#pragma unroll 1
    for (int l = 0; l < 3; ++l) {
        const int32_t* list = l == 0 ? a.need_list : (l == 1 ? a.done_list : a.bf_list);
        const int* cnt = l == 0 ? a.need_count : (l == 1 ? a.done_count : a.bf_count);
        if (!list) continue;
        const int n = *cnt;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
            const int64_t srow = list[i];
            const RowCtx c = load_row(a, srow);
            float d;
            if (keep_basic(a, c, d)) {
                const uint32_t bit = 1u << (srow & 31);
                const uint32_t old = atomicOr(a.keep_bits + (srow >> 5), bit);
                if (a.fp.distinct && !(old & bit)) atomicAdd(a.train_cnt + a.t_prefix[c.p] + c.t.i0, 1);
            }
        }
    }
}

// distinct filter, second pass: kept rows whose best train row is shared lose their bit (SfM.cpp:547-564)
__global__ void __launch_bounds__(256) distinct_prune_kernel(FilterArgs a) {
#pragma unroll 1
    for (int l = 0; l < 3; ++l) {
        const int32_t* list = l == 0 ? a.need_list : (l == 1 ? a.done_list : a.bf_list);
        const int* cnt = l == 0 ? a.need_count : (l == 1 ? a.done_count : a.bf_count);
        if (!list) continue;
        const int n = *cnt;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
            const int64_t srow = list[i];
            if (!keep_bit(a.keep_bits, srow)) continue;
            const RowCtx c = load_row(a, srow);
            if (a.train_cnt[a.t_prefix[c.p] + c.t.i0] != 1) atomicAnd(a.keep_bits + (srow >> 5), ~(1u << (srow & 31)));
        }
    }
}

__global__ void __launch_bounds__(256) filter_mark_kernel(FilterArgs a) {
    const int64_t srow = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
    if (srow >= a.staged_rows) return;
    const RowCtx c = load_row(a, srow);
    float d;
    if (keep_basic(a, c, d)) atomicAdd(a.train_cnt + a.t_prefix[c.p] + c.t.i0, 1);
}

__global__ void __launch_bounds__(256) filter_count_kernel(FilterArgs a, int32_t* __restrict__ chunk_counts) {
    const int64_t srow = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
    bool k = false;
    if (srow < a.staged_rows) {
        const RowCtx c = load_row(a, srow);
        float d;
        k = keep_final(a, c, d);
    }
    const int n = __syncthreads_count(k);
    if (threadIdx.x == 0) chunk_counts[blockIdx.x] = n;
}

__global__ void __launch_bounds__(256) compact_kernel(FilterArgs a, const int64_t* __restrict__ chunk_excl,
                                                      const int64_t* __restrict__ pair_offsets,
                                                      const uint8_t* __restrict__ pair_dropped, DMatch* __restrict__ out,
                                                      int64_t capacity, int* __restrict__ overflow) {
    __shared__ int warp_cnt[8];
    const int64_t srow = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    bool k = false;
    float d = 0.f;
    RowCtx c;
    c.p = 0; c.row = 0; c.valid = false;
    if (srow < a.staged_rows) {
        c = load_row(a, srow);
        k = keep_final(a, c, d);
    }
    const unsigned bal = __ballot_sync(0xffffffffu, k);
    if (lane == 0) warp_cnt[warp] = __popc(bal);
    __syncthreads();
    int before = __popc(bal & ((1u << lane) - 1));
    for (int w = 0; w < warp; ++w) before += warp_cnt[w];
    if (!k) return;
    // a chunk never straddles two pairs (staging rows of a pair start at a multiple of 256)
    if (pair_dropped[c.p]) return;
    const int64_t first_chunk = a.out_prefix[c.p] >> 8;
    const int64_t pos = pair_offsets[c.p] + (chunk_excl[blockIdx.x] - chunk_excl[first_chunk]) + before;
    if (pos >= capacity) { atomicOr(overflow, 1); return; }
    DMatch m;
    m.queryIdx = c.row; m.trainIdx = c.t.i0; m.imgIdx = 0; m.distance = d;
    out[pos] = m;
}

// sparse count: one thread per 256-row staging block = 8 keep words
__global__ void __launch_bounds__(256) count_keep_bits_kernel(const uint32_t* __restrict__ keep_bits, int64_t n_blocks,
                                                              int32_t* __restrict__ chunk_counts) {
    const int64_t b = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (b >= n_blocks) return;
    const uint4* w = reinterpret_cast<const uint4*>(keep_bits + b * 8);
    const uint4 x = __ldg(w), y = __ldg(w + 1);
    chunk_counts[b] = __popc(x.x) + __popc(x.y) + __popc(x.z) + __popc(x.w) + __popc(y.x) + __popc(y.y) + __popc(y.z) + __popc(y.w);
}

// sparse compaction: one warp per 256-row staging block (persistent grid); lane l < 8 owns keep word l and emits the DMatch
// records of its set bits in ascending row order (a block holds ~1 kept row on an all-pairs list)
__global__ void __launch_bounds__(256) compact_keep_bits_kernel(FilterArgs a, int64_t n_blocks, const int64_t* __restrict__ chunk_excl,
                                                                const int64_t* __restrict__ pair_offsets,
                                                                const uint8_t* __restrict__ pair_dropped, DMatch* __restrict__ out,
                                                                int64_t capacity, int* __restrict__ overflow) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
    for (int64_t b = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); b < n_blocks; b += warps) {
        uint32_t w = lane < 8 ? __ldg(a.keep_bits + b * 8 + lane) : 0u;
        if (__ballot_sync(0xffffffffu, w != 0) == 0) continue;
        // kept rows in the words before mine
        int before = 0;
        const int cnt = __popc(w);
#pragma unroll
        for (int l = 0; l < 7; ++l) {
            const int cl = __shfl_sync(0xffffffffu, cnt, l);
            if (lane > l) before += cl;
        }
        if (w == 0) continue;
        const int p = a.blk_pair[b];
        if (pair_dropped[p]) continue;
        const int64_t first_chunk = a.out_prefix[p] >> 8;
        int64_t pos = pair_offsets[p] + (chunk_excl[b] - chunk_excl[first_chunk]) + before;
        const int64_t row0 = a.out_prefix[p];
        while (w) {
            const int bit = __ffs(w) - 1;
            w &= w - 1;
            const int64_t srow = b * 256 + lane * 32 + bit;
            if (pos >= capacity) { atomicOr(overflow, 1); break; }
            const Top2 t = a.top2[srow];
            DMatch m;
            m.queryIdx = static_cast<int>(srow - row0); m.trainIdx = t.i0; m.imgIdx = 0; m.distance = final_distance(a.fp.norm, t.d0);
            out[pos++] = m;
        }
    }
}

// ------------------------------------------------------------------------------------------------ refine
// The tcgen05 kernel reports rank 1 exactly and, as rank 2, the best train row OUTSIDE the 32-row chunk that
// holds rank 1.  The true second neighbour is the smaller of that and the second-best row inside the chunk.
// A row needs the recomputation only if it could still survive the ratio test: with the provisional
// d1' >= d1(true), d0 >= ratio*d1' already implies rejection.  For every row that needs it one warp recomputes
// the chunk: lane l takes train row chunk0 + l (128-byte row, 32 x __dp4a), then a warp min over the packed
// (distance, index) keys.  Exact integer arithmetic, ties -> lowest trainIdx.
template <bool kHamming>
__global__ void __launch_bounds__(256) refine_second_kernel(RefineArgs a) {
    const int64_t srow = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
    const int lane = threadIdx.x & 31;
    bool need = false, changed = false;
    Top2 t;
    t.i0 = -1; t.i1 = -1; t.d0 = 0.f; t.d1 = 0.f;
    int q_bank_row = 0, t_row0 = 0, nt = 0;
    if (srow < a.staged_rows) {
        const int p = pair_of_row(a.blk_pair, a.out_prefix, a.n_pairs, srow);
        const PairDesc pd = a.pairs[p];
        const int row = static_cast<int>(srow - a.out_prefix[p]);
        if (row < pd.nq) {
            t = a.top2[srow];
            q_bank_row = pd.q_row0 + row; t_row0 = pd.t_row0; nt = pd.nt;
            if (t.i0 >= 0) {
                if (a.all_rows || t.i1 < 0) need = true;
                else {
                    const float x0 = kHamming ? t.d0 : __fsqrt_rn(t.d0), x1 = kHamming ? t.d1 : __fsqrt_rn(t.d1);
                    need = static_cast<double>(x0) < static_cast<double>(x1) * a.ratio;
                }
            }
        }
    }
    unsigned mask = __ballot_sync(0xffffffffu, need);
    while (mask) {
        const int src = __ffs(mask) - 1;
        mask &= mask - 1;
        const int qrow = __shfl_sync(0xffffffffu, q_bank_row, src);
        const int tr0 = __shfl_sync(0xffffffffu, t_row0, src);
        const int ntr = __shfl_sync(0xffffffffu, nt, src);
        const int best = __shfl_sync(0xffffffffu, t.i0, src);
        const int j = (best & ~31) + lane;
        int32_t d;
        if (kHamming) {
            // 32-byte ORB rows: the distance itself, 8 x __popc
            const uint4* qv = reinterpret_cast<const uint4*>(a.bank + static_cast<size_t>(qrow) * 32);
            const uint4* tv = reinterpret_cast<const uint4*>(a.bank + (static_cast<size_t>(tr0) + j) * 32);
            const uint4 x0 = __ldg(qv), x1 = __ldg(qv + 1), y0 = __ldg(tv), y1 = __ldg(tv + 1);
            d = __popc(x0.x ^ y0.x) + __popc(x0.y ^ y0.y) + __popc(x0.z ^ y0.z) + __popc(x0.w ^ y0.w) +
                __popc(x1.x ^ y1.x) + __popc(x1.y ^ y1.y) + __popc(x1.z ^ y1.z) + __popc(x1.w ^ y1.w);
        } else {
            const uint4* qv = reinterpret_cast<const uint4*>(a.bank + static_cast<size_t>(qrow) * 128);
            const uint4* tv = reinterpret_cast<const uint4*>(a.bank + (static_cast<size_t>(tr0) + j) * 128);
            uint32_t dot = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint4 x = __ldg(qv + i), y = __ldg(tv + i);
                dot = __dp4a(x.x, y.x, dot); dot = __dp4a(x.y, y.y, dot);
                dot = __dp4a(x.z, y.z, dot); dot = __dp4a(x.w, y.w, dot);
            }
            d = a.norm2[qrow] + a.norm2[tr0 + j] - 2 * static_cast<int32_t>(dot);
        }
        long long key = (j < ntr && j != best) ? ((static_cast<long long>(d) << 32) | static_cast<unsigned>(j)) : LLONG_MAX;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) key = min(key, __shfl_xor_sync(0xffffffffu, key, o));
        if (lane == src && key != LLONG_MAX) {
            const float cd = static_cast<float>(static_cast<int32_t>(key >> 32));
            const int cj = static_cast<int>(key & 0xFFFFFFFFll);
            if (t.i1 < 0 || cd < t.d1 || (cd == t.d1 && cj < t.i1)) { t.i1 = cj; t.d1 = cd; changed = true; }
        }
    }
    if (changed) a.top2[srow] = t;
}

// ------------------------------------------------------------------------------------------------ refine (value-only)
// Input: per query row the two best 32-row train chunks by D = ab - (|b|^2 >> 1) and their D values, as written by
// knn2_l2_u8_tcv_kernel (i0 = chunk1 | (chunk3 + 1) << 16, i1 = chunk2 | ambiguous << 30, d0/d1 = D1/D2 as int bits;
// chunk3 is present when a third chunk's maximum ties the second, ambiguous when a fourth does too).
// Output: the exact Top2 {idx0, idx1, d0^2, d1^2} for rows that can still pass the ratio test, i0 = -1 for the rest.
//   bounds: d0^2 >= |a|^2 - 2 D1 and d1^2 <= |a|^2 - 2 D2 + 1  (the parity bit of |b|^2 is in [0,1]); sqrtf and the
//   double product are monotonic, so  sqrtf(lo0) >= ratio * sqrtf(hi1)  implies the real test fails.
// One warp per row that needs work: lane = train row of the chunk, 32 x __dp4a, lexicographic (d^2, idx) keys.
// one warp per row that survived the fused ratio bound of the knn kernel (persistent grid over need_list)
__global__ void __launch_bounds__(256, 4) refine_value_rows_kernel(RefineArgs a) {
    const int lane = threadIdx.x & 31;
    const int n = *a.need_count;
    if (blockIdx.x == 0 && threadIdx.x == 0 && a.stats) atomicAdd(a.stats, static_cast<unsigned long long>(n));   // feedback for the host
    const int warps = gridDim.x * (blockDim.x >> 5);
    const float inf = __int_as_float(0x7f800000);
    for (int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += warps) {
        const int64_t srow = a.need_list[i];
        const int p = pair_of_row(a.blk_pair, a.out_prefix, a.n_pairs, srow);
        const PairDesc pd = a.pairs[p];
        const int qrow = pd.q_row0 + static_cast<int>(srow - a.out_prefix[p]);
        const int tr0 = pd.t_row0, ntr = pd.nt;
        const Top2 t = a.top2[srow];
        const int c1raw = t.i0, c2raw = t.i1;
        const int c1 = c1raw & 0xFFFF, c3 = (c1raw >> 16) - 1;      // c3 >= 0: a third chunk ties the second
        const int na = a.norm2[qrow];
        uint4 q[8];
        const uint4* qv = reinterpret_cast<const uint4*>(a.bank + static_cast<size_t>(qrow) * 128);
#pragma unroll
        for (int k = 0; k < 8; ++k) q[k] = __ldg(qv + k);
        long long a1 = LLONG_MAX, a2 = LLONG_MAX;
        const int sub0 = a.chunk_rows / 32;
        bool done = false;
        // ---- stage A: the best chunk alone.  Every train row outside it has D <= D2, i.e. d^2 >= |a|^2 - 2 D2 =: lb2, and the best
        // row of chunk 2 has d^2 in [lb2, lb2 + 1] (the parity bit of |b|^2).  A real match has e0 < lb2: it is certified as THE
        // nearest neighbour, and the ratio test is decided from d1^2 in [min(e1', lb2), min(e1', lb2 + 1)] (e1' = second best inside
        // the chunk) unless the two ends disagree — a third of the exact distances of the full path (dense pair lists: C4).
        if (!a.all_rows && c2raw >= 0 && !(c2raw & 0x40000000)) {
            for (int k = 0; k < sub0; ++k) warp_chunk_candidates(a.bank, a.norm2, q, na, tr0, ntr, c1 * sub0 + k, lane, a1, a2);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const long long b1 = __shfl_xor_sync(0xffffffffu, a1, off), b2 = __shfl_xor_sync(0xffffffffu, a2, off);
                a2 = min(max(a1, b1), min(a2, b2));
                a1 = min(a1, b1);
            }
            if (a1 != LLONG_MAX) {
                const long long e0 = a1 >> 32;
                const long long lb2 = static_cast<long long>(na) - 2ll * __float_as_int(t.d1);
                if (e0 < lb2) {
                    const long long e1 = a2 != LLONG_MAX ? (a2 >> 32) : LLONG_MAX;
                    const long long lo1 = min(e1, lb2), hi1 = min(e1, lb2 + 1);
                    const float s0 = __fsqrt_rn(static_cast<float>(static_cast<int32_t>(e0)));
                    const bool pass_lo = static_cast<double>(s0) < static_cast<double>(__fsqrt_rn(static_cast<float>(static_cast<int32_t>(lo1)))) * a.ratio;
                    const bool pass_hi = static_cast<double>(s0) < static_cast<double>(__fsqrt_rn(static_cast<float>(static_cast<int32_t>(hi1)))) * a.ratio;
                    if (pass_lo == pass_hi) {
                        __syncwarp();
                        if (lane == 0) {
                            Top2 o;
                            o.i0 = static_cast<int>(a1 & 0xFFFFFFFFll); o.d0 = static_cast<float>(static_cast<int32_t>(e0));
                            // the second neighbour itself is not reported by the ratio-filtered stage: any index >= 0 marks "a
                            // second neighbour exists", d1 = the end of the interval that was tested (keep_basic re-runs the test)
                            o.i1 = a2 != LLONG_MAX ? static_cast<int>(a2 & 0xFFFFFFFFll) : ntr;
                            o.d1 = static_cast<float>(static_cast<int32_t>(pass_lo ? lo1 : hi1));
                            a.top2[srow] = o;
                        }
                        done = true;
                    }
                }
            }
            a1 = LLONG_MAX; a2 = LLONG_MAX;
        }
        if (done) continue;
        if (c2raw >= 0 && (c2raw & 0x40000000)) {
            // ambiguous: a fourth chunk ties the second one -> exact brute force over the whole train image
            warp_brute_force(a.bank, a.norm2, q, na, tr0, ntr, lane, a1, a2);
        } else {
            const int sub = a.chunk_rows / 32;                      // a candidate chunk = sub groups of 32 train rows
            for (int k = 0; k < sub; ++k) {
                warp_chunk_candidates(a.bank, a.norm2, q, na, tr0, ntr, c1 * sub + k, lane, a1, a2);
                if (c2raw >= 0) warp_chunk_candidates(a.bank, a.norm2, q, na, tr0, ntr, c2raw * sub + k, lane, a1, a2);
                if (c3 >= 0) warp_chunk_candidates(a.bank, a.norm2, q, na, tr0, ntr, c3 * sub + k, lane, a1, a2);
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const long long b1 = __shfl_xor_sync(0xffffffffu, a1, off), b2 = __shfl_xor_sync(0xffffffffu, a2, off);
            a2 = min(max(a1, b1), min(a2, b2));
            a1 = min(a1, b1);
        }
        __syncwarp();                                               // every lane has read the candidate record
        if (lane == 0) {
            Top2 o;
            o.i0 = -1; o.i1 = -1; o.d0 = inf; o.d1 = inf;
            if (a1 != LLONG_MAX) { o.i0 = static_cast<int>(a1 & 0xFFFFFFFFll); o.d0 = static_cast<float>(static_cast<int32_t>(a1 >> 32)); }
            if (a2 != LLONG_MAX) { o.i1 = static_cast<int>(a2 & 0xFFFFFFFFll); o.d1 = static_cast<float>(static_cast<int32_t>(a2 >> 32)); }
            a.top2[srow] = o;
        }
    }
}

// ------------------------------------------------------------------------------------------------ refine (norm-less)
// Input per query row, from knn2_l2_u8_tcv_kernel<.., kNorm = false>: up to four candidate chunks (i0 = c0 | c1 << 16,
// i1 = c2 | c3 << 16, 0xFFFF = none; only chunks whose maximum is > 0, so every candidate chunk holds a real row) ranked
// by chunk maximum of the raw dot product, V1 = d0, V2 = d1 (int bits, -1 = none) and aux = V5 >= 0, an upper bound of
// a.b for every train row outside the candidate chunks (zero-padding rows and orthogonal rows have a.b = 0).
// With N- <= |b|^2 <= N+ over the train image (per-block ranges from norms_ckeys_kernel):
//     every train row                   : d^2 >= |a|^2 + N- - 2 V1
//     the best rows of chunks 1 and 2   : d^2 <= |a|^2 + N+ - 2 V2          (two different rows)
//     every row outside the four chunks : d^2 >= |a|^2 + N- - 2 V5  =: lbo
// (1) rows with sqrtf(|a|^2 + N- - 2 V1) >= ratio * sqrtf(|a|^2 + N+ - 2 V2) cannot pass the ratio test: rejected.
// (2a) the others first recompute the best chunk alone (stage A in the kernel): a planted match is certified against
//     lb2 = |a|^2 + N- - 2 V2 and its ratio test decided from d1^2 in [min(e1', lb2), min(e1', |a|^2 + N+ - 2 V2)].
// (2) what is left recomputes all candidate rows exactly (__dp4a) -> (e0, j0), (e1, j1), exact within the chunks.
//     e0 < lbo certifies (e0, j0) as THE nearest neighbour (ties resolve to the lowest index inside the chunks; a tie
//     with an outside row is excluded by the strict '<').  The second neighbour's d1^2 lies in [min(e1, lbo), e1]; the
//     ratio test  sqrtf(e0) < ratio * sqrtf(d1^2)  is monotone in d1^2, so it is decided if both ends agree, and the
//     reported (d0, idx0) never depends on d1.  In that case d1 is written as the end that was used (keep_basic re-runs
//     the same comparison); the lists equal cv::BFMatcher + ratio test bit for bit.
// (3) anything else (not certified, or the two ends disagree) is brute-forced over the whole train image.
__device__ __forceinline__ RefineCtx refine_ctx_of(const RefineArgs& a) {
    RefineCtx c;
    c.bank = a.bank; c.norm2 = a.norm2; c.top2 = a.top2; c.stats = a.stats; c.bf_list = a.bf_list; c.bf_count = a.bf_count;
    c.chunk_rows = a.chunk_rows; c.all_rows = a.all_rows; c.ratio = a.ratio;
    return c;
}

// |b|^2 range of the train image of every pair of the batch (from the per-256-row-block ranges): one warp per pair
__global__ void pair_norm_range_kernel(RefineArgs a) {
    const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (p >= a.n_pairs) return;
    const PairDesc pd = a.pairs[p];
    const int b0 = pd.t_row0 / kRowAlign, nb = (pd.nt + kRowAlign - 1) / kRowAlign;
    int mn = INT_MAX, mx = 0;
    for (int b = lane; b < nb; b += 32) { mn = min(mn, a.blk_min[b0 + b]); mx = max(mx, a.blk_max[b0 + b]); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
    if (lane == 0) { a.pair_nb[2 * p] = mn; a.pair_nb[2 * p + 1] = mx; }
}

// one warp per row that survived the fused ratio bound of the knn kernel (persistent grid over need_list).  The row's metadata is a chain of five dependent
// loads (list -> block table -> pair -> candidates / norm / bounds): it is fetched one row ahead of the row being refined.
struct DotRowMeta { int64_t srow; Top2 t; int v5, na, qrow, tr0, nt, nbmin, nbmax; };
__device__ __forceinline__ DotRowMeta load_dot_row(const RefineArgs& a, int i) {
    DotRowMeta m;
    m.srow = a.need_list[i];
    const int p = pair_of_row(a.blk_pair, a.out_prefix, a.n_pairs, m.srow);
    const PairDesc pd = a.pairs[p];
    m.qrow = pd.q_row0 + static_cast<int>(m.srow - a.out_prefix[p]);
    m.t = a.top2[m.srow];
    m.v5 = a.aux[m.srow];
    m.na = a.norm2[m.qrow];
    m.tr0 = pd.t_row0; m.nt = pd.nt;
    m.nbmin = a.pair_nb[2 * p]; m.nbmax = a.pair_nb[2 * p + 1];
    return m;
}

__global__ void __launch_bounds__(256, 4) refine_dot_rows_kernel(RefineArgs a) {
    const int lane = threadIdx.x & 31;
    const int n = *a.need_count;
    if (blockIdx.x == 0 && threadIdx.x == 0 && a.stats) atomicAdd(a.stats, static_cast<unsigned long long>(n));   // feedback for the host
    const int warps = gridDim.x * (blockDim.x >> 5);
    int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n) return;
    DotRowMeta cur = load_dot_row(a, i);
    while (true) {
        const int ni = i + warps;
        DotRowMeta nxt = cur;
        if (ni < n) nxt = load_dot_row(a, ni);
        refine_dot_row(refine_ctx_of(a), cur.srow, lane, cur.t, cur.v5, cur.na, cur.qrow, cur.tr0, cur.nt, cur.nbmin, cur.nbmax);
        if (ni >= n) break;
        cur = nxt; i = ni;
    }
}

// Rows queued by refine_dot_row: exact top-2 over the whole train image, one CTA per row (8 warps take every eighth
// group of 4 x 32 train rows), lexicographic (d^2, idx) keys, block reduction through shared memory.
__global__ void __launch_bounds__(256) brute_force_rows_kernel(RefineArgs a) {
    __shared__ long long s_k[16];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = *a.bf_count;
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        const int64_t srow = a.bf_list[i];
        const int p = pair_of_row(a.blk_pair, a.out_prefix, a.n_pairs, srow);
        const PairDesc pd = a.pairs[p];
        const int qrow = pd.q_row0 + static_cast<int>(srow - a.out_prefix[p]);
        const int na = a.norm2[qrow];
        uint4 q[8];
        const uint4* qv = reinterpret_cast<const uint4*>(a.bank + static_cast<size_t>(qrow) * 128);
#pragma unroll
        for (int k = 0; k < 8; ++k) q[k] = __ldg(qv + k);
        long long a1 = LLONG_MAX, a2 = LLONG_MAX;
        // warp w takes train rows [w * 128 + 1024 * m, +128): a sub-range view of the image for warp_brute_force
        for (int r0 = warp * 128; r0 < pd.nt; r0 += 1024) {
            long long b1 = LLONG_MAX, b2 = LLONG_MAX;
            warp_brute_force(a.bank, a.norm2, q, na, pd.t_row0 + r0, min(128, pd.nt - r0), lane, b1, b2);
            // indices are relative to r0: rebase, then merge
            if (b1 != LLONG_MAX) b1 += r0;
            if (b2 != LLONG_MAX) b2 += r0;
            a2 = min(max(a1, b1), min(a2, b2));
            a1 = min(a1, b1);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const long long b1 = __shfl_xor_sync(0xffffffffu, a1, off), b2 = __shfl_xor_sync(0xffffffffu, a2, off);
            a2 = min(max(a1, b1), min(a2, b2));
            a1 = min(a1, b1);
        }
        __syncthreads();                                  // previous row's s_k has been read
        if (lane == 0) { s_k[2 * warp] = a1; s_k[2 * warp + 1] = a2; }
        __syncthreads();
        if (threadIdx.x == 0) {
            long long m1 = LLONG_MAX, m2 = LLONG_MAX;
            for (int w = 0; w < 8; ++w) {
                const long long b1 = s_k[2 * w], b2 = s_k[2 * w + 1];
                m2 = min(max(m1, b1), min(m2, b2));
                m1 = min(m1, b1);
            }
            const float inf = __int_as_float(0x7f800000);
            Top2 o;
            o.i0 = -1; o.i1 = -1; o.d0 = inf; o.d1 = inf;
            if (m1 != LLONG_MAX) { o.i0 = static_cast<int>(m1 & 0xFFFFFFFFll); o.d0 = static_cast<float>(static_cast<int32_t>(m1 >> 32)); }
            if (m2 != LLONG_MAX) { o.i1 = static_cast<int>(m2 & 0xFFFFFFFFll); o.d1 = static_cast<float>(static_cast<int32_t>(m2 >> 32)); }
            a.top2[srow] = o;
        }
    }
}

cudaError_t launch_refine_dot(const RefineArgs& a, cudaStream_t s) {
    if (a.staged_rows == 0) return cudaSuccess;
    pair_norm_range_kernel<<<static_cast<unsigned>((a.n_pairs + 7) / 8), 256, 0, s>>>(a);
    refine_dot_rows_kernel<<<148 * 4, 256, 0, s>>>(a);       // persistent over the rows that survived the fused ratio bound
    brute_force_rows_kernel<<<592, 256, 0, s>>>(a);          // persistent over the queue (usually a few dozen rows)
    return cudaGetLastError();
}

static inline unsigned chunks_of(int64_t rows) { return static_cast<unsigned>((rows + 255) / 256); }

cudaError_t launch_refine_second(const RefineArgs& a, cudaStream_t s) {
    if (a.staged_rows == 0) return cudaSuccess;
    if (a.hamming) refine_second_kernel<true><<<chunks_of(a.staged_rows), 256, 0, s>>>(a);
    else refine_second_kernel<false><<<chunks_of(a.staged_rows), 256, 0, s>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_refine_value(const RefineArgs& a, cudaStream_t s) {
    if (a.staged_rows == 0) return cudaSuccess;
    refine_value_rows_kernel<<<148 * 4, 256, 0, s>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_count_keep_bits(const uint32_t* keep_bits, int64_t n_blocks, int32_t* chunk_counts, cudaStream_t s) {
    if (n_blocks == 0) return cudaSuccess;
    count_keep_bits_kernel<<<static_cast<unsigned>((n_blocks + 255) / 256), 256, 0, s>>>(keep_bits, n_blocks, chunk_counts);
    return cudaGetLastError();
}
cudaError_t launch_compact_keep_bits(const FilterArgs& a, const int64_t* chunk_excl, const int64_t* pair_offsets,
                                     const uint8_t* pair_dropped, DMatch* out, int64_t out_capacity, int* overflow_flag,
                                     cudaStream_t s) {
    if (a.staged_rows == 0) return cudaSuccess;
    compact_keep_bits_kernel<<<148 * 2, 256, 0, s>>>(a, a.staged_rows / 256, chunk_excl, pair_offsets, pair_dropped, out, out_capacity,
                                                      overflow_flag);
    return cudaGetLastError();
}
cudaError_t launch_mark_keep(const FilterArgs& a, cudaStream_t s) {
    if (a.staged_rows == 0) return cudaSuccess;
    mark_keep_kernel<<<148 * 2, 256, 0, s>>>(a);
    if (a.fp.distinct) distinct_prune_kernel<<<148 * 2, 256, 0, s>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_filter_mark(const FilterArgs& a, cudaStream_t s) {
    if (a.staged_rows == 0) return cudaSuccess;
    filter_mark_kernel<<<chunks_of(a.staged_rows), 256, 0, s>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_filter_count(const FilterArgs& a, int32_t* chunk_counts, cudaStream_t s) {
    if (a.staged_rows == 0) return cudaSuccess;
    filter_count_kernel<<<chunks_of(a.staged_rows), 256, 0, s>>>(a, chunk_counts);
    return cudaGetLastError();
}
cudaError_t launch_compact(const FilterArgs& a, const int64_t* chunk_excl, const int64_t* pair_offsets,
                           const uint8_t* pair_dropped, DMatch* out, int64_t out_capacity, int* overflow_flag,
                           cudaStream_t s) {
    if (a.staged_rows == 0) return cudaSuccess;
    compact_kernel<<<chunks_of(a.staged_rows), 256, 0, s>>>(a, chunk_excl, pair_offsets, pair_dropped, out, out_capacity,
                                                            overflow_flag);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ scan
// Single-CTA exclusive scan over `n` values produced by `get(i)`, results through `put(i, excl)`; returns the
// total to every thread.  n is at most a few hundred thousand (one value per 256 staged rows).
template <class Get, class Put>
__device__ int64_t block_exclusive_scan(int64_t n, Get get, Put put) {
    __shared__ int64_t warp_sums[32];
    __shared__ int64_t total_s;
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int64_t per = (n + nthr - 1) / nthr;
    const int64_t lo = min(n, per * tid), hi = min(n, lo + per);
    int64_t sum = 0;
    for (int64_t i = lo; i < hi; ++i) sum += get(i);
    int64_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int64_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int64_t w = lane < (nthr >> 5) ? warp_sums[lane] : 0;
        int64_t wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int64_t v = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += v;
        }
        warp_sums[lane] = wi - w;                 // exclusive prefix of warp totals
        if (lane == 31) total_s = wi;
    }
    __syncthreads();
    int64_t run = warp_sums[warp] + (incl - sum);
    for (int64_t i = lo; i < hi; ++i) { const int64_t v = get(i); put(i, run); run += v; }
    __syncthreads();
    return total_s;
}

__global__ void __launch_bounds__(1024) scan_offsets_kernel(const int32_t* __restrict__ chunk_counts, int64_t n_chunks,
                                                            const int64_t* __restrict__ out_prefix, int n_pairs,
                                                            int min_match_count, int64_t* __restrict__ chunk_excl,
                                                            int64_t* __restrict__ pair_counts,
                                                            int64_t* __restrict__ pair_offsets,
                                                            uint8_t* __restrict__ pair_dropped,
                                                            int64_t* __restrict__ running_total) {
    const int64_t total = block_exclusive_scan(
        n_chunks, [&](int64_t i) { return static_cast<int64_t>(chunk_counts[i]); },
        [&](int64_t i, int64_t e) { chunk_excl[i] = e; });
    if (threadIdx.x == 0) chunk_excl[n_chunks] = total;
    __syncthreads();
    for (int p = threadIdx.x; p < n_pairs; p += blockDim.x) {
        const int64_t c = chunk_excl[out_prefix[p + 1] >> 8] - chunk_excl[out_prefix[p] >> 8];
        const bool drop = c < min_match_count;
        pair_dropped[p] = drop ? 1 : 0;
        pair_counts[p] = drop ? 0 : c;
    }
    __syncthreads();
    const int64_t base = *running_total;
    const int64_t kept = block_exclusive_scan(
        n_pairs, [&](int64_t i) { return pair_counts[i]; },
        [&](int64_t i, int64_t e) { pair_offsets[i] = base + e; });
    if (threadIdx.x == 0) *running_total = base + kept;
}

cudaError_t launch_scan_offsets(const int32_t* chunk_counts, int64_t n_chunks, const int64_t* out_prefix, int n_pairs,
                                int min_match_count, int64_t* chunk_excl, int64_t* pair_counts_tmp, int64_t* pair_offsets,
                                uint8_t* pair_dropped, int64_t* running_total, cudaStream_t s) {
    if (n_pairs == 0) return cudaSuccess;
    scan_offsets_kernel<<<1, 1024, 0, s>>>(chunk_counts, n_chunks, out_prefix, n_pairs, min_match_count, chunk_excl,
                                           pair_counts_tmp, pair_offsets, pair_dropped, running_total);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ float (3xTF32) path
// Split every fp32 bank row into hi = tf32(x) and lo = x - hi, its squared norm, and the three tf32-exact pieces of
// -|b|^2/2 that the norm K step of knn2_l2_f32_tc3_kernel multiplies with {1,1,1}.  One warp per 128-float row.
__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

__global__ void f32_split_kernel(const float* __restrict__ bank, int64_t padded_rows,
                                 const int32_t* __restrict__ valid_in_block, float* __restrict__ hi, float* __restrict__ lo,
                                 float* __restrict__ fnorm2, float* __restrict__ ext, int* __restrict__ max_norm_bits) {
    __shared__ int s_max;                                         // largest |b|^2 of the CTA's rows (float bits)
    if (threadIdx.x == 0) s_max = 0;
    __syncthreads();
    const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row < padded_rows) {
    const float4 x = reinterpret_cast<const float4*>(bank + row * 128)[lane];
    float4 h, l;
    h.x = tf32_trunc(x.x); h.y = tf32_trunc(x.y); h.z = tf32_trunc(x.z); h.w = tf32_trunc(x.w);
    l.x = x.x - h.x; l.y = x.y - h.y; l.z = x.z - h.z; l.w = x.w - h.w;
    reinterpret_cast<float4*>(hi + row * 128)[lane] = h;
    reinterpret_cast<float4*>(lo + row * 128)[lane] = l;
    float s = fmaf(x.x, x.x, fmaf(x.y, x.y, fmaf(x.z, x.z, x.w * x.w)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
        const bool valid = static_cast<int>(row & (kRowAlign - 1)) < valid_in_block[row / kRowAlign];
        fnorm2[row] = s;
        const float g = valid ? 0.5f * s : 1e30f;                 // padded rows: D~ = -1e30, never a candidate
        const float g0 = tf32_trunc(g), g1 = tf32_trunc(g - g0), g2 = tf32_trunc(g - g0 - g1);
        float4* e = reinterpret_cast<float4*>(ext + row * 8);
        e[0] = make_float4(-g0, -g1, -g2, 0.f);
        e[1] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid && s >= 0.f) atomicMax(&s_max, __float_as_int(s));          // non-negative floats order like ints
    }
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_max > 0) atomicMax(max_norm_bits, s_max);      // one global atomic per CTA, not per row
}

cudaError_t launch_f32_split(const float* bank, int64_t padded_rows, const int32_t* valid_in_block, float* hi, float* lo,
                             float* fnorm2, float* ext, int* max_norm_bits, cudaStream_t s) {
    if (padded_rows == 0) return cudaSuccess;
    const int64_t threads = padded_rows * 32;
    f32_split_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, s>>>(bank, padded_rows, valid_in_block, hi, lo,
                                                                                   fnorm2, ext, max_norm_bits);
    return cudaGetLastError();
}

// Exact fp32 re-rank + certificate for the 3xTF32 candidate kernel.  Input per query row: up to four candidate
// chunks (i0 = c0 | c1 << 16, i1 = c2 | c3 << 16, 0xFFFF = none), v0 = d0, v1 = d1 (best / second-best chunk maximum
// of D~ = a.b - |b|^2/2) and aux = v4 (fifth-best chunk maximum: an upper bound of D~ for every row outside the four
// chunks).  |D~ - D| <= eps(row).  Output: exact Top2 (d^2 = sum (a-b)^2 in fp32), i0 = -1 for rows rejected by the
// provisional ratio test.
__device__ __forceinline__ void warp_chunk_candidates_f32(const float* __restrict__ bank, const float4 (&q)[32], int tr0,
                                                          int ntr, int chunk, int lane, unsigned long long& a1,
                                                          unsigned long long& a2) {
    const int j = chunk * 32 + lane;
    if (j < ntr) {
        const float4* tv = reinterpret_cast<const float4*>(bank + (static_cast<size_t>(tr0) + j) * 128);
        float d = 0.f;
#pragma unroll 8
        for (int i = 0; i < 32; ++i) {
            const float4 y = __ldg(tv + i);
            const float e0 = q[i].x - y.x, e1 = q[i].y - y.y, e2 = q[i].z - y.z, e3 = q[i].w - y.w;
            d = fmaf(e0, e0, d); d = fmaf(e1, e1, d); d = fmaf(e2, e2, d); d = fmaf(e3, e3, d);
        }
        if (d == d) {                                               // NaN never becomes a neighbour
            const unsigned long long key = (static_cast<unsigned long long>(__float_as_uint(d)) << 32) | static_cast<unsigned>(j);
            a2 = min(a2, max(a1, key));
            a1 = min(a1, key);
        }
    }
}

__global__ void __launch_bounds__(256) refine_f32_kernel(RefineF32Args a) {
    const int64_t srow = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
    const int lane = threadIdx.x & 31;
    bool need = false, valid = false;
    Top2 t;
    t.i0 = -1; t.i1 = -1; t.d0 = 0.f; t.d1 = 0.f;
    float v4 = 0.f, eps = 0.f, na = 0.f;
    int q_bank_row = 0, t_row0 = 0, nt = 0;
    if (srow < a.staged_rows) {
        const int p = pair_of_row(a.blk_pair, a.out_prefix, a.n_pairs, srow);
        PairDesc pd = a.pairs[p];
        if (a.swap_roles) {
            const PairDesc f = pd;
            pd.q_row0 = f.t_row0; pd.nq = f.nt; pd.t_row0 = f.q_row0; pd.nt = f.nq;
        }
        const int row = static_cast<int>(srow - a.out_prefix[p]);
        if (row < pd.nq) {
            valid = true;
            t = a.top2[srow];
            v4 = a.aux[srow];
            q_bank_row = pd.q_row0 + row; t_row0 = pd.t_row0; nt = pd.nt;
            na = a.fnorm2[q_bank_row];
            // 3xTF32 + fp32 accumulation + fp32 norms: |D~ - D| well below 2^-13 of the magnitudes involved
            eps = 1.220703125e-4f * (sqrtf(na * a.nb_max) + na + a.nb_max);
            if ((t.i0 & 0xFFFF) != 0xFFFF) {
                if (a.all_rows || (t.i0 >> 16) == 0xFFFF) need = true;        // fewer than two chunks known: look inside
                else {
                    const float lo0 = fmaxf(0.f, na - 2.f * (t.d0 + eps)), hi1 = fmaxf(0.f, na - 2.f * (t.d1 - eps));
                    need = static_cast<double>(__fsqrt_rn(lo0)) * (1.0 - 1e-6) < static_cast<double>(__fsqrt_rn(hi1)) * a.ratio;
                }
            }
        }
    }
    const float inf = __int_as_float(0x7f800000);
    Top2 o;
    o.i0 = -1; o.i1 = -1; o.d0 = inf; o.d1 = inf;
    unsigned mask = __ballot_sync(0xffffffffu, need);
    if (lane == 0 && mask) atomicAdd(a.stats, static_cast<unsigned long long>(__popc(mask)));
    while (mask) {
        const int src = __ffs(mask) - 1;
        mask &= mask - 1;
        const int qrow = __shfl_sync(0xffffffffu, q_bank_row, src);
        const int tr0 = __shfl_sync(0xffffffffu, t_row0, src);
        const int ntr = __shfl_sync(0xffffffffu, nt, src);
        const int i0 = __shfl_sync(0xffffffffu, t.i0, src), i1 = __shfl_sync(0xffffffffu, t.i1, src);
        const float rv4 = __shfl_sync(0xffffffffu, v4, src), reps = __shfl_sync(0xffffffffu, eps, src);
        const float rna = __shfl_sync(0xffffffffu, na, src);
        float4 q[32];
        const float4* qv = reinterpret_cast<const float4*>(a.bank + static_cast<size_t>(qrow) * 128);
#pragma unroll
        for (int i = 0; i < 32; ++i) q[i] = __ldg(qv + i);
        const int cand[4] = {i0 & 0xFFFF, (i0 >> 16) & 0xFFFF, i1 & 0xFFFF, (i1 >> 16) & 0xFFFF};
        unsigned long long a1 = ~0ull, a2 = ~0ull;
        for (int k = 0; k < 4; ++k)
            if (cand[k] != 0xFFFF) warp_chunk_candidates_f32(a.bank, q, tr0, ntr, cand[k], lane, a1, a2);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const unsigned long long b1 = __shfl_xor_sync(0xffffffffu, a1, off), b2 = __shfl_xor_sync(0xffffffffu, a2, off);
            a2 = min(max(a1, b1), min(a2, b2));
            a1 = min(a1, b1);
        }
        // certificate: rows outside the candidate chunks have D~ <= v4, i.e. d^2 >= |a|^2 - 2 (v4 + eps)
        const bool all_chunks_known = cand[3] == 0xFFFF || rv4 == __int_as_float(0xff800000);
        const float lb_other = rna - 2.f * (rv4 + reps);
        const bool certified = all_chunks_known ||
                               (a2 != ~0ull && __uint_as_float(static_cast<unsigned>(a2 >> 32)) * (1.f + 1e-6f) < lb_other);
        if (!certified) {                                           // exact brute force over the whole train image
            if (lane == 0) atomicAdd(a.stats + 1, 1ull);
            a1 = ~0ull; a2 = ~0ull;
            for (int c = 0; c * 32 < ntr; ++c) warp_chunk_candidates_f32(a.bank, q, tr0, ntr, c, lane, a1, a2);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const unsigned long long b1 = __shfl_xor_sync(0xffffffffu, a1, off), b2 = __shfl_xor_sync(0xffffffffu, a2, off);
                a2 = min(max(a1, b1), min(a2, b2));
                a1 = min(a1, b1);
            }
        }
        if (lane == src) {
            if (a1 != ~0ull) { o.i0 = static_cast<int>(a1 & 0xFFFFFFFFull); o.d0 = __uint_as_float(static_cast<unsigned>(a1 >> 32)); }
            if (a2 != ~0ull) { o.i1 = static_cast<int>(a2 & 0xFFFFFFFFull); o.d1 = __uint_as_float(static_cast<unsigned>(a2 >> 32)); }
        }
    }
    if (valid) a.top2[srow] = o;
}

cudaError_t launch_refine_f32(const RefineF32Args& a, cudaStream_t s) {
    if (a.staged_rows == 0) return cudaSuccess;
    refine_f32_kernel<<<static_cast<unsigned>((a.staged_rows + 255) / 256), 256, 0, s>>>(a);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ reorder
// The pipelined host path schedules pairs by availability of their images, so the compacted lists come out in
// schedule order; these three passes bring them back to INPUT pair order (order[k] = input index of scheduled pair k).
__global__ void reorder_counts_kernel(const int64_t* __restrict__ off_s, const int64_t* __restrict__ total,
                                      const int64_t* __restrict__ order, int64_t n, int64_t* __restrict__ cnt_in) {
    const int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int64_t end = k + 1 < n ? off_s[k + 1] : *total;
    cnt_in[order[k]] = end - off_s[k];
}
__global__ void __launch_bounds__(1024) reorder_scan_kernel(const int64_t* __restrict__ cnt_in, int64_t n,
                                                            int64_t* __restrict__ off_in) {
    block_exclusive_scan(n, [&](int64_t i) { return cnt_in[i]; }, [&](int64_t i, int64_t e) { off_in[i] = e; });
}
__global__ void reorder_copy_kernel(const DMatch* __restrict__ src, const int64_t* __restrict__ off_s,
                                    const int64_t* __restrict__ total, const int64_t* __restrict__ order,
                                    const int64_t* __restrict__ off_in, const uint8_t* __restrict__ drop_s, int64_t n,
                                    DMatch* __restrict__ dst, uint8_t* __restrict__ drop_in) {
    const int64_t k = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;   // one warp per pair
    const int lane = threadIdx.x & 31;
    if (k >= n) return;
    const int64_t p = order[k];
    const int64_t s0 = off_s[k], cnt = (k + 1 < n ? off_s[k + 1] : *total) - s0, d0 = off_in[p];
    for (int64_t i = lane; i < cnt; i += 32) dst[d0 + i] = src[s0 + i];
    if (lane == 0) drop_in[p] = drop_s[k];
}

cudaError_t launch_reorder(const DMatch* src, const int64_t* off_s, const int64_t* total, const int64_t* order,
                           const uint8_t* drop_s, int64_t n, int64_t* cnt_tmp, int64_t* off_in, DMatch* dst,
                           uint8_t* drop_in, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    reorder_counts_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(off_s, total, order, n, cnt_tmp);
    reorder_scan_kernel<<<1, 1024, 0, s>>>(cnt_tmp, n, off_in);
    reorder_copy_kernel<<<static_cast<unsigned>((n * 32 + 255) / 256), 256, 0, s>>>(src, off_s, total, order, off_in, drop_s,
                                                                                    n, dst, drop_in);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ knnMatch arrays
__global__ void top2_to_arrays_kernel(const Top2* __restrict__ top2, int nq, int k, int norm, int32_t* __restrict__ nidx,
                                      float* __restrict__ dist) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nq) return;
    const Top2 t = top2[r];
    const float inf = __int_as_float(0x7f800000);
    nidx[r * k] = t.i0;
    dist[r * k] = t.i0 >= 0 ? final_distance(norm, t.d0) : inf;
    if (k > 1) {
        nidx[r * k + 1] = t.i1;
        dist[r * k + 1] = t.i1 >= 0 ? final_distance(norm, t.d1) : inf;
    }
}

cudaError_t launch_top2_to_arrays(const Top2* top2, int nq, int k, int norm, int32_t* nidx, float* dist, cudaStream_t s) {
    if (nq == 0) return cudaSuccess;
    top2_to_arrays_kernel<<<(nq + 255) / 256, 256, 0, s>>>(top2, nq, k, norm, nidx, dist);
    return cudaGetLastError();
}

}  // namespace sfm
