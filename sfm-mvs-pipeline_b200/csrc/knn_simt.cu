// knn_simt.cu — CUDA-core knn(k=2) kernels.
//
//  * knn2_hamming_popc : ORB / NORM_HAMMING.  One query descriptor (256 bit) per thread in registers,
//    train tiles staged in shared memory with 128-bit cp.async, XOR + __popc, running top-2 per thread,
//    train range split over the CTA's warp groups and merged through shared memory.
//  * knn2_l2_u8_dp4a   : exact u8 squared-L2 on CUDA cores (__dp4a).  On-GPU cross-check of the tcgen05
//    kernel and the engine used when SFM_ENGINE_SIMT is requested.
//  * knn2_l2_f32       : non-integer float descriptors (never produced by the reference pipeline's
//    SIFT/ORB, kept so that knnMatch on arbitrary CV_32F data is not refused): fp32 sum of squared
//    differences, the same formulation cv::batchDistance uses.
//
// Replaces cv::batchDistance under BFMatcher::knnMatch (call sites UnorderedFeatureMatchingStrategy.cpp:51,
// VideoFeatureMatchingStrategy.cpp:62, GridFeatureMatchingStrategy.cpp:105).  Semantics: ascending
// distance, ties -> lowest trainIdx for rank 1 and 2: every thread scans train rows in
// ascending order and replaces only on strict '<'.
#include "common.cuh"
#include "kernels.h"

namespace sfm {

constexpr int kSimtThreads = 128;   // query rows per unit (CTA)

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ------------------------------------------------------------------------------------------------
// Generic driver: a unit = (pair, block of 128 query rows).  ROW_BYTES per descriptor, TILE_ROWS train
// rows per smem stage.  Dist functor computes the distance of the thread's query to one smem row.
// ------------------------------------------------------------------------------------------------
template <int ROW_BYTES, int TILE_ROWS, class Traits>
__device__ __forceinline__ void knn2_simt_body(const uint8_t* __restrict__ bank, const int32_t* __restrict__ norm2,
                                               const PairDesc* __restrict__ pairs, const int64_t* __restrict__ unit_prefix,
                                               int n_pairs, Top2* __restrict__ out) {
    constexpr int kVecPerRow = ROW_BYTES / 16;
    constexpr int kTileBytes = TILE_ROWS * ROW_BYTES;
    __shared__ __align__(16) uint8_t tile[2][kTileBytes];
    __shared__ int32_t tile_norm[2][TILE_ROWS];

    const int64_t unit = blockIdx.x;
    const int p = find_segment(unit_prefix, n_pairs, unit);
    const PairDesc pd = pairs[p];
    const int rb = static_cast<int>(unit - unit_prefix[p]);
    const int row = rb * kSimtThreads + threadIdx.x;

    typename Traits::Query q;
    // padded rows exist (zero descriptors) up to the next multiple of kRowAlign: always safe to read
    Traits::load_query(q, bank + (static_cast<size_t>(pd.q_row0) + row) * ROW_BYTES,
                       norm2 ? norm2[pd.q_row0 + row] : 0);

    float d0 = __int_as_float(0x7f800000), d1 = d0;
    int j0 = -1, j1 = -1;

    const int n_tiles = (pd.nt + TILE_ROWS - 1) / TILE_ROWS;
    const uint8_t* tbase = bank + static_cast<size_t>(pd.t_row0) * ROW_BYTES;
    auto issue = [&](int t, int buf) {
        const uint8_t* src = tbase + static_cast<size_t>(t) * kTileBytes;
        for (int v = threadIdx.x; v < TILE_ROWS * kVecPerRow; v += kSimtThreads)
            cp_async16(&tile[buf][v * 16], src + v * 16);
        if (norm2) {
            for (int v = threadIdx.x; v < TILE_ROWS; v += kSimtThreads)
                tile_norm[buf][v] = norm2[pd.t_row0 + t * TILE_ROWS + v];
        }
        cp_async_commit();
    };
    if (n_tiles > 0) issue(0, 0);
    for (int t = 0; t < n_tiles; ++t) {
        const int buf = t & 1;
        if (t + 1 < n_tiles) { issue(t + 1, buf ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
        __syncthreads();
        const int rows_here = min(TILE_ROWS, pd.nt - t * TILE_ROWS);
#pragma unroll 2
        for (int r = 0; r < rows_here; ++r) {
            const float d = Traits::dist(q, &tile[buf][r * ROW_BYTES], tile_norm[buf][r]);
            if (d < d1) top2_insert(d, t * TILE_ROWS + r, d0, j0, d1, j1);
        }
        __syncthreads();
    }
    if (row < pd.nq) {
        Top2 o;
        o.i0 = j0; o.i1 = j1; o.d0 = d0; o.d1 = d1;
        out[pd.out_row0 + row] = o;
    }
}

struct HammingTraits {
    struct Query { uint32_t w[8]; };
    static __device__ __forceinline__ void load_query(Query& q, const uint8_t* p, int) {
        const uint4* v = reinterpret_cast<const uint4*>(p);
        uint4 a = __ldg(v), b = __ldg(v + 1);
        q.w[0] = a.x; q.w[1] = a.y; q.w[2] = a.z; q.w[3] = a.w;
        q.w[4] = b.x; q.w[5] = b.y; q.w[6] = b.z; q.w[7] = b.w;
    }
    static __device__ __forceinline__ float dist(const Query& q, const uint8_t* row, int) {
        const uint4* v = reinterpret_cast<const uint4*>(row);      // smem broadcast, 128-bit
        const uint4 a = v[0], b = v[1];
        int s = __popc(q.w[0] ^ a.x) + __popc(q.w[1] ^ a.y) + __popc(q.w[2] ^ a.z) + __popc(q.w[3] ^ a.w) +
                __popc(q.w[4] ^ b.x) + __popc(q.w[5] ^ b.y) + __popc(q.w[6] ^ b.z) + __popc(q.w[7] ^ b.w);
        return static_cast<float>(s);
    }
};

struct L2U8Traits {
    struct Query { uint32_t w[32]; int32_t n2; };
    static __device__ __forceinline__ void load_query(Query& q, const uint8_t* p, int n2) {
        const uint4* v = reinterpret_cast<const uint4*>(p);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            uint4 a = __ldg(v + i);
            q.w[4 * i] = a.x; q.w[4 * i + 1] = a.y; q.w[4 * i + 2] = a.z; q.w[4 * i + 3] = a.w;
        }
        q.n2 = n2;
    }
    static __device__ __forceinline__ float dist(const Query& q, const uint8_t* row, int n2) {
        const uint4* v = reinterpret_cast<const uint4*>(row);
        uint32_t dot = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint4 a = v[i];
            dot = __dp4a(q.w[4 * i], a.x, dot);
            dot = __dp4a(q.w[4 * i + 1], a.y, dot);
            dot = __dp4a(q.w[4 * i + 2], a.z, dot);
            dot = __dp4a(q.w[4 * i + 3], a.w, dot);
        }
        // exact: |a|^2 + |b|^2 - 2ab <= 128*255^2 < 2^24, representable in float
        return static_cast<float>(q.n2 + n2 - 2 * static_cast<int32_t>(dot));
    }
};

__global__ void __launch_bounds__(kSimtThreads) knn2_hamming_popc(const uint8_t* bank, const PairDesc* pairs,
                                                                  const int64_t* unit_prefix, int n_pairs, Top2* out) {
    knn2_simt_body<32, 256, HammingTraits>(bank, nullptr, pairs, unit_prefix, n_pairs, out);
}

__global__ void __launch_bounds__(kSimtThreads) knn2_l2_u8_dp4a(const uint8_t* bank, const int32_t* norm2,
                                                                const PairDesc* pairs, const int64_t* unit_prefix,
                                                                int n_pairs, Top2* out) {
    knn2_simt_body<128, 64, L2U8Traits>(bank, norm2, pairs, unit_prefix, n_pairs, out);
}

// fp32 descriptors of arbitrary width `cols` (multiple of 4, <= 512): one query row per thread is too
// wide for registers, so the CTA stages 32 query rows in smem and each of its 4 warps scans a quarter of
// every train tile; lanes = query rows, partial top-2 merged through smem in train order.
constexpr int kF32Q = 32;
__global__ void __launch_bounds__(128) knn2_l2_f32(const float* __restrict__ bank, int cols,
                                                   const PairDesc* __restrict__ pairs,
                                                   const int64_t* __restrict__ unit_prefix, int n_pairs,
                                                   Top2* __restrict__ out) {
    extern __shared__ __align__(16) float sm[];
    constexpr int kTile = 32;                    // train rows per stage
    float* qs = sm;                              // [kF32Q][cols+1]
    float* ts = sm + kF32Q * (cols + 1);         // [kTile][cols]
    __shared__ float md[4][kF32Q][2];
    __shared__ int mj[4][kF32Q][2];

    const int64_t unit = blockIdx.x;
    const int p = find_segment(unit_prefix, n_pairs, unit);
    const PairDesc pd = pairs[p];
    const int rb = static_cast<int>(unit - unit_prefix[p]);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int qrow0 = rb * kF32Q;

    for (int i = threadIdx.x; i < kF32Q * cols; i += 128) {
        const int r = i / cols, c = i - r * cols;
        qs[r * (cols + 1) + c] = bank[(static_cast<size_t>(pd.q_row0) + qrow0 + r) * cols + c];
    }
    float d0 = __int_as_float(0x7f800000), d1 = d0;
    int j0 = -1, j1 = -1;
    const float* qrow = qs + lane * (cols + 1);
    for (int t0 = 0; t0 < pd.nt; t0 += kTile) {
        __syncthreads();
        const int rows_here = min(kTile, pd.nt - t0);
        for (int i = threadIdx.x; i < rows_here * cols; i += 128)
            ts[i] = bank[(static_cast<size_t>(pd.t_row0) + t0) * cols + i];
        __syncthreads();
        // warp w handles tile rows [w*8, w*8+8): train order inside a warp stays ascending
        for (int r = warp * (kTile / 4); r < min(rows_here, (warp + 1) * (kTile / 4)); ++r) {
            const float* tr = ts + r * cols;
            float s = 0.f;
            for (int c = 0; c < cols; ++c) { const float e = qrow[c] - tr[c]; s = fmaf(e, e, s); }
            if (s < d1) top2_insert(s, t0 + r, d0, j0, d1, j1);
        }
    }
    md[warp][lane][0] = d0; md[warp][lane][1] = d1;
    mj[warp][lane][0] = j0; mj[warp][lane][1] = j1;
    __syncthreads();
    if (warp == 0 && qrow0 + lane < pd.nq) {
        // merge the four partial lists lexicographically on (distance, index)
        float bd0 = __int_as_float(0x7f800000), bd1 = bd0;
        int bj0 = -1, bj1 = -1;
        for (int w = 0; w < 4; ++w)
            for (int k = 0; k < 2; ++k) {
                const float d = md[w][lane][k];
                const int j = mj[w][lane][k];
                if (j < 0) continue;
                if (d < bd0 || (d == bd0 && j < bj0) || bj0 < 0) { bd1 = bd0; bj1 = bj0; bd0 = d; bj0 = j; }
                else if (d < bd1 || (d == bd1 && j < bj1) || bj1 < 0) { bd1 = d; bj1 = j; }
            }
        Top2 o; o.i0 = bj0; o.i1 = bj1; o.d0 = bd0; o.d1 = bd1;
        out[pd.out_row0 + qrow0 + lane] = o;
    }
}

// ------------------------------------------------------------------------------------------------ launchers
cudaError_t launch_knn2_hamming_popc(const uint8_t* bank, const PairDesc* pairs, const int64_t* unit_prefix,
                                     int n_pairs, int64_t n_units, Top2* out, cudaStream_t s) {
    if (n_units == 0) return cudaSuccess;
    knn2_hamming_popc<<<static_cast<unsigned>(n_units), kSimtThreads, 0, s>>>(bank, pairs, unit_prefix, n_pairs, out);
    return cudaGetLastError();
}
cudaError_t launch_knn2_l2_u8_dp4a(const uint8_t* bank, const int32_t* norm2, const PairDesc* pairs,
                                   const int64_t* unit_prefix, int n_pairs, int64_t n_units, Top2* out, cudaStream_t s) {
    if (n_units == 0) return cudaSuccess;
    knn2_l2_u8_dp4a<<<static_cast<unsigned>(n_units), kSimtThreads, 0, s>>>(bank, norm2, pairs, unit_prefix, n_pairs, out);
    return cudaGetLastError();
}
cudaError_t launch_knn2_l2_f32(const float* bank, int cols, const PairDesc* pairs, const int64_t* unit_prefix,
                               int n_pairs, int64_t n_units, Top2* out, cudaStream_t s) {
    if (n_units == 0) return cudaSuccess;
    const size_t smem = (static_cast<size_t>(kF32Q) * (cols + 1) + 32 * static_cast<size_t>(cols)) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(knn2_l2_f32, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    knn2_l2_f32<<<static_cast<unsigned>(n_units), 128, smem, s>>>(bank, cols, pairs, unit_prefix, n_pairs, out);
    return cudaGetLastError();
}

}  // namespace sfm
