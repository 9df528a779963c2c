// common.cuh — shared device-side types and PTX wrappers (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sfm {

// Rows of every image are padded to a multiple of kRowAlign inside the device bank, with zero
// descriptors, so that a TMA box / smem tile never straddles two images.
constexpr int kRowAlign = 256;
// Sentinel column key for padded train rows in the tcgen05 kernel: larger than any real key
// (max real key = 128*255^2*256 + 255 = 0x7F0100FF), still < 2^31.
constexpr int32_t kSentinelKey = 0x7FFFFF00;

// Value-only tcgen05 kernel (knn_l2_tcv.cu): every bank row carries 32 signed "norm digits" e_k so that one extra
// u8 x s8 K-step against the constant weight row W = {255 x12, 1 x4, 255 x12, 1 x4} adds  -(||b||^2 >> 1)  to the
// accumulator:  sum_k W_k e_k = -(||b||^2 >> 1).  Usable iff every ||b||^2 <= kExtMaxNorm2 (SIFT: ~2.6e5).
constexpr int kExtBytes = 32;
constexpr int32_t kExtMaxNorm2 = 1500000;
// padded (invalid) train rows get all digits = -128: D = 0 - kExtPadValue, below every valid D (>= -750000)
constexpr int32_t kExtPadValue = 24 * 255 * 128 + 8 * 128;

// One image pair of a batch, device-resident.  left <-> query, right <-> train (Scene.h:47-51).
struct PairDesc {
    int32_t q_row0;     // first bank row of the query (left) image, multiple of kRowAlign
    int32_t nq;         // real query rows
    int32_t t_row0;     // first bank row of the train (right) image
    int32_t nt;         // real train rows
    int64_t out_row0;   // first staging row of this pair (multiple of kRowAlign)
};

// Raw knn(k=2) result of one query row: what cv::batchDistance(..., K=2) yields.
// d = squared L2 (exact integer held in float when the data is u8-valued) or Hamming count.
struct __align__(16) Top2 {
    int32_t i0, i1;     // train indices, -1 = missing neighbour
    float   d0, d1;
};

struct __align__(16) DMatch {   // byte-compatible with cv::DMatch / sfm_dmatch
    int32_t queryIdx, trainIdx, imgIdx;
    float   distance;
};

// largest p with prefix[p] <= x, prefix ascending, prefix[0] == 0, n entries (+1 sentinel end)
__device__ __forceinline__ int find_segment(const int64_t* __restrict__ prefix, int n, int64_t x) {
    int lo = 0, hi = n;            // invariant: prefix[lo] <= x < prefix[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (__ldg(prefix + mid) <= x) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ void top2_insert(float d, int j, float& d0, int& j0, float& d1, int& j1) {
    // callers scan j ascending and use strict '<': ties keep the lowest train index
    if (d < d0) { d1 = d0; j1 = j0; d0 = d; j0 = j; }
    else        { d1 = d;  j1 = j; }
}

// ---------------------------------------------------------------- PTX: mbarrier / TMA / tcgen05
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// non-blocking: has the phase of that parity completed?
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) { }
}
// 2-D tiled TMA load (global -> smem), completion on an mbarrier
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
// 1-D bulk copy (global -> smem), size multiple of 16 B
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

// one lane of a converged warp (elect.sync): unlike `lane == 0`, ptxas then KNOWS a single thread runs the guarded
// tcgen05 instructions and drops the per-thread elect / R2UR.BROADCAST / BRA.U.ANY loop around each of them
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after()  { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], u8 x u8 -> s32
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// this warp's 32 TMEM lanes x 32 consecutive 32-bit columns -> 32 registers per thread
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
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

// UMMA shared-memory descriptor: K-major operand, 128-byte swizzle, rows of 128 B, 8-row groups
// 1024 B apart (SBO = 64 x 16 B), LBO = 1 (ignored for swizzled K-major), descriptor version 1.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);        // start address  [0,14)
    d |= static_cast<uint64_t>(1) << 16;                              // LBO            [16,30)
    d |= static_cast<uint64_t>(1024 >> 4) << 32;                      // SBO            [32,46)
    d |= static_cast<uint64_t>(1) << 46;                              // version = 1    [46,48)
    d |= static_cast<uint64_t>(2) << 61;                              // SWIZZLE_128B   [61,64)
    return d;
}
// instruction descriptor, kind::i8: D = s32, A = B = u8, both K-major, shape M x N
__host__ __device__ constexpr uint32_t umma_idesc_u8(int m, int n) {
    return (2u << 4) | (0u << 7) | (0u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
           (static_cast<uint32_t>(m >> 4) << 24);
}
// same with a signed B operand (A = u8, B = s8)
__host__ __device__ constexpr uint32_t umma_idesc_u8s8(int m, int n) {
    return umma_idesc_u8(m, n) | (1u << 10);
}
// UMMA shared-memory descriptor: K-major operand with 32-byte rows, 32-byte swizzle; 8-row groups 256 B apart.
__device__ __forceinline__ uint64_t umma_desc_sw32(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(256 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(6) << 61;                              // SWIZZLE_32B
    return d;
}

}  // namespace sfm
