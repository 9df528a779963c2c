// ctx_internal.h — host-side state of one context (one GPU) shared by the translation units that implement the C ABI
// (sfmmatch.cu: single-GPU entry points; dist.cu: the multi-GPU group).  Not part of the public boundary.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>      // header-only: ranges show up in nsys / `ncu --nvtx`, no cost without a tool attached

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/sfmmatch.h"
#include "kernels.h"

namespace sfmhost {
using namespace sfm;

extern std::string g_create_error;

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    // growth that keeps the first `used` bytes (feature store: images are appended one by one)
    cudaError_t ensure_keep(size_t bytes, size_t used, cudaStream_t s) {
        if (bytes <= cap) return cudaSuccess;
        const size_t want = std::max(bytes, cap + cap / 2);
        void* q = nullptr;
        cudaError_t e = cudaMalloc(&q, want);
        if (e != cudaSuccess) return e;
        if (p && used) {
            e = cudaMemcpyAsync(q, p, used, cudaMemcpyDeviceToDevice, s);
            if (e == cudaSuccess) e = cudaStreamSynchronize(s);
            if (e != cudaSuccess) { cudaFree(q); return e; }
        }
        if (p) cudaFree(p);
        p = q; cap = want;
        return cudaSuccess;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return static_cast<T*>(p); }
};
struct PinBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMallocHost(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return static_cast<T*>(p); }
};

inline int64_t pad_rows(int64_t n) { return (n + kRowAlign - 1) / kRowAlign * kRowAlign; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// Device-resident descriptor bank.  Image i occupies padded rows [row0[i], row0[i] + pad(n_rows[i])).
struct Bank {
    int n_images = 0, cols = 0, depth = 0;
    bool u8_valued = false;          // d_u8 holds the descriptors (ORB bytes, or u8-valued SIFT)
    bool have_f32 = false;           // d_f32 holds float descriptors (non-integer data only)
    std::vector<int32_t> n_rows;
    std::vector<int64_t> row0;
    int64_t padded_rows = 0;
    DevBuf d_kp;                     // KeyPoint.pt (float2) per bank row, sfm_keypoints_upload
    bool have_kp = false;
    DevBuf d_blkmin, d_blkmax;       // |b|^2 range per 256-row block (norm-less value-only path)
    int32_t nb_min = 0, nb_max = 0;  // |b|^2 range over the valid rows of the bank
    DevBuf d_u8, d_f32, d_norm2, d_ckey, d_valid, d_ext, d_bits;     // d_bits: ORB bits expanded to bytes (256 B rows)
    alignas(64) CUtensorMap tmap_a, tmap_b, tmap_e;
    bool ext_ok = false;             // every |b|^2 <= kExtMaxNorm2: the value-only tcgen05 kernel may be used
    // non-integer float descriptors, 128 wide: hi/lo split for the 3xTF32 tcgen05 kernel
    DevBuf d_fhi, d_flo, d_fnorm, d_fext;
    alignas(64) CUtensorMap tmaps_f[5];   // hi (A box), lo (A box), hi (B box), lo (B box), norm rows
    bool f_tc_ok = false;
    float f_nb_max = 0.f;
    bool have_tmap = false;
    void release() { d_kp.release(); d_blkmin.release(); d_blkmax.release(); d_u8.release(); d_f32.release(); d_norm2.release(); d_ckey.release(); d_valid.release(); d_ext.release(); d_bits.release();
                     d_fhi.release(); d_flo.release(); d_fnorm.release(); d_fext.release(); }
};

// NVTX range for the phases of the stage (bank upload, enqueue, collect, knnMatch, homography)
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

// Per-batch device buffers, two sets: the knn kernel of batch i + 1 runs on the compute stream while the post kernels
// (exact re-rank, filters, scan, compaction) of batch i run on the post stream.
struct BatchSlot {
    DevBuf top2, rev;                // raw top-2 / candidate records per staging row (rev: roles swapped, cross-check)
    DevBuf aux, aux_rev;             // fifth-best chunk maximum per staging row (norm-less and 3xTF32 paths)
    DevBuf need, bf, done;           // value-only path: rows that survived the fused ratio bound (for the post pass); rows that need
                                     // the whole train image; rows re-ranked by the refine warps inside the knn kernel
    DevBuf keep;                     // value-only path: one keep bit per staging row
    DevBuf counters;                 // [0] brute-force queue length, [1] need-list length, [2] done-list length
    DevBuf blk_pair;                 // pair index of every 256-row staging block
    DevBuf chunk_counts, chunk_excl, pair_counts, pair_nb, train_cnt;
    cudaEvent_t knn_done = nullptr, post_done = nullptr;
    void release() {
        DevBuf* all[] = {&top2, &rev, &aux, &aux_rev, &need, &bf, &done, &keep, &counters, &blk_pair, &chunk_counts, &chunk_excl, &pair_counts, &pair_nb, &train_cnt};
        for (DevBuf* b : all) b->release();
        if (knn_done) cudaEventDestroy(knn_done);
        if (post_done) cudaEventDestroy(post_done);
        knn_done = post_done = nullptr;
    }
};

struct RunState {                    // what collect() needs from the last enqueue
    bool valid = false;
    int64_t n_pairs = 0;
    int64_t total_query_rows = 0;
    std::vector<int32_t> pairs;      // kept for the automatic capacity retry
    sfm_opts opts{};
};

}  // namespace sfmhost

using namespace sfmhost;      // internal header: the two ABI structs below are global (extern "C" handles)

struct sfm_result {
    int64_t n_pairs = 0;
    PinBuf offsets, matches, dropped;
    sfm_ctx* owner = nullptr;
};

struct sfm_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    std::mutex mu;
    EncodeTiledFn encode = nullptr;
    Bank bank, scratch;
    // workspace
    DevBuf d_pairs, d_rev_pairs, d_unit_prefix, d_rev_unit_prefix, d_out_prefix, d_t_prefix;
    BatchSlot slot[2];
    cudaStream_t post_stream = nullptr;
    DevBuf d_pair_offsets, d_dropped;
    DevBuf d_scalars;                // [0..7] int64 running_total, [8..11] int overflow, [12..15] int not_integer
    DevBuf d_out, d_knn;
    DevBuf d_hom;                    // homography stage: row0[2n] int64 | thresholds | inliers | best hypothesis
    DevBuf d_out2, d_pair_offsets2, d_dropped2, d_order, d_cnt_tmp;   // reorder targets of the pipelined host path
    cudaStream_t copy_stream = nullptr;                               // uploads of the pipelined host path
    std::vector<cudaEvent_t> group_ev;                                // image group g is resident + packed
    int64_t out_capacity = 0;
    PinBuf h_meta, h_stage[2], h_scalars, h_knn, h_valid;
    std::vector<sfm_result*> result_pool;   // recycled results (pinned buffers are expensive to allocate)
    cudaEvent_t stage_ev[2] = {nullptr, nullptr};
    cudaEvent_t meta_ev = nullptr;   // h_meta may be rewritten once this has fired
    cudaEvent_t valid_ev = nullptr;  // same for h_valid
    RunState run;
    int64_t stat_launches = 0, stat_h2d = 0, stat_d2h = 0;
    // optional per-kernel timing of the last enqueue (sfm_set_profiling)
    bool profiling = false;
    std::vector<cudaEvent_t> prof_ev;     // triples per batch: knn begin, knn end, post end
    int prof_used = 0;
    size_t staging_budget_rows = 0;
    int tcv_layout_run = 12;
    int tcv_spread_div = 2;          // norm-less variant allowed when (max - min |b|^2) * div <= max (SFM_TCV_SPREAD_DIV):
                                     // measured a win at the 36 % spread of the 200-image SURVEY 8d bank, cv::SIFT has ~1 %
    int32_t prev_nb_min = 0, prev_nb_max = 1;   // |b|^2 range of the previous bank of this context (pipelined path)
    int tcv_normless = 1;            // norm-less variant of the value-only kernel: SFM_TCV_NORMLESS = 0 never | 1 auto | 2 always
    int tcv_chunk = 0;               // train rows per candidate chunk of the value-only kernels: 0 = adaptive, SFM_TCV_CHUNK = 32 | 64
    int tcv_chunk_run = 64;
    // adaptive choice between (norm-less kernel, 64-row chunks) and (kernel with the norm K-step, 32-row chunks): the share
    // of query rows the previous run had to re-rank exactly.  Sparse matches (C3: 0.5 %) favour the first, dense matches
    // (C4 grid neighbours: 7 %) the second, whose refine pass reads a quarter of the train rows per re-ranked row.
    bool dense_matches = false;
    bool tune_pending = false;
    int64_t tune_rows = 0;
    PinBuf h_tune;
    cudaEvent_t tune_ev = nullptr;
    // feature extraction stage (sift.cu): images extracted so far, device-resident
    SiftWorkspace* sift = nullptr;
    OrbWorkspace* orb = nullptr;
    int feat_cols = 0;               // descriptor bytes of the images extracted so far: 128 (SIFT) / 32 (ORB), 0 = none yet
    DevBuf feat_kp, feat_desc;       // sfm_keypoint[total], u8[total][feat_cols]
    std::vector<int64_t> feat_off{0};
    int feat_counts[3] = {0, 0, 0};
    int tcv_issuers = 2;             // MMA-issuing warps of the value-only kernel (SFM_TCV_ISSUERS = 1 | 2)
    int tcv_layout = 0;              // epilogue organisation of the value-only kernel (10 * parity + halves): 0 = auto,
                                     // SFM_TCV_LAYOUT = 12 | 14 | 21 forces
    bool tcv_inkernel_refine = true; // SFM_TCV_INKERNEL_REFINE = 0: survivors of the fused ratio bound go to the post pass instead of
                                     // the knn CTA's own finishing warps
    bool tcv_divert_test = false;    // SFM_TCV_DIVERT_TEST = 1 (tests): every finishing warp re-ranks one row per unit and diverts the rest
    int tcv_backpressure = 1;        // SFM_TCV_BACKPRESSURE: 1 = on dense lists the epilogue waits for the finishing warps, 0 = a warp that
                                     // falls behind always diverts its rows to the post pass, 2 = the epilogue always waits
    int knn_grid_limit = 0;          // > 0: the persistent knn kernels launch at most this many CTAs (dist from-host path: SMs left
                                     // free for the NCCL broadcasts of the chunks still in flight)
    int min_batches = 6;             // a long pair list is cut into at least this many batches (SFM_MIN_BATCHES): post kernels of
                                     // batch i overlap the knn kernel of batch i + 1
    void* dist = nullptr;            // multi-GPU group membership (sfmhost::DistState, csrc/dist.cu)
};

namespace sfmhost {

inline int fail(sfm_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg; else g_create_error = msg;
    return code;
}
#define CU_TRY(ctx, expr)                                                                            \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess)                                                                       \
            return fail(ctx, SFM_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));      \
    } while (0)

// ---- implemented in sfmmatch.cu, used by dist.cu
// Optional processing schedule of enqueue_impl (pipelined host path): order[k] = input index of the k-th scheduled
// pair, avail[k] = index of the event (non-decreasing in k) the batch containing k has to wait for.
struct Schedule {
    const int64_t* order = nullptr;
    const int* avail = nullptr;
    const cudaEvent_t* events = nullptr;
    int n_groups = 0;                 // > 0: number of availability groups (batches of the earlier ones may leave SMs free)
};
int make_tmaps(sfm_ctx* c, Bank& b);
int bank_layout(sfm_ctx* c, Bank& b, int n_images, const int32_t* n_rows, int cols, int depth);
int bank_finish(sfm_ctx* c, Bank& b);
int bank_upload_host(sfm_ctx* c, Bank& b, int n_images, const void* const* rows, const int32_t* n_rows, int cols,
                     const size_t* step_bytes, int depth);
int enqueue_impl(sfm_ctx* c, const int32_t* pairs_in, int64_t n_pairs, const sfm_opts* o, const Schedule* sched = nullptr);
int collect_impl(sfm_ctx* c, sfm_result** out);
void dist_state_destroy(sfm_ctx* c);      // dist.cu
int from_host_impl(sfm_ctx* c, int n_images, const void* const* rows, const int32_t* n_rows, int cols,
                   const size_t* step_bytes, int depth, const int32_t* pairs, int64_t n_pairs, const sfm_opts* o,
                   sfm_result** out);

}  // namespace sfmhost
