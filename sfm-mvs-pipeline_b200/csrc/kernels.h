// kernels.h — host-side launcher prototypes of every hand-written kernel in csrc/.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "common.cuh"

namespace sfm {

// ---- knn_simt.cu
constexpr int kSimtRowsPerUnit = 128;
constexpr int kF32RowsPerUnit = 32;
cudaError_t launch_knn2_hamming_popc(const uint8_t* bank, const PairDesc* pairs, const int64_t* unit_prefix,
                                     int n_pairs, int64_t n_units, Top2* out, cudaStream_t s);
cudaError_t launch_knn2_l2_u8_dp4a(const uint8_t* bank, const int32_t* norm2, const PairDesc* pairs,
                                   const int64_t* unit_prefix, int n_pairs, int64_t n_units, Top2* out, cudaStream_t s);
cudaError_t launch_knn2_l2_f32(const float* bank, int cols, const PairDesc* pairs, const int64_t* unit_prefix,
                               int n_pairs, int64_t n_units, Top2* out, cudaStream_t s);

// ---- knn_l2_tc.cu  (tcgen05 / TMA / TMEM)
constexpr int kTcRowsPerUnit = 128;
struct TcLaunchInfo { int grid; int smem_bytes; };
// tmap: CUtensorMap (128 B, 64-B aligned, passed by value as __grid_constant__) over the u8 bank
cudaError_t launch_knn2_l2_u8_tc(const void* tmap_a_host, const void* tmap_b_host, const int32_t* ckey,
                                 const int32_t* norm2, const PairDesc* pairs, const int64_t* unit_prefix,
                                 int n_pairs, int64_t n_units, Top2* out, int sm_count, int slabs, cudaStream_t s);

// ---- knn_l2_tcv.cu  (tcgen05, value-only epilogue; needs every |b|^2 <= kExtMaxNorm2)
// What the fused ratio-test bound of the epilogue needs: |a|^2 of every bank row, the |b|^2 range per 256-row bank block
// (norm-less variant), the ratio, and the list the surviving staging rows are appended to.
struct TcvFuse {
    const int32_t* norm2;
    const int32_t* blk_min;
    const int32_t* blk_max;
    int32_t* need_list;          // capacity = staged rows of the batch
    int* need_count;
    double ratio;
    // in-kernel unit end and re-rank (norm-less variant, 8 epilogue warps): four extra "finishing" warps of the CTA merge the
    // epilogue warps' keys at the end of a unit, apply the ratio bound, re-rank the surviving rows exactly (refine_dot_row) and
    // append them to done_list; rows they cannot decide go to bf_list, rows they have no time for to need_list.
    // refine = 0: the epilogue warps finish the unit themselves and all survivors go to need_list for the post pass.
    int refine;                  // 1: divert when behind (sparse lists), 2: never divert (dense lists), 3: test mode, see knn_l2_tcv.cu
    int uniform_units;           // > 0: every pair of the launch has this many 128-row units (pair of a unit = a division, no search)
    const uint8_t* bank;
    int32_t* done_list;
    int* done_count;
    int32_t* bf_list;
    int* bf_count;
    unsigned long long* stats;
};
cudaError_t launch_knn2_l2_u8_tcv(const void* tmap_a_host, const void* tmap_b_host, const void* tmap_e_host,
                                  const PairDesc* pairs, const int64_t* unit_prefix, int n_pairs, int64_t n_units,
                                  Top2* out, int32_t* aux, int sm_count, int layout, int tile_rows, int issuers, int chunk_rows,
                                  const TcvFuse& fz, cudaStream_t s);

// ---- knn_l2_tf32.cu  (tcgen05 kind::tf32, 3xTF32 candidate search for non-integer float descriptors)
cudaError_t launch_knn2_l2_f32_tc3(const void* tmaps /* 5 CUtensorMap: hi_a, lo_a, hi_b, lo_b, ext */, const PairDesc* pairs,
                                   const int64_t* unit_prefix, int n_pairs, int64_t n_units, Top2* out, float* aux,
                                   int sm_count, cudaStream_t s);

// ---- post.cu
cudaError_t launch_pack_f32_to_u8(const float* src, size_t src_stride_elems, int n_rows, int cols,
                                  const int32_t* valid_in_block, uint8_t* dst, int* not_integer_flag, cudaStream_t s);
cudaError_t launch_expand_bits(const uint8_t* bank32, int64_t padded_rows, const int32_t* valid_in_block, uint8_t* bits,
                               int32_t* norm2, int32_t* ckey, cudaStream_t s);
cudaError_t launch_zero_padding(void* bank, int row_bytes, int64_t padded_rows, const int32_t* valid_in_block, cudaStream_t s);
cudaError_t launch_norms_ckeys(const uint8_t* bank, int64_t padded_rows, const int32_t* row_valid_end /*per 256-row block*/,
                               int32_t* norm2, int32_t* ckey, int8_t* ext, int* max_norm2 /* [0] max, [1] min over valid rows */,
                               int32_t* blk_min, int32_t* blk_max /* |b|^2 range per 256-row block, pre-initialised */,
                               cudaStream_t s);
struct FilterParams {
    int norm;            // SFM_NORM_*
    int k;               // 1 or 2
    double ratio;
    int cross_check;
    int distinct;
};
// Everything the filter passes need.  Staging rows of pair p start at out_prefix[p] (multiple of 256);
// per-train-row side arrays (reverse knn for cross-check, best-match counters for distinct) of pair p
// start at t_prefix[p] (multiple of 256).
struct FilterArgs {
    const Top2* top2;
    const Top2* rev;             // cross-check: knn of train rows against query rows (top-1 used)
    const PairDesc* pairs;
    const int64_t* out_prefix;   // n_pairs + 1
    const int64_t* t_prefix;     // n_pairs + 1
    int n_pairs;
    int64_t staged_rows;         // out_prefix[n_pairs]
    FilterParams fp;
    int32_t* train_cnt;          // distinct: zeroed by the caller
    const int32_t* blk_pair;     // pair of every 256-row staging block (launch_block_pairs), or null: binary search
    // sparse form (value-only tcgen05 path, see post.cu): one keep bit per staging row, set for rows of need_list only
    uint32_t* keep_bits;         // null: dense form (every staging row carries a Top2 record)
    const int32_t* need_list;    // rows re-ranked by the post pass
    const int* need_count;
    const int32_t* done_list;    // rows re-ranked by the refine warps inside the knn kernel (may be null)
    const int* done_count;
    const int32_t* bf_list;      // rows finished by brute_force_rows_kernel (may be null)
    const int* bf_count;
};
// tcgen05 path only: tighten the provisional second neighbour (see refine_second_kernel in post.cu)
struct RefineArgs {
    Top2* top2;
    const PairDesc* pairs;
    const int64_t* out_prefix;   // n_pairs + 1
    int n_pairs;
    int64_t staged_rows;
    const uint8_t* bank;         // u8 bank, 128-byte rows
    const int32_t* norm2;
    int all_rows;                // 1: every row (raw knnMatch output), 0: only rows that can pass the ratio test
    double ratio;
    int hamming;                 // bank = 32-byte ORB rows, distances by __popc (no sqrt in the provisional test)
    // norm-less value-only path (refine_dot): fifth-best chunk maximum per staged row, |b|^2 range per 256-row bank block
    const int32_t* aux;
    const int32_t* blk_min;
    const int32_t* blk_max;
    unsigned long long* stats;   // [0] rows re-ranked, [1] rows brute-forced
    int chunk_rows;              // train rows per candidate chunk of the value-only kernels (32 or 64)
    // refine_dot: rows whose answer needs the whole train image are queued here and finished by brute_force_rows_kernel
    // (one CTA per row) instead of stalling the warp that found them
    int32_t* bf_list;            // staged row numbers, capacity = staged_rows
    int* bf_count;
    int32_t* need_list;          // refine_dot pass 1 -> pass 2: staged rows that survived the quick reject (capacity = staged_rows)
    int* need_count;
    int32_t* pair_nb;            // [2 * n_pairs] |b|^2 range of the train image of every pair of the batch
    const int32_t* blk_pair;     // as in FilterArgs
};
// pair index of every 256-row staging block: one binary search per block instead of one per thread and post kernel
cudaError_t launch_block_pairs(const int64_t* out_prefix, int n_pairs, int64_t n_blocks, int32_t* blk_pair, cudaStream_t s);
cudaError_t launch_refine_second(const RefineArgs& a, cudaStream_t s);
// value-only tcgen05 path: (chunk, D) pairs -> exact Top2 for rows that can pass the ratio test (see post.cu)
cudaError_t launch_refine_value(const RefineArgs& a, cudaStream_t s);
// norm-less value-only path: four candidate chunks + bounds -> exact best match and a decided ratio test (see post.cu)
cudaError_t launch_refine_dot(const RefineArgs& a, cudaStream_t s);
// sparse form: keep decision of the rows of need_list -> keep bits (+ the distinct filter's two passes)
cudaError_t launch_mark_keep(const FilterArgs& a, cudaStream_t s);
cudaError_t launch_count_keep_bits(const uint32_t* keep_bits, int64_t n_blocks, int32_t* chunk_counts, cudaStream_t s);
cudaError_t launch_compact_keep_bits(const FilterArgs& a, const int64_t* chunk_excl, const int64_t* pair_offsets,
                                     const uint8_t* pair_dropped, DMatch* out, int64_t out_capacity, int* overflow_flag,
                                     cudaStream_t s);
// pass 1 (only with distinct): count how often each train row is the best match of a kept query row
cudaError_t launch_filter_mark(const FilterArgs& a, cudaStream_t s);
// pass 2: number of surviving matches per 256-row chunk
cudaError_t launch_filter_count(const FilterArgs& a, int32_t* chunk_counts, cudaStream_t s);
// scan: chunk counts -> chunk offsets; per-pair counts with min_match_count applied -> absolute pair
// offsets continuing *running_total (device scalar, updated), dropped flags
cudaError_t launch_scan_offsets(const int32_t* chunk_counts, int64_t n_chunks, const int64_t* out_prefix, int n_pairs,
                                int min_match_count, int64_t* chunk_excl /*n_chunks+1*/, int64_t* pair_counts_tmp,
                                int64_t* pair_offsets /*n_pairs*/, uint8_t* pair_dropped, int64_t* running_total,
                                cudaStream_t s);
// pass 3: ordered compaction into the DMatch output (ascending queryIdx inside every pair)
cudaError_t launch_compact(const FilterArgs& a, const int64_t* chunk_excl, const int64_t* pair_offsets,
                           const uint8_t* pair_dropped, DMatch* out, int64_t out_capacity, int* overflow_flag,
                           cudaStream_t s);
// float path: hi/lo split + norms + norm K-step rows; exact fp32 re-rank with certificate
cudaError_t launch_f32_split(const float* bank, int64_t padded_rows, const int32_t* valid_in_block, float* hi, float* lo,
                             float* fnorm2, float* ext, int* max_norm_bits, cudaStream_t s);
struct RefineF32Args {
    Top2* top2;
    const float* aux;            // fifth-best chunk maximum per staged row
    const PairDesc* pairs;
    const int64_t* out_prefix;
    int n_pairs;
    int64_t staged_rows;
    const float* bank;           // fp32 bank, 128 floats per row
    const float* fnorm2;
    float nb_max;                // largest |b|^2 of the bank (error bound)
    const int32_t* blk_pair;     // as in FilterArgs (staging blocks of THIS pass), or null
    int all_rows;
    unsigned long long* stats;   // [0] rows re-ranked, [1] rows brute-forced (certificate failed)
    int swap_roles;              // cross-check pass: train rows query the left image (rows laid out by t_prefix)
    double ratio;
};
cudaError_t launch_refine_f32(const RefineF32Args& a, cudaStream_t s);
// ---- homography.cu: SfM::calculateHomography (SfM.cpp:599-637) on the device-resident match lists
struct HomographyArgs {
    const float2* keypoints;     // KeyPoint.pt of every bank row (same row numbering as the descriptor bank)
    const int64_t* row0;         // [2 * n_pairs]: first bank row of the left / right image of every pair
    const DMatch* matches;       // compacted lists of the last run, input pair order
    const int64_t* pair_offsets; // [n_pairs] start of every list
    const int64_t* total;        // end of the last list
    const uint8_t* dropped;      // pairs removed by min_match_count (may be null)
    int n_pairs;
    const double* thresholds;    // reprojection threshold in pixels: one value, or one per pair
    int n_thresholds;
    int max_iters;               // hypotheses per pair (cv::findHomography default maxIters = 2000), <= 16384
    double confidence;           // cv::findHomography default 0.995; >= 1 evaluates every hypothesis
    uint64_t seed;
    int refine;                  // 1: least-squares re-estimation on the consensus set, mask of the refined model (cv::findHomography)
    int32_t* inliers;            // out: inlier count cv::findHomography's mask would hold, -1 = no homography (< 4 matches / dropped)
    int32_t* ransac_inliers;     // out: consensus size of the best minimal model (may be null)
    int32_t* best_hyp;           // out: hypothesis number that produced it (may be null)
};
cudaError_t launch_homography_ransac(const HomographyArgs& a, cudaStream_t s);
// ---- sift.cu: feature extraction, SfM::extractFeatures (SfM.cpp:577-597) with cv::SIFT (PhotogrammetrieCli.cpp:342-357)
struct SiftParams {
    int n_layers;                // nOctaveLayers
    int n_features;              // nfeatures (retainBest), 0 = all
    double contrast_threshold, edge_threshold, sigma;
};
constexpr int kSiftMaxOctaves = 16;
struct SiftWorkspace;            // device buffers of the stage: grey image, pyramid, candidate / keypoint lists, descriptors
SiftWorkspace* sift_workspace_create();
void sift_workspace_destroy(SiftWorkspace* w);
// One grey image (host memory) -> keypoints (sorted, deduplicated, input-image coordinates) and 128-byte descriptors in the
// workspace's device buffers.  Synchronises the stream once at the end (the keypoint count is needed on the host).
// counts_out: [0] DoG extrema, [1] keypoints before removeDuplicatedSorted, [2] keypoints.
cudaError_t sift_extract(SiftWorkspace* w, const uint8_t* gray, int rows, int cols, size_t step, const SiftParams& prm,
                         int max_keypoints, int sm_count, cudaStream_t s, int* n_keypoints, int* n_launches, int counts_out[3],
                         std::string* err);
const void* sift_keypoints_device_raw(const SiftWorkspace* w);       // sift::Keypoint[n] = sfm_keypoint[n]
const uint8_t* sift_descriptors_device(const SiftWorkspace* w);      // n x 128
// device time of the last extraction (CUDA events on the stream): pyramid construction / whole image; algorithmic HBM bytes of the pyramid
void sift_last_profile(const SiftWorkspace* w, double* pyramid_ms, double* total_ms, double* pyramid_bytes);
// test hook: geometry of the pyramid of the last extraction (float offsets of level 0 of every octave) and its base pointer
int sift_pyramid_geometry(const SiftWorkspace* w, int* n_layers, int* widths, int* heights, int64_t* offsets, const float** base);
cudaError_t launch_keypoint_xy(const void* keypoints, int n, float2* xy, cudaStream_t s);
// ---- orb.cu: feature extraction with cv::ORB::create(featureLimit) (PhotogrammetrieCli.cpp:347-348)
struct OrbWorkspace;
OrbWorkspace* orb_workspace_create();
void orb_workspace_destroy(OrbWorkspace* w);
// One grey image (host memory) -> keypoints ordered by (level, y, x) in input-image coordinates and 32-byte descriptors in the
// workspace's device buffers; synchronises the stream once at the end.
cudaError_t orb_extract(OrbWorkspace* w, const uint8_t* gray, int rows, int cols, size_t step, int n_features, int max_keypoints,
                        cudaStream_t s, int* n_keypoints, int* n_launches, std::string* err);
const void* orb_keypoints_device_raw(const OrbWorkspace* w);         // sfm_keypoint[n]
const uint8_t* orb_descriptors_device(const OrbWorkspace* w);        // n x 32
float orb_last_ms(const OrbWorkspace* w);
// test hook: per-pixel map of the last extraction; returns the element size (1 or 4 bytes) or -1
int orb_level_map(const OrbWorkspace* w, int what, int level, const void** ptr, int* width, int* height);
// schedule order -> input pair order (pipelined host path)
cudaError_t launch_reorder(const DMatch* src, const int64_t* off_s, const int64_t* total, const int64_t* order,
                           const uint8_t* drop_s, int64_t n, int64_t* cnt_tmp, int64_t* off_in, DMatch* dst,
                           uint8_t* drop_in, cudaStream_t s);
// raw knn result -> cv::batchDistance-shaped arrays (sfm_knn_match)
cudaError_t launch_top2_to_arrays(const Top2* top2, int nq, int k, int norm, int32_t* nidx, float* dist, cudaStream_t s);

}  // namespace sfm
