/*
 * sfmmatch.h — C ABI of the B200-native pairwise descriptor matcher (libsfmmatch.so).
 *
 * This is the drop-in boundary for the feature-matching stage of brunothg/sfm-mvs-pipeline.
 * Every entry point names the reference interface it replaces (paths into the reference tree).
 * No C++ / torch types cross this boundary: plain pointers, sizes and status codes.
 *
 * Status codes: 0 = ok, < 0 = error (sfm_last_error(ctx) gives the text).  No exception
 * crosses the ABI; the C++ adapter turns a non-zero status into cv::Exception
 * so the reference's catch-branch (UnorderedFeatureMatchingStrategy.cpp:66-72) still works.
 * There is NO CPU fallback: every compute entry point fails with SFM_ERR_CUDA when no sm_100
 * device is usable.
 */
#ifndef SFMMATCH_H
#define SFMMATCH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SFM_OK              0
#define SFM_ERR_INVALID    -1   /* bad argument / shape mismatch (cv::batchDistance asserts) */
#define SFM_ERR_CUDA       -2   /* CUDA runtime/driver error, or no usable device */
#define SFM_ERR_CAPACITY   -3   /* caller buffer too small / train rows >= 2^18 (IMGIDX_ONE) */
#define SFM_ERR_UNSUPPORTED -4  /* e.g. radius match, k > 2, cross_check with k != 1 semantics */
#define SFM_ERR_STATE      -5   /* call order violated (e.g. match before bank upload) */
#define SFM_ERR_NCCL       -6   /* NCCL not loadable, or a collective failed (multi-GPU group only) */

/* cv::NormTypes values used by the reference (PhotogrammetrieCli.cpp:378, :387) */
#define SFM_NORM_L2        4
#define SFM_NORM_HAMMING   6
/* cv::Mat depth codes of the descriptor matrices (CameraShot.h:39-42) */
#define SFM_CV_8U          0
#define SFM_CV_32F         5
/* OpenCV's IMGIDX_ONE limit, surfaced by the reference as "max 262144" (PhotogrammetrieCli.cpp:430) */
#define SFM_MAX_ROWS       (1 << 18)

/* engine selector (tests / profiling): which hand-written kernel computes the distances */
#define SFM_ENGINE_AUTO    0    /* tcgen05 kernels (L2 on u8-valued 128-d data; Hamming on 256-bit descriptors) */
#define SFM_ENGINE_TENSOR  1    /* force tcgen05/TMA/TMEM kernels (Hamming: bits expanded to u8, K = 256, exact) */
#define SFM_ENGINE_SIMT    2    /* force CUDA-core kernels (dp4a L2 / popc Hamming / fp32 L2) */
#define SFM_ENGINE_TENSOR_IMAD 3 /* force the tcgen05 kernel with the packed-key (IMAD) epilogue even where the
                                   value-only tcgen05 kernel would be chosen */

/* Byte-compatible with cv::DMatch {int queryIdx, trainIdx, imgIdx; float distance;}.
 * left shot <-> queryIdx, right shot <-> trainIdx (Scene.h:47-51). */
typedef struct sfm_dmatch {
    int32_t queryIdx;
    int32_t trainIdx;
    int32_t imgIdx;
    float   distance;
} sfm_dmatch;

/* Options of one matching run.  Defaults = the reference's behaviour:
 * k=2 knnMatch + Lowe ratio 0.7 (UnorderedFeatureMatchingStrategy.cpp:51-59), cross-check off
 * (BFMatcher::create(normType), PhotogrammetrieCli.cpp:378/:387), distinct off
 * (PhotogrammetrieCli.cpp:112), min_match_count 0 here (SfM applies 20, PhotogrammetrieCli.cpp:95). */
typedef struct sfm_opts {
    int32_t norm;             /* SFM_NORM_L2 | SFM_NORM_HAMMING */
    int32_t k;                /* 2 (ratio test) ; 1 = DescriptorMatcher::match() semantics, keep all */
    double  ratio;            /* 0.7 */
    int32_t cross_check;      /* 1: mutual-nearest-neighbour filter instead of the ratio test
                                 (cv::BFMatcher(crossCheck=true) semantics) */
    int32_t distinct;         /* 1: SfM.cpp:547-564 one-to-one trainIdx filter */
    int32_t min_match_count;  /* pairs with fewer matches are dropped (SfM.cpp:566-570); 0 keeps all */
    int32_t engine;           /* SFM_ENGINE_* */
} sfm_opts;

typedef struct sfm_ctx sfm_ctx;
typedef struct sfm_result sfm_result;

/* Fill *o with the reference defaults for `norm`. */
void sfm_opts_default(sfm_opts *o, int32_t norm);

/* Library / context ------------------------------------------------------------------------
 * One context per GPU (one process per GPU under torchrun, or one context per device in a
 * single C++ process).  Replaces: construction of the matcher object,
 * PhotogrammetrieCli::configureFeatureMatcher (PhotogrammetrieCli.cpp:359-392). */
int  sfm_ctx_create(sfm_ctx **out, int device);
void sfm_ctx_destroy(sfm_ctx *ctx);
const char *sfm_last_error(const sfm_ctx *ctx);      /* ctx may be NULL: last create error */
int  sfm_device_sm_count(const sfm_ctx *ctx);
/* The CUDA stream (cudaStream_t) all work of this context is enqueued on — lets the caller
 * bracket work with CUDA events. */
void *sfm_ctx_stream(sfm_ctx *ctx);

/* Descriptor bank -----------------------------------------------------------------------------
 * Replaces: reading CameraShot::getFeatures().descriptors for every shot of the scene
 * (UnorderedFeatureMatchingStrategy.cpp:51, CameraShot.h:39-42).  rows[i] points at image i's
 * host cv::Mat data (n_rows[i] x cols, row stride step_bytes[i] >= cols*elemsize; step_bytes
 * may be NULL for dense rows).  depth SFM_CV_32F with cols=128 (SIFT) or SFM_CV_8U (ORB: cols=32;
 * SIFT already packed: cols=128).  Inputs are copied; the caller's buffers are not retained.
 * CV_32F data that is integer-valued in 0..255 (what cv::SIFT emits) is packed to u8 on the
 * device and matched exactly; other float data takes the fp32 path. */
int sfm_bank_upload(sfm_ctx *ctx, int n_images, const void *const *rows, const int32_t *n_rows,
                    int cols, const size_t *step_bytes, int cv_depth);
/* Same, but the descriptors already live in device memory of this context's GPU (e.g. a replica
 * received through ncclBroadcast): one dense buffer, image i occupying rows
 * [row_offset[i], row_offset[i]+n_rows[i]) with row stride cols*elemsize. */
int sfm_bank_upload_device(sfm_ctx *ctx, int n_images, const void *dev_rows, const int64_t *row_offset,
                           const int32_t *n_rows, int cols, int cv_depth);
int sfm_bank_info(const sfm_ctx *ctx, int *n_images, int *cols, int *is_u8_valued);
/* Device address of image `image`'s packed u8 rows inside the resident bank (valid until the next upload).  Lets a
 * multi-GPU host upload 1/N of the images per GPU and all-gather the packed bank over NVLink instead of sending the
 * whole CV_32F scene through every GPU's PCIe link. */
int sfm_bank_device_ptr(const sfm_ctx *ctx, int image, const void **dev_rows, int32_t *n_rows);

/* Pair selection --------------------------------------------------------------------------------
 * Replaces the matchPairs construction of the three strategies, chosen exactly like
 * PhotogrammetrieCli::configureFeatureMatcherStrategy (PhotogrammetrieCli.cpp:320-340):
 * feature_sequence >= 2 && feature_gridlength >= 1 -> Grid (GridFeatureMatchingStrategy.cpp:48-85),
 * feature_sequence >= 2 -> Video (VideoFeatureMatchingStrategy.cpp:43-48), else Unordered
 * (UnorderedFeatureMatchingStrategy.cpp:32-37).  Writes up to `capacity` pairs (left,right) into
 * `pairs` and the full count into *n_pairs (call with capacity 0 to size the buffer). */
int sfm_select_pairs(int n_shots, int feature_sequence, int feature_gridlength,
                     int32_t *pairs, int64_t capacity, int64_t *n_pairs);

/* The stage ----------------------------------------------------------------------------------------
 * Replaces IFeatureMatchingStrategy::calculateShotMatches (IFeatureMatchingStrategy.h:45-46) as
 * invoked from SfM::calculateShotMatches (SfM.cpp:545) plus that function's post-filters
 * (SfM.cpp:547-570): for every pair (left,right) of the list, knnMatch(k=2) of left's descriptors
 * against right's, Lowe ratio filter, optional distinct / min-match-count filters.
 * Results come back in INPUT PAIR ORDER, each list ascending by queryIdx (what OpenCV emits). */
int sfm_match_pairs(sfm_ctx *ctx, const int32_t *pairs /* n_pairs x 2 */, int64_t n_pairs,
                    const sfm_opts *opts, sfm_result **out);
/* Two-phase form used for device-side timing: enqueue runs every kernel on the context stream and
 * leaves the compacted match lists in device memory; collect copies them to pinned host memory. */
int sfm_match_pairs_enqueue(sfm_ctx *ctx, const int32_t *pairs, int64_t n_pairs, const sfm_opts *opts);
int sfm_match_pairs_collect(sfm_ctx *ctx, sfm_result **out);
/* Device-resident view of the last enqueue's result (valid until the next enqueue/upload on this context), for a
 * multi-GPU host that gathers match lists GPU-to-GPU (ncclGather / torch.distributed) instead of through host memory:
 * d_matches = total_matches sfm_dmatch records, d_pair_offsets = n_pairs int64 start offsets, d_dropped = n_pairs bytes.
 * Waits for the enqueued kernels. */
int sfm_match_pairs_device_view(sfm_ctx *ctx, const void **d_matches, const void **d_pair_offsets,
                                const void **d_dropped, int64_t *n_pairs, int64_t *total_matches);

/* sfm_bank_upload + sfm_match_pairs as ONE call = the whole of IFeatureMatchingStrategy::calculateShotMatches
 * (descriptors of every shot in host memory in, per-pair DMatch lists out).  When the descriptor matrices are
 * page-locked and 128 columns wide, the upload of image groups overlaps the matching of the pairs that are already
 * resident; results are identical to the two-call form (input pair order).  The bank stays resident afterwards. */
int sfm_match_pairs_from_host(sfm_ctx *ctx, int n_images, const void *const *rows, const int32_t *n_rows, int cols,
                              const size_t *step_bytes, int cv_depth, const int32_t *pairs, int64_t n_pairs,
                              const sfm_opts *opts, sfm_result **out);

int64_t           sfm_result_n_pairs(const sfm_result *r);
const int64_t    *sfm_result_offsets(const sfm_result *r);   /* n_pairs + 1 entries */
const sfm_dmatch *sfm_result_matches(const sfm_result *r);   /* offsets[n_pairs] entries */
const uint8_t    *sfm_result_dropped(const sfm_result *r);   /* n_pairs flags: 1 = erased by min_match_count */
void              sfm_result_free(sfm_result *r);       /* call before sfm_ctx_destroy of the producing context */

/* Multi-GPU group ------------------------------------------------------------------------------------
 * Replaces the `#pragma omp parallel for` over image pairs of the three strategies
 * (UnorderedFeatureMatchingStrategy.cpp:40, VideoFeatureMatchingStrategy.cpp:51, GridFeatureMatchingStrategy.cpp:94)
 * when more than one GPU is used: every GPU holds a replica of the descriptor bank, the pair list is dealt by cost
 * Nq*Nt, NCCL is used only to exchange the packed descriptors and to gather the match lists on participant 0
 * (SURVEY 8e).  Results on participant 0 are BYTE-IDENTICAL to a single-GPU run of the same list (same kernels per
 * pair, input pair order).  NCCL (libnccl.so.2) is opened with dlopen at first use.
 *
 * (a) one process, one thread per GPU — what the reference's single-process pipeline uses:  sfm_mgpu_*.
 *     The calls mirror sfm_bank_upload / sfm_match_pairs / sfm_match_pairs_from_host; the result is released with
 *     sfm_result_free before sfm_mgpu_destroy.
 * (b) one process per GPU (torchrun / MPI launchers): participant 0 makes an id (sfm_dist_unique_id), the launcher
 *     hands its SFM_DIST_ID_BYTES bytes to every participant by its own means, each calls sfm_dist_init on its context.
 *     sfm_dist_match_pairs* are COLLECTIVE: every participant calls them with the same pair list / scene description;
 *     *out is the gathered result on participant 0 and NULL elsewhere.  With world == 1 (or without sfm_dist_init) they
 *     behave like the single-GPU calls.
 * from_host: participant r uploads only the images sfm_dist_upload_share names, over its own PCIe link, and the packed
 * chunks are exchanged over NVLink while matching of the already-complete pairs runs (128-column descriptors, NORM_L2,
 * k = 2); for any other configuration every participant reads the WHOLE scene, so `rows` must then be valid for all
 * images on every participant. */
#define SFM_DIST_ID_BYTES 128
typedef struct sfm_mgpu sfm_mgpu;
int  sfm_mgpu_create(sfm_mgpu **out, const int *devices, int n_devices);
void sfm_mgpu_destroy(sfm_mgpu *g);
int  sfm_mgpu_device_count(const sfm_mgpu *g);
sfm_ctx *sfm_mgpu_ctx(sfm_mgpu *g, int participant);        /* per-GPU context (stats, profiling, homography stage) */
const char *sfm_mgpu_last_error(const sfm_mgpu *g);
int sfm_mgpu_bank_upload(sfm_mgpu *g, int n_images, const void *const *rows, const int32_t *n_rows, int cols,
                         const size_t *step_bytes, int cv_depth);
int sfm_mgpu_match_pairs(sfm_mgpu *g, const int32_t *pairs, int64_t n_pairs, const sfm_opts *opts, sfm_result **out);
int sfm_mgpu_match_pairs_from_host(sfm_mgpu *g, int n_images, const void *const *rows, const int32_t *n_rows, int cols,
                                   const size_t *step_bytes, int cv_depth, const int32_t *pairs, int64_t n_pairs,
                                   const sfm_opts *opts, sfm_result **out);

int sfm_dist_unique_id(uint8_t *id /* SFM_DIST_ID_BYTES */);
int sfm_dist_init(sfm_ctx *ctx, const uint8_t *id, int rank, int world);
int sfm_dist_info(const sfm_ctx *ctx, int *rank, int *world);
int sfm_dist_match_pairs(sfm_ctx *ctx, const int32_t *pairs, int64_t n_pairs, const sfm_opts *opts, sfm_result **out);
int sfm_dist_match_pairs_from_host(sfm_ctx *ctx, int n_images, const void *const *rows, const int32_t *n_rows, int cols,
                                   const size_t *step_bytes, int cv_depth, const int32_t *pairs, int64_t n_pairs,
                                   const sfm_opts *opts, sfm_result **out);
/* The deal (host arithmetic, no GPU): owner[p] = participant that matches pair p — stable sort by descending cost
 * Nq*Nt, dealt in snake order, ascending pair index inside a participant. */
int sfm_dist_assign_pairs(const int32_t *pairs, int64_t n_pairs, const int32_t *n_rows, int n_images, int world,
                          int32_t *owner);
/* Images [*first_image, *end_image) are the ones participant `rank` uploads in the from_host exchange path. */
int sfm_dist_upload_share(const int32_t *n_rows, int n_images, int world, int rank, int *first_image, int *end_image);
/* The images of a scene were extracted on different participants (the loop of SfM::extractFeatures, SfM.cpp:577-597, split
 * over the GPUs): collective exchange after which EVERY participant's feature set holds all n_images_total images in scene order
 * (then sfm_bank_from_features on each).  global_index[k] = scene position of the k-th image this participant extracted. */
int sfm_dist_features_allgather(sfm_ctx *ctx, const int32_t *global_index, int n_local, int n_images_total);
/* One process: image i is extracted on device i % n (detector: SFM_DETECTOR_SIFT with sfm_sift_opts, SFM_DETECTOR_ORB with
 * sfm_orb_opts; opts may be NULL), the sets are exchanged, every device adopts the scene as its bank + keypoint table.
 * n_keypoints[n_images] may be NULL; features are read back through sfm_features_download(sfm_mgpu_ctx(g, 0), ...). */
#define SFM_DETECTOR_SIFT 0
#define SFM_DETECTOR_ORB  1
int sfm_mgpu_extract_features(sfm_mgpu *g, int detector, int n_images, const uint8_t *const *gray, const int32_t *rows,
                              const int32_t *cols, const size_t *step_bytes, const void *opts, int32_t *n_keypoints);
/* Host-clock phases (ms) of the last collective on this participant: [0] upload + exchange enqueued / deal,
 * [1] kernels enqueued, [3] totals all-gather incl. waiting for the kernels, [4] send/recv + reorder + D2H. */
int sfm_dist_last_phases(const sfm_ctx *ctx, double *ms /* 8 */);

/* Homography stage -------------------------------------------------------------------------
 * Replaces SfM::calculateHomography (SfM.cpp:599-637): per ShotMatches of the last sfm_match_pairs* run,
 * cv::findHomography(left points, right points, cv::RANSAC, threshold, mask) -> inlier count / match count.
 * The match lists never leave the GPU for this.
 *
 * sfm_keypoints_upload: KeyPoint.pt of every descriptor row of the bank (same images, same row counts; call after the
 * bank upload).  pts[i] points at the first (x, y) float pair of image i, step_bytes[i] = distance between consecutive
 * points (8 for a packed float2 array, sizeof(cv::KeyPoint) = 28 for &keypoints[0].pt of a std::vector<cv::KeyPoint>;
 * NULL = 8).
 * sfm_homography_inlier_ratios: thresholds in pixels, one value or one per pair (SfM.cpp:617-620 derives it from
 * ransacReprojectionMatchingThreshold and the image sizes).  opts (NULL = defaults): max_iters / confidence =
 * cv::findHomography's maxIters (2000) and confidence (0.995), i.e. the budget of 4-point hypotheses per pair and the
 * early-stopping rule of OpenCV's RANSAC loop, replayed on the device (confidence >= 1: every hypothesis counts);
 * refine = 1: like cv::findHomography, re-estimate the model on the consensus set (normalised least squares) and
 * report the matches within the threshold of THAT model; seed: same seed -> same result.
 * ratios[p] = inliers / matches, or -1 for pairs with < 4 matches or dropped by min_match_count (the reference leaves
 * ShotMatches::homographyInlierRatio at -1 there); exactly 4 matches -> 1 (OpenCV runs no RANSAC on four points).
 * inliers / ransac_inliers (consensus size of the best minimal model) / best_hypothesis may be NULL.
 * Parity with cv::findHomography is statistical (different random minimal sets): tolerance in tests/test_homography.py. */
typedef struct sfm_homography_opts {
    int32_t  max_iters;
    int32_t  refine;
    double   confidence;
    uint64_t seed;
} sfm_homography_opts;
void sfm_homography_opts_default(sfm_homography_opts *o);
int sfm_keypoints_upload(sfm_ctx *ctx, int n_images, const void *const *pts, const int32_t *n_rows,
                         const size_t *step_bytes);
int sfm_homography_inlier_ratios(sfm_ctx *ctx, const double *thresholds, int64_t n_thresholds,
                                 const sfm_homography_opts *opts, double *ratios, int32_t *inliers,
                                 int32_t *ransac_inliers, int32_t *best_hypothesis);

/* Feature extraction stage -----------------------------------------------------------------------
 * Replaces the body of SfM::extractFeatures' loop (SfM.cpp:584-590),
 *     featureDetector->detect(image, keypoints); descriptorExtractor->compute(image, keypoints, descriptors);
 * for the detector PhotogrammetrieCli.cpp:342-357 configures, cv::SIFT::create(featureLimit, 3, 0.09): Gaussian / DoG
 * pyramid, scale-space extrema, sub-pixel refinement, orientation histograms, removeDuplicatedSorted, retainBest and
 * the 4 x 4 x 8 descriptors, all on the device (csrc/sift.cu).  Keypoints come back in cv::SIFT's order (sorted by x, y, ...), in
 * input-image coordinates; descriptors are the u8 values cv::SIFT stores in its CV_32F rows.
 *
 * The extracted images accumulate in the context (image 0, 1, ... in call order) and stay device-resident;
 * sfm_bank_from_features turns them into the descriptor bank AND the keypoint table of the matching / homography
 * stages without a host round trip (it replaces sfm_bank_upload + sfm_keypoints_upload);
 * sfm_features_download copies one image's keypoints / descriptors to the host (Shot::setFeatures needs them for the
 * later pipeline stages).  Parity with cv::SIFT is a tolerance (float arithmetic with data-dependent decisions):
 * tests/_sift_compare.py states it.  With n_features > 0 the survivors of retainBest keep the sorted order (OpenCV
 * leaves them in the order std::nth_element produced; the set is the same).  Descriptors come from the detection pyramid
 * (cv::SIFT::compute on its own rebuilds it without the 2x upsampling when no keypoint lies in octave -1: DESIGN.md section 8).
 * n_octave_layers (1..8), edge_threshold and sigma are honoured; on the GPU only the reference's values (3, 10, 1.6) have
 * been exercised so far, the other values through the host build of the same per-keypoint code (tests/test_sift_core_host.py).
 * max_keypoints bounds the per-image lists (0 = 262143, the matcher's per-image limit); more -> SFM_ERR_CAPACITY. */
typedef struct sfm_keypoint {      /* cv::KeyPoint without class_id */
    float   x, y, size, angle, response;
    int32_t octave;
} sfm_keypoint;
typedef struct sfm_sift_opts {
    int32_t n_octave_layers;       /* cv::SIFT::create arguments; defaults 3, 0.04, 10, 1.6 */
    int32_t max_keypoints;
    double  contrast_threshold;
    double  edge_threshold;
    double  sigma;
    int32_t n_features;            /* cv::SIFT::create's nfeatures: 0 = keep all (default); the reference passes its
                                      feature-limit (default 10000, PhotogrammetrieCli.cpp:345,355) */
    int32_t reserved;
} sfm_sift_opts;
void sfm_sift_opts_default(sfm_sift_opts *o);
/* The reference's other detector: cv::ORB::create(featureLimit) (PhotogrammetrieCli.cpp:347-348; -Pfeature-detector=ORB,
 * run-scripts/run-orb-sequence.sh:4), detect() + compute() on the device (csrc/orb.cu): 8-level pyramid (INTER_LINEAR_EXACT),
 * FAST-9 + non-maximum suppression, Harris ranking with retainBest's tie rule per level, intensity-centroid orientation,
 * 7 x 7 Gaussian blur, rotated 256-bit BRIEF.  Device == the numpy restatement (oracle/orb_np.py), which equals cv2 (keypoint
 * SET, responses and descriptors identical, angles within 1e-3 degrees); keypoints come back ordered by (level, y, x) — cv::ORB's
 * own order is a by-product of std::nth_element.  Only n_features may differ from cv::ORB::create's defaults (the reference
 * passes nothing else); descriptors are 32 bytes, one feature set holds SIFT or ORB images, not both. */
typedef struct sfm_orb_opts {
    int32_t n_features;            /* cv::ORB::create's nfeatures (default 500); the reference passes its feature-limit */
    int32_t max_keypoints;         /* capacity of the per-image lists, 0 = 262143; retainBest keeps ties, so slightly more than
                                      n_features keypoints can come back */
    int32_t n_levels, edge_threshold, patch_size, fast_threshold;   /* 8, 31, 31, 20: other values -> SFM_ERR_UNSUPPORTED */
    float   scale_factor;          /* 1.2f */
    int32_t reserved;
} sfm_orb_opts;
void sfm_orb_opts_default(sfm_orb_opts *o);
int sfm_features_extract_orb(sfm_ctx *ctx, const uint8_t *gray, int rows, int cols, size_t step_bytes,
                             const sfm_orb_opts *opts, int32_t *n_keypoints);
/* Test aid: one per-pixel map of the last ORB extraction at pyramid level `level` (what: 0 level image, 1 blurred image, 2 FAST
 * score, 3 score after non-maximum suppression + border filter — uint8; 4 Harris response — float); out may be NULL to query the size. */
int sfm_features_orb_level(sfm_ctx *ctx, int what, int level, void *out, int32_t *width, int32_t *height);
/* Descriptor bytes per keypoint of the images extracted so far: 128 (SIFT), 32 (ORB), 0 = none yet. */
int sfm_features_descriptor_bytes(const sfm_ctx *ctx, int *bytes);
/* Host-side helper in front of the extractor: the grey conversion cv::SIFT applies to a colour photograph
 * (cvtColor COLOR_BGR2GRAY on 8-bit data: (B*3735 + G*19235 + R*9798 + 2^14) >> 15).  channels = 3 or 4 (alpha ignored),
 * rgb_order != 0 for R,G,B[,A] input; steps in bytes, 0 = dense.  No GPU involved. */
int sfm_gray_from_bgr(const uint8_t *src, int rows, int cols, size_t step_bytes, int channels, int rgb_order,
                      uint8_t *gray, size_t gray_step_bytes);
int sfm_features_clear(sfm_ctx *ctx);
int sfm_features_extract_sift(sfm_ctx *ctx, const uint8_t *gray, int rows, int cols, size_t step_bytes,
                              const sfm_sift_opts *opts, int32_t *n_keypoints);
int sfm_features_count(const sfm_ctx *ctx, int *n_images);
/* descriptors: n_keypoints x sfm_features_descriptor_bytes() */
int sfm_features_download(sfm_ctx *ctx, int image, int32_t *n_keypoints, sfm_keypoint *keypoints, uint8_t *descriptors);
int sfm_bank_from_features(sfm_ctx *ctx);
/* Measurement / test aids: [0] DoG extrema, [1] keypoints before removeDuplicatedSorted, [2] keypoints of the last
 * extraction; one level of its Gaussian pyramid (out may be NULL to query the size). */
int sfm_features_last_counts(const sfm_ctx *ctx, int32_t counts[3]);
int sfm_features_pyramid_level(sfm_ctx *ctx, int octave, int level, float *out, int32_t *width, int32_t *height);
/* Device time of the last extraction (CUDA events on the context stream): Gaussian pyramid construction and the whole
 * image (ms), and the algorithmic HBM bytes of the pyramid (8 B per pixel and level) — bench.py's roofline block. */
int sfm_features_last_profile(const sfm_ctx *ctx, double *pyramid_ms, double *total_ms, double *pyramid_bytes);

/* Counters of the last enqueue: kernels launched / bytes moved, for bench.py's gpu_launches etc. */
int sfm_last_stats(const sfm_ctx *ctx, int64_t *kernel_launches, int64_t *h2d_bytes, int64_t *d2h_bytes);
/* Non-integer CV_32F descriptors (3xTF32 tcgen05 candidate search + exact fp32 re-rank): how many query rows the
 * last run re-ranked in fp32 and how many of those failed the exactness certificate and were brute-forced. */
int sfm_last_float_stats(sfm_ctx *ctx, int64_t *rows_reranked, int64_t *rows_brute_forced);

/* Per-kernel device timing of the last enqueue (CUDA events on the context stream around the knn kernel and
 * around the filter/scan/compact kernels of every batch).  Measurement aid for bench.py's roofline block. */
int sfm_set_profiling(sfm_ctx *ctx, int on);
int sfm_last_profile(sfm_ctx *ctx, double *knn_ms, double *post_ms, int *knn_launches);

/* Operator level ---------------------------------------------------------------------------------
 * Replaces cv::DescriptorMatcher::knnMatch(query, train, matches, k) (call sites
 * UnorderedFeatureMatchingStrategy.cpp:51, VideoFeatureMatchingStrategy.cpp:62,
 * GridFeatureMatchingStrategy.cpp:105) and ::match() (…:68, :79, :122) for one host query/train pair.
 * Output has the shape of cv::batchDistance(..., K=k): nidx[nq*k] (-1 where fewer than k train rows)
 * and dist[nq*k] (L2: sqrtf; Hamming: popcount as float), ascending distance, ties -> lowest trainIdx.
 * Re-entrant per context (serialised internally). */
int sfm_knn_match(sfm_ctx *ctx, const void *query, int nq, size_t q_step,
                  const void *train, int nt, size_t t_step, int cols, int cv_depth,
                  int norm, int k, int engine, int32_t *nidx, float *dist);

#ifdef __cplusplus
}
#endif
#endif /* SFMMATCH_H */
