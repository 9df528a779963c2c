#include "matching.h"

#include <algorithm>

#include <cmath>

namespace sfmhost {

static void check(sfm_ctx* ctx, int rc) {
    if (rc != SFM_OK) throw MatcherError(rc, sfm_last_error(ctx));
}

GpuDescriptorMatcher::GpuDescriptorMatcher(int normType, bool crossCheck, int device, bool flannMode)
    : norm_(normType), crossCheck_(crossCheck), flann_(flannMode) {
    if (normType != SFM_NORM_L2 && normType != SFM_NORM_HAMMING) throw std::invalid_argument("unsupported normType");
    const int rc = sfm_ctx_create(&ctx_, device);
    if (rc != SFM_OK) throw MatcherError(rc, sfm_last_error(nullptr));
}

GpuDescriptorMatcher::GpuDescriptorMatcher(int normType, bool crossCheck, const std::vector<int>& devices, bool flannMode)
    : norm_(normType), crossCheck_(crossCheck), flann_(flannMode) {
    if (normType != SFM_NORM_L2 && normType != SFM_NORM_HAMMING) throw std::invalid_argument("unsupported normType");
    if (devices.empty()) throw std::invalid_argument("empty device list");
    if (devices.size() == 1) {
        const int rc = sfm_ctx_create(&ctx_, devices[0]);
        if (rc != SFM_OK) throw MatcherError(rc, sfm_last_error(nullptr));
        return;
    }
    const int rc = sfm_mgpu_create(&group_, devices.data(), static_cast<int>(devices.size()));
    if (rc != SFM_OK) throw MatcherError(rc, sfm_last_error(nullptr));
    ctx_ = sfm_mgpu_ctx(group_, 0);
}

GpuDescriptorMatcher::~GpuDescriptorMatcher() {
    if (group_) sfm_mgpu_destroy(group_);        // owns its contexts
    else sfm_ctx_destroy(ctx_);
}

void GpuDescriptorMatcher::knnMatch(const DescriptorMat& q, const DescriptorMat& t,
                                    std::vector<std::vector<DMatch>>& matches, int k) const {
    matches.clear();
    // cv::batchDistance asserts type == src2.type() && src1.cols == src2.cols
    if (q.depth != t.depth || q.cols != t.cols || q.cols == 0)
        throw MatcherError(SFM_ERR_INVALID, "knnMatch: descriptor type/width mismatch");
    if (crossCheck_ && k != 1) throw MatcherError(SFM_ERR_INVALID, "knnMatch: crossCheck requires k == 1");
    if (k < 1) return;
    const int kk = k > 2 ? -1 : k;
    if (kk < 0) throw MatcherError(SFM_ERR_UNSUPPORTED, "knnMatch: k > 2 is not supported");
    std::vector<int32_t> nidx(static_cast<std::size_t>(q.rows) * kk);
    std::vector<float> dist(static_cast<std::size_t>(q.rows) * kk);
    check(ctx_, sfm_knn_match(ctx_, q.data, q.rows, q.step, t.data, t.rows, t.step, q.cols, q.depth, norm_, kk,
                              SFM_ENGINE_AUTO, nidx.data(), dist.data()));
    std::vector<int32_t> rev;
    if (crossCheck_ && t.rows > 0 && q.rows > 0) {
        std::vector<float> rdist(t.rows);
        rev.resize(t.rows);
        check(ctx_, sfm_knn_match(ctx_, t.data, t.rows, t.step, q.data, q.rows, q.step, q.cols, q.depth, norm_, 1,
                                  SFM_ENGINE_AUTO, rev.data(), rdist.data()));
    }
    matches.resize(q.rows);
    for (int r = 0; r < q.rows; ++r)
        for (int j = 0; j < kk; ++j) {
            const int32_t ti = nidx[static_cast<std::size_t>(r) * kk + j];
            if (ti < 0) break;
            if (crossCheck_ && rev[ti] != r) break;      // rows without a mutual partner return an empty list
            matches[r].push_back(DMatch{r, ti, 0, dist[static_cast<std::size_t>(r) * kk + j]});
        }
}

void GpuDescriptorMatcher::match(const DescriptorMat& q, const DescriptorMat& t, std::vector<DMatch>& out) const {
    std::vector<std::vector<DMatch>> knn;
    knnMatch(q, t, knn, 1);
    out.clear();
    for (auto& m : knn)
        if (!m.empty()) out.push_back(m[0]);
}

void IFeatureMatchingStrategy::calculateShotMatches(const Scene& scene, std::shared_ptr<GpuDescriptorMatcher>& matcher,
                                                    std::vector<ShotMatches>& out) {
    if (!matcher) throw std::invalid_argument("matcher must not be null");
    const auto& shots = scene.getShots();
    const PairList pl = matchPairs(shots.size());
    dropped_.assign(pl.size(), 0);
    if (shots.empty()) return;
    sfm_ctx* ctx = matcher->context();
    // one bank upload for the whole scene (replaces reading every shot's cv::Mat once per pair)
    std::vector<const void*> rows(shots.size());
    std::vector<int32_t> nrows(shots.size());
    std::vector<std::size_t> steps(shots.size());
    int cols = 0, depth = -1;
    for (std::size_t i = 0; i < shots.size(); ++i) {
        const DescriptorMat& d = shots[i]->descriptors;
        if (!d.empty()) {
            if (depth < 0) { depth = d.depth; cols = d.cols; }
            else if (d.depth != depth || d.cols != cols) throw MatcherError(SFM_ERR_INVALID, "descriptor type/width mismatch");
        }
    }
    if (depth < 0) { depth = SFM_CV_8U; cols = 128; }
    const std::size_t esz = depth == SFM_CV_32F ? 4 : 1;
    for (std::size_t i = 0; i < shots.size(); ++i) {
        const DescriptorMat& d = shots[i]->descriptors;
        rows[i] = d.empty() ? nullptr : d.data;
        nrows[i] = d.empty() ? 0 : d.rows;
        steps[i] = d.step ? d.step : static_cast<std::size_t>(cols) * esz;
    }
    std::vector<int32_t> flat(pl.size() * 2);
    for (std::size_t p = 0; p < pl.size(); ++p) { flat[2 * p] = pl[p].first; flat[2 * p + 1] = pl[p].second; }
    sfm_opts o;
    sfm_opts_default(&o, matcher->normType());
    o.ratio = stage_.ratio;
    o.cross_check = matcher->crossCheck() ? 1 : 0;
    o.k = matcher->crossCheck() ? 1 : 2;
    o.distinct = stage_.distinct ? 1 : 0;
    o.min_match_count = stage_.minMatchCount;
    sfm_result* res = nullptr;
    sfm_mgpu* group = matcher->group();
    if (group && !scene.bankResident) {
        // several GPUs: every device uploads its share of the scene, NVLink exchange, pairs dealt by cost, lists gathered in
        // pair-list order on devices[0] (byte-identical to the single-GPU result)
        const int rc = sfm_mgpu_match_pairs_from_host(group, static_cast<int>(shots.size()), rows.data(), nrows.data(), cols, steps.data(),
                                                      depth, flat.data(), static_cast<int64_t>(pl.size()), &o, &res);
        if (rc != SFM_OK) throw MatcherError(rc, sfm_mgpu_last_error(group));
    } else if (scene.bankResident) {
        // the feature extractor left descriptors + keypoints of these shots on the device(s): nothing to upload
        int n_bank = 0;
        check(ctx, sfm_bank_info(ctx, &n_bank, nullptr, nullptr));
        if (n_bank != static_cast<int>(shots.size())) throw MatcherError(SFM_ERR_STATE, "resident bank does not belong to this scene");
        if (group) {
            const int rc = sfm_mgpu_match_pairs(group, flat.data(), static_cast<int64_t>(pl.size()), &o, &res);
            if (rc != SFM_OK) throw MatcherError(rc, sfm_mgpu_last_error(group));
        } else
            check(ctx, sfm_match_pairs(ctx, flat.data(), static_cast<int64_t>(pl.size()), &o, &res));
    } else {
        // one call for the whole scene: bank upload (pipelined with the matching when the Mats are page-locked) + all pairs
        check(ctx, sfm_match_pairs_from_host(ctx, static_cast<int>(shots.size()), rows.data(), nrows.data(), cols, steps.data(),
                                             depth, flat.data(), static_cast<int64_t>(pl.size()), &o, &res));
    }
    const int64_t* off = sfm_result_offsets(res);
    const sfm_dmatch* m = sfm_result_matches(res);
    const uint8_t* dr = sfm_result_dropped(res);
    for (std::size_t p = 0; p < pl.size(); ++p) {
        dropped_[p] = dr[p];
        ShotMatches sm;
        sm.left = shots[pl[p].first];
        sm.right = shots[pl[p].second];
        sm.matches.assign(m + off[p], m + off[p + 1]);
        out.push_back(std::move(sm));
    }
    sfm_result_free(res);
}

MatchingStage::MatchingStage()
    : strategy_(std::make_shared<UnorderedFeatureMatchingStrategy>()) {}

std::vector<ShotMatches> MatchingStage::calculateShotMatches(const Scene& scene) {
    if (!matcher_) throw std::invalid_argument("Der Feature Matching Algorithmus darf nicht null sein.");
    StageOptions so;
    so.distinct = distinct_;
    so.minMatchCount = minMatchCount_;
    strategy_->setStageOptions(so);
    std::vector<ShotMatches> all, kept;
    strategy_->calculateShotMatches(scene, matcher_, all);
    const auto& dropped = strategy_->lastDropped();
    keptPair_.clear();
    lastShots_ = scene.getShots();
    lastBankResident_ = scene.bankResident;
    lastPairs_ = strategy_->matchPairs(lastShots_.size());
    for (std::size_t p = 0; p < all.size(); ++p)
        if (!dropped[p]) { kept.push_back(std::move(all[p])); keptPair_.push_back(p); }
    return kept;
}

void MatchingStage::calculateHomography(std::vector<ShotMatches>& shotMatches) {
    if (!matcher_) throw std::invalid_argument("Der Feature Matching Algorithmus darf nicht null sein.");
    if (shotMatches.size() != keptPair_.size())
        throw std::invalid_argument("calculateHomography expects the ShotMatches of the last calculateShotMatches");
    if (lastPairs_.empty()) return;
    sfm_ctx* ctx = matcher_->context();
    const std::size_t n = lastShots_.size();
    std::vector<const void*> pts(n);
    std::vector<int32_t> nrows(n);
    std::vector<std::size_t> steps(n);
    for (std::size_t i = 0; i < n; ++i) {
        const Shot& s = *lastShots_[i];
        nrows[i] = s.descriptors.empty() ? 0 : s.descriptors.rows;
        if (nrows[i] > 0 && !s.keypointPts && !lastBankResident_) throw std::invalid_argument("calculateHomography: shot without keypoints");
        pts[i] = s.keypointPts;
        steps[i] = s.keypointStep ? s.keypointStep : 8;
    }
    if (!lastBankResident_)          // sfm_bank_from_features already placed the keypoints next to the descriptors
        check(ctx, sfm_keypoints_upload(ctx, static_cast<int>(n), pts.data(), nrows.data(), steps.data()));
    // SfM.cpp:615-620 (the reference takes rightSize.height twice; mirrored)
    std::vector<double> thr(lastPairs_.size());
    for (std::size_t p = 0; p < lastPairs_.size(); ++p) {
        const Shot& l = *lastShots_[lastPairs_[p].first];
        const Shot& r = *lastShots_[lastPairs_[p].second];
        const double t = ransacReprojectionMatchingThreshold_;
        thr[p] = t < 0 ? -t : std::max(std::max(l.imageWidth, l.imageHeight), r.imageHeight) * t;
        if (!(thr[p] > 0)) throw std::invalid_argument("calculateHomography: image sizes are needed for a relative threshold");
    }
    std::vector<double> ratios(lastPairs_.size());
    check(ctx, sfm_homography_inlier_ratios(ctx, thr.data(), static_cast<int64_t>(thr.size()), nullptr, ratios.data(), nullptr,
                                            nullptr, nullptr));
    for (std::size_t k = 0; k < shotMatches.size(); ++k) shotMatches[k].homographyInlierRatio = ratios[keptPair_[k]];
}

GpuSiftFeatureDetector::GpuSiftFeatureDetector(const std::shared_ptr<GpuDescriptorMatcher>& matcher, int nfeatures, int nOctaveLayers,
                                               double contrastThreshold, double edgeThreshold, double sigma)
    : matcher_(matcher) {
    if (!matcher) throw std::invalid_argument("Der Feature Detektor braucht den GPU Kontext des Matchers.");
    sfm_sift_opts_default(&opts_);
    opts_.n_features = nfeatures;
    opts_.n_octave_layers = nOctaveLayers;
    opts_.contrast_threshold = contrastThreshold;
    opts_.edge_threshold = edgeThreshold;
    opts_.sigma = sigma;
}

// the loop of SfM::extractFeatures (SfM.cpp:577-597) around one detector; `extractOne` runs detect + compute of image i on the device
template <class ExtractOne>
static void extractAll(const GpuDescriptorMatcher& matcher, const std::vector<GrayImage>& images, Scene& scene, std::vector<Features>& features,
                       int descBytes, int detector, const void* opts, ExtractOne extractOne) {
    sfm_ctx* ctx = matcher.context();
    if (scene.shots.empty())
        for (std::size_t i = 0; i < images.size(); ++i) scene.shots.push_back(std::make_shared<Shot>());
    if (scene.shots.size() != images.size()) throw std::invalid_argument("extractFeatures: one image per shot");
    scene.bankResident = false;
    features.assign(images.size(), Features{});
    std::vector<int32_t> counts(images.size(), 0);
    sfm_mgpu* group = matcher.group();
    if (group) {
        // several GPUs: image i on device i % n, feature sets exchanged over NCCL, every device adopts the scene as its bank
        std::vector<const uint8_t*> gray(images.size());
        std::vector<int32_t> rows(images.size()), cols(images.size());
        std::vector<std::size_t> steps(images.size());
        for (std::size_t i = 0; i < images.size(); ++i) { gray[i] = images[i].data; rows[i] = images[i].rows; cols[i] = images[i].cols; steps[i] = images[i].step; }
        const int rc = sfm_mgpu_extract_features(group, detector, static_cast<int>(images.size()), gray.data(), rows.data(), cols.data(),
                                                 steps.data(), opts, counts.data());
        if (rc != SFM_OK) throw MatcherError(rc, sfm_mgpu_last_error(group));
    } else
        check(ctx, sfm_features_clear(ctx));
    for (std::size_t i = 0; i < images.size(); ++i) {
        const GrayImage& im = images[i];
        int32_t n = counts[i];
        if (!group) check(ctx, extractOne(im, &n));
        Features& f = features[i];
        f.descriptorBytes = descBytes;
        f.keypoints.resize(static_cast<std::size_t>(n));
        f.descriptors.resize(static_cast<std::size_t>(n) * descBytes);
        if (n > 0) check(ctx, sfm_features_download(ctx, static_cast<int>(i), &n, f.keypoints.data(), f.descriptors.data()));
        Shot& s = *scene.shots[i];
        s.descriptors = DescriptorMat{f.descriptors.data(), n, descBytes, static_cast<std::size_t>(descBytes), SFM_CV_8U};
        s.keypointPts = n > 0 ? &f.keypoints[0].x : nullptr;
        s.keypointStep = sizeof(sfm_keypoint);
        s.imageWidth = im.cols;
        s.imageHeight = im.rows;
    }
    if (!group) check(ctx, sfm_bank_from_features(ctx));
    scene.bankResident = true;
}

void GpuSiftFeatureDetector::extractFeatures(const std::vector<GrayImage>& images, Scene& scene, std::vector<Features>& features) {
    sfm_ctx* ctx = matcher_->context();
    extractAll(*matcher_, images, scene, features, 128, SFM_DETECTOR_SIFT, &opts_, [&](const GrayImage& im, int32_t* n) {
        return sfm_features_extract_sift(ctx, im.data, im.rows, im.cols, im.step, &opts_, n);
    });
}

GpuOrbFeatureDetector::GpuOrbFeatureDetector(const std::shared_ptr<GpuDescriptorMatcher>& matcher, int nfeatures) : matcher_(matcher) {
    if (!matcher) throw std::invalid_argument("Der Feature Detektor braucht den GPU Kontext des Matchers.");
    sfm_orb_opts_default(&opts_);
    opts_.n_features = nfeatures;
}

void GpuOrbFeatureDetector::extractFeatures(const std::vector<GrayImage>& images, Scene& scene, std::vector<Features>& features) {
    sfm_ctx* ctx = matcher_->context();
    extractAll(*matcher_, images, scene, features, 32, SFM_DETECTOR_ORB, &opts_, [&](const GrayImage& im, int32_t* n) {
        return sfm_features_extract_orb(ctx, im.data, im.rows, im.cols, im.step, &opts_, n);
    });
}

std::shared_ptr<GpuDescriptorMatcher> configureFeatureMatcher(const std::string& det, const std::string& mat, int device,
                                                              std::vector<std::string>* warnings) {
    const bool flann = mat == "FLANN";
    if (!flann && mat != "BF" && !mat.empty() && warnings)
        warnings->push_back("Unbekannter Algorithmus fuer den Merkmalsvergleich: " + mat + ". Benutze BF.");
    const int norm = det == "ORB" ? SFM_NORM_HAMMING : SFM_NORM_L2;      // anything else than ORB -> SIFT
    return std::make_shared<GpuDescriptorMatcher>(norm, false, device, flann);
}

std::shared_ptr<GpuDescriptorMatcher> configureFeatureMatcher(const std::string& det, const std::string& mat, const std::vector<int>& devices,
                                                              std::vector<std::string>* warnings) {
    const bool flann = mat == "FLANN";
    if (!flann && mat != "BF" && !mat.empty() && warnings)
        warnings->push_back("Unbekannter Algorithmus fuer den Merkmalsvergleich: " + mat + ". Benutze BF.");
    const int norm = det == "ORB" ? SFM_NORM_HAMMING : SFM_NORM_L2;
    return std::make_shared<GpuDescriptorMatcher>(norm, false, devices, flann);
}

std::shared_ptr<IFeatureMatchingStrategy> configureFeatureMatcherStrategy(int seq, int grid, std::vector<std::string>* warnings) {
    if (seq >= 2) {
        if (grid >= 1) return std::make_shared<GridFeatureMatchingStrategy>(seq, grid);
        return std::make_shared<VideoFeatureMatchingStrategy>(seq);
    }
    if (seq != 0 && warnings)
        warnings->push_back("Ungueltige Sequenzlaenge: " + std::to_string(seq) + ". Benutze default (Ungeordnet).");
    return std::make_shared<UnorderedFeatureMatchingStrategy>();
}

}  // namespace sfmhost
