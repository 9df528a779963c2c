// matching.h — C++ host side above the C ABI, mirroring the reference's matching-stage interface
// (same names, argument meaning and error behaviour) without depending on OpenCV:
//
//   DescriptorMat                <-> cv::Mat Features::descriptors            (CameraShot.h:39-42)
//   DMatch                       <-> cv::DMatch (byte-compatible)
//   GpuDescriptorMatcher         <-> cv::Ptr<cv::DescriptorMatcher>           (SfM.h:78, knnMatch/match call sites
//                                    UnorderedFeatureMatchingStrategy.cpp:51/:68)
//   Shot / Scene / ShotMatches   <-> CameraShot / Scene / ShotMatches         (Scene.h:35-135)
//   IFeatureMatchingStrategy + Unordered/Video/Grid strategies                (IFeatureMatchingStrategy.h:34-48 ...)
//   MatchingStage                <-> SfM::calculateShotMatches + its setters  (SfM.cpp:542-575, :52-79) and
//                                    SfM::calculateHomography                 (SfM.cpp:599-637)
//   GpuSiftFeatureDetector       <-> cv::Ptr<cv::Feature2D> of PhotogrammetrieCli::configureFeatureDetector
//                                    (PhotogrammetrieCli.cpp:342-357) + the loop of SfM::extractFeatures (SfM.cpp:577-597)
//
// The strategies here do not loop over pairs calling knnMatch: they hand the whole pair list to
// sfm_match_pairs (one bank upload, batched kernels), which is the throughput path.
#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/sfmmatch.h"
#include "pair_selector.h"

namespace sfmhost {

using DMatch = sfm_dmatch;

struct DescriptorMat {          // subset of cv::Mat the stage touches
    const void* data = nullptr;
    int rows = 0, cols = 0;
    std::size_t step = 0;       // bytes between rows (0 = dense)
    int depth = SFM_CV_32F;     // SFM_CV_8U | SFM_CV_32F
    bool empty() const { return rows == 0 || cols == 0; }
};

// Thrown where the reference would see cv::Exception out of OpenCV (caught at *Strategy.cpp:66).
struct MatcherError : std::runtime_error {
    int code;
    MatcherError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

class GpuDescriptorMatcher {
public:
    // normType: SFM_NORM_L2 (BFMatcher::create(NORM_L2)) or SFM_NORM_HAMMING; crossCheck as in cv::BFMatcher.
    // `flannMode` records that -Pfeature-matcher=FLANN was requested: the exact GPU matcher replaces it.
    explicit GpuDescriptorMatcher(int normType = SFM_NORM_L2, bool crossCheck = false, int device = 0, bool flannMode = false);
    // several GPUs from this one process (sfm_mgpu_*: one worker thread per device, NCCL inside the library): the strategies
    // deal the pair list over the devices; knnMatch / match and the homography stage run on devices[0]
    GpuDescriptorMatcher(int normType, bool crossCheck, const std::vector<int>& devices, bool flannMode = false);
    ~GpuDescriptorMatcher();
    GpuDescriptorMatcher(const GpuDescriptorMatcher&) = delete;
    GpuDescriptorMatcher& operator=(const GpuDescriptorMatcher&) = delete;

    // cv::DescriptorMatcher::knnMatch(query, train, matches, k): per query row up to k matches, ascending distance.
    void knnMatch(const DescriptorMat& query, const DescriptorMat& train, std::vector<std::vector<DMatch>>& matches, int k) const;
    // cv::DescriptorMatcher::match(query, train, matches): best match per query row (cross-check honoured).
    void match(const DescriptorMat& query, const DescriptorMat& train, std::vector<DMatch>& matches) const;

    int normType() const { return norm_; }
    bool crossCheck() const { return crossCheck_; }
    bool flannMode() const { return flann_; }
    sfm_ctx* context() const { return ctx_; }
    sfm_mgpu* group() const { return group_; }          // null with one device

private:
    sfm_ctx* ctx_ = nullptr;
    sfm_mgpu* group_ = nullptr;
    int norm_;
    bool crossCheck_, flann_;
};

struct Shot {
    std::string imagePath;
    DescriptorMat descriptors;      // Features::descriptors
    // Features::keypoints (CameraShot.h:39-42): pointer to the first KeyPoint.pt (two floats) and the distance in bytes
    // between consecutive points (sizeof(cv::KeyPoint) = 28 for a std::vector<cv::KeyPoint>, 8 for packed float pairs);
    // one point per descriptor row.  Needed by MatchingStage::calculateHomography only.
    const float* keypointPts = nullptr;
    std::size_t keypointStep = 0;
    int imageWidth = 0, imageHeight = 0;    // CameraShot::getImageSize()
};

struct Scene {
    std::vector<std::shared_ptr<Shot>> shots;
    // set by GpuSiftFeatureDetector::extractFeatures: descriptors and keypoints of exactly these shots already form the
    // matcher's device bank (sfm_bank_from_features), the strategies and the homography stage upload nothing
    bool bankResident = false;
    const std::vector<std::shared_ptr<Shot>>& getShots() const { return shots; }
};

// Features (CameraShot.h:39-42) as the device extractor returns them: cv::KeyPoint minus class_id, CV_8U descriptor rows
struct Features {
    std::vector<sfm_keypoint> keypoints;
    std::vector<uint8_t> descriptors;       // keypoints.size() x descriptorBytes
    int descriptorBytes = 128;              // 128: cv::SIFT rows (as CV_8U), 32: cv::ORB rows
};

struct GrayImage {              // subset of the CV_8UC1 cv::Mat that Shot::loadImage + cv::SIFT's grey conversion yield
    const uint8_t* data = nullptr;
    int rows = 0, cols = 0;
    std::size_t step = 0;       // bytes between rows (0 = dense)
};

struct ShotMatches {
    std::shared_ptr<Shot> left, right;    // left -> queryIdx, right -> trainIdx
    std::vector<DMatch> matches;
    double homographyInlierRatio = -1;
};

struct StageOptions {               // what SfM::calculateShotMatches applies after the strategy
    bool distinct = false;          // --distinct-matches (SfM.cpp:547-564)
    int minMatchCount = 0;          // SfM.cpp:566-570 (0 = strategies alone, like the reference's strategy classes)
    double ratio = 0.7;             // Lowe ratio (UnorderedFeatureMatchingStrategy.cpp:53)
};

class IFeatureMatchingStrategy {
public:
    virtual ~IFeatureMatchingStrategy() = default;
    virtual PairList matchPairs(std::size_t nShots) const = 0;
    // Same contract as the reference: appends one ShotMatches per pair.  Order = pair-list order (the reference's
    // order depends on thread timing).
    virtual void calculateShotMatches(const Scene& scene, std::shared_ptr<GpuDescriptorMatcher>& matcher,
                                      std::vector<ShotMatches>& matches);
    void setStageOptions(const StageOptions& o) { stage_ = o; }
    const std::vector<uint8_t>& lastDropped() const { return dropped_; }

protected:
    StageOptions stage_;
    std::vector<uint8_t> dropped_;
};

class UnorderedFeatureMatchingStrategy : public IFeatureMatchingStrategy {
public:
    PairList matchPairs(std::size_t n) const override { return unorderedPairs(n); }
};

class VideoFeatureMatchingStrategy : public IFeatureMatchingStrategy {
public:
    explicit VideoFeatureMatchingStrategy(int sequenceLength) { setSequenceLength(sequenceLength); }
    void setSequenceLength(int s) {
        if (s < 2) throw std::invalid_argument("Die Sequenzlaenge darf nicht kleiner als 2 sein.");
        sequenceLength_ = s;
    }
    PairList matchPairs(std::size_t n) const override { return videoPairs(n, sequenceLength_); }
private:
    int sequenceLength_ = 2;
};

class GridFeatureMatchingStrategy : public IFeatureMatchingStrategy {
public:
    GridFeatureMatchingStrategy(int sequenceLength, int rowLength) { setSequenceLength(sequenceLength); setRowLength(rowLength); }
    void setSequenceLength(int s) {
        if (s < 2) throw std::invalid_argument("Die Sequenzlaenge darf nicht kleiner als 2 sein.");
        sequenceLength_ = s;
    }
    void setRowLength(int r) {
        if (r < 1) throw std::invalid_argument("Die Zeilenlaenge darf nicht kleiner als 1 sein.");
        rowLength_ = r;
    }
    PairList matchPairs(std::size_t n) const override { return gridPairs(n, sequenceLength_, rowLength_); }
private:
    int sequenceLength_ = 2, rowLength_ = 1;
};

// SfM::calculateShotMatches: strategy + distinct filter + min-match-count drop.
class MatchingStage {
public:
    MatchingStage();
    void setMatchingAlgorithm(const std::shared_ptr<GpuDescriptorMatcher>& m) {
        if (!m) throw std::invalid_argument("Der Feature Matching Algorithmus darf nicht null sein.");
        matcher_ = m;
    }
    void setFeatureMatchingStrategy(const std::shared_ptr<IFeatureMatchingStrategy>& s) {
        if (!s) throw std::invalid_argument("Die Feature Matching Strategie darf nicht null sein.");
        strategy_ = s;
    }
    void setMinMatchCount(int n) {
        if (n < 4) throw std::invalid_argument("Der minimale Match Count darf nicht < 4 sein.");
        minMatchCount_ = n;
    }
    void setUseDistinctFeatureMatchTest(bool b) { distinct_ = b; }
    // SfM::setRansacReprojectionMatchingThreshold (SfM.cpp:108-115): < 0 = pixels, > 0 = fraction of the image size
    void setRansacReprojectionMatchingThreshold(double t) {
        if (t == 0) throw std::invalid_argument("Der RANSAC Schwellwert darf nicht 0 sein.");
        ransacReprojectionMatchingThreshold_ = t;
    }
    // returns the surviving ShotMatches (pairs below minMatchCount erased), pair-list order
    std::vector<ShotMatches> calculateShotMatches(const Scene& scene);
    // SfM::calculateHomography (SfM.cpp:599-637) for the ShotMatches returned by the last calculateShotMatches: sets
    // homographyInlierRatio from the match lists that are still resident on the GPU (pairs with < 4 matches keep -1).
    // Every shot needs keypointPts and an image size.
    void calculateHomography(std::vector<ShotMatches>& shotMatches);

private:
    std::shared_ptr<GpuDescriptorMatcher> matcher_;
    std::shared_ptr<IFeatureMatchingStrategy> strategy_;
    int minMatchCount_ = 20;        // SfM.h default
    bool distinct_ = false;
    double ransacReprojectionMatchingThreshold_ = -3.0;      // SfM.h:50
    std::vector<std::size_t> keptPair_;                      // pair-list position of every ShotMatches returned last
    std::vector<std::shared_ptr<Shot>> lastShots_;
    PairList lastPairs_;
    bool lastBankResident_ = false;
};

// cv::SIFT::create(featureLimit, 3, 0.09) (PhotogrammetrieCli.cpp:355) running on the matcher's GPU context, and the loop of
// SfM::extractFeatures around it.  The extracted descriptors stay on the device and become the matcher's bank.
class GpuSiftFeatureDetector {
public:
    explicit GpuSiftFeatureDetector(const std::shared_ptr<GpuDescriptorMatcher>& matcher, int nfeatures = 10000, int nOctaveLayers = 3,
                                    double contrastThreshold = 0.09, double edgeThreshold = 10.0, double sigma = 1.6);
    // SfM::extractFeatures: detect + compute for every image (call order = shot order).  features[i] receives what
    // Shot::setFeatures stores; the shots of `scene` (one per image, created if the scene is empty) get their descriptor /
    // keypoint views pointed at features[i], their image size set, and scene.bankResident = true.
    void extractFeatures(const std::vector<GrayImage>& images, Scene& scene, std::vector<Features>& features);

private:
    std::shared_ptr<GpuDescriptorMatcher> matcher_;
    sfm_sift_opts opts_;
};

// cv::ORB::create(featureLimit) (PhotogrammetrieCli.cpp:347-348) on the matcher's GPU context; same contract as the SIFT detector,
// the bank it leaves behind holds 32-byte descriptors for NORM_HAMMING matching.
class GpuOrbFeatureDetector {
public:
    explicit GpuOrbFeatureDetector(const std::shared_ptr<GpuDescriptorMatcher>& matcher, int nfeatures = 500);
    void extractFeatures(const std::vector<GrayImage>& images, Scene& scene, std::vector<Features>& features);

private:
    std::shared_ptr<GpuDescriptorMatcher> matcher_;
    sfm_orb_opts opts_;
};

// PhotogrammetrieCli::configureFeatureMatcher / configureFeatureMatcherStrategy (PhotogrammetrieCli.cpp:320-392)
std::shared_ptr<GpuDescriptorMatcher> configureFeatureMatcher(const std::string& featureDetector,
                                                              const std::string& featureMatcher, int device,
                                                              std::vector<std::string>* warnings);
std::shared_ptr<GpuDescriptorMatcher> configureFeatureMatcher(const std::string& featureDetector,
                                                              const std::string& featureMatcher, const std::vector<int>& devices,
                                                              std::vector<std::string>* warnings);
std::shared_ptr<IFeatureMatchingStrategy> configureFeatureMatcherStrategy(int featureSequence, int featureGridLength,
                                                                          std::vector<std::string>* warnings);

}  // namespace sfmhost
