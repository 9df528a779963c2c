// cv_adapter.h — drop-in cv::DescriptorMatcher over the C ABI (header-only; needs OpenCV C++ headers, which are
// NOT present in the build container, so this file is compiled only inside the reference's own build).
//
// Usage in the reference (PhotogrammetrieCli::configureFeatureMatcher, PhotogrammetrieCli.cpp:359-392):
//     matcher = cv::makePtr<sfmhost::GpuMatcher>(cv::NORM_L2);        // instead of cv::BFMatcher::create(cv::NORM_L2)
// Every knnMatch / match the strategies issue (UnorderedFeatureMatchingStrategy.cpp:51/:68,
// VideoFeatureMatchingStrategy.cpp:62/:79, GridFeatureMatchingStrategy.cpp:105/:122) then runs on the GPU.
#pragma once
#include <opencv2/features2d.hpp>

#include <mutex>

#include "../../include/sfmmatch.h"

namespace sfmhost {

class GpuMatcher : public cv::DescriptorMatcher {
public:
    explicit GpuMatcher(int normType = cv::NORM_L2, bool crossCheck = false, int device = 0)
        : normType_(normType), crossCheck_(crossCheck), device_(device) {
        if (sfm_ctx_create(&ctx_, device) != SFM_OK) CV_Error(cv::Error::GpuNotSupported, sfm_last_error(nullptr));
    }
    ~GpuMatcher() override { sfm_ctx_destroy(ctx_); }

    bool isMaskSupported() const override { return false; }
    cv::Ptr<cv::DescriptorMatcher> clone(bool /*emptyTrainData*/ = false) const override {
        return cv::makePtr<GpuMatcher>(normType_, crossCheck_, device_);
    }

protected:
    // knnMatch(query, train, matches, k) lands here after DescriptorMatcher::knnMatch cloned the matcher and add()ed
    // the train descriptors (one train image, imgIdx 0), exactly as for cv::BFMatcher.
    void knnMatchImpl(cv::InputArray queryDescriptors, std::vector<std::vector<cv::DMatch>>& matches, int k,
                      cv::InputArrayOfArrays masks, bool /*compactResult*/) override {
        CV_Assert(masks.empty() && trainDescCollection.size() == 1);
        const cv::Mat q = queryDescriptors.getMat(), t = trainDescCollection[0];
        matches.clear();
        CV_Assert(q.type() == t.type() && q.cols == t.cols);             // cv::batchDistance's own assert
        CV_Assert(!crossCheck_ || k == 1);
        CV_Assert(k == 1 || k == 2);
        const int depth = q.depth() == CV_8U ? SFM_CV_8U : SFM_CV_32F;
        CV_Assert(q.depth() == CV_8U || q.depth() == CV_32F);
        std::vector<int32_t> nidx(static_cast<size_t>(q.rows) * k), rev;
        std::vector<float> dist(static_cast<size_t>(q.rows) * k), rdist;
        int rc = sfm_knn_match(ctx_, q.data, q.rows, q.step, t.data, t.rows, t.step, q.cols, depth, normType_, k,
                               SFM_ENGINE_AUTO, nidx.data(), dist.data());
        if (rc != SFM_OK) CV_Error(cv::Error::StsError, sfm_last_error(ctx_));      // -> caught at *Strategy.cpp:66
        if (crossCheck_ && q.rows && t.rows) {
            rev.resize(t.rows); rdist.resize(t.rows);
            rc = sfm_knn_match(ctx_, t.data, t.rows, t.step, q.data, q.rows, q.step, q.cols, depth, normType_, 1,
                               SFM_ENGINE_AUTO, rev.data(), rdist.data());
            if (rc != SFM_OK) CV_Error(cv::Error::StsError, sfm_last_error(ctx_));
        }
        matches.resize(q.rows);
        for (int r = 0; r < q.rows; ++r)
            for (int j = 0; j < k; ++j) {
                const int ti = nidx[static_cast<size_t>(r) * k + j];
                if (ti < 0 || (crossCheck_ && rev[ti] != r)) break;
                matches[r].emplace_back(r, ti, 0, dist[static_cast<size_t>(r) * k + j]);
            }
    }
    void radiusMatchImpl(cv::InputArray, std::vector<std::vector<cv::DMatch>>&, float, cv::InputArrayOfArrays,
                         bool) override {
        CV_Error(cv::Error::StsNotImplemented, "radiusMatch is not used by the pipeline and not provided");
    }

private:
    sfm_ctx* ctx_ = nullptr;
    int normType_;
    bool crossCheck_;
    int device_;
};

}  // namespace sfmhost
