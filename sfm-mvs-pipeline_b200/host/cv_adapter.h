// cv_adapter.h — drop-in cv::DescriptorMatcher over the C ABI (header-only; needs OpenCV C++ headers: compiled inside
// the reference's own build, and in this repository against tests/stubs/opencv2/features2d.hpp by
// tests/host_adapter_harness.cpp, which drives it from an OpenMP loop exactly like the reference's strategies).
//
// Usage in the reference (PhotogrammetrieCli::configureFeatureMatcher, PhotogrammetrieCli.cpp:359-392):
//     matcher = cv::makePtr<sfmhost::GpuMatcher>(cv::NORM_L2);        // instead of cv::BFMatcher::create(cv::NORM_L2)
// Every knnMatch / match the strategies issue (UnorderedFeatureMatchingStrategy.cpp:51/:68,
// VideoFeatureMatchingStrategy.cpp:62/:79, GridFeatureMatchingStrategy.cpp:105/:122) then runs on the GPU.
//
// OpenCV's two-argument knnMatch CLONES the matcher on every call, and the reference calls it from an OpenMP loop on one
// shared matcher (UnorderedFeatureMatchingStrategy.cpp:40,51).  A clone therefore must be cheap: all clones share one
// small pool of library contexts (streams, pinned and device buffers live there); a call borrows a context, so up to
// `contexts` calls overlap (the upload of one pair with the kernels of another) and nothing is created per pair.
#pragma once
#include <opencv2/features2d.hpp>

#include <condition_variable>
#include <cstdlib>
#include <memory>
#include <mutex>
#include <vector>

#include "../../include/sfmmatch.h"

namespace sfmhost {

class GpuMatcher : public cv::DescriptorMatcher {
    // contexts shared by a matcher and all of its clones; created lazily, destroyed with the last owner
    struct Pool {
        Pool(int device_, int max_) : device(device_), max(max_ < 1 ? 1 : max_) {}
        ~Pool() { for (sfm_ctx* c : all) sfm_ctx_destroy(c); }
        sfm_ctx* acquire() {
            std::unique_lock<std::mutex> lk(mu);
            for (;;) {
                if (!idle.empty()) { sfm_ctx* c = idle.back(); idle.pop_back(); return c; }
                if (static_cast<int>(all.size()) < max) {
                    sfm_ctx* c = nullptr;
                    if (sfm_ctx_create(&c, device) != SFM_OK) {
                        if (all.empty()) return nullptr;             // no usable GPU at all: the caller raises cv::Exception
                    } else { all.push_back(c); return c; }
                }
                cv.wait(lk);
            }
        }
        void release(sfm_ctx* c) { { std::lock_guard<std::mutex> lk(mu); idle.push_back(c); } cv.notify_one(); }
        const int device, max;
        std::mutex mu;
        std::condition_variable cv;
        std::vector<sfm_ctx*> all, idle;
    };
    struct Lease {
        Lease(Pool& p) : pool(p), ctx(p.acquire()) {}
        ~Lease() { if (ctx) pool.release(ctx); }
        Pool& pool;
        sfm_ctx* ctx;
    };

public:
    explicit GpuMatcher(int normType = cv::NORM_L2, bool crossCheck = false, int device = 0, int contexts = 0)
        : normType_(normType), crossCheck_(crossCheck) {
        if (contexts <= 0) { const char* e = std::getenv("SFM_ADAPTER_CONTEXTS"); contexts = e ? std::atoi(e) : 2; }
        pool_ = std::make_shared<Pool>(device, contexts);
    }

    bool isMaskSupported() const override { return false; }
    cv::Ptr<cv::DescriptorMatcher> clone(bool emptyTrainData = false) const override {
        cv::Ptr<GpuMatcher> m = cv::makePtr<GpuMatcher>(*this);       // shares pool_: no CUDA object is created here
        if (emptyTrainData) m->clear();
        return m;
    }

protected:
    // knnMatch(query, train, matches, k) lands here after DescriptorMatcher::knnMatch cloned the matcher and add()ed
    // the train descriptors (one train image, imgIdx 0), exactly as for cv::BFMatcher.
    void knnMatchImpl(cv::InputArray queryDescriptors, std::vector<std::vector<cv::DMatch>>& matches, int k,
                      cv::InputArrayOfArrays masks, bool compactResult) override {
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
        Lease lease(*pool_);
        if (!lease.ctx) CV_Error(cv::Error::GpuNotSupported, sfm_last_error(nullptr));
        int rc = sfm_knn_match(lease.ctx, q.data, q.rows, q.step, t.data, t.rows, t.step, q.cols, depth, normType_, k,
                               SFM_ENGINE_AUTO, nidx.data(), dist.data());
        if (rc != SFM_OK) CV_Error(cv::Error::StsError, sfm_last_error(lease.ctx));      // -> caught at *Strategy.cpp:66
        if (crossCheck_ && q.rows && t.rows) {
            rev.resize(t.rows); rdist.resize(t.rows);
            rc = sfm_knn_match(lease.ctx, t.data, t.rows, t.step, q.data, q.rows, q.step, q.cols, depth, normType_, 1,
                               SFM_ENGINE_AUTO, rev.data(), rdist.data());
            if (rc != SFM_OK) CV_Error(cv::Error::StsError, sfm_last_error(lease.ctx));
        }
        matches.reserve(q.rows);
        for (int r = 0; r < q.rows; ++r) {
            std::vector<cv::DMatch> row;
            for (int j = 0; j < k; ++j) {
                const int ti = nidx[static_cast<size_t>(r) * k + j];
                if (ti < 0 || (crossCheck_ && rev[ti] != r)) break;
                row.emplace_back(r, ti, 0, dist[static_cast<size_t>(r) * k + j]);
            }
            if (!compactResult || !row.empty()) matches.push_back(std::move(row));
        }
    }
    void radiusMatchImpl(cv::InputArray, std::vector<std::vector<cv::DMatch>>&, float, cv::InputArrayOfArrays,
                         bool) override {
        CV_Error(cv::Error::StsNotImplemented, "radiusMatch is not used by the pipeline and not provided");
    }

private:
    std::shared_ptr<Pool> pool_;
    int normType_;
    bool crossCheck_;
};

}  // namespace sfmhost
