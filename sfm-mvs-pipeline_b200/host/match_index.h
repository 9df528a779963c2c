// match_index.h — hash indices over the ShotMatches of a scene.
//
// Once matching takes milliseconds, the reference's consumers of the match lists become the bottleneck: they scan
// linearly, per lookup, first the vector of ShotMatches and then one match list:
//   Scene::addShotMatches            (Scene.cpp:327-338)  existing entry with the same (left, right)?
//   Scene::find3d2dMatches           (Scene.cpp:369-424)  first ShotMatches joining two shots in either orientation, then the
//                                                         first match whose keypoint on one side equals a given 2-D point
//   Scene::mergePointcloudElement3d2d (Scene.cpp:470-561) same ShotMatches lookup, then the first match whose LEFT and RIGHT
//                                                         keypoints equal two given points and |distance| <= a bound
// This header answers the same questions in O(1) expected time with identical results, including the reference's
// "first in vector / list order wins" semantics and its point comparison (cv::Point2d(KeyPoint.pt) == point: exact
// equality of the float coordinates, -0 == +0, NaN equals nothing).  Host-only, header-only, no OpenCV and no CUDA.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../../include/sfmmatch.h"

namespace sfmhost {

struct IndexedShotMatches {               // what the index needs to know about one ShotMatches (Scene.h:35-135)
    const void* left = nullptr;           // identity of the left / right CameraShot (the reference compares shared_ptrs)
    const void* right = nullptr;
    const sfm_dmatch* matches = nullptr;  // byte-compatible with cv::DMatch
    std::size_t n_matches = 0;
    const float* leftPts = nullptr;       // KeyPoint.pt of the left shot: first (x, y) pair and byte stride (28 for cv::KeyPoint)
    std::size_t leftStep = 8;
    const float* rightPts = nullptr;
    std::size_t rightStep = 8;
};

class ShotMatchIndex {
public:
    static constexpr int kNone = -1;

    void clear() { entries_.clear(); ordered_.clear(); leftPt_.clear(); rightPt_.clear(); }
    std::size_t size() const { return entries_.size(); }

    // Scene::addShotMatches: index of the existing entry with the same ordered (left, right), else appends and returns the
    // new index.  `inserted` tells which happened.
    int add(const IndexedShotMatches& sm, bool* inserted = nullptr) {
        const int found = find(sm.left, sm.right);
        if (found != kNone) { if (inserted) *inserted = false; return found; }
        const int idx = static_cast<int>(entries_.size());
        entries_.push_back(sm);
        ordered_.emplace(PairKey{sm.left, sm.right}, idx);
        leftPt_.emplace_back();
        rightPt_.emplace_back();
        auto& lm = leftPt_.back();
        auto& rm = rightPt_.back();
        lm.reserve(sm.n_matches * 2);
        rm.reserve(sm.n_matches * 2);
        for (std::size_t i = 0; i < sm.n_matches; ++i) {
            uint64_t k;
            if (pointKey(pt(sm.leftPts, sm.leftStep, sm.matches[i].queryIdx), &k)) lm[k].push_back(static_cast<int>(i));
            if (pointKey(pt(sm.rightPts, sm.rightStep, sm.matches[i].trainIdx), &k)) rm[k].push_back(static_cast<int>(i));
        }
        if (inserted) *inserted = true;
        return idx;
    }

    // ordered lookup (addShotMatches)
    int find(const void* left, const void* right) const {
        auto it = ordered_.find(PairKey{left, right});
        return it == ordered_.end() ? kNone : it->second;
    }
    // first entry in insertion order joining a and b in either orientation (find3d2dMatches, mergePointcloudElement3d2d)
    int findEither(const void* a, const void* b) const {
        const int x = find(a, b), y = find(b, a);
        if (x == kNone) return y;
        if (y == kNone) return x;
        return x < y ? x : y;
    }
    // first match (list order) of entry `sm` whose keypoint on the given side is exactly (x, y)  (find3d2dMatches)
    int firstMatchWithPoint(int sm, bool leftSide, double x, double y) const {
        uint64_t k;
        if (!doubleKey(x, y, &k)) return kNone;
        const auto& m = leftSide ? leftPt_[sm] : rightPt_[sm];
        auto it = m.find(k);
        return it == m.end() ? kNone : it->second.front();
    }
    // first match whose left keypoint is (lx, ly), right keypoint is (rx, ry) and |distance| <= maxAbsDistance
    // (mergePointcloudElement3d2d)
    int firstMatchWithPoints(int sm, double lx, double ly, double rx, double ry, double maxAbsDistance) const {
        uint64_t kl, kr;
        if (!doubleKey(lx, ly, &kl) || !doubleKey(rx, ry, &kr)) return kNone;
        auto it = leftPt_[sm].find(kl);
        if (it == leftPt_[sm].end()) return kNone;
        const IndexedShotMatches& e = entries_[sm];
        for (int i : it->second) {                                  // ascending list order
            uint64_t k;
            if (!pointKey(pt(e.rightPts, e.rightStep, e.matches[i].trainIdx), &k) || k != kr) continue;
            if (std::fabs(static_cast<double>(e.matches[i].distance)) <= maxAbsDistance) return i;
        }
        return kNone;
    }
    const IndexedShotMatches& entry(int i) const { return entries_[i]; }

private:
    struct PairKey {
        const void* a; const void* b;
        bool operator==(const PairKey& o) const { return a == o.a && b == o.b; }
    };
    struct PairHash {
        std::size_t operator()(const PairKey& k) const {
            uint64_t x = reinterpret_cast<uintptr_t>(k.a) * 0x9E3779B97F4A7C15ull;
            x ^= (reinterpret_cast<uintptr_t>(k.b) + 0x7F4A7C15ull) * 0xBF58476D1CE4E5B9ull;
            return static_cast<std::size_t>(x ^ (x >> 29));
        }
    };
    using PointMap = std::unordered_map<uint64_t, std::vector<int>>;

    static const float* pt(const float* base, std::size_t step, int idx) {
        return reinterpret_cast<const float*>(reinterpret_cast<const char*>(base) + static_cast<std::size_t>(idx) * step);
    }
    // bit pattern of an (x, y) float pair as the hash key; -0 -> +0, NaN has no key (it equals nothing)
    static bool pointKey(const float* p, uint64_t* key) {
        float x = p[0], y = p[1];
        if (x != x || y != y) return false;
        if (x == 0.0f) x = 0.0f;
        if (y == 0.0f) y = 0.0f;
        uint32_t bx, by;
        std::memcpy(&bx, &x, 4);
        std::memcpy(&by, &y, 4);
        *key = (static_cast<uint64_t>(bx) << 32) | by;
        return true;
    }
    // the query point is a cv::Point2d: it can only equal a float keypoint if both coordinates are exactly floats
    static bool doubleKey(double x, double y, uint64_t* key) {
        const float fx = static_cast<float>(x), fy = static_cast<float>(y);
        if (static_cast<double>(fx) != x || static_cast<double>(fy) != y) return false;     // also rejects NaN
        const float p[2] = {fx, fy};
        return pointKey(p, key);
    }

    std::vector<IndexedShotMatches> entries_;
    std::unordered_map<PairKey, int, PairHash> ordered_;
    std::vector<PointMap> leftPt_, rightPt_;
};

}  // namespace sfmhost
