// sift_host_harness.cpp — TEST INFRASTRUCTURE, never linked into libsfmmatch.so.
//
// Runs the per-keypoint code of csrc/sift_core.cuh (the functions the CUDA kernels of csrc/sift.cu call, compiled here
// for the host) serially over a Gaussian pyramid handed in by the test, in cv::SIFT's loop order: extremum test,
// adjustLocalExtrema, orientation peaks, removeDuplicatedSorted, firstOctave correction, descriptors.  The CPU test
// suite compares its output with the numpy restatement (oracle/sift_np.py) on the SAME pyramid, which checks the
// shared arithmetic in a container without a GPU; the kernels' own plumbing (pyramid, lists, sort) is covered by the
// -m gpu tests.  Built by tests/test_sift_core_host.py:  g++ -O2 -ffp-contract=off -shared -fPIC.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../sfm-mvs-pipeline_b200/csrc/sift_core.cuh"

using namespace sfm::sift;

extern "C" int harness_sift(const float* pyr, int n_octaves, int n_layers, const int* w, const int* h, const int64_t* off,
                            double contrast_threshold, double edge_threshold, double sigma, int capacity, Keypoint* kps_out,
                            uint8_t* desc_out, int* n_candidates) {
    PyramidView P{};
    P.base = pyr; P.n_octaves = n_octaves; P.n_layers = n_layers;
    for (int o = 0; o < n_octaves; ++o) { P.w[o] = w[o]; P.h[o] = h[o]; P.off[o] = off[o]; }
    const float threshold = static_cast<float>(static_cast<int>(std::floor(0.5 * contrast_threshold / n_layers * 255)));
    std::vector<Keypoint> kps;
    int cands = 0;
    for (int o = 0; o < n_octaves; ++o)
        for (int i = 1; i <= n_layers; ++i)
            for (int r = kImgBorder; r < P.h[o] - kImgBorder; ++r)
                for (int c = kImgBorder; c < P.w[o] - kImgBorder; ++c) {
                    if (!is_extremum(P, o, i, r, c, threshold)) continue;
                    ++cands;
                    int layer = i, r1 = r, c1 = c;
                    Keypoint kp;
                    if (!adjust_local_extrema(P, o, layer, r1, c1, static_cast<float>(contrast_threshold),
                                              static_cast<float>(edge_threshold), static_cast<float>(sigma), kp))
                        continue;
                    const float scl_octv = kp.size * 0.5f / (1 << o);
                    float angles[kOriBins];
                    const int m = orientation_peaks(P, o, layer, r1, c1, cv_round(4.5f * scl_octv), 1.5f * scl_octv, angles);
                    for (int a = 0; a < m; ++a) { kp.angle = angles[a]; kps.push_back(kp); }
                }
    if (n_candidates) *n_candidates = cands;
    std::stable_sort(kps.begin(), kps.end(), [](const Keypoint& a, const Keypoint& b) { return keypoint_less(a, b); });
    std::vector<Keypoint> out;
    for (size_t i = 0; i < kps.size(); ++i)
        if (i == 0 || !keypoint_duplicate(kps[i - 1], kps[i])) out.push_back(kps[i]);
    if (static_cast<int>(out.size()) > capacity) return -1;
    std::vector<float> hist(kDescHistLen);
    for (size_t i = 0; i < out.size(); ++i) {
        Keypoint& kp = out[i];
        kp.octave = (kp.octave & ~255) | ((kp.octave - 1) & 255);
        kp.x *= 0.5f; kp.y *= 0.5f; kp.size *= 0.5f;
        int octave, layer;
        float scale;
        unpack_octave(kp.octave, octave, layer, scale);
        const int o = octave + 1;
        float angle = 360.f - kp.angle;
        if (std::fabs(angle - 360.f) < 1.1920929e-07f) angle = 0.f;
        sift_descriptor(P.level(o, layer), P.w[o], P.h[o], kp.x * scale, kp.y * scale, angle, kp.size * scale * 0.5f, hist.data(), 1,
                        desc_out + i * kDescLen);
        kps_out[i] = kp;
    }
    return static_cast<int>(out.size());
}
