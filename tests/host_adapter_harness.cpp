// tests/host_adapter_harness.cpp — drives host/cv_adapter.h (compiled against tests/stubs/opencv2) the way the
// reference's strategies drive their injected matcher: UnorderedFeatureMatchingStrategy.cpp:32-91 restated — all pairs
// i < j, `#pragma omp parallel for` over the pairs on ONE shared cv::Ptr<cv::DescriptorMatcher>, knnMatch(k = 2),
// Lowe ratio in double, cv::Exception -> match() fallback, push under a critical section.  Test infrastructure.
//
// usage: host_adapter_harness <in.bin> <out.bin> [threads]
//   in : int32 norm, depth, cols, n_images, n_rows[n_images], then the descriptor rows of every image (dense)
//   out: int64 n_pairs, then per pair (input pair order): int32 left, right, int64 count, count x cv::DMatch (16 B)
#include <omp.h>

#include <chrono>
#include <cstdio>
#include <fstream>
#include <vector>

#include "../sfm-mvs-pipeline_b200/host/cv_adapter.h"

int main(int argc, char** argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: %s in.bin out.bin [threads]\n", argv[0]); return 2; }
    std::ifstream in(argv[1], std::ios::binary);
    int32_t hdr[4];
    in.read(reinterpret_cast<char*>(hdr), sizeof hdr);
    const int norm = hdr[0], depth = hdr[1], cols = hdr[2], n_images = hdr[3];
    std::vector<int32_t> n_rows(n_images);
    in.read(reinterpret_cast<char*>(n_rows.data()), n_images * 4);
    const size_t esz = depth == CV_32F ? 4 : 1;
    std::vector<std::vector<uint8_t>> store(n_images);
    std::vector<cv::Mat> desc(n_images);
    for (int i = 0; i < n_images; ++i) {
        store[i].resize(static_cast<size_t>(n_rows[i]) * cols * esz + 16);
        in.read(reinterpret_cast<char*>(store[i].data()), static_cast<std::streamsize>(n_rows[i]) * cols * esz);
        // an image without features is an EMPTY Mat (0 x 0), as Features::descriptors is in the reference
        desc[i] = n_rows[i] ? cv::Mat(n_rows[i], cols, CV_MAKETYPE(depth, 1), store[i].data()) : cv::Mat();
    }
    if (argc > 3) omp_set_num_threads(std::atoi(argv[3]));
    std::vector<std::pair<int, int>> matchPairs;
    for (int i = 0; i < n_images; ++i)
        for (int j = i + 1; j < n_images; ++j) matchPairs.emplace_back(i, j);
    cv::Ptr<cv::DescriptorMatcher> matcher = cv::makePtr<sfmhost::GpuMatcher>(norm);
    std::vector<std::vector<cv::DMatch>> result(matchPairs.size());
    int fallbacks = 0, failed = 0;
    const auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic) shared(matcher, result, desc, matchPairs)
    for (size_t p = 0; p < matchPairs.size(); ++p) {
        const cv::Mat& dl = desc[matchPairs[p].first];
        const cv::Mat& dr = desc[matchPairs[p].second];
        std::vector<cv::DMatch> good;
        try {
            std::vector<std::vector<cv::DMatch>> knn;
            matcher->knnMatch(dl, dr, knn, 2);
            for (const auto& m : knn) {
                if (m.size() >= 2) {
                    if (static_cast<double>(m[0].distance) < static_cast<double>(m[1].distance) * 0.7) good.push_back(m[0]);
                } else if (m.size() == 1) good.push_back(m[0]);
            }
        } catch (const cv::Exception&) {
            try {
#pragma omp atomic
                ++fallbacks;
                matcher->match(dl, dr, good);
            } catch (const cv::Exception&) {
#pragma omp atomic
                ++failed;                          // the reference would std::terminate here
                good.clear();
            }
        }
#pragma omp critical
        result[p] = std::move(good);
    }
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    std::ofstream out(argv[2], std::ios::binary);
    const int64_t np = static_cast<int64_t>(matchPairs.size());
    out.write(reinterpret_cast<const char*>(&np), 8);
    for (size_t p = 0; p < matchPairs.size(); ++p) {
        const int32_t lr[2] = {matchPairs[p].first, matchPairs[p].second};
        const int64_t cnt = static_cast<int64_t>(result[p].size());
        out.write(reinterpret_cast<const char*>(lr), 8);
        out.write(reinterpret_cast<const char*>(&cnt), 8);
        static_assert(sizeof(cv::DMatch) == 16, "cv::DMatch layout");
        out.write(reinterpret_cast<const char*>(result[p].data()), cnt * 16);
    }
    std::printf("pairs %lld threads %d ms %.2f fallbacks %d failed %d\n", static_cast<long long>(np), omp_get_max_threads(), ms,
                fallbacks, failed);
    return 0;
}
