// tests/stubs/opencv2/features2d.hpp — TEST INFRASTRUCTURE, not OpenCV.
//
// The build container has no OpenCV C++ headers, so host/cv_adapter.h — the cv::DescriptorMatcher a
// maintainer of the reference would include (PhotogrammetrieCli.cpp:378/:387) — could never go through a compiler.
// This stub declares just enough of the OpenCV 4 interface for that: cv::Mat as a strided matrix view, cv::DMatch,
// cv::Ptr, cv::Exception / CV_Error / CV_Assert, InputArray, and cv::DescriptorMatcher with the call flow of
// modules/features2d/src/matchers.cpp: the two-argument knnMatch(query, train, matches, k) CLONES the matcher
// (clone(true)), add()s the train descriptors and calls knnMatchImpl; match() is knnMatch(k = 1) flattened.
// Names, signatures and that flow follow OpenCV's public headers so that the adapter compiles unchanged against the
// real library; nothing here is shipped.
#pragma once
#include <cstdint>
#include <cstring>
#include <exception>
#include <memory>
#include <string>
#include <vector>

#define CV_8U 0
#define CV_32F 5
#define CV_MAT_DEPTH(t) ((t) & 7)
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn) - 1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)

namespace cv {

enum NormTypes { NORM_L2 = 4, NORM_HAMMING = 6 };
namespace Error { enum Code { StsError = -2, StsNotImplemented = -213, StsAssert = -215, GpuNotSupported = -216 }; }

class Exception : public std::exception {
public:
    Exception(int code_, const std::string& err_, const std::string& func_, const std::string& file_, int line_)
        : code(code_), err(err_), func(func_), file(file_), line(line_) {
        msg = file + ":" + std::to_string(line) + ": error: (" + std::to_string(code) + ") " + err + " in function '" + func + "'";
    }
    const char* what() const noexcept override { return msg.c_str(); }
    int code;
    std::string err, func, file, msg;
    int line;
};
[[noreturn]] inline void error(int code, const std::string& err, const char* func, const char* file, int line) {
    throw Exception(code, err, func, file, line);
}
#define CV_Error(code, msg) ::cv::error(code, msg, __func__, __FILE__, __LINE__)
#define CV_Assert(expr) do { if (!(expr)) ::cv::error(::cv::Error::StsAssert, #expr, __func__, __FILE__, __LINE__); } while (0)

template <class T> using Ptr = std::shared_ptr<T>;
template <class T, class... A> Ptr<T> makePtr(A&&... a) { return std::make_shared<T>(std::forward<A>(a)...); }

struct DMatch {
    DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(3.402823466e+38f) {}
    DMatch(int q, int t, int i, float d) : queryIdx(q), trainIdx(t), imgIdx(i), distance(d) {}
    int queryIdx, trainIdx, imgIdx;
    float distance;
};

// a non-owning (or vector-owning) strided matrix: what the adapter touches of cv::Mat
class Mat {
public:
    Mat() = default;
    Mat(int rows_, int cols_, int type_, void* data_, size_t step_ = 0)
        : rows(rows_), cols(cols_), data(static_cast<uint8_t*>(data_)), type_v(type_),
          step(step_ ? step_ : static_cast<size_t>(cols_) * (CV_MAT_DEPTH(type_) == CV_32F ? 4 : 1)) {}
    int type() const { return type_v; }
    int depth() const { return CV_MAT_DEPTH(type_v); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int rows = 0, cols = 0;
    uint8_t* data = nullptr;
    int type_v = 0;
    size_t step = 0;
};

class _InputArray {
public:
    _InputArray() = default;
    _InputArray(const Mat& m) : mat(&m) {}
    Mat getMat() const { return mat ? *mat : Mat(); }
    bool empty() const { return mat == nullptr || mat->empty(); }
private:
    const Mat* mat = nullptr;
};
typedef const _InputArray& InputArray;
typedef InputArray InputArrayOfArrays;
inline InputArray noArray() { static _InputArray none; return none; }

class DescriptorMatcher {
public:
    virtual ~DescriptorMatcher() = default;
    virtual void add(const std::vector<Mat>& descriptors) { trainDescCollection.insert(trainDescCollection.end(), descriptors.begin(), descriptors.end()); }
    virtual void clear() { trainDescCollection.clear(); }
    virtual bool isMaskSupported() const = 0;
    virtual Ptr<DescriptorMatcher> clone(bool emptyTrainData = false) const = 0;

    // matchers.cpp: DescriptorMatcher::knnMatch(query, train, matches, k, mask, compactResult)
    void knnMatch(InputArray queryDescriptors, InputArray trainDescriptors, std::vector<std::vector<DMatch>>& matches, int k) const {
        Ptr<DescriptorMatcher> tempMatcher = clone(true);
        tempMatcher->add(std::vector<Mat>(1, trainDescriptors.getMat()));
        tempMatcher->knnMatchImpl(queryDescriptors, matches, k, noArray(), false);
    }
    // matchers.cpp: DescriptorMatcher::match(query, train, matches, mask) = knnMatch(k = 1, compactResult = true) flattened
    void match(InputArray queryDescriptors, InputArray trainDescriptors, std::vector<DMatch>& matches) const {
        Ptr<DescriptorMatcher> tempMatcher = clone(true);
        tempMatcher->add(std::vector<Mat>(1, trainDescriptors.getMat()));
        std::vector<std::vector<DMatch>> knn;
        tempMatcher->knnMatchImpl(queryDescriptors, knn, 1, noArray(), true);
        matches.clear();
        for (const auto& row : knn) for (const auto& m : row) matches.push_back(m);
    }

protected:
    virtual void knnMatchImpl(InputArray queryDescriptors, std::vector<std::vector<DMatch>>& matches, int k, InputArrayOfArrays masks,
                              bool compactResult) = 0;
    virtual void radiusMatchImpl(InputArray queryDescriptors, std::vector<std::vector<DMatch>>& matches, float maxDistance,
                                 InputArrayOfArrays masks, bool compactResult) = 0;
    std::vector<Mat> trainDescCollection;
};

}  // namespace cv
