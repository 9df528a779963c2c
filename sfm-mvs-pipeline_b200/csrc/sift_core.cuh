// sift_core.cuh — per-keypoint arithmetic of the feature-extraction stage (SfM::extractFeatures, SfM.cpp:577-597, with the
// detector PhotogrammetrieCli.cpp:342-357 configures: cv::SIFT).  The functions follow the published algorithm of OpenCV's
// sift.simd.hpp (adjustLocalExtrema, calcOrientationHist, calcSIFTDescriptor) in float32 and in the same operation and
// accumulation order; they are __host__ __device__ so that the kernels of sift.cu and the host test harness
// (tests/sift_host_harness.cpp) run the very same code.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define SIFT_HD __host__ __device__ __forceinline__
#else
#define SIFT_HD inline
#endif

namespace sfm {
namespace sift {

constexpr int kMaxOctaves = 16;
constexpr int kMaxBlurRadius = 32;
constexpr int kImgBorder = 5;            // SIFT_IMG_BORDER
constexpr int kMaxInterpSteps = 5;       // SIFT_MAX_INTERP_STEPS
constexpr int kOriBins = 36;             // SIFT_ORI_HIST_BINS
constexpr int kDescWidth = 4;            // SIFT_DESCR_WIDTH
constexpr int kDescBins = 8;             // SIFT_DESCR_HIST_BINS
constexpr int kDescHistLen = (kDescWidth + 2) * (kDescWidth + 2) * (kDescBins + 2);   // 360
constexpr int kDescLen = kDescWidth * kDescWidth * kDescBins;                          // 128

// Gaussian pyramid of one image: level i of octave o is a dense w[o] x h[o] float image at
// base + off[o] + i * w[o] * h[o]; n_layers + 3 levels per octave.  DoG values are differences taken on the fly.
struct PyramidView {
    const float* base;
    int n_octaves, n_layers;
    int w[kMaxOctaves], h[kMaxOctaves];
    int64_t off[kMaxOctaves];
    SIFT_HD const float* level(int o, int i) const { return base + off[o] + static_cast<int64_t>(i) * w[o] * h[o]; }
    SIFT_HD float gauss(int o, int i, int r, int c) const { return level(o, i)[static_cast<int64_t>(r) * w[o] + c]; }
    SIFT_HD float dog(int o, int i, int r, int c) const {            // buildDoGPyramid: level i + 1 minus level i
        const int64_t p = static_cast<int64_t>(r) * w[o] + c;
        return level(o, i + 1)[p] - level(o, i)[p];
    }
};

struct Keypoint {            // cv::KeyPoint without class_id
    float x, y, size, angle, response;
    int32_t octave;
};

struct Candidate {           // a 26-neighbour extremum of the DoG pyramid
    int32_t octave, layer, r, c;
};

SIFT_HD int cv_round(float v) { return static_cast<int>(rintf(v)); }      // cvRound: round half to even

// cv::hal::fastAtan2, degrees in [0, 360)
SIFT_HD float fast_atan2_deg(float y, float x) {
    const float p1 = 0.9997878412794807f * static_cast<float>(180 / 3.14159265358979323846);
    const float p3 = -0.3258083974640975f * static_cast<float>(180 / 3.14159265358979323846);
    const float p5 = 0.1555786518463281f * static_cast<float>(180 / 3.14159265358979323846);
    const float p7 = -0.04432655554792128f * static_cast<float>(180 / 3.14159265358979323846);
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + 2.220446049250313e-16f);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + 2.220446049250313e-16f);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

// Matx33f::solve(b, DECOMP_LU): closed form, singular -> zero vector
SIFT_HD void solve3(const float a[3][3], const float b[3], float x[3]) {
    float d = a[0][0] * (a[1][1] * a[2][2] - a[2][1] * a[1][2]) - a[0][1] * (a[1][0] * a[2][2] - a[2][0] * a[1][2]) +
              a[0][2] * (a[1][0] * a[2][1] - a[2][0] * a[1][1]);
    if (d == 0.f) { x[0] = x[1] = x[2] = 0.f; return; }
    d = 1.f / d;
    x[0] = d * (b[0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) - a[0][1] * (b[1] * a[2][2] - a[1][2] * b[2]) +
                a[0][2] * (b[1] * a[2][1] - a[1][1] * b[2]));
    x[1] = d * (a[0][0] * (b[1] * a[2][2] - a[1][2] * b[2]) - b[0] * (a[1][0] * a[2][2] - a[1][2] * a[2][0]) +
                a[0][2] * (a[1][0] * b[2] - b[1] * a[2][0]));
    x[2] = d * (a[0][0] * (a[1][1] * b[2] - b[1] * a[2][1]) - a[0][1] * (a[1][0] * b[2] - b[1] * a[2][0]) +
                b[0] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]));
}

// findScaleSpaceExtremaT::process, the test of one pixel: |val| > threshold and val is the maximum (val > 0) or the
// minimum (val < 0) of its 26 neighbours, ties included
SIFT_HD bool is_extremum(const PyramidView& P, int o, int i, int r, int c, float threshold) {
    const float val = P.dog(o, i, r, c);
    if (!(fabsf(val) > threshold)) return false;
    if (val > 0) {
        for (int l = i - 1; l <= i + 1; ++l)
            for (int dr = -1; dr <= 1; ++dr)
                for (int dc = -1; dc <= 1; ++dc)
                    if (val < P.dog(o, l, r + dr, c + dc)) return false;
        return true;
    }
    if (val < 0) {
        for (int l = i - 1; l <= i + 1; ++l)
            for (int dr = -1; dr <= 1; ++dr)
                for (int dc = -1; dc <= 1; ++dc)
                    if (val > P.dog(o, l, r + dr, c + dc)) return false;
        return true;
    }
    return false;
}

// adjustLocalExtrema: sub-pixel / sub-scale position by up to five Newton steps, contrast and edge rejection.
// On success fills kpt (octave coordinates scaled to the doubled base image, as OpenCV does before the firstOctave
// correction) and the integer position (layer, r, c) the orientation histogram is taken at.
SIFT_HD bool adjust_local_extrema(const PyramidView& P, int octv, int& layer, int& r, int& c, float contrast_threshold,
                                  float edge_threshold, float sigma, Keypoint& kpt) {
    const float img_scale = 1.f / 255.f;
    const float deriv_scale = img_scale * 0.5f;
    const float second_deriv_scale = img_scale;
    const float cross_deriv_scale = img_scale * 0.25f;
    const int n_layers = P.n_layers;
    const int cols = P.w[octv], rows = P.h[octv];
    float xi = 0, xr = 0, xc = 0;
    int i = 0;
    for (; i < kMaxInterpSteps; ++i) {
        const float dD[3] = {(P.dog(octv, layer, r, c + 1) - P.dog(octv, layer, r, c - 1)) * deriv_scale,
                             (P.dog(octv, layer, r + 1, c) - P.dog(octv, layer, r - 1, c)) * deriv_scale,
                             (P.dog(octv, layer + 1, r, c) - P.dog(octv, layer - 1, r, c)) * deriv_scale};
        const float v2 = P.dog(octv, layer, r, c) * 2.f;
        const float dxx = (P.dog(octv, layer, r, c + 1) + P.dog(octv, layer, r, c - 1) - v2) * second_deriv_scale;
        const float dyy = (P.dog(octv, layer, r + 1, c) + P.dog(octv, layer, r - 1, c) - v2) * second_deriv_scale;
        const float dss = (P.dog(octv, layer + 1, r, c) + P.dog(octv, layer - 1, r, c) - v2) * second_deriv_scale;
        const float dxy = (P.dog(octv, layer, r + 1, c + 1) - P.dog(octv, layer, r + 1, c - 1) - P.dog(octv, layer, r - 1, c + 1) +
                           P.dog(octv, layer, r - 1, c - 1)) * cross_deriv_scale;
        const float dxs = (P.dog(octv, layer + 1, r, c + 1) - P.dog(octv, layer + 1, r, c - 1) - P.dog(octv, layer - 1, r, c + 1) +
                           P.dog(octv, layer - 1, r, c - 1)) * cross_deriv_scale;
        const float dys = (P.dog(octv, layer + 1, r + 1, c) - P.dog(octv, layer + 1, r - 1, c) - P.dog(octv, layer - 1, r + 1, c) +
                           P.dog(octv, layer - 1, r - 1, c)) * cross_deriv_scale;
        const float H[3][3] = {{dxx, dxy, dxs}, {dxy, dyy, dys}, {dxs, dys, dss}};
        float X[3];
        solve3(H, dD, X);
        xi = -X[2]; xr = -X[1]; xc = -X[0];
        if (fabsf(xi) < 0.5f && fabsf(xr) < 0.5f && fabsf(xc) < 0.5f) break;
        const float lim = static_cast<float>(2147483647 / 3);
        if (!(fabsf(xi) <= lim) || !(fabsf(xr) <= lim) || !(fabsf(xc) <= lim)) return false;     // also rejects NaN
        c += cv_round(xc);
        r += cv_round(xr);
        layer += cv_round(xi);
        if (layer < 1 || layer > n_layers || c < kImgBorder || c >= cols - kImgBorder || r < kImgBorder || r >= rows - kImgBorder)
            return false;
    }
    if (i >= kMaxInterpSteps) return false;
    {
        const float dD[3] = {(P.dog(octv, layer, r, c + 1) - P.dog(octv, layer, r, c - 1)) * deriv_scale,
                             (P.dog(octv, layer, r + 1, c) - P.dog(octv, layer, r - 1, c)) * deriv_scale,
                             (P.dog(octv, layer + 1, r, c) - P.dog(octv, layer - 1, r, c)) * deriv_scale};
        const float t = dD[0] * xc + dD[1] * xr + dD[2] * xi;
        const float contr = P.dog(octv, layer, r, c) * img_scale + t * 0.5f;
        if (fabsf(contr) * n_layers < contrast_threshold) return false;
        const float v2 = P.dog(octv, layer, r, c) * 2.f;
        const float dxx = (P.dog(octv, layer, r, c + 1) + P.dog(octv, layer, r, c - 1) - v2) * second_deriv_scale;
        const float dyy = (P.dog(octv, layer, r + 1, c) + P.dog(octv, layer, r - 1, c) - v2) * second_deriv_scale;
        const float dxy = (P.dog(octv, layer, r + 1, c + 1) - P.dog(octv, layer, r + 1, c - 1) - P.dog(octv, layer, r - 1, c + 1) +
                           P.dog(octv, layer, r - 1, c - 1)) * cross_deriv_scale;
        const float tr = dxx + dyy;
        const float det = dxx * dyy - dxy * dxy;
        if (det <= 0 || tr * tr * edge_threshold >= (edge_threshold + 1) * (edge_threshold + 1) * det) return false;
        kpt.response = fabsf(contr);
    }
    const float scale = static_cast<float>(1 << octv);
    kpt.x = (c + xc) * scale;
    kpt.y = (r + xr) * scale;
    kpt.octave = octv + (layer << 8) + (static_cast<int>(rint((static_cast<double>(xi) + 0.5) * 255)) << 16);
    kpt.size = sigma * powf(2.f, (layer + xi) / n_layers) * scale * 2;
    kpt.angle = -1.f;
    return true;
}

// calcOrientationHist, one sample: g enumerates the (2 radius + 1)^2 window row by row (OpenCV's k order once the samples
// outside the image are skipped); false = outside the image
SIFT_HD bool orientation_sample(const float* img, int cols, int rows, int pr, int pc, int radius, float expf_scale, int g,
                                int& bin, float& val) {
    const int side = 2 * radius + 1;
    const int i = g / side - radius, j = g % side - radius;
    const int y = pr + i, x = pc + j;
    if (y <= 0 || y >= rows - 1 || x <= 0 || x >= cols - 1) return false;
    const float* p = img + static_cast<int64_t>(y) * cols + x;
    const float dx = p[1] - p[-1];
    const float dy = p[-cols] - p[cols];
    const float w = expf(static_cast<float>(i * i + j * j) * expf_scale);
    const float ori = fast_atan2_deg(dy, dx);
    const float mag = sqrtf(dx * dx + dy * dy);
    bin = cv_round((kOriBins / 360.f) * ori);
    if (bin >= kOriBins) bin -= kOriBins;
    if (bin < 0) bin += kOriBins;
    val = w * mag;
    return true;
}

// [1 4 6 4 1] / 16 smoothing of the raw histogram + the peak search of findScaleSpaceExtremaT::process: angles[] receives
// the orientation of every peak >= 0.8 * maximum (degrees, OpenCV's 360 - bin convention); returns their number
SIFT_HD int orientation_finish(const float* temphist, float* angles) {
    const int n = kOriBins;
    float hist[kOriBins];
    float maxval = 0.f;
    for (int k = 0; k < n; ++k) {
        const float m2 = temphist[(k + n - 2) % n], m1 = temphist[(k + n - 1) % n], p1 = temphist[(k + 1) % n], p2 = temphist[(k + 2) % n];
        hist[k] = (m2 + p2) * (1.f / 16.f) + (m1 + p1) * (4.f / 16.f) + temphist[k] * (6.f / 16.f);
        if (k == 0 || hist[k] > maxval) maxval = hist[k];
    }
    const float mag_thr = maxval * 0.8f;
    int count = 0;
    for (int j = 0; j < n; ++j) {
        const int l = j > 0 ? j - 1 : n - 1;
        const int r2 = j < n - 1 ? j + 1 : 0;
        if (hist[j] > hist[l] && hist[j] > hist[r2] && hist[j] >= mag_thr) {
            float bin = j + 0.5f * (hist[l] - hist[r2]) / (hist[l] - 2 * hist[j] + hist[r2]);
            bin = bin < 0 ? n + bin : (bin >= n ? bin - n : bin);
            float angle = 360.f - (360.f / n) * bin;
            if (fabsf(angle - 360.f) < 1.1920929e-07f) angle = 0.f;
            angles[count++] = angle;
        }
    }
    return count;
}

// the serial form (host harness; the kernel walks the same samples with one warp, see sift.cu)
SIFT_HD int orientation_peaks(const PyramidView& P, int octv, int layer, int pr, int pc, int radius, float sigma, float* angles) {
    const float* img = P.level(octv, layer);
    const int cols = P.w[octv], rows = P.h[octv];
    const float expf_scale = -1.f / (2.f * sigma * sigma);
    float temphist[kOriBins];
    for (int k = 0; k < kOriBins; ++k) temphist[k] = 0.f;
    const int total = (2 * radius + 1) * (2 * radius + 1);
    for (int g = 0; g < total; ++g) {
        int bin;
        float val;
        if (orientation_sample(img, cols, rows, pr, pc, radius, expf_scale, g, bin, val)) temphist[bin] += val;
    }
    return orientation_finish(temphist, angles);
}

// KeyPoint12_LessThan of KeyPointsFilter::removeDuplicatedSorted (keypoint.cpp); class_id is -1 everywhere
SIFT_HD bool keypoint_less(const Keypoint& a, const Keypoint& b) {
    if (a.x != b.x) return a.x < b.x;
    if (a.y != b.y) return a.y < b.y;
    if (a.size != b.size) return a.size > b.size;
    if (a.angle != b.angle) return a.angle < b.angle;
    if (a.response != b.response) return a.response > b.response;
    if (a.octave != b.octave) return a.octave > b.octave;
    return false;
}
SIFT_HD bool keypoint_duplicate(const Keypoint& a, const Keypoint& b) {
    return a.x == b.x && a.y == b.y && a.size == b.size && a.angle == b.angle;
}

// unpackOctave (sift.dispatch.cpp)
SIFT_HD void unpack_octave(int32_t field, int& octave, int& layer, float& scale) {
    octave = field & 255;
    layer = (field >> 8) & 255;
    octave = octave < 128 ? octave : (-128 | octave);
    scale = octave >= 0 ? 1.f / (1 << octave) : static_cast<float>(1 << -octave);
}

// calcSIFTDescriptor, split into the per-keypoint frame, the votes of one sample and the final normalisation
struct DescFrame {
    int px, py, radius;
    float cos_t, sin_t, ori, exp_scale, bins_per_rad;
};

SIFT_HD DescFrame descriptor_frame(int cols, int rows, float ptx, float pty, float ori, float scl) {
    const int d = kDescWidth, n = kDescBins;
    DescFrame F;
    F.px = cv_round(ptx);
    F.py = cv_round(pty);
    F.ori = ori;
    F.cos_t = cosf(ori * static_cast<float>(3.14159265358979323846 / 180));
    F.sin_t = sinf(ori * static_cast<float>(3.14159265358979323846 / 180));
    F.bins_per_rad = n / 360.f;
    F.exp_scale = -1.f / (d * d * 0.5f);
    const float hist_width = 3.f * scl;                                    // SIFT_DESCR_SCL_FCTR
    int radius = cv_round(hist_width * 1.4142135623730951f * (d + 1) * 0.5f);
    const int diag = static_cast<int>(sqrt(static_cast<double>(cols) * cols + static_cast<double>(rows) * rows));
    F.radius = radius < diag ? radius : diag;
    F.cos_t /= hist_width;
    F.sin_t /= hist_width;
    return F;
}

// offsets of the eight trilinear votes of a sample from its base histogram index, in OpenCV's update order
SIFT_HD int descriptor_vote_offset(int v) {
    const int d = kDescWidth, n = kDescBins;
    return ((v >> 2) & 1) * (d + 2) * (n + 2) + ((v >> 1) & 1) * (n + 2) + (v & 1);
}

// sample (i, j) of the window: false = outside the rotated 4 x 4 grid or the image; else base index + eight votes
SIFT_HD bool descriptor_sample(const float* img, int cols, int rows, const DescFrame& F, int i, int j, int& idx, float v[8]) {
    const int d = kDescWidth, n = kDescBins;
    const float c_rot = j * F.cos_t - i * F.sin_t;
    const float r_rot = j * F.sin_t + i * F.cos_t;
    float rbin = r_rot + d / 2 - 0.5f;
    float cbin = c_rot + d / 2 - 0.5f;
    const int r = F.py + i, c = F.px + j;
    if (!(rbin > -1 && rbin < d && cbin > -1 && cbin < d && r > 0 && r < rows - 1 && c > 0 && c < cols - 1)) return false;
    const float* p = img + static_cast<int64_t>(r) * cols + c;
    const float dx = p[1] - p[-1];
    const float dy = p[-cols] - p[cols];
    const float w = expf((c_rot * c_rot + r_rot * r_rot) * F.exp_scale);
    const float angle = fast_atan2_deg(dy, dx);
    const float mag = sqrtf(dx * dx + dy * dy) * w;
    float obin = (angle - F.ori) * F.bins_per_rad;
    const int r0 = static_cast<int>(floorf(rbin)), c0 = static_cast<int>(floorf(cbin));
    int o0 = static_cast<int>(floorf(obin));
    rbin -= r0; cbin -= c0; obin -= o0;
    if (o0 < 0) o0 += n;
    if (o0 >= n) o0 -= n;
    const float v_r1 = mag * rbin, v_r0 = mag - v_r1;
    const float v_rc11 = v_r1 * cbin, v_rc10 = v_r1 - v_rc11;
    const float v_rc01 = v_r0 * cbin, v_rc00 = v_r0 - v_rc01;
    const float v_rco111 = v_rc11 * obin, v_rco110 = v_rc11 - v_rco111;
    const float v_rco101 = v_rc10 * obin, v_rco100 = v_rc10 - v_rco101;
    const float v_rco011 = v_rc01 * obin, v_rco010 = v_rc01 - v_rco011;
    const float v_rco001 = v_rc00 * obin, v_rco000 = v_rc00 - v_rco001;
    idx = ((r0 + 1) * (d + 2) + c0 + 1) * (n + 2) + o0;
    v[0] = v_rco000; v[1] = v_rco001; v[2] = v_rco010; v[3] = v_rco011;
    v[4] = v_rco100; v[5] = v_rco101; v[6] = v_rco110; v[7] = v_rco111;
    return true;
}

// histogram index of descriptor element e = (i * 4 + j) * 8 + k
SIFT_HD int descriptor_element_index(int e) {
    const int d = kDescWidth, n = kDescBins;
    const int k = e % n, j = (e / n) % d, i = e / (n * d);
    return ((i + 1) * (d + 2) + (j + 1)) * (n + 2) + k;
}

// circular orientation bins folded in; returns the clipping threshold 0.2 * |raw| (sums in OpenCV's element order)
SIFT_HD float descriptor_fold_and_threshold(float* hist, int hstride) {
    const int d = kDescWidth, n = kDescBins;
    float nrm2 = 0.f;
    for (int i = 0; i < d; ++i)
        for (int j = 0; j < d; ++j) {
            const int idx = ((i + 1) * (d + 2) + (j + 1)) * (n + 2);
            hist[idx * hstride] += hist[(idx + n) * hstride];
            hist[(idx + 1) * hstride] += hist[(idx + n + 1) * hstride];
            for (int k = 0; k < n; ++k) { const float v = hist[(idx + k) * hstride]; nrm2 += v * v; }
        }
    return sqrtf(nrm2) * 0.2f;                                             // SIFT_DESCR_MAG_THR
}
// clipped values written back; returns the scale factor 512 / max(|clipped|, FLT_EPSILON)
SIFT_HD float descriptor_clip_and_scale(float* hist, int hstride, float thr) {
    float nrm2 = 0.f;
    for (int e = 0; e < kDescLen; ++e) {
        const int idx = descriptor_element_index(e);
        float v = hist[idx * hstride];
        v = v < thr ? v : thr;
        hist[idx * hstride] = v;
        nrm2 += v * v;
    }
    const float s = sqrtf(nrm2);
    return 512.f / (s > 1.1920929e-07f ? s : 1.1920929e-07f);             // SIFT_INT_DESCR_FCTR
}
SIFT_HD uint8_t descriptor_quantise(float v, float f) {                   // saturate_cast<uchar>(v * f)
    const float q = rintf(v * f);
    return static_cast<uint8_t>(q < 0.f ? 0.f : (q > 255.f ? 255.f : q));
}

// the serial form (host harness; the kernel walks the same samples with one warp, see sift.cu): hist = kDescHistLen floats
// of scratch with element stride hstride; dst = 128 bytes
SIFT_HD void sift_descriptor(const float* img, int cols, int rows, float ptx, float pty, float ori, float scl, float* hist,
                             int hstride, uint8_t* dst) {
    const DescFrame F = descriptor_frame(cols, rows, ptx, pty, ori, scl);
    for (int k = 0; k < kDescHistLen; ++k) hist[k * hstride] = 0.f;
    for (int i = -F.radius; i <= F.radius; ++i)
        for (int j = -F.radius; j <= F.radius; ++j) {
            int idx;
            float v[8];
            if (!descriptor_sample(img, cols, rows, F, i, j, idx, v)) continue;
            for (int k = 0; k < 8; ++k) hist[(idx + descriptor_vote_offset(k)) * hstride] += v[k];
        }
    const float thr = descriptor_fold_and_threshold(hist, hstride);
    const float f = descriptor_clip_and_scale(hist, hstride, thr);
    for (int e = 0; e < kDescLen; ++e) dst[e] = descriptor_quantise(hist[descriptor_element_index(e) * hstride], f);
}

}  // namespace sift
}  // namespace sfm
