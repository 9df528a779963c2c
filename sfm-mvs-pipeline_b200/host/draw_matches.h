// draw_matches.h — the matching stage's only on-disk artifact in the reference: one picture per ShotMatches with both shots side
// by side and a line per match (PhotogrammetrieCli.cpp:174-199: cv::drawMatches(left, kpL, right, kpR, matches, out,
// Scalar::all(-1), Scalar::all(-1), {}, NOT_DRAW_SINGLE_POINTS) -> matches/<i><left>-<right>.jpg).  Header-only host code, no
// OpenCV and no GPU: the canvas layout (left image at x = 0, right image at x = left.cols, canvas height = max of the two, a
// circle at both keypoints and a line between them, single points not drawn) follows cv::drawMatches; OpenCV picks a RANDOM
// colour per match (Scalar::all(-1)), here the colour is a fixed function of the match index so that two runs write the same
// bytes.  Written as binary PPM (the reference writes JPEG through cv::imwrite; there is no JPEG encoder in this tree).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/sfmmatch.h"

namespace sfmhost {

struct RgbImage {
    int rows = 0, cols = 0;
    std::vector<uint8_t> px;      // rows x cols x 3, R G B
    void resize(int r, int c) { rows = r; cols = c; px.assign(static_cast<size_t>(r) * c * 3, 0); }
    void set(int x, int y, const uint8_t rgb[3]) {
        if (x < 0 || y < 0 || x >= cols || y >= rows) return;
        uint8_t* p = &px[(static_cast<size_t>(y) * cols + x) * 3];
        p[0] = rgb[0]; p[1] = rgb[1]; p[2] = rgb[2];
    }
};

// grey (channels = 1) or interleaved colour (channels = 3, `bgr` != 0 for OpenCV's channel order) into the canvas at (x0, 0)
inline void blit(RgbImage& canvas, int x0, const uint8_t* img, int rows, int cols, size_t step, int channels, int bgr) {
    for (int y = 0; y < rows && y < canvas.rows; ++y)
        for (int x = 0; x < cols && x0 + x < canvas.cols; ++x) {
            const uint8_t* s = img + static_cast<size_t>(y) * step + static_cast<size_t>(x) * channels;
            uint8_t rgb[3];
            if (channels == 1) rgb[0] = rgb[1] = rgb[2] = s[0];
            else { rgb[0] = s[bgr ? 2 : 0]; rgb[1] = s[1]; rgb[2] = s[bgr ? 0 : 2]; }
            canvas.set(x0 + x, y, rgb);
        }
}

inline void draw_line(RgbImage& c, int x0, int y0, int x1, int y1, const uint8_t rgb[3]) {
    const int dx = std::abs(x1 - x0), sx = x0 < x1 ? 1 : -1, dy = -std::abs(y1 - y0), sy = y0 < y1 ? 1 : -1;
    int err = dx + dy;
    for (;;) {
        c.set(x0, y0, rgb);
        if (x0 == x1 && y0 == y1) break;
        const int e2 = 2 * err;
        if (e2 >= dy) { err += dy; x0 += sx; }
        if (e2 <= dx) { err += dx; y0 += sy; }
    }
}

inline void draw_circle(RgbImage& c, int cx, int cy, int radius, const uint8_t rgb[3]) {
    int x = radius, y = 0, err = 1 - radius;                // midpoint circle
    while (x >= y) {
        const int pts[8][2] = {{x, y}, {y, x}, {-y, x}, {-x, y}, {-x, -y}, {-y, -x}, {y, -x}, {x, -y}};
        for (auto& p : pts) c.set(cx + p[0], cy + p[1], rgb);
        ++y;
        if (err < 0) err += 2 * y + 1;
        else { --x; err += 2 * (y - x) + 1; }
    }
}

// colour of match k: a fixed walk through hue space (OpenCV draws a random colour per match)
inline void match_colour(size_t k, uint8_t rgb[3]) {
    const double h = std::fmod(0.61803398875 * static_cast<double>(k), 1.0) * 6.0;
    const int i = static_cast<int>(h);
    const double f = h - i;
    const uint8_t v = 255, p = 40, q = static_cast<uint8_t>(255 - 215 * f), t = static_cast<uint8_t>(40 + 215 * f);
    const uint8_t tab[6][3] = {{v, t, p}, {q, v, p}, {p, v, t}, {p, q, v}, {t, p, v}, {v, p, q}};
    rgb[0] = tab[i % 6][0]; rgb[1] = tab[i % 6][1]; rgb[2] = tab[i % 6][2];
}

// cv::drawMatches with NOT_DRAW_SINGLE_POINTS: kp = (x, y) float pairs, `stride` bytes apart (8 packed, 24 = sfm_keypoint,
// 28 = cv::KeyPoint)
inline RgbImage draw_matches(const uint8_t* left, int lrows, int lcols, size_t lstep, int lch, const void* kp_left, size_t lstride,
                             const uint8_t* right, int rrows, int rcols, size_t rstep, int rch, const void* kp_right, size_t rstride,
                             const sfm_dmatch* matches, size_t n_matches, int bgr = 0) {
    RgbImage out;
    out.resize(std::max(lrows, rrows), lcols + rcols);
    blit(out, 0, left, lrows, lcols, lstep, lch, bgr);
    blit(out, lcols, right, rrows, rcols, rstep, rch, bgr);
    for (size_t k = 0; k < n_matches; ++k) {
        const float* a = reinterpret_cast<const float*>(static_cast<const uint8_t*>(kp_left) + lstride * matches[k].queryIdx);
        const float* b = reinterpret_cast<const float*>(static_cast<const uint8_t*>(kp_right) + rstride * matches[k].trainIdx);
        uint8_t rgb[3];
        match_colour(k, rgb);
        const int ax = static_cast<int>(std::lrint(a[0])), ay = static_cast<int>(std::lrint(a[1]));
        const int bx = static_cast<int>(std::lrint(b[0])) + lcols, by = static_cast<int>(std::lrint(b[1]));
        draw_circle(out, ax, ay, 3, rgb);
        draw_circle(out, bx, by, 3, rgb);
        draw_line(out, ax, ay, bx, by, rgb);
    }
    return out;
}

inline bool write_ppm(const std::string& path, const RgbImage& img) {
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    std::fprintf(f, "P6\n%d %d\n255\n", img.cols, img.rows);
    const bool ok = std::fwrite(img.px.data(), 1, img.px.size(), f) == img.px.size();
    std::fclose(f);
    return ok;
}

}  // namespace sfmhost
