// image.cpp — host-side grey conversion in front of the feature extractor.
//
// Shot::loadImage (CameraShot.cpp) returns the photograph as OpenCV decoded it (8-bit BGR); cv::SIFT::detectAndCompute
// converts it with cvtColor(COLOR_BGR2GRAY) before anything else (sift.dispatch.cpp createInitialImage).  For 8-bit data
// that conversion is integer arithmetic, restated here:  (B * 3735 + G * 19235 + R * 9798 + 2^14) >> 15
// (BT.601 weights in 15-bit fixed point; equal to cv2 4.13's cvtColor on every (B, G, R), tests/test_cabi_cpu.py).
#include <cstddef>
#include <cstdint>

#include "../../include/sfmmatch.h"

extern "C" int sfm_gray_from_bgr(const uint8_t* src, int rows, int cols, size_t step_bytes, int channels, int rgb_order,
                                 uint8_t* gray, size_t gray_step_bytes) {
    if (rows < 0 || cols < 0 || (channels != 3 && channels != 4)) return SFM_ERR_INVALID;
    if (rows == 0 || cols == 0) return SFM_OK;
    if (!src || !gray) return SFM_ERR_INVALID;
    if (step_bytes == 0) step_bytes = static_cast<size_t>(cols) * channels;
    if (gray_step_bytes == 0) gray_step_bytes = static_cast<size_t>(cols);
    if (step_bytes < static_cast<size_t>(cols) * channels || gray_step_bytes < static_cast<size_t>(cols)) return SFM_ERR_INVALID;
    const int ib = rgb_order ? 2 : 0, ir = rgb_order ? 0 : 2;
    for (int y = 0; y < rows; ++y) {
        const uint8_t* p = src + static_cast<size_t>(y) * step_bytes;
        uint8_t* q = gray + static_cast<size_t>(y) * gray_step_bytes;
        for (int x = 0; x < cols; ++x, p += channels)
            q[x] = static_cast<uint8_t>((p[ib] * 3735 + p[1] * 19235 + p[ir] * 9798 + (1 << 14)) >> 15);
    }
    return SFM_OK;
}
