// orb.cu — the reference's second detector on the device: cv::ORB::create(featureLimit) (PhotogrammetrieCli.cpp:347-348,
// run-scripts/run-orb-sequence.sh:4), detect() + compute() as SfM::extractFeatures calls them (SfM.cpp:584-590).
//
// cv::ORB lives in OpenCV (un-vendored); what is built here is its published algorithm, operation for operation like the
// numpy restatement the tests check it against (pinned against cv2 golden vectors: identical keypoint sets, responses and
// descriptors).  Everything that decides a keypoint is integer or order-fixed float arithmetic, so device == restatement:
//
//   orb_coeff_kernel / orb_resize_kernel   8-level pyramid, scale 1.2^level, each level from the PREVIOUS one with
//                                          INTER_LINEAR_EXACT (8.8 fixed-point weights, (sum + 2^15) >> 16)
//   orb_fast_kernel                        FAST-9/16 score (cornerScore<16>: largest threshold that keeps the corner) per pixel
//   orb_nms_kernel                         3 x 3 non-maximum suppression (strict >), edge-threshold border, score histogram
//   orb_select_score_kernel                retainBest(2 x quota) by FAST score: threshold from the histogram (ties kept)
//   orb_harris_kernel                      Harris response (7 x 7 block, k = 0.04) of the survivors
//   orb_select_harris_kernel               retainBest(quota) by response: radix select on the float bits (ties kept)
//   orb_row_count / orb_row_scan / orb_emit_kernel   ordered compaction: keypoints come out sorted by (level, y, x)
//                                          (cv::ORB's own order is a by-product of std::nth_element; the SET is the same)
//   orb_blur_rows / orb_blur_cols_kernel   GaussianBlur(7 x 7, sigma 2) of every level as the generic separable float filter
//                                          (the level is a submatrix of OpenCV's pyramid buffer, so the 8-bit fixed-point
//                                          path is not taken): row pass in tap order, column pass centre + symmetric pairs
//   orb_describe_kernel                    one warp per keypoint: intensity-centroid angle (integer moments over the circular
//                                          patch, cv::fastAtan2), then 256 comparisons of the rotated bit pattern, one
//                                          descriptor byte per lane
// Compiled with --fmad=false (multiply and add round separately, like numpy and the non-FMA OpenCV build).
#include <cmath>
#include <cstdio>
#include <string>
#include <vector>

#include "kernels.h"
#include "sift_core.cuh"

namespace sfm {

namespace {

constexpr int kOrbLevels = 8, kOrbEdge = 31, kOrbHalfPatch = 15, kOrbFastThreshold = 20, kOrbPatch = 31;
constexpr int kOrbPatchPixels = 749;           // pixels of the circular patch (umax table)

struct OrbLevels {
    int w[kOrbLevels], h[kOrbLevels];
    int64_t off[kOrbLevels];                   // pixel offset of the level in the flat per-pixel buffers
    int64_t row_off[kOrbLevels];               // first global row number of the level
    float scale[kOrbLevels];
};

__constant__ int c_pattern[256][4] = {
#include "orb_pattern.inc"
};
__constant__ int8_t c_patch_u[kOrbPatchPixels], c_patch_v[kOrbPatchPixels];
__constant__ float c_gauss[7];
// Bresenham circle of radius 3 in OpenCV's order
__constant__ int c_circle_dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
__constant__ int c_circle_dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

__device__ __forceinline__ int reflect101_orb(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * (n - 1) - p;
    return p;
}

// interpolationLinear<ufixedpoint16>::getCoeffs for one axis: offset + weight of the right tap in 1/256
__global__ void orb_coeff_kernel(int dst_n, int src_n, int* __restrict__ ofs, int* __restrict__ c1) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= dst_n) return;
    const double scale = static_cast<double>(src_n) / static_cast<double>(dst_n);
    const double f = scale * (static_cast<double>(i) + 0.5) - 0.5;
    const int fi = static_cast<int>(floor(f));
    int o = 0, c = 0;
    if (fi >= 0 && src_n > 1) {
        if (fi < src_n - 1) { o = fi; c = static_cast<int>(rint((f - static_cast<double>(fi)) * 256.0)); }
        else o = src_n - 1;
    } else if (fi >= 0) o = src_n - 1;
    ofs[i] = o;
    c1[i] = c;
}

__global__ void __launch_bounds__(256) orb_resize_kernel(const uint8_t* __restrict__ src, int sw, int sh, uint8_t* __restrict__ dst,
                                                         int dw, int dh, const int* __restrict__ ox, const int* __restrict__ cx,
                                                         const int* __restrict__ oy, const int* __restrict__ cy) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= dw || y >= dh) return;
    const int x0 = ox[x], x1 = min(x0 + 1, sw - 1), y0 = oy[y], y1 = min(y0 + 1, sh - 1);
    const int wx = cx[x], wy = cy[y];
    const uint8_t* r0 = src + static_cast<size_t>(y0) * sw;
    const uint8_t* r1 = src + static_cast<size_t>(y1) * sw;
    const int h0 = r0[x0] * (256 - wx) + r0[x1] * wx;           // 8.8 fixed point
    const int h1 = r1[x0] * (256 - wx) + r1[x1] * wx;
    const int v = (h0 * (256 - wy) + h1 * wy + 32768) >> 16;
    dst[static_cast<size_t>(y) * dw + x] = static_cast<uint8_t>(min(max(v, 0), 255));
}

// cornerScore<16> where the pixel is a FAST-9 corner for the threshold, 0 elsewhere.  Window minima over the 16-cycle by
// doubling: m9[k] = min(d[k .. k+8]);  bright score = max_k m9 on d = centre - circle, dark score the same on circle - centre.
__global__ void __launch_bounds__(256) orb_fast_kernel(const uint8_t* __restrict__ img, int w, int h, int threshold,
                                                       uint8_t* __restrict__ score) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= w || y >= h) return;
    int s = 0;
    if (x >= 3 && y >= 3 && x < w - 3 && y < h - 3) {
        const uint8_t* p = img + static_cast<size_t>(y) * w + x;
        const int v = p[0];
        // d = centre - circle (bright centre), e = circle - centre (dark centre): two min-chains.  (Written without negating a
        // max-chain: nvcc 12.9 folded  max(lo9, -hi9)  into one VIMNMX3 and lost the negation on sm_100a.)
        int d[16], e[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int q = p[c_circle_dy[k] * w + c_circle_dx[k]];
            d[k] = v - q;
            e[k] = q - v;
        }
        int d2[16], e2[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) { d2[k] = min(d[k], d[(k + 1) & 15]); e2[k] = min(e[k], e[(k + 1) & 15]); }            // 2
        int d4[16], e4[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) { d4[k] = min(d2[k], d2[(k + 2) & 15]); e4[k] = min(e2[k], e2[(k + 2) & 15]); }        // 4
        int best = -1000000;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int d9 = min(min(d4[k], d4[(k + 4) & 15]), d[(k + 8) & 15]);                                               // 9
            const int e9 = min(min(e4[k], e4[(k + 4) & 15]), e[(k + 8) & 15]);
            best = max(best, max(d9, e9));
        }
        if (best > threshold) s = best - 1;
    }
    score[static_cast<size_t>(y) * w + x] = static_cast<uint8_t>(s);
}

// non-maximum suppression + KeyPointsFilter::runByImageBorder(edgeThreshold); survivors keep their score in `cand`
__global__ void __launch_bounds__(256) orb_nms_kernel(const uint8_t* __restrict__ score, int w, int h, uint8_t* __restrict__ cand,
                                                      unsigned* __restrict__ hist /* 256 bins of this level */) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= w || y >= h) return;
    int keep = 0;
    const size_t i = static_cast<size_t>(y) * w + x;
    if (x >= kOrbEdge && y >= kOrbEdge && x < w - kOrbEdge && y < h - kOrbEdge) {
        const int s = score[i];
        if (s > 0) {
            const uint8_t* p = score + i;
            if (s > p[-1] && s > p[1] && s > p[-w - 1] && s > p[-w] && s > p[-w + 1] && s > p[w - 1] && s > p[w] && s > p[w + 1]) {
                keep = s;
                atomicAdd(hist + s, 1u);
            }
        }
    }
    cand[i] = static_cast<uint8_t>(keep);
}

// retainBest(n) on integer scores: the n-th best value (everything >= it stays); 0 = keep all
__global__ void orb_select_score_kernel(const unsigned* __restrict__ hist, const int* __restrict__ quota2, int* __restrict__ thr) {
    const int lv = blockIdx.x;
    if (threadIdx.x != 0) return;
    const unsigned* hs = hist + lv * 256;
    const int n = quota2[lv];
    unsigned total = 0;
    for (int s = 255; s >= 1; --s) total += hs[s];
    int t = 1;
    if (n <= 0) t = 256;                                    // nothing survives
    else if (total > static_cast<unsigned>(n)) {
        unsigned acc = 0;
        for (int s = 255; s >= 1; --s) { acc += hs[s]; if (acc >= static_cast<unsigned>(n)) { t = s; break; } }
    }
    thr[lv] = t;
}

__device__ __forceinline__ unsigned float_key(float f) {          // monotone: larger float -> larger key
    const unsigned b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// HarrisResponses(blockSize 7, k 0.04) of the pixels that survived the score selection; others get key 0 in `hkey`
__global__ void __launch_bounds__(256) orb_harris_kernel(const uint8_t* __restrict__ img, const uint8_t* __restrict__ cand, int w, int h,
                                                         const int* __restrict__ thr, int lv, float* __restrict__ resp,
                                                         unsigned* __restrict__ hkey) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= w || y >= h) return;
    const size_t i = static_cast<size_t>(y) * w + x;
    unsigned key = 0;
    float r = 0.f;
    const int s = cand[i];
    if (s > 0 && s >= thr[lv]) {
        int a = 0, b = 0, c = 0;
        for (int dy = -3; dy <= 3; ++dy) {
            const uint8_t* p = img + static_cast<size_t>(y + dy) * w + x;
            for (int dx = -3; dx <= 3; ++dx) {
                const uint8_t* q = p + dx;
                const int ix = (q[1] - q[-1]) * 2 + (q[-w + 1] - q[-w - 1]) + (q[w + 1] - q[w - 1]);
                const int iy = (q[w] - q[-w]) * 2 + (q[w - 1] - q[-w - 1]) + (q[w + 1] - q[-w + 1]);
                a += ix * ix; b += iy * iy; c += ix * iy;
            }
        }
        const float scale = 1.f / (4.f * 7.f * 255.f);
        const float s4 = scale * scale * scale * scale;
        const float fa = static_cast<float>(a), fb = static_cast<float>(b), fc = static_cast<float>(c);
        r = (fa * fb - fc * fc - 0.04f * (fa + fb) * (fa + fb)) * s4;
        key = float_key(r);
        if (key == 0) key = 1;                              // key 0 is reserved for "not a candidate" (only -NaN maps there)
    }
    resp[i] = r;
    hkey[i] = key;
}

// retainBest(quota) by Harris response: the n-th largest key of the level by a 4-pass radix select (one CTA per level);
// keep_key[lv] = that key (everything >= it stays), 1 = keep all candidates
__global__ void __launch_bounds__(1024) orb_select_harris_kernel(const unsigned* __restrict__ hkey, OrbLevels L,
                                                                 const int* __restrict__ quota, unsigned* __restrict__ keep_key) {
    __shared__ unsigned hist[256];
    __shared__ unsigned s_prefix, s_remaining, s_total;
    const int lv = blockIdx.x;
    const int64_t n_px = static_cast<int64_t>(L.w[lv]) * L.h[lv];
    const unsigned* keys = hkey + L.off[lv];
    const int n = quota[lv];
    if (threadIdx.x == 0) { s_prefix = 0; s_remaining = static_cast<unsigned>(max(n, 0)); s_total = 0; }
    __syncthreads();
    // candidates of the level
    unsigned cnt = 0;
    for (int64_t i = threadIdx.x; i < n_px; i += blockDim.x) cnt += keys[i] != 0;
    atomicAdd(&s_total, cnt);
    __syncthreads();
    if (n <= 0) { if (threadIdx.x == 0) keep_key[lv] = 0xFFFFFFFFu; return; }
    if (s_total <= static_cast<unsigned>(n)) { if (threadIdx.x == 0) keep_key[lv] = 1u; return; }
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        for (int b = threadIdx.x; b < 256; b += blockDim.x) hist[b] = 0;
        __syncthreads();
        const unsigned prefix = s_prefix;
        const unsigned mask = pass == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
        for (int64_t i = threadIdx.x; i < n_px; i += blockDim.x) {
            const unsigned k = keys[i];
            if (k != 0 && (k & mask) == prefix) atomicAdd(&hist[(k >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned rem = s_remaining, acc = 0;
            int b = 255;
            for (; b >= 0; --b) { if (acc + hist[b] >= rem) break; acc += hist[b]; }
            s_prefix = prefix | (static_cast<unsigned>(b) << shift);
            s_remaining = rem - acc;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) keep_key[lv] = s_prefix;
}

// ordered compaction, pass 1: kept pixels per image row (one warp per row of any level)
__global__ void __launch_bounds__(256) orb_row_count_kernel(const unsigned* __restrict__ hkey, OrbLevels L, const unsigned* __restrict__ keep_key,
                                                            int total_rows, int* __restrict__ row_count) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= total_rows) return;
    int lv = 0;
    while (lv + 1 < kOrbLevels && row >= L.row_off[lv + 1]) ++lv;
    const int y = row - static_cast<int>(L.row_off[lv]), w = L.w[lv];
    const unsigned* k = hkey + L.off[lv] + static_cast<int64_t>(y) * w;
    const unsigned kk = keep_key[lv];
    int n = 0;
    for (int x = lane; x < w; x += 32) n += k[x] != 0 && k[x] >= kk;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    if (lane == 0) row_count[row] = n;
}

// pass 2: exclusive scan over all rows (single CTA), total -> counts[0]
__global__ void __launch_bounds__(1024) orb_row_scan_kernel(const int* __restrict__ row_count, int total_rows, int* __restrict__ row_start,
                                                            int* __restrict__ counts) {
    __shared__ int warp_sum[32];
    __shared__ int carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < total_rows; base += 1024) {
        const int i = base + tid;
        const int v = i < total_rows ? row_count[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) warp_sum[warp] = incl;
        __syncthreads();
        int before = 0, total = 0;
        for (int wv = 0; wv < 32; ++wv) { const int s = warp_sum[wv]; if (wv < warp) before += s; total += s; }
        if (i < total_rows) row_start[i] = carry + before + incl - v;
        __syncthreads();
        if (tid == 0) carry += total;
        __syncthreads();
    }
    if (tid == 0) counts[0] = carry;
}

// pass 3: the keypoints in (level, y, x) order; level coordinates are kept in `lxy` for the descriptor pass
__global__ void __launch_bounds__(256) orb_emit_kernel(const unsigned* __restrict__ hkey, const float* __restrict__ resp, OrbLevels L,
                                                       const unsigned* __restrict__ keep_key, int total_rows, const int* __restrict__ row_start,
                                                       int capacity, sift::Keypoint* __restrict__ kps, int* __restrict__ lxy) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= total_rows) return;
    int lv = 0;
    while (lv + 1 < kOrbLevels && row >= L.row_off[lv + 1]) ++lv;
    const int y = row - static_cast<int>(L.row_off[lv]), w = L.w[lv];
    const int64_t base = L.off[lv] + static_cast<int64_t>(y) * w;
    const unsigned kk = keep_key[lv];
    int pos = row_start[row];
    for (int x0 = 0; x0 < w; x0 += 32) {
        const int x = x0 + lane;
        const bool keep = x < w && hkey[base + x] != 0 && hkey[base + x] >= kk;
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const int p = pos + __popc(bal & ((1u << lane) - 1u));
            if (p < capacity) {
                sift::Keypoint k;
                k.x = static_cast<float>(x) * L.scale[lv];
                k.y = static_cast<float>(y) * L.scale[lv];
                k.size = static_cast<float>(kOrbPatch) * L.scale[lv];
                k.angle = -1.f;
                k.response = resp[base + x];
                k.octave = lv;
                kps[p] = k;
                lxy[2 * p] = x; lxy[2 * p + 1] = y;
            }
        }
        pos += __popc(bal);
    }
}

// GaussianBlur(7 x 7, 2, 2, BORDER_REFLECT_101), generic separable float filter: row pass s = k0 x0 + k1 x1 + ... (tap order)
__global__ void __launch_bounds__(256) orb_blur_rows_kernel(const uint8_t* __restrict__ img, int w, int h, float* __restrict__ tmp) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= w || y >= h) return;
    const uint8_t* row = img + static_cast<size_t>(y) * w;
    float s = c_gauss[0] * static_cast<float>(row[reflect101_orb(x - 3, w)]);
#pragma unroll
    for (int i = 1; i < 7; ++i) s = s + c_gauss[i] * static_cast<float>(row[reflect101_orb(x - 3 + i, w)]);
    tmp[static_cast<size_t>(y) * w + x] = s;
}
// column pass: s = k3 r3 + k4 (r4 + r2) + k5 (r5 + r1) + k6 (r6 + r0), cvRound, saturate
__global__ void __launch_bounds__(256) orb_blur_cols_kernel(const float* __restrict__ tmp, int w, int h, uint8_t* __restrict__ out) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= w || y >= h) return;
    auto at = [&](int yy) { return tmp[static_cast<size_t>(reflect101_orb(yy, h)) * w + x]; };
    float s = c_gauss[3] * at(y);
#pragma unroll
    for (int j = 1; j <= 3; ++j) s = s + c_gauss[3 + j] * (at(y + j) + at(y - j));
    const int v = __float2int_rn(s);
    out[static_cast<size_t>(y) * w + x] = static_cast<uint8_t>(min(max(v, 0), 255));
}

constexpr int kOrbDescWarps = 8;
// one warp per keypoint: IC_Angle on the level image, then the 32 descriptor bytes from the blurred level
__global__ void __launch_bounds__(kOrbDescWarps * 32) orb_describe_kernel(const uint8_t* __restrict__ img, const uint8_t* __restrict__ blur,
                                                                          OrbLevels L, sift::Keypoint* __restrict__ kps,
                                                                          const int* __restrict__ lxy, const int* __restrict__ n_kps,
                                                                          int capacity, uint8_t* __restrict__ desc) {
    const int lane = threadIdx.x & 31;
    const int n = min(*n_kps, capacity);
    for (int i = blockIdx.x * kOrbDescWarps + (threadIdx.x >> 5); i < n; i += gridDim.x * kOrbDescWarps) {
        const int lv = kps[i].octave, x = lxy[2 * i], y = lxy[2 * i + 1], w = L.w[lv];
        const uint8_t* c = img + L.off[lv] + static_cast<int64_t>(y) * w + x;
        int m10 = 0, m01 = 0;
        for (int k = lane; k < kOrbPatchPixels; k += 32) {
            const int u = c_patch_u[k], v = c_patch_v[k];
            const int val = c[v * w + u];
            m10 += u * val;
            m01 += v * val;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { m10 += __shfl_xor_sync(0xffffffffu, m10, o); m01 += __shfl_xor_sync(0xffffffffu, m01, o); }
        const float angle = sift::fast_atan2_deg(static_cast<float>(m01), static_cast<float>(m10));
        if (lane == 0) kps[i].angle = angle;
        // computeOrbDescriptors: a = cos, b = sin of the angle in radians (double functions, rounded to float)
        const float rad = angle * static_cast<float>(3.14159265358979323846 / 180.0);
        const float a = static_cast<float>(cos(static_cast<double>(rad))), b = static_cast<float>(sin(static_cast<double>(rad)));
        const uint8_t* cb = blur + L.off[lv] + static_cast<int64_t>(y) * w + x;
        unsigned byte = 0;
#pragma unroll
        for (int bit = 0; bit < 8; ++bit) {
            const int* p = c_pattern[lane * 8 + bit];
            const float x0 = static_cast<float>(p[0]), y0 = static_cast<float>(p[1]), x1 = static_cast<float>(p[2]), y1 = static_cast<float>(p[3]);
            const int ix0 = __float2int_rn(x0 * a - y0 * b), iy0 = __float2int_rn(x0 * b + y0 * a);
            const int ix1 = __float2int_rn(x1 * a - y1 * b), iy1 = __float2int_rn(x1 * b + y1 * a);
            const int t0 = cb[iy0 * w + ix0], t1 = cb[iy1 * w + ix1];
            byte |= static_cast<unsigned>(t0 < t1) << bit;
        }
        desc[static_cast<size_t>(i) * 32 + lane] = static_cast<uint8_t>(byte);
    }
}

template <class T>
cudaError_t grow_buf(T*& p, size_t& cap, size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&p), n * sizeof(T));
    if (e == cudaSuccess) cap = n;
    return e;
}

}  // namespace

struct OrbWorkspace {
    uint8_t* d_img = nullptr; size_t img_cap = 0;          // all levels, flat
    uint8_t* d_blur = nullptr; size_t blur_cap = 0;
    uint8_t* d_score = nullptr; size_t score_cap = 0;
    uint8_t* d_cand = nullptr; size_t cand_cap = 0;
    float* d_resp = nullptr; size_t resp_cap = 0;
    unsigned* d_hkey = nullptr; size_t hkey_cap = 0;
    float* d_tmp = nullptr; size_t tmp_cap = 0;
    int* d_coeff = nullptr; size_t coeff_cap = 0;          // ox | cx | oy | cy of the level being resized
    int* d_rows = nullptr; size_t rows_cap = 0;            // row_count | row_start
    unsigned* d_hist = nullptr;                            // 8 x 256
    int* d_small = nullptr;                                // quota2[8] | quota[8] | thr[8] | keep_key[8] | counts[4]
    int* h_counts = nullptr;
    sift::Keypoint* d_kp = nullptr; size_t kp_cap = 0;
    int* d_lxy = nullptr; size_t lxy_cap = 0;
    uint8_t* d_desc = nullptr; size_t desc_cap = 0;
    bool tables_set = false;
    OrbLevels last{};                                      // geometry of the last extraction (test hook)
    cudaEvent_t ev[2] = {nullptr, nullptr};
    float total_ms = 0.f;
};

OrbWorkspace* orb_workspace_create() { return new OrbWorkspace(); }
void orb_workspace_destroy(OrbWorkspace* w) {
    if (!w) return;
    cudaFree(w->d_img); cudaFree(w->d_blur); cudaFree(w->d_score); cudaFree(w->d_cand); cudaFree(w->d_resp); cudaFree(w->d_hkey);
    cudaFree(w->d_tmp); cudaFree(w->d_coeff); cudaFree(w->d_rows); cudaFree(w->d_hist); cudaFree(w->d_small); cudaFree(w->d_kp);
    cudaFree(w->d_lxy); cudaFree(w->d_desc);
    if (w->h_counts) cudaFreeHost(w->h_counts);
    for (cudaEvent_t e : w->ev) if (e) cudaEventDestroy(e);
    delete w;
}
// test hook: one per-pixel map of the last extraction (0 image, 1 blurred image, 2 FAST score, 3 candidates after the
// non-maximum suppression: all u8; 4 Harris response: float) and the level geometry
int orb_level_map(const OrbWorkspace* w, int what, int level, const void** ptr, int* width, int* height) {
    if (level < 0 || level >= kOrbLevels || w->last.w[level] <= 0) return -1;
    const int64_t off = w->last.off[level];
    *width = w->last.w[level]; *height = w->last.h[level];
    switch (what) {
        case 0: *ptr = w->d_img + off; return 1;
        case 1: *ptr = w->d_blur + off; return 1;
        case 2: *ptr = w->d_score + off; return 1;
        case 3: *ptr = w->d_cand + off; return 1;
        case 4: *ptr = w->d_resp + off; return 4;
        default: return -1;
    }
}
const void* orb_keypoints_device_raw(const OrbWorkspace* w) { return w->d_kp; }
const uint8_t* orb_descriptors_device(const OrbWorkspace* w) { return w->d_desc; }
float orb_last_ms(const OrbWorkspace* w) { return w->total_ms; }

#define ORB_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return _e; } while (0)

cudaError_t orb_extract(OrbWorkspace* ws, const uint8_t* gray, int rows, int cols, size_t step, int n_features, int max_keypoints,
                        cudaStream_t s, int* n_keypoints, int* n_launches, std::string* err) {
    *n_keypoints = 0;
    int launches = 0;
    // ---- geometry (ORB_Impl::detectAndCompute: layerScale, layerInfo)
    const double scale_factor = static_cast<double>(1.2f);          // ORB::create(.., float scaleFactor = 1.2f, ..) held in a double
    OrbLevels L{};
    int64_t total_px = 0, total_rows = 0;
    int max_w = 0;
    for (int lv = 0; lv < kOrbLevels; ++lv) {
        L.scale[lv] = static_cast<float>(std::pow(scale_factor, static_cast<double>(lv)));
        L.w[lv] = static_cast<int>(std::lrint(static_cast<float>(cols) / L.scale[lv]));
        L.h[lv] = static_cast<int>(std::lrint(static_cast<float>(rows) / L.scale[lv]));
        if (L.w[lv] < 1 || L.h[lv] < 1) { if (err) *err = "ORB: image too small for 8 pyramid levels"; return cudaErrorInvalidValue; }
        L.off[lv] = total_px; L.row_off[lv] = total_rows;
        total_px += static_cast<int64_t>(L.w[lv]) * L.h[lv];
        total_rows += L.h[lv];
        max_w = std::max(max_w, L.w[lv]);
    }
    ws->last = L;
    // nfeaturesPerLevel (computeKeyPoints)
    int quota[kOrbLevels], small_host[40] = {0};
    {
        const float factor = static_cast<float>(1.0 / scale_factor);
        float nd = static_cast<float>(n_features) * (1.f - factor) / (1.f - static_cast<float>(std::pow(static_cast<double>(factor), static_cast<double>(kOrbLevels))));
        int sum = 0;
        for (int lv = 0; lv < kOrbLevels - 1; ++lv) { quota[lv] = static_cast<int>(std::lrint(nd)); sum += quota[lv]; nd *= factor; }
        quota[kOrbLevels - 1] = std::max(n_features - sum, 0);
        for (int lv = 0; lv < kOrbLevels; ++lv) { small_host[lv] = 2 * quota[lv]; small_host[8 + lv] = quota[lv]; }
    }
    ORB_TRY(grow_buf(ws->d_img, ws->img_cap, static_cast<size_t>(total_px)));
    ORB_TRY(grow_buf(ws->d_blur, ws->blur_cap, static_cast<size_t>(total_px)));
    ORB_TRY(grow_buf(ws->d_score, ws->score_cap, static_cast<size_t>(total_px)));
    ORB_TRY(grow_buf(ws->d_cand, ws->cand_cap, static_cast<size_t>(total_px)));
    ORB_TRY(grow_buf(ws->d_resp, ws->resp_cap, static_cast<size_t>(total_px)));
    ORB_TRY(grow_buf(ws->d_hkey, ws->hkey_cap, static_cast<size_t>(total_px)));
    ORB_TRY(grow_buf(ws->d_tmp, ws->tmp_cap, static_cast<size_t>(L.w[0]) * L.h[0]));
    ORB_TRY(grow_buf(ws->d_coeff, ws->coeff_cap, static_cast<size_t>(2 * (L.w[0] + L.h[0]) + 16)));
    ORB_TRY(grow_buf(ws->d_rows, ws->rows_cap, static_cast<size_t>(2 * total_rows + 16)));
    ORB_TRY(grow_buf(ws->d_kp, ws->kp_cap, static_cast<size_t>(max_keypoints)));
    ORB_TRY(grow_buf(ws->d_lxy, ws->lxy_cap, static_cast<size_t>(max_keypoints) * 2));
    ORB_TRY(grow_buf(ws->d_desc, ws->desc_cap, static_cast<size_t>(max_keypoints) * 32));
    if (!ws->d_hist) ORB_TRY(cudaMalloc(reinterpret_cast<void**>(&ws->d_hist), kOrbLevels * 256 * sizeof(unsigned)));
    if (!ws->d_small) ORB_TRY(cudaMalloc(reinterpret_cast<void**>(&ws->d_small), 64 * sizeof(int)));
    if (!ws->h_counts) ORB_TRY(cudaMallocHost(reinterpret_cast<void**>(&ws->h_counts), 16));
    for (cudaEvent_t& e : ws->ev) if (!e) ORB_TRY(cudaEventCreate(&e));
    if (!ws->tables_set) {
        // circular patch (umax table of computeKeyPoints) and the Gaussian kernel (getGaussianKernel(7, 2, CV_32F))
        int umax[kOrbHalfPatch + 2] = {0};
        const int vmax = static_cast<int>(std::floor(kOrbHalfPatch * std::sqrt(2.0) / 2 + 1)), vmin = static_cast<int>(std::ceil(kOrbHalfPatch * std::sqrt(2.0) / 2));
        for (int v = 0; v <= vmax; ++v) umax[v] = static_cast<int>(std::lrint(std::sqrt(static_cast<double>(kOrbHalfPatch * kOrbHalfPatch - v * v))));
        for (int v = kOrbHalfPatch, v0 = 0; v >= vmin; --v) { while (umax[v0] == umax[v0 + 1]) ++v0; umax[v] = v0; ++v0; }
        std::vector<int8_t> pu, pv;
        for (int v = -kOrbHalfPatch; v <= kOrbHalfPatch; ++v)
            for (int u = -umax[std::abs(v)]; u <= umax[std::abs(v)]; ++u) { pu.push_back(static_cast<int8_t>(u)); pv.push_back(static_cast<int8_t>(v)); }
        if (static_cast<int>(pu.size()) != kOrbPatchPixels) { if (err) *err = "ORB: patch table size"; return cudaErrorInvalidValue; }
        ORB_TRY(cudaMemcpyToSymbol(c_patch_u, pu.data(), kOrbPatchPixels));
        ORB_TRY(cudaMemcpyToSymbol(c_patch_v, pv.data(), kOrbPatchPixels));
        double t[7], sum = 0;
        for (int i = 0; i < 7; ++i) { const double x = i - 3; t[i] = std::exp(-(x * x) / (2 * 2.0 * 2.0)); sum += t[i]; }
        float g[7];
        for (int i = 0; i < 7; ++i) g[i] = static_cast<float>(t[i] / sum);
        ORB_TRY(cudaMemcpyToSymbol(c_gauss, g, sizeof g));
        ws->tables_set = true;
    }
    int* d_quota2 = ws->d_small; int* d_quota = ws->d_small + 8; int* d_thr = ws->d_small + 16;
    unsigned* d_keep = reinterpret_cast<unsigned*>(ws->d_small + 24); int* d_counts = ws->d_small + 32;
    ORB_TRY(cudaEventRecord(ws->ev[0], s));
    ORB_TRY(cudaMemcpyAsync(ws->d_small, small_host, sizeof small_host, cudaMemcpyHostToDevice, s));
    ORB_TRY(cudaMemsetAsync(ws->d_hist, 0, kOrbLevels * 256 * sizeof(unsigned), s));
    ORB_TRY(cudaMemcpy2DAsync(ws->d_img, cols, gray, step, cols, rows, cudaMemcpyHostToDevice, s));
    const dim3 blk(32, 8);
    auto grid2 = [](int w, int h) { return dim3((w + 31) / 32, (h + 7) / 8); };
    for (int lv = 0; lv < kOrbLevels; ++lv) {
        const int w = L.w[lv], h = L.h[lv];
        uint8_t* img = ws->d_img + L.off[lv];
        if (lv > 0) {
            const int sw = L.w[lv - 1], sh = L.h[lv - 1];
            int* ox = ws->d_coeff; int* cx = ox + w; int* oy = cx + w; int* cy = oy + h;
            orb_coeff_kernel<<<(w + 255) / 256, 256, 0, s>>>(w, sw, ox, cx);
            orb_coeff_kernel<<<(h + 255) / 256, 256, 0, s>>>(h, sh, oy, cy);
            orb_resize_kernel<<<grid2(w, h), blk, 0, s>>>(ws->d_img + L.off[lv - 1], sw, sh, img, w, h, ox, cx, oy, cy);
            launches += 3;
        }
        orb_fast_kernel<<<grid2(w, h), blk, 0, s>>>(img, w, h, kOrbFastThreshold, ws->d_score + L.off[lv]);
        orb_nms_kernel<<<grid2(w, h), blk, 0, s>>>(ws->d_score + L.off[lv], w, h, ws->d_cand + L.off[lv], ws->d_hist + lv * 256);
        orb_blur_rows_kernel<<<grid2(w, h), blk, 0, s>>>(img, w, h, ws->d_tmp);
        orb_blur_cols_kernel<<<grid2(w, h), blk, 0, s>>>(ws->d_tmp, w, h, ws->d_blur + L.off[lv]);
        launches += 4;
    }
    orb_select_score_kernel<<<kOrbLevels, 32, 0, s>>>(ws->d_hist, d_quota2, d_thr);
    for (int lv = 0; lv < kOrbLevels; ++lv)
        orb_harris_kernel<<<grid2(L.w[lv], L.h[lv]), blk, 0, s>>>(ws->d_img + L.off[lv], ws->d_cand + L.off[lv], L.w[lv], L.h[lv], d_thr, lv,
                                                                  ws->d_resp + L.off[lv], ws->d_hkey + L.off[lv]);
    orb_select_harris_kernel<<<kOrbLevels, 1024, 0, s>>>(ws->d_hkey, L, d_quota, d_keep);
    int* row_count = ws->d_rows; int* row_start = ws->d_rows + total_rows;
    const int tr = static_cast<int>(total_rows);
    orb_row_count_kernel<<<(tr + 7) / 8, 256, 0, s>>>(ws->d_hkey, L, d_keep, tr, row_count);
    orb_row_scan_kernel<<<1, 1024, 0, s>>>(row_count, tr, row_start, d_counts);
    orb_emit_kernel<<<(tr + 7) / 8, 256, 0, s>>>(ws->d_hkey, ws->d_resp, L, d_keep, tr, row_start, max_keypoints, ws->d_kp, ws->d_lxy);
    orb_describe_kernel<<<148 * 4, kOrbDescWarps * 32, 0, s>>>(ws->d_img, ws->d_blur, L, ws->d_kp, ws->d_lxy, d_counts, max_keypoints, ws->d_desc);
    launches += 6 + kOrbLevels;
    ORB_TRY(cudaGetLastError());
    ORB_TRY(cudaEventRecord(ws->ev[1], s));
    ORB_TRY(cudaMemcpyAsync(ws->h_counts, d_counts, 4, cudaMemcpyDeviceToHost, s));
    ORB_TRY(cudaStreamSynchronize(s));
    ORB_TRY(cudaEventElapsedTime(&ws->total_ms, ws->ev[0], ws->ev[1]));
    if (ws->h_counts[0] > max_keypoints) {
        if (err) *err = "feature extraction: more ORB keypoints (" + std::to_string(ws->h_counts[0]) + ") than the capacity: raise max_keypoints";
        return cudaErrorMemoryAllocation;
    }
    *n_keypoints = ws->h_counts[0];
    if (n_launches) *n_launches = launches;
    return cudaSuccess;
}

}  // namespace sfm
