// sift.cu — feature extraction on the device: SfM::extractFeatures (SfM.cpp:577-597) with the
// detector of PhotogrammetrieCli.cpp:342-357, cv::SIFT::create(featureLimit, 3, 0.09).  One grey image in, sorted + deduplicated keypoints and their
// 128-byte descriptors out, both device-resident so that the descriptor bank of the matching stage is filled without a
// host round trip (sfm_bank_from_features).
//
// Kernels (fp32, compiled with --fmad=false so that the pyramid is reproducible to the bit):
//   upsample2x_kernel        u8 grey -> float, 2x INTER_LINEAR (createInitialImage)
//   gauss_blur_kernel        separable Gaussian, BORDER_REFLECT_101, both passes fused through shared memory:
//                            every pyramid level is read once and written once
//   downsample_kernel        octave base = every second pixel of level [nOctaveLayers] of the previous octave
//   extrema_kernel           26-neighbour extrema of the DoG (differences taken on the fly) -> candidate list
//   refine_orient_kernel     adjustLocalExtrema + calcOrientationHist, one warp per candidate -> keypoint list
//   bucket_* / scatter_kernel / dedupe_kernel      KeyPointsFilter::removeDuplicatedSorted + firstOctave correction
//   retain_best_kernel       KeyPointsFilter::retainBest (nfeatures = the reference's feature-limit)
//   descriptor_kernel        calcSIFTDescriptor, one warp per keypoint -> u8 rows
// The per-keypoint arithmetic is sift_core.cuh (shared with the host test harness).  The warps walk the samples of a
// keypoint in OpenCV's order and add the contributions to every histogram bin in that order, which keeps the float sums —
// and with them every borderline decision — identical to the serial algorithm.
#include <cmath>
#include <cstdio>
#include <string>
#include <vector>

#include "kernels.h"
#include "sift_core.cuh"

namespace sfm {

using namespace sift;
static_assert(kSiftMaxOctaves == kMaxOctaves && sizeof(Keypoint) == 24, "sift_core.cuh / kernels.h disagree");

namespace {

struct BlurWeights {
    int radius;
    float w[2 * kMaxBlurRadius + 1];
};

__device__ __forceinline__ int reflect101(int p, int n) {      // cv::borderInterpolate(p, n, BORDER_REFLECT_101)
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * (n - 1) - p;
    return p;
}

__global__ void __launch_bounds__(256) upsample2x_kernel(const uint8_t* __restrict__ src, int w, int h, size_t step,
                                                         float* __restrict__ dst) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= 2 * w || y >= 2 * h) return;
    // sample position (d + 0.5) / 2 - 0.5: even d -> (k - 1, 0.75), odd d -> (k, 0.25); clamped taps have weight 0
    int x0 = (x & 1) ? (x >> 1) : (x >> 1) - 1;
    float fx = (x & 1) ? 0.25f : 0.75f;
    if (x0 < 0) { x0 = 0; fx = 0.f; }
    if (x0 >= w - 1) { x0 = w - 1; fx = 0.f; }
    const int x1 = min(x0 + 1, w - 1);
    int y0 = (y & 1) ? (y >> 1) : (y >> 1) - 1;
    float fy = (y & 1) ? 0.25f : 0.75f;
    if (y0 < 0) { y0 = 0; fy = 0.f; }
    if (y0 >= h - 1) { y0 = h - 1; fy = 0.f; }
    const int y1 = min(y0 + 1, h - 1);
    const uint8_t* r0 = src + static_cast<size_t>(y0) * step;
    const uint8_t* r1 = src + static_cast<size_t>(y1) * step;
    const float a = static_cast<float>(r0[x0]) * (1.f - fx) + static_cast<float>(r0[x1]) * fx;
    const float b = static_cast<float>(r1[x0]) * (1.f - fx) + static_cast<float>(r1[x1]) * fx;
    dst[static_cast<size_t>(y) * (2 * w) + x] = a * (1.f - fy) + b * fy;
}

// createInitialImage without the 2x upsampling (cv::SIFT::compute when no keypoint lies in octave -1): u8 grey -> float
__global__ void __launch_bounds__(256) gray_to_float_kernel(const uint8_t* __restrict__ src, int w, int h, size_t step,
                                                            float* __restrict__ dst) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x < w && y < h) dst[static_cast<size_t>(y) * w + x] = static_cast<float>(src[static_cast<size_t>(y) * step + x]);
}

// octave range of the final keypoints: range[0] = min octave + 128 (starts at 0x7fffffff), range[1] = max octave + 128
__global__ void __launch_bounds__(256) octave_range_kernel(const Keypoint* __restrict__ kps, const int* __restrict__ n_kps, int capacity,
                                                           int* __restrict__ range) {
    const int n = min(*n_kps, capacity);
    int mn = 0x7fffffff, mx = 0;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        int octave, layer;
        float scale;
        unpack_octave(kps[i].octave, octave, layer, scale);
        mn = min(mn, octave + 128);
        mx = max(mx, octave + 128);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
    if ((threadIdx.x & 31) == 0 && mx > 0) { atomicMin(range, mn); atomicMax(range + 1, mx); }
}

constexpr int kBlurTW = 64, kBlurTH = 32;
constexpr int kBlurMidStride = kBlurTW + 1;      // odd strides: a warp walking down a column of the tile hits 32 banks
__host__ __device__ inline int blur_in_stride(int radius) { return (kBlurTW + 2 * radius) | 1; }

// eight consecutive outputs along the filter direction per thread: a sliding window of eight inputs in registers, one
// shared-memory load per tap.  Every output is still  ((0 + w0 x0) + w1 x1) + ...  in tap order, multiply and add rounded
// separately (--fmad=false): the same floats as the one-output-at-a-time form.
__device__ __forceinline__ void blur_eight_outputs(const float* __restrict__ p, int stride, int taps, const BlurWeights& k,
                                                   float acc[8]) {
    float win[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) { acc[o] = 0.f; win[o] = p[o * stride]; }
    for (int t8 = 0; t8 < taps; t8 += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (t8 + u < taps) {
                const float wt = k.w[t8 + u];
#pragma unroll
                for (int o = 0; o < 8; ++o) acc[o] += wt * win[(o + u) & 7];
                win[u] = p[(t8 + u + 8) * stride];
            }
        }
    }
}

// taps known at compile time: the tap loop is fully unrolled and the weights become constant-bank operands
// (tools/blur_ab.cu v2: bit-identical to the generic form, 1.5-1.6x faster; profiles/r2_blur_ab.txt)
template <int R>
__device__ __forceinline__ void blur_eight_outputs_fixed(const float* __restrict__ p, int stride, const BlurWeights& k, float acc[8]) {
    constexpr int taps = 2 * R + 1;
    float win[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) { acc[o] = 0.f; win[o] = p[o * stride]; }
#pragma unroll
    for (int t = 0; t < taps; ++t) {
        const float wt = k.w[t];
#pragma unroll
        for (int o = 0; o < 8; ++o) acc[o] += wt * win[(o + t) & 7];
        win[t & 7] = p[(t + 8) * stride];
    }
}

// source tile + halo into shared memory: 2-D thread mapping (no integer division per element); interior tiles skip the
// reflect-101 index arithmetic (only the frame of the image needs it)
__device__ __forceinline__ void blur_load_tile(const float* __restrict__ src, int w, int h, int x0, int y0, int R, float* in) {
    const int IW = kBlurTW + 2 * R, IH = kBlurTH + 2 * R, IS = blur_in_stride(R);
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const bool interior = x0 - R >= 0 && x0 + kBlurTW + R <= w && y0 - R >= 0 && y0 + kBlurTH + R <= h;
    if (interior) {
        const float* base = src + static_cast<size_t>(y0 - R) * w + (x0 - R);
        for (int iy = ty; iy < IH; iy += 8) {
            const float* row = base + static_cast<size_t>(iy) * w;
            float* d = in + iy * IS;
            for (int ix = tx; ix < IW; ix += 32) d[ix] = row[ix];
        }
    } else {
        for (int iy = ty; iy < IH; iy += 8) {
            const float* row = src + static_cast<size_t>(reflect101(y0 - R + iy, h)) * w;
            float* d = in + iy * IS;
            for (int ix = tx; ix < IW; ix += 32) d[ix] = row[reflect101(x0 - R + ix, w)];
        }
    }
}

// kR > 0: radius fixed at compile time (the radii cv::SIFT's blur schedule produces); kR = 0: any radius <= 32
template <int kR>
__global__ void __launch_bounds__(256) gauss_blur_kernel(const float* __restrict__ src, float* __restrict__ dst, int w, int h,
                                                         const BlurWeights k) {
    extern __shared__ float sm[];
    const int R = kR > 0 ? kR : k.radius;
    const int IH = kBlurTH + 2 * R;
    const int IS = blur_in_stride(R);
    float* in = sm;                  // IH x IW source tile with halo (row stride IS), 8 floats of slack behind it
    float* mid = sm + IH * IS + 8;   // IH x kBlurTW, filtered along x (row stride kBlurMidStride), 8 rows of slack behind it
    const int x0 = blockIdx.x * kBlurTW, y0 = blockIdx.y * kBlurTH;
    const int tid = threadIdx.x;
    blur_load_tile(src, w, h, x0, y0, R, in);
    __syncthreads();
    const int taps = 2 * R + 1;
    // along x: unit = (eight outputs, one tile row); consecutive lanes take consecutive rows
    for (int unit = tid; unit < IH * (kBlurTW / 8); unit += 256) {
        const int iy = unit % IH, ix0 = (unit / IH) * 8;
        const float* p = in + iy * IS + ix0;
        float acc[8];
        if (kR > 0) blur_eight_outputs_fixed<kR>(p, 1, k, acc);
        else blur_eight_outputs(p, 1, taps, k, acc);
        float* q = mid + iy * kBlurMidStride + ix0;
#pragma unroll
        for (int o = 0; o < 8; ++o) q[o] = acc[o];
    }
    __syncthreads();
    // along y: unit = (eight output rows, one column); consecutive lanes take consecutive columns
    {
        const int ox = tid % kBlurTW, oy0 = (tid / kBlurTW) * 8;
        const float* p = mid + oy0 * kBlurMidStride + ox;
        float acc[8];
        if (kR > 0) blur_eight_outputs_fixed<kR>(p, kBlurMidStride, k, acc);
        else blur_eight_outputs(p, kBlurMidStride, taps, k, acc);
        if (x0 + ox < w) {
#pragma unroll
            for (int o = 0; o < 8; ++o)
                if (y0 + oy0 + o < h) dst[static_cast<size_t>(y0 + oy0 + o) * w + x0 + ox] = acc[o];
        }
    }
}

typedef void (*BlurKernel)(const float*, float*, int, int, const BlurWeights);
// the radii of cv::SIFT(nOctaveLayers = 3, sigma = 1.6): 5 (first blur of the doubled image), 6, 5, 6, 8, 10, 13 (levels 1..5)
static BlurKernel blur_kernel_for(int radius) {
    switch (radius) {
        case 3: return gauss_blur_kernel<3>;
        case 4: return gauss_blur_kernel<4>;
        case 5: return gauss_blur_kernel<5>;
        case 6: return gauss_blur_kernel<6>;
        case 7: return gauss_blur_kernel<7>;
        case 8: return gauss_blur_kernel<8>;
        case 9: return gauss_blur_kernel<9>;
        case 10: return gauss_blur_kernel<10>;
        case 11: return gauss_blur_kernel<11>;
        case 13: return gauss_blur_kernel<13>;
        default: return gauss_blur_kernel<0>;
    }
}

__global__ void __launch_bounds__(256) downsample_kernel(const float* __restrict__ src, int sw, float* __restrict__ dst, int dw,
                                                         int dh) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x < dw && y < dh) dst[static_cast<size_t>(y) * dw + x] = src[static_cast<size_t>(2 * y) * sw + 2 * x];
}

__global__ void __launch_bounds__(256) extrema_kernel(const PyramidView P, int o, float threshold, Candidate* __restrict__ cand,
                                                      int* __restrict__ count, int capacity) {
    const int c = blockIdx.x * 32 + threadIdx.x + kImgBorder, r = blockIdx.y * 8 + threadIdx.y + kImgBorder;
    const int i = blockIdx.z + 1;
    if (r >= P.h[o] - kImgBorder || c >= P.w[o] - kImgBorder) return;
    if (!is_extremum(P, o, i, r, c, threshold)) return;
    const int k = atomicAdd(count, 1);
    if (k < capacity) cand[k] = Candidate{o, i, r, c};
}

// One warp per candidate (grid-stride).  Lane 0 runs adjustLocalExtrema; the orientation histogram is filled by all 32
// lanes, 32 consecutive samples per step, and the contributions to one bin are added in sample order (lanes that hit
// the same bin take turns by their rank among those lanes): the float sums equal the serial loop's bit for bit.
constexpr int kRefineWarps = 4;
__global__ void __launch_bounds__(kRefineWarps * 32) refine_orient_kernel(const PyramidView P, const Candidate* __restrict__ cand,
                                                                         const int* __restrict__ n_cand, int cand_capacity,
                                                                         float contrast, float edge, float sigma,
                                                                         Keypoint* __restrict__ kps, int* __restrict__ n_kps,
                                                                         int kp_capacity) {
    __shared__ float hist_s[kRefineWarps][kOriBins];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* hist = hist_s[warp];
    const int n = min(*n_cand, cand_capacity);
    const int stride = gridDim.x * kRefineWarps;
    for (int k = blockIdx.x * kRefineWarps + warp; k < n; k += stride) {
        const Candidate cd = cand[k];
        int layer = cd.layer, r = cd.r, c = cd.c;
        Keypoint kp{};
        int ok = 0;
        if (lane == 0) ok = adjust_local_extrema(P, cd.octave, layer, r, c, contrast, edge, sigma, kp) ? 1 : 0;
        ok = __shfl_sync(0xffffffffu, ok, 0);
        if (!ok) continue;
        layer = __shfl_sync(0xffffffffu, layer, 0);
        r = __shfl_sync(0xffffffffu, r, 0);
        c = __shfl_sync(0xffffffffu, c, 0);
        const float size = __shfl_sync(0xffffffffu, kp.size, 0);
        const float scl_octv = size * 0.5f / (1 << cd.octave);
        const int radius = cv_round(4.5f * scl_octv);
        const float osigma = 1.5f * scl_octv;
        const float expf_scale = -1.f / (2.f * osigma * osigma);
        const float* img = P.level(cd.octave, layer);
        const int cols = P.w[cd.octave], rows = P.h[cd.octave];
        for (int b = lane; b < kOriBins; b += 32) hist[b] = 0.f;
        __syncwarp();
        const int total = (2 * radius + 1) * (2 * radius + 1);
        for (int g0 = 0; g0 < total; g0 += 32) {
            const int g = g0 + lane;
            int bin = 0;
            float val = 0.f;
            const bool valid = g < total && orientation_sample(img, cols, rows, r, c, radius, expf_scale, g, bin, val);
            const unsigned peers = __match_any_sync(0xffffffffu, valid ? bin : 64 + lane);
            const int rank = valid ? __popc(peers & ((1u << lane) - 1u)) : -1;
            const int rounds = __reduce_max_sync(0xffffffffu, rank) + 1;
            for (int t = 0; t < rounds; ++t) {
                if (rank == t) hist[bin] += val;
                __syncwarp();
            }
        }
        if (lane == 0) {
            float temphist[kOriBins], angles[kOriBins];
            for (int b = 0; b < kOriBins; ++b) temphist[b] = hist[b];
            const int m = orientation_finish(temphist, angles);
            for (int a = 0; a < m; ++a) {
                kp.angle = angles[a];
                const int idx = atomicAdd(n_kps, 1);
                if (idx < kp_capacity) kps[idx] = kp;
            }
        }
        __syncwarp();
    }
}

// Position of every keypoint in KeyPoint12_LessThan order (ties between identical keypoints: list position).  The primary
// key is x: keypoints are grouped by floor(x) (counting sort), every keypoint below a bucket precedes every keypoint in
// it, and only the few members of one bucket are compared field by field.
__device__ __forceinline__ int x_bucket(const Keypoint& k, int n_buckets) {
    const int b = static_cast<int>(k.x);
    return b < 0 ? 0 : (b >= n_buckets ? n_buckets - 1 : b);
}
__global__ void __launch_bounds__(256) bucket_count_kernel(const Keypoint* __restrict__ kps, const int* __restrict__ n_kps, int capacity,
                                                           int n_buckets, int* __restrict__ count) {
    const int n = min(*n_kps, capacity);
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) atomicAdd(count + x_bucket(kps[i], n_buckets), 1);
}
// exclusive scan of the bucket sizes -> start[0 .. n_buckets]; cursor = copy of start for the fill pass
__global__ void __launch_bounds__(1024) bucket_scan_kernel(const int* __restrict__ count, int n_buckets, int* __restrict__ start,
                                                           int* __restrict__ cursor) {
    __shared__ int warp_sum[32];
    __shared__ int running;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) running = 0;
    __syncthreads();
    for (int base = 0; base < n_buckets; base += 1024) {
        const int i = base + tid;
        const int v = i < n_buckets ? count[i] : 0;
        int incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) warp_sum[warp] = incl;
        __syncthreads();
        int before = 0, total = 0;
        for (int wv = 0; wv < 32; ++wv) {
            const int sv = warp_sum[wv];
            if (wv < warp) before += sv;
            total += sv;
        }
        if (i < n_buckets) {
            const int excl = running + before + incl - v;
            start[i] = excl;
            cursor[i] = excl;
        }
        __syncthreads();
        if (tid == 0) running += total;
        __syncthreads();
    }
    if (tid == 0) start[n_buckets] = running;
}
__global__ void __launch_bounds__(256) bucket_fill_kernel(const Keypoint* __restrict__ kps, const int* __restrict__ n_kps, int capacity,
                                                          int n_buckets, int* __restrict__ cursor, int* __restrict__ grouped) {
    const int n = min(*n_kps, capacity);
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256)
        grouped[atomicAdd(cursor + x_bucket(kps[i], n_buckets), 1)] = i;
}
__global__ void __launch_bounds__(256) bucket_rank_kernel(const Keypoint* __restrict__ kps, const int* __restrict__ n_kps, int capacity,
                                                          int n_buckets, const int* __restrict__ start, const int* __restrict__ grouped,
                                                          int* __restrict__ rank) {
    const int n = min(*n_kps, capacity);
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        const Keypoint me = kps[i];
        const int b = x_bucket(me, n_buckets);
        const int p0 = start[b], p1 = start[b + 1];
        int cnt = 0;
        for (int p = p0; p < p1; ++p) {
            const int j = grouped[p];
            const Keypoint other = kps[j];
            if (keypoint_less(other, me) || (!keypoint_less(me, other) && j < i)) ++cnt;
        }
        rank[i] = p0 + cnt;
    }
}

__global__ void __launch_bounds__(256) scatter_kernel(const Keypoint* __restrict__ kps, const int* __restrict__ n_kps, int capacity,
                                                      const int* __restrict__ rank, Keypoint* __restrict__ sorted) {
    const int n = min(*n_kps, capacity);
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) sorted[rank[i]] = kps[i];
}

// keep the first of every run of equal (x, y, size, angle); coordinates back to the input image (firstOctave = -1)
__global__ void __launch_bounds__(1024) dedupe_kernel(const Keypoint* __restrict__ sorted, const int* __restrict__ n_kps,
                                                      int capacity, Keypoint* __restrict__ out, int* __restrict__ n_out) {
    __shared__ int warp_sum[32];
    __shared__ int running;
    const int n = min(*n_kps, capacity);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) running = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + tid;
        Keypoint kp{};
        bool keep = false;
        if (i < n) {
            kp = sorted[i];
            keep = i == 0 || !keypoint_duplicate(sorted[i - 1], kp);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) warp_sum[warp] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
        for (int wv = 0; wv < 32; ++wv) {
            const int s = warp_sum[wv];
            if (wv < warp) before += s;
            total += s;
        }
        if (keep) {
            const int pos = running + before + __popc(bal & ((1u << lane) - 1u));
            kp.octave = (kp.octave & ~255) | ((kp.octave - 1) & 255);
            kp.x *= 0.5f;
            kp.y *= 0.5f;
            kp.size *= 0.5f;
            out[pos] = kp;
        }
        __syncthreads();
        if (tid == 0) running += total;
        __syncthreads();
    }
    if (tid == 0) *n_out = running;
}

// KeyPointsFilter::retainBest (nfeatures > 0): keep every keypoint whose response is >= the n_features-th largest one
// (boundary ties stay), in place and in the order of the list.  One CTA: radix select over the bit pattern of the
// responses (non-negative floats order like their bits; 11 + 11 + 10 bits), then an ordered compaction.
__global__ void __launch_bounds__(1024) retain_best_kernel(Keypoint* __restrict__ kps, int* __restrict__ n_io, int capacity,
                                                           int n_features) {
    __shared__ unsigned hist[2048];
    __shared__ unsigned sel_prefix;
    __shared__ int sel_remaining;
    __shared__ int warp_sum[32];
    __shared__ int running;
    const int n = min(*n_io, capacity);
    if (n_features <= 0 || n <= n_features) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned prefix = 0, prefix_mask = 0;
    int remaining = n_features;
    for (int pass = 0; pass < 3; ++pass) {
        const int shift = pass == 0 ? 21 : (pass == 1 ? 10 : 0);
        const int nb = pass == 2 ? 1024 : 2048;
        for (int b = tid; b < nb; b += 1024) hist[b] = 0;
        __syncthreads();
        for (int i = tid; i < n; i += 1024) {
            const unsigned bits = __float_as_uint(kps[i].response);
            if ((bits & prefix_mask) == prefix) atomicAdd(&hist[(bits >> shift) & (nb - 1)], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            int rem = remaining, b = nb - 1;
            for (; b > 0; --b) {
                if (hist[b] >= static_cast<unsigned>(rem)) break;
                rem -= static_cast<int>(hist[b]);
            }
            sel_prefix = prefix | (static_cast<unsigned>(b) << shift);
            sel_remaining = rem;
        }
        __syncthreads();
        prefix = sel_prefix;
        remaining = sel_remaining;
        prefix_mask |= static_cast<unsigned>(nb - 1) << shift;
        __syncthreads();
    }
    const unsigned threshold = prefix;
    if (tid == 0) running = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + tid;
        Keypoint kp{};
        bool keep = false;
        if (i < n) {
            kp = kps[i];
            keep = __float_as_uint(kp.response) >= threshold;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) warp_sum[warp] = __popc(bal);
        __syncthreads();                                   // every element of this chunk has been read
        int before = 0, total = 0;
        for (int wv = 0; wv < 32; ++wv) {
            const int sv = warp_sum[wv];
            if (wv < warp) before += sv;
            total += sv;
        }
        if (keep) kps[running + before + __popc(bal & ((1u << lane) - 1u))] = kp;      // position <= i
        __syncthreads();
        if (tid == 0) running += total;
        __syncthreads();
    }
    if (tid == 0) *n_io = running;
}

// One warp per keypoint (grid-stride).  The 32 lanes compute the votes of 32 consecutive samples of the window (row by
// row, OpenCV's k order) and park them in shared memory; the histogram is then updated sample by sample, the eight
// votes of a sample (eight distinct bins) by eight lanes at once: every bin receives its contributions in the serial
// order.  The norms are summed by one lane in element order; clipping and quantisation are element-wise.
constexpr int kDescWarps = 4;
struct DescScratch {
    float hist[kDescHistLen];
    float votes[32][9];          // padded: lane-major writes and sample-major reads both conflict-free
    int base[32];
};
__global__ void __launch_bounds__(kDescWarps * 32) descriptor_kernel(const PyramidView P, const Keypoint* __restrict__ kps,
                                                                    const int* __restrict__ n_kps, int capacity,
                                                                    uint8_t* __restrict__ desc, int first_octave) {
    __shared__ DescScratch scratch[kDescWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    DescScratch& S = scratch[warp];
    const int n = min(*n_kps, capacity);
    const int stride = gridDim.x * kDescWarps;
    const int my_off = descriptor_vote_offset(lane & 7);
    for (int i = blockIdx.x * kDescWarps + warp; i < n; i += stride) {
        const Keypoint kp = kps[i];
        int octave, layer;
        float scale;
        unpack_octave(kp.octave, octave, layer, scale);
        const int o = octave - first_octave;     // detection pyramid: firstOctave = -1
        float angle = 360.f - kp.angle;
        if (fabsf(angle - 360.f) < 1.1920929e-07f) angle = 0.f;
        const float* img = P.level(o, layer);
        const int cols = P.w[o], rows = P.h[o];
        const DescFrame F = descriptor_frame(cols, rows, kp.x * scale, kp.y * scale, angle, kp.size * scale * 0.5f);
        for (int b = lane; b < kDescHistLen; b += 32) S.hist[b] = 0.f;
        __syncwarp();
        const int side = 2 * F.radius + 1, total = side * side;
        for (int g0 = 0; g0 < total; g0 += 32) {
            const int g = g0 + lane;
            int idx = 0;
            float v[8];
            const bool valid = g < total && descriptor_sample(img, cols, rows, F, g / side - F.radius, g % side - F.radius, idx, v);
            unsigned mask = __ballot_sync(0xffffffffu, valid);
            if (mask == 0) continue;
            if (valid) {
                S.base[lane] = idx;
#pragma unroll
                for (int k = 0; k < 8; ++k) S.votes[lane][k] = v[k];
            }
            __syncwarp();
            while (mask) {
                const int l = __ffs(mask) - 1;
                mask &= mask - 1;
                if (lane < 8) S.hist[S.base[l] + my_off] += S.votes[l][lane];
                __syncwarp();
            }
        }
        float thr = 0.f;
        if (lane == 0) thr = descriptor_fold_and_threshold(S.hist, 1);
        thr = __shfl_sync(0xffffffffu, thr, 0);
        float f = 0.f;
        if (lane == 0) f = descriptor_clip_and_scale(S.hist, 1, thr);
        f = __shfl_sync(0xffffffffu, f, 0);
        __syncwarp();
        uint32_t packed = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            packed |= static_cast<uint32_t>(descriptor_quantise(S.hist[descriptor_element_index(lane * 4 + k)], f)) << (8 * k);
        reinterpret_cast<uint32_t*>(desc + static_cast<size_t>(i) * kDescLen)[lane] = packed;
        __syncwarp();
    }
}

__global__ void __launch_bounds__(256) keypoint_xy_kernel(const Keypoint* __restrict__ kps, int n, float2* __restrict__ xy) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < n) xy[i] = make_float2(kps[i].x, kps[i].y);
}

BlurWeights make_weights(double sigma) {
    // cv::GaussianBlur(src, dst, Size(), sigma) on CV_32F: ksize = cvRound(sigma * 4 * 2 + 1) | 1, exp kernel normalised in double
    BlurWeights k{};
    const int ksize = static_cast<int>(std::lrint(sigma * 8 + 1)) | 1;
    k.radius = (ksize - 1) / 2;
    if (k.radius > kMaxBlurRadius) { k.radius = -1; return k; }
    std::vector<double> v(ksize);
    double sum = 0;
    for (int i = 0; i < ksize; ++i) {
        const double x = i - k.radius;
        v[i] = std::exp(-(x * x) / (2.0 * sigma * sigma));
        sum += v[i];
    }
    for (int i = 0; i < ksize; ++i) k.w[i] = static_cast<float>(v[i] / sum);
    return k;
}

template <class T>
cudaError_t grow(T*& p, size_t& cap, size_t need) {
    if (need <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&p), need * sizeof(T));
    if (e == cudaSuccess) cap = need;
    return e;
}

}  // namespace

struct SiftWorkspace {
    uint8_t* d_gray = nullptr;   size_t gray_cap = 0;
    float* d_up = nullptr;       size_t up_cap = 0;        // upsampled grey image before the first blur
    float* d_pyr = nullptr;      size_t pyr_cap = 0;
    Candidate* d_cand = nullptr; size_t cand_cap = 0;
    Keypoint* d_kp_raw = nullptr; Keypoint* d_kp_sorted = nullptr; Keypoint* d_kp = nullptr; size_t kp_cap = 0, kp_cap2 = 0, kp_cap3 = 0;
    int* d_rank = nullptr;       size_t rank_cap = 0;
    int* d_grouped = nullptr;    size_t grouped_cap = 0;
    int* d_bucket = nullptr;     size_t bucket_cap = 0;   // count [nb + 1] | start [nb + 1] | cursor [nb + 1]
    uint8_t* d_desc = nullptr;   size_t desc_cap = 0;
    int* d_counts = nullptr;     // [0] candidates, [1] raw keypoints, [2] final keypoints, [4] / [5] min / max octave + 128 of the final keypoints
    int* h_counts = nullptr;     // pinned
    bool recomputed = false;     // the last extraction rebuilt the pyramid without upsampling for the descriptors (compute() semantics)
    PyramidView view{};
    bool smem_set = false;
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};      // start | pyramid built | descriptors written
    float pyramid_ms = 0.f, total_ms = 0.f;
};

SiftWorkspace* sift_workspace_create() { return new SiftWorkspace(); }

void sift_workspace_destroy(SiftWorkspace* w) {
    if (!w) return;
    cudaFree(w->d_gray); cudaFree(w->d_up); cudaFree(w->d_pyr); cudaFree(w->d_cand); cudaFree(w->d_kp_raw);
    cudaFree(w->d_kp_sorted); cudaFree(w->d_kp); cudaFree(w->d_rank); cudaFree(w->d_grouped); cudaFree(w->d_bucket); cudaFree(w->d_desc); cudaFree(w->d_counts);
    if (w->h_counts) cudaFreeHost(w->h_counts);
    for (cudaEvent_t e : w->ev) if (e) cudaEventDestroy(e);
    delete w;
}

const void* sift_keypoints_device_raw(const SiftWorkspace* w) { return w->d_kp; }
const uint8_t* sift_descriptors_device(const SiftWorkspace* w) { return w->d_desc; }
void sift_last_profile(const SiftWorkspace* w, double* pyramid_ms, double* total_ms, double* pyramid_bytes) {
    if (pyramid_ms) *pyramid_ms = w->pyramid_ms;
    if (total_ms) *total_ms = w->total_ms;
    // algorithmic HBM traffic of the pyramid: the u8 image in, every level written once and (all but the last of an octave) read once
    if (pyramid_bytes) {
        const PyramidView& P = w->view;
        double b = 0;
        for (int o = 0; o < P.n_octaves; ++o) b += 8.0 * (P.n_layers + 3) * static_cast<double>(P.w[o]) * P.h[o];
        *pyramid_bytes = b + (P.n_octaves ? 0.25 * P.w[0] * P.h[0] : 0.0);
    }
}
int sift_pyramid_geometry(const SiftWorkspace* w, int* n_layers, int* widths, int* heights, int64_t* offsets, const float** base) {
    const PyramidView& P = w->view;
    if (n_layers) *n_layers = P.n_layers;
    for (int o = 0; o < P.n_octaves; ++o) {
        if (widths) widths[o] = P.w[o];
        if (heights) heights[o] = P.h[o];
        if (offsets) offsets[o] = P.off[o];
    }
    if (base) *base = P.base;
    return P.n_octaves;
}

#define SIFT_TRY(expr)                                                                                  \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess) { if (err) *err = std::string(#expr) + ": " + cudaGetErrorString(_e); return _e; } \
    } while (0)

cudaError_t sift_extract(SiftWorkspace* ws, const uint8_t* gray, int rows, int cols, size_t step, const SiftParams& prm,
                         int max_keypoints, int sm_count, cudaStream_t s, int* n_keypoints, int* n_launches, int counts_out[3], std::string* err) {
    *n_keypoints = 0;
    if (counts_out) counts_out[0] = counts_out[1] = counts_out[2] = 0;
    int launches = 0;
    // ---- geometry: doubled base image, octave count of SIFT_Impl::detectAndCompute (firstOctave = -1)
    const int bw = 2 * cols, bh = 2 * rows;
    const int n_oct = static_cast<int>(std::lrint(std::log(static_cast<double>(std::min(bw, bh))) / std::log(2.0) - 2)) + 1;
    if (n_oct < 1) return cudaSuccess;                                     // image too small for a single octave
    if (n_oct > kMaxOctaves) { if (err) *err = "image too large: more than 16 octaves"; return cudaErrorInvalidValue; }
    const int L = prm.n_layers, levels = L + 3;
    PyramidView& P = ws->view;
    P.n_octaves = n_oct; P.n_layers = L;
    int64_t total = 0;
    for (int o = 0; o < n_oct; ++o) {
        P.w[o] = o == 0 ? bw : P.w[o - 1] / 2;
        P.h[o] = o == 0 ? bh : P.h[o - 1] / 2;
        if (P.w[o] < 1 || P.h[o] < 1) { P.n_octaves = o; break; }
        P.off[o] = total;
        total += static_cast<int64_t>(levels) * P.w[o] * P.h[o];
    }
    const int n_octaves = P.n_octaves;
    SIFT_TRY(grow(ws->d_gray, ws->gray_cap, static_cast<size_t>(rows) * cols));
    SIFT_TRY(grow(ws->d_up, ws->up_cap, static_cast<size_t>(bw) * bh));
    SIFT_TRY(grow(ws->d_pyr, ws->pyr_cap, static_cast<size_t>(total)));
    P.base = ws->d_pyr;
    const size_t cand_capacity = std::max<size_t>(1 << 16, static_cast<size_t>(max_keypoints) * 4);
    SIFT_TRY(grow(ws->d_cand, ws->cand_cap, cand_capacity));
    SIFT_TRY(grow(ws->d_kp_raw, ws->kp_cap, static_cast<size_t>(max_keypoints)));
    SIFT_TRY(grow(ws->d_kp_sorted, ws->kp_cap2, static_cast<size_t>(max_keypoints)));
    SIFT_TRY(grow(ws->d_kp, ws->kp_cap3, static_cast<size_t>(max_keypoints)));
    SIFT_TRY(grow(ws->d_rank, ws->rank_cap, static_cast<size_t>(max_keypoints)));
    SIFT_TRY(grow(ws->d_grouped, ws->grouped_cap, static_cast<size_t>(max_keypoints)));
    const int n_buckets = bw + 1;                 // floor(x) of a keypoint in base-image coordinates
    SIFT_TRY(grow(ws->d_bucket, ws->bucket_cap, 3 * static_cast<size_t>(n_buckets + 1)));
    SIFT_TRY(grow(ws->d_desc, ws->desc_cap, static_cast<size_t>(max_keypoints) * kDescLen));
    if (!ws->d_counts) SIFT_TRY(cudaMalloc(reinterpret_cast<void**>(&ws->d_counts), 32));
    if (!ws->h_counts) SIFT_TRY(cudaMallocHost(reinterpret_cast<void**>(&ws->h_counts), 32));
    if (!ws->smem_set) {
        for (int r : {0, 3, 4, 5, 6, 7, 8, 9, 10, 11, 13})
            SIFT_TRY(cudaFuncSetAttribute(blur_kernel_for(r), cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        ws->smem_set = true;
    }
    for (cudaEvent_t& e : ws->ev) if (!e) SIFT_TRY(cudaEventCreate(&e));
    SIFT_TRY(cudaMemsetAsync(ws->d_counts, 0, 32, s));
    SIFT_TRY(cudaMemsetAsync(ws->d_counts + 4, 0x7f, 4, s));          // [4] min octave + 128, [5] max octave + 128 of the final keypoints
    SIFT_TRY(cudaEventRecord(ws->ev[0], s));
    // ---- grey image to the device, doubled, first blur (createInitialImage)
    SIFT_TRY(cudaMemcpy2DAsync(ws->d_gray, cols, gray, step, cols, rows, cudaMemcpyHostToDevice, s));
    const dim3 blk(32, 8);
    upsample2x_kernel<<<dim3((bw + 31) / 32, (bh + 7) / 8), blk, 0, s>>>(ws->d_gray, cols, rows, cols, ws->d_up);
    ++launches;
    auto blur = [&](const float* src, float* dst, int w, int h, double sigma) -> cudaError_t {
        const BlurWeights k = make_weights(sigma);
        if (k.radius < 0) { if (err) *err = "Gaussian kernel radius above 32 (sigma / nOctaveLayers out of the supported range)"; return cudaErrorInvalidValue; }
        const size_t ih = static_cast<size_t>(kBlurTH + 2 * k.radius);
        const size_t smem = (ih * blur_in_stride(k.radius) + 8 + (ih + 8) * kBlurMidStride) * sizeof(float);
        blur_kernel_for(k.radius)<<<dim3((w + kBlurTW - 1) / kBlurTW, (h + kBlurTH - 1) / kBlurTH), 256, smem, s>>>(src, dst, w, h, k);
        ++launches;
        return cudaGetLastError();
    };
    const float sigma_f = prm.sigma;
    const float sig_diff = sqrtf(std::max(sigma_f * sigma_f - 0.5f * 0.5f * 4, 0.01f));
    SIFT_TRY(blur(ws->d_up, ws->d_pyr + P.off[0], bw, bh, static_cast<double>(sig_diff)));
    // ---- buildGaussianPyramid
    std::vector<double> sig(levels);
    sig[0] = prm.sigma;
    const double kfac = std::pow(2.0, 1.0 / L);
    for (int i = 1; i < levels; ++i) {
        const double sig_prev = std::pow(kfac, static_cast<double>(i - 1)) * prm.sigma;
        const double sig_total = sig_prev * kfac;
        sig[i] = std::sqrt(sig_total * sig_total - sig_prev * sig_prev);
    }
    // level 0 of octave 0 is in place: the remaining levels and octaves (buildGaussianPyramid)
    auto build_levels = [&](int octaves) -> cudaError_t {
        for (int o = 0; o < octaves; ++o) {
            const int w = P.w[o], h = P.h[o];
            const size_t plane = static_cast<size_t>(w) * h;
            float* lv0 = ws->d_pyr + P.off[o];
            if (o > 0) {
                const float* src = ws->d_pyr + P.off[o - 1] + static_cast<size_t>(L) * P.w[o - 1] * P.h[o - 1];
                downsample_kernel<<<dim3((w + 31) / 32, (h + 7) / 8), blk, 0, s>>>(src, P.w[o - 1], lv0, w, h);
                ++launches;
            }
            for (int i = 1; i < levels; ++i) {
                const cudaError_t e = blur(lv0 + (i - 1) * plane, lv0 + i * plane, w, h, sig[i]);
                if (e != cudaSuccess) return e;
            }
        }
        return cudaSuccess;
    };
    SIFT_TRY(build_levels(n_octaves));
    SIFT_TRY(cudaEventRecord(ws->ev[1], s));
    // ---- findScaleSpaceExtrema
    const float threshold = static_cast<float>(static_cast<int>(std::floor(0.5 * prm.contrast_threshold / L * 255)));
    for (int o = 0; o < n_octaves; ++o) {
        const int iw = P.w[o] - 2 * kImgBorder, ih = P.h[o] - 2 * kImgBorder;
        if (iw <= 0 || ih <= 0) continue;
        extrema_kernel<<<dim3((iw + 31) / 32, (ih + 7) / 8, L), blk, 0, s>>>(P, o, threshold, ws->d_cand, ws->d_counts,
                                                                              static_cast<int>(cand_capacity));
        ++launches;
    }
    const unsigned persistent = static_cast<unsigned>(sm_count > 0 ? sm_count : 148) * 8;
    refine_orient_kernel<<<persistent, kRefineWarps * 32, 0, s>>>(
        P, ws->d_cand, ws->d_counts, static_cast<int>(cand_capacity), static_cast<float>(prm.contrast_threshold),
        static_cast<float>(prm.edge_threshold), static_cast<float>(prm.sigma), ws->d_kp_raw, ws->d_counts + 1, max_keypoints);
    // ---- removeDuplicatedSorted + firstOctave correction
    int* b_count = ws->d_bucket;
    int* b_start = ws->d_bucket + (n_buckets + 1);
    int* b_cursor = ws->d_bucket + 2 * (n_buckets + 1);
    SIFT_TRY(cudaMemsetAsync(b_count, 0, static_cast<size_t>(n_buckets + 1) * sizeof(int), s));
    bucket_count_kernel<<<persistent, 256, 0, s>>>(ws->d_kp_raw, ws->d_counts + 1, max_keypoints, n_buckets, b_count);
    bucket_scan_kernel<<<1, 1024, 0, s>>>(b_count, n_buckets, b_start, b_cursor);
    bucket_fill_kernel<<<persistent, 256, 0, s>>>(ws->d_kp_raw, ws->d_counts + 1, max_keypoints, n_buckets, b_cursor, ws->d_grouped);
    bucket_rank_kernel<<<persistent, 256, 0, s>>>(ws->d_kp_raw, ws->d_counts + 1, max_keypoints, n_buckets, b_start, ws->d_grouped,
                                                  ws->d_rank);
    scatter_kernel<<<persistent, 256, 0, s>>>(ws->d_kp_raw, ws->d_counts + 1, max_keypoints, ws->d_rank, ws->d_kp_sorted);
    dedupe_kernel<<<1, 1024, 0, s>>>(ws->d_kp_sorted, ws->d_counts + 1, max_keypoints, ws->d_kp, ws->d_counts + 2);
    if (prm.n_features > 0) {
        retain_best_kernel<<<1, 1024, 0, s>>>(ws->d_kp, ws->d_counts + 2, max_keypoints, prm.n_features);
        ++launches;
    }
    // ---- calcDescriptors
    descriptor_kernel<<<persistent, kDescWarps * 32, 0, s>>>(P, ws->d_kp, ws->d_counts + 2, max_keypoints, ws->d_desc, -1);
    octave_range_kernel<<<64, 256, 0, s>>>(ws->d_kp, ws->d_counts + 2, max_keypoints, ws->d_counts + 4);
    launches += 9;
    SIFT_TRY(cudaGetLastError());
    SIFT_TRY(cudaEventRecord(ws->ev[2], s));
    SIFT_TRY(cudaMemcpyAsync(ws->h_counts, ws->d_counts, 24, cudaMemcpyDeviceToHost, s));
    SIFT_TRY(cudaStreamSynchronize(s));
    // The reference calls detect() and compute() separately (SfM.cpp:586-587), and cv::SIFT::compute rebuilds the pyramid from
    // the octave range of the keypoints it is given: when NONE lies in octave -1 the image is not upsampled
    // (createInitialImage(img, doubleImageSize = false): grey -> float -> blur sqrt(sigma^2 - 0.25)) and only max octave + 1
    // octaves are built, so the descriptors come from slightly different pixels.  Photographs always have octave -1 keypoints;
    // a picture of a few wide blobs does not: mirror compute() then (one more pass, rare).
    {
        const int n_final = std::min(ws->h_counts[2], max_keypoints);
        const int oct_min = ws->h_counts[4] - 128, oct_max = ws->h_counts[5] - 128;
        if (n_final > 0 && ws->h_counts[1] <= max_keypoints && oct_min >= 0) {
            const int n_oct2 = oct_max + 1;
            int64_t total2 = 0;
            P.n_octaves = 0;
            for (int o = 0; o < n_oct2; ++o) {
                P.w[o] = o == 0 ? cols : P.w[o - 1] / 2;
                P.h[o] = o == 0 ? rows : P.h[o - 1] / 2;
                if (P.w[o] < 1 || P.h[o] < 1) break;
                P.off[o] = total2;
                total2 += static_cast<int64_t>(levels) * P.w[o] * P.h[o];
                P.n_octaves = o + 1;
            }
            if (P.n_octaves != n_oct2) { if (err) *err = "feature extraction: keypoint octave beyond the non-doubled pyramid"; return cudaErrorInvalidValue; }
            // d_pyr / d_up are large enough: the doubled pyramid had four times the pixels per octave
            gray_to_float_kernel<<<dim3((cols + 31) / 32, (rows + 7) / 8), blk, 0, s>>>(ws->d_gray, cols, rows, cols, ws->d_up);
            ++launches;
            const float sig_diff1 = sqrtf(std::max(sigma_f * sigma_f - 0.5f * 0.5f, 0.01f));
            SIFT_TRY(blur(ws->d_up, ws->d_pyr + P.off[0], cols, rows, static_cast<double>(sig_diff1)));
            SIFT_TRY(build_levels(n_oct2));
            descriptor_kernel<<<persistent, kDescWarps * 32, 0, s>>>(P, ws->d_kp, ws->d_counts + 2, max_keypoints, ws->d_desc, 0);
            ++launches;
            SIFT_TRY(cudaGetLastError());
            SIFT_TRY(cudaEventRecord(ws->ev[2], s));
            SIFT_TRY(cudaStreamSynchronize(s));
            ws->recomputed = true;
        } else ws->recomputed = false;
    }
    SIFT_TRY(cudaEventElapsedTime(&ws->pyramid_ms, ws->ev[0], ws->ev[1]));
    SIFT_TRY(cudaEventElapsedTime(&ws->total_ms, ws->ev[0], ws->ev[2]));
    if (counts_out) { counts_out[0] = ws->h_counts[0]; counts_out[1] = ws->h_counts[1]; counts_out[2] = ws->h_counts[2]; }
    if (ws->h_counts[0] > static_cast<int>(cand_capacity) || ws->h_counts[1] > max_keypoints) {
        if (err) *err = "feature extraction: more candidates / keypoints than the capacity (" + std::to_string(ws->h_counts[0]) + " / " +
                        std::to_string(ws->h_counts[1]) + "): raise max_keypoints";
        return cudaErrorMemoryAllocation;
    }
    *n_keypoints = ws->h_counts[2];
    if (n_launches) *n_launches = launches;
    return cudaSuccess;
}

cudaError_t launch_keypoint_xy(const void* keypoints, int n, float2* xy, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    keypoint_xy_kernel<<<(n + 255) / 256, 256, 0, s>>>(static_cast<const Keypoint*>(keypoints), n, xy);
    return cudaGetLastError();
}

}  // namespace sfm
