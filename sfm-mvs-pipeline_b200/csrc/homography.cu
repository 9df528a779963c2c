// homography.cu — RANSAC homography inlier ratio per image pair, on the device-resident match lists.
//
// Replaces SfM::calculateHomography (SfM.cpp:599-637): for every ShotMatches with >= 4 matches
//     cv::findHomography(left points, right points, cv::RANSAC, ransacThreshold, inlierMask)
//     homographyInlierRatio = countNonZero(inlierMask) / matches.size()
// (left <-> queryIdx, right <-> trainIdx, ShotMatches::alignFeatures, Scene.cpp:95-110).  OpenCV's mask is the consensus
// set of the best minimal (4-point) model its RANSAC loop found (the later least-squares / LM refinement does not touch
// the mask), i.e. the number the pipeline consumes is  max over sampled 4-point models of #{i : |m_i - H M_i|^2 <= thr^2}.
//
// Here: one CTA per pair, one thread per hypothesis (max_iters hypotheses, OpenCV's default 2000 = the upper bound of its
// adaptive loop), all threads walk the pair's matches together (shared-memory broadcast).  Minimal sets come from a
// counter-based generator (splitmix64 of seed, pair, hypothesis, draw) so that the numpy restatement used by the
// tests reproduces every hypothesis; degenerate sets are rejected like cv::HomographyEstimatorCallback::
// checkSubset (collinear triples, orientation of the four triples must agree between the two images).  The 4-point model is
// the closed form  H = S_dst * adj(S_src)  with S = unit-square-to-quadrilateral map (Heckbert), in double; the error is the
// forward transfer error in the right image, compared with thr^2 exactly as computeError / findInliers do.
// Not bit-exact with OpenCV (different random minimal sets): parity is the inlier RATIO within a stated tolerance
// (tests/test_gpu_homography.py), and bit-level agreement with the restatement of THIS algorithm.
// Compiled with --fmad=false so that the double arithmetic matches the numpy restatement operation for operation.
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "kernels.h"

namespace sfm {

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// unit square (0,0),(1,0),(1,1),(0,1) -> quadrilateral p0..p3; returns false if the quadrilateral is degenerate
__device__ __forceinline__ bool square_to_quad(const double (&x)[4], const double (&y)[4], double (&S)[9]) {
    const double dx1 = x[1] - x[2], dx2 = x[3] - x[2], sx = x[0] - x[1] + x[2] - x[3];
    const double dy1 = y[1] - y[2], dy2 = y[3] - y[2], sy = y[0] - y[1] + y[2] - y[3];
    const double den = dx1 * dy2 - dy1 * dx2;
    if (den == 0.0) return false;
    const double g = (sx * dy2 - sy * dx2) / den;
    const double h = (dx1 * sy - dy1 * sx) / den;
    S[0] = x[1] - x[0] + g * x[1]; S[1] = x[3] - x[0] + h * x[3]; S[2] = x[0];
    S[3] = y[1] - y[0] + g * y[1]; S[4] = y[3] - y[0] + h * y[3]; S[5] = y[0];
    S[6] = g; S[7] = h; S[8] = 1.0;
    return true;
}

__device__ __forceinline__ double cross3(double ax, double ay, double bx, double by, double cx, double cy) {
    return (bx - ax) * (cy - ay) - (by - ay) * (cx - ax);
}

// cv::HomographyEstimatorCallback::checkSubset: no collinear triple in either image, and the orientation of every triple
// is preserved (all four sign products agree)
__device__ __forceinline__ bool subset_ok(const double (&x1)[4], const double (&y1)[4], const double (&x2)[4], const double (&y2)[4]) {
    int negative = 0;
#pragma unroll
    for (int skip = 0; skip < 4; ++skip) {
        const int a = skip == 0 ? 1 : 0, b = skip <= 1 ? 2 : 1, c = skip <= 2 ? 3 : 2;
        const double c1 = cross3(x1[a], y1[a], x1[b], y1[b], x1[c], y1[c]);
        const double c2 = cross3(x2[a], y2[a], x2[b], y2[b], x2[c], y2[c]);
        const double s1 = fabs(x1[b] - x1[a]) + fabs(y1[b] - y1[a]) + fabs(x1[c] - x1[a]) + fabs(y1[c] - y1[a]);
        const double s2 = fabs(x2[b] - x2[a]) + fabs(y2[b] - y2[a]) + fabs(x2[c] - x2[a]) + fabs(y2[c] - y2[a]);
        if (fabs(c1) <= 1.1920928955078125e-7 * s1 || fabs(c2) <= 1.1920928955078125e-7 * s2) return false;
        negative += (c1 * c2 < 0.0) ? 1 : 0;
    }
    return negative == 0 || negative == 4;
}

constexpr int kHomThreads = 256;
constexpr int kHomSmemPoints = 3072;            // matches staged in shared memory (4 floats each); more stream from L2

__global__ void __launch_bounds__(kHomThreads) homography_ransac_kernel(HomographyArgs a) {
    extern __shared__ float4 s_pts[];           // (x1, y1, x2, y2) per match
    __shared__ unsigned long long s_best[kHomThreads / 32];
    const int p = blockIdx.x;
    const int64_t m0 = a.pair_offsets[p];
    const int64_t m1 = (p + 1 < a.n_pairs) ? a.pair_offsets[p + 1] : *a.total;
    const int M = static_cast<int>(m1 - m0);
    const bool skip = M < 4 || (a.dropped && a.dropped[p]);
    if (skip) {                                 // "Homographie kann nicht gefunden werden": ratio stays -1 (SfM.cpp:606-609)
        if (threadIdx.x == 0) {
            a.inliers[p] = -1;
            if (a.best_hyp) a.best_hyp[p] = -1;
        }
        return;
    }
    const float2* kq = a.keypoints + a.row0[2 * p];
    const float2* kt = a.keypoints + a.row0[2 * p + 1];
    const DMatch* mm = a.matches + m0;
    const int staged = min(M, kHomSmemPoints);
    for (int i = threadIdx.x; i < staged; i += kHomThreads) {
        const DMatch d = mm[i];
        const float2 q = kq[d.queryIdx], t = kt[d.trainIdx];
        s_pts[i] = make_float4(q.x, q.y, t.x, t.y);
    }
    __syncthreads();
    auto point = [&](int i) -> float4 {
        if (i < staged) return s_pts[i];
        const DMatch d = mm[i];
        const float2 q = kq[d.queryIdx], t = kt[d.trainIdx];
        return make_float4(q.x, q.y, t.x, t.y);
    };
    const double thr = a.n_thresholds > 1 ? a.thresholds[p] : a.thresholds[0];
    const double thr2 = thr * thr;
    unsigned long long best = 0;                // (count << 32) | (0xFFFFFFFF - hypothesis): max count, lowest hypothesis
    for (int h0 = 0; h0 < a.max_iters; h0 += kHomThreads) {
        const int h = h0 + threadIdx.x;
        bool ok = h < a.max_iters;
        double H[9];
        if (ok) {
            // ---- minimal set: four distinct matches
            int idx[4];
            const uint64_t key = splitmix64(a.seed ^ (static_cast<uint64_t>(p) << 32) ^ static_cast<uint64_t>(h));
            int got = 0;
            for (int draw = 0; draw < 16 && got < 4; ++draw) {
                const uint64_t r = splitmix64(key + static_cast<uint64_t>(draw));
                const int cand = static_cast<int>(((r >> 32) * static_cast<uint64_t>(M)) >> 32);
                bool dup = false;
                for (int j = 0; j < got; ++j) dup |= idx[j] == cand;
                if (!dup) idx[got++] = cand;
            }
            ok = got == 4;
            double x1[4], y1[4], x2[4], y2[4];
            if (ok) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 pt = point(idx[j]);
                    x1[j] = pt.x; y1[j] = pt.y; x2[j] = pt.z; y2[j] = pt.w;
                }
                ok = subset_ok(x1, y1, x2, y2);
            }
            // ---- model: H = S_dst * adj(S_src)
            double S1[9], S2[9];
            if (ok) ok = square_to_quad(x1, y1, S1) && square_to_quad(x2, y2, S2);
            if (ok) {
                double A[9];                    // adjugate of S1
                A[0] = S1[4] * S1[8] - S1[5] * S1[7]; A[1] = S1[2] * S1[7] - S1[1] * S1[8]; A[2] = S1[1] * S1[5] - S1[2] * S1[4];
                A[3] = S1[5] * S1[6] - S1[3] * S1[8]; A[4] = S1[0] * S1[8] - S1[2] * S1[6]; A[5] = S1[2] * S1[3] - S1[0] * S1[5];
                A[6] = S1[3] * S1[7] - S1[4] * S1[6]; A[7] = S1[1] * S1[6] - S1[0] * S1[7]; A[8] = S1[0] * S1[4] - S1[1] * S1[3];
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        H[3 * r + c] = S2[3 * r] * A[c] + S2[3 * r + 1] * A[3 + c] + S2[3 * r + 2] * A[6 + c];
                ok = H[8] != 0.0 && isfinite(H[8]);
                if (ok) {
                    const double inv = 1.0 / H[8];                    // OpenCV scales the model to H[8] = 1
#pragma unroll
                    for (int i = 0; i < 9; ++i) H[i] *= inv;
                }
            }
        }
        // ---- consensus: every thread walks the matches (a warp reads the same point: broadcast)
        int count = 0;
        if (__any_sync(0xffffffffu, ok)) {
            for (int i = 0; i < M; ++i) {
                const float4 pt = point(i);
                if (ok) {
                    const double X = pt.x, Y = pt.y;
                    const double w = H[6] * X + H[7] * Y + 1.0;
                    const double ww = 1.0 / w;
                    const double dx = (H[0] * X + H[1] * Y + H[2]) * ww - static_cast<double>(pt.z);
                    const double dy = (H[3] * X + H[4] * Y + H[5]) * ww - static_cast<double>(pt.w);
                    const double err = dx * dx + dy * dy;
                    count += (err <= thr2) ? 1 : 0;                   // NaN compares false
                }
            }
        }
        if (ok && count > 0) {
            const unsigned long long k = (static_cast<unsigned long long>(count) << 32) | (0xFFFFFFFFu - static_cast<unsigned>(h));
            best = max(best, k);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
    if ((threadIdx.x & 31) == 0) s_best[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kHomThreads / 32; ++w) best = max(best, s_best[w]);
        a.inliers[p] = static_cast<int32_t>(best >> 32);
        if (a.best_hyp) a.best_hyp[p] = best ? static_cast<int32_t>(0xFFFFFFFFu - static_cast<unsigned>(best & 0xFFFFFFFFu)) : -1;
    }
}

cudaError_t launch_homography_ransac(const HomographyArgs& a, cudaStream_t s) {
    if (a.n_pairs == 0) return cudaSuccess;
    const size_t smem = static_cast<size_t>(kHomSmemPoints) * sizeof(float4);
    cudaError_t e = cudaFuncSetAttribute(homography_ransac_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    homography_ransac_kernel<<<static_cast<unsigned>(a.n_pairs), kHomThreads, smem, s>>>(a);
    return cudaGetLastError();
}

}  // namespace sfm
