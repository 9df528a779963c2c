// homography.cu — RANSAC homography inlier ratio per image pair, on the device-resident match lists.
//
// Replaces SfM::calculateHomography (SfM.cpp:599-637): for every ShotMatches with >= 4 matches
//     cv::findHomography(left points, right points, cv::RANSAC, ransacThreshold, inlierMask)
//     homographyInlierRatio = countNonZero(inlierMask) / matches.size()
// (left <-> queryIdx, right <-> trainIdx, ShotMatches::alignFeatures, Scene.cpp:95-110).  OpenCV's mask is the consensus
// set of the best minimal (4-point) model its RANSAC loop found (the later least-squares / LM refinement does not touch
// the mask), i.e. the number the pipeline consumes is  max over sampled 4-point models of #{i : |m_i - H M_i|^2 <= thr^2}.
//
// Here: one CTA per pair, one thread per hypothesis (max_iters hypotheses, OpenCV's default 2000), all threads walk the
// pair's matches together (shared-memory broadcast); the consensus sizes land in shared memory and one thread replays
// OpenCV's adaptive loop over them in hypothesis order (RANSACPointSetRegistrator::run: a new best model shrinks the
// iteration budget through RANSACUpdateNumIters(confidence, outlier ratio, 4, niters); rejected minimal sets do not count
// as iterations because getSubset redraws them), so the answer has the same early-stopping behaviour as the reference.  Minimal sets come from a
// counter-based generator (nested splitmix64 of seed, pair, hypothesis, draw) so that the numpy restatement used by the
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

// deterministic block sum: strided per-thread partials -> xor-butterfly inside each warp -> warp partials added in warp
// order by every thread (the numpy restatement performs the same additions in the same order)
__device__ __forceinline__ double block_sum(double v, double* s_red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();                            // previous use of s_red is over
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = s_red[0];
#pragma unroll
    for (int w = 1; w < kHomThreads / 32; ++w) t += s_red[w];
    return t;
}

__device__ __forceinline__ bool inlier_under(const double* H, const float4 pt, double thr2) {
    const double X = pt.x, Y = pt.y;
    const double w = H[6] * X + H[7] * Y + 1.0;
    const double ww = 1.0 / w;
    const double dx = (H[0] * X + H[1] * Y + H[2]) * ww - static_cast<double>(pt.z);
    const double dy = (H[3] * X + H[4] * Y + H[5]) * ww - static_cast<double>(pt.w);
    return dx * dx + dy * dy <= thr2;           // NaN compares false
}

__global__ void __launch_bounds__(kHomThreads) homography_ransac_kernel(HomographyArgs a) {
    extern __shared__ float4 s_pts[];           // (x1, y1, x2, y2) per match, then int32 consensus size per hypothesis
    int32_t* s_cnt = reinterpret_cast<int32_t*>(s_pts + kHomSmemPoints);
    __shared__ double s_red[kHomThreads / 32];
    __shared__ double s_H[9];
    __shared__ int s_best[2];
    const int p = blockIdx.x;
    const int64_t m0 = a.pair_offsets[p];
    const int64_t m1 = (p + 1 < a.n_pairs) ? a.pair_offsets[p + 1] : *a.total;
    const int M = static_cast<int>(m1 - m0);
    const bool skip = M < 4 || (a.dropped && a.dropped[p]);
    if (skip || M == 4) {
        // < 4 matches / dropped pair: "Homographie kann nicht gefunden werden", the ratio stays -1 (SfM.cpp:606-609).
        // Exactly four point pairs: cv::findHomography runs no RANSAC, the 4-point solution comes back with an all-ones
        // mask (calib3d fundam.cpp: `if (method == 0 || npoints == 4)`), i.e. 4 inliers of 4.
        if (threadIdx.x == 0) {
            a.inliers[p] = skip ? -1 : 4;
            if (a.ransac_inliers) a.ransac_inliers[p] = skip ? -1 : 4;
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
    // minimal model of hypothesis h (H[8] = 1); false if the minimal set is rejected
    auto minimal_model = [&](int h, double (&H)[9]) -> bool {
        int idx[4];
        const uint64_t key = splitmix64(splitmix64(splitmix64(a.seed) ^ static_cast<uint64_t>(p)) + static_cast<uint64_t>(h));
        int got = 0;
        for (int draw = 0; draw < 16 && got < 4; ++draw) {
            const uint64_t r = splitmix64(key + static_cast<uint64_t>(draw));
            const int cand = static_cast<int>(((r >> 32) * static_cast<uint64_t>(M)) >> 32);
            bool dup = false;
            for (int j = 0; j < got; ++j) dup |= idx[j] == cand;
            if (!dup) idx[got++] = cand;
        }
        if (got != 4) return false;
        double x1[4], y1[4], x2[4], y2[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float4 pt = point(idx[j]);
            x1[j] = pt.x; y1[j] = pt.y; x2[j] = pt.z; y2[j] = pt.w;
        }
        if (!subset_ok(x1, y1, x2, y2)) return false;
        double S1[9], S2[9];                    // H = S_dst * adj(S_src)
        if (!(square_to_quad(x1, y1, S1) && square_to_quad(x2, y2, S2))) return false;
        double A[9];
        A[0] = S1[4] * S1[8] - S1[5] * S1[7]; A[1] = S1[2] * S1[7] - S1[1] * S1[8]; A[2] = S1[1] * S1[5] - S1[2] * S1[4];
        A[3] = S1[5] * S1[6] - S1[3] * S1[8]; A[4] = S1[0] * S1[8] - S1[2] * S1[6]; A[5] = S1[2] * S1[3] - S1[0] * S1[5];
        A[6] = S1[3] * S1[7] - S1[4] * S1[6]; A[7] = S1[1] * S1[6] - S1[0] * S1[7]; A[8] = S1[0] * S1[4] - S1[1] * S1[3];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c)
                H[3 * r + c] = S2[3 * r] * A[c] + S2[3 * r + 1] * A[3 + c] + S2[3 * r + 2] * A[6 + c];
        if (!(H[8] != 0.0 && isfinite(H[8]))) return false;
        const double inv = 1.0 / H[8];          // OpenCV scales the model to H[8] = 1
#pragma unroll
        for (int i = 0; i < 9; ++i) H[i] *= inv;
        return true;
    };
    const double thr = a.n_thresholds > 1 ? a.thresholds[p] : a.thresholds[0];
    const double thr2 = thr * thr;

    // ---- stages 1 + 2 in rounds of 256 hypotheses: consensus sizes (a warp reads the same point: shared-memory broadcast),
    // then one thread continues the replay of RANSACPointSetRegistrator::run over the new sizes in hypothesis order -- a better
    // model shrinks the budget through RANSACUpdateNumIters(confidence, outlier ratio, 4, niters), rejected minimal sets do
    // not count as iterations.  The loop ends as soon as the replay has consumed its budget: a pair with 99 % inliers is done
    // after the first round (OpenCV: after ~3 iterations) instead of after all max_iters hypotheses.
    __shared__ double s_niters;
    __shared__ int s_iter, s_stop;
    if (threadIdx.x == 0) { s_best[0] = 0; s_best[1] = -1; s_niters = static_cast<double>(a.max_iters); s_iter = 0; s_stop = 0; }
    __syncthreads();
    for (int h0 = 0; h0 < a.max_iters; h0 += kHomThreads) {
        const int h = h0 + threadIdx.x;
        double H[9];
        const bool ok = h < a.max_iters && minimal_model(h, H);
        int count = 0;
        if (__any_sync(0xffffffffu, ok)) {
            for (int i = 0; i < M; ++i) {
                const float4 pt = point(i);
                if (ok) count += inlier_under(H, pt, thr2) ? 1 : 0;
            }
        }
        if (h < a.max_iters) s_cnt[h] = ok ? count : -1;           // -1: minimal set rejected (redrawn by OpenCV)
        __syncthreads();
        if (threadIdx.x == 0) {
            int best = s_best[0], best_h = s_best[1], iter = s_iter;
            double niters = s_niters;
            const int h_end = min(a.max_iters, h0 + kHomThreads);
            int hh = h0;
            for (; hh < h_end && iter < niters; ++hh) {
                const int cnt = s_cnt[hh];
                if (cnt < 0) continue;
                if (cnt > max(best, 3)) {                           // goodCount > max(maxGoodCount, modelPoints - 1)
                    best = cnt; best_h = hh;
                    const double ep = static_cast<double>(M - cnt) / static_cast<double>(M);
                    const double num = fmax(1.0 - a.confidence, 2.2250738585072014e-308);
                    const double q = 1.0 - ep;
                    const double denom = 1.0 - q * q * q * q;
                    if (denom < 2.2250738585072014e-308) niters = 0.0;
                    else {
                        const double ln = log(num), ld = log(denom);
                        niters = (ld >= 0.0 || -ln >= niters * (-ld)) ? niters : rint(ln / ld);
                    }
                }
                ++iter;
            }
            s_best[0] = best; s_best[1] = best_h; s_iter = iter; s_niters = niters;
            s_stop = !(iter < niters) ? 1 : 0;                      // budget consumed: later hypotheses are never drawn
        }
        __syncthreads();
        if (s_stop) break;
    }
    if (threadIdx.x == 0 && s_best[1] >= 0) {
        double H[9];
        minimal_model(s_best[1], H);
        for (int i = 0; i < 9; ++i) s_H[i] = H[i];
    }
    __syncthreads();
    const int ransac_count = s_best[0], best_h = s_best[1];
    int final_count = ransac_count;
    // ---- stage 3: least-squares re-estimation on the consensus set and the mask of the refined model.
    // cv::findHomography (npoints > 4): runKernel on the inliers (Hartley-normalised DLT), LM polish, and the returned
    // mask is the set of matches within the threshold of THAT model (checked against cv2 on every fixture).  Here: the
    // same normalisation and the same normal matrix LtL, solved with h33 = 1 in the normalised frame (no LM).
    if (best_h >= 0 && a.refine) {
        // pass A: centroids of the inliers
        double sx1 = 0, sy1 = 0, sx2 = 0, sy2 = 0, sn = 0;
        for (int i = threadIdx.x; i < M; i += kHomThreads) {
            const float4 pt = point(i);
            if (inlier_under(s_H, pt, thr2)) { sx1 += pt.x; sy1 += pt.y; sx2 += pt.z; sy2 += pt.w; sn += 1.0; }
        }
        const double n = block_sum(sn, s_red);
        const double c1x = block_sum(sx1, s_red) / n, c1y = block_sum(sy1, s_red) / n;
        const double c2x = block_sum(sx2, s_red) / n, c2y = block_sum(sy2, s_red) / n;
        // pass B: mean absolute deviation per axis (cv::HomographyEstimatorCallback::runKernel)
        double ax1 = 0, ay1 = 0, ax2 = 0, ay2 = 0;
        for (int i = threadIdx.x; i < M; i += kHomThreads) {
            const float4 pt = point(i);
            if (inlier_under(s_H, pt, thr2)) {
                ax1 += fabs(pt.x - c1x); ay1 += fabs(pt.y - c1y); ax2 += fabs(pt.z - c2x); ay2 += fabs(pt.w - c2y);
            }
        }
        const double d1x = block_sum(ax1, s_red), d1y = block_sum(ay1, s_red);
        const double d2x = block_sum(ax2, s_red), d2y = block_sum(ay2, s_red);
        const double eps = 2.220446049250313e-16;
        const bool scale_ok = d1x > eps && d1y > eps && d2x > eps && d2y > eps;     // runKernel returns 0 otherwise
        if (scale_ok) {
            const double s1x = n / d1x, s1y = n / d1y, s2x = n / d2x, s2y = n / d2y;
            // pass C: upper triangle of LtL (9 x 9) over the inliers, normalised coordinates
            double L[45];
#pragma unroll
            for (int k = 0; k < 45; ++k) L[k] = 0.0;
            for (int i = threadIdx.x; i < M; i += kHomThreads) {
                const float4 pt = point(i);
                if (!inlier_under(s_H, pt, thr2)) continue;
                const double X = (pt.x - c1x) * s1x, Y = (pt.y - c1y) * s1y;
                const double x = (pt.z - c2x) * s2x, y = (pt.w - c2y) * s2y;
                const double Lx[9] = {X, Y, 1.0, 0.0, 0.0, 0.0, -x * X, -x * Y, -x};
                const double Ly[9] = {0.0, 0.0, 0.0, X, Y, 1.0, -y * X, -y * Y, -y};
                int k = 0;
#pragma unroll
                for (int r = 0; r < 9; ++r)
#pragma unroll
                    for (int c = r; c < 9; ++c) { L[k] += Lx[r] * Lx[c] + Ly[r] * Ly[c]; ++k; }
            }
            double T[45];
#pragma unroll
            for (int k = 0; k < 45; ++k) T[k] = block_sum(L[k], s_red);
            if (threadIdx.x == 0) {
                // minimise h^T LtL h with h[8] = 1: LtL[0:8, 0:8] h = -LtL[0:8, 8]; Gaussian elimination, partial pivoting
                double Mx[8][9];
                int k = 0;
                for (int r = 0; r < 9; ++r)
                    for (int c = r; c < 9; ++c) {
                        if (r < 8 && c < 8) { Mx[r][c] = T[k]; Mx[c][r] = T[k]; }
                        else if (r < 8) Mx[r][8] = -T[k];
                        ++k;
                    }
                bool solved = true;
                for (int col = 0; col < 8 && solved; ++col) {
                    int piv = col;
                    for (int r = col + 1; r < 8; ++r) if (fabs(Mx[r][col]) > fabs(Mx[piv][col])) piv = r;
                    if (!(fabs(Mx[piv][col]) > 0.0)) { solved = false; break; }
                    if (piv != col) for (int c = 0; c < 9; ++c) { const double t = Mx[col][c]; Mx[col][c] = Mx[piv][c]; Mx[piv][c] = t; }
                    for (int r = col + 1; r < 8; ++r) {
                        const double f = Mx[r][col] / Mx[col][col];
                        for (int c = col; c < 9; ++c) Mx[r][c] -= f * Mx[col][c];
                    }
                }
                double hn[9];
                hn[8] = 1.0;
                if (solved) {
                    for (int r = 7; r >= 0; --r) {
                        double acc = Mx[r][8];
                        for (int c = r + 1; c < 8; ++c) acc -= Mx[r][c] * hn[c];
                        hn[r] = acc / Mx[r][r];
                    }
                    // H = inv(T2) * Hn * T1,  T = [s_x 0 -s_x c_x; 0 s_y -s_y c_y; 0 0 1]
                    double G[9];                                    // Hn * T1
                    for (int r = 0; r < 3; ++r) {
                        G[3 * r] = hn[3 * r] * s1x;
                        G[3 * r + 1] = hn[3 * r + 1] * s1y;
                        G[3 * r + 2] = hn[3 * r + 2] - hn[3 * r] * s1x * c1x - hn[3 * r + 1] * s1y * c1y;
                    }
                    double R[9];                                    // inv(T2) * G: x = x_n / s_x + c_x * w
                    for (int c = 0; c < 3; ++c) {
                        R[c] = G[c] / s2x + c2x * G[6 + c];
                        R[3 + c] = G[3 + c] / s2y + c2y * G[6 + c];
                        R[6 + c] = G[6 + c];
                    }
                    solved = R[8] != 0.0 && isfinite(R[8]);
                    if (solved) {
                        const double inv = 1.0 / R[8];
                        for (int i = 0; i < 9; ++i) s_H[i] = R[i] * inv;
                    }
                }
                s_best[0] = solved ? 1 : 0;
            }
            __syncthreads();
            if (s_best[0]) {
                double c = 0;
                for (int i = threadIdx.x; i < M; i += kHomThreads) c += inlier_under(s_H, point(i), thr2) ? 1.0 : 0.0;
                final_count = static_cast<int>(block_sum(c, s_red));
            }
        }
    }
    if (threadIdx.x == 0) {
        a.inliers[p] = final_count;
        if (a.ransac_inliers) a.ransac_inliers[p] = ransac_count;
        if (a.best_hyp) a.best_hyp[p] = best_h;
    }
}

cudaError_t launch_homography_ransac(const HomographyArgs& a, cudaStream_t s) {
    if (a.n_pairs == 0) return cudaSuccess;
    const size_t smem = static_cast<size_t>(kHomSmemPoints) * sizeof(float4) + static_cast<size_t>(a.max_iters) * 4;
    cudaError_t e = cudaFuncSetAttribute(homography_ransac_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    homography_ransac_kernel<<<static_cast<unsigned>(a.n_pairs), kHomThreads, smem, s>>>(a);
    return cudaGetLastError();
}

}  // namespace sfm
