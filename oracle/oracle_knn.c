/*
 * oracle_knn.c — plain-C restatement of the reference's matching stage (CPU, OpenMP).
 *
 * TEST INFRASTRUCTURE ONLY: built into oracle/liboracle_knn.so by oracle/Makefile and used by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference arm as the checker.
 * The product (libsfmmatch.so) never links or loads it.
 *
 * Parity: UNPINNED by the reference's own tests (it has none); pinned instead
 * against cv2 golden vectors in tests/golden (see oracle/oracle_np.py header).
 *
 * Follows:
 *   pair lists      UnorderedFeatureMatchingStrategy.cpp:32-37, VideoFeatureMatchingStrategy.cpp:43-48,
 *                   GridFeatureMatchingStrategy.cpp:48-85
 *   knn k=2         knnMatch call sites Unordered...cpp:51 / Video...cpp:62 / Grid...cpp:105
 *                   (cv::batchDistance semantics: ascending distance, ties -> lowest trainIdx)
 *   ratio filter    UnorderedFeatureMatchingStrategy.cpp:55-65 (double compare, ratio 0.7)
 *   threads         #pragma omp parallel for over pairs, Unordered...cpp:40
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { int32_t queryIdx, trainIdx, imgIdx; float distance; } orc_dmatch;

/* ---- pair selection: write into out[2*cap]; return number of pairs (may exceed cap: call twice) ---- */
long orc_pairs_unordered(int n, int32_t *out, long cap) {
    long k = 0;
    for (int i = 0; i < n; i++)
        for (int j = i + 1; j < n; j++) { if (k < cap) { out[2*k] = i; out[2*k+1] = j; } k++; }
    return k;
}

long orc_pairs_video(int n, int seq, int32_t *out, long cap) {
    if (seq < 2) return -1;
    long k = 0;
    for (int i = 0; i < n; i++)
        for (int j = i + 1; j < n && (j - (i + 1)) < (seq - 1); j++) { if (k < cap) { out[2*k] = i; out[2*k+1] = j; } k++; }
    return k;
}

long orc_pairs_grid(int n, int seq, int rowlen, int32_t *out, long cap) {
    if (seq < 2 || rowlen < 1) return -1;
    int rows = n / rowlen;                       /* integer division inside ceil(): floor */
    long k = 0;
    for (int r = 0; r < rows; r++)
        for (int c = 0; c < rowlen; c++)
            for (int dr = 0; dr < seq; dr++)
                for (int dc = 0; dc < seq; dc++) {
                    int rr = r + dr, cc = c + dc;
                    if ((dr == 0 && dc == 0) || dr + dc >= seq || rr >= rows || cc >= rowlen) continue;
                    if (k < cap) { out[2*k] = r * rowlen + c; out[2*k+1] = rr * rowlen + cc; }
                    k++;
                }
    return k;
}

/* ---- knn k=2 ---- */
static inline void top2_push(int64_t d, int32_t j, int64_t *d0, int32_t *j0, int64_t *d1, int32_t *j1) {
    /* strict '<' while scanning j ascending == lowest index first on ties, rank 1 and rank 2 */
    if (d < *d0) { *d1 = *d0; *j1 = *j0; *d0 = d; *j0 = j; }
    else if (d < *d1) { *d1 = d; *j1 = j; }
}

/* squared L2 on uint8 rows (SIFT descriptors are integer-valued, SURVEY finding 4) */
void orc_knn2_l2_u8(const uint8_t *q, int nq, const uint8_t *t, int nt, int dim,
                    int32_t *idx /*nq*2*/, float *dist /*nq*2*/) {
    for (int i = 0; i < nq; i++) {
        const uint8_t *a = q + (size_t)i * dim;
        int64_t d0 = INT64_MAX, d1 = INT64_MAX; int32_t j0 = -1, j1 = -1;
        for (int j = 0; j < nt; j++) {
            const uint8_t *b = t + (size_t)j * dim;
            int32_t s = 0;
            for (int k = 0; k < dim; k++) { int32_t e = (int32_t)a[k] - (int32_t)b[k]; s += e * e; }
            top2_push(s, j, &d0, &j0, &d1, &j1);
        }
        idx[2*i] = j0; idx[2*i+1] = j1;
        dist[2*i]   = j0 >= 0 ? sqrtf((float)d0) : INFINITY;
        dist[2*i+1] = j1 >= 0 ? sqrtf((float)d1) : INFINITY;
    }
}

void orc_knn2_hamming(const uint8_t *q, int nq, const uint8_t *t, int nt, int nbytes,
                      int32_t *idx, float *dist) {
    for (int i = 0; i < nq; i++) {
        const uint8_t *a = q + (size_t)i * nbytes;
        int64_t d0 = INT64_MAX, d1 = INT64_MAX; int32_t j0 = -1, j1 = -1;
        for (int j = 0; j < nt; j++) {
            const uint8_t *b = t + (size_t)j * nbytes;
            int32_t s = 0;
            int k = 0;
            for (; k + 8 <= nbytes; k += 8) {
                uint64_t x, y; memcpy(&x, a + k, 8); memcpy(&y, b + k, 8);
                s += __builtin_popcountll(x ^ y);
            }
            for (; k < nbytes; k++) s += __builtin_popcount((unsigned)(a[k] ^ b[k]));
            top2_push(s, j, &d0, &j0, &d1, &j1);
        }
        idx[2*i] = j0; idx[2*i+1] = j1;
        dist[2*i]   = j0 >= 0 ? (float)d0 : INFINITY;
        dist[2*i+1] = j1 >= 0 ? (float)d1 : INFINITY;
    }
}

/* Lowe ratio filter: returns number kept, writes DMatch rows ascending queryIdx */
int orc_ratio_filter(const int32_t *idx, const float *dist, int nq, double ratio, orc_dmatch *out) {
    int n = 0;
    for (int i = 0; i < nq; i++) {
        if (idx[2*i] < 0) continue;
        int keep = idx[2*i+1] < 0 ? 1 : ((double)dist[2*i] < (double)dist[2*i+1] * ratio);
        if (keep) { out[n].queryIdx = i; out[n].trainIdx = idx[2*i]; out[n].imgIdx = 0; out[n].distance = dist[2*i]; n++; }
    }
    return n;
}

/* Whole stage over a pair list, threads over pairs like the reference.  norm: 4 = L2 (u8 rows), 6 = Hamming.
 * counts[p] = number of ratio survivors of pair p; if out != NULL, out + p*stride receives them. */
int orc_match_pairs(const uint8_t *const *rows, const int32_t *n_rows, int row_bytes, int norm,
                    const int32_t *pairs, long n_pairs, double ratio, int32_t *counts,
                    orc_dmatch *out, long stride, int threads) {
    int err = 0;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (long p = 0; p < n_pairs; p++) {
        int l = pairs[2*p], r = pairs[2*p+1];
        int nq = n_rows[l], nt = n_rows[r];
        int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)(nq > 0 ? nq : 1));
        float *dist = (float *)malloc(sizeof(float) * 2 * (size_t)(nq > 0 ? nq : 1));
        orc_dmatch *tmp = (orc_dmatch *)malloc(sizeof(orc_dmatch) * (size_t)(nq > 0 ? nq : 1));
        if (!idx || !dist || !tmp) { err = 1; free(idx); free(dist); free(tmp); continue; }
        if (norm == 4) orc_knn2_l2_u8(rows[l], nq, rows[r], nt, row_bytes, idx, dist);
        else orc_knn2_hamming(rows[l], nq, rows[r], nt, row_bytes, idx, dist);
        int n = orc_ratio_filter(idx, dist, nq, ratio, tmp);
        counts[p] = n;
        if (out) memcpy(out + p * stride, tmp, sizeof(orc_dmatch) * (size_t)(n < stride ? n : stride));
        free(idx); free(dist); free(tmp);
    }
    return err;
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
