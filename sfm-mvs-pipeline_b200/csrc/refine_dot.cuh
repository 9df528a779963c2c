// refine_dot.cuh — exact re-rank of one query row from the candidate chunks of the value-only tcgen05 kernel (warp-level
// device code).  Shared by the post pass (post.cu: refine_dot_rows_kernel, refine_value_rows_kernel, brute_force_rows_kernel)
// and by the refine warps INSIDE knn2_l2_u8_tcv_kernel (knn_l2_tcv.cu), which re-rank the rows that survive the fused
// ratio bound while the tensor pipe works on the next tiles.
#pragma once
#include <climits>

#include "common.cuh"

namespace sfm {

// what refine_dot_row needs besides the row itself
struct RefineCtx {
    const uint8_t* bank;         // u8 bank, 128-byte rows
    const int32_t* norm2;
    Top2* top2;                  // staging rows: the final record of the row is written here
    unsigned long long* stats;   // [0] rows re-ranked, [1] rows brute-forced
    int32_t* bf_list;            // rows whose answer needs the whole train image (finished by brute_force_rows_kernel)
    int* bf_count;
    int chunk_rows;              // train rows per candidate chunk (32 or 64)
    int all_rows;
    double ratio;
};

__device__ __forceinline__ void warp_chunk_candidates(const uint8_t* __restrict__ bank, const int32_t* __restrict__ norm2,
                                                      const uint4 (&q)[8], int na, int tr0, int ntr, int chunk, int lane,
                                                      long long& a1, long long& a2) {
    const int j = chunk * 32 + lane;
    const uint4* tv = reinterpret_cast<const uint4*>(bank + (static_cast<size_t>(tr0) + j) * 128);
    uint32_t dot = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint4 y = __ldg(tv + i);
        dot = __dp4a(q[i].x, y.x, dot); dot = __dp4a(q[i].y, y.y, dot);
        dot = __dp4a(q[i].z, y.z, dot); dot = __dp4a(q[i].w, y.w, dot);
    }
    if (j < ntr) {
        const int32_t d = na + norm2[tr0 + j] - 2 * static_cast<int32_t>(dot);
        const long long key = (static_cast<long long>(d) << 32) | static_cast<unsigned>(j);
        a2 = min(a2, max(a1, key));
        a1 = min(a1, key);
    }
}

// exact top-2 over the WHOLE train image for one query row, one warp: four 32-row groups per step so that their loads
// are in flight together (a warp that has to brute-force a row is the tail of the refine kernels)
__device__ __forceinline__ void warp_brute_force(const uint8_t* __restrict__ bank, const int32_t* __restrict__ norm2,
                                                 const uint4 (&q)[8], int na, int tr0, int ntr, int lane, long long& a1,
                                                 long long& a2) {
    const int groups = (ntr + 31) / 32;
    int c = 0;
    for (; c + 4 <= groups; c += 4) {
        uint32_t dot[4] = {0, 0, 0, 0};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            uint4 y[4];
#pragma unroll
            for (int g = 0; g < 4; ++g)
                y[g] = __ldg(reinterpret_cast<const uint4*>(bank + (static_cast<size_t>(tr0) + (c + g) * 32 + lane) * 128) + i);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                dot[g] = __dp4a(q[i].x, y[g].x, dot[g]); dot[g] = __dp4a(q[i].y, y[g].y, dot[g]);
                dot[g] = __dp4a(q[i].z, y[g].z, dot[g]); dot[g] = __dp4a(q[i].w, y[g].w, dot[g]);
            }
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const int j = (c + g) * 32 + lane;
            if (j < ntr) {
                const int32_t d = na + norm2[tr0 + j] - 2 * static_cast<int32_t>(dot[g]);
                const long long key = (static_cast<long long>(d) << 32) | static_cast<unsigned>(j);
                a2 = min(a2, max(a1, key));
                a1 = min(a1, key);
            }
        }
    }
    for (; c < groups; ++c) warp_chunk_candidates(bank, norm2, q, na, tr0, ntr, c, lane, a1, a2);
}

// one warp, one query row that survived the quick reject: stages (2a), (2b), (3) of the comment above.  All arguments are
// warp-uniform; lane 0 writes the row (or queues it for brute_force_rows_kernel).
__device__ __forceinline__ void refine_dot_row(const RefineCtx& a, int64_t srow, int lane, const Top2 t, int rv5, int rna,
                                               int qrow, int tr0, int ntr, int nbmin, int nbmax) {
    const float inf = __int_as_float(0x7f800000);
    const int i0 = t.i0, i1 = t.i1;
    uint4 q[8];
    const uint4* qv = reinterpret_cast<const uint4*>(a.bank + static_cast<size_t>(qrow) * 128);
#pragma unroll
    for (int i = 0; i < 8; ++i) q[i] = __ldg(qv + i);
    const int cand[4] = {i0 & 0xFFFF, (i0 >> 16) & 0xFFFF, i1 & 0xFFFF, (i1 >> 16) & 0xFFFF};
    const int rV2 = __float_as_int(t.d1);
    const int sub = a.chunk_rows / 32;                          // a candidate chunk = sub groups of 32 train rows
    long long a1 = LLONG_MAX, a2 = LLONG_MAX;
    // ---- stage A: the best chunk alone.  Every row outside it has a.b <= V2, i.e. d^2 >= |a|^2 + N- - 2 V2 =: lb2,
    // and chunk 2 holds a real row with d^2 <= |a|^2 + N+ - 2 V2 =: ub2.  A planted match has e0 far below lb2: the
    // nearest neighbour is certified and d1^2 lies in [min(e1', lb2), min(e1', ub2)] (e1' = second best inside the
    // chunk) -- if the ratio test agrees at both ends the row is finished after 32-64 exact distances.
    if (!a.all_rows && cand[0] != 0xFFFF && rV2 > 0) {
        for (int h = 0; h < sub; ++h)
            warp_chunk_candidates(a.bank, a.norm2, q, rna, tr0, ntr, cand[0] * sub + h, lane, a1, a2);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const long long b1 = __shfl_xor_sync(0xffffffffu, a1, off), b2 = __shfl_xor_sync(0xffffffffu, a2, off);
            a2 = min(max(a1, b1), min(a2, b2));
            a1 = min(a1, b1);
        }
        if (a1 != LLONG_MAX) {
            const long long e0 = a1 >> 32;
            const long long lb2 = static_cast<long long>(rna) + nbmin - 2ll * rV2;
            const long long ub2 = static_cast<long long>(rna) + nbmax - 2ll * rV2;
            if (e0 < lb2) {
                const long long e1 = a2 != LLONG_MAX ? (a2 >> 32) : LLONG_MAX;
                const long long lo1 = min(e1, lb2), hi1 = min(e1, ub2);
                const float s0 = __fsqrt_rn(static_cast<float>(static_cast<int32_t>(e0)));
                const bool pass_lo = static_cast<double>(s0) < static_cast<double>(__fsqrt_rn(static_cast<float>(static_cast<int32_t>(lo1)))) * a.ratio;
                const bool pass_hi = static_cast<double>(s0) < static_cast<double>(__fsqrt_rn(static_cast<float>(static_cast<int32_t>(hi1)))) * a.ratio;
                if (pass_lo == pass_hi) {
                    if (lane == 0) {
                        Top2 o;
                        o.i0 = static_cast<int>(a1 & 0xFFFFFFFFll); o.d0 = static_cast<float>(static_cast<int32_t>(e0));
                        // the second neighbour itself is not reported by the ratio-filtered stage: any index >= 0 marks
                        // "a second neighbour exists", d1 = the end of the interval that was tested
                        o.i1 = a2 != LLONG_MAX ? static_cast<int>(a2 & 0xFFFFFFFFll) : ntr;
                        o.d1 = static_cast<float>(static_cast<int32_t>(pass_lo ? lo1 : hi1));
                        a.top2[srow] = o;
                    }
                    return;
                }
            }
        }
        a1 = LLONG_MAX; a2 = LLONG_MAX;
    }
    // ---- stage B: all candidate chunks
    int covered = 0;                                                // real train rows inside the candidate chunks
    for (int k = 0; k < 4; ++k)
        if (cand[k] != 0xFFFF) {
            for (int h = 0; h < sub; ++h)
                warp_chunk_candidates(a.bank, a.norm2, q, rna, tr0, ntr, cand[k] * sub + h, lane, a1, a2);
            covered += max(0, min(a.chunk_rows, ntr - a.chunk_rows * cand[k]));
        }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const long long b1 = __shfl_xor_sync(0xffffffffu, a1, off), b2 = __shfl_xor_sync(0xffffffffu, a2, off);
        a2 = min(max(a1, b1), min(a2, b2));
        a1 = min(a1, b1);
    }
    // decide (all lanes hold the same a1, a2)
    bool done = false;
    float out_d1 = inf;
    int out_i1 = -1;
    const bool outside = covered < ntr;                             // train rows exist outside the candidate chunks
    const long long lbo = outside ? static_cast<long long>(rna) + nbmin - 2ll * rv5 : LLONG_MAX;
    if (a1 != LLONG_MAX) {
        const long long e0 = a1 >> 32;
        if (!outside) {                                             // the chunks cover the whole train image: exact
            done = true;
            if (a2 != LLONG_MAX) { out_i1 = static_cast<int>(a2 & 0xFFFFFFFFll); out_d1 = static_cast<float>(static_cast<int32_t>(a2 >> 32)); }
        } else if (e0 < lbo && !a.all_rows) {
            const float s0 = __fsqrt_rn(static_cast<float>(static_cast<int32_t>(e0)));
            const long long e1 = a2 != LLONG_MAX ? (a2 >> 32) : LLONG_MAX;
            const long long lb1 = min(e1, lbo);                     // <= true d1^2 <= e1 (e1 = none: only the bound)
            const bool pass_lo = static_cast<double>(s0) < static_cast<double>(__fsqrt_rn(static_cast<float>(static_cast<int32_t>(lb1)))) * a.ratio;
            if (e1 == LLONG_MAX) {
                // a second neighbour exists outside the chunks (outside == true): only 'pass at the lower end' decides
                if (pass_lo) { done = true; out_i1 = ntr; out_d1 = static_cast<float>(static_cast<int32_t>(lb1)); }
            } else {
                const bool pass_hi = static_cast<double>(s0) < static_cast<double>(__fsqrt_rn(static_cast<float>(static_cast<int32_t>(e1)))) * a.ratio;
                if (pass_lo == pass_hi) {
                    done = true;
                    out_i1 = static_cast<int>(a2 & 0xFFFFFFFFll);
                    out_d1 = static_cast<float>(static_cast<int32_t>(pass_lo ? lb1 : e1));
                }
            }
        }
    }
    // not certified, but perhaps certainly failing: the true d0^2 is >= min(e0, lbo) and the true d1^2 is <= e1 (the
    // second best candidate is a real row), so  sqrtf(min(e0, lbo)) >= ratio * sqrtf(e1)  means the ratio test fails
    // whatever lies outside the candidates -- the usual case of a row without a planted match
    bool rejected = false;
    if (!done && !a.all_rows && a1 != LLONG_MAX && a2 != LLONG_MAX && outside) {
        const long long lb0 = max(0ll, min(a1 >> 32, lbo)), e1 = a2 >> 32;
        const float s0 = __fsqrt_rn(static_cast<float>(static_cast<int32_t>(lb0)));
        const float s1 = __fsqrt_rn(static_cast<float>(static_cast<int32_t>(e1)));
        if (!(static_cast<double>(s0) < static_cast<double>(s1) * a.ratio)) { done = true; rejected = true; }
    }
    if (!done) {
        // exact brute force over the whole train image: queued for brute_force_rows_kernel (a whole CTA per row)
        if (lane == 0) {
            atomicAdd(a.stats + 1, 1ull);
            a.bf_list[atomicAdd(a.bf_count, 1)] = static_cast<int32_t>(srow);
        }
        rejected = true;                                            // placeholder until that kernel writes the row
    }
    if (lane == 0) {
        Top2 o;
        o.i0 = -1; o.i1 = -1; o.d0 = inf; o.d1 = inf;
        if (!rejected) {
            if (a1 != LLONG_MAX) { o.i0 = static_cast<int>(a1 & 0xFFFFFFFFll); o.d0 = static_cast<float>(static_cast<int32_t>(a1 >> 32)); }
            o.i1 = out_i1; o.d1 = out_d1;
        }
        a.top2[srow] = o;
    }
}


}  // namespace sfm
