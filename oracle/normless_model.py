"""ORACLE-SIDE MODEL (test infrastructure): the decision logic of the norm-less value-only path restated in numpy.

csrc/knn_l2_tcv.cu<kNorm = false> ranks the 32/64-row chunks of a train image by the maximum raw dot product and reports
the four best chunks with a positive maximum, V1, V2 and V5; refine_dot_* (csrc/post.cu) turns that into the answer of
cv::BFMatcher + Lowe ratio test using only the train image's norm range [N-, N+].  This file replays exactly those
decisions on the CPU (exact integer dot products instead of the tensor cores) so that the BOUNDS and CERTIFICATES can be
property-tested against the plain oracle on inputs the GPU tests rarely hit (huge norm spreads, zero rows, duplicates,
tiny images).  It is not used by the product and says nothing about the CUDA code itself.
"""
from __future__ import annotations

import numpy as np

from .oracle_np import DMATCH_DTYPE


def _f32sqrt(x):
    return np.sqrt(np.float32(x), dtype=np.float32)


def _ratio_pass(d0_sq, d1_sq, ratio):
    """float(sqrtf(d0^2)) < float(sqrtf(d1^2)) * ratio in double, as keep_basic evaluates it."""
    return float(_f32sqrt(d0_sq)) < float(_f32sqrt(d1_sq)) * float(ratio)


def match_pair_normless(q: np.ndarray, t: np.ndarray, ratio: float = 0.7, chunk: int = 64, stats: dict | None = None):
    """Ratio-filtered DMatch list of one pair, decided the way the norm-less GPU path decides it."""
    nq, nt = q.shape[0], t.shape[0]
    out = []
    if nq == 0 or nt == 0:
        return np.zeros(0, DMATCH_DTYPE)
    q64, t64 = q.astype(np.int64), t.astype(np.int64)
    na_all, nb = (q64 * q64).sum(1), (t64 * t64).sum(1)
    nbmin, nbmax = int(nb.min()), int(nb.max())
    ab_all = q64 @ t64.T
    pad = (-nt) % 256                                   # images are zero-padded to 256 rows: padded rows give a.b = 0
    n_chunks = (nt + pad) // chunk
    st = stats if stats is not None else {}
    for k in ("rejected", "stage_a", "stage_b", "proved_fail", "brute"):
        st.setdefault(k, 0)
    for r in range(nq):
        na = int(na_all[r])
        ab = np.concatenate([ab_all[r], np.zeros(pad, np.int64)])
        cmax = ab.reshape(n_chunks, chunk).max(1)
        # chunks ordered by (maximum desc, chunk index asc); only positive maxima are candidates
        order = sorted(range(n_chunks), key=lambda c: (-int(cmax[c]), c))
        vals = [int(cmax[c]) for c in order]
        cand = [c for c in order[:4] if cmax[c] > 0]
        V1 = vals[0] if vals[0] > 0 else -1
        V2 = vals[1] if len(vals) > 1 and vals[1] > 0 else -1
        V5 = vals[4] if len(vals) > 4 and vals[4] > 0 else 0
        d2 = na + nb - 2 * ab_all[r]                    # exact squared distances (what __dp4a recomputes)

        def top2_in(rows):
            rows = np.asarray([j for j in rows if j < nt], np.int64)
            if len(rows) == 0:
                return None, None
            keys = sorted((int(d2[j]), int(j)) for j in rows)
            return keys[0], (keys[1] if len(keys) > 1 else None)

        def brute():
            st["brute"] += 1
            return top2_in(range(nt))

        # (1) quick reject
        if V2 > 0:
            lo0 = max(0, na + nbmin - 2 * V1)
            hi1 = max(0, na + nbmax - 2 * V2)
            if not _ratio_pass(lo0, hi1, ratio):
                st["rejected"] += 1
                continue
        e0 = e1 = None
        decided = False
        keep = False
        d1_used = None
        # (2a) stage A: best chunk alone
        if cand and V2 > 0:
            a0, a1 = top2_in(range(cand[0] * chunk, cand[0] * chunk + chunk))
            if a0 is not None:
                lb2, ub2 = na + nbmin - 2 * V2, na + nbmax - 2 * V2
                if a0[0] < lb2:
                    e1v = a1[0] if a1 is not None else None
                    lo1 = lb2 if e1v is None else min(e1v, lb2)
                    hi1 = ub2 if e1v is None else min(e1v, ub2)
                    p_lo, p_hi = _ratio_pass(a0[0], lo1, ratio), _ratio_pass(a0[0], hi1, ratio)
                    if p_lo == p_hi:
                        st["stage_a"] += 1
                        decided, keep, e0 = True, p_lo, a0
        if not decided:
            # (2b) stage B: all candidate chunks
            rows = [j for c in cand for j in range(c * chunk, c * chunk + chunk)]
            covered = sum(max(0, min(chunk, nt - chunk * c)) for c in cand)
            b0, b1 = top2_in(rows)
            outside = covered < nt
            lbo = na + nbmin - 2 * V5 if outside else None
            if b0 is not None:
                if not outside:
                    decided, e0, e1 = True, b0, b1
                    keep = True if b1 is None else _ratio_pass(b0[0], b1[0], ratio)
                    st["stage_b"] += 1
                elif b0[0] < lbo:
                    lb1 = lbo if b1 is None else min(b1[0], lbo)
                    p_lo = _ratio_pass(b0[0], lb1, ratio)
                    if b1 is None:
                        if p_lo:
                            decided, keep, e0 = True, True, b0
                            st["stage_b"] += 1
                    else:
                        p_hi = _ratio_pass(b0[0], b1[0], ratio)
                        if p_lo == p_hi:
                            decided, keep, e0 = True, p_lo, b0
                            st["stage_b"] += 1
            if not decided and b0 is not None and b1 is not None and outside:
                lb0 = max(0, min(b0[0], lbo))
                if not _ratio_pass(lb0, b1[0], ratio):
                    decided, keep = True, False
                    st["proved_fail"] += 1
            if not decided:
                e0, e1 = brute()
                keep = e0 is not None and (e1 is None or _ratio_pass(e0[0], e1[0], ratio))
        if keep and e0 is not None:
            out.append((r, e0[1], 0, float(_f32sqrt(e0[0]))))
    return np.array(out, DMATCH_DTYPE) if out else np.zeros(0, DMATCH_DTYPE)


def match_pair_value(q: np.ndarray, t: np.ndarray, ratio: float = 0.7, chunk: int = 64, stats: dict | None = None):
    """Same for the variant WITH the norm K-step (knn2_l2_u8_tcv_kernel<kNorm = true> + refine_value_kernel): chunks are
    ranked by D = a.b - (|b|^2 >> 1) (|a-b|^2 = |a|^2 - 2 D + (|b|^2 & 1)); the kernel reports the two best chunks, a third
    if its maximum ties the second, and 'ambiguous' if a fourth ties too (-> brute force)."""
    nq, nt = q.shape[0], t.shape[0]
    out = []
    if nq == 0 or nt == 0:
        return np.zeros(0, DMATCH_DTYPE)
    q64, t64 = q.astype(np.int64), t.astype(np.int64)
    na_all, nb = (q64 * q64).sum(1), (t64 * t64).sum(1)
    ab_all = q64 @ t64.T
    pad = (-nt) % 256
    n_chunks = (nt + pad) // chunk
    st = stats if stats is not None else {}
    for k in ("rejected", "chunks", "brute"):
        st.setdefault(k, 0)
    NEG = -(1 << 40)                                      # padded rows: below every valid D
    for r in range(nq):
        na = int(na_all[r])
        D = np.concatenate([ab_all[r] - (nb >> 1), np.full(pad, NEG, np.int64)])
        cmax = D.reshape(n_chunks, chunk).max(1)
        order = [c for c in sorted(range(n_chunks), key=lambda c: (-int(cmax[c]), c)) if cmax[c] > NEG]
        d2 = na + nb - 2 * ab_all[r]

        def top2_in(rows):
            keys = sorted((int(d2[j]), int(j)) for j in rows if j < nt)
            return (keys[0] if keys else None), (keys[1] if len(keys) > 1 else None)

        if not order:
            continue
        c1 = order[0]
        c2 = order[1] if len(order) > 1 else None
        c3 = order[2] if len(order) > 2 and cmax[order[2]] == cmax[order[1]] else None
        ambiguous = c3 is not None and len(order) > 3 and cmax[order[3]] == cmax[order[1]]
        if c2 is not None and not ambiguous:
            lo0 = max(0, na - 2 * int(cmax[c1]))
            hi1 = na - 2 * int(cmax[c2]) + 1
            if not _ratio_pass(lo0, hi1, ratio):
                st["rejected"] += 1
                continue
        if ambiguous:
            st["brute"] += 1
            e0, e1 = top2_in(range(nt))
        else:
            st["chunks"] += 1
            rows = [j for c in (c1, c2, c3) if c is not None for j in range(c * chunk, c * chunk + chunk)]
            e0, e1 = top2_in(rows)
        if e0 is not None and (e1 is None or _ratio_pass(e0[0], e1[0], ratio)):
            out.append((r, e0[1], 0, float(_f32sqrt(e0[0]))))
    return np.array(out, DMATCH_DTYPE) if out else np.zeros(0, DMATCH_DTYPE)
