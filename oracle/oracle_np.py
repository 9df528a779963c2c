"""CPU restatement (numpy) of the reference's feature-matching stage.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (the package
``sfm-mvs-pipeline_b200/`` and the C-ABI library it builds) may import, call or
link anything in ``oracle/``.  Allowed users: ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / reference arm.

Parity status: the reference (``/root/reference``) ships NO tests and NO golden
vectors for this path -> *parity is unpinned by the
reference's own tests*.  The arithmetic of the path lives in an un-vendored
third-party dependency, OpenCV 4.5.1 (``build.sh:105``).  This restatement is
therefore pinned against (a) the one known-answer vector the reference holds
(the grid pairing example in ``GridFeatureMatchingStrategy.h:31-39``) and
(b) outputs of the same OpenCV routines the reference calls, run in the build
container through Python ``cv2`` 4.13.0 and committed as fixtures under
``tests/golden/`` by ``tests/golden/make_golden.py``.

What is restated, with the reference lines each function follows:

* pair selection      UnorderedFeatureMatchingStrategy.cpp:32-37,
                      VideoFeatureMatchingStrategy.cpp:43-48 (+ :31-36),
                      GridFeatureMatchingStrategy.cpp:48-85 (+ :30-42)
* knnMatch(k=2)       call sites Unordered...cpp:51, Video...cpp:62, Grid...cpp:105
                      -> cv::BFMatcher::knnMatchImpl -> cv::batchDistance
                      (OpenCV 4.5.1, modules/core/src/batch_distance.cpp)
* Lowe ratio filter   UnorderedFeatureMatchingStrategy.cpp:55-65
* match() fallback    UnorderedFeatureMatchingStrategy.cpp:66-72 (k=1, keep all)
* distinct filter     SfM.cpp:547-564
* min-match-count     SfM.cpp:566-570
* cross-check         cv::BFMatcher(crossCheck=true) semantics, SURVEY Appendix A.6
                      (never enabled by the reference; optional feature of the build)
"""
from __future__ import annotations

import numpy as np

NORM_L2 = 4        # cv::NORM_L2
NORM_HAMMING = 6   # cv::NORM_HAMMING

DMATCH_DTYPE = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"),
                         ("imgIdx", "<i4"), ("distance", "<f4")])

# OpenCV refuses train sets with >= 2**18 rows (IMGIDX_ONE, matchers.cpp); the
# reference surfaces this as "max 262144" in PhotogrammetrieCli.cpp:430.
MAX_TRAIN_ROWS = 1 << 18


# --------------------------------------------------------------------------
# pair selection
# --------------------------------------------------------------------------
def pairs_unordered(n_shots: int) -> np.ndarray:
    """UnorderedFeatureMatchingStrategy.cpp:32-37 — all (i, j) with i < j, i-major."""
    out = [(i, j) for i in range(n_shots) for j in range(i + 1, n_shots)]
    return np.asarray(out, dtype=np.int32).reshape(-1, 2)


def pairs_video(n_shots: int, sequence_length: int) -> np.ndarray:
    """VideoFeatureMatchingStrategy.cpp:43-48 — i < j and j - (i+1) < sequenceLength-1.

    ``sequence_length < 2`` raises like the reference's setter (:31-36)."""
    if sequence_length < 2:
        raise ValueError("sequence length must not be smaller than 2")
    out = []
    for i in range(n_shots):
        j = i + 1
        while j < n_shots and (j - (i + 1)) < (sequence_length - 1):
            out.append((i, j))
            j += 1
    return np.asarray(out, dtype=np.int32).reshape(-1, 2)


def pairs_grid(n_shots: int, sequence_length: int, row_length: int) -> np.ndarray:
    """GridFeatureMatchingStrategy.cpp:48-85.

    rowCount = n_shots / rowLength is an INTEGER division inside ceil() (:48),
    i.e. floor: shots of a trailing partial row are never paired.
    Emission order: cells row-major; per cell iMRow outer, iMCol inner (:67-83)."""
    if sequence_length < 2:
        raise ValueError("sequence length must not be smaller than 2")
    if row_length < 1:
        raise ValueError("row length must not be smaller than 1")
    row_count = n_shots // row_length
    out = []
    for r in range(row_count):
        for c in range(row_length):
            for dr in range(sequence_length):
                for dc in range(sequence_length):
                    rr, cc = r + dr, c + dc
                    same = (dr == 0 and dc == 0)
                    triangular = (dr + dc) < sequence_length
                    in_grid = rr < row_count and cc < row_length
                    if same or not triangular or not in_grid:
                        continue
                    out.append((r * row_length + c, rr * row_length + cc))
    return np.asarray(out, dtype=np.int32).reshape(-1, 2)


def select_pairs(n_shots: int, feature_sequence: int = 0, feature_gridlength: int = 0) -> np.ndarray:
    """PhotogrammetrieCli.cpp:320-340 — switch values -> strategy."""
    if feature_sequence >= 2:
        if feature_gridlength >= 1:
            return pairs_grid(n_shots, feature_sequence, feature_gridlength)
        return pairs_video(n_shots, feature_sequence)
    return pairs_unordered(n_shots)


# --------------------------------------------------------------------------
# knn (k = 2) — exact restatements of cv::batchDistance as BFMatcher uses it
# --------------------------------------------------------------------------
def _top2_from_int_matrix(d: np.ndarray, chunk_offset: int, best: np.ndarray | None, nt_total: int):
    """Two smallest (value, index) per row with lowest-index-first ties.

    Uses the unique int64 key value * nt_total + index, so ties on value are
    broken by the lower train index for rank 1 AND rank 2."""
    n, m = d.shape
    key = d.astype(np.int64) * np.int64(nt_total) + (np.arange(m, dtype=np.int64) + chunk_offset)[None, :]
    if m >= 2:
        part = np.partition(key, 1, axis=1)[:, :2]
    else:
        part = np.concatenate([key, np.full((n, 1), np.iinfo(np.int64).max)], axis=1)
    if best is not None:
        part = np.concatenate([best, part], axis=1)
        part = np.sort(part, axis=1)[:, :2]
    else:
        part = np.sort(part, axis=1)
    return part


def is_integer_valued_u8(desc: np.ndarray) -> bool:
    if desc.dtype == np.uint8:
        return True
    d = np.asarray(desc)
    return bool(np.all(d == np.rint(d)) and d.min(initial=0) >= 0 and d.max(initial=0) <= 255)


def knn2_sqdist_int(q: np.ndarray, t: np.ndarray, block: int = 2048):
    """Exact integer squared-L2 top-2.  q, t integer-valued (uint8 or float32 holding 0..255).

    Returns (idx[Nq,2] int32, d2[Nq,2] int64); missing neighbours are idx -1, d2 -1."""
    nq, nt = q.shape[0], t.shape[0]
    idx = np.full((nq, 2), -1, np.int32)
    d2 = np.full((nq, 2), -1, np.int64)
    if nq == 0 or nt == 0:
        return idx, d2
    qf = np.ascontiguousarray(q, dtype=np.float64)
    tf = np.ascontiguousarray(t, dtype=np.float64)
    qn = np.einsum("ij,ij->i", qf, qf)
    tn = np.einsum("ij,ij->i", tf, tf)
    big = np.iinfo(np.int64).max
    for q0 in range(0, nq, block):
        q1 = min(nq, q0 + block)
        best = None
        for t0 in range(0, nt, 8192):
            t1 = min(nt, t0 + 8192)
            # float64 GEMM on integer values < 2**53 is exact
            dm = qn[q0:q1, None] + tn[None, t0:t1] - 2.0 * (qf[q0:q1] @ tf[t0:t1].T)
            best = _top2_from_int_matrix(np.rint(dm).astype(np.int64), t0, best, nt)
        have2 = best[:, 1] != big
        idx[q0:q1, 0] = (best[:, 0] % nt).astype(np.int32)
        d2[q0:q1, 0] = best[:, 0] // nt
        idx[q0:q1, 1] = np.where(have2, best[:, 1] % nt, -1).astype(np.int32)
        d2[q0:q1, 1] = np.where(have2, best[:, 1] // nt, -1)
    return idx, d2


def knn2_l2(q: np.ndarray, t: np.ndarray):
    """cv::BFMatcher(NORM_L2).knnMatch(q, t, k=2) as arrays.

    Integer-valued descriptors (what cv::SIFT emits): exact; distance =
    float32(sqrt(float32(sum (a-b)^2))).  Other float data: float32 direct
    differences (indices can differ from OpenCV only on float near-ties)."""
    if q.ndim != 2 or t.ndim != 2 or q.shape[1] != t.shape[1]:
        raise ValueError("descriptor shape mismatch (cv::batchDistance asserts src1.cols == src2.cols)")
    if t.shape[0] >= MAX_TRAIN_ROWS:
        raise ValueError("train rows >= 2**18 (IMGIDX_ONE)")
    if is_integer_valued_u8(q) and is_integer_valued_u8(t):
        idx, d2 = knn2_sqdist_int(q, t)
        dist = np.sqrt(np.maximum(d2, 0).astype(np.float32)).astype(np.float32)
        dist[idx < 0] = np.float32(np.inf)
        return idx, dist
    return _knn2_l2_float(q, t)


def _knn2_l2_float(q, t, block=256):
    nq, nt = q.shape[0], t.shape[0]
    idx = np.full((nq, 2), -1, np.int32)
    dist = np.full((nq, 2), np.inf, np.float32)
    if nq == 0 or nt == 0:
        return idx, dist
    q32 = np.asarray(q, np.float32)
    t32 = np.asarray(t, np.float32)
    for q0 in range(0, nq, block):
        q1 = min(nq, q0 + block)
        diff = q32[q0:q1, None, :] - t32[None, :, :]
        d2 = np.einsum("ijk,ijk->ij", diff, diff).astype(np.float32)
        order = np.argsort(d2, axis=1, kind="stable")[:, :2]
        k = order.shape[1]
        idx[q0:q1, :k] = order
        dist[q0:q1, :k] = np.sqrt(np.take_along_axis(d2, order, axis=1))
    return idx, dist


def knn2_hamming(q: np.ndarray, t: np.ndarray):
    """cv::BFMatcher(NORM_HAMMING).knnMatch(q, t, k=2) as arrays.

    Distance = integer popcount(a xor b) converted to float; ties -> lowest trainIdx."""
    if q.dtype != np.uint8 or t.dtype != np.uint8:
        raise ValueError("NORM_HAMMING needs CV_8U descriptors (cv::error otherwise)")
    if q.ndim != 2 or t.ndim != 2 or q.shape[1] != t.shape[1]:
        raise ValueError("descriptor shape mismatch")
    if t.shape[0] >= MAX_TRAIN_ROWS:
        raise ValueError("train rows >= 2**18 (IMGIDX_ONE)")
    nq, nt = q.shape[0], t.shape[0]
    idx = np.full((nq, 2), -1, np.int32)
    dist = np.full((nq, 2), np.inf, np.float32)
    if nq == 0 or nt == 0:
        return idx, dist
    qb = np.unpackbits(q, axis=1).astype(np.float32)
    tb = np.unpackbits(t, axis=1).astype(np.float32)
    qc, tc = qb.sum(1), tb.sum(1)
    big = np.iinfo(np.int64).max
    for q0 in range(0, nq, 4096):
        q1 = min(nq, q0 + 4096)
        best = None
        for t0 in range(0, nt, 8192):
            t1 = min(nt, t0 + 8192)
            # popcount(a^b) = |a| + |b| - 2 a.b ; all values <= 256*... exact in float32
            dm = qc[q0:q1, None] + tc[None, t0:t1] - 2.0 * (qb[q0:q1] @ tb[t0:t1].T)
            best = _top2_from_int_matrix(np.rint(dm).astype(np.int64), t0, best, nt)
        have2 = best[:, 1] != big
        idx[q0:q1, 0] = (best[:, 0] % nt).astype(np.int32)
        dist[q0:q1, 0] = (best[:, 0] // nt).astype(np.float32)
        idx[q0:q1, 1] = np.where(have2, best[:, 1] % nt, -1).astype(np.int32)
        dist[q0:q1, 1] = np.where(have2, (best[:, 1] // nt).astype(np.float32), np.inf)
    return idx, dist


def knn2(q, t, norm):
    if norm == NORM_L2:
        return knn2_l2(q, t)
    if norm == NORM_HAMMING:
        return knn2_hamming(q, t)
    raise ValueError(f"unsupported norm {norm}")


# --------------------------------------------------------------------------
# filters
# --------------------------------------------------------------------------
def ratio_filter(idx: np.ndarray, dist: np.ndarray, ratio: float = 0.7) -> np.ndarray:
    """UnorderedFeatureMatchingStrategy.cpp:55-65 as arrays.

    keep m[0] iff (double)m[0].distance < (double)m[1].distance * ratio;
    rows with a single neighbour are kept unconditionally (:62-64); rows with no
    neighbour are skipped (documented deviation from the reference's UB, App. C)."""
    d0 = dist[:, 0].astype(np.float64)
    d1 = dist[:, 1].astype(np.float64)
    has1 = idx[:, 0] >= 0
    has2 = idx[:, 1] >= 0
    keep = has1 & (~has2 | (d0 < d1 * np.float64(ratio)))
    rows = np.nonzero(keep)[0]
    out = np.zeros(rows.size, DMATCH_DTYPE)
    out["queryIdx"] = rows
    out["trainIdx"] = idx[rows, 0]
    out["imgIdx"] = 0
    out["distance"] = dist[rows, 0]
    return out


def match_k1(idx: np.ndarray, dist: np.ndarray) -> np.ndarray:
    """DescriptorMatcher::match() (k=1, keep everything) — the catch-branch at
    UnorderedFeatureMatchingStrategy.cpp:66-72."""
    rows = np.nonzero(idx[:, 0] >= 0)[0]
    out = np.zeros(rows.size, DMATCH_DTYPE)
    out["queryIdx"] = rows
    out["trainIdx"] = idx[rows, 0]
    out["distance"] = dist[rows, 0]
    return out


def cross_check_filter(idx_qt: np.ndarray, dist_qt: np.ndarray, idx_tq: np.ndarray) -> np.ndarray:
    """cv::BFMatcher(crossCheck=true): keep q iff NN(q)=t and NN(t)=q (App. A.6).

    idx_qt: query->train top-1 (column 0 used); idx_tq: train->query top-1."""
    q = np.arange(idx_qt.shape[0])
    t = idx_qt[:, 0]
    ok = t >= 0
    mutual = np.zeros_like(ok)
    mutual[ok] = idx_tq[t[ok], 0] == q[ok]
    rows = np.nonzero(mutual)[0]
    out = np.zeros(rows.size, DMATCH_DTYPE)
    out["queryIdx"] = rows
    out["trainIdx"] = t[rows]
    out["distance"] = dist_qt[rows, 0]
    return out


def distinct_filter(m: np.ndarray) -> np.ndarray:
    """SfM.cpp:547-564 — drop every match whose trainIdx is shared with another
    match of a different queryIdx."""
    if m.size == 0:
        return m
    _, inv, cnt = np.unique(m["trainIdx"], return_inverse=True, return_counts=True)
    return m[cnt[inv] == 1]


def match_pairs(bank, pairs, norm, ratio=0.7, cross_check=False, distinct=False, min_match_count=0):
    """The whole stage for a pair list: the strategy loop (e.g. Unordered...cpp:40-91)
    followed by SfM::calculateShotMatches' post-filters (SfM.cpp:547-570).

    Returns a list with one DMATCH array per input pair, in pair order; pairs
    dropped by ``min_match_count`` yield ``None`` (the reference erases them)."""
    out = []
    for (l, r) in np.asarray(pairs).reshape(-1, 2):
        q, t = bank[int(l)], bank[int(r)]
        if q.shape[0] == 0 or t.shape[0] == 0:
            m = np.zeros(0, DMATCH_DTYPE)       # documented deviation (App. C)
        else:
            idx, dist = knn2(q, t, norm)
            if cross_check:
                idx_tq, _ = knn2(t, q, norm)
                m = cross_check_filter(idx, dist, idx_tq)
            else:
                m = ratio_filter(idx, dist, ratio)
        if distinct:
            m = distinct_filter(m)
        if m.size < min_match_count:
            out.append(None)
        else:
            out.append(m)
    return out


def dmatch_equal(a: np.ndarray, b: np.ndarray, ignore_distance: bool = False) -> bool:
    """OpenCvUtils::equals / equalsIgnoreDistance (OpenCvUtils.cpp:85-91) lifted to lists."""
    if a.shape != b.shape:
        return False
    same = (np.array_equal(a["queryIdx"], b["queryIdx"]) and
            np.array_equal(a["trainIdx"], b["trainIdx"]) and
            np.array_equal(a["imgIdx"], b["imgIdx"]))
    if ignore_distance:
        return same
    return same and np.array_equal(a["distance"].view(np.uint32), b["distance"].view(np.uint32))
