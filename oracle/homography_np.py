"""ORACLE (test infrastructure, never imported by the product): numpy restatement of the homography stage.

Reference behaviour: SfM::calculateHomography (SfM.cpp:599-637) =
    cv::findHomography(left pts, right pts, cv::RANSAC, thr, mask);  ratio = countNonZero(mask) / matches.size()
with left <-> queryIdx, right <-> trainIdx (Scene.cpp:95-110 alignFeatures), pairs with < 4 matches skipped.
cv::findHomography lives in OpenCV 4.5.1 calib3d (un-vendored, build.sh:105): RANSACPointSetRegistrator::run keeps
the consensus set of the best 4-point model; its random minimal sets cannot be reproduced outside OpenCV, so parity
with it is statistical (cv2_ref.find_homography_ratio gives the golden numbers, tests state the tolerance).

ransac_inliers() below restates the algorithm of csrc/homography.cu operation for operation (same counter-based
generator, same degeneracy test, same closed-form 4-point model, float64 without FMA), so the CUDA path can be
checked hypothesis by hypothesis.
"""
from __future__ import annotations

import numpy as np

U64 = np.uint64
EPS = 1.1920928955078125e-7        # FLT_EPSILON, as in cv::HomographyEstimatorCallback::checkSubset


def splitmix64(x):
    with np.errstate(over="ignore"):
        x = (np.asarray(x, U64) + U64(0x9E3779B97F4A7C15)).astype(U64)
        x = ((x ^ (x >> U64(30))) * U64(0xBF58476D1CE4E5B9)).astype(U64)
        x = ((x ^ (x >> U64(27))) * U64(0x94D049BB133111EB)).astype(U64)
        return x ^ (x >> U64(31))


def minimal_sets(seed: int, pair: int, n_hyp: int, m: int):
    """idx[n_hyp, 4], ok[n_hyp]: four distinct match indices per hypothesis (at most 16 draws)."""
    h = np.arange(n_hyp, dtype=U64)
    with np.errstate(over="ignore"):
        key = splitmix64(splitmix64(splitmix64(U64(seed)) ^ U64(pair)) + h)
    idx = np.full((n_hyp, 4), -1, np.int64)
    got = np.zeros(n_hyp, np.int64)
    for draw in range(16):
        with np.errstate(over="ignore"):
            r = splitmix64(key + U64(draw))
        cand = (((r >> U64(32)) * U64(m)) >> U64(32)).astype(np.int64)
        active = got < 4
        dup = (idx == cand[:, None]).any(1)
        take = active & ~dup
        rows = np.nonzero(take)[0]
        idx[rows, got[rows]] = cand[rows]
        got[rows] += 1
    return idx, got == 4


def _cross3(ax, ay, bx, by, cx, cy):
    return (bx - ax) * (cy - ay) - (by - ay) * (cx - ax)


def _subset_ok(x1, y1, x2, y2):
    ok = np.ones(x1.shape[0], bool)
    negative = np.zeros(x1.shape[0], np.int64)
    for skip in range(4):
        a, b, c = [j for j in range(4) if j != skip]
        c1 = _cross3(x1[:, a], y1[:, a], x1[:, b], y1[:, b], x1[:, c], y1[:, c])
        c2 = _cross3(x2[:, a], y2[:, a], x2[:, b], y2[:, b], x2[:, c], y2[:, c])
        s1 = np.abs(x1[:, b] - x1[:, a]) + np.abs(y1[:, b] - y1[:, a]) + np.abs(x1[:, c] - x1[:, a]) + np.abs(y1[:, c] - y1[:, a])
        s2 = np.abs(x2[:, b] - x2[:, a]) + np.abs(y2[:, b] - y2[:, a]) + np.abs(x2[:, c] - x2[:, a]) + np.abs(y2[:, c] - y2[:, a])
        ok &= ~((np.abs(c1) <= EPS * s1) | (np.abs(c2) <= EPS * s2))
        negative += (c1 * c2 < 0.0)
    return ok & ((negative == 0) | (negative == 4))


def _square_to_quad(x, y):
    dx1, dx2, sx = x[:, 1] - x[:, 2], x[:, 3] - x[:, 2], x[:, 0] - x[:, 1] + x[:, 2] - x[:, 3]
    dy1, dy2, sy = y[:, 1] - y[:, 2], y[:, 3] - y[:, 2], y[:, 0] - y[:, 1] + y[:, 2] - y[:, 3]
    den = dx1 * dy2 - dy1 * dx2
    ok = den != 0.0
    with np.errstate(all="ignore"):
        g = (sx * dy2 - sy * dx2) / den
        h = (dx1 * sy - dy1 * sx) / den
        S = np.stack([x[:, 1] - x[:, 0] + g * x[:, 1], x[:, 3] - x[:, 0] + h * x[:, 3], x[:, 0],
                      y[:, 1] - y[:, 0] + g * y[:, 1], y[:, 3] - y[:, 0] + h * y[:, 3], y[:, 0],
                      g, h, np.ones_like(g)], 1)
    return S, ok


def models(pts: np.ndarray, idx: np.ndarray, ok: np.ndarray):
    """pts[M, 4] float32 (x1, y1, x2, y2) -> H[n_hyp, 9] (H[8] = 1) and validity."""
    safe = np.where(idx < 0, 0, idx)
    p = pts[safe].astype(np.float64)                       # [n_hyp, 4 points, 4 coords]
    x1, y1, x2, y2 = p[..., 0], p[..., 1], p[..., 2], p[..., 3]
    ok = ok & _subset_ok(x1, y1, x2, y2)
    S1, ok1 = _square_to_quad(x1, y1)
    S2, ok2 = _square_to_quad(x2, y2)
    ok = ok & ok1 & ok2
    with np.errstate(all="ignore"):
        A = np.stack([S1[:, 4] * S1[:, 8] - S1[:, 5] * S1[:, 7], S1[:, 2] * S1[:, 7] - S1[:, 1] * S1[:, 8], S1[:, 1] * S1[:, 5] - S1[:, 2] * S1[:, 4],
                      S1[:, 5] * S1[:, 6] - S1[:, 3] * S1[:, 8], S1[:, 0] * S1[:, 8] - S1[:, 2] * S1[:, 6], S1[:, 2] * S1[:, 3] - S1[:, 0] * S1[:, 5],
                      S1[:, 3] * S1[:, 7] - S1[:, 4] * S1[:, 6], S1[:, 1] * S1[:, 6] - S1[:, 0] * S1[:, 7], S1[:, 0] * S1[:, 4] - S1[:, 1] * S1[:, 3]], 1)
        H = np.empty_like(A)
        for r in range(3):
            for c in range(3):
                H[:, 3 * r + c] = S2[:, 3 * r] * A[:, c] + S2[:, 3 * r + 1] * A[:, 3 + c] + S2[:, 3 * r + 2] * A[:, 6 + c]
        ok = ok & (H[:, 8] != 0.0) & np.isfinite(H[:, 8])
        inv = 1.0 / H[:, 8]
        H = H * inv[:, None]
    return H, ok


def consensus(pts: np.ndarray, H: np.ndarray, ok: np.ndarray, thr: float):
    X, Y = pts[:, 0].astype(np.float64)[None, :], pts[:, 1].astype(np.float64)[None, :]
    with np.errstate(all="ignore"):
        w = H[:, 6:7] * X + H[:, 7:8] * Y + 1.0
        ww = 1.0 / w
        dx = (H[:, 0:1] * X + H[:, 1:2] * Y + H[:, 2:3]) * ww - pts[:, 2].astype(np.float64)[None, :]
        dy = (H[:, 3:4] * X + H[:, 4:5] * Y + H[:, 5:6]) * ww - pts[:, 3].astype(np.float64)[None, :]
        err = dx * dx + dy * dy
        inl = err <= thr * thr
    inl &= ok[:, None]
    return inl


def update_num_iters(confidence: float, ep: float, niters: float) -> float:
    """cv::RANSACUpdateNumIters(confidence, ep, modelPoints = 4, maxIters = niters) (calib3d ptsetreg.cpp)."""
    num = max(1.0 - confidence, 2.2250738585072014e-308)
    q = 1.0 - ep
    denom = 1.0 - q * q * q * q
    if denom < 2.2250738585072014e-308:
        return 0.0
    ln, ld = np.log(num), np.log(denom)
    return niters if (ld >= 0.0 or -ln >= niters * (-ld)) else float(np.rint(ln / ld))


NT = 256        # threads per CTA of csrc/homography.cu: the summation order below is that kernel's


def _block_sum(vals: np.ndarray, mask: np.ndarray) -> float:
    """Sum of vals[mask] in the order of the kernel's block_sum: thread t adds its points t, t+256, ... in order (starting
    from 0.0), a xor-butterfly (16, 8, 4, 2, 1) inside each warp, then the 8 warp partials in warp order."""
    m = len(vals)
    pad = (-m) % NT
    v = np.concatenate([np.where(mask, vals, 0.0), np.zeros(pad)]).reshape(-1, NT)
    part = np.zeros(NT)
    for row in v:                                  # masked-out points contribute +0.0, which leaves the sum bits unchanged
        part = part + row
    lanes = part.reshape(NT // 32, 32)
    for o in (16, 8, 4, 2, 1):
        lanes = lanes + lanes[:, np.arange(32) ^ o]
    t = lanes[0, 0]
    for w in range(1, NT // 32):
        t = t + lanes[w, 0]
    return float(t)


def refit(pts: np.ndarray, H: np.ndarray, thr: float):
    """Stage 3 of the kernel: normalised least-squares homography on the inliers of H and the mask of that model
    (cv::findHomography: runKernel on the consensus set + mask of the refined model).  Returns (count, mask) or None."""
    p64 = pts.astype(np.float64)
    inl = consensus(pts, H[None, :], np.ones(1, bool), thr)[0]
    n = _block_sum(np.ones(len(pts)), inl)
    c1x, c1y = _block_sum(p64[:, 0], inl) / n, _block_sum(p64[:, 1], inl) / n
    c2x, c2y = _block_sum(p64[:, 2], inl) / n, _block_sum(p64[:, 3], inl) / n
    d1x, d1y = _block_sum(np.abs(p64[:, 0] - c1x), inl), _block_sum(np.abs(p64[:, 1] - c1y), inl)
    d2x, d2y = _block_sum(np.abs(p64[:, 2] - c2x), inl), _block_sum(np.abs(p64[:, 3] - c2y), inl)
    eps = 2.220446049250313e-16
    if not (d1x > eps and d1y > eps and d2x > eps and d2y > eps):
        return None
    s1x, s1y, s2x, s2y = n / d1x, n / d1y, n / d2x, n / d2y
    X, Y = (p64[:, 0] - c1x) * s1x, (p64[:, 1] - c1y) * s1y
    x, y = (p64[:, 2] - c2x) * s2x, (p64[:, 3] - c2y) * s2y
    one, zero = np.ones_like(X), np.zeros_like(X)
    Lx = [X, Y, one, zero, zero, zero, -x * X, -x * Y, -x]
    Ly = [zero, zero, zero, X, Y, one, -y * X, -y * Y, -y]
    Mx = np.zeros((8, 9))
    for r in range(9):
        for c in range(r, 9):
            t = _block_sum(Lx[r] * Lx[c] + Ly[r] * Ly[c], inl)
            if r < 8 and c < 8:
                Mx[r, c] = Mx[c, r] = t
            elif r < 8:
                Mx[r, 8] = -t
    for col in range(8):                           # Gaussian elimination, partial pivoting (first maximum wins)
        piv = col
        for r in range(col + 1, 8):
            if abs(Mx[r, col]) > abs(Mx[piv, col]):
                piv = r
        if not abs(Mx[piv, col]) > 0.0:
            return None
        if piv != col:
            Mx[[col, piv]] = Mx[[piv, col]]
        for r in range(col + 1, 8):
            f = Mx[r, col] / Mx[col, col]
            for c in range(col, 9):
                Mx[r, c] = Mx[r, c] - f * Mx[col, c]
    hn_ = np.ones(9)
    for r in range(7, -1, -1):
        acc = Mx[r, 8]
        for c in range(r + 1, 8):
            acc = acc - Mx[r, c] * hn_[c]
        hn_[r] = acc / Mx[r, r]
    G = np.zeros(9)
    for r in range(3):
        G[3 * r] = hn_[3 * r] * s1x
        G[3 * r + 1] = hn_[3 * r + 1] * s1y
        G[3 * r + 2] = hn_[3 * r + 2] - hn_[3 * r] * s1x * c1x - hn_[3 * r + 1] * s1y * c1y
    R = np.zeros(9)
    for c in range(3):
        R[c] = G[c] / s2x + c2x * G[6 + c]
        R[3 + c] = G[3 + c] / s2y + c2y * G[6 + c]
        R[6 + c] = G[6 + c]
    if not (R[8] != 0.0 and np.isfinite(R[8])):
        return None
    Hr = R * (1.0 / R[8])
    mask = consensus(pts, Hr[None, :], np.ones(1, bool), thr)[0]
    return int(mask.sum()), mask, Hr


def ransac_inliers(pts: np.ndarray, thr: float, pair: int, seed: int = 0, max_iters: int = 2000, confidence: float = 0.995,
                   refine: bool = True, details: bool = False):
    """(inlier count, winning hypothesis, inlier mask) for one pair, as csrc/homography.cu computes it; (-1, -1, None)
    when there are fewer than 4 matches.  details=True also returns the consensus size of the minimal model."""
    def ret(cnt, hyp, mask, rcnt):
        return (cnt, hyp, mask, rcnt) if details else (cnt, hyp, mask)
    pts = np.ascontiguousarray(pts, np.float32)
    m = pts.shape[0]
    if m < 4:
        return ret(-1, -1, None, -1)
    if m == 4:      # cv::findHomography: `if (method == 0 || npoints == 4)` -> no RANSAC, mask = all ones
        return ret(4, -1, np.ones(4, bool), 4)
    idx, ok = minimal_sets(seed, pair, max_iters, m)
    H, ok = models(pts, idx, ok)
    inl = consensus(pts, H, ok, thr)
    counts = inl.sum(1)
    # RANSACPointSetRegistrator::run replayed over the hypotheses in order: a better model shrinks the iteration budget,
    # rejected minimal sets are redrawn by getSubset and do not count as iterations
    best, hyp, niters, it = 0, -1, float(max_iters), 0
    for h in range(max_iters):
        if not it < niters:
            break
        if not ok[h]:
            continue
        c = int(counts[h])
        if c > max(best, 3):
            best, hyp = c, h
            niters = update_num_iters(confidence, (m - c) / m, niters)
        it += 1
    if hyp < 0:
        return ret(0, -1, np.zeros(m, bool), 0)
    if refine:
        r = refit(pts, H[hyp], thr)
        if r is not None:
            return ret(r[0], hyp, r[1], best)
    return ret(best, hyp, inl[hyp], best)


def aligned_points(kp_left: np.ndarray, kp_right: np.ndarray, matches: np.ndarray) -> np.ndarray:
    """ShotMatches::alignFeatures (Scene.cpp:95-110): left <-> queryIdx, right <-> trainIdx -> [M, 4] float32."""
    return np.concatenate([kp_left[matches["queryIdx"]], kp_right[matches["trainIdx"]]], 1).astype(np.float32)


def homography_ratios(keypoints, pairs, match_lists, thr, seed=0, max_iters=2000, confidence=0.995, refine=True):
    """The stage for a list of pairs: ratio = inliers / matches, -1 where no homography is attempted."""
    out = np.full(len(pairs), -1.0)
    cnt = np.full(len(pairs), -1, np.int64)
    for p, (l, r) in enumerate(np.asarray(pairs).reshape(-1, 2)):
        m = match_lists[p]
        if m is None or len(m) < 4:
            continue
        pts = aligned_points(keypoints[int(l)], keypoints[int(r)], m)
        k, _, _ = ransac_inliers(pts, thr if np.isscalar(thr) else thr[p], p, seed, max_iters, confidence, refine)
        cnt[p] = k
        out[p] = k / len(m)
    return out, cnt
