"""CPU restatement of cv::ORB (detect + compute) — TEST INFRASTRUCTURE, never imported by the product path.

The reference's second detector is ``cv::ORB::create(featureLimit)`` (PhotogrammetrieCli.cpp:347-348, used by
run-scripts/run-orb-sequence.sh:4), called as ``detect()`` then ``compute()`` (SfM.cpp:586-587).  cv::ORB lives in OpenCV
(un-vendored, pinned 4.5.1); this file restates its published algorithm (modules/features2d/src/orb.cpp, fast.cpp,
fast_score.cpp, imgproc resize.cpp / filter.simd.hpp) operation by operation in numpy:

  pyramid     8 levels, scale 1.2**level, level size = cvRound(size / scale), each level resized from the PREVIOUS one with
              INTER_LINEAR_EXACT (8.8 fixed-point coefficients, (sum + 2**15) >> 16), borders REFLECT_101
  FAST        FAST-9/16, threshold 20, score = largest threshold that keeps the corner (cornerScore<16>), 3 x 3 non-maximum
              suppression with strict '>', keypoints closer than edgeThreshold = 31 to the level border removed
  retainBest  per level: 2 x quota by FAST score, then quota by Harris response; as a SET: everything >= the n-th best
              (cv::KeyPointsFilter::retainBest keeps ties; its ORDER comes out of std::nth_element and is not restated)
  Harris      7 x 7 block of Sobel-like integer gradients, k = 0.04, float arithmetic in OpenCV's order
  angle       intensity centroid over the circular patch of radius 15 (umax table), cv::fastAtan2
  blur        GaussianBlur(7 x 7, sigma 2) of every level: the level is a SUBMATRIX of the pyramid buffer, so OpenCV takes the
              generic separable float filter (row pass in tap order, column pass centre + symmetric pairs, cvRound), not the
              8-bit fixed-point path — found by trying the candidates against cv2's descriptors (tests/golden/make_golden_orb.py)
  descriptor  256 intensity comparisons of the rotated bit pattern (cvRound of the rotated coordinates)

Pinned against cv2 4.13 golden vectors (tests/golden/orb_extract.npz, tests/test_orb_oracle.py): identical keypoint SETS
(x, y, octave), identical responses and descriptors, angles within 1e-4 degrees.  The 256 x 4 sampling pattern
(bit_pattern_31_) is OpenCV data, not part of the reference tree: oracle/orb_pattern.py holds it, extracted from the installed
cv2 binary by tests/golden/make_golden_orb.py (which also re-checks it against cv2's descriptors).
"""
from __future__ import annotations

import math

import numpy as np

from .orb_pattern import BIT_PATTERN_31

f32 = np.float32
KEYPOINT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"), ("octave", "<i4")])
HARRIS_K = f32(0.04)
# scaleFactor is a float argument of ORB::create (1.2f) held in a double: 1.2000000476837158
PATCH_SIZE, EDGE_THRESHOLD, N_LEVELS, SCALE_FACTOR, FAST_THRESHOLD, HARRIS_BLOCK = 31, 31, 8, float(np.float32(1.2)), 20, 7
HALF_PATCH = PATCH_SIZE // 2
# Bresenham circle of radius 3 in OpenCV's order (fast_score.cpp, makeOffsets, patternSize 16)
CIRCLE = ((0, 3), (1, 3), (2, 2), (3, 1), (3, 0), (3, -1), (2, -2), (1, -3), (0, -3), (-1, -3), (-2, -2), (-3, -1), (-3, 0), (-3, 1), (-2, 2), (-1, 3))


# ------------------------------------------------------------------------------------------------ pyramid
def level_scale(level: int) -> np.float32:
    return f32(math.pow(SCALE_FACTOR, float(level)))          # getScale: (float)std::pow(scaleFactor, level - firstLevel)


def level_size(rows: int, cols: int, level: int):
    s = level_scale(level)
    return int(np.rint(f32(rows) / s)), int(np.rint(f32(cols) / s))


def linear_exact_coeffs(dst_n: int, src_n: int):
    """interpolationLinear<ufixedpoint16>::getCoeffs: source offset and the weight of the RIGHT tap in 1/256 per destination
    index (left weight = 256 - right); positions before the first / after the last source sample clamp."""
    scale = np.float64(src_n) / np.float64(dst_n)
    f = scale * (np.arange(dst_n, dtype=np.float64) + 0.5) - 0.5
    i = np.floor(f).astype(np.int64)
    ofs = np.zeros(dst_n, np.int64)
    c1 = np.zeros(dst_n, np.int64)
    inside = (i >= 0) & (i < src_n - 1) & (src_n > 1)
    ofs[inside] = i[inside]
    c1[inside] = np.rint((f[inside] - i[inside]) * 256.0).astype(np.int64)
    ofs[(i >= src_n - 1) | ((i >= 0) & (src_n <= 1))] = src_n - 1
    return ofs, c1


def resize_linear_exact(src: np.ndarray, rows: int, cols: int) -> np.ndarray:
    """cv::resize(src, dst, Size(cols, rows), 0, 0, INTER_LINEAR_EXACT) for CV_8UC1."""
    sh, sw = src.shape
    ox, cx = linear_exact_coeffs(cols, sw)
    oy, cy = linear_exact_coeffs(rows, sh)
    ox1, oy1 = np.minimum(ox + 1, sw - 1), np.minimum(oy + 1, sh - 1)
    s = src.astype(np.int64)
    h = s[:, ox] * (256 - cx)[None, :] + s[:, ox1] * cx[None, :]
    v = h[oy, :] * (256 - cy)[:, None] + h[oy1, :] * cy[:, None]
    return np.clip((v + 32768) >> 16, 0, 255).astype(np.uint8)


def build_pyramid(gray: np.ndarray, n_levels: int = N_LEVELS):
    levels, prev = [], gray
    for lv in range(n_levels):
        cur = gray if lv == 0 else resize_linear_exact(prev, *level_size(*gray.shape, lv))
        levels.append(cur)
        prev = cur
    return levels


def border_width() -> int:
    desc_patch = int(math.ceil(HALF_PATCH * math.sqrt(2.0)))
    return max(EDGE_THRESHOLD, max(desc_patch, HARRIS_BLOCK // 2)) + 1      # 32


# ------------------------------------------------------------------------------------------------ FAST
def fast9_scores(img: np.ndarray, threshold: int = FAST_THRESHOLD) -> np.ndarray:
    """cornerScore<16> at FAST-9 corners (threshold), 0 elsewhere: max over the 16 arcs of 9 contiguous circle pixels of
    min(v - p) (bright centre) / min(p - v) (dark centre), minus 1; a corner iff that maximum exceeds the threshold."""
    h, w = img.shape
    v = img.astype(np.int32)
    d = np.zeros((16, h, w), np.int32)
    for k, (dx, dy) in enumerate(CIRCLE):
        sh = np.zeros_like(v)
        ys, ye, xs, xe = max(0, -dy), min(h, h - dy), max(0, -dx), min(w, w - dx)
        sh[ys:ye, xs:xe] = v[ys + dy:ye + dy, xs + dx:xe + dx]
        d[k] = v - sh
    best = np.full((h, w), -(10 ** 6), np.int32)
    for start in range(16):
        arc = d[[(start + i) % 16 for i in range(9)]]
        best = np.maximum(best, arc.min(0))
        best = np.maximum(best, (-arc).min(0))
    score = np.where(best > threshold, best - 1, 0).astype(np.int32)
    score[:3] = 0
    score[-3:] = 0
    score[:, :3] = 0
    score[:, -3:] = 0
    return score


def fast9_corners(img: np.ndarray, threshold: int = FAST_THRESHOLD):
    """cv::FAST(img, kps, threshold, nonmaxSuppression = true): (xs, ys, scores) in row-major order."""
    s = fast9_scores(img, threshold)
    h, w = s.shape
    p = np.pad(s, 1)
    keep = s > 0
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            if dx or dy:
                keep &= s > p[1 + dy:1 + dy + h, 1 + dx:1 + dx + w]
    ys, xs = np.nonzero(keep)
    return xs, ys, s[ys, xs]


# ------------------------------------------------------------------------------------------------ selection
def features_per_level(nfeatures: int, n_levels: int = N_LEVELS):
    factor = f32(1.0 / SCALE_FACTOR)
    nd = f32(nfeatures) * (f32(1) - factor) / (f32(1) - f32(math.pow(float(factor), n_levels)))
    out, total = [], 0
    for _ in range(n_levels - 1):
        n = int(np.rint(nd))
        out.append(n)
        total += n
        nd = nd * factor
    out.append(max(nfeatures - total, 0))
    return out


def retain_best(resp: np.ndarray, n: int) -> np.ndarray:
    """cv::KeyPointsFilter::retainBest as a set: indices of everything >= the n-th best response."""
    if n <= 0:
        return np.zeros(0, np.int64)
    if len(resp) <= n:
        return np.arange(len(resp))
    thr = np.sort(resp)[::-1][n - 1]
    return np.nonzero(resp >= thr)[0]


def harris_responses(ext: np.ndarray, border: int, xs: np.ndarray, ys: np.ndarray) -> np.ndarray:
    """HarrisResponses(..., blockSize 7, k 0.04): integer a = sum Ix^2, b = sum Iy^2, c = sum IxIy over the block, then
    (a b - c^2 - k (a + b)^2) * scale^4 in float, scale = 1 / (4 * 7 * 255)."""
    p = ext.astype(np.int64)
    r = HARRIS_BLOCK // 2
    dy, dx = np.mgrid[-r:r + 1, -r:r + 1]
    yy = (ys[:, None] + border + dy.reshape(1, -1)).astype(np.int64)
    xx = (xs[:, None] + border + dx.reshape(1, -1)).astype(np.int64)
    ix = (p[yy, xx + 1] - p[yy, xx - 1]) * 2 + (p[yy - 1, xx + 1] - p[yy - 1, xx - 1]) + (p[yy + 1, xx + 1] - p[yy + 1, xx - 1])
    iy = (p[yy + 1, xx] - p[yy - 1, xx]) * 2 + (p[yy + 1, xx - 1] - p[yy - 1, xx - 1]) + (p[yy + 1, xx + 1] - p[yy - 1, xx + 1])
    a = (ix * ix).sum(1).astype(np.float32)      # (float)a: exact integers well below 2**24? no: rounded like the C cast
    b = (iy * iy).sum(1).astype(np.float32)
    c = (ix * iy).sum(1).astype(np.float32)
    scale = f32(1.0) / (f32(1 << 2) * f32(HARRIS_BLOCK) * f32(255.0))
    s4 = scale * scale * scale * scale
    return ((a * b - c * c - HARRIS_K * (a + b) * (a + b)) * s4).astype(np.float32)


def umax_table(hp: int = HALF_PATCH):
    vmax = int(math.floor(hp * math.sqrt(2.0) / 2 + 1))
    vmin = int(math.ceil(hp * math.sqrt(2.0) / 2))
    umax = [0] * (hp + 2)
    for v in range(vmax + 1):
        umax[v] = int(np.rint(math.sqrt(float(hp * hp - v * v))))
    v0 = 0
    for v in range(hp, vmin - 1, -1):
        while umax[v0] == umax[v0 + 1]:
            v0 += 1
        umax[v] = v0
        v0 += 1
    return umax[:hp + 1]


def fast_atan2_deg(y: np.ndarray, x: np.ndarray) -> np.ndarray:
    """cv::fastAtan2 on float32 arrays (the polynomial of mathfuncs_core; degrees in [0, 360))."""
    y, x = y.astype(np.float32), x.astype(np.float32)
    ax, ay = np.abs(x), np.abs(y)
    p1, p3 = f32(0.9997878412794807 * (180 / np.pi)), f32(-0.3258083974640975 * (180 / np.pi))
    p5, p7 = f32(0.1555786518463281 * (180 / np.pi)), f32(-0.04432655554792128 * (180 / np.pi))
    eps = f32(2.220446049250313e-16)
    with np.errstate(divide="ignore", invalid="ignore"):
        c1 = ay / (ax + eps)
        c2 = ax / (ay + eps)
    c = np.where(ax >= ay, c1, c2).astype(np.float32)
    cc = c * c
    a = (((p7 * cc + p5) * cc + p3) * cc + p1) * c
    a = np.where(ax >= ay, a, f32(90) - a).astype(np.float32)
    a = np.where(x < 0, f32(180) - a, a).astype(np.float32)
    a = np.where(y < 0, f32(360) - a, a).astype(np.float32)
    return a


def ic_angles(ext: np.ndarray, border: int, xs: np.ndarray, ys: np.ndarray) -> np.ndarray:
    """IC_Angle: m10 = sum u I, m01 = sum v I over the circular patch (integers), fastAtan2((float)m01, (float)m10)."""
    um = umax_table()
    us, vs = [], []
    for v in range(-HALF_PATCH, HALF_PATCH + 1):
        d = um[abs(v)]
        for u in range(-d, d + 1):
            us.append(u)
            vs.append(v)
    us, vs = np.array(us, np.int64), np.array(vs, np.int64)
    p = ext.astype(np.int64)
    vals = p[(ys[:, None] + border + vs[None, :]), (xs[:, None] + border + us[None, :])]
    m10 = (vals * us[None, :]).sum(1)
    m01 = (vals * vs[None, :]).sum(1)
    return fast_atan2_deg(m01.astype(np.float32), m10.astype(np.float32))


# ------------------------------------------------------------------------------------------------ descriptors
def gaussian_kernel_f32(ksize: int = 7, sigma: float = 2.0) -> np.ndarray:
    x = np.arange(ksize, dtype=np.float64) - (ksize - 1) / 2
    t = np.exp(-(x * x) / (2 * sigma * sigma))
    return (t / t.sum()).astype(np.float32)


def blur_level(level: np.ndarray) -> np.ndarray:
    """GaussianBlur(level, level, Size(7, 7), 2, 2, BORDER_REFLECT_101) through the generic separable float filter: row pass
    s = k0 x0 + k1 x1 + ... in tap order, column pass s = k3 r3 + k4 (r4 + r2) + k5 (r5 + r1) + k6 (r6 + r0), cvRound, all float32
    with separately rounded multiply and add."""
    k = gaussian_kernel_f32()
    h, w = level.shape
    p = np.pad(level.astype(np.float32), 3, mode="reflect")
    rows = k[0] * p[:, 0:w]
    for i in range(1, 7):
        rows = rows + k[i] * p[:, i:i + w]
    out = k[3] * rows[3:3 + h]
    for j in (1, 2, 3):
        out = out + k[3 + j] * (rows[3 + j:3 + j + h] + rows[3 - j:3 - j + h])
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


def descriptors(ext_blur: np.ndarray, border: int, xs: np.ndarray, ys: np.ndarray, angles_deg: np.ndarray) -> np.ndarray:
    ang = angles_deg.astype(np.float32) * f32(np.pi / 180.0)
    a = np.cos(ang.astype(np.float64)).astype(np.float32)
    b = np.sin(ang.astype(np.float64)).astype(np.float32)
    pat = BIT_PATTERN_31.astype(np.float32)
    px, py = pat[:, [0, 2]].reshape(1, 512), pat[:, [1, 3]].reshape(1, 512)
    rx = np.rint(px * a[:, None] - py * b[:, None]).astype(np.int64)
    ry = np.rint(px * b[:, None] + py * a[:, None]).astype(np.int64)
    vals = ext_blur[ys[:, None] + border + ry, xs[:, None] + border + rx].astype(np.int32).reshape(len(xs), 256, 2)
    bits = (vals[:, :, 0] < vals[:, :, 1]).astype(np.uint8)
    return np.packbits(bits.reshape(len(xs), 32, 8)[:, :, ::-1], axis=2).reshape(len(xs), 32)


# ------------------------------------------------------------------------------------------------ the detector
def detect_and_compute(gray: np.ndarray, nfeatures: int = 500):
    """cv::ORB::create(nfeatures): detect() then compute().  Keypoints come back ordered by (level, y, x) — cv::ORB's own order
    is whatever std::nth_element left inside retainBest; the SET is the same."""
    gray = np.ascontiguousarray(gray, np.uint8)
    border = border_width()
    quota = features_per_level(nfeatures)
    kps, descs = [], []
    levels = build_pyramid(gray)
    for lv, img in enumerate(levels):
        h, w = img.shape
        if h <= 2 * EDGE_THRESHOLD or w <= 2 * EDGE_THRESHOLD or quota[lv] <= 0:
            continue
        xs, ys, sc = fast9_corners(img)
        inb = (xs >= EDGE_THRESHOLD) & (xs < w - EDGE_THRESHOLD) & (ys >= EDGE_THRESHOLD) & (ys < h - EDGE_THRESHOLD)
        xs, ys, sc = xs[inb], ys[inb], sc[inb]
        sel = retain_best(sc.astype(np.float32), 2 * quota[lv])
        xs, ys = xs[sel], ys[sel]
        ext = np.pad(img, border, mode="reflect")
        resp = harris_responses(ext, border, xs, ys)
        sel = retain_best(resp, quota[lv])
        xs, ys, resp = xs[sel], ys[sel], resp[sel]
        if len(xs) == 0:
            continue
        ang = ic_angles(ext, border, xs, ys)
        ext_blur = np.pad(blur_level(img), border, mode="reflect")
        d = descriptors(ext_blur, border, xs, ys, ang)
        s = level_scale(lv)
        k = np.zeros(len(xs), KEYPOINT_DTYPE)
        k["x"], k["y"] = xs.astype(np.float32) * s, ys.astype(np.float32) * s
        k["size"], k["angle"], k["response"], k["octave"] = f32(PATCH_SIZE) * s, ang, resp, lv
        kps.append(k)
        descs.append(d)
    if not kps:
        return np.zeros(0, KEYPOINT_DTYPE), np.zeros((0, 32), np.uint8)
    return np.concatenate(kps), np.concatenate(descs)
