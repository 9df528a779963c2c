"""The reference's own matcher classes, reached through Python ``cv2``.

TEST INFRASTRUCTURE / CPU BASELINE ONLY (see oracle/oracle_np.py header).

The reference owns no matcher arithmetic: every distance on the hot path is
computed by OpenCV's ``cv::BFMatcher`` / ``cv::FlannBasedMatcher``
(``PhotogrammetrieCli.cpp:371-389``; call sites ``Unordered...cpp:51``,
``Video...cpp:62``, ``Grid...cpp:105``).  The reference binary cannot be built
here (OpenCV/Ceres/PCL/OpenMVS/CGAL/Boost C++ are absent, ``CMakeLists.txt:31-65``)
but the same library routines are importable as ``cv2`` 4.13.0 (reference pins
4.5.1, ``build.sh:105``; for integer-valued SIFT and for Hamming the results
are exact and version-independent).  This module is used to (1) pin
``oracle_np`` and generate ``tests/golden`` and (2) time the CPU baseline /
``bench.py --impl reference``.
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

try:
    import cv2
except Exception:  # pragma: no cover - cv2 is in the image; gate anyway
    cv2 = None

from .oracle_np import DMATCH_DTYPE, NORM_HAMMING, NORM_L2


def available() -> bool:
    return cv2 is not None


def batch_distance_k2(q: np.ndarray, t: np.ndarray, norm: int):
    """cv::batchDistance(..., K=2): the array-level routine BFMatcher::knnMatchImpl calls."""
    if norm == NORM_L2:
        q = np.ascontiguousarray(q, np.float32)
        t = np.ascontiguousarray(t, np.float32)
        dist, nidx = cv2.batchDistance(q, t, cv2.CV_32F, None, None, cv2.NORM_L2, K=2)
    else:
        dist, nidx = cv2.batchDistance(q, t, cv2.CV_32S, None, None, cv2.NORM_HAMMING, K=2)
        dist = dist.astype(np.float32)
    return nidx.astype(np.int32), dist.astype(np.float32)


def bf_knn_match(q: np.ndarray, t: np.ndarray, norm: int, k: int = 2):
    """cv::BFMatcher::create(norm)->knnMatch(q, t, matches, k) (PhotogrammetrieCli.cpp:378/:387)."""
    if norm == NORM_L2:
        q = np.ascontiguousarray(q, np.float32)
        t = np.ascontiguousarray(t, np.float32)
    m = cv2.BFMatcher(cv2.NORM_L2 if norm == NORM_L2 else cv2.NORM_HAMMING)
    return m.knnMatch(q, t, k=k)


def ratio_good(knn, ratio: float = 0.7) -> np.ndarray:
    """The literal loop of UnorderedFeatureMatchingStrategy.cpp:55-65 over cv2 DMatch lists."""
    rows = []
    for m in knn:
        if len(m) >= 2:
            if float(m[0].distance) < float(m[1].distance) * ratio:
                rows.append((m[0].queryIdx, m[0].trainIdx, m[0].imgIdx, m[0].distance))
        elif len(m) == 1:
            rows.append((m[0].queryIdx, m[0].trainIdx, m[0].imgIdx, m[0].distance))
    return np.array(rows, dtype=DMATCH_DTYPE)


def cross_check_match(q, t, norm):
    if norm == NORM_L2:
        q = np.ascontiguousarray(q, np.float32)
        t = np.ascontiguousarray(t, np.float32)
    m = cv2.BFMatcher(cv2.NORM_L2 if norm == NORM_L2 else cv2.NORM_HAMMING, crossCheck=True)
    res = m.match(q, t)
    return np.array([(x.queryIdx, x.trainIdx, x.imgIdx, x.distance) for x in res], dtype=DMATCH_DTYPE)


def flann_knn_ratio(q, t, norm, ratio=0.7):
    """FlannBasedMatcher exactly as PhotogrammetrieCli.cpp:371-384 configures it; the index is
    built inside every two-argument knnMatch call, as in the reference."""
    if norm == NORM_L2:
        q = np.ascontiguousarray(q, np.float32)
        t = np.ascontiguousarray(t, np.float32)
        fm = cv2.FlannBasedMatcher(dict(algorithm=1, trees=5), dict(checks=100))
    else:
        fm = cv2.FlannBasedMatcher(dict(algorithm=6, table_number=6, key_size=12, multi_probe_level=1),
                                   dict(checks=100))
    knn = fm.knnMatch(q, t, k=2)
    top1 = np.array([m[0].trainIdx if len(m) else -1 for m in knn], np.int32)
    return top1, ratio_good(knn, ratio)


def _pair_good_count(args):
    q, t, norm, ratio = args
    nidx, dist = batch_distance_k2(q, t, norm)
    keep = dist[:, 0].astype(np.float64) < dist[:, 1].astype(np.float64) * ratio
    return int(keep.sum())


def time_pairs(bank, pairs, norm, ratio=0.7, topology="inner", threads=None):
    """Time the CPU matcher over ``pairs``.  topology 'inner': pairs serial, OpenCV threads inside
    knnMatch; 'outer': a thread per pair with single-threaded OpenCV — mirrors the reference's
    ``#pragma omp parallel for`` over pairs (Unordered...cpp:40).  Returns (seconds, total_good)."""
    import time
    threads = threads or os.cpu_count() or 1
    work = [((np.ascontiguousarray(bank[l], np.float32) if norm == NORM_L2 else bank[l]),
             (np.ascontiguousarray(bank[r], np.float32) if norm == NORM_L2 else bank[r]), norm, ratio)
            for l, r in np.asarray(pairs).reshape(-1, 2)]
    if topology == "inner":
        cv2.setNumThreads(threads)
        t0 = time.perf_counter()
        good = sum(_pair_good_count(w) for w in work)
        dt = time.perf_counter() - t0
    else:
        cv2.setNumThreads(1)
        t0 = time.perf_counter()
        with ThreadPoolExecutor(threads) as ex:
            good = sum(ex.map(_pair_good_count, work))
        dt = time.perf_counter() - t0
        cv2.setNumThreads(threads)
    return dt, good


def find_homography_inliers(pts_left: np.ndarray, pts_right: np.ndarray, threshold: float = 3.0):
    """cv::findHomography(left, right, cv::RANSAC, threshold, mask) as SfM::calculateHomography calls it
    (SfM.cpp:622-625, default maxIters 2000 / confidence 0.995) -> (inlier count, mask); (0, None) if no model."""
    H, mask = cv2.findHomography(np.ascontiguousarray(pts_left, np.float32), np.ascontiguousarray(pts_right, np.float32),
                                 cv2.RANSAC, float(threshold))
    if H is None or mask is None:
        return 0, None
    return int(np.count_nonzero(mask)), mask.reshape(-1).astype(bool)
