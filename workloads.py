"""Synthetic descriptor banks for bench.py and the parity tests.

Not part of the product path and not part of the oracle: it only manufactures
inputs.  Recipes follow SURVEY.md §8(d) "Synthetic inputs".
"""
from __future__ import annotations

import numpy as np


def sift_like_image(i: int, n_rows: int, prev: np.ndarray | None = None, planted: float = 0.30,
                    seed_base: int = 1000) -> np.ndarray:
    """One image's SIFT-like descriptors: uint8 [n_rows, 128], integer-valued like cv::SIFT.

    gamma(0.6) -> L2-normalise -> clip 0.2 -> renormalise -> rint(512 x) clipped to 0..255;
    the first ``planted`` fraction of rows are noisy copies (sigma 6) of random rows of
    ``prev`` so that the ratio test has survivors."""
    rng = np.random.Generator(np.random.PCG64(seed_base + i))
    x = rng.gamma(0.6, size=(n_rows, 128)).astype(np.float32)
    x /= np.maximum(np.linalg.norm(x, axis=1, keepdims=True), 1e-12)
    np.minimum(x, 0.2, out=x)
    x /= np.maximum(np.linalg.norm(x, axis=1, keepdims=True), 1e-12)
    d = np.clip(np.rint(512.0 * x), 0, 255).astype(np.uint8)
    if prev is not None and prev.shape[0] > 0 and planted > 0:
        k = int(n_rows * planted)
        src = rng.integers(0, prev.shape[0], size=k)
        noisy = prev[src].astype(np.float32) + rng.normal(0.0, 6.0, size=(k, 128)).astype(np.float32)
        d[:k] = np.clip(np.rint(noisy), 0, 255).astype(np.uint8)
    return d


def sift_like_keypoints(n_images: int, n_rows: int, planted: float = 0.30, seed_base: int = 1000, width: float = 4000.0,
                        height: float = 3000.0, noise: float = 0.7):
    """KeyPoint.pt arrays (float32 [n_rows, 2]) consistent with sift_like_bank: the planted rows of image i sit where a
    random homography maps the keypoints of their source rows in image i-1 (+ Gaussian noise in pixels), every other
    row is uniform in the image.  Replays sift_like_image's generator to recover the source rows."""
    kps, prev_kp = [], None
    for i in range(n_images):
        rng = np.random.Generator(np.random.PCG64(seed_base + i))
        rng.gamma(0.6, size=(n_rows, 128))                                    # same stream position as sift_like_image
        krng = np.random.Generator(np.random.PCG64(seed_base + 500000 + i))
        kp = np.stack([krng.uniform(0, width, n_rows), krng.uniform(0, height, n_rows)], 1)
        if prev_kp is not None and planted > 0:
            k = int(n_rows * planted)
            src = rng.integers(0, prev_kp.shape[0], size=k)
            H = np.array([[1 + krng.normal(0, 0.02), krng.normal(0, 0.02), krng.normal(0, 40)],
                          [krng.normal(0, 0.02), 1 + krng.normal(0, 0.02), krng.normal(0, 40)],
                          [krng.normal(0, 3e-6), krng.normal(0, 3e-6), 1.0]])
            q = np.c_[prev_kp[src], np.ones(k)] @ H.T
            kp[:k] = q[:, :2] / q[:, 2:3] + krng.normal(0, noise, (k, 2))
        prev_kp = kp
        kps.append(kp.astype(np.float32))
    return kps


def sift_like_bank(n_images: int, n_rows: int, planted: float = 0.30, seed_base: int = 1000):
    bank, prev = [], None
    for i in range(n_images):
        prev = sift_like_image(i, n_rows, prev, planted, seed_base)
        bank.append(prev)
    return bank


def orb_like_image(i: int, n_rows: int, prev: np.ndarray | None = None, planted: float = 0.30,
                   flips: int = 20, seed_base: int = 5000) -> np.ndarray:
    """ORB-like descriptors: uint8 [n_rows, 32] uniform bits; ``planted`` rows are copies of
    rows of ``prev`` with ``flips`` random bit flips."""
    rng = np.random.Generator(np.random.PCG64(seed_base + i))
    d = rng.integers(0, 256, size=(n_rows, 32), dtype=np.uint8)
    if prev is not None and prev.shape[0] > 0 and planted > 0:
        k = int(n_rows * planted)
        src = rng.integers(0, prev.shape[0], size=k)
        c = prev[src].copy()
        bit = rng.integers(0, 256, size=(k, flips))
        for f in range(flips):
            c[np.arange(k), bit[:, f] >> 3] ^= (1 << (bit[:, f] & 7)).astype(np.uint8)
        d[:k] = c
    return d


def orb_like_bank(n_images: int, n_rows: int, planted: float = 0.30, flips: int = 20, seed_base: int = 5000):
    bank, prev = [], None
    for i in range(n_images):
        prev = orb_like_image(i, n_rows, prev, planted, flips, seed_base)
        bank.append(prev)
    return bank


def adversarial_sift(seed: int = 3):
    """Small descriptor sets exercising ties, duplicates, zero rows and ragged sizes."""
    rng = np.random.Generator(np.random.PCG64(seed))
    base = rng.integers(0, 120, size=(257, 128), dtype=np.uint8)
    dup = np.concatenate([base[:40], base[:40], base[10:20], np.zeros((3, 128), np.uint8), base[40:]])
    sat = np.full((5, 128), 255, np.uint8)
    return {
        "base": base,
        "dup": dup,                                  # exact ties at ranks 1 and 2
        "zeros": np.zeros((9, 128), np.uint8),
        "sat": np.concatenate([sat, np.zeros((2, 128), np.uint8), base[:7]]),  # max distance 128*255^2
        "one": base[:1],
        "two": base[:2],
        "n127": base[:127],
        "n129": base[:129],
        "empty": np.zeros((0, 128), np.uint8),
    }


def synthetic_photo(seed: int = 0, height: int = 480, width: int = 640) -> np.ndarray:
    """Synthetic grey image with structure at many scales (blobs, edges, fine texture): the input of the feature
    extraction stage (SfM::extractFeatures) when no photographs are at hand.  uint8, deterministic in `seed`."""
    rng = np.random.default_rng(77000 + seed)
    yy, xx = np.mgrid[0:height, 0:width].astype(np.float32)
    img = np.zeros((height, width), np.float32)
    n_blobs = max(40, height * width // 160)
    cx = rng.uniform(0, width, n_blobs).astype(np.float32)
    cy = rng.uniform(0, height, n_blobs).astype(np.float32)
    sg = np.exp(rng.uniform(np.log(1.2), np.log(12.0), n_blobs) - rng.exponential(0.0, n_blobs)).astype(np.float32)
    am = rng.uniform(-1.0, 1.0, n_blobs).astype(np.float32)
    for k in range(n_blobs):
        r = int(4 * sg[k]) + 1
        x0, x1 = max(0, int(cx[k]) - r), min(width, int(cx[k]) + r + 1)
        y0, y1 = max(0, int(cy[k]) - r), min(height, int(cy[k]) + r + 1)
        if x0 >= x1 or y0 >= y1:
            continue
        d2 = (xx[y0:y1, x0:x1] - cx[k]) ** 2 + (yy[y0:y1, x0:x1] - cy[k]) ** 2
        img[y0:y1, x0:x1] += am[k] * np.exp(-d2 / (2 * sg[k] * sg[k]))
    for _ in range(6):                                                    # a few straight edges
        a = rng.uniform(0, np.pi)
        off = rng.uniform(0.2, 0.8)
        img += rng.uniform(-0.5, 0.5) * ((np.cos(a) * xx / width + np.sin(a) * yy / height) > off)
    img += 0.05 * rng.standard_normal((height, width)).astype(np.float32)  # fine texture
    img = 128.0 + 56.0 * (img - img.mean(dtype=np.float64)) / max(float(img.std(dtype=np.float64)), 1e-6)
    return np.rint(np.clip(img, 0.0, 255.0)).astype(np.uint8)
