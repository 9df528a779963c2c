"""Golden fixtures for the homography stage (SfM::calculateHomography, SfM.cpp:599-637), from cv2 in the BUILD CONTAINER:
    python tests/golden/make_golden_homography.py
Keypoints of images/insel (same detectors as make_golden.py, so rows line up with insel_{sift,orb}.npz), and
cv2.findHomography(RANSAC, 3.0) inlier counts / masks on the golden ratio-test survivors of every pair, plus synthetic
planted-homography cases.  cv2's RANSAC uses a fixed RNG seed, so these numbers are reproducible."""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import cv2_ref  # noqa: E402

IMG = "/root/reference/images/insel"


def keypoints(det):
    kps, descs, sizes = [], [], []
    for i in (1, 2, 3):
        img = cv2.imread(f"{IMG}/{i}.jpg", cv2.IMREAD_GRAYSCALE)
        d = cv2.SIFT_create(0, 3, 0.09) if det == "SIFT" else cv2.ORB_create(30000)
        kp = d.detect(img, None)
        kp, desc = d.compute(img, kp)
        kps.append(np.array([k.pt for k in kp], np.float32).reshape(-1, 2))
        descs.append(desc)
        sizes.append((img.shape[1], img.shape[0]))
    return kps, descs, sizes


def synthetic_case(seed, n, outlier_frac, noise):
    rng = np.random.default_rng(seed)
    H = np.array([[1.0 + rng.normal(0, 0.03), rng.normal(0, 0.03), rng.normal(0, 20)],
                  [rng.normal(0, 0.03), 1.0 + rng.normal(0, 0.03), rng.normal(0, 20)],
                  [rng.normal(0, 2e-5), rng.normal(0, 2e-5), 1.0]])
    p1 = rng.uniform(0, 720, (n, 2)).astype(np.float32)
    q = np.c_[p1, np.ones(n)] @ H.T
    p2 = (q[:, :2] / q[:, 2:3] + rng.normal(0, noise, (n, 2))).astype(np.float32)
    out = rng.random(n) < outlier_frac
    p2[out] = rng.uniform(0, 720, (int(out.sum()), 2)).astype(np.float32)
    return p1, p2, ~out


def main():
    out = {}
    for det in ("SIFT", "ORB"):
        kps, descs, sizes = keypoints(det)
        gold = np.load(os.path.join(HERE, f"insel_{det.lower()}.npz"))
        tag = det.lower()
        for i in range(3):
            ref = gold[f"desc{i}"]
            assert np.array_equal(descs[i].astype(ref.dtype), ref), "keypoints do not line up with the descriptor fixtures"
            out[f"{tag}_kp{i}"] = kps[i]
        out[f"{tag}_sizes"] = np.array(sizes, np.int32)
        for a, b in ((0, 1), (0, 2), (1, 2)):
            good = gold[f"p{a}{b}_good"]
            n, mask = cv2_ref.find_homography_inliers(kps[a][good["queryIdx"]], kps[b][good["trainIdx"]], 3.0)
            out[f"{tag}_h{a}{b}_count"] = np.int64(n)
            out[f"{tag}_h{a}{b}_mask"] = mask
            print(det, a, b, "matches", len(good), "cv2 inliers", n, "ratio", n / len(good))
    for k, (seed, n, frac, noise) in enumerate(((1, 400, 0.4, 0.5), (2, 60, 0.2, 0.3), (3, 2500, 0.7, 0.8), (4, 12, 0.3, 0.2),
                                                (5, 5000, 0.5, 1.0))):
        p1, p2, inl = synthetic_case(seed, n, frac, noise)
        cnt, mask = cv2_ref.find_homography_inliers(p1, p2, 3.0)
        out[f"syn{k}_p1"], out[f"syn{k}_p2"], out[f"syn{k}_planted"] = p1, p2, inl
        out[f"syn{k}_count"] = np.int64(cnt)
        print("syn", k, "n", n, "planted inliers", int(inl.sum()), "cv2 inliers", cnt)
    np.savez_compressed(os.path.join(HERE, "insel_homography.npz"), **out)


if __name__ == "__main__":
    main()
