"""Homography stage on the C3 workload (200 x 8192 SIFT-like, 19 900 pairs): device time of
sfm_homography_inlier_ratios on the device-resident match lists vs cv2.findHomography (the routine
SfM::calculateHomography calls) on a sample of the same pairs, all host threads over pairs like the reference's
`#pragma omp parallel for` (SfM.cpp:603).  usage: python tests/measure/homography_perf.py [images] [rows] [cpu_sample_pairs]"""
import json, os, sys, time
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import __graft_entry__ as ge
import workloads

sfm = ge.load_package()
n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 200
n_rows = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
n_cpu = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
m = sfm.Matcher(0)
bank = workloads.sift_like_bank(n_img, n_rows)
kps = workloads.sift_like_keypoints(n_img, n_rows)
pairs = sfm.select_pairs(n_img, 0, 0)
m.upload_bank(bank)
m.upload_keypoints(kps)
res = m.match_pairs(pairs, sfm.NORM_L2)
counts = res.counts()
stream = torch.cuda.ExternalStream(m.stream)
best = None
for rep in range(4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = m.homography_inlier_ratios(3.0, 2000, seed=rep)
    ratios, inl, hyp = r["ratio"], r["inliers"], r["hypothesis"]
    dt = time.perf_counter() - t0
    best = dt if best is None or dt < best else best
attempted = int((counts >= 4).sum())
out = {"pairs": int(len(pairs)), "pairs_with_ge4_matches": attempted, "matches": int(counts.sum()),
       "gpu_ms_incl_d2h": best * 1e3, "gpu_pairs_per_s": len(pairs) / best,
       "ratio_adjacent_pairs_mean": float(ratios[[i for i, (a, b) in enumerate(pairs) if b == a + 1]].mean())}
# ---- CPU: cv2.findHomography on a sample, threads over pairs
from oracle import cv2_ref
from oracle import homography_np as hn
rng = np.random.default_rng(7)
cand = np.nonzero(counts >= 4)[0]
sel = rng.choice(cand, size=min(n_cpu, len(cand)), replace=False)
pts = [hn.aligned_points(kps[pairs[p][0]], kps[pairs[p][1]], res[p]) for p in sel]
def one(x):
    return cv2_ref.find_homography_inliers(x[:, :2], x[:, 2:], 3.0)[0]
cores = os.cpu_count() or 1
cv2_ref.cv2.setNumThreads(1)
t0 = time.perf_counter()
with ThreadPoolExecutor(cores) as ex:
    cv_counts = list(ex.map(one, pts))
dt = time.perf_counter() - t0
gpu_r = ratios[sel]
cv_r = np.array(cv_counts) / counts[sel]
out.update({"cpu_sample_pairs": int(len(sel)), "cpu_s": dt, "cpu_pairs_per_s": len(sel) / dt, "cpu_threads": cores,
            "max_abs_ratio_diff_vs_cv2": float(np.abs(gpu_r - cv_r).max()),
            "mean_abs_ratio_diff_vs_cv2": float(np.abs(gpu_r - cv_r).mean()),
            "frac_pairs_within_0.05": float((np.abs(gpu_r - cv_r) <= 0.05).mean())})
dc = np.abs(inl[sel] - np.array(cv_counts))
tol = np.maximum(1, np.ceil(0.05 * counts[sel]))
out["frac_pairs_within_max(1 match, 5%)"] = float((dc <= tol).mean())
out["pairs_beyond_tolerance"] = [(int(counts[p]), int(inl[p]), int(c)) for p, c, d, t in zip(sel, cv_counts, dc, tol) if d > t][:20]
big = counts[sel] >= 40
out["pairs_ge40_matches"] = int(big.sum())
out["max_abs_ratio_diff_ge40_matches"] = float(np.abs(gpu_r - cv_r)[big].max()) if big.any() else None
print(json.dumps(out))
