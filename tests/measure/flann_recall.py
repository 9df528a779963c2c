"""FLANN-mode report (BASELINE config 5): the reference's approximate matcher (cv2.FlannBasedMatcher, KDTree(5),
checks=100, index built inside every knnMatch call as the reference does) against the exact GPU result on a fixed
random sample of pairs: time, top-1 recall, recall/precision of the ratio-test survivors."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as ge
import workloads
from oracle import cv2_ref
from oracle.oracle_np import NORM_L2

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
n_pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 12
sfm = ge.load_package()
n_img = 12
bank = workloads.sift_like_bank(n_img, rows)
rng = np.random.Generator(np.random.PCG64(7))
allp = sfm.select_pairs(n_img, 0, 0)
sel = allp[rng.choice(len(allp), size=min(n_pairs, len(allp)), replace=False)]
sel = np.concatenate([sel, [[0, 1], [1, 2]]])            # two pairs with planted matches
m = sfm.Matcher(0)
m.upload_bank(bank)
t0 = time.perf_counter(); res = m.match_pairs(sel, sfm.NORM_L2); t_gpu = time.perf_counter() - t0
out = {"rows": rows, "pairs": len(sel), "gpu_exact_s_total": t_gpu, "per_pair": []}
t_flann = 0.0
tp = fp = fn = 0
top1_hits = top1_n = 0
for p, (l, r) in enumerate(sel):
    idx, _ = m.knn_match(bank[l], bank[r], sfm.NORM_L2, 1)
    t0 = time.perf_counter()
    top1, good = cv2_ref.flann_knn_ratio(bank[l], bank[r], NORM_L2)
    dt = time.perf_counter() - t0
    t_flann += dt
    exact = {(int(a), int(b)) for a, b in zip(res[p]["queryIdx"], res[p]["trainIdx"])}
    approx = {(int(a), int(b)) for a, b in zip(good["queryIdx"], good["trainIdx"])}
    tp += len(exact & approx); fp += len(approx - exact); fn += len(exact - approx)
    top1_hits += int((top1 == idx[:, 0]).sum()); top1_n += len(top1)
    out["per_pair"].append({"pair": [int(l), int(r)], "flann_s": dt, "exact_good": len(exact), "flann_good": len(approx)})
out.update({"flann_s_per_pair": t_flann / len(sel), "flann_top1_recall": top1_hits / max(1, top1_n),
            "survivor_recall": tp / max(1, tp + fn), "survivor_precision": tp / max(1, tp + fp),
            "cores": os.cpu_count()})
print(json.dumps(out))
