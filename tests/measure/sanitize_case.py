"""Small end-to-end case for compute-sanitizer (memcheck): every engine once, ragged sizes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as ge
import workloads
from oracle import oracle_np as orc
sfm = ge.load_package()
m = sfm.Matcher(0)
sizes = [300, 257, 129, 0, 33]
bank, prev = [], None
for i, n in enumerate(sizes):
    d = workloads.sift_like_image(i, n, prev if prev is not None and len(prev) else None); bank.append(d); prev = d
pairs = sfm.select_pairs(len(sizes), 0, 0)
m.upload_bank([b.astype(np.float32) for b in bank])
exp = orc.match_pairs(bank, pairs, orc.NORM_L2)
for eng in (sfm.ENGINE_TENSOR, sfm.ENGINE_TENSOR_IMAD, sfm.ENGINE_SIMT):
    r = m.match_pairs(pairs, sfm.NORM_L2, engine=eng)
    assert all(orc.dmatch_equal(r[p], exp[p]) for p in range(len(pairs)))
r = m.match_pairs(pairs, sfm.NORM_L2, k=1, cross_check=True)
r = m.match_pairs(pairs, sfm.NORM_L2, distinct=True, min_match_count=5)
ob = [workloads.orb_like_image(i, n) for i, n in enumerate(sizes)]
m.upload_bank(ob)
for eng in (sfm.ENGINE_TENSOR, sfm.ENGINE_SIMT):
    r = m.match_pairs(pairs, sfm.NORM_HAMMING, engine=eng)
idx, dist = m.knn_match(bank[0], bank[1], sfm.NORM_L2, 2)
q = np.random.default_rng(0).random((70, 128), dtype=np.float32)
idx, dist = m.knn_match(q, q[:33], sfm.NORM_L2, 2)
m.close()
print("sanitize case ok")
