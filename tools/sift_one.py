"""One device feature extraction of a synthetic photograph (for `ncu`: kernel list of the stage).
    python tools/sift_one.py [height width [repeats]]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
import workloads  # noqa: E402

h, w = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1200, 1600)
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
sfm = ge.load_package()
m = sfm.Matcher(0)
img = workloads.synthetic_photo(1, h, w)
for _ in range(reps):
    m.features_clear()
    t = time.perf_counter()
    n = m.extract_sift(img, contrast_threshold=0.09)
    dt = time.perf_counter() - t
print((h, w), "keypoints", n, m.features_last_counts(), "ms", round(dt * 1e3, 3))
m.close()
