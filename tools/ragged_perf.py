"""Device-side pairs/s on a RAGGED scene (every image a different number of descriptors): the case where the knn kernel cannot
use the uniform-unit decode and finds each unit's pair by search (development aid; SFMMATCH_LIB selects the library build).

    python tools/ragged_perf.py [n_images] [max_rows]
"""
import hashlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import __graft_entry__ as ge
import workloads

sfm = ge.load_package()
n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 100
max_rows = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
m = sfm.Matcher(0)
full = workloads.sift_like_bank(n_img, max_rows)
rng = np.random.Generator(np.random.PCG64(11))
rows = rng.integers(max_rows // 3, max_rows + 1, size=n_img)
bank = [b[:r] for b, r in zip(full, rows)]
pairs = sfm.select_pairs(n_img, 0, 0)
m.upload_bank(bank)
stream = torch.cuda.ExternalStream(m.stream)
work = sum(2.0 * len(bank[i]) * len(bank[j]) * 128 for i, j in pairs)
best = 1e30
for rep in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    m.enqueue(pairs, sfm.NORM_L2)
    e1.record(stream)
    e1.synchronize()
    ms = e0.elapsed_time(e1)
    r = m.collect()
    best = min(best, ms)
    h = hashlib.sha1(r.offsets.tobytes() + r.matches.tobytes()).hexdigest()[:12]
    print(f"rep{rep}: {ms:.3f} ms for {len(pairs)} ragged pairs -> {len(pairs)/ms*1e3:.0f} pairs/s, {work/ms/1e9:.0f} TOP/s, sha1 {h}")
print(f"lib {os.path.basename(sfm.LIB_PATH)} best {best:.3f} ms {work/best/1e9:.0f} TOP/s")
