"""Device-side timing of the non-integer CV_32F path (3xTF32 tcgen05 + fp32 re-rank) vs the CUDA-core fp32 kernel
on a RootSIFT-like all-pairs workload (development aid; prints per-kernel times via sfm_set_profiling)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import __graft_entry__ as ge
import workloads

sfm = ge.load_package()
n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 12
n_rows = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
m = sfm.Matcher(0)
ub = workloads.sift_like_bank(n_img, n_rows)
# RootSIFT of the synthetic SIFT bank: sqrt(x / sum x) -> non-integer, unit L2 norm, same planted structure
bank = [np.sqrt(b.astype(np.float32) / np.maximum(b.astype(np.float32).sum(1, keepdims=True), 1)).astype(np.float32) for b in ub]
pairs = sfm.select_pairs(n_img, 0, 0)
m.upload_bank(bank)
assert not m.bank_info()["u8_valued"]
m.set_profiling(True)
stream = torch.cuda.ExternalStream(m.stream)
ref = None
for name, eng in (("tf32x3", sfm.ENGINE_AUTO), ("simt_f32", sfm.ENGINE_SIMT)):
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); m.enqueue(pairs, sfm.NORM_L2, engine=eng); e1.record(stream); e1.synchronize()
        ms = e0.elapsed_time(e1)
        r = m.collect()
        prof = m.last_profile()
        print(f"{name} rep{rep}: {ms:.3f} ms for {len(pairs)} pairs -> {len(pairs)/ms*1e3:.1f} pairs/s, "
              f"{2*n_rows*n_rows*128*len(pairs)/ms/1e9:.1f} TFLOP/s-equiv, matches={int(r.offsets[-1])}, "
              f"knn {prof['knn_ms']:.3f} ms post {prof['post_ms']:.3f} ms, float_stats={m.float_stats() if eng == sfm.ENGINE_AUTO else None}")
    key = [(int(r.offsets[p]), int(r.offsets[p + 1])) for p in range(len(pairs))]
    sets = [set(zip(r[p]["queryIdx"].tolist(), r[p]["trainIdx"].tolist())) for p in range(len(pairs))]
    if ref is None:
        ref = sets
    else:
        diff = sum(len(a ^ b) for a, b in zip(ref, sets))
        print("symmetric difference of match lists tf32x3 vs simt_f32:", diff, "of", sum(len(a) for a in ref))
