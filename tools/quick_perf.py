"""Quick device-side timing of the matching kernels on a small all-pairs workload (development aid)."""
import os, sys, time
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
bank = workloads.sift_like_bank(n_img, n_rows)
pairs = sfm.select_pairs(n_img, 0, 0)
t0 = time.time(); m.upload_bank([b.astype(np.float32) for b in bank]); print("upload f32 s", time.time() - t0)
t0 = time.time(); m.upload_bank(bank); print("upload u8 s", time.time() - t0)
stream = torch.cuda.ExternalStream(m.stream)
for name, eng in (("tensor", sfm.ENGINE_TENSOR), ("simt", sfm.ENGINE_SIMT)):
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        m.enqueue(pairs, sfm.NORM_L2, engine=eng)
        e1.record(stream)
        e1.synchronize()
        ms = e0.elapsed_time(e1)
        r = m.collect()
        print(f"{name} rep{rep}: {ms:.3f} ms for {len(pairs)} pairs -> {len(pairs)/ms*1e3:.1f} pairs/s, "
              f"{2*n_rows*n_rows*128*len(pairs)/ms/1e9:.1f} TOP/s, matches={int(r.offsets[-1])}")
ob = workloads.orb_like_bank(4, 30000)
m.upload_bank(ob)
op = sfm.select_pairs(4, 2, 0)
for name, eng in (("tensor", sfm.ENGINE_TENSOR), ("popc", sfm.ENGINE_SIMT)):
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); m.enqueue(op, sfm.NORM_HAMMING, engine=eng); e1.record(stream); e1.synchronize()
        ms = e0.elapsed_time(e1); r = m.collect()
        print(f"orb {name} rep{rep}: {ms:.3f} ms for {len(op)} pairs (30000x30000) -> {len(op)/ms*1e3:.1f} pairs/s "
              f"matches={int(r.offsets[-1])}")
