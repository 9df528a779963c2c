"""ORB / Hamming engines on an ORB-like bank (30 000 x 256-bit descriptors per image, run-orb-sequence.sh's feature limit),
sequence pairing: device time per engine + achieved bytes/popcounts (development aid and the ncu target for K2)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as ge
import workloads

sfm = ge.load_package()
n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 12
n_rows = int(sys.argv[2]) if len(sys.argv) > 2 else 30000
seq = int(sys.argv[3]) if len(sys.argv) > 3 else 3
m = sfm.Matcher(0)
bank = workloads.orb_like_bank(n_img, n_rows)
m.upload_bank(bank)
pairs = sfm.select_pairs(n_img, seq, 0)
m.set_profiling(True)
ref = None
for name, eng in (("tensor (tcgen05 on bit-expanded rows)", sfm.ENGINE_TENSOR), ("popc (knn2_hamming_popc)", sfm.ENGINE_SIMT)):
    best = None
    for rep in range(4):
        m.enqueue(pairs, sfm.NORM_HAMMING, engine=eng)
        pr = m.last_profile()
        r = m.collect()
        if rep and (best is None or pr["knn_ms"] < best):
            best = pr["knn_ms"]
    per_pair = best / len(pairs)
    print(f"{name}: knn {best:.3f} ms for {len(pairs)} pairs of {n_rows} x {n_rows} -> {len(pairs)/best*1e3:.1f} pairs/s; "
          f"popcount-equivalents {8*n_rows*n_rows/per_pair/1e9:.2f} T/s; algorithmic bytes 32*(Nq+Nt) = {64*n_rows/1e6:.2f} MB/pair "
          f"-> {64*n_rows/per_pair/1e6:.2f} GB/s; matches={int(r.offsets[-1])}")
    lists = [r[p].tobytes() for p in range(len(pairs))]
    if ref is None:
        ref = lists
    else:
        print("byte-identical match lists:", ref == lists)
