"""Wall-clock breakdown of one end-to-end step (development aid): upload / enqueue+device / collect."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import __graft_entry__ as ge
import workloads

sfm = ge.load_package()
n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 200
n_rows = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
m = sfm.Matcher(0)
host = torch.empty((n_img * n_rows, 128), dtype=torch.float32, pin_memory=True)
bank = workloads.sift_like_bank(min(n_img, 16), n_rows)
hn = host.numpy()
for i in range(n_img):
    hn[i * n_rows:(i + 1) * n_rows] = bank[i % len(bank)]
lst = [hn[i * n_rows:(i + 1) * n_rows] for i in range(n_img)]
pairs = sfm.select_pairs(n_img, 0, 0)
for rep in range(4):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); m.upload_bank(lst); torch.cuda.synchronize(); t1 = time.perf_counter()
    m.enqueue(pairs, sfm.NORM_L2); t2 = time.perf_counter(); torch.cuda.synchronize(); t3 = time.perf_counter()
    r = m.collect(); t4 = time.perf_counter()
    print(f"rep{rep}: upload {1e3*(t1-t0):.1f} ms | enqueue host {1e3*(t2-t1):.1f} ms | device wait {1e3*(t3-t2):.1f} ms | "
          f"collect {1e3*(t4-t3):.1f} ms | total {1e3*(t4-t0):.1f} ms | matches {int(r.offsets[-1])}")
u8 = [x.astype(np.uint8) for x in lst[:n_img]]
t0 = time.perf_counter(); m.upload_bank(u8); torch.cuda.synchronize(); print("upload pageable u8 ms", 1e3 * (time.perf_counter() - t0))
