import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import __graft_entry__ as ge, workloads
sfm = ge.load_package()
m = sfm.Matcher(0)
n_img, n_rows = 200, 8192
dev = torch.device("cuda", 0)
bank = torch.randint(0, 120, (n_img * n_rows, 128), dtype=torch.uint8, device=dev)
offs = [i * n_rows for i in range(n_img)]; rows = [n_rows] * n_img
for rep in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    m.upload_bank_device(bank.data_ptr(), offs, rows, 128, sfm.CV_8U)
    torch.cuda.synchronize(); print("adopt ms", (time.perf_counter() - t0) * 1e3)
host = torch.empty((25 * n_rows, 128), dtype=torch.float32, pin_memory=True); host.random_(0, 120)
hl = [host.numpy()[i * n_rows:(i + 1) * n_rows] for i in range(25)]
p = sfm.Matcher(0)
for rep in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    p.upload_bank(hl)
    torch.cuda.synchronize(); print("upload 1/8 slice (105 MB f32) ms", (time.perf_counter() - t0) * 1e3)
