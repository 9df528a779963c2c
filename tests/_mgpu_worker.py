"""Worker for tests/test_multi_gpu.py (launched with torch.distributed.run, one rank per GPU): replicate the bank
over NCCL, shard the pair list, gather on rank 0 and compare byte-for-byte with rank 0's own single-GPU run."""
import importlib.util
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
import workloads  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    sfm = ge.load_package()
    spec = importlib.util.spec_from_file_location("sfm_shard", os.path.join(ge.PKG_DIR, "shard.py"))
    shard = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(shard)
    sizes = [1500, 1024, 777, 2048, 300, 1300, 0, 640]
    total = sum(sizes)
    bank_dev = torch.zeros((max(total, 1), 128), dtype=torch.uint8, device=dev)
    if rank == 0:
        bank, prev = [], None
        for i, n in enumerate(sizes):
            d = workloads.sift_like_image(i, n, prev if prev is not None and len(prev) else None)
            bank.append(d)
            prev = d
        bank_dev.copy_(torch.from_numpy(np.concatenate(bank)))
    shard.broadcast_bank(bank_dev, 0)
    offs = np.concatenate([[0], np.cumsum(sizes)[:-1]]).tolist()
    m = sfm.Matcher(local)
    m.upload_bank_device(bank_dev.data_ptr(), offs, sizes, 128, sfm.CV_8U)
    pairs = sfm.select_pairs(len(sizes), 0, 0)[:-1]      # 27 pairs: ranks get unequal shares (padding in the gather)
    mine = shard.assign_pairs(pairs, sizes, world)[rank]
    res = m.match_pairs(pairs[mine], sfm.NORM_L2, min_match_count=20)
    g = shard.gather_matches(mine, res.counts(), res.matches, res.dropped, len(pairs), dev, 0)
    # the GPU-to-GPU gather of the device-resident result must give the same answer
    all_mine = shard.assign_pairs(pairs, sizes, world)
    m.enqueue(pairs[mine], sfm.NORM_L2, min_match_count=20)
    g2 = shard.gather_matches_device(m, mine, all_mine, len(pairs), dev, 0)
    # the native group (sfm_dist_*: NCCL inside the library): bank resident, and from host matrices (exchange path needs
    # >= 2 images per participant; ragged sizes, an empty image, unequal shares)
    uid = torch.zeros(sfm.DIST_ID_BYTES, dtype=torch.uint8, device=dev)
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(sfm.dist_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    m.dist_init(uid.cpu().numpy().tobytes(), rank, world)
    assert m.dist_info() == (rank, world)
    g3 = m.dist_match_pairs(pairs, sfm.NORM_L2, min_match_count=20)
    host_bank = []
    prev = None
    for i, n in enumerate(sizes):
        d = workloads.sift_like_image(i, n, prev if prev is not None and len(prev) else None)
        host_bank.append(d.astype(np.float32))
        prev = d
    g4 = m.dist_match_pairs_from_host(host_bank, pairs, sfm.NORM_L2, min_match_count=20)
    g5 = m.dist_match_pairs(pairs, sfm.NORM_L2, min_match_count=20, distinct=True)        # bank stays resident after from_host
    # a configuration the exchange path does not take (cross-check): every participant uploads the whole scene
    g6 = m.dist_match_pairs_from_host(host_bank, pairs, sfm.NORM_L2, k=1, cross_check=True)
    ok = True
    if rank != 0:
        ok = g3 is None and g4 is None and g5 is None and g6 is None
    if rank == 0:
        full = m.match_pairs(pairs, sfm.NORM_L2, min_match_count=20)
        for name, gd in (("dist_match_pairs", g3), ("dist_match_pairs_from_host", g4)):
            same = (np.array_equal(gd.offsets, full.offsets) and gd.matches.tobytes() == full.matches.tobytes()
                    and np.array_equal(gd.dropped, full.dropped))
            if not same:
                print("MISMATCH in", name, flush=True)
            ok = ok and same
        f5 = m.match_pairs(pairs, sfm.NORM_L2, min_match_count=20, distinct=True)
        f6 = m.match_pairs(pairs, sfm.NORM_L2, k=1, cross_check=True)
        for name, gd, f in (("distinct", g5, f5), ("cross-check", g6, f6)):
            same = np.array_equal(gd.offsets, f.offsets) and gd.matches.tobytes() == f.matches.tobytes()
            if not same:
                print("MISMATCH in", name, int(gd.offsets[-1]), int(f.offsets[-1]), flush=True)
            ok = ok and same
        for name, gg in (("torch gather (host)", g), ("torch gather (device)", g2)):
            same = (np.array_equal(gg[0], full.offsets) and gg[1].tobytes() == full.matches.tobytes()
                    and np.array_equal(gg[2], full.dropped))
            if not same:
                print("MISMATCH in", name, flush=True)
            ok = ok and same
        print("MGPU_IDENTICAL" if ok else "MGPU_MISMATCH", int(full.offsets[-1]), flush=True)
    m.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
