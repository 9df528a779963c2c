"""N > 1 host logic on CPU: world_size-2 gloo run of the pair-list sharding + match-list gather.  The per-rank
compute is stood in for by the oracle (the GPU kernels are covered by the -m gpu tests); what is checked here
is that sharding + gather reproduce the single-process result byte for byte."""
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _shard():
    spec = importlib.util.spec_from_file_location("sfm_shard", os.path.join(ROOT, "sfm-mvs-pipeline_b200", "shard.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import workloads
    from oracle import oracle_np as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shard = _shard()
    sizes = (300, 257, 120, 400, 64, 0, 33)
    bank, prev = [], None
    for i, n in enumerate(sizes):
        d = workloads.sift_like_image(i, n, prev if prev is not None and len(prev) else None)
        bank.append(d)
        prev = d
    pairs = orc.pairs_unordered(len(sizes))
    mine = shard.assign_pairs(pairs, sizes, world)[rank]
    local = orc.match_pairs(bank, pairs[mine], orc.NORM_L2, min_match_count=5)
    counts = np.array([0 if m is None else len(m) for m in local], np.int64)
    dropped = np.array([m is None for m in local], np.uint8)
    matches = (np.concatenate([m for m in local if m is not None]) if any(m is not None for m in local)
               else np.zeros(0, orc.DMATCH_DTYPE))
    g = shard.gather_matches(mine, counts, matches, dropped, len(pairs), "cpu", 0)
    if rank == 0:
        full = orc.match_pairs(bank, pairs, orc.NORM_L2, min_match_count=5)
        offsets, allm, alld = g
        ok = True
        for p, m in enumerate(full):
            got = allm[offsets[p]:offsets[p + 1]]
            if m is None:
                ok = ok and alld[p] == 1 and len(got) == 0
            else:
                ok = ok and alld[p] == 0 and got.tobytes() == m.tobytes()
        q.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


def test_assign_pairs_balances_and_partitions():
    shard = _shard()
    rng = np.random.default_rng(0)
    sizes = rng.integers(0, 9000, size=40)
    pairs = np.array([(i, j) for i in range(40) for j in range(i + 1, 40)], np.int64)
    for world in (1, 2, 4, 8):
        parts = shard.assign_pairs(pairs, sizes, world)
        allidx = np.sort(np.concatenate(parts))
        assert np.array_equal(allidx, np.arange(len(pairs)))           # a partition
        cost = sizes[pairs[:, 0]] * sizes[pairs[:, 1]]
        loads = np.array([cost[p].sum() for p in parts], float)
        assert loads.max() <= loads.mean() + cost.max()                 # snake deal: within one max-cost pair
        assert all(np.all(np.diff(p) > 0) for p in parts if len(p) > 1)  # ascending inside a rank
    eq = shard.assign_pairs(pairs, np.full(40, 8192), 8)
    assert max(len(p) for p in eq) - min(len(p) for p in eq) <= 1


def test_world2_gloo_shard_and_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    assert all(p.exitcode == 0 for p in procs)
    assert q.get(timeout=5) is True
