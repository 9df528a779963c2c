"""Pair-list sharding across GPUs and gathering of match lists (one process per GPU, torch.distributed).

The stage shards naturally: every image pair is independent (the reference already runs them as an OpenMP
``parallel for`` over pairs, UnorderedFeatureMatchingStrategy.cpp:40).  Each rank holds a replica of the
descriptor bank, takes a cost-balanced share of the pair list, and only the compacted match lists travel:
NCCL (or gloo in the CPU tests) is used to broadcast the bank and to gather (counts, matches) on rank 0.
No data-path collective runs while the kernels work.

This module is the torch.distributed form of the group, kept for launchers that already own a process group and for
the gloo tests; the product path is native (csrc/dist.cu: sfm_dist_* / sfm_mgpu_*, NCCL inside the library), and
``assign_pairs`` here is the same deal as ``sfm_dist_assign_pairs`` (tests/test_cabi_cpu.py checks they agree).
"""
from __future__ import annotations

import numpy as np


def assign_pairs(pairs: np.ndarray, n_rows, world_size: int):
    """Deal pairs to ranks by cost Nq*Nt: sort by cost (descending, stable) and deal round-robin in a
    snake order, then restore ascending pair order inside each rank (keeps pairs that share the left image
    adjacent, which keeps the train images L2-resident).  Equal-cost lists degenerate to a strided deal.
    Returns a list of int64 index arrays (positions in ``pairs``), one per rank."""
    pairs = np.asarray(pairs, np.int64).reshape(-1, 2)
    n_rows = np.asarray(n_rows, np.int64)
    if world_size <= 1 or len(pairs) == 0:
        return [np.arange(len(pairs), dtype=np.int64)] + [np.zeros(0, np.int64) for _ in range(max(0, world_size - 1))]
    cost = n_rows[pairs[:, 0]] * n_rows[pairs[:, 1]]
    order = np.argsort(-cost, kind="stable")
    k = np.arange(len(order))
    rnd, pos = k // world_size, k % world_size
    rank_of = np.where(rnd % 2 == 0, pos, world_size - 1 - pos)      # snake: 0..W-1, W-1..0, ...
    out = []
    for r in range(world_size):
        out.append(np.sort(order[rank_of == r]).astype(np.int64))
    return out


def broadcast_bank(bank_tensor, src: int = 0, group=None):
    """Broadcast the packed descriptor bank (a torch tensor, on the GPU under NCCL) from ``src``."""
    import torch.distributed as dist
    dist.broadcast(bank_tensor, src=src, group=group)
    return bank_tensor


def gather_matches(local_idx: np.ndarray, counts: np.ndarray, matches: np.ndarray, dropped: np.ndarray,
                   n_pairs_total: int, device, dst: int = 0, group=None):
    """Gather per-rank results on ``dst`` and reassemble them in global pair order.

    local_idx : positions (in the global pair list) of this rank's pairs, ascending
    counts    : matches per local pair;  matches: concatenated DMatch records (structured, 16 B)
    Returns (offsets[n_pairs_total+1], matches, dropped) on ``dst`` and None elsewhere."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n_local, n_match = int(len(local_idx)), int(len(matches))
    sizes = torch.tensor([n_local, n_match], dtype=torch.int64, device=device)
    all_sizes = [torch.zeros(2, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    all_sizes = torch.stack(all_sizes).cpu().numpy()
    max_local, max_match = int(all_sizes[:, 0].max()), int(all_sizes[:, 1].max())
    # one int32 payload per rank: [idx | counts | dropped | matches as 4 x int32]
    width = 3 * max_local + 4 * max_match
    payload = np.zeros(max(width, 1), np.int32)
    payload[:n_local] = local_idx
    payload[max_local:max_local + n_local] = counts
    payload[2 * max_local:2 * max_local + n_local] = dropped
    if n_match:
        payload[3 * max_local:3 * max_local + 4 * n_match] = np.ascontiguousarray(matches).view(np.int32).reshape(-1)
    t = torch.from_numpy(payload).to(device)
    gathered = [torch.zeros_like(t) for _ in range(world)] if rank == dst else None
    dist.gather(t, gathered, dst=dst, group=group)
    if rank != dst:
        return None
    total_counts = np.zeros(n_pairs_total, np.int64)
    total_dropped = np.zeros(n_pairs_total, np.uint8)
    per_rank = []
    for r in range(world):
        buf = gathered[r].cpu().numpy()
        nl, nm = int(all_sizes[r, 0]), int(all_sizes[r, 1])
        idx = buf[:nl].astype(np.int64)
        cnt = buf[max_local:max_local + nl].astype(np.int64)
        total_counts[idx] = cnt
        total_dropped[idx] = buf[2 * max_local:2 * max_local + nl].astype(np.uint8)
        m = buf[3 * max_local:3 * max_local + 4 * nm].copy().view(_dmatch_dtype())
        per_rank.append((idx, cnt, m))
    offsets = np.zeros(n_pairs_total + 1, np.int64)
    np.cumsum(total_counts, out=offsets[1:])
    out = np.zeros(int(offsets[-1]), _dmatch_dtype())
    for idx, cnt, m in per_rank:
        if len(m) == 0:
            continue
        src_off = np.zeros(len(cnt) + 1, np.int64)
        np.cumsum(cnt, out=src_off[1:])
        # destination of every record: start of its pair in the global list + position inside the pair
        dst = np.repeat(offsets[idx] - src_off[:-1], cnt) + np.arange(len(m), dtype=np.int64)
        out[dst] = m
    return offsets, out, total_dropped


def cuda_view(ptr: int, nbytes: int, device, dtype=None):
    """torch view (no copy) of library-owned device memory."""
    import torch

    class _Holder:
        pass
    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (max(nbytes, 0),), "typestr": "|u1", "data": (ptr, False), "version": 3}
    t = torch.as_tensor(h, device=device) if nbytes > 0 else torch.zeros(0, dtype=torch.uint8, device=device)
    return t if dtype is None else t.view(dtype)


_PINNED = {}
LAST_D2H_BYTES = 0        # bytes of the last gather_matches_device's device-to-host copies (bench.py's e2e accounting)


def _pinned(name, shape, dtype):
    """Reusable page-locked host buffer (device-to-host copies into pageable memory run at a fraction of PCIe speed)."""
    import torch
    n = int(np.prod(shape))
    buf = _PINNED.get(name)
    if buf is None or buf.dtype != dtype or buf.numel() < n:
        buf = torch.empty(max(n, 1), dtype=dtype, pin_memory=True)
        _PINNED[name] = buf
    return buf[:n].view(*shape)


def gather_matches_device(matcher, local_idx, all_local_idx, n_pairs_total: int, device, dst: int = 0, group=None):
    """GPU-to-GPU gather of the last enqueue's match lists (NCCL): per-rank device views from the library
    (sfm_match_pairs_device_view) -> dist.gather of padded int32 payloads -> reassembly in global pair order on the
    destination GPU (one scatter for all ranks) -> ONE device-to-host copy into pinned memory on ``dst``.
    ``all_local_idx`` = assign_pairs(...) for every rank (known everywhere, so only the match totals need an all_gather).
    Returns (offsets, matches, dropped) on ``dst`` — numpy views of reusable page-locked buffers, valid until the next
    call — and None elsewhere."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    pm, po, pd, n_local, n_match = matcher.device_view()
    lib_stream = torch.cuda.ExternalStream(matcher.stream, device=device)
    torch.cuda.current_stream(device).wait_stream(lib_stream)          # the views below are written on the library stream
    tot = torch.tensor([n_match], dtype=torch.int64, device=device)
    all_tot_t = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(all_tot_t, tot, group=group)
    all_tot = all_tot_t.cpu().numpy()
    max_local = max(len(x) for x in all_local_idx)
    max_match = int(all_tot.max())
    width = max(2 * max_local + 4 * max_match, 1)
    off = cuda_view(po, n_local * 8, device, torch.int64)
    payload = torch.empty(width, dtype=torch.int32, device=device)
    payload[:2 * max_local].zero_()                # ranks with fewer pairs than max_local: padded counts must be 0
    payload[:n_local] = torch.diff(off, append=tot)
    payload[max_local:max_local + n_local] = cuda_view(pd, n_local, device)
    payload[2 * max_local:2 * max_local + 4 * n_match] = cuda_view(pm, n_match * 16, device, torch.int32)
    gathered = torch.empty((world, width), dtype=torch.int32, device=device) if rank == dst else None
    dist.gather(payload, list(gathered.unbind(0)) if rank == dst else None, dst=dst, group=group)
    # the payload copies read library-owned buffers on torch's stream: the library's next enqueue must not overwrite them
    lib_stream.wait_stream(torch.cuda.current_stream(device))
    if rank != dst:
        return None
    # global pair index of every (rank, local position), padded positions point at a dummy slot (rebuilt on every call:
    # the deal depends on the row counts, not only on the pair count)
    idx = np.full((world, max_local), n_pairs_total, np.int64)
    for r in range(world):
        idx[r, :len(all_local_idx[r])] = all_local_idx[r]
    idx_dev = torch.from_numpy(idx).to(device)
    cnt = gathered[:, :max_local].to(torch.int64)                       # [world, max_local], 0 in padded positions
    total_counts = torch.zeros(n_pairs_total + 1, dtype=torch.int64, device=device)
    total_counts[idx_dev.reshape(-1)] = cnt.reshape(-1)
    total_dropped = torch.zeros(n_pairs_total + 1, dtype=torch.uint8, device=device)
    total_dropped[idx_dev.reshape(-1)] = gathered[:, max_local:2 * max_local].reshape(-1).to(torch.uint8)
    offsets = torch.zeros(n_pairs_total + 1, dtype=torch.int64, device=device)
    offsets[1:] = torch.cumsum(total_counts[:n_pairs_total], 0)
    # destination of every gathered record: start of its pair in the global list + position inside the pair
    src_off = torch.cumsum(cnt, 1) - cnt                                # start of each local pair in its rank's records
    n_all = int(all_tot.sum())
    shift = (offsets[idx_dev.clamp(max=n_pairs_total - 1)] - src_off).reshape(-1)      # per (rank, local pair)
    rec_rank_pos = torch.repeat_interleave(torch.arange(world * max_local, device=device), cnt.reshape(-1), output_size=n_all)
    within = torch.cat([torch.arange(int(all_tot[r]), device=device) for r in range(world)]) if n_all else torch.zeros(0, dtype=torch.int64, device=device)
    dst_idx = shift[rec_rank_pos] + within
    recs = torch.cat([gathered[r, 2 * max_local:2 * max_local + 4 * int(all_tot[r])] for r in range(world)]).view(n_all, 4)
    out = torch.empty((n_all, 4), dtype=torch.int32, device=device)
    out[dst_idx] = recs
    h_out = _pinned("matches", (n_all, 4), torch.int32)
    h_off = _pinned("offsets", (n_pairs_total + 1,), torch.int64)
    h_drop = _pinned("dropped", (n_pairs_total,), torch.uint8)
    h_out.copy_(out, non_blocking=True)
    h_off.copy_(offsets, non_blocking=True)
    h_drop.copy_(total_dropped[:n_pairs_total], non_blocking=True)
    torch.cuda.current_stream(device).synchronize()
    global LAST_D2H_BYTES
    LAST_D2H_BYTES = int(h_out.numel() * 4 + h_off.numel() * 8 + h_drop.numel() + all_tot_t.numel() * 8)
    return h_off.numpy(), h_out.numpy().view(_dmatch_dtype()).reshape(-1), h_drop.numpy()


def _dmatch_dtype():
    return np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"), ("distance", "<f4")])
