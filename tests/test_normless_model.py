"""CPU property test of the norm-less path's decision logic (bounds, certificates, fallbacks) against the plain oracle.
The model (oracle/normless_model.py) restates csrc/post.cu's refine_dot_* decisions with exact integer dot products; the
CUDA code itself is checked on the GPU (tests/test_gpu_parity.py), this guards the MATH on inputs that are rare there."""
import numpy as np
import pytest

import workloads
from oracle import normless_model as nm
from oracle import oracle_np as orc
from oracle.oracle_np import NORM_L2


def _check(q, t, ratio, chunk, stats):
    exp = orc.match_pairs([q, t], [[0, 1]], NORM_L2, ratio=ratio)[0]
    got = nm.match_pair_normless(q, t, ratio, chunk, stats)
    assert orc.dmatch_equal(got, exp), ("norm-less", q.shape, t.shape, ratio, chunk)
    got = nm.match_pair_value(q, t, ratio, chunk, stats.setdefault("value", {}))
    assert orc.dmatch_equal(got, exp), ("norm K-step", q.shape, t.shape, ratio, chunk)


@pytest.mark.parametrize("chunk", [32, 64])
def test_sift_like_banks(chunk):
    stats = {}
    bank = workloads.sift_like_bank(4, 700)
    for a, b in ((1, 0), (0, 1), (2, 0), (3, 2), (0, 3)):
        for ratio in (0.7, 0.95):
            _check(bank[a], bank[b], ratio, chunk, stats)
    assert stats["stage_a"] > 100 and stats["rejected"] > 1000          # the cheap paths carry the load on SIFT-like data


@pytest.mark.parametrize("chunk", [32, 64])
def test_adversarial_and_wide_norm_spread(chunk):
    """Zero rows, duplicates, saturated rows, one- and two-row images, random bytes (norm spread ~100 %): the bounds are
    useless here, everything has to come out of stage B / the proved-fail rule / the brute-force fallback -- exactly."""
    stats = {}
    adv = workloads.adversarial_sift()
    rng = np.random.default_rng(5)
    rnd = rng.integers(0, 256, size=(300, 128), dtype=np.uint8)
    sparse = (rng.random((257, 128)) < 0.03).astype(np.uint8) * rng.integers(1, 255, size=(257, 128), dtype=np.uint8)
    tiled = np.concatenate([adv["base"][:40]] * 14)              # every chunk holds the same rows: chunk maxima tie
    sets = [adv["base"], adv["dup"], adv["zeros"], adv["sat"], adv["one"], adv["two"], adv["n129"], rnd, sparse, tiled]
    for i, q in enumerate(sets):
        for j, t in enumerate(sets):
            if i != j:
                for ratio in (0.7, 1.0):
                    _check(q[:120], t, ratio, chunk, stats)
    assert stats["brute"] > 0 and stats["stage_b"] > 0 and stats["proved_fail"] > 0
    assert stats["value"]["brute"] > 0 and stats["value"]["chunks"] > 0              # tied chunk maxima -> 'ambiguous' path


def test_random_small_images_fuzz():
    rng = np.random.default_rng(11)
    stats = {}
    for trial in range(25):
        nq, nt = int(rng.integers(1, 90)), int(rng.integers(1, 400))
        scale = int(rng.choice([8, 40, 255]))
        q = rng.integers(0, scale + 1, size=(nq, 128), dtype=np.uint8)
        t = rng.integers(0, scale + 1, size=(nt, 128), dtype=np.uint8)
        k = min(nq, nt) // 2
        if k:
            q[:k] = np.clip(t[rng.integers(0, nt, k)].astype(np.int32) + rng.integers(-2, 3, size=(k, 128)), 0, 255).astype(np.uint8)
        _check(q, t, float(rng.choice([0.5, 0.7, 0.9])), int(rng.choice([32, 64])), stats)
