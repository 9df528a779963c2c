"""host/cv_adapter.h — the cv::DescriptorMatcher a maintainer of the reference would inject (INTEGRATION.md level 1) —
compiled against tests/stubs/opencv2 and driven by tests/host_adapter_harness.cpp from an OpenMP loop over pairs on ONE
shared matcher, like UnorderedFeatureMatchingStrategy.cpp:40-91 (knnMatch(k=2), ratio filter, cv::Exception -> match())."""
import os
import re
import subprocess

import numpy as np
import pytest

import workloads
from oracle import oracle_np as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = os.path.join(ROOT, "tests", "host_adapter_harness")


def test_adapter_compiles_against_the_opencv_stub():
    """No GPU needed: the adapter header goes through a compiler (build() makes the harness)."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "sfm-mvs-pipeline_b200"), "../tests/host_adapter_harness"])
    assert os.path.exists(HARNESS)
    src = open(os.path.join(ROOT, "sfm-mvs-pipeline_b200", "host", "cv_adapter.h")).read()
    # a clone must not create CUDA objects: OpenCV clones the matcher on every two-argument knnMatch
    clone = src[src.index("clone(bool emptyTrainData"):src.index("protected:")]
    assert "sfm_ctx_create" not in clone


def _write_scene(path, bank, norm, depth):
    with open(path, "wb") as f:
        cols = bank[0].shape[1] if len(bank) else 0
        np.array([norm, depth, cols, len(bank)], np.int32).tofile(f)
        np.array([len(b) for b in bank], np.int32).tofile(f)
        for b in bank:
            np.ascontiguousarray(b).tofile(f)


def _read_result(path):
    raw = open(path, "rb").read()
    n = int(np.frombuffer(raw, np.int64, 1)[0])
    off, out = 8, []
    dm = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"), ("distance", "<f4")])
    for _ in range(n):
        l, r = np.frombuffer(raw, np.int32, 2, off)
        cnt = int(np.frombuffer(raw, np.int64, 1, off + 8)[0])
        out.append(((int(l), int(r)), np.frombuffer(raw, dm, cnt, off + 16).copy()))
        off += 16 + 16 * cnt
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["sift_f32", "orb"])
def test_openmp_loop_through_the_adapter_equals_the_oracle(tmp_path, kind):
    if kind == "sift_f32":
        sizes = [700, 517, 300, 1, 1030, 0]
        bank = workloads.sift_like_bank(len(sizes), 1100)
        bank = [b[:n] for b, n in zip(bank, sizes)]
        scene, norm, depth = [b.astype(np.float32) for b in bank], orc.NORM_L2, 5
    else:
        bank = workloads.orb_like_bank(4, 900)
        scene, norm, depth = bank, orc.NORM_HAMMING, 0
    inp, outp = tmp_path / "scene.bin", tmp_path / "out.bin"
    _write_scene(inp, scene, norm, depth)
    r = subprocess.run([HARNESS, str(inp), str(outp), "6"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    got = _read_result(outp)
    pairs = orc.select_pairs(len(bank), 0, 0)
    assert [g[0] for g in got] == [tuple(p) for p in pairs.tolist()]
    n_empty_pairs = 0
    for (l, rr), m in got:
        if len(bank[l]) == 0 or len(bank[rr]) == 0:
            # an empty Mat makes knnMatch AND the match() fallback throw, like cv::batchDistance:
            # the harness records the pair as failed where the reference would terminate
            n_empty_pairs += 1
            assert len(m) == 0
            continue
        exp = orc.match_pairs(bank, [[l, rr]], norm)[0]
        assert orc.dmatch_equal(m, exp), (l, rr)
    stats = dict(re.findall(r"(\w+) (\d+)", r.stdout))
    assert int(stats["failed"]) == n_empty_pairs and int(stats["fallbacks"]) == n_empty_pairs
    assert sum(len(m) for _, m in got) > 100
