"""CPU-only checks of the boundary: the C-ABI library loads, exports every symbol include/sfmmatch.h declares,
its host-side logic (pair selection) equals the oracle, and compute entry points refuse to run without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import oracle_np as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "sfmmatch.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sfm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(sfm):
    lib = C.CDLL(sfm.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/sfmmatch.h but not exported"
    assert sorted(sfm.EXPORTS) == names


def test_struct_layouts(sfm):
    assert sfm.DMATCH_DTYPE.itemsize == 16 and sfm.DMATCH_DTYPE == orc.DMATCH_DTYPE   # == cv::DMatch
    assert C.sizeof(sfm.Opts) == 32
    o = sfm.Opts()
    C.CDLL(sfm.LIB_PATH).sfm_opts_default(C.byref(o), C.c_int32(4))
    assert (o.norm, o.k, o.ratio, o.cross_check, o.distinct, o.min_match_count, o.engine) == (4, 2, 0.7, 0, 0, 0, 0)
    # homography stage: cv::findHomography's defaults (maxIters 2000, confidence 0.995), refinement on
    assert C.sizeof(sfm.HomographyOpts) == 24
    h = sfm.HomographyOpts()
    lib = C.CDLL(sfm.LIB_PATH)
    lib.sfm_homography_opts_default.restype = None
    lib.sfm_homography_opts_default(C.byref(h))
    assert (h.max_iters, h.refine, h.confidence, h.seed) == (2000, 1, 0.995, 0)


@pytest.mark.parametrize("n,seq,grid", [(0, 0, 0), (1, 0, 0), (7, 0, 0), (200, 0, 0), (3, 2, 0), (3, 3, 0), (50, 5, 0),
                                        (1000, 3, 40), (1000, 2, 40), (1000, 4, 40), (20, 3, 5), (3, 2, 2), (3, 2, 3),
                                        (3, 2, 1), (3, 3, 3), (17, 4, 4), (9, 1, 0), (9, -3, 2), (12, 2, 12), (12, 9, 5)])
def test_select_pairs_equals_oracle(sfm, n, seq, grid):
    got = sfm.select_pairs(n, seq, grid)
    exp = orc.select_pairs(n, seq, grid)
    assert got.tolist() == exp.tolist()


def test_grid_known_answer_through_the_abi(sfm):
    # GridFeatureMatchingStrategy.h:31-39
    p = sfm.select_pairs(20, 3, 5)
    assert sorted(int(b) + 1 for a, b in p if a == 0) == [2, 3, 6, 7, 11]


def test_c_oracle_pair_lists_agree():
    from oracle import oracle_c as oc
    assert oc.pairs_grid(1000, 3, 40).tolist() == orc.pairs_grid(1000, 3, 40).tolist()
    assert oc.pairs_video(50, 5).tolist() == orc.pairs_video(50, 5).tolist()
    assert oc.pairs_unordered(30).tolist() == orc.pairs_unordered(30).tolist()


def test_c_oracle_matches_numpy_oracle(insel_sift):
    from oracle import oracle_c as oc
    import workloads
    idx, dist = oc.knn2(insel_sift["desc0"], insel_sift["desc1"], 4)
    assert np.array_equal(idx, insel_sift["p01_nidx"])
    assert np.array_equal(dist.view(np.uint32), insel_sift["p01_dist"].view(np.uint32))
    ob = workloads.orb_like_bank(2, 500)
    i1, d1 = oc.knn2(ob[0], ob[1], 6)
    i2, d2 = orc.knn2_hamming(ob[0], ob[1])
    assert np.array_equal(i1, i2) and np.array_equal(d1, d2)
    bank = workloads.sift_like_bank(3, 400)
    a = oc.match_pairs(bank, orc.pairs_unordered(3), 4)
    b = orc.match_pairs(bank, orc.pairs_unordered(3), 4)
    assert all(orc.dmatch_equal(x, y) for x, y in zip(a, b))


def test_no_cpu_fallback_without_gpu(sfm):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(sfm.SfmError) as e:
        sfm.Matcher(0)
    assert e.value.code == sfm.ERR_CUDA


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "sfm-mvs-pipeline_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) or f == "Makefile":
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.lower(), f"{f} mentions the oracle: the product must not depend on it"


def test_host_match_index_equals_linear_scans():
    """hash indices over the ShotMatches (host/match_index.h) answer Scene::addShotMatches,
    find3d2dMatches and mergePointcloudElement3d2d's lookups exactly like the reference's linear scans (restated in
    host/test_match_index.cpp; random scenes with duplicated points, both orientations, -0 and NaN coordinates)."""
    import subprocess
    pkg = os.path.join(ROOT, "sfm-mvs-pipeline_b200")
    subprocess.check_call(["make", "-s", "-C", pkg, "test_match_index"])
    for seed in (1, 2, 3):
        out = subprocess.run([os.path.join(pkg, "test_match_index"), str(seed)], capture_output=True, text=True, timeout=120)
        assert out.returncode == 0 and out.stdout.startswith("OK "), out.stdout + out.stderr
        assert int(out.stdout.split()[1]) > 10000


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the reference's CPU matcher through cv2) prints ONE JSON line with the contract's keys;
    small shape so that the CPU suite stays fast."""
    import json
    import subprocess
    import sys
    from oracle import cv2_ref
    if not cv2_ref.available():
        pytest.skip("cv2 not importable")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--images", "6", "--rows", "512",
                          "--steps", "2", "--warmup", "1", "--cpu-sample-pairs", "8"], capture_output=True, text=True, timeout=300)
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout + out.stderr
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "pairs/s" and d["value"] > 0 and d["vs_baseline"] is None
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1


def test_bench_extract_reference_arm_contract():
    """The secondary bench line of the feature-extraction stage: `bench.py --workload extract --impl reference` (cv2's SIFT on
    the host threads) prints ONE JSON line with the contract's keys."""
    import json
    import subprocess
    import sys
    from oracle import cv2_ref
    if not cv2_ref.available():
        pytest.skip("cv2 not importable")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "extract", "--impl", "reference", "--images", "3",
                          "--photo", "120x160", "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout + out.stderr
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "images/s" and d["value"] > 0 and "SIFT" in d["metric"]


def test_gray_conversion_equals_cvtcolor(sfm):
    """sfm_gray_from_bgr (host arithmetic in front of the extractor) = cv2.cvtColor(BGR2GRAY) on 8-bit data = the restatement,
    on every 5th value of each channel plus random images; RGB order, 4 channels and strided input included."""
    from oracle import sift_np as S
    vals = np.arange(0, 256, 5, dtype=np.uint8)
    grid = np.stack(np.meshgrid(vals, vals, vals, indexing="ij"), -1).reshape(-1, 52, 3)
    rng = np.random.default_rng(0)
    rnd = rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)
    for img in (grid, rnd):
        got = sfm.gray_from_bgr(img)
        assert np.array_equal(got, S.bgr_to_gray(img))
        try:
            import cv2
            assert np.array_equal(got, cv2.cvtColor(img, cv2.COLOR_BGR2GRAY))
        except ImportError:
            pass
    assert np.array_equal(sfm.gray_from_bgr(rnd[..., ::-1], rgb_order=True), sfm.gray_from_bgr(rnd))
    rgba = np.concatenate([rnd, np.full(rnd.shape[:2] + (1,), 7, np.uint8)], -1)
    assert np.array_equal(sfm.gray_from_bgr(rgba), sfm.gray_from_bgr(rnd))
    wide = rng.integers(0, 256, (20, 64, 3), dtype=np.uint8)
    assert np.array_equal(sfm.gray_from_bgr(wide[:, 8:40]), sfm.gray_from_bgr(np.ascontiguousarray(wide[:, 8:40])))
    assert sfm.gray_from_bgr(np.zeros((0, 5, 3), np.uint8)).shape == (0, 5)
    with pytest.raises(sfm.SfmError):
        sfm.gray_from_bgr(np.zeros((4, 4), np.uint8))


def test_native_deal_equals_the_python_deal(sfm):
    """sfm_dist_assign_pairs (csrc/dist.cu) and shard.assign_pairs are the same deal (no GPU involved)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("sfm_shard", os.path.join(os.path.dirname(sfm.LIB_PATH), "shard.py"))
    shard = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(shard)
    rng = np.random.default_rng(5)
    for world in (1, 2, 3, 8):
        for n_img in (2, 9, 33):
            for equal in (False, True):
                n_rows = np.full(n_img, 8192) if equal else rng.integers(0, 5000, n_img)
                pairs = sfm.select_pairs(n_img, 0, 0)
                owner = sfm.dist_assign_pairs(pairs, n_rows, world)
                ref = shard.assign_pairs(pairs, n_rows, world)
                for r in range(world):
                    assert np.array_equal(np.nonzero(owner == r)[0], ref[r])
    # upload shares: contiguous, cover every image once, balanced by padded rows
    n_rows = rng.integers(0, 9000, 57)
    prev_end = 0
    for r in range(5):
        lo, hi = sfm.dist_upload_share(n_rows, 5, r)
        assert lo == prev_end and hi >= lo
        prev_end = hi
    assert prev_end == 57
