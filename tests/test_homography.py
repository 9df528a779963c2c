"""Homography stage (SfM::calculateHomography, SfM.cpp:599-637).

CPU tests: the numpy restatement of the GPU algorithm (oracle/homography_np.py) against the cv2.findHomography golden
numbers (tests/golden/insel_homography.npz).  GPU tests (-m gpu): csrc/homography.cu through the C ABI against the
restatement (same hypotheses -> same inlier counts) and against the cv2 golden numbers.

Stated tolerance against cv::findHomography (different random minimal sets, so no bit parity): the inlier RATIO agrees
within 0.03 absolute on every fixture, for every seed tried.  Both run the same procedure — RANSAC over 4-point models
with OpenCV's early-stopping rule, least-squares re-estimation on the consensus set, mask of the re-estimated model —
and the re-estimation makes the final count nearly independent of which good minimal set was drawn (insel ORB 0-1:
6076..6087 over six seeds vs cv2's 6074 of 6170)."""
import os

import numpy as np
import pytest

from oracle import homography_np as hn
from oracle import oracle_np as orc
from oracle.oracle_np import NORM_HAMMING, NORM_L2

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RATIO_TOL = 0.03
PAIRS = ((0, 1), (0, 2), (1, 2))


@pytest.fixture(scope="module")
def hom():
    return dict(np.load(os.path.join(GOLDEN, "insel_homography.npz")))


def test_oracle_vs_cv2_synthetic(hom):
    for k in range(5):
        p1, p2 = hom[f"syn{k}_p1"], hom[f"syn{k}_p2"]
        pts = np.concatenate([p1, p2], 1)
        cv = int(hom[f"syn{k}_count"])
        for seed in (7, 8, 9):
            cnt, hyp, mask = hn.ransac_inliers(pts, 3.0, pair=k, seed=seed)
            assert abs(cnt - cv) / len(pts) <= RATIO_TOL, (k, seed, cnt, cv)
        # the consensus set contains (almost) only planted inliers
        planted = hom[f"syn{k}_planted"]
        assert (mask & ~planted).sum() <= max(2, 0.02 * len(pts))
        assert hyp >= 0 and mask.sum() == cnt


def test_oracle_vs_cv2_insel(hom, insel_sift, insel_orb):
    for tag, gold in (("sift", insel_sift), ("orb", insel_orb)):
        for a, b in PAIRS:
            good = gold[f"p{a}{b}_good"]
            pts = hn.aligned_points(hom[f"{tag}_kp{a}"], hom[f"{tag}_kp{b}"], good)
            cv = int(hom[f"{tag}_h{a}{b}_count"])
            for seed in (1, 2, 3):
                cnt, _, _ = hn.ransac_inliers(pts, 3.0, pair=0, seed=seed)
                assert abs(cnt - cv) / len(good) <= RATIO_TOL, (tag, a, b, seed, cnt, cv)
            # without the re-estimation the count is the consensus of one minimal model: never larger than the exhaustive one
            c_min = hn.ransac_inliers(pts, 3.0, pair=0, seed=1, refine=False)[0]
            c_all = hn.ransac_inliers(pts, 3.0, pair=0, seed=1, refine=False, confidence=1.0)[0]
            assert 4 <= c_min <= c_all <= len(good)


def test_oracle_degenerate_inputs():
    assert hn.ransac_inliers(np.zeros((3, 4), np.float32), 3.0, 0)[0] == -1          # < 4 matches: no homography
    # all points identical / collinear: every minimal set is rejected -> 0 inliers, no model
    same = np.ones((10, 4), np.float32)
    assert hn.ransac_inliers(same, 3.0, 0)[0] == 0
    line = np.stack([np.arange(20), np.arange(20), np.arange(20), np.arange(20)], 1).astype(np.float32)
    assert hn.ransac_inliers(line, 3.0, 0)[0] == 0
    assert hn.ransac_inliers(np.random.default_rng(0).uniform(0, 99, (4, 4)).astype(np.float32), 3.0, 0)[0] == 4   # exactly 4: no RANSAC
    # generator: indices in range and distinct
    idx, ok = hn.minimal_sets(3, 5, 500, 7)
    assert ok.all() and idx.min() >= 0 and idx.max() < 7
    assert all(len(set(r)) == 4 for r in idx.tolist())


# ------------------------------------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def matcher(sfm):
    m = sfm.Matcher(0)
    yield m
    m.close()


@pytest.mark.gpu
@pytest.mark.parametrize("tag,norm", [("sift", NORM_L2), ("orb", NORM_HAMMING)])
def test_gpu_insel_pairs(sfm, matcher, hom, insel_sift, insel_orb, tag, norm):
    """BASELINE configs C1 / C2 carried through the next stage: match on the GPU, homography on the device-resident lists."""
    gold = insel_sift if tag == "sift" else insel_orb
    bank = [gold[f"desc{i}"] for i in range(3)]
    kps = [hom[f"{tag}_kp{i}"] for i in range(3)]
    pairs = np.array(PAIRS, np.int32)
    matcher.upload_bank(bank)
    matcher.upload_keypoints(kps)
    res = matcher.match_pairs(pairs, norm)
    for seed, refine, conf in ((11, True, 0.995), (12, True, 0.995), (11, False, 0.995), (11, False, 1.0)):
        r = matcher.homography_inlier_ratios(3.0, 2000, seed=seed, refine=refine, confidence=conf)
        for p, (a, b) in enumerate(PAIRS):
            good = gold[f"p{a}{b}_good"]
            assert orc.dmatch_equal(res[p], good)
            # the restatement of the same algorithm: identical counts and winning hypothesis
            pts = hn.aligned_points(kps[a], kps[b], good)
            cnt, h, _, rcnt = hn.ransac_inliers(pts, 3.0, pair=p, seed=seed, refine=refine, confidence=conf, details=True)
            assert (int(r["ransac_inliers"][p]), int(r["hypothesis"][p])) == (rcnt, h)
            assert int(r["inliers"][p]) == cnt
            assert r["ratio"][p] == cnt / len(good)
            if refine:      # cv::findHomography golden number
                cv = int(hom[f"{tag}_h{a}{b}_count"])
                assert abs(r["ratio"][p] - cv / len(good)) <= RATIO_TOL, (tag, a, b, int(r["inliers"][p]), cv)


@pytest.mark.gpu
def test_gpu_keypoint_stride_thresholds_and_skips(sfm, matcher, hom, insel_sift):
    bank = [insel_sift[f"desc{i}"] for i in range(3)] + [insel_sift["desc0"][:3]]
    kps = [hom[f"sift_kp{i}"] for i in range(3)] + [hom["sift_kp0"][:3]]
    # packed cv::KeyPoint-like records (28 bytes): pt is the first field
    recs = []
    for k in kps:
        r = np.zeros((len(k), 7), np.float32)
        r[:, :2] = k
        r[:, 2:] = 123.0
        recs.append(r)
    pairs = np.array([(0, 1), (3, 0), (1, 2), (0, 2)], np.int32)      # (3, 0): 3 query rows -> < 4 matches
    matcher.upload_bank(bank)
    matcher.upload_keypoints([r[:, :2] for r in recs])
    res = matcher.match_pairs(pairs, NORM_L2, min_match_count=180)    # drops pair (0, 2) (163 matches)
    thr = np.array([3.0, 3.0, 1.0, 3.0])
    r = matcher.homography_inlier_ratios(thr, 500, seed=5)
    ratios, inl, hyp = r["ratio"], r["inliers"], r["hypothesis"]
    assert ratios[1] == -1.0 and inl[1] == -1                          # fewer than 4 matches
    assert ratios[3] == -1.0 and res[3] is None                        # dropped by min_match_count
    for p in (0, 2):
        a, b = pairs[p]
        pts = hn.aligned_points(kps[a], kps[b], res[p])
        cnt, h, _ = hn.ransac_inliers(pts, float(thr[p]), pair=p, seed=5, max_iters=500)
        assert (int(inl[p]), int(hyp[p])) == (cnt, h)
    assert ratios[2] < ratios[0]                                       # 1-pixel threshold keeps fewer matches
    with pytest.raises(sfm.SfmError):
        matcher.homography_inlier_ratios(-1.0)
    with pytest.raises(sfm.SfmError):
        matcher.upload_keypoints(kps[:2])                              # image count differs from the bank


@pytest.mark.gpu
def test_gpu_synthetic_planted_homographies(sfm, matcher, hom):
    """Many pairs at once, match lists larger than the shared-memory staging (5000 > 3072): counts equal the restatement,
    ratios within tolerance of cv2, consensus sets are the planted inliers."""
    import workloads
    ks = [0, 1, 2, 3, 4]
    # one image pair per synthetic case: descriptors are one-hot-ish unique rows so that match i <-> i survives the ratio test
    bank, kps, pairs = [], [], []
    for j, k in enumerate(ks):
        p1, p2 = hom[f"syn{k}_p1"], hom[f"syn{k}_p2"]
        n = len(p1)
        rng = np.random.default_rng(100 + k)
        d = rng.integers(0, 256, size=(n, 128), dtype=np.uint8)
        bank += [d, d.copy()]
        kps += [p1, p2]
        pairs.append((2 * j, 2 * j + 1))
    pairs = np.array(pairs, np.int32)
    matcher.upload_bank(bank)
    matcher.upload_keypoints(kps)
    res = matcher.match_pairs(pairs, NORM_L2)
    r = matcher.homography_inlier_ratios(3.0, 2000, seed=3)
    ratios, inl, hyp = r["ratio"], r["inliers"], r["hypothesis"]
    for j, k in enumerate(ks):
        m = res[j]
        n = len(hom[f"syn{k}_p1"])
        assert len(m) == n and np.array_equal(m["queryIdx"], m["trainIdx"])      # identical rows: distance 0 vs random
        pts = np.concatenate([hom[f"syn{k}_p1"], hom[f"syn{k}_p2"]], 1)
        cnt, h, mask = hn.ransac_inliers(pts, 3.0, pair=j, seed=3)
        assert (int(inl[j]), int(hyp[j])) == (cnt, h), (k, int(inl[j]), cnt)
        assert abs(ratios[j] - int(hom[f"syn{k}_count"]) / n) <= RATIO_TOL


@pytest.mark.gpu
def test_gpu_host_mirror_cli_homography(sfm, hom, insel_sift, tmp_path):
    """C++ host mirror (MatchingStage::calculateHomography) through the CLI with the reference's switch spelling
    (-Pransac-matching-threshold, PhotogrammetrieCli.cpp:98-99; relative threshold = max(image side) * t)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cli = os.path.join(root, "sfm-mvs-pipeline_b200", "sfm_match_cli")
    if not os.path.exists(cli):
        pytest.skip("cli not built")
    bank = [insel_sift[f"desc{i}"].astype(np.float32) for i in range(3)]
    path, kpath = tmp_path / "bank.sfmd", tmp_path / "bank.sfmk"
    with open(path, "wb") as f:
        f.write(b"SFMD" + np.array([1, 3, 128, 5], np.uint32).tobytes())
        for d in bank:
            f.write(np.uint32(d.shape[0]).tobytes() + d.tobytes())
    with open(kpath, "wb") as f:
        f.write(b"SFMK" + np.array([1, 3], np.uint32).tobytes())
        for i in range(3):
            k = hom[f"sift_kp{i}"]
            w, h = hom["sift_sizes"][i]
            f.write(np.array([len(k), w, h], np.uint32).tobytes() + np.ascontiguousarray(k, np.float32).tobytes())
    r = subprocess.run([cli, f"-Pdescriptors={path}", f"-Pkeypoints={kpath}", "-Pfeature-detector=SIFT",
                        "-Pmatch-threshold=20", "-Pransac-matching-threshold=0.006"], capture_output=True, text=True, timeout=120)
    assert "pairs=3 kept=3 matches=583" in r.stdout, r.stdout + r.stderr          # 205 + 163 + 215
    ratios = [float(line.rsplit(" ", 1)[1]) for line in r.stdout.splitlines() if "homographyInlierRatio" in line]
    assert len(ratios) == 3
    thr = 720 * 0.006                                                              # insel images are 720 x 405
    for (a, b), got in zip(PAIRS, ratios):
        good = insel_sift[f"p{a}{b}_good"]
        pts = hn.aligned_points(hom[f"sift_kp{a}"], hom[f"sift_kp{b}"], good)
        from oracle import cv2_ref
        if cv2_ref.available():
            cv, _ = cv2_ref.find_homography_inliers(pts[:, :2], pts[:, 2:], thr)
            assert abs(got - cv / len(good)) <= RATIO_TOL
        assert 0.5 < got <= 1.0
