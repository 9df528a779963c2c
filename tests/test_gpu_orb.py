"""cv::ORB on the device (csrc/orb.cu through the C ABI) against the numpy restatement (oracle/orb_np.py) and against cv2's golden
vectors (SfM::extractFeatures with -Pfeature-detector=ORB, PhotogrammetrieCli.cpp:347-348).  Everything that decides a keypoint is
integer or order-fixed float arithmetic: device == restatement exactly (same keypoints in the same (level, y, x) order, same
responses, same descriptors; angle within 1e-3 degrees), and restatement == cv2 as a set (tests/test_orb_oracle.py)."""
import os
import sys

import numpy as np
import pytest

import workloads
from oracle import oracle_np as orc
from oracle import orb_np as O

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from test_orb_oracle import assert_same_features  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def m(sfm):
    mt = sfm.Matcher(0)
    yield mt
    mt.close()


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(HERE, "golden", "orb_extract.npz"))


def assert_equals_restatement(kp, desc, kp_o, desc_o, what):
    assert len(kp) == len(kp_o), (what, len(kp), len(kp_o))
    for f in ("x", "y", "size", "response", "octave"):
        assert np.array_equal(kp[f], kp_o[f]), (what, f, np.nonzero(kp[f] != kp_o[f])[0][:5])
    da = np.abs(kp["angle"].astype(np.float64) - kp_o["angle"].astype(np.float64))
    assert np.minimum(da, 360 - da).max(initial=0) < 1e-3, what
    assert np.array_equal(desc, desc_o), (what, int((desc != desc_o).any(1).sum()), len(desc))


@pytest.mark.parametrize("shape,n", [((240, 320), 500), ((240, 320), 30000), ((200, 264), 150), ((131, 517), 2000), ((90, 100), 500)])
def test_device_equals_restatement(m, shape, n):
    img = workloads.synthetic_photo(6, *shape)
    m.features_clear()
    got = m.extract_orb(img, n_features=n)
    kp, desc = m.features_download(0)
    assert got == len(kp) and desc.shape == (len(kp), 32) and m.features_descriptor_bytes() == 32
    kp_o, desc_o = O.detect_and_compute(img, n)
    assert_equals_restatement(kp, desc, kp_o, desc_o, f"{shape} {n}")
    if shape[0] >= 200:
        assert len(kp) > 100
        order = np.lexsort((kp["x"], kp["y"], kp["octave"]))
        assert np.array_equal(order, np.arange(len(kp)))                    # (level, y, x)


@pytest.mark.parametrize("name,n", [("insel_crop", 30000), ("insel_crop", 500), ("syn4", 300)])
def test_device_against_cv2_golden(m, gold, name, n):
    m.features_clear()
    m.extract_orb(gold[name], n_features=n)
    kp, desc = m.features_download(0)
    assert_same_features(gold[f"{name}_kp_{n}"], gold[f"{name}_desc_{n}"], kp, desc, f"{name} {n} vs cv2")


def test_reference_photograph_feature_count(m, gold):
    """cv2.ORB_create(30000) on images/insel/1.jpg -> 14 655 features (run-orb-sequence.sh's feature limit)."""
    m.features_clear()
    n = m.extract_orb(gold["insel1_gray"], n_features=30000)
    assert n == 14655
    kp, desc = m.features_download(0)
    assert_same_features(gold["insel1_kp_30000"], gold["insel1_desc_30000"], kp, desc, "insel 1 vs cv2")


def test_orb_features_feed_the_hamming_matcher(m, sfm):
    a = workloads.synthetic_photo(11, 240, 320)
    b = np.roll(a, (4, 6), axis=(0, 1))
    m.features_clear()
    m.extract_orb(a, n_features=2000)
    m.extract_orb(b, n_features=2000)
    m.extract_orb(np.full((80, 80), 7, np.uint8), n_features=2000)           # a shot without features
    m.bank_from_features()
    assert m.bank_info()["cols"] == 32
    feats = [m.features_download(i) for i in range(3)]
    pairs = sfm.select_pairs(3, 0, 0)
    for eng in (sfm.ENGINE_AUTO, sfm.ENGINE_SIMT):
        res = m.match_pairs(pairs, sfm.NORM_HAMMING, engine=eng)
        exp = orc.match_pairs([f[1] for f in feats], pairs, orc.NORM_HAMMING)
        for p in range(len(pairs)):
            assert orc.dmatch_equal(res[p], exp[p]), pairs[p]
    assert len(res[0]) > 100 and len(res[1]) == 0
    ka, kb = feats[0][0], feats[1][0]
    dx = kb["x"][res[0]["trainIdx"]] - ka["x"][res[0]["queryIdx"]]
    dy = kb["y"][res[0]["trainIdx"]] - ka["y"][res[0]["queryIdx"]]
    assert abs(np.median(dx) - 6) < 0.5 and abs(np.median(dy) - 4) < 0.5
    h = m.homography_inlier_ratios(3.0, seed=1)
    assert h["ratio"][0] > 0.7


def test_orb_edge_cases(m, sfm):
    m.features_clear()
    assert m.extract_orb(np.full((100, 100), 50, np.uint8)) == 0
    assert m.extract_orb(workloads.synthetic_photo(1, 62, 62)) == 0         # no level is wider than 2 x edgeThreshold
    wide = workloads.synthetic_photo(5, 150, 400)
    m.features_clear()
    m.extract_orb(wide[:, 40:300], n_features=800)
    m.extract_orb(np.ascontiguousarray(wide[:, 40:300]), n_features=800)     # strided input == contiguous copy
    a, b = m.features_download(0), m.features_download(1)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and len(a[0]) > 50
    with pytest.raises(sfm.SfmError) as e:
        m.extract_orb(wide, n_levels=4)
    assert e.value.code == sfm.ERR_UNSUPPORTED
    with pytest.raises(sfm.SfmError) as e:
        m.extract_orb(wide, n_features=3000, max_keypoints=10)
    assert e.value.code == sfm.ERR_CAPACITY
    with pytest.raises(sfm.SfmError) as e:                                   # one feature set holds SIFT or ORB images, not both
        m.extract_sift(wide)
    assert e.value.code == sfm.ERR_STATE
    m.features_clear()
    assert m.extract_orb(np.zeros((4, 4), np.uint8)) == 0                    # 8 levels down to 1 x 1 pixel, no keypoints


def test_pyramid_score_and_blur_maps_equal_restatement(m):
    """The per-pixel maps behind the keypoints (test hook sfm_features_orb_level): every pyramid level (INTER_LINEAR_EXACT from the
    previous level), the FAST score map, the candidates after non-maximum suppression + border filter and the blurred level."""
    img = workloads.synthetic_photo(6, 240, 320)
    m.features_clear()
    m.extract_orb(img, n_features=30000)
    levels = O.build_pyramid(img)
    for lv in range(8):
        assert np.array_equal(m.orb_level(0, lv), levels[lv]), lv
        assert np.array_equal(m.orb_level(1, lv), O.blur_level(levels[lv])), lv
        assert np.array_equal(m.orb_level(2, lv), O.fast9_scores(levels[lv])), lv
        xs, ys, s = O.fast9_corners(levels[lv])
        h, w = levels[lv].shape
        inb = (xs >= 31) & (xs < w - 31) & (ys >= 31) & (ys < h - 31)
        ref = np.zeros((h, w), np.uint8)
        ref[ys[inb], xs[inb]] = s[inb]
        assert np.array_equal(m.orb_level(3, lv), ref), lv
