"""The numpy restatement of cv::ORB (oracle/orb_np.py) against cv2's golden vectors (tests/golden/orb_extract.npz, written by
tests/golden/make_golden_orb.py) — CPU only.  Keypoints are compared as SETS keyed by (octave, x, y) in level coordinates:
cv::ORB's own order is a by-product of std::nth_element inside KeyPointsFilter::retainBest."""
import os

import numpy as np
import pytest

from oracle import orb_np as O

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(HERE, "golden", "orb_extract.npz"))


def keyed(kp, desc):
    out = {}
    for k, d in zip(kp, desc):
        s = O.level_scale(int(k["octave"]))
        out[(int(k["octave"]), int(round(float(k["x"]) / float(s))), int(round(float(k["y"]) / float(s))))] = (k, d)
    assert len(out) == len(kp)
    return out


def assert_same_features(kp_ref, desc_ref, kp, desc, what, min_exact_desc=1.0):
    a, b = keyed(kp_ref, desc_ref), keyed(kp, desc)
    assert set(a) == set(b), (what, len(a), len(b), sorted(set(a) ^ set(b))[:5])
    exact = 0
    for key in a:
        (k0, d0), (k1, d1) = a[key], b[key]
        assert k0["x"] == k1["x"] and k0["y"] == k1["y"] and k0["size"] == k1["size"], (what, key)
        assert k0["response"] == k1["response"], (what, key, k0["response"], k1["response"])
        da = abs(float(k0["angle"]) - float(k1["angle"]))
        assert min(da, 360 - da) < 1e-3, (what, key, k0["angle"], k1["angle"])
        exact += int(np.array_equal(d0, d1))
    assert exact >= min_exact_desc * len(a), (what, exact, len(a))
    return exact, len(a)


def test_pattern_is_the_one_cv2_uses(gold):
    assert np.array_equal(gold["pattern"], O.BIT_PATTERN_31) and O.BIT_PATTERN_31.shape == (256, 4)
    assert np.abs(O.BIT_PATTERN_31).max() <= 13                 # every sample stays inside the 31 x 31 patch when rotated


def test_primitives():
    assert O.border_width() == 32
    assert O.umax_table() == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    q = O.features_per_level(500)
    assert q[0] == 109 and sum(q) == 500 and all(a >= b for a, b in zip(q[:-2], q[1:-1])) and sum(O.features_per_level(30000)) == 30000
    assert [O.level_size(405, 720, lv) for lv in (0, 1, 7)] == [(405, 720), (338, 600), (113, 201)]
    # FAST: one bright pixel on a dark ground is a corner (every circle pixel is darker by 190 -> score 189); a 3 x 3 blob is
    # nine corners of EQUAL score, which the strict '>' of the non-maximum suppression removes altogether (as cv::FAST does)
    img = np.full((40, 40), 10, np.uint8)
    img[20, 20] = 200
    xs, ys, sc = O.fast9_corners(img)
    assert list(zip(xs.tolist(), ys.tolist(), sc.tolist())) == [(20, 20, 189)]
    img[19:22, 19:22] = 200
    assert len(O.fast9_corners(img)[0]) == 0


@pytest.mark.parametrize("name,n", [("insel_crop", 30000), ("insel_crop", 500), ("syn4", 300), ("syn4", 5000)])
def test_restatement_equals_cv2(gold, name, n):
    kp, desc = O.detect_and_compute(gold[name], n)
    exact, total = assert_same_features(gold[f"{name}_kp_{n}"], gold[f"{name}_desc_{n}"], kp, desc, f"{name} {n}")
    assert total > 200


def test_reference_photograph_counts(gold):
    """ORB_create(30000) on images/insel/1.jpg gives 14 655 features."""
    kp, desc = O.detect_and_compute(gold["insel1_gray"], 30000)
    assert len(kp) == 14655 == int(gold["insel_counts_30000"][0])
    assert_same_features(gold["insel1_kp_30000"], gold["insel1_desc_30000"], kp, desc, "insel 1")
