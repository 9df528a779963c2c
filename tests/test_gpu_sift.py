"""Feature extraction on the device (csrc/sift.cu through the C ABI) against the numpy restatement of cv::SIFT and
against cv2's golden vectors (SfM::extractFeatures, SfM.cpp:577-597).

Parity bar: the Gaussian pyramid is float arithmetic in a fixed order -> compared to the restatement at 1e-4 absolute
(values 0..255); keypoints / descriptors under the tolerance of tests/_sift_compare.py (borderline float decisions)."""
import os
import sys

import numpy as np
import pytest

import workloads
from oracle import oracle_np as orc
from oracle import sift_np as S

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _sift_compare as sc  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def m(sfm):
    mt = sfm.Matcher(0)
    yield mt
    mt.close()


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(HERE, "golden", "sift_extract.npz"))


def test_pyramid_equals_restatement(m):
    img = workloads.synthetic_photo(0, 240, 320)
    m.features_clear()
    m.extract_sift(img)
    base = S.create_initial_image(img)
    n_oct = S.n_octaves_for(base.shape)
    gpyr = S.build_gaussian_pyramid(base, n_oct)
    for o in range(n_oct):
        for i in (0, 3, 5):
            got = m.pyramid_level(o, i)
            exp = gpyr[o * 6 + i]
            assert got.shape == exp.shape, (o, i)
            assert np.abs(got - exp).max() < 1e-4, (o, i, float(np.abs(got - exp).max()))


@pytest.mark.parametrize("shape,ct", [((240, 320), 0.04), ((240, 320), 0.09), ((101, 67), 0.04), ((64, 513), 0.04)])
def test_keypoints_and_descriptors_equal_restatement(m, shape, ct):
    img = workloads.synthetic_photo(3, *shape)
    m.features_clear()
    n = m.extract_sift(img, contrast_threshold=ct)
    kp, desc = m.features_download(0)
    assert n == len(kp)
    kp_o, desc_o = S.detect_and_compute(img, contrast_threshold=ct)
    r = sc.assert_close(kp_o, desc_o.astype(np.uint8), kp, desc, f"{shape} {ct}")
    assert abs(r["n_a"] - r["n_b"]) <= max(2, r["n_a"] // 100)
    assert np.all(np.diff(kp["x"]) >= 0)          # removeDuplicatedSorted order


def test_reference_photo_against_cv2_golden(m, gold):
    m.features_clear()
    m.extract_sift(gold["insel1_gray"], contrast_threshold=0.09)          # PhotogrammetrieCli.cpp:355 (limit not reached)
    kp, desc = m.features_download(0)
    r = sc.assert_close(gold["insel1_kp_009"], gold["insel1_desc_009"], kp, desc, "insel vs cv2")
    assert r["n_b"] in range(316, 325)
    syn = workloads.synthetic_photo(0, 240, 320)
    m.extract_sift(syn)
    kp, desc = m.features_download(1)
    sc.assert_close(gold["syn0_kp_004"], gold["syn0_desc_004"], kp, desc, "synthetic vs cv2")


def test_feature_limit(m, gold):
    """cv::SIFT::create(featureLimit, 3, 0.09) (PhotogrammetrieCli.cpp:345,355): retainBest keeps the strongest responses."""
    m.features_clear()
    n = m.extract_sift(gold["insel1_gray"], contrast_threshold=0.09, n_features=100)
    kp, desc = m.features_download(0)
    assert n == 101 and np.all(np.diff(kp["x"]) >= 0)
    sc.assert_close(gold["insel1_kp_009_n100"], gold["insel1_desc_009_n100"], kp, desc, "insel limit 100 vs cv2")
    img = workloads.synthetic_photo(3, 240, 320)
    for limit in (1, 37, 150, 100000):
        m.extract_sift(img, n_features=limit)
        kp, desc = m.features_download(m.features_count() - 1)
        kp_o, desc_o = S.detect_and_compute(img, nfeatures=limit)
        assert abs(len(kp) - len(kp_o)) <= 2 and len(kp) >= min(limit, len(kp_o) - 2)
        sc.assert_close(kp_o, desc_o.astype(np.uint8), kp, desc, f"limit {limit}")


def test_edge_cases(m, sfm):
    m.features_clear()
    assert m.extract_sift(np.full((64, 64), 100, np.uint8)) == 0
    assert m.extract_sift(np.arange(36, dtype=np.uint8).reshape(6, 6) * 7) == 0
    assert m.extract_sift(np.zeros((1, 40), np.uint8)) == 0
    assert m.features_count() == 3
    kp, desc = m.features_download(1)
    assert len(kp) == 0 and desc.shape == (0, 128)
    # strided input (a column window of a wider image) equals the contiguous copy
    wide = workloads.synthetic_photo(5, 120, 400)
    m.features_clear()
    m.extract_sift(wide[:, 40:300])
    m.extract_sift(np.ascontiguousarray(wide[:, 40:300]))
    a, b = m.features_download(0), m.features_download(1)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and len(a[0]) > 20
    with pytest.raises(sfm.SfmError) as e:
        m.extract_sift(wide.astype(np.float32))
    assert e.value.code == sfm.ERR_INVALID
    with pytest.raises(sfm.SfmError) as e:
        m.extract_sift(wide, n_octave_layers=9)
    assert e.value.code == sfm.ERR_UNSUPPORTED
    with pytest.raises(sfm.SfmError) as e:
        m.extract_sift(wide, max_keypoints=10)
    assert e.value.code == sfm.ERR_CAPACITY
    with pytest.raises(sfm.SfmError):
        m.features_download(7)


def test_run_to_run_identical(m):
    img = workloads.synthetic_photo(7, 200, 300)
    m.features_clear()
    m.extract_sift(img)
    m.extract_sift(img)
    a, b = m.features_download(0), m.features_download(1)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_extracted_features_feed_the_matching_stage(m, sfm):
    """extract -> bank_from_features -> match_pairs -> homography, nothing uploaded by the host in between."""
    a = workloads.synthetic_photo(11, 240, 320)
    b = np.roll(a, (3, 5), axis=(0, 1))                    # the same scene shifted by (5, 3) pixels
    m.features_clear()
    m.extract_sift(a)
    m.extract_sift(b)
    m.extract_sift(np.full((32, 32), 9, np.uint8))          # a shot without features
    m.bank_from_features()
    feats = [m.features_download(i) for i in range(3)]
    pairs = sfm.select_pairs(3, 0, 0)
    res = m.match_pairs(pairs, sfm.NORM_L2)
    exp = orc.match_pairs([f[1] for f in feats], pairs, orc.NORM_L2)
    for p in range(len(pairs)):
        assert orc.dmatch_equal(res[p], exp[p]), pairs[p]
    assert len(res[0]) > 50 and len(res[1]) == 0 and len(res[2]) == 0
    # the matches of the shifted pair agree with the shift, and the homography stage sees the keypoints
    ka, kb = feats[0][0], feats[1][0]
    dx = kb["x"][res[0]["trainIdx"]] - ka["x"][res[0]["queryIdx"]]
    dy = kb["y"][res[0]["trainIdx"]] - ka["y"][res[0]["queryIdx"]]
    assert np.median(np.abs(dx - 5)) < 0.1 and np.median(np.abs(dy - 3)) < 0.1
    h = m.homography_inlier_ratios(3.0, seed=1)
    assert h["ratio"][0] > 0.8 and h["ratio"][1] == -1


@pytest.mark.parametrize("n_layers,sigma,edge,ct", [(2, 1.6, 10.0, 0.04), (4, 1.2, 5.0, 0.04), (5, 2.0, 20.0, 0.09)])
def test_non_default_detector_parameters_equal_restatement(m, n_layers, sigma, edge, ct):
    """cv::SIFT::create(nfeatures, nOctaveLayers, contrastThreshold, edgeThreshold, sigma) with values other than the
    reference's (3, 10, 1.6): pyramid depth, blur schedule and the edge test all change.  Device == restatement with the
    SAME keypoint count."""
    img = workloads.synthetic_photo(9, 180, 260)
    m.features_clear()
    n = m.extract_sift(img, contrast_threshold=ct, n_octave_layers=n_layers, edge_threshold=edge, sigma=sigma)
    kp, desc = m.features_download(0)
    kp_o, desc_o = S.detect_and_compute(img, n_layers=n_layers, contrast_threshold=ct, edge_threshold=edge, sigma=sigma)
    assert n == len(kp) == len(kp_o), (n, len(kp_o))
    assert len(kp) > 20
    sc.assert_close(kp_o, desc_o.astype(np.uint8), kp, desc, f"layers {n_layers} sigma {sigma} edge {edge}")
    base = S.create_initial_image(img, sigma=sigma)
    gpyr = S.build_gaussian_pyramid(base, S.n_octaves_for(base.shape), n_layers=n_layers, sigma=sigma)
    for o in (0, 1):
        for i in (0, n_layers + 2):
            got = m.pyramid_level(o, i)
            exp = gpyr[o * (n_layers + 3) + i]
            assert got.shape == exp.shape and np.abs(got - exp).max() < 1e-4, (o, i)


def test_compute_semantics_without_octave_minus_one(m):
    """The reference calls detect() then compute() (SfM.cpp:586-587); cv::SIFT::compute rebuilds the pyramid WITHOUT the 2x
    upsampling when no keypoint lies in octave -1.  A picture of a few wide blobs has none: the device then takes its
    descriptors from a second, non-doubled pyramid and equals the restatement of compute() exactly."""
    yy, xx = np.mgrid[0:160, 0:200]
    img = np.zeros((160, 200), np.float32)
    rng = np.random.default_rng(5)
    for _ in range(6):
        cx, cy, sg = rng.uniform(40, 160), rng.uniform(40, 120), rng.uniform(6, 12)
        img += rng.uniform(0.5, 1) * np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * sg * sg))
    img = np.rint(255 * img / img.max()).astype(np.uint8)
    kp_o = S.detect(img)
    assert len(kp_o) >= 10 and min(S.unpack_octave(int(o))[0] for o in kp_o["octave"]) >= 0
    desc_o = S.compute(img, kp_o)                                     # own pyramid, not doubled
    m.features_clear()
    n = m.extract_sift(img)
    kp, desc = m.features_download(0)
    assert n == len(kp) == len(kp_o)
    sc.assert_close(kp_o, desc_o.astype(np.uint8), kp, desc, "compute() without octave -1")
    assert np.abs(desc.astype(np.int32) - desc_o.astype(np.int32)).max() <= 1
    # the pyramid left behind is the non-doubled one: octave 0 has the size of the input image
    assert m.pyramid_level(0, 0).shape == img.shape
    # and a photograph right after it uses the doubled pyramid again
    photo = workloads.synthetic_photo(3, 120, 160)
    m.extract_sift(photo)
    assert m.pyramid_level(0, 0).shape == (240, 320)
