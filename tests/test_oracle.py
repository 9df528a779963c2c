"""Pin the oracle (oracle/oracle_np.py) against the committed cv2 golden vectors, the
reference's only known-answer vector, and (when cv2 is importable) live cv2 on seeded inputs."""
import hashlib

import numpy as np
import pytest

import workloads
from oracle import oracle_np as orc
from oracle.oracle_np import NORM_HAMMING, NORM_L2

PAIRS = ((0, 1), (0, 2), (1, 2))


def _check_knn(idx, dist, g_idx, g_dist):
    assert np.array_equal(idx, g_idx)
    assert np.array_equal(dist.view(np.uint32), g_dist.view(np.uint32))   # every distance bit


def test_grid_known_answer_from_reference_header():
    # GridFeatureMatchingStrategy.h:31-39: seq 3, row length 5, image 01 pairs with 02,03,06,11,07
    p = orc.pairs_grid(20, 3, 5)
    first = sorted(int(b) + 1 for a, b in p if a == 0)
    assert first == sorted([2, 3, 6, 11, 7])
    # emission order of the nested loops (Grid...cpp:67-83)
    assert [int(b) for a, b in p if a == 0] == [1, 2, 5, 6, 10]


def test_grid_pair_counts_and_floor_quirk():
    assert len(orc.pairs_grid(1000, 3, 40)) == 4741
    assert len(orc.pairs_grid(1000, 2, 40)) == 1935
    assert len(orc.pairs_grid(1000, 4, 40)) == 8355
    assert orc.pairs_grid(3, 2, 3).tolist() == [[0, 1], [1, 2]]
    assert orc.pairs_grid(3, 2, 1).tolist() == [[0, 1], [1, 2]]
    assert orc.pairs_grid(3, 3, 3).tolist() == [[0, 1], [0, 2], [1, 2]]
    assert orc.pairs_grid(3, 2, 2).tolist() == [[0, 1]]   # trailing partial row dropped (floor)
    with pytest.raises(ValueError):
        orc.pairs_grid(4, 1, 2)
    with pytest.raises(ValueError):
        orc.pairs_grid(4, 2, 0)


def test_video_and_unordered_pairs():
    assert orc.pairs_video(3, 2).tolist() == [[0, 1], [1, 2]]
    assert orc.pairs_video(3, 3).tolist() == [[0, 1], [0, 2], [1, 2]]
    assert len(orc.pairs_video(1000, 5)) == 3990
    assert len(orc.pairs_unordered(200)) == 19900
    assert orc.pairs_unordered(0).shape == (0, 2)
    with pytest.raises(ValueError):
        orc.pairs_video(5, 1)
    assert orc.select_pairs(5, 0, 0).shape[0] == 10
    assert orc.select_pairs(5, 2, 0).tolist() == orc.pairs_video(5, 2).tolist()
    assert orc.select_pairs(6, 2, 3).tolist() == orc.pairs_grid(6, 2, 3).tolist()


@pytest.mark.parametrize("a,b", PAIRS)
def test_insel_sift_matches_cv2_golden(insel_sift, a, b):
    q, t = insel_sift[f"desc{a}"], insel_sift[f"desc{b}"]
    idx, dist = orc.knn2_l2(q, t)
    _check_knn(idx, dist, insel_sift[f"p{a}{b}_nidx"], insel_sift[f"p{a}{b}_dist"])
    good = orc.ratio_filter(idx, dist, 0.7)
    assert orc.dmatch_equal(good, insel_sift[f"p{a}{b}_good"])


@pytest.mark.parametrize("a,b", PAIRS)
def test_insel_orb_matches_cv2_golden(insel_orb, a, b):
    q, t = insel_orb[f"desc{a}"], insel_orb[f"desc{b}"]
    idx, dist = orc.knn2_hamming(q, t)
    _check_knn(idx, dist, insel_orb[f"p{a}{b}_nidx"], insel_orb[f"p{a}{b}_dist"])
    good = orc.ratio_filter(idx, dist, 0.7)
    assert orc.dmatch_equal(good, insel_orb[f"p{a}{b}_good"])


def test_insel_sha1_appendix_b(insel_sift, insel_orb):
    exp = {"SIFT_01": (205, "8db69caf1819"), "SIFT_02": (163, "bad86f99d227"), "SIFT_12": (215, "92fbc19a8882"),
           "ORB_01": (6170, "6d8dfe32e5d3"), "ORB_02": (3989, "a273bbf1b448"), "ORB_12": (6797, "261eb561e2ea")}
    for det, fx in (("SIFT", insel_sift), ("ORB", insel_orb)):
        for a, b in PAIRS:
            g = fx[f"p{a}{b}_good"]
            arr = np.stack([g["queryIdx"], g["trainIdx"]], 1).astype(np.int32)
            assert (len(g), hashlib.sha1(arr.tobytes()).hexdigest()[:12]) == exp[f"{det}_{a}{b}"]


def test_adversarial_and_synthetic_vs_cv2_golden(synthetic_cv2):
    adv = workloads.adversarial_sift()
    cases = {"dup_base": (adv["dup"], adv["base"]), "base_dup": (adv["base"], adv["dup"]),
             "zeros_dup": (adv["zeros"], adv["dup"]), "sat_sat": (adv["sat"], adv["sat"]),
             "base_two": (adv["base"], adv["two"]),
             "n127_n129": (adv["n127"], adv["n129"]), "n129_n127": (adv["n129"], adv["n127"])}
    bank = workloads.sift_like_bank(3, 700)
    cases["syn01"] = (bank[0], bank[1])
    cases["syn12"] = (bank[1], bank[2])
    for name, (q, t) in cases.items():
        idx, dist = orc.knn2_l2(q, t)
        _check_knn(idx, dist, synthetic_cv2[f"{name}_nidx"], synthetic_cv2[f"{name}_dist"])
        assert orc.dmatch_equal(orc.ratio_filter(idx, dist), synthetic_cv2[f"{name}_good"])
    # one train row -> one-element lists, all kept (Unordered...cpp:62-64)
    idx, dist = orc.knn2_l2(adv["base"], adv["one"])
    assert np.array_equal(idx[:, 0], synthetic_cv2["base_one_nidx"][:, 0]) and np.all(idx[:, 1] == -1)
    assert np.array_equal(dist[:, 0].view(np.uint32), synthetic_cv2["base_one_dist"][:, 0].view(np.uint32))
    assert len(orc.ratio_filter(idx, dist)) == adv["base"].shape[0]
    assert orc.dmatch_equal(orc.ratio_filter(idx, dist), synthetic_cv2["base_one_good"])


def test_hamming_synthetic_vs_cv2_golden(synthetic_cv2):
    ob = workloads.orb_like_bank(3, 900)
    obd = np.concatenate([ob[1][:50], ob[1][:50], ob[1]])
    for name, q, t in (("orb01", ob[0], ob[1]), ("orb12", ob[1], ob[2]), ("orb_dup", ob[0], obd)):
        idx, dist = orc.knn2_hamming(q, t)
        _check_knn(idx, dist, synthetic_cv2[f"{name}_nidx"], synthetic_cv2[f"{name}_dist"])
        assert orc.dmatch_equal(orc.ratio_filter(idx, dist), synthetic_cv2[f"{name}_good"])
    idx, dist = orc.knn2_hamming(ob[0], ob[1][:1])
    assert orc.dmatch_equal(orc.ratio_filter(idx, dist), synthetic_cv2["orb_one_good"])


def test_cross_check_vs_cv2_golden(insel_sift, synthetic_cv2):
    q, t = insel_sift["desc0"], insel_sift["desc1"]
    i1, d1 = orc.knn2_l2(q, t)
    i2, _ = orc.knn2_l2(t, q)
    assert orc.dmatch_equal(orc.cross_check_filter(i1, d1, i2), insel_sift["p01_cross"])
    bank = workloads.sift_like_bank(3, 700)
    i1, d1 = orc.knn2_l2(bank[0], bank[1])
    i2, _ = orc.knn2_l2(bank[1], bank[0])
    assert orc.dmatch_equal(orc.cross_check_filter(i1, d1, i2), synthetic_cv2["syn01_cross"])
    ob = workloads.orb_like_bank(3, 900)
    i1, d1 = orc.knn2_hamming(ob[0], ob[1])
    i2, _ = orc.knn2_hamming(ob[1], ob[0])
    assert orc.dmatch_equal(orc.cross_check_filter(i1, d1, i2), synthetic_cv2["orb01_cross"])


def test_degenerate_shapes():
    adv = workloads.adversarial_sift()
    idx, dist = orc.knn2_l2(adv["base"], adv["empty"])      # train 0x128 -> N empty lists
    assert np.all(idx == -1) and len(orc.ratio_filter(idx, dist)) == 0
    idx, dist = orc.knn2_l2(adv["empty"], adv["base"])      # query 0x128 -> empty result
    assert idx.shape == (0, 2)
    with pytest.raises(ValueError):
        orc.knn2_l2(adv["base"], np.zeros((4, 64), np.uint8))
    with pytest.raises(ValueError):
        orc.knn2_hamming(adv["base"].astype(np.float32), adv["base"].astype(np.float32))


def test_distinct_and_min_count():
    m = np.zeros(5, orc.DMATCH_DTYPE)
    m["queryIdx"] = [0, 1, 2, 3, 4]
    m["trainIdx"] = [7, 8, 7, 9, 8]
    assert orc.distinct_filter(m)["queryIdx"].tolist() == [3]      # SfM.cpp:547-564
    bank = workloads.sift_like_bank(3, 300)
    res = orc.match_pairs(bank, orc.pairs_unordered(3), NORM_L2, min_match_count=20)
    assert res[0] is not None and res[1] is None or res[1] is not None   # planted pairs survive
    assert all(r is None or len(r) >= 20 for r in res)


def test_live_cv2_agrees_on_seeded_inputs():
    cv2_ref = pytest.importorskip("oracle.cv2_ref")
    if not cv2_ref.available():
        pytest.skip("cv2 not importable")
    bank = workloads.sift_like_bank(2, 1500, seed_base=77)
    idx, dist = orc.knn2_l2(bank[0], bank[1])
    nidx, d = cv2_ref.batch_distance_k2(bank[0], bank[1], NORM_L2)
    _check_knn(idx, dist, nidx, d)
    ob = workloads.orb_like_bank(2, 1500, seed_base=78)
    idx, dist = orc.knn2_hamming(ob[0], ob[1])
    nidx, d = cv2_ref.batch_distance_k2(ob[0], ob[1], NORM_HAMMING)
    _check_knn(idx, dist, nidx, d)
