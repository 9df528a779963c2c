"""GPU parity tests (run with -m gpu on the B200): the CUDA path through the C ABI against the oracle and the
committed cv2 golden vectors.  Bit-exact for u8-valued SIFT and for ORB (indices AND distance bits); stated
tolerance only for non-integer float descriptors."""
import os
import subprocess
import sys

import numpy as np
import pytest

import workloads
from oracle import oracle_np as orc
from oracle.oracle_np import NORM_HAMMING, NORM_L2

pytestmark = pytest.mark.gpu
PAIRS = ((0, 1), (0, 2), (1, 2))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def matcher(sfm):
    m = sfm.Matcher(0)
    yield m
    m.close()


def _same_knn(idx, dist, g_idx, g_dist):
    assert np.array_equal(idx, g_idx), f"index mismatch at rows {np.nonzero((idx != g_idx).any(1))[0][:10]}"
    assert np.array_equal(dist.view(np.uint32), g_dist.view(np.uint32))


def _engines(sfm):
    # "tensor" = tcgen05 (value-only kernel where allowed), "tensor_imad" = tcgen05 with the packed-key epilogue
    return [("tensor", sfm.ENGINE_TENSOR), ("tensor_imad", sfm.ENGINE_TENSOR_IMAD), ("simt", sfm.ENGINE_SIMT)]


# ------------------------------------------------------------------ operator level: knnMatch
@pytest.mark.parametrize("a,b", PAIRS)
@pytest.mark.parametrize("eng", ["tensor", "simt"])
@pytest.mark.parametrize("as_float", [True, False])
def test_insel_sift_knn_bit_exact(sfm, matcher, insel_sift, a, b, eng, as_float):
    q, t = insel_sift[f"desc{a}"], insel_sift[f"desc{b}"]
    if as_float:      # the reference hands CV_32F integer-valued Mats
        q, t = q.astype(np.float32), t.astype(np.float32)
    idx, dist = matcher.knn_match(q, t, NORM_L2, 2, dict(_engines(sfm))[eng])
    _same_knn(idx, dist, insel_sift[f"p{a}{b}_nidx"], insel_sift[f"p{a}{b}_dist"])


@pytest.mark.parametrize("a,b", PAIRS)
@pytest.mark.parametrize("eng", ["tensor", "simt"])     # tcgen05 on bit-expanded rows / the __popc kernel
def test_insel_orb_knn_bit_exact(sfm, matcher, insel_orb, a, b, eng):
    idx, dist = matcher.knn_match(insel_orb[f"desc{a}"], insel_orb[f"desc{b}"], NORM_HAMMING, 2, dict(_engines(sfm))[eng])
    _same_knn(idx, dist, insel_orb[f"p{a}{b}_nidx"], insel_orb[f"p{a}{b}_dist"])


@pytest.mark.parametrize("eng", ["tensor", "simt"])
def test_adversarial_ties_and_ragged_sizes(sfm, matcher, synthetic_cv2, eng):
    adv = workloads.adversarial_sift()
    e = dict(_engines(sfm))[eng]
    cases = {"dup_base": (adv["dup"], adv["base"]), "base_dup": (adv["base"], adv["dup"]),
             "zeros_dup": (adv["zeros"], adv["dup"]), "sat_sat": (adv["sat"], adv["sat"]),
             "base_two": (adv["base"], adv["two"]), "n127_n129": (adv["n127"], adv["n129"]),
             "n129_n127": (adv["n129"], adv["n127"])}
    for name, (q, t) in cases.items():
        idx, dist = matcher.knn_match(q, t, NORM_L2, 2, e)
        _same_knn(idx, dist, synthetic_cv2[f"{name}_nidx"], synthetic_cv2[f"{name}_dist"])
    # one train row -> single neighbour, second slot -1 / inf
    idx, dist = matcher.knn_match(adv["base"], adv["one"], NORM_L2, 2, e)
    assert np.array_equal(idx[:, 0], synthetic_cv2["base_one_nidx"][:, 0]) and np.all(idx[:, 1] == -1)
    assert np.array_equal(dist[:, 0].view(np.uint32), synthetic_cv2["base_one_dist"][:, 0].view(np.uint32))
    # zero train rows -> N empty lists; zero query rows -> empty result
    idx, dist = matcher.knn_match(adv["base"], adv["empty"], NORM_L2, 2, e)
    assert np.all(idx == -1)
    idx, dist = matcher.knn_match(adv["empty"], adv["base"], NORM_L2, 2, e)
    assert idx.shape == (0, 2)
    # k = 1 (DescriptorMatcher::match)
    idx, dist = matcher.knn_match(adv["dup"], adv["base"], NORM_L2, 1, e)
    assert np.array_equal(idx[:, 0], synthetic_cv2["dup_base_nidx"][:, 0])


def test_hamming_ties_and_synthetic(sfm, matcher, synthetic_cv2):
    ob = workloads.orb_like_bank(3, 900)
    obd = np.concatenate([ob[1][:50], ob[1][:50], ob[1]])
    for name, q, t in (("orb01", ob[0], ob[1]), ("orb12", ob[1], ob[2]), ("orb_dup", ob[0], obd),
                       ("orb_one", ob[0], ob[1][:1])):
        for eng in (sfm.ENGINE_TENSOR, sfm.ENGINE_SIMT):
            idx, dist = matcher.knn_match(q, t, NORM_HAMMING, 2, eng)
            g_idx, g_dist = synthetic_cv2[f"{name}_nidx"], synthetic_cv2[f"{name}_dist"]
            if name == "orb_one":
                assert np.array_equal(idx[:, 0], g_idx[:, 0]) and np.all(idx[:, 1] == -1)
            else:
                _same_knn(idx, dist, g_idx, g_dist)
    # all-zero / all-one descriptors and ragged sizes
    z = np.zeros((70, 32), np.uint8)
    o = np.full((45, 32), 255, np.uint8)
    mix = np.concatenate([z[:3], ob[0][:200], o[:2], ob[0][:50]])
    for q, t in ((mix, ob[1][:257]), (ob[1][:129], mix), (z, o), (mix, mix)):
        e_idx, e_dist = orc.knn2_hamming(q, t)
        for eng in (sfm.ENGINE_TENSOR, sfm.ENGINE_SIMT):
            idx, dist = matcher.knn_match(q, t, NORM_HAMMING, 2, eng)
            _same_knn(idx, dist, e_idx, e_dist)


def test_error_behaviour_mirrors_opencv(sfm, matcher):
    adv = workloads.adversarial_sift()
    with pytest.raises(sfm.SfmError):       # width mismatch -> cv::error
        matcher.knn_match(adv["base"], np.zeros((4, 64), np.uint8), NORM_L2)
    with pytest.raises(sfm.SfmError):       # float descriptors with NORM_HAMMING -> cv::error
        matcher.knn_match(adv["base"].astype(np.float32), adv["base"].astype(np.float32), NORM_HAMMING)
    with pytest.raises(sfm.SfmError):       # k > 2 unsupported
        matcher.knn_match(adv["base"], adv["base"], NORM_L2, 3)
    with pytest.raises(sfm.SfmError) as e:  # >= 2^18 train rows (IMGIDX_ONE)
        matcher.knn_match(adv["base"], np.zeros((1 << 18, 128), np.uint8), NORM_L2)
    assert e.value.code == sfm.ERR_CAPACITY


def test_float_descriptors_tolerance(sfm, matcher):
    """Non-integer CV_32F data: fp32 path.  Tolerance (north_star): identical match indices except on
    distance ties within 1e-5 relative; distances within 1e-5 relative."""
    rng = np.random.default_rng(5)
    q = rng.random((300, 128), dtype=np.float32) * 3
    t = rng.random((411, 128), dtype=np.float32) * 3
    idx, dist = matcher.knn_match(q, t, NORM_L2, 2)
    eidx, edist = orc.knn2_l2(q, t)
    assert np.allclose(dist, edist, rtol=1e-5)
    diff = idx != eidx
    if diff.any():      # only allowed where the two candidates are tied within tolerance
        r, k = np.nonzero(diff)
        d_alt = np.sqrt(((q[r] - t[idx[r, k]]) ** 2).sum(1))
        assert np.allclose(d_alt, edist[r, k], rtol=1e-5)
    # 64-wide float descriptors (SURF-like) also run
    idx, dist = matcher.knn_match(q[:, :64].copy(), t[:, :64].copy(), NORM_L2, 2)
    eidx, edist = orc.knn2_l2(q[:, :64], t[:, :64])
    assert np.allclose(dist, edist, rtol=1e-5) and (idx == eidx).mean() > 0.999


def test_strided_rows_like_cv_mat_step(sfm, matcher, insel_sift):
    q = insel_sift["desc0"]
    big = np.zeros((q.shape[0], 160), np.uint8)
    big[:, :128] = q
    idx, dist = matcher.knn_match(big[:, :128], insel_sift["desc1"], NORM_L2, 2)
    _same_knn(idx, dist, insel_sift["p01_nidx"], insel_sift["p01_dist"])


# ------------------------------------------------------------------ plugin level: match_pairs
@pytest.mark.parametrize("eng", ["tensor", "tensor_imad", "simt"])
def test_insel_sift_match_pairs_c1(sfm, matcher, insel_sift, eng):
    """BASELINE config C1: insel SIFT, grid pairing (seq,row) in {(2,3),(3,3),(2,1)}."""
    bank = [insel_sift[f"desc{i}"].astype(np.float32) for i in range(3)]
    matcher.upload_bank(bank)
    assert matcher.bank_info()["u8_valued"]
    for seq, row in ((2, 3), (3, 3), (2, 1)):
        pairs = sfm.select_pairs(3, seq, row)
        res = matcher.match_pairs(pairs, NORM_L2, engine=dict(_engines(sfm))[eng])
        for p, (a, b) in enumerate(pairs):
            assert orc.dmatch_equal(res[p], insel_sift[f"p{a}{b}_good"])


def test_insel_orb_match_pairs_c2(sfm, matcher, insel_orb):
    """BASELINE config C2: insel ORB, sequence pairing seq in {2,3}, bit-exact."""
    matcher.upload_bank([insel_orb[f"desc{i}"] for i in range(3)])
    for seq in (2, 3):
        pairs = sfm.select_pairs(3, seq, 0)
        for eng in (sfm.ENGINE_AUTO, sfm.ENGINE_SIMT):
            res = matcher.match_pairs(pairs, NORM_HAMMING, engine=eng)
            for p, (a, b) in enumerate(pairs):
                assert orc.dmatch_equal(res[p], insel_orb[f"p{a}{b}_good"])


def test_synthetic_bank_all_pairs_vs_oracle(sfm, matcher):
    bank = []
    prev = None
    for i, n in enumerate((700, 517, 300, 1024, 0, 1, 257)):    # ragged, empty and single-row images
        prev_ok = prev if prev is not None and prev.shape[0] else None
        d = workloads.sift_like_image(i, n, prev_ok)
        bank.append(d)
        prev = d
    pairs = sfm.select_pairs(len(bank), 0, 0)
    matcher.upload_bank(bank)
    exp = orc.match_pairs(bank, pairs, NORM_L2)
    for eng in (sfm.ENGINE_TENSOR, sfm.ENGINE_TENSOR_IMAD, sfm.ENGINE_SIMT):
        res = matcher.match_pairs(pairs, NORM_L2, engine=eng)
        assert len(res) == len(pairs)
        for p in range(len(pairs)):
            assert orc.dmatch_equal(res[p], exp[p]), (eng, pairs[p])
    # post filters of SfM::calculateShotMatches
    exp = orc.match_pairs(bank, pairs, NORM_L2, distinct=True, min_match_count=20)
    res = matcher.match_pairs(pairs, NORM_L2, distinct=True, min_match_count=20)
    for p in range(len(pairs)):
        if exp[p] is None:
            assert res[p] is None
        else:
            assert orc.dmatch_equal(res[p], exp[p])
    # match() semantics (k = 1 keeps everything)
    res = matcher.match_pairs(pairs[:3], NORM_L2, k=1)
    for p in range(3):
        idx, dist = orc.knn2_l2(bank[pairs[p][0]], bank[pairs[p][1]])
        assert orc.dmatch_equal(res[p], orc.match_k1(idx, dist))


def test_value_only_kernel_ties_and_fallback(sfm, matcher):
    """The value-only tcgen05 kernel orders rows by D = ab - (|b|^2 >> 1); exact ties in D (duplicated rows, zero
    rows, several chunks with the same maximum) must be resolved exactly by the refine pass, and banks whose norms
    exceed the digit range must fall back to the packed-key kernel.  Bit-exact against the oracle either way."""
    adv = workloads.adversarial_sift()
    rng = np.random.default_rng(11)
    base = adv["base"]
    # many chunks containing the same rows: every chunk maximum ties -> 'ambiguous' brute-force path
    tiled = np.concatenate([base[:40]] * 12 + [np.zeros((5, 128), np.uint8)] + [base[:33]] * 3)
    near = np.clip(base.astype(np.int32) + rng.integers(-1, 2, size=base.shape), 0, 255).astype(np.uint8)
    bank = [base, tiled, near, adv["dup"], adv["zeros"], adv["one"], adv["two"], adv["n129"]]
    pairs = sfm.select_pairs(len(bank), 0, 0)
    pairs = np.concatenate([pairs, pairs[:, ::-1]])            # both directions
    matcher.upload_bank(bank)
    exp = orc.match_pairs(bank, pairs, NORM_L2, ratio=0.95)
    for eng in (sfm.ENGINE_TENSOR, sfm.ENGINE_TENSOR_IMAD):
        res = matcher.match_pairs(pairs, NORM_L2, ratio=0.95, engine=eng)
        for p in range(len(pairs)):
            assert orc.dmatch_equal(res[p], exp[p]), (eng, pairs[p].tolist())
    # the norm-less variant forced onto this bank (zero rows, duplicates, one-row images: its bounds are useless here,
    # its certificate / brute-force fallback must still give the exact lists)
    import os
    os.environ["SFM_TCV_NORMLESS"] = "2"
    try:
        m2 = sfm.Matcher(0)
        m2.upload_bank(bank)
        for ratio in (0.95, 0.7, 1.0, 0.0):
            exp_r = exp if ratio == 0.95 else orc.match_pairs(bank, pairs, NORM_L2, ratio=ratio)
            res = m2.match_pairs(pairs, NORM_L2, ratio=ratio)
            for p in range(len(pairs)):
                assert orc.dmatch_equal(res[p], exp_r[p]), ("normless", ratio, pairs[p].tolist())
        assert m2.float_stats()["rows_reranked"] > 0
        m2.close()
    finally:
        del os.environ["SFM_TCV_NORMLESS"]
    # ratio 1.0 / 0.0 edge cases of the provisional bound
    for ratio in (1.0, 0.0, 0.3):
        exp = orc.match_pairs(bank, pairs[:10], NORM_L2, ratio=ratio)
        res = matcher.match_pairs(pairs[:10], NORM_L2, ratio=ratio)
        for p in range(10):
            assert orc.dmatch_equal(res[p], exp[p]), (ratio, pairs[p].tolist())
    # large norms (|b|^2 up to 128*255^2): digit range exceeded -> packed-key kernel, still exact
    big = [rng.integers(0, 256, size=(300, 128), dtype=np.uint8), rng.integers(0, 256, size=(257, 128), dtype=np.uint8),
           adv["sat"]]
    bp = sfm.select_pairs(3, 0, 0)
    matcher.upload_bank(big)
    exp = orc.match_pairs(big, bp, NORM_L2, ratio=0.99)
    res = matcher.match_pairs(bp, NORM_L2, ratio=0.99)
    for p in range(len(bp)):
        assert orc.dmatch_equal(res[p], exp[p])


@pytest.mark.parametrize("layout,issuers,normless,chunk", [(12, 2, 1, 64), (12, 2, 0, 64), (12, 2, 0, 32), (12, 1, 1, 32),
                                                           (14, 2, 1, 64), (14, 2, 0, 32), (21, 2, 1, 32), (21, 1, 0, 64)])
def test_value_only_kernel_epilogue_layouts(sfm, layout, issuers, normless, chunk, monkeypatch):
    """Every epilogue organisation of the value-only kernel (SFM_TCV_LAYOUT: column halves / column quarters of every
    tile, alternate tiles) and the two MMA-issuing warps give the oracle's lists bit for bit, including units with a
    single train tile (one issuing warp has nothing to do), odd tile counts and > 32768-row train images (layout 14)."""
    monkeypatch.setenv("SFM_TCV_LAYOUT", str(layout))
    monkeypatch.setenv("SFM_TCV_ISSUERS", str(issuers))
    monkeypatch.setenv("SFM_TCV_NORMLESS", str(normless))
    monkeypatch.setenv("SFM_TCV_CHUNK", str(chunk))
    m = sfm.Matcher(0)
    try:
        sizes = [700, 256, 1, 130, 2049, 513, 300]               # 1, 2, 3, 9 train tiles; units of 1..17 query blocks
        bank = workloads.sift_like_bank(len(sizes), 2100)
        bank = [b[:n] for b, n in zip(bank, sizes)]
        pairs = sfm.select_pairs(len(bank), 0, 0)
        pairs = np.concatenate([pairs, pairs[:, ::-1]])
        m.upload_bank(bank)
        exp = orc.match_pairs(bank, pairs, NORM_L2, ratio=0.8)
        res = m.match_pairs(pairs, NORM_L2, ratio=0.8)
        for p in range(len(pairs)):
            assert orc.dmatch_equal(res[p], exp[p]), (layout, pairs[p].tolist())
        if layout == 14:                                         # 40 000-row train image: 157 tiles, 4 epilogue groups
            big = workloads.sift_like_image(1, 40000, bank[0])
            q = workloads.sift_like_image(2, 600, big)
            m.upload_bank([q, big])
            exp = orc.match_pairs([q, big], [[0, 1], [1, 0]], NORM_L2)
            res = m.match_pairs([[0, 1], [1, 0]], NORM_L2)
            assert orc.dmatch_equal(res[0], exp[0]) and orc.dmatch_equal(res[1], exp[1])
            assert len(res[0]) > 100
    finally:
        m.close()


def test_adaptive_variant_switch_keeps_results(sfm):
    """The library switches between (norm-less kernel, 64-row chunks) and (norm K-step, 32-row chunks) from the share of
    rows the previous run re-ranked; dense and sparse pair lists alternate here and every run equals the oracle."""
    m = sfm.Matcher(0)
    try:
        bank = workloads.sift_like_bank(6, 1500)
        dense = np.array([(0, 1), (1, 2), (2, 3), (3, 4), (4, 5)], np.int32)       # 30 % planted matches per pair
        sparse = np.array([(0, 5), (5, 0), (1, 5), (0, 4)], np.int32)              # (almost) none
        m.upload_bank(bank)
        exp = {"dense": orc.match_pairs(bank, dense, NORM_L2), "sparse": orc.match_pairs(bank, sparse, NORM_L2)}
        for name, pl in (("dense", dense), ("dense", dense), ("sparse", sparse), ("sparse", sparse), ("dense", dense),
                         ("sparse", sparse), ("dense", dense)):
            res = m.match_pairs(pl, NORM_L2)
            for p in range(len(pl)):
                assert orc.dmatch_equal(res[p], exp[name][p]), (name, pl[p].tolist())
    finally:
        m.close()


def test_cross_check_vs_cv2_golden(sfm, matcher, insel_sift, synthetic_cv2):
    matcher.upload_bank([insel_sift[f"desc{i}"] for i in range(3)])
    res = matcher.match_pairs([[0, 1]], NORM_L2, k=1, cross_check=True)
    assert orc.dmatch_equal(res[0], insel_sift["p01_cross"])
    bank = workloads.sift_like_bank(3, 700)
    matcher.upload_bank(bank)
    res = matcher.match_pairs([[0, 1]], NORM_L2, k=1, cross_check=True)
    assert orc.dmatch_equal(res[0], synthetic_cv2["syn01_cross"])
    ob = workloads.orb_like_bank(3, 900)
    matcher.upload_bank(ob)
    for eng in (sfm.ENGINE_TENSOR, sfm.ENGINE_SIMT):
        res = matcher.match_pairs([[0, 1]], NORM_HAMMING, k=1, cross_check=True, engine=eng)
        assert orc.dmatch_equal(res[0], synthetic_cv2["orb01_cross"])


def test_batching_is_invisible(sfm):
    """A tiny staging budget forces many batches; results must be byte-identical."""
    bank = workloads.sift_like_bank(6, 600)
    pairs = sfm.select_pairs(6, 0, 0)
    out = []
    for mb in ("1", "512"):
        os.environ["SFM_STAGING_MB"] = mb
        m = sfm.Matcher(0)
        m.upload_bank(bank)
        r = m.match_pairs(pairs, NORM_L2)
        out.append((r.offsets.copy(), r.matches.copy()))
        m.close()
    os.environ.pop("SFM_STAGING_MB")
    assert np.array_equal(out[0][0], out[1][0]) and out[0][1].tobytes() == out[1][1].tobytes()


def test_full_size_pair_c3_shape(sfm, matcher):
    """One 8192 x 8192 pair of the C3 workload, full oracle comparison, plus size-independent properties."""
    a = workloads.sift_like_image(0, 8192)
    b = workloads.sift_like_image(1, 8192, a)
    idx, dist = matcher.knn_match(a, b, NORM_L2, 2, sfm.ENGINE_TENSOR)
    eidx, edist = orc.knn2_l2(a, b)
    _same_knn(idx, dist, eidx, edist)
    assert np.all(dist[:, 0] <= dist[:, 1])                      # sortedness
    sidx, sdist = matcher.knn_match(a, a, NORM_L2, 1, sfm.ENGINE_TENSOR)
    # self-match: distance 0 at (the lowest index of) an identical row
    assert np.all(sdist[:, 0] == 0) and np.all(sidx[:, 0] <= np.arange(8192))
    matcher.upload_bank([a, b])
    r1 = matcher.match_pairs([[0, 1]], NORM_L2, engine=sfm.ENGINE_TENSOR)
    r2 = matcher.match_pairs([[0, 1]], NORM_L2, engine=sfm.ENGINE_SIMT)          # two independent kernels agree
    assert r1.matches.tobytes() == r2.matches.tobytes()
    assert orc.dmatch_equal(r1[0], orc.ratio_filter(eidx, edist))
    r3 = matcher.match_pairs([[0, 1]], NORM_L2)                                   # idempotence
    assert r1.matches.tobytes() == r3.matches.tobytes()


def test_orb_30000_rows_properties(sfm, matcher):
    """N = 30000 (run-orb-sequence.sh -Pfeature-limit=30000): sampled oracle rows + planted-match recovery."""
    a = workloads.orb_like_image(0, 30000)
    b = workloads.orb_like_image(1, 30000, a)
    eidx, edist = orc.knn2_hamming(b[:2048], a)
    for eng in (sfm.ENGINE_TENSOR, sfm.ENGINE_SIMT):
        idx, dist = matcher.knn_match(b[:2048], a, NORM_HAMMING, 2, eng)
        _same_knn(idx, dist, eidx, edist)
    matcher.upload_bank([a, b])
    r = matcher.match_pairs([[1, 0]], NORM_HAMMING)
    r_simt = matcher.match_pairs([[1, 0]], NORM_HAMMING, engine=sfm.ENGINE_SIMT)
    assert r.matches.tobytes() == r_simt.matches.tobytes()        # tcgen05 and __popc kernels agree on all 30000 rows
    got = r[0]
    # planted copies (first 30 % of b, 20 bit flips) must survive the ratio test
    assert (got["queryIdx"] < 9000).sum() >= 8900 and np.all(got["distance"][got["queryIdx"] < 9000] <= 20)


def test_cli_keeps_reference_switches(sfm, insel_sift, tmp_path):
    cli = os.path.join(ROOT, "sfm-mvs-pipeline_b200", "sfm_match_cli")
    if not os.path.exists(cli):
        pytest.skip("cli not built")
    bank = [insel_sift[f"desc{i}"].astype(np.float32) for i in range(3)]
    path = tmp_path / "bank.sfmd"
    with open(path, "wb") as f:
        f.write(b"SFMD" + np.array([1, 3, 128, 5], np.uint32).tobytes())
        for d in bank:
            f.write(np.uint32(d.shape[0]).tobytes() + d.tobytes())
    out = tmp_path / "m.bin"
    r = subprocess.run([cli, f"-Pdescriptors={path}", "-Pfeature-detector=SIFT", "-Pfeature-matcher=FLANN",
                        "-Pfeature-sequence=2", "-Pfeature-gridlength=3", "-Pmatch-threshold=20", f"-Pout={out}"],
                       capture_output=True, text=True, timeout=120)
    assert "pairs=2 kept=2 matches=420" in r.stdout, r.stdout + r.stderr     # 205 + 215 (Appendix B)
    raw = open(out, "rb").read()
    assert np.frombuffer(raw[:8], np.uint64)[0] == 2
    n0 = int(np.frombuffer(raw[16:24], np.uint64)[0])
    m0 = np.frombuffer(raw[24:24 + 16 * n0], orc.DMATCH_DTYPE)
    assert orc.dmatch_equal(m0, insel_sift["p01_good"])


def test_randomised_ragged_banks_all_engines(sfm, matcher):
    """Fuzz: random image sizes (0 .. ~2500 rows, around every tile/chunk boundary), random pair lists, every engine,
    against the C restatement of the oracle (threads over pairs).  Bit-exact."""
    from oracle import oracle_c as oc
    rng = np.random.default_rng(2024)
    edge = [0, 1, 2, 31, 32, 33, 127, 128, 129, 255, 256, 257, 511, 513, 1023, 1025]
    for trial in range(4):
        n_img = 7
        sizes = [int(rng.choice(edge)) if rng.random() < 0.5 else int(rng.integers(0, 2500)) for _ in range(n_img)]
        bank, prev = [], None
        for i, n in enumerate(sizes):
            d = workloads.sift_like_image(100 * trial + i, n, prev if prev is not None and len(prev) else None)
            bank.append(d)
            prev = d
        pairs = sfm.select_pairs(n_img, 0, 0)
        pairs = pairs[rng.permutation(len(pairs))]
        pairs = np.concatenate([pairs, pairs[:5, ::-1]])
        ne = [p for p in pairs if sizes[p[0]] > 0 and sizes[p[1]] > 0]
        exp = dict(zip(map(tuple, ne), oc.match_pairs(bank, np.array(ne), 4, 0.8))) if ne else {}
        matcher.upload_bank(bank)
        for eng in (sfm.ENGINE_TENSOR, sfm.ENGINE_TENSOR_IMAD, sfm.ENGINE_SIMT):
            res = matcher.match_pairs(pairs, NORM_L2, ratio=0.8, engine=eng)
            for k, p in enumerate(map(tuple, pairs)):
                want = exp.get(p, np.zeros(0, orc.DMATCH_DTYPE))
                assert orc.dmatch_equal(res[k], want), (trial, eng, p, sizes)
        # ORB-like with the same ragged sizes, both Hamming engines
        ob, prev = [], None
        for i, n in enumerate(sizes):
            d = workloads.orb_like_image(100 * trial + i, n, prev if prev is not None and len(prev) else None)
            ob.append(d)
            prev = d
        exp = dict(zip(map(tuple, ne), oc.match_pairs(ob, np.array(ne), 6, 0.8))) if ne else {}
        matcher.upload_bank(ob)
        for eng in (sfm.ENGINE_TENSOR, sfm.ENGINE_SIMT):
            res = matcher.match_pairs(pairs, NORM_HAMMING, ratio=0.8, engine=eng)
            for k, p in enumerate(map(tuple, pairs)):
                want = exp.get(p, np.zeros(0, orc.DMATCH_DTYPE))
                assert orc.dmatch_equal(res[k], want), (trial, eng, p, sizes)


def test_pipelined_from_host_equals_two_call_form(sfm, matcher):
    """sfm_match_pairs_from_host (upload of image groups overlapped with matching, pairs scheduled by availability,
    lists re-ordered on the device) must return byte-identical results to sfm_bank_upload + sfm_match_pairs."""
    import torch
    rng = np.random.default_rng(5)
    sizes = [int(x) for x in rng.integers(0, 900, size=21)] + [0, 1, 257]
    bank, prev = [], None
    for i, n in enumerate(sizes):
        d = workloads.sift_like_image(i, n, prev if prev is not None and len(prev) else None)
        bank.append(d)
        prev = d
    pairs = sfm.select_pairs(len(bank), 0, 0)
    pairs = pairs[rng.permutation(len(pairs))]
    matcher.upload_bank(bank)
    ref = matcher.match_pairs(pairs, NORM_L2, min_match_count=3)
    for dtype in (np.float32, np.uint8):
        pinned = [torch.from_numpy(b.astype(dtype)).pin_memory() for b in bank]
        arrs = [p.numpy() for p in pinned]
        got = matcher.match_pairs_from_host(arrs, pairs, NORM_L2, min_match_count=3)          # pipelined path
        assert np.array_equal(got.offsets, ref.offsets) and got.matches.tobytes() == ref.matches.tobytes()
        assert np.array_equal(got.dropped, ref.dropped)
        again = matcher.match_pairs(pairs, NORM_L2, min_match_count=3)                        # the bank stays resident
        assert again.matches.tobytes() == ref.matches.tobytes()
    got = matcher.match_pairs_from_host([b.astype(np.float32) for b in bank], pairs, NORM_L2, min_match_count=3)
    assert got.matches.tobytes() == ref.matches.tobytes()                                    # pageable -> sequential path
    # optimistic assumption violated (non-integer floats / huge norms): falls back, still equals the two-call result
    fl = [torch.from_numpy((b.astype(np.float32) * 0.37)).pin_memory() for b in bank]
    matcher.upload_bank([f.numpy() for f in fl])
    ref2 = matcher.match_pairs(pairs[:40], NORM_L2)
    got2 = matcher.match_pairs_from_host([f.numpy() for f in fl], pairs[:40], NORM_L2)
    assert got2.matches.tobytes() == ref2.matches.tobytes()
    big = [torch.from_numpy(rng.integers(0, 256, size=(max(n, 1), 128), dtype=np.uint8)).pin_memory() for n in sizes]
    matcher.upload_bank([x.numpy() for x in big])
    ref3 = matcher.match_pairs(pairs[:40], NORM_L2, ratio=0.95)
    got3 = matcher.match_pairs_from_host([x.numpy() for x in big], pairs[:40], NORM_L2, ratio=0.95)
    assert got3.matches.tobytes() == ref3.matches.tobytes()


# ------------------------------------------------------------------ BASELINE shapes C5 / C4 under the default settings
def test_c5_shape_pair_default_layout_vs_oracle(sfm):
    """C5's shape: 16 384-row images under the DEFAULT epilogue layout (12: two groups, 32-bit keys number 512 chunks per
    warp) against the C oracle — both directions of one pair, plus the pair against a ragged 16 001-row image."""
    from oracle import oracle_c
    m = sfm.Matcher(0)
    try:
        bank = workloads.sift_like_bank(3, 16384)
        bank[2] = bank[2][:16001]
        pairs = np.array([(0, 1), (1, 0), (1, 2), (2, 1)], np.int32)
        m.upload_bank([b.astype(np.float32) for b in bank])
        res = m.match_pairs(pairs, NORM_L2)
        exp = oracle_c.match_pairs(bank, pairs, NORM_L2)
        for p in range(len(pairs)):
            assert orc.dmatch_equal(res[p], exp[p]), pairs[p].tolist()
        assert len(res[0]) > 1000
    finally:
        m.close()


@pytest.mark.parametrize("settings", [{}, {"SFM_TCV_NORMLESS": "0", "SFM_TCV_CHUNK": "32"}, {"SFM_TCV_NORMLESS": "0", "SFM_TCV_CHUNK": "64"},
                                      {"SFM_TCV_NORMLESS": "2", "SFM_TCV_CHUNK": "32"}, {"SFM_TCV_BACKPRESSURE": "0"}, {"SFM_TCV_DIVERT_TEST": "1"}])
def test_c4_shape_grid_list_dense_branch(sfm, settings, monkeypatch):
    """C4's shape: 4 096-row images, grid pairing (neighbouring shots share 30 % planted rows: the DENSE branch, 7 % of the
    rows re-ranked).  64 images as an 8 x 8 grid, sequenceLength 3: every variant the adaptive choice can land on gives the
    same bytes; a sample of the pairs is checked against the C oracle, the whole list against the independent CUDA-core
    (dp4a) engine."""
    from oracle import oracle_c
    for k, v in settings.items():
        monkeypatch.setenv(k, v)
    m = sfm.Matcher(0)
    try:
        bank = workloads.sift_like_bank(64, 4096)
        pairs = sfm.select_pairs(64, 3, 8)
        assert len(pairs) == len(orc.select_pairs(64, 3, 8)) and len(pairs) > 200
        m.upload_bank(bank)
        res = m.match_pairs(pairs, NORM_L2)
        res2 = m.match_pairs(pairs, NORM_L2)                    # second run: the adaptive feedback of the first has arrived
        assert np.array_equal(res.offsets, res2.offsets) and res.matches.tobytes() == res2.matches.tobytes()
        simt = m.match_pairs(pairs, NORM_L2, engine=sfm.ENGINE_SIMT)
        assert np.array_equal(res.offsets, simt.offsets) and res.matches.tobytes() == simt.matches.tobytes()
        rng = np.random.default_rng(4)
        sample = np.sort(rng.choice(len(pairs), 12, replace=False))
        exp = oracle_c.match_pairs(bank, pairs[sample], NORM_L2)
        for k, p in enumerate(sample):
            assert orc.dmatch_equal(res[p], exp[k]), pairs[p].tolist()
        assert int(res.offsets[-1]) > 20000                     # the neighbours really share matches
    finally:
        m.close()


def test_from_host_failure_leaves_no_half_built_bank(sfm):
    """A failing sfm_match_pairs_from_host must not leave a bank that later calls would match garbage against."""
    import torch
    m = sfm.Matcher(0)
    try:
        bank = [torch.from_numpy(b.astype(np.float32)).pin_memory().numpy() for b in workloads.sift_like_bank(16, 300)]
        bad = np.array([(0, 1), (3, 99)], np.int32)
        with pytest.raises(sfm.SfmError):
            m.match_pairs_from_host(bank, bad, NORM_L2)
        with pytest.raises(sfm.SfmError) as e:
            m.match_pairs([[0, 1]], NORM_L2)
        assert e.value.code == sfm.ERR_STATE
        good = m.match_pairs_from_host(bank, [[0, 1], [1, 2]], NORM_L2)
        exp = orc.match_pairs([b.astype(np.uint8) for b in bank], [[0, 1], [1, 2]], NORM_L2)
        assert orc.dmatch_equal(good[0], exp[0]) and orc.dmatch_equal(good[1], exp[1])
    finally:
        m.close()


@pytest.mark.parametrize("settings", [{}, {"SFM_TCV_NORMLESS": "2"}, {"SFM_TCV_NORMLESS": "0"}, {"SFM_TCV_INKERNEL_REFINE": "0"},
                                      {"SFM_MIN_BATCHES": "1"}, {"SFM_TCV_BACKPRESSURE": "0"}, {"SFM_TCV_DIVERT_TEST": "1"}])
def test_ragged_scene_first_run_every_variant(sfm, settings, monkeypatch):
    """The scene of the multi-GPU worker (ragged sizes, an empty image, 30 % planted neighbours, min-match-count 20): the FIRST
    run on a fresh context and the following ones (adaptive feedback switches the kernel variant) against the C oracle."""
    from oracle import oracle_c
    for k, v in settings.items():
        monkeypatch.setenv(k, v)
    sizes = [1500, 1024, 777, 2048, 300, 1300, 0, 640]
    bank, prev = [], None
    for i, n in enumerate(sizes):
        d = workloads.sift_like_image(i, n, prev if prev is not None and len(prev) else None)
        bank.append(d)
        prev = d
    pairs = sfm.select_pairs(len(sizes), 0, 0)
    ne = np.array([p for p in pairs if sizes[p[0]] and sizes[p[1]]], np.int32)
    exp = dict(zip(map(tuple, ne.tolist()), oracle_c.match_pairs(bank, ne, NORM_L2)))
    m = sfm.Matcher(0)
    try:
        m.upload_bank(bank)
        for run in range(3):
            res = m.match_pairs(pairs, NORM_L2, min_match_count=20)
            for p, (l, r) in enumerate(pairs.tolist()):
                e = exp.get((l, r), np.zeros(0, orc.DMATCH_DTYPE))
                if len(e) < 20:
                    assert res.dropped[p] == 1 and res[p] is None, (run, l, r)
                else:
                    assert orc.dmatch_equal(res[p], e), (run, l, r, len(res[p]), len(e))
    finally:
        m.close()
