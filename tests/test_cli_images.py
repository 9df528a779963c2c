"""sfm_match_cli -Pimage=...: the C++ host mirror of SfM::extractFeatures -> calculateShotMatches -> calculateHomography
(host/matching.{h,cpp}: GpuSiftFeatureDetector, MatchingStage) driven with the reference's switch grammar."""
import os
import re
import subprocess

import numpy as np
import pytest

import workloads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "sfm-mvs-pipeline_b200", "sfm_match_cli")


def write_pgm(path, img):
    with open(path, "wb") as f:
        f.write(b"P5\n# written by the test\n%d %d\n255\n" % (img.shape[1], img.shape[0]))
        f.write(np.ascontiguousarray(img, np.uint8).tobytes())


def run_cli(*args):
    return subprocess.run([CLI, *args], capture_output=True, text=True, timeout=300)


def test_pgm_errors_are_reported_like_the_reference_reports_exceptions(tmp_path):
    # main.cpp:26-34: the exception text is printed, the process returns 0
    bad = tmp_path / "bad.pgm"
    bad.write_bytes(b"P3\n4 4\n255\n" + bytes(48))
    out = run_cli(f"-Pimage={bad}")
    assert out.returncode == 0 and "not a binary PGM" in out.stderr
    short = tmp_path / "short.pgm"
    short.write_bytes(b"P5\n40 40\n255\n" + bytes(100))
    out = run_cli(f"-Pimage={short}")
    assert out.returncode == 0 and "truncated PNM" in out.stderr
    good = tmp_path / "good.pgm"
    write_pgm(good, workloads.synthetic_photo(0, 60, 80))
    colour = tmp_path / "colour.ppm"                       # binary PPM: converted with cv::SIFT's grey conversion on the host
    rgb = np.stack([workloads.synthetic_photo(s, 60, 80) for s in (1, 2, 3)], -1)
    colour.write_bytes(b"P6\n80 60\n255\n" + rgb.tobytes())
    import torch
    if not torch.cuda.is_available():
        out = run_cli(f"-Pimage={colour}")
        assert out.returncode == 0 and "no CUDA device usable" in out.stderr
        out = run_cli(f"-Pimage={good}")
        assert out.returncode == 0 and "no CUDA device usable" in out.stderr      # parsed the image, then no silent CPU path


@pytest.mark.gpu
def test_cli_extracts_matches_and_rates_the_pairs(tmp_path, sfm):
    a = workloads.synthetic_photo(11, 240, 320)
    imgs = [a, np.roll(a, (3, 5), axis=(0, 1)), workloads.synthetic_photo(12, 200, 280)]
    paths = []
    for i, im in enumerate(imgs):
        paths.append(str(tmp_path / f"shot{i}.pgm"))
        write_pgm(paths[-1], im)
    out = run_cli(*[f"-Pimage={p}" for p in paths], "-Pmatch-threshold=4", "-Pransac-matching-threshold=-3")
    assert out.returncode == 0 and "[ERROR]" not in out.stderr, out.stderr
    head = re.search(r"images=(\d+) keypoints=(\d+)", out.stdout)
    tail = re.search(r"pairs=(\d+) kept=(\d+) matches=(\d+)", out.stdout)
    assert head and tail, out.stdout
    # the same through the Python binding of the same ABI
    m = sfm.Matcher(0)
    m.features_clear()
    n_kp = sum(m.extract_sift(im, contrast_threshold=0.09, n_features=10000) for im in imgs)
    m.bank_from_features()
    res = m.match_pairs(sfm.select_pairs(3, 0, 0), sfm.NORM_L2, min_match_count=4)
    kept = [p for p in range(3) if not res.dropped[p]]
    assert (int(head.group(1)), int(head.group(2))) == (3, n_kp)
    assert (int(tail.group(1)), int(tail.group(2)), int(tail.group(3))) == (3, len(kept), sum(len(res[p]) for p in kept))
    ratios = {(l, r): float(v) for l, r, v in re.findall(r"shot(\d)\.pgm : \S*shot(\d)\.pgm -> \d+ homographyInlierRatio: (\S+)", out.stdout)}
    assert ratios[("0", "1")] > 0.8
    # the stage's on-disk artifact (PhotogrammetrieCli.cpp:174-199): one side-by-side picture per kept pair, deterministic bytes
    mdir = tmp_path / "matches"
    mdir.mkdir()
    out2 = run_cli(*[f"-Pimage={p}" for p in paths], "-Pmatch-threshold=4", "-Pransac-matching-threshold=-3", f"-Pout-matches-dir={mdir}")
    assert f"match_pictures={len(kept)}" in out2.stdout, out2.stdout + out2.stderr
    pic = (mdir / "0shot0.pgm-shot1.pgm.ppm").read_bytes()
    assert pic.startswith(b"P6\n640 240\n255\n") and len(pic) == 15 + 640 * 240 * 3
    px = np.frombuffer(pic[15:], np.uint8).reshape(240, 640, 3)
    coloured = (px[..., 0] != px[..., 1]) | (px[..., 1] != px[..., 2])                 # the grey shots carry coloured lines
    assert coloured.sum() > 50 * len(res[0]) and coloured[:, :320].any() and coloured[:, 320:].any()
    m.close()


@pytest.mark.gpu
def test_cli_orb_detector(tmp_path, sfm):
    """-Pfeature-detector=ORB -Pimage=...: cv::ORB::create(feature-limit) + NORM_HAMMING matching on the device (run-orb-sequence.sh)."""
    a = workloads.synthetic_photo(11, 240, 320)
    imgs = [a, np.roll(a, (3, 5), axis=(0, 1))]
    paths = []
    for i, im in enumerate(imgs):
        paths.append(str(tmp_path / f"shot{i}.pgm"))
        write_pgm(paths[-1], im)
    out = run_cli(*[f"-Pimage={p}" for p in paths], "-Pfeature-detector=ORB", "-Pfeature-limit=3000", "-Pfeature-sequence=2",
                  "-Pmatch-threshold=4", "-Pransac-matching-threshold=-3")
    assert out.returncode == 0 and "[ERROR]" not in out.stderr, out.stderr
    head = re.search(r"images=(\d+) keypoints=(\d+)", out.stdout)
    tail = re.search(r"pairs=(\d+) kept=(\d+) matches=(\d+)", out.stdout)
    m = sfm.Matcher(0)
    m.features_clear()
    n_kp = sum(m.extract_orb(im, n_features=3000) for im in imgs)
    m.bank_from_features()
    res = m.match_pairs([[0, 1]], sfm.NORM_HAMMING, min_match_count=4)
    assert (int(head.group(1)), int(head.group(2))) == (2, n_kp)
    assert (int(tail.group(1)), int(tail.group(2)), int(tail.group(3))) == (1, 1, len(res[0])) and len(res[0]) > 100
    m.close()
