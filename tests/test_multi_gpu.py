"""1-GPU vs N-GPU determinism on real GPUs: needs >= 2 visible GPUs, skipped otherwise
(run it with `gpurun --gpus 2 -- python -m pytest tests -m gpu -k multi_gpu`)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_gpu_result_is_byte_identical_to_one_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    n = min(torch.cuda.device_count(), 4)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29611",
                        os.path.join(ROOT, "tests", "_mgpu_worker.py")], capture_output=True, text=True, timeout=600)
    diag = "\n".join(ln[:600] for ln in (r.stdout + "\n" + r.stderr).splitlines()
                     if "MISMATCH" in ln or "MGPU_" in ln or "Error" in ln or "rror:" in ln or "line " in ln)
    assert r.returncode == 0, diag[-6000:]
    assert "MGPU_IDENTICAL" in r.stdout


def test_single_process_group_is_byte_identical_to_one_gpu():
    """sfm_mgpu_*: one process, one worker thread per GPU, NCCL inside the library."""
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    import workloads
    sfm = ge.load_package()
    n = min(torch.cuda.device_count(), 4)
    sizes = [1500, 1024, 777, 2048, 300, 1300, 0, 640, 900, 33]
    bank, prev = [], None
    for i, k in enumerate(sizes):
        d = workloads.sift_like_image(i, k, prev if prev is not None and len(prev) else None)
        bank.append(d)
        prev = d
    pairs = sfm.select_pairs(len(sizes), 0, 0)[:-2]
    one = sfm.Matcher(0)
    one.upload_bank(bank)
    full = one.match_pairs(pairs, sfm.NORM_L2, min_match_count=5)
    g = sfm.MultiGpuMatcher(list(range(n)))
    for attempt in range(2):                       # second pass: cached deal, recycled buffers
        r1 = g.match_pairs_from_host([b.astype(np.float32) for b in bank], pairs, sfm.NORM_L2, min_match_count=5)
        r2 = g.match_pairs(pairs, sfm.NORM_L2, min_match_count=5)
        for r in (r1, r2):
            assert np.array_equal(r.offsets, full.offsets)
            assert r.matches.tobytes() == full.matches.tobytes()
            assert np.array_equal(r.dropped, full.dropped)
    # another scene with the same pair count and other row counts: the cached deal must not be reused
    bank2 = [b[: max(0, len(b) - 100 * (i % 3))] for i, b in enumerate(bank)]
    one.upload_bank(bank2)
    full2 = one.match_pairs(pairs, sfm.NORM_L2)
    g.upload_bank(bank2)
    r3 = g.match_pairs(pairs, sfm.NORM_L2)
    assert np.array_equal(r3.offsets, full2.offsets) and r3.matches.tobytes() == full2.matches.tobytes()
    # ORB through the group (whole-scene upload on every participant)
    ob = workloads.orb_like_bank(5, 700)
    op = sfm.select_pairs(5, 0, 0)
    one.upload_bank(ob)
    fo = one.match_pairs(op, sfm.NORM_HAMMING)
    ro = g.match_pairs_from_host(ob, op, sfm.NORM_HAMMING)
    assert np.array_equal(ro.offsets, fo.offsets) and ro.matches.tobytes() == fo.matches.tobytes()
    g.close()
    one.close()


def test_cli_devices_switch_is_byte_identical(tmp_path):
    """sfm_match_cli -Pdevices=0,1 (C++ host, one process, sfm_mgpu_*) writes the same artifact as -Pdevice=0, including
    the homography stage that runs on the gathered lists of the first device."""
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sys.path.insert(0, ROOT)
    import workloads
    cli = os.path.join(ROOT, "sfm-mvs-pipeline_b200", "sfm_match_cli")
    n_img, n_rows = 8, 900
    bank = workloads.sift_like_bank(n_img, n_rows)
    kps = workloads.sift_like_keypoints(n_img, n_rows)
    dpath, kpath = tmp_path / "bank.sfmd", tmp_path / "bank.sfmk"
    with open(dpath, "wb") as f:
        f.write(b"SFMD" + np.array([1, n_img, 128, 5], np.uint32).tobytes())
        for d in bank:
            f.write(np.uint32(d.shape[0]).tobytes() + d.astype(np.float32).tobytes())
    with open(kpath, "wb") as f:
        f.write(b"SFMK" + np.array([1, n_img], np.uint32).tobytes())
        for k in kps:
            f.write(np.array([len(k), 4000, 3000], np.uint32).tobytes() + np.ascontiguousarray(k, np.float32).tobytes())
    outs = []
    for dev in ("-Pdevice=0", "-Pdevices=0,1"):
        out = tmp_path / f"m{len(outs)}.bin"
        r = subprocess.run([cli, f"-Pdescriptors={dpath}", f"-Pkeypoints={kpath}", "-Pmatch-threshold=20", f"-Pout={out}", dev],
                           capture_output=True, text=True, timeout=300)
        assert "pairs=28" in r.stdout and "[ERROR]" not in r.stderr, r.stdout + r.stderr
        outs.append((open(out, "rb").read(), [ln for ln in r.stdout.splitlines() if "homographyInlierRatio" in ln]))
    assert outs[0][0] == outs[1][0] and len(outs[0][0]) > 10000
    assert outs[0][1] == outs[1][1] and len(outs[0][1]) >= 7


@pytest.mark.parametrize("detector", ["SIFT", "ORB"])
def test_extraction_split_over_gpus_equals_one_gpu(detector):
    """SfM::extractFeatures split over the devices (sfm_mgpu_extract_features: image i on device i % n, NCCL exchange of the feature
    sets, every device adopts the scene): every device ends up with the single-GPU feature set, and matching on top of it is
    byte-identical."""
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    import workloads
    sfm = ge.load_package()
    n = min(torch.cuda.device_count(), 4)
    base = workloads.synthetic_photo(11, 200, 280)
    imgs = [base, np.roll(base, (3, 5), axis=(0, 1)), workloads.synthetic_photo(12, 160, 240), np.full((70, 90), 9, np.uint8),
            np.roll(base, (-4, 2), axis=(0, 1))]
    one = sfm.Matcher(0)
    one.features_clear()
    kw = dict(n_features=1500) if detector == "ORB" else dict(contrast_threshold=0.09, n_features=10000)
    counts1 = [(one.extract_orb if detector == "ORB" else one.extract_sift)(im, **kw) for im in imgs]
    one.bank_from_features()
    norm = sfm.NORM_HAMMING if detector == "ORB" else sfm.NORM_L2
    pairs = sfm.select_pairs(len(imgs), 0, 0)
    full = one.match_pairs(pairs, norm)
    g = sfm.MultiGpuMatcher(list(range(n)))
    counts = g.extract_features(imgs, detector, **kw)
    assert counts == counts1
    for dev in range(n):
        c = g.ctx(dev)
        for i in range(len(imgs)):
            a, b = one.features_download(i), c.features_download(i)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), (dev, i)
    res = g.match_pairs(pairs, norm)
    assert np.array_equal(res.offsets, full.offsets) and res.matches.tobytes() == full.matches.tobytes() and int(full.offsets[-1]) > 100
    g.close()
    one.close()


def test_cli_images_over_two_devices(tmp_path):
    """sfm_match_cli -Pimage=... -Pdevices=0,1: extractFeatures, calculateShotMatches and calculateHomography with the scene split over the
    GPUs of one process print exactly what the single-device run prints."""
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sys.path.insert(0, ROOT)
    import workloads
    cli = os.path.join(ROOT, "sfm-mvs-pipeline_b200", "sfm_match_cli")
    base = workloads.synthetic_photo(11, 200, 280)
    imgs = [base, np.roll(base, (3, 5), axis=(0, 1)), np.roll(base, (-2, 4), axis=(0, 1)), workloads.synthetic_photo(12, 160, 240)]
    paths = []
    for i, im in enumerate(imgs):
        paths.append(str(tmp_path / f"shot{i}.pgm"))
        with open(paths[-1], "wb") as f:
            f.write(b"P5\n%d %d\n255\n" % (im.shape[1], im.shape[0]) + np.ascontiguousarray(im, np.uint8).tobytes())
    outs = []
    for det in ("SIFT", "ORB"):
        for dev in ("-Pdevice=0", "-Pdevices=0,1"):
            r = subprocess.run([cli, *[f"-Pimage={p}" for p in paths], f"-Pfeature-detector={det}", "-Pfeature-limit=2000", "-Pmatch-threshold=4",
                                "-Pransac-matching-threshold=-3", dev], capture_output=True, text=True, timeout=300)
            assert "[ERROR]" not in r.stderr and "pairs=6" in r.stdout, r.stdout + r.stderr
            lines = [ln for ln in r.stdout.splitlines() if "seconds" not in ln and not ln.startswith("NCCL version")]
            lines.append(r.stdout.split("keypoints=")[1].split()[0])
            outs.append(lines)
        assert outs[-2] == outs[-1], (det, outs[-2], outs[-1])
        assert any("homographyInlierRatio" in ln for ln in outs[-1])
