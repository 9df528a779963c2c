"""csrc/sift_core.cuh — the per-keypoint arithmetic the CUDA kernels of csrc/sift.cu run — compiled for the host
(tests/sift_host_harness.cpp) and compared with the numpy restatement of cv::SIFT on the same Gaussian pyramid.
No GPU involved: this checks the shared code here; tests/test_gpu_sift.py checks the kernels on a B200."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import workloads
from oracle import sift_np as S

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _sift_compare as sc  # noqa: E402


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("sift_harness") / "libsift_harness.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-o", out,
                           os.path.join(HERE, "sift_host_harness.cpp")])
    lib = C.CDLL(out)
    lib.harness_sift.restype = C.c_int
    return lib


def run_harness(lib, gpyr, n_oct, n_layers, ct, et=10.0, sigma=1.6, cap=20000):
    w = np.array([gpyr[o * (n_layers + 3)].shape[1] for o in range(n_oct)], np.int32)
    h = np.array([gpyr[o * (n_layers + 3)].shape[0] for o in range(n_oct)], np.int32)
    sizes = (w.astype(np.int64) * h) * (n_layers + 3)
    off = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64)
    flat = np.concatenate([np.ascontiguousarray(g, np.float32).ravel() for g in gpyr])
    kps = np.zeros(cap, S.KEYPOINT_DTYPE)
    desc = np.zeros((cap, 128), np.uint8)
    ncand = C.c_int(0)
    n = lib.harness_sift(flat.ctypes.data_as(C.c_void_p), C.c_int(n_oct), C.c_int(n_layers), w.ctypes.data_as(C.c_void_p),
                         h.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p), C.c_double(ct), C.c_double(et),
                         C.c_double(sigma), C.c_int(cap), kps.ctypes.data_as(C.c_void_p), desc.ctypes.data_as(C.c_void_p),
                         C.byref(ncand))
    assert n >= 0
    return kps[:n], desc[:n], ncand.value


@pytest.mark.parametrize("ct", [0.04, 0.09])
def test_shared_code_equals_restatement_on_the_same_pyramid(harness, ct):
    img = workloads.synthetic_photo(0, 240, 320)
    kp_o, gpyr = S.detect(img, contrast_threshold=ct, return_pyramid=True)
    desc_o = S.compute(img, kp_o, gpyr=gpyr)
    n_oct = len(gpyr) // 6
    kp_h, desc_h, _ = run_harness(harness, gpyr, n_oct, 3, ct)
    # same pyramid, same operation order: the keypoint lists agree one to one (libm vs numpy exp / pow: last-bit effects only)
    assert len(kp_h) == len(kp_o)
    assert np.array_equal(kp_h["octave"], kp_o["octave"])
    for f in ("x", "y", "response"):
        assert np.array_equal(kp_h[f], kp_o[f]), f
    assert np.allclose(kp_h["size"], kp_o["size"], rtol=1e-6) and np.allclose(kp_h["angle"], kp_o["angle"], atol=1e-3)
    d = np.abs(desc_h.astype(np.int32) - desc_o.astype(np.int32))
    assert d.max() <= 1 and (d > 0).mean() < 1e-3
    sc.assert_close(kp_o, desc_o.astype(np.uint8), kp_h, desc_h, "harness vs oracle")


def test_shared_code_against_cv2_golden(harness):
    gold = np.load(os.path.join(HERE, "golden", "sift_extract.npz"))
    img = gold["insel1_gray"]
    base = S.create_initial_image(img)
    n_oct = S.n_octaves_for(base.shape)
    gpyr = S.build_gaussian_pyramid(base, n_oct)
    kp_h, desc_h, ncand = run_harness(harness, gpyr, n_oct, 3, 0.09)
    r = sc.assert_close(gold["insel1_kp_009"], gold["insel1_desc_009"], kp_h, desc_h, "harness vs cv2")
    assert r["n_b"] == 320 and ncand > 320


@pytest.mark.parametrize("seed,shape", [(1, (97, 131)), (2, (64, 200)), (3, (150, 75))])
def test_shared_code_on_random_images_of_odd_sizes(harness, seed, shape):
    """Smoothed noise, odd sizes (ragged octave sizes, borders): the host build of the shared code and the restatement walk the
    same pyramid to the same keypoints."""
    rng = np.random.default_rng(seed)
    img = rng.random(shape).astype(np.float32)
    img = S.gaussian_blur(img, 1.5)
    img = np.rint(255 * (img - img.min()) / (img.max() - img.min())).astype(np.uint8)
    kp_o, gpyr = S.detect(img, contrast_threshold=0.03, return_pyramid=True)
    desc_o = S.compute(img, kp_o, gpyr=gpyr)
    kp_h, desc_h, ncand = run_harness(harness, gpyr, len(gpyr) // 6, 3, 0.03)
    assert len(kp_o) > 5 and len(kp_h) == len(kp_o) and ncand >= len(kp_o) // 2
    assert np.array_equal(kp_h["octave"], kp_o["octave"]) and np.array_equal(kp_h["x"], kp_o["x"]) and np.array_equal(kp_h["y"], kp_o["y"])
    assert np.abs(desc_h.astype(np.int32) - desc_o.astype(np.int32)).max() <= 1


@pytest.mark.parametrize("n_layers,ct,et,sigma", [(4, 0.04, 10.0, 1.6), (2, 0.03, 5.0, 1.6), (3, 0.04, 10.0, 1.2)])
def test_shared_code_with_other_detector_parameters(harness, n_layers, ct, et, sigma):
    """nOctaveLayers / edgeThreshold / sigma other than the reference's (3, 10, 1.6): restatement and shared code still agree."""
    img = workloads.synthetic_photo(4, 200, 260)
    kp_o, gpyr = S.detect(img, n_layers, ct, et, sigma, return_pyramid=True)
    desc_o = S.compute(img, kp_o, n_layers, sigma, gpyr=gpyr)
    kp_h, desc_h, _ = run_harness(harness, gpyr, len(gpyr) // (n_layers + 3), n_layers, ct, et, sigma)
    assert len(kp_o) > 100 and len(kp_h) == len(kp_o)
    assert np.array_equal(kp_h["octave"], kp_o["octave"]) and np.array_equal(kp_h["x"], kp_o["x"]) and np.array_equal(kp_h["y"], kp_o["y"])
    assert np.abs(desc_h.astype(np.int32) - desc_o.astype(np.int32)).max() <= 1
