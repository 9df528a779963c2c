"""Golden vectors for the feature-extraction stage (SfM::extractFeatures, SfM.cpp:577-597).

Run in the BUILD CONTAINER only (needs /root/reference/images/insel and cv2 4.13.0):
    python tests/golden/make_golden_sift.py
Writes tests/golden/sift_extract.npz with, for the reference's own photograph images/insel/1.jpg (grey, as
Shot::loadImage + cv::SIFT see it) and for workloads.synthetic_photo(0, 240, 320):
    the grey image (insel only; the synthetic one is regenerated from its seed),
    cv2's keypoints after detect() and descriptors after compute(), called separately like the reference does,
    with the detector of PhotogrammetrieCli.cpp:342-357, cv::SIFT::create(featureLimit, 3, 0.09) (limit not reached, and a
    limit of 100 that is), and with OpenCV's defaults.
Nothing here comes from the numpy restatement (oracle/sift_np.py): the tests pin the restatement — and through it the
CUDA path — against these arrays on a box where /root/reference does not exist.
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import workloads  # noqa: E402

KP_FIELDS = ("x", "y", "size", "angle", "response", "octave")


def extract(img, contrast, nfeatures=0):
    det = cv2.SIFT_create(nfeatures, 3, contrast)
    kp = det.detect(img, None)
    kp, desc = det.compute(img, kp)
    arr = np.array([(k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave) for k in kp],
                   dtype=[("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"), ("octave", "<i4")])
    assert desc.dtype == np.float32 and np.array_equal(desc, np.rint(desc)) and desc.max() <= 255
    return arr, desc.astype(np.uint8)


def main():
    out = {}
    insel = cv2.imread("/root/reference/images/insel/1.jpg", cv2.IMREAD_GRAYSCALE)
    out["insel1_gray"] = insel
    out["insel1_kp_009"], out["insel1_desc_009"] = extract(insel, 0.09)
    # the reference passes its feature-limit as nfeatures (default 10000): retainBest, exercised here with a small limit
    out["insel1_kp_009_n100"], out["insel1_desc_009_n100"] = extract(insel, 0.09, 100)
    syn = workloads.synthetic_photo(0, 240, 320)
    out["syn0_sha_probe"] = np.array([int(syn.astype(np.int64).sum()), int((syn.astype(np.int64) * np.arange(320)).sum())])
    out["syn0_kp_004"], out["syn0_desc_004"] = extract(syn, 0.04)
    out["syn0_kp_009"], out["syn0_desc_009"] = extract(syn, 0.09)
    out["cv2_version"] = np.array(cv2.__version__)
    path = os.path.join(HERE, "sift_extract.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
