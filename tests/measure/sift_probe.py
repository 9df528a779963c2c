"""Stage-by-stage probe of the device feature extraction against the numpy restatement (uses the CPU checker, hence
under tests/measure): pyramid differences per level, list sizes, keypoint / descriptor agreement, timing.
    python tests/measure/sift_probe.py [height width]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402
import workloads  # noqa: E402
from oracle import sift_np as S  # noqa: E402
import _sift_compare as sc  # noqa: E402


def main():
    sfm = ge.load_package()
    m = sfm.Matcher(0)
    h, w = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (240, 320)
    img = workloads.synthetic_photo(0, h, w)
    n = m.extract_sift(img)
    print("device keypoints", n, m.features_last_counts())
    base = S.create_initial_image(img)
    n_oct = S.n_octaves_for(base.shape)
    gpyr = S.build_gaussian_pyramid(base, n_oct)
    for o in range(n_oct):
        d = []
        for i in range(6):
            got = m.pyramid_level(o, i)
            d.append(float(np.abs(got - gpyr[o * 6 + i]).max()) if got.shape == gpyr[o * 6 + i].shape else -1.0)
        print("octave", o, gpyr[o * 6].shape, "max |diff| per level", ["%.2e" % x for x in d])
    kp, desc = m.features_download(0)
    kp_o, desc_o = S.detect_and_compute(img)
    print("oracle keypoints", len(kp_o))
    print(sc.compare(kp_o, desc_o.astype(np.uint8), kp, desc))
    if len(kp) and len(kp_o):
        print("first device", kp[:3])
        print("first oracle", kp_o[:3])
    for shape in ((1200, 1600),) + (((3000, 4000),) if os.environ.get("SIFT_PROBE_BIG") else ()):
        big = workloads.synthetic_photo(1, *shape)
        m.features_clear()
        m.extract_sift(big, contrast_threshold=0.09)
        t = time.perf_counter()
        for _ in range(3):
            nk = m.extract_sift(big, contrast_threshold=0.09)
        dt = (time.perf_counter() - t) / 3
        print(shape, "keypoints", nk, m.features_last_counts(), "ms per image (host wall, incl. H2D)", round(dt * 1e3, 2))
    m.close()


if __name__ == "__main__":
    main()
