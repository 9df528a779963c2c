"""Throughput of the device feature-extraction stage against cv2's SIFT on the host cores (the reference's
SfM::extractFeatures runs cv::SIFT under `#pragma omp parallel for`, SfM.cpp:582): prints one JSON line per image size.
Lives under tests/measure because the CPU side is the checker's reference (cv2), not the product.
    python tests/measure/sift_perf.py [n_images]"""
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
import workloads  # noqa: E402


def cpu_rate(images, contrast, threads):
    import cv2
    cv2.setNumThreads(1)                       # one image per thread, like the reference's omp loop

    def one(img):
        det = cv2.SIFT_create(0, 3, contrast)
        kp = det.detect(img, None)
        kp, d = det.compute(img, kp)
        return len(kp)
    t = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        n = list(ex.map(one, images))
    return len(images) / (time.perf_counter() - t), n


def main():
    n_images = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    sfm = ge.load_package()
    m = sfm.Matcher(0)
    threads = len(os.sched_getaffinity(0))
    shapes = ((1200, 1600), (3000, 4000)) if os.environ.get("SIFT_PERF_BIG") else ((1200, 1600),)
    for shape in shapes:
        base = [workloads.synthetic_photo(s, *shape) for s in range(2)]
        images = [base[i % 2] if i < 2 else np.ascontiguousarray(np.roll(base[i % 2], 17 * i, axis=1)) for i in range(n_images)]
        m.features_clear()
        m.extract_sift(images[0], contrast_threshold=0.09)          # warm-up: buffers, module load
        m.features_clear()
        t = time.perf_counter()
        counts = [m.extract_sift(im, contrast_threshold=0.09) for im in images]
        m.bank_from_features()                                       # descriptors + keypoints become the matcher's bank
        gpu_s = time.perf_counter() - t
        cpu_sample = images[:max(2, min(n_images, threads))]
        cpu_ips, cpu_counts = cpu_rate(cpu_sample, 0.09, threads)
        print(json.dumps({"stage": "extractFeatures (cv::SIFT(0, 3, 0.09))", "image": list(shape), "n_images": n_images,
                          "keypoints_per_image": int(np.mean(counts)), "gpu_images_per_s": round(n_images / gpu_s, 1),
                          "gpu_ms_per_image": round(1e3 * gpu_s / n_images, 3), "includes": "H2D of the grey image, bank adoption",
                          "cpu_images_per_s": round(cpu_ips, 2), "cpu_threads": threads, "cpu_sample_images": len(cpu_sample),
                          "cpu_keypoints_per_image": int(np.mean(cpu_counts)), "speedup": round(n_images / gpu_s / cpu_ips, 1)}))
    m.close()


if __name__ == "__main__":
    main()
