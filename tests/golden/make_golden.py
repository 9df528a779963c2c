"""Generate the committed golden fixtures under tests/golden/.

Run in the BUILD CONTAINER only (needs /root/reference/images/insel and cv2):
    python tests/golden/make_golden.py
Everything written here comes from the OpenCV routines the reference calls
(cv2 4.13.0: SIFT_create(0,3,0.09) / ORB_create(30000) as in
PhotogrammetrieCli.cpp:345-354, detect() then compute() as in SfM.cpp:584-588,
BFMatcher.knnMatch / batchDistance as in Unordered...cpp:51) — NOT from the
numpy oracle — so tests can pin the oracle and the CUDA path against them on a
box where /root/reference does not exist.
"""
import hashlib
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import cv2_ref  # noqa: E402
from oracle.oracle_np import NORM_HAMMING, NORM_L2  # noqa: E402
import workloads  # noqa: E402

IMG = "/root/reference/images/insel"


def extract(det):
    descs = []
    for i in (1, 2, 3):
        img = cv2.imread(f"{IMG}/{i}.jpg", cv2.IMREAD_GRAYSCALE)
        d = cv2.SIFT_create(0, 3, 0.09) if det == "SIFT" else cv2.ORB_create(30000)
        kp = d.detect(img, None)
        kp, desc = d.compute(img, kp)
        descs.append(desc)
    return descs


def pair_record(q, t, norm, out, tag):
    nidx, dist = cv2_ref.batch_distance_k2(q, t, norm)
    knn = cv2_ref.bf_knn_match(q, t, norm)
    # batchDistance == BFMatcher.knnMatch: assert while generating
    for r, m in enumerate(knn):
        for k, x in enumerate(m):
            assert x.trainIdx == nidx[r, k] and np.float32(x.distance) == dist[r, k]
    good = cv2_ref.ratio_good(knn, 0.7)
    out[f"{tag}_nidx"] = nidx
    out[f"{tag}_dist"] = dist
    out[f"{tag}_good"] = good
    return good


def main():
    sha = {}
    # --- insel fixtures (BASELINE configs C1 / C2)
    for det, norm in (("SIFT", NORM_L2), ("ORB", NORM_HAMMING)):
        descs = extract(det)
        out = {}
        for i, d in enumerate(descs):
            if det == "SIFT":
                assert np.all(d == np.rint(d)) and d.min() >= 0 and d.max() <= 255
                out[f"desc{i}"] = d.astype(np.uint8)          # integer-valued CV_32F -> stored as u8
            else:
                out[f"desc{i}"] = d
        for a, b in ((0, 1), (0, 2), (1, 2)):
            good = pair_record(descs[a], descs[b], norm, out, f"p{a}{b}")
            arr = np.stack([good["queryIdx"], good["trainIdx"]], 1).astype(np.int32)
            sha[f"{det}_{a}{b}"] = (len(good), hashlib.sha1(arr.tobytes()).hexdigest()[:12])
        if det == "SIFT":
            out["p01_cross"] = cv2_ref.cross_check_match(descs[0], descs[1], norm)
        np.savez_compressed(os.path.join(HERE, f"insel_{det.lower()}.npz"), **out)
    # --- small synthetic + adversarial cases straight from cv2
    out = {}
    adv = workloads.adversarial_sift()
    for name, q, t in (("dup_base", adv["dup"], adv["base"]), ("base_dup", adv["base"], adv["dup"]),
                       ("zeros_dup", adv["zeros"], adv["dup"]), ("sat_sat", adv["sat"], adv["sat"]),
                       ("base_one", adv["base"], adv["one"]), ("base_two", adv["base"], adv["two"]),
                       ("n127_n129", adv["n127"], adv["n129"]), ("n129_n127", adv["n129"], adv["n127"])):
        pair_record(q.astype(np.float32), t.astype(np.float32), NORM_L2, out, name)
    bank = workloads.sift_like_bank(3, 700)
    for a, b in ((0, 1), (1, 2)):
        pair_record(bank[a].astype(np.float32), bank[b].astype(np.float32), NORM_L2, out, f"syn{a}{b}")
    out["syn01_cross"] = cv2_ref.cross_check_match(bank[0], bank[1], NORM_L2)
    ob = workloads.orb_like_bank(3, 900)
    obd = np.concatenate([ob[1][:50], ob[1][:50], ob[1]])   # duplicated rows -> Hamming ties
    for name, q, t in (("orb01", ob[0], ob[1]), ("orb12", ob[1], ob[2]), ("orb_dup", ob[0], obd),
                       ("orb_one", ob[0], ob[1][:1])):
        pair_record(q, t, NORM_HAMMING, out, name)
    out["orb01_cross"] = cv2_ref.cross_check_match(ob[0], ob[1], NORM_HAMMING)
    np.savez_compressed(os.path.join(HERE, "synthetic_cv2.npz"), **out)
    with open(os.path.join(HERE, "insel_sha1.txt"), "w") as f:
        for k, (n, h) in sorted(sha.items()):
            f.write(f"{k} {n} {h}\n")
    print(sha)


if __name__ == "__main__":
    main()
