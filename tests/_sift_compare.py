"""Shared comparison of two SIFT outputs (keypoints + descriptors) under the tolerance stated for the
feature-extraction stage.  cv::SIFT is float arithmetic with data-dependent decisions (thresholds, extremum ties,
orientation peaks): two correct implementations differ in the last bits of the pyramid and may flip a borderline
decision, so parity is a tolerance, written here once:

    keypoints   matched one-to-one when |dx|, |dy| <= POS_TOL px, |dsize| <= SIZE_TOL * size, the angle agrees within
                ANGLE_TOL degrees (cyclic) and the packed octave field is equal;
                at least MIN_MATCHED of each side must be matched
    descriptors of matched keypoints: at most BAD_DESC of the rows may contain an element that differs by more than
                DESC_TOL (a one-unit flip of the u8 rounding is common, a borderline histogram vote can move a few units)
"""
import numpy as np

POS_TOL = 0.02
SIZE_TOL = 1e-3
ANGLE_TOL = 0.05
MIN_MATCHED = 0.98
DESC_TOL = 2
BAD_DESC = 0.02


def match_keypoints(a, b):
    """Greedy one-to-one matching of structured keypoint arrays (fields x, y, size, angle, octave); returns index pairs."""
    if len(a) == 0 or len(b) == 0:
        return np.zeros((0, 2), np.int64)
    order = np.argsort(b["x"], kind="stable")
    bx = b["x"][order]
    used = np.zeros(len(b), bool)
    out = []
    for i in range(len(a)):
        lo = np.searchsorted(bx, a["x"][i] - POS_TOL, "left")
        hi = np.searchsorted(bx, a["x"][i] + POS_TOL, "right")
        for j in order[lo:hi]:
            if used[j] or a["octave"][i] != b["octave"][j]:
                continue
            da = abs(float(a["angle"][i]) - float(b["angle"][j]))
            da = min(da, 360.0 - da)
            if (abs(float(a["y"][i]) - float(b["y"][j])) <= POS_TOL and da <= ANGLE_TOL and
                    abs(float(a["size"][i]) - float(b["size"][j])) <= SIZE_TOL * float(a["size"][i])):
                used[j] = True
                out.append((i, j))
                break
    return np.array(out, np.int64).reshape(-1, 2)


def compare(kp_a, desc_a, kp_b, desc_b):
    """Returns a dict of the measured agreement; assert_close() applies the tolerance."""
    m = match_keypoints(kp_a, kp_b)
    res = {"n_a": len(kp_a), "n_b": len(kp_b), "matched": len(m)}
    if len(m) and desc_a is not None and desc_b is not None:
        d = np.abs(desc_a[m[:, 0]].astype(np.int32) - desc_b[m[:, 1]].astype(np.int32))
        res["desc_max"] = int(d.max())
        res["desc_bad_rows"] = float((d.max(axis=1) > DESC_TOL).mean())
        res["desc_diff_elems"] = float((d > 0).mean())
    return res


def assert_close(kp_a, desc_a, kp_b, desc_b, what=""):
    r = compare(kp_a, desc_a, kp_b, desc_b)
    assert r["matched"] >= MIN_MATCHED * r["n_a"] and r["matched"] >= MIN_MATCHED * r["n_b"], (what, r)
    if "desc_bad_rows" in r:
        assert r["desc_bad_rows"] <= BAD_DESC, (what, r)
    return r
