"""ctypes binding over libsfmmatch.so (the C ABI in include/sfmmatch.h).

This module contains no compute: every call lands in the hand-written CUDA kernels through the
C ABI.  It fails loudly (ImportError / SfmError) when the library is not built or no sm_100 GPU is
usable — there is no CPU fallback.  Names mirror the reference's matching interface:
``knn_match`` <-> cv::DescriptorMatcher::knnMatch (Unordered...cpp:51), ``match_pairs`` <->
IFeatureMatchingStrategy::calculateShotMatches + SfM::calculateShotMatches' post filters
(IFeatureMatchingStrategy.h:45-46, SfM.cpp:542-575), ``select_pairs`` <-> the three strategies'
pair lists (PhotogrammetrieCli.cpp:320-340).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SFMMATCH_LIB") or os.path.join(HERE, "libsfmmatch.so")   # override: kernel experiments

NORM_L2, NORM_HAMMING = 4, 6
CV_8U, CV_32F = 0, 5
ENGINE_AUTO, ENGINE_TENSOR, ENGINE_SIMT, ENGINE_TENSOR_IMAD = 0, 1, 2, 3
MAX_ROWS = 1 << 18
OK, ERR_INVALID, ERR_CUDA, ERR_CAPACITY, ERR_UNSUPPORTED, ERR_STATE, ERR_NCCL = 0, -1, -2, -3, -4, -5, -6
DIST_ID_BYTES = 128

DMATCH_DTYPE = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"), ("distance", "<f4")])

EXPORTS = ["sfm_opts_default", "sfm_ctx_create", "sfm_ctx_destroy", "sfm_last_error", "sfm_device_sm_count",
           "sfm_ctx_stream", "sfm_bank_upload", "sfm_bank_upload_device", "sfm_bank_info", "sfm_select_pairs",
           "sfm_match_pairs", "sfm_match_pairs_enqueue", "sfm_match_pairs_collect", "sfm_result_n_pairs",
           "sfm_result_offsets", "sfm_result_matches", "sfm_result_dropped", "sfm_result_free", "sfm_last_stats",
           "sfm_knn_match", "sfm_set_profiling", "sfm_last_profile", "sfm_bank_device_ptr", "sfm_match_pairs_device_view", "sfm_match_pairs_from_host",
           "sfm_last_float_stats", "sfm_keypoints_upload", "sfm_homography_inlier_ratios",
           "sfm_homography_opts_default", "sfm_sift_opts_default", "sfm_features_clear", "sfm_features_extract_sift",
           "sfm_features_count", "sfm_features_download", "sfm_bank_from_features", "sfm_features_last_counts",
           "sfm_features_pyramid_level", "sfm_features_last_profile", "sfm_gray_from_bgr",
           "sfm_orb_opts_default", "sfm_features_extract_orb", "sfm_features_descriptor_bytes", "sfm_features_orb_level",
           "sfm_mgpu_create", "sfm_mgpu_destroy", "sfm_mgpu_device_count", "sfm_mgpu_ctx", "sfm_mgpu_last_error",
           "sfm_mgpu_bank_upload", "sfm_mgpu_match_pairs", "sfm_mgpu_match_pairs_from_host", "sfm_dist_unique_id",
           "sfm_dist_init", "sfm_dist_info", "sfm_dist_match_pairs", "sfm_dist_match_pairs_from_host",
           "sfm_dist_assign_pairs", "sfm_dist_upload_share", "sfm_dist_last_phases", "sfm_dist_features_allgather",
           "sfm_mgpu_extract_features"]

KEYPOINT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                           ("octave", "<i4")])          # sfm_keypoint = cv::KeyPoint without class_id


class SfmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"sfmmatch error {code}: {msg}")
        self.code = code


class HomographyOpts(C.Structure):
    _fields_ = [("max_iters", C.c_int32), ("refine", C.c_int32), ("confidence", C.c_double), ("seed", C.c_uint64)]


class SiftOpts(C.Structure):
    _fields_ = [("n_octave_layers", C.c_int32), ("max_keypoints", C.c_int32), ("contrast_threshold", C.c_double),
                ("edge_threshold", C.c_double), ("sigma", C.c_double), ("n_features", C.c_int32), ("reserved", C.c_int32)]


class OrbOpts(C.Structure):
    _fields_ = [("n_features", C.c_int32), ("max_keypoints", C.c_int32), ("n_levels", C.c_int32), ("edge_threshold", C.c_int32),
                ("patch_size", C.c_int32), ("fast_threshold", C.c_int32), ("scale_factor", C.c_float), ("reserved", C.c_int32)]


class Opts(C.Structure):
    _fields_ = [("norm", C.c_int32), ("k", C.c_int32), ("ratio", C.c_double), ("cross_check", C.c_int32),
                ("distinct", C.c_int32), ("min_match_count", C.c_int32), ("engine", C.c_int32)]


def load_library():
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is not built (run `make -C {HERE}` or __graft_entry__.build()); "
                          "there is no fallback implementation")
    lib = C.CDLL(LIB_PATH)
    lib.sfm_last_error.restype = C.c_char_p
    lib.sfm_ctx_stream.restype = C.c_void_p
    lib.sfm_result_n_pairs.restype = C.c_int64
    lib.sfm_result_offsets.restype = C.POINTER(C.c_int64)
    lib.sfm_result_matches.restype = C.c_void_p
    lib.sfm_result_dropped.restype = C.POINTER(C.c_uint8)
    lib.sfm_result_free.restype = None
    lib.sfm_ctx_destroy.restype = None
    lib.sfm_opts_default.restype = None
    lib.sfm_homography_opts_default.restype = None
    lib.sfm_orb_opts_default.restype = None
    lib.sfm_mgpu_destroy.restype = None
    lib.sfm_mgpu_ctx.restype = C.c_void_p
    lib.sfm_mgpu_last_error.restype = C.c_char_p
    return lib


_lib = load_library()


def select_pairs(n_shots: int, feature_sequence: int = 0, feature_gridlength: int = 0) -> np.ndarray:
    n = C.c_int64(0)
    rc = _lib.sfm_select_pairs(C.c_int(n_shots), C.c_int(feature_sequence), C.c_int(feature_gridlength), None,
                               C.c_int64(0), C.byref(n))
    if rc != OK:
        raise SfmError(rc, "invalid pairing parameters")
    out = np.zeros((n.value, 2), np.int32)
    _lib.sfm_select_pairs(C.c_int(n_shots), C.c_int(feature_sequence), C.c_int(feature_gridlength),
                          out.ctypes.data_as(C.c_void_p), C.c_int64(n.value), C.byref(n))
    return out


def gray_from_bgr(img: np.ndarray, rgb_order: bool = False) -> np.ndarray:
    """The grey image cv::SIFT derives from a colour photograph (cvtColor BGR2GRAY, 8 bit); host arithmetic, no GPU."""
    img = np.asarray(img)
    if img.dtype != np.uint8 or img.ndim != 3 or img.shape[2] not in (3, 4):
        raise SfmError(ERR_INVALID, "gray_from_bgr needs a uint8 [rows, cols, 3 or 4] image")
    if img.size and (img.strides[2] != 1 or img.strides[1] != img.shape[2]):
        img = np.ascontiguousarray(img)
    out = np.empty(img.shape[:2], np.uint8)
    rc = _lib.sfm_gray_from_bgr(C.c_void_p(img.ctypes.data if img.size else None), C.c_int(img.shape[0]), C.c_int(img.shape[1]),
                                C.c_size_t(img.strides[0] if img.shape[0] > 1 else 0), C.c_int(img.shape[2]), C.c_int(int(rgb_order)),
                                C.c_void_p(out.ctypes.data if out.size else None), C.c_size_t(0))
    if rc != OK:
        raise SfmError(rc, "sfm_gray_from_bgr failed")
    return out


def dist_unique_id() -> bytes:
    """Id of a new multi-GPU group (made by participant 0, handed to the others by the launcher)."""
    buf = (C.c_uint8 * DIST_ID_BYTES)()
    rc = _lib.sfm_dist_unique_id(buf)
    if rc != OK:
        raise SfmError(rc, _lib.sfm_last_error(None).decode())
    return bytes(buf)


def dist_assign_pairs(pairs, n_rows, world: int) -> np.ndarray:
    """owner[p] = participant that matches pair p (the library's deal; host arithmetic, no GPU)."""
    pairs = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)
    n_rows = np.ascontiguousarray(n_rows, np.int32)
    owner = np.zeros(len(pairs), np.int32)
    rc = _lib.sfm_dist_assign_pairs(pairs.ctypes.data_as(C.c_void_p), C.c_int64(len(pairs)), n_rows.ctypes.data_as(C.c_void_p),
                                    C.c_int(len(n_rows)), C.c_int(world), owner.ctypes.data_as(C.c_void_p))
    if rc != OK:
        raise SfmError(rc, "sfm_dist_assign_pairs: bad arguments")
    return owner


def dist_upload_share(n_rows, world: int, rank: int):
    n_rows = np.ascontiguousarray(n_rows, np.int32)
    a, b = C.c_int(), C.c_int()
    rc = _lib.sfm_dist_upload_share(n_rows.ctypes.data_as(C.c_void_p), C.c_int(len(n_rows)), C.c_int(world), C.c_int(rank),
                                    C.byref(a), C.byref(b))
    if rc != OK:
        raise SfmError(rc, "sfm_dist_upload_share: bad arguments")
    return a.value, b.value


def _depth_of(a: np.ndarray) -> int:
    if a.dtype == np.uint8:
        return CV_8U
    if a.dtype == np.float32:
        return CV_32F
    raise SfmError(ERR_INVALID, f"descriptor dtype {a.dtype} is neither CV_8U nor CV_32F")


class MatchResult:
    """Per-pair DMatch lists in input pair order (ShotMatches without the shot pointers)."""

    def __init__(self, offsets, matches, dropped):
        self.offsets, self.matches, self.dropped = offsets, matches, dropped

    def __len__(self):
        return len(self.offsets) - 1

    def __getitem__(self, p):
        if self.dropped[p]:
            return None
        return self.matches[self.offsets[p]:self.offsets[p + 1]]

    def counts(self):
        return np.diff(self.offsets)


class MatchResultView:
    """The library-owned result itself: numpy VIEWS of the pinned host buffers the lists were copied into (no second copy).
    Valid until release() — call it (or use `with`) before the producing context is closed."""

    def __init__(self, res):
        self._res = res
        n = _lib.sfm_result_n_pairs(res)
        self.offsets = np.ctypeslib.as_array(_lib.sfm_result_offsets(res), shape=(n + 1,))
        total = int(self.offsets[n])
        self.dropped = np.ctypeslib.as_array(_lib.sfm_result_dropped(res), shape=(n,)) if n else np.zeros(0, np.uint8)
        if total:
            buf = (C.c_char * (total * 16)).from_address(_lib.sfm_result_matches(res))
            self.matches = np.frombuffer(buf, dtype=DMATCH_DTYPE, count=total)
        else:
            self.matches = np.zeros(0, DMATCH_DTYPE)

    def release(self):
        if self._res:
            self.offsets = self.matches = self.dropped = None
            _lib.sfm_result_free(self._res)
            self._res = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.release()

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


class Matcher:
    """One context = one GPU.  Mirrors the role of the injected cv::Ptr<cv::DescriptorMatcher> plus the
    strategy object of the reference (SfM.cpp:52-65)."""

    def __init__(self, device: int = 0):
        self._ctx = C.c_void_p()
        rc = _lib.sfm_ctx_create(C.byref(self._ctx), C.c_int(device))
        if rc != OK:
            raise SfmError(rc, _lib.sfm_last_error(None).decode())
        self.device = device

    def close(self):
        if self._ctx:
            _lib.sfm_ctx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != OK:
            raise SfmError(rc, _lib.sfm_last_error(self._ctx).decode())

    @property
    def sm_count(self):
        return int(_lib.sfm_device_sm_count(self._ctx))

    @property
    def stream(self) -> int:
        return int(_lib.sfm_ctx_stream(self._ctx) or 0)

    def stats(self):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        _lib.sfm_last_stats(self._ctx, C.byref(a), C.byref(b), C.byref(c))
        return {"kernel_launches": a.value, "h2d_bytes": b.value, "d2h_bytes": c.value}

    def float_stats(self):
        """Non-integer CV_32F runs: rows re-ranked in fp32 / rows that failed the certificate (brute-forced)."""
        a, b = C.c_int64(), C.c_int64()
        self._check(_lib.sfm_last_float_stats(self._ctx, C.byref(a), C.byref(b)))
        return {"rows_reranked": a.value, "rows_brute_forced": b.value}

    def set_profiling(self, on: bool):
        self._check(_lib.sfm_set_profiling(self._ctx, C.c_int(1 if on else 0)))

    def last_profile(self):
        a, b, n = C.c_double(), C.c_double(), C.c_int()
        self._check(_lib.sfm_last_profile(self._ctx, C.byref(a), C.byref(b), C.byref(n)))
        return {"knn_ms": a.value, "post_ms": b.value, "knn_launches": n.value}

    # ---- bank
    def upload_bank(self, descriptors):
        """descriptors: list of 2-D numpy arrays (one per shot), all uint8 or all float32, same width;
        rows may be strided (like a cv::Mat with step > cols*elemSize)."""
        n = len(descriptors)
        if n == 0:
            self._check(_lib.sfm_bank_upload(self._ctx, 0, None, None, C.c_int(128), None, C.c_int(CV_8U)))
            return
        depth = _depth_of(descriptors[0])
        cols = descriptors[0].shape[1]
        keep = []
        for d in descriptors:
            if d.ndim != 2 or d.shape[1] != cols or _depth_of(d) != depth:
                raise SfmError(ERR_INVALID, "all descriptor matrices must share dtype and width")
            if d.strides[1] != d.itemsize:
                d = np.ascontiguousarray(d)
            keep.append(d)
        ptrs = (C.c_void_p * n)(*[d.ctypes.data if d.shape[0] else None for d in keep])
        nrows = (C.c_int32 * n)(*[d.shape[0] for d in keep])
        steps = (C.c_size_t * n)(*[d.strides[0] if d.shape[0] > 1 else cols * d.itemsize for d in keep])
        self._check(_lib.sfm_bank_upload(self._ctx, C.c_int(n), ptrs, nrows, C.c_int(cols), steps, C.c_int(depth)))

    def upload_bank_device(self, dev_ptr: int, row_offset, n_rows, cols: int, depth: int):
        n = len(n_rows)
        ro = (C.c_int64 * n)(*[int(x) for x in row_offset])
        nr = (C.c_int32 * n)(*[int(x) for x in n_rows])
        self._check(_lib.sfm_bank_upload_device(self._ctx, C.c_int(n), C.c_void_p(dev_ptr), ro, nr, C.c_int(cols),
                                                C.c_int(depth)))

    def bank_device_ptr(self, image: int):
        p, n = C.c_void_p(), C.c_int32()
        self._check(_lib.sfm_bank_device_ptr(self._ctx, C.c_int(image), C.byref(p), C.byref(n)))
        return int(p.value or 0), n.value

    def bank_info(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        _lib.sfm_bank_info(self._ctx, C.byref(a), C.byref(b), C.byref(c))
        return {"n_images": a.value, "cols": b.value, "u8_valued": bool(c.value)}

    # ---- the stage
    def _opts(self, norm, k, ratio, cross_check, distinct, min_match_count, engine):
        o = Opts()
        _lib.sfm_opts_default(C.byref(o), C.c_int32(norm))
        o.k, o.ratio, o.cross_check, o.distinct = k, ratio, int(cross_check), int(distinct)
        o.min_match_count, o.engine = min_match_count, engine
        return o

    def enqueue(self, pairs, norm, k=2, ratio=0.7, cross_check=False, distinct=False, min_match_count=0,
                engine=ENGINE_AUTO):
        pairs = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)
        o = self._opts(norm, k, ratio, cross_check, distinct, min_match_count, engine)
        self._check(_lib.sfm_match_pairs_enqueue(self._ctx, pairs.ctypes.data_as(C.c_void_p), C.c_int64(len(pairs)),
                                                 C.byref(o)))

    def _bank_args(self, descriptors):
        n = len(descriptors)
        depth = _depth_of(descriptors[0])
        cols = descriptors[0].shape[1]
        keep = []
        for d in descriptors:
            if d.ndim != 2 or d.shape[1] != cols or _depth_of(d) != depth:
                raise SfmError(ERR_INVALID, "all descriptor matrices must share dtype and width")
            if d.strides[1] != d.itemsize:
                d = np.ascontiguousarray(d)
            keep.append(d)
        ptrs = (C.c_void_p * n)(*[d.ctypes.data if d.shape[0] else None for d in keep])
        nrows = (C.c_int32 * n)(*[d.shape[0] for d in keep])
        steps = (C.c_size_t * n)(*[d.strides[0] if d.shape[0] > 1 else cols * d.itemsize for d in keep])
        return keep, ptrs, nrows, steps, cols, depth

    def match_pairs_from_host(self, descriptors, pairs, norm, k=2, ratio=0.7, cross_check=False, distinct=False,
                              min_match_count=0, engine=ENGINE_AUTO) -> MatchResult:
        """Upload + match + collect in one call (pipelined when the arrays are page-locked, 128 wide)."""
        pairs = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)
        keep, ptrs, nrows, steps, cols, depth = self._bank_args(descriptors)
        o = self._opts(norm, k, ratio, cross_check, distinct, min_match_count, engine)
        res = C.c_void_p()
        self._check(_lib.sfm_match_pairs_from_host(self._ctx, C.c_int(len(keep)), ptrs, nrows, C.c_int(cols), steps,
                                                   C.c_int(depth), pairs.ctypes.data_as(C.c_void_p), C.c_int64(len(pairs)),
                                                   C.byref(o), C.byref(res)))
        return self._wrap_result(res)

    # ---- multi-GPU group, one process per GPU (sfm_dist_*): collective calls, result on participant 0 only
    def dist_init(self, uid: bytes, rank: int, world: int):
        buf = (C.c_uint8 * DIST_ID_BYTES).from_buffer_copy(uid) if uid is not None else None
        self._check(_lib.sfm_dist_init(self._ctx, buf, C.c_int(rank), C.c_int(world)))

    def dist_features_allgather(self, global_index, n_images_total: int):
        """Collective: afterwards this context's feature set holds every image of the scene in scene order."""
        gi = np.ascontiguousarray(global_index, np.int32)
        self._check(_lib.sfm_dist_features_allgather(self._ctx, gi.ctypes.data_as(C.c_void_p), C.c_int(len(gi)), C.c_int(n_images_total)))

    def dist_info(self):
        a, b = C.c_int(), C.c_int()
        _lib.sfm_dist_info(self._ctx, C.byref(a), C.byref(b))
        return a.value, b.value

    def dist_last_phases(self):
        ms = (C.c_double * 8)()
        _lib.sfm_dist_last_phases(self._ctx, ms)
        return list(ms)

    def dist_match_pairs(self, pairs, norm, k=2, ratio=0.7, cross_check=False, distinct=False, min_match_count=0,
                         engine=ENGINE_AUTO, view=False):
        pairs = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)
        o = self._opts(norm, k, ratio, cross_check, distinct, min_match_count, engine)
        res = C.c_void_p()
        self._check(_lib.sfm_dist_match_pairs(self._ctx, pairs.ctypes.data_as(C.c_void_p), C.c_int64(len(pairs)), C.byref(o),
                                              C.byref(res)))
        return (self._view_result(res) if view else self._wrap_result(res)) if res else None

    def dist_match_pairs_from_host(self, descriptors, pairs, norm, k=2, ratio=0.7, cross_check=False, distinct=False,
                                   min_match_count=0, engine=ENGINE_AUTO, view=False):
        pairs = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)
        keep, ptrs, nrows, steps, cols, depth = self._bank_args(descriptors)
        o = self._opts(norm, k, ratio, cross_check, distinct, min_match_count, engine)
        res = C.c_void_p()
        self._check(_lib.sfm_dist_match_pairs_from_host(self._ctx, C.c_int(len(keep)), ptrs, nrows, C.c_int(cols), steps,
                                                        C.c_int(depth), pairs.ctypes.data_as(C.c_void_p), C.c_int64(len(pairs)),
                                                        C.byref(o), C.byref(res)))
        return (self._view_result(res) if view else self._wrap_result(res)) if res else None

    def collect(self) -> MatchResult:
        res = C.c_void_p()
        self._check(_lib.sfm_match_pairs_collect(self._ctx, C.byref(res)))
        return self._wrap_result(res)

    @staticmethod
    def _view_result(res) -> "MatchResultView":
        return MatchResultView(res)

    @staticmethod
    def _wrap_result(res) -> MatchResult:
        try:
            n = _lib.sfm_result_n_pairs(res)
            offsets = np.ctypeslib.as_array(_lib.sfm_result_offsets(res), shape=(n + 1,)).copy()
            total = int(offsets[n])
            dropped = (np.ctypeslib.as_array(_lib.sfm_result_dropped(res), shape=(n,)).copy() if n else
                       np.zeros(0, np.uint8))
            if total:
                buf = (C.c_char * (total * 16)).from_address(_lib.sfm_result_matches(res))
                matches = np.frombuffer(buf, dtype=DMATCH_DTYPE, count=total).copy()
            else:
                matches = np.zeros(0, DMATCH_DTYPE)
        finally:
            _lib.sfm_result_free(res)
        return MatchResult(offsets, matches, dropped)

    def device_view(self):
        """(d_matches_ptr, d_pair_offsets_ptr, d_dropped_ptr, n_pairs, total_matches) of the last enqueue."""
        a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        n, t = C.c_int64(), C.c_int64()
        self._check(_lib.sfm_match_pairs_device_view(self._ctx, C.byref(a), C.byref(b), C.byref(c), C.byref(n), C.byref(t)))
        return int(a.value or 0), int(b.value or 0), int(c.value or 0), n.value, t.value

    def match_pairs(self, pairs, norm, **kw) -> MatchResult:
        self.enqueue(pairs, norm, **kw)
        return self.collect()

    # ---- homography stage (SfM::calculateHomography, SfM.cpp:599-637)
    def upload_keypoints(self, keypoints):
        """keypoints: one float32 [n_rows, 2] array (KeyPoint.pt) per shot of the bank; rows may be strided
        (e.g. a view into packed cv::KeyPoint records, step 28)."""
        n = len(keypoints)
        keep = []
        for k in keypoints:
            k = np.asarray(k)
            if k.dtype != np.float32 or k.ndim != 2 or k.shape[1] != 2:
                raise SfmError(ERR_INVALID, "keypoints must be float32 [n, 2]")
            if k.shape[0] and k.strides[1] != 4:
                k = np.ascontiguousarray(k)
            keep.append(k)
        ptrs = (C.c_void_p * n)(*[k.ctypes.data if k.shape[0] else None for k in keep])
        nrows = (C.c_int32 * n)(*[k.shape[0] for k in keep])
        steps = (C.c_size_t * n)(*[k.strides[0] if k.shape[0] > 1 else 8 for k in keep])
        self._check(_lib.sfm_keypoints_upload(self._ctx, C.c_int(n), ptrs, nrows, steps))

    def homography_inlier_ratios(self, threshold=3.0, max_iters=2000, seed=0, confidence=0.995, refine=True):
        """Per pair of the last match_pairs run: dict(ratio, inliers, ransac_inliers, hypothesis); ratio = -1 where the
        reference attempts no homography (< 4 matches, dropped pairs).  threshold: pixels, scalar or one per pair."""
        a, b, c, n, t = self.device_view()
        thr = np.ascontiguousarray(np.atleast_1d(np.asarray(threshold, np.float64)))
        o = HomographyOpts()
        _lib.sfm_homography_opts_default(C.byref(o))
        o.max_iters, o.refine, o.confidence, o.seed = max_iters, int(bool(refine)), confidence, seed
        ratios = np.zeros(n, np.float64)
        inl, rinl, hyp = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.int32)
        self._check(_lib.sfm_homography_inlier_ratios(self._ctx, thr.ctypes.data_as(C.c_void_p), C.c_int64(len(thr)), C.byref(o),
                                                      ratios.ctypes.data_as(C.c_void_p), inl.ctypes.data_as(C.c_void_p),
                                                      rinl.ctypes.data_as(C.c_void_p), hyp.ctypes.data_as(C.c_void_p)))
        return {"ratio": ratios, "inliers": inl, "ransac_inliers": rinl, "hypothesis": hyp}

    # ---- feature extraction stage (SfM::extractFeatures, SfM.cpp:577-597, cv::SIFT)
    def features_clear(self):
        self._check(_lib.sfm_features_clear(self._ctx))

    def extract_sift(self, gray: np.ndarray, contrast_threshold: float = 0.04, n_octave_layers: int = 3,
                     edge_threshold: float = 10.0, sigma: float = 1.6, max_keypoints: int = 0, n_features: int = 0) -> int:
        """detect() + compute() of cv::SIFT::create(n_features, n_octave_layers, contrast_threshold, edge_threshold, sigma) on one grey
        uint8 image; the result is appended to the context's device-resident feature set.  Returns the keypoint count."""
        gray = np.asarray(gray)
        if gray.dtype == np.uint8 and gray.ndim == 3 and gray.shape[2] in (3, 4):
            gray = gray_from_bgr(gray)                     # a colour photograph as cv::imread returns it (BGR)
        if gray.dtype != np.uint8 or gray.ndim != 2:
            raise SfmError(ERR_INVALID, "extract_sift needs a uint8 grey [rows, cols] or BGR [rows, cols, 3] image")
        if gray.size and gray.strides[1] != 1:
            gray = np.ascontiguousarray(gray)
        o = SiftOpts()
        _lib.sfm_sift_opts_default(C.byref(o))
        o.n_octave_layers, o.max_keypoints, o.n_features = n_octave_layers, max_keypoints, n_features
        o.contrast_threshold, o.edge_threshold, o.sigma = contrast_threshold, edge_threshold, sigma
        n = C.c_int32(0)
        step = gray.strides[0] if gray.shape[0] > 1 else gray.shape[1]
        self._check(_lib.sfm_features_extract_sift(self._ctx, C.c_void_p(gray.ctypes.data if gray.size else None),
                                                   C.c_int(gray.shape[0]), C.c_int(gray.shape[1]), C.c_size_t(step), C.byref(o),
                                                   C.byref(n)))
        return n.value

    def extract_orb(self, gray: np.ndarray, n_features: int = 500, max_keypoints: int = 0, **other) -> int:
        """detect() + compute() of cv::ORB::create(n_features) on one grey uint8 image, appended to the context's feature set
        (32-byte descriptors).  `other`: n_levels / edge_threshold / patch_size / fast_threshold / scale_factor (defaults only)."""
        gray = np.asarray(gray)
        if gray.dtype == np.uint8 and gray.ndim == 3 and gray.shape[2] in (3, 4):
            gray = gray_from_bgr(gray)
        if gray.dtype != np.uint8 or gray.ndim != 2:
            raise SfmError(ERR_INVALID, "extract_orb needs a uint8 grey [rows, cols] or BGR [rows, cols, 3] image")
        if gray.size and gray.strides[1] != 1:
            gray = np.ascontiguousarray(gray)
        o = OrbOpts()
        _lib.sfm_orb_opts_default(C.byref(o))
        o.n_features, o.max_keypoints = n_features, max_keypoints
        for k, v in other.items():
            setattr(o, k, v)
        n = C.c_int32(0)
        step = gray.strides[0] if gray.shape[0] > 1 else gray.shape[1]
        self._check(_lib.sfm_features_extract_orb(self._ctx, C.c_void_p(gray.ctypes.data if gray.size else None),
                                                  C.c_int(gray.shape[0]), C.c_int(gray.shape[1]), C.c_size_t(step), C.byref(o), C.byref(n)))
        return n.value

    def orb_level(self, what: int, level: int) -> np.ndarray:
        """Test aid: 0 level image, 1 blurred, 2 FAST score, 3 candidates (uint8), 4 Harris response (float32)."""
        w, h = C.c_int32(0), C.c_int32(0)
        self._check(_lib.sfm_features_orb_level(self._ctx, C.c_int(what), C.c_int(level), None, C.byref(w), C.byref(h)))
        out = np.zeros((h.value, w.value), np.float32 if what == 4 else np.uint8)
        self._check(_lib.sfm_features_orb_level(self._ctx, C.c_int(what), C.c_int(level), out.ctypes.data_as(C.c_void_p), C.byref(w), C.byref(h)))
        return out

    def features_descriptor_bytes(self) -> int:
        b = C.c_int(0)
        self._check(_lib.sfm_features_descriptor_bytes(self._ctx, C.byref(b)))
        return b.value

    def features_count(self) -> int:
        n = C.c_int(0)
        self._check(_lib.sfm_features_count(self._ctx, C.byref(n)))
        return n.value

    def features_download(self, image: int):
        """(keypoints [n] KEYPOINT_DTYPE, descriptors [n, 128] uint8) of one extracted image."""
        n = C.c_int32(0)
        self._check(_lib.sfm_features_download(self._ctx, C.c_int(image), C.byref(n), None, None))
        kps = np.zeros(n.value, KEYPOINT_DTYPE)
        desc = np.zeros((n.value, self.features_descriptor_bytes() or 128), np.uint8)
        if n.value:
            self._check(_lib.sfm_features_download(self._ctx, C.c_int(image), C.byref(n), kps.ctypes.data_as(C.c_void_p),
                                                   desc.ctypes.data_as(C.c_void_p)))
        return kps, desc

    def bank_from_features(self):
        """The extracted images become the descriptor bank and the keypoint table (device to device)."""
        self._check(_lib.sfm_bank_from_features(self._ctx))

    def features_last_counts(self):
        c = (C.c_int32 * 3)()
        self._check(_lib.sfm_features_last_counts(self._ctx, c))
        return {"extrema": c[0], "keypoints_raw": c[1], "keypoints": c[2]}

    def features_last_profile(self):
        a, b, c = C.c_double(0), C.c_double(0), C.c_double(0)
        self._check(_lib.sfm_features_last_profile(self._ctx, C.byref(a), C.byref(b), C.byref(c)))
        return {"pyramid_ms": a.value, "total_ms": b.value, "pyramid_bytes": c.value}

    def pyramid_level(self, octave: int, level: int) -> np.ndarray:
        w, h = C.c_int32(0), C.c_int32(0)
        self._check(_lib.sfm_features_pyramid_level(self._ctx, C.c_int(octave), C.c_int(level), None, C.byref(w), C.byref(h)))
        out = np.zeros((h.value, w.value), np.float32)
        self._check(_lib.sfm_features_pyramid_level(self._ctx, C.c_int(octave), C.c_int(level), out.ctypes.data_as(C.c_void_p),
                                                    C.byref(w), C.byref(h)))
        return out

    # ---- operator level
    def knn_match(self, query: np.ndarray, train: np.ndarray, norm: int, k: int = 2, engine: int = ENGINE_AUTO):
        """cv::batchDistance-shaped result: (nidx[nq,k] int32, dist[nq,k] float32)."""
        dq, dt = _depth_of(query), _depth_of(train)
        if dq != dt or query.ndim != 2 or train.ndim != 2 or query.shape[1] != train.shape[1]:
            raise SfmError(ERR_INVALID, "type == src2.type() && src1.cols == src2.cols (cv::batchDistance assert)")
        if query.strides[1] != query.itemsize:
            query = np.ascontiguousarray(query)
        if train.strides[1] != train.itemsize:
            train = np.ascontiguousarray(train)
        nq, nt, cols = query.shape[0], train.shape[0], query.shape[1]
        nidx = np.full((nq, k), -1, np.int32)
        dist = np.full((nq, k), np.inf, np.float32)
        qs = query.strides[0] if nq > 1 else cols * query.itemsize
        ts = train.strides[0] if nt > 1 else cols * train.itemsize
        self._check(_lib.sfm_knn_match(self._ctx, C.c_void_p(query.ctypes.data if nq else None), C.c_int(nq),
                                       C.c_size_t(qs), C.c_void_p(train.ctypes.data if nt else None), C.c_int(nt),
                                       C.c_size_t(ts), C.c_int(cols), C.c_int(dq), C.c_int(norm), C.c_int(k),
                                       C.c_int(engine), nidx.ctypes.data_as(C.c_void_p), dist.ctypes.data_as(C.c_void_p)))
        return nidx, dist


class _CtxView(Matcher):
    """A context owned by a MultiGpuMatcher (stats / profiling / homography per participant); never destroyed here."""

    def __init__(self, ctx_ptr, device):
        self._ctx = C.c_void_p(ctx_ptr)
        self.device = device

    def close(self):
        self._ctx = C.c_void_p()


class MultiGpuMatcher:
    """One process, one worker thread per GPU (sfm_mgpu_*): the multi-GPU form of Matcher for a single-process
    pipeline like the reference's (UnorderedFeatureMatchingStrategy.cpp:40 runs pairs as an OpenMP parallel for)."""

    def __init__(self, devices):
        devices = [int(d) for d in devices]
        arr = (C.c_int * len(devices))(*devices)
        self._g = C.c_void_p()
        rc = _lib.sfm_mgpu_create(C.byref(self._g), arr, C.c_int(len(devices)))
        if rc != OK:
            raise SfmError(rc, _lib.sfm_last_error(None).decode())
        self.devices = devices

    def close(self):
        if self._g:
            _lib.sfm_mgpu_destroy(self._g)
            self._g = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != OK:
            raise SfmError(rc, _lib.sfm_mgpu_last_error(self._g).decode())

    def ctx(self, i) -> Matcher:
        return _CtxView(_lib.sfm_mgpu_ctx(self._g, C.c_int(i)), self.devices[i])

    def extract_features(self, images, detector: str = "SIFT", **opts):
        """SfM::extractFeatures over all devices (image i on device i % n), feature sets exchanged, every device adopts the scene
        as its bank.  Returns the keypoint counts; read features back with ctx(0).features_download(i)."""
        imgs = [np.ascontiguousarray(im, np.uint8) for im in images]
        n = len(imgs)
        ptrs = (C.c_void_p * n)(*[im.ctypes.data for im in imgs])
        rows = (C.c_int32 * n)(*[im.shape[0] for im in imgs])
        cols = (C.c_int32 * n)(*[im.shape[1] for im in imgs])
        steps = (C.c_size_t * n)(*[im.strides[0] for im in imgs])
        if detector.upper() == "ORB":
            o = OrbOpts()
            _lib.sfm_orb_opts_default(C.byref(o))
            det = 1
        else:
            o = SiftOpts()
            _lib.sfm_sift_opts_default(C.byref(o))
            det = 0
        for k, v in opts.items():
            setattr(o, k, v)
        counts = (C.c_int32 * n)()
        self._check(_lib.sfm_mgpu_extract_features(self._g, C.c_int(det), C.c_int(n), ptrs, rows, cols, steps, C.byref(o), counts))
        return list(counts)

    def upload_bank(self, descriptors):
        keep, ptrs, nrows, steps, cols, depth = Matcher._bank_args(None, descriptors)
        self._check(_lib.sfm_mgpu_bank_upload(self._g, C.c_int(len(keep)), ptrs, nrows, C.c_int(cols), steps, C.c_int(depth)))

    def match_pairs(self, pairs, norm, k=2, ratio=0.7, cross_check=False, distinct=False, min_match_count=0,
                    engine=ENGINE_AUTO, view=False) -> MatchResult:
        pairs = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)
        o = Matcher._opts(None, norm, k, ratio, cross_check, distinct, min_match_count, engine)
        res = C.c_void_p()
        self._check(_lib.sfm_mgpu_match_pairs(self._g, pairs.ctypes.data_as(C.c_void_p), C.c_int64(len(pairs)), C.byref(o),
                                              C.byref(res)))
        return Matcher._view_result(res) if view else Matcher._wrap_result(res)

    def match_pairs_from_host(self, descriptors, pairs, norm, k=2, ratio=0.7, cross_check=False, distinct=False,
                              min_match_count=0, engine=ENGINE_AUTO, view=False) -> MatchResult:
        pairs = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)
        keep, ptrs, nrows, steps, cols, depth = Matcher._bank_args(None, descriptors)
        o = Matcher._opts(None, norm, k, ratio, cross_check, distinct, min_match_count, engine)
        res = C.c_void_p()
        self._check(_lib.sfm_mgpu_match_pairs_from_host(self._g, C.c_int(len(keep)), ptrs, nrows, C.c_int(cols), steps,
                                                        C.c_int(depth), pairs.ctypes.data_as(C.c_void_p), C.c_int64(len(pairs)),
                                                        C.byref(o), C.byref(res)))
        return Matcher._view_result(res) if view else Matcher._wrap_result(res)
