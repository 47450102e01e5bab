"""Drop-in mirror of final_project/algorithms/matching.py for the AKAZE/Hamming configuration.

`Matcher` stands in for the cv2.BFMatcher objects built by get_akaze_matcher_lr_matcher()
(matching.py:19-24): same `.match(q, t)` / `.knnMatch(q, t, k)` signatures, same return types
(tuples of real cv2.DMatch, imgIdx=0, float distance), same ordering and tie-breaks, same
error behaviour (cv2.error on dtype / width mismatch, empty tuple on empty input — SURVEY.md
section 8b).  The arithmetic runs in libslamfe's sm_100a kernel; host arrays are staged
through pinned memory.  AKAZE detection itself (matching.py:42-43) stays with OpenCV.
"""
from __future__ import annotations

import numpy as np

from . import _cabi, _objects, ops

try:  # cv2 provides the DMatch / KeyPoint container types the reference's callers consume
    import cv2
except Exception:  # pragma: no cover - cv2 is part of the image
    cv2 = None

NORM_HAMMING = 6  # == cv2.NORM_HAMMING


class _Staging:
    """Cached pinned host buffers + device buffers for one stream of calls."""

    def __init__(self):
        self._pinned = {}
        self._dev = {}

    def _buf(self, store, key, nbytes, **kw):
        torch = _cabi.require_cuda()
        cur = store.get(key)
        if cur is None or cur.numel() < nbytes:
            cap = max(nbytes, 1 << 16)
            cap = 1 << (cap - 1).bit_length()
            cur = torch.empty((cap,), dtype=torch.uint8, **kw)
            store[key] = cur
        return cur

    def to_device(self, key, arr: np.ndarray):
        """numpy (C-contiguous) -> device tensor of the same dtype/shape via pinned staging."""
        torch = _cabi.require_cuda()
        arr = np.ascontiguousarray(arr)
        nbytes = arr.nbytes
        pin = self._buf(self._pinned, key, nbytes, pin_memory=True)
        dev = self._buf(self._dev, key, nbytes, device="cuda")
        if nbytes:
            pin[:nbytes].numpy()[:] = arr.reshape(-1).view(np.uint8)
            dev[:nbytes].copy_(pin[:nbytes], non_blocking=True)
        tdtype = {np.dtype(np.uint8): torch.uint8, np.dtype(np.float32): torch.float32,
                  np.dtype(np.float64): torch.float64, np.dtype(np.int32): torch.int32}[arr.dtype]
        return dev[:nbytes].view(tdtype).view(arr.shape)

    def to_host(self, key, t) -> np.ndarray:
        """device tensor -> numpy copy via pinned staging (synchronises the current stream)."""
        torch = _cabi.require_cuda()
        t = t.contiguous()
        nbytes = t.numel() * t.element_size()
        pin = self._buf(self._pinned, key, nbytes, pin_memory=True)
        if nbytes:
            pin[:nbytes].copy_(t.view(-1).view(torch.uint8), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        npdtype = {torch.uint8: np.uint8, torch.int32: np.int32, torch.float32: np.float32,
                   torch.float64: np.float64}[t.dtype]
        return pin[:nbytes].numpy().view(npdtype).reshape(tuple(t.shape)).copy()


class _ThreadLocalStaging:
    """One _Staging per Python thread: the module-level entry points (extract_inliers_outliers,
    triangulate_links, transformation_agreement, ...) may be called from several threads, each on its
    own CUDA stream, and must not share pinned buffers."""

    def __init__(self):
        import threading
        self._tls = threading.local()

    def _get(self):
        st = getattr(self._tls, "st", None)
        if st is None:
            st = self._tls.st = _Staging()
        return st

    def to_device(self, key, arr):
        return self._get().to_device(key, arr)

    def to_host(self, key, t):
        return self._get().to_host(key, t)



def _cv2_error(msg):
    if cv2 is not None:
        return cv2.error(msg)
    return ValueError(msg)


def _check_descriptors(q, t):
    """Return (q, t) as 2-D uint8 arrays, or None if the call is the empty case."""
    if q is None or t is None:
        return None
    q, t = np.asarray(q), np.asarray(t)
    if q.size == 0 or t.size == 0:
        return None
    if q.dtype != np.uint8 or t.dtype != np.uint8:
        raise _cv2_error("slamfe.Matcher: NORM_HAMMING needs uint8 descriptors (batch_distance.cpp type assert)")
    if q.ndim != 2 or t.ndim != 2 or q.shape[1] != t.shape[1]:
        raise _cv2_error("slamfe.Matcher: query and train descriptors must have the same width "
                         "(batch_distance.cpp size assert)")
    if q.shape[1] > _cabi.MAX_DESC_BYTES:
        raise _cv2_error("slamfe.Matcher: descriptors wider than 64 bytes are not supported")
    return q, t


def _dmatches(qidx, tidx, dist):
    """Arrays -> tuple of cv2.DMatch(queryIdx, trainIdx, imgIdx=0, distance) (the 4-argument form:
    matcher output has imgIdx 0).  .tolist() first: building the objects from
    Python ints/floats is ~3x faster than from numpy scalars."""
    fast = _objects.dmatch_tuple(qidx, tidx, dist)   # C helper: 0.15 us per object instead of 1 us
    if fast is not None:
        return fast
    n = len(qidx)
    return tuple(map(cv2.DMatch, np.asarray(qidx).tolist(), np.asarray(tidx).tolist(), [0] * n,
                     np.asarray(dist, dtype=np.float64).tolist()))


class Matcher:
    """cv2.BFMatcher(normType=cv2.NORM_HAMMING, crossCheck=...) on the GPU."""

    def __init__(self, normType=NORM_HAMMING, crossCheck=False):
        if normType != NORM_HAMMING:
            raise ValueError("slamfe.Matcher implements NORM_HAMMING only (SIFT/L2 is out of scope)")
        self.normType = normType
        self.crossCheck = bool(crossCheck)
        self._st = _ThreadLocalStaging()  # the module-global MATCHER objects may be shared by threads

    # -- array-level API (no per-match Python objects) ------------------------------------
    def match_arrays(self, queryDescriptors, trainDescriptors):
        """(queryIdx, trainIdx, distance) int32 arrays with .match() semantics."""
        chk = _check_descriptors(queryDescriptors, trainDescriptors)
        if chk is None:
            z = np.zeros(0, dtype=np.int32)
            return z, z.copy(), z.copy()
        q, t = chk
        qd = self._st.to_device("q", q)
        td = self._st.to_device("t", t)
        row_keys, col_keys = ops.hamming_top2(qd, td, want_cols=self.crossCheck, best_only=True)
        if self.crossCheck:
            mt, md = ops.cross_check(row_keys, col_keys)
            both = self._st.to_host("o", mt._base)   # (2, nq): match_t and match_dist share one buffer
            keep = both[0] >= 0
            return np.nonzero(keep)[0].astype(np.int32), both[0][keep], both[1][keep]
        keys = self._st.to_host("o", row_keys)
        idx, dist = ops.keys_to_numpy(keys[:, 0])
        return np.arange(q.shape[0], dtype=np.int32), idx, dist

    def knn_arrays(self, queryDescriptors, trainDescriptors):
        """(idx2, dist2) int32 (Nq, 2) arrays with knnMatch(k=2) semantics (-1 = missing)."""
        chk = _check_descriptors(queryDescriptors, trainDescriptors)
        if chk is None:
            z = np.zeros((0, 2), dtype=np.int32)
            return z, z.copy()
        q, t = chk
        row_keys, _ = ops.hamming_top2(self._st.to_device("q", q), self._st.to_device("t", t))
        keys = self._st.to_host("o", row_keys)
        return ops.keys_to_numpy(keys)

    # -- cv2.BFMatcher API ------------------------------------------------------------------
    def match(self, queryDescriptors, trainDescriptors, mask=None):
        if mask is not None:
            raise NotImplementedError("match masks are not used by the reference")
        qi, ti, d = self.match_arrays(queryDescriptors, trainDescriptors)
        return _dmatches(qi, ti, d)

    def knnMatch(self, queryDescriptors, trainDescriptors, k=2, mask=None, compactResult=False):
        if mask is not None:
            raise NotImplementedError("match masks are not used by the reference")
        if k not in (1, 2):
            raise NotImplementedError("slamfe.Matcher.knnMatch supports k in {1, 2}")
        if self.crossCheck:
            if k != 1:
                raise _cv2_error("slamfe.Matcher: crossCheck requires k == 1")
            return tuple((m,) for m in self.match(queryDescriptors, trainDescriptors))
        idx2, dist2 = self.knn_arrays(queryDescriptors, trainDescriptors)
        nq = idx2.shape[0]
        if nq == 0:
            return ()
        qi = np.arange(nq)
        first = _dmatches(qi, idx2[:, 0], dist2[:, 0])
        if k == 1:
            return tuple((m,) for m in first)
        has2 = idx2[:, 1] >= 0  # a single train row yields length-1 inner tuples (cv2 behaviour)
        seconds = _dmatches(qi[has2], idx2[has2, 1], dist2[has2, 1])
        fast = _objects.knn_tuples(first, seconds, has2)
        if fast is not None:
            return fast
        second = iter(seconds)
        return tuple((m, next(second)) if h else (m,) for m, h in zip(first, has2.tolist()))


def ratio_test_mask(dist2, ratio_num=5, ratio_den=3):
    """VAN_ex/code/ex1.py:118-122 on knn distances: d1 < 0.6 * d2 as the exact integer test
    5 * d1 < 3 * d2 (identical for all 0 <= d <= 512)."""
    d = np.asarray(dist2)
    return (d[:, 1] >= 0) & (ratio_num * d[:, 0].astype(np.int64) < ratio_den * d[:, 1].astype(np.int64))


def get_akaze_matcher_lr_matcher():
    """matching.py:19-24 with the two BFMatchers replaced by GPU matchers."""
    feature = cv2.AKAZE_create(threshold=0.0008, nOctaves=4, nOctaveLayers=4)
    matcher = Matcher(normType=NORM_HAMMING, crossCheck=False)
    matcher_left_right = Matcher(normType=NORM_HAMMING, crossCheck=True)
    return feature, matcher_left_right, matcher


_filter_staging = _ThreadLocalStaging()


def extract_inliers_outliers(kp_left, kp_right, matches):
    """matching.py:48-69: indices (into `matches`) passing / failing the rectified-stereo test
    abs(yl - yr) < 2 and xl > xr + 2.  Accepts cv2.KeyPoint sequences or (N, 2) float arrays."""
    n = len(matches)
    if n == 0:
        return np.array([]), np.array([])
    pl = _keypoint_array(kp_left)
    pr = _keypoint_array(kp_right)
    fast = _objects.dmatch_indices(matches)
    if fast is not None:
        mq, mt = fast
    else:
        mq = np.fromiter((m.queryIdx for m in matches), dtype=np.int32, count=n)
        mt = np.fromiter((m.trainIdx for m in matches), dtype=np.int32, count=n)
    mask = stereo_filter_mask(pl, pr, mq, mt)
    return np.array(np.nonzero(mask)[0].tolist()), np.array(np.nonzero(~mask)[0].tolist())


def stereo_filter_mask(pts_left, pts_right, match_q, match_t):
    """Array form of the row filter: bool mask over matches."""
    st = _filter_staging
    mask = ops.stereo_filter(st.to_device("pl", pts_left.astype(np.float32, copy=False)),
                             st.to_device("pr", pts_right.astype(np.float32, copy=False)),
                             st.to_device("mq", match_q.astype(np.int32, copy=False)),
                             st.to_device("mt", match_t.astype(np.int32, copy=False)))
    return st.to_host("mask", mask).astype(bool)


def _keypoint_array(kps):
    if isinstance(kps, np.ndarray):
        return np.ascontiguousarray(kps, dtype=np.float32).reshape(-1, 2)
    if len(kps) == 0:
        return np.zeros((0, 2), dtype=np.float32)
    return np.ascontiguousarray(cv2.KeyPoint_convert(kps), dtype=np.float32).reshape(-1, 2)


# Module globals the reference's callers import by name (matching.py:72-73, AKAZE line).
FEATURE = None
MATCHER_LEFT_RIGHT = None
MATCHER = None


def init_globals():
    """Instantiate FEATURE / MATCHER_LEFT_RIGHT / MATCHER (kept lazy so that importing the
    package needs neither a GPU nor cv2)."""
    global FEATURE, MATCHER_LEFT_RIGHT, MATCHER
    FEATURE, MATCHER_LEFT_RIGHT, MATCHER = get_akaze_matcher_lr_matcher()
    return FEATURE, MATCHER_LEFT_RIGHT, MATCHER


def extract_kps_descs_matches(img_0, img1):
    """matching.py:38-45: AKAZE detect+describe on the CPU (out of scope), L<->R crossCheck match
    on the GPU."""
    if FEATURE is None:
        init_globals()
    kp0, desc0 = FEATURE.detectAndCompute(img_0, None)
    kp1, desc1 = FEATURE.detectAndCompute(img1, None)
    matches = MATCHER_LEFT_RIGHT.match(desc0, desc1)
    return kp0, kp1, desc0, desc1, matches
