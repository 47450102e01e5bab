"""ORACLE — test infrastructure only.  NOT part of the product path.

CPU restatement (numpy + the C library built from oracle/hamming_oracle.c) of the
reference's front-end hot path.  Every function cites the reference file:line it follows.
Array-in / array-out: the reference's per-object Python containers (cv2.DMatch, cv2.KeyPoint,
Link) are flattened to ndarrays; `oracle/refshim.py` + `oracle/make_golden.py` check these
restatements against the unmodified reference and against cv2 4.13.0 in the build container,
and freeze the outputs under tests/golden/ (parity is pinned by those generated vectors; the
reference itself ships no tests — SURVEY.md section 4).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile oracle/hamming_oracle.c -> oracle/liboracle.so (gcc, OpenMP)."""
    src = os.path.join(_HERE, "hamming_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-msse4.2", "-mpopcnt", "-fopenmp", "-shared", "-fPIC",
                               src, "-o", _LIB_PATH])
    return _LIB_PATH


_P3P_LIB_PATH = os.path.join(_HERE, "libp3p_host.so")
_p3p_lib = None


def build_p3p_shim(force: bool = False) -> str:
    """Compile oracle/p3p_host_shim.cpp (the product's csrc/p3p.cuh solver header, built for the
    HOST) -> oracle/libp3p_host.so, so that CPU tests can pin the solver's arithmetic."""
    src = os.path.join(_HERE, "p3p_host_shim.cpp")
    hdr = os.path.join(os.path.dirname(_HERE), "67604-slam---video-navigation_b200", "csrc", "p3p.cuh")
    newest = max(os.path.getmtime(src), os.path.getmtime(hdr))
    if force or not os.path.exists(_P3P_LIB_PATH) or os.path.getmtime(_P3P_LIB_PATH) < newest:
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-x", "c++", src, "-o",
                               _P3P_LIB_PATH])
    return _P3P_LIB_PATH


class P3PHost:
    """ctypes view of the host build of csrc/p3p.cuh (test infrastructure)."""

    def __init__(self):
        global _p3p_lib
        if _p3p_lib is None:
            build_p3p_shim()
            _p3p_lib = ctypes.CDLL(_P3P_LIB_PATH)
        self.lib = _p3p_lib
        self._dp = ctypes.POINTER(ctypes.c_double)

    def solve(self, pts4, pix4, K):
        """(T (3,4), ok) for one 4-point sample."""
        T = np.zeros(12)
        P = np.ascontiguousarray(pts4, dtype=np.float64)
        uv = np.ascontiguousarray(pix4, dtype=np.float64)
        Kc = np.ascontiguousarray(K, dtype=np.float64)
        Ki = np.ascontiguousarray(np.linalg.inv(Kc))
        ok = self.lib.p3p_host_solve(P.ctypes.data_as(self._dp), uv.ctypes.data_as(self._dp),
                                     Kc.ctypes.data_as(self._dp), Ki.ctypes.data_as(self._dp),
                                     T.ctypes.data_as(self._dp))
        return T.reshape(3, 4), bool(ok)

    def quartic(self, A):
        A = np.ascontiguousarray(A, dtype=np.float64)
        r = np.zeros(4)
        n = self.lib.p3p_host_quartic(A.ctypes.data_as(self._dp), r.ctypes.data_as(self._dp))
        return r[:n]

    def sample4(self, seed, frame, hyp, n):
        idx = (ctypes.c_int * 4)()
        self.lib.p3p_host_sample4(ctypes.c_uint64(seed), ctypes.c_uint32(frame), ctypes.c_uint32(hyp), int(n), idx)
        return np.array(list(idx), dtype=np.int32)


_RANSAC_LIB_PATH = os.path.join(_HERE, "libransac_host.so")
_ransac_lib = None


def build_ransac_shim(force: bool = False) -> str:
    """Compile oracle/ransac_host_shim.cpp (the product's csrc/ransac_core.cuh, built for the HOST
    with -ffp-contract=off) -> oracle/libransac_host.so."""
    src = os.path.join(_HERE, "ransac_host_shim.cpp")
    csrc = os.path.join(os.path.dirname(_HERE), "67604-slam---video-navigation_b200", "csrc")
    newest = max(os.path.getmtime(src), os.path.getmtime(os.path.join(csrc, "ransac_core.cuh")),
                 os.path.getmtime(os.path.join(csrc, "hd.cuh")))
    if force or not os.path.exists(_RANSAC_LIB_PATH) or os.path.getmtime(_RANSAC_LIB_PATH) < newest:
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-x", "c++", src, "-o",
                               _RANSAC_LIB_PATH])
    return _RANSAC_LIB_PATH


class ScorerHost:
    """ctypes view of the host build of csrc/ransac_core.cuh (test infrastructure)."""

    def __init__(self):
        global _ransac_lib
        if _ransac_lib is None:
            build_ransac_shim()
            _ransac_lib = ctypes.CDLL(_RANSAC_LIB_PATH)
        self.lib = _ransac_lib

    def score(self, T, pts, l_pix, r_pix, K, M1, M2):
        """Returns bool arrays (agrees, exact, fast, certified) for one hypothesis, plus `shared`: uint8 per
        point, 1 = the rectified-rig shortcut reproduces the full right-camera rows bit for bit, 0 = it
        does not (a bug), 2 = the hypothesis does not qualify for the shortcut."""
        dp, up = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_ubyte)
        c = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        arrs = [c(K), c(M1), c(M2), c(T), c(pts), c(l_pix), c(r_pix)]
        n = arrs[4].shape[0]
        outs = [np.zeros(n, np.uint8) for _ in range(5)]
        self.lib.ransac_host_score(*[a.ctypes.data_as(dp) for a in arrs], ctypes.c_long(n),
                                   *[o.ctypes.data_as(up) for o in outs])
        return tuple(o.astype(bool) for o in outs[:4]) + (outs[4],)

    def prune(self, T, pts, l_pix, r_pix, K, M1, M2):
        """The kernel's fp32 pre-filter for one hypothesis: bool arrays (far_v, far_u, rows, exact) — dropped by
        the v / u test, agrees_rows() (what the kernel counts for kept pairs), the reference's verdict."""
        dp, up = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_ubyte)
        c = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        arrs = [c(K), c(M1), c(M2), c(T), c(pts), c(l_pix), c(r_pix)]
        n = arrs[4].shape[0]
        outs = [np.zeros(n, np.uint8) for _ in range(4)]
        self.lib.ransac_host_prune(*[a.ctypes.data_as(dp) for a in arrs], ctypes.c_long(n),
                                   *[o.ctypes.data_as(up) for o in outs])
        return tuple(o.astype(bool) for o in outs)

    def matrices(self, T, K, M1, M2):
        dp = ctypes.POINTER(ctypes.c_double)
        c = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        out = np.zeros(24)
        self.lib.ransac_host_matrices(c(K).ctypes.data_as(dp), c(M1).ctypes.data_as(dp), c(M2).ctypes.data_as(dp),
                                      c(T).ctypes.data_as(dp), out.ctypes.data_as(dp))
        return out[:12].reshape(3, 4), out[12:].reshape(3, 4)


_REFIT_LIB_PATH = os.path.join(_HERE, "librefit_host.so")
_refit_lib = None


def build_refit_shim(force: bool = False) -> str:
    """Compile oracle/refit_host_shim.cpp (the product's csrc/refit_core.cuh, built for the HOST) ->
    oracle/librefit_host.so."""
    src = os.path.join(_HERE, "refit_host_shim.cpp")
    hdr = os.path.join(os.path.dirname(_HERE), "67604-slam---video-navigation_b200", "csrc", "refit_core.cuh")
    newest = max(os.path.getmtime(src), os.path.getmtime(hdr))
    if force or not os.path.exists(_REFIT_LIB_PATH) or os.path.getmtime(_REFIT_LIB_PATH) < newest:
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-x", "c++", src, "-o",
                               _REFIT_LIB_PATH])
    return _REFIT_LIB_PATH


def refit_host_build(T_seed, K, pts, pix, mask=None, max_iter=20, tol=1e-12):
    """Host build of the product's refit (same control flow as pnp_refit_kernel).
    Returns (T (3,4), status, rms)."""
    global _refit_lib
    if _refit_lib is None:
        build_refit_shim()
        _refit_lib = ctypes.CDLL(_REFIT_LIB_PATH)
    dp = ctypes.POINTER(ctypes.c_double)
    c = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    T0, Kc, P, uv = c(T_seed), c(K), c(pts), c(pix)
    m = np.ones(len(P), np.uint8) if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
    out, rms = np.zeros(12), ctypes.c_double(0.0)
    st = _refit_lib.refit_host(T0.ctypes.data_as(dp), Kc.ctypes.data_as(dp), P.ctypes.data_as(dp),
                               uv.ctypes.data_as(dp), m.ctypes.data_as(ctypes.POINTER(ctypes.c_ubyte)),
                               ctypes.c_long(len(P)), int(max_iter), ctypes.c_double(tol), out.ctypes.data_as(dp),
                               ctypes.byref(rms))
    return out.reshape(3, 4), int(st), float(rms.value)


def pnp_refit(T_seed, K, pts, pix, iters=30):
    """Independent numpy restatement of the refit's objective (ransac.py:185-193: the pose that fits
    the left-image pixels of the consensus set): Gauss-Newton on the left reprojection error with a
    numerical-free analytic Jacobian in the axis-angle + translation chart at the current pose.
    Returns (T (3,4), rms)."""
    T = np.array(T_seed, dtype=np.float64).reshape(3, 4)
    K = np.asarray(K, dtype=np.float64)
    X = np.asarray(pts, dtype=np.float64)
    uv = np.asarray(pix, dtype=np.float64)

    def residual(T):
        Y = X @ T[:, :3].T + T[:, 3]
        q = Y @ K.T
        return (q[:, :2] / q[:, 2:3] - uv), Y, q

    for _ in range(iters):
        r, Y, q = residual(T)
        W = Y - T[:, 3]
        J = np.zeros((len(X), 2, 6))
        for row in range(2):
            a = (K[row][None, :] - (q[:, row] / q[:, 2])[:, None] * K[2][None, :]) / q[:, 2:3]
            J[:, row, :3] = np.cross(W, a)          # a . (w x W) = w . (W x a)
            J[:, row, 3:] = a
        Jf, rf = J.reshape(-1, 6), r.reshape(-1)
        d = np.linalg.lstsq(Jf, -rf, rcond=None)[0]
        T = np.hstack([rodriguez_to_mat(d[:3].reshape(3, 1), np.zeros((3, 1)))[:, :3] @ T[:, :3],
                       (T[:, 3] + d[3:])[:, None]])
        if np.linalg.norm(d) < 1e-13:
            break
    r, _, _ = residual(T)
    return T, float(np.sqrt((r ** 2).sum() / len(X)))


_TRI_LIB_PATH = os.path.join(_HERE, "libtriangulate_host.so")
_tri_lib = None


def build_triangulate_shim(force: bool = False) -> str:
    """Compile oracle/triangulate_host_shim.cpp (csrc/triangulate_core.cuh for the HOST)."""
    src = os.path.join(_HERE, "triangulate_host_shim.cpp")
    csrc = os.path.join(os.path.dirname(_HERE), "67604-slam---video-navigation_b200", "csrc")
    newest = max(os.path.getmtime(src), os.path.getmtime(os.path.join(csrc, "triangulate_core.cuh")))
    if force or not os.path.exists(_TRI_LIB_PATH) or os.path.getmtime(_TRI_LIB_PATH) < newest:
        subprocess.check_call(["g++", "-O2", "-ffp-contract=fast", "-shared", "-fPIC", "-x", "c++", src, "-o",
                               _TRI_LIB_PATH])
    return _TRI_LIB_PATH


def triangulate_links_host_build(links, P, Q):
    """(M, 3) links -> (M, 3) points through the HOST build of the product's triangulation core."""
    global _tri_lib
    if _tri_lib is None:
        build_triangulate_shim()
        _tri_lib = ctypes.CDLL(_TRI_LIB_PATH)
    dp = ctypes.POINTER(ctypes.c_double)
    c = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    links, Pc, Qc = c(links).reshape(-1, 3), c(P), c(Q)
    out = np.zeros_like(links)
    _tri_lib.triangulate_host_links(links.ctypes.data_as(dp), ctypes.c_long(len(links)), Pc.ctypes.data_as(dp),
                                    Qc.ctypes.data_as(dp), out.ctypes.data_as(dp))
    return out


_HAM_LIB_PATH = os.path.join(_HERE, "libhamming_host.so")
_ham_lib = None


def build_hamming_shim(force: bool = False) -> str:
    """Compile oracle/hamming_host_shim.cpp (csrc/hamming_core.cuh for the HOST)."""
    src = os.path.join(_HERE, "hamming_host_shim.cpp")
    root = os.path.dirname(_HERE)
    csrc = os.path.join(root, "67604-slam---video-navigation_b200", "csrc")
    newest = max(os.path.getmtime(src), os.path.getmtime(os.path.join(csrc, "hamming_core.cuh")))
    if force or not os.path.exists(_HAM_LIB_PATH) or os.path.getmtime(_HAM_LIB_PATH) < newest:
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", "-I", os.path.join(root, "include"), src,
                               "-o", _HAM_LIB_PATH])
    return _HAM_LIB_PATH


def hamming_keys_host_build(q, t, desc_bytes=None, cs=9):
    """Row-wise distances of q[i], t[i] through the HOST build of the matcher's carry-save arithmetic.
    Returns the keys (distance << 22) exactly as the kernel forms them."""
    global _ham_lib
    if _ham_lib is None:
        build_hamming_shim()
        _ham_lib = ctypes.CDLL(_HAM_LIB_PATH)
    q, qp = _u8(q)
    t, tp = _u8(t)
    desc_bytes = q.shape[1] if desc_bytes is None else desc_bytes
    keys = np.zeros(q.shape[0], dtype=np.uint32)
    _ham_lib.hamming_host_keys(qp, tp, ctypes.c_long(q.shape[0]), int(q.shape[1]), int(desc_bytes), int(cs),
                               keys.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)))
    return keys


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        u8p = ctypes.POINTER(ctypes.c_uint8)
        i32p = ctypes.POINTER(ctypes.c_int32)
        for name in ("oracle_hamming_top2", "oracle_hamming_colmin", "oracle_hamming_matrix"):
            fn = getattr(_lib, name)
            fn.restype = None
            fn.argtypes = [u8p, ctypes.c_int, ctypes.c_int, u8p, ctypes.c_int, ctypes.c_int,
                           ctypes.c_int] + ([i32p, i32p] if name != "oracle_hamming_matrix" else [i32p])
    return _lib


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))


def _i32(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


# --------------------------------------------------------------------------------------
# Hamming matcher — cv2.BFMatcher(NORM_HAMMING) semantics
# --------------------------------------------------------------------------------------
def hamming_matrix(q, t):
    """Exact (Nq, Nt) Hamming distance matrix (small cases)."""
    q, qp = _u8(q)
    t, tp = _u8(t)
    D = np.empty((q.shape[0], t.shape[0]), dtype=np.int32)
    _load().oracle_hamming_matrix(qp, q.shape[0], q.shape[1], tp, t.shape[0], t.shape[1],
                                  q.shape[1], _i32(D))
    return D


def knn2(q, t):
    """knnMatch(q, t, k=2) (VAN_ex/code/ex1.py:189-190): per query the two nearest train rows in
    ascending (distance, trainIdx) order.  Returns idx2, dist2 of shape (Nq, 2), -1 = missing."""
    q, qp = _u8(q)
    t, tp = _u8(t)
    idx2 = np.empty((q.shape[0], 2), dtype=np.int32)
    dist2 = np.empty((q.shape[0], 2), dtype=np.int32)
    _load().oracle_hamming_top2(qp, q.shape[0], q.shape[1], tp, t.shape[0], t.shape[1],
                                q.shape[1], _i32(idx2), _i32(dist2))
    return idx2, dist2


def match(q, t):
    """MATCHER.match(q, t), crossCheck=False (database.py:54-55, loop_closure.py:422):
    one entry per query row in query order, first-min train index.  Returns (trainIdx, distance)."""
    idx2, dist2 = knn2(q, t)
    return idx2[:, 0].copy(), dist2[:, 0].copy()


def colmin(q, t):
    """First-min of each column = MATCHER.match(t, q) seen from the (q, t) distance matrix."""
    q, qp = _u8(q)
    t, tp = _u8(t)
    idx = np.empty(t.shape[0], dtype=np.int32)
    dist = np.empty(t.shape[0], dtype=np.int32)
    _load().oracle_hamming_colmin(qp, q.shape[0], q.shape[1], tp, t.shape[0], t.shape[1],
                                  q.shape[1], _i32(idx), _i32(dist))
    return idx, dist


def match_crosscheck(l, r):
    """MATCHER_LEFT_RIGHT.match(l, r), crossCheck=True (matching.py:22,44): keep (i, j) iff j is
    the first-min of row i and i is the first-min of column j; sorted by queryIdx.
    Returns (queryIdx, trainIdx, distance)."""
    ti, td = match(l, r)
    ci, _ = colmin(l, r)
    qi = np.arange(l.shape[0], dtype=np.int32)
    keep = (ti >= 0) & (ci[np.maximum(ti, 0)] == qi)
    return qi[keep], ti[keep], td[keep]


def ratio_test(dist2, ratio=0.6):
    """ex1.py:118-122 with GOOD_RATIO=0.6 (ex1.py:9): keep iff m.distance < ratio * n.distance,
    evaluated in float64 exactly as the reference does."""
    d1 = dist2[:, 0].astype(np.float64)
    d2 = dist2[:, 1].astype(np.float64)
    return (dist2[:, 1] >= 0) & (d1 < ratio * d2)


def mutual_forward_backward(prev, cur):
    """database.py:54-77: forward match prev->cur, backward match cur->prev, keep forward match j
    iff backward[trainIdx].trainIdx == queryIdx.  Returns (fwd_idx, fwd_dist, good_idx)."""
    fi, fd = match(prev, cur)
    bi, _ = match(cur, prev)
    good = np.nonzero(bi[fi] == np.arange(prev.shape[0]))[0]
    return fi, fd, good


# --------------------------------------------------------------------------------------
# Rectified-stereo row filter + links
# --------------------------------------------------------------------------------------
def extract_inliers_outliers(pt_left, pt_right, match_q, match_t):
    """matching.py:48-69.  pt_* are (N, 2) arrays of KeyPoint.pt (float32 values); inlier iff
    abs(yl - yr) < 2 and xl > xr + 2 (matching.py:62-63), compared in float64 like Python
    floats.  Returns (inlier_idx, outlier_idx) into the match list."""
    pl = np.asarray(pt_left, dtype=np.float64)[np.asarray(match_q)]
    pr = np.asarray(pt_right, dtype=np.float64)[np.asarray(match_t)]
    good = (np.abs(pl[:, 1] - pr[:, 1]) < 2) & (pl[:, 0] > pr[:, 0] + 2)
    return np.nonzero(good)[0], np.nonzero(~good)[0]


def create_links(pt_left, pt_right, match_q, match_t):
    """tracking_database.py:224-246 with all matches inliers: Link(xl, xr, (yl + yr) / 2)
    (tracking_database.py:243) in match order; is_valid mask over left keypoints.
    Returns (is_valid (Nl,) bool, links (M, 3) float64 = [x_left, x_right, y])."""
    pl = np.asarray(pt_left, dtype=np.float64)[np.asarray(match_q)]
    pr = np.asarray(pt_right, dtype=np.float64)[np.asarray(match_t)]
    links = np.stack([pl[:, 0], pr[:, 0], (pl[:, 1] + pr[:, 1]) / 2], axis=1)
    is_valid = np.zeros(len(pt_left), dtype=bool)
    is_valid[np.asarray(match_q)] = True
    return is_valid, links


# --------------------------------------------------------------------------------------
# Cameras / small math
# --------------------------------------------------------------------------------------
KITTI00_P0 = np.array([[7.188560000000e+02, 0.0, 6.071928000000e+02, 0.0],
                       [0.0, 7.188560000000e+02, 1.852157000000e+02, 0.0],
                       [0.0, 0.0, 1.0, 0.0]])
KITTI00_P1 = np.array([[7.188560000000e+02, 0.0, 6.071928000000e+02, -3.861448000000e+02],
                       [0.0, 7.188560000000e+02, 1.852157000000e+02, 0.0],
                       [0.0, 0.0, 1.0, 0.0]])


def read_cameras(p0=KITTI00_P0, p1=KITTI00_P1):
    """utils.py:36-51 / Inputs.py:22-37 minus the file read: k = P0[:, :3];
    m1 = inv(k) @ P0; m2 = inv(k) @ P1."""
    k = p0[:, :3]
    m1 = np.linalg.inv(k) @ p0
    m2 = np.linalg.inv(k) @ p1
    return k, m1, m2


def rodriguez_to_mat(rvec, tvec):
    """utils.py:16-18: hstack(cv2.Rodrigues(rvec)[0], tvec)."""
    import cv2
    rot, _ = cv2.Rodrigues(rvec)
    return np.hstack((rot, tvec))


# --------------------------------------------------------------------------------------
# Triangulation
# --------------------------------------------------------------------------------------
def linear_least_squares_triangulation(P, Q, kp_left, kp_right):
    """triangulation.py:5-24: 4x4 DLT, null vector from np.linalg.svd, w==0 guard."""
    A = np.zeros((4, 4))
    p_x, p_y = kp_left
    q_x, q_y = kp_right
    A[0] = P[2] * p_x - P[0]
    A[1] = P[2] * p_y - P[1]
    A[2] = Q[2] * q_x - Q[0]
    A[3] = Q[2] * q_y - Q[1]
    _, _, V = np.linalg.svd(A)
    if V[-1, 3] == 0:
        return V[-1, :3] / (V[-1, 3] + 1e-20)
    return V[-1, :3] / V[-1, 3]


def triangulate_links(links, p, q):
    """triangulation.py:41-50 (and :27-38): loop of the DLT over links (x_left, y), (x_right, y).
    `links` is (M, 3) float64 [x_left, x_right, y]."""
    links = np.asarray(links, dtype=np.float64)
    x = np.zeros((len(links), 3))
    for i in range(len(links)):
        xl, xr, y = links[i]
        x[i] = linear_least_squares_triangulation(p, q, (xl, y), (xr, y))
    return x


def triangulate_points(P, Q, pxy, qxy):
    """General form of triangulation.py:5-24 over arrays of left/right pixels (analysis.py:400,
    VAN_ex/code/ex2.py:209 call it with distinct y)."""
    out = np.zeros((len(pxy), 3))
    for i in range(len(pxy)):
        out[i] = linear_least_squares_triangulation(P, Q, pxy[i], qxy[i])
    return out


# --------------------------------------------------------------------------------------
# RANSAC-PnP scoring
# --------------------------------------------------------------------------------------
def transformation_agreement(T, pts, l_pix, r_pix, K, M1, M2):
    """ransac.py:28-56, same association order ((K @ T @ M_h) @ X), true division, strict < 2 on
    both cameras.  Quirks kept: right camera is K @ T @ [M2; 0001] (M2 applied before T,
    ransac.py:39) and there is no positive-depth test."""
    points_4d = np.hstack((pts, np.ones((pts.shape[0], 1)))).T
    l1 = ((K @ T @ np.vstack((M1, np.array([0, 0, 0, 1])))) @ points_4d)[:3, :].T
    r1 = ((K @ T @ np.vstack((M2, np.array([0, 0, 0, 1])))) @ points_4d)[:3, :].T
    with np.errstate(divide="ignore", invalid="ignore"):
        tl = l1 / l1[:, 2][:, np.newaxis]
        tr = r1 / r1[:, 2][:, np.newaxis]
        agree_l = np.logical_and(np.abs(tl[:, 1] - l_pix[:, 1]) < 2, np.abs(tl[:, 0] - l_pix[:, 0]) < 2)
        agree_r = np.logical_and(np.abs(tr[:, 1] - r_pix[:, 1]) < 2, np.abs(tr[:, 0] - r_pix[:, 0]) < 2)
    return np.logical_and(agree_l, agree_r)


def calc_ransac_iteration(inliers_percent, suc_prob=0.9999999999):
    """ransac.py:59-67 (SUCCESS_PROBABILITY ransac.py:9)."""
    outliers_prob = 1 - (inliers_percent / 100) + 0.0000000001
    return int(np.log(1 - suc_prob) / np.log(1 - np.power(1 - outliers_prob, 4))) + 1


def score_hypotheses(Ts, pts, l_pix, r_pix, K, M1, M2):
    """The scoring half of the RANSAC loops (ransac.py:106-112, :174-182) over GIVEN hypotheses:
    per-hypothesis inlier counts, the first hypothesis with the strictly-largest count
    (ransac.py:110 keeps only strictly better -> lowest index on ties) and its inlier mask.
    best = -1 when no hypothesis scores > 0 inliers."""
    counts = np.zeros(len(Ts), dtype=np.int64)
    best, best_cnt, best_mask = -1, 0, np.zeros(len(pts), dtype=bool)
    for h, T in enumerate(Ts):
        m = transformation_agreement(T, pts, l_pix, r_pix, K, M1, M2)
        c = int(np.sum(m))
        counts[h] = c
        if c > best_cnt:
            best, best_cnt, best_mask = h, c, m
    return counts, best, best_mask


def generate_hypotheses(points_3d, l_pix, K, n_iter):
    """The sampling half of the loops (ransac.py:94-104, :155-171): np.random.choice(n, 4,
    replace=False) on the GLOBAL numpy RNG, cv2.solvePnP(EPNP) on the 4 points, rodriguez_to_mat.
    Returns (Ts (H, 3, 4), ok (H,) bool)."""
    import cv2
    dist = np.zeros((5, 1))
    Ts = np.zeros((n_iter, 3, 4))
    ok = np.zeros(n_iter, dtype=bool)
    for i in range(n_iter):
        idx = np.random.choice(len(points_3d), 4, replace=False)
        s, rvec, tvec = cv2.solvePnP(points_3d[idx], l_pix[idx], K, distCoeffs=dist,
                                     flags=cv2.SOLVEPNP_EPNP)
        if s:
            Ts[i] = rodriguez_to_mat(rvec, tvec)
            ok[i] = True
    return Ts, ok


def ransac_pnp_for_tracking_db(match_q, match_t, prev_links, cur_links, inliers_percent, K, M1, M2):
    """ransac.py:70-113 on arrays: gather links by queryIdx/trainIdx (:76-81), triangulate prev
    (:83), loop sample/solve/score keeping strictly-better counts (:94-112).
    Returns best_matches_idx (int64) or None."""
    n_iter = calc_ransac_iteration(inliers_percent)
    fprev = np.asarray(prev_links)[np.asarray(match_q)]
    fcur = np.asarray(cur_links)[np.asarray(match_t)]
    P, Q = K @ M1, K @ M2
    pts = triangulate_links(fprev, P, Q)
    l_pix = np.stack([fcur[:, 0], fcur[:, 2]], axis=1)
    r_pix = np.stack([fcur[:, 1], fcur[:, 2]], axis=1)
    Ts, ok = generate_hypotheses(pts, l_pix, K, n_iter)
    best_inliers, best_idx = 0, None
    for h in range(n_iter):
        if not ok[h]:
            continue
        m = transformation_agreement(Ts[h], pts, l_pix, r_pix, K, M1, M2)
        c = np.sum(m)
        if c > best_inliers:
            best_inliers, best_idx = c, np.where(m)[0]
    return best_idx


# --------------------------------------------------------------------------------------
# The real third-party engine (cv2), as the reference calls it — used for goldens and as the
# CPU baseline on the GPU box (cv2 is part of the image; /root/reference is not needed).
# --------------------------------------------------------------------------------------
def cv2_match(q, t, cross_check=False):
    import cv2
    m = cv2.BFMatcher(normType=cv2.NORM_HAMMING, crossCheck=cross_check).match(q, t)
    return (np.array([x.queryIdx for x in m], dtype=np.int32),
            np.array([x.trainIdx for x in m], dtype=np.int32),
            np.array([x.distance for x in m], dtype=np.float32))


def cv2_knn2(q, t):
    import cv2
    m = cv2.BFMatcher(normType=cv2.NORM_HAMMING, crossCheck=False).knnMatch(q, t, k=2)
    idx2 = np.full((len(m), 2), -1, dtype=np.int32)
    dist2 = np.full((len(m), 2), -1, dtype=np.int32)
    for i, pair in enumerate(m):
        for k, x in enumerate(pair):
            idx2[i, k] = x.trainIdx
            dist2[i, k] = int(x.distance)
    return idx2, dist2


# ---------------------------------------------------------------------------------------------------
# Loop-closure candidate gating (backend/loop/loop_closure.py:140-228, backend/loop/graph.py)
# ---------------------------------------------------------------------------------------------------
def so3_logmap(R):
    """GTSAM SO3::Logmap (published algorithm; gtsam is not installed here — pinned against scipy.linalg.logm
    by tests/test_oracle.py)."""
    tr = np.trace(R)
    if tr + 1.0 < 1e-3:
        k = int(np.argmax(np.diag(R)))
        d = R[k, k]
        w = (np.pi / np.sqrt(2.0 + 2.0 * d)) * (R[:, k] + np.eye(3)[k])
        skew = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
        return -w if skew[k] < 0 else w
    tr3 = tr - 3.0
    mag = np.arccos((tr - 1.0) / 2.0) / (2.0 * np.sin(np.arccos((tr - 1.0) / 2.0))) if tr3 < -1e-7 else 0.5 - tr3 / 12.0
    return mag * np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])


def pose3_logmap(R, t):
    """GTSAM Pose3::Logmap: (rotation vector, V^-1 t)."""
    w = so3_logmap(R)
    th = np.linalg.norm(w)
    if th < 1e-10:
        return np.concatenate([w, t])
    W = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]) / th
    WT = W @ t
    u = t - (0.5 * th) * WT + (1 - th / (2.0 * np.tan(0.5 * th))) * (W @ WT)
    return np.concatenate([w, u])


def dijkstra_path(adj, start, end):
    """backend/loop/graph.py:57-97 restated: heapq Dijkstra, (distance, node) order, strict-< relaxation.
    adj: {node: {neighbour: weight}}."""
    import heapq
    if start == end:
        return [start]
    pq = [(0, start)]
    dist = {v: float("inf") for v in adj}
    dist[start] = 0
    pred = {v: None for v in adj}
    while pq:
        d, u = heapq.heappop(pq)
        if u == end:
            break
        if d > dist[u]:
            continue
        for v, w in adj[u].items():
            nd = d + w
            if nd < dist[v]:
                dist[v], pred[v] = nd, u
                heapq.heappush(pq, (nd, v))
    if dist[end] == float("inf"):
        return None
    path, v = [], end
    while v is not None:
        path.insert(0, v)
        v = pred[v]
    return path


def gate_distances(poses, edges, n, gap=10, graph_cls=None):
    """check_candidate (loop_closure.py:164-196) for query keyframe n against every i < n - gap.
    poses (K, 3, 4) camera-to-world; edges: list of (v1, v2, cov (6,6)).  With graph_cls (the reference's own
    Graph class, backend/loop/graph.py) the shortest paths come from the unmodified reference."""
    K = len(poses)
    if graph_cls is not None:
        g = graph_cls()
        for a, b, c in edges:
            g.add_edge(a, b, np.array(c, dtype=np.float64))
        path_of = lambda i: g.get_shortest_path(i, n)
        cov_of = lambda a, b: g.get_cov(a, b)
    else:
        adj, covs = {}, {}
        for a, b, c in edges:
            w = np.linalg.det(np.asarray(c, dtype=np.float64))
            adj.setdefault(a, {})[b] = w
            adj.setdefault(b, {})[a] = w
            covs[(a, b)] = covs[(b, a)] = np.asarray(c, dtype=np.float64)
        path_of = lambda i: dijkstra_path(adj, i, n)
        cov_of = lambda a, b: covs[(a, b)]
    out = np.full(K, np.inf)
    Rn, tn = poses[n][:3, :3], poses[n][:3, 3]
    for i in range(0, n - gap):
        path = path_of(i)
        if path is None:
            continue
        cov = None
        for a, b in zip(path[:-1], path[1:]):
            cov = np.array(cov_of(a, b), dtype=np.float64) if cov is None else cov + cov_of(a, b)
        xi = pose3_logmap(Rn.T @ poses[i][:3, :3], Rn.T @ (poses[i][:3, 3] - tn))
        out[i] = np.sqrt(xi @ np.linalg.solve(cov, xi))
    return out


def gate_distances_typed(poses, adj, cov_dict, index_list, n, gap=10):
    """get_good_candidates' distances (loop_closure.py:164-228) with the reference's own bookkeeping: `adj` is
    cov_dijkstra_graph.graph ({frame: {frame: weight}}, insertion order), `cov_dict` relative_covariance_dict
    (str((a, b)) entries for consecutive keyframes, :264-279, int entries per keyframe, :282,:286).  The
    covariance of the hop a -> b is looked up as get_relative_covariance_along_path does (:121-130):
    cov_dict[str((a, b))], else what get_relative_consecutive_covariance returns first (:94-97): cov_dict[b].
    poses: (K, 3, 4) camera-to-world by POSITION in index_list.  Returns (K,) distances, inf where not a candidate."""
    K = len(index_list)
    out = np.full(K, np.inf)
    Rn, tn = poses[n][:3, :3], poses[n][:3, 3]
    for i in range(0, n - gap):
        path = dijkstra_path(adj, index_list[i], index_list[n])
        if path is None:
            continue
        cov = None
        for a, b in zip(path[:-1], path[1:]):
            c = np.asarray(cov_dict[str((a, b))] if str((a, b)) in cov_dict else cov_dict[b], dtype=np.float64)
            cov = c.copy() if cov is None else cov + c
        xi = pose3_logmap(Rn.T @ poses[i][:3, :3], Rn.T @ (poses[i][:3, 3] - tn))
        out[i] = np.sqrt(xi @ np.linalg.solve(cov, xi))
    return out

