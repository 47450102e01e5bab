"""Drop-in mirror of final_project/algorithms/ransac.py (same names, signatures, returns).

What changes: the reference scores ONE hypothesis per Python iteration with NumPy
(ransac.py:106, :174); here all hypotheses of a call are generated first (same global-RNG call
sequence: np.random.choice + cv2.solvePnP(EPNP) on the host, ransac.py:95-98 — the minimal solver
is a "next" row) and then scored against all correspondences in a single
kernel launch that also returns the winner (first strictly-best) and its inlier mask.  With the
same np.random seed the outputs equal the reference's.
"""
from __future__ import annotations

import numpy as np

from . import ops
from . import utils as _utils
from .matching import _ThreadLocalStaging
from .triangulation import links_to_array, triangulate_link_array
from .utils import rodriguez_to_mat

SUCCESS_PROBABILITY = 0.9999999999  # ransac.py:9

K, M1, M2 = _utils.K, _utils.M1, _utils.M2  # ransac.py:11
P, Q = K @ M1, K @ M2                        # ransac.py:12

_st = _ThreadLocalStaging()


def set_cameras(k, m1, m2):
    """Rebind the module-level cameras (the reference fixes them at import, ransac.py:11-12)."""
    global K, M1, M2, P, Q
    K, M1, M2 = np.asarray(k, float), np.asarray(m1, float), np.asarray(m2, float)
    P, Q = K @ M1, K @ M2


def get_pixels_from_links(links):
    """ransac.py:15-25."""
    pixels_first = []
    pixels_second = []
    for link in links:
        pixels_first.append((link.x_left, link.y))
        pixels_second.append((link.x_right, link.y))
    return pixels_first, pixels_second


def score_hypotheses(Ts, pts, l_pix, r_pix, hyp_valid=None):
    """All hypotheses x all points on the GPU.
    Returns (counts (H,) int32, best_index or -1, best_count, best_mask (N,) bool)."""
    Ts = np.ascontiguousarray(Ts, dtype=np.float64).reshape(-1, 3, 4)
    pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 3)
    l_pix = np.ascontiguousarray(l_pix, dtype=np.float64).reshape(-1, 2)
    r_pix = np.ascontiguousarray(r_pix, dtype=np.float64).reshape(-1, 2)
    if Ts.shape[0] == 0:
        return np.zeros(0, np.int32), -1, 0, np.zeros(pts.shape[0], bool)
    hv = None if hyp_valid is None else _st.to_device("hv", np.asarray(hyp_valid, dtype=np.uint8))
    ptsd = _st.to_device("pts", pts)
    H = Ts.shape[0]
    import torch
    packed = torch.empty((H + 2,), dtype=torch.int32, device=ptsd.device)   # [counts | best index, best count]
    counts, best, mask = ops.ransac_score(_st.to_device("T", Ts), ptsd, _st.to_device("lp", l_pix),
                                          _st.to_device("rp", r_pix), K, M1, M2, hyp_valid=hv,
                                          out={"counts": packed[:H].view(1, H), "best": packed[H:].view(1, 2)})
    host = _st.to_host("cb", packed)
    mask_h = _st.to_host("mask", mask).astype(bool)
    return host[:-2].copy(), int(host[-2]), int(host[-1]), mask_h


def transformation_agreement(T, traingulated_pts, ordered_cur_left_pix_values, ordered_cur_right_pix_values):
    """ransac.py:28-56 -> (N,) bool."""
    pts = np.asarray(traingulated_pts, dtype=np.float64)
    _, _, _, mask = score_hypotheses(np.asarray(T, dtype=np.float64).reshape(1, 3, 4), pts,
                                     ordered_cur_left_pix_values, ordered_cur_right_pix_values)
    return mask


def calc_ransac_iteration(inliers_percent):
    """ransac.py:59-67."""
    suc_prob = SUCCESS_PROBABILITY
    outliers_prob = 1 - (inliers_percent / 100) + 0.0000000001
    min_set_size = 4
    ransac_iterations = int(np.log(1 - suc_prob) / np.log(1 - np.power(1 - outliers_prob, min_set_size))) + 1
    return ransac_iterations


def _gather(matches_l_l, prev_links, cur_links):
    """ransac.py:76-81 / :132-138 + :83,:85-88: points and pixel arrays for the matched links."""
    n = len(matches_l_l)
    qi = np.fromiter((m.queryIdx for m in matches_l_l), dtype=np.int64, count=n)
    ti = np.fromiter((m.trainIdx for m in matches_l_l), dtype=np.int64, count=n)
    prev = links_to_array(prev_links)[qi]
    cur = links_to_array(cur_links)[ti]
    points_3d = triangulate_link_array(prev, P, Q)
    l_pix = np.ascontiguousarray(cur[:, [0, 2]])
    r_pix = np.ascontiguousarray(cur[:, [1, 2]])
    return points_3d, l_pix, r_pix


def generate_hypotheses(points_3d, l_pix, n_iter):
    """ransac.py:94-104: the host half of the loop, same RNG/solver call sequence.
    Returns (Ts (H, 3, 4), ok (H,) uint8)."""
    import cv2
    diff_coeff = np.zeros((5, 1))
    Ts = np.zeros((n_iter, 3, 4))
    ok = np.zeros(n_iter, dtype=np.uint8)
    for i in range(n_iter):
        random_idx = np.random.choice(len(points_3d), 4, replace=False)
        success, rvec, tvec = cv2.solvePnP(points_3d[random_idx], l_pix[random_idx], K,
                                           distCoeffs=diff_coeff, flags=cv2.SOLVEPNP_EPNP)
        if success:
            Ts[i] = rodriguez_to_mat(rvec, tvec)
            ok[i] = 1
    return Ts, ok


# Hypothesis generator of the RANSAC entry points below:
#   "cv2"      host loop with the reference's exact call sequence (np.random.choice + cv2 EPnP):
#              with the same np.random seed the outputs equal the reference's (default);
#   "p3p_gpu"  slamfe_ransac_hypotheses: sampling + minimal solver on the GPU, the whole RANSAC stays
#              on the device.  Different (exact, minimal) solver and RNG, so individual hypotheses
#              differ from cv2's; the result is statistically equivalent.
HYPOTHESIS_GENERATOR = "cv2"
_gpu_seed = [0x5EED]


def set_hypothesis_generator(name, seed=None):
    global HYPOTHESIS_GENERATOR
    if name not in ("cv2", "p3p_gpu"):
        raise ValueError("generator must be 'cv2' or 'p3p_gpu'")
    HYPOTHESIS_GENERATOR = name
    if seed is not None:
        _gpu_seed[0] = int(seed)


def ransac_device(points_3d, l_pix, r_pix, n_iter, seed=None, sample_idx=None, refit=False):
    """Generate n_iter hypotheses on the GPU and score them, one H2D and one D2H.
    Returns (best_index or -1, best_count, best_mask (N,) bool, T_best (3,4) or None); with refit=True
    T_best is the pose refit on the consensus set (slamfe_pnp_refit, ransac.py:185-193) when it has at
    least 4 inliers, else the winning hypothesis."""
    import torch
    pts = _st.to_device("pts", np.ascontiguousarray(points_3d, dtype=np.float64).reshape(-1, 3))
    lp = _st.to_device("lp", np.ascontiguousarray(l_pix, dtype=np.float64).reshape(-1, 2))
    rp = _st.to_device("rp", np.ascontiguousarray(r_pix, dtype=np.float64).reshape(-1, 2))
    if seed is None:
        _gpu_seed[0] += 1
        seed = _gpu_seed[0]
    si = None if sample_idx is None else _st.to_device("si", np.ascontiguousarray(sample_idx, dtype=np.int32))
    T, valid = ops.ransac_hypotheses(pts, lp, K, n_iter, seed=seed, sample_idx=si)
    _, best, mask = ops.ransac_score(T, pts, lp, rp, K, M1, M2, hyp_valid=valid)
    b = _st.to_host("cb", best.view(-1))
    mask_h = _st.to_host("mask", mask).astype(bool)
    bi = int(b[0])
    if refit and bi >= 0 and int(b[1]) >= 4:
        Tr, status, _ = ops.pnp_refit(T, best, pts, lp, mask, K)
        Tb = _st.to_host("Tb", Tr[0]).copy()
        if int(_st.to_host("rs", status)[0]) == 0:
            Tb = None
        return bi, int(b[1]), mask_h, Tb
    Tb = _st.to_host("Tb", T[bi]).copy() if bi >= 0 else None
    return bi, int(b[1]), mask_h, Tb


def ransac_pnp_for_tracking_db(matches_l_l, prev_links, cur_links, inliers_percent):
    """ransac.py:70-113 -> best_matches_idx (int64 array) or None."""
    ransac_iterations = calc_ransac_iteration(inliers_percent)
    points_3d, l_pix, r_pix = _gather(matches_l_l, prev_links, cur_links)
    if HYPOTHESIS_GENERATOR == "p3p_gpu":
        if len(points_3d) < 4:  # np.random.choice(n < 4, 4, replace=False) raises in the reference
            raise ValueError("Cannot take a larger sample than population when 'replace=False'")
        best, _, mask, _ = ransac_device(points_3d, l_pix, r_pix, ransac_iterations)
    else:
        Ts, ok = generate_hypotheses(points_3d, l_pix, ransac_iterations)
        _, best, _, mask = score_hypotheses(Ts, points_3d, l_pix, r_pix, hyp_valid=ok)
    if best < 0:
        return None
    return np.where(mask)[0]


class Pose3:
    """Minimal gtsam.Pose3 stand-in used when gtsam is not installed (ransac.py:199-200)."""

    def __init__(self, R, t):
        self._R = np.asarray(R, dtype=np.float64).reshape(3, 3)
        self._t = np.asarray(t, dtype=np.float64).reshape(3)

    def inverse(self):
        return Pose3(self._R.T, -self._R.T @ self._t)

    def rotation(self):
        return _Rot3(self._R)

    def translation(self):
        return self._t

    def matrix(self):
        m = np.eye(4)
        m[:3, :3] = self._R
        m[:3, 3] = self._t
        return m


class _Rot3:
    def __init__(self, R):
        self._R = R

    def matrix(self):
        return self._R


def _pose3(T):
    try:
        import gtsam
        return gtsam.Pose3(gtsam.Rot3(T[:3, :3]), gtsam.Point3(T[:3, 3]))
    except ImportError:
        return Pose3(T[:3, :3], T[:3, 3])


def ransac_pnp(matches_l_l, prev_links, cur_links, inliers_percent=50):
    """ransac.py:116-204 -> (camera_to_world Pose3, best_matches_idx, best_inliers)."""
    import cv2
    ransac_iterations = calc_ransac_iteration(inliers_percent)
    points_3d, l_pix, r_pix = _gather(matches_l_l, prev_links, cur_links)
    if HYPOTHESIS_GENERATOR == "p3p_gpu":
        # the whole call on the device: sampling + P3P, scoring, refit on the consensus set; no host solve
        if len(points_3d) < 4:
            raise ValueError("Cannot take a larger sample than population when 'replace=False'")
        best, best_inliers, mask, T = ransac_device(points_3d, l_pix, r_pix, ransac_iterations, refit=True)
        best_matches_idx = np.where(mask)[0] if best >= 0 else []
        if len(best_matches_idx) < 4:
            return None, [], []
        if T is None:
            return None, None, None
        return _pose3(T).inverse(), best_matches_idx, np.int64(best_inliers)
    else:
        Ts, ok = generate_hypotheses(points_3d, l_pix, ransac_iterations)
        _, best, best_inliers, mask = score_hypotheses(Ts, points_3d, l_pix, r_pix, hyp_valid=ok)
    best_matches_idx = np.where(mask)[0] if best >= 0 else []
    if len(best_matches_idx) < 4:
        return None, [], []
    success, rvec, tvec = cv2.solvePnP(points_3d[best_matches_idx], l_pix[best_matches_idx], K,
                                       distCoeffs=np.zeros((5, 1)), flags=cv2.SOLVEPNP_EPNP)
    if success:
        T = rodriguez_to_mat(rvec, tvec)
        world_to_camera = _pose3(T)
        return world_to_camera.inverse(), best_matches_idx, np.int64(best_inliers)
    return None, None, None
