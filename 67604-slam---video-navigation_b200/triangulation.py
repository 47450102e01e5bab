"""Drop-in mirror of final_project/algorithms/triangulation.py (same names and signatures).

The per-link Python loop + np.linalg.svd of the reference (triangulation.py:27-50) becomes one
kernel launch over all links.
"""
from __future__ import annotations

import numpy as np

from . import ops
from .matching import _ThreadLocalStaging

_st = _ThreadLocalStaging()


def links_to_array(links) -> np.ndarray:
    """List of Link-like objects (x_left, x_right, y) or an (M, 3) array -> (M, 3) float64."""
    if isinstance(links, np.ndarray):
        return np.ascontiguousarray(links, dtype=np.float64).reshape(-1, 3)
    out = np.empty((len(links), 3), dtype=np.float64)
    for i, ln in enumerate(links):
        out[i, 0] = ln.x_left
        out[i, 1] = ln.x_right
        out[i, 2] = ln.y
    return out


def triangulate_link_array(links_xyz: np.ndarray, p, q) -> np.ndarray:
    """(M, 3) [x_left, x_right, y] float64 -> (M, 3) float64 points, on the GPU."""
    arr = np.ascontiguousarray(links_xyz, dtype=np.float64).reshape(-1, 3)
    if arr.shape[0] == 0:
        return np.zeros((0, 3))
    if ops.same_rows_stereo(p, q):
        xyz = ops.triangulate_links(_st.to_device("links", arr), p, q)
    else:   # general P, Q: the (x_left, y) / (x_right, y) pixel pairs are laid out on the host
        pxy = _st.to_device("pxy", np.ascontiguousarray(arr[:, [0, 2]]))
        qxy = _st.to_device("qxy", np.ascontiguousarray(arr[:, [1, 2]]))
        xyz = ops.triangulate_dlt(pxy, qxy, p, q)
    return _st.to_host("xyz", xyz)


def linear_least_squares_triangulation(P, Q, kp_left, kp_right):
    """triangulation.py:5-24: DLT of one left/right pixel pair -> (3,) float64."""
    pxy = np.asarray([[kp_left[0], kp_left[1]]], dtype=np.float64)
    qxy = np.asarray([[kp_right[0], kp_right[1]]], dtype=np.float64)
    return triangulate_points(P, Q, pxy, qxy)[0]


def triangulate_points(P, Q, pxy, qxy) -> np.ndarray:
    """Batched general DLT: (n, 2) left pixels, (n, 2) right pixels -> (n, 3)."""
    pxy = np.ascontiguousarray(pxy, dtype=np.float64).reshape(-1, 2)
    qxy = np.ascontiguousarray(qxy, dtype=np.float64).reshape(-1, 2)
    if pxy.shape[0] == 0:
        return np.zeros((0, 3))
    xyz = ops.triangulate_dlt(_st.to_device("pxy", pxy), _st.to_device("qxy", qxy), P, Q)
    return _st.to_host("xyz", xyz)


def triangulate_last_frame(tracking_db, p, q, links=None):
    """triangulation.py:27-38."""
    if links is None:
        links = tracking_db.all_last_frame_links()
    return triangulate_link_array(links_to_array(links), p, q)


def triangulate_links(links, p, q):
    """triangulation.py:41-50."""
    return triangulate_link_array(links_to_array(links), p, q)
