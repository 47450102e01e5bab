"""Mirror of the hot-path parts of final_project/utils.py (rodriguez_to_mat :16-18,
read_cameras :36-51, module constants K, M1, M2, P, Q :137-138).

The reference reads calib.txt from a hard-coded path at import time; here the KITTI-00 rows are
the default and `read_cameras(path)` accepts any calib.txt.  I/O, plotting and pickle helpers of
the reference's utils are out of scope.
"""
from __future__ import annotations

import numpy as np

from .synth import KITTI00_P0, KITTI00_P1


def rodriguez_to_mat(rvec, tvec):
    """utils.py:16-18: hstack(cv2.Rodrigues(rvec)[0], tvec) -> (3, 4) float64."""
    import cv2
    rot, _ = cv2.Rodrigues(rvec)
    return np.hstack((rot, tvec))


def cameras_from_projection(p0, p1):
    """utils.py:45-51: k = P0[:, :3]; m1 = inv(k) @ P0; m2 = inv(k) @ P1."""
    p0 = np.asarray(p0, dtype=np.float64).reshape(3, 4)
    p1 = np.asarray(p1, dtype=np.float64).reshape(3, 4)
    k = p0[:, :3]
    m1 = np.linalg.inv(k) @ p0
    m2 = np.linalg.inv(k) @ p1
    return k, m1, m2


def read_cameras(calib_path=None):
    """utils.py:36-51 / Inputs.py:22-37.  calib_path=None -> KITTI sequence 00."""
    if calib_path is None:
        return cameras_from_projection(KITTI00_P0, KITTI00_P1)
    with open(calib_path) as f:
        l1 = f.readline().split()[1:]
        l2 = f.readline().split()[1:]
    return cameras_from_projection([float(i) for i in l1], [float(i) for i in l2])


K, M1, M2 = read_cameras()
P, Q = K @ M1, K @ M2
