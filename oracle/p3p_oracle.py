"""ORACLE — test infrastructure only.  NOT part of the product path.

numpy restatement of the minimal-sample pose solver behind `slamfe_ransac_hypotheses` (SURVEY.md
section 8f rank 1: the reference's `cv2.solvePnP(..., flags=cv2.SOLVEPNP_EPNP)` on 4 sampled points,
final_project/algorithms/ransac.py:95-104 and :157-171, moved to the GPU).

The reference's solver is OpenCV's EPnP, which for exactly 4 points is implementation-defined (the
12x12 moment matrix has a 4-dimensional null space whose basis depends on the SVD routine), so it is
NOT reproduced bit for bit; hypothesis generation in the reference is also unseeded
(`np.random.choice` on the global RNG).  The GPU generator instead solves the exact minimal problem:

  P3P (Grunert's formulation: law of cosines -> quartic in v = s3/s1) on the first three sampled
  points, up to four poses, and the fourth sampled point picks the pose with the smallest left-image
  reprojection error — the same design as cv2.SOLVEPNP_P3P, against which this file is pinned
  (tests/test_oracle.py) together with ground-truth recovery on noise-free samples.

The quartic coefficients were derived with sympy (elimination of s1 and u = s2/s1 from the three
cosine-law equations); the quartic is solved here with numpy.roots, i.e. independently of the
closed-form Ferrari solver the CUDA kernel uses.

Only tests/ may import this module.
"""
from __future__ import annotations

import numpy as np


def quartic_coefficients(a2, b2, c2, ca, cb, cg):
    """A4..A0 of the quartic in v = s3/s1.  a2 = |P2-P3|^2, b2 = |P1-P3|^2, c2 = |P1-P2|^2,
    ca = j2.j3, cb = j1.j3, cg = j1.j2 (unit bearings)."""
    x9 = a2 * a2 + b2 * b2 + c2 * c2 - 2 * a2 * c2
    x15 = (-ca * cg * b2 * b2 - cb * a2 * a2 - cb * c2 * c2 + 2 * cb * a2 * c2 + a2 * b2 * ca * cg + ca * cg * b2 * c2)
    A4 = b2 * (-2 * a2 * b2 - 4 * b2 * c2 * ca * ca + 2 * b2 * c2 + x9)
    A3 = -4 * b2 * (-cb * 2 * b2 * c2 * ca * ca - cb * a2 * b2 + cb * b2 * c2 - x15)
    A2 = 2 * b2 * (-4 * a2 * c2 * cb * cb - 4 * ca * cg * cb * a2 * b2 - 4 * ca * cg * cb * b2 * c2
                   - 2 * b2 * c2 * ca * ca + 2 * cg * cg * b2 * b2 - 2 * a2 * b2 * cg * cg + 2 * cb * cb * a2 * a2
                   + 2 * cb * cb * c2 * c2 + 2 * ca * ca * b2 * b2 + a2 * a2 - b2 * b2 + c2 * c2 - 2 * a2 * c2)
    A1 = -4 * b2 * (-cb * 2 * a2 * b2 * cg * cg + cb * a2 * b2 - cb * b2 * c2 - x15)
    A0 = b2 * (2 * a2 * b2 - 4 * a2 * b2 * cg * cg - 2 * b2 * c2 + x9)
    return A4, A3, A2, A1, A0


def _frame(p0, p1, p2):
    e1 = p1 - p0
    n1 = np.linalg.norm(e1)
    e3 = np.cross(e1, p2 - p0)
    n3 = np.linalg.norm(e3)
    if n1 == 0 or n3 == 0:
        return None
    e1, e3 = e1 / n1, e3 / n3
    return np.stack([e1, np.cross(e3, e1), e3], axis=1)  # columns


def p3p(P, J):
    """All poses (R, t) with R @ P[i] + t = s_i * J[i], s_i > 0.  P (3,3) world points (rows),
    J (3,3) unit bearing vectors (rows)."""
    a2 = float(np.sum((P[1] - P[2]) ** 2)); b2 = float(np.sum((P[0] - P[2]) ** 2)); c2 = float(np.sum((P[0] - P[1]) ** 2))
    ca, cb, cg = float(J[1] @ J[2]), float(J[0] @ J[2]), float(J[0] @ J[1])
    if min(a2, b2, c2) <= 0:
        return []
    A = quartic_coefficients(a2, b2, c2, ca, cb, cg)
    if not np.all(np.isfinite(A)) or A[0] == 0:
        return []
    out = []
    Fw = _frame(P[0], P[1], P[2])
    if Fw is None:
        return []
    for v in np.roots(A):
        if abs(v.imag) > 1e-7 * max(1.0, abs(v.real)) or v.real <= 0:
            continue
        v = float(v.real)
        for _ in range(2):  # Newton polish on the quartic
            f = (((A[0] * v + A[1]) * v + A[2]) * v + A[3]) * v + A[4]
            df = ((4 * A[0] * v + 3 * A[1]) * v + 2 * A[2]) * v + A[3]
            if df != 0:
                v -= f / df
        den = 2 * b2 * (cg - ca * v)
        if abs(den) < 1e-12 * b2:
            continue
        u = ((a2 - c2) * (1 + v * v - 2 * cb * v) - b2 * (v * v - 1)) / den
        w = 1 + v * v - 2 * v * cb
        if u <= 0 or v <= 0 or w <= 0:
            continue
        s1 = np.sqrt(b2 / w)
        C = np.stack([s1 * J[0], u * s1 * J[1], v * s1 * J[2]])
        Fc = _frame(C[0], C[1], C[2])
        if Fc is None:
            continue
        R = Fc @ Fw.T
        t = C[0] - R @ P[0]
        out.append((R, t))
    return out


def solve_sample(pts4, pix4, K):
    """Pose hypothesis from 4 correspondences: P3P on the first three, the fourth picks the
    solution.  Returns (T (3,4) = [R|t] world -> camera, ok)."""
    Kinv = np.linalg.inv(K)
    J = (Kinv @ np.hstack([pix4[:3], np.ones((3, 1))]).T).T
    J /= np.linalg.norm(J, axis=1, keepdims=True)
    best, best_err = None, np.inf
    for R, t in p3p(np.asarray(pts4[:3], float), J):
        c = R @ pts4[3] + t
        if c[2] <= 0:
            continue
        uv = (K @ c)[:2] / c[2]
        err = float(np.sum((uv - pix4[3]) ** 2))
        if err < best_err:
            best, best_err = np.hstack([R, t[:, None]]), err
    if best is None:
        return np.zeros((3, 4)), False
    return best, True
