"""Seeded synthetic KITTI-00-shaped front-end inputs.

The reference ships no data (KITTI is expected at hard-coded paths,
final_project/arguments.py:3-14), so every test / bench input is generated here:
61-byte MLDB-shaped descriptors (byte 60 uses its low 6 bits, like cv2 AKAZE output),
fp32-representable keypoints on a 1241x376 image, KITTI-00 calibration.
"""
from __future__ import annotations

import numpy as np

DESC_BYTES = 61
IMG_W, IMG_H = 1241, 376

# KITTI sequence-00 projection matrices (calib.txt rows P0 / P1).
KITTI00_P0 = np.array([[7.188560000000e+02, 0.0, 6.071928000000e+02, 0.0],
                       [0.0, 7.188560000000e+02, 1.852157000000e+02, 0.0],
                       [0.0, 0.0, 1.0, 0.0]])
KITTI00_P1 = np.array([[7.188560000000e+02, 0.0, 6.071928000000e+02, -3.861448000000e+02],
                       [0.0, 7.188560000000e+02, 1.852157000000e+02, 0.0],
                       [0.0, 0.0, 1.0, 0.0]])


def descriptors(rng: np.random.Generator, n: int) -> np.ndarray:
    """(n, 61) uint8, i.i.d. bits, byte 60 masked to 6 bits."""
    d = rng.integers(0, 256, size=(n, DESC_BYTES), dtype=np.uint8)
    d[:, 60] &= 0x3F
    return d


def flip_bits(rng: np.random.Generator, d: np.ndarray, p: float = 0.08) -> np.ndarray:
    """Copy of d with every bit flipped independently with probability p."""
    bits = rng.random(size=(d.shape[0], d.shape[1], 8)) < p
    mask = np.packbits(bits, axis=2, bitorder="little")[:, :, 0]
    out = d ^ mask
    out[:, 60] &= 0x3F
    return out


def paired_descriptors(rng: np.random.Generator, base: np.ndarray, n_out: int | None = None,
                       match_frac: float = 0.6, flip: float = 0.08, dup_frac: float = 0.01):
    """A second descriptor set related to `base`: a `match_frac` subset are bit-flipped copies
    (true matches), the rest fresh random rows, row-permuted; `dup_frac` of the rows are then
    overwritten with exact duplicates of other rows to force distance ties.
    Returns (out, src) where src[j] = base row that out[j] came from, or -1."""
    n = base.shape[0]
    n_out = n if n_out is None else n_out
    out = descriptors(rng, n_out)
    src = np.full(n_out, -1, dtype=np.int64)
    n_match = min(int(match_frac * n), n_out)
    from_rows = rng.permutation(n)[:n_match]
    to_rows = rng.permutation(n_out)[:n_match]
    out[to_rows] = flip_bits(rng, base[from_rows], flip)
    src[to_rows] = from_rows
    n_dup = int(dup_frac * n_out)
    if n_dup and n_out > 1:
        a = rng.integers(0, n_out, size=n_dup)
        b = rng.integers(0, n_out, size=n_dup)
        out[a] = out[b]
        src[a] = src[b]
    return out, src


def stereo_frame(rng: np.random.Generator, n: int, outlier_frac: float = 0.15):
    """One rectified stereo frame: (desc_l, desc_r, pts_l, pts_r), pts float32 (n, 2) = (x, y).
    Right keypoint of a true match sits at x_r = x_l - d, d ~ U(2.5, 120), y_r = y_l + N(0, 0.5);
    `outlier_frac` of them violate the row / disparity test."""
    desc_l = descriptors(rng, n)
    desc_r, src = paired_descriptors(rng, desc_l)
    pts_l = np.stack([rng.uniform(20, 1220, n), rng.uniform(5, 370, n)], axis=1).astype(np.float32)
    pts_r = np.stack([rng.uniform(20, 1220, n), rng.uniform(5, 370, n)], axis=1).astype(np.float32)
    has = src >= 0
    d = rng.uniform(2.5, 120, n)
    xr = pts_l[src[has], 0] - d[has]
    yr = pts_l[src[has], 1] + rng.normal(0, 0.5, has.sum())
    bad = rng.random(has.sum()) < outlier_frac
    yr = np.where(bad & (rng.random(has.sum()) < 0.5), yr + 5.0, yr)
    xr = np.where(bad, xr + d[has] + 1.0, xr)
    pts_r[has, 0] = xr.astype(np.float32)
    pts_r[has, 1] = yr.astype(np.float32)
    return desc_l, desc_r, pts_l, pts_r


def next_frame_descriptors(rng: np.random.Generator, prev: np.ndarray, n: int):
    """Frame t+1 left descriptors: re-flipped copies of frame t rows plus fresh rows."""
    out, _ = paired_descriptors(rng, prev, n_out=n)
    return out


def links(rng: np.random.Generator, n: int) -> np.ndarray:
    """(n, 3) float64 [x_left, x_right, y] with fp32-representable values and valid disparity."""
    xl = rng.uniform(150, 1220, n).astype(np.float32)
    d = rng.uniform(2.5, 120, n).astype(np.float32)
    xr = (xl - d).astype(np.float32)
    yl = rng.uniform(5, 370, n).astype(np.float32)
    yr = (yl + rng.normal(0, 0.5, n)).astype(np.float32)
    y = (yl.astype(np.float64) + yr.astype(np.float64)) / 2
    return np.stack([xl.astype(np.float64), xr.astype(np.float64), y], axis=1)


def cameras():
    """K, M1, M2 exactly as final_project/utils.py:36-51 derives them from calib.txt."""
    k = KITTI00_P0[:, :3]
    m1 = np.linalg.inv(k) @ KITTI00_P0
    m2 = np.linalg.inv(k) @ KITTI00_P1
    return k, m1, m2


def _rodrigues(rvec):
    th = np.linalg.norm(rvec)
    if th < 1e-12:
        return np.eye(3)
    k = rvec / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * (Kx @ Kx)


def pnp_problem(rng: np.random.Generator, n: int, n_hyp: int, outlier_frac: float = 0.4,
                hyp_noise: float = 0.02):
    """RANSAC-PnP scoring inputs (config 3): 3-D points in frame t, their (noisy) left/right
    pixels in frame t+1 under a ground-truth motion, `outlier_frac` gross outliers, and `n_hyp`
    pose hypotheses (perturbed ground truth; a few exact, a few far off).
    Returns (Ts (H,3,4), pts (n,3), l_pix (n,2), r_pix (n,2)) float64."""
    k, m1, m2 = cameras()
    z = rng.uniform(4, 60, n)
    x = (rng.uniform(20, 1220, n) - k[0, 2]) * z / k[0, 0]
    y = (rng.uniform(5, 370, n) - k[1, 2]) * z / k[1, 1]
    pts = np.stack([x, y, z], axis=1)
    rvec = rng.normal(0, 0.01, 3)
    tvec = np.array([0.0, 0.0, -0.9]) + rng.normal(0, 0.05, 3)
    T_gt = np.hstack([_rodrigues(rvec), tvec[:, None]])
    X4 = np.hstack([pts, np.ones((n, 1))]).T
    pl = (k @ T_gt @ np.vstack([m1, [0, 0, 0, 1]]) @ X4)[:3].T
    pr = (k @ T_gt @ np.vstack([m2, [0, 0, 0, 1]]) @ X4)[:3].T
    l_pix = pl[:, :2] / pl[:, 2:3] + rng.normal(0, 0.5, (n, 2))
    r_pix = pr[:, :2] / pr[:, 2:3] + rng.normal(0, 0.5, (n, 2))
    out = rng.random(n) < outlier_frac
    l_pix[out] += rng.uniform(-80, 80, (out.sum(), 2))
    Ts = np.zeros((n_hyp, 3, 4))
    for h in range(n_hyp):
        s = hyp_noise * (0.02 if h % 7 == 0 else (5.0 if h % 11 == 0 else 1.0))
        Ts[h] = np.hstack([_rodrigues(rvec + rng.normal(0, s, 3)),
                           (tvec + rng.normal(0, 10 * s, 3))[:, None]])
    return Ts, pts, l_pix, r_pix


# ----------------------------------------------------------------------------------------------
# Whole-sequence generator on the GPU (torch is plumbing here: RNG + memory), BASELINE config 2.
# Every frame depends only on (seed, frame index), so any rank can generate any frame range and
# neighbouring ranks agree on their shared halo frame.
# ----------------------------------------------------------------------------------------------
def _fresh_count(seed, f, lo=1000, hi=2500):
    """Number of 'new' descriptors frame f introduces (pure function of seed and f)."""
    if f < 0:
        return 0
    return int(np.random.default_rng([seed, f, 17]).integers(lo, hi + 1))


def sequence_sizes(n_frames, first_frame=0, seed=1, lo=1000, hi=2500):
    """Per-frame keypoint counts: n_f = fresh_{f-1} + fresh_f in [2*lo, 2*hi] (KITTI-like 2-5k)."""
    return np.array([_fresh_count(seed, f - 1, lo, hi) + _fresh_count(seed, f, lo, hi) if f > 0
                     else 2 * _fresh_count(seed, 0, lo, hi)
                     for f in range(first_frame, first_frame + n_frames)], dtype=np.int32)


def frame_motion(seed, f):
    """World->camera motion from frame f-1 to frame f (pure function of seed and f): small rotation
    rvec ~ N(0, 0.01), forward translation t = (0, 0, -0.9) +- 0.05 m (KITTI-like)."""
    rng = np.random.default_rng([seed, f, 23])
    rvec = rng.normal(0, 0.01, 3)
    tvec = np.array([0.0, 0.0, -0.9]) + rng.normal(0, 0.05, 3)
    return _rodrigues(rvec), tvec


def torch_sequence(n_frames, first_frame=0, seed=1, device="cuda", lo=1000, hi=2500, row_align=16):
    """Synthetic stereo sequence in the packed layout of frontend.PackedSequence, generated on
    `device`.  Frame f's left descriptors are bit-flipped copies (p = 1/16 per bit) of the
    descriptors frame f-1 introduced plus fresh random ones, row-permuted; its right descriptors
    are flipped copies of 60 % of the left rows (permuted) plus random rows; 1 % exact duplicates
    force distance ties.  Geometry is consistent in 3-D: every fresh keypoint gets a pixel and a
    disparity (i.e. a 3-D point in its frame's camera coordinates, KITTI-00 calibration); in the next
    frame the carried-over keypoints are those points moved by frame_motion() and re-projected with
    N(0, 0.5) px noise, so frame-to-frame RANSAC-PnP finds a real consensus set; right keypoints of
    true stereo matches sit at x_left - disparity except for 15 % gross outliers.
    Returns a dict of tensors + numpy offset arrays."""
    import torch
    fx, cx, cy = float(KITTI00_P0[0, 0]), float(KITTI00_P0[0, 2]), float(KITTI00_P0[1, 2])
    fxb = -float(KITTI00_P1[0, 3])  # fx * baseline

    def gen(f, tag):
        g = torch.Generator(device=device)
        g.manual_seed((seed * 1000003 + f) * 31 + tag)
        return g

    def rand_desc(n, g):
        d = torch.randint(0, 256, (n, DESC_BYTES), dtype=torch.uint8, device=device, generator=g)
        d[:, 60] &= 0x3F
        return d

    def flip(d, g):
        m = torch.randint(0, 256, (4,) + tuple(d.shape), dtype=torch.uint8, device=device, generator=g)
        out = d ^ (m[0] & m[1] & m[2] & m[3])
        out[:, 60] &= 0x3F
        return out

    def fresh(f):
        """Descriptors, left pixels (n, 2) and disparities (n,) of the keypoints frame f introduces."""
        n = _fresh_count(seed, f, lo, hi)
        g = gen(f, 1)
        d = rand_desc(n, g)
        g3 = gen(f, 3)
        pix = torch.rand((n, 2), device=device, generator=g3, dtype=torch.float64) * \
            torch.tensor([1200.0, 365.0], device=device, dtype=torch.float64) + \
            torch.tensor([20.0, 5.0], device=device, dtype=torch.float64)
        disp = torch.rand((n,), device=device, generator=g3, dtype=torch.float64) * 117.5 + 2.5
        return d, pix, disp

    def carry(pix, disp, f, g):
        """Pixels / disparities in frame f of points seen at (pix, disp) in frame f-1."""
        R, t = frame_motion(seed, f)
        R = torch.from_numpy(R).to(device)
        t = torch.from_numpy(t).to(device)
        Z = fxb / disp
        P = torch.stack([(pix[:, 0] - cx) * Z / fx, (pix[:, 1] - cy) * Z / fx, Z], dim=1)
        Pn = P @ R.T + t
        ok = Pn[:, 2] > 1.0
        Zn = torch.where(ok, Pn[:, 2], torch.ones_like(Pn[:, 2]))
        new_pix = torch.stack([fx * Pn[:, 0] / Zn + cx, fx * Pn[:, 1] / Zn + cy], dim=1) + \
            0.5 * torch.randn(pix.shape, device=device, generator=g, dtype=torch.float64)
        new_disp = fxb / Zn
        # points that passed the camera (or would have an absurd disparity) become unrelated keypoints
        ok = ok & (new_disp < 400.0)
        rnd = torch.rand(pix.shape, device=device, generator=g, dtype=torch.float64) * \
            torch.tensor([1200.0, 365.0], device=device, dtype=torch.float64) + \
            torch.tensor([20.0, 5.0], device=device, dtype=torch.float64)
        new_pix = torch.where(ok[:, None], new_pix, rnd)
        new_disp = torch.where(ok, new_disp, torch.full_like(new_disp, 30.0))
        return new_pix, new_disp

    sizes = sequence_sizes(n_frames, first_frame, seed, lo, hi)
    off = np.zeros(n_frames + 1, dtype=np.int64)
    np.cumsum((sizes.astype(np.int64) + row_align - 1) // row_align * row_align, out=off[1:])
    total = int(off[-1])
    desc_l = torch.zeros((total, DESC_BYTES), dtype=torch.uint8, device=device)
    desc_r = torch.zeros((total, DESC_BYTES), dtype=torch.uint8, device=device)
    pts_l = torch.zeros((total, 2), dtype=torch.float32, device=device)
    pts_r = torch.zeros((total, 2), dtype=torch.float32, device=device)
    prev_fresh = fresh(first_frame - 1) if first_frame > 0 else None
    for i in range(n_frames):
        f = first_frame + i
        g = gen(f, 2)
        cur_d, cur_pix, cur_disp = fresh(f)
        if prev_fresh is not None:
            old_d = flip(prev_fresh[0], g)
            old_pix, old_disp = carry(prev_fresh[1], prev_fresh[2], f, g)
        else:
            old_d = rand_desc(cur_d.shape[0], g)
            old_pix = torch.rand((cur_d.shape[0], 2), device=device, generator=g, dtype=torch.float64) * \
                torch.tensor([1200.0, 365.0], device=device, dtype=torch.float64) + \
                torch.tensor([20.0, 5.0], device=device, dtype=torch.float64)
            old_disp = torch.rand((cur_d.shape[0],), device=device, generator=g, dtype=torch.float64) * 117.5 + 2.5
        perm = torch.randperm(int(sizes[i]), device=device, generator=g)
        left = torch.cat([old_d, cur_d])[perm]
        pl = torch.cat([old_pix, cur_pix])[perm].to(torch.float32)
        disp = torch.cat([old_disp, cur_disp])[perm]
        n = left.shape[0]
        assert n == int(sizes[i])
        n_match = int(0.6 * n)
        src = torch.randperm(n, device=device, generator=g)[:n_match]
        dst = torch.randperm(n, device=device, generator=g)[:n_match]
        right = rand_desc(n, g)
        right[dst] = flip(left[src], g)
        n_dup = max(1, n // 100)
        a = torch.randint(0, n, (n_dup,), device=device, generator=g)
        b = torch.randint(0, n, (n_dup,), device=device, generator=g)
        right[a] = right[b]
        pr = torch.rand((n, 2), device=device, generator=g) * torch.tensor([1200.0, 365.0], device=device) \
            + torch.tensor([20.0, 5.0], device=device)
        d = disp[src].to(torch.float32)
        bad = torch.rand((n_match,), device=device, generator=g) < 0.15
        xr = pl[src, 0] - torch.where(bad, -torch.ones_like(d), d)
        yr = pl[src, 1] + 0.5 * torch.randn((n_match,), device=device, generator=g)
        pr[dst, 0] = xr
        pr[dst, 1] = yr
        o = int(off[i])
        desc_l[o:o + n] = left
        desc_r[o:o + n] = right
        pts_l[o:o + n] = pl
        pts_r[o:o + n] = pr
        prev_fresh = (cur_d, cur_pix, cur_disp)
    off32 = off.astype(np.int32)
    return {"desc_l": desc_l, "desc_r": desc_r, "pts_l": pts_l, "pts_r": pts_r,
            "l_off": off32, "r_off": off32.copy(), "n_l": sizes, "n_r": sizes.copy()}
