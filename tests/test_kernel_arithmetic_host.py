"""CPU fuzz of the RANSAC scorer's arithmetic (csrc/ransac_core.cuh, compiled for the host by
oracle/ransac_host_shim.cpp) against the reference formula of transformation_agreement
(final_project/algorithms/ransac.py:28-56, restated in numpy by the oracle).

The kernel decides |num/den - pix| < 2 without dividing and certifies the verdict with a rounding
bound; uncertified cases take the literal IEEE division.  What must hold, bit for bit:
  * agrees() == the reference verdict, everywhere;
  * wherever the certificate fires, the division-free verdict alone already equals the reference
    (soundness of the certificate) — checked on millions of cases placed within a few ulps of the
    +-2 threshold, over many magnitudes, behind the camera, at z = 0, with NaN / inf."""
import numpy as np
import pytest


def _ref(oracle, T, pts, lp, rp, K, M1, M2):
    with np.errstate(all="ignore"):
        return oracle.transformation_agreement(T, pts, lp, rp, K, M1, M2)


def _project(T, pts, K, M1, M2):
    X = np.hstack([pts, np.ones((len(pts), 1))]).T
    with np.errstate(all="ignore"):
        l = ((K @ T) @ np.vstack([M1, [0, 0, 0, 1]])) @ X
        r = ((K @ T) @ np.vstack([M2, [0, 0, 0, 1]])) @ X
        return (l[:2] / l[2]).T, (r[:2] / r[2]).T


def _nudge(v, ulps):
    out = v.copy()
    for _ in range(int(np.abs(ulps).max())):
        up = ulps > 0
        dn = ulps < 0
        out = np.where(up, np.nextafter(out, np.inf), np.where(dn, np.nextafter(out, -np.inf), out))
        ulps = ulps - np.sign(ulps)
    return out


def test_hypothesis_matrices_match_numpy_bits(oracle):
    from slamfe import synth
    host = oracle.ScorerHost()
    K, M1, M2 = synth.cameras()
    rng = np.random.default_rng(400)
    for _ in range(200):
        T = np.hstack([synth._rodrigues(rng.normal(0, 0.5, 3)), rng.normal(0, 5, (3, 1))])
        PL, PR = host.matrices(T, K, M1, M2)
        assert np.array_equal(PL, (K @ T) @ np.vstack([M1, [0, 0, 0, 1]]))
        assert np.array_equal(PR, (K @ T) @ np.vstack([M2, [0, 0, 0, 1]]))


@pytest.mark.parametrize("scale", [1.0, 1e-4, 1e5])
def test_certified_scoring_equals_reference_on_borderline_cases(oracle, scale):
    from slamfe import synth
    host = oracle.ScorerHost()
    K, M1, M2 = synth.cameras()
    rng = np.random.default_rng(401)
    total = uncert = 0
    for rep in range(6):
        Ts, pts, _, _ = synth.pnp_problem(rng, 60000, 2)
        pts = pts * scale
        pts[::7, 2] *= -1.0                                   # behind the camera: no cheirality test
        T = Ts[rep % 2].copy()
        T[:, 3] *= scale
        uvl, uvr = _project(T, pts, K, M1, M2)
        lp, rp = uvl.copy(), uvr.copy()
        n = len(pts)
        which = rng.integers(0, 5, n)                         # 4 = all four coordinates at once
        sign = rng.choice([-2.0, 2.0], (n, 4))
        ulps = rng.integers(-6, 7, (n, 4))
        for col, (arr, c) in enumerate(((lp, 0), (lp, 1), (rp, 0), (rp, 1))):
            m = (which == col) | (which == 4)
            arr[m, c] = _nudge(arr[m, c] + sign[m, col], ulps[m, col])
        a, e, f, cert, shared = host.score(T, pts, lp, rp, K, M1, M2)
        ref = _ref(oracle, T, pts, lp, rp, K, M1, M2)
        assert np.array_equal(a, ref) and np.array_equal(e, ref)
        assert np.array_equal(f[cert], ref[cert])            # the certificate is sound
        assert (shared == 1).all()                           # KITTI rig: the shared-accumulator shortcut is exact
        total += n
        uncert += int((~cert).sum())
    assert uncert > 0.2 * total                              # the borderline cases really hit the fallback
    assert 0.2 < ref.mean() < 0.8                            # ... and the verdicts are genuinely split


def test_certified_scoring_random_magnitudes_and_degenerates(oracle):
    from slamfe import synth
    host = oracle.ScorerHost()
    K, M1, M2 = synth.cameras()
    rng = np.random.default_rng(402)
    for rep in range(8):
        n = 50000
        pts = rng.normal(0, 1, (n, 3)) * 10 ** rng.uniform(-6, 6, (n, 1))
        T = np.hstack([synth._rodrigues(rng.normal(0, 1.0, 3)), (rng.normal(0, 1, 3) * 10 ** rng.uniform(-3, 3))[:, None]])
        uvl, uvr = _project(T, pts, K, M1, M2)
        jit = 10 ** rng.uniform(-14, 2, (n, 1)) * rng.normal(0, 1, (n, 2))
        lp = np.nan_to_num(uvl, nan=0.0, posinf=1e300, neginf=-1e300) + jit
        rp = np.nan_to_num(uvr, nan=0.0, posinf=1e300, neginf=-1e300) + jit[:, ::-1]
        pts[:5] = [[0, 0, 0], [1, 2, 0], [np.nan, 1, 5], [np.inf, 1, 5], [1e308, 1e308, 1e308]]
        lp[5:10, 0] = [np.nan, np.inf, -np.inf, 1e308, -1e308]
        a, e, f, cert, shared = host.score(T, pts, lp, rp, K, M1, M2)
        ref = _ref(oracle, T, pts, lp, rp, K, M1, M2)
        assert np.array_equal(a, ref) and np.array_equal(e, ref)
        assert np.array_equal(f[cert], ref[cert])
        assert (shared == 1).all()
    # a non-rectified rig (rotated right camera) does not qualify for the shortcut and stays exact
    M2r = np.hstack([synth._rodrigues(np.array([0.0, 0.02, 0.0])), M2[:, 3:4]])
    Ts, pts, lp, rp = synth.pnp_problem(rng, 20000, 1)
    a, e, f, cert, shared = host.score(Ts[0], pts, lp, rp, K, M1, M2r)
    assert (shared == 2).all() and np.array_equal(a, _ref(oracle, Ts[0], pts, lp, rp, K, M1, M2r))


def test_triangulation_core_vs_svd_on_extreme_links(oracle):
    """csrc/triangulate_core.cuh (host build) against the reference's per-link np.linalg.svd
    (triangulation.py:5-24 via the oracle): KITTI cameras with disparities down to the filter's limit
    (xl > xr + 2, matching.py:63) and up to the image width, and random shared-row stereo pairs."""
    from slamfe import synth, utils
    rng = np.random.default_rng(403)
    n = 4000
    xl = rng.uniform(3, 1240, n).astype(np.float32).astype(np.float64)
    d = np.concatenate([rng.uniform(2.0001, 2.5, n // 4), rng.uniform(2.5, 120, n // 2), rng.uniform(120, 1200, n // 4)])
    links = np.stack([xl, (xl - d).astype(np.float32).astype(np.float64), rng.uniform(0, 376, n)], axis=1)
    links = links[links[:, 0] > links[:, 1] + 2]
    got = oracle.triangulate_links_host_build(links, utils.P, utils.Q)
    ref = oracle.triangulate_links(links, utils.P, utils.Q)
    rel = np.linalg.norm(got - ref, axis=1) / np.linalg.norm(ref, axis=1)
    assert rel.max() < 1e-10, rel.max()
    assert (got[:, 2] > 0).all()
    for _ in range(20):  # other rectified rigs: rows 1 and 2 shared, row 0 differs
        fx, fy = rng.uniform(300, 1500, 2)
        cx, cy = rng.uniform(200, 900), rng.uniform(100, 500)
        b = rng.uniform(0.05, 1.5)
        P = np.array([[fx, rng.normal(0, 0.5), cx, 0.0], [0, fy, cy, 0.0], [0, 0, 1, 0.0]])
        Q = P.copy()
        Q[0, 3] = -fx * b
        lk = synth.links(rng, 300)
        got = oracle.triangulate_links_host_build(lk, P, Q)
        ref = oracle.triangulate_links(lk, P, Q)
        rel = np.linalg.norm(got - ref, axis=1) / np.linalg.norm(ref, axis=1)
        assert rel.max() < 1e-10, rel.max()


@pytest.mark.parametrize("cs", [7, 8, 9, 10])
def test_carry_save_distance_equals_popcount(oracle, cs):
    """csrc/hamming_core.cuh (host build): prefix form + full-adder chain + weighted popcount
    accumulation == popcount(q XOR t), for every single-bit and adjacent/arbitrary double-bit pattern of
    a 64-byte row, the extremes, random rows and every descriptor width 1..64 (zero padded)."""
    rng = np.random.default_rng(500 + cs)
    bits = np.arange(512)
    single = np.zeros((512, 64), np.uint8)
    single[bits, bits // 8] = 1 << (bits % 8)
    zero = np.zeros_like(single)
    assert np.array_equal(oracle.hamming_keys_host_build(single, zero, cs=cs) >> 22, np.ones(512))
    assert np.array_equal(oracle.hamming_keys_host_build(zero, single, cs=cs) >> 22, np.ones(512))
    i, j = np.triu_indices(512, k=1)                       # all 130 816 two-bit patterns
    double = np.zeros((len(i), 64), np.uint8)
    np.bitwise_or.at(double, (np.arange(len(i)), i // 8), (1 << (i % 8)).astype(np.uint8))
    np.bitwise_or.at(double, (np.arange(len(j)), j // 8), (1 << (j % 8)).astype(np.uint8))
    base = rng.integers(0, 256, (len(i), 64), dtype=np.uint8)
    assert np.array_equal(oracle.hamming_keys_host_build(base, base ^ double, cs=cs) >> 22, np.full(len(i), 2))
    ones = np.full((3, 64), 255, np.uint8)
    assert (oracle.hamming_keys_host_build(ones, np.zeros((3, 64), np.uint8), cs=cs) >> 22).tolist() == [512] * 3
    assert (oracle.hamming_keys_host_build(ones, ones, cs=cs)).tolist() == [0] * 3
    popc = np.array([bin(v).count("1") for v in range(256)], np.int64)
    for width in list(range(1, 65)):
        q = rng.integers(0, 256, (400, 64), dtype=np.uint8)
        t = rng.integers(0, 256, (400, 64), dtype=np.uint8)
        k = oracle.hamming_keys_host_build(q, t, desc_bytes=width, cs=cs)
        assert (k & 0x3FFFFF == 0).all()
        assert np.array_equal(k >> 22, popc[q[:, :width] ^ t[:, :width]].sum(axis=1))
    # correlated rows (true matches: few differing bits) and heavy rows
    q = rng.integers(0, 256, (20000, 64), dtype=np.uint8)
    flips = (rng.random((20000, 64, 8)) < rng.uniform(0, 1, (20000, 1, 1))).astype(np.uint8)
    t = q ^ np.packbits(flips, axis=2, bitorder="little")[:, :, 0]
    assert np.array_equal(oracle.hamming_keys_host_build(q, t, cs=cs) >> 22, popc[q ^ t].sum(axis=1))


# ---------------------------------------------------------------------------------------------------
# PnP refit (csrc/refit_core.cuh, host build): ransac.py:185-193
# ---------------------------------------------------------------------------------------------------
def _refit_problem(rng, n, noise=0.5, seed_err=1.0):
    from slamfe import synth
    K, _, _ = synth.cameras()
    z = rng.uniform(4, 60, n)
    X = np.stack([(rng.uniform(20, 1220, n) - K[0, 2]) * z / K[0, 0], (rng.uniform(5, 370, n) - K[1, 2]) * z / K[1, 1], z], 1)
    rvec, tvec = rng.normal(0, 0.02, 3), np.array([0.0, 0.0, -0.9]) + rng.normal(0, 0.05, 3)
    T_gt = np.hstack([synth._rodrigues(rvec), tvec[:, None]])
    q = (X @ T_gt[:, :3].T + T_gt[:, 3]) @ K.T
    pix = q[:, :2] / q[:, 2:3] + rng.normal(0, noise, (n, 2))
    T_seed = np.hstack([synth._rodrigues(rvec + rng.normal(0, 0.003 * seed_err, 3)),
                        (tvec + rng.normal(0, 0.05 * seed_err, 3))[:, None]])
    return K, X, pix, T_gt, T_seed


def _pose_err(Ta, Tb):
    dR = Ta[:, :3] @ Tb[:, :3].T
    ang = np.arcsin(min(1.0, np.linalg.norm(dR - dR.T) / (2 * np.sqrt(2))))   # accurate near 0, unlike arccos
    return ang, np.linalg.norm(Ta[:, 3] - Tb[:, 3])


def test_refit_host_build_equals_numpy_gauss_newton_and_beats_the_seed():
    """The product's LM refit (host build of csrc/refit_core.cuh) lands on the same minimum as an independent
    numpy Gauss-Newton of the same objective (1e-9), the minimum is closer to the true pose than the
    seed, and with noise-free pixels it IS the true pose."""
    from oracle import ref_oracle as ora
    rng = np.random.default_rng(91)
    for n, noise in [(12, 0.5), (200, 0.5), (1500, 0.5), (300, 0.0), (5, 0.2)]:
        K, X, pix, T_gt, T_seed = _refit_problem(rng, n, noise)
        T, status, rms = ora.refit_host_build(T_seed, K, X, pix)
        assert status > 0, status
        T_np, rms_np = ora.pnp_refit(T_seed, K, X, pix)
        ang, dt = _pose_err(T, T_np)
        assert ang < 1e-9 and dt < 1e-8, (n, ang, dt)
        assert abs(rms - rms_np) < 1e-9
        assert np.allclose(T[:, :3] @ T[:, :3].T, np.eye(3), atol=1e-12)
        if noise == 0.0:
            ang, dt = _pose_err(T, T_gt)
            assert ang < 1e-9 and dt < 1e-8 and rms < 1e-8
        elif n >= 200:
            a0, d0 = _pose_err(T_seed, T_gt)
            a1, d1 = _pose_err(T, T_gt)
            assert a1 < a0 and d1 < d0 and d1 < 0.01 and a1 < 1e-3, (n, a0, d0, a1, d1)


def test_refit_host_build_against_cv2_solvepnp():
    """Contract against what the reference calls (cv2.solvePnP EPNP on the consensus set,
    ransac.py:190): same pose within rotation 1e-3 rad / translation 1 cm on consensus-sized problems,
    and never a larger reprojection error; the mask restricts the sum to the inliers."""
    import cv2
    from oracle import ref_oracle as ora
    rng = np.random.default_rng(92)
    for n in (150, 600, 2000):
        K, X, pix, T_gt, T_seed = _refit_problem(rng, n)
        mask = (rng.random(n) < 0.7).astype(np.uint8)
        mask[:4] = 1
        garbage = pix.copy()
        garbage[mask == 0] += 500.0     # outliers must not matter
        T, status, rms = ora.refit_host_build(T_seed, K, X, garbage, mask=mask)
        assert status > 0
        ok, rvec, tvec = cv2.solvePnP(X[mask == 1], pix[mask == 1], K, np.zeros((5, 1)), flags=cv2.SOLVEPNP_EPNP)
        assert ok
        T_cv = ora.rodriguez_to_mat(rvec, tvec)
        ang, dt = _pose_err(T, T_cv)
        assert ang < 1e-3 and dt < 0.01, (n, ang, dt)
        q = (X[mask == 1] @ T_cv[:, :3].T + T_cv[:, 3]) @ K.T
        rms_cv = np.sqrt((((q[:, :2] / q[:, 2:3]) - pix[mask == 1]) ** 2).sum() / mask.sum())
        assert rms <= rms_cv + 1e-9


def test_refit_host_build_degenerate_inputs():
    from oracle import ref_oracle as ora
    rng = np.random.default_rng(93)
    K, X, pix, T_gt, T_seed = _refit_problem(rng, 50)
    T, status, _ = ora.refit_host_build(T_seed, K, X, pix, mask=np.r_[np.ones(3, np.uint8), np.zeros(47, np.uint8)])
    assert status == 0                      # fewer than 4 inliers: no pose (ransac.py:187-188)
    Xc = np.tile(X[:1], (50, 1))            # all points coincide: singular normal equations, seed returned
    T, status, _ = ora.refit_host_build(T_seed, K, Xc, np.tile(pix[:1], (50, 1)))
    assert status != 0 and np.isfinite(T).all()
    far = T_seed.copy(); far[:, 3] += [0.5, -0.3, 0.8]    # a poor seed still converges (LM damping)
    T, status, _ = ora.refit_host_build(far, K, X, pix, max_iter=50)
    T_np, _ = ora.pnp_refit(T_seed, K, X, pix)
    ang, dt = _pose_err(T, T_np)
    assert status > 0 and ang < 1e-7 and dt < 1e-6


def _prefilter_case(oracle, host, T, pts, lp, rp, K, M1, M2):
    """Runs the host build of the fp32 pre-filter + agrees_rows for one hypothesis and checks what the kernel
    relies on: a dropped pair is never one the reference accepts, and for every pair the kernel's count
    (kept && agrees_rows) equals the reference's verdict.  Returns (dropped, reference verdict)."""
    with np.errstate(all="ignore"):
        far_v, far_u, rows, exact = host.prune(T, pts, lp, rp, K, M1, M2)
    ref = _ref(oracle, T, pts, lp, rp, K, M1, M2)
    dropped = far_v | far_u
    assert np.array_equal(exact, ref)
    assert not (dropped & ref).any()
    assert np.array_equal(~dropped & rows, ref)
    return dropped, ref


@pytest.mark.parametrize("scale", [1.0, 1e-4, 1e5])
def test_fp32_prefilter_never_drops_an_inlier_near_its_own_threshold(oracle, scale):
    """The scorer's fp32 pre-filter (prune_* in csrc/ransac_core.cuh, derivation there) against the reference's
    transformation_agreement (ransac.py:28-56): pixels placed around the reference's threshold (2 px), around
    the filter's own (2.001 px +- its slack) and far away, at three scene scales, points behind the camera."""
    from slamfe import synth
    host = oracle.ScorerHost()
    K, M1, M2 = synth.cameras()
    rng = np.random.default_rng(98)
    n_drop = n_in = total = 0
    for rep in range(6):
        Ts, pts, _, _ = synth.pnp_problem(rng, 60000, 2)
        pts = pts * scale
        pts[::7, 2] *= -1.0
        T = Ts[rep % 2].copy()
        T[:, 3] *= scale
        uvl, uvr = _project(T, pts, K, M1, M2)
        n = len(pts)
        off = np.concatenate([rng.choice([-2.0, 2.0], n // 3) * (1 + rng.integers(-8, 9, n // 3) * 2.0 ** -52),
                              rng.choice([-1.0, 1.0], n // 3) * rng.uniform(1.99, 2.02, n // 3),
                              rng.normal(0, 30, n - 2 * (n // 3))])
        coord = rng.integers(0, 2, n)
        lp = uvl + rng.uniform(-1.5, 1.5, (n, 2))
        lp[np.arange(n), coord] = uvl[np.arange(n), coord] + off
        rp = uvr + rng.uniform(-1.5, 1.5, (n, 2))
        dropped, ref = _prefilter_case(oracle, host, T, pts, lp, rp, K, M1, M2)
        n_drop += int(dropped.sum()); n_in += int(ref.sum()); total += n
        # everything clearly outside on the left image is dropped (the filter does its job)
        with np.errstate(all="ignore"):
            clear = (np.abs(lp - uvl).max(axis=1) > 2.1) & np.isfinite(uvl).all(axis=1)
        # (a scene shrunk to millimetres keeps the homogeneous "+ 1" of the bound B: conservative, less sharp; one
        # blown up to 10^5 km trips the 10^9 range guard and is not filtered at all)
        if scale <= 1.0:
            assert dropped[clear].mean() > (0.999 if scale == 1.0 else 0.9)
    assert (n_drop > 0.25 * total or scale > 1.0) and n_in > 0.1 * total


def test_fp32_prefilter_random_magnitudes_and_degenerates(oracle):
    """Magnitudes over 24 decades (fp32 underflow and the 1e9 range guards on both sides), z = 0, NaN, inf,
    1e308, hypotheses with huge / tiny / NaN translations, a non-rectified rig."""
    from slamfe import synth
    host = oracle.ScorerHost()
    K, M1, M2 = synth.cameras()
    rng = np.random.default_rng(99)
    kept_any = dropped_any = 0
    for rep in range(10):
        n = 50000
        pts = rng.normal(0, 1, (n, 3)) * 10 ** rng.uniform(-12, 12, (n, 1))
        tscale = [1.0, 1e-3, 1e3, 1e-30, 1e12, 1e-45, 1e6, 1.0, 1.0, 1.0][rep]
        T = np.hstack([synth._rodrigues(rng.normal(0, 1.0, 3)), (rng.normal(0, 1, 3) * tscale)[:, None]])
        if rep == 7:
            T[1, 3] = np.nan
        if rep == 8:
            T[0, 3] = np.inf
        uvl, uvr = _project(T, pts, K, M1, M2)
        jit = 10 ** rng.uniform(-14, 2, (n, 1)) * rng.normal(0, 1, (n, 2))
        lp = np.nan_to_num(uvl, nan=0.0, posinf=1e300, neginf=-1e300) + jit
        rp = np.nan_to_num(uvr, nan=0.0, posinf=1e300, neginf=-1e300) + jit[:, ::-1]
        pts[:6] = [[0, 0, 0], [1, 2, 0], [np.nan, 1, 5], [np.inf, 1, 5], [1e308, 1e308, 1e308], [1e-320, 1e-320, 1e-320]]
        lp[6:13, 0] = [np.nan, np.inf, -np.inf, 1e308, -1e308, 1e-320, 3e9]
        lp[13:15, 1] = [np.nan, 2e9]
        dropped, ref = _prefilter_case(oracle, host, T, pts, lp, rp, K, M1, M2)
        kept_any += int(ref.sum()); dropped_any += int(dropped.sum())
    assert kept_any > 1000 and dropped_any > 1000
    M2r = np.hstack([synth._rodrigues(np.array([0.0, 0.02, 0.0])), M2[:, 3:4]])
    Ts, pts, lp, rp = synth.pnp_problem(rng, 20000, 1)
    _prefilter_case(oracle, host, Ts[0], pts, lp, rp, K, M1, M2r)
