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
