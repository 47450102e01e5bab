"""Host-side logic of the mirror modules (no GPU): error behaviour, key decoding, sharding."""
import numpy as np
import pytest


def test_calc_ransac_iteration_table(slamfe, golden):
    from slamfe import ransac
    for p, n in golden("ransac")["iters"]:
        assert ransac.calc_ransac_iteration(int(p)) == int(n)
    assert ransac.calc_ransac_iteration(40) == 888  # loop_closure.py:425


def test_cameras_match_reference(slamfe, golden):
    from slamfe import ransac, utils
    g = golden("ransac")
    assert np.array_equal(utils.K, g["K"]) and np.array_equal(utils.M1, g["M1"]) and np.array_equal(utils.M2, g["M2"])
    assert np.array_equal(ransac.P, g["K"] @ g["M1"]) and np.array_equal(ransac.Q, g["K"] @ g["M2"])


def test_key_decode(slamfe):
    from slamfe import ops
    keys = np.array([(37 << 22) | 5, 0xFFFFFFFF, (488 << 22) | 4194302, 0], dtype=np.uint32)
    idx, dist = ops.keys_to_numpy(keys.view(np.int32))
    assert idx.tolist() == [5, -1, 4194302, 0] and dist.tolist() == [37, -1, 488, 0]


def test_ratio_mask(slamfe, golden):
    from slamfe import matching
    g = golden("matching")
    for name in ("akaze", "ties", "one_train"):
        assert np.array_equal(matching.ratio_test_mask(g[f"{name}_knn_dist"]), g[f"{name}_ratio"])


def test_matcher_error_behaviour_matches_cv2(slamfe):
    """cv2 4.13 behaviour: empty / None -> (), wrong dtype or width -> cv2.error,
    all decided before the GPU is touched."""
    import cv2
    from slamfe import matching
    m = matching.Matcher(crossCheck=False)
    ref = cv2.BFMatcher(cv2.NORM_HAMMING)
    a = np.zeros((4, 61), np.uint8)
    assert m.match(np.zeros((0, 61), np.uint8), a) == () == tuple(ref.match(np.zeros((0, 61), np.uint8), a))
    assert m.match(a, np.zeros((0, 61), np.uint8)) == ()
    assert m.match(None, a) == () and m.knnMatch(a, None, k=2) == ()
    with pytest.raises(cv2.error):
        m.match(a.astype(np.float32), a)
    with pytest.raises(cv2.error):
        ref.match(a.astype(np.float32), a)
    with pytest.raises(cv2.error):
        m.match(a, np.zeros((4, 32), np.uint8))
    with pytest.raises(cv2.error):
        ref.match(a, np.zeros((4, 32), np.uint8))
    with pytest.raises(ValueError):
        matching.Matcher(normType=cv2.NORM_L2)


def test_no_cpu_fallback(slamfe):
    """Without a CUDA device the product path must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from slamfe import matching, ransac, triangulation
    a = np.zeros((4, 61), np.uint8)
    with pytest.raises(slamfe.SlamfeError):
        matching.Matcher().match(a, a)
    with pytest.raises(slamfe.SlamfeError):
        triangulation.triangulate_links(np.ones((3, 3)), ransac.P, ransac.Q)
    with pytest.raises(slamfe.SlamfeError):
        ransac.transformation_agreement(np.eye(3, 4), np.ones((3, 3)), np.ones((3, 2)), np.ones((3, 2)))
    assert matching.extract_inliers_outliers((), (), ())[0].size == 0


def test_product_does_not_import_oracle(slamfe):
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/: not the package,
    not the C sources, not the development scripts."""
    import os
    import re
    pkg = slamfe.__path__[0]
    root = os.path.dirname(pkg)
    files = [os.path.join(pkg, fn) for fn in os.listdir(pkg) if fn.endswith(".py")]
    files += [os.path.join(pkg, "csrc", fn) for fn in os.listdir(os.path.join(pkg, "csrc"))]
    files += [os.path.join(root, "scripts", fn) for fn in os.listdir(os.path.join(root, "scripts")) if fn.endswith(".py")]
    files += [os.path.join(root, "slamfe.py")]
    for path in files:
        src = open(path).read()
        assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), path
        assert "liboracle" not in src and "libp3p_host" not in src, path
    # and the library itself must not depend on the oracle's shared objects
    import subprocess
    deps = subprocess.run(["ldd", os.path.join(pkg, "libslamfe.so")], capture_output=True, text=True).stdout
    assert "oracle" not in deps and "p3p_host" not in deps


def test_offsets_are_tma_aligned(slamfe):
    from slamfe import frontend
    off = frontend.plan_offsets([3000, 17, 0, 4999])
    assert off.tolist() == [0, 3008, 3040, 3040, 8048]
    assert all((int(o) * 61) % 16 == 0 for o in off)


def test_balanced_ranges_and_slices(slamfe):
    from slamfe import dist
    b = dist.balanced_ranges(np.ones(10), 4)
    assert b[0] == 0 and b[-1] == 10 and np.all(np.diff(b) >= 2)
    b = dist.balanced_ranges([1, 1, 1, 1, 10, 1, 1, 1], 3)
    assert b.tolist() == [0, 4, 5, 8]
    assert dist.balanced_ranges([], 4).tolist() == [0, 0, 0, 0, 0]
    s = dist.train_slices(20000, 8)
    assert s[0] == 0 and s[-1] == 20000 and all(int(x) % 16 == 0 for x in s[:-1])
    pairs = dist.candidate_pairs(5)
    assert len(pairs) == 10 and np.all(pairs[:, 0] > pairs[:, 1])
    blocks = dist.candidate_blocks(pairs, np.array([100, 200, 300, 400, 500]), 2)
    w = (np.array([100, 200, 300, 400, 500])[pairs[:, 0]] * np.array([100, 200, 300, 400, 500])[pairs[:, 1]])
    assert abs(w[: blocks[1]].sum() - w[blocks[1]:].sum()) <= w.max()


def test_merge_top2_host(slamfe):
    from slamfe import dist
    rng = np.random.default_rng(0)
    keys = rng.integers(0, 2 ** 31, (4, 50, 2), dtype=np.int64).astype(np.uint32)
    keys.sort(axis=2)
    merged = dist.merge_top2_host(keys)
    flat = np.sort(keys.transpose(1, 0, 2).reshape(50, 8), axis=1)
    assert np.array_equal(merged, flat[:, :2])


def test_chunk_schedule_of_the_host_pipeline(slamfe):
    from slamfe import frontend
    for F in (0, 1, 2, 7, 100, 599, 600, 2000, 4541, 20000):
        for c in (1, 2, 16, 100, 576, 5000):
            b = frontend.chunk_bounds(F, c)
            assert b[0] == 0 and b[-1] == F if F else b == [0]
            sizes = np.diff(b)
            assert (sizes > 0).all() and sizes.max(initial=0) <= max(c, 1)
    sizes = np.diff(frontend.chunk_bounds(4541, 576))
    assert sizes[0] == 72 and sizes[-1] == 72 and (sizes == 576).sum() >= 4   # ramp up, full speed, ramp down


def test_sharding_properties(slamfe):
    """Property tests of the host-side sharding arithmetic (slamfe.dist)."""
    from hypothesis import given, settings, strategies as st
    from slamfe import dist

    @settings(max_examples=200, deadline=None)
    @given(st.lists(st.integers(0, 10_000), min_size=0, max_size=60), st.integers(1, 9))
    def ranges(work, world):
        b = dist.balanced_ranges(work, world)
        assert len(b) == world + 1 and b[0] == 0 and b[-1] == len(work)
        assert all(x <= y for x, y in zip(b, b[1:]))
        if sum(work) > 0 and len(work) >= world:
            shares = [sum(work[b[r]:b[r + 1]]) for r in range(world)]
            assert max(shares) <= sum(work) / world + max(work)      # never worse than one unit off

    @settings(max_examples=100, deadline=None)
    @given(st.integers(0, 100_000), st.integers(1, 9))
    def slices(n, world):
        b = dist.train_slices(n, world)
        assert b[0] == 0 and b[-1] == n and all(x <= y for x, y in zip(b, b[1:]))
        assert all(int(x) % 16 == 0 for x in b[:-1] if x < n)      # every non-empty slice starts 16-row aligned

    @settings(max_examples=100, deadline=None)
    @given(st.integers(1, 6), st.integers(1, 40), st.integers(0, 2 ** 32 - 1))
    def merge(shards, nq, seed):
        rng = np.random.default_rng(seed)
        keys = rng.integers(0, 2 ** 32, (shards, nq, 2), dtype=np.uint64).astype(np.uint32)
        keys.sort(axis=2)
        got = dist.merge_top2_host(keys)
        want = np.sort(keys.transpose(1, 0, 2).reshape(nq, -1), axis=1)[:, :2]
        assert np.array_equal(got, want)

    ranges(); slices(); merge()
    pairs = dist.candidate_pairs(7)
    assert len(pairs) == 21 and (pairs[:, 0] > pairs[:, 1]).all() and pairs[0].tolist() == [1, 0]
