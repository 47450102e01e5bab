"""Batched loop-closure candidate verification (slamfe.loop) against the oracle's restatement of
check_candidate_match (loop_closure.py:405-436): match tables and RANSAC inputs bit-exact, scoring
bit-exact on the device-generated hypotheses, planted revisits accepted."""
import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("matcher_kernel")]


def _keyframes(rng, sizes, revisits):
    """Keyframe pool: random descriptors and valid stereo links; revisits {a: b} make keyframe a a
    re-observation of keyframe b (flipped descriptors, links = b's 3-D points under a small motion)."""
    from slamfe import synth, utils
    K = utils.K
    fx, cx, cy, fxb = K[0, 0], K[0, 2], K[1, 2], -utils.KITTI00_P1[0, 3]
    desc, links = [], []
    for n in sizes:
        desc.append(synth.descriptors(rng, n))
        xl = rng.uniform(20, 1220, n).astype(np.float32).astype(np.float64)
        d = rng.uniform(2.5, 120, n)
        xr = (xl - d).astype(np.float32).astype(np.float64)
        y = rng.uniform(5, 370, n).astype(np.float32).astype(np.float64)
        links.append(np.stack([xl, xr, y], axis=1))
    for a, b in revisits.items():
        n = min(sizes[a], sizes[b])
        perm = rng.permutation(sizes[b])[:n]
        desc[a][:n] = synth.flip_bits(rng, desc[b][perm], 0.06)
        lb = links[b][perm]
        Z = fxb / (lb[:, 0] - lb[:, 1])
        P = np.stack([(lb[:, 0] - cx) * Z / fx, (lb[:, 2] - cy) * Z / fx, Z], axis=1)
        R = synth._rodrigues(rng.normal(0, 0.02, 3))
        t = np.array([0.3, -0.05, -0.6])
        # keyframe a sees the points at P; keyframe b (the candidate) sees them at R P + t: build a's links
        # from the inverse so that T = [R|t] maps a's triangulated points to b's camera
        Pa = (P - t) @ R          # R^T (P - t)
        ok = Pa[:, 2] > 2.0
        xl = fx * Pa[:, 0] / Pa[:, 2] + cx + rng.normal(0, 0.3, n)
        y = fx * Pa[:, 1] / Pa[:, 2] + cy + rng.normal(0, 0.3, n)
        xr = xl - fxb / Pa[:, 2]
        la = np.stack([xl, xr, y], axis=1).astype(np.float32).astype(np.float64)
        links[a][:n][ok] = la[ok]
    return desc, links


def test_candidate_verification_vs_oracle(slamfe, oracle):
    import torch
    from slamfe import dist, loop, utils
    rng = np.random.default_rng(81)
    sizes = [300, 280, 310, 5, 290, 300]
    desc, links = _keyframes(rng, sizes, {4: 1, 5: 0})
    off = np.zeros(len(sizes) + 1, np.int64)
    np.cumsum([-(-n // 16) * 16 for n in sizes], out=off[1:])
    pool_d = np.zeros((off[-1], 61), np.uint8)
    pool_l = np.ones((off[-1], 3), np.float64) * [30.0, 10.0, 50.0]
    for k, n in enumerate(sizes):
        pool_d[off[k]:off[k] + n] = desc[k]
        pool_l[off[k]:off[k] + n] = links[k]
    pairs = dist.candidate_pairs(len(sizes))
    H = 96
    ver = loop.CandidateVerifier(block_pairs=4)          # several blocks
    res = ver.verify(torch.from_numpy(pool_d).cuda(), torch.from_numpy(pool_l).cuda(), off[:-1], np.array(sizes), pairs,
                     n_iter=H, seed=3, want_masks=True)
    K, M1, M2 = utils.K, utils.M1, utils.M2
    for p, (a, b) in enumerate(pairs):
        oi, od = oracle.match(desc[a], desc[b])
        k = res["keys"][p]
        assert np.array_equal(k & 0x3FFFFF, oi.astype(np.uint32)) and np.array_equal(k >> 22, od.astype(np.uint32))
        assert res["n_matches"][p] == sizes[a]
    # scoring on the device-generated hypotheses of the LAST block, bit-exact against the oracle
    nb = len(pairs) % 4 or 4
    last = pairs[-nb:]
    T = ver._buf["T"][:nb * H].cpu().numpy(); valid = ver._buf["hyp_valid"][:nb * H].cpu().numpy().astype(bool)
    for q, (a, b) in enumerate(last):
        p = len(pairs) - nb + q
        oi, _ = oracle.match(desc[a], desc[b])
        pts = oracle.triangulate_links(links[a], utils.P, utils.Q)
        cur = links[b][oi]
        lp, rp = cur[:, [0, 2]], cur[:, [1, 2]]
        counts = np.zeros(H, np.int64); masks = {}
        for h in np.nonzero(valid[q * H:(q + 1) * H])[0]:
            # the device triangulates with its own fp64 kernel (1e-13 from the SVD): score the oracle on
            # the oracle's points and require the same verdicts except within 1e-9 px of the threshold
            masks[h] = oracle.transformation_agreement(T[q * H + h], pts, lp, rp, K, M1, M2)
            counts[h] = masks[h].sum()
        assert abs(int(counts.max()) - int(res["inliers"][p])) <= 1
        if counts.max() > 0 and counts.max() - np.sort(counts)[-2] > 2:
            assert res["best_hyp"][p] == int(np.argmax(counts))
            assert (res["mask"][p] != masks[int(np.argmax(counts))]).sum() <= 1
    # planted revisits are accepted, unrelated keyframes are not
    acc = {tuple(pr): bool(a) for pr, a in zip(pairs.tolist(), res["accepted"])}
    assert acc[(4, 1)] and acc[(5, 0)]
    assert sum(acc.values()) == 2
    i41 = pairs.tolist().index([4, 1])
    assert res["inliers"][i41] > 150 and 0.5 < res["percentage"][i41] <= 1.0


def test_candidates_against_the_reference_golden(slamfe, golden):
    """tests/golden/loop_candidates.npz: the UNMODIFIED reference's check_candidate_match
    (loop_closure.py:405-436) on keyframes of the TrackingDB its own create_db built
    (tests/golden/create_db.npz).  Match tables must be identical; where the reference's randomised
    888-iteration RANSAC found a consensus, the GPU verifier's consensus contains it; the unrelated
    pair stays far below the acceptance threshold on both sides."""
    import torch
    from slamfe import loop
    db, g = golden("create_db"), golden("loop_candidates")
    n = int(db["n_frames"])
    feats = [db[f"features{f}"] for f in range(n)]
    links = [db[f"links{f}"] for f in range(n)]
    sizes = [len(x) for x in feats]
    off = np.zeros(n + 1, np.int64)
    np.cumsum([-(-s // 16) * 16 for s in sizes], out=off[1:])
    pool_d = np.zeros((off[-1], 61), np.uint8)
    pool_l = np.ones((off[-1], 3)) * [30.0, 10.0, 50.0]
    for f in range(n):
        pool_d[off[f]:off[f] + sizes[f]] = feats[f]
        pool_l[off[f]:off[f] + sizes[f]] = links[f]
    pairs = g["pairs"]
    res = loop.CandidateVerifier().verify(torch.from_numpy(pool_d).cuda(), torch.from_numpy(pool_l).cuda(), off[:-1],
                                          np.array(sizes), pairs, inliers_percent=40, seed=4, want_masks=True)
    found = 0
    for k, (a, b) in enumerate(pairs):
        keys = res["keys"][k]
        assert np.array_equal(keys & 0x3FFFFF, g[f"match_t{k}"].astype(np.uint32))
        assert np.array_equal((keys >> 22).astype(np.float32), g[f"match_d{k}"])
        ref_in = np.zeros(sizes[a], bool)
        ref_in[g[f"inlier_q{k}"]] = True
        got_in = res["mask"][k]
        assert got_in.sum() == res["inliers"][k] and res["n_matches"][k] == sizes[a]
        assert got_in.sum() >= 0.8 * ref_in.sum()
        if ref_in.sum() >= 20:
            found += 1
            assert (got_in & ref_in).sum() / ref_in.sum() >= 0.75
            assert abs(res["percentage"][k] - float(g[f"percentage{k}"])) < 0.1
    assert found >= 2
    unrelated = [k for k, (a, b) in enumerate(pairs) if abs(int(a) - int(b)) > 1]
    assert all(res["inliers"][k] < 15 and len(g[f"inlier_q{k}"]) < 15 for k in unrelated)
    assert not res["accepted"].any()       # tiny keyframes: nobody reaches the 120-inlier threshold


class _L:
    def __init__(self, a, b, c):
        self.x_left, self.x_right, self.y = a, b, c


class _GoldenDB:
    """TrackingDB-shaped view of tests/golden/create_db.npz: features(f) / all_frame_links(f)."""

    def __init__(self, g):
        self.g = g

    def features(self, f):
        return self.g[f"features{f}"]

    def all_frame_links(self, f):
        return [_L(*r) for r in self.g[f"links{f}"]]


def test_loop_closure_dropins_against_the_reference_golden(slamfe, golden, monkeypatch):
    """slamfe.loop.check_candidate_match / consensus_matches (same signatures and returns as
    loop_closure.py:405-436, :572-599) on the reference's own TrackingDB content, against the reference's
    recorded check_candidate_match results."""
    from slamfe import loop
    db, g = _GoldenDB(golden("create_db")), golden("loop_candidates")
    np.random.seed(8)
    for k, (a, b) in enumerate(g["pairs"]):
        ref_q = set(g[f"inlier_q{k}"].tolist())
        matches, pct, pose = loop.check_candidate_match(int(a), int(b), db)
        got_q = {m.queryIdx for m in matches}
        assert all(m.imgIdx == 0 and g[f"match_t{k}"][m.queryIdx] == m.trainIdx and
                   g[f"match_d{k}"][m.queryIdx] == m.distance for m in matches)
        assert len(got_q) >= 0.8 * len(ref_q)
        if len(ref_q) >= 20:
            assert len(got_q & ref_q) / len(ref_q) >= 0.75 and abs(pct - float(g[f"percentage{k}"])) < 0.1
            P_ref, P_got = g[f"pose{k}"], pose.matrix()
            assert np.abs(P_got[:3, 3] - P_ref[:3, 3]).max() < 0.15 and np.abs(P_got[:3, :3] - P_ref[:3, :3]).max() < 0.02
        if abs(int(a) - int(b)) > 1:
            assert len(matches) < 15
    # consensus_matches: first candidate in list order above the threshold wins; none -> (None, [], last rel_T)
    cand, ms, rel = loop.consensus_matches(4, [1, 3], db)
    assert cand is None and ms == []
    monkeypatch.setattr(loop, "INLIERS_THRESHOLD", 30)
    cand, ms, rel = loop.consensus_matches(4, [1, 3], db)
    assert cand == 3 and len(ms) > 30 and rel is not None
    cand2, ms2, _ = loop.consensus_matches(4, [3, 1], db)
    assert cand2 == 3 and abs(len(ms2) - len(ms)) <= 0.2 * len(ms)
    assert loop.consensus_matches(4, [], db) == (None, [], None)


def test_loop_closure_pose_needs_no_host_solve(slamfe, golden, monkeypatch):
    """check_candidate_match's pose comes from slamfe_pnp_refit inside the batch (loop.REFIT = "gpu"); with
    the same seed (same consensus set) it agrees with the reference's recipe, cv2.solvePnP(EPNP) on the
    consensus set (loop.REFIT = "cv2", ransac.py:185-204).  The golden keyframes are small (300-520
    keypoints, consensus sets of 60-150): EPnP (algebraic + Gauss-Newton on its betas) and the
    reprojection-error minimum differ by centimetres there, so the bound is 5 mrad / 5 cm; on
    consensus sets of >= 100 well-spread points the two agree to 1 mrad / 1 cm
    (tests/test_gpu_ransac.py::test_pnp_refit_kernel_equals_host_build_and_cv2)."""
    import cv2
    from slamfe import loop
    db, g = _GoldenDB(golden("create_db")), golden("loop_candidates")
    checked = 0
    for k, (a, b) in enumerate(g["pairs"]):
        poses = {}
        for mode in ("gpu", "cv2"):
            monkeypatch.setattr(loop, "REFIT", mode)
            np.random.seed(8 + k)
            calls = []
            if mode == "gpu":
                monkeypatch.setattr(cv2, "solvePnP", lambda *a, **kw: calls.append(1) or (_ for _ in ()).throw(
                    AssertionError("host solve on the device-refit path")))
            matches, pct, pose = loop.check_candidate_match(int(a), int(b), db)
            monkeypatch.undo()
            poses[mode] = (None if pose is None else pose.matrix(), len(matches))
        assert poses["gpu"][1] == poses["cv2"][1]
        if poses["gpu"][1] >= 40:
            Pg, Pc = poses["gpu"][0], poses["cv2"][0]
            dR = Pg[:3, :3] @ Pc[:3, :3].T
            assert np.linalg.norm(dR - dR.T) / (2 * np.sqrt(2)) < 5e-3 and np.linalg.norm(Pg[:3, 3] - Pc[:3, 3]) < 0.05
            checked += 1
    assert checked >= 1


def test_candidate_gating_vs_oracle(slamfe, oracle):
    """slamfe_gate_candidates (get_good_candidates / check_candidate, loop_closure.py:164-228) for many query
    keyframes in one launch, against the oracle's restatement (pinned on the CPU against scipy's matrix
    logarithm and the unmodified reference's Graph class): same shortest paths (hop counts), Mahalanobis
    distances to 1e-9, the same <= 15 candidates below the threshold in the same order."""
    from slamfe import loop
    import sys
    sys.path.insert(0, __import__("os").path.dirname(__file__))
    from test_oracle import _random_pose_graph
    rng = np.random.default_rng(96)
    poses, edges = _random_pose_graph(rng, K=300, n_loops=8)
    g = loop.CovarianceGraph(len(poses))
    for a, b, c in edges:
        g.add_edge(a, b, c)
    queries = [12, 40, 41, 150, 151, 299] + rng.integers(11, 300, 30).tolist()
    dist, hops = loop.gate_distances(poses, g, queries)
    adj = {}
    for a, b, c in edges:
        w = np.linalg.det(c)
        adj.setdefault(a, {})[b] = w; adj.setdefault(b, {})[a] = w
    shortcuts = 0
    for q, n in enumerate(queries):
        want = oracle.gate_distances(poses, edges, n)
        fin = np.isfinite(want)
        assert np.array_equal(np.isfinite(dist[q]), fin)
        assert np.allclose(dist[q][fin], want[fin], rtol=1e-9, atol=0), (n, np.abs(dist[q][fin] / want[fin] - 1).max())
        for i in (0, max(0, n - 11)):
            if i < n - loop.KEY_FRAME_GAP:
                path = oracle.dijkstra_path(adj, i, n)
                assert hops[q, i] == len(path) - 1
                shortcuts += len(path) - 1 < n - i
        assert (hops[q][~fin] == -1).all()
        sw = np.sort(want[fin])
        thr = float((sw[len(sw) // 2 - 1] + sw[len(sw) // 2]) / 2) if len(sw) > 1 else 1.0   # splits this graph, hits no value
        got_sel = loop.select_candidates(dist[q], threshold=thr)
        ref = sorted([(want[i], i) for i in np.nonzero(want < thr)[0]])[:loop.MAX_CANDIDATES]
        assert got_sel == [i for _, i in ref] and len(got_sel) <= loop.MAX_CANDIDATES
    assert shortcuts > 0                                     # the loop edges are used by some shortest paths
    assert loop.get_good_candidates(5, poses, g) == []      # nothing is KEY_FRAME_GAP behind keyframe 5


class _FakePose3:
    def __init__(self, T):
        self._T = np.vstack([T, [0, 0, 0, 1.0]])

    def matrix(self):
        return self._T


class _FakeValues:
    """result.atPose3(key) of gtsam.Values, keys = ('c', frame) tuples from the stand-in symbol()."""
    def __init__(self, poses_by_frame):
        self.p = poses_by_frame

    def atPose3(self, key):
        return _FakePose3(self.p[key[1]])


class _FakeGraph:
    """The reference's backend/loop/graph.py Graph, as far as the gating reads it (.graph, det weights)."""
    def __init__(self):
        self.graph = {}

    def add_edge(self, a, b, cov):
        w = np.linalg.det(cov)
        self.graph.setdefault(a, {})[b] = w
        self.graph.setdefault(b, {})[a] = w


@pytest.mark.gpu
def test_typed_gating_dropin_follows_the_reference_bookkeeping(slamfe, oracle):
    """loop.get_good_candidates_typed — the drop-in behind loop_closure.get_good_candidates' own signature
    (marginals, result, index_list) — on stand-ins for the GTSAM objects (gtsam is not in this image) and for
    the module globals cov_dijkstra_graph / relative_covariance_dict filled as
    init_dijksra_graph_relative_covariance_dict does (loop_closure.py:246-287): direction-dependent covariances
    for consecutive keyframes, per-frame entries, loop-closure edges that only live in the graph.  Expected
    distances: the oracle's restatement of check_candidate with the same dictionaries."""
    from slamfe import loop
    import sys
    sys.path.insert(0, __import__("os").path.dirname(__file__))
    from test_oracle import _random_pose_graph
    rng = np.random.default_rng(197)
    K = 180
    poses, edges = _random_pose_graph(rng, K=K, n_loops=6)
    index_list = sorted(rng.choice(np.arange(3 * K), K, replace=False).tolist())       # keyframe frame numbers
    symbol = lambda ch, i: (ch, int(i))

    def spd():
        a = rng.normal(0, 1, (6, 6))
        return (a @ a.T + 6 * np.eye(6)) * 10.0 ** rng.uniform(-4, -2)

    ref_graph, cov_dict = _FakeGraph(), {}
    for a, b, c in edges:
        fa, fb = index_list[a], index_list[b]
        ref_graph.add_edge(fa, fb, c)                          # graph weights: det of the stored covariance
        if b == a + 1:                                         # consecutive keyframes: loop_closure.py:264-282
            cov_dict[str((fa, fb))] = c
            cov_dict[str((fb, fa))] = spd()                    # c1 given c2: a different matrix
            cov_dict[fb] = c
    cov_dict[index_list[0]] = spd()                            # :286
    values = _FakeValues({index_list[i]: poses[i] for i in range(K)})
    checked = used_loop = 0
    for n in [11, 60, 61, 140, K - 1]:
        want = oracle.gate_distances_typed(poses, ref_graph.graph, cov_dict, index_list, n)
        graph = loop.covariance_graph_from_reference(ref_graph, index_list, cov_dict, None, symbol)
        dist, hops = loop.gate_distances(poses, graph, [n])
        fin = np.isfinite(want)
        assert np.array_equal(np.isfinite(dist[0]), fin)
        assert np.allclose(dist[0][fin], want[fin], rtol=1e-9, atol=0)
        used_loop += int((hops[0][fin] < n - np.nonzero(fin)[0]).sum())
        got = loop.get_good_candidates_typed(n, None, values, index_list, ref_graph, cov_dict, symbol=symbol)
        ref = sorted([(want[i], i) for i in np.nonzero(want < loop.MAHALANOBIS_THRESHOLD)[0]])[:loop.MAX_CANDIDATES]
        assert got == [index_list[i] for _, i in ref]
        checked += len(got)
    assert used_loop > 0 and checked > 0   # loop edges are used by some shortest paths; some candidates pass the gate
    # an edge no dictionary covers falls back to the joint marginal information, as loop_closure.py:99-106

    class _Info:
        def __init__(self, m):
            self.m = m

        def at(self, a, b):
            return self.m

    class _Marg:
        def jointMarginalInformation(self, keys):
            return _Info(np.diag([4.0] * 6))

    c = loop.traversal_covariance(7, 9, {}, _Marg(), symbol)
    assert np.allclose(c, np.eye(6) / 4)
