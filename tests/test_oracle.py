"""The oracle (oracle/ref_oracle.py + hamming_oracle.c) against the golden vectors generated from
the unmodified reference + cv2 4.13.0 (oracle/make_golden.py), and — when /root/reference is
present — against the reference itself through the import shim.  CPU only."""
import os

import numpy as np
import pytest

MATCH_CASES = ("akaze", "ragged", "orb", "ties", "one_train")


@pytest.mark.parametrize("name", MATCH_CASES)
def test_match_first_min(oracle, golden, name):
    g = golden("matching")
    idx, dist = oracle.match(g[f"{name}_q"], g[f"{name}_t"])
    assert np.array_equal(idx, g[f"{name}_match_t"])
    assert np.array_equal(dist, g[f"{name}_match_d"].astype(np.int32))


@pytest.mark.parametrize("name", MATCH_CASES)
def test_crosscheck_and_knn(oracle, golden, name):
    g = golden("matching")
    q, t = g[f"{name}_q"], g[f"{name}_t"]
    cq, ct, cd = oracle.match_crosscheck(q, t)
    assert np.array_equal(cq, g[f"{name}_cc_q"]) and np.array_equal(ct, g[f"{name}_cc_t"])
    assert np.array_equal(cd, g[f"{name}_cc_d"].astype(np.int32))
    assert np.all(np.diff(cq) > 0)  # sorted by queryIdx
    i2, d2 = oracle.knn2(q, t)
    assert np.array_equal(i2, g[f"{name}_knn_idx"]) and np.array_equal(d2, g[f"{name}_knn_dist"])
    assert np.array_equal(oracle.ratio_test(d2), g[f"{name}_ratio"])


def test_matrix_matches_numpy_bits(oracle):
    rng = np.random.default_rng(5)
    q = rng.integers(0, 256, (40, 61), dtype=np.uint8)
    t = rng.integers(0, 256, (33, 61), dtype=np.uint8)
    D = oracle.hamming_matrix(q, t)
    ref = np.unpackbits(q[:, None, :] ^ t[None, :, :], axis=2).sum(axis=2)
    assert np.array_equal(D, ref)
    idx, dist = oracle.match(q, t)
    assert np.array_equal(idx, ref.argmin(axis=1)) and np.array_equal(dist, ref.min(axis=1))
    ci, cd = oracle.colmin(q, t)
    assert np.array_equal(ci, ref.argmin(axis=0)) and np.array_equal(cd, ref.min(axis=0))


def test_ratio_integer_equivalence(oracle):
    # 5*d1 < 3*d2  <=>  d1 < 0.6*d2 in float64 for every reachable distance pair
    d1, d2 = np.meshgrid(np.arange(0, 513), np.arange(0, 513), indexing="ij")
    assert np.array_equal(5 * d1 < 3 * d2, d1.astype(np.float64) < 0.6 * d2.astype(np.float64))


def test_stereo_filter_and_links(oracle, golden):
    g = golden("stereo")
    inl, outl = oracle.extract_inliers_outliers(g["pts_l"], g["pts_r"], g["match_q"], g["match_t"])
    assert np.array_equal(inl, g["inliers"]) and np.array_equal(outl, g["outliers"])
    ident = np.arange(6)
    i6, o6 = oracle.extract_inliers_outliers(g["pts_l"], g["pts_r"], ident, ident)
    assert np.array_equal(i6, g["ident_inliers"]) and np.array_equal(o6, g["ident_outliers"])
    valid, links = oracle.create_links(g["pts_l"], g["pts_r"], g["match_q"][inl], g["match_t"][inl])
    assert np.array_equal(links, g["links"])
    assert np.array_equal(g["desc_l"][valid], g["features"])


def test_triangulation(oracle, golden):
    g = golden("triangulation")
    assert np.array_equal(oracle.triangulate_links(g["links"], g["P"], g["Q"]), g["xyz"])
    assert np.array_equal(oracle.triangulate_points(g["P"], g["Q"], g["pxy"], g["qxy"]), g["xyz_dlt"])
    assert np.array_equal(oracle.triangulate_points(g["P"], g["Q2"], g["pxy2"], g["qxy2"]), g["xyz_gen"])
    # closed form Z = fx*b/(xl-xr) (SURVEY 7 "triangulation precision")
    k, m1, m2 = oracle.read_cameras()
    assert np.array_equal(k, g["K"]) and np.array_equal(m1, g["M1"]) and np.array_equal(m2, g["M2"])
    z = -g["Q"][0, 3] / (g["links"][:, 0] - g["links"][:, 1])
    assert np.allclose(z, g["xyz"][:, 2], rtol=1e-11)


def test_ransac_scoring(oracle, golden):
    g = golden("ransac")
    masks = np.unpackbits(g["masks"], axis=1)[:, : g["pts"].shape[0]].astype(bool)
    counts, best, mask = oracle.score_hypotheses(g["Ts"], g["pts"], g["l_pix"], g["r_pix"], g["K"], g["M1"], g["M2"])
    assert np.array_equal(counts, g["counts"])
    assert best == int(np.argmax(g["counts"])) and np.array_equal(mask, masks[best])
    for h in (0, 3, 17, 95):
        assert np.array_equal(
            oracle.transformation_agreement(g["Ts"][h], g["pts"], g["l_pix"], g["r_pix"], g["K"], g["M1"], g["M2"]),
            masks[h])
    for p, n in g["iters"]:
        assert oracle.calc_ransac_iteration(int(p)) == int(n)


def test_seeded_ransac_loop(oracle, golden):
    g = golden("ransac")
    n = g["prev_links"].shape[0]
    np.random.seed(7)
    best = oracle.ransac_pnp_for_tracking_db(np.arange(n), g["match_t"], g["prev_links"], g["cur_links"], 55,
                                             g["K"], g["M1"], g["M2"])
    assert np.array_equal(best, g["tracking_best_idx"])


def test_database_mutual(oracle, golden):
    g = golden("database")
    fi, fd, good = oracle.mutual_forward_backward(g["prev"], g["cur"])
    assert np.array_equal(fi, g["fwd_t"]) and np.array_equal(fd, g["fwd_d"].astype(np.int32))
    assert np.array_equal(good, g["good_idx"])
    bi, _ = oracle.match(g["cur"], g["prev"])
    assert np.array_equal(bi, g["bwd_t"])
    # one pass gives both directions: column minima of (prev, cur) == match(cur, prev)
    ci, _ = oracle.colmin(g["prev"], g["cur"])
    assert np.array_equal(ci, bi)


def test_oracle_vs_cv2_live(oracle):
    """cv2 is part of the image (also on the GPU box): re-check the C restatement against it."""
    import slamfe
    from slamfe import synth
    rng = np.random.default_rng(77)
    q = synth.descriptors(rng, 500)
    t, _ = synth.paired_descriptors(rng, q, n_out=470, dup_frac=0.05)
    _, ti, td = oracle.cv2_match(q, t)
    oi, od = oracle.match(q, t)
    assert np.array_equal(oi, ti) and np.array_equal(od, td.astype(np.int32))
    cq, ct, cd = oracle.cv2_match(q, t, cross_check=True)
    xq, xt, xd = oracle.match_crosscheck(q, t)
    assert np.array_equal(cq, xq) and np.array_equal(ct, xt) and np.array_equal(cd.astype(np.int32), xd)
    k_i, k_d = oracle.cv2_knn2(q, t)
    o_i, o_d = oracle.knn2(q, t)
    assert np.array_equal(k_i, o_i) and np.array_equal(k_d, o_d)


@pytest.mark.skipif(not os.path.isdir("/root/reference/final_project"), reason="reference tree not present")
def test_oracle_vs_reference_shim(oracle):
    """Build container only: the restatement against the unmodified reference functions."""
    import cv2
    from oracle import refshim
    import slamfe
    from slamfe import synth
    ref = refshim.load()
    rng = np.random.default_rng(9)
    links = synth.links(rng, 50)
    objs = [ref.tracking_database.Link(*r) for r in links]
    assert np.array_equal(ref.triangulation.triangulate_links(objs, ref.ransac.P, ref.ransac.Q),
                          oracle.triangulate_links(links, ref.ransac.P, ref.ransac.Q))
    Ts, pts, lp, rp = synth.pnp_problem(rng, 200, 8)
    for T in Ts:
        assert np.array_equal(ref.ransac.transformation_agreement(T, pts, lp, rp),
                              oracle.transformation_agreement(T, pts, lp, rp, ref.ransac.K, ref.ransac.M1, ref.ransac.M2))
    dl, dr, pl, pr = synth.stereo_frame(rng, 150)
    ms = ref.matching.MATCHER_LEFT_RIGHT.match(dl, dr)
    kpl = tuple(cv2.KeyPoint(float(x), float(y), 1.0) for x, y in pl)
    kpr = tuple(cv2.KeyPoint(float(x), float(y), 1.0) for x, y in pr)
    inl, outl = ref.matching.extract_inliers_outliers(kpl, kpr, ms)
    mq = np.array([m.queryIdx for m in ms]); mt = np.array([m.trainIdx for m in ms])
    oi, oo = oracle.extract_inliers_outliers(pl, pr, mq, mt)
    assert np.array_equal(inl, oi) and np.array_equal(outl, oo)


# ---------------------------------------------------------------------------------------------
# Minimal-sample pose solver (GPU hypothesis generation): the numpy oracle and the
# HOST build of the product's solver header are pinned against ground truth and cv2.SOLVEPNP_P3P.
# ---------------------------------------------------------------------------------------------
def _pose_samples(rng, n, noise=0.0):
    from slamfe import synth
    K, _, _ = synth.cameras()
    out = []
    for _ in range(n):
        z = rng.uniform(4, 60, 4)
        P = np.stack([(rng.uniform(20, 1220, 4) - K[0, 2]) * z / K[0, 0],
                      (rng.uniform(5, 370, 4) - K[1, 2]) * z / K[1, 1], z], axis=1)
        R = synth._rodrigues(rng.normal(0, 0.05, 3))
        t = np.array([0.0, 0.0, -0.9]) + rng.normal(0, 0.2, 3)
        C = (R @ P.T).T + t
        pix = (K @ C.T).T
        pix = pix[:, :2] / pix[:, 2:3] + rng.normal(0, noise, (4, 2)) if noise else pix[:, :2] / pix[:, 2:3]
        out.append((P, pix, np.hstack([R, t[:, None]])))
    return K, out


def test_p3p_oracle_vs_ground_truth_and_cv2():
    import cv2
    from oracle import p3p_oracle as p3
    rng = np.random.default_rng(300)
    K, samples = _pose_samples(rng, 400)
    gt_ok = cv_ok = cv_n = 0
    for P, pix, T_gt in samples:
        T, ok = p3.solve_sample(P, pix, K)
        assert ok
        gt_ok += np.abs(T - T_gt).max() < 1e-6
        s, rv, tv = cv2.solvePnP(P, pix, K, np.zeros((5, 1)), flags=cv2.SOLVEPNP_P3P)
        if s:
            cv_n += 1
            cv_ok += np.abs(T - np.hstack([cv2.Rodrigues(rv)[0], tv])).max() < 1e-5
    assert gt_ok >= 0.985 * len(samples) and cv_ok >= 0.985 * cv_n and cv_n > 350


def test_p3p_host_build_of_product_header(oracle):
    from oracle import p3p_oracle as p3
    host = oracle.P3PHost()
    rng = np.random.default_rng(301)
    K, samples = _pose_samples(rng, 1500, noise=0.5)
    both = agree = 0
    for P, pix, _ in samples:
        T, ok = host.solve(P, pix, K)
        To, oko = p3.solve_sample(P, pix, K)
        if ok and oko:
            both += 1
            agree += np.abs(T - To).max() < 1e-6 * max(1.0, np.abs(To).max())
        if ok:  # a returned pose reproduces the three minimal points exactly and is a rotation
            c = (T[:, :3] @ P[:3].T).T + T[:, 3]
            uv = (K @ c.T).T
            assert np.abs(uv[:, :2] / uv[:, 2:3] - pix[:3]).max() < 1e-3  # px; ill-conditioned samples lose digits
            assert np.abs(T[:, :3] @ T[:, :3].T - np.eye(3)).max() < 1e-9 and np.linalg.det(T[:, :3]) > 0
    assert both >= 0.97 * len(samples) and agree >= 0.995 * both


def test_quartic_solver_has_no_wrong_roots(oracle):
    host = oracle.P3PHost()
    rng = np.random.default_rng(302)
    missing = 0
    for _ in range(4000):
        A = rng.normal(0, 1, 5) * 10 ** rng.uniform(-3, 3, 5)
        ref = np.roots(A)
        ref = np.sort(ref[np.abs(ref.imag) < 1e-9 * np.maximum(1, np.abs(ref.real))].real)
        got = np.sort(host.quartic(A))
        for g in got:  # every returned root is a true root
            assert np.min(np.abs(ref - g) / np.maximum(1e-300, np.abs(g))) < 1e-6, (A, g, ref)
        missing += len(got) < len(ref)
    assert missing < 0.03 * 4000
    assert np.allclose(np.sort(host.quartic([1, -10, 35, -50, 24])), [1, 2, 3, 4])
    assert np.allclose(np.sort(host.quartic([1, 0, -5, 0, 4])), [-2, -1, 1, 2])  # biquadratic branch
    assert len(host.quartic([1, 0, 2, 0, 5])) == 0


def test_sample4_is_a_uniform_distinct_subset(oracle):
    host = oracle.P3PHost()
    seen = {}
    for h in range(8400):
        s = host.sample4(7, 3, h, 9)
        assert len(set(s.tolist())) == 4 and s.min() >= 0 and s.max() < 9
        seen[frozenset(s.tolist())] = seen.get(frozenset(s.tolist()), 0) + 1
    assert len(seen) == 126 and min(seen.values()) > 30 and max(seen.values()) < 110  # 8400 / 126 = 66.7
    assert np.array_equal(host.sample4(7, 3, 5, 9), host.sample4(7, 3, 5, 9))
    assert not np.array_equal(host.sample4(7, 3, 5, 1000), host.sample4(8, 3, 5, 1000))
    assert sorted(host.sample4(1, 0, 0, 4).tolist()) == [0, 1, 2, 3]


def test_ransac_with_minimal_solver_is_statistically_equivalent_to_epnp(oracle):
    """The contract of the GPU generator: same RANSAC outcome as with cv2 EPnP hypotheses up to the
    randomness the reference itself has (unseeded sampling): best inlier count within 5 % and the
    best inlier sets overlap (Jaccard >= 0.9)."""
    import cv2
    from slamfe import synth
    host = oracle.P3PHost()
    K, M1, M2 = synth.cameras()
    rng = np.random.default_rng(303)
    for _ in range(3):
        _, pts, lp, rp = synth.pnp_problem(rng, 700, 1, outlier_frac=0.4)
        Te, Tp = [], []
        for h in range(120):
            idx = rng.choice(len(pts), 4, replace=False)
            s, rv, tv = cv2.solvePnP(pts[idx], lp[idx], K, np.zeros((5, 1)), flags=cv2.SOLVEPNP_EPNP)
            Te.append(np.hstack([cv2.Rodrigues(rv)[0], tv]) if s else np.zeros((3, 4)))
            Tp.append(host.solve(pts[idx], lp[idx], K)[0])
        ce, _, me = oracle.score_hypotheses(np.array(Te), pts, lp, rp, K, M1, M2)
        cp, _, mp = oracle.score_hypotheses(np.array(Tp), pts, lp, rp, K, M1, M2)
        assert abs(int(ce.max()) - int(cp.max())) <= 0.05 * ce.max()
        assert (me & mp).sum() / (me | mp).sum() >= 0.9


def test_p3p_degenerate_samples_never_yield_invalid_poses(oracle):
    """Collinear / coincident / behind-camera / non-finite samples: the solver either reports failure or
    returns a finite proper rotation that reproduces the minimal points — never NaN flagged as valid."""
    host = oracle.P3PHost()
    rng = np.random.default_rng(304)
    K, samples = _pose_samples(rng, 300, noise=0.3)
    checked = valid = 0
    for P, pix, _ in samples:
        cases = []
        Pc = P.copy(); Pc[2] = Pc[0] + 2.0 * (Pc[1] - Pc[0]); cases.append((Pc, pix))        # collinear 3-D points
        Pd = P.copy(); Pd[1] = Pd[0]; cases.append((Pd, pix))                                   # coincident points
        px = pix.copy(); px[1] = px[0]; cases.append((P, px))                                   # coincident pixels
        Pn = P.copy(); Pn[0, 0] = np.nan; cases.append((Pn, pix))                               # NaN
        pi_ = pix.copy(); pi_[2, 1] = np.inf; cases.append((P, pi_))                            # inf pixel
        cases.append((P * 1e-9, pix))                                                           # tiny scene
        cases.append((P * 1e9, pix))                                                            # huge scene
        cases.append((P, pix[::-1].copy()))                                                     # mismatched pixels
        for Pq, pq in cases:
            with np.errstate(all="ignore"):
                T, ok = host.solve(Pq, pq, K)
            checked += 1
            if not ok:
                continue
            valid += 1
            assert np.isfinite(T).all()
            R = T[:, :3]
            assert np.abs(R @ R.T - np.eye(3)).max() < 1e-6 and np.linalg.det(R) > 0
    assert checked == 300 * 8 and valid > 0


def test_oracle_on_real_akaze_descriptors(oracle, golden):
    """tests/golden/akaze_real.npz: REAL cv2.AKAZE descriptors (correlated bits, > 3k keypoints per image)
    through the reference's own extract_kps_descs_matches / extract_inliers_outliers / MATCHER.match /
    knnMatch (matching.py:38-69, database.py:54-55, ex1.py:189-190)."""
    g = golden("akaze_real")
    d0, d1, d2 = g["desc_l"], g["desc_r"], g["desc_l2"]
    assert d0.shape[1] == 61 and min(len(d0), len(d1)) > 2000 and int(d0[:, 60].max()) <= 63
    cq, ct, cd = oracle.match_crosscheck(d0, d1)
    assert np.array_equal(cq, g["cross_q"]) and np.array_equal(ct, g["cross_t"]) and np.array_equal(cd, g["cross_d"])
    inl, outl = oracle.extract_inliers_outliers(g["pts_l"], g["pts_r"], cq, ct)
    assert np.array_equal(inl, g["inliers"]) and np.array_equal(outl, g["outliers"]) and len(inl) > 500
    mi, md = oracle.match(d0, d2)
    assert np.array_equal(mi, g["match_t"]) and np.array_equal(md, g["match_d"])
    bi, bd = oracle.match(d2, d0)
    assert np.array_equal(bi, g["back_t"]) and np.array_equal(bd, g["back_d"])
    ki, kd = oracle.knn2(d0, d2)
    assert np.array_equal(ki, g["knn_idx"]) and np.array_equal(kd, g["knn_dist"])


def test_track_ids_restatement_on_the_48_frame_reference_run(golden):
    """tests/golden/create_db_48.npz: the unmodified reference's create_db + TrackingDB on the bench
    workload's first 48 frames.  The host restatement of slamfe_track_ids (slamfe.trackdb.track_ids_host),
    fed the reference's own forward matches and inlier flags, reproduces frameId_to_trackIds_list of every
    frame (tracking_database.py:273-337) and the track count."""
    from slamfe import trackdb
    g = golden("create_db_48")
    F = int(g["n_frames"])
    n_links = g["n_links"]
    fwd = [g[f"match_t{f + 1}"].astype(np.int64) for f in range(F - 1)]
    inl = [np.unpackbits(g[f"inliers{f + 1}"])[:n_links[f]].astype(bool) for f in range(F - 1)]
    ids, n_tracks = trackdb.track_ids_host(fwd, inl, n_links)
    assert n_tracks == int(g["n_tracks"]) > 10000
    for f in range(F):
        assert np.array_equal(ids[f], g[f"track_ids{f}"]), f
    assert sum(int((a >= 0).sum()) for a in ids) == int(g["n_links_total"])


def _random_pose_graph(rng, K=120, n_loops=5):
    """A keyframe chain with a few loop edges, random-walk poses and SPD 6x6 edge covariances."""
    from slamfe import synth
    poses = np.zeros((K, 3, 4))
    R, t = np.eye(3), np.zeros(3)
    for k in range(K):
        poses[k] = np.hstack([R, t[:, None]])
        R = R @ synth._rodrigues(rng.normal(0, 0.05, 3))
        t = t + R @ np.array([0.0, 0.0, 1.0]) * rng.uniform(5, 15) + rng.normal(0, 0.2, 3)
    def spd(scale):
        a = rng.normal(0, 1, (6, 6))
        return (a @ a.T + 6 * np.eye(6)) * scale
    edges = [(k, k + 1, spd(rng.uniform(1e-3, 3e-3))) for k in range(K - 1)]
    for _ in range(n_loops):
        a, b = sorted(rng.choice(K, 2, replace=False).tolist())
        if b - a > 1:
            edges.append((a, b, spd(rng.uniform(2e-4, 8e-4))))
    return poses, edges


def test_gating_restatement_against_scipy_and_the_reference_graph(oracle):
    """Candidate gating (loop_closure.py:164-228): gtsam is not installed, so the SE(3) logarithm is pinned
    against scipy.linalg.logm of the homogeneous matrix (an independent algorithm) and the shortest paths
    against the UNMODIFIED reference's Graph class (backend/loop/graph.py), which imports without gtsam."""
    from scipy.linalg import logm
    from slamfe import synth
    rng = np.random.default_rng(95)
    for scale in (1e-9, 1e-3, 0.3, 1.5, 3.0):
        for _ in range(6):
            R = synth._rodrigues(rng.normal(0, 1, 3) * scale / np.sqrt(3))
            t = rng.normal(0, 5, 3)
            T = np.eye(4); T[:3, :3] = R; T[:3, 3] = t
            L = np.real(logm(T))
            want = np.array([L[2, 1], L[0, 2], L[1, 0], L[0, 3], L[1, 3], L[2, 3]])
            got = oracle.pose3_logmap(R, t)
            assert np.allclose(got, want, rtol=1e-7, atol=1e-9), (scale, got, want)
    from oracle import refshim
    poses, edges = _random_pose_graph(rng)
    mine = oracle.gate_distances(poses, edges, 110)
    assert np.isfinite(mine[:100]).all() and np.isinf(mine[100:]).all()
    if refshim.available():
        import importlib
        sys_path_added = False
        import sys
        if "/root/reference" not in sys.path:
            sys.path.insert(0, "/root/reference"); sys_path_added = True
        try:
            Graph = importlib.import_module("final_project.backend.loop.graph").Graph
        finally:
            if sys_path_added:
                sys.path.remove("/root/reference")
        ref = oracle.gate_distances(poses, edges, 110, graph_cls=Graph)
        assert np.array_equal(ref, mine)                     # same paths, same summation order
        g = Graph()
        adj = {}
        for a, b, c in edges:
            g.add_edge(a, b, c)
            w = np.linalg.det(c)
            adj.setdefault(a, {})[b] = w; adj.setdefault(b, {})[a] = w
        for _ in range(200):
            a, b = rng.choice(len(poses), 2, replace=False).tolist()
            assert g.get_shortest_path(a, b) == oracle.dijkstra_path(adj, a, b)
