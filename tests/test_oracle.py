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
