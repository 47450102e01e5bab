"""Generate tests/golden/*.npz from the UNMODIFIED reference (through oracle/refshim.py) and from
cv2 4.13.0, in the build container.  Test infrastructure only.

    python oracle/make_golden.py

The reference ships no tests or golden vectors, so parity is pinned by
these generated fixtures: every array below is the output of the reference's own function (or
of the cv2 call the reference makes) on seeded synthetic inputs.  The script also asserts that
the oracle restatement (oracle/ref_oracle.py + hamming_oracle.c) reproduces each of them, so a
green run pins the oracle as well.
"""
from __future__ import annotations

import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_oracle as ora  # noqa: E402
from oracle import refshim  # noqa: E402
import slamfe  # noqa: E402,F401
from slamfe import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def dm_arrays(ms):
    return (np.array([m.queryIdx for m in ms], np.int32), np.array([m.trainIdx for m in ms], np.int32),
            np.array([m.distance for m in ms], np.float32), np.array([m.imgIdx for m in ms], np.int32))


def golden_matching(ref):
    rng = np.random.default_rng(100)
    cases = {}
    # (name, nq, nt, width): 61 = AKAZE MLDB, 32 = ORB-like, 2 = tie torture
    for name, nq, nt, w in (("akaze", 300, 280, 61), ("ragged", 37, 515, 61), ("orb", 130, 90, 32),
                            ("ties", 96, 80, 2), ("one_train", 9, 1, 61)):
        q = rng.integers(0, 256, (nq, w), dtype=np.uint8)
        t = rng.integers(0, 256, (nt, w), dtype=np.uint8)
        if w == 61:
            q = synth.descriptors(rng, nq)
            t, _ = synth.paired_descriptors(rng, q, n_out=nt, dup_frac=0.05)
        if w == 2:
            q[:, 1] &= 0x03
            t[:, 1] &= 0x03
        M, MLR = ref.matching.MATCHER, ref.matching.MATCHER_LEFT_RIGHT
        mq, mt, md, mi = dm_arrays(M.match(q, t))
        cq, ct, cd, _ = dm_arrays(MLR.match(q, t))
        knn = M.knnMatch(q, t, k=2)
        k_idx = np.full((nq, 2), -1, np.int32)
        k_dist = np.full((nq, 2), -1, np.int32)
        for i, pair in enumerate(knn):
            for c, m in enumerate(pair):
                k_idx[i, c], k_dist[i, c] = m.trainIdx, int(m.distance)
        # the restatement must agree with cv2
        oi, od = ora.match(q, t)
        assert np.array_equal(oi, mt) and np.array_equal(od, md.astype(np.int32)), name
        assert np.array_equal(mq, np.arange(nq)) and np.all(mi == 0)
        xq, xt, xd = ora.match_crosscheck(q, t)
        assert np.array_equal(xq, cq) and np.array_equal(xt, ct) and np.array_equal(xd, cd.astype(np.int32)), name
        o2i, o2d = ora.knn2(q, t)
        assert np.array_equal(o2i, k_idx) and np.array_equal(o2d, k_dist), name
        # ratio test of ex1.py:118-122 evaluated the reference's way (float compare on DMatch.distance)
        ratio = np.array([len(p) == 2 and p[0].distance < 0.6 * p[1].distance for p in knn])
        assert np.array_equal(ratio, ora.ratio_test(k_dist))
        cases.update({f"{name}_q": q, f"{name}_t": t, f"{name}_match_t": mt, f"{name}_match_d": md,
                      f"{name}_cc_q": cq, f"{name}_cc_t": ct, f"{name}_cc_d": cd,
                      f"{name}_knn_idx": k_idx, f"{name}_knn_dist": k_dist, f"{name}_ratio": ratio})
    np.savez_compressed(os.path.join(OUT, "matching.npz"), **cases)


def golden_stereo(ref):
    rng = np.random.default_rng(101)
    dl, dr, pl, pr = synth.stereo_frame(rng, 400)
    # a few exact-threshold cases for the strict comparisons of matching.py:62-63
    pl[:6] = np.array([[100, 50], [100, 50], [100, 50], [100, 50], [300.5, 20.25], [300.5, 20.25]], np.float32)
    pr[:6] = np.array([[98, 50], [97.99, 50], [90, 52], [90, 51.99], [298.5, 22.25], [298.25, 18.5]], np.float32)
    kpl = tuple(cv2.KeyPoint(float(x), float(y), 1.0) for x, y in pl)
    kpr = tuple(cv2.KeyPoint(float(x), float(y), 1.0) for x, y in pr)
    matches = ref.matching.MATCHER_LEFT_RIGHT.match(dl, dr)
    mq, mt, md, _ = dm_arrays(matches)
    inl, outl = ref.matching.extract_inliers_outliers(kpl, kpr, matches)
    oi, oo = ora.extract_inliers_outliers(pl, pr, mq, mt)
    assert np.array_equal(oi, inl) and np.array_equal(oo, outl)
    # identity matches over the handcrafted rows (queryIdx == trainIdx) to hit the thresholds
    ident = tuple(cv2.DMatch(i, i, 0, 0.0) for i in range(6))
    inl6, outl6 = ref.matching.extract_inliers_outliers(kpl, kpr, ident)
    # TrackingDB.create_links on the filtered matches (database.py:21-25)
    filt = [matches[i] for i in inl]
    feats, links = ref.tracking_database.TrackingDB.create_links(dl, kpl, kpr, filt, np.array([True] * len(filt)))
    links_arr = np.array([[l.x_left, l.x_right, l.y] for l in links])
    valid, olinks = ora.create_links(pl, pr, mq[inl], mt[inl])
    assert np.array_equal(olinks, links_arr) and np.array_equal(dl[valid], feats)
    np.savez_compressed(os.path.join(OUT, "stereo.npz"), desc_l=dl, desc_r=dr, pts_l=pl, pts_r=pr,
                        match_q=mq, match_t=mt, match_d=md, inliers=inl, outliers=outl,
                        ident_inliers=inl6, ident_outliers=outl6, links=links_arr, features=feats)


def golden_triangulation(ref):
    rng = np.random.default_rng(102)
    K, M1, M2 = ref.ransac.K, ref.ransac.M1, ref.ransac.M2
    P, Q = ref.ransac.P, ref.ransac.Q
    links = synth.links(rng, 300)
    Link = ref.tracking_database.Link
    link_objs = [Link(*row) for row in links]
    xyz = ref.triangulation.triangulate_links(link_objs, P, Q)
    assert np.allclose(ora.triangulate_links(links, P, Q), xyz, rtol=0, atol=0)
    # general DLT: distinct y and a non-rectified second camera (analysis.py:400 style use)
    n = 120
    pxy = np.stack([rng.uniform(100, 1200, n), rng.uniform(10, 360, n)], axis=1)
    qxy = pxy + np.stack([-rng.uniform(3, 90, n), rng.normal(0, 0.7, n)], axis=1)
    xyz_dlt = np.array([ref.triangulation.linear_least_squares_triangulation(P, Q, pxy[i], qxy[i]) for i in range(n)])
    th = 0.05
    R = np.array([[np.cos(th), 0, np.sin(th)], [0, 1, 0], [-np.sin(th), 0, np.cos(th)]])
    Q2 = K @ np.hstack([R, np.array([[-0.5], [0.02], [0.01]])])
    X = np.stack([rng.uniform(-10, 10, n), rng.uniform(-2, 2, n), rng.uniform(5, 50, n), np.ones(n)])
    pp, qq = (P @ X), (Q2 @ X)
    pxy2 = (pp[:2] / pp[2]).T + rng.normal(0, 0.3, (n, 2))
    qxy2 = (qq[:2] / qq[2]).T + rng.normal(0, 0.3, (n, 2))
    xyz_gen = np.array([ref.triangulation.linear_least_squares_triangulation(P, Q2, pxy2[i], qxy2[i])
                        for i in range(n)])
    assert np.array_equal(ora.triangulate_points(P, Q2, pxy2, qxy2), xyz_gen)
    np.savez_compressed(os.path.join(OUT, "triangulation.npz"), K=K, M1=M1, M2=M2, P=P, Q=Q, links=links, xyz=xyz,
                        pxy=pxy, qxy=qxy, xyz_dlt=xyz_dlt, Q2=Q2, pxy2=pxy2, qxy2=qxy2, xyz_gen=xyz_gen)


def golden_ransac(ref):
    rng = np.random.default_rng(103)
    K, M1, M2 = ref.ransac.K, ref.ransac.M1, ref.ransac.M2
    Ts, pts, l_pix, r_pix = synth.pnp_problem(rng, 700, 96)
    pts[5] = [0.3, -0.2, -4.0]          # behind the camera: the reference has no cheirality test
    pts[6] = [1.0, 1.0, 0.0]
    Ts[3] = np.hstack([np.eye(3), np.zeros((3, 1))])  # with pts[6]: z == 0 -> inf/nan -> not an inlier
    masks = np.array([ref.ransac.transformation_agreement(T, pts, l_pix, r_pix) for T in Ts])
    counts = masks.sum(axis=1)
    oc, ob, om = ora.score_hypotheses(Ts, pts, l_pix, r_pix, K, M1, M2)
    assert np.array_equal(oc, counts) and ob == int(np.argmax(counts)) and np.array_equal(om, masks[ob])
    iters = np.array([[p, ref.ransac.calc_ransac_iteration(p)] for p in (100, 90, 80, 70, 60, 50, 40, 30, 27, 25)])
    assert all(ora.calc_ransac_iteration(p) == n for p, n in iters)

    # a full seeded ransac_pnp_for_tracking_db run (ransac.py:70-113) on Link / DMatch objects
    Link = ref.tracking_database.Link
    n = 260
    T_gt = Ts[0]
    prev = synth.links(rng, n)
    P, Q = K @ M1, K @ M2
    X = ora.triangulate_links(prev, P, Q)
    X4 = np.hstack([X, np.ones((n, 1))]).T
    pl = K @ T_gt @ np.vstack([M1, [0, 0, 0, 1]]) @ X4
    pr = K @ T_gt @ np.vstack([M2, [0, 0, 0, 1]]) @ X4
    xl = pl[0] / pl[2] + rng.normal(0, 0.3, n)
    xr = pr[0] / pr[2] + rng.normal(0, 0.3, n)
    y = pl[1] / pl[2] + rng.normal(0, 0.3, n)
    cur = np.stack([xl, xr, y], axis=1)
    bad = rng.random(n) < 0.35
    cur[bad, 0] += rng.uniform(-60, 60, bad.sum())
    perm = rng.permutation(n)
    cur_links = [Link(*cur[j]) for j in perm]            # cur link k = cur[perm[k]]
    inv = np.argsort(perm)
    prev_links = [Link(*row) for row in prev]
    matches = [cv2.DMatch(i, int(inv[i]), 0, 10.0) for i in range(n)]
    np.random.seed(7)
    best_idx = ref.ransac.ransac_pnp_for_tracking_db(matches, prev_links, cur_links, 55)
    np.random.seed(7)
    obest = ora.ransac_pnp_for_tracking_db(np.arange(n), inv, prev, np.array([[l.x_left, l.x_right, l.y] for l in cur_links]),
                                           55, K, M1, M2)
    assert np.array_equal(best_idx, obest)
    np.random.seed(11)
    pose, idx2, cnt2 = ref.ransac.ransac_pnp(matches, prev_links, cur_links, inliers_percent=50)
    np.savez_compressed(os.path.join(OUT, "ransac.npz"), K=K, M1=M1, M2=M2, Ts=Ts, pts=pts, l_pix=l_pix, r_pix=r_pix,
                        masks=np.packbits(masks, axis=1), counts=counts, iters=iters,
                        prev_links=prev, cur_links=np.array([[l.x_left, l.x_right, l.y] for l in cur_links]),
                        match_t=inv.astype(np.int32), tracking_best_idx=best_idx, pnp_best_idx=idx2,
                        pnp_best_inliers=np.int64(cnt2), pnp_pose=pose.matrix())


def golden_database(ref):
    """database.py:54-77 forward/backward mutual check on synthetic consecutive-frame features."""
    rng = np.random.default_rng(104)
    prev = synth.descriptors(rng, 350)
    cur = synth.next_frame_descriptors(rng, prev, 330)
    M = ref.matching.MATCHER
    fwd = np.array(M.match(prev, cur))
    bwd = np.array(M.match(cur, prev))
    good = []
    for j, m in enumerate(fwd):
        if bwd[m.trainIdx].trainIdx != m.queryIdx:
            continue
        good.append(j)
    fi, fd, og = ora.mutual_forward_backward(prev, cur)
    assert np.array_equal(og, np.array(good)) and np.array_equal(fi, [m.trainIdx for m in fwd])
    np.savez_compressed(os.path.join(OUT, "database.npz"), prev=prev, cur=cur,
                        fwd_t=np.array([m.trainIdx for m in fwd], np.int32),
                        fwd_d=np.array([m.distance for m in fwd], np.float32),
                        bwd_t=np.array([m.trainIdx for m in bwd], np.int32), good_idx=np.array(good))


def golden_create_db(ref):
    """The UNMODIFIED reference's create_db (database.py:30-89) + TrackingDB on 5 synthetic frames
    (3-D-consistent motion), with its image reader / AKAZE detector replaced by a provider of the
    synthetic keypoints and descriptors (inputs, not code under test) and np.random seeded.  Every
    add_frame call is recorded: links, features, forward matches, inlier flags."""
    import torch  # noqa: F401  (synth.torch_sequence)
    n = 5
    seq = synth.torch_sequence(n, first_frame=3, seed=5, device="cpu", lo=150, hi=260)
    frames = []
    for f in range(n):
        lo, k = int(seq["l_off"][f]), int(seq["n_l"][f])
        frames.append((seq["pts_l"][lo:lo + k].numpy(), seq["pts_r"][lo:lo + k].numpy(),
                       seq["desc_l"][lo:lo + k].numpy(), seq["desc_r"][lo:lo + k].numpy()))

    class Provider:
        def detectAndCompute(self, token, mask):
            side, f = token
            pts = frames[f][0 if side == "L" else 1]
            return tuple(cv2.KeyPoint(float(x), float(y), 1.0) for x, y in pts), frames[f][2 if side == "L" else 3]

    old_feature, old_reader = ref.matching.FEATURE, ref.inputs.read_images
    ref.matching.FEATURE = Provider()
    ref.inputs.read_images = lambda idx: (("L", idx), ("R", idx))
    calls = []
    try:
        db = ref.tracking_database.TrackingDB()
        real_add = db.add_frame

        def spy(links, left_features, matches_to_previous_left=None, inliers=None):
            calls.append((links, left_features, matches_to_previous_left, inliers))
            return real_add(links, left_features, matches_to_previous_left, inliers)

        db.add_frame = spy
        np.random.seed(3)
        ref.database.create_db(start_frame=0, num_frames=n, db=db)
    finally:
        ref.matching.FEATURE, ref.inputs.read_images = old_feature, old_reader
    out = {"n_frames": np.array(n)}
    for f, (pl, pr, dl, dr) in enumerate(frames):
        out[f"pts_l{f}"], out[f"pts_r{f}"], out[f"desc_l{f}"], out[f"desc_r{f}"] = pl, pr, dl, dr
    for f, (links, feats, ms, inl) in enumerate(calls):
        out[f"links{f}"] = np.array([(l.x_left, l.x_right, l.y) for l in links], np.float64).reshape(-1, 3)
        out[f"features{f}"] = feats
        out[f"inliers_percent{f}"] = np.array(db.frameID_to_inliers_percent[f])
        if ms is not None:
            out[f"match_t{f}"] = np.array([m.trainIdx for m in ms], np.int32)
            out[f"match_d{f}"] = np.array([m.distance for m in ms], np.float32)
            out[f"inliers{f}"] = np.asarray(inl, dtype=bool)
    out["n_tracks"] = np.array(len(db.trackId_to_frames))
    np.savez_compressed(os.path.join(OUT, "create_db.npz"), **out)
    return db


def golden_loop_candidates(db):
    """The UNMODIFIED reference's check_candidate_match (backend/loop/loop_closure.py:405-436:
    MATCHER.match at :422 + ransac_pnp(..., inliers_percent=40) at :425, 888 iterations) on keyframes of
    the TrackingDB built by golden_create_db: four overlapping pairs and one unrelated pair."""
    lc = refshim.load_loop_closure().loop_closure
    pairs = [(1, 0), (2, 1), (3, 2), (4, 3), (4, 1)]
    out = {"pairs": np.array(pairs, np.int32)}
    for k, (a, b) in enumerate(pairs):
        np.random.seed(20 + k)
        fa, fb = db.features(a), db.features(b)
        ms = lc.MATCHER.match(fa, fb)
        inl, pct, pose = lc.check_candidate_match(a, b, db)
        out[f"match_t{k}"] = np.array([m.trainIdx for m in ms], np.int32)
        out[f"match_d{k}"] = np.array([m.distance for m in ms], np.float32)
        out[f"inlier_q{k}"] = np.array([m.queryIdx for m in inl], np.int32)
        out[f"percentage{k}"] = np.array(float(pct))
        out[f"pose{k}"] = pose.matrix() if pose is not None else np.zeros((0, 0))
    np.savez_compressed(os.path.join(OUT, "loop_candidates.npz"), **out)


def _texture(rng, h=376, w=1241):
    """A corner-rich synthetic image (random rectangles and discs at three scales, blurred)."""
    img = np.zeros((h, w), np.float32)
    for s, n, a in ((3, 4000, 60), (7, 1500, 80), (15, 400, 90)):
        for x, y in zip(rng.integers(0, w, n), rng.integers(0, h, n)):
            c = float(rng.uniform(-a, a))
            if rng.random() < 0.5:
                cv2.rectangle(img, (int(x), int(y)), (int(x + rng.integers(2, s * 3)), int(y + rng.integers(2, s * 3))), c, -1)
            else:
                cv2.circle(img, (int(x), int(y)), int(rng.integers(1, s * 2)), c, -1)
    img = cv2.GaussianBlur(img, (0, 0), 1.0)
    img -= img.min()
    return (img / img.max() * 255).astype(np.uint8)


def golden_akaze(ref):
    """REAL cv2.AKAZE output (correlated bits, cv2-owned (N, 61) arrays, > 2k keypoints) through the
    reference's own extract_kps_descs_matches (matching.py:38-45: detectAndCompute x2 + MATCHER_LEFT_RIGHT
    crossCheck match) and extract_inliers_outliers (:48-69), on a synthetic textured rectified stereo pair
    (smooth disparity field) and a second left view (small zoom + shift) for MATCHER.match / knnMatch."""
    rng = np.random.default_rng(7)
    left = _texture(rng)
    h, w = left.shape
    xs, ys = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32))
    disp = 6.0 + 40.0 * ys / h + 8.0 * np.sin(xs / 180.0)
    right = cv2.remap(left, xs + disp, ys, cv2.INTER_LINEAR)
    left2 = cv2.warpAffine(left, np.float32([[1.03, 0.004, -12.0], [-0.004, 1.03, -4.0]]), (w, h))
    kp0, kp1, d0, d1, ms = ref.matching.extract_kps_descs_matches(left, right)
    assert d0.shape[1] == 61 and d0.dtype == np.uint8 and len(kp0) > 2000 and len(kp1) > 2000
    inl, outl = ref.matching.extract_inliers_outliers(kp0, kp1, ms)
    cq, ct, cd, _ = dm_arrays(ms)
    p0 = np.array([k.pt for k in kp0], np.float32)
    p1 = np.array([k.pt for k in kp1], np.float32)
    _, d2 = ref.matching.FEATURE.detectAndCompute(left2, None)
    d2 = d2[:2500]
    mq, mt, md, _ = dm_arrays(ref.matching.MATCHER.match(d0, d2))
    bq, bt, bd, _ = dm_arrays(ref.matching.MATCHER.match(d2, d0))
    knn = ref.matching.MATCHER.knnMatch(d0, d2, k=2)
    k_idx = np.array([[m.trainIdx for m in pr] for pr in knn], np.int32)
    k_dist = np.array([[int(m.distance) for m in pr] for pr in knn], np.int32)
    # the oracle restatement must agree with cv2 on real descriptors too
    oq, ot, od = ora.match_crosscheck(d0, d1)
    assert np.array_equal(oq, cq) and np.array_equal(ot, ct) and np.array_equal(od, cd.astype(np.int64))
    oi, odist = ora.match(d0, d2)
    assert np.array_equal(oi, mt) and np.array_equal(odist, md.astype(np.int64))
    ki, kd = ora.knn2(d0, d2)
    assert np.array_equal(ki, k_idx) and np.array_equal(kd, k_dist)
    o_in, o_out = ora.extract_inliers_outliers(p0, p1, cq, ct)
    assert np.array_equal(o_in, inl) and np.array_equal(o_out, outl)
    ties = int((k_dist[:, 0] == k_dist[:, 1]).sum())
    np.savez_compressed(os.path.join(OUT, "akaze_real.npz"), desc_l=d0, desc_r=d1, desc_l2=d2, pts_l=p0, pts_r=p1,
                        cross_q=cq, cross_t=ct, cross_d=cd, inliers=inl, outliers=outl, match_t=mt, match_d=md,
                        back_t=bt, back_d=bd, knn_idx=k_idx, knn_dist=k_dist, n_exact_ties=np.array(ties))
    print("  akaze_real:", d0.shape, d1.shape, d2.shape, "crossCheck matches", len(cq), "stereo inliers", len(inl),
          "top-2 ties", ties)


def golden_create_db_48(ref):
    """The UNMODIFIED reference's create_db + TrackingDB on 48 frames at the BENCH's keypoint counts
    (2000-5000 per image, synth.torch_sequence(seed=1), the bench workload's first frames).  The inputs are
    regenerated from the seed on the test side (CPU generator: deterministic), so only the reference's
    OUTPUTS are stored: per frame the stereo survivors, the forward matches, the RANSAC inlier flags
    (np.random seeded) and the track id of every feature."""
    import torch  # noqa: F401
    n = 48
    seq = synth.torch_sequence(n, first_frame=0, seed=1, device="cpu")
    frames = []
    for f in range(n):
        lo, k = int(seq["l_off"][f]), int(seq["n_l"][f])
        frames.append((seq["pts_l"][lo:lo + k].numpy(), seq["pts_r"][lo:lo + k].numpy(),
                       seq["desc_l"][lo:lo + k].numpy(), seq["desc_r"][lo:lo + k].numpy()))

    class Provider:
        def detectAndCompute(self, token, mask):
            side, f = token
            pts = frames[f][0 if side == "L" else 1]
            return tuple(cv2.KeyPoint(float(x), float(y), 1.0) for x, y in pts), frames[f][2 if side == "L" else 3]

    old_feature, old_reader = ref.matching.FEATURE, ref.inputs.read_images
    ref.matching.FEATURE = Provider()
    ref.inputs.read_images = lambda idx: (("L", idx), ("R", idx))
    calls = []
    try:
        db = ref.tracking_database.TrackingDB()
        real_add = db.add_frame
        db.add_frame = lambda links, left_features, matches_to_previous_left=None, inliers=None: (
            calls.append((links, left_features, matches_to_previous_left, inliers)),
            real_add(links, left_features, matches_to_previous_left, inliers))[1]
        np.random.seed(11)
        ref.database.create_db(start_frame=0, num_frames=n, db=db)
    finally:
        ref.matching.FEATURE, ref.inputs.read_images = old_feature, old_reader
    db._check_consistency()
    out = {"n_frames": np.array(n), "seed": np.array(1), "n_tracks": np.array(db.track_num()),
           "n_links_total": np.array(db.link_num()),
           "n_links": np.array([len(c[0]) for c in calls], np.int32),
           "inliers_percent": np.array([db.frameID_to_inliers_percent[f] for f in range(n)])}
    for f, (links, feats, ms, inl) in enumerate(calls):
        # link f,k sits on left keypoint link_src: recover it from the descriptor rows (features[is_valid])
        out[f"x_left{f}"] = np.array([l.x_left for l in links], np.float32)
        out[f"y{f}"] = np.array([l.y for l in links], np.float64)
        out[f"track_ids{f}"] = np.array(db.frameId_to_trackIds_list[f], np.int32)
        if ms is not None:
            out[f"match_t{f}"] = np.array([m.trainIdx for m in ms], np.uint16)
            out[f"match_d{f}"] = np.array([int(m.distance) for m in ms], np.uint16)
            out[f"inliers{f}"] = np.packbits(np.asarray(inl, dtype=bool))
    np.savez_compressed(os.path.join(OUT, "create_db_48.npz"), **out)
    print("  create_db_48:", n, "frames,", int(out["n_links"].sum()), "links,", db.track_num(), "tracks,",
          db.link_num(), "links on tracks")


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = refshim.load()
    golden_matching(ref)
    golden_stereo(ref)
    golden_triangulation(ref)
    golden_ransac(ref)
    golden_database(ref)
    golden_loop_candidates(golden_create_db(ref))
    golden_akaze(ref)
    golden_create_db_48(ref)
    print("golden vectors written to", OUT, "cv2", cv2.__version__, "numpy", np.__version__)
    for f in sorted(os.listdir(OUT)):
        print(" ", f, os.path.getsize(os.path.join(OUT, f)), "bytes")


if __name__ == "__main__":
    main()
