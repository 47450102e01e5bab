"""The batched device-resident pipeline against the oracle, frame by frame."""
import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("matcher_kernel")]


def make_sequence(rng, sizes):
    from slamfe import synth
    frames = []
    prev = None
    for n in sizes:
        dl, dr, pl, pr = synth.stereo_frame(rng, n)
        if prev is not None:  # temporal structure: re-flipped copies of the previous left frame
            dl = synth.next_frame_descriptors(rng, prev, n)
            dr, src = synth.paired_descriptors(rng, dl)
            has = src >= 0
            pr[has, 0] = pl[src[has], 0] - rng.uniform(2.5, 120, has.sum()).astype(np.float32)
            pr[has, 1] = pl[src[has], 1] + rng.normal(0, 0.5, has.sum()).astype(np.float32)
        frames.append((dl, dr, pl, pr))
        prev = dl
    return frames


def test_sequence_pipeline_vs_oracle(slamfe, oracle):
    import torch
    from slamfe import frontend, ops, ransac
    rng = np.random.default_rng(71)
    frames = make_sequence(rng, [900, 1200, 33, 1500, 800, 1000])
    seq = frontend.pack_sequence(frames)
    ds = frontend.to_device(seq)
    fe = frontend.FrontEnd()
    out = fe.run(ds)
    out = fe.run(ds)  # buffers are reused: a second pass must give the same tables
    host, nbytes, _ = frontend.results_to_host(out)
    assert nbytes > 0
    feats, links_ref = [], []
    for f, (dl, dr, pl, pr) in enumerate(frames):
        cq, ct, _ = oracle.match_crosscheck(dl, dr)
        inl, _ = oracle.extract_inliers_outliers(pl, pr, cq, ct)
        valid, links = oracle.create_links(pl, pr, cq[inl], ct[inl])
        feats.append(dl[valid]); links_ref.append(links)
        lo, k = seq.l_off[f], host["n_links"][f]
        assert k == len(inl) and host["n_matches"][f] == len(cq)
        assert np.array_equal(host["link_src"][lo:lo + k], cq[inl])
        xyz = oracle.triangulate_links(links, ransac.P, ransac.Q)
        got = host["xyz"][lo:lo + k].astype(np.float64)
        rel = np.linalg.norm(got - xyz, axis=1) / np.linalg.norm(xyz, axis=1)
        assert rel.max() < 1e-5  # fp32 pipeline tolerance (north_star)
    for f in range(len(frames) - 1):
        lo, k = seq.l_off[f], len(feats[f])
        fi, fd = ops.keys_to_numpy(host["fwd_keys"][lo:lo + k])
        oi, od = oracle.match(feats[f], feats[f + 1])  # the pipeline keeps the best neighbour only
        assert fi.shape == (k,) and np.array_equal(fi, oi) and np.array_equal(fd, od)   # compact keys: one per row
        lo1, k1 = seq.l_off[f + 1], len(feats[f + 1])
        obi, obd = oracle.match(feats[f + 1], feats[f])
        got_b = ops.keys_to_numpy(host["bwd_keys"][lo1:lo1 + k1])
        assert np.array_equal(got_b[0], obi) and np.array_equal(got_b[1], obd)
    pairs = frontend.descriptor_pairs(seq.n_l, seq.n_r, host["n_links"])
    assert pairs == sum(len(a) * len(b) for a, b, _, _ in frames) + sum(
        len(feats[f]) * len(feats[f + 1]) for f in range(len(frames) - 1))


@pytest.mark.parametrize("chunk", [1, 2, 4, 100])
def test_host_pipeline_equals_resident_run(slamfe, chunk):
    """run_host (chunked, copies overlapped on three streams) must give the tables of run()."""
    import torch
    from slamfe import frontend
    rng = np.random.default_rng(72)
    frames = make_sequence(rng, [700, 900, 40, 1100, 650, 800, 500])
    seq = frontend.pack_sequence(frames)
    fe = frontend.FrontEnd()
    ref, _, _ = frontend.results_to_host(fe.run(frontend.to_device(seq)), keys=frontend.ALL_RESULT_KEYS)
    ref = {k: v.copy() for k, v in ref.items()}
    fe2 = frontend.FrontEnd()
    for _ in range(2):  # second pass reuses every buffer
        got, h2d, d2h = fe2.run_host(seq, chunk_frames=chunk, keys=frontend.ALL_RESULT_KEYS)
        assert h2d >= seq.h2d_bytes() and d2h > 0
        for f in range(seq.n_frames):
            lo, k = seq.l_off[f], ref["n_links"][f]
            assert got["n_links"][f] == k and got["n_matches"][f] == ref["n_matches"][f]
            n = seq.n_l[f]
            assert np.array_equal(got["match_t"][lo:lo + n], ref["match_t"][lo:lo + n])
            for key in ("link_src", "links", "xyz", "fwd_keys", "bwd_keys"):
                assert np.array_equal(got[key][lo:lo + k], ref[key][lo:lo + k]), (key, f)
            # `links` is not in the default D2H set: the host derives it, bit for bit
            assert np.array_equal(frontend.links_from_tables(seq, got, f), ref["links"][lo:lo + k]), f
    assert "links" not in fe2.run_host(seq, chunk_frames=chunk)[0]
    torch.cuda.synchronize()
    # with the tracking stages: same samples (RNG keyed by global pair index), same tables
    trk = fe.track(frontend.to_device(seq), h_max=40, seed=9)
    ref_t = {k: trk[k].cpu().numpy() for k in frontend.TRACK_KEYS}
    got, _, _ = fe2.run_host(seq, chunk_frames=chunk, track=True, h_max=40, seed=9)
    n_pairs = seq.n_frames - 1
    for key in ("best", "n_good", "n_hyp"):
        assert np.array_equal(got[key][:n_pairs], ref_t[key][:n_pairs]), key
    for f in range(seq.n_frames):
        lo, k = seq.l_off[f], ref["n_links"][f]
        assert np.array_equal(got["inlier_fwd"][lo:lo + k], ref_t["inlier_fwd"][lo:lo + k]), f


def test_patch_rebinds_reference_style_modules(slamfe):
    """patch() swaps by-value imports of a reference-shaped module tree (the real tree is only
    present in the build container)."""
    import types
    from slamfe import patch
    mods = {}
    for name in ("final_project.algorithms.matching", "final_project.backend.database.database",
                 "final_project.algorithms.ransac", "final_project.backend.loop.loop_closure"):
        m = types.ModuleType(name)
        for a in ("MATCHER", "MATCHER_LEFT_RIGHT", "extract_inliers_outliers", "ransac_pnp_for_tracking_db",
                  "ransac_pnp", "triangulate_links", "transformation_agreement"):
            setattr(m, a, "reference")
        mods[name] = m
    tok = patch.patch(mods)
    assert mods["final_project.backend.database.database"].MATCHER is mods["final_project.algorithms.matching"].MATCHER
    assert type(mods["final_project.backend.loop.loop_closure"].MATCHER).__name__ == "Matcher"
    assert callable(mods["final_project.algorithms.ransac"].transformation_agreement)
    patch.unpatch(tok)
    assert mods["final_project.algorithms.ransac"].transformation_agreement == "reference"


OTHER_CALIB = (np.array([[707.0912, 0.0, 601.8873], [0.0, 707.0912, 183.1104], [0.0, 0.0, 1.0]]),
               np.hstack([np.eye(3), np.zeros((3, 1))]),
               np.hstack([np.eye(3), np.array([[-0.4731], [0.0], [0.0]])]))  # not KITTI-00: other fx, cx, cy, baseline


@pytest.fixture(params=["kitti00", "other_calib"])
def cameras(request):
    """The cameras of the drop-in entry points: the default (KITTI 00) or another calibration installed the
    way patch() does it (ransac.set_cameras); FrontEnd() / create_db must follow."""
    from slamfe import ransac
    old = (ransac.K, ransac.M1, ransac.M2)
    if request.param == "other_calib":
        ransac.set_cameras(*OTHER_CALIB)
    yield request.param
    ransac.set_cameras(*old)


def test_tracking_stages_vs_oracle(slamfe, oracle, cameras):
    """FrontEnd.track: mutual check, link gather, fp64 triangulation, per-pair iteration counts and the
    scoring of the device-generated hypotheses, against the oracle's restatement of
    database.py:54-85 / ransac.py:59-113 on the same frames — for the default cameras and for a
    calibration installed through ransac.set_cameras (what patch() does with the reference's)."""
    from slamfe import frontend, ransac
    rng = np.random.default_rng(73)
    frames = make_sequence(rng, [800, 1000, 30, 1200, 700])
    seq = frontend.pack_sequence(frames)
    ds = frontend.to_device(seq)
    fe = frontend.FrontEnd()
    assert np.array_equal(fe.K, ransac.K) and np.array_equal(fe.P, ransac.K @ ransac.M1)
    if cameras == "other_calib":
        assert not np.array_equal(fe.K, frontend.FrontEnd(*__import__("slamfe").utils.read_cameras()).K)
    H = 48
    out = fe.track(ds, h_max=H, seed=5)
    out = fe.track(ds, h_max=H, seed=5)      # buffers reused
    g = {k: v.cpu().numpy() for k, v in out.items() if k in ("good_j", "good_t", "n_good", "n_hyp", "pts", "lpix", "rpix", "T",
                                                            "hyp_valid", "counts", "best", "best_mask", "inlier_fwd",
                                                            "n_links", "n_matches")}
    feats, links = [], []
    for dl, dr, pl, pr in frames:
        cq, ct, _ = oracle.match_crosscheck(dl, dr)
        inl, _ = oracle.extract_inliers_outliers(pl, pr, cq, ct)
        valid, ln = oracle.create_links(pl, pr, cq[inl], ct[inl])
        feats.append(dl[valid]); links.append(np.asarray(ln, dtype=np.float64).reshape(-1, 3))
    K, M1, M2 = ransac.K, ransac.M1, ransac.M2
    xyz_tab = out["xyz"].cpu().numpy().astype(np.float64)
    for f in range(len(frames)):   # the fp32 triangulation stage uses the same cameras
        if len(links[f]):
            ref = oracle.triangulate_links(links[f], K @ M1, K @ M2)
            got = xyz_tab[seq.l_off[f]:seq.l_off[f] + len(links[f])]
            assert (np.linalg.norm(got - ref, axis=1) / np.linalg.norm(ref, axis=1)).max() < 1e-5
    for f in range(len(frames) - 1):
        lo = seq.l_off[f]
        fi, fd, good = oracle.mutual_forward_backward(feats[f], feats[f + 1])
        n = len(good)
        assert g["n_good"][f] == n
        assert np.array_equal(g["good_j"][lo:lo + n], good) and np.array_equal(g["good_t"][lo:lo + n], fi[good])
        cur = links[f + 1][fi[good]]
        assert np.array_equal(g["lpix"][lo:lo + n], cur[:, [0, 2]]) and np.array_equal(g["rpix"][lo:lo + n], cur[:, [1, 2]])
        if n:
            ref = oracle.triangulate_links(links[f][good], ransac.P, ransac.Q)
            rel = np.linalg.norm(g["pts"][lo:lo + n] - ref, axis=1) / np.linalg.norm(ref, axis=1)
            assert rel.max() < 1e-11
        pct = 100 * (g["n_links"][f + 1] / g["n_matches"][f + 1])
        assert g["n_hyp"][f] == min(H, ransac.calc_ransac_iteration(pct))
        # scoring of the generated hypotheses: bit-exact against the oracle on the same T and points
        T = g["T"][f * H:(f + 1) * H]; ok = g["hyp_valid"][f * H:(f + 1) * H].astype(bool)
        assert not ok[g["n_hyp"][f]:].any()
        if n >= 4:
            assert ok[:g["n_hyp"][f]].sum() >= 0.4 * g["n_hyp"][f]  # random geometry: many samples admit no pose
        pts, lp, rp = g["pts"][lo:lo + n], g["lpix"][lo:lo + n], g["rpix"][lo:lo + n]
        counts = np.zeros(H, np.int64)
        masks = {}
        for h in np.nonzero(ok)[0]:
            masks[h] = oracle.transformation_agreement(T[h], pts, lp, rp, K, M1, M2)
            counts[h] = masks[h].sum()
        assert np.array_equal(g["counts"][f], counts)
        best = int(np.argmax(counts)) if counts.max() > 0 else -1
        assert g["best"][f, 0] == best and g["best"][f, 1] == counts.max()
        flags = np.zeros(len(feats[f]), np.uint8)
        if best >= 0:
            assert np.array_equal(g["best_mask"][lo:lo + n].astype(bool), masks[best])
            flags[good[masks[best]]] = 1
        else:
            flags[good] = 1                     # good_idx[None] quirk, database.py:82
        assert np.array_equal(g["inlier_fwd"][lo:lo + len(feats[f])], flags)


def test_full_ransac_does_not_depend_on_h_max(slamfe):
    """h_max only sizes the batched RANSAC launch (csrc/tracking.cu caps n_hyp at it); the reference
    always runs calc_ransac_iteration in full (ransac.py:59-67,94).  With full_ransac the truncated pairs
    are re-run at their full count from the same (seed, pair, hypothesis) samples, so a tiny h_max gives
    exactly the tables of an h_max that truncates nothing."""
    from slamfe import frontend, ransac
    rng = np.random.default_rng(76)
    frames = make_sequence(rng, [700, 900, 650, 800, 750])
    seq = frontend.pack_sequence(frames)
    n_pairs = seq.n_frames - 1
    big = frontend.FrontEnd()
    ref = big.track(frontend.to_device(seq), h_max=256, seed=4)
    ref = {k: ref[k].cpu().numpy() for k in ("best", "inlier_fwd", "n_hyp", "n_hyp_full", "n_links", "n_matches", "n_good")}
    assert (ref["n_hyp_full"][:n_pairs] == ref["n_hyp"][:n_pairs]).all() and big.last_truncated == 0
    for f in range(n_pairs):
        pct = 100 * (ref["n_links"][f + 1] / ref["n_matches"][f + 1])
        assert ref["n_hyp_full"][f] == ransac.calc_ransac_iteration(pct) > 8
    small = frontend.FrontEnd()
    cut = small.track(frontend.to_device(seq), h_max=8, seed=4)
    assert (cut["n_hyp"][:n_pairs].cpu().numpy() == 8).all()
    assert np.array_equal(cut["n_hyp_full"][:n_pairs].cpu().numpy(), ref["n_hyp_full"][:n_pairs])
    got = small.track(frontend.to_device(seq), h_max=8, seed=4, full_ransac=True)
    assert small.last_truncated == n_pairs
    assert np.array_equal(got["best"][:n_pairs].cpu().numpy(), ref["best"][:n_pairs])
    assert np.array_equal(got["inlier_fwd"].cpu().numpy(), ref["inlier_fwd"])
    # the host pipeline: default full_ransac=True patches the host tables too; False only reports
    host, _, _ = frontend.FrontEnd().run_host(seq, chunk_frames=2, track=True, h_max=8, seed=4)
    assert np.array_equal(host["best"][:n_pairs], ref["best"][:n_pairs])
    assert np.array_equal(host["inlier_fwd"], ref["inlier_fwd"])
    fe3 = frontend.FrontEnd()
    fe3.run_host(seq, chunk_frames=2, track=True, h_max=8, seed=4, full_ransac=False)
    assert fe3.last_truncated == n_pairs


def test_create_db_raises_like_the_reference_on_tiny_pairs(slamfe, oracle):
    """A frame pair with fewer than 4 mutual matches: the reference dies in np.random.choice(n, 4,
    replace=False) with ValueError (ransac.py:95); the batched builder must not silently flag every
    match as an inlier instead."""
    from slamfe import database as sdb, synth
    rng = np.random.default_rng(77)
    dl, dr, pl, pr = synth.stereo_frame(rng, 500)
    cq, ct, _ = oracle.match_crosscheck(dl, dr)
    inl, _ = oracle.extract_inliers_outliers(pl, pr, cq, ct)
    keep_l, keep_r = cq[inl][:3], ct[inl][:3]          # second frame: exactly 3 stereo links
    small = (pl[keep_l], pr[keep_r], dl[keep_l], dr[keep_r])

    class FakeDB:
        frameID_to_inliers_percent = {}

        def add_frame(self, *a, **k):
            pass

    with pytest.raises(ValueError, match="larger sample than population"):
        sdb.create_db([(pl, pr, dl, dr), small], FakeDB(), h_max=16)


def test_tracking_edge_cases(slamfe):
    """Empty frames, frames with fewer than 4 links, a single-frame sequence: no crash, invalid
    pairs report best = -1 / zero flags, and the chunked host pipeline agrees with the resident one."""
    from slamfe import frontend, synth
    rng = np.random.default_rng(74)
    empty = (np.zeros((0, 61), np.uint8), np.zeros((0, 61), np.uint8), np.zeros((0, 2), np.float32),
             np.zeros((0, 2), np.float32))
    frames = [empty, synth.stereo_frame(rng, 40), synth.stereo_frame(rng, 3), empty, synth.stereo_frame(rng, 500),
              synth.stereo_frame(rng, 450)]
    seq = frontend.pack_sequence(frames)
    fe = frontend.FrontEnd()
    out = fe.track(frontend.to_device(seq), h_max=16, seed=3)
    host = {k: out[k].cpu().numpy() for k in ("n_links", "n_matches", "n_good", "n_hyp", "best", "inlier_fwd", "hyp_valid")}
    assert host["n_links"][0] == 0 and host["n_links"][3] == 0 and host["n_matches"][0] == 0
    n_pairs = len(frames) - 1
    for f in range(n_pairs):
        if host["n_good"][f] < 4:
            assert host["best"][f, 0] == -1 and host["best"][f, 1] == 0
            assert not host["hyp_valid"][f * 16:(f + 1) * 16].any()
    assert host["n_hyp"][2] == 0            # frame 3 has no stereo matches: no iteration count
    got, _, _ = frontend.FrontEnd().run_host(seq, chunk_frames=2, track=True, h_max=16, seed=3)
    assert np.array_equal(got["best"][:n_pairs], host["best"][:n_pairs])
    assert np.array_equal(got["n_good"][:n_pairs], host["n_good"][:n_pairs])
    for f in range(len(frames)):
        lo, k = seq.l_off[f], host["n_links"][f]
        assert np.array_equal(got["inlier_fwd"][lo:lo + k], host["inlier_fwd"][lo:lo + k])
    one = frontend.pack_sequence([synth.stereo_frame(rng, 300)])
    o1 = frontend.FrontEnd().track(frontend.to_device(one), h_max=8)
    assert int(o1["n_links"][0]) > 0
    g1, _, _ = frontend.FrontEnd().run_host(one, track=True, h_max=8)
    assert g1["n_links"][0] == int(o1["n_links"][0])


def test_create_db_adapter_on_the_gpu(slamfe, oracle):
    """slamfe.database.create_db: the add_frame arguments built from the GPU tables equal the oracle's
    restatement of database.py:12-27,54-66 (links, features, forward matches); the conversion itself is
    checked against the real reference TrackingDB by tests/test_reference_db.py on the CPU."""
    from slamfe import database as sdb

    class FakeDB:
        def __init__(self):
            self.calls, self.frameID_to_inliers_percent = [], {}

        def add_frame(self, links, left_features, matches_to_previous_left=None, inliers=None):
            self.calls.append((links, left_features, matches_to_previous_left, inliers))

    rng = np.random.default_rng(75)
    frames = make_sequence(rng, [600, 750, 500, 640])
    db = sdb.create_db([(pl, pr, dl, dr) for dl, dr, pl, pr in frames], FakeDB(), chunk_frames=2, h_max=32)
    assert len(db.calls) == len(frames)
    prev_feat = None
    for f, ((dl, dr, pl, pr), (links, feats, ms, inl)) in enumerate(zip(frames, db.calls)):
        cq, ct, _ = oracle.match_crosscheck(dl, dr)
        ii, _ = oracle.extract_inliers_outliers(pl, pr, cq, ct)
        valid, ref_links = oracle.create_links(pl, pr, cq[ii], ct[ii])
        assert np.array_equal(np.array([(l.x_left, l.x_right, l.y) for l in links]).reshape(-1, 3),
                              np.asarray(ref_links, dtype=np.float64).reshape(-1, 3))
        assert np.array_equal(feats, dl[valid])
        assert db.frameID_to_inliers_percent[f] == 100 * (len(ii) / len(cq))
        if f == 0:
            assert ms is None and inl is None
        else:
            fi, fd = oracle.match(prev_feat, dl[valid])
            assert [m.queryIdx for m in ms] == list(range(len(prev_feat))) and all(m.imgIdx == 0 for m in ms)
            assert np.array_equal([m.trainIdx for m in ms], fi) and np.array_equal([int(m.distance) for m in ms], fd)
            assert inl.dtype == bool and len(inl) == len(ms)
        prev_feat = dl[valid]


def test_create_db_against_the_reference_golden(slamfe, golden):
    """tests/golden/create_db.npz holds every add_frame call of the UNMODIFIED reference's create_db on 5
    synthetic frames (oracle/make_golden.py: golden_create_db).  slamfe.database.create_db on the same
    frames must pass identical links, features and forward matches; the inlier flags come from a
    randomised RANSAC on both sides (the reference's run is seeded, ~20-60 iterations) and must agree
    as consensus sets."""
    from slamfe import database as sdb
    g = golden("create_db")
    n = int(g["n_frames"])
    frames = [(g[f"pts_l{f}"], g[f"pts_r{f}"], g[f"desc_l{f}"], g[f"desc_r{f}"]) for f in range(n)]

    class FakeDB:
        def __init__(self):
            self.calls, self.frameID_to_inliers_percent = [], {}

        def add_frame(self, links, left_features, matches_to_previous_left=None, inliers=None):
            self.calls.append((links, left_features, matches_to_previous_left, inliers))

    db = sdb.create_db(frames, FakeDB(), chunk_frames=3, h_max=128, seed=2)
    assert len(db.calls) == n
    jac = []
    for f, (links, feats, ms, inl) in enumerate(db.calls):
        assert np.array_equal(np.array([(l.x_left, l.x_right, l.y) for l in links], np.float64).reshape(-1, 3),
                              g[f"links{f}"])
        assert np.array_equal(feats, g[f"features{f}"])
        assert db.frameID_to_inliers_percent[f] == float(g[f"inliers_percent{f}"])
        if f == 0:
            assert ms is None
            continue
        assert np.array_equal([m.trainIdx for m in ms], g[f"match_t{f}"])
        assert np.array_equal(np.array([m.distance for m in ms], np.float32), g[f"match_d{f}"])
        a, b = np.asarray(inl, bool), g[f"inliers{f}"]
        assert a.shape == b.shape
        # the reference's seeded run missed the consensus on one pair (3 inliers); the GPU run (128
        # exact minimal solutions instead of ~25 EPnP ones) must find at least what the reference found
        assert a.sum() >= 0.8 * b.sum(), (f, int(a.sum()), int(b.sum()))
        if b.sum() >= 20:  # ... and its consensus set contains the reference's
            jac.append((a & b).sum() / b.sum())
    assert len(jac) >= 3 and min(jac) >= 0.75, jac


def test_batched_create_db_through_patch(slamfe, golden):
    """patch(batched_db=True) on a reference-shaped module tree (the real tree is not on the GPU box; the
    wiring against the real one is tested on the CPU by tests/test_reference_db.py): `database.create_db`
    reads / describes the frames through the tree's Inputs + FEATURE and builds the DB with the GPU
    pipeline — same add_frame arguments as the reference's own run frozen in tests/golden/create_db.npz."""
    import types
    import cv2
    from slamfe import patch
    g = golden("create_db")
    n = int(g["n_frames"])

    class Provider:
        def detectAndCompute(self, token, mask):
            side, f = token
            pts = g[f"pts_{'l' if side == 'L' else 'r'}{f}"]
            return tuple(cv2.KeyPoint(float(x), float(y), 1.0) for x, y in pts), g[f"desc_{'l' if side == 'L' else 'r'}{f}"]

    class TrackingDB:
        def __init__(self):
            self.calls, self.frameID_to_inliers_percent = [], {}

        def add_frame(self, links, left_features, matches_to_previous_left=None, inliers=None):
            self.calls.append((links, left_features, matches_to_previous_left, inliers))

    matching = types.ModuleType("final_project.algorithms.matching")
    matching.FEATURE = Provider()
    database = types.ModuleType("final_project.backend.database.database")
    database.Inputs = types.SimpleNamespace(read_images=lambda idx: (("L", idx), ("R", idx)))
    database.TrackingDB = TrackingDB
    database.create_db = lambda start_frame=0, num_frames=200, db=None: "reference loop"
    mods = {matching.__name__: matching, database.__name__: database}
    tok = patch.patch(mods, batched_db=True, h_max=64, chunk_frames=2)
    try:
        db = database.create_db(num_frames=n)
        assert database.create_db(start_frame=2, num_frames=n, db=db) == "reference loop"   # resumed build
    finally:
        patch.unpatch(tok)
    assert len(db.calls) == n
    for f, (links, feats, ms, inl) in enumerate(db.calls):
        assert np.array_equal(np.array([(l.x_left, l.x_right, l.y) for l in links], np.float64).reshape(-1, 3),
                              g[f"links{f}"])
        assert np.array_equal(feats, g[f"features{f}"])
        if f:
            assert np.array_equal([m.trainIdx for m in ms], g[f"match_t{f}"]) and len(inl) == len(ms)


def _bench_frames(n, seed=1):
    from slamfe import synth
    seq = synth.torch_sequence(n, first_frame=0, seed=seed, device="cpu")
    frames = []
    for f in range(n):
        lo, k = int(seq["l_off"][f]), int(seq["n_l"][f])
        frames.append((seq["desc_l"][lo:lo + k].numpy(), seq["desc_r"][lo:lo + k].numpy(),
                       seq["pts_l"][lo:lo + k].numpy(), seq["pts_r"][lo:lo + k].numpy()))
    return frames


def test_48_frames_at_bench_size_against_the_reference_run(slamfe, golden):
    """tests/golden/create_db_48.npz: the unmodified reference's create_db + TrackingDB on the bench
    workload's first 48 frames (2000-5000 keypoints per image).  The pipeline's stereo survivors and forward
    matches are identical; slamfe_track_ids fed the REFERENCE's inlier flags reproduces its track ids
    exactly; the pipeline's own RANSAC consensus (different sampler and solver) contains most of the
    reference's and builds a consistent store of similar size."""
    import torch
    from slamfe import frontend, ops, trackdb
    g = golden("create_db_48")
    F = int(g["n_frames"])
    seq = frontend.pack_sequence(_bench_frames(F, int(g["seed"])))
    fe = frontend.FrontEnd()
    tables, _, _ = fe.run_host(seq, chunk_frames=16, track=True, h_max=128, seed=3, track_ids=True)
    n_links = g["n_links"]
    assert np.array_equal(tables["n_links"][:F], n_links)
    ref_flags = np.zeros(seq.desc_l.shape[0], np.uint8)
    jac = []
    for f in range(F):
        lo, k = int(seq.l_off[f]), int(n_links[f])
        src = tables["link_src"][lo:lo + k]
        assert np.array_equal(seq.pts_l[lo + src, 0], g[f"x_left{f}"])             # the same links in the same order
        assert np.allclose(frontend.links_from_tables(seq, tables, f)[:, 2], g[f"y{f}"], rtol=0, atol=1e-4)
        assert abs(100 * k / tables["n_matches"][f] - g["inliers_percent"][f]) < 1e-9
        if f + 1 < F:
            idx, dist = ops.keys_to_numpy(tables["fwd_keys"][lo:lo + k])
            assert np.array_equal(idx, g[f"match_t{f + 1}"]) and np.array_equal(dist, g[f"match_d{f + 1}"])
            want = np.unpackbits(g[f"inliers{f + 1}"])[:k].astype(bool)
            got = tables["inlier_fwd"][lo:lo + k].astype(bool)
            ref_flags[lo:lo + k] = want
            jac.append((want & got).sum() / max(1, (want | got).sum()))
    assert np.median(jac) > 0.5, np.median(jac)   # two independent ~60-iteration RANSAC runs (measured 0.66)
    # the reference's flags through the device kernel -> the reference's track ids, exactly
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    tid, n_tr, _ = ops.track_ids(dev(tables["fwd_keys"]), dev(ref_flags), dev(seq.l_off), dev(tables["n_links"]), F)
    tid = tid.cpu().numpy()
    assert int(n_tr.item()) == int(g["n_tracks"])
    for f in range(F):
        lo = int(seq.l_off[f])
        assert np.array_equal(tid[lo:lo + int(n_links[f])], g[f"track_ids{f}"]), f
        assert (tid[lo + int(n_links[f]):int(seq.l_off[f + 1])] == -1).all()
    # the pipeline's own store
    db = trackdb.build(seq, tables)
    assert db.check_consistency() and db.frame_num() == F
    # P3P hypotheses reach larger consensus sets than the reference's 4-point EPnP run: more, not fewer, tracks
    assert 0.9 * int(g["n_tracks"]) < db.track_num() < 1.4 * int(g["n_tracks"])
    host_ids = trackdb.build(seq, {k: v for k, v in tables.items() if k not in ("track_id", "n_tracks")})
    assert host_ids == db                                     # device ids == host restatement
    db2 = trackdb.create_db(seq, chunk_frames=100, h_max=128, seed=3)
    assert db2 == db                                          # chunking does not matter


def test_track_ids_edge_cases(slamfe):
    """No frames, one frame, frames without links, a pair without inliers: ids stay NO_ID, counts are 0."""
    import torch
    from slamfe import frontend, trackdb
    rng = np.random.default_rng(78)
    frames = make_sequence(rng, [600])
    out = frontend.FrontEnd().track(frontend.to_device(frontend.pack_sequence(frames)), track_ids=True)
    assert int(out["n_tracks"].item()) == 0
    frames = make_sequence(rng, [500, 30, 520, 510])
    seq = frontend.pack_sequence(frames)
    fe = frontend.FrontEnd()
    out = fe.track(frontend.to_device(seq), h_max=32, seed=2, track_ids=True)
    tid = out["track_id"].cpu().numpy()
    fwd = out["fwd_keys"].cpu().numpy()
    inl = out["inlier_fwd"].cpu().numpy()
    n_links = out["n_links"].cpu().numpy()
    ids, n_tracks = trackdb.track_ids_host(
        [fwd[int(seq.l_off[f]):int(seq.l_off[f]) + int(n_links[f])].view(np.uint32) & 0x3FFFFF for f in range(3)],
        [inl[int(seq.l_off[f]):int(seq.l_off[f]) + int(n_links[f])] for f in range(3)], n_links)
    assert int(out["n_tracks"].item()) == n_tracks
    for f in range(4):
        assert np.array_equal(tid[int(seq.l_off[f]):int(seq.l_off[f]) + int(n_links[f])], ids[f])
