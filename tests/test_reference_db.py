"""The batched create_db adapter (slamfe.database) against the UNMODIFIED reference's create_db +
TrackingDB (backend/database/database.py:30-89, tracking_database.py), run through oracle/refshim.py.

CPU only; skipped where /root/reference is absent (the GPU box).  The reference runs with its own
cv2 path on synthetic frames (its image reader and AKAZE detector are replaced by a provider of
synthetic keypoints / descriptors — inputs, not code under test); every `add_frame` call is
recorded.  The adapter is fed host tables in the exact format FrontEnd.run_host(track=True) returns,
built here from the oracle (the product's CUDA path cannot run in this container; the same tables
are checked against the oracle on the GPU by tests/test_gpu_frontend.py), with the reference's own
inlier flags, and must drive a fresh reference TrackingDB into an identical state."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402

pytestmark = pytest.mark.skipif(not refshim.available(), reason="reference tree not present")


def _synthetic_frames(n):
    from slamfe import synth
    seq = synth.torch_sequence(n, first_frame=3, seed=5, device="cpu", lo=150, hi=260)
    frames = []
    for f in range(n):
        lo, k = int(seq["l_off"][f]), int(seq["n_l"][f])
        frames.append((seq["pts_l"][lo:lo + k].numpy(), seq["pts_r"][lo:lo + k].numpy(),
                       seq["desc_l"][lo:lo + k].numpy(), seq["desc_r"][lo:lo + k].numpy()))
    return frames


def _oracle_tables(seq, frames, oracle, inlier_flags):
    """Host tables in FrontEnd.run_host(track=True) format from the oracle."""
    from slamfe import _cabi
    L, F = seq.desc_l.shape[0], seq.n_frames
    t = {"match_t": np.full(L, -1, np.int32), "n_matches": np.zeros(F, np.int32), "n_links": np.zeros(F, np.int32),
         "link_src": np.full(L, -1, np.int32), "fwd_keys": np.full(L, -1, np.int32),
         "inlier_fwd": np.zeros(L, np.uint8)}
    feats = []
    for f, (pl, pr, dl, dr) in enumerate(frames):
        lo = int(seq.l_off[f])
        cq, ct, _ = oracle.match_crosscheck(dl, dr)
        inl, _ = oracle.extract_inliers_outliers(pl, pr, cq, ct)
        t["match_t"][lo + cq] = ct
        t["n_matches"][f], t["n_links"][f] = len(cq), len(inl)
        t["link_src"][lo:lo + len(inl)] = cq[inl]
        feats.append(dl[cq[inl]])
    for f in range(F - 1):
        lo, k = int(seq.l_off[f]), len(feats[f])
        fi, fd = oracle.match(feats[f], feats[f + 1])
        t["fwd_keys"][lo:lo + k] = ((fd.astype(np.uint32) << _cabi.KEY_IDX_BITS) | fi.astype(np.uint32)).view(np.int32)
        t["inlier_fwd"][lo:lo + k] = inlier_flags[f + 1]
    return t


def test_adapter_drives_the_reference_tracking_db_identically(oracle):
    import cv2
    ref = refshim.load()
    from slamfe import database as sdb
    frames = _synthetic_frames(5)

    class Provider:  # stands in for cv2.AKAZE: keypoints / descriptors of the synthetic frame
        def detectAndCompute(self, token, mask):
            side, f = token
            pts = frames[f][0 if side == "L" else 1]
            desc = frames[f][2 if side == "L" else 3]
            return tuple(cv2.KeyPoint(float(x), float(y), 1.0) for x, y in pts), desc

    old_feature, old_reader = ref.matching.FEATURE, ref.inputs.read_images
    ref.matching.FEATURE = Provider()
    ref.inputs.read_images = lambda idx: (("L", idx), ("R", idx))
    calls = []
    try:
        db_ref = ref.tracking_database.TrackingDB()
        real_add = db_ref.add_frame

        def spy(links, left_features, matches_to_previous_left=None, inliers=None):
            calls.append((links, left_features, matches_to_previous_left, inliers))
            return real_add(links, left_features, matches_to_previous_left, inliers)

        db_ref.add_frame = spy
        np.random.seed(3)
        ref.database.create_db(start_frame=0, num_frames=len(frames), db=db_ref)
    finally:
        ref.matching.FEATURE, ref.inputs.read_images = old_feature, old_reader
    assert len(calls) == len(frames)

    seq = sdb.pack_frames([(pl, pr, dl, dr) for pl, pr, dl, dr in frames], pin=False)
    flags = [None] + [np.asarray(c[3], dtype=bool) for c in calls[1:]]
    tables = _oracle_tables(seq, frames, oracle, flags)
    db_new = ref.tracking_database.TrackingDB()
    replay = []
    real_add2 = db_new.add_frame
    db_new.add_frame = lambda links, left_features, matches_to_previous_left=None, inliers=None: (
        replay.append((links, left_features, matches_to_previous_left, inliers)),
        real_add2(links, left_features, matches_to_previous_left, inliers))[1]
    for fr in sdb.frames_from_tables(seq, tables):  # what create_db() does after run_host()
        links = [ref.tracking_database.Link(float(a), float(b), float(c)) for a, b, c in fr["links"]]
        db_new.frameID_to_inliers_percent[fr["frame"]] = fr["inliers_percent"]
        if fr["frame"] == 0:
            db_new.add_frame(links=links, left_features=fr["features"], matches_to_previous_left=None, inliers=None)
            continue
        n = len(fr["fwd_idx"])
        ms = np.empty(n, dtype=object)
        ms[:] = list(map(cv2.DMatch, range(n), fr["fwd_idx"].tolist(), [0] * n, fr["fwd_dist"].tolist()))
        db_new.add_frame(links, fr["features"], ms, fr["inliers"])

    # every add_frame argument equals the reference's own
    for (l0, f0, m0, i0), (l1, f1, m1, i1) in zip(calls, replay):
        assert [(a.x_left, a.x_right, a.y) for a in l0] == [(a.x_left, a.x_right, a.y) for a in l1]
        assert np.array_equal(f0, f1)
        if m0 is None:
            assert m1 is None
        else:
            assert [(m.queryIdx, m.trainIdx, m.imgIdx, m.distance) for m in m0] == \
                   [(m.queryIdx, m.trainIdx, m.imgIdx, m.distance) for m in m1]
            assert np.array_equal(np.asarray(i0, bool), np.asarray(i1, bool))
    # ... and so does the resulting database
    assert db_ref.frameID_to_inliers_percent == db_new.frameID_to_inliers_percent
    assert db_ref.trackId_to_frames == db_new.trackId_to_frames and len(db_ref.trackId_to_frames) > 20
    assert db_ref.frameId_to_trackIds_list == db_new.frameId_to_trackIds_list
    assert sorted(db_ref.linkId_to_link) == sorted(db_new.linkId_to_link)
    for k, ln in db_ref.linkId_to_link.items():
        other = db_new.linkId_to_link[k]
        assert (ln.x_left, ln.x_right, ln.y) == (other.x_left, other.x_right, other.y)
    db_new._check_consistency() if hasattr(db_new, "_check_consistency") else None


def _reference_run(ref, frames, seed=3):
    """The unmodified reference's create_db + TrackingDB on synthetic frames; returns (db, add_frame calls)."""
    import cv2

    class Provider:
        def detectAndCompute(self, token, mask):
            side, f = token
            pts = frames[f][0 if side == "L" else 1]
            desc = frames[f][2 if side == "L" else 3]
            return tuple(cv2.KeyPoint(float(x), float(y), 1.0) for x, y in pts), desc

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
        np.random.seed(seed)
        ref.database.create_db(start_frame=0, num_frames=len(frames), db=db)
    finally:
        ref.matching.FEATURE, ref.inputs.read_images = old_feature, old_reader
    return db, calls


def _links_eq(a, b):
    return [(x.x_left, x.x_right, x.y) for x in a] == [(x.x_left, x.x_right, x.y) for x in b]


def test_soa_tracking_db_equals_the_reference_tracking_db(oracle, tmp_path):
    """slamfe.trackdb.SoATrackingDB (flat columns + track ids) built from run_host-format tables, against the
    TrackingDB the UNMODIFIED reference builds frame by frame (tracking_database.py:273-337): the read API
    answers the same, to_reference_db() rebuilds an equal dict-of-lists object (all seven maps), the .npz
    round trip is lossless, _check_consistency (tracking_database.py:442-471) passes on both forms."""
    ref = refshim.load()
    from slamfe import database as sdb, trackdb
    frames = _synthetic_frames(7)
    db_ref, calls = _reference_run(ref, frames)
    seq = sdb.pack_frames([(pl, pr, dl, dr) for pl, pr, dl, dr in frames], pin=False)
    flags = [None] + [np.asarray(c[3], dtype=bool) for c in calls[1:]]
    tables = _oracle_tables(seq, frames, oracle, flags)
    tables["n_good"] = np.array([max(4, int(f.sum())) for f in flags[1:]], np.int32)
    soa = trackdb.build(seq, tables, link_factory=ref.tracking_database.Link)
    assert soa.check_consistency()
    # --- read API
    assert soa.frame_num() == db_ref.frame_num() and soa.track_num() == db_ref.track_num() > 20
    assert soa.link_num() == db_ref.link_num() and sorted(soa.all_tracks()) == sorted(db_ref.all_tracks())
    for f in db_ref.all_frames():
        assert soa.tracks(f) == db_ref.tracks(f)
        assert np.array_equal(soa.features(f), db_ref.features(f))
        assert _links_eq(soa.all_frame_links(f), db_ref.all_frame_links(f))
        got, want = soa.links(f), db_ref.links(f)
        assert sorted(got) == sorted(want) and all(_links_eq([got[t]], [want[t]]) for t in want)
    for t in db_ref.all_tracks():
        assert soa.frames(t) == db_ref.frames(t) and soa.last_frame_of_track(t) == db_ref.last_frame_of_track(t)
        got, want = soa.track(t), db_ref.track(t)
        assert sorted(got) == sorted(want) and all(_links_eq([got[f]], [want[f]]) for f in want)
        f0 = want and sorted(want)[0]
        assert _links_eq([soa.link(f0, t)], [db_ref.link(f0, t)])
    assert soa.link(0, 10 ** 6) is None and soa.features(99) is None and soa.tracks(99) == []
    assert _links_eq(soa.all_last_frame_links(), db_ref.all_last_frame_links())
    # --- the reference's own object, rebuilt
    db2 = soa.to_reference_db(ref.tracking_database.TrackingDB, ref.tracking_database.Link)
    assert db2.last_frameId == db_ref.last_frameId and db2.last_trackId == db_ref.last_trackId
    assert db2.trackId_to_frames == db_ref.trackId_to_frames
    assert db2.frameId_to_trackIds_list == db_ref.frameId_to_trackIds_list
    assert db2.frameID_to_inliers_percent == db_ref.frameID_to_inliers_percent
    assert sorted(db2.linkId_to_link) == sorted(db_ref.linkId_to_link)
    assert all(_links_eq([db2.linkId_to_link[k]], [v]) for k, v in db_ref.linkId_to_link.items())
    assert sorted(db2.leftover_links) == sorted(db_ref.leftover_links)
    assert all(_links_eq(db2.leftover_links[f], v) for f, v in db_ref.leftover_links.items())
    assert _links_eq(db2.prev_frame_links, db_ref.prev_frame_links)
    assert all(np.array_equal(db2.frameId_to_lfeature[f], v) for f, v in db_ref.frameId_to_lfeature.items())
    db2._check_consistency()
    for f in db_ref.all_frames():
        assert _links_eq(db2.all_frame_links(f), db_ref.all_frame_links(f))
    # --- on-disk format
    path = tmp_path / "db.npz"
    soa.save(path)
    back = trackdb.SoATrackingDB.load(path, link_factory=ref.tracking_database.Link)
    assert back == soa and back.check_consistency() and back.tracks(3) == db_ref.tracks(3)
    import zipfile
    assert all(n.endswith(".npy") for n in zipfile.ZipFile(path).namelist())     # plain arrays, no pickle
    # --- the device table path: track ids supplied as a table (slamfe_track_ids' output format)
    ids, n_tracks = trackdb.track_ids_host(
        [tables["fwd_keys"][int(seq.l_off[f]):int(seq.l_off[f]) + int(tables["n_links"][f])].view(np.uint32) & 0x3FFFFF
         for f in range(seq.n_frames - 1)],
        [tables["inlier_fwd"][int(seq.l_off[f]):int(seq.l_off[f]) + int(tables["n_links"][f])] for f in range(seq.n_frames - 1)],
        tables["n_links"])
    table_ids = np.full(seq.desc_l.shape[0], -1, np.int32)
    for f, a in enumerate(ids):
        table_ids[int(seq.l_off[f]):int(seq.l_off[f]) + len(a)] = a
    tables2 = dict(tables, track_id=table_ids, n_tracks=np.array([n_tracks], np.int32))
    assert trackdb.build(seq, tables2, link_factory=ref.tracking_database.Link) == soa


def test_patch_batched_db_rebinds_create_db_on_the_real_reference(monkeypatch):
    """patch(batched_db=True) on the real reference module tree: `database.create_db` becomes the batched
    builder, which reads and describes the frames through the reference's own `Inputs.read_images` /
    `FEATURE.detectAndCompute` in order and hands them to slamfe.database.create_db (stubbed here: the
    CUDA pipeline cannot run in this container); a resumed build goes to the reference's loop; unpatch
    restores everything."""
    import sys
    import cv2
    ref = refshim.load()
    from slamfe import database as sdb, patch
    frames = _synthetic_frames(3)

    class Provider:
        def detectAndCompute(self, token, mask):
            side, f = token
            pts = frames[f][0 if side == "L" else 1]
            return tuple(cv2.KeyPoint(float(x), float(y), 1.0) for x, y in pts), frames[f][2 if side == "L" else 3]

    seen = {}

    def fake_create_db(frs, db, link_factory=None, **kw):
        seen["frames"], seen["db"], seen["link"], seen["kw"] = frs, db, link_factory, kw
        return db

    monkeypatch.setattr(sdb, "create_db", fake_create_db)
    monkeypatch.setattr(patch, "replacements", lambda: {})      # the GPU objects cannot be built here
    monkeypatch.setattr(ref.matching, "FEATURE", Provider())
    monkeypatch.setattr(ref.inputs, "read_images", lambda idx: (("L", idx), ("R", idx)))
    original = ref.database.create_db
    mods = {k: v for k, v in sys.modules.items() if k.startswith("final_project")}
    # only the batched_db part of patch() is exercised: the per-attribute rebinding needs the GPU objects
    monkeypatch.setattr(patch, "_REBINDS", {})
    tok = patch.patch(mods, batched_db=True, h_max=64)
    try:
        assert ref.database.create_db is not original and ref.database.create_db.__wrapped__ is original
        db = ref.database.create_db(start_frame=0, num_frames=3, db=None)
        assert isinstance(db, ref.tracking_database.TrackingDB) and seen["db"] is db
        assert seen["link"] is ref.tracking_database.Link and seen["kw"] == {"h_max": 64}
        assert len(seen["frames"]) == 3
        for f, (kl, kr, dl, dr) in enumerate(seen["frames"]):
            assert np.array_equal(dl, frames[f][2]) and np.array_equal(dr, frames[f][3])
            assert np.allclose([k.pt for k in kl], frames[f][0]) and np.allclose([k.pt for k in kr], frames[f][1])
        seq = sdb.pack_frames(seen["frames"], pin=False)     # what the batched builder packs for the GPU
        assert seq.n_frames == 3 and np.array_equal(seq.pts_l[:len(frames[0][0])], frames[0][0])
    finally:
        patch.unpatch(tok)
    assert ref.database.create_db is original


def test_patch_rebinds_a_non_binary_feature_and_pack_frames_validates(monkeypatch):
    """the checked-in reference selects SIFT (matching.py:72).  patch() rebinds matching.FEATURE to
    the AKAZE detector of get_akaze_matcher_lr_matcher() (matching.py:19-21) or, with rebind_feature=False,
    refuses; pack_frames rejects descriptors that are not (n, 61) uint8 instead of casting them."""
    import sys
    import cv2
    ref = refshim.load()
    from slamfe import database as sdb, patch
    monkeypatch.setattr(patch, "replacements", lambda: {})
    monkeypatch.setattr(patch, "_REBINDS", {})
    mods = {k: v for k, v in sys.modules.items() if k.startswith("final_project")}
    sift = cv2.SIFT_create()
    monkeypatch.setattr(ref.matching, "FEATURE", sift)
    token = patch.patch(mods)
    try:
        assert "FEATURE" in token["final_project.algorithms.matching"]
        f = ref.matching.FEATURE
        assert f is not sift and f.descriptorType() == 0 and f.descriptorSize() == 61
        assert abs(f.getThreshold() - 0.0008) < 1e-9 and f.getNOctaves() == 4 and f.getNOctaveLayers() == 4
    finally:
        patch.unpatch(token)
    assert ref.matching.FEATURE is sift
    with pytest.raises(TypeError, match="non-binary"):
        patch.patch(mods, rebind_feature=False)
    akaze = cv2.AKAZE_create()
    monkeypatch.setattr(ref.matching, "FEATURE", akaze)
    token = patch.patch(mods)
    assert ref.matching.FEATURE is akaze          # a binary detector is left alone
    patch.unpatch(token)
    pts = np.zeros((5, 2), np.float32)
    with pytest.raises(TypeError, match="uint8"):
        sdb.pack_frames([(pts, pts, np.zeros((5, 128), np.float32), np.zeros((5, 128), np.float32))], pin=False)
    with pytest.raises(TypeError, match="uint8"):
        sdb.pack_frames([(pts, pts, np.zeros((5, 32), np.uint8), np.zeros((5, 32), np.uint8))], pin=False)


def test_patch_batched_gating_reads_the_reference_module_globals(monkeypatch):
    """patch(batched_gating=True) on the real reference module tree: `loop_closure.get_good_candidates`
    (loop_closure.py:199-228) is rebound to slamfe.loop.get_good_candidates_typed, which reads the module's own
    `cov_dijkstra_graph` (the unmodified Graph class, graph.py) and `relative_covariance_dict` at call time.
    The device launch is replaced by a NumPy walk over the very arrays the kernel would get
    (CovarianceGraph.arrays(): directed rows), so this checks the translation of the reference's structures;
    expected values: the oracle's restatement of check_candidate over the reference's own adjacency dict."""
    import sys
    from oracle import ref_oracle as oracle
    from test_oracle import _random_pose_graph
    ref = refshim.load_loop_closure()
    from slamfe import loop as sloop, patch
    lc = ref.loop_closure
    rng = np.random.default_rng(11)
    K = 60
    poses, edges = _random_pose_graph(rng, K=K, n_loops=4)
    index_list = sorted(rng.choice(np.arange(4 * K), K, replace=False).tolist())

    def spd():
        a = rng.normal(0, 1, (6, 6))
        return (a @ a.T + 6 * np.eye(6)) * 1e-3

    graph, cov_dict = type(lc.cov_dijkstra_graph)(), {}
    for a, b, c in edges:
        graph.add_edge(index_list[a], index_list[b], np.array(c))
        if b == a + 1:
            cov_dict[str((index_list[a], index_list[b]))] = c
            cov_dict[str((index_list[b], index_list[a]))] = spd()
            cov_dict[index_list[b]] = c
    cov_dict[index_list[0]] = spd()
    monkeypatch.setattr(lc, "cov_dijkstra_graph", graph)
    monkeypatch.setattr(lc, "relative_covariance_dict", cov_dict)
    monkeypatch.setattr(sys.modules["gtsam"], "symbol", lambda ch, i: (ch, int(i)), raising=False)

    class Values:
        def atPose3(self, key):
            T = np.vstack([poses[index_list.index(key[1])], [0, 0, 0, 1.0]])
            return type("P", (), {"matrix": lambda self: T})()

    def gate_on_host(poses_, g, queries, gap=sloop.KEY_FRAME_GAP):
        off, node, edge, w, cov = g.arrays()
        out = np.full((len(queries), len(poses_)), np.inf)
        for qi, n in enumerate(queries):                     # Dijkstra from the query, as the kernel runs it
            import heapq
            dist, pred, pe = {n: 0.0}, {}, {}
            pq = [(0.0, n)]
            while pq:
                d, u = heapq.heappop(pq)
                if d > dist.get(u, np.inf):
                    continue
                for e in range(off[u], off[u + 1]):
                    v, nd = int(node[e]), d + w[edge[e]]
                    if nd < dist.get(v, np.inf):
                        dist[v], pred[v], pe[v] = nd, u, int(edge[e])
                        heapq.heappush(pq, (nd, v))
            Rn, tn = poses_[n][:3, :3], poses_[n][:3, 3]
            for i in range(0, n - gap):
                c, v = None, i
                while v != n:
                    c = cov[pe[v]].reshape(6, 6).copy() if c is None else c + cov[pe[v]].reshape(6, 6)
                    v = pred[v]
                xi = oracle.pose3_logmap(Rn.T @ poses_[i][:3, :3], Rn.T @ (poses_[i][:3, 3] - tn))
                out[qi, i] = np.sqrt(xi @ np.linalg.solve(c, xi))
        return out, None

    monkeypatch.setattr(sloop, "gate_distances", gate_on_host)
    monkeypatch.setattr(patch, "replacements", lambda: {})
    monkeypatch.setattr(patch, "_REBINDS", {})
    original = lc.get_good_candidates
    mods = {k: v for k, v in sys.modules.items() if k.startswith("final_project")}
    tok = patch.patch(mods, batched_gating=True)
    try:
        assert lc.get_good_candidates is not original
        some = 0
        for n in (K - 1, 40, 12):
            got = lc.get_good_candidates(n, None, Values(), index_list)
            want = oracle.gate_distances_typed(poses, graph.graph, cov_dict, index_list, n)
            ref_sel = sorted([(want[i], i) for i in np.nonzero(want < lc.MAHALANOBIS_THRESHOLD)[0]])[:lc.MAX_CANDIDATES]
            assert got == [index_list[i] for _, i in ref_sel]
            some += len(got)
        assert some > 0
    finally:
        patch.unpatch(tok)
    assert lc.get_good_candidates is original
