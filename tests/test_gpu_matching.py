"""Parity of the CUDA Hamming matcher against the golden vectors (cv2 4.13.0 / reference) and the
oracle: bit-exact indices, distances, ordering and tie-breaks."""
import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("matcher_kernel")]

CASES = ("akaze", "ragged", "orb", "ties", "one_train")


def _keys(ops, t):
    return ops.keys_to_numpy(t.cpu().numpy())


@pytest.mark.parametrize("name", CASES)
def test_match_golden(slamfe, golden, name):
    from slamfe import matching
    g = golden("matching")
    q, t = g[f"{name}_q"], g[f"{name}_t"]
    ms = matching.Matcher(crossCheck=False).match(q, t)
    assert isinstance(ms, tuple) and len(ms) == q.shape[0]
    assert [m.queryIdx for m in ms] == list(range(q.shape[0]))
    assert [m.trainIdx for m in ms] == g[f"{name}_match_t"].tolist()
    assert [m.distance for m in ms] == g[f"{name}_match_d"].tolist()
    assert all(m.imgIdx == 0 for m in ms) and type(ms[0]).__name__ == "DMatch"
    arr = np.array(ms)  # database.py:57-65 relies on object-array fancy indexing
    assert arr.dtype == object and arr[[0, len(ms) - 1]][1].queryIdx == len(ms) - 1


@pytest.mark.parametrize("name", CASES)
def test_crosscheck_golden(slamfe, golden, name):
    from slamfe import matching
    g = golden("matching")
    ms = matching.Matcher(crossCheck=True).match(g[f"{name}_q"], g[f"{name}_t"])
    assert [m.queryIdx for m in ms] == g[f"{name}_cc_q"].tolist()
    assert [m.trainIdx for m in ms] == g[f"{name}_cc_t"].tolist()
    assert [m.distance for m in ms] == g[f"{name}_cc_d"].tolist()


@pytest.mark.parametrize("name", CASES)
def test_knn_and_ratio_golden(slamfe, golden, name):
    import torch
    from slamfe import matching, ops
    g = golden("matching")
    q, t = g[f"{name}_q"], g[f"{name}_t"]
    knn = matching.Matcher().knnMatch(q, t, k=2)
    assert len(knn) == q.shape[0] and type(knn[0]) is tuple
    idx = np.full((len(knn), 2), -1, np.int32)
    dist = np.full((len(knn), 2), -1, np.int32)
    for i, pair in enumerate(knn):
        for c, m in enumerate(pair):
            assert m.queryIdx == i and m.imgIdx == 0
            idx[i, c], dist[i, c] = m.trainIdx, int(m.distance)
    assert np.array_equal(idx, g[f"{name}_knn_idx"]) and np.array_equal(dist, g[f"{name}_knn_dist"])
    row_keys, _ = ops.hamming_top2(torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda())
    mask = ops.ratio_test(row_keys).cpu().numpy().astype(bool)
    assert np.array_equal(mask, g[f"{name}_ratio"])
    assert np.array_equal(matching.ratio_test_mask(dist), g[f"{name}_ratio"])
    ui, ud = ops.unpack_keys(row_keys)
    assert np.array_equal(ui.cpu().numpy(), idx) and np.array_equal(ud.cpu().numpy(), dist)


def test_database_forward_backward_golden(slamfe, golden):
    """database.py:54-77 through the drop-in Matcher, plus the one-pass column minima."""
    import torch
    from slamfe import matching, ops
    g = golden("database")
    M = matching.Matcher()
    fwd = np.array(M.match(g["prev"], g["cur"]))
    bwd = np.array(M.match(g["cur"], g["prev"]))
    assert [m.trainIdx for m in fwd] == g["fwd_t"].tolist() and [m.trainIdx for m in bwd] == g["bwd_t"].tolist()
    good = [j for j, m in enumerate(fwd) if bwd[m.trainIdx].trainIdx == m.queryIdx]
    assert good == g["good_idx"].tolist()
    rk, ck = ops.hamming_top2(torch.from_numpy(g["prev"]).cuda(), torch.from_numpy(g["cur"]).cuda(), want_cols=True)
    ci, _ = _keys(ops, ck)
    assert np.array_equal(ci, g["bwd_t"])
    mt, md = ops.cross_check(rk, ck)
    assert np.array_equal(np.nonzero(mt.cpu().numpy() >= 0)[0], g["good_idx"])


@pytest.mark.parametrize("nq,nt", [(3000, 3000), (1, 5000), (4999, 1), (130, 129), (513, 2049)])
def test_random_vs_oracle(slamfe, oracle, nq, nt):
    import torch
    from slamfe import ops, synth
    rng = np.random.default_rng(nq * 7 + nt)
    q = synth.descriptors(rng, nq)
    t, _ = synth.paired_descriptors(rng, q, n_out=nt, dup_frac=0.05)
    rk, ck = ops.hamming_top2(torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda(), want_cols=True)
    idx, dist = _keys(ops, rk)
    oi, od = oracle.knn2(q, t)
    assert np.array_equal(idx, oi) and np.array_equal(dist, od)
    ci, cd = _keys(ops, ck)
    oci, ocd = oracle.colmin(q, t)
    assert np.array_equal(ci, oci) and np.array_equal(cd, ocd)


def test_layouts_padded_misaligned_and_short_descriptors(slamfe, oracle):
    """64-byte padded rows (garbage in the pad bytes must be ignored), a base pointer that is not
    16-byte aligned (TMA fallback path) and 32-byte descriptors."""
    import torch
    from slamfe import ops, synth
    rng = np.random.default_rng(3)
    q, t = synth.descriptors(rng, 700), synth.descriptors(rng, 900)
    oi, od = oracle.knn2(q, t)
    qp = rng.integers(0, 256, (700, 64), dtype=np.uint8); qp[:, :61] = q
    tp = rng.integers(0, 256, (900, 64), dtype=np.uint8); tp[:, :61] = t
    rk, _ = ops.hamming_top2(torch.from_numpy(qp).cuda(), torch.from_numpy(tp).cuda(), desc_bytes=61)
    i, d = _keys(ops, rk)
    assert np.array_equal(i, oi) and np.array_equal(d, od)
    big_q = torch.from_numpy(np.concatenate([np.zeros((1, 61), np.uint8), q])).cuda()
    big_t = torch.from_numpy(np.concatenate([np.zeros((3, 61), np.uint8), t])).cuda()
    rk, _ = ops.hamming_top2(big_q[1:], big_t[3:])
    i, d = _keys(ops, rk)
    assert np.array_equal(i, oi) and np.array_equal(d, od)
    q32, t32 = np.ascontiguousarray(q[:, :32]), np.ascontiguousarray(t[:, :32])
    rk, _ = ops.hamming_top2(torch.from_numpy(q32).cuda(), torch.from_numpy(t32).cuda())
    i, d = _keys(ops, rk)
    o32i, o32d = oracle.knn2(q32, t32)
    assert np.array_equal(i, o32i) and np.array_equal(d, o32d)


def test_batched_ragged_vs_oracle(slamfe, oracle):
    """Ragged batch with device-side counts, empty problems, both CTA shapes."""
    import torch
    from slamfe import frontend, ops, synth
    rng = np.random.default_rng(11)
    for sizes in ([(300, 280), (0, 50), (17, 0), (1, 1), (515, 700)], [(1100, 900)] * 70 + [(5, 2000)]):
        qs = [synth.descriptors(rng, a) for a, _ in sizes]
        ts = [synth.paired_descriptors(rng, synth.descriptors(rng, max(b, 1)), n_out=b, dup_frac=0.05)[0] for _, b in sizes]
        q_off = frontend.plan_offsets([a for a, _ in sizes]); t_off = frontend.plan_offsets([b for _, b in sizes])
        Q = np.zeros((q_off[-1], 61), np.uint8); T = np.zeros((t_off[-1], 61), np.uint8)
        for p, (a, b) in enumerate(sizes):
            Q[q_off[p]:q_off[p] + a] = qs[p]; T[t_off[p]:t_off[p] + b] = ts[p]
        dev = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
        rk, ck = ops.hamming_top2_batched(
            dev(Q), dev(q_off), dev(T), dev(t_off), len(sizes), max(a for a, _ in sizes), max(b for _, b in sizes), 61,
            q_cnt=dev(np.array([a for a, _ in sizes], np.int32)), t_cnt=dev(np.array([b for _, b in sizes], np.int32)),
            want_cols=True)
        ri, rd = _keys(ops, rk); ci, cd = _keys(ops, ck)
        for p, (a, b) in enumerate(sizes):
            if a == 0:
                continue
            if b == 0:
                assert np.all(ri[q_off[p]:q_off[p] + a] == -1)
                continue
            oi, od = oracle.knn2(qs[p], ts[p])
            assert np.array_equal(ri[q_off[p]:q_off[p] + a], oi), p
            assert np.array_equal(rd[q_off[p]:q_off[p] + a], od), p
            oci, ocd = oracle.colmin(qs[p], ts[p])
            assert np.array_equal(ci[t_off[p]:t_off[p] + b], oci) and np.array_equal(cd[t_off[p]:t_off[p] + b], ocd)
        # rows in the alignment padding stay NONE
        pad = np.ones(q_off[-1], bool)
        for p, (a, _) in enumerate(sizes):
            pad[q_off[p]:q_off[p] + a] = False
        assert np.all(ri[pad] == -1)


def test_train_shards_merge_exactly(slamfe, oracle):
    """Config 5 on one GPU: train slices with global index bases + slamfe_merge_top2 == full sweep."""
    import torch
    from slamfe import dist, ops, synth
    rng = np.random.default_rng(21)
    q = synth.descriptors(rng, 1500)
    t, _ = synth.paired_descriptors(rng, q, n_out=4100, dup_frac=0.05)
    qd, td = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
    b = dist.train_slices(t.shape[0], 4)
    shards = torch.stack([ops.hamming_top2(qd, td[b[r]:b[r + 1]], t_index_base=int(b[r]))[0] for r in range(4)])
    merged = ops.merge_top2(shards)
    i, d = _keys(ops, merged)
    oi, od = oracle.knn2(q, t)
    assert np.array_equal(i, oi) and np.array_equal(d, od)
    assert np.array_equal(dist.merge_top2_host(shards.cpu().numpy()).view(np.int32), merged.cpu().numpy())


def test_dense_20k_sweep_vs_oracle(slamfe, oracle):
    """BASELINE config 5 at full size (20k x 20k): bit-exact against the C oracle, plus the
    size-independent properties (self-match, row/column duality)."""
    import torch
    from slamfe import ops, synth
    rng = np.random.default_rng(4)
    q = synth.descriptors(rng, 20000)
    t, _ = synth.paired_descriptors(rng, q, n_out=20000)
    qd, td = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
    rk, ck = ops.hamming_top2(qd, td, want_cols=True)
    i, d = _keys(ops, rk)
    oi, od = oracle.knn2(q, t)
    assert np.array_equal(i, oi) and np.array_equal(d, od)
    rk2, ck2 = ops.hamming_top2(td, qd, want_cols=True)  # duality: columns of (q,t) == rows of (t,q)
    assert np.array_equal(ck.cpu().numpy(), rk2[:, 0].cpu().numpy())
    assert np.array_equal(ck2.cpu().numpy(), rk[:, 0].cpu().numpy())
    rs, _ = ops.hamming_top2(qd, qd)  # self sweep: best distance is 0 at an index <= i (duplicates)
    si, sd = _keys(ops, rs)
    assert np.all(sd[:, 0] == 0) and np.all(si[:, 0] <= np.arange(20000))


@pytest.mark.parametrize("best_only", [False, True])
def test_candidate_pairs_vs_oracle(slamfe, oracle, best_only):
    """Loop-closure form (loop_closure.py:422): a pool of keyframes, every keyframe against all
    earlier ones, results indexed by candidate pair; the same keyframe is query in many pairs."""
    import torch
    from slamfe import dist, ops, synth
    rng = np.random.default_rng(41)
    sizes = [260, 40, 513, 129, 300, 1]
    pool = [synth.descriptors(rng, n) for n in sizes]
    pool[3][:60] = pool[0][:60]            # revisit: exact duplicates force index tie-breaks
    pool[4][10:200] = synth.flip_bits(rng, pool[2][10:200], 0.05)
    off = np.zeros(len(sizes) + 1, np.int64)
    np.cumsum([-(-n // 16) * 16 for n in sizes], out=off[1:])
    flat = np.zeros((off[-1], 61), np.uint8)
    for k, d in enumerate(pool):
        flat[off[k]:off[k] + sizes[k]] = d
    pairs = dist.candidate_pairs(len(sizes))
    nq = np.array([sizes[i] for i, _ in pairs], np.int32)
    out_off = np.zeros(len(pairs) + 1, np.int64)
    np.cumsum(nq, out=out_off[1:])
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).cuda()
    fd = torch.from_numpy(flat).cuda()
    keys = ops.hamming_pairs(fd, dev(off[pairs[:, 0]]), dev(nq), fd, dev(off[pairs[:, 1]]),
                             dev([sizes[j] for _, j in pairs]), dev(out_off[:-1]), len(pairs), max(sizes), max(sizes),
                             61, out_rows_total=int(out_off[-1]), best_only=best_only)
    idx, dst = ops.keys_to_numpy(keys.cpu().numpy())
    for p, (i, j) in enumerate(pairs):
        oi, od = oracle.knn2(pool[i], pool[j])
        got_i, got_d = idx[out_off[p]:out_off[p + 1]], dst[out_off[p]:out_off[p + 1]]
        assert np.array_equal(got_i[:, 0], oi[:, 0]) and np.array_equal(got_d[:, 0], od[:, 0]), (i, j)
        if best_only:
            assert (got_i[:, 1] == -1).all()
        else:
            assert np.array_equal(got_i[:, 1], oi[:, 1]) and np.array_equal(got_d[:, 1], od[:, 1]), (i, j)


def test_compact_keys_and_batched_index_base(slamfe, oracle):
    """SLAMFE_MATCH_COMPACT_KEYS (one key per row) and the batched t_index_base (train shards)."""
    import torch
    from slamfe import ops, synth
    rng = np.random.default_rng(43)
    F, nq, nt = 3, 700, 1040
    q = np.concatenate([synth.descriptors(rng, nq) for _ in range(F)])
    t = np.concatenate([synth.paired_descriptors(rng, q[f * nq:(f + 1) * nq], n_out=nt, dup_frac=0.05)[0]
                        for f in range(F)])
    qd, td = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
    i32 = lambda a: torch.tensor(a, dtype=torch.int32, device="cuda")
    base, cnt = 512, nt - 512            # second train shard of every frame
    keys, _ = ops.hamming_top2_batched(qd, i32([f * nq for f in range(F)]), td, i32([f * nt + base for f in range(F)]),
                                       F, nq, cnt, 61, q_cnt=i32([nq] * F), t_cnt=i32([cnt] * F), t_index_base=base)
    comp, _ = ops.hamming_top2_batched(qd, i32([f * nq for f in range(F)]), td, i32([f * nt + base for f in range(F)]),
                                       F, nq, cnt, 61, q_cnt=i32([nq] * F), t_cnt=i32([cnt] * F), t_index_base=base,
                                       best_only=True, compact=True)
    assert comp.shape == (F * nq,)
    ki, kd = ops.keys_to_numpy(keys.cpu().numpy())
    ci, cd = ops.keys_to_numpy(comp.cpu().numpy())
    for f in range(F):
        oi, od = oracle.knn2(q[f * nq:(f + 1) * nq], t[f * nt + base:(f + 1) * nt])
        sl = slice(f * nq, (f + 1) * nq)
        assert np.array_equal(ki[sl], oi + base) and np.array_equal(kd[sl], od)
        assert np.array_equal(ci[sl], oi[:, 0] + base) and np.array_equal(cd[sl], od[:, 0])
    with pytest.raises(ValueError):
        ops.hamming_top2_batched(qd, i32([0]), td, i32([0]), 1, nq, nt, 61, q_cnt=i32([nq]), t_cnt=i32([nt]),
                                 compact=True)


def test_knn_inner_tuple_lengths(slamfe):
    """cv2: knnMatch(k=2) against a single train row returns length-1 inner tuples; k=1 likewise."""
    from slamfe import matching, synth
    rng = np.random.default_rng(44)
    q, t = synth.descriptors(rng, 5), synth.descriptors(rng, 1)
    m = matching.Matcher()
    knn = m.knnMatch(q, t, k=2)
    assert len(knn) == 5 and all(len(p) == 1 and p[0].trainIdx == 0 and p[0].queryIdx == i for i, p in enumerate(knn))
    t3 = synth.descriptors(rng, 3)
    k1 = m.knnMatch(q, t3, k=1)
    k2 = m.knnMatch(q, t3, k=2)
    assert all(len(p) == 1 for p in k1) and all(len(p) == 2 for p in k2)
    assert [p[0].trainIdx for p in k1] == [p[0].trainIdx for p in k2] == [x.trainIdx for x in m.match(q, t3)]


def test_shared_matcher_from_several_threads(slamfe, oracle):
    """The module-global MATCHER objects of the reference are shared; here each Python thread gets its
    own pinned staging buffers and runs on its own CUDA stream (the C-ABI keeps no global state)."""
    import threading
    import torch
    from slamfe import matching, synth
    rng = np.random.default_rng(45)
    m = matching.Matcher(crossCheck=True)
    jobs = []
    for _ in range(6):
        dl, dr, _, _ = synth.stereo_frame(rng, int(rng.integers(300, 900)))
        jobs.append((dl, dr, oracle.match_crosscheck(dl, dr)))
    errors = []

    def work(k):
        try:
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream):
                for _ in range(5):
                    dl, dr, (cq, ct, cd) = jobs[k]
                    qi, ti, d = m.match_arrays(dl, dr)
                    assert np.array_equal(qi, cq) and np.array_equal(ti, ct) and np.array_equal(d, cd)
        except Exception as exc:  # pragma: no cover - reported below
            errors.append((k, repr(exc)))

    threads = [threading.Thread(target=work, args=(k,)) for k in range(len(jobs))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_real_akaze_descriptors_through_the_dropin_matchers(slamfe, golden):
    """tests/golden/akaze_real.npz: real cv2.AKAZE descriptors (correlated bits, > 3k per image) through the
    drop-in Matcher objects, against what the reference's extract_kps_descs_matches / MATCHER.match /
    knnMatch returned (matching.py:38-45, database.py:54-55).  The arrays are also passed the way cv2 hands
    them around: as non-contiguous views of a wider buffer."""
    from slamfe import matching
    g = golden("akaze_real")
    d0, d1, d2 = g["desc_l"], g["desc_r"], g["desc_l2"]
    wide = np.zeros((len(d0), 80), np.uint8)
    wide[:, 7:68] = d0
    view0 = wide[:, 7:68]
    assert not view0.flags["C_CONTIGUOUS"]
    lr, mm = matching.Matcher(crossCheck=True), matching.Matcher(crossCheck=False)
    for q in (d0, view0):
        ms = lr.match(q, d1)
        assert [m.queryIdx for m in ms] == g["cross_q"].tolist() and [m.trainIdx for m in ms] == g["cross_t"].tolist()
        assert [m.distance for m in ms] == g["cross_d"].tolist()
        fw = mm.match(q, d2)
        assert [m.trainIdx for m in fw] == g["match_t"].tolist() and [m.distance for m in fw] == g["match_d"].tolist()
    inl, outl = matching.extract_inliers_outliers(g["pts_l"], g["pts_r"], ms)
    assert np.array_equal(inl, g["inliers"]) and np.array_equal(outl, g["outliers"])
    bw = mm.match(d2, d0)
    assert [m.trainIdx for m in bw] == g["back_t"].tolist() and [m.distance for m in bw] == g["back_d"].tolist()
    knn = mm.knnMatch(d0, d2, k=2)
    assert [[m.trainIdx for m in p] for p in knn] == g["knn_idx"].tolist()
    assert [[int(m.distance) for m in p] for p in knn] == g["knn_dist"].tolist()
    assert int(g["n_exact_ties"]) >= 0


def test_tensor_core_and_int_kernels_agree_on_random_shapes(slamfe, oracle):
    """Fuzz: the tcgen05 kernel (int8 contraction, packed 16-bit keys, spare-K popc / invalid-row markers) and
    the INT kernel (XOR / carry-save POPC) must produce IDENTICAL key tables — rows, second neighbours and
    column minima — on random sizes around every tile boundary (128 / 256 query rows, 64 / 128 train rows),
    every descriptor width the tensor kernel takes (1..63 bytes), padded strides, misaligned bases and heavy
    duplication; a sample is also checked against the C oracle."""
    import torch
    from slamfe import ops
    rng = np.random.default_rng(97)
    edges = [1, 2, 31, 63, 64, 65, 127, 128, 129, 191, 192, 193, 255, 256, 257, 383, 384, 385, 511, 513]
    for case in range(70):
        nq = int(rng.choice(edges)) if rng.random() < 0.6 else int(rng.integers(1, 900))
        nt = int(rng.choice(edges)) if rng.random() < 0.6 else int(rng.integers(1, 900))
        db = int(rng.integers(1, 64)) if case % 3 else 61
        stride = db if rng.random() < 0.5 else int(rng.integers(db, 65))
        lead_q, lead_t = int(rng.integers(0, 3)), int(rng.integers(0, 3))
        Q = rng.integers(0, 256, (nq + lead_q, stride), dtype=np.uint8)
        T = rng.integers(0, 256, (nt + lead_t, stride), dtype=np.uint8)
        if rng.random() < 0.5 and nt > 1:   # duplicates: ties on distance, decided by index
            T[rng.integers(lead_t, lead_t + nt, nt // 2)] = T[rng.integers(lead_t, lead_t + nt, nt // 2)]
        if rng.random() < 0.3:              # near-identical rows: small distances incl. d = 0
            T[lead_t:lead_t + min(nq, nt), :db] = Q[lead_q:lead_q + min(nq, nt), :db]
        qd, td = torch.from_numpy(Q).cuda()[lead_q:], torch.from_numpy(T).cuda()[lead_t:]
        cols, best_only = bool(rng.random() < 0.6), bool(rng.random() < 0.5)
        base = int(rng.integers(0, 1000)) if rng.random() < 0.3 else 0
        res = {}
        for kind in ("mma", "int"):
            old = ops.set_matcher_kernel(kind)
            try:
                res[kind] = ops.hamming_top2(qd, td, desc_bytes=db, want_cols=cols, t_index_base=base, best_only=best_only)
            finally:
                ops.set_matcher_kernel(old)
        tag = (case, nq, nt, db, stride, cols, best_only, base)
        assert torch.equal(res["mma"][0], res["int"][0]), tag
        if cols:
            assert torch.equal(res["mma"][1], res["int"][1]), tag
        if case % 7 == 0 and not best_only:
            oi, od = oracle.knn2(np.ascontiguousarray(Q[lead_q:, :db]), np.ascontiguousarray(T[lead_t:, :db]))
            ki, kd = ops.keys_to_numpy(res["mma"][0].cpu().numpy())
            assert np.array_equal(ki, np.where(oi >= 0, oi + base, -1)) and np.array_equal(kd, od), tag
