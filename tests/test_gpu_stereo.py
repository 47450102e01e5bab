"""Row filter / links kernels against the golden vectors and the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_extract_inliers_outliers_golden(slamfe, golden):
    import cv2
    from slamfe import matching
    g = golden("stereo")
    kpl = tuple(cv2.KeyPoint(float(x), float(y), 1.0) for x, y in g["pts_l"])
    kpr = tuple(cv2.KeyPoint(float(x), float(y), 1.0) for x, y in g["pts_r"])
    ms = tuple(cv2.DMatch(int(a), int(b), 0, float(c)) for a, b, c in zip(g["match_q"], g["match_t"], g["match_d"]))
    inl, outl = matching.extract_inliers_outliers(kpl, kpr, ms)
    assert np.array_equal(inl, g["inliers"]) and np.array_equal(outl, g["outliers"])
    assert inl.dtype == g["inliers"].dtype
    ident = tuple(cv2.DMatch(i, i, 0, 0.0) for i in range(6))  # exact-threshold rows
    i6, o6 = matching.extract_inliers_outliers(kpl, kpr, ident)
    assert np.array_equal(i6, g["ident_inliers"]) and np.array_equal(o6, g["ident_outliers"])
    e_in, e_out = matching.extract_inliers_outliers(kpl, kpr, ())
    assert e_in.size == 0 and e_out.size == 0


def test_stereo_links_batched_vs_oracle(slamfe, oracle, golden):
    import torch
    from slamfe import frontend, ops, synth
    rng = np.random.default_rng(31)
    g = golden("stereo")
    frames = [(g["desc_l"], g["desc_r"], g["pts_l"], g["pts_r"])]
    for n in (700, 1, 333, 2049):
        frames.append(synth.stereo_frame(rng, n))
    dl0, dr0, pl0, pr0 = synth.stereo_frame(rng, 64)
    frames.append((dl0, dr0[:40], pl0, pr0[:40]))  # Nl != Nr
    seq = frontend.pack_sequence(frames, pin=False)
    ds = frontend.to_device(seq)
    F = seq.n_frames
    rk, ck = ops.hamming_top2_batched(ds.desc_l, ds.l_off, ds.desc_r, ds.r_off, F, ds.max_nl, ds.max_nr, 61,
                                      q_cnt=ds.n_l, t_cnt=ds.n_r, want_cols=True)
    out = ops.stereo_links_batched(rk, ck, ds.l_off, ds.r_off, F, ds.pts_l, ds.pts_r, desc_left=ds.desc_l,
                                   desc_bytes=61, n_l=ds.n_l, n_r=ds.n_r)
    host = {k: v.cpu().numpy() for k, v in out.items()}
    for f, (dl, dr, pl, pr) in enumerate(frames):
        cq, ct, cd = oracle.match_crosscheck(dl, dr)
        lo = seq.l_off[f]
        mt = host["match_t"][lo:lo + len(dl)]
        assert np.array_equal(np.nonzero(mt >= 0)[0], cq) and np.array_equal(mt[mt >= 0], ct)
        assert host["n_matches"][f] == len(cq)
        inl, _ = oracle.extract_inliers_outliers(pl, pr, cq, ct)
        valid, links = oracle.create_links(pl, pr, cq[inl], ct[inl])
        k = host["n_links"][f]
        assert k == len(inl)
        assert np.array_equal(host["link_src"][lo:lo + k], cq[inl])
        assert np.array_equal(host["links"][lo:lo + k], links.astype(np.float32))
        assert np.all(host["link_src"][lo + k:lo + len(dl)] == -1)
        feat = host["feat"][lo:lo + k]
        assert np.array_equal(feat[:, :61], dl[valid]) and np.all(feat[:, 61:] == 0)
    if g is not None:  # frame 0 is the golden frame: links must equal the reference's create_links
        k = host["n_links"][0]
        assert np.array_equal(host["links"][:k], g["links"].astype(np.float32))
        assert np.array_equal(host["feat"][:k, :61], g["features"])
