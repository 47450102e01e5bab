"""Device-level operators: torch tensors in, torch tensors out, libslamfe kernels underneath.

PyTorch is used for memory ownership and streams only; every computation is a hand-written
sm_100a kernel reached through the C-ABI (include/slamfe.h).  All functions are asynchronous on
the current CUDA stream and raise if the library or the GPU is missing (no CPU fallback).
"""
from __future__ import annotations

import numpy as np

from . import _cabi
from ._cabi import KEY_IDX_BITS, KEY_IDX_MASK, KEY_NONE, check, load_library, ptr, stream_handle


def _torch():
    return _cabi.require_cuda()


_MATCHER_KERNEL = "mma"


def set_matcher_kernel(kind):
    """Which kernel the matcher entry points run: "int" = the INT-pipe carry-save kernel
    (csrc/hamming.cu), "mma" = the tcgen05 tensor-core kernel (csrc/hamming_mma.cu, SLAMFE_MATCH_MMA).
    Both produce identical keys.  Returns the previous setting.  (The environment variable
    SLAMFE_MATCH_MMA=0/1, read by the library, overrides this per process.)"""
    global _MATCHER_KERNEL
    if kind not in ("int", "mma"):
        raise ValueError("matcher kernel must be 'int' or 'mma'")
    old, _MATCHER_KERNEL = _MATCHER_KERNEL, kind
    return old


def matcher_kernel():
    return _MATCHER_KERNEL


def _flags(best_only, compact=False):
    return ((_cabi.MATCH_BEST_ONLY if best_only else 0) | (_cabi.MATCH_COMPACT_KEYS if compact else 0)
            | (_cabi.MATCH_MMA if _MATCHER_KERNEL == "mma" else 0))


def _desc_tensor(d):
    torch = _torch()
    if d.dtype != torch.uint8 or d.dim() != 2 or not d.is_cuda:
        raise ValueError("descriptors must be a 2-D uint8 CUDA tensor")
    if d.stride(1) != 1:
        d = d.contiguous()
    return d


def hamming_top2(q, t, desc_bytes=None, want_cols=False, t_index_base=0, best_only=False):
    """Top-2 Hamming neighbours of every row of q among the rows of t.

    q, t: (N, stride) uint8 CUDA tensors (stride 61 = cv2 layout, or 64-byte padded rows);
    desc_bytes defaults to the row width.  Returns (row_keys (Nq, 2) int32-viewed-uint32,
    col_keys (Nt,) or None).  Key = distance << 22 | index; see include/slamfe.h.
    best_only=True skips the second neighbour (row_keys[:, 1] = KEY_NONE): enough for .match()
    and crossCheck, and cheaper.
    """
    torch = _torch()
    q, t = _desc_tensor(q), _desc_tensor(t)
    nq, nt = q.shape[0], t.shape[0]
    desc_bytes = q.shape[1] if desc_bytes is None else desc_bytes
    row_keys = torch.empty((nq, 2), dtype=torch.int32, device=q.device)
    col_keys = torch.empty((nt,), dtype=torch.int32, device=q.device) if want_cols else None
    with torch.cuda.device(q.device):
        check(load_library().slamfe_hamming_top2(
            ptr(q), nq, q.stride(0), ptr(t), nt, t.stride(0) if nt else max(desc_bytes, 1), desc_bytes, t_index_base,
            ptr(row_keys), ptr(col_keys), _flags(best_only), stream_handle()),
            "slamfe_hamming_top2")
    return row_keys, col_keys


def hamming_top2_batched(q, q_off, t, t_off, n_problems, max_nq, max_nt, desc_bytes,
                         q_cnt=None, t_cnt=None, want_cols=False, row_keys=None, col_keys=None, best_only=False,
                         t_index_base=0, compact=False):
    """Ragged batch of matching problems in one launch (slamfe_hamming_top2_batched).

    q_off / t_off / q_cnt / t_cnt are int32 CUDA tensors (see include/slamfe.h); keys are indexed
    by global row and hold problem-local indices (+ t_index_base).  compact=True (with best_only)
    returns one uint32 key per row instead of (best, second).
    """
    torch = _torch()
    q, t = _desc_tensor(q), _desc_tensor(t)
    if compact and not best_only:
        raise ValueError("compact keys need best_only=True")
    if row_keys is None:
        row_keys = torch.empty((q.shape[0],) if compact else (q.shape[0], 2), dtype=torch.int32, device=q.device)
    if want_cols and col_keys is None:
        col_keys = torch.empty((t.shape[0],), dtype=torch.int32, device=q.device)
    with torch.cuda.device(q.device):
        check(load_library().slamfe_hamming_top2_batched(
            ptr(q), q.stride(0), ptr(q_off), ptr(q_cnt), ptr(t), t.stride(0), ptr(t_off), ptr(t_cnt),
            n_problems, max_nq, max_nt, desc_bytes, int(t_index_base),
            ptr(row_keys), q.shape[0], ptr(col_keys) if want_cols else 0, t.shape[0],
            _flags(best_only, compact), stream_handle()),
            "slamfe_hamming_top2_batched")
    return row_keys, (col_keys if want_cols else None)


def hamming_pairs(q, q_off, q_cnt, t, t_off, t_cnt, out_off, n_problems, max_nq, max_nt, desc_bytes=None,
                  row_keys=None, out_rows_total=None, best_only=False, compact=False):
    """Candidate-pair matching (slamfe_hamming_top2_pairs): problem p = rows q_off[p].. of q against
    rows t_off[p].. of t, results at row_keys[out_off[p] + i].  All index tensors int32 CUDA.
    compact=True (with best_only): row_keys is (rows,) — one key per query row."""
    torch = _torch()
    q, t = _desc_tensor(q), _desc_tensor(t)
    desc_bytes = q.shape[1] if desc_bytes is None else desc_bytes
    if row_keys is None:
        if out_rows_total is None:
            raise ValueError("give row_keys or out_rows_total")
        row_keys = torch.empty((out_rows_total,) if compact else (out_rows_total, 2), dtype=torch.int32,
                               device=q.device)
    if out_rows_total is None:
        out_rows_total = row_keys.shape[0]
    with torch.cuda.device(q.device):
        check(load_library().slamfe_hamming_top2_pairs(
            ptr(q), q.stride(0), ptr(q_off), ptr(q_cnt), ptr(t), t.stride(0), ptr(t_off), ptr(t_cnt), ptr(out_off),
            n_problems, max_nq, max_nt, desc_bytes, ptr(row_keys), out_rows_total,
            _flags(best_only, compact), stream_handle()), "slamfe_hamming_top2_pairs")
    return row_keys


def unpack_keys(keys):
    """keys (...,) -> (idx, dist) int32 tensors of the same shape, -1 where no neighbour."""
    torch = _torch()
    keys = keys.contiguous()
    idx = torch.empty_like(keys)
    dist = torch.empty_like(keys)
    with torch.cuda.device(keys.device):
        check(load_library().slamfe_unpack_keys(ptr(keys), keys.numel(), ptr(idx), ptr(dist), stream_handle()),
              "slamfe_unpack_keys")
    return idx, dist


def merge_top2(shard_keys):
    """(n_shards, nq, 2) keys -> (nq, 2) global top-2 (exact: keys carry global indices)."""
    torch = _torch()
    shard_keys = shard_keys.contiguous()
    n_shards, nq = shard_keys.shape[0], shard_keys.shape[1]
    out = torch.empty((nq, 2), dtype=torch.int32, device=shard_keys.device)
    with torch.cuda.device(out.device):
        check(load_library().slamfe_merge_top2(ptr(shard_keys), n_shards, nq, ptr(out), stream_handle()),
              "slamfe_merge_top2")
    return out


def cross_check(row_keys, col_keys):
    """crossCheck epilogue: (match_t, match_dist) int32 (nq,), -1 where the pair is not mutual.  The two
    are rows of ONE (2, nq) buffer (`match_t._base`), so a caller can bring both to the host in one copy."""
    torch = _torch()
    nq, nt = row_keys.shape[0], col_keys.shape[0]
    both = torch.empty((2, nq), dtype=torch.int32, device=row_keys.device)
    mt, md = both[0], both[1]
    with torch.cuda.device(mt.device):
        check(load_library().slamfe_cross_check(ptr(row_keys), ptr(col_keys), nq, nt, ptr(mt), ptr(md),
                                                stream_handle()), "slamfe_cross_check")
    return mt, md


def ratio_test(row_keys, num=5, den=3):
    """mask[i] = num * d1 < den * d2  (5*d1 < 3*d2 == d1 < 0.6*d2, VAN_ex/code/ex1.py:118-122)."""
    torch = _torch()
    nq = row_keys.shape[0]
    mask = torch.empty((nq,), dtype=torch.uint8, device=row_keys.device)
    with torch.cuda.device(mask.device):
        check(load_library().slamfe_ratio_test(ptr(row_keys), nq, num, den, ptr(mask), stream_handle()),
              "slamfe_ratio_test")
    return mask


def stereo_filter(pts_left, pts_right, match_q, match_t):
    """matching.py:48-69 mask over matches; pts (N,2) float32, match_* int32 CUDA tensors."""
    torch = _torch()
    n = match_q.shape[0]
    mask = torch.empty((n,), dtype=torch.uint8, device=match_q.device)
    with torch.cuda.device(mask.device):
        check(load_library().slamfe_stereo_filter(ptr(pts_left), ptr(pts_right), ptr(match_q), ptr(match_t), n,
                                                  ptr(mask), stream_handle()), "slamfe_stereo_filter")
    return mask


def stereo_links_batched(row_keys, col_keys, l_off, r_off, n_frames, pts_left, pts_right,
                         desc_left=None, desc_bytes=61, out=None, n_l=None, n_r=None):
    """Fused crossCheck + row filter + create_links + feature compaction over all frames.

    Returns a dict of CUDA tensors: match_t, n_matches, n_links, link_src, links, feat.
    """
    torch = _torch()
    dev = row_keys.device
    l_rows = row_keys.shape[0]
    if out is None:
        out = {
            "match_t": torch.empty((l_rows,), dtype=torch.int32, device=dev),
            "n_matches": torch.empty((n_frames,), dtype=torch.int32, device=dev),
            "n_links": torch.empty((n_frames,), dtype=torch.int32, device=dev),
            "link_src": torch.empty((l_rows,), dtype=torch.int32, device=dev),
            "links": torch.empty((l_rows, 3), dtype=torch.float32, device=dev),
            "feat": torch.empty((l_rows, 64), dtype=torch.uint8, device=dev) if desc_left is not None else None,
        }
    with torch.cuda.device(dev):
        check(load_library().slamfe_stereo_links_batched(
            ptr(row_keys), ptr(col_keys), ptr(l_off), ptr(n_l), ptr(r_off), ptr(n_r), n_frames,
            ptr(pts_left), ptr(pts_right),
            ptr(desc_left), desc_left.stride(0) if desc_left is not None else 0, desc_bytes,
            ptr(out["match_t"]), ptr(out["n_matches"]), ptr(out["n_links"]), ptr(out["link_src"]),
            ptr(out["links"]), ptr(out["feat"]), stream_handle()), "slamfe_stereo_links_batched")
    return out


def triangulate_links(links, P, Q, out=None):
    """links (n, 3) [x_left, x_right, y] float32 or float64 CUDA tensor -> xyz (n, 3), same dtype."""
    torch = _torch()
    links = links.contiguous()
    n = links.shape[0]
    if out is None:
        out = torch.empty((n, 3), dtype=links.dtype, device=links.device)
    fn = {torch.float64: "slamfe_triangulate_links_f64", torch.float32: "slamfe_triangulate_links_f32"}[links.dtype]
    Pb, Qb = _cabi.host_doubles(P, 12), _cabi.host_doubles(Q, 12)
    with torch.cuda.device(links.device):
        check(getattr(load_library(), fn)(ptr(links), n, Pb, Qb, ptr(out), stream_handle()), fn)
    return out


def triangulate_dlt(pxy, qxy, P, Q):
    """General 4x4 DLT (distinct y, arbitrary P/Q): pxy, qxy (n, 2) float64 -> xyz (n, 3) float64."""
    torch = _torch()
    pxy, qxy = pxy.contiguous(), qxy.contiguous()
    n = pxy.shape[0]
    out = torch.empty((n, 3), dtype=torch.float64, device=pxy.device)
    Pb, Qb = _cabi.host_doubles(P, 12), _cabi.host_doubles(Q, 12)
    with torch.cuda.device(pxy.device):
        check(load_library().slamfe_triangulate_dlt_f64(ptr(pxy), ptr(qxy), n, Pb, Qb, ptr(out), stream_handle()),
              "slamfe_triangulate_dlt_f64")
    return out


def same_rows_stereo(P, Q) -> bool:
    """True when P[1:] == Q[1:] bitwise (rectified pair): the links kernel applies."""
    P, Q = np.asarray(P, dtype=np.float64), np.asarray(Q, dtype=np.float64)
    return bool(np.array_equal(P[1:], Q[1:]))


def ransac_score(T, pts, l_pix, r_pix, K, M1, M2, hyp_valid=None, pt_off=None, n_frames=1, max_points=None,
                 pt_cnt=None, out=None):
    """All hypotheses x all correspondences (x all frames) in one launch.

    T (n_frames*H, 3, 4) or (n_frames, H, 3, 4) float64; pts (Ntot, 3), l_pix / r_pix (Ntot, 2)
    float64; pt_off (n_frames+1,) int32 CUDA tensor for n_frames > 1.
    Returns (counts (n_frames, H) int32, best (n_frames, 2) int32 [index, count], best_mask (Ntot,) uint8).
    """
    torch = _torch()
    dev = pts.device
    T = T.contiguous()
    H = T.numel() // (12 * n_frames) if n_frames else 0
    pts, l_pix, r_pix = pts.contiguous(), l_pix.contiguous(), r_pix.contiguous()
    n_tot = pts.shape[0]
    if max_points is None:
        max_points = n_tot
    if out is None:
        out = {}
    counts = out.get("counts")
    if counts is None or counts.shape != (n_frames, H):
        counts = torch.empty((n_frames, H), dtype=torch.int32, device=dev)
    best = out.get("best", torch.empty((n_frames, 2), dtype=torch.int32, device=dev))
    mask = out.get("best_mask", torch.empty((n_tot,), dtype=torch.uint8, device=dev))
    work = out.get("work", torch.empty((n_frames,), dtype=torch.int32, device=dev))
    Kb, M1b, M2b = _cabi.host_doubles(K, 9), _cabi.host_doubles(M1, 12), _cabi.host_doubles(M2, 12)
    with torch.cuda.device(dev):
        check(load_library().slamfe_ransac_score(
            ptr(T), ptr(hyp_valid), H, ptr(pts), ptr(l_pix), ptr(r_pix), ptr(pt_off), ptr(pt_cnt), n_tot, n_frames,
            max_points,
            Kb, M1b, M2b, ptr(counts), ptr(best), ptr(mask), ptr(work), stream_handle()), "slamfe_ransac_score")
    return counts, best, mask


def track_ids(fwd_keys, inlier_fwd, l_off, n_links, n_frames, out=None):
    """Track id of every link row of a sequence (slamfe_track_ids; TrackingDB.add_frame's bookkeeping,
    tracking_database.py:273-337).  fwd_keys (L,) compact best keys / inlier_fwd (L,) as FrontEnd.track leaves them.
    Returns (track_id (L,) int32 with -1 = NO_ID, n_tracks (1,) int32, head_base (n_frames + 1,) int32)."""
    torch = _torch()
    dev = fwd_keys.device
    L = fwd_keys.shape[0]
    out = out if out is not None else {}
    i32 = dict(dtype=torch.int32, device=dev)

    def buf(name, shape):
        t = out.get(name)
        if t is None or tuple(t.shape) != tuple(shape):
            t = out[name] = torch.empty(shape, **i32)
        return t

    pred, rank, track_id = buf("trk_pred", (L,)), buf("trk_rank", (L,)), buf("track_id", (L,))
    head_cnt, head_base, n_tracks = buf("trk_head_cnt", (max(n_frames, 1),)), buf("head_base", (n_frames + 1,)), \
        buf("n_tracks", (1,))
    with torch.cuda.device(dev):
        check(load_library().slamfe_track_ids(ptr(fwd_keys), ptr(inlier_fwd), ptr(l_off), ptr(n_links), n_frames, L,
                                              ptr(pred), ptr(rank), ptr(head_cnt), ptr(head_base), ptr(track_id),
                                              ptr(n_tracks), stream_handle()), "slamfe_track_ids")
    return track_id, n_tracks, head_base


def pack_db(o, l_off, r_off, pts_l, pts_r, n_frames, desc_bytes, track_id=None, out=None):
    """Dense tracking-database columns on the device (slamfe_pack_db) from the pipeline tables `o`
    (n_links, link_src, match_t, feat).  Returns a dict: link_off (F + 1,) int32, x_left / x_right (L,)
    float32, y (L,) float64, feat (L, desc_bytes) uint8, track_id (L,) int32 — the first link_off[-1] rows
    of each are the store."""
    torch = _torch()
    dev = o["feat"].device
    L = o["feat"].shape[0]
    out = out if out is not None else {}

    def buf(name, shape, dtype):
        t = out.get(name)
        if t is None or tuple(t.shape) != tuple(shape):
            t = out[name] = torch.empty(shape, dtype=dtype, device=dev)
        return t

    res = {"link_off": buf("db_link_off", (n_frames + 1,), torch.int32), "x_left": buf("db_x_left", (L,), torch.float32),
           "x_right": buf("db_x_right", (L,), torch.float32), "y": buf("db_y", (L,), torch.float64),
           "feat": buf("db_feat", (L, desc_bytes), torch.uint8), "track_id": buf("db_track_id", (L,), torch.int32)}
    with torch.cuda.device(dev):
        check(load_library().slamfe_pack_db(
            ptr(l_off), ptr(r_off), ptr(o["n_links"]), ptr(o["link_src"]), ptr(o["match_t"]), ptr(pts_l), ptr(pts_r),
            ptr(o["feat"]), int(desc_bytes), ptr(track_id), n_frames, ptr(res["link_off"]), ptr(res["x_left"]),
            ptr(res["x_right"]), ptr(res["y"]), ptr(res["feat"]), ptr(res["track_id"]), stream_handle()),
            "slamfe_pack_db")
    return res


def gate_candidates(poses, adj_off, adj_node, adj_edge, edge_w, edge_cov, queries, gap):
    """Mahalanobis gating distances (slamfe_gate_candidates; loop_closure.py:164-228) for the query
    keyframes `queries` (int32 CUDA) against all earlier keyframes.  Returns (dist (Q, K) float64 with +inf
    where i is no candidate, hops (Q, K) int32)."""
    torch = _torch()
    dev = poses.device
    K, Q = poses.shape[0], queries.shape[0]
    dist = torch.empty((Q, K), dtype=torch.float64, device=dev)
    hops = torch.empty((Q, K), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(load_library().slamfe_gate_candidates(
            ptr(poses.contiguous()), K, ptr(adj_off), ptr(adj_node), ptr(adj_edge), ptr(edge_w), ptr(edge_cov.contiguous()),
            ptr(queries), Q, int(gap), ptr(dist), ptr(hops), stream_handle()), "slamfe_gate_candidates")
    return dist, hops


def pnp_refit(T, best, pts, l_pix, mask, K, pt_off=None, pt_cnt=None, n_frames=1, max_iter=20, tol=1e-12, out=None):
    """Refit of the pose on the consensus set (ransac.py:185-193) for n_frames problems in one launch
    (slamfe_pnp_refit): T / best / mask as returned by ransac_hypotheses / ransac_score.
    Returns (T_refit (n_frames, 3, 4) float64, status (n_frames,) int32, rms (n_frames,) float64)."""
    torch = _torch()
    dev = pts.device
    T = T.contiguous()
    H = T.numel() // (12 * n_frames) if n_frames else 0
    out = out or {}
    T_out = out.get("T_refit")
    if T_out is None:
        T_out = torch.empty((n_frames, 3, 4), dtype=torch.float64, device=dev)
    status = out.get("refit_status")
    if status is None:
        status = torch.empty((n_frames,), dtype=torch.int32, device=dev)
    rms = out.get("refit_rms")
    if rms is None:
        rms = torch.empty((n_frames,), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        check(load_library().slamfe_pnp_refit(
            ptr(T), H, ptr(best), ptr(pts.contiguous()), ptr(l_pix.contiguous()), ptr(mask), ptr(pt_off), ptr(pt_cnt),
            pts.shape[0], n_frames, _cabi.host_doubles(K, 9), int(max_iter), float(tol), ptr(T_out), ptr(status),
            ptr(rms), stream_handle()), "slamfe_pnp_refit")
    return T_out, status, rms


def ransac_hypotheses(pts, l_pix, K, H, seed=0, pt_off=None, pt_cnt=None, n_frames=1, n_hyp=None, sample_idx=None,
                      out=None, frame_index_base=0):
    """Pose hypotheses on the GPU (slamfe_ransac_hypotheses): P3P + 4th-point disambiguation on 4
    sampled correspondences per hypothesis.  Returns (T (n_frames*H, 3, 4) float64, valid (n_frames*H,)
    uint8) — directly usable as the T / hyp_valid arguments of ransac_score."""
    torch = _torch()
    dev = pts.device
    pts, l_pix = pts.contiguous(), l_pix.contiguous()
    if out is not None:
        T, valid = out
    else:
        T = torch.empty((n_frames * H, 3, 4), dtype=torch.float64, device=dev)
        valid = torch.empty((n_frames * H,), dtype=torch.uint8, device=dev)
    Kb = _cabi.host_doubles(K, 9)
    with torch.cuda.device(dev):
        check(load_library().slamfe_ransac_hypotheses(
            ptr(pts), ptr(l_pix), ptr(pt_off), ptr(pt_cnt), pts.shape[0], n_frames, H, ptr(n_hyp), ptr(sample_idx),
            int(seed) & 0xFFFFFFFFFFFFFFFF, int(frame_index_base), Kb, ptr(T), ptr(valid), stream_handle()),
            "slamfe_ransac_hypotheses")
    return T, valid


def track_gather(o, ds_l_off, ds_r_off, pts_left, pts_right, n_pairs, P, Q, h_max, out):
    """slamfe_track_gather on the pipeline tables `o` (FrontEnd buffers); fills out[good_j, good_t,
    n_good, n_hyp, pts, lpix, rpix]."""
    torch = _torch()
    Pb, Qb = _cabi.host_doubles(P, 12), _cabi.host_doubles(Q, 12)
    with torch.cuda.device(pts_left.device):
        check(load_library().slamfe_track_gather(
            ptr(o["fwd_keys"]), ptr(o["bwd_keys"]), ptr(ds_l_off), ptr(ds_r_off), ptr(o["n_links"]),
            ptr(o["n_matches"]), ptr(pts_left), ptr(pts_right), ptr(o["link_src"]), ptr(o["match_t"]), n_pairs,
            Pb, Qb, int(h_max), ptr(out["good_j"]), ptr(out["good_t"]), ptr(out["n_good"]), ptr(out["n_hyp"]),
            ptr(out.get("n_hyp_full")), ptr(out["pts"]), ptr(out["lpix"]), ptr(out["rpix"]), stream_handle()),
            "slamfe_track_gather")
    return out


def pairs_gather(keys, q_off, q_cnt, t_off, out_off, n_problems, max_nq, kf_pts, kf_links, pts, lpix, rpix):
    """slamfe_pairs_gather: RANSAC inputs of loop-closure candidates from the compact match keys."""
    torch = _torch()
    with torch.cuda.device(keys.device):
        check(load_library().slamfe_pairs_gather(ptr(keys), ptr(q_off), ptr(q_cnt), ptr(t_off), ptr(out_off),
                                                 n_problems, max_nq, ptr(kf_pts), ptr(kf_links), ptr(pts), ptr(lpix),
                                                 ptr(rpix), stream_handle()), "slamfe_pairs_gather")


def scatter_inliers(best_mask, good_j, l_off, n_good, best, n_pairs, inlier_fwd):
    torch = _torch()
    with torch.cuda.device(inlier_fwd.device):
        check(load_library().slamfe_scatter_inliers(ptr(best_mask), ptr(good_j), ptr(l_off), ptr(n_good), ptr(best),
                                                    n_pairs, ptr(inlier_fwd), inlier_fwd.numel(), stream_handle()),
              "slamfe_scatter_inliers")
    return inlier_fwd


def measure_peak(mode: int, iters: int = 4096, ctas_per_sm: int = 8, block: int = 256):
    """Run a pipe-peak micro-benchmark; returns ops/s (popc/s for modes 0-1, fp64 FMA/s for mode 2)."""
    import ctypes
    torch = _torch()
    dev = torch.cuda.current_device()
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    grid = sms * ctas_per_sm
    sink = torch.empty((grid * block,), dtype=torch.int32, device=torch.device("cuda", dev))
    ops = ctypes.c_int(0)
    lib = load_library()
    best = 0.0
    for _ in range(3):
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        check(lib.slamfe_peak_kernel(mode, iters, grid, block, ptr(sink), ctypes.byref(ops), stream_handle()),
              "slamfe_peak_kernel")
        stop.record()
        stop.synchronize()
        ms = start.elapsed_time(stop)
        best = max(best, grid * block * float(iters) * ops.value / (ms * 1e-3))
    return best


def keys_to_numpy(keys_host: np.ndarray):
    """Host-side decode of packed keys: returns (idx int32, dist int32), -1 for KEY_NONE."""
    k = keys_host.view(np.uint32)
    none = k == KEY_NONE
    idx = (k & KEY_IDX_MASK).astype(np.int32)
    dist = (k >> KEY_IDX_BITS).astype(np.int32)
    idx[none] = -1
    dist[none] = -1
    return idx, dist
