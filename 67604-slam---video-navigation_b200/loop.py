"""Batched loop-closure candidate verification — the hot half of backend/loop/loop_closure.py.

The reference verifies one candidate at a time (`check_candidate_match`, loop_closure.py:405-436):
`MATCHER.match(kf1_features, kf2_features)` (:422, crossCheck=False) followed by
`ransac_pnp(matches, kf1_links, kf2_links, inliers_percent=40)` (:425 -> ransac.py:116-204: 888
iterations of sample / EPnP / transformation_agreement), and accepts the first candidate with more
than INLIERS_THRESHOLD = 120 inliers (`consensus_matches`, :572-599).  Here a whole block of
(keyframe, candidate) pairs goes through each stage in one launch, device resident:

    slamfe_hamming_top2_pairs -> slamfe_pairs_gather -> slamfe_ransac_hypotheses -> slamfe_ransac_score
    -> slamfe_pnp_refit (the final solve on the consensus set, ransac.py:185-193)

Candidate gating (Mahalanobis distance on the pose graph) and the GTSAM bundle stay with the
reference (out of scope).  Hypotheses come from the GPU generator
(DESIGN.md section 2.6), so inlier counts are statistically — not bit-wise — equal to a run of the
reference, which is itself unseeded.
"""
from __future__ import annotations

import numpy as np

from . import _cabi, ops
from .ransac import calc_ransac_iteration

INLIERS_THRESHOLD = 120   # loop_closure.py:17
LOOP_INLIERS_PERCENT = 40  # loop_closure.py:425
MAHALANOBIS_THRESHOLD = 220   # loop_closure.py:15
MAX_CANDIDATES = 15           # loop_closure.py:18
KEY_FRAME_GAP = 10            # loop_closure.py:20


class CandidateVerifier:
    """Owns the device buffers of one block of candidates and reuses them across calls."""

    def __init__(self, K=None, M1=None, M2=None, block_pairs=8192):
        from . import utils
        self.K = utils.K if K is None else np.asarray(K, float)
        self.M1 = utils.M1 if M1 is None else np.asarray(M1, float)
        self.M2 = utils.M2 if M2 is None else np.asarray(M2, float)
        self.P, self.Q = self.K @ self.M1, self.K @ self.M2
        self.block_pairs = int(block_pairs)
        self._buf = None
        self._key = None
        self.last_launches = 0

    def _buffers(self, rows, n_pairs, H, dev):
        torch = _cabi.require_cuda()
        key = (rows, n_pairs, H, dev)
        if self._key != key:
            f64 = dict(dtype=torch.float64, device=dev)
            i32 = dict(dtype=torch.int32, device=dev)
            self._buf = {
                "keys": torch.empty((rows,), **i32), "pts": torch.empty((rows, 3), **f64),
                "lpix": torch.empty((rows, 2), **f64), "rpix": torch.empty((rows, 2), **f64),
                "T": torch.empty((n_pairs * H, 3, 4), **f64),
                "hyp_valid": torch.empty((n_pairs * H,), dtype=torch.uint8, device=dev),
                "counts": torch.empty((n_pairs, H), **i32), "best": torch.empty((n_pairs, 2), **i32),
                "best_mask": torch.empty((rows,), dtype=torch.uint8, device=dev), "work": torch.empty((n_pairs,), **i32),
                "refit_status": torch.empty((n_pairs,), **i32), "refit_rms": torch.empty((n_pairs,), **f64),
            }
            self._key = key
        return self._buf

    def verify(self, pool_desc, pool_links, kf_off, kf_cnt, pairs, inliers_percent=LOOP_INLIERS_PERCENT, n_iter=None,
               seed=1, want_masks=False, pair_base=0, key_table=None, sync=True):
        """pool_desc (R, 61|64) uint8 CUDA: filtered features of all keyframes; pool_links (R, 3) float64
        CUDA [x_left, x_right, y] of the same rows; kf_off / kf_cnt: row offset / link count per keyframe
        (numpy int); pairs (P, 2) numpy: (reference keyframe, candidate keyframe).
        Returns a dict of numpy arrays per pair: n_matches, inliers (best count), best_hyp, percentage
        (= inliers / n_matches, loop_closure.py:429), accepted (inliers > 120); with want_masks also
        `keys` (best match key per query row) and `mask` (inlier flag per match) as ragged lists.
        key_table: optional (sum of n_matches,) int32 CUDA tensor that receives the best-match keys of
        all pairs (pair-major), e.g. for the multi-GPU all-gather of match tables.  sync=False leaves
        the per-pair [best_hyp, inliers] table on the device (res["best_dev"], (P, 2) int32) and skips
        the host copies.  res["pose"] (P, 3, 4) / res["pose_status"] (P,): the world-to-camera [R|t] refit
        on each candidate's consensus set (slamfe_pnp_refit; status > 0 = converged, 0 = fewer than 4
        inliers) — `pose_dev` / `pose_status_dev` with sync=False."""
        torch = _cabi.require_cuda()
        dev = pool_desc.device
        pairs = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
        kf_off = np.asarray(kf_off, dtype=np.int64)
        kf_cnt = np.asarray(kf_cnt, dtype=np.int64)
        H = int(n_iter) if n_iter is not None else calc_ransac_iteration(inliers_percent)
        kf_pts = ops.triangulate_links(pool_links, self.P, self.Q)       # once per keyframe pool
        res = {"n_matches": kf_cnt[pairs[:, 0]].copy(), "inliers": np.zeros(len(pairs), np.int64),
               "best_hyp": np.full(len(pairs), -1, np.int64)}
        best_dev = torch.empty((len(pairs), 2), dtype=torch.int32, device=dev)
        pose_dev = torch.empty((len(pairs), 3, 4), dtype=torch.float64, device=dev)
        pose_ok_dev = torch.empty((len(pairs),), dtype=torch.int32, device=dev)
        row0 = 0
        keys_out, mask_out = [], []
        self.last_launches = 1
        max_n = int(kf_cnt.max()) if len(kf_cnt) else 0
        blk = max(1, min(self.block_pairs, 65535))
        rows_cap = blk * max_n
        for b0 in range(0, len(pairs), blk):
            pb = pairs[b0:b0 + blk]
            n = len(pb)
            q_cnt_h = kf_cnt[pb[:, 0]]
            out_off_h = np.zeros(n + 1, np.int64)
            np.cumsum(q_cnt_h, out=out_off_h[1:])
            rows = int(out_off_h[-1])
            buf = self._buffers(rows_cap, blk, H, dev)
            small = np.concatenate([kf_off[pb[:, 0]], q_cnt_h, kf_off[pb[:, 1]], kf_cnt[pb[:, 1]], out_off_h[:-1]])
            sd = torch.from_numpy(small.astype(np.int32)).to(dev, non_blocking=True)
            q_off, q_cnt, t_off, t_cnt, out_off = (sd[i * n:(i + 1) * n] for i in range(5))
            v = {k: buf[k][:rows] for k in ("keys", "pts", "lpix", "rpix", "best_mask")}
            if key_table is not None:
                v["keys"] = key_table[row0:row0 + rows]
            row0 += rows
            ops.hamming_pairs(pool_desc, q_off, q_cnt, pool_desc, t_off, t_cnt, out_off, n, max_n, max_n,
                              row_keys=v["keys"], out_rows_total=rows, best_only=True, compact=True)
            ops.pairs_gather(v["keys"], q_off, q_cnt, t_off, out_off, n, max_n, kf_pts, pool_links, v["pts"],
                             v["lpix"], v["rpix"])
            T, valid = buf["T"][:n * H], buf["hyp_valid"][:n * H]
            ops.ransac_hypotheses(v["pts"], v["lpix"], self.K, H, seed=seed, pt_off=out_off, pt_cnt=q_cnt, n_frames=n,
                                  out=(T, valid), frame_index_base=pair_base + b0)
            out = {"counts": buf["counts"][:n], "best": best_dev[b0:b0 + n], "best_mask": v["best_mask"],
                   "work": buf["work"][:n]}
            ops.ransac_score(T, v["pts"], v["lpix"], v["rpix"], self.K, self.M1, self.M2, hyp_valid=valid,
                             pt_off=out_off, pt_cnt=q_cnt, n_frames=n, max_points=max_n, out=out)
            ops.pnp_refit(T, out["best"], v["pts"], v["lpix"], v["best_mask"], self.K, pt_off=out_off, pt_cnt=q_cnt,
                          n_frames=n, out={"T_refit": pose_dev[b0:b0 + n], "refit_status": pose_ok_dev[b0:b0 + n],
                                           "refit_rms": buf["refit_rms"][:n]})
            self.last_launches += 5
            if want_masks:
                kh = v["keys"].cpu().numpy().view(np.uint32)
                mh = v["best_mask"].cpu().numpy().astype(bool)
                for p in range(n):
                    keys_out.append(kh[out_off_h[p]:out_off_h[p + 1]].copy())
                    mask_out.append(mh[out_off_h[p]:out_off_h[p + 1]].copy())
        res["best_dev"], res["pose_dev"], res["pose_status_dev"] = best_dev, pose_dev, pose_ok_dev
        if not sync:
            return res
        best = best_dev.cpu().numpy()
        res["pose"], res["pose_status"] = pose_dev.cpu().numpy(), pose_ok_dev.cpu().numpy()
        res["best_hyp"][:] = best[:, 0]
        res["inliers"][:] = best[:, 1]
        with np.errstate(divide="ignore", invalid="ignore"):
            res["percentage"] = np.where(res["n_matches"] > 0, res["inliers"] / np.maximum(res["n_matches"], 1), 0.0)
        res["accepted"] = res["inliers"] > INLIERS_THRESHOLD
        if want_masks:
            res["keys"], res["mask"] = keys_out, mask_out
        return res


# ------------------------------------------------------------------------------------------------
# Drop-ins for the two callers in backend/loop/loop_closure.py (same names, signatures, returns)
# ------------------------------------------------------------------------------------------------
_default_verifier = None


def _verifier():
    """The verifier behind the drop-in entry points; rebuilt when slamfe.ransac.set_cameras (patch())
    has changed the cameras since it was made."""
    global _default_verifier
    from . import ransac
    v = _default_verifier
    if v is None or not (np.array_equal(v.K, ransac.K) and np.array_equal(v.M1, ransac.M1)
                         and np.array_equal(v.M2, ransac.M2)):
        v = _default_verifier = CandidateVerifier(K=ransac.K, M1=ransac.M1, M2=ransac.M2, block_pairs=64)
    return v


def _keyframe_pool(db, frames):
    """Features and links of the given keyframes of a TrackingDB as one padded pool (host arrays)."""
    from .triangulation import links_to_array
    feats = [np.ascontiguousarray(db.features(f), dtype=np.uint8) for f in frames]
    links = [links_to_array(db.all_frame_links(f)) for f in frames]
    cnt = np.array([len(x) for x in feats], dtype=np.int64)
    off = np.zeros(len(frames) + 1, dtype=np.int64)
    np.cumsum((cnt + 15) // 16 * 16, out=off[1:])
    width = feats[0].shape[1] if len(feats) and feats[0].ndim == 2 else 61
    pool_d = np.zeros((max(int(off[-1]), 1), width), dtype=np.uint8)
    pool_l = np.tile(np.array([[30.0, 10.0, 50.0]]), (max(int(off[-1]), 1), 1))
    for k in range(len(frames)):
        pool_d[off[k]:off[k] + cnt[k]] = feats[k]
        pool_l[off[k]:off[k] + cnt[k]] = links[k]
    return pool_d, pool_l, off[:-1], cnt, links


def _verify_from_db(reference_key_frame, candidates, db, seed=None):
    torch = _cabi.require_cuda()
    frames = [reference_key_frame] + list(candidates)
    pool_d, pool_l, off, cnt, links = _keyframe_pool(db, frames)
    pairs = np.array([(0, k + 1) for k in range(len(candidates))], dtype=np.int64)
    ver = _verifier()
    if seed is None:
        seed = int(np.random.randint(0, 2 ** 31 - 1))   # follows np.random.seed like the reference's sampling
    res = ver.verify(torch.from_numpy(pool_d).cuda(), torch.from_numpy(pool_l).cuda(), off, cnt, pairs,
                     inliers_percent=LOOP_INLIERS_PERCENT, seed=seed, want_masks=True)
    return res, links


REFIT = "gpu"   # "gpu": slamfe_pnp_refit inside the batch (default); "cv2": host cv2.solvePnP(EPNP) per candidate


def _candidate_result(res, k, links_ref, ver):
    """(inlier DMatch list, percentage, camera_to_world pose or None) of candidate k, as
    check_candidate_match returns them (loop_closure.py:425-436 + ransac.py:185-204)."""
    import cv2
    from . import ransac
    keys, mask = res["keys"][k], res["mask"][k]
    idx = np.nonzero(mask)[0]
    if len(idx) < 4:  # ransac.py:187-188: (None, [], []) -> percentage_inliers = 0 via the except branch
        return [], 0, None
    ti = (keys & _cabi.KEY_IDX_MASK).astype(np.int64)
    td = (keys >> _cabi.KEY_IDX_BITS).astype(np.float64)
    if REFIT == "gpu" and res["pose_status"][k] != 0:
        T = res["pose"][k]
    else:
        from .triangulation import triangulate_link_array
        pts = triangulate_link_array(links_ref[idx], ver.P, ver.Q)
        cur = res["_links"][k + 1][ti[idx]]
        ok, rvec, tvec = cv2.solvePnP(pts, np.ascontiguousarray(cur[:, [0, 2]]), ver.K, distCoeffs=np.zeros((5, 1)),
                                      flags=cv2.SOLVEPNP_EPNP)
        if not ok:  # ransac.py:204 -> (None, None, None): the reference then fails on `for i in None`
            raise TypeError("'NoneType' object is not iterable")
        T = ransac.rodriguez_to_mat(rvec, tvec)
    pose = ransac._pose3(T).inverse()
    matches = list(map(cv2.DMatch, idx.tolist(), ti[idx].tolist(), [0] * len(idx), td[idx].tolist()))
    return matches, len(idx) / len(keys), pose


def check_candidate_match(reference_key_frame, candiate_keyframe, db):
    """loop_closure.py:405-436 -> ([inlier DMatch], percentage_inliers, camera_to_world Pose3 or None)."""
    res, links = _verify_from_db(reference_key_frame, [candiate_keyframe], db)
    res["_links"] = links
    return _candidate_result(res, 0, links[0], _verifier())


def consensus_matches(reference_key_frame, candidates_index_lst, data_base):
    """loop_closure.py:572-599: the first candidate (in list order) with more than INLIERS_THRESHOLD inlier
    matches wins.  All candidates are verified in ONE batch (match + 888-hypothesis RANSAC each); the
    return value — (best_candidate or None, best_matches, rel_T of the winner, else of the last candidate
    evaluated) — is what the reference's sequential loop produces."""
    candidates = list(candidates_index_lst)
    if not candidates:
        return None, [], None
    res, links = _verify_from_db(reference_key_frame, candidates, data_base)
    res["_links"] = links
    ver = _verifier()
    rel_T = None
    for k, cand in enumerate(candidates):
        if res["inliers"][k] > INLIERS_THRESHOLD:
            matches, _, rel_T = _candidate_result(res, k, links[0], ver)
            return cand, matches, rel_T
    _, _, rel_T = _candidate_result(res, len(candidates) - 1, links[0], ver)
    return None, [], rel_T


# ------------------------------------------------------------------------------------------------
# Candidate gating (loop_closure.py:164-228) on arrays: the GTSAM objects stay with the caller
# ------------------------------------------------------------------------------------------------
class CovarianceGraph:
    """The covariance graph of backend/loop/graph.py (undirected, edge weight det(cov)) as arrays, plus the
    keyframe poses: what get_good_candidates needs from the pose graph.  Nodes are keyframe INDICES
    0..K-1 (positions in the reference's index_list); `add_edge` mirrors Graph.add_edge (:15-25; adding an
    existing edge replaces it, as update_edge does)."""

    def __init__(self, n_nodes):
        self.n_nodes = int(n_nodes)
        self._edges = {}          # (min, max) -> (cov for min->max, cov for max->min, weight)

    def add_edge(self, v1, v2, cov, cov_back=None, weight=None):
        """cov is added to a path that traverses the edge v1 -> v2, cov_back (default: cov) to one that
        traverses it v2 -> v1 — the reference keeps direction-dependent covariances for consecutive keyframes
        (loop_closure.py:264-279: cov_dict[str((i, i+1))] and cov_dict[str((i+1, i))]); weight defaults to
        det(cov) (graph.py:11-13)."""
        a, b = int(v1), int(v2)
        c = np.array(cov, dtype=np.float64).reshape(6, 6)
        cb = c if cov_back is None else np.array(cov_back, dtype=np.float64).reshape(6, 6)
        w = float(np.linalg.det(c)) if weight is None else float(weight)
        self._edges[(min(a, b), max(a, b))] = (c, cb, w) if a <= b else (cb, c, w)

    def remove_edge(self, v1, v2):
        return self._edges.pop((min(int(v1), int(v2)), max(int(v1), int(v2))), None) is not None

    def arrays(self):
        """CSR adjacency + per-DIRECTED-edge weight / covariance (rows 2e: min -> max, 2e + 1: max -> min).
        Neighbours are listed in insertion order of the edges, as the reference's dict-of-dicts iterates them.
        The device Dijkstra runs from the query keyframe c_n outwards, the reference sums covariances along
        the path c_i -> c_n: the adjacency entry "u lists v" therefore carries the row of the traversal v -> u."""
        E = len(self._edges)
        ends = np.array(list(self._edges), dtype=np.int32).reshape(E, 2)
        cov = np.zeros((2 * E, 36), dtype=np.float64)
        w = np.zeros(2 * E, dtype=np.float64)
        for e, (c_ab, c_ba, wt) in enumerate(self._edges.values()):
            cov[2 * e], cov[2 * e + 1] = c_ab.reshape(36), c_ba.reshape(36)
            w[2 * e] = w[2 * e + 1] = wt
        deg = np.zeros(self.n_nodes + 1, dtype=np.int64)
        for a, b in ends:
            deg[a + 1] += 1
            deg[b + 1] += 1
        off = np.cumsum(deg).astype(np.int32)
        fill = off[:-1].astype(np.int64).copy()
        node = np.zeros(2 * E, dtype=np.int32)
        edge = np.zeros(2 * E, dtype=np.int32)
        for e, (a, b) in enumerate(ends):
            node[fill[a]], edge[fill[a]] = b, 2 * e + 1      # hop a -> b from the query = traversal b -> a
            fill[a] += 1
            node[fill[b]], edge[fill[b]] = a, 2 * e          # hop b -> a from the query = traversal a -> b
            fill[b] += 1
        return off, node, edge, w, cov


def gate_distances(poses, graph: CovarianceGraph, queries, gap=KEY_FRAME_GAP):
    """Mahalanobis distance of every earlier keyframe to each query keyframe (check_candidate,
    loop_closure.py:164-196), all queries in one launch.  poses: (K, 3, 4) or (K, 4, 4) camera-to-world.
    Returns (dist (Q, K) float64 numpy: +inf where not a candidate, hops (Q, K))."""
    torch = _cabi.require_cuda()
    P = np.ascontiguousarray(np.asarray(poses, dtype=np.float64)[:, :3, :4]).reshape(-1, 12)
    off, node, edge, w, cov = graph.arrays()
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    q = np.ascontiguousarray(np.asarray(queries, dtype=np.int32).reshape(-1))
    dist, hops = ops.gate_candidates(dev(P), dev(off), dev(node), dev(edge), dev(w), dev(cov), dev(q), gap)
    return dist.cpu().numpy(), hops.cpu().numpy()


def select_candidates(dist_row, index_list=None, threshold=MAHALANOBIS_THRESHOLD, max_candidates=MAX_CANDIDATES):
    """get_good_candidates' selection (loop_closure.py:214-228): distances below the threshold, ascending
    (stable: ties keep index order, as list.sort), the best `max_candidates`; mapped through index_list."""
    d = np.asarray(dist_row, dtype=np.float64)
    cand = np.nonzero(d < threshold)[0]
    cand = cand[np.argsort(d[cand], kind="stable")][:max_candidates]
    return [int(index_list[c]) if index_list is not None else int(c) for c in cand]


def get_good_candidates(c_n_index, poses, graph, index_list=None, gap=KEY_FRAME_GAP):
    """loop_closure.py:199-228 on arrays: the candidate keyframes of keyframe c_n_index."""
    dist, _ = gate_distances(poses, graph, [c_n_index], gap=gap)
    return select_candidates(dist[0], index_list)


# -- the same behind the reference's own signature (GTSAM-typed arguments, duck-typed here) --------
def traversal_covariance(a, b, cov_dict, marginals=None, symbol=None):
    """The covariance get_relative_covariance_along_path (loop_closure.py:110-135) adds for the hop a -> b
    (frame numbers): the cached str((a, b)) entry (consecutive keyframes, :264-279); otherwise what
    get_relative_consecutive_covariance (:84-107) returns — the cached per-frame entry cov_dict[b] if there is
    one (there is for every keyframe after init, :282,:286 — so a loop-closure edge contributes the covariance of
    b given its PREDECESSOR keyframe: a quirk of the reference, kept), else the inverse of b's block of the
    joint marginal information of (a, b)."""
    k = str((a, b))
    if k in cov_dict:
        return np.asarray(cov_dict[k], dtype=np.float64)
    if b in cov_dict:
        return np.asarray(cov_dict[b], dtype=np.float64)
    c1, c2 = symbol("c", a), symbol("c", b)
    keys = [c1, c2]
    try:  # gtsam.KeyVector when the real module is around (loop_closure.py:99-101)
        import gtsam
        kv = gtsam.KeyVector()
        kv.append(c1)
        kv.append(c2)
        keys = kv
    except Exception:
        pass
    return np.linalg.inv(np.asarray(marginals.jointMarginalInformation(keys).at(c2, c2), dtype=np.float64))


def covariance_graph_from_reference(ref_graph, index_list, cov_dict, marginals=None, symbol=None):
    """CovarianceGraph (nodes = positions in index_list) from the reference's `cov_dijkstra_graph`
    (backend/loop/graph.py: adjacency dict in insertion order, weights as stored) and its
    `relative_covariance_dict`, with the direction-dependent covariances of traversal_covariance."""
    pos = {int(f): i for i, f in enumerate(index_list)}
    g = CovarianceGraph(len(index_list))
    for a, nbrs in ref_graph.graph.items():
        for b, w in nbrs.items():
            if a in pos and b in pos and (min(pos[a], pos[b]), max(pos[a], pos[b])) not in g._edges:
                g.add_edge(pos[a], pos[b], traversal_covariance(a, b, cov_dict, marginals, symbol),
                           cov_back=traversal_covariance(b, a, cov_dict, marginals, symbol), weight=w)
    return g


def get_good_candidates_typed(c_n_index, marginals, result, index_list, ref_graph, cov_dict, symbol=None):
    """Drop-in for loop_closure.get_good_candidates(c_n_index, marginals, result, index_list)
    (loop_closure.py:199-228; check_candidate :164-196): `result` needs atPose3(key).matrix(), `marginals`
    jointMarginalInformation (only for edges the dictionaries do not cover); ref_graph / cov_dict are the
    module globals cov_dijkstra_graph / relative_covariance_dict (loop_closure.py:27,30), read at call time by
    patch(batched_gating=True).  All candidates of the keyframe are gated in one launch."""
    if symbol is None:
        import gtsam
        symbol = gtsam.symbol
    graph = covariance_graph_from_reference(ref_graph, index_list, cov_dict, marginals, symbol)
    poses = np.stack([np.asarray(result.atPose3(symbol("c", int(f))).matrix(), dtype=np.float64)[:3, :4]
                      for f in index_list])
    dist, _ = gate_distances(poses, graph, [c_n_index], gap=KEY_FRAME_GAP)
    return select_candidates(dist[0], index_list)
