"""Device-resident batched front-end over a whole sequence (BASELINE config 2).

The reference walks the sequence one frame at a time (backend/database/database.py:30-89):
per frame a crossCheck L<->R match (matching.py:44), the row filter (matching.py:48-69),
create_links (tracking_database.py:224-246), then forward + backward matches against the previous
frame's filtered features (database.py:54-55) and the mutual check (database.py:67-77), plus
triangulation of the links (ransac.py:83 / triangulation.py:41-50).  Those steps are independent
per frame / frame pair once descriptors exist, so here ALL frames go through each step in one
ragged launch:

    1. L<->R Hamming top-2 + column minima, all frames          (slamfe_hamming_top2_batched)
    2. crossCheck + row filter + links + feature compaction     (slamfe_stereo_links_batched)
    3. triangulation of every link                              (slamfe_triangulate_links_f32)
    4. frame t <-> t+1 match on the compacted features, forward rows + backward columns from one
       pass                                                     (slamfe_hamming_top2_batched)

Layout in HBM: descriptors of all frames concatenated as cv2 lays them out ((rows, 61) uint8);
every frame starts at a row offset that is a multiple of 16 so that each frame's byte range is
16-byte aligned (TMA bulk copies); keypoints (rows, 2) float32 with the same offsets; results are
indexed by the same global rows (per-frame capacity = that frame's left row count).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import _cabi, ops
from .synth import DESC_BYTES

ROW_ALIGN = 16


def _round_up(x, m):
    return (x + m - 1) // m * m


@dataclass
class PackedSequence:
    """Host-side (pinned when `pin=True`) concatenation of per-frame inputs."""
    desc_l: np.ndarray   # (L, 61) uint8
    desc_r: np.ndarray   # (R, 61) uint8
    pts_l: np.ndarray    # (L, 2) float32
    pts_r: np.ndarray    # (R, 2) float32
    l_off: np.ndarray    # (F+1,) int32 row offsets (multiples of 16); frame f has n_l[f] valid rows
    r_off: np.ndarray
    n_l: np.ndarray      # (F,) int32
    n_r: np.ndarray
    tensors: dict | None = None  # pinned torch tensors backing the arrays (kept alive)

    @property
    def n_frames(self):
        return len(self.n_l)

    def h2d_bytes(self):
        return int(self.desc_l.nbytes + self.desc_r.nbytes + self.pts_l.nbytes + self.pts_r.nbytes
                   + self.l_off.nbytes + self.r_off.nbytes + self.n_l.nbytes + self.n_r.nbytes)


def plan_offsets(counts):
    counts = np.asarray(counts, dtype=np.int64)
    off = np.zeros(len(counts) + 1, dtype=np.int64)
    np.cumsum([_round_up(int(c), ROW_ALIGN) for c in counts], out=off[1:])
    if off[-1] >= 2 ** 31:
        raise ValueError("sequence too large for int32 row offsets")
    return off.astype(np.int32)


def pack_sequence(frames, pin=True) -> PackedSequence:
    """frames: iterable of (desc_l (nl,61) u8, desc_r (nr,61) u8, pts_l (nl,2) f32, pts_r (nr,2) f32)."""
    frames = list(frames)
    n_l = np.array([f[0].shape[0] for f in frames], dtype=np.int32)
    n_r = np.array([f[1].shape[0] for f in frames], dtype=np.int32)
    l_off, r_off = plan_offsets(n_l), plan_offsets(n_r)
    L, R = int(l_off[-1]), int(r_off[-1])
    tensors = None
    if pin:
        torch = _cabi.require_cuda()
        tensors = {
            "desc_l": torch.zeros((L, DESC_BYTES), dtype=torch.uint8, pin_memory=True),
            "desc_r": torch.zeros((R, DESC_BYTES), dtype=torch.uint8, pin_memory=True),
            "pts_l": torch.zeros((L, 2), dtype=torch.float32, pin_memory=True),
            "pts_r": torch.zeros((R, 2), dtype=torch.float32, pin_memory=True),
        }
        arrs = {k: v.numpy() for k, v in tensors.items()}
    else:
        arrs = {"desc_l": np.zeros((L, DESC_BYTES), np.uint8), "desc_r": np.zeros((R, DESC_BYTES), np.uint8),
                "pts_l": np.zeros((L, 2), np.float32), "pts_r": np.zeros((R, 2), np.float32)}
    for f, (dl, dr, pl, pr) in enumerate(frames):
        arrs["desc_l"][l_off[f]:l_off[f] + n_l[f]] = dl
        arrs["desc_r"][r_off[f]:r_off[f] + n_r[f]] = dr
        arrs["pts_l"][l_off[f]:l_off[f] + n_l[f]] = pl
        arrs["pts_r"][r_off[f]:r_off[f] + n_r[f]] = pr
    return PackedSequence(arrs["desc_l"], arrs["desc_r"], arrs["pts_l"], arrs["pts_r"], l_off, r_off, n_l, n_r, tensors)


@dataclass
class DeviceSequence:
    desc_l: object
    desc_r: object
    pts_l: object
    pts_r: object
    l_off: object      # (F+1,) int32 CUDA
    r_off: object
    n_l: object        # (F,) int32 CUDA
    n_r: object
    n_frames: int
    max_nl: int
    max_nr: int


def to_device(seq: PackedSequence, device="cuda", non_blocking=True) -> DeviceSequence:
    torch = _cabi.require_cuda()

    def up(name, arr):
        src = seq.tensors[name] if seq.tensors and name in seq.tensors else torch.from_numpy(arr)
        return src.to(device, non_blocking=non_blocking)

    return DeviceSequence(
        up("desc_l", seq.desc_l), up("desc_r", seq.desc_r), up("pts_l", seq.pts_l), up("pts_r", seq.pts_r),
        torch.from_numpy(seq.l_off).to(device, non_blocking=non_blocking),
        torch.from_numpy(seq.r_off).to(device, non_blocking=non_blocking),
        torch.from_numpy(seq.n_l).to(device, non_blocking=non_blocking),
        torch.from_numpy(seq.n_r).to(device, non_blocking=non_blocking),
        seq.n_frames, int(seq.n_l.max()) if seq.n_frames else 0, int(seq.n_r.max()) if seq.n_frames else 0)


# Tables run_host / results_to_host bring back by default.  `links` (x_left, x_right, y per link) stays on the
# device unless asked for: the host derives it from its own keypoints (links_from_tables), 12 bytes per row less.
RESULT_KEYS = ("match_t", "n_matches", "n_links", "link_src", "xyz", "fwd_keys", "bwd_keys")
ALL_RESULT_KEYS = RESULT_KEYS + ("links",)
TRACK_KEYS = ("inlier_fwd", "best", "n_good", "n_hyp", "n_hyp_full", "pose", "pose_status")


def chunk_bounds(n_frames, chunk_frames):
    """Frame boundaries of the host pipeline's chunks.  The first chunk's H2D copy and the last chunk's
    D2H copy are the only transfers that cannot hide behind kernels, so the sequence starts (and ends)
    with small chunks — chunk/8, chunk/4, chunk/2 — and runs at full size in between."""
    c = max(1, int(chunk_frames))
    ramp = [s for s in (c // 8, c // 4, c // 2) if s >= 16]
    bounds = [0]
    for s in ramp:
        if bounds[-1] + s + sum(ramp) + c > n_frames:   # keep at least one full chunk and the ramp-down
            break
        bounds.append(bounds[-1] + s)
    tail = []
    end = n_frames
    if len(bounds) > 1:
        for s in ramp:
            tail.append(end)
            end -= s
    pos = bounds[-1]
    while pos + c < end:
        pos += c
        bounds.append(pos)
    if end > bounds[-1] and end != n_frames:
        bounds.append(end)
    bounds += sorted(tail)
    if bounds[-1] != n_frames:
        bounds.append(n_frames)
    return sorted(set(bounds))


class FrontEnd:
    """Runs the four batched stages over a DeviceSequence; owns (and reuses) the output buffers.

    run(ds)          inputs already resident in HBM, everything on the current stream;
    run_host(seq)    host (pinned) inputs -> host (pinned) result tables: the sequence is cut into
                     chunks of frames and the H2D copy of chunk c+1, the kernels of chunk c and
                     the D2H copy of chunk c-1 run concurrently (copy-in, 2 compute, copy-out streams).
    """

    def __init__(self, K=None, M1=None, M2=None):
        """Cameras default to the ones the drop-in entry points score with at construction time
        (slamfe.ransac.K / M1 / M2, which patch() copies from the reference, ransac.py:11); the projection
        matrices of the triangulation stages are always derived from the same three (P = K @ M1,
        Q = K @ M2, ransac.py:12), so matching, triangulation, hypotheses and scoring cannot disagree."""
        from . import ransac as _ransac
        self.K = np.asarray(_ransac.K if K is None else K, dtype=np.float64).reshape(3, 3)
        self.M1 = np.asarray(_ransac.M1 if M1 is None else M1, dtype=np.float64).reshape(3, 4)
        self.M2 = np.asarray(_ransac.M2 if M2 is None else M2, dtype=np.float64).reshape(3, 4)
        self.P, self.Q = self.K @ self.M1, self.K @ self.M2
        self._out = None
        self._key = None
        self._trk = None
        self._trk_key = None
        self._in = None
        self._in_key = None
        self._pinned_out = None
        self._pinned_ids = None
        self._pinned_db = None
        self._streams = None
        # kernels launched by one run(): 2 matcher launches, stereo epilogue, triangulation
        # (cudaMemsetAsync initialisation of the key tables is not counted)
        self.launches_per_run = 4
        self.last_launches = 0
        self.last_truncated = 0   # pairs whose RANSAC iteration count exceeded h_max in the last run
        self.trace = None         # set to [] to collect (stage, chunk, start event, end event) of run_host

    def _buffers(self, L, R, F, dev):
        torch = _cabi.require_cuda()
        key = (L, R, F, dev)
        if self._key != key:
            i32 = dict(dtype=torch.int32, device=dev)
            self._out = {
                "lr_row_keys": torch.empty((L, 2), **i32), "lr_col_keys": torch.empty((R,), **i32),
                "match_t": torch.empty((L,), **i32), "n_matches": torch.empty((F,), **i32),
                "n_links": torch.empty((F,), **i32), "link_src": torch.empty((L,), **i32),
                "links": torch.empty((L, 3), dtype=torch.float32, device=dev),
                "feat": torch.empty((L, 64), dtype=torch.uint8, device=dev),
                "xyz": torch.empty((L, 3), dtype=torch.float32, device=dev),
                "fwd_keys": torch.empty((L,), **i32), "bwd_keys": torch.empty((L,), **i32),
            }
            self._key = key
        return self._out

    # -- stages on a frame range; all tensors are views whose offsets are relative to the view --
    @staticmethod
    def _stereo_stages(P, Q, desc_l, desc_r, pts_l, pts_r, l_off, r_off, n_l, n_r, n, max_nl, max_nr, o):
        """Stages 1-3 for n frames: o holds views of the output tables for the same rows/frames."""
        # 1. stereo match, every frame (best neighbour + column minima: crossCheck needs no more)
        ops.hamming_top2_batched(desc_l, l_off, desc_r, r_off, n, max_nl, max_nr, DESC_BYTES,
                                 q_cnt=n_l, t_cnt=n_r, want_cols=True, best_only=True,
                                 row_keys=o["lr_row_keys"], col_keys=o["lr_col_keys"])
        # 2. crossCheck + row filter + links + features[is_valid]
        ops.stereo_links_batched(o["lr_row_keys"], o["lr_col_keys"], l_off, r_off, n, pts_l, pts_r,
                                 desc_left=desc_l, desc_bytes=DESC_BYTES, out=o, n_l=n_l, n_r=n_r)
        # 3. triangulate every link
        ops.triangulate_links(o["links"], P, Q, out=o["xyz"])

    @staticmethod
    def _pair_stage(feat_q, q_off, q_cnt, feat_t, t_off, t_cnt, n_pairs, max_links, fwd_keys, bwd_keys):
        """Stage 4: problem p matches the filtered features of frame p (rows of feat_q) against
        those of frame p+1 (rows of feat_t); forward rows + backward columns from one pass."""
        ops.hamming_top2_batched(feat_q, q_off, feat_t, t_off, n_pairs, max_links, max_links, DESC_BYTES,
                                 q_cnt=q_cnt, t_cnt=t_cnt, want_cols=True, best_only=True, compact=True,
                                 row_keys=fwd_keys, col_keys=bwd_keys)

    def run(self, ds: DeviceSequence):
        """All stages, asynchronous on the current stream.  Returns the dict of output tensors."""
        o = self._buffers(ds.desc_l.shape[0], ds.desc_r.shape[0], ds.n_frames, ds.desc_l.device)
        F = ds.n_frames
        self.last_launches = 0
        if F == 0:
            return o
        self._stereo_stages(self.P, self.Q, ds.desc_l, ds.desc_r, ds.pts_l, ds.pts_r, ds.l_off, ds.r_off,
                            ds.n_l, ds.n_r, F, ds.max_nl, ds.max_nr, o)
        self.last_launches += 3
        # 4. consecutive frames on the filtered features: problem f = (frame f, frame f+1)
        if F > 1:
            max_links = min(ds.max_nl, ds.max_nr)
            self._pair_stage(o["feat"], ds.l_off, o["n_links"], o["feat"], ds.l_off[1:], o["n_links"][1:],
                             F - 1, max_links, o["fwd_keys"], o["bwd_keys"])
            self.last_launches += 1
        return o

    # -- stages 5-8: the rest of the create_db loop body (database.py:54-85), device resident -----
    def _track_buffers(self, L, F, h_max, dev):
        torch = _cabi.require_cuda()
        key = (L, F, h_max, dev)
        if self._trk_key != key:
            i32 = dict(dtype=torch.int32, device=dev)
            f64 = dict(dtype=torch.float64, device=dev)
            n = max(F - 1, 1)
            self._trk = {
                "good_j": torch.empty((L,), **i32), "good_t": torch.empty((L,), **i32),
                "n_good": torch.empty((n,), **i32), "n_hyp": torch.empty((n,), **i32),
                "n_hyp_full": torch.empty((n,), **i32),
                "pts": torch.empty((L, 3), **f64), "lpix": torch.empty((L, 2), **f64),
                "rpix": torch.empty((L, 2), **f64),
                "T": torch.empty((n * h_max, 3, 4), **f64), "hyp_valid": torch.empty((n * h_max,), dtype=torch.uint8,
                                                                                     device=dev),
                "counts": torch.empty((n, h_max), **i32), "best": torch.empty((n, 2), **i32),
                "best_mask": torch.empty((L,), dtype=torch.uint8, device=dev), "work": torch.empty((n,), **i32),
                "inlier_fwd": torch.empty((L,), dtype=torch.uint8, device=dev),
                "pose": torch.empty((n, 3, 4), **f64), "pose_status": torch.empty((n,), **i32),
                "pose_rms": torch.empty((n,), **f64),
            }
            self._trk_key = key
        return self._trk

    def _track_stages(self, o, t, l_off, r_off, pts_l, pts_r, n_pairs, max_links, h_max, seed, pair_base=0):
        """Stages 5-8 for n_pairs consecutive pairs; o / t hold (views of) the pipeline and tracking
        tables whose row 0 is the first row of the first pair's previous frame."""
        ops.track_gather(o, l_off, r_off, pts_l, pts_r, n_pairs, self.P, self.Q, h_max, t)
        ops.ransac_hypotheses(t["pts"], t["lpix"], self.K, h_max, seed=seed, pt_off=l_off, pt_cnt=t["n_good"],
                              n_frames=n_pairs, n_hyp=t["n_hyp"], out=(t["T"], t["hyp_valid"]),
                              frame_index_base=pair_base)
        ops.ransac_score(t["T"], t["pts"], t["lpix"], t["rpix"], self.K, self.M1, self.M2, hyp_valid=t["hyp_valid"],
                         pt_off=l_off, pt_cnt=t["n_good"], n_frames=n_pairs, max_points=max_links, out=t)
        ops.scatter_inliers(t["best_mask"], t["good_j"], l_off, t["n_good"], t["best"], n_pairs, t["inlier_fwd"])
        # relative pose of every pair: refit on the consensus set (ransac.py:185-193), no host solve
        ops.pnp_refit(t["T"], t["best"], t["pts"], t["lpix"], t["best_mask"], self.K, pt_off=l_off, pt_cnt=t["n_good"],
                      n_frames=n_pairs, out={"T_refit": t["pose"], "refit_status": t["pose_status"],
                                             "refit_rms": t["pose_rms"]})

    def track(self, ds: DeviceSequence, h_max=256, seed=1, full_ransac=False, track_ids=False):
        """run(ds) followed by the frame-to-frame tracking of database.py:54-85 for every consecutive
        pair, without leaving the device: mutual forward/backward check + link gather + fp64
        triangulation of the previous links (slamfe_track_gather), RANSAC-PnP hypothesis generation
        (slamfe_ransac_hypotheses; per-pair iteration count from calc_ransac_iteration, capped at
        h_max), scoring of all hypotheses of all pairs in one launch (slamfe_ransac_score) and the
        inlier flags per forward match (in_prev_cur, database.py:84-85).  Returns the output dict of
        run() extended with good_j, good_t, n_good, n_hyp, pts, lpix, rpix, T, hyp_valid, counts, best,
        best_mask, inlier_fwd, n_hyp_full, pose (world-to-camera [R|t] of frame f+1 relative to frame f, refit
        on the consensus set by slamfe_pnp_refit), pose_status, pose_rms.

        h_max caps the hypotheses per pair of the batched launch; n_hyp_full holds the reference's
        uncapped count (calc_ransac_iteration, ransac.py:59-67).  full_ransac=True synchronises and re-runs
        every pair with n_hyp_full > h_max at its full count (rescore_truncated), so the result never
        depends on h_max; with False the caller can inspect n_hyp_full itself.  track_ids=True adds track_id
        (L,) / n_tracks (1,) / head_base (F + 1,): the track every link belongs to, numbered as
        TrackingDB.add_frame numbers them (slamfe_track_ids)."""
        o = self.run(ds)
        F = ds.n_frames
        L = ds.desc_l.shape[0]
        t = self._track_buffers(L, F, h_max, ds.desc_l.device)
        out = dict(o)
        out.update(t)
        if F < 2 or L == 0:  # nothing to track: report "no mutual matches, no hypothesis" for every pair
            t["n_good"].zero_(); t["n_hyp"].zero_(); t["n_hyp_full"].zero_(); t["best"].fill_(-1)
            t["best"][:, 1].zero_()
            t["inlier_fwd"].zero_(); t["pose"].zero_(); t["pose_status"].zero_()
            self.last_truncated = 0
            if track_ids:   # nothing continues into a next frame: no tracks
                torch = _cabi.require_cuda()
                dev = ds.desc_l.device
                out["track_id"] = torch.full((L,), -1, dtype=torch.int32, device=dev)
                out["n_tracks"] = torch.zeros((1,), dtype=torch.int32, device=dev)
                out["head_base"] = torch.zeros((F + 1,), dtype=torch.int32, device=dev)
            return out
        self._track_stages(o, t, ds.l_off, ds.r_off, ds.pts_l, ds.pts_r, F - 1, min(ds.max_nl, ds.max_nr), h_max, seed)
        self.last_launches += 5
        if full_ransac:
            self.rescore_truncated(ds.l_off.cpu().numpy(), F - 1, h_max, seed)
        if track_ids:   # TrackingDB.add_frame's bookkeeping for the whole sequence (tracking_database.py:273-337)
            ops.track_ids(o["fwd_keys"], t["inlier_fwd"], ds.l_off, o["n_links"], F, out=t)
            out.update({k: t[k] for k in ("track_id", "n_tracks", "head_base")})
            self.last_launches += 4
        return out

    def rescore_truncated(self, l_off, n_pairs, h_max, seed=1, host_tables=None, h_cap=1 << 16):
        """Re-run RANSAC-PnP at the reference's FULL iteration count for every pair the batched launch
        truncated (n_hyp_full > h_max), on the device tables of the last track() / run_host():
        hypotheses 0..n_hyp_full-1 of pair f are the same pure function of (seed, f, h) the batched
        launch used, so the outcome equals a run with h_max >= n_hyp_full.  Synchronises.  Patches
        best / best_mask / inlier_fwd on the device and, when given, in the host tables.  Returns the
        list of re-run pairs.  A count above h_cap raises instead (calc_ransac_iteration grows without
        bound as the stereo inlier rate falls: 14 389 at 20 %, 230 257 at 10 %)."""
        torch = _cabi.require_cuda()
        t = self._trk
        full = t["n_hyp_full"][:n_pairs].cpu().numpy()
        n_good = t["n_good"][:n_pairs].cpu().numpy()
        todo = [int(f) for f in np.nonzero(full > h_max)[0]]
        self.last_truncated = len(todo)
        if not todo:
            return todo
        worst = int(full[todo].max())
        if worst > h_cap:
            raise _cabi.SlamfeError(f"calc_ransac_iteration asks for {worst} hypotheses (> {h_cap}) on pair "
                                    f"{todo[int(np.argmax(full[todo]))]}: stereo inlier rate too low")
        l_off = np.asarray(l_off, dtype=np.int64)
        zero = torch.zeros((1,), dtype=torch.int32, device=t["pts"].device)
        for f in todo:
            l0, n, cap = int(l_off[f]), int(n_good[f]), int(l_off[f + 1] - l_off[f])
            if n < 4:
                continue
            H = int(full[f])
            pts, lp, rp = t["pts"][l0:l0 + n], t["lpix"][l0:l0 + n], t["rpix"][l0:l0 + n]
            T, valid = ops.ransac_hypotheses(pts, lp, self.K, H, seed=seed, n_frames=1, frame_index_base=f)
            _, best, mask = ops.ransac_score(T, pts, lp, rp, self.K, self.M1, self.M2, hyp_valid=valid)
            t["best"][f:f + 1].copy_(best)
            t["best_mask"][l0:l0 + n].copy_(mask)
            ops.scatter_inliers(mask, t["good_j"][l0:l0 + n], zero, t["n_good"][f:f + 1], best, 1,
                                t["inlier_fwd"][l0:l0 + cap])
            ops.pnp_refit(T, best, pts, lp, mask, self.K, out={"T_refit": t["pose"][f:f + 1],
                                                               "refit_status": t["pose_status"][f:f + 1],
                                                               "refit_rms": t["pose_rms"][f:f + 1]})
            self.last_launches += 4
            if host_tables is not None:
                for k in ("pose", "pose_status"):
                    if k in host_tables:
                        host_tables[k][f] = t[k][f].cpu().numpy()
                if "best" in host_tables:
                    host_tables["best"][f] = best[0].cpu().numpy()
                if "inlier_fwd" in host_tables:
                    host_tables["inlier_fwd"][l0:l0 + cap] = t["inlier_fwd"][l0:l0 + cap].cpu().numpy()
        torch.cuda.synchronize()
        return todo

    # -- host in, host out ---------------------------------------------------------------------
    def _input_buffers(self, seq: PackedSequence, dev):
        torch = _cabi.require_cuda()
        L, R, F = seq.desc_l.shape[0], seq.desc_r.shape[0], seq.n_frames
        key = (L, R, F, dev)
        if self._in_key != key:
            self._in = {
                "desc_l": torch.empty((L, DESC_BYTES), dtype=torch.uint8, device=dev),
                "desc_r": torch.empty((R, DESC_BYTES), dtype=torch.uint8, device=dev),
                "pts_l": torch.empty((L, 2), dtype=torch.float32, device=dev),
                "pts_r": torch.empty((R, 2), dtype=torch.float32, device=dev),
            }
            self._in_key = key
            self._pinned_out = None
        return self._in

    def run_host(self, seq: PackedSequence, chunk_frames=288, device="cuda", keys=RESULT_KEYS, track=False,
                 h_max=256, seed=1, full_ransac=True, track_ids=False, track_keys=TRACK_KEYS, pack_db=False):
        """Pinned host inputs -> pinned host result tables, copies overlapped with the kernels.

        Returns (dict of numpy views of the pinned result tables, h2d_bytes, d2h_bytes).  The call
        returns after the last table has landed in host memory.  track=True adds the tracking stages
        of track() per chunk and the tables TRACK_KEYS (inlier flags per forward match, best
        hypothesis / inlier count, mutual-match count and capped / uncapped RANSAC iteration count per
        pair).  full_ransac=True (default) re-runs the pairs h_max truncated at their full count before
        returning (rescore_truncated; self.last_truncated = how many).  track_ids=True (needs track) adds the
        tables track_id (L,) int32 and n_tracks (1,): computed once over the whole sequence after the last
        chunk (tracks cross chunk boundaries), after any re-run of truncated pairs.  track_keys: which of the
        TRACK_KEYS tables travel back (n_hyp_full always does).  pack_db=True (needs track_ids) adds the DENSE
        tracking-database columns db_link_off / db_x_left / db_x_right / db_y / db_feat / db_track_id
        (slamfe_pack_db) — with keys=("n_matches", "n_links") and a short track_keys list that is all
        slamfe.trackdb needs, and far less D2H than the padded tables."""
        torch = _cabi.require_cuda()
        if track_ids and not track:
            raise ValueError("track_ids needs track=True")
        if pack_db and not track_ids:
            raise ValueError("pack_db needs track_ids=True")
        if seq.tensors is None:
            raise ValueError("run_host needs a pinned PackedSequence (pack_sequence(..., pin=True))")
        dev = torch.device(device)
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        F = seq.n_frames
        L, R = seq.desc_l.shape[0], seq.desc_r.shape[0]
        din = self._input_buffers(seq, dev)
        o = self._buffers(L, R, F, dev)
        trk = self._track_buffers(L, F, h_max, dev) if track else None
        if track:
            keys = tuple(keys) + tuple(k for k in tuple(track_keys) + ("n_hyp_full",) if k not in keys)
            o = dict(o)
            o.update(trk)
        if self._pinned_out is None or any(k not in self._pinned_out or self._pinned_out[k].shape != o[k].shape
                                           for k in keys):
            self._pinned_out = {k: torch.zeros(o[k].shape, dtype=o[k].dtype, pin_memory=True) for k in keys}
            if "best" in self._pinned_out:  # pairs that are never computed (single-frame sequence) read "no hypothesis"
                self._pinned_out["best"][:, 0] = -1
        hout = self._pinned_out
        if self._streams is None:
            self._streams = tuple(torch.cuda.Stream(device=dev) for _ in range(4))
        s_in, s_cmp_a, s_cmp_b, s_out = self._streams
        self.last_launches = 0
        if F == 0:
            return {k: hout[k].numpy() for k in keys}, 0, 0
        l_off, r_off = seq.l_off.astype(np.int64), seq.r_off.astype(np.int64)
        max_nl, max_nr = int(seq.n_l.max()), int(seq.n_r.max())
        max_links = min(max_nl, max_nr)
        bounds = chunk_bounds(F, chunk_frames)
        # chunk-relative offset tables for every chunk, one small upload
        parts, index, pos = [], [], 0
        for c in range(len(bounds) - 1):
            f0, f1 = bounds[c], bounds[c + 1]
            p0 = max(f0 - 1, 0)
            tabs = (l_off[f0:f1 + 1] - l_off[f0], r_off[f0:f1 + 1] - r_off[f0],
                    l_off[p0:f1 - 1] - l_off[p0], l_off[p0 + 1:f1] - l_off[p0 + 1],
                    l_off[p0:f1] - l_off[p0], r_off[p0:f1] - r_off[p0])
            loc = []
            for t in tabs:
                loc.append((pos, pos + len(t)))
                parts.append(t)
                pos += len(t)
            index.append(loc)
        small = np.concatenate(parts + [seq.n_l.astype(np.int64), seq.n_r.astype(np.int64)]).astype(np.int32)
        small_pin = torch.from_numpy(small).pin_memory()
        cur = torch.cuda.current_stream(dev)
        for s in self._streams:
            s.wait_stream(cur)
        h2d = d2h = 0
        with torch.cuda.stream(s_in):
            small_dev = small_pin.to(dev, non_blocking=True)
            h2d += small_pin.numel() * 4
        n_l_dev, n_r_dev = small_dev[pos:pos + F], small_dev[pos + F:pos + 2 * F]
        ev_done = []
        ev_links_prev = None
        tr = self.trace

        def mark(stream):
            if tr is None:
                return None
            e = torch.cuda.Event(enable_timing=True)
            e.record(stream)
            return e

        # The copy-in stream is the critical resource (the step is PCIe-bound): all of its copies are enqueued
        # before any kernel, so it never waits for the host to finish launching a chunk's kernels.
        ev_in_all = []
        with torch.cuda.stream(s_in):
            for c in range(len(bounds) - 1):
                f0, f1 = bounds[c], bounds[c + 1]
                a, b = int(l_off[f0]), int(l_off[f1])
                ra, rb = int(r_off[f0]), int(r_off[f1])
                e0 = mark(s_in)
                for k, lo, hi in (("desc_l", a, b), ("pts_l", a, b), ("desc_r", ra, rb), ("pts_r", ra, rb)):
                    src = seq.tensors[k][lo:hi]
                    din[k][lo:hi].copy_(src, non_blocking=True)
                    h2d += src.numel() * src.element_size()
                ev_in = torch.cuda.Event(enable_timing=tr is not None)
                ev_in.record(s_in)
                ev_in_all.append(ev_in)
                if tr is not None:
                    tr.append(("h2d", c, e0, ev_in))
        for c in range(len(bounds) - 1):
            # chunks alternate between two compute streams, so the tail of one chunk's matcher
            # launch (few long CTAs left) overlaps the head of the next chunk's
            s_cmp = s_cmp_a if c % 2 == 0 else s_cmp_b
            f0, f1 = bounds[c], bounds[c + 1]
            n = f1 - f0
            a, b = int(l_off[f0]), int(l_off[f1])
            ra, rb = int(r_off[f0]), int(r_off[f1])
            ev_in = ev_in_all[c]
            (l0, l1), (r0, r1), (q0, q1), (t0, t1), (lp0, lp1), (rp0, rp1) = index[c]
            with torch.cuda.stream(s_cmp):
                s_cmp.wait_event(ev_in)
                e0 = mark(s_cmp)
                view = {k: o[k][a:b] for k in ("lr_row_keys", "match_t", "link_src", "links", "feat", "xyz")}
                view["lr_col_keys"] = o["lr_col_keys"][ra:rb]
                view["n_matches"], view["n_links"] = o["n_matches"][f0:f1], o["n_links"][f0:f1]
                self._stereo_stages(self.P, self.Q, din["desc_l"][a:b], din["desc_r"][ra:rb], din["pts_l"][a:b],
                                    din["pts_r"][ra:rb], small_dev[l0:l1], small_dev[r0:r1], n_l_dev[f0:f1],
                                    n_r_dev[f0:f1], n, max_nl, max_nr, view)
                self.last_launches += 3
                ev_links = torch.cuda.Event()
                ev_links.record(s_cmp)
                if ev_links_prev is not None:  # frame f0-1's features come from the other stream
                    s_cmp.wait_event(ev_links_prev)
                ev_links_prev = ev_links
                p0 = max(f0 - 1, 0)
                n_pairs = f1 - 1 - p0
                if n_pairs > 0:
                    qa, qb = int(l_off[p0]), int(l_off[f1 - 1])   # rows of frames p0 .. f1-2
                    ta = int(l_off[p0 + 1])                         # rows of frames p0+1 .. f1-1
                    self._pair_stage(o["feat"][qa:qb], small_dev[q0:q1], o["n_links"][p0:f1 - 1],
                                     o["feat"][ta:b], small_dev[t0:t1], o["n_links"][p0 + 1:f1],
                                     n_pairs, max_links, o["fwd_keys"][qa:qb], o["bwd_keys"][ta:b])
                    self.last_launches += 1
                    if track:  # rows / frames based at frame p0; pairs p0 .. f1-2
                        ra0 = int(r_off[p0])
                        ov = {"fwd_keys": o["fwd_keys"][qa:b], "bwd_keys": o["bwd_keys"][qa:b],
                              "n_links": o["n_links"][p0:f1], "n_matches": o["n_matches"][p0:f1],
                              "link_src": o["link_src"][qa:b], "match_t": o["match_t"][qa:b]}
                        tv = {k: trk[k][qa:qb] for k in ("good_j", "good_t", "pts", "lpix", "rpix", "best_mask",
                                                         "inlier_fwd")}
                        tv.update({k: trk[k][p0:f1 - 1] for k in ("n_good", "n_hyp", "n_hyp_full", "counts", "best", "work",
                                                                  "pose", "pose_status", "pose_rms")})
                        tv["T"] = trk["T"][p0 * h_max:(f1 - 1) * h_max]
                        tv["hyp_valid"] = trk["hyp_valid"][p0 * h_max:(f1 - 1) * h_max]
                        self._track_stages(ov, tv, small_dev[lp0:lp1], small_dev[rp0:rp1], din["pts_l"][qa:b],
                                           din["pts_r"][ra0:rb], n_pairs, max_links, h_max, seed, pair_base=p0)
                        self.last_launches += 5
                if f1 == F:  # the last frame has no successor, frame 0 no predecessor
                    o["fwd_keys"][int(l_off[F - 1]):].fill_(-1)
                    if track:
                        trk["inlier_fwd"][int(l_off[F - 1]):].zero_()
                if f0 == 0:
                    o["bwd_keys"][:int(l_off[1])].fill_(-1)
                ev_cmp = torch.cuda.Event(enable_timing=tr is not None)
                ev_cmp.record(s_cmp)
                if tr is not None:
                    tr.append(("compute", c, e0, ev_cmp))
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_cmp)
                e0 = mark(s_out)
                # complete after this chunk: per-row tables of its frames, forward keys up to f1-2
                fa = int(l_off[max(f0 - 1, 0)])
                fb = b if f1 == F else int(l_off[f1 - 1])
                for k in keys:
                    if k in ("n_matches", "n_links"):
                        lo, hi = f0, f1
                    elif k in ("best", "n_good", "n_hyp", "n_hyp_full", "pose", "pose_status"):   # per pair: p0 .. f1-2
                        lo, hi = max(f0 - 1, 0), f1 - 1
                    elif k in ("fwd_keys", "inlier_fwd"):
                        lo, hi = fa, fb
                    else:
                        lo, hi = a, b
                    if hi > lo:
                        hout[k][lo:hi].copy_(o[k][lo:hi], non_blocking=True)
                        d2h += (hi - lo) * o[k][0:1].numel() * o[k].element_size()
                ev = torch.cuda.Event(enable_timing=tr is not None)
                ev.record(s_out)
                ev_done.append(ev)
                if tr is not None:
                    tr.append(("d2h", c, e0, ev))
        ev_done[-1].synchronize()
        for s in self._streams:
            cur.wait_stream(s)
        tables = {k: hout[k].numpy() for k in keys}
        self.last_truncated = 0
        if track and F > 1:
            if full_ransac and bool((tables["n_hyp_full"][:F - 1] > h_max).any()):
                self.rescore_truncated(l_off, F - 1, h_max, seed, host_tables=tables)
            else:
                self.last_truncated = int((tables["n_hyp_full"][:F - 1] > h_max).sum())
        if track_ids:
            l_off_dev = small_dev.new_tensor(seq.l_off) if F else None
            tid, n_tr, _ = ops.track_ids(o["fwd_keys"], trk["inlier_fwd"], l_off_dev, o["n_links"], F, out=trk)
            self.last_launches += 4
            pin = self._pinned_ids
            if pin is None or pin[0].shape != tid.shape:
                pin = self._pinned_ids = (torch.empty(tid.shape, dtype=tid.dtype, pin_memory=True),
                                          torch.empty((1,), dtype=torch.int32, pin_memory=True))
            pin[0].copy_(tid, non_blocking=True)
            pin[1].copy_(n_tr, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
            tables["track_id"], tables["n_tracks"] = pin[0].numpy(), pin[1].numpy()
            d2h += tid.numel() * 4 + 4
            if pack_db:
                r_off_dev = small_dev.new_tensor(seq.r_off)
                db = ops.pack_db(o, l_off_dev, r_off_dev, din["pts_l"], din["pts_r"], F, DESC_BYTES, track_id=tid, out=trk)
                self.last_launches += 2
                n_total = int(np.asarray(tables["n_links"][:F], dtype=np.int64).sum())
                if self._pinned_db is None or self._pinned_db["x_left"].shape[0] < n_total:
                    cap = max(n_total, 1)
                    self._pinned_db = {k: torch.empty((cap,) + tuple(v.shape[1:]), dtype=v.dtype, pin_memory=True)
                                       for k, v in db.items() if k != "link_off"}
                    self._pinned_db["link_off"] = torch.empty((0,), dtype=torch.int32, pin_memory=True)
                if self._pinned_db["link_off"].shape[0] != F + 1:
                    self._pinned_db["link_off"] = torch.empty((F + 1,), dtype=torch.int32, pin_memory=True)
                for k, v in db.items():
                    n_rows = F + 1 if k == "link_off" else n_total
                    self._pinned_db[k][:n_rows].copy_(v[:n_rows], non_blocking=True)
                    tables["db_" + k] = self._pinned_db[k][:n_rows].numpy()
                    d2h += n_rows * v[0:1].numel() * v.element_size()
                torch.cuda.current_stream(dev).synchronize()
        return tables, int(h2d), int(d2h)


def links_from_tables(seq: PackedSequence, tables, f):
    """The `links` rows of frame f — [x_left, x_right, (yl + yr) / 2] per link, float32 as the device table
    holds them (tracking_database.py:243) — from the host's own keypoints and the match_t / link_src tables, so
    the (L, 3) table need not travel back."""
    lo, ro, k = int(seq.l_off[f]), int(seq.r_off[f]), int(tables["n_links"][f])
    src = np.asarray(tables["link_src"][lo:lo + k], dtype=np.int64)
    dst = np.asarray(tables["match_t"][lo:lo + int(seq.n_l[f])], dtype=np.int64)[src]
    pl, pr = np.asarray(seq.pts_l[lo:lo + int(seq.n_l[f])]), np.asarray(seq.pts_r[ro:ro + int(seq.n_r[f])])
    y = (pl[src, 1].astype(np.float64) + pr[dst, 1].astype(np.float64)) / 2
    return np.stack([pl[src, 0], pr[dst, 0], y.astype(np.float32)], axis=1).astype(np.float32) if k else \
        np.zeros((0, 3), np.float32)


def descriptor_pairs(n_l, n_r, n_links=None) -> int:
    """Algorithmic descriptor pairs of one pass: sum Nl*Nr (stereo; one pass yields both
    directions) + sum links_f * links_{f+1} (consecutive frames on the filtered features)."""
    total = int(np.sum(np.asarray(n_l, dtype=np.int64) * np.asarray(n_r, dtype=np.int64)))
    if n_links is not None and len(n_links) > 1:
        k = np.asarray(n_links, dtype=np.int64)
        total += int(np.sum(k[:-1] * k[1:]))
    return total


def results_to_host(o, keys=RESULT_KEYS, pinned=None):
    """Copy the result tables to (pinned) host memory; returns (dict of numpy arrays, bytes, pinned)."""
    torch = _cabi.require_cuda()
    if pinned is None:
        pinned = {}
    nbytes = 0
    for k in keys:
        t = o[k]
        if k not in pinned or pinned[k].shape != t.shape:
            pinned[k] = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        pinned[k].copy_(t, non_blocking=True)
        nbytes += t.numel() * t.element_size()
    torch.cuda.current_stream().synchronize()
    return {k: pinned[k].numpy() for k in keys}, nbytes, pinned
