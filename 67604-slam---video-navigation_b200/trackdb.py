"""TrackingDB as flat arrays.

The reference's store (backend/database/tracking_database.py:75-471) is a set of Python dicts —
(frameId, trackId) -> Link, trackId -> [frameId], frameId -> descriptor array / track-id list /
leftover links — filled one `add_frame` at a time (:273-337) and pickled (:340-408).  Everything it
holds is a function of four per-link columns, which the batched pipeline already has as tables:

    frame_off  (F + 1,)   first link of every frame (links of a frame = its stereo-filtered features,
                          in the order `create_links` emits them, tracking_database.py:224-246)
    x_left, x_right (N,)  float32 keypoint abscissae;  y (N,) float64 = (yl + yr) / 2   (:243)
    features   (N, 61)    uint8 left descriptors (`features[is_valid]`, :235)
    track_id   (N,)       int32, -1 = NO_ID: computed on the device for the whole sequence by
                          slamfe_track_ids (csrc/trackdb.cu) with add_frame's own numbering

`SoATrackingDB` keeps those columns, answers the reference's read API from them (tracks(), frames(),
track(), link(), links(), features(), all_frame_links(), ... same names and return shapes), writes /
reads one .npz file, and `to_reference_db()` rebuilds the reference's dict-of-lists object — proven
equal to the UNMODIFIED TrackingDB built by the reference's own create_db loop
(tests/test_reference_db.py).  `build()` goes from FrontEnd.run_host's host tables to the store without
a per-match Python object.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import _cabi

NO_ID = -1  # tracking_database.py:9
FORMAT_VERSION = 1


class Link:
    """tracking_database.Link (:12-29): same fields and helpers."""
    __slots__ = ("x_left", "x_right", "y")

    def __init__(self, x_left, x_right, y):
        self.x_left, self.x_right, self.y = x_left, x_right, y

    def left_keypoint(self):
        return np.array([self.x_left, self.y])

    def right_keypoint(self):
        return np.array([self.x_right, self.y])

    def __str__(self):
        return f"Link (xl={self.x_left}, xr={self.x_right}, y={self.y})"

    def __eq__(self, other):
        return (self.x_left, self.x_right, self.y) == (other.x_left, other.x_right, other.y)


@dataclass
class SoATrackingDB:
    frame_off: np.ndarray         # (F + 1,) int64
    x_left: np.ndarray            # (N,) float32
    x_right: np.ndarray           # (N,) float32
    y: np.ndarray                 # (N,) float64
    feat: np.ndarray              # (N, W) uint8
    track_id: np.ndarray          # (N,) int32, NO_ID = -1
    inliers_percent: np.ndarray   # (F,) float64: frameID_to_inliers_percent (database.py:26)
    n_tracks: int = 0
    link_factory: type = Link
    _csr: tuple | None = None     # (track_off, rows sorted by (track, frame)) built on first use

    # ---- sizes (tracking_database.py:125-143) -------------------------------------------------
    def frame_num(self) -> int:
        return len(self.frame_off) - 1

    def all_frames(self):
        return range(self.frame_num())

    def track_num(self) -> int:
        return int(self.n_tracks)

    def all_tracks(self):
        return list(range(self.n_tracks))

    def link_num(self) -> int:
        return int(np.count_nonzero(self.track_id != NO_ID))

    @property
    def last_frameId(self) -> int:
        return self.frame_num() - 1

    @property
    def last_trackId(self) -> int:
        return self.n_tracks - 1

    # ---- per frame ----------------------------------------------------------------------------
    def _rows(self, frameId):
        return int(self.frame_off[frameId]), int(self.frame_off[frameId + 1])

    def _link(self, row):
        return self.link_factory(float(self.x_left[row]), float(self.x_right[row]), float(self.y[row]))

    def features(self, frameId):
        if not 0 <= frameId < self.frame_num():
            return None
        a, b = self._rows(frameId)
        return self.feat[a:b]

    def last_features(self):
        return self.features(self.last_frameId)

    def tracks(self, frameId):
        """Sorted track ids on frameId (tracking_database.py:118-122)."""
        if not 0 <= frameId < self.frame_num():
            return []
        a, b = self._rows(frameId)
        t = self.track_id[a:b]
        return sorted(t[t != NO_ID].tolist())

    def frame_track_ids(self, frameId):
        """frameId_to_trackIds_list[frameId]: one id (or NO_ID) per feature row."""
        a, b = self._rows(frameId)
        return self.track_id[a:b]

    def all_frame_links(self, frameId):
        """Every link of the frame, on a track or not, in feature order (tracking_database.py:171-188)."""
        a, b = self._rows(frameId)
        return [self._link(r) for r in range(a, b)]

    def all_last_frame_links(self):
        return self.all_frame_links(self.last_frameId)

    def frame_link_array(self, frameId):
        """(k, 3) float64 [x_left, x_right, y] of the frame: the array form the drop-in triangulation and
        RANSAC entry points take directly."""
        a, b = self._rows(frameId)
        return np.stack([self.x_left[a:b].astype(np.float64), self.x_right[a:b].astype(np.float64), self.y[a:b]], axis=1)

    def links(self, frameId):
        """trackId -> Link for the links of frameId that are on a track (tracking_database.py:157-163)."""
        a, b = self._rows(frameId)
        t = self.track_id[a:b]
        return {int(t[k]): self._link(a + k) for k in np.nonzero(t != NO_ID)[0]}

    def link(self, frameId, trackId):
        a, b = self._rows(frameId)
        k = np.nonzero(self.track_id[a:b] == trackId)[0]
        return self._link(a + int(k[0])) if len(k) and trackId != NO_ID else None

    # ---- per track ----------------------------------------------------------------------------
    def _track_csr(self):
        if self._csr is None:
            rows = np.nonzero(self.track_id != NO_ID)[0]
            order = np.argsort(self.track_id[rows], kind="stable")   # rows are frame-major already
            rows = rows[order]
            off = np.zeros(self.n_tracks + 1, dtype=np.int64)
            np.cumsum(np.bincount(self.track_id[rows], minlength=self.n_tracks), out=off[1:])
            self._csr = (off, rows)
        return self._csr

    def _frame_of_rows(self, rows):
        return np.searchsorted(self.frame_off, rows, side="right") - 1

    def track_rows(self, trackId):
        off, rows = self._track_csr()
        if not 0 <= trackId < self.n_tracks:
            return np.zeros(0, dtype=np.int64)
        return rows[off[trackId]:off[trackId + 1]]

    def frames(self, trackId):
        """Frames the track appears on, ascending (tracking_database.py:102-103)."""
        return self._frame_of_rows(self.track_rows(trackId)).tolist()

    def track(self, trackId):
        """frameId -> Link (tracking_database.py:106-111)."""
        rows = self.track_rows(trackId)
        return {int(f): self._link(int(r)) for f, r in zip(self._frame_of_rows(rows), rows)}

    def last_frame_of_track(self, trackId):
        return self.frames(trackId)[-1]

    def track_lengths(self):
        off, _ = self._track_csr()
        return np.diff(off)

    # ---- invariants (tracking_database.py:442-471, vectorised) ---------------------------------
    def check_consistency(self):
        n = self.link_num()
        off, rows = self._track_csr()
        assert off[-1] == n and len(rows) == n
        lengths = np.diff(off)
        assert (lengths >= 2).all(), "every track has at least two links"
        frames = self._frame_of_rows(rows)
        for a, b in zip(off[:-1], off[1:]):   # one link per frame, consecutive frames
            assert (np.diff(frames[a:b]) == 1).all()
        per_frame = [len(self.tracks(f)) for f in self.all_frames()]
        assert sum(per_frame) == n
        ids = self.track_id[self.track_id != NO_ID]
        assert ids.min(initial=0) >= 0 and ids.max(initial=-1) < self.n_tracks
        assert len(np.unique(ids)) == self.n_tracks
        return True

    # ---- on-disk format -------------------------------------------------------------------------
    def save(self, path):
        """One .npz (numpy's zip of .npy members, no pickle): the columns above + a format version.
        Replaces TrackingDB.serialize's pickle of the dicts (tracking_database.py:340-356)."""
        np.savez(path, format_version=np.int32(FORMAT_VERSION), frame_off=self.frame_off, x_left=self.x_left,
                 x_right=self.x_right, y=self.y, feat=self.feat, track_id=self.track_id,
                 inliers_percent=self.inliers_percent, n_tracks=np.int64(self.n_tracks))

    @classmethod
    def load(cls, path, link_factory=Link):
        with np.load(path if str(path).endswith(".npz") else str(path) + ".npz", allow_pickle=False) as z:
            if int(z["format_version"]) != FORMAT_VERSION:
                raise ValueError(f"unsupported SoATrackingDB format {int(z['format_version'])}")
            return cls(z["frame_off"], z["x_left"], z["x_right"], z["y"], z["feat"], z["track_id"],
                       z["inliers_percent"], int(z["n_tracks"]), link_factory)

    def __eq__(self, other):
        return (self.n_tracks == other.n_tracks and
                all(np.array_equal(getattr(self, k), getattr(other, k), equal_nan=k == "inliers_percent")
                    for k in ("frame_off", "x_left", "x_right", "y", "feat", "track_id", "inliers_percent")))

    # ---- the reference's object -----------------------------------------------------------------
    def to_reference_db(self, tracking_db_cls, link_cls=None):
        """The reference's TrackingDB (dicts of lists of Link objects) with exactly the content its own
        add_frame loop produces (tracking_database.py:273-337), for code that needs the original type."""
        mk = link_cls or self.link_factory
        db = tracking_db_cls()
        F = self.frame_num()
        db.last_frameId = F - 1
        db.last_trackId = self.n_tracks - 1
        off, rows = self._track_csr()
        frames = self._frame_of_rows(rows)
        db.trackId_to_frames = {t: frames[off[t]:off[t + 1]].tolist() for t in range(self.n_tracks)}
        db.frameId_to_lfeature = {f: self.features(f) for f in range(F)}
        db.frameId_to_trackIds_list = {f: self.frame_track_ids(f).tolist() for f in range(F)}
        db.frameID_to_inliers_percent = {f: float(p) for f, p in enumerate(self.inliers_percent) if not np.isnan(p)}
        xl, xr, y = self.x_left.astype(np.float64).tolist(), self.x_right.astype(np.float64).tolist(), self.y.tolist()
        link_of = {}
        db.leftover_links = {}
        for f in range(F):
            a, b = self._rows(f)
            ids = self.track_id[a:b].tolist()
            left = []
            for k, t in enumerate(ids):
                ln = mk(xl[a + k], xr[a + k], y[a + k])
                if t != NO_ID:
                    link_of[(f, t)] = ln
                elif f < F - 1:
                    left.append(ln)
            if f < F - 1:
                db.leftover_links[f] = left
            else:
                db.prev_frame_links = [link_of[(f, t)] if t != NO_ID else mk(xl[a + k], xr[a + k], y[a + k])
                                       for k, t in enumerate(ids)]
        # insertion order of linkId_to_link in the reference: frame pair by pair, ascending previous index;
        # dict equality does not depend on it, so frame-major order is used here
        db.linkId_to_link = link_of
        return db


def track_ids_host(fwd_idx, inliers, n_links):
    """Host (numpy) restatement of slamfe_track_ids for per-frame arrays: fwd_idx[f][i] = link of frame
    f + 1 that link i of frame f matches, inliers[f][i] = in_prev_cur.  Returns (list of per-frame id
    arrays, n_tracks).  Used by build() when the device table is not supplied, and by the tests."""
    F = len(n_links)
    ids = [np.full(int(n), NO_ID, dtype=np.int32) for n in n_links]
    nxt = 0
    for f in range(F - 1):
        sel = np.nonzero(np.asarray(inliers[f], dtype=bool))[0]
        if len(sel) == 0:
            continue
        j = np.asarray(fwd_idx[f], dtype=np.int64)[sel]
        new = ids[f][sel] == NO_ID
        k = int(new.sum())
        ids[f][sel[new]] = np.arange(nxt, nxt + k, dtype=np.int32)
        nxt += k
        ids[f + 1][j] = ids[f][sel]
    return ids, nxt


def build(seq, tables, link_factory=Link):
    """SoATrackingDB from a PackedSequence and the host tables of FrontEnd.run_host(seq, track=True,
    track_ids=True) — vectorised over the whole sequence, no per-match objects.  Raises where the
    reference's loop would (see database.create_db)."""
    F = seq.n_frames
    l_off = seq.l_off.astype(np.int64)
    r_off = seq.r_off.astype(np.int64)
    n_links = tables["n_links"][:F].astype(np.int64)
    n_matches = tables["n_matches"][:F].astype(np.int64)
    if (n_matches == 0).any():
        raise ZeroDivisionError("division by zero")          # database.py:26
    if "db_track_id" in tables:   # dense columns straight from the device (slamfe_pack_db): nothing left to do
        if F > 1:
            if (n_links == 0).any():
                raise IndexError("tuple index out of range")    # database.py:56 on an empty match list
            if (tables["n_good"][:F - 1] < 4).any():            # ransac.py:95
                raise ValueError("Cannot take a larger sample than population when 'replace=False'")
        with np.errstate(invalid="ignore", divide="ignore"):
            pct = 100 * (n_links / n_matches)
        return SoATrackingDB(tables["db_link_off"].astype(np.int64), tables["db_x_left"].copy(),
                             tables["db_x_right"].copy(), tables["db_y"].copy(), tables["db_feat"].copy(),
                             tables["db_track_id"].copy(), pct.astype(np.float64), int(tables["n_tracks"][0]), link_factory)
    width = np.diff(l_off)
    within = np.arange(int(l_off[-1]), dtype=np.int64) - np.repeat(l_off[:-1], width)
    frame_of = np.repeat(np.arange(F, dtype=np.int64), width)
    valid = within < np.repeat(n_links, width)
    rows = np.nonzero(valid)[0]
    f_of = frame_of[rows]
    src = tables["link_src"][rows].astype(np.int64)                       # left keypoint of the link
    left_row = l_off[f_of] + src
    right_row = r_off[f_of] + tables["match_t"][left_row].astype(np.int64)
    frame_off = np.zeros(F + 1, dtype=np.int64)
    np.cumsum(n_links, out=frame_off[1:])
    if F > 1:
        n_good = tables["n_good"][:F - 1]
        keys = tables["fwd_keys"][rows[f_of < F - 1]].view(np.uint32) if len(rows) else np.zeros(0, np.uint32)
        if (n_links[:-1] == 0).any() or (keys == _cabi.KEY_NONE).any():
            raise IndexError("tuple index out of range")    # database.py:56 on an empty match list
        if (n_good < 4).any():                                # ransac.py:95
            raise ValueError("Cannot take a larger sample than population when 'replace=False'")
    if "track_id" in tables:
        track_id = np.ascontiguousarray(tables["track_id"][rows], dtype=np.int32)
        n_tracks = int(tables["n_tracks"][0])
    else:
        per_frame_idx, per_frame_inl = [], []
        for f in range(F - 1):
            a = int(l_off[f])
            k = tables["fwd_keys"][a:a + int(n_links[f])].view(np.uint32)
            per_frame_idx.append(k & _cabi.KEY_IDX_MASK)
            per_frame_inl.append(tables["inlier_fwd"][a:a + int(n_links[f])])
        ids, n_tracks = track_ids_host(per_frame_idx, per_frame_inl, n_links)
        track_id = np.concatenate(ids) if ids else np.zeros(0, np.int32)
    pl, pr = seq.pts_l, seq.pts_r
    y = (pl[left_row, 1].astype(np.float64) + pr[right_row, 1].astype(np.float64)) / 2     # tracking_database.py:243
    with np.errstate(invalid="ignore", divide="ignore"):
        pct = 100 * (n_links / n_matches)
    return SoATrackingDB(frame_off, np.ascontiguousarray(pl[left_row, 0]), np.ascontiguousarray(pr[right_row, 0]), y,
                         np.ascontiguousarray(seq.desc_l[left_row]), track_id, pct.astype(np.float64), n_tracks,
                         link_factory)


def create_db(frames, chunk_frames=288, h_max=256, seed=1, front_end=None, link_factory=Link):
    """database.py:30-89 for a whole sequence, straight into the flat store: pack -> FrontEnd.run_host
    (track + track ids on the device) -> build().  `frames` as for database.pack_frames."""
    from . import database, frontend
    seq = frames if isinstance(frames, frontend.PackedSequence) else database.pack_frames(frames)
    fe = front_end or frontend.FrontEnd()
    tables, _, _ = fe.run_host(seq, chunk_frames=chunk_frames, track=True, h_max=h_max, seed=seed, full_ransac=True,
                               track_ids=True, pack_db=True, keys=("n_matches", "n_links"),
                               track_keys=("n_good", "n_hyp", "best", "pose", "pose_status"))
    return build(seq, tables, link_factory=link_factory)
