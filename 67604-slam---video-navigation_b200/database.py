"""Batched drop-in for the reference's tracking-DB builder loop (backend/database/database.py:30-89).

The reference walks the sequence frame by frame: `first_operation` (detect + crossCheck L<->R match +
row filter + `db.create_links`, :12-27), forward and backward `MATCHER.match` against the previous
frame (:54-55), the mutual check (:67-77), `ransac_pnp_for_tracking_db` (:80) and `db.add_frame` (:87).
Everything between feature extraction and `add_frame` is independent per frame / frame pair, so
`create_db` here runs it for the whole sequence through `FrontEnd.run_host(track=True)` and then
replays `add_frame` with exactly the arguments the reference would have passed:

    links                      Link(x_left, x_right, (yl + yr) / 2)   (tracking_database.py:243)
    left_features              desc_left[is_valid]                    (tracking_database.py:235)
    matches_to_previous_left   one cv2.DMatch per previous-frame feature (database.py:54)
    inliers                    in_prev_cur                            (database.py:84-85)

Feature extraction (AKAZE, `matching.py:42-43`) and the TrackingDB itself are the reference's /
OpenCV's (out of scope): the caller passes keypoints + descriptors per frame
and a `db` object with the reference's `add_frame` / `frameID_to_inliers_percent` interface.
"""
from __future__ import annotations

import numpy as np

from . import frontend
from .matching import _keypoint_array

try:
    import cv2
except Exception:  # pragma: no cover
    cv2 = None


class Link:
    """Stand-in for tracking_database.Link (:12-29) when the reference class is not supplied."""
    __slots__ = ("x_left", "x_right", "y")

    def __init__(self, x_left, x_right, y):
        self.x_left, self.x_right, self.y = x_left, x_right, y

    def __repr__(self):
        return f"Link (xl={self.x_left}, xr={self.x_right}, y={self.y})"


def pack_frames(frames, pin=True):
    """frames: iterable of (kp_left, kp_right, desc_left, desc_right) as
    `extract_kps_descs_matches` returns them (matching.py:38-45; keypoints as cv2.KeyPoint sequences or
    (n, 2) float arrays) -> frontend.PackedSequence."""
    def desc(d):
        d = np.asarray(d)
        if d.dtype != np.uint8 or d.ndim != 2 or d.shape[1] != frontend.DESC_BYTES:
            # e.g. SIFT's float32 (n, 128): casting would silently produce garbage (cv2 raises cv2.error here)
            raise TypeError(f"descriptors must be (n, {frontend.DESC_BYTES}) uint8 (AKAZE MLDB), got {d.dtype} {d.shape}")
        return np.ascontiguousarray(d)

    return frontend.pack_sequence([(desc(dl), desc(dr), _keypoint_array(kl), _keypoint_array(kr))
                                   for kl, kr, dl, dr in frames], pin=pin)


def frames_from_tables(seq, tables):
    """Per-frame `add_frame` arguments from the host tables of FrontEnd.run_host(seq, track=True).
    Yields dicts: frame, links (k, 3) float64 [x_left, x_right, y], features (k, 61) uint8,
    inliers_percent, and for frames > 0: fwd_idx / fwd_dist (one entry per PREVIOUS-frame feature) and
    inliers (bool per forward match)."""
    F = seq.n_frames
    prev_k = 0
    for f in range(F):
        lo, ro = int(seq.l_off[f]), int(seq.r_off[f])
        k = int(tables["n_links"][f])
        src = tables["link_src"][lo:lo + k].astype(np.int64)
        dst = tables["match_t"][lo:lo + int(seq.n_l[f])][src].astype(np.int64)
        pl = seq.pts_l[lo:lo + int(seq.n_l[f])].astype(np.float64)
        pr = seq.pts_r[ro:ro + int(seq.n_r[f])].astype(np.float64)
        links = np.stack([pl[src, 0], pr[dst, 0], (pl[src, 1] + pr[dst, 1]) / 2], axis=1) if k else np.zeros((0, 3))
        n_matches = int(tables["n_matches"][f])
        out = {"frame": f, "links": links, "features": seq.desc_l[lo:lo + int(seq.n_l[f])][src],
               "inliers_percent": 100 * (k / n_matches) if n_matches else None}
        if f > 0:
            plo = int(seq.l_off[f - 1])
            keys = tables["fwd_keys"][plo:plo + prev_k].view(np.uint32)
            out["fwd_idx"] = (keys & frontend._cabi.KEY_IDX_MASK).astype(np.int64)
            out["fwd_dist"] = (keys >> frontend._cabi.KEY_IDX_BITS).astype(np.float64)
            out["fwd_valid"] = keys != frontend._cabi.KEY_NONE
            out["inliers"] = tables["inlier_fwd"][plo:plo + prev_k].astype(bool)
        prev_k = k
        yield out


def create_db(frames, db, link_factory=None, chunk_frames=288, h_max=256, seed=1, front_end=None):
    """database.py:30-89 for a whole sequence.  `frames` as for pack_frames (or a PackedSequence),
    `db` a TrackingDB-like object, `link_factory` the reference's Link class (default: the stand-in
    above).  Returns db.  Raises where the reference does: IndexError for a frame pair without any
    forward match (`matches_l_l[0]`, database.py:56), ValueError for a pair with fewer than 4 mutual
    matches (`np.random.choice(n < 4, 4, replace=False)`, ransac.py:95).  h_max only sizes the batched
    RANSAC launch: pairs whose calc_ransac_iteration exceeds it are re-run at their full count
    (FrontEnd.rescore_truncated), so every pair gets the reference's number of hypotheses."""
    mk_link = link_factory or Link
    seq = frames if isinstance(frames, frontend.PackedSequence) else pack_frames(frames)
    fe = front_end or frontend.FrontEnd()
    tables, _, _ = fe.run_host(seq, chunk_frames=chunk_frames, track=True, h_max=h_max, seed=seed, full_ransac=True)
    for fr in frames_from_tables(seq, tables):
        links = [mk_link(float(a), float(b), float(c)) for a, b, c in fr["links"]]
        if fr["inliers_percent"] is None:
            raise ZeroDivisionError("division by zero")  # database.py:26: no stereo matches in this frame
        if hasattr(db, "frameID_to_inliers_percent"):
            db.frameID_to_inliers_percent[fr["frame"]] = fr["inliers_percent"]
        if fr["frame"] == 0:
            db.add_frame(links=links, left_features=fr["features"], matches_to_previous_left=None, inliers=None)
            continue
        if len(fr["fwd_idx"]) == 0 or not fr["fwd_valid"].all():
            raise IndexError("tuple index out of range")  # database.py:56 on an empty match list
        if int(tables["n_good"][fr["frame"] - 1]) < 4:   # ransac.py:95: np.random.choice(n < 4, 4, replace=False)
            raise ValueError("Cannot take a larger sample than population when 'replace=False'")
        n = len(fr["fwd_idx"])
        matches = np.empty(n, dtype=object)  # database.py:60: np.array(matches_l_l)
        matches[:] = list(map(cv2.DMatch, range(n), fr["fwd_idx"].tolist(), [0] * n, fr["fwd_dist"].tolist()))
        db.add_frame(links, fr["features"], matches, fr["inliers"])
    return db


def make_reference_create_db(reference_database_module, inputs_module=None, matching_module=None, **pipeline_kw):
    """A function with the signature of the reference's `create_db(start_frame=0, num_frames=200, db=None)`
    (database.py:30) that reads and describes the frames exactly as `first_operation` does
    (`Inputs.read_images`, `FEATURE.detectAndCompute`, database.py:17-18 / matching.py:42-43 — CPU work that
    stays with OpenCV) and then runs everything else of the loop as ONE batched GPU pipeline
    (`create_db` above).  `patch.patch(batched_db=True)` binds it to `database.create_db`, so
    `database.run(path)` / `project.run_project` build the TrackingDB through it.

    A resumed build (start_frame != 0, database.py:46-47) needs the previous frame from `db` and is handed
    to the reference's own loop (which, once patched, still uses the GPU matchers frame by frame)."""
    ref_create_db = reference_database_module.create_db
    inputs = inputs_module or reference_database_module.Inputs
    if matching_module is None:
        import sys
        matching_module = sys.modules.get("final_project.algorithms.matching")
    tracking_db_cls = reference_database_module.TrackingDB
    link_cls = None
    try:
        import sys
        link_cls = sys.modules[tracking_db_cls.__module__].Link
    except Exception:
        link_cls = None

    def create_db_batched(start_frame=0, num_frames=200, db=None):
        if start_frame != 0:
            return ref_create_db(start_frame=start_frame, num_frames=num_frames, db=db)
        if db is None:
            db = tracking_db_cls()
        feature = matching_module.FEATURE
        frames = []
        for idx in range(num_frames):
            img_l, img_r = inputs.read_images(idx)
            kp_l, desc_l = feature.detectAndCompute(img_l, None)
            kp_r, desc_r = feature.detectAndCompute(img_r, None)
            frames.append((kp_l, kp_r, desc_l, desc_r))
        return create_db(frames, db, link_factory=link_cls, **pipeline_kw)

    create_db_batched.__wrapped__ = ref_create_db
    return create_db_batched
