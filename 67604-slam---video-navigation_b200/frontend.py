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


class FrontEnd:
    """Runs the four batched stages over a DeviceSequence; owns (and reuses) the output buffers."""

    def __init__(self, P=None, Q=None):
        from . import utils
        self.P = utils.P if P is None else np.asarray(P, dtype=np.float64)
        self.Q = utils.Q if Q is None else np.asarray(Q, dtype=np.float64)
        self._out = None
        self._key = None
        # kernels launched by one run(): 2 matcher launches, stereo epilogue, triangulation
        # (cudaMemsetAsync initialisation of the key tables is not counted)
        self.launches_per_run = 4

    def _buffers(self, ds: DeviceSequence):
        torch = _cabi.require_cuda()
        L, R, F = ds.desc_l.shape[0], ds.desc_r.shape[0], ds.n_frames
        key = (L, R, F, ds.desc_l.device)
        if self._key != key:
            dev = ds.desc_l.device
            i32 = dict(dtype=torch.int32, device=dev)
            self._out = {
                "lr_row_keys": torch.empty((L, 2), **i32), "lr_col_keys": torch.empty((R,), **i32),
                "match_t": torch.empty((L,), **i32), "n_matches": torch.empty((F,), **i32),
                "n_links": torch.empty((F,), **i32), "link_src": torch.empty((L,), **i32),
                "links": torch.empty((L, 3), dtype=torch.float32, device=dev),
                "feat": torch.empty((L, 64), dtype=torch.uint8, device=dev),
                "xyz": torch.empty((L, 3), dtype=torch.float32, device=dev),
                "fwd_keys": torch.empty((L, 2), **i32), "bwd_keys": torch.empty((L,), **i32),
            }
            self._key = key
        return self._out

    def run(self, ds: DeviceSequence):
        """All stages, asynchronous on the current stream.  Returns the dict of output tensors."""
        o = self._buffers(ds)
        F = ds.n_frames
        if F == 0:
            return o
        # 1. stereo match, every frame
        ops.hamming_top2_batched(ds.desc_l, ds.l_off, ds.desc_r, ds.r_off, F, ds.max_nl, ds.max_nr, DESC_BYTES,
                                 q_cnt=ds.n_l, t_cnt=ds.n_r, want_cols=True,
                                 row_keys=o["lr_row_keys"], col_keys=o["lr_col_keys"])
        # 2. crossCheck + row filter + links + features[is_valid]
        ops.stereo_links_batched(o["lr_row_keys"], o["lr_col_keys"], ds.l_off, ds.r_off, F, ds.pts_l, ds.pts_r,
                                 desc_left=ds.desc_l, desc_bytes=DESC_BYTES, out=o, n_l=ds.n_l, n_r=ds.n_r)
        # 3. triangulate every link
        ops.triangulate_links(o["links"], self.P, self.Q, out=o["xyz"])
        # 4. consecutive frames on the filtered features: problem f = (frame f, frame f+1)
        if F > 1:
            max_links = min(ds.max_nl, ds.max_nr)
            ops.hamming_top2_batched(o["feat"], ds.l_off, o["feat"], ds.l_off[1:], F - 1, max_links, max_links,
                                     DESC_BYTES, q_cnt=o["n_links"], t_cnt=o["n_links"][1:], want_cols=True,
                                     row_keys=o["fwd_keys"], col_keys=o["bwd_keys"])
        return o



def descriptor_pairs(n_l, n_r, n_links=None) -> int:
    """Algorithmic descriptor pairs of one pass: sum Nl*Nr (stereo; one pass yields both
    directions) + sum links_f * links_{f+1} (consecutive frames on the filtered features)."""
    total = int(np.sum(np.asarray(n_l, dtype=np.int64) * np.asarray(n_r, dtype=np.int64)))
    if n_links is not None and len(n_links) > 1:
        k = np.asarray(n_links, dtype=np.int64)
        total += int(np.sum(k[:-1] * k[1:]))
    return total


def results_to_host(o, keys=("match_t", "n_matches", "n_links", "link_src", "links", "xyz", "fwd_keys", "bwd_keys"),
                    pinned=None):
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
