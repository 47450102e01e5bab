"""What the column minima cost on the sequence-shaped ragged stereo launch (shipped matcher): the same launch with and
without them, best-only.  python scripts/stereo_cols_cost.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import slamfe
from slamfe import ops, frontend
ops.set_matcher_kernel("mma")
rng = np.random.default_rng(0)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for F in (256, 1024):
    nl = rng.integers(2000, 5001, F); nr = rng.integers(2000, 5001, F)
    lo, ro = frontend.plan_offsets(nl), frontend.plan_offsets(nr)
    DL = torch.randint(0, 256, (int(lo[-1]), 61), dtype=torch.uint8, device="cuda")
    DR = torch.randint(0, 256, (int(ro[-1]), 61), dtype=torch.uint8, device="cuda")
    lod, rod, nld, nrd = dev(lo), dev(ro), dev(nl.astype(np.int32)), dev(nr.astype(np.int32))
    pairs = float(np.sum(nl.astype(np.int64) * nr))
    padded = float(np.sum((np.ceil(nl / 128) * 128) * (np.ceil(nr / 128) * 128)))
    floor_ms = padded / (16 * 148 * 1.965e9) * 1e3
    for cols in (False, True):
        for compact in (False, True):
            ms = timeit(lambda: ops.hamming_top2_batched(DL, lod, DR, rod, F, 5000, 5000, 61, q_cnt=nld, t_cnt=nrd, want_cols=cols, best_only=True, compact=compact))
            print(f"stereo {F} frames cols={cols} compact={compact}: {ms:.3f} ms, {pairs/ms/1e6:.1f} G pairs/s, padded MMA floor {floor_ms:.3f} ms ({floor_ms/ms:.2f})", flush=True)
