"""ncu target: a few launches of the tcgen05 matcher (dense 20k with / without column minima, and a
sequence-shaped ragged stereo launch).  python scripts/prof_mma.py [int|mma]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import slamfe
from slamfe import ops, synth, frontend
kind = sys.argv[1] if len(sys.argv) > 1 else "mma"
ops.set_matcher_kernel(kind)
rng = np.random.default_rng(0)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
n = 20000
q = dev(synth.descriptors(rng, n)); t = dev(synth.descriptors(rng, n))
F = 64
nl = rng.integers(2000, 5001, F); nr = rng.integers(2000, 5001, F)
lo, ro = frontend.plan_offsets(nl), frontend.plan_offsets(nr)
DL = torch.randint(0, 256, (int(lo[-1]), 61), dtype=torch.uint8, device="cuda")
DR = torch.randint(0, 256, (int(ro[-1]), 61), dtype=torch.uint8, device="cuda")
lod, rod, nld, nrd = dev(lo), dev(ro), dev(nl.astype(np.int32)), dev(nr.astype(np.int32))
for _ in range(2):
    ops.hamming_top2(q, t, want_cols=False, best_only=True)
    ops.hamming_top2(q, t, want_cols=True, best_only=True)
    ops.hamming_top2_batched(DL, lod, DR, rod, F, 5000, 5000, 61, q_cnt=nld, t_cnt=nrd, want_cols=True, best_only=True)
torch.cuda.synchronize()
print("done")
