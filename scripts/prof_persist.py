"""Development only: clock64 breakdown of the persistent tcgen05 matcher (library built with
SLAMFE_NVCC_DEFINES="-DSLAMFE_MMA_DEV -DSLAMFE_MMA_PROF").  Prints per-CTA averages for a stereo-shaped ragged launch."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import slamfe
from slamfe import ops, frontend, _cabi

lib = _cabi.load_library()
prof = lib.slamfe_dev_mma_prof
prof.argtypes = [ctypes.c_void_p, ctypes.c_int]
rng = np.random.default_rng(0)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
ops.set_matcher_kernel("mma")

def report(name, fn, pairs):
    fn(); torch.cuda.synchronize()
    prof(None, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    buf = np.zeros((256, 16), np.uint64)
    prof(buf.ctypes.data, 0)
    b = buf[:148].astype(np.float64)
    names = ["mma_total", "read_job", "a_ready0", "b_full", "d_empty", "mid_wait(all)", "jobs", "mma_jobs", "a_ready1",
             "g0_dfull_wait", "g0_build", "g0_last_ld+build", "g1_dfull_wait", "g1_build", "g1_last_ld+build", "g0_peek"]
    print(f"--- {name}: {ms:.3f} ms = {ms*1e-3*1.965e9:.0f} cycles, {pairs/ms/1e6:.1f} G pairs/s")
    for i, nme in enumerate(names):
        print(f"   {nme:18s} mean {b[:, i].mean():12.0f}  min {b[:, i].min():12.0f} max {b[:, i].max():12.0f}")
    jobs = b[:, 6].mean(); mm = b[:, 7].mean()
    print(f"   per job: boundary read {b[:,1].mean()/jobs:.0f} a_ready0 {b[:,2].mean()/jobs:.0f} a_ready1 {b[:,8].mean()/jobs:.0f}; "
          f"per MMA job: total {b[:,0].mean()/mm:.0f} (floor 1024) b_full {b[:,3].mean()/mm:.0f} d_empty {b[:,4].mean()/mm:.0f}")

F = 256
nl = rng.integers(2000, 5001, F); nr = rng.integers(2000, 5001, F)
lo, ro = frontend.plan_offsets(nl), frontend.plan_offsets(nr)
DL = torch.randint(0, 256, (int(lo[-1]), 61), dtype=torch.uint8, device="cuda")
DR = torch.randint(0, 256, (int(ro[-1]), 61), dtype=torch.uint8, device="cuda")
lod, rod, nld, nrd = dev(lo), dev(ro), dev(nl.astype(np.int32)), dev(nr.astype(np.int32))
pairs = float(np.sum(nl.astype(np.int64) * nr))
report("stereo 256 frames, cols, best-only", lambda: ops.hamming_top2_batched(DL, lod, DR, rod, F, 5000, 5000, 61, q_cnt=nld, t_cnt=nrd, want_cols=True, best_only=True), pairs)
report("stereo 256 frames, rows only, best-only", lambda: ops.hamming_top2_batched(DL, lod, DR, rod, F, 5000, 5000, 61, q_cnt=nld, t_cnt=nrd, want_cols=False, best_only=True), pairs)
n = 20000
q = torch.randint(0, 256, (n, 61), dtype=torch.uint8, device="cuda"); t = torch.randint(0, 256, (n, 61), dtype=torch.uint8, device="cuda")
report("dense 20k, rows only", lambda: ops.hamming_top2(q, t, want_cols=False, best_only=True), float(n) * n)
