"""Where does the host pipeline's time go?  e2e step time by chunk size, and the same with the copy-out
disabled / the kernels disabled.  python scripts/e2e_probe.py"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import slamfe
from slamfe import frontend, synth
dev = torch.device("cuda", 0)
F = 4541
seq_t = synth.torch_sequence(F, first_frame=0, seed=1, device=dev)
pinned = {k: torch.empty(seq_t[k].shape, dtype=seq_t[k].dtype, pin_memory=True).copy_(seq_t[k]) for k in ("desc_l", "desc_r", "pts_l", "pts_r")}
seq = frontend.PackedSequence(pinned["desc_l"].numpy(), pinned["desc_r"].numpy(), pinned["pts_l"].numpy(), pinned["pts_r"].numpy(),
                              seq_t["l_off"], seq_t["r_off"], seq_t["n_l"], seq_t["n_r"], pinned)
def timeit(fn, reps=5):
    fn(); fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
fe = frontend.FrontEnd()
for c in (144, 288, 576, 1152, 2304):
    print(f"chunk {c}: {timeit(lambda: fe.run_host(seq, chunk_frames=c, track=True, h_max=128, seed=1)):.2f} ms", flush=True)
print(f"keys=n_links only (tiny D2H): {timeit(lambda: fe.run_host(seq, chunk_frames=576, track=True, h_max=128, seed=1, keys=('n_links',), track_keys=('n_good',))):.2f} ms")
# host-side issue cost of one step: same calls, but measure until the Python call returns vs until the GPU is done
t0 = time.perf_counter(); fe.run_host(seq, chunk_frames=576, track=True, h_max=128, seed=1); t1 = time.perf_counter()
print(f"run_host wall (includes final sync): {(t1 - t0) * 1e3:.2f} ms")
# pure H2D of the inputs in 8 big copies
d = {k: torch.empty_like(v, device=dev) for k, v in pinned.items()}
def h2d():
    for k, v in pinned.items(): d[k].copy_(v, non_blocking=True)
print(f"plain H2D: {timeit(h2d):.2f} ms")
# timeline of one step (CUDA events on the three kinds of streams)
for c in (288, 576):
    fe.trace = []
    fe.run_host(seq, chunk_frames=c, track=True, h_max=128, seed=1)
    torch.cuda.synchronize()
    tr, fe.trace = fe.trace, None
    t0 = tr[0][2]
    print(f"--- timeline chunk_frames={c} (ms since the first H2D started)")
    for kind in ("h2d", "compute", "d2h"):
        rows = [(ch, t0.elapsed_time(a), t0.elapsed_time(b)) for k, ch, a, b in tr if k == kind]
        busy = sum(e - s for _, s, e in rows)
        print(f"{kind:8s} busy {busy:6.2f} ms, first start {rows[0][1]:6.2f}, last end {rows[-1][2]:6.2f}: " +
              " ".join(f"[{s:.1f}-{e:.1f}]" for _, s, e in rows))
