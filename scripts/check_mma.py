"""Development check: tcgen05 matcher (SLAMFE_MATCH_MMA) vs the INT-pipe kernel — identical keys on a
shape matrix, then timings.  Run on a B200: python scripts/check_mma.py"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import slamfe
from slamfe import ops, synth

def run(kind, fn):
    ops.set_matcher_kernel(kind)
    out = fn()
    torch.cuda.synchronize()
    ops.set_matcher_kernel("int")
    return out

def compare(name, fn):
    a = run("int", fn)
    b = run("mma", fn)
    ok = True
    for x, y in zip(a, b):
        if x is None:
            continue
        same = bool(torch.equal(x, y))
        if not same:
            xa, ya = x.cpu().numpy().view(np.uint32), y.cpu().numpy().view(np.uint32)
            bad = np.argwhere(xa != ya)
            print("   first diffs", bad[:5].tolist(), [hex(int(xa[tuple(i)])) for i in bad[:5]], [hex(int(ya[tuple(i)])) for i in bad[:5]], "of", len(bad))
        ok &= same
    print(("OK  " if ok else "FAIL"), name, flush=True)
    return ok

rng = np.random.default_rng(0)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
allok = True
for nq, nt in [(128, 192), (1, 1), (100, 50), (130, 400), (256, 64), (257, 65), (129, 63), (300, 128), (513, 2049), (3000, 3000), (4999, 1), (1, 5000)]:
    q = synth.descriptors(rng, nq)
    t, _ = synth.paired_descriptors(rng, q, n_out=nt, dup_frac=0.05)
    qd, td = dev(q), dev(t)
    for best_only in (False, True):
        allok &= compare(f"single {nq}x{nt} best_only={best_only} cols", lambda: ops.hamming_top2(qd, td, want_cols=True, best_only=best_only))
# padded 64-byte rows, misaligned base, 32-byte descriptors
q, t = synth.descriptors(rng, 700), synth.descriptors(rng, 900)
qp = rng.integers(0, 256, (700, 64), dtype=np.uint8); qp[:, :61] = q
tp = rng.integers(0, 256, (900, 64), dtype=np.uint8); tp[:, :61] = t
qpd, tpd = dev(qp), dev(tp)
allok &= compare("padded 64B rows", lambda: ops.hamming_top2(qpd, tpd, desc_bytes=61, want_cols=True))
bq = dev(np.concatenate([np.zeros((1, 61), np.uint8), q])); bt = dev(np.concatenate([np.zeros((3, 61), np.uint8), t]))
allok &= compare("misaligned base", lambda: ops.hamming_top2(bq[1:], bt[3:], want_cols=True))
q32, t32 = dev(q[:, :32]), dev(t[:, :32])
allok &= compare("32-byte descriptors", lambda: ops.hamming_top2(q32, t32, want_cols=True))
q64 = dev(rng.integers(0, 256, (300, 64), dtype=np.uint8)); t64 = dev(rng.integers(0, 256, (500, 64), dtype=np.uint8))
t64[7] = 255; q64[3] = 0; q64[4] = 255; t64[9] = 0
allok &= compare("64-byte descriptors incl. d=512", lambda: ops.hamming_top2(q64, t64, want_cols=True))
# ragged batch
from slamfe import frontend
sizes = [(300, 280), (0, 50), (17, 0), (1, 1), (515, 700)] + [(1100, 900)] * 40
q_off = frontend.plan_offsets([a for a, _ in sizes]); t_off = frontend.plan_offsets([b for _, b in sizes])
Q = rng.integers(0, 256, (q_off[-1], 61), dtype=np.uint8); T = rng.integers(0, 256, (t_off[-1], 61), dtype=np.uint8)
Qd, Td = dev(Q), dev(T)
qo, to = dev(q_off), dev(t_off)
qc, tc = dev(np.array([a for a, _ in sizes], np.int32)), dev(np.array([b for _, b in sizes], np.int32))
for bo, comp in ((False, False), (True, False), (True, True)):
    allok &= compare(f"ragged batch best_only={bo} compact={comp}", lambda: ops.hamming_top2_batched(
        Qd, qo, Td, to, len(sizes), 1100, 900, 61, q_cnt=qc, t_cnt=tc, want_cols=True, best_only=bo, compact=comp))
print("ALL OK" if allok else "SOME FAILED", flush=True)

# timings
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

n = 20000
q = dev(synth.descriptors(rng, n)); t = dev(synth.descriptors(rng, n))
for kind in ("int", "mma"):
    for cols, bo in ((False, True), (True, True), (False, False), (True, False)):
        ops.set_matcher_kernel(kind)
        ms = timeit(lambda: ops.hamming_top2(q, t, want_cols=cols, best_only=bo))
        print(f"dense 20k {kind} cols={cols} best_only={bo}: {ms:.3f} ms, {n*n/ms/1e6:.1f} G pairs/s", flush=True)
# sequence-shaped stereo launch: 256 frames x ~3500
F = 256
nl = rng.integers(2000, 5001, F); nr = rng.integers(2000, 5001, F)
lo, ro = frontend.plan_offsets(nl), frontend.plan_offsets(nr)
DL = torch.randint(0, 256, (int(lo[-1]), 61), dtype=torch.uint8, device="cuda")
DR = torch.randint(0, 256, (int(ro[-1]), 61), dtype=torch.uint8, device="cuda")
lod, rod, nld, nrd = dev(lo), dev(ro), dev(nl.astype(np.int32)), dev(nr.astype(np.int32))
pairs = float(np.sum(nl.astype(np.int64) * nr))
for kind in ("int", "mma"):
    ops.set_matcher_kernel(kind)
    ms = timeit(lambda: ops.hamming_top2_batched(DL, lod, DR, rod, F, 5000, 5000, 61, q_cnt=nld, t_cnt=nrd, want_cols=True, best_only=True))
    print(f"stereo {F} frames {kind}: {ms:.3f} ms, {pairs/ms/1e6:.1f} G pairs/s", flush=True)
ops.set_matcher_kernel("int")
