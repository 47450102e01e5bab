"""What does an empty CTA of the tcgen05 matcher cost?  A ragged batch whose problems are all empty launches
grid (max_nq / 256, 1, n_problems) CTAs that load their problem's counts and exit.  python scripts/noop_probe.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import slamfe
from slamfe import ops, frontend
F = 4541
n = np.full(F, 5000)
off = frontend.plan_offsets(n)
D = torch.randint(0, 256, (int(off[-1]), 61), dtype=torch.uint8, device="cuda")
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
offd = dev(off)
for cnt, label in ((0, "all problems empty"), (256, "one real CTA per problem (256 x 256)"), (2000, "2000 x 2000"), (5000, "5000 x 5000")):
    c = dev(np.full(F, cnt, np.int32))
    for _ in range(3):
        ops.hamming_top2_batched(D, offd, D, offd, F, 5000, 5000, 61, q_cnt=c, t_cnt=c, want_cols=True, best_only=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        ops.hamming_top2_batched(D, offd, D, offd, F, 5000, 5000, 61, q_cnt=c, t_cnt=c, want_cols=True, best_only=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    ctas = 20 * F
    real = ((cnt + 255) // 256) * F
    print(f"{label}: {ms:.3f} ms for {ctas} CTAs ({real} with work): {ms * 1e3 / (ctas / 148):.2f} us per CTA slot per SM")

# the pair stage's shape: 64-byte feature rows, 61 descriptor bytes, link counts of the bench sequence
rng = np.random.default_rng(0)
k = rng.integers(1000, 2551, F)
offk = frontend.plan_offsets(np.full(F, 5000))
Fe = torch.randint(0, 256, (int(offk[-1]), 64), dtype=torch.uint8, device="cuda")
Fe[:, 61:] = 0
offkd = dev(offk)
qc, tc = dev(k[:-1].astype(np.int32)), dev(k[1:].astype(np.int32))
qo, to = offkd[:-1].contiguous(), offkd[1:].contiguous()
pairs = float((k[:-1].astype(np.int64) * k[1:]).sum())
for max_n, label in ((5000, "grid sized for 5000 rows (as the pipeline launches it)"), (2560, "grid sized for 2560 rows")):
    run = lambda: ops.hamming_top2_batched(Fe, qo, Fe, to, F - 1, max_n, max_n, 61, q_cnt=qc, t_cnt=tc, want_cols=True,
                                           best_only=True, compact=True)
    for _ in range(3):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"pair-stage shape, {label}: {ms:.3f} ms, {pairs / ms / 1e9:.3f} T pairs/s = {pairs / ms / 1e9 / 4.653:.2f} of the MMA issue floor")
