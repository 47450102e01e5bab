"""Development tool: time the batched matcher launch for every (full adders per pair, rows per
thread) variant and check that all variants return identical key tables."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import slamfe  # noqa: E402,F401
from slamfe import ops, synth  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 512
rs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [2, 3, 4]
css = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [7, 8, 9, 10]
dev = torch.device("cuda", 0)
seq = synth.torch_sequence(frames, first_frame=0, seed=1, device=dev)
l_off = torch.from_numpy(seq["l_off"]).to(dev)
r_off = torch.from_numpy(seq["r_off"]).to(dev)
n_l = torch.from_numpy(seq["n_l"]).to(dev)
n_r = torch.from_numpy(seq["n_r"]).to(dev)
pairs = float(np.sum(seq["n_l"].astype(np.int64) * seq["n_r"].astype(np.int64)))
ref = None
ref2 = None
modes = os.environ.get("TUNE_MODES", "11,10,01,00").split(",")
for want_cols, best_only in [(m[0] == "1", m[1] == "1") for m in modes]:
    for r in rs:
        for cs in css:
            os.environ["SLAMFE_HAMMING_CS"] = str(cs)
            os.environ["SLAMFE_HAMMING_R"] = str(r)

            def run():
                return ops.hamming_top2_batched(seq["desc_l"], l_off, seq["desc_r"], r_off, frames,
                                                int(seq["n_l"].max()), int(seq["n_r"].max()), 61,
                                                q_cnt=n_l, t_cnt=n_r, want_cols=want_cols, best_only=best_only)

            rk, ck = run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                rk, ck = run()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            if ref is None:
                ref = (rk.clone(), ck.clone())
            same = bool(torch.equal(rk[:, 0], ref[0][:, 0])) and (ck is None or bool(torch.equal(ck, ref[1])))
            if not best_only:
                if ref2 is None:
                    ref2 = rk.clone()
                same = same and bool(torch.equal(rk, ref2))
            print(f"cols={int(want_cols)} best_only={int(best_only)} R={r} CS={cs}: {ms:8.3f} ms  {pairs / ms / 1e6:8.1f} Gpairs/s  "
                  f"({16 * pairs / ms / 1e6:7.0f} Gpopc-equiv/s)  identical={same}", flush=True)
