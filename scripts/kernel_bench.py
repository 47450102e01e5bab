"""Per-kernel device timings with their rooflines (development + profiles/ evidence).

    python scripts/kernel_bench.py [--out gpurun_out/kernels.json]

Every kernel is timed alone with CUDA events after warm-up on inputs larger than L2 where the
kernel is memory-bound.  Algorithmic work per unit follows SURVEY.md section 8(d) / DESIGN.md.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import slamfe  # noqa: E402,F401
from slamfe import dist as sdist, ops, synth, utils  # noqa: E402


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "kernels.json"))
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = peaks.get("hbm_gbs", 6650.0)
    res = {"hbm_peak_gbs": hbm, "hbm_peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}
    res["popc_peak_gpopc"] = ops.measure_peak(0) / 1e9
    res["fp64_fma_peak_gfma"] = ops.measure_peak(2) / 1e9
    g = torch.Generator(device=dev).manual_seed(7)

    # ---- triangulation (rectified links): HBM-bound, 24 B/match fp32, 48 B/match fp64 ----
    n = 64 * 1024 * 1024
    xl = torch.rand(n, device=dev, generator=g) * 1200 + 20
    links = torch.stack([xl, xl - (torch.rand(n, device=dev, generator=g) * 117 + 2.5),
                         torch.rand(n, device=dev, generator=g) * 365 + 5], dim=1).contiguous()
    out32 = torch.empty_like(links)
    ms = timed(lambda: ops.triangulate_links(links, utils.P, utils.Q, out=out32))
    res["triangulate_links_f32"] = {"n": n, "ms": ms, "bytes_per_match": 24, "gbs": 24 * n / ms / 1e6,
                                    "frac_hbm": 24 * n / ms / 1e6 / hbm, "gmatches_per_s": n / ms / 1e6}
    n64 = n // 2
    links64 = links[:n64].double().contiguous()
    out64 = torch.empty_like(links64)
    ms = timed(lambda: ops.triangulate_links(links64, utils.P, utils.Q, out=out64))
    res["triangulate_links_f64"] = {"n": n64, "ms": ms, "bytes_per_match": 48, "gbs": 48 * n64 / ms / 1e6,
                                    "frac_hbm": 48 * n64 / ms / 1e6 / hbm, "gmatches_per_s": n64 / ms / 1e6}
    nd = 8 * 1024 * 1024
    pxy = torch.stack([links64[:nd, 0], links64[:nd, 2]], dim=1).contiguous()
    qxy = torch.stack([links64[:nd, 1], links64[:nd, 2] + 0.3], dim=1).contiguous()
    ms = timed(lambda: ops.triangulate_dlt(pxy, qxy, utils.P, utils.Q))
    res["triangulate_dlt_f64"] = {"n": nd, "ms": ms, "bytes_per_match": 56, "gbs": 56 * nd / ms / 1e6,
                                  "gmatches_per_s": nd / ms / 1e6, "bound": "fp64 (Jacobi sweeps), not HBM"}
    del links, links64, out32, out64, pxy, qxy, xl

    # ---- RANSAC scoring: config 3, 4096 hypotheses x 5000 correspondences per frame ----
    rng = np.random.default_rng(2)
    H, N, F = 4096, 5000, 64
    Ts, pts, lp, rp = synth.pnp_problem(rng, N, H)
    T = torch.from_numpy(np.ascontiguousarray(Ts)).to(dev).repeat(F, 1, 1).contiguous()
    ptsd = torch.from_numpy(pts).to(dev).repeat(F, 1).contiguous()
    lpd = torch.from_numpy(lp).to(dev).repeat(F, 1).contiguous()
    rpd = torch.from_numpy(rp).to(dev).repeat(F, 1).contiguous()
    pt_off = torch.arange(0, (F + 1) * N, N, dtype=torch.int32, device=dev)
    ms = timed(lambda: ops.ransac_score(T, ptsd, lpd, rpd, utils.K, utils.M1, utils.M2, pt_off=pt_off, n_frames=F,
                                        max_points=N), reps=3)
    fma = 24.0 * H * N * F
    res["ransac_score"] = {"H": H, "N": N, "frames": F, "ms": ms, "hyp_point_pairs_per_s": H * N * F / ms * 1e3,
                           "frames_per_s": F / ms * 1e3, "gfma_per_s": fma / ms / 1e6,
                           "frac_fp64_fma_peak": fma / ms / 1e6 / res["fp64_fma_peak_gfma"],
                           "note": "24 fp64 FMA + 4 IEEE fp64 divisions + 8 compares per (hypothesis, point); the "
                                   "divisions are not counted in gfma_per_s"}
    del T, ptsd, lpd, rpd

    # ---- dense sweep, config 5: 20k x 20k top-2 ----
    d = synth.descriptors(rng, 20000)
    q = torch.from_numpy(d).to(dev)
    t = torch.from_numpy(synth.flip_bits(rng, d[rng.permutation(20000)], 0.08)).to(dev)
    for name, kw in (("dense_20k_top2", {}), ("dense_20k_best_only", {"best_only": True}),
                     ("dense_20k_best_cols", {"best_only": True, "want_cols": True})):
        ms = timed(lambda: ops.hamming_top2(q, t, **kw), reps=10)
        res[name] = {"nq": 20000, "nt": 20000, "ms": ms, "gpairs_per_s": 4e8 / ms / 1e6,
                     "gpopc_equiv_per_s": 16 * 4e8 / ms / 1e6,
                     "frac_popc_peak": 16 * 4e8 / ms / 1e6 / res["popc_peak_gpopc"]}

    # ---- loop closure, config 4: 450 keyframes x 2000 descriptors, every keyframe vs all prior ----
    K, n_kf = 450, 2000
    kf = torch.randint(0, 256, (K * n_kf, 61), dtype=torch.uint8, device=dev, generator=g)
    kf[:, 60] &= 0x3F
    pairs = sdist.candidate_pairs(K)
    q_off = torch.from_numpy((pairs[:, 0].astype(np.int64) * n_kf).astype(np.int32)).to(dev)
    t_off = torch.from_numpy((pairs[:, 1].astype(np.int64) * n_kf).astype(np.int32)).to(dev)
    cnt = torch.full((len(pairs),), n_kf, dtype=torch.int32, device=dev)
    blk = 16384  # candidate pairs per launch: row_keys of one launch = blk * 2000 * 8 B = 262 MB
    out_off = torch.arange(0, blk, dtype=torch.int32, device=dev) * n_kf
    row_keys = torch.empty((blk * n_kf, 2), dtype=torch.int32, device=dev)

    def loop_pass():
        for p0 in range(0, len(pairs), blk):
            p1 = min(len(pairs), p0 + blk)
            ops.hamming_pairs(kf, q_off[p0:p1], cnt[p0:p1], kf, t_off[p0:p1], cnt[p0:p1], out_off, p1 - p0,
                              n_kf, n_kf, 61, row_keys=row_keys, out_rows_total=(p1 - p0) * n_kf, best_only=True)
    ms = timed(loop_pass, reps=2, warm=1)
    npairs = float(len(pairs)) * n_kf * n_kf
    res["loop_closure_all_prior"] = {"keyframes": K, "descriptors_per_keyframe": n_kf, "candidate_pairs": len(pairs),
                                     "ms": ms, "candidate_pairs_per_s": len(pairs) / ms * 1e3,
                                     "gpairs_per_s": npairs / ms / 1e6,
                                     "frac_popc_peak": 16 * npairs / ms / 1e6 / res["popc_peak_gpopc"]}
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
