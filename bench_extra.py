"""Additional bench workloads (BASELINE.json configs[2..4]); `bench.py --workload ransac|loop|dense`.

Same JSON-line contract as the default (sequence) workload.  These are the parity-test configurations
run at full size so that their 1/2/4/8-GPU scaling can be recorded (profiles/); the driver's headline
line is the default workload.

  ransac  configs[2]: 4096 hypotheses x 5000 correspondences per frame, frames sharded across GPUs
          (weak scaling), all-gather of the per-frame [best hypothesis, inlier count] tables
  loop    configs[3]: 450 keyframes x 2000 descriptors, every keyframe against ALL prior keyframes
          (101 025 candidate pairs), candidate blocks balanced by Nq*Nt across GPUs (strong
          scaling), all-gather of the per-pair best-match tables
  dense   configs[4]: 20 000 x 20 000 all-pairs top-2 per frame, train set sliced across GPUs (strong
          scaling), all-gather of the per-query top-2 keys + exact min-merge
"""
from __future__ import annotations

import json
import os
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))


def _peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


class Workload:
    metric = unit = ""
    scaling = "strong"
    dtype = "u8"
    launches_per_step = 1

    def step(self):  # inputs resident
        raise NotImplementedError

    def e2e_step(self):  # host buffers in, host tables out; returns (h2d, d2h) bytes
        raise NotImplementedError


def matcher_roofline(kernel_role, pairs_per_launch, launch_ms, peak_popc_g, sm_mhz=None, sms=148, desc_bytes=61):
    """Roofline record of one matcher launch for whichever kernel is selected (ops.matcher_kernel()).

    tcgen05 kernel (default): tensor bound.  Algorithmic work = one int8 MAC per descriptor bit and pair
    (8*desc_bytes MAC = 2 ops each); peak = int8 dense rate = 2 x the bf16 rate in MEASURED_PEAKS.json (the
    file has no int8 figure; tcgen05 kind::i8 runs at twice the kind::f16 rate).  The kernel's own ceiling is
    the MMA issue floor: M128 x N192 x K32 per 96 clocks, 16 K-steps per tile -> 16 pairs/clk/SM.
    INT kernel: popc bound as in round 1 (16 popc32 per pair against the measured POPC-pipe peak).
    Both records carry the popc-equivalent figures BASELINE.json's metric names."""
    from slamfe import ops
    secs = launch_ms * 1e-3
    pairs_s = pairs_per_launch / secs
    popc = 16.0 * pairs_s / 1e9
    common = {"launch_ms": launch_ms, "descriptor_pairs_per_launch": pairs_per_launch,
              "gdesc_pairs_per_s": pairs_s / 1e9, "gpopc32_equiv_per_s": popc,
              "popc_peak_measured_gpopc32": peak_popc_g, "frac_of_popc_peak": popc / peak_popc_g if peak_popc_g else None,
              "traffic": None, "traffic_source": None}
    if ops.matcher_kernel() == "mma":
        peaks = _peaks()
        bf16 = peaks.get("bf16_tflops", 1647.8)
        peak = 2.0 * bf16
        ach = 2.0 * 8 * desc_bytes * pairs_s / 1e12
        mhz = sm_mhz or peaks.get("sm_max_mhz", 1965.0)
        floor = sms * 16.0 * mhz * 1e6
        common.update({
            "kernel": f"hamming_mma_kernel ({kernel_role}): tcgen05.mma kind::i8, M128 N192 K32, A in TMEM",
            "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TOP/s", "frac": ach / peak,
            "peak_source": ("2 x bf16_tflops of MEASURED_PEAKS.json (int8 dense = twice the bf16 rate; the file "
                            "has no int8 figure)" if peaks else "fallback 2 x 1647.8 TF"),
            "mma_issue_floor_gdesc_pairs_per_s": floor / 1e9, "frac_of_mma_issue_floor": pairs_s / floor,
            "peak_caveat": "MEASURED_PEAKS.json's bf16 figure is cuBLAS under its own power limit (SM clock median "
                           "1327 MHz under load there); this kernel holds the full clock, so frac against 2 x that "
                           "figure can exceed 1 — frac_of_mma_issue_floor (16 descriptor pairs/clk/SM, reached by "
                           "scripts/probe_tcgen05.cu rate3) is the hardware bound",
            "note": "achieved = 2 ops x 8*desc_bytes MACs per descriptor pair (one MAC per descriptor bit; the "
                    "kernel pads K to 512).  d = popc(q) + (1-2q).t accumulates exactly in int32, keys are "
                    "bit-identical to the XOR/POPC kernel.  The binding resource is not the tensor pipe but the "
                    "INT work around it: bit->byte expansion of the train tile (ALU) and the key fold of the "
                    "128 x 192 accumulators (DESIGN.md 2.1)"})
    else:
        common.update({
            "kernel": f"hamming_top2_kernel<256,2,9 adders> ({kernel_role})", "bound": "popc", "achieved": popc,
            "peak": peak_popc_g, "unit": "Gpopc32/s", "frac": popc / peak_popc_g if peak_popc_g else None,
            "peak_source": "measured on this GPU: slamfe_peak_kernel mode 0 (pure POPC chains)",
            "note": "achieved counts the ALGORITHMIC 16 popc32 per descriptor pair (SURVEY 8d); the carry-save "
                    "adders execute 7, so frac > 1 against the plain POPC-pipe peak is expected"})
    return common


# ------------------------------------------------------------------------------------------------
class RansacWorkload(Workload):
    metric = "frames/s of RANSAC-PnP inlier scoring (4096 hypotheses x 5000 correspondences per frame)"
    unit = "frames/s"
    scaling = "weak"
    dtype = "f64"

    def __init__(self, args, rank, world, dev):
        import torch
        from slamfe import synth, utils
        self.torch, self.dev, self.world, self.rank = torch, dev, world, rank
        self.H, self.N, self.F = 4096, 5000, args.frames if args.frames != 4541 else 128
        rng = np.random.default_rng(args.seed + 1000 * rank)
        base = [synth.pnp_problem(rng, self.N, self.H) for _ in range(4)]  # 4 distinct frames, tiled
        reps = -(-self.F // 4)
        cat = lambda k: np.concatenate([b[k] for b in base] * reps)[: self.F * base[0][k].shape[0]]
        self.host = {"T": cat(0), "pts": cat(1), "lp": cat(2), "rp": cat(3)}
        self.pinned = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in self.host.items()}
        self.d = {k: v.to(dev) for k, v in self.pinned.items()}
        self.pt_off = torch.arange(0, (self.F + 1) * self.N, self.N, dtype=torch.int32, device=dev)
        self.cams = (utils.K, utils.M1, utils.M2)
        self.units_local = self.F
        self.base = base
        self.config = {"workload": "configs[2]: RANSAC-PnP inlier scoring, 4096 hypotheses x 5000 correspondences "
                                   "per frame, batched over the sequence", "frames_per_gpu": self.F,
                       "hypotheses": self.H, "correspondences": self.N, "outlier_frac": 0.4,
                       "l2_policy": f"hypotheses+points {sum(v.nbytes for v in self.host.values()) >> 20} MB per GPU "
                                    f"(> L2 for >= 256 frames); operands are reused on chip by design",
                       "sharding": f"frames, {world} rank(s)"}

    def _kernel(self, d):
        from slamfe import ops
        self.out = ops.ransac_score(d["T"], d["pts"], d["lp"], d["rp"], *self.cams, pt_off=self.pt_off,
                                    n_frames=self.F, max_points=self.N)
        return self.out

    def _run(self, d):
        from slamfe import dist as sdist
        counts, best, mask = self._kernel(d)
        if self.world > 1:  # collective: every rank must call _run the same number of times
            sdist.all_gather_padded(best, lengths=np.full(self.world, self.F))
        return counts, best, mask

    def step(self):
        return self._run(self.d)

    def e2e_step(self):
        torch = self.torch
        d = {k: self.d[k].copy_(self.pinned[k], non_blocking=True) for k in self.pinned}
        counts, best, mask = self._run(d)
        hb = best.to("cpu", non_blocking=True)
        hm = mask.to("cpu", non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return sum(v.numel() * v.element_size() for v in self.pinned.values()), hb.numel() * 4 + hm.numel()

    def roofline(self, launch_ms, peak_fp64_gfma):
        fma = 24.0 * self.H * self.N * self.F
        return {"kernel": "ransac_score_kernel (all hypotheses x all correspondences x all frames, one launch)",
                "bound": "fp64", "achieved": fma / launch_ms / 1e6, "peak": peak_fp64_gfma, "unit": "GFMA64/s",
                "frac": fma / launch_ms / 1e6 / peak_fp64_gfma, "traffic": None, "launch_ms": launch_ms,
                "note": "algorithmic 24 fp64 FMA per (hypothesis, correspondence) (SURVEY 8d); the perspective "
                        "divisions are replaced by certified multiplication-form tests, and a one-sided fp32 "
                        "pre-filter drops hopeless pairs before any fp64 work (this workload's hypotheses are "
                        "mostly good, so nearly every pair takes the fp64 path); peak measured on this GPU "
                        "(slamfe_peak_kernel mode 2); not HBM-bound: operands are reused H or N times",
                "hyp_point_pairs_per_s": self.H * self.N * self.F / launch_ms * 1e3}

    def time_kernel(self):  # rank 0 only: must not issue collectives
        return lambda: self._kernel(self.d)

    def cpu_baseline(self):
        from oracle import ref_oracle as ora
        Ts, pts, lp, rp = self.base[0]
        n_h = 256
        t0 = time.perf_counter()
        oc, ob, om = ora.score_hypotheses(Ts[:n_h], pts, lp, rp, *self.cams)
        dt = time.perf_counter() - t0
        counts = self.out[0][0, :n_h].cpu().numpy()
        ok = bool(np.array_equal(counts, oc))
        return ({"value": 1.0 / (dt * self.H / n_h), "unit": self.unit, "cores": 1, "kind": "port",
                 "sample": f"{n_h} of the 4096 hypotheses of one frame ({dt:.1f} s), oracle restatement of the "
                           f"reference's NumPy transformation_agreement loop (single thread, as the reference)"},
                {"hypotheses_checked": n_h, "inlier_counts_bit_exact": ok})


# ------------------------------------------------------------------------------------------------
class LoopWorkload(Workload):
    metric = ("candidate keyframe pairs/s (loop-closure verification: every keyframe vs all prior keyframes, "
              "match + 888-iteration RANSAC-PnP per candidate)")
    unit = "candidate_pairs/s"
    scaling = "strong"

    def __init__(self, args, rank, world, dev):
        import torch
        from slamfe import dist as sdist, loop, synth, utils
        self.torch, self.dev, self.world, self.rank = torch, dev, world, rank
        self.K, self.n = args.keyframes, 2000
        n = self.n
        g = torch.Generator(device=dev).manual_seed(args.seed + 3)
        pool = torch.randint(0, 256, (self.K * n, 61), dtype=torch.uint8, device=dev, generator=g)
        pool[:, 60] &= 0x3F
        rng = np.random.default_rng(args.seed + 3)
        Kc = utils.K
        fx, cx, cy, fxb = Kc[0, 0], Kc[0, 2], Kc[1, 2], -synth.KITTI00_P1[0, 3]
        f32 = lambda a: a.astype(np.float32).astype(np.float64)
        xl = f32(rng.uniform(20, 1220, self.K * n))
        links = np.stack([xl, f32(xl - rng.uniform(2.5, 120, self.K * n)), f32(rng.uniform(5, 370, self.K * n))], axis=1)
        # planted revisits (cf. project.py:109-119): keyframe a re-observes keyframe b from a nearby pose
        self.revisits = [(a, b) for a, b in ((self.K - 3, 5), (self.K // 2, 11), (self.K - 40, self.K // 3),
                                             (self.K // 3 + 7, 2)) if 0 <= b < a < self.K]
        for a, b in self.revisits:
            flip = (torch.rand((n, 61, 8), device=dev, generator=g) < 0.06)
            bits = (flip * (2 ** torch.arange(8, device=dev))).sum(-1).to(torch.uint8)
            perm = torch.randperm(n, device=dev, generator=g)
            pool[a * n:(a + 1) * n] = pool[b * n:(b + 1) * n][perm] ^ bits
            pool[a * n:(a + 1) * n, 60] &= 0x3F
            lb = links[b * n:(b + 1) * n][perm.cpu().numpy()]
            Z = fxb / (lb[:, 0] - lb[:, 1])
            P3 = np.stack([(lb[:, 0] - cx) * Z / fx, (lb[:, 2] - cy) * Z / fx, Z], axis=1)
            R = synth._rodrigues(rng.normal(0, 0.02, 3))
            t = np.array([0.3, -0.05, -0.6]) + rng.normal(0, 0.05, 3)
            Pa = (P3 - t) @ R
            ok = Pa[:, 2] > 2.0
            xa = fx * Pa[:, 0] / Pa[:, 2] + cx + rng.normal(0, 0.3, n)
            ya = fx * Pa[:, 1] / Pa[:, 2] + cy + rng.normal(0, 0.3, n)
            la = f32(np.stack([xa, xa - fxb / Pa[:, 2], ya], axis=1))
            blk_l = links[a * n:(a + 1) * n]
            blk_l[ok] = la[ok]
        self.links_host = links
        self.pool = pool
        self.links = torch.from_numpy(links).to(dev)
        self.pool_pinned = torch.empty(pool.shape, dtype=torch.uint8, pin_memory=True).copy_(pool)
        self.links_pinned = torch.from_numpy(links).pin_memory()
        pairs = sdist.candidate_pairs(self.K)
        self.pairs = pairs
        b = sdist.candidate_blocks(pairs, np.full(self.K, n), world)
        self.lo, self.hi = int(b[rank]), int(b[rank + 1])
        self.bounds = b
        self.lengths = (b[1:] - b[:-1]) * n
        self.pair_lengths = b[1:] - b[:-1]
        self.kf_off = np.arange(self.K, dtype=np.int64) * n
        self.kf_cnt = np.full(self.K, n, dtype=np.int64)
        self.H = 888  # calc_ransac_iteration(40), loop_closure.py:425
        self.ver = loop.CandidateVerifier(block_pairs=args.loop_block)
        self.table = torch.empty(((self.hi - self.lo) * n,), dtype=torch.int32, device=dev)
        self.units_total = len(pairs)
        self.launches_per_step = 1 + 4 * -(-(self.hi - self.lo) // args.loop_block)
        self.desc_pairs_total = float(len(pairs)) * n * n
        self.config = {"workload": "configs[3]: loop-closure candidate verification, each keyframe vs all prior "
                                   "keyframes: Hamming match + RANSAC-PnP (888 hypotheses) per candidate",
                       "keyframes": self.K, "descriptors_per_keyframe": n, "candidate_pairs": len(pairs),
                       "ransac_iterations": self.H, "planted_revisits": self.revisits,
                       "l2_policy": "keyframe pool 55 MB is L2-resident by design (every keyframe is re-read ~K/2 "
                                    "times); match tables (808 MB total) and RANSAC inputs stream through HBM",
                       "sharding": f"candidate blocks balanced by Nq*Nt, {world} rank(s), pool replicated"}

    def _run(self, pool, links):
        from slamfe import dist as sdist
        res = self.ver.verify(pool, links, self.kf_off, self.kf_cnt, self.pairs[self.lo:self.hi], n_iter=self.H,
                              seed=1, pair_base=self.lo, key_table=self.table, sync=False)
        self.best = res["best_dev"]
        if self.world > 1:  # the path's only collectives: best-match tables and inlier tables
            self.gathered, _ = sdist.all_gather_padded(self.table, lengths=self.lengths)
            self.gathered_best, _ = sdist.all_gather_padded(self.best, lengths=self.pair_lengths)
        return self.best

    def step(self):
        return self._run(self.pool, self.links)

    def e2e_step(self):
        torch = self.torch
        self.pool.copy_(self.pool_pinned, non_blocking=True)
        self.links.copy_(self.links_pinned, non_blocking=True)
        best = self._run(self.pool, self.links)
        if not hasattr(self, "host_table"):
            self.host_table = torch.empty(self.table.shape, dtype=self.table.dtype, pin_memory=True)
            self.host_best = torch.empty(best.shape, dtype=best.dtype, pin_memory=True)
        self.host_table.copy_(self.table, non_blocking=True)
        self.host_best.copy_(best, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self.pool_pinned.numel() + self.links_pinned.numel() * 8, self.table.numel() * 4 + best.numel() * 4

    def time_kernel(self):
        from slamfe import ops
        torch = self.torch
        n_mine = min(self.hi - self.lo, 16384)
        mine = self.pairs[self.lo:self.lo + n_mine]
        dev = self.dev
        q_off = torch.from_numpy((mine[:, 0].astype(np.int64) * self.n).astype(np.int32)).to(dev)
        t_off = torch.from_numpy((mine[:, 1].astype(np.int64) * self.n).astype(np.int32)).to(dev)
        cnt = torch.full((n_mine,), self.n, dtype=torch.int32, device=dev)
        out_off = torch.arange(0, n_mine, dtype=torch.int32, device=dev) * self.n

        def fn():
            ops.hamming_pairs(self.pool, q_off, cnt, self.pool, t_off, cnt, out_off, n_mine, self.n, self.n, 61,
                              row_keys=self.table[:n_mine * self.n], out_rows_total=n_mine * self.n, best_only=True,
                              compact=True)
        self._kernel_pairs = float(n_mine) * self.n * self.n
        return fn

    def roofline(self, launch_ms, peak_popc):
        r = matcher_roofline("rows only, best-only, compact keys, via slamfe_hamming_top2_pairs", self._kernel_pairs,
                             launch_ms, peak_popc)
        r["note"] += "; the RANSAC stages (hypotheses + scoring, instruction-issue-bound behind an fp32 pre-filter) are the rest of a step"
        return r

    def cpu_baseline(self):
        """Rank 0: the reference's check_candidate_match on a few candidates of EVERY rank's block (cv2 match +
        the oracle's 888-iteration RANSAC loop), compared with the all-gathered tables (world > 1) or the
        local ones: match tables bit-exact, consensus sizes side by side."""
        import cv2
        from oracle import ref_oracle as ora
        from slamfe import utils
        cv2.setNumThreads(os.cpu_count() or 1)
        mm = cv2.BFMatcher(normType=cv2.NORM_HAMMING, crossCheck=False)
        pool = self.pool_pinned.numpy()
        n = self.n
        kf = lambda i: pool[i * n:(i + 1) * n]
        lk = lambda i: self.links_host[i * n:(i + 1) * n]
        per_rank = max(1, 6 // self.world)
        sample_idx = [int(self.bounds[r]) + k for r in range(self.world) for k in range(per_rank)
                      if int(self.bounds[r]) + k < int(self.bounds[r + 1])]
        index_of = {tuple(p): k for k, p in enumerate(self.pairs.tolist())}
        sample_idx += [index_of[tuple(r)] for r in self.revisits[:2] if tuple(r) in index_of]
        mm.match(kf(1), kf(0))
        t0 = time.perf_counter()
        out = []
        for g in sample_idx:  # check_candidate_match, loop_closure.py:405-436
            i, j = (int(v) for v in self.pairs[g])
            ms = mm.match(kf(i), kf(j))
            ti = np.fromiter((m.trainIdx for m in ms), np.int64, n)
            idx = ora.ransac_pnp_for_tracking_db(np.arange(n), ti, lk(i), lk(j), 40, utils.K, utils.M1, utils.M2)
            out.append((ti, np.fromiter((int(m.distance) for m in ms), np.int64, n), 0 if idx is None else len(idx)))
        dt = time.perf_counter() - t0
        if self.world > 1:   # what the all-gather delivered to this rank
            tables = self.gathered.cpu().numpy().view(np.uint32)          # (world, max rows)
            bests = self.gathered_best.cpu().numpy()                       # (world, max pairs, 2)
        else:
            tables = self.table.cpu().numpy().view(np.uint32)[None]
            bests = self.best.cpu().numpy()[None]
        ok, ratios = True, []
        for g, (ti, td, cnt) in zip(sample_idx, out):
            r = int(np.searchsorted(self.bounds, g, side="right") - 1)
            k = g - int(self.bounds[r])
            row = tables[r, k * n:(k + 1) * n]
            ok &= bool(np.array_equal(row & 0x3FFFFF, ti.astype(np.uint32)) and np.array_equal(row >> 22, td.astype(np.uint32)))
            ratios.append((int(bests[r, k, 1]), cnt))
        accepted, n_acc = {}, 0
        for r in range(self.world):
            cnt_r = int(self.bounds[r + 1] - self.bounds[r])
            n_acc += int((bests[r, :cnt_r, 1] > 120).sum())
        for rv in self.revisits:
            g = index_of.get(tuple(rv))
            if g is not None:
                r = int(np.searchsorted(self.bounds, g, side="right") - 1)
                accepted[str(tuple(rv))] = int(bests[r, g - int(self.bounds[r]), 1])
        return ({"value": len(sample_idx) / dt, "unit": self.unit, "cores": cv2.getNumThreads(), "kind": "port",
                 "sample": f"{len(sample_idx)} candidate pairs ({dt:.1f} s): cv2 {cv2.__version__} BFMatcher.match "
                           f"(all host threads) + the oracle's restatement of ransac_pnp's 888-iteration loop with "
                           f"cv2 EPnP (single thread, as the reference), loop_closure.py:405-436"},
                {"pairs_checked": len(sample_idx), "ranks_covered": self.world, "match_tables_bit_exact": ok,
                 "inliers_gpu_vs_cpu": ratios, "planted_revisit_inliers": accepted,
                 "candidates_accepted_over_120_inliers": n_acc, "planted": len(self.revisits),
                 "bytes_all_gathered_per_step": int((self.gathered.numel() + self.gathered_best.numel()) * 4)
                 if self.world > 1 else 0})


# ------------------------------------------------------------------------------------------------
class DenseWorkload(Workload):
    metric = "descriptor pairs/s (dense all-pairs Hamming top-2 sweep, 20000 x 20000 per frame)"
    unit = "descriptor_pairs/s"
    scaling = "strong"

    def __init__(self, args, rank, world, dev):
        import torch
        from slamfe import dist as sdist
        self.torch, self.dev, self.world, self.rank = torch, dev, world, rank
        self.n, self.F = 20000, args.frames if args.frames != 4541 else 64
        g = torch.Generator(device=dev).manual_seed(args.seed + 4)
        q = torch.randint(0, 256, (self.F * self.n, 61), dtype=torch.uint8, device=dev, generator=g)
        q[:, 60] &= 0x3F
        flip = torch.rand((self.F * self.n, 61), device=dev, generator=g) < 0.3   # ~8 % of bits via random bytes
        noise = torch.randint(0, 256, q.shape, dtype=torch.uint8, device=dev, generator=g) & \
            torch.randint(0, 256, q.shape, dtype=torch.uint8, device=dev, generator=g)
        t = q ^ (noise * flip)
        t[:, 60] &= 0x3F
        perm = torch.stack([torch.randperm(self.n, device=dev, generator=g) + f * self.n for f in range(self.F)])
        t = t[perm.reshape(-1)].contiguous()
        t[1::97] = t[0::97][: t[1::97].shape[0]]  # exact duplicate train rows: index tie-breaks matter
        self.q, self.t = q, t
        self.pinned = {"q": torch.empty(q.shape, dtype=torch.uint8, pin_memory=True).copy_(q),
                       "t": torch.empty(t.shape, dtype=torch.uint8, pin_memory=True).copy_(t)}
        b = sdist.train_slices(self.n, world)
        self.base, self.cnt_slice = int(b[rank]), int(b[rank + 1] - b[rank])
        fr = torch.arange(self.F, dtype=torch.int32, device=dev) * self.n
        self.q_off = fr
        self.t_off = fr + self.base
        self.q_cnt = torch.full((self.F,), self.n, dtype=torch.int32, device=dev)
        self.t_cnt = torch.full((self.F,), self.cnt_slice, dtype=torch.int32, device=dev)
        self.keys = torch.empty((self.F * self.n, 2), dtype=torch.int32, device=dev)
        self.units_total = float(self.F) * self.n * self.n
        self.config = {"workload": "configs[4]: dense-keypoint stress, 20k descriptors/frame all-pairs Hamming kNN "
                                   "(top-2) sweep", "frames": self.F, "descriptors": self.n,
                       "l2_policy": "per step every frame's 2 x 1.2 MB is read once from HBM; tiles are reused from "
                                    "L2/shared memory by design",
                       "sharding": f"train set sliced across {world} rank(s), queries replicated"}

    def _run(self, q, t):
        from slamfe import ops, dist as sdist
        ops.hamming_top2_batched(q, self.q_off, t, self.t_off, self.F, self.n, self.cnt_slice, 61, q_cnt=self.q_cnt,
                                 t_cnt=self.t_cnt, row_keys=self.keys, t_index_base=self.base)
        self.merged = sdist.gather_and_merge_top2(self.keys)
        return self.merged

    def step(self):
        return self._run(self.q, self.t)

    def e2e_step(self):
        """Host buffers in, host table out.  Every rank needs all queries (replicated by design of the
        train-slice sharding) but only ITS slice of every frame's train rows: one strided copy per frame."""
        torch = self.torch
        self.q.copy_(self.pinned["q"], non_blocking=True)
        t3, p3 = self.t.view(self.F, self.n, 61), self.pinned["t"].view(self.F, self.n, 61)
        lo, hi = self.base, self.base + self.cnt_slice
        for f in range(self.F):
            t3[f, lo:hi].copy_(p3[f, lo:hi], non_blocking=True)
        m = self._run(self.q, self.t)
        if not hasattr(self, "host_keys"):
            self.host_keys = torch.empty(m.shape, dtype=m.dtype, pin_memory=True)
        self.host_keys.copy_(m, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self.pinned["q"].numel() + self.F * self.cnt_slice * 61, m.numel() * 4

    def time_kernel(self):
        from slamfe import ops
        self._kernel_pairs = float(self.F) * self.n * self.cnt_slice
        return lambda: ops.hamming_top2_batched(self.q, self.q_off, self.t, self.t_off, self.F, self.n, self.cnt_slice,
                                                61, q_cnt=self.q_cnt, t_cnt=self.t_cnt, row_keys=self.keys,
                                                t_index_base=self.base)

    def roofline(self, launch_ms, peak_popc):
        return matcher_roofline("rows only, top-2, batched over frames", self._kernel_pairs, launch_ms, peak_popc)

    def cpu_baseline(self):
        import cv2
        from oracle import ref_oracle as ora
        cv2.setNumThreads(os.cpu_count() or 1)
        mm = cv2.BFMatcher(normType=cv2.NORM_HAMMING, crossCheck=False)
        q = self.pinned["q"].numpy()[: self.n]
        t = self.pinned["t"].numpy()[: self.n]
        mm.knnMatch(q[:500], t[:500], k=2)
        t0 = time.perf_counter()
        res = mm.knnMatch(q, t, k=2)
        dt = time.perf_counter() - t0
        k = self.merged[: self.n].cpu().numpy().view(np.uint32)
        ti = np.array([[m.trainIdx for m in r] for r in res], np.uint32)
        td = np.array([[int(m.distance) for m in r] for r in res], np.uint32)
        ok = bool(np.array_equal(k & 0x3FFFFF, ti) and np.array_equal(k >> 22, td))
        # the table every rank holds after all-gather + slamfe_merge_top2 (world > 1) against the C oracle
        # (oracle/hamming_oracle.c) on sampled query rows of the last frame, which cv2 above did not see
        f, rows = self.F - 1, np.arange(0, self.n, max(1, self.n // 1024))
        qf = self.pinned["q"].numpy()[f * self.n:(f + 1) * self.n][rows]
        tf = self.pinned["t"].numpy()[f * self.n:(f + 1) * self.n]
        oi, od = ora.knn2(qf, tf)
        kf = self.merged[f * self.n:(f + 1) * self.n].cpu().numpy().view(np.uint32)[rows]
        ok_oracle = bool(np.array_equal(kf & 0x3FFFFF, oi.astype(np.uint32)) and
                         np.array_equal(kf >> 22, od.astype(np.uint32)))
        return ({"value": float(self.n) * self.n / dt, "unit": self.unit, "cores": cv2.getNumThreads(),
                 "kind": "reference",
                 "sample": f"frame 0 (20000 x 20000, {dt:.2f} s): cv2 {cv2.__version__} BFMatcher.knnMatch(k=2), all "
                           f"host threads"},
                {"frames_checked": 2, "top2_tables_bit_exact": ok and ok_oracle,
                 "merged_over_ranks": self.world, "frame0_vs_cv2_knnMatch": ok,
                 f"frame{f}_{len(rows)}_rows_vs_c_oracle": ok_oracle,
                 "bytes_all_gathered_per_step": int(self.keys.numel() * 4 * self.world) if self.world > 1 else 0})


# ------------------------------------------------------------------------------------------------
def best_of(fn, reps=7, warm=2):
    for _ in range(warm):
        fn()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3


def run_dropin(args):
    """bench.py --workload dropin: per-call timings of the drop-in entry points against the reference's CPU
    calls on BASELINE configs[0] (one KITTI-00-shaped stereo pair, ~3k keypoints).  Both sides get HOST numpy
    inputs and return the reference's Python types; the CPU side (cv2 with all host threads + the oracle's
    restatement of the reference's Python) is this bench's cpu_baseline leg."""
    import cv2
    from slamfe import matching, ransac, synth, triangulation
    from oracle import ref_oracle as ora
    cv2.setNumThreads(os.cpu_count() or 1)
    rng = np.random.default_rng(0)
    n = 3000
    dl, dr, pl, pr = synth.stereo_frame(rng, n)
    nxt = synth.next_frame_descriptors(rng, dl, n)
    res = {"n_keypoints": n, "cv2": cv2.__version__, "cv2_threads": cv2.getNumThreads(), "host_cores": os.cpu_count()}
    cpu_m = cv2.BFMatcher(normType=cv2.NORM_HAMMING, crossCheck=False)
    cpu_lr = cv2.BFMatcher(normType=cv2.NORM_HAMMING, crossCheck=True)
    gpu_m, gpu_lr = matching.Matcher(crossCheck=False), matching.Matcher(crossCheck=True)

    def row(name, cpu_fn, gpu_fn, check=None):
        c, g = best_of(cpu_fn), best_of(gpu_fn)
        res[name] = {"cpu_ms": c, "slamfe_ms": g, "speedup": c / g}
        if check is not None:
            res[name]["identical"] = bool(check())

    same = lambda a, b: [(m.queryIdx, m.trainIdx, m.distance) for m in a] == [(m.queryIdx, m.trainIdx, m.distance) for m in b]
    row("MATCHER.match (database.py:54)", lambda: cpu_m.match(dl, nxt), lambda: gpu_m.match(dl, nxt),
        lambda: same(cpu_m.match(dl, nxt), gpu_m.match(dl, nxt)))
    row("MATCHER_LEFT_RIGHT.match crossCheck (matching.py:44)", lambda: cpu_lr.match(dl, dr), lambda: gpu_lr.match(dl, dr),
        lambda: same(cpu_lr.match(dl, dr), gpu_lr.match(dl, dr)))
    row("knnMatch k=2 (ex1.py:190)", lambda: cpu_m.knnMatch(dl, nxt, k=2), lambda: gpu_m.knnMatch(dl, nxt, k=2),
        lambda: all(same(a, b) for a, b in zip(cpu_m.knnMatch(dl, nxt, k=2), gpu_m.knnMatch(dl, nxt, k=2))))
    row("match_arrays (no DMatch objects)", lambda: cpu_m.match(dl, nxt), lambda: gpu_m.match_arrays(dl, nxt))
    ms = cpu_lr.match(dl, dr)
    mq = np.fromiter((m.queryIdx for m in ms), np.int32, len(ms))
    mt = np.fromiter((m.trainIdx for m in ms), np.int32, len(ms))
    row("extract_inliers_outliers (matching.py:48-69)", lambda: ora.extract_inliers_outliers(pl, pr, mq, mt),
        lambda: matching.extract_inliers_outliers(pl, pr, ms),
        lambda: np.array_equal(ora.extract_inliers_outliers(pl, pr, mq, mt)[0],
                               matching.extract_inliers_outliers(pl, pr, ms)[0]))
    inl, _ = ora.extract_inliers_outliers(pl, pr, mq, mt)
    _, links = ora.create_links(pl, pr, mq[inl], mt[inl])
    res["n_links"] = int(len(links))
    row("triangulate_links (triangulation.py:41-50)", lambda: ora.triangulate_links(links, ransac.P, ransac.Q),
        lambda: triangulation.triangulate_links(links, ransac.P, ransac.Q),
        lambda: np.allclose(ora.triangulate_links(links, ransac.P, ransac.Q),
                            triangulation.triangulate_links(links, ransac.P, ransac.Q), rtol=1e-9, atol=0))
    Ts, pts, lp, rp = synth.pnp_problem(rng, 2500, 888)
    row("transformation_agreement x1 (ransac.py:28-56)",
        lambda: ora.transformation_agreement(Ts[0], pts, lp, rp, ransac.K, ransac.M1, ransac.M2),
        lambda: ransac.transformation_agreement(Ts[0], pts, lp, rp),
        lambda: np.array_equal(ora.transformation_agreement(Ts[0], pts, lp, rp, ransac.K, ransac.M1, ransac.M2),
                               ransac.transformation_agreement(Ts[0], pts, lp, rp)))
    row("score 888 hypotheses x 2500 points (the loop ransac.py:155-182 at inliers_percent=40)",
        lambda: ora.score_hypotheses(Ts, pts, lp, rp, ransac.K, ransac.M1, ransac.M2),
        lambda: ransac.score_hypotheses(Ts, pts, lp, rp),
        lambda: np.array_equal(ora.score_hypotheses(Ts, pts, lp, rp, ransac.K, ransac.M1, ransac.M2)[0],
                               ransac.score_hypotheses(Ts, pts, lp, rp)[0]))
    print(json.dumps({"metric": "per-call latency of the drop-in entry points (ms), configs[0]", "unit": "ms",
                      "n_gpus": 1, "data": "synthetic", "config": {"workload": "configs[0]: one stereo pair, "
                                                                             "3000 keypoints per image"},
                      "calls": res}))




def run_createdb(args, ClockSampler):
    """bench.py --workload createdb: frames/s from host keypoints + descriptors to a filled tracking database
    (backend/database/database.py:30-89 + tracking_database.py:273-337), three ways on the same frames:
      soa     slamfe.trackdb.create_db: FrontEnd.run_host (track + device track ids) -> SoATrackingDB, no
              per-match Python object (the product's store);
      typed   slamfe.database.create_db: the same pipeline, then one add_frame per frame with the reference's
              argument types (Link and cv2.DMatch objects per match) into a counting stand-in for TrackingDB
              (the reference class is not on this box) — what patch(batched_db=True) costs;
      cpu     the reference's loop body on the host (cv2 + the oracle's restatement, no TrackingDB insertion)
              on a bounded sample — the cpu_baseline leg."""
    import torch
    import slamfe
    from slamfe import database as sdb, frontend, synth, trackdb, utils
    import bench
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    slamfe.load_library()
    F = args.frames if args.frames != 4541 else 1024
    seq_t = synth.torch_sequence(F, first_frame=0, seed=args.seed, device=dev)
    pinned = {k: torch.empty(seq_t[k].shape, dtype=seq_t[k].dtype, pin_memory=True).copy_(seq_t[k])
              for k in ("desc_l", "desc_r", "pts_l", "pts_r")}
    seq = frontend.PackedSequence(pinned["desc_l"].numpy(), pinned["desc_r"].numpy(), pinned["pts_l"].numpy(),
                                  pinned["pts_r"].numpy(), seq_t["l_off"], seq_t["r_off"], seq_t["n_l"], seq_t["n_r"],
                                  pinned)
    fe = frontend.FrontEnd()
    kw = dict(chunk_frames=min(args.chunk_frames, max(16, F // 4)), h_max=args.h_max, seed=args.seed, front_end=fe)
    db = trackdb.create_db(seq, **kw)
    with ClockSampler(0) as clocks:
        t0 = time.perf_counter()
        for _ in range(args.steps):
            db = trackdb.create_db(seq, **kw)
        dt_soa = (time.perf_counter() - t0) / args.steps
    t0 = time.perf_counter()
    tables, h2d, d2h = fe.run_host(seq, chunk_frames=kw["chunk_frames"], track=True, h_max=args.h_max, seed=args.seed,
                                   track_ids=True)
    dt_pipe = time.perf_counter() - t0
    t0 = time.perf_counter()
    trackdb.build(seq, tables)
    dt_build = time.perf_counter() - t0

    class CountingDB:
        def __init__(self):
            self.frameID_to_inliers_percent, self.frames, self.links, self.matches = {}, 0, 0, 0

        def add_frame(self, links, left_features, matches_to_previous_left=None, inliers=None):
            self.frames += 1
            self.links += len(links)
            self.matches += 0 if matches_to_previous_left is None else len(matches_to_previous_left)

    n_typed = min(F, 256)
    sub = frontend.pack_sequence([(seq.desc_l[int(seq.l_off[f]):int(seq.l_off[f]) + int(seq.n_l[f])],
                                   seq.desc_r[int(seq.r_off[f]):int(seq.r_off[f]) + int(seq.n_r[f])],
                                   seq.pts_l[int(seq.l_off[f]):int(seq.l_off[f]) + int(seq.n_l[f])],
                                   seq.pts_r[int(seq.r_off[f]):int(seq.r_off[f]) + int(seq.n_r[f])]) for f in range(n_typed)])
    cdb = CountingDB()
    sdb.create_db(sub, cdb, chunk_frames=kw["chunk_frames"], h_max=args.h_max, seed=args.seed)
    t0 = time.perf_counter()
    cdb = CountingDB()
    sdb.create_db(sub, cdb, chunk_frames=kw["chunk_frames"], h_max=args.h_max, seed=args.seed)
    dt_typed = time.perf_counter() - t0
    ok = db.check_consistency() and cdb.frames == n_typed and cdb.links == int(db.frame_off[n_typed])
    cpu = None
    if not args.no_cpu_baseline:
        n = max(2, min(args.cpu_sample_frames, F))
        frames = bench.host_frames(seq_t, 0, n)
        threads = bench.cpu_threads()
        t0 = time.perf_counter()
        bench.cpu_frames_pass(frames, utils.P, utils.Q)
        dtc = time.perf_counter() - t0
        cpu = {"value": n / dtc, "unit": "frames/s", "cores": threads, "kind": "port",
               "sample": f"first {n} frames ({dtc:.1f} s): the loop body of the reference's create_db (cv2 matchers with "
                         f"all host threads + the oracle's restatement of the Python loops), without the TrackingDB "
                         f"insertion; {bench.REFERENCE_ARM_NOTE}"}
    value = F / dt_soa
    print(json.dumps({
        "metric": "frames/s from host keypoints + descriptors to a filled tracking database (create_db)",
        "value": value, "unit": "frames/s", "n_gpus": 1, "steps": args.steps, "warmup": 1, "ms_per_step": dt_soa * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "create_db on a synthetic KITTI-00-shaped sequence (configs[1] frames)", "frames": F,
                   "keypoints_per_image": "2000-5000", "seed": args.seed},
        "clocks": clocks.summary(),
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "api": "slamfe.trackdb.create_db(PackedSequence) -> SoATrackingDB"},
        "gpu_launches": fe.last_launches * args.steps,
        "stages": {"pipeline_run_host_ms": dt_pipe * 1e3, "host_build_of_the_store_ms": dt_build * 1e3,
                   "links": int(db.frame_off[-1]), "tracks": db.track_num(), "links_on_tracks": db.link_num()},
        "typed_api": {"value": n_typed / dt_typed, "unit": "frames/s", "frames": n_typed,
                      "api": "slamfe.database.create_db(frames, db): Link + cv2.DMatch objects per match, one "
                             "add_frame per frame (patch(batched_db=True))",
                      "objects_built": cdb.links + cdb.matches},
        "cpu_baseline": cpu, "parity": {"store_consistent": bool(ok)}}))


WORKLOADS = {"ransac": RansacWorkload, "loop": LoopWorkload, "dense": DenseWorkload}


# ------------------------------------------------------------------------------------------------
def run_workload(w, args, ClockSampler, rank, world, local_rank, dev, steps=None, warm=None):
    """W warm-up steps, K timed steps between barriers with CUDA events (max over ranks), the e2e leg with
    host buffers, the roofline of the dominant kernel and (rank 0) the CPU baseline + in-run parity.
    Collective: every rank must call it.  Returns the record on rank 0, None elsewhere."""
    import torch
    import torch.distributed as tdist
    from slamfe import ops
    steps = args.steps if steps is None else steps
    warm = max(args.warmup, 3) if warm is None else warm

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(warm):
        w.step()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        ev0.record()
        for _ in range(steps):
            w.step()
        ev1.record()
        barrier()
    ms_per_step = allmax(ev0.elapsed_time(ev1)) / steps
    units = getattr(w, "units_total", None)
    if units is None:
        units = w.units_local * world
    value = units / (ms_per_step * 1e-3)

    e2e = None
    if not args.no_e2e:
        for _ in range(2):
            h2d, d2h = w.e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            h2d, d2h = w.e2e_step()
        barrier()
        e_ms = allmax((time.perf_counter() - t0) * 1e3) / steps
        e2e = {"value": units / (e_ms * 1e-3), "unit": w.unit, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": e_ms, "bytes_are": "per rank"}

    if rank != 0:
        return None
    fn = w.time_kernel()
    fn()
    torch.cuda.synchronize()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(2, steps)
    k0.record()
    for _ in range(reps):
        fn()
    k1.record()
    torch.cuda.synchronize()
    launch_ms = k0.elapsed_time(k1) / reps
    roofline = w.roofline(launch_ms, ops.measure_peak(2 if w.dtype == "f64" else 0) / 1e9)
    roofline.setdefault("peak_source", "measured on this GPU by slamfe_peak_kernel")
    cpu_baseline = parity = None
    if not args.no_cpu_baseline:  # results of the last (collective) step are still in place
        cpu_baseline, parity = w.cpu_baseline()
    line = {"metric": w.metric, "value": value, "unit": w.unit, "n_gpus": world, "steps": steps,
            "warmup": warm, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": w.scaling,
            "vs_baseline": None, "dtype": w.dtype, "data": "synthetic", "config": w.config,
            "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": w.launches_per_step * steps,
            "roofline": roofline, "cpu_baseline": cpu_baseline, "parity": parity}
    if hasattr(w, "desc_pairs_total"):
        line["descriptor_pairs_per_s"] = w.desc_pairs_total / (ms_per_step * 1e-3)
    return line


def run(args, ClockSampler):
    """bench.py --workload ransac|loop|dense|dropin: one workload, one JSON line on rank 0."""
    import torch
    import torch.distributed as tdist
    import slamfe
    from slamfe import dist as sdist
    if args.workload == "dropin":
        if int(os.environ.get("RANK", "0")) == 0:
            run_dropin(args)
        return
    if args.workload == "createdb":
        if int(os.environ.get("RANK", "0")) == 0:
            run_createdb(args, ClockSampler)
        return
    rank, world, local_rank = sdist.init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    dev = torch.device("cuda", local_rank)
    slamfe.load_library()
    w = WORKLOADS[args.workload](args, rank, world, dev)
    line = run_workload(w, args, ClockSampler, rank, world, local_rank, dev)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        tdist.barrier()
        tdist.destroy_process_group()
