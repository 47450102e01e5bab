#!/usr/bin/env python
"""Front-end hot-path benchmark (driver contract: one JSON line on rank 0).

Workload = BASELINE.json configs[1]: a synthetic KITTI-00-shaped stereo sequence of 4541 frames
PER GPU (1241x376, 2-5k 61-byte AKAZE descriptors per image): L<->R crossCheck matching + row
filter + links, triangulation of every link, and frame t<->t+1 forward/backward matching on the
filtered features.  One "step" = one pass over the whole sequence.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--frames F] [--impl slamfe|reference]

  value         frame pairs/s, whole job, inputs resident in HBM when the timed region starts
  e2e           same metric through the host-buffer API: pinned host inputs -> H2D -> kernels ->
                D2H of the result tables, all inside the timed region
  roofline      the matcher kernel (dominant): algorithmic 16 popc32 per descriptor pair over the
                live CUDA-event duration of its launch, against the popc-pipe peak measured on
                this GPU by slamfe_peak_kernel (MEASURED_PEAKS.json has no INT-pipe figure);
                HBM figures beside it
  cpu_baseline  the reference's CPU path (cv2.BFMatcher + the oracle's restatement of the
                reference's Python loops) on a bounded sample of the same frames, host cores
  --impl reference   the CPU path alone, same metric/config, bounded sample per step
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REFERENCE_ARM_NOTE = ("port of the reference's CPU path: triangulates ALL links of every frame (configs[1]'s stated "
                      "workload) where the reference's own create_db loop only triangulates the mutual-matched links "
                      "inside ransac_pnp_for_tracking_db (ransac.py:83) - about 30 % of this arm's time; the row filter "
                      "and create_links are vectorised NumPy where the reference loops in Python (conservative: "
                      "faster than the reference)")
METRIC = "frame pairs/s (stereo + consecutive-frame Hamming matching, row filter, triangulation, RANSAC-PnP)"
UNIT = "frame_pairs/s"


# --------------------------------------------------------------------------------------------
def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--frames", type=int, default=4541, help="frames per GPU (KITTI 00 has 4541)")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--impl", default="slamfe", choices=["slamfe", "reference"])
    ap.add_argument("--workload", default="sequence", choices=["sequence", "ransac", "loop", "dense", "dropin", "createdb"],
                    help="sequence = BASELINE configs[1] (the headline line); the others are configs[2..4], see "
                         "bench_extra.py")
    ap.add_argument("--keyframes", type=int, default=450, help="--workload loop: number of keyframes")
    ap.add_argument("--loop-block", type=int, default=8192, help="--workload loop: candidate pairs per launch")
    ap.add_argument("--h-max", type=int, default=128,
                    help="cap of RANSAC-PnP hypotheses per frame pair (calc_ransac_iteration gives ~57 at the "
                         "workload's 76 %% stereo inlier rate)")
    ap.add_argument("--cpu-sample-frames", type=int, default=48)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--chunk-frames", type=int, default=288, help="frames per chunk of the host pipeline (e2e)")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the loop-closure (configs[3]) and dense (configs[4]) strong-scaling sub-records")
    ap.add_argument("--extra-steps", type=int, default=3, help="timed steps of each sub-record")
    return ap.parse_args()


class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU during the timed region (NVML)."""

    def __init__(self, index=0, period=0.1):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        self.index, self.period = index, period
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {}
        for n in dir(nv):
            if n.startswith("nvmlClocksEventReason") or n.startswith("nvmlClocksThrottleReason"):
                v = getattr(nv, n)
                if isinstance(v, int) and v:
                    names.setdefault(v, n.replace("nvmlClocksEventReason", "").replace("nvmlClocksThrottleReason", ""))
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit and bit & (bit - 1) == 0:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr:
            self._thr.join()

    def summary(self):
        idle = {"GpuIdle", "None", "ApplicationsClocksSetting"}
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(r for r in self.reasons if r not in idle),
                "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------
# CPU path of the reference on host cores (oracle = checker / baseline only)
# --------------------------------------------------------------------------------------------
def cpu_frames_pass(frames, P, Q, collect=False):
    """The reference's per-frame work (database.py:12-27, :54-85; ransac.py:70-113) on the CPU:
    cv2 crossCheck match, row filter, create_links, np.linalg.svd triangulation per link,
    forward+backward cv2 matches, mutual check, RANSAC-PnP (cv2 EPnP hypotheses + NumPy scoring).
    Returns per-frame results when collect=True."""
    import cv2
    from oracle import ref_oracle as ora
    lr = cv2.BFMatcher(normType=cv2.NORM_HAMMING, crossCheck=True)
    mm = cv2.BFMatcher(normType=cv2.NORM_HAMMING, crossCheck=False)
    from slamfe import utils as _u
    K, M1, M2 = _u.K, _u.M1, _u.M2
    out, prev_feat, prev_links = [], None, None
    for dl, dr, pl, pr in frames:
        ms = lr.match(dl, dr)
        mq = np.fromiter((m.queryIdx for m in ms), np.int32, len(ms))
        mt = np.fromiter((m.trainIdx for m in ms), np.int32, len(ms))
        inl, _ = ora.extract_inliers_outliers(pl, pr, mq, mt)
        valid, links = ora.create_links(pl, pr, mq[inl], mt[inl])
        feat = dl[valid]
        xyz = ora.triangulate_links(links, P, Q)
        fwd_t = fwd_d = bwd_t = None
        if prev_feat is not None and len(prev_feat) and len(feat):
            fwd = mm.match(prev_feat, feat)
            bwd = mm.match(feat, prev_feat)
            fwd_t = np.fromiter((m.trainIdx for m in fwd), np.int32, len(fwd))
            fwd_d = np.fromiter((m.distance for m in fwd), np.int32, len(fwd))
            bwd_t = np.fromiter((m.trainIdx for m in bwd), np.int32, len(bwd))
        good = best_idx = None
        if fwd_t is not None:
            # database.py:67-85: mutual check, then ransac_pnp_for_tracking_db with the frame's stereo inlier rate
            good = np.nonzero(bwd_t[fwd_t] == np.arange(len(fwd_t)))[0]
            if len(good) >= 4:
                pct = 100 * (len(inl) / len(ms))
                best_idx = ora.ransac_pnp_for_tracking_db(good, fwd_t[good], prev_links, links, pct, K, M1, M2)
        if collect:
            out.append({"mq": mq, "mt": mt, "inl": inl, "links": links, "xyz": xyz, "fwd_t": fwd_t, "fwd_d": fwd_d,
                        "bwd_t": bwd_t, "good": good, "best_idx": best_idx,
                        "ransac_inliers": None if best_idx is None else len(best_idx)})
        prev_feat, prev_links = feat, links
    return out


def host_frames(seq_t, first, count):
    """Slice `count` frames out of the packed (torch, possibly CUDA) sequence as numpy arrays."""
    frames = []
    for f in range(first, first + count):
        lo, n = int(seq_t["l_off"][f]), int(seq_t["n_l"][f])
        ro, m = int(seq_t["r_off"][f]), int(seq_t["n_r"][f])
        frames.append((seq_t["desc_l"][lo:lo + n].cpu().numpy(), seq_t["desc_r"][ro:ro + m].cpu().numpy(),
                       seq_t["pts_l"][lo:lo + n].cpu().numpy(), seq_t["pts_r"][ro:ro + m].cpu().numpy()))
    return frames


def cpu_threads():
    import cv2
    cv2.setNumThreads(os.cpu_count() or 1)
    return cv2.getNumThreads()


def run_reference(args, rank):
    """--impl reference: the CPU path alone.  Rank 0 only; other ranks exit 0 without work."""
    if rank != 0:
        return
    import torch
    import cv2
    from slamfe import synth, utils
    threads = cpu_threads()
    n = max(2, min(args.cpu_sample_frames, args.frames))
    seq = synth.torch_sequence(n, first_frame=0, seed=args.seed, device="cpu")
    frames = host_frames(seq, 0, n)
    for _ in range(max(args.warmup, 0)):
        cpu_frames_pass(frames[:4], utils.P, utils.Q)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_frames_pass(frames, utils.P, utils.Q)
    dt = (time.perf_counter() - t0) / args.steps
    value = (n - 1) / dt
    sample = (f"{n} consecutive frames of the workload per step (cv2 {cv2.__version__}, numpy {np.__version__}); "
              f"{REFERENCE_ARM_NOTE}")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def workload_config(args, world):
    return {"workload": "configs[1]: synthetic 4541-frame KITTI-00-shaped stereo sequence per GPU "
                        "(stereo + consecutive-frame matching + triangulation + RANSAC-PnP per frame pair)",
            "frames_per_gpu": args.frames, "frames_total": args.frames * world,
            "keypoints_per_image": "2000-5000 (mean 3500)", "descriptor_bytes": 61, "image": "1241x376",
            "l2_policy": "inputs (~2 GB/GPU) are larger than L2; no flush needed", "seed": args.seed,
            "sharding": f"contiguous frame-pair blocks, {world} rank(s), +1 halo frame per rank"}


# --------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    from slamfe import dist as sdist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.workload != "sequence":
        import bench_extra
        bench_extra.run(args, ClockSampler)
        return

    import torch
    import torch.distributed as tdist
    import slamfe
    from slamfe import frontend, ops, synth, utils
    rank, world, local_rank = sdist.init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU path")
    dev = torch.device("cuda", local_rank)
    slamfe.load_library()
    old_affinity = sdist.bind_to_gpu_numa(local_rank) if os.environ.get("SLAMFE_BIND_NUMA") == "1" else None

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()

    # ---- synthetic inputs, generated on the GPU; rank r owns frames [r*F, (r+1)*F) + 1 halo ----
    F = args.frames
    halo = 1 if world > 1 and rank < world - 1 else 0
    seq_t = synth.torch_sequence(F + halo, first_frame=rank * F, seed=args.seed, device=dev)
    ds = frontend.DeviceSequence(
        seq_t["desc_l"], seq_t["desc_r"], seq_t["pts_l"], seq_t["pts_r"],
        torch.from_numpy(seq_t["l_off"]).to(dev), torch.from_numpy(seq_t["r_off"]).to(dev),
        torch.from_numpy(seq_t["n_l"]).to(dev), torch.from_numpy(seq_t["n_r"]).to(dev),
        F + halo, int(seq_t["n_l"].max()), int(seq_t["n_r"].max()))
    fe = frontend.FrontEnd()
    pairs_total = F * world - 1  # the last frame of the job has no successor

    n_rows_own = int(seq_t["l_off"][F])
    row_lengths = None
    if world > 1:  # static per-rank table sizes, exchanged once outside the timed region
        _, row_lengths = sdist.all_gather_padded(torch.zeros((n_rows_own, 1), dtype=torch.int8, device=dev))

    def gather_tables(o):
        """The only collective on the path: all-gather of the per-shard result tables (fixed
        stride) — per left row [stereo match, best forward key], per frame [n_matches, n_links]."""
        if world == 1:
            return None
        rows = torch.stack([o["match_t"][:n_rows_own], o["fwd_keys"][:n_rows_own]], dim=1)
        g_rows, _ = sdist.all_gather_padded(rows, lengths=row_lengths)
        summ = torch.stack([o["n_matches"][:F], o["n_links"][:F]], dim=1)
        g_frames, _ = sdist.all_gather_padded(summ, lengths=np.full(world, F))
        return g_rows, g_frames

    def step():
        o = fe.track(ds, h_max=args.h_max, seed=args.seed)
        gather_tables(o)
        return o

    for _ in range(max(args.warmup, 3)):
        out = step()
    barrier()

    # ---- timed region: K steps, inputs resident in HBM ----
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        ev0.record()
        for _ in range(args.steps):
            out = step()
        ev1.record()
        barrier()
    ms = ev0.elapsed_time(ev1)
    t_ms = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        tdist.all_reduce(t_ms, op=tdist.ReduceOp.MAX)
    ms_per_step = float(t_ms.item()) / args.steps
    value = pairs_total / (ms_per_step * 1e-3)

    launches_per_step, matcher_kind = fe.last_launches, ops.matcher_kernel()
    n_links = out["n_links"].cpu().numpy()
    # algorithmic descriptor pairs of this rank: its F stereo frames + its consecutive pairs
    desc_pairs_local = frontend.descriptor_pairs(seq_t["n_l"][:F], seq_t["n_r"][:F], n_links[:F + halo])
    dp = torch.tensor([desc_pairs_local], device=dev, dtype=torch.float64)
    if world > 1:
        tdist.all_reduce(dp, op=tdist.ReduceOp.SUM)
    desc_pairs_total = float(dp.item())

    # ---- e2e: host buffers in, host tables out, copies inside the timed region ----
    # FrontEnd.run_host is the host-buffer entry point: pinned inputs are copied in chunks of
    # frames, chunk c+1's H2D, chunk c's kernels and chunk c-1's D2H overlap on three streams.
    e2e = None
    if not args.no_e2e:
        pinned_in = {k: torch.empty(seq_t[k].shape, dtype=seq_t[k].dtype, pin_memory=True).copy_(seq_t[k])
                     for k in ("desc_l", "desc_r", "pts_l", "pts_r")}
        host_seq = frontend.PackedSequence(
            pinned_in["desc_l"].numpy(), pinned_in["desc_r"].numpy(), pinned_in["pts_l"].numpy(),
            pinned_in["pts_r"].numpy(), seq_t["l_off"], seq_t["r_off"], seq_t["n_l"], seq_t["n_r"], pinned_in)
        fe2 = frontend.FrontEnd()
        h2d = d2h = 0

        e2e_tables = None

        def e2e_step():
            nonlocal h2d, d2h, e2e_tables
            e2e_tables, h2d, d2h = fe2.run_host(host_seq, chunk_frames=args.chunk_frames, device=dev, track=True,
                                       h_max=args.h_max, seed=args.seed)
            gather_tables(fe2._out)

        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        barrier()
        e_ms = torch.tensor([(time.perf_counter() - t0) * 1e3], device=dev, dtype=torch.float64)
        if world > 1:
            tdist.all_reduce(e_ms, op=tdist.ReduceOp.MAX)
        e2e = {"value": pairs_total / (float(e_ms.item()) / args.steps * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": float(e_ms.item()) / args.steps,
               "api": f"FrontEnd.run_host(PackedSequence, chunk_frames={args.chunk_frames}, track=True)",
               "gpu_launches_per_step": fe2.last_launches, "ransac_pairs_rerun_at_full_count": fe2.last_truncated}
        # what the link itself delivers: the same bytes as plain pinned copies, no kernels (per rank, all
        # ranks at once: at N > 1 the ranks share the host's memory system and PCIe root complexes)
        dev_in = {k: torch.empty_like(v, device=dev) for k, v in pinned_in.items()}
        outs = {k: fe2._out[k] for k in frontend.RESULT_KEYS}
        pin_out = {k: torch.empty(v.shape, dtype=v.dtype, pin_memory=True) for k, v in outs.items()}
        barrier()
        t0 = time.perf_counter()
        for k, v in pinned_in.items():
            dev_in[k].copy_(v, non_blocking=True)
        torch.cuda.synchronize()
        t_h2d = time.perf_counter() - t0
        barrier()
        t0 = time.perf_counter()
        for k, v in outs.items():
            pin_out[k].copy_(v, non_blocking=True)
        torch.cuda.synchronize()
        t_d2h = time.perf_counter() - t0
        b_in = sum(v.numel() * v.element_size() for v in pinned_in.values())
        b_out = sum(v.numel() * v.element_size() for v in outs.values())
        link = torch.tensor([b_in / t_h2d / 1e9, b_out / t_d2h / 1e9], device=dev, dtype=torch.float64)
        if world > 1:
            tdist.all_reduce(link, op=tdist.ReduceOp.MIN)
        e2e["pcie"] = {"h2d_gbs_per_rank_min": float(link[0].item()), "d2h_gbs_per_rank_min": float(link[1].item()),
                       "h2d_ms_at_that_rate": b_in / float(link[0].item()) / 1e6,
                       "note": "plain pinned copies of the step's inputs / result tables, all ranks concurrently, "
                               "slowest rank; the e2e step cannot be shorter than the H2D time: the inputs are "
                               "61-byte descriptors of both images (485 KB per frame) and cannot be made smaller"}
        del dev_in, pin_out, outs
        # the host pipeline (chunked, 4 streams) must deliver exactly the tables of the resident run
        torch.cuda.synchronize()
        resident = {k: out[k].cpu().numpy() for k in e2e_tables if k in out}
        FFh = F + halo
        n_pairs_loc = FFh - 1
        lo_all = seq_t["l_off"].astype(np.int64)
        width = np.diff(lo_all)
        within = np.arange(int(lo_all[-1])) - np.repeat(lo_all[:-1], width)       # row index inside its frame
        link_rows = within < np.repeat(resident["n_links"][:FFh].astype(np.int64), width)
        left_rows = within < np.repeat(seq_t["n_l"][:FFh].astype(np.int64), width)
        trunc = resident["n_hyp_full"][:n_pairs_loc] > args.h_max
        same = {}
        for k, v in resident.items():
            g = e2e_tables[k]
            if k in ("n_matches", "n_links"):
                same[k] = bool(np.array_equal(v[:FFh], g[:FFh]))
            elif k in ("best", "n_good", "n_hyp", "n_hyp_full", "pose", "pose_status"):
                if k in ("best", "pose", "pose_status") and trunc.any():
                    continue   # run_host re-ran the truncated pairs at their full count; the resident run did not
                same[k] = bool(np.array_equal(v[:n_pairs_loc], g[:n_pairs_loc]))
            elif k == "match_t":
                same[k] = bool(np.array_equal(v[left_rows], g[left_rows]))
            elif k == "inlier_fwd" and trunc.any():
                continue
            else:
                rows = link_rows.copy()
                if k in ("inlier_fwd", "fwd_keys"):     # the last frame has no successor
                    rows[int(lo_all[FFh - 1]):] = False
                same[k] = bool(np.array_equal(v[rows], g[rows]))
        e2e["tables_equal_resident_run"] = all(same.values())
        e2e["tables_compared"] = {k: v for k, v in sorted(same.items())}
        del pinned_in, host_seq, fe2, resident, e2e_tables

    # ---- roofline of the dominant kernel: the stereo matcher launch, timed alone with events ----
    roofline = cpu_baseline = parity = None
    if rank == 0:
        FF = ds.n_frames
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(2, args.steps)
        torch.cuda.synchronize()
        k0.record()
        for _ in range(reps):
            ops.hamming_top2_batched(ds.desc_l, ds.l_off, ds.desc_r, ds.r_off, FF, ds.max_nl, ds.max_nr, 61,
                                     q_cnt=ds.n_l, t_cnt=ds.n_r, want_cols=True, best_only=True,
                                     row_keys=out["lr_row_keys"], col_keys=out["lr_col_keys"])
        k1.record()
        torch.cuda.synchronize()
        launch_ms = k0.elapsed_time(k1) / reps
        nl, nr = seq_t["n_l"][:FF].astype(np.int64), seq_t["n_r"][:FF].astype(np.int64)
        pairs_launch = float(np.sum(nl * nr))
        peak_popc = ops.measure_peak(0) / 1e9
        peak_mix = ops.measure_peak(1) / 1e9
        peak_fp64 = ops.measure_peak(2) / 1e9
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        alg_bytes = float(61 * np.sum(nl + nr) + 8 * np.sum(nl) + 4 * np.sum(nr))
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        sm_mhz = clocks.summary()["sm_mhz"] or 1965.0
        import bench_extra
        roofline = bench_extra.matcher_roofline("stereo L<->R launch, all frames: rows + column minima, best-only",
                                                pairs_launch, launch_ms, peak_popc, sm_mhz=sm_mhz, sms=sms)
        # DRAM traffic of this launch: from the committed ncu capture of the same command at full size (the
        # bench cannot run under ncu); null for any other size / world
        kind = ops.matcher_kernel()
        traffic_file = os.path.join("profiles", "r02_matcher_traffic.json" if kind == "mma" else "r01_matcher_traffic.json")
        try:
            if F == 4541 and world == 1:
                roofline["traffic"] = json.load(open(os.path.join(ROOT, traffic_file)))["dram_bytes_per_launch"]
                roofline["traffic_source"] = (f"{traffic_file}: ncu dram__bytes_read.sum + dram__bytes_write.sum of "
                                              f"this launch, captured in a separate run of this command (not live)")
        except Exception:
            pass
        roofline.update({
            "peak_matcher_mix_gops": peak_mix, "fp64_fma_peak_gfma": peak_fp64,
            "hbm": {"algorithmic_bytes_per_launch": alg_bytes, "achieved_gbs": alg_bytes / (launch_ms * 1e-3) / 1e9,
                    "peak_gbs": hbm_peak, "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                    "frac": alg_bytes / (launch_ms * 1e-3) / 1e9 / hbm_peak},
            "share_of_step": launch_ms / ms_per_step,
        })
        if kind == "int":
            # Executed-instruction model of the INT kernel (csrc/hamming.cu, kAdders = 9): per descriptor pair
            # and lane 7 POPC on the XU pipe (16 lanes/clk/SM) and ~30 ALU ops (64 lanes/clk/SM)
            pipe_bound_pairs = sms * sm_mhz * 1e6 / max(7.0 / 16.0, 30.0 / 64.0)
            roofline["executed"] = {"popc_per_pair": 7, "alu_ops_per_pair": 30, "sm_mhz": sm_mhz,
                                    "pipe_bound_gdesc_pairs_per_s": pipe_bound_pairs / 1e9,
                                    "frac_of_executed_pipe_bound": pairs_launch / (launch_ms * 1e-3) / pipe_bound_pairs}

        # ---- CPU baseline on a bounded sample + parity of the GPU tables on those frames ----
        if old_affinity is not None:  # the CPU baseline uses every host thread
            os.sched_setaffinity(0, old_affinity)
        if not args.no_cpu_baseline:
            import cv2
            from oracle import ref_oracle as ora
            # windows of consecutive frames spread over the whole sequence, centred on chunk boundaries of the
            # host pipeline (where a bug in the chunking would show), plus the first and the last frames
            n_total = max(2, min(args.cpu_sample_frames, FF))
            wlen = max(2, min(8, n_total))
            n_win = max(1, n_total // wlen)
            bounds = frontend.chunk_bounds(FF, args.chunk_frames)
            anchors = [0, FF - wlen] + [b - wlen // 2 for b in bounds[1:-1]]
            step_a = max(1, len(anchors) // n_win)
            starts = sorted({int(min(max(a, 0), FF - wlen)) for a in ([0, FF - wlen] + anchors[2::step_a])})[:n_win]
            threads = cpu_threads()
            cpu_frames_pass(host_frames(seq_t, 0, min(3, FF)), utils.P, utils.Q)
            host, _, _ = frontend.results_to_host(out)
            trk = {k: out[k].cpu().numpy() for k in ("good_j", "n_good", "best", "best_mask", "pts", "lpix", "rpix",
                                                     "n_hyp", "n_hyp_full", "pose", "pose_status")}
            pose_gpu, pose_cpu = [], []

            def pose_error(T, Rg, tg):
                dR = T[:, :3] @ Rg.T
                return float(np.linalg.norm(dR - dR.T) / (2 * np.sqrt(2))), float(np.linalg.norm(T[:, 3] - tg))

            ok, worst, mutual_ok, dt, n_pairs_cpu, frames_checked = True, 0.0, True, 0.0, 0, 0
            ratios, rec_gpu, rec_cpu, gt_sizes = [], [], [], []
            for s0 in starts:
                frames = host_frames(seq_t, s0, wlen)
                t0 = time.perf_counter()
                res = cpu_frames_pass(frames, utils.P, utils.Q, collect=True)
                dt += time.perf_counter() - t0
                n_pairs_cpu += wlen - 1
                frames_checked += wlen
                for i, r in enumerate(res):
                    f = s0 + i
                    lo, k = int(seq_t["l_off"][f]), len(r["inl"])
                    mt = host["match_t"][lo:lo + int(seq_t["n_l"][f])]
                    ok &= bool(np.array_equal(np.nonzero(mt >= 0)[0], r["mq"]) and np.array_equal(mt[mt >= 0], r["mt"]))
                    ok &= bool(host["n_links"][f] == k and np.array_equal(host["link_src"][lo:lo + k], r["mq"][r["inl"]]))
                    if k:
                        got = host["xyz"][lo:lo + k].astype(np.float64)
                        worst = max(worst, float((np.linalg.norm(got - r["xyz"], axis=1) /
                                                  np.linalg.norm(r["xyz"], axis=1)).max()))
                    if i + 1 < wlen and res[i + 1]["fwd_t"] is not None:
                        nxt = res[i + 1]
                        fi, fd = ops.keys_to_numpy(host["fwd_keys"][lo:lo + k])
                        ok &= bool(np.array_equal(fi, nxt["fwd_t"]) and np.array_equal(fd, nxt["fwd_d"]))
                        lo1, k1n = int(seq_t["l_off"][f + 1]), len(nxt["inl"])
                        bi, _ = ops.keys_to_numpy(host["bwd_keys"][lo1:lo1 + k1n])
                        ok &= bool(np.array_equal(bi, nxt["bwd_t"]))
                        if nxt["good"] is None:
                            continue
                        # tracking stages: mutual matches bit-exact ...
                        ng = len(nxt["good"])
                        mutual_ok &= bool(trk["n_good"][f] == ng and np.array_equal(trk["good_j"][lo:lo + ng], nxt["good"]))
                        if nxt["ransac_inliers"]:
                            ratios.append(float(trk["best"][f, 1]) / nxt["ransac_inliers"])
                        # ... and the RANSAC consensus of both arms against the GROUND-TRUTH motion of the
                        # synthetic sequence (synth.frame_motion): recall of the correspondences that agree
                        # with the true pose under the reference's own test (ransac.py:28-56)
                        Rg, tg = synth.frame_motion(args.seed, rank * F + f + 1)
                        gt = ora.transformation_agreement(np.hstack([Rg, tg[:, None]]), trk["pts"][lo:lo + ng],
                                                          trk["lpix"][lo:lo + ng], trk["rpix"][lo:lo + ng],
                                                          utils.K, utils.M1, utils.M2)
                        if gt.sum() >= 4:
                            gmask = trk["best_mask"][lo:lo + ng].astype(bool) if trk["best"][f, 0] >= 0 else np.zeros(ng, bool)
                            cmask = np.zeros(ng, bool)
                            if nxt["best_idx"] is not None:
                                cmask[nxt["best_idx"]] = True
                            rec_gpu.append(float((gmask & gt).sum()) / gt.sum())
                            rec_cpu.append(float((cmask & gt).sum()) / gt.sum())
                            gt_sizes.append(int(gt.sum()))
                            # pose: device refit on the device consensus vs the reference's recipe (cv2 EPnP
                            # on the CPU arm's consensus, ransac.py:185-193), both against the true motion
                            if trk["pose_status"][f] > 0:
                                pose_gpu.append(pose_error(trk["pose"][f], Rg, tg))
                            if cmask.sum() >= 4:
                                okc, rv, tv = cv2.solvePnP(trk["pts"][lo:lo + ng][cmask], trk["lpix"][lo:lo + ng][cmask],
                                                           utils.K, np.zeros((5, 1)), flags=cv2.SOLVEPNP_EPNP)
                                if okc:
                                    pose_cpu.append(pose_error(utils.rodriguez_to_mat(rv, tv), Rg, tg))
            cpu_baseline = {"value": n_pairs_cpu / dt, "unit": UNIT, "cores": threads, "kind": "port",
                            "sample": f"{len(starts)} windows of {wlen} consecutive frames spread over the sequence "
                                      f"(starts {starts}; {dt:.1f} s): cv2 {cv2.__version__} BFMatcher crossCheck + "
                                      f"fwd/bwd match with all host threads, oracle restatement of the reference's "
                                      f"Python row filter / create_links / per-link np.linalg.svd triangulation / "
                                      f"mutual check / RANSAC-PnP loop with cv2 EPnP (single thread, as the "
                                      f"reference); {REFERENCE_ARM_NOTE}"}
            rg, rc = np.array(rec_gpu), np.array(rec_cpu)
            n_pairs_loc = FF - 1
            parity = {"frames_checked": frames_checked, "windows": starts, "match_tables_bit_exact": ok,
                      "xyz_max_rel_err": worst, "xyz_tolerance": 1e-5, "mutual_matches_bit_exact": mutual_ok,
                      "ransac_vs_ground_truth": {
                          "pairs": len(rec_gpu), "median_ground_truth_inliers": float(np.median(gt_sizes)) if gt_sizes else None,
                          "gpu_p3p_frac_pairs_recall_ge_0.9": float((rg >= 0.9).mean()) if len(rg) else None,
                          "cpu_epnp_frac_pairs_recall_ge_0.9": float((rc >= 0.9).mean()) if len(rc) else None,
                          "gpu_p3p_median_recall": float(np.median(rg)) if len(rg) else None,
                          "cpu_epnp_median_recall": float(np.median(rc)) if len(rc) else None,
                          "gpu_at_least_cpu": bool(len(rg) == 0 or (rg >= 0.9).mean() >= (rc >= 0.9).mean()),
                          "pose_error_median_rad_m": {
                              "gpu_refit": [float(x) for x in np.median(np.array(pose_gpu), axis=0)] if pose_gpu else None,
                              "cpu_epnp_refit": [float(x) for x in np.median(np.array(pose_cpu), axis=0)] if pose_cpu else None},
                          "frac_pairs_pose_within_1cm_1mrad": {
                              "gpu_refit": float(np.mean([a < 1e-3 and b < 0.01 for a, b in pose_gpu])) if pose_gpu else None,
                              "cpu_epnp_refit": float(np.mean([a < 1e-3 and b < 0.01 for a, b in pose_cpu])) if pose_cpu else None},
                          "note": "ground truth = correspondences agreeing with the sequence's true motion under "
                                  "transformation_agreement; both arms run the reference's own iteration count "
                                  "(calc_ransac_iteration, ~57 here) from different random samples and solvers "
                                  "(GPU: P3P minimal solver, CPU: cv2 EPnP on 4 points)"},
                      "ransac_inlier_count_ratio_gpu_over_cpu": {"median": float(np.median(ratios)) if ratios else None,
                                                                 "min": min(ratios) if ratios else None,
                                                                 "max": max(ratios) if ratios else None},
                      "ransac_pairs_over_h_max": int((trk["n_hyp_full"][:n_pairs_loc] > args.h_max).sum()),
                      "ransac_max_iterations_asked": int(trk["n_hyp_full"][:n_pairs_loc].max()) if n_pairs_loc else 0,
                      "h_max": args.h_max}

    # ---- sub-records: north_star's multi-GPU targets (configs[3] loop closure, configs[4] dense sweep), strong
    # scaling at the current world size, in-run parity against the oracle (collective: all ranks) ----
    extra = None
    if not args.no_extras:
        import bench_extra
        del ds, seq_t, out, fe
        torch.cuda.empty_cache()
        extra = {}
        sub = argparse.Namespace(**vars(args))
        sub.frames = 4541   # the sub-workloads' own default sizes (64 dense frames, 450 keyframes)
        for name in ("loop", "dense"):
            w = bench_extra.WORKLOADS[name](sub, rank, world, dev)
            rec = bench_extra.run_workload(w, sub, ClockSampler, rank, world, local_rank, dev,
                                           steps=max(1, args.extra_steps), warm=3)
            if rank == 0:
                extra[name] = rec
            del w
            torch.cuda.empty_cache()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args, world),
            "descriptor_pairs_per_s": desc_pairs_total / (ms_per_step * 1e-3),
            "descriptor_pairs_per_step": desc_pairs_total,
            "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": launches_per_step * args.steps,
            "matcher_kernel": matcher_kind,
            "roofline": roofline, "cpu_baseline": cpu_baseline, "parity": parity, "extra": extra,
        }
        print(json.dumps(line))
    if world > 1:
        tdist.barrier()
        tdist.destroy_process_group()


if __name__ == "__main__":
    main()
