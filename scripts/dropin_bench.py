"""Per-call timings of the drop-in entry points against the reference's CPU calls (BASELINE
configs[0]: one KITTI-00-shaped stereo pair, ~3k AKAZE keypoints, matched + triangulated).

    python scripts/dropin_bench.py [--out gpurun_out/dropin.json]

Both sides get HOST numpy inputs and return the reference's Python return types (cv2.DMatch tuples,
float64 arrays, bool masks): the slamfe side therefore includes pinned staging, H2D, the kernels, D2H
and the construction of the Python objects.  The CPU side is cv2 4.13 with all host threads and the
oracle's restatement of the reference's Python loops (test infrastructure, used here as the baseline).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cv2  # noqa: E402
import slamfe  # noqa: E402,F401
from slamfe import matching, ransac, synth, triangulation  # noqa: E402
from oracle import ref_oracle as ora  # noqa: E402


def best_of(fn, reps=7, warm=2):
    for _ in range(warm):
        fn()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "dropin.json"))
    ap.add_argument("--n", type=int, default=3000)
    args = ap.parse_args()
    cv2.setNumThreads(os.cpu_count() or 1)
    rng = np.random.default_rng(0)
    n = args.n
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
        print(name, res[name], flush=True)

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
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
