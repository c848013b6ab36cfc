"""Times the drop-in template surface: the reference's own GCRANSAC<Estimator, Graph, ScoringFunction> (GCRANSAC.h compiled
verbatim, oracle/_ref) with gcr::CudaMSACScoringFunction (host/gcr_scoring_function.h: one model per getScore call, one
H2D + launch + D2H + stream synchronise each) as its _ScoringFunction argument, against the stock MSACScoringFunction, on
BASELINE configs[0] (N = 762) and on synthetic homography problems up to N = 2^20.  A measurement tool (it runs the
reference driver, so it lives beside the tests, not in the product):  python tools/adapter_bench.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "graph-cut-ransac_b200"))
import oracle as O  # noqa: E402
from pygcransac import synth  # noqa: E402


def run(ref, pts, sizes, thr, weight, kind, seed, max_iters):
    h1, w1, h2, w2 = sizes
    eps = np.finfo(np.float64).eps
    cs = np.array([(w1 + eps) / 4.0, (h1 + eps) / 4.0, (w2 + eps) / 4.0, (h2 + eps) / 4.0])
    t0 = time.perf_counter()
    out = ref.run(O.H, pts, cs, 4, thr, conf=0.99, spatial_weight=weight, max_iters=max_iters, min_iters=50, lo_number=50, seed=seed,
                  scoring_kind=kind)
    return time.perf_counter() - t0, out


def main():
    cpu, gpu = O.Ref(), O.RefCuda()
    g = np.load(os.path.join(ROOT, "tests", "golden", "cfg1_matches.npz"))
    cases = [("cfg1 (N=762, bundled pair)", np.ascontiguousarray(g["pts"], dtype=np.float64), tuple(int(v) for v in g["size"]), 3.0, 0.975, 5000, 20)]
    for n, reps in ((4096, 10), (65536, 5), (1 << 20, 2)):
        # spatial weight 0 at the large sizes: the reference's labeling keeps an N x N used-edges matrix (GCRANSAC.h:963)
        cases.append(("synthetic H, N=%d, 50%% outliers" % n, synth.homography_pair(n, 0.5, seed=3)[0], (1024,) * 4, 3.0,
                      0.975 if n <= 4096 else 0.0, 500, reps))
    print("%-40s %12s %12s %9s %s" % ("case", "stock ms", "adapter ms", "speed-up", "iterations / equal"))
    for name, pts, sizes, thr, weight, max_iters, reps in cases:
        run(gpu, pts, sizes, thr, weight, 1, 1, max_iters)                       # warm-up: context, point upload
        tc = tg = 0.0
        same = True
        for s in range(reps):
            a, oc = run(cpu, pts, sizes, thr, weight, 0, s + 1, max_iters)
            b, og = run(gpu, pts, sizes, thr, weight, 1, s + 1, max_iters)
            tc, tg = tc + a, tg + b
            same &= oc["stats"]["iterations"] == og["stats"]["iterations"] and int((oc["mask"] != og["mask"]).sum()) <= 2
        print("%-40s %12.2f %12.2f %9.2f %d / %s" % (name, 1e3 * tc / reps, 1e3 * tg / reps, tc / tg, oc["stats"]["iterations"], same), flush=True)


if __name__ == "__main__":
    main()
