"""Does K2 (FP64 7-point solver) run beside K3 (FP32 scoring) on one GPU?  Two contexts / streams, wall clock over many
iterations: sequential vs concurrent, for several K3 CTAs-per-SM caps.  python tools/overlap_probe.py"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "graph-cut-ransac_b200"))
from pygcransac import _cabi, synth  # noqa: E402

K, N = 262144, 4096
pts = synth.epipolar_pair(N, 0.5, seed=1)[0]
os.environ["GCR_STREAM_PRIORITY"] = os.environ.get("PRIO_A", "1")
a = _cabi.Context(0)
os.environ["GCR_STREAM_PRIORITY"] = os.environ.get("PRIO_B", "0")
b = _cabi.Context(0)
print("priority a", os.environ.get("PRIO_A", "1"), "b", os.environ.get("PRIO_B", "0"), flush=True)
CARVE = int(os.environ.get("CARVEOUT", "-1"))
for c in (a, b):
    c.points_upload(pts)
    c.set_tuning(_cabi.TUNE_SMEM_CARVEOUT, CARVE)
print("carveout", CARVE, flush=True)
vp = C.c_void_p


def prep(c):
    ds = c.malloc(K * 7 * 4)
    c.check(c.lib.gcr_sample(c.h, _cabi.UNIFORM, 7, 0, K, 12345, vp(ds)))
    dm, dv = c.malloc(K * 3 * 72), c.malloc(K * 3)
    dsc = c.malloc(K * 3 * 4)
    return ds, dm, dv, dsc


A, B = prep(a), prep(b)


def solve(c, bufs):
    c.check(c.lib.gcr_solve_minimal(c.h, _cabi.F, vp(bufs[0]), K, vp(bufs[1]), vp(bufs[2])))


def score(c, bufs):
    c.check(c.lib.gcr_score(c.h, _cabi.F, _cabi.MSAC, vp(bufs[1]), vp(bufs[2]), K * 3, 0.75, 0, vp(bufs[3]), None))


solve(a, A)
solve(b, B)
a.sync(); b.sync()
IT = 30
for cap in (0, 3, 2, 1):
    a.set_tuning(_cabi.TUNE_K3_CTAS_PER_SM, cap)
    for mode in ("k3 only", "k2 only", "sequential", "concurrent"):
        for rep in range(2):
            a.sync(); b.sync()
            t0 = time.perf_counter()
            for it in range(IT):
                if mode in ("k3 only", "sequential", "concurrent"):
                    score(a, A)
                if mode == "k2 only" or mode == "sequential":
                    solve(a, B if mode == "k2 only" else A)
                if mode == "concurrent":
                    solve(b, B)
            a.sync(); b.sync()
            dt = (time.perf_counter() - t0) / IT
        print("K3 CTAs/SM cap %d  %-11s %.1f us per iteration" % (cap, mode, dt * 1e6), flush=True)
