"""Wall time of the non-minimal fit kernels on one B200: `sets` index sets of `size` points (an LO round) and one set of all
inliers (IRLS / final fit), F with and without the LM bundle adjustment, H, E.  python tools/refit_bench.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "graph-cut-ransac_b200"))
from pygcransac import _cabi, synth  # noqa: E402

ctx = _cabi.Context(0)
rng = np.random.default_rng(1)


def timed(problem, sets, ba_frame, reps=20):
    ctx.refit(problem, sets, ba_frame=ba_frame)
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        ctx.refit(problem, sets, ba_frame=ba_frame)
    ctx.sync()
    return 1e6 * (time.perf_counter() - t0) / reps


for name, problem, pts, inl, m in (("F", _cabi.F, *synth.epipolar_pair(4096, 0.5, 1)[::4], 7),
                                   ("H", _cabi.H, *synth.homography_pair(2048, 0.3, 1)[::2], 4),
                                   ("E", _cabi.E, None, None, 5)):
    if pts is None:
        px, Fm, Em, K, inl = synth.epipolar_pair(2048, 0.5, 2)
        pts = np.concatenate([(px[:, :2] - K[0, 2]) / K[0, 0], (px[:, 2:] - K[0, 2]) / K[0, 0]], axis=1)
    ctx.points_upload(pts)
    idx = np.flatnonzero(inl).astype(np.uint32)
    lo = [rng.choice(idx, 7 * m, replace=False) for _ in range(50)]
    for ba in ((0, _cabi.BA_HARTLEY) if name == "F" else (0,)):
        print("%s ba_frame %d: 50 sets x %d points %.1f us per call (host wall, upload + launch + read-back); one set of %d points %.1f us" % (
            name, ba, 7 * m, timed(problem, lo, ba), len(idx), timed(problem, [idx], ba)))
