#!/usr/bin/env python
"""Per-CTA timeline of one K3 launch (gcr_set_score_trace): start / end spread, per-SM CTA counts, segment counts.
    python tools/k3_trace.py [f4k|h2k]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "graph-cut-ransac_b200"))
import ctypes as C
from pygcransac import _cabi, synth

wl = sys.argv[1] if len(sys.argv) > 1 else "f4k"
ctx = _cabi.Context(0)
if wl == "f4k":
    pts = synth.epipolar_pair(4096, 0.5, 1)[0][None]
    prob, samples, thr = _cabi.F, 262144, 0.75
else:
    pts = np.stack([synth.homography_pair(2048, 0.3, s + 1)[0] for s in range(512)])
    prob, samples, thr = _cabi.H, 1024, 3.0
ctx.points_upload(pts, n_pairs=len(pts))
for it in range(3):
    ctx.round(prob, _cabi.MSAC, _cabi.UNIFORM, it * samples, samples, seed=1, thr=thr)
d = ctx.malloc(4 * 8 * 2048)
ctx.lib.gcr_memset(ctx.h, C.c_void_p(d), 0, 4 * 8 * 2048)
ctx.lib.gcr_set_score_trace(ctx.h, C.c_void_p(d))
ctx.round(prob, _cabi.MSAC, _cabi.UNIFORM, 3 * samples, samples, seed=1, thr=thr)
ctx.lib.gcr_set_score_trace(ctx.h, None)
t = ctx.to_host(d, (2048, 4), np.uint64)
t = t[t[:, 2] > 0]
t0 = t[:, 1].min()
start, end = (t[:, 1] - t0) / 1e3, (t[:, 2] - t0) / 1e3
dur = end - start
print("CTAs %d  SMs %d  kernel span %.1f us" % (len(t), len(np.unique(t[:, 0])), end.max()))
print("start us: min %.1f p50 %.1f p90 %.1f max %.1f" % (start.min(), *np.percentile(start, [50, 90]), start.max()))
print("end   us: min %.1f p10 %.1f p50 %.1f p90 %.1f max %.1f" % (end.min(), *np.percentile(end, [10, 50, 90]), end.max()))
print("dur   us: min %.1f p10 %.1f p50 %.1f p90 %.1f max %.1f" % (dur.min(), *np.percentile(dur, [10, 50, 90]), dur.max()))
cnt = np.bincount(t[:, 0].astype(int))
print("CTAs per SM: histogram", np.bincount(cnt[cnt > 0]))
for k in np.unique(t[:, 3]):
    m = t[:, 3] == k
    print("segments %d: %d CTAs, mean dur %.1f us" % (k, m.sum(), dur[m].mean()))
# per-SM: when does its last CTA end vs its first
for q in (0, 1, 2, 3, 4):
    pass
order = np.argsort(end)
print("slowest 5 CTAs:", [(int(t[i, 0]), round(float(start[i]), 1), round(float(end[i]), 1), int(t[i, 3])) for i in order[-5:]])
print("fastest 5 CTAs:", [(int(t[i, 0]), round(float(start[i]), 1), round(float(end[i]), 1), int(t[i, 3])) for i in order[:5]])
