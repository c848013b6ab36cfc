"""Dumps gcr_refit(E) outputs of fixed index sets (LO-sized sets, one large set, weighted sets, a 5-point set) to an .npz:
python tools/refit_e_dump.py out.npz -- for bitwise A/B comparisons of two builds (GCR_LIB_PATH)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "graph-cut-ransac_b200"))
from pygcransac import _cabi, synth  # noqa: E402

ctx = _cabi.Context(0)
rng = np.random.default_rng(3)
px, Fm, Em, K, inl = synth.epipolar_pair(2048, 0.5, 2)
pts = np.concatenate([(px[:, :2] - K[0, 2]) / K[0, 0], (px[:, 2:] - K[0, 2]) / K[0, 0]], axis=1)
ctx.points_upload(pts)
idx = np.flatnonzero(inl).astype(np.uint32)
out = {}
lo = [rng.choice(idx, 35, replace=False) for _ in range(200)] + [rng.choice(2048, 35, replace=False).astype(np.uint32) for _ in range(50)]
out["lo_m"], out["lo_ok"] = ctx.refit(_cabi.E, lo)
out["all_m"], out["all_ok"] = ctx.refit(_cabi.E, [idx])
out["five_m"], out["five_ok"] = ctx.refit(_cabi.E, [rng.choice(idx, 5, replace=False) for _ in range(64)])
w = rng.random(2048)
out["w_m"], out["w_ok"] = ctx.refit(_cabi.E, lo[:64], weights=w, weight_mode=_cabi.WEIGHT_BY_POINT)
np.savez(sys.argv[1], **out)
print("dumped", {k: v.shape for k, v in out.items()})
