"""Generates tests/golden/*.npz from the reference's own code compiled verbatim (oracle/_ref/libgcr_ref.so).

Run here (needs /root/reference):  python oracle/gen_golden.py
The fixtures pin the self-contained oracle (oracle/libgcr_oracle.so) on machines where /root/reference is absent.
What is "reference" in each fixture is stated in the fixture's `provenance` string: the scoring loops, gamma table,
PROSAC growth function / subset schedule, grid graph and termination formula are the reference's own compiled code;
residual functions come from the restated estimators (the reference's estimator headers need Eigen, absent here).
"""
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "graph-cut-ransac_b200"))
import oracle as O  # noqa: E402
from pygcransac import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


IMPORTANCE_CASES = ((762, 4), (4096, 7), (2048, 5), (2, 4), (1, 4))


def importance_probabilities(n):
    """Skewed inlier probabilities with a fifth of the points at exactly zero (never drawn)."""
    r = np.random.default_rng(1000 + n)
    p = r.random(n) ** 3
    if n > 10:
        p[r.integers(0, n, n // 5)] = 0.0
    return p


def importance(ref):
    # ---- importance sampler (importance_sampler.h verbatim: its own std::discrete_distribution object) ----
    imp = {}
    for n, m in IMPORTANCE_CASES:
        imp["samples_%d_%d" % (n, m)] = ref.sample_importance(importance_probabilities(n), m, 11, 0, 3000)
    imp["provenance"] = np.array("ImportanceSampler::multinomial_distribution (built by the reference's constructor, "
                                 "compiled verbatim) drawn with the Philox word stream (seed 11, stream 3)")
    np.savez_compressed(os.path.join(OUT, "importance.npz"), **imp)


def main():
    ref = O.Ref()
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "importance":   # only this fixture (the others stay byte-identical)
        importance(ref)
        return
    importance(ref)
    rng = np.random.default_rng(2024)

    # ---- scoring (scoring_function.h verbatim) ----
    cases = {}
    pts_h, Hm, _ = synth.homography_pair(700, 0.3, seed=1)
    pts_f, Fm, Em, Kc, _ = synth.epipolar_pair(600, 0.5, seed=2)
    pts_p, Pm, _ = synth.pnp_scene(500, 0.4, seed=3)
    for name, prob, pts, model, thr in (("h", O.H, pts_h, Hm, 2.0), ("f", O.F, pts_f, Fm, 0.75),
                                        ("pnp", O.PNP, pts_p, Pm, 2.0 / synth.FOCAL)):
        models = synth.perturbed_models(model, 24, rng)
        for sc, scn in ((O.MSAC, "msac"), (O.MAGSAC, "magsac")):
            scores, counts, lists = [], [], []
            for mdl in models:
                s, c, idx = ref.score(prob, sc, pts, mdl, thr, want_inliers=True)
                scores.append(s)
                counts.append(c)
                lists.append(idx)
            cases["%s_%s_scores" % (name, scn)] = np.array(scores)
            cases["%s_%s_counts" % (name, scn)] = np.array(counts, dtype=np.uint64)
            cases["%s_%s_inl0" % (name, scn)] = lists[0]
            cases["%s_%s_inl5" % (name, scn)] = lists[5]
        cases["%s_pts" % name] = pts
        cases["%s_models" % name] = models
        cases["%s_thr" % name] = np.array(thr)
    cases["provenance"] = np.array("scoring_function.h MSAC/MAGSAC compiled verbatim from /root/reference; residuals restated")
    np.savez_compressed(os.path.join(OUT, "scoring.npz"), **cases)

    # ---- gamma table (gamma_values.cpp:2012) ----
    tab = ref.gamma_table()
    np.savez_compressed(os.path.join(OUT, "gamma.npz"), first=tab[:5001], every64=tab[::64], size=np.array(len(tab)),
                        provenance=np.array("stored_gamma_values[] of gamma_values.cpp compiled verbatim"))

    # ---- PROSAC (prosac_sampler.h verbatim) ----
    pro = {}
    for n, m in ((762, 4), (4096, 7), (2048, 5), (64, 4), (12, 7)):
        growth, sizes, last = ref.prosac(n, m, 100000, 4000)
        pro["growth_%d_%d" % (n, m)] = growth
        pro["sizes_%d_%d" % (n, m)] = sizes
        pro["last_%d_%d" % (n, m)] = last
    pro["provenance"] = np.array("ProsacSampler growth_function / subset_size per call, compiled verbatim")
    np.savez_compressed(os.path.join(OUT, "prosac.npz"), **pro)

    # ---- grid (grid_neighborhood_graph.h verbatim) ----
    grid = {}
    pts4 = np.ascontiguousarray(np.random.default_rng(7).uniform(0, 1024, (3000, 4)))
    grid["pts"] = pts4
    for Cn in (3, 4, 8):
        cs = np.array([(1024 + 1e-9) / Cn] * 4)
        g = ref.grid(pts4, cs, Cn)
        for key in ("keys", "sorted_idx", "cell_offsets"):
            grid["%s_%d" % (key, Cn)] = g[key]
        grid["meta_%d" % Cn] = np.array([g["n_cells"], g["neighbor_number"]], dtype=np.uint64)
        grid["cell_sizes_%d" % Cn] = cs
    grid["provenance"] = np.array("GridNeighborhoodGraph<4>::initialize compiled verbatim")
    np.savez_compressed(os.path.join(OUT, "grid.npz"), **grid)

    # ---- termination (GCRANSAC.h:158-172 verbatim) ----
    rows = []
    for inl, n, m, conf in ((10, 762, 4, 0.99), (400, 762, 4, 0.99), (762, 762, 4, 0.99), (1, 4096, 7, 0.95),
                            (2000, 4096, 7, 0.99), (300, 2048, 5, 0.99), (7, 4096, 7, 0.99), (0, 100, 4, 0.99)):
        lp = math.log(1 - conf)
        rows.append((inl, n, m, lp, ref.iteration_number(inl, n, m, lp)))
    np.savez_compressed(os.path.join(OUT, "iteration.npz"), inl=np.array([r[0] for r in rows], dtype=np.uint64),
                        n=np.array([r[1] for r in rows], dtype=np.uint64), m=np.array([r[2] for r in rows]),
                        logp=np.array([r[3] for r in rows]), out=np.array([r[4] for r in rows], dtype=np.uint64),
                        provenance=np.array("GCRANSAC<>::getIterationNumber compiled verbatim"))
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
