"""The self-contained oracle against the golden vectors generated from the reference's own compiled code
(oracle/gen_golden.py -> tests/golden/).  CPU only."""
import os

import numpy as np
import oracle as O
from conftest import GOLDEN


def _load(name):
    return np.load(os.path.join(GOLDEN, name))


def test_scoring_matches_reference_fixtures(oracle):
    g = _load("scoring.npz")
    for name, prob in (("h", O.H), ("f", O.F), ("pnp", O.PNP)):
        pts, models, thr = g["%s_pts" % name], g["%s_models" % name], float(g["%s_thr" % name])
        for sc, scn in ((O.MSAC, "msac"), (O.MAGSAC, "magsac")):
            want_s, want_c = g["%s_%s_scores" % (name, scn)], g["%s_%s_counts" % (name, scn)]
            got_s, got_c = oracle.score_batch(prob, sc, pts, models, thr)
            assert np.array_equal(got_c, want_c), (name, scn)
            if sc == O.MSAC:
                assert np.array_equal(got_s, want_s), (name, scn)  # bit-exact
            else:  # LUT printed with 13-16 digits in the reference; closed form here
                assert np.allclose(got_s, want_s, rtol=1e-12, atol=1e-12), (name, scn)
            for j in (0, 5):
                _, _, idx = oracle.score(prob, sc, pts, models[j], thr, want_inliers=True)
                assert np.array_equal(idx, g["%s_%s_inl%d" % (name, scn, j)])
            assert want_c.max() > 50  # the fixture is not vacuous


def test_gamma_table(oracle):
    g = _load("gamma.npz")
    mine = np.array([oracle.gamma_lut(k) for k in range(5001)])
    assert np.abs(mine - g["first"]).max() < 5e-14
    every = np.array([oracle.gamma_lut(k) for k in range(0, int(g["size"]), 64)])
    assert np.abs(every - g["every64"]).max() < 5e-14


def test_prosac_growth_and_schedule(oracle):
    g = _load("prosac.npz")
    for n, m in ((762, 4), (4096, 7), (2048, 5), (64, 4), (12, 7)):
        growth = oracle.prosac_growth(n, m)
        assert np.array_equal(growth, g["growth_%d_%d" % (n, m)])
        sizes = np.array([oracle.prosac_subset_size(growth, n, m, k) for k in range(1, 4001)], dtype=np.uint32)
        assert np.array_equal(sizes, g["sizes_%d_%d" % (n, m)])
        samples = oracle.sample_prosac(5, 1, 4000, n, m, growth)
        assert np.array_equal(samples[:, -1], g["last_%d_%d" % (n, m)])
        # free indices come from [0, n_k - 2] and are distinct (k >= 2)
        free = samples[1:, :-1]
        assert (free < (sizes[1:, None] - 1)).all()
        assert all(len(set(r)) == m - 1 for r in free[:500])


def test_prosac_first_draw_quirk(oracle):
    # prosac_sampler.h:171-172 + :232: the first sample may repeat index m-1
    n, m = 762, 4
    growth = oracle.prosac_growth(n, m)
    firsts = np.array([oracle.sample_prosac(seed, 1, 1, n, m, growth)[0] for seed in range(200)])
    assert (firsts[:, -1] == m - 1).all() and (firsts[:, :-1] < m).all()
    assert any((r[:-1] == m - 1).any() for r in firsts)


def test_importance_sampler_matches_reference_fixture(oracle):
    """samplers/importance_sampler.h: the restated table + draw against the reference's own std::discrete_distribution
    object (fixture from oracle/_ref), bit-exact; draws are with replacement and never hit a zero-probability point."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_golden", os.path.join(os.path.dirname(GOLDEN), "..", "oracle", "gen_golden.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    g = _load("importance.npz")
    for n, m in gen.IMPORTANCE_CASES:
        p = gen.importance_probabilities(n)
        cdf = oracle.importance_cdf(p)
        assert cdf[-1] == 1.0 and np.all(np.diff(cdf) >= 0)
        got = oracle.sample_importance(11, 0, 3000, m, cdf)
        assert np.array_equal(got, g["samples_%d_%d" % (n, m)]), (n, m)
        assert np.all(p[got] > 0)
        # counter based: any sub-range reproduces
        assert np.array_equal(oracle.sample_importance(11, 1000, 10, m, cdf), got[1000:1010])
    p = gen.importance_probabilities(762)
    s = oracle.sample_importance(3, 0, 50000, 4, oracle.importance_cdf(p))
    assert any(len(set(r)) < 4 for r in s)                       # with replacement (importance_sampler.h:134-135)
    freq = np.bincount(s.ravel(), minlength=762) / s.size
    assert np.abs(freq - p / p.sum()).max() < 2e-3


def test_grid(oracle):
    g = _load("grid.npz")
    for Cn in (3, 4, 8):
        got = oracle.grid_build(g["pts"], g["cell_sizes_%d" % Cn], Cn)
        for key in ("keys", "sorted_idx", "cell_offsets"):
            assert np.array_equal(got[key], g["%s_%d" % (key, Cn)]), (Cn, key)
        assert [got["n_cells"], got["neighbor_number"]] == list(g["meta_%d" % Cn])


def test_iteration_number(oracle):
    g = _load("iteration.npz")
    for inl, n, m, lp, want in zip(g["inl"], g["n"], g["m"], g["logp"], g["out"]):
        assert oracle.iteration_number(int(inl), int(n), int(m), float(lp)) == int(want)


def test_uniform_sampler_properties(oracle):
    s = oracle.sample_uniform(123, 0, 5000, 100, 7)
    assert s.max() < 100 and all(len(set(r)) == 7 for r in s)
    # counter based: any sub-range reproduces
    assert np.array_equal(oracle.sample_uniform(123, 1000, 10, 100, 7), s[1000:1010])
    # roughly uniform
    hist = np.bincount(s.ravel(), minlength=100)
    assert hist.min() > 250 and hist.max() < 450
    # Philox4x32-10 known-answer (Random123 kat_vectors: zero counter / zero key)
    assert [hex(v) for v in oracle.philox([0, 0, 0, 0], [0, 0])] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    assert [hex(v) for v in oracle.philox([0xffffffff] * 4, [0xffffffff] * 2)] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]


def test_h4_solver_and_refit(oracle):
    from pygcransac import synth
    pts, Hm, inl = synth.homography_pair(300, 0.0, seed=11, sigma=0.0)
    rng = np.random.default_rng(0)
    for _ in range(20):
        s = rng.choice(300, 4, replace=False)
        Hs = oracle.solve_h4(pts, s)
        assert Hs is not None and np.allclose(Hs / Hs[2, 2], Hm, rtol=1e-6, atol=1e-6)
    idx = rng.choice(300, 40, replace=False)
    Hr = oracle.refit(O.H, pts, idx)
    assert np.allclose(Hr / Hr[2, 2], Hm, rtol=1e-8, atol=1e-8)
    # against numpy least squares on noisy data (normal equations of the same normalised system)
    ptsn, Hm2, _ = synth.homography_pair(300, 0.0, seed=12, sigma=0.5)
    Hr = oracle.refit(O.H, ptsn, idx)
    res = np.array([oracle.sqres(O.H, ptsn[i], Hr) for i in idx])
    assert res.mean() < 1.0
    assert oracle.h_valid_model(Hm) and not oracle.h_valid_model(np.diag([1.0, 1.0, 1e-3]))
