"""K1 / K2 / K4 parity through the C ABI: sample indices bit-exact with the oracle's counter-based stream,
H4 models bit-exact (same FP64 operation order), argmax key semantics."""
import numpy as np
import oracle as O
import pytest
from pygcransac import _cabi, synth

pytestmark = pytest.mark.gpu
PAIR_STRIDE = 0x9E3779B97F4A7C15
MASK = (1 << 64) - 1


def test_uniform_samples_bit_exact(gcr_ctx, oracle):
    pts = np.stack([synth.homography_pair(333, 0.3, seed=s)[0] for s in (1, 2)])
    gcr_ctx.points_upload(pts, n_pairs=2)
    for m in (3, 4, 5, 7):
        got = gcr_ctx.sample(_cabi.UNIFORM, m, 1000, 3000, seed=77)
        for p in range(2):
            want = oracle.sample_uniform((77 + p * PAIR_STRIDE) & MASK, 1000, 3000, 333, m)
            assert np.array_equal(got[p], want), (m, p)


def test_uniform_tiny_pool(gcr_ctx, oracle):
    pts = synth.homography_pair(7, 0.0, seed=3)[0]
    gcr_ctx.points_upload(pts)
    got = gcr_ctx.sample(_cabi.UNIFORM, 7, 0, 500, seed=1)[0]
    assert np.array_equal(got, oracle.sample_uniform(1, 0, 500, 7, 7))
    assert all(sorted(r) == list(range(7)) for r in got)
    with pytest.raises(_cabi.GcrError):
        gcr_ctx.sample(_cabi.UNIFORM, 8, 0, 10, seed=1)  # uniform_sampler.h:87-88: sample larger than the pool


def test_prosac_samples_bit_exact(gcr_ctx, oracle):
    for n, m in ((762, 4), (4096, 7), (64, 5)):
        pts = synth.homography_pair(n, 0.3, seed=5)[0]
        gcr_ctx.points_upload(pts)
        growth = oracle.prosac_growth(n, m)
        got = gcr_ctx.sample(_cabi.PROSAC, m, 1, 5000, seed=9)[0]
        assert np.array_equal(got, oracle.sample_prosac(9, 1, 5000, n, m, growth)), (n, m)
        # beyond the convergence point PROSAC degenerates to uniform draws over all points (prosac_sampler.h:224-229)
        got = gcr_ctx.sample(_cabi.PROSAC, m, 99990, 30, seed=9)[0]
        assert np.array_equal(got, oracle.sample_prosac(9, 99990, 30, n, m, growth)), (n, m)


def test_importance_samples_bit_exact(gcr_ctx, oracle):
    """GCR_IMPORTANCE (samplers/importance_sampler.h, NG-RANSAC): the device draws select the indices the reference's own
    std::discrete_distribution object selects from the same word stream (fixture from oracle/_ref) and the oracle's."""
    import os
    from conftest import GOLDEN
    sys_path_gen = os.path.join(os.path.dirname(GOLDEN), "..", "oracle", "gen_golden.py")
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_golden", sys_path_gen)
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    g = np.load(os.path.join(GOLDEN, "importance.npz"))
    for n, m in gen.IMPORTANCE_CASES:
        if n < m:
            continue                                   # gcr_sample refuses m > n like the uniform sampler
        p = gen.importance_probabilities(n)
        pts = np.stack([synth.homography_pair(n, 0.3, seed=s)[0] for s in (1, 2)])
        gcr_ctx.points_upload(pts, n_pairs=2)
        with pytest.raises(_cabi.GcrError):
            gcr_ctx.sample(_cabi.IMPORTANCE, m, 0, 10, seed=11)         # no probabilities for these points yet
        gcr_ctx.set_point_probabilities(p)
        got = gcr_ctx.sample(_cabi.IMPORTANCE, m, 0, 3000, seed=11)
        assert np.array_equal(got[0], g["samples_%d_%d" % (n, m)]), (n, m)
        cdf = oracle.importance_cdf(p)
        assert np.array_equal(got[1], oracle.sample_importance((11 + PAIR_STRIDE) & MASK, 0, 3000, m, cdf)), (n, m)
        got = gcr_ctx.sample(_cabi.IMPORTANCE, m, 123456789012, 100, seed=5)[0]     # 64-bit sample counters
        assert np.array_equal(got, oracle.sample_importance(5, 123456789012, 100, m, cdf))
    with pytest.raises(_cabi.GcrError):
        gcr_ctx.set_point_probabilities(np.ones(5))                     # importance_sampler.h:72-73: one per point
    with pytest.raises(_cabi.GcrError):
        gcr_ctx.set_point_probabilities(np.zeros(gcr_ctx.n))


def test_round_with_importance_sampler(gcr_ctx, oracle):
    """gcr_round draws the importance samples inside the solver kernel: the winner is the sample gcr_sample returns for
    that index, solved and scored like any other."""
    pts, Hm, inl = synth.homography_pair(1500, 0.4, seed=21)
    prob = np.where(inl, 0.9, 0.1) * np.random.default_rng(0).uniform(0.5, 1.0, len(pts))
    gcr_ctx.points_upload(pts)
    gcr_ctx.set_point_probabilities(prob)
    keys, best, evals = gcr_ctx.round(_cabi.H, _cabi.MSAC, _cabi.IMPORTANCE, 0, 2048, seed=3, thr=3.0)
    score, idx = gcr_ctx.key_decode(keys[0])
    sample = oracle.sample_importance(3, idx, 1, 4, oracle.importance_cdf(prob))[0]
    want = oracle.solve_h4(pts, sample)
    assert want is not None and np.array_equal(best[0].reshape(3, 3), want)
    want_score, _ = oracle.score(O.H, O.MSAC, pts, want, 3.0)
    assert abs(score - want_score) <= 1e-5 * want_score
    # guided sampling pays: far more all-inlier samples than uniform draws give
    s = gcr_ctx.sample(_cabi.IMPORTANCE, 4, 0, 2048, seed=3)[0]
    u = gcr_ctx.sample(_cabi.UNIFORM, 4, 0, 2048, seed=3)[0]
    assert inl[s].all(axis=1).mean() > 3 * inl[u].all(axis=1).mean()


def test_pool_sampler(gcr_ctx, oracle):
    pool = np.arange(10, 500, 3, dtype=np.uint32)
    got = gcr_ctx.sample_from_pool(pool, 28, 0, 50, seed=4, stream=2)
    want = oracle.sample_uniform(4, 0, 50, len(pool), 28, stream=2)
    assert np.array_equal(got, pool[want])


def test_h4_models_bit_exact(gcr_ctx, oracle):
    pts, Hm, _ = synth.homography_pair(500, 0.3, seed=13)
    gcr_ctx.points_upload(pts)
    samples = oracle.sample_uniform(3, 0, 4000, 500, 4)
    models, valid = gcr_ctx.solve_minimal(O.H, samples)
    models, valid = models[0, :, 0], valid[0, :, 0]
    n_valid = 0
    for s, mdl, v in zip(samples, models, valid):
        ok = oracle.h_valid_sample(pts, s)
        want = oracle.solve_h4(pts, s) if ok else None
        assert bool(v) == (want is not None)
        if want is not None:
            n_valid += 1
            assert np.array_equal(mdl.reshape(3, 3), want)  # bit-exact FP64
    assert 500 < n_valid < 4000  # both branches exercised


def test_f7_models_match_oracle(gcr_ctx, oracle):
    """K2 7-point: same null-space construction and closed-form cubic as the oracle -> same models (slot = root index);
    CUDA's cbrt / acos / cos differ from libm's by a few ulp, which near-multiple roots amplify: compared at 1e-6."""
    pts, Fm, Em, Kc, _ = synth.epipolar_pair(800, 0.4, seed=14)
    gcr_ctx.points_upload(pts)
    samples = oracle.sample_uniform(8, 0, 3000, 800, 7)
    models, valid = gcr_ctx.solve_minimal(O.F, samples)
    models, valid = models[0], valid[0]
    n_models = 0
    for s, mdl, v in zip(samples, models, valid):
        want = oracle.estimate_f7(pts, s)                      # 7-pt + oriented-constraint filter, ascending roots
        got = mdl[v.astype(bool)].reshape(-1, 3, 3)
        assert len(got) == len(want), (s, len(got), len(want))
        for g, w in zip(got, want):
            assert np.abs(g - w).max() <= 1e-6 * np.abs(w).max()
        n_models += len(want)
    assert 300 < n_models < 9000


def test_e5_models_match_oracle(gcr_ctx, oracle):
    """K2 5-point (Nister): statement-for-statement the oracle's arithmetic -> the same models, sample by sample."""
    pts, Fm, Em, Kc, _ = synth.epipolar_pair(600, 0.5, seed=15)
    npts = oracle.normalize_correspondences(pts, Kc, Kc)
    gcr_ctx.points_upload(npts)
    samples = oracle.sample_uniform(9, 0, 2000, 600, 5)
    models, valid = gcr_ctx.solve_minimal(O.E, samples)
    models, valid = models[0], valid[0]
    n_models = 0
    for s, mdl, v in zip(samples, models, valid):
        want = oracle.estimate_e5(npts, s)                     # Nister + oriented-constraint filter, ascending roots
        got = mdl[v.astype(bool)].reshape(-1, 3, 3)
        assert len(got) == len(want), (s, len(got), len(want))
        for g, w in zip(got, want):
            assert np.abs(g - w).max() <= 1e-9 * np.abs(w).max(), (s, np.abs(g - w).max())
        n_models += len(want)
    assert 500 < n_models < 20000


def test_round_e_magsac_matches_cpu_round(gcr_ctx, oracle):
    """BASELINE configs[2] shape: 5-point + MAGSAC++ on K-normalised correspondences, threshold / focal length."""
    pts, Fm, Em, Kc, _ = synth.epipolar_pair(800, 0.6, seed=16)
    npts = oracle.normalize_correspondences(pts, Kc, Kc)
    thr = 0.75 / synth.FOCAL
    gcr_ctx.points_upload(npts)
    keys, best, evals = gcr_ctx.round(O.E, _cabi.MAGSAC, _cabi.UNIFORM, 0, 1500, seed=11, thr=thr,
                                      flags=_cabi.SCORE_EXACT_COUNTS)
    score, idx = gcr_ctx.key_decode(keys[0])
    want_score, want_idx, scored, want_model = oracle.round(O.E, O.MAGSAC, npts, 11, 0, 1500, thr)
    assert idx // 10 == want_idx // 10
    r64, rc = gcr_ctx.round_results()
    assert abs(r64[0] - want_score) <= 1e-9 * want_score
    err = np.abs(best[0] / np.linalg.norm(best[0]) - want_model / np.linalg.norm(want_model)).max()
    assert err < 1e-9
    assert gcr_ctx.counters()[1] == 0


def test_p3p_models_match_oracle(gcr_ctx, oracle):
    """K2 P3P (Lambda-Twist): same statements as the oracle; cbrt / acos / cos are the only libm calls."""
    pts, Pm, _ = synth.pnp_scene(500, 0.4, seed=17)
    gcr_ctx.points_upload(pts)
    samples = oracle.sample_uniform(10, 0, 3000, 500, 3)
    models, valid = gcr_ctx.solve_minimal(O.PNP, samples)
    models, valid = models[0], valid[0]
    n_models = 0
    for s, mdl, v in zip(samples, models, valid):
        want = oracle.solve_p3p(pts, s)
        got = mdl[v.astype(bool)].reshape(-1, 3, 4)
        assert len(got) == len(want), (s, len(got), len(want))
        for g, w in zip(got, want):
            assert np.abs(g - w).max() <= 1e-7 * max(1.0, np.abs(w).max()), (s, np.abs(g - w).max())
        n_models += len(want)
    assert 3000 < n_models < 12000


def test_round_pnp_matches_cpu_round(gcr_ctx, oracle):
    """P3P + reprojection-error MSAC (find6DPose's inner loop, py/gcransac_python.cpp:1026-1272)."""
    pts, Pm, inl = synth.pnp_scene(900, 0.5, seed=18)
    thr = 2.0 / synth.FOCAL
    gcr_ctx.points_upload(pts)
    keys, best, evals = gcr_ctx.round(O.PNP, _cabi.MSAC, _cabi.UNIFORM, 0, 2000, seed=13, thr=thr,
                                      flags=_cabi.SCORE_EXACT_COUNTS)
    score, idx = gcr_ctx.key_decode(keys[0])
    want_score, want_idx, scored, want_model = oracle.round(O.PNP, O.MSAC, pts, 13, 0, 2000, thr)
    assert idx // 4 == want_idx // 4
    r64, rc = gcr_ctx.round_results()
    assert abs(r64[0] - want_score) <= 1e-6 * want_score
    assert np.abs(best[0].reshape(3, 4) - want_model.reshape(3, 4)).max() < 1e-7
    assert np.abs(best[0].reshape(3, 4) - Pm).max() < 5e-3      # and it is the true pose
    assert gcr_ctx.counters()[1] == 0


def test_argmax_first_wins(gcr_ctx):
    pts = np.stack([synth.homography_pair(64, 0.3, seed=s)[0] for s in (1, 2)])
    gcr_ctx.points_upload(pts, n_pairs=2)
    scores = np.zeros((2, 5000), dtype=np.float32)
    scores[0, [700, 4000]] = 12.5   # tie: the lower index must win (GCRANSAC.h:440 strict '<')
    scores[1, :] = -1.0
    scores[1, 4999] = 0.0
    keys = gcr_ctx.argmax(scores, index_base=100)
    assert gcr_ctx.key_decode(keys[0]) == (12.5, 800)
    assert gcr_ctx.key_decode(keys[1]) == (0.0, 5099)


def test_round_h_matches_stagewise_oracle(gcr_ctx, oracle):
    pts, Hm, _ = synth.homography_pair(1500, 0.3, seed=17)
    gcr_ctx.points_upload(pts)
    keys, best, evals = gcr_ctx.round(O.H, _cabi.MSAC, _cabi.UNIFORM, 0, 2048, seed=21, thr=3.0,
                                      flags=_cabi.SCORE_EXACT_COUNTS)
    score, idx = gcr_ctx.key_decode(keys[0])
    samples = oracle.sample_uniform(21, 0, 2048, 1500, 4)
    scores = np.full(2048, -1.0)
    for i, s in enumerate(samples):
        if oracle.h_valid_sample(pts, s):
            mdl = oracle.solve_h4(pts, s)
            if mdl is not None:
                scores[i] = oracle.score(O.H, O.MSAC, pts, mdl, 3.0)[0]
    order = np.argsort(-scores, kind="stable")
    # the near-best candidates of the FP32 ranking are re-scored in FP64: the winner is the FP64 winner (first wins)
    assert idx == order[0]
    r64, rc = gcr_ctx.round_results()
    assert abs(r64[0] - scores[idx]) <= 1e-12 * scores[idx]
    assert rc[0] == oracle.score(O.H, O.MSAC, pts, oracle.solve_h4(pts, samples[idx]), 3.0)[1]
    assert score == np.float32(r64[0])
    assert np.array_equal(best[0].reshape(3, 3), oracle.solve_h4(pts, samples[idx]))
    assert evals == 2048 * 1500


def test_round_f_matches_cpu_round(gcr_ctx, oracle):
    """Whole fixed-K pass (K1 -> K2 -> K3 -> K4) for the 7-point problem against the oracle's sequential round."""
    pts, Fm, Em, Kc, _ = synth.epipolar_pair(1200, 0.5, seed=19)
    gcr_ctx.points_upload(pts)
    keys, best, evals = gcr_ctx.round(O.F, _cabi.MSAC, _cabi.UNIFORM, 0, 3000, seed=5, thr=0.75,
                                      flags=_cabi.SCORE_EXACT_COUNTS)
    score, idx = gcr_ctx.key_decode(keys[0])
    want_score, want_idx, scored, want_model = oracle.round(O.F, O.MSAC, pts, 5, 0, 3000, 0.75)
    # oracle index = sample * 3 + position among the KEPT models; ours = sample * 3 + root slot: compare the sample
    assert idx // 3 == want_idx // 3
    r64, rc = gcr_ctx.round_results()
    assert abs(r64[0] - want_score) <= 1e-9 * want_score
    err = np.abs(best[0] / np.linalg.norm(best[0]) - want_model / np.linalg.norm(want_model)).max()
    assert err < 1e-9
    assert gcr_ctx.counters()[1] == 0


def test_plane_parallax_models_bit_exact(gcr_ctx, oracle):
    """DEGENSAC's inner minimal solver (solver_fundamental_matrix_plane_and_parallax.h:109-160 + the oriented
    constraint): same statement order as the oracle, only + - * -> bit-exact."""
    pts, Fm, inl, on_plane = synth.epipolar_pair_planar(600, 0.3, 0.6, seed=23)
    Hm = oracle.refit(O.H, pts, np.flatnonzero(on_plane)[:50].astype(np.uint32))
    gcr_ctx.points_upload(pts)
    samples = oracle.sample_uniform(5, 0, 2000, 600, 2)
    models, valid = gcr_ctx.solve_plane_parallax(samples, Hm)
    n_valid = 0
    for s, mdl, v in zip(samples, models[0], valid[0]):
        want = oracle.solve_f_plane_parallax(pts, s, Hm, oriented=True)
        assert bool(v) == (want is not None)
        if want is not None:
            n_valid += 1
            assert np.array_equal(mdl.reshape(3, 3), want)
    assert 200 < n_valid < 2000
