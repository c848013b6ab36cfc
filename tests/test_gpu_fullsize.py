"""The hot path at BASELINE.json's FULL sizes.  The oracle is too slow to replay these whole, so the checks are
(a) the oracle on the part it can afford (the winner, a few pairs of the batch) and (b) size-independent properties:
additivity of the score over a split of the points, invariance of the winner under a split of the hypotheses (what
the multi-GPU shard relies on), idempotence (bitwise equal keys / models on a second run), pair independence."""
import numpy as np
import oracle as O
import pytest
from pygcransac import _cabi, synth

pytestmark = pytest.mark.gpu
PAIR_STRIDE = 0x9E3779B97F4A7C15
MASK = (1 << 64) - 1


def test_cfg2_f_4096_points_262144_samples(gcr_ctx, oracle):
    """configs[1] as benchmarked: one pair, N = 4096, 262144 7-point samples in ONE round -> the oracle's sequential round."""
    pts = synth.epipolar_pair(4096, 0.5, seed=1)[0]
    gcr_ctx.points_upload(pts)
    keys, best, evals = gcr_ctx.round(O.F, _cabi.MSAC, _cabi.UNIFORM, 0, 262144, seed=12345, thr=0.75)
    score, idx = gcr_ctx.key_decode(keys[0])
    r64, rc = gcr_ctx.round_results()
    want_score, want_idx, scored, want_model = oracle.round(O.F, O.MSAC, pts, 12345, 0, 262144, 0.75)
    assert idx // 3 == want_idx // 3                                  # the same minimal sample wins
    assert abs(r64[0] - want_score) <= 1e-9 * want_score              # FP64 decision score
    assert abs(score - want_score) <= 5e-5 * want_score               # FP32 ranking score
    assert rc[0] == oracle.score(O.F, O.MSAC, pts, want_model, 0.75)[1]
    assert np.abs(best[0] / np.linalg.norm(best[0]) - want_model / np.linalg.norm(want_model)).max() < 1e-9
    keys2, best2, _ = gcr_ctx.round(O.F, _cabi.MSAC, _cabi.UNIFORM, 0, 262144, seed=12345, thr=0.75)
    assert np.array_equal(keys, keys2) and np.array_equal(best, best2)  # idempotent, bit for bit
    assert gcr_ctx.counters()[1] == 0


def test_cfg4_batched_h_2048_pairs(gcr_ctx, oracle):
    """configs[3] per GPU: 2048 pairs x 2048 points x 1024 samples in one round; a few pairs against the oracle, every
    pair's key sane, a second run bitwise equal."""
    n_pairs, n, k = 2048, 2048, 1024
    rng = np.random.default_rng(4)
    base = [synth.homography_pair(n, 0.3, seed=s)[0] for s in range(1, 9)]
    pts = np.stack([base[p % 8] + (p // 8) * 1e-3 for p in range(n_pairs)])   # distinct pairs, cheap to generate
    gcr_ctx.points_upload(pts, n_pairs=n_pairs)
    keys, best, evals = gcr_ctx.round(O.H, _cabi.MSAC, _cabi.UNIFORM, 0, k, seed=77, thr=3.0)
    r64, rc = gcr_ctx.round_results()
    assert evals == n_pairs * n * k
    for p in (0, 1, 777, 2047):
        score, idx = gcr_ctx.key_decode(keys[p])
        samples = oracle.sample_uniform((77 + p * PAIR_STRIDE) & MASK, 0, k, n, 4)
        scores = np.full(k, -1.0)
        for i, s in enumerate(samples):
            if oracle.h_valid_sample(pts[p], s):
                mdl = oracle.solve_h4(pts[p], s)
                if mdl is not None:
                    scores[i] = oracle.score(O.H, O.MSAC, pts[p], mdl, 3.0)[0]
        want = int(np.argsort(-scores, kind="stable")[0])
        assert idx == want, (p, idx, want)
        assert abs(r64[p] - scores[want]) <= 1e-12 * scores[want]
        assert np.array_equal(best[p].reshape(3, 3), oracle.solve_h4(pts[p], samples[want]))
    # the tail of configs[3]: inliers of every pair's winner and one least-squares refit per pair, two launches
    d_best = gcr_ctx.to_device(best)
    d_idx, d_cnt = gcr_ctx.malloc(n_pairs * n * 4), gcr_ctx.malloc(n_pairs * 4)
    d_sc, d_refit, d_ok = gcr_ctx.malloc(n_pairs * 8), gcr_ctx.malloc(n_pairs * 72), gcr_ctx.malloc(n_pairs)
    try:
        gcr_ctx.final_refit_pairs(_cabi.H, d_best, 3.0, d_idx, d_cnt, d_sc, d_refit, d_ok)
        cnt = gcr_ctx.to_host(d_cnt, (n_pairs,), np.uint32)
        sc = gcr_ctx.to_host(d_sc, (n_pairs,), np.float64)
        refit = gcr_ctx.to_host(d_refit, (n_pairs, 3, 3), np.float64)
        ok = gcr_ctx.to_host(d_ok, (n_pairs,), np.uint8)
        idx_all = gcr_ctx.to_host(d_idx, (n_pairs, n), np.uint32)
    finally:
        for d in (d_best, d_idx, d_cnt, d_sc, d_refit, d_ok):
            gcr_ctx.free(d)
    assert ok.all() and np.array_equal(cnt, rc) and np.allclose(sc, r64, rtol=1e-12)
    for p in (0, 1, 777, 2047):
        s_p, c_p, inl = oracle.score(O.H, O.MSAC, pts[p], best[p].reshape(3, 3), 3.0, want_inliers=True)
        assert c_p == cnt[p] and np.array_equal(inl, idx_all[p, : cnt[p]])
        want = oracle.refit(O.H, pts[p], inl)
        g, w = refit[p] / np.linalg.norm(refit[p]), want / np.linalg.norm(want)
        assert min(np.abs(g - w).max(), np.abs(g + w).max()) < 1e-5
    dec = [gcr_ctx.key_decode(kk) for kk in keys]
    assert all(0 <= i < k for _, i in dec) and min(s for s, _ in dec) > 0.5 * 0.7 * n   # 70 % inliers: every pair found a model
    keys2, best2, _ = gcr_ctx.round(O.H, _cabi.MSAC, _cabi.UNIFORM, 0, k, seed=77, thr=3.0)
    assert np.array_equal(keys, keys2) and np.array_equal(best, best2)
    assert gcr_ctx.counters()[1] == 0


def test_cfg5_single_f_one_million_points(gcr_ctx, oracle):
    """configs[4] shape on one GPU: N = 2^20 correspondences, 65536-sample chunks (the sharded driver's chunk size).
    Properties: the winner of [0, K) is the better of the winners of [0, K/2) and [K/2, K) (first wins); the winner's
    FP64 score and inlier count are additive over a split of the points; the winner's score matches the oracle."""
    n = 1 << 20
    pts = synth.epipolar_pair(n, 0.5, seed=9)[0]
    gcr_ctx.points_upload(pts)
    K = 65536
    whole = gcr_ctx.round(O.F, _cabi.MSAC, _cabi.UNIFORM, 0, K, seed=3, thr=0.75)
    (s_w, i_w), (r_w, c_w) = gcr_ctx.key_decode(whole[0][0]), gcr_ctx.round_results()
    lo = gcr_ctx.round(O.F, _cabi.MSAC, _cabi.UNIFORM, 0, K // 2, seed=3, thr=0.75)
    (s_lo, i_lo), (r_lo, _) = gcr_ctx.key_decode(lo[0][0]), gcr_ctx.round_results()
    hi = gcr_ctx.round(O.F, _cabi.MSAC, _cabi.UNIFORM, K // 2, K // 2, seed=3, thr=0.75, index_base=3 * (K // 2))
    (s_hi, i_hi), (r_hi, _) = gcr_ctx.key_decode(hi[0][0]), gcr_ctx.round_results()
    want_i, want_r, want_m = (i_lo, r_lo[0], lo[1][0]) if r_lo[0] >= r_hi[0] else (i_hi, r_hi[0], hi[1][0])
    assert i_w == want_i and r_w[0] == want_r and np.array_equal(whole[1][0], want_m)
    assert int(whole[0][0]) == max(int(lo[0][0]), int(hi[0][0]))      # the packed keys order the same way (NCCL max)
    # additivity over the points: score / count of the winner on the two halves of the correspondences
    F = whole[1][0]
    full = oracle.score(O.F, O.MSAC, pts, F, 0.75)
    assert abs(r_w[0] - full[0]) <= 1e-9 * full[0] and c_w[0] == full[1]
    parts = []
    for sl in (slice(0, n // 2), slice(n // 2, n)):
        gcr_ctx.points_upload(np.ascontiguousarray(pts[sl]))
        s64, c64 = gcr_ctx.score_exact(_cabi.F, _cabi.MSAC, F, 0.75, want_inliers=False)
        parts.append((s64[0], c64[0]))
    assert abs(parts[0][0] + parts[1][0] - r_w[0]) <= 1e-9 * r_w[0]
    assert parts[0][1] + parts[1][1] == c_w[0]
