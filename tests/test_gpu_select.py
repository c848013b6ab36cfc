"""K4 parity: the winner of gcr_round is the model the reference's running best would keep -- strict '<' on the FP64
score (GCRANSAC.h:440, scoring_function.h:63-66), first model among equals -- and it is decided in FP64 everywhere:
inside a round (every slot within the FP32 margin is re-scored, no bucketing), across the chunks of a problem and across
GPUs (gcr_winner records).  Also the block-cache regression of the host wrappers (round-1 ADVICE)."""
import threading

import numpy as np
import oracle as O
import pytest
from pygcransac import _cabi, synth

pytestmark = pytest.mark.gpu


def _brute_force_winner(ctx, prob, scoring, m, mm, count, seed, thr, first=0):
    """FP64 score of EVERY valid model of the round (gcr_sample -> gcr_solve_minimal -> gcr_score_exact): the (score,
    lowest index) argmax, plus all scores."""
    samples = ctx.sample(_cabi.UNIFORM, m, first, count, seed)
    models, valid = ctx.solve_minimal(prob, samples)
    models, valid = models[0].reshape(count * mm, -1), valid[0].reshape(-1)
    slots = np.nonzero(valid)[0]
    scores = np.full(count * mm, -1.0)
    for c0 in range(0, len(slots), 8192):
        sl = slots[c0:c0 + 8192]
        s64, _ = ctx.score_exact(prob, scoring, models[sl], thr, want_inliers=False)
        scores[sl] = s64
    best = int(np.argmax(scores))            # argmax returns the FIRST maximum: the reference's first-wins rule
    return best, scores, models


@pytest.mark.parametrize("seed", [12345, 7, 99])
def test_round_winner_is_fp64_argmax_cfg2_full_k(gcr_ctx, seed):
    """configs[1] at the benchmarked K = 262144 samples (786432 slots, ~86 k valid models): the round's winner is the
    brute-force FP64 argmax over ALL of them (the round-1 design kept one FP32 candidate per slot % 64 class and could
    lose a near-tied true winner), and the selection's premise holds: FP32 score error < margin / 2."""
    pts = synth.epipolar_pair(4096, 0.5, seed=1)[0]
    gcr_ctx.points_upload(pts)
    gcr_ctx.counters(reset=True)
    K = 262144
    keys, best, _ = gcr_ctx.round(_cabi.F, _cabi.MSAC, _cabi.UNIFORM, 0, K, seed=seed, thr=0.75)
    r64, rc = gcr_ctx.round_results()
    rescored, max_err = gcr_ctx.select_stats()
    want, scores, models = _brute_force_winner(gcr_ctx, _cabi.F, _cabi.MSAC, 7, 3, K, seed, 0.75)
    _, idx = gcr_ctx.key_decode(keys[0])
    assert idx == want, (idx, want, scores[idx], scores[want])
    assert abs(r64[0] - scores[want]) <= 1e-12 * scores[want]
    assert np.array_equal(best[0], models[want])
    # every model within the margin of the top was re-scored (the candidate set is complete) and the margin is safe
    near = int((scores >= scores[want] * (1 - 1e-4)).sum())
    assert rescored >= near >= 1
    assert 0 < max_err < 1e-4                      # margin 2e-4 (MSAC) needs FP32 error < 1e-4
    print("K4: %d FP64 re-scores, %d models within 1e-4 of the winner, max FP32 score error %.2e" % (rescored, near, max_err))


def test_round_winner_is_fp64_argmax_magsac_e(gcr_ctx):
    """configs[2]'s scoring (5-point + MAGSAC++, 85 % outliers): same statement, MAGSAC margin 1e-3."""
    px = synth.epipolar_pair(2048, 0.85, seed=3)[0]
    pts = (px - synth.SIZE / 2) / synth.FOCAL
    thr = 0.75 / synth.FOCAL
    gcr_ctx.points_upload(pts)
    gcr_ctx.counters(reset=True)
    K = 16384
    keys, best, _ = gcr_ctx.round(_cabi.E, _cabi.MAGSAC, _cabi.UNIFORM, 0, K, seed=5, thr=thr)
    r64, _ = gcr_ctx.round_results()
    rescored, max_err = gcr_ctx.select_stats()
    want, scores, models = _brute_force_winner(gcr_ctx, _cabi.E, _cabi.MAGSAC, 5, 10, K, 5, thr)
    _, idx = gcr_ctx.key_decode(keys[0])
    assert idx == want and abs(r64[0] - scores[want]) <= 1e-12 * scores[want]
    assert np.array_equal(best[0], models[want])
    assert rescored >= 1 and max_err < 5e-4


def test_batched_pairs_winner_is_fp64_argmax(gcr_ctx):
    """Many pairs (one selection CTA per pair): every pair's winner against the brute-force FP64 argmax of that pair."""
    n_pairs, n, K = 320, 512, 256
    pts = np.stack([synth.homography_pair(n, 0.3, seed=s + 1)[0] for s in range(n_pairs)])
    gcr_ctx.points_upload(pts, n_pairs=n_pairs)
    keys, best, _ = gcr_ctx.round(_cabi.H, _cabi.MSAC, _cabi.UNIFORM, 0, K, seed=11, thr=3.0)
    r64, rc = gcr_ctx.round_results()
    samples = gcr_ctx.sample(_cabi.UNIFORM, 4, 0, K, 11)
    models, valid = gcr_ctx.solve_minimal(_cabi.H, samples)
    for p in range(0, n_pairs, 37):
        one = _cabi.Context(0)
        try:
            one.points_upload(pts[p])
            sl = np.nonzero(valid[p].reshape(-1))[0]
            s64, c64 = one.score_exact(_cabi.H, _cabi.MSAC, models[p].reshape(K, 9)[sl], 3.0, want_inliers=False)
        finally:
            one.close()
        want = sl[int(np.argmax(s64))]
        _, idx = gcr_ctx.key_decode(keys[p])
        assert idx == want and abs(r64[p] - s64.max()) <= 1e-12 * s64.max() and rc[p] == c64[int(np.argmax(s64))]


def test_chunked_accumulation_equals_one_round(gcr_ctx):
    """A problem walked in chunks keeps its running best on the device as an FP64 record: the result is bit-identical to
    the unchunked pass, for every chunk size, and indices are 64-bit (index_base beyond 2^32 does not wrap)."""
    pts = synth.epipolar_pair(8192, 0.5, seed=4)[0]
    gcr_ctx.points_upload(pts)
    total = 24000
    whole, _ = gcr_ctx.round_sharded(_cabi.F, _cabi.MSAC, _cabi.UNIFORM, 0, total, 31, 0.75, chunk_samples=total)
    assert whole["valid"][0] == 1 and whole["score"][0] > 1000
    keys, best, _ = gcr_ctx.round(_cabi.F, _cabi.MSAC, _cabi.UNIFORM, 0, total, seed=31, thr=0.75)
    r64, rc = gcr_ctx.round_results()
    assert whole["index"][0] == gcr_ctx.key_decode(keys[0])[1] and whole["score"][0] == r64[0] and whole["count"][0] == rc[0]
    assert np.array_equal(whole["model"][0][:9], best[0]) and not whole["model"][0][9:].any()
    for chunk in (1000, 4096, 23999):
        got, _ = gcr_ctx.round_sharded(_cabi.F, _cabi.MSAC, _cabi.UNIFORM, 0, total, 31, 0.75, chunk_samples=chunk)
        assert got.tobytes() == whole.tobytes(), chunk
    # 64-bit global indices
    d = gcr_ctx.malloc(128)
    try:
        gcr_ctx.round_winners_device(_cabi.F, _cabi.MSAC, _cabi.UNIFORM, 0, 4096, 31, 0.75, 0, 5 * 2 ** 32, False, d)
        far = gcr_ctx.to_host(d, (1,), _cabi.WINNER_DTYPE)
        gcr_ctx.round_winners_device(_cabi.F, _cabi.MSAC, _cabi.UNIFORM, 0, 4096, 31, 0.75, 0, 0, False, d)
        near = gcr_ctx.to_host(d, (1,), _cabi.WINNER_DTYPE)
    finally:
        gcr_ctx.free(d)
    assert far["index"][0] == near["index"][0] + 5 * 2 ** 32 and far["score"][0] == near["score"][0]


def test_winner_reduce_kernel_matches_host_rule(gcr_ctx):
    """k_winners_reduce (what every rank runs after the all-gather) against the host statement of the rule, on the
    records of tests/test_sharding_gloo.py: FP64 decides where FP32 ties, 64-bit indices, ranks without a model."""
    import test_sharding_gloo as T
    gcr_ctx.points_upload(np.zeros((5, 8, 4)) + np.arange(8)[None, :, None], n_pairs=5)
    allrec = np.stack([np.stack(T.LOCAL[0]), np.stack(T.LOCAL[1])])
    d_all, d_out = gcr_ctx.to_device(allrec.view(np.uint8)), gcr_ctx.malloc(5 * 128)
    try:
        gcr_ctx.check(gcr_ctx.lib.gcr_winners_reduce(gcr_ctx.h, d_all, 2, d_out))
        got = gcr_ctx.to_host(d_out, (5,), _cabi.WINNER_DTYPE)
    finally:
        gcr_ctx.free(d_all)
        gcr_ctx.free(d_out)
    assert got.tobytes() == np.stack(T.EXPECT).tobytes()


def _device_count():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_round_equals_single_gpu_bit_for_bit(world):
    """SURVEY 8(e) row 2 on hardware: `world` GPUs (one context + one host thread each, NCCL all-gather of the winner
    records over NVLink) return, on EVERY rank, the record one GPU returns -- score bits, 64-bit index, inlier count
    and model bits.  Skips itself when the box has fewer GPUs."""
    if _device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    n, total, seed, thr = 65536, 50000, 17, 0.75
    pts = synth.epipolar_pair(n, 0.5, seed=6)[0]
    single = _cabi.Context(0)
    try:
        single.points_upload(pts)
        want, _ = single.round_sharded(_cabi.F, _cabi.MSAC, _cabi.UNIFORM, 0, total, seed, thr, chunk_samples=8192)
    finally:
        single.close()
    uid = _cabi.comm_unique_id()
    got, errs = [None] * world, []

    def rank_main(r):
        try:
            c = _cabi.Context(r)
            try:
                c.points_upload(pts)
                c.comm_init(uid, r, world)
                for rep in range(2):                                    # the communicator is reusable
                    got[r], _ = c.round_sharded(_cabi.F, _cabi.MSAC, _cabi.UNIFORM, 0, total, seed, thr, chunk_samples=8192)
                c.comm_destroy()
            finally:
                c.close()
        except Exception as e:                                          # noqa: BLE001
            errs.append((r, repr(e)))

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not errs, errs
    assert want["valid"][0] == 1
    for r in range(world):
        assert got[r] is not None and got[r].tobytes() == want.tobytes(), (r, got[r], want)


def test_host_wrappers_keep_the_block_cache_consistent():
    """gcr_score_host / gcr_score_exact_host return their temporaries with gcr_free (round 1 paired gcr_malloc with a raw
    cudaFree: dangling cache entries, a double free at gcr_ctx_destroy): 1000 calls on one context leave the cache
    bounded with nothing handed out, and the results never change."""
    ctx = _cabi.Context(0)
    try:
        pts, Hm, _ = synth.homography_pair(600, 0.3, seed=2)
        ctx.points_upload(pts)
        rng = np.random.default_rng(0)
        models = synth.perturbed_models(Hm, 24, rng)
        s0, c0 = ctx.score(_cabi.H, _cabi.MSAC, models, 2.0, flags=_cabi.SCORE_EXACT_COUNTS)
        e0 = ctx.score_exact(_cabi.H, _cabi.MSAC, models[:3], 2.0)
        base = ctx.cache_stats()
        assert base["in_use"] == 0
        for i in range(1000):
            s, c = ctx.score(_cabi.H, _cabi.MSAC, models, 2.0, flags=_cabi.SCORE_EXACT_COUNTS)
            assert np.array_equal(s, s0) and np.array_equal(c, c0)
            if i % 10 == 0:
                e = ctx.score_exact(_cabi.H, _cabi.MSAC, models[:3], 2.0)
                assert np.array_equal(e[0], e0[0]) and np.array_equal(e[1], e0[1])
                # interleave allocations of other sizes, as a driver would
                d = ctx.malloc(1000 + 64 * (i % 7))
                ctx.free(d)
        st = ctx.cache_stats()
        assert st["in_use"] == 0 and st["blocks"] <= base["blocks"] + 8 and st["bytes"] <= base["bytes"] + (1 << 20), (base, st)
    finally:
        ctx.close()                                                     # a double free would fail here (compute-sanitizer clean)


def test_float32_upload_equals_widened_float64_upload(gcr_ctx):
    """gcr_points_upload_f32 (the float32 arrays the reference's tests pass; widened on the device) gives bit-identical
    rounds to uploading the same values as float64 (pybind's host-side widening in the reference), batched pairs included."""
    n_pairs, n = 6, 777
    pts = np.stack([synth.homography_pair(n, 0.3, seed=s + 40)[0] for s in range(n_pairs)]).astype(np.float32)
    gcr_ctx.points_upload(pts.astype(np.float64), n_pairs=n_pairs)
    want = gcr_ctx.round(_cabi.H, _cabi.MSAC, _cabi.UNIFORM, 0, 512, seed=3, thr=3.0)
    want_r = gcr_ctx.round_results()
    gcr_ctx.points_upload_f32(pts, n_pairs=n_pairs)
    got = gcr_ctx.round(_cabi.H, _cabi.MSAC, _cabi.UNIFORM, 0, 512, seed=3, thr=3.0)
    got_r = gcr_ctx.round_results()
    assert np.array_equal(want[0], got[0]) and np.array_equal(want[1], got[1])
    assert np.array_equal(want_r[0], got_r[0]) and np.array_equal(want_r[1], got_r[1])
