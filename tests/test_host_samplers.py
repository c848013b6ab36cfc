"""The driver's host-side sequential samplers (graph-cut-ransac_b200/host/gcr_host_samplers.h: the AR-Sampler and
Progressive NAPSAC sequences drawn ahead of the GPU batches) against the reference's own AdaptiveReorderingSampler and
ProgressiveNapsacSampler<4> classes compiled verbatim (oracle/_ref; their random numbers redirected to the same
counter-based words).  CPU only: the sequences involve no kernel."""
import ctypes as C
import os

import numpy as np
import pytest
from conftest import ROOT


@pytest.fixture(scope="module")
def host_ar():
    path = os.path.join(ROOT, "graph-cut-ransac_b200", "host", "_ref", "libgcr_driver.so")
    if not os.path.exists(path):
        pytest.skip("host driver not built (needs the reference tree for its BK max-flow)")
    C.CDLL(os.path.join(ROOT, "graph-cut-ransac_b200", "libgcr_b200.so"), mode=C.RTLD_GLOBAL)
    f = C.CDLL(path).gcr_host_sample_adaptive_reordering
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_uint64, C.c_uint64, C.c_int, C.c_longlong, C.c_void_p]

    def draw(prob, m, variance, seed, first, count, repeat_at=-1):
        prob = np.ascontiguousarray(prob, dtype=np.float64)
        out = np.empty((count, m), dtype=np.uint32)
        assert f(prob.ctypes.data, len(prob), m, variance, seed, first, count, repeat_at, out.ctypes.data) == 0
        return out
    return draw


def test_adaptive_reordering_sequence_matches_reference_class(host_ar, ref):
    rng = np.random.default_rng(1)
    for n, m, variance, repeat_at in ((762, 4, 0.1, -1), (4096, 7, 0.1, -1), (2048, 5, 0.9765, -1), (50, 4, 0.1, 17),
                                      (9, 7, 0.1, -1), (300, 7, 0.1, 0)):
        noisy = rng.random(n)
        if n > 100:
            noisy[::7] = 1.0                      # probability 1 -> 1 - 1e-6 (adaptive_reordering_sampler.h:86-87)
            noisy[3] = 0.0                        # probability 0 -> NaN Beta parameters, clamped to 0.999 on its first update
        ranks = 1.0 - np.arange(n) / n            # the reference's tests: get_probabilities
        for prob in (noisy, ranks):
            want = ref.sample_adaptive_reordering(prob, m, variance, 5, 3000, repeat_at)
            assert np.array_equal(host_ar(prob, m, variance, 5, 0, 3000, repeat_at), want), (n, m, variance, repeat_at)
            # any later range without the earlier ones having been asked for
            assert np.array_equal(host_ar(prob, m, variance, 5, 2500, 100, repeat_at), want[2500:2600])
            if repeat_at < 0:                                   # m distinct points per sample (each point is in the heap once;
                assert all(len(set(r)) == m for r in want)      # a repeated update pushes its points twice, as in the reference)
    # the seed only drives the noise: same first sample, different sequence later
    a, b = host_ar(ranks, 7, 0.1, 1, 0, 400), host_ar(ranks, 7, 0.1, 2, 0, 400)
    assert np.array_equal(a[0], b[0]) and not np.array_equal(a, b)


def test_progressive_napsac_sequence_matches_reference_class(ref):
    """samplers/progressive_napsac_sampler.h through its three regimes: local sampling in the grid layers (hit counters,
    growing neighbourhoods, layer fall-through), PROSAC over all points after 0.5 n samples, uniform after n."""
    from conftest import GOLDEN
    from pygcransac import synth
    path = os.path.join(ROOT, "graph-cut-ransac_b200", "host", "_ref", "libgcr_driver.so")
    if not os.path.exists(path):
        pytest.skip("host driver not built (needs the reference tree for its BK max-flow)")
    C.CDLL(os.path.join(ROOT, "graph-cut-ransac_b200", "libgcr_b200.so"), mode=C.RTLD_GLOBAL)
    f = C.CDLL(path).gcr_host_sample_progressive_napsac
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p]
    g = np.load(os.path.join(GOLDEN, "cfg1_matches.npz"))
    real = np.ascontiguousarray(g["pts"].astype(np.float64))
    h1, w1, h2, w2 = (float(v) for v in g["size"])
    square = (1024.0,) * 4
    cases = [(real, (w1, h1, w2, h2), 4), (real, (h1, w1, h2, w1), 7), (real[:40].copy(), (w1, h1, w2, h2), 5),
             (synth.homography_pair(3000, 0.5, seed=3)[0], square, 4), (synth.epipolar_pair(500, 0.5, seed=3)[0], square, 5)]
    for pts, sizes, m in cases:
        n = len(pts)
        count = min(3 * n, 6000)
        want = ref.sample_pnapsac(pts, m, sizes, 7, count)
        got = np.empty((count, m), dtype=np.uint32)
        sz = np.array(sizes, dtype=np.float64)
        assert f(pts.ctypes.data, n, m, sz.ctypes.data, 7, 0, count, got.ctypes.data) == 0
        assert np.array_equal(got, want), (n, m)
        sub = np.empty((50, m), dtype=np.uint32)
        assert f(pts.ctypes.data, n, m, sz.ctypes.data, 7, count - 50, 50, sub.ctypes.data) == 0
        assert np.array_equal(sub, want[-50:])
        assert (want[: n // 2, -1] == np.arange(n // 2)).all()      # local phase: the k-th best point is the centre
