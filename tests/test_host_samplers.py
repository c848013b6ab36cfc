"""The driver's host-side sequential sampler (graph-cut-ransac_b200/host/gcr_host_samplers.h: the AR-Sampler sequence
drawn ahead of the GPU batches) against the reference's own AdaptiveReorderingSampler class compiled verbatim
(oracle/_ref; its rand() redirected to the same counter-based words).  CPU only: the sequence involves no kernel."""
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
