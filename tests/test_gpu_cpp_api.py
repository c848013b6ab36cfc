"""The reference's C++ driver entry points (include/gcransac_python.h: findHomography_, findFundamentalMatrix_,
findEssentialMatrix_) as graph-cut-ransac_b200/host/gcr_gcransac_python.h provides them, called from a C++ translation
unit with std::vector arguments (tests/cpp/cpp_api_shim.cpp): same results as the Python entry points (which are pinned
against the verbatim reference driver elsewhere), the reference's output conventions and failure behaviour."""
import ctypes as C
import os

import numpy as np
import pytest
from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def shim():
    import pygcransac
    try:
        pygcransac._load_driver()
    except Exception as e:
        pytest.skip("host driver not available: %s" % e)
    path = os.path.join(ROOT, "graph-cut-ransac_b200", "host", "_ref", "libgcr_cpp_api_shim.so")
    if not os.path.exists(path):
        pytest.skip("C++ API shim not built")
    f = C.CDLL(path).shim_find
    f.restype = C.c_int
    dp = C.c_void_p
    f.argtypes = [C.c_int, dp, C.c_int, dp, C.c_int, dp, dp] + [C.c_int] * 4 + [C.c_double] * 3 + [C.c_int] * 4 + \
                 [C.c_double, C.c_int, C.c_double, C.c_int, C.c_uint64, dp, dp, C.POINTER(C.c_int), C.POINTER(C.c_int)]

    def call(problem, pts, prob, sizes, weight, thr, sampler, neighborhood, nsize, seed, K1=None, K2=None, max_iters=5000,
             use_sp=True, variance=0.1):
        pts = np.ascontiguousarray(pts, dtype=np.float64)
        prob = np.ascontiguousarray(prob, dtype=np.float64)
        model, mask = np.full(12, np.nan), np.zeros(len(pts) + 1, dtype=np.uint8)
        ml, kl = C.c_int(), C.c_int()
        K1 = np.ascontiguousarray(K1 if K1 is not None else np.eye(3), dtype=np.float64)
        K2 = np.ascontiguousarray(K2 if K2 is not None else np.eye(3), dtype=np.float64)
        count = f(problem, pts.ctypes.data, len(pts), prob.ctypes.data, len(prob), K1.ctypes.data, K2.ctypes.data, *sizes,
                  weight, thr, 0.99, max_iters, 50, sampler, neighborhood, nsize, 1 if use_sp else 0, variance, 50, seed,
                  model.ctypes.data, mask.ctypes.data, C.byref(ml), C.byref(kl))
        return count, model[: ml.value], mask[: kl.value].astype(bool)
    return pygcransac, call


def test_cpp_entry_points_equal_python_entry_points(shim):
    pyg, call = shim
    from pygcransac import _cabi
    g = np.load(os.path.join(GOLDEN, "cfg1_matches.npz"))
    pts = np.ascontiguousarray(g["pts"], dtype=np.float64)
    sizes = tuple(int(v) for v in g["size"])
    prob = 1.0 - np.arange(len(pts)) / len(pts)
    # H: the reference's test call (P-NAPSAC, defaults: grid 4, weight 0.975, space partitioning on)
    count, H, mask = call(_cabi.H, pts, [], sizes, 0.975, 3.0, 2, 0, 4.0, 1)
    Hp, maskp = pyg.findHomography(pts, *sizes, threshold=3.0, sampler=2, max_iters=5000, min_iters=50, probabilities=[])
    assert count == mask.sum() == maskp.sum() > 250 and np.array_equal(mask, maskp) and np.array_equal(H.reshape(3, 3), Hp)
    # use_space_partitioning forces the grid neighbourhood (py/gcransac_python.cpp:4282-4286)
    count2, H2, mask2 = call(_cabi.H, pts, [], sizes, 0.975, 3.0, 2, 1, 4.0, 1)
    assert count2 == count and np.array_equal(H2, H)
    # F: the reference's test call (AR-Sampler, radius neighbourhood 20, default estimator)
    count, F, mask = call(_cabi.F, pts, prob, sizes, 0.0, 0.75, 4, 1, 20.0, 1)
    Fp, maskp = pyg.findFundamentalMatrix(pts, *sizes, threshold=0.75, sampler=4, max_iters=5000, min_iters=50, probabilities=prob)
    assert count == mask.sum() > 380 and np.array_equal(mask, maskp) and np.array_equal(F.reshape(3, 3), Fp)
    # E: the reference's test call on the fountain pair
    g = np.load(os.path.join(GOLDEN, "fountain_matches.npz"))
    pts, K1, K2 = np.ascontiguousarray(g["pts"], dtype=np.float64), g["K1"], g["K2"]
    sizes = tuple(int(v) for v in g["size"])
    prob = 1.0 - np.arange(len(pts)) / len(pts)
    count, E, mask = call(_cabi.E, pts, prob, sizes, 0.1, 0.75, 4, 1, 20.0, 1, K1=K1, K2=K2)
    Ep, maskp = pyg.findEssentialMatrix(pts, K1, K2, *sizes, probabilities=prob, threshold=0.75, conf=0.99, min_iters=50,
                                        max_iters=5000, sampler=4)
    assert count == mask.sum() > 450 and np.array_equal(mask, maskp) and np.array_equal(E.reshape(3, 3), Ep)


def test_cpp_entry_points_failure_conventions(shim):
    """fprintf(stderr) + return 0, outputs untouched (py/gcransac_python.cpp:4360-4368 and alike)."""
    pyg, call = shim
    from pygcransac import _cabi, synth
    pts = synth.homography_pair(500, 0.3, seed=2)[0]
    sizes = (1024,) * 4
    count, H, mask = call(_cabi.H, pts, [], sizes, 0.975, 3.0, 9, 0, 4.0, 1)           # unknown sampler
    assert count == 0 and len(H) == 0 and len(mask) == 0
    count, H, mask = call(_cabi.H, pts, [0.5] * 10, sizes, 0.975, 3.0, 4, 0, 4.0, 1)   # AR-Sampler: one probability per point
    assert count == 0 and len(H) == 0
    bad = np.full(len(pts), 0.5)
    bad[7] = np.nan
    for sampler in (3, 4):                                                             # probabilities that are no probabilities
        count, H, mask = call(_cabi.H, pts, bad, sizes, 0.975, 3.0, sampler, 0, 4.0, 1)
        assert count == 0 and len(H) == 0
        count, H, mask = call(_cabi.H, pts, np.zeros(len(pts)), sizes, 0.975, 3.0, sampler, 0, 4.0, 1)
        assert count == 0 and len(H) == 0
    count, H, mask = call(_cabi.H, pts, [], (0, 1024, 1024, 1024), 0.975, 3.0, 2, 0, 4.0, 1)   # P-NAPSAC without an image size
    assert count == 0
    count, H, mask = call(_cabi.H, pts, [], sizes, 0.975, 3.0, 1, 5, 4.0, 1, use_sp=False)     # unknown neighbourhood
    assert count == 0
    count, H, mask = call(_cabi.H, pts[:3], [], sizes, 0.975, 3.0, 1, 0, 4.0, 1)       # fewer points than a minimal sample
    assert count == 0
    rng = np.random.default_rng(0)                                                     # no model: count 0, all-false mask
    count, H, mask = call(_cabi.H, rng.uniform(0, 1000, (50, 4)), [], (1000,) * 4, 0.975, 1e-6, 1, 0, 4.0, 1, max_iters=200)
    assert count == 0 and len(mask) == 50 and not mask.any()


def test_cpp_find6dpose_equals_python(shim):
    """find6DPose_ (include/gcransac_python.h:177-211) called from the C++ translation unit == pygcransac.find6DPose."""
    pyg, call = shim
    from pygcransac import _cabi, synth
    pts, P0, inl = synth.pnp_scene(600, 0.5, seed=9)
    count, P, mask = call(_cabi.PNP, pts, [], (0, 0, 0, 0), 0.975, 2.0 / synth.FOCAL, 0, 0, 0.6, 3)
    Pp, maskp = pyg.find6DPose(pts, [], threshold=2.0 / synth.FOCAL, sampler=0, neighborhood_size=0.6, max_iters=5000, seed=3)
    assert count == mask.sum() == maskp.sum() > 0.9 * inl.sum() and np.array_equal(mask, maskp)
    assert P.shape == (12,) and np.array_equal(P.reshape(3, 4), Pp) and np.abs(Pp - P0).max() < 5e-3
    assert call(_cabi.PNP, pts, [], (0, 0, 0, 0), 0.975, 2.0 / synth.FOCAL, 0, 1, 0.6, 3)[0] == 0      # unknown neighbourhood id
