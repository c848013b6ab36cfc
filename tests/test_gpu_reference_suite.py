"""The reference's own test suite (tests/test_homography.py, test_fundamental_matrix.py, test_essential_matrix.py) on
its own images, through the same Python calls: the bundled 004520/004523 pair for H and F, the fountain pair with its
intrinsics for E -- SIFT tentatives extracted exactly as those tests do (committed fixtures: oracle/gen_cfg1_fixture.py,
oracle/gen_fountain_fixture.py).  Each test keeps the reference's inlier-count floor (> 250, > 380, > 450) and adds what
the reference cannot check: equality with its verbatim GCRANSAC<> driver on the same sample streams.  Differences from
the reference's calls: the first tests pass sampler=1 (PROSAC); test_reference_calls_with_their_own_samplers makes the
calls with the reference's own sampler ids (2 P-NAPSAC for H, 4 AR-Sampler for F and E); for neighborhood=1 the exact
radius graph stands where the reference asks FLANN for an approximate one."""
import os

import numpy as np
import oracle as O
import pytest
from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _up_to_scale(A, B):
    A = A / np.linalg.norm(A)
    B = B / np.linalg.norm(B)
    if (A * B).sum() < 0:
        B = -B
    return np.abs(A - B).max() / np.abs(A).max()


@pytest.fixture(scope="module")
def pyg():
    import pygcransac
    try:
        pygcransac._load_driver()
    except Exception as e:
        pytest.skip("host driver not available: %s" % e)
    return pygcransac


def test_find_homography(pyg):
    """tests/test_homography.py:39-85 (threshold 3.0, the defaults otherwise): > 250 inliers."""
    g = np.load(os.path.join(GOLDEN, "cfg1_matches.npz"))
    h1, w1, h2, w2 = (int(v) for v in g["size"])
    H, mask = pyg.findHomography(np.ascontiguousarray(g["pts"]), h1, w1, h2, w2, threshold=3.0, sampler=1, max_iters=5000,
                                 min_iters=50, probabilities=[])
    assert H is not None and H.shape == (3, 3) and mask.dtype == bool and mask.shape == (len(g["pts"]),)
    assert mask.sum() > 250


def test_find_fundamental_matrix(pyg, ref):
    """tests/test_fundamental_matrix.py:37-90 (threshold 0.75, max_iters 5000, min_iters 50): > 380 inliers; and the
    reference's default pipeline (PROSAC, BA refits, DEGENSAC, post-run BA) run verbatim on the same streams."""
    g = np.load(os.path.join(GOLDEN, "cfg1_matches.npz"))
    pts = np.ascontiguousarray(g["pts"])
    h1, w1, h2, w2 = (int(v) for v in g["size"])
    F, mask = pyg.findFundamentalMatrix(pts, h1, w1, h2, w2, threshold=0.75, sampler=1, max_iters=5000, min_iters=50,
                                        probabilities=[])
    st = pyg.last_statistics
    assert mask.sum() > 380
    cells = 20
    cs = np.array([w1 / cells, h1 / cells, w2 / cells, h2 / cells])
    want = ref.run(O.F, pts.astype(np.float64), cs, cells, 0.75, conf=0.99, spatial_weight=0.0, max_iters=5000, min_iters=50,
                   lo_number=50, seed=1, use_degensac=True, f_ba=3, sampler=1)
    assert st["iteration_number"] == want["stats"]["iterations"], (st, want["stats"])
    assert (mask != want["mask"]).sum() <= 2 and _up_to_scale(want["model"], F) < 1e-5


def test_find_essential_matrix(pyg, ref, oracle):
    """tests/test_essential_matrix.py:32-98 (fountain pair, threshold 0.75, max_iters 5000): > 450 inliers; the run equals
    the verbatim reference driver over the K-normalised points (grid neighbourhood on the pixel coordinates)."""
    g = np.load(os.path.join(GOLDEN, "fountain_matches.npz"))
    pts, K1, K2 = np.ascontiguousarray(g["pts"]), g["K1"], g["K2"]
    h1, w1, h2, w2 = (int(v) for v in g["size"])
    cells = 8
    E, mask = pyg.findEssentialMatrix(pts, K1, K2, h1, w1, h2, w2, probabilities=[], threshold=0.75, conf=0.99, min_iters=50,
                                      max_iters=5000, sampler=1, neighborhood=0, neighborhood_size=float(cells))
    st = pyg.last_statistics
    assert E is not None and mask.sum() > 450
    npts = oracle.normalize_correspondences(pts.astype(np.float64), K1, K2)
    cs = np.array([w1 / cells, h1 / cells, w2 / cells, h2 / cells])
    thr = 0.75 / (0.25 * (K1[0, 0] + K1[1, 1] + K2[0, 0] + K2[1, 1]))
    want = ref.run(O.E, npts, cs, cells, thr, conf=0.99, spatial_weight=0.1, max_iters=5000, min_iters=50, lo_number=50,
                   seed=1, grid_pts=pts.astype(np.float64), sampler=1, f_ba=4)
    assert st["iteration_number"] == want["stats"]["iterations"], (st, want["stats"])
    assert st["lo_samples"] == want["stats"]["lo_samples"]
    assert (mask != want["mask"]).sum() <= 2, (int(mask.sum()), int(want["mask"].sum()))
    assert _up_to_scale(want["model"], E) < 1e-5


def test_find_essential_matrix_default_arguments(pyg, ref, oracle):
    """The reference's call verbatim but for the sampler: every default of findEssentialMatrix, i.e. neighborhood=1 (radius
    20 px around each correspondence in the 4-D space) with spatial_coherence_weight 0.1."""
    g = np.load(os.path.join(GOLDEN, "fountain_matches.npz"))
    pts, K1, K2 = np.ascontiguousarray(g["pts"]), g["K1"], g["K2"]
    h1, w1, h2, w2 = (int(v) for v in g["size"])
    E, mask = pyg.findEssentialMatrix(pts, K1, K2, h1, w1, h2, w2, probabilities=[], threshold=0.75, conf=0.99, min_iters=50,
                                      max_iters=5000, sampler=1)
    st = pyg.last_statistics
    assert E is not None and mask.sum() > 450
    npts = oracle.normalize_correspondences(pts.astype(np.float64), K1, K2)
    thr = 0.75 / (0.25 * (K1[0, 0] + K1[1, 1] + K2[0, 0] + K2[1, 1]))
    want = ref.run(O.E, npts, np.ones(4), 20, thr, conf=0.99, spatial_weight=0.1, max_iters=5000, min_iters=50, lo_number=50,
                   seed=1, grid_pts=pts.astype(np.float64), sampler=1, radius=20.0, f_ba=4)
    assert st["iteration_number"] == want["stats"]["iterations"], (st, want["stats"])
    assert st["graph_cut_number"] == want["stats"]["graph_cuts"] and st["lo_samples"] == want["stats"]["lo_samples"]
    assert (mask != want["mask"]).sum() <= 2 and _up_to_scale(want["model"], E) < 1e-5


def test_reference_calls_with_their_own_samplers(pyg, ref, oracle):
    """The reference's tests verbatim, sampler ids included: P-NAPSAC (2) for H, AR-Sampler (4) with rank-based inlier
    probabilities for F and E (tests/test_*.py: get_probabilities).  The calls run the reference's own sampler sequences
    (host/gcr_host_samplers.h) and are compared with the verbatim reference driver over the reference's own
    ProgressiveNapsacSampler<4> / AdaptiveReorderingSampler classes."""
    g = np.load(os.path.join(GOLDEN, "cfg1_matches.npz"))
    pts = np.ascontiguousarray(g["pts"])
    h1, w1, h2, w2 = (int(v) for v in g["size"])
    prob = [1.0 - i / len(pts) for i in range(len(pts))]
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("error")                                       # every sampler id is native: nothing is substituted
        # tests/test_homography.py:75-85
        H, mask = pyg.findHomography(pts, h1, w1, h2, w2, threshold=3.0, sampler=2, max_iters=5000, min_iters=50, probabilities=[])
        st = pyg.last_statistics
        assert mask.sum() > 250                                              # tests/test_homography.py:85
        eps = np.finfo(np.float64).eps
        cs = np.array([(w1 + eps) / 4, (h1 + eps) / 4, (w2 + eps) / 4, (h2 + eps) / 4])
        want = ref.run(O.H, pts.astype(np.float64), cs, 4, 3.0, conf=0.99, spatial_weight=0.975, max_iters=5000, min_iters=50,
                       lo_number=50, seed=1, sampler=2, image_sizes=(w1, h1, w2, h2))
        assert st["iteration_number"] == want["stats"]["iterations"], (st, want["stats"])
        assert st["lo_samples"] == want["stats"]["lo_samples"]
        assert np.array_equal(mask, want["mask"]) and _up_to_scale(want["model"], H) < 1e-5
        F, mask = pyg.findFundamentalMatrix(pts, h1, w1, h2, w2, threshold=0.75, sampler=4, max_iters=5000, min_iters=50,
                                            probabilities=prob)
    st = pyg.last_statistics
    assert mask.sum() > 380                                                  # tests/test_fundamental_matrix.py:90
    cells = 20
    cs = np.array([w1 / cells, h1 / cells, w2 / cells, h2 / cells])
    want = ref.run(O.F, pts.astype(np.float64), cs, cells, 0.75, conf=0.99, spatial_weight=0.0, max_iters=5000, min_iters=50,
                   lo_number=50, seed=1, use_degensac=True, f_ba=3, sampler=4, probabilities=prob, sampler_variance=0.1)
    assert st["iteration_number"] == want["stats"]["iterations"], (st, want["stats"])
    assert (mask != want["mask"]).sum() <= 2 and _up_to_scale(want["model"], F) < 1e-5
    # H on the same pair with the AR-Sampler, linear pipeline: exact masks
    Hm, mask = pyg.findHomography(pts, h1, w1, h2, w2, prob, threshold=3.0, sampler=4, max_iters=5000, min_iters=50,
                                  neighborhood=0, neighborhood_size=4.0, sampler_variance=0.3, seed=2)
    st = pyg.last_statistics
    want = ref.run(O.H, pts.astype(np.float64), g["cell_sizes"], 4, 3.0, conf=0.99, spatial_weight=0.975, max_iters=5000,
                   min_iters=50, lo_number=50, seed=2, sampler=4, probabilities=prob, sampler_variance=0.3)
    assert st["iteration_number"] == want["stats"]["iterations"], (st, want["stats"])
    assert st["lo_samples"] == want["stats"]["lo_samples"]
    assert np.array_equal(mask, want["mask"]) and _up_to_scale(want["model"], Hm) < 1e-5 and mask.sum() > 250
    g = np.load(os.path.join(GOLDEN, "fountain_matches.npz"))
    pts, K1, K2 = np.ascontiguousarray(g["pts"]), g["K1"], g["K2"]
    h1, w1, h2, w2 = (int(v) for v in g["size"])
    prob = [1.0 - i / len(pts) for i in range(len(pts))]
    E, mask = pyg.findEssentialMatrix(pts, K1, K2, h1, w1, h2, w2, probabilities=prob, threshold=0.75, conf=0.99, min_iters=50,
                                      max_iters=5000, sampler=4)
    st = pyg.last_statistics
    assert mask.sum() > 450                                                  # tests/test_essential_matrix.py:98
    npts = oracle.normalize_correspondences(pts.astype(np.float64), K1, K2)
    thr = 0.75 / (0.25 * (K1[0, 0] + K1[1, 1] + K2[0, 0] + K2[1, 1]))
    want = ref.run(O.E, npts, np.ones(4), 20, thr, conf=0.99, spatial_weight=0.1, max_iters=5000, min_iters=50, lo_number=50,
                   seed=1, grid_pts=pts.astype(np.float64), sampler=4, radius=20.0, probabilities=prob, sampler_variance=0.1, f_ba=4)
    assert st["iteration_number"] == want["stats"]["iterations"], (st, want["stats"])
    assert st["lo_samples"] == want["stats"]["lo_samples"]
    assert (mask != want["mask"]).sum() <= 2 and _up_to_scale(want["model"], E) < 1e-5


def test_essential_with_progressive_napsac_and_importance_sampler(pyg, ref, oracle):
    """findEssentialMatrix with the two samplers the reference's E test does not use: P-NAPSAC builds its grid layers on the
    PIXEL coordinates while the run is on the K-normalised ones (py/gcransac_python.cpp:2650); the importance sampler
    draws from rank-based probabilities.  Both against the verbatim reference driver."""
    g = np.load(os.path.join(GOLDEN, "fountain_matches.npz"))
    pts, K1, K2 = np.ascontiguousarray(g["pts"]), g["K1"], g["K2"]
    h1, w1, h2, w2 = (int(v) for v in g["size"])
    prob = 1.0 - np.arange(len(pts)) / len(pts)
    npts = oracle.normalize_correspondences(pts.astype(np.float64), K1, K2)
    thr = 0.75 / (0.25 * (K1[0, 0] + K1[1, 1] + K2[0, 0] + K2[1, 1]))
    cells = 8
    cs = np.array([w1 / cells, h1 / cells, w2 / cells, h2 / cells])
    for sampler in (2, 3):
        E, mask = pyg.findEssentialMatrix(pts, K1, K2, h1, w1, h2, w2, probabilities=prob, threshold=0.75, conf=0.99, min_iters=50,
                                          max_iters=5000, sampler=sampler, neighborhood=0, neighborhood_size=float(cells), seed=2)
        st = pyg.last_statistics
        want = ref.run(O.E, npts, cs, cells, thr, conf=0.99, spatial_weight=0.1, max_iters=5000, min_iters=50, lo_number=50, seed=2,
                       grid_pts=pts.astype(np.float64), sampler=sampler, probabilities=prob, image_sizes=(w1, h1, w2, h2), f_ba=4)
        assert st["iteration_number"] == want["stats"]["iterations"], (sampler, st, want["stats"])
        assert st["lo_samples"] == want["stats"]["lo_samples"]
        assert (mask != want["mask"]).sum() <= 2 and _up_to_scale(want["model"], E) < 1e-5, sampler
        assert mask.sum() > 450
