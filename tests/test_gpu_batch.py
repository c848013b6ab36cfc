"""The lock-step multi-pair driver (graph-cut-ransac_b200/host/gcr_batch_engine.h): P image pairs resident on one context,
every kind of device operation executed as one launch for all the pairs that asked.  Pair p of a batch must return what
the single-pair driver returns for it with the same seed -- iteration counts, local optimisations, masks, models -- and
the single-pair driver is pinned against the verbatim reference run in tests/test_gpu_driver.py."""
import os

import numpy as np
import pytest
from conftest import GOLDEN
from pygcransac import synth

pytestmark = pytest.mark.gpu

STAT_KEYS = ("iteration_number", "graph_cut_number", "local_optimization_number", "main_samples", "lo_samples", "degensac_runs",
             "degensac_updates")


@pytest.fixture(scope="module")
def pyg():
    import pygcransac
    try:
        pygcransac._load_driver()
    except Exception as e:  # built only where the reference tree (BK max-flow) exists
        pytest.skip("host driver not available: %s" % e)
    yield pygcransac
    pygcransac.release_batch_engines()


def _check(pyg, single, batch_out, P, exact_models=True):
    models, masks, counts = batch_out
    stats = pyg.last_batch_statistics
    for p in range(P):
        M, mask = single(p)
        st = pyg.last_statistics
        got = stats[p].as_dict()
        for k in STAT_KEYS:
            assert got[k] == st[k], (p, k, got, st)
        assert np.array_equal(masks[p], mask), (p, int(masks[p].sum()), int(mask.sum()))
        assert counts[p] == mask.sum()
        if M is None:
            assert counts[p] == 0
            continue
        if exact_models:
            assert np.array_equal(models[p], M), (p, np.abs(models[p] - M).max())
        else:
            assert np.abs(models[p] - M).max() <= 1e-9 * np.abs(M).max()
        assert got["score"] == st["score"] or abs(got["score"] - st["score"]) <= 1e-12 * abs(st["score"])


@pytest.mark.parametrize("seats,engines", [(4, 1), (6, 2), (12, 3)])
def test_homography_batch_equals_single_pairs(pyg, seats, engines):
    P, n = 12, 1000
    pts = np.stack([synth.homography_pair(n, 0.3 + 0.02 * p, seed=50 + p)[0] for p in range(P)])
    seeds = np.arange(101, 101 + P, dtype=np.uint64)
    kw = dict(threshold=3.0, conf=0.99, spatial_coherence_weight=0.975, max_iters=10000, min_iters=50, use_sprt=False, sampler=0,
              neighborhood=0, neighborhood_size=4.0, lo_number=50)
    out = pyg.findHomographyBatch(pts, 1024, 1024, 1024, 1024, seeds=seeds, seats=seats, engines=engines, **kw)
    _check(pyg, lambda p: pyg.findHomography(pts[p], 1024, 1024, 1024, 1024, [], seed=int(seeds[p]), **kw), out, P)


def test_homography_batch_cfg1_progressive_napsac_and_prosac(pyg):
    """BASELINE configs[0] as a batch: the bundled pair with the reference test's own call (P-NAPSAC: the index sets are
    drawn on the host and go up per batch), one seed per seat; PROSAC (the Python default) as well."""
    g = np.load(os.path.join(GOLDEN, "cfg1_matches.npz"))
    h1, w1, h2, w2 = (int(v) for v in g["size"])
    P = 8
    pts = np.stack([g["pts"].astype(np.float64)] * P)
    seeds = np.arange(1, P + 1, dtype=np.uint64)
    for sampler in (2, 1):
        kw = dict(threshold=3.0, conf=0.99, spatial_coherence_weight=0.975, max_iters=5000, min_iters=50, use_sprt=False,
                  sampler=sampler, neighborhood=0, neighborhood_size=4.0, lo_number=50)
        out = pyg.findHomographyBatch(pts, h1, w1, h2, w2, seeds=seeds, seats=8, engines=2, **kw)
        _check(pyg, lambda p: pyg.findHomography(pts[p], h1, w1, h2, w2, [], seed=int(seeds[p]), **kw), out, P)
        assert (out[2] > 250).all()                                         # the reference's own floor (tests/test_homography.py:85)


def test_fundamental_batch_default_pipeline_equals_single_pairs(pyg):
    """Every default of findFundamentalMatrix: PROSAC, 8-point + LM bundle adjustment refits, DEGENSAC (one plane-dominated
    scene: the inner plane-and-parallax run works on the worker's own context), post-run polish."""
    scenes = [synth.epipolar_pair(1500, 0.5, seed=36)[0], synth.epipolar_pair_planar(1000, 0.4, 0.9, seed=12)[0][:1000],
              synth.epipolar_pair(1500, 0.3, seed=37)[0], synth.epipolar_pair_planar(1000, 0.4, 0.95, seed=13)[0]]
    n = min(len(s) for s in scenes)
    pts = np.stack([s[:n] for s in scenes] * 2)
    P = len(pts)
    seeds = np.arange(1, P + 1, dtype=np.uint64)
    kw = dict(threshold=0.75, spatial_coherence_weight=0.0, neighborhood=0, neighborhood_size=4.0, use_sprt=False)
    out = pyg.findFundamentalMatrixBatch(pts, 1024, 1024, 1024, 1024, seeds=seeds, seats=8, engines=2, **kw)
    _check(pyg, lambda p: pyg.findFundamentalMatrix(pts[p], 1024, 1024, 1024, 1024, [], seed=int(seeds[p]), **kw), out, P)
    assert sum(pyg.last_batch_statistics[p].degensac_runs for p in range(P)) >= 1
    # graph-cut on (spatial weight > 0, grid neighbourhood), linear 8-point refits
    kw = dict(threshold=0.75, spatial_coherence_weight=0.14, neighborhood=0, neighborhood_size=4.0, use_sprt=False, sampler=0,
              bundle_adjustment=False, use_degensac=False)
    out = pyg.findFundamentalMatrixBatch(pts, 1024, 1024, 1024, 1024, seeds=seeds, seats=5, engines=1, **kw)
    _check(pyg, lambda p: pyg.findFundamentalMatrix(pts[p], 1024, 1024, 1024, 1024, [], seed=int(seeds[p]), **kw), out, P)


def test_essential_batch_equals_single_pairs(pyg):
    """findEssentialMatrix: 5-point + Stewenius refits (up to 10 models per fit), grid on the pixel coordinates, MAGSAC++
    scoring, the post-run bundle adjustment; the radius neighbourhood (the Python default) on a second pass."""
    P, n = 6, 800
    scenes = [synth.epipolar_pair(n, 0.5 + 0.03 * p, seed=60 + p) for p in range(P)]
    K = scenes[0][3]
    pts = np.stack([s[0] for s in scenes])
    seeds = np.arange(7, 7 + P, dtype=np.uint64)
    for kw in (dict(threshold=0.75, spatial_coherence_weight=0.1, neighborhood=0, neighborhood_size=4.0, use_sprt=False, sampler=0, scoring=1),
               dict(threshold=0.75, use_sprt=False, max_iters=2000)):
        out = pyg.findEssentialMatrixBatch(pts, K, K, 1024, 1024, 1024, 1024, seeds=seeds, seats=6, engines=2, **kw)
        _check(pyg, lambda p: pyg.findEssentialMatrix(pts[p], K, K, 1024, 1024, 1024, 1024, [], seed=int(seeds[p]), **kw), out, P)


def test_batch_argument_errors(pyg):
    pts = np.zeros((2, 100, 4))
    with pytest.raises(ValueError):
        pyg.findHomographyBatch(np.zeros((100, 4)), 10, 10, 10, 10)
    with pytest.raises(NotImplementedError):
        pyg.findHomographyBatch(pts, 10, 10, 10, 10, sampler=3)
    with pytest.raises(ValueError):
        pyg.findHomographyBatch(pts, 10, 10, 10, 10, seeds=[1, 2, 3])
