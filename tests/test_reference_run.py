"""BASELINE configs[0] (findHomography on the bundled pair, threshold 3.0 px, on CPU) through the reference's own
GCRANSAC<>::run compiled verbatim (oracle/_ref), and the drop-in scoring adapter inside that same driver on the GPU.

The committed fixture (oracle/gen_cfg1_fixture.py) holds the SIFT tentatives of tests/test_data/004520 <-> 004523 and the
reference driver's results for three sample-stream seeds."""
import os

import numpy as np
import oracle as O
import pytest
from conftest import GOLDEN
from pygcransac import synth


def _cfg1():
    g = np.load(os.path.join(GOLDEN, "cfg1_matches.npz"))
    return g, g["pts"].astype(np.float64)


def _run_cfg1(impl, pts, cs, seed, **kw):
    return impl.run(O.H, pts, cs, 4, 3.0, conf=0.99, spatial_weight=0.975, max_iters=10000, min_iters=50, lo_number=50,
                    seed=seed, **kw)


def test_cfg1_reference_run_reproduces_fixture(ref):
    g, pts = _cfg1()
    assert pts.shape == (762, 4)
    for seed in (1, 2, 3):
        r = _run_cfg1(ref, pts, g["cell_sizes"], seed)
        assert np.array_equal(r["mask"], g["mask_%d" % seed])
        assert np.array_equal(r["model"], g["model_%d" % seed])
        assert r["inliers"] > 300
        # the three seeds agree on the scene's dominant plane
        H0 = g["model_1"] / g["model_1"][2, 2]
        H = r["model"] / r["model"][2, 2]
        common = g["mask_1"] & r["mask"]
        assert common.sum() > 0.9 * min(g["mask_1"].sum(), r["mask"].sum())
        assert np.abs(H - H0).max() < 0.2 * np.abs(H0).max()


def test_reference_run_synthetic_f(ref):
    pts, Fm, Em, K, inl = synth.epipolar_pair(1000, 0.4, seed=8)
    cs = np.array([synth.SIZE / 4.0] * 4)
    r = ref.run(O.F, pts, cs, 4, 0.75, seed=3, spatial_weight=0.1)
    assert (r["mask"] & inl).sum() > 0.95 * inl.sum() and (r["mask"] & ~inl).sum() < 20


@pytest.mark.gpu
def test_cuda_scoring_adapter_inside_reference_driver(ref):
    """gcr::CudaMSACScoringFunction as GCRANSAC<>'s _ScoringFunction: same sample stream, same best-model sequence, same
    LO and IRLS decisions -> identical inlier mask and model as the reference's own MSACScoringFunction."""
    try:
        cuda = O.RefCuda()
    except (FileNotFoundError, OSError) as e:
        pytest.skip("oracle/_ref/libgcr_ref_cuda.so not available: %s" % e)
    g, pts = _cfg1()
    for seed in (1, 2):
        want = _run_cfg1(ref, pts, g["cell_sizes"], seed)
        got = _run_cfg1(cuda, pts, g["cell_sizes"], seed, scoring_kind=1)
        assert np.array_equal(got["mask"], want["mask"])
        assert got["stats"] == want["stats"]
        assert np.abs(got["model"] - want["model"]).max() <= 1e-12 * np.abs(want["model"]).max()
        assert abs(got["score"] - want["score"]) <= 1e-9 * want["score"]
    ptsf, Fm, Em, K, inl = synth.epipolar_pair(1500, 0.5, seed=9)
    cs = np.array([synth.SIZE / 4.0] * 4)
    want = ref.run(O.F, ptsf, cs, 4, 0.75, seed=5, spatial_weight=0.1)
    got = cuda.run(O.F, ptsf, cs, 4, 0.75, seed=5, spatial_weight=0.1, scoring_kind=1)
    assert np.array_equal(got["mask"], want["mask"]) and got["stats"] == want["stats"]
    assert np.abs(got["model"] - want["model"]).max() <= 1e-12 * np.abs(want["model"]).max()
