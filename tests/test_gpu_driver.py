"""The batched host driver (graph-cut-ransac_b200/host/gcr_driver.cpp) through pygcransac.findHomography /
findFundamentalMatrix against the reference's GCRANSAC<>::run compiled verbatim (oracle/_ref) on the SAME counter-based
sample streams: same iteration counts, same local optimisations, same inlier mask; models to 1e-5 (north star)."""
import os

import numpy as np
import oracle as O
import pytest
from conftest import GOLDEN
from pygcransac import synth

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
    except Exception as e:  # built only where the reference tree (BK max-flow) exists
        pytest.skip("host driver not available: %s" % e)
    return pygcransac


def test_findhomography_cfg1_matches_reference_run(pyg, ref):
    """BASELINE configs[0]: the bundled pair with the call of the reference's own test -- sampler=1 is PROSAC over the
    SNN-ratio order (py/gcransac_python.cpp:4320-4323); sampler=0 (uniform) as well."""
    g = np.load(os.path.join(GOLDEN, "cfg1_matches.npz"))
    pts = g["pts"]                                   # float32 N x 4, like the reference's tests pass it
    h1, w1, h2, w2 = (int(v) for v in g["size"])
    for sampler, prefix in ((1, "prosac_"), (0, "")):
        for seed in (1, 2, 3):
            H, mask = pyg.findHomography(pts, h1, w1, h2, w2, [], threshold=3.0, conf=0.99, spatial_coherence_weight=0.975,
                                         max_iters=10000, min_iters=50, use_sprt=False, sampler=sampler, neighborhood=0,
                                         neighborhood_size=4.0, use_space_partitioning=False, lo_number=50, seed=seed)
            st = pyg.last_statistics
            want = ref.run(O.H, pts.astype(np.float64), g["cell_sizes"], 4, 3.0, conf=0.99, spatial_weight=0.975,
                           max_iters=10000, min_iters=50, lo_number=50, seed=seed, sampler=sampler)
            assert st["iteration_number"] == want["stats"]["iterations"], (sampler, seed, st, want["stats"])
            assert st["graph_cut_number"] == want["stats"]["graph_cuts"]
            assert st["local_optimization_number"] == want["stats"]["local_optimizations"]
            assert st["lo_samples"] == want["stats"]["lo_samples"]
            assert np.array_equal(mask, want["mask"]), (sampler, seed, int(mask.sum()), int(want["mask"].sum()))
            assert _up_to_scale(want["model"], H) < 1e-5
            assert abs(st["score"] - want["score"]) <= 1e-9 * want["score"]
            assert np.array_equal(mask, g["%smask_%d" % (prefix, seed)])      # and the committed fixture
            assert mask.sum() > 250                                            # the reference's own floor (tests/test_homography.py:85)


def test_importance_sampler_matches_reference_run(pyg, ref):
    """sampler=3 (NG-RANSAC's ImportanceSampler, py/gcransac_python.cpp:4343-4346): the same run as the verbatim reference
    driver over the reference's own sampler distribution, on the bundled pair (rank-based probabilities like the
    reference's tests build them) and on synthetic F with noisy inlier probabilities."""
    g = np.load(os.path.join(GOLDEN, "cfg1_matches.npz"))
    pts = g["pts"]
    h1, w1, h2, w2 = (int(v) for v in g["size"])
    n = len(pts)
    prob = 1.0 - np.arange(n) / n                    # SNN-ratio rank -> probability (tests/test_*.py: get_probabilities)
    for seed in (1, 2):
        H, mask = pyg.findHomography(pts, h1, w1, h2, w2, prob, threshold=3.0, conf=0.99, spatial_coherence_weight=0.975,
                                     max_iters=10000, min_iters=50, use_sprt=False, sampler=3, neighborhood=0,
                                     neighborhood_size=4.0, use_space_partitioning=False, lo_number=50, seed=seed)
        st = pyg.last_statistics
        want = ref.run(O.H, pts.astype(np.float64), g["cell_sizes"], 4, 3.0, conf=0.99, spatial_weight=0.975,
                       max_iters=10000, min_iters=50, lo_number=50, seed=seed, sampler=3, probabilities=prob)
        assert st["iteration_number"] == want["stats"]["iterations"], (seed, st, want["stats"])
        assert st["lo_samples"] == want["stats"]["lo_samples"]
        assert np.array_equal(mask, want["mask"]), (seed, int(mask.sum()), int(want["mask"].sum()))
        assert _up_to_scale(want["model"], H) < 1e-5
        assert mask.sum() > 250
    pts, Fm, Em, K, inl = synth.epipolar_pair(2000, 0.6, seed=8)
    rng = np.random.default_rng(4)
    prob = np.clip(np.where(inl, 0.7, 0.3) + rng.normal(0, 0.2, len(pts)), 0.0, 1.0)
    F, mask = pyg.findFundamentalMatrix(pts, 1024, 1024, 1024, 1024, prob, threshold=0.75, spatial_coherence_weight=0.1,
                                        neighborhood=0, neighborhood_size=8.0, use_sprt=False, seed=6, sampler=3,
                                        use_degensac=False, bundle_adjustment=False)
    st = pyg.last_statistics
    want = ref.run(O.F, pts, np.array([1024 / 8] * 4), 8, 0.75, conf=0.99, spatial_weight=0.1, max_iters=10000, min_iters=50,
                   lo_number=50, seed=6, sampler=3, probabilities=prob)
    assert st["iteration_number"] == want["stats"]["iterations"], (st, want["stats"])
    assert np.array_equal(mask, want["mask"]) and _up_to_scale(want["model"], F) < 1e-5
    assert (mask & inl).sum() > 0.9 * inl.sum()
    with pytest.raises(ValueError):
        pyg.findHomography(pts, 1024, 1024, 1024, 1024, prob[:-1], sampler=3)      # one probability per correspondence
    with pytest.raises(ValueError):
        pyg.findHomography(pts, 1024, 1024, 1024, 1024, np.zeros(len(pts)), sampler=3)


def test_progressive_napsac_gpu_built_layers_match_reference_run(pyg, ref):
    """sampler=2 on a point set large enough for the four grid layers to come from gcr_grid_build (smaller sets build
    them on the host): the same run as the verbatim reference driver over the reference's own P-NAPSAC class."""
    pts, Hm, inl = synth.homography_pair(20000, 0.5, seed=41)
    H, mask = pyg.findHomography(pts, 1024, 1024, 1024, 1024, [], threshold=2.0, spatial_coherence_weight=0.0, max_iters=2000,
                                 use_sprt=False, use_space_partitioning=False, sampler=2, seed=3)
    st = pyg.last_statistics
    cs = np.array([(1024 + np.finfo(np.float64).eps) / 4] * 4)
    want = ref.run(O.H, pts, cs, 4, 2.0, conf=0.99, spatial_weight=0.0, max_iters=2000, min_iters=50, lo_number=50, seed=3,
                   sampler=2, image_sizes=(1024, 1024, 1024, 1024))
    assert st["iteration_number"] == want["stats"]["iterations"], (st, want["stats"])
    assert st["lo_samples"] == want["stats"]["lo_samples"]
    assert np.array_equal(mask, want["mask"]) and _up_to_scale(want["model"], H) < 1e-5
    assert (mask & inl).sum() > 0.97 * inl.sum()


def test_magsac_scoring_pipeline_matches_reference_run(pyg, ref, oracle):
    """GCRANSAC<Estimator, Graph, MAGSACScoringFunction<Estimator>> (scoring_function.h:237-349; BASELINE configs[2] is
    E + MAGSAC++): the driver with gcr_run_settings.scoring = GCR_MAGSAC against the verbatim reference driver
    instantiated with the reference's own MAGSACScoringFunction, E and H.  The mask marks residual <= threshold
    (MAGSAC's inlier rule, :338-344), not MSAC's 1.5 x threshold."""
    from pygcransac import _cabi
    K = synth.K_MAT
    for seed, outliers in ((6, 0.5), (7, 0.7)):
        pts, Fm, Em, Kc, inl = synth.epipolar_pair(1500, outliers, seed=seed + 2)
        E, mask = pyg.findEssentialMatrix(pts, K, K, 1024, 1024, 1024, 1024, [], threshold=0.75, max_iters=5000, sampler=0,
                                          neighborhood=0, neighborhood_size=4.0, seed=seed, scoring=_cabi.MAGSAC)
        st = pyg.last_statistics
        npts = oracle.normalize_correspondences(pts, K, K)
        thr = 0.75 / (0.25 * (K[0, 0] + K[1, 1] + K[0, 0] + K[1, 1]))
        want = ref.run(O.E, npts, np.array([256.0] * 4), 4, thr, conf=0.99, spatial_weight=0.1, max_iters=5000, min_iters=50,
                       lo_number=50, seed=seed, grid_pts=pts, scoring_kind=2, f_ba=4)
        assert st["iteration_number"] == want["stats"]["iterations"], (seed, st, want["stats"])
        assert st["lo_samples"] == want["stats"]["lo_samples"]
        assert (mask != want["mask"]).sum() <= 2 and _up_to_scale(want["model"], E) < 1e-5
        assert abs(st["score"] - want["score"]) <= 1e-6 * want["score"]
        assert (mask & inl).sum() > 0.8 * inl.sum()
    # F, plain (8-point refits) and the default estimator (BA refits, DEGENSAC with its MSAC-scored inner run, post-run BA)
    pts, Fm, Em, Kc, inl = synth.epipolar_pair(2000, 0.5, seed=36)
    cs4 = np.array([1024 / 4.0] * 4)
    for full in (False, True):
        F, mask = pyg.findFundamentalMatrix(pts, 1024, 1024, 1024, 1024, [], threshold=0.75, spatial_coherence_weight=0.0,
                                            neighborhood=0, neighborhood_size=4.0, seed=4, sampler=0, use_degensac=full,
                                            bundle_adjustment=full, scoring=_cabi.MAGSAC)
        st = pyg.last_statistics
        want = ref.run(O.F, pts, cs4, 4, 0.75, conf=0.99, spatial_weight=0.0, max_iters=10000, min_iters=50, lo_number=50, seed=4,
                       use_degensac=full, f_ba=3 if full else 0, scoring_kind=2)
        assert st["iteration_number"] == want["stats"]["iterations"], (full, st, want["stats"])
        assert (mask != want["mask"]).sum() <= (2 if full else 0) and _up_to_scale(want["model"], F) < 1e-5, full
    pts, Hm, inl = synth.homography_pair(2000, 0.4, seed=8)
    H, mask = pyg.findHomography(pts, 1024, 1024, 1024, 1024, [], threshold=2.0, seed=6, scoring=_cabi.MAGSAC)
    st = pyg.last_statistics
    cs = np.array([(1024 + np.finfo(np.float64).eps) / 4] * 4)
    want = ref.run(O.H, pts, cs, 4, 2.0, seed=6, scoring_kind=2, sampler=1)
    assert st["iteration_number"] == want["stats"]["iterations"], (st, want["stats"])
    assert np.array_equal(mask, want["mask"]) and _up_to_scale(want["model"], H) < 1e-5
    assert abs(st["score"] - want["score"]) <= 1e-9 * want["score"]


def test_ar_sampler_when_model_generation_gives_up(pyg, ref):
    """Two thirds of the correspondences are one and the same point and carry the highest probabilities: the AR-Sampler
    keeps proposing degenerate samples, GCRANSAC::run gives up after 100 of them in a row (GCRANSAC.h:297-339) and its
    end-of-iteration update() then runs on the last sample a second time (:474) -- which changes every later sample.  The
    driver rebuilds its drawn-ahead sequence at that point; several such events, then a model is found."""
    n = 600
    cs = np.array([(1024 + np.finfo(np.float64).eps) / 4] * 4)
    prob = 1.0 - np.arange(n) / n
    for degenerate, variance in ((400, 0.1), (500, 0.9765)):
        pts, Hm, inl = synth.homography_pair(n, 0.2, seed=5)
        pts[:degenerate] = pts[0]
        H, mask = pyg.findHomography(pts, 1024, 1024, 1024, 1024, prob, threshold=3.0, max_iters=5000, use_sprt=False,
                                     use_space_partitioning=False, sampler=4, sampler_variance=variance, seed=3)
        st = pyg.last_statistics
        want = ref.run(O.H, pts, cs, 4, 3.0, conf=0.99, spatial_weight=0.975, max_iters=5000, min_iters=50, lo_number=50, seed=3,
                       sampler=4, probabilities=prob, sampler_variance=variance)
        give_ups = want["stats"]["iterations"] - want["stats"]["main_samples"]     # an exhausted iteration counts 101 for 100 samples
        assert give_ups >= 5, want["stats"]
        assert st["iteration_number"] == want["stats"]["iterations"], (st, want["stats"])
        assert st["main_samples"] == want["stats"]["main_samples"]
        assert np.array_equal(mask, want["mask"]) and _up_to_scale(want["model"], H) < 1e-5 and mask.sum() > 500


def test_findhomography_synthetic_and_errors(pyg, ref):
    pts, Hm, inl = synth.homography_pair(3000, 0.5, seed=33)
    H, mask = pyg.findHomography(pts, 1024, 1024, 1024, 1024, [], threshold=2.0, use_sprt=False, use_space_partitioning=False, seed=5)
    cs = np.array([(1024 + np.finfo(np.float64).eps) / 4] * 4)
    want = ref.run(O.H, pts, cs, 4, 2.0, conf=0.99, spatial_weight=0.975, max_iters=10000, min_iters=50, lo_number=50, seed=5,
                   sampler=1)                                       # the Python default sampler is PROSAC
    assert np.array_equal(mask, want["mask"]) and _up_to_scale(want["model"], H) < 1e-5
    assert (mask & inl).sum() > 0.97 * inl.sum()
    with pytest.raises(ValueError):
        pyg.findHomography(np.zeros((3, 4)), 10, 10, 10, 10, [])
    with pytest.raises(ValueError):
        pyg.findHomography(np.zeros((10, 3)), 10, 10, 10, 10, [])
    with pytest.raises(NotImplementedError):
        pyg.findHomography(pts, 1024, 1024, 1024, 1024, [], sampler=7)
    with pytest.raises(NotImplementedError):                         # SIFT / affine correspondence solvers are other estimators
        pyg.findHomography(pts, 1024, 1024, 1024, 1024, [], solver=1)
    with pytest.raises(ValueError):                                  # NG-RANSAC / AR need one probability per correspondence
        pyg.findHomography(pts, 1024, 1024, 1024, 1024, [], sampler=4)
    # no model: pure noise, tiny threshold -> (None, all-false mask) like bindings.cpp:973-975
    rng = np.random.default_rng(0)
    Hn, maskn = pyg.findHomography(rng.uniform(0, 1000, (50, 4)), 1000, 1000, 1000, 1000, [], threshold=1e-6, max_iters=200)
    assert Hn is None and not maskn.any()


def test_findfundamental_matches_reference_run(pyg, ref):
    pts, Fm, Em, K, inl = synth.epipolar_pair(2000, 0.5, seed=34)
    for seed, weight in ((7, 0.0), (8, 0.1)):
        F, mask = pyg.findFundamentalMatrix(pts, 1024, 1024, 1024, 1024, [], threshold=0.75, spatial_coherence_weight=weight,
                                            neighborhood=0, neighborhood_size=4.0, use_sprt=False, seed=seed,
                                            use_degensac=False, bundle_adjustment=False, sampler=0)
        st = pyg.last_statistics
        cs = np.array([1024 / 4.0] * 4)
        want = ref.run(O.F, pts, cs, 4, 0.75, conf=0.99, spatial_weight=weight, max_iters=10000, min_iters=50, lo_number=50, seed=seed)
        assert st["iteration_number"] == want["stats"]["iterations"], (st, want["stats"])
        assert st["lo_samples"] == want["stats"]["lo_samples"]
        assert np.array_equal(mask, want["mask"])
        assert _up_to_scale(want["model"], F) < 1e-5
        assert (mask & inl).sum() > 0.95 * inl.sum()


def test_findfundamental_degensac_matches_reference_run(pyg, ref):
    """A scene dominated by one plane: the best 7-point samples are H-degenerate, DEGENSAC (fundamental_estimator.h
    :341-572) refits the plane homography and runs the inner plane-and-parallax GC-RANSAC.  The driver's nested run
    against the reference's verbatim GCRANSAC<> (outer AND inner) over the oracle estimators, same sample streams."""
    cs = np.array([1024 / 4.0] * 4)
    runs = updates = 0
    for seed, plane in ((1, 0.8), (3, 0.8), (1, 0.95), (2, 0.95), (3, 0.95), (4, 0.0)):
        pts, Fm, inl, on_plane = synth.epipolar_pair_planar(1000, 0.4, plane, seed + 10)
        F, mask = pyg.findFundamentalMatrix(pts, 1024, 1024, 1024, 1024, [], threshold=0.75, spatial_coherence_weight=0.0,
                                            neighborhood=0, neighborhood_size=4.0, use_sprt=False, seed=seed,
                                            bundle_adjustment=False, sampler=0)
        st = pyg.last_statistics
        want = ref.run(O.F, pts, cs, 4, 0.75, conf=0.99, spatial_weight=0.0, max_iters=10000, min_iters=50, lo_number=50,
                       seed=seed, use_degensac=True)
        assert st["degensac_runs"] == want["stats"]["degensac_runs"], (seed, plane, st, want["stats"])
        assert st["degensac_updates"] == want["stats"]["degensac_updates"]
        assert st["iteration_number"] == want["stats"]["iterations"], (seed, plane, st, want["stats"])
        assert st["lo_samples"] == want["stats"]["lo_samples"]
        assert np.array_equal(mask, want["mask"]), (seed, plane, int(mask.sum()), int(want["mask"].sum()))
        assert _up_to_scale(want["model"], F) < 1e-5
        assert (mask & inl).sum() > 0.9 * inl.sum()
        runs += st["degensac_runs"]
        updates += st["degensac_updates"]
    assert runs >= 3 and updates >= 2      # the degenerate branch and the model replacement were both exercised


def test_findfundamental_default_estimator_matches_reference_run(pyg, ref):
    """The reference's DEFAULT F pipeline: non-minimal fits by 8-point + LM bundle adjustment (types.h:78-80), DEGENSAC,
    and the post-run bundle adjustment of py/gcransac_python.cpp:1832-1887 -- the GPU driver against the verbatim
    GCRANSAC<> over the oracle estimator with the same parts switched on.  The LM sums are ordered differently on the
    two sides (block tree vs sequential), so masks may differ on points at the threshold: a few bits of slack."""
    cs = np.array([1024 / 4.0] * 4)
    general = lambda t: (t[0], t[1], t[4])                                  # (pts, F, E, K, inliers) -> (pts, F, inliers)
    for seed, (pts, Fm, inl) in enumerate([general(synth.epipolar_pair(2000, 0.5, seed=36)),
                                           synth.epipolar_pair_planar(1000, 0.4, 0.9, seed=12)[:3],
                                           general(synth.epipolar_pair(1200, 0.3, seed=37))], start=1):
        F, mask = pyg.findFundamentalMatrix(pts, 1024, 1024, 1024, 1024, [], threshold=0.75, spatial_coherence_weight=0.0,
                                            neighborhood=0, neighborhood_size=4.0, use_sprt=False, seed=seed)
        st = pyg.last_statistics
        want = ref.run(O.F, pts, cs, 4, 0.75, conf=0.99, spatial_weight=0.0, max_iters=10000, min_iters=50, lo_number=50,
                       seed=seed, use_degensac=True, f_ba=3, sampler=1)  # every Python default: PROSAC, BA, DEGENSAC
        assert st["iteration_number"] == want["stats"]["iterations"], (seed, st, want["stats"])
        assert st["lo_samples"] == want["stats"]["lo_samples"]
        assert st["degensac_runs"] == want["stats"]["degensac_runs"]
        assert (mask != want["mask"]).sum() <= 2, (seed, int(mask.sum()), int(want["mask"].sum()))
        assert _up_to_scale(want["model"], F) < 1e-5
        assert abs(st["score"] - want["score"]) <= 1e-5 * want["score"]
        assert (mask & inl).sum() > 0.9 * inl.sum()


def test_findessential_matches_reference_run(pyg, ref, oracle):
    """findEssentialMatrix (5-point minimal, Stewenius non-minimal refits with up to 10 models each, MSAC on K-normalised
    points, grid on the pixel coordinates) against the reference's verbatim GCRANSAC<>::run over the oracle's E estimator."""
    pts, Fm, Em, K, inl = synth.epipolar_pair(1500, 0.5, seed=35)
    npts = oracle.normalize_correspondences(pts, K, K)
    for seed, weight in ((7, 0.0), (8, 0.1)):
        E, mask = pyg.findEssentialMatrix(pts, K, K, 1024, 1024, 1024, 1024, [], threshold=0.75, spatial_coherence_weight=weight,
                                          neighborhood=0, neighborhood_size=4.0, use_sprt=False, seed=seed, sampler=0)
        st = pyg.last_statistics
        cs = np.array([1024 / 4.0] * 4)
        thr = 0.75 / (0.25 * (K[0, 0] + K[1, 1] + K[0, 0] + K[1, 1]))          # py/gcransac_python.cpp:2624-2626, 2737
        want = ref.run(O.E, npts, cs, 4, thr, conf=0.99, spatial_weight=weight, max_iters=10000, min_iters=50, lo_number=50,
                       seed=seed, grid_pts=pts, f_ba=4)
        assert st["iteration_number"] == want["stats"]["iterations"], (st, want["stats"])
        assert st["graph_cut_number"] == want["stats"]["graph_cuts"]
        assert st["lo_samples"] == want["stats"]["lo_samples"]
        assert (mask != want["mask"]).sum() <= 2, (int(mask.sum()), int(want["mask"].sum()))
        assert _up_to_scale(want["model"], E) < 1e-5
        assert abs(st["score"] - want["score"]) <= 1e-6 * want["score"]
        assert (mask & inl).sum() > 0.95 * inl.sum()
    with pytest.raises(ValueError):
        pyg.findEssentialMatrix(pts, np.eye(2), K, 10, 10, 10, 10, [])
    with pytest.raises(ValueError):
        pyg.findEssentialMatrix(np.zeros((4, 4)), K, K, 10, 10, 10, 10, [])


def test_find6dpose_matches_reference_run(pyg, ref, oracle):
    """find6DPose (py/gcransac_python.cpp:1026-1272: GCRANSAC<DefaultPnPEstimator>, P3P minimal solver, radius neighbourhood
    over the 5-column rows, the non-minimal fit on the inliers replacing the pose at the end) against the reference's
    verbatim GCRANSAC<>::run over the oracle's PnP estimator (same stated non-minimal solver: DLT + the LM of refine_pnp):
    iteration / graph-cut / LO-sample counts, masks, poses -- a synthetic scene (with and without the spatial term) and
    the reference's own example scene with the notebook's 75 % outliers."""
    import os
    GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    g = np.load(os.path.join(GOLDEN, "pose6dscene.npz"))
    spts, P0, sinl = synth.pnp_scene(800, 0.5, seed=21)
    cases = [(spts, 2.0 / synth.FOCAL, 0.975, 0.6, 3, P0, sinl, 5e-3),
             (spts, 2.0 / synth.FOCAL, 0.0, 0.6, 4, P0, sinl, 5e-3),
             (g["pts"], float(g["threshold"]), 0.975, 8.0, 5, g["gt_pose"], g["inlier"], 2e-2)]
    for pts, thr, weight, radius, seed, P_true, inl, tol in cases:
        P, mask = pyg.find6DPose(pts, [], threshold=thr, conf=0.99, spatial_coherence_weight=weight, max_iters=5000,
                                 min_iters=50, sampler=0, neighborhood=0, neighborhood_size=radius, seed=seed)
        st = pyg.last_statistics
        want = ref.run(O.PNP, pts, np.ones(4), 1, thr, conf=0.99, spatial_weight=weight, max_iters=5000, min_iters=50,
                       lo_number=50, seed=seed, radius=radius if weight > 0 else 0.0, f_ba=8)
        assert P is not None and P.shape == (3, 4)
        assert st["iteration_number"] == want["stats"]["iterations"], (st, want["stats"])
        assert st["graph_cut_number"] == want["stats"]["graph_cuts"] and st["lo_samples"] == want["stats"]["lo_samples"]
        assert (mask != want["mask"]).sum() <= 2, (int(mask.sum()), int(want["mask"].sum()))
        W = want["model"].reshape(3, 4)
        assert np.abs(P[:, :3] - W[:, :3]).max() < 1e-5 and np.abs(P[:, 3] - W[:, 3]).max() < 1e-5 * max(1.0, np.abs(W[:, 3]).max())
        assert (mask & inl).sum() > 0.9 * inl.sum()
        assert np.abs(P[:, :3] - P_true[:, :3]).max() < tol
    # binding behaviour: shape errors, unknown neighbourhood id, P-NAPSAC
    with pytest.raises(ValueError):
        pyg.find6DPose(np.zeros((10, 4)), [])
    with pytest.raises(ValueError):
        pyg.find6DPose(np.zeros((3, 5)), [])
    assert pyg.find6DPose(spts, [], neighborhood=1)[0] is None and pyg.find6DPose(spts, [], sampler=2)[0] is None
