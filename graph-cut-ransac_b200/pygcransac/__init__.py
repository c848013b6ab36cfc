"""pygcransac (B200) -- the reference's Python entry points on top of libgcr_b200.so.

    findHomography / findFundamentalMatrix / findEssentialMatrix
                                same names, arguments, defaults and return values as danini/graph-cut-ransac's
                                pybind11 module (py/bindings.cpp:1003-1178)

The data-parallel loop body runs on the GPU through the batched host driver (graph-cut-ransac_b200/host/gcr_driver.cpp:
GCRANSAC<>::run's sequential semantics replayed over GPU batches; the graph-cut min-cut is the reference's own BK
max-flow on the host).  There is no CPU fallback: without a CUDA device / the built libraries these functions raise.

Accepted-but-mapped arguments: `use_sprt` and `use_space_partitioning` are CPU early-exit / pruning
tricks of the reference; the GPU scores every point of every model, which yields the scores those tricks approximate,
so the flags are accepted and ignored.  `sampler` takes every id of the reference: 0 uniform, 1 PROSAC (the
reference's default: the correspondences are taken to be ordered by quality), 3 the importance sampler of NG-RANSAC
(independent draws from `probabilities`) -- these three drawn on the GPU -- and 2 Progressive NAPSAC / 4 the AR-Sampler,
whose sample k depends on samples 1..k-1 (hit counters resp. a heap) but on no verdict: their sequences are drawn ahead
on the host (host/gcr_host_samplers.h) and uploaded per batch.  All on the counter-based word stream.
`neighborhood` is 0 (grid, `neighborhood_size` cells per axis) or 1 (the reference's FLANN radius graph,
`neighborhood_size` = radius in pixels: built here as the EXACT radius graph on the GPU, where OpenCV's FLANN returns an
approximate, run-dependent one).  Not reference arguments: `seed` (the counter-based sample streams), `device`, and
`scoring` (0 = MSAC, what the reference's entry points instantiate; 1 = MAGSAC++, the reference's
MAGSACScoringFunction template argument, scoring_function.h:237-349).
"""
import ctypes as C
import os
import threading

# One context (= stream) per host thread (`_context`): give every stream a hardware work queue of its own.  With CUDA's
# default of 8 connections the streams of sixteen threads share queues and one pair's latency-bound kernels hold up another
# pair's launches.  Only effective if CUDA has not been initialised in this process yet; a value set by the user stays.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import numpy as np

from . import _cabi

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_DRIVER_PATH = os.path.join(_PKG, "host", "_ref", "libgcr_driver.so")
_driver = None
_contexts = {}


class RunSettings(C.Structure):
    _fields_ = [("threshold", C.c_double), ("confidence", C.c_double), ("spatial_coherence_weight", C.c_double),
                ("max_iteration_number", C.c_uint64), ("min_iteration_number", C.c_uint64),
                ("max_local_optimization_number", C.c_uint64), ("min_iteration_number_before_lo", C.c_uint64),
                ("max_graph_cut_number", C.c_uint64), ("max_least_squares_iterations", C.c_uint64),
                ("max_unsuccessful_model_generations", C.c_uint64), ("do_local_optimization", C.c_int),
                ("do_final_iterated_least_squares", C.c_int), ("cell_sizes", C.c_double * 4),
                ("cells_per_axis", C.c_uint64), ("seed", C.c_uint64), ("batch", C.c_int), ("use_degensac", C.c_int), ("f_bundle_adjustment", C.c_int),
                ("sampler", C.c_int), ("neighborhood", C.c_int), ("neighborhood_radius", C.c_double),
                ("point_probabilities", C.c_void_p), ("sampler_variance", C.c_double), ("image_sizes", C.c_double * 4), ("scoring", C.c_int),
                ("e_bundle_adjustment", C.c_int), ("pnp_bundle_adjustment", C.c_int)]


class RunStats(C.Structure):
    _fields_ = [("iteration_number", C.c_uint64), ("graph_cut_number", C.c_uint64),
                ("local_optimization_number", C.c_uint64), ("main_samples", C.c_uint64), ("lo_samples", C.c_uint64),
                ("gpu_batches", C.c_uint64), ("exact_rescores", C.c_uint64), ("models_scored", C.c_uint64),
                ("score", C.c_double), ("graph_cut_seconds", C.c_double), ("degensac_runs", C.c_uint64),
                ("degensac_updates", C.c_uint64), ("phase_seconds", C.c_double * 6)]

    def as_dict(self):
        return {name: (list(getattr(self, name)) if name == "phase_seconds" else getattr(self, name)) for name, _ in self._fields_}


def _load_driver():
    global _driver
    if _driver is None:
        _cabi.load_library()
        if not os.path.exists(_DRIVER_PATH):
            raise _cabi.GcrError("%s is missing: build it with graph-cut-ransac_b200/host/build_driver.sh (needs the "
                                 "reference tree for its BK max-flow); there is no CPU fallback" % _DRIVER_PATH)
        lib = C.CDLL(_DRIVER_PATH)
        lib.gcr_run_settings_default.restype = None
        lib.gcr_run_settings_default.argtypes = [C.POINTER(RunSettings)]
        lib.gcr_find_model.restype = C.c_int
        lib.gcr_find_model.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(RunSettings), C.c_void_p,
                                       C.c_void_p, C.POINTER(RunStats)]
        lib.gcr_find_essential.restype = C.c_int
        lib.gcr_find_essential.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(RunSettings),
                                           C.c_void_p, C.c_void_p, C.POINTER(RunStats)]
        u64p = C.c_void_p
        lib.gcr_find_model_batch.restype = C.c_int
        lib.gcr_find_model_batch.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.POINTER(RunSettings), u64p, C.c_int,
                                             C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.gcr_find_essential_batch.restype = C.c_int
        lib.gcr_find_essential_batch.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                                 C.POINTER(RunSettings), u64p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                                 C.c_void_p, C.c_void_p]
        lib.gcr_batch_engines_release.restype = None
        lib.gcr_batch_engines_release.argtypes = []
        lib.gcr_batch_launch_count.restype = C.c_uint64
        lib.gcr_batch_launch_count.argtypes = []
        lib.gcr_driver_last_error.restype = C.c_char_p
        lib.gcr_driver_last_error.argtypes = []
        _driver = lib
    return _driver


def _context(device=0):
    """One gcr_ctx (stream + workspaces) per (device, calling thread): independent image pairs can be processed from
    several host threads at once -- ctypes drops the GIL during the driver call, the GPU overlaps their small kernels
    and copies (the reference's counterpart is one problem per host core)."""
    key = (device, threading.get_ident())
    ctx = _contexts.get(key)
    if ctx is None:
        ctx = _contexts[key] = _cabi.Context(device)
    return ctx


_tls = threading.local()  # _tls.last_statistics: the calling thread's most recent RunStats
last_statistics = None  # RunStats of the most recent call (the reference fills utils::RANSACStatistics; not returned by its API either)


def _settings(lib, sizes, threshold, conf, spatial_coherence_weight, max_iters, min_iters, sampler, neighborhood,
              neighborhood_size, lo_number, eps_in_cell_size, seed, use_degensac, f_bundle_adjustment, prob, sampler_variance,
              scoring, e_bundle_adjustment, pnp_bundle_adjustment):
    """gcr_run_settings as the reference's drivers fill utils::Settings (py/gcransac_python.cpp:4498-4504)."""
    s = RunSettings()
    lib.gcr_run_settings_default(C.byref(s))
    s.threshold, s.confidence, s.spatial_coherence_weight = threshold, conf, spatial_coherence_weight
    s.max_iteration_number, s.min_iteration_number = max_iters, min_iters
    s.max_local_optimization_number = lo_number
    cells = int(neighborhood_size)                                          # static_cast<size_t>(neighborhood_size)
    h1, w1, h2, w2 = sizes
    eps = np.finfo(np.float64).eps if eps_in_cell_size else 0.0            # gcransac_python.cpp:4290-4297 vs :1672-1677
    s.cells_per_axis = max(cells, 1)
    for i, v in enumerate((w1, h1, w2, h2)):
        s.cell_sizes[i] = (v + eps) / max(cells, 1)
    s.seed = seed
    s.use_degensac = 1 if use_degensac else 0
    s.sampler = sampler
    # neighborhood=1: the reference builds a FLANN radius graph with radius neighborhood_size (py/gcransac_python.cpp
    # :123-125); here the exact radius graph (gcr_radius_graph_*).  The grid settings below are then unused.
    s.neighborhood = neighborhood
    s.neighborhood_radius = float(neighborhood_size)
    s.f_bundle_adjustment = f_bundle_adjustment
    s.point_probabilities = prob.ctypes.data if sampler in (3, 4) else None
    s.sampler_variance = float(sampler_variance)
    s.scoring = int(scoring)
    s.e_bundle_adjustment = int(e_bundle_adjustment)
    s.pnp_bundle_adjustment = int(pnp_bundle_adjustment)
    for i, v in enumerate((w1, h1, w2, h2)):                                # sampler 2: ProgressiveNapsacSampler's `sizes_`
        s.image_sizes[i] = float(v)
    return s


def _find(problem, correspondences, sizes, threshold, conf, spatial_coherence_weight, max_iters, min_iters, sampler,
          neighborhood, neighborhood_size, lo_number, eps_in_cell_size, seed, device, intrinsics=None, use_degensac=False, f_bundle_adjustment=0,
          probabilities=(), sampler_variance=0.1, scoring=0, e_bundle_adjustment=0, pnp_bundle_adjustment=0):
    global last_statistics
    pts = np.ascontiguousarray(correspondences, dtype=np.float64)           # bindings.cpp: py::array_t<double>
    min_n = {_cabi.H: 4, _cabi.F: 7, _cabi.E: 5, _cabi.PNP: 4}[problem]
    if problem == _cabi.PNP:
        if pts.ndim != 2 or pts.shape[1] != 5 or pts.shape[0] < min_n:
            raise ValueError("x1y1x2y2z2 should be an array with dims [n,5], n>=4")                   # bindings.cpp:223-228
    elif pts.ndim != 2 or pts.shape[1] != 4:
        raise ValueError("correspondences should be an array with dims [n,4], n>=%d" % min_n)
    elif pts.shape[0] < min_n:
        raise ValueError("correspondences should be an array with dims [n,4], n>=%d" % min_n)        # bindings.cpp:881-889
    prob = None
    if sampler in (3, 4):
        prob = np.ascontiguousarray(probabilities, dtype=np.float64).ravel()
        if len(prob) != len(pts):
            raise ValueError("The number of correspondences and inlier probabilities should be the same")   # bindings.cpp
        # 3: NG-RANSAC's ImportanceSampler (py/gcransac_python.cpp:4343-4346): stateless, m independent draws from the
        # probabilities per sample -- drawn on the GPU (GCR_IMPORTANCE).  4: the AR-Sampler (:4348-4359): probabilities
        # divided by their maximum, heap recurrence on the host (host/gcr_host_samplers.h)
        if not (np.all(np.isfinite(prob)) and np.all(prob >= 0) and prob.sum() > 0):
            raise ValueError("probabilities should be finite, non-negative and not all zero")
    elif sampler not in (0, 1, 2):
        raise NotImplementedError("sampler=%d is not a sampler of the three entry points (0-4)" % sampler)
    if neighborhood not in (0, 1):
        raise ValueError("neighborhood should be 0 (grid) or 1 (radius)")
    lib = _load_driver()
    ctx = _context(device)
    s = _settings(lib, sizes, threshold, conf, spatial_coherence_weight, max_iters, min_iters, sampler, neighborhood,
                  neighborhood_size, lo_number, eps_in_cell_size, seed, use_degensac, f_bundle_adjustment, prob, sampler_variance,
                  scoring, e_bundle_adjustment, pnp_bundle_adjustment)
    model = np.zeros(12 if problem == _cabi.PNP else 9)
    mask = np.zeros(len(pts), dtype=np.uint8)
    stats = RunStats()
    if intrinsics is None:
        n_inl = lib.gcr_find_model(ctx.h, problem, pts.ctypes.data, len(pts), C.byref(s), model.ctypes.data,
                                   mask.ctypes.data, C.byref(stats))
    else:
        K1, K2 = intrinsics
        n_inl = lib.gcr_find_essential(ctx.h, pts.ctypes.data, len(pts), K1.ctypes.data, K2.ctypes.data, C.byref(s),
                                       model.ctypes.data, mask.ctypes.data, C.byref(stats))
    if n_inl < 0:
        raise _cabi.GcrError("gcr_find_model failed (%d): %s" % (n_inl, ctx.lib.gcr_last_error(ctx.h).decode()))
    last_statistics = _tls.last_statistics = stats.as_dict()
    mask = mask.astype(bool)
    if n_inl == 0:
        return None, mask                                                   # bindings.cpp:973-975
    return model.reshape(3, -1), mask


def findHomography(correspondences, h1, w1, h2, w2, probabilities=(), threshold=1.0, conf=0.99,
                   spatial_coherence_weight=0.975, max_iters=10000, min_iters=50, use_sprt=True,
                   min_inlier_ratio_for_sprt=0.00001, sampler=1, neighborhood=0, neighborhood_size=4.0,
                   use_space_partitioning=True, lo_number=50, sampler_variance=0.1, solver=0, seed=1, device=0, scoring=0):
    """py/bindings.cpp:859-983, 1158-1178 -> findHomography_ (py/gcransac_python.cpp:4221-4550).  Returns (H 3x3 or
    None, mask bool[N]); the mask marks points within 1.5 x threshold (MSAC's truncated threshold)."""
    if solver != 0:
        raise NotImplementedError("solver=%d: only point correspondences (solver 0) are on this path" % solver)
    if use_space_partitioning and neighborhood != 0:
        # py/gcransac_python.cpp:4282-4286: "Space Partitioning only works with Grid neighorbood yet. Thus, setting
        # neighborhood_id = 0." -- the flag itself has no effect here, its side effect on the graph is kept
        neighborhood = 0
    return _find(_cabi.H, correspondences, (h1, w1, h2, w2), threshold, conf, spatial_coherence_weight, max_iters,
                 min_iters, sampler, neighborhood, neighborhood_size, lo_number, True, seed, device, probabilities=probabilities,
                 sampler_variance=sampler_variance, scoring=scoring)


def findFundamentalMatrix(correspondences, h1, w1, h2, w2, probabilities=(), threshold=1.0, conf=0.99,
                          spatial_coherence_weight=0.0, max_iters=10000, min_iters=50, use_sprt=True,
                          min_inlier_ratio_for_sprt=0.00001, sampler=1, neighborhood=1, neighborhood_size=20.0,
                          lo_number=50, sampler_variance=0.1, solver=0, seed=1, device=0, use_degensac=True,
                          bundle_adjustment=True, scoring=0):
    """py/bindings.cpp:279-400, 1003-1022 -> findFundamentalMatrix_ (py/gcransac_python.cpp:1600-1921) with the
    reference's default estimator (types.h:78-80): 7-point minimal solver, non-minimal fits = 8-point + LM bundle
    adjustment (gcr_refit_f_ba), DEGENSAC (fundamental_estimator.h:341-572), and the post-run bundle adjustment on the
    inliers (py/gcransac_python.cpp:1832-1887).  `use_degensac=False` / `bundle_adjustment=False` (not reference
    arguments) switch those parts off (linear 8-point non-minimal fits, no polish)."""
    if solver != 0:
        raise NotImplementedError("solver=%d: only point correspondences (solver 0) are on this path" % solver)
    return _find(_cabi.F, correspondences, (h1, w1, h2, w2), threshold, conf, spatial_coherence_weight, max_iters,
                 min_iters, sampler, neighborhood, neighborhood_size, lo_number, False, seed, device,
                 use_degensac=use_degensac, f_bundle_adjustment=3 if bundle_adjustment else 0, probabilities=probabilities,
                 sampler_variance=sampler_variance, scoring=scoring)


def findEssentialMatrix(correspondences, K1, K2, h1, w1, h2, w2, probabilities=(), threshold=1.0, conf=0.99,
                        spatial_coherence_weight=0.1, max_iters=10000, min_iters=50, use_sprt=True,
                        min_inlier_ratio_for_sprt=0.01, sampler=1, neighborhood=1, neighborhood_size=20.0, lo_number=50,
                        sampler_variance=0.1, solver=0, seed=1, device=0, scoring=0, bundle_adjustment=True):
    """py/bindings.cpp:710-857, 1091-1112 -> findEssentialMatrix_ (py/gcransac_python.cpp:2519-2855): 5-point (Nister)
    minimal solver, Stewenius non-minimal refits, MSAC on the K-normalised points with threshold / mean focal length,
    grid neighbourhood on the pixel coordinates, and the final relative-pose bundle adjustment on the inliers
    (:2766-2821; gcr_refit_e_ba).  `bundle_adjustment=False` (not a reference argument) returns GC-RANSAC's own result."""
    if solver != 0:
        raise NotImplementedError("solver=%d: only point correspondences (solver 0) are on this path" % solver)
    K1 = np.ascontiguousarray(K1, dtype=np.float64)
    K2 = np.ascontiguousarray(K2, dtype=np.float64)
    if K1.shape != (3, 3):
        raise ValueError("K1 shape should be [3x3]")                        # bindings.cpp:752-754
    if K2.shape != (3, 3):
        raise ValueError("K2 shape should be [3x3]")
    return _find(_cabi.E, correspondences, (h1, w1, h2, w2), threshold, conf, spatial_coherence_weight, max_iters,
                 min_iters, sampler, neighborhood, neighborhood_size, lo_number, False, seed, device, intrinsics=(K1, K2),
                 probabilities=probabilities, sampler_variance=sampler_variance, scoring=scoring,
                 e_bundle_adjustment=1 if bundle_adjustment else 0)


def find6DPose(correspondences, probabilities=(), threshold=0.001, conf=0.99, spatial_coherence_weight=0.975, max_iters=10000,
               min_iters=50, use_sprt=True, min_inlier_ratio_for_sprt=0.00001, sampler=1, neighborhood=0, neighborhood_size=8.0,
               lo_number=50, sampler_variance=0.1, seed=1, device=0, scoring=0, bundle_adjustment=True):
    """py/bindings.cpp:203-277, 1058-1072 -> find6DPose_ (py/gcransac_python.cpp:1026-1272): absolute pose [R | t] from
    2D-3D matches, rows (u, v, X, Y, Z) with K-normalised image points.  GCRANSAC<DefaultPnPEstimator> = P3P minimal solver
    (Lambda-Twist, estimators/solver_p3p.h) + PnPBundleAdjustment non-minimal fits; the neighbourhood (id 0) is the radius
    graph over the 5-column rows (FLANN in the reference, the exact graph here); after the run the non-minimal fit on the
    inliers replaces the pose (:1227-1238; `bundle_adjustment=False`, not a reference argument, skips it).
    Returns (pose 3x4 or None, mask bool[N]).

    The reference's non-minimal solver has no defined behaviour to match (its LM accumulator is constructed with the loss
    where its camera belongs, loops over all rows while indexing the sample and multiplies by weights passed as nullptr;
    its initialiser is OpenCV's EPnP, outside the tree) -- stated here as: normalised DLT on the index set (>= 6 points)
    + the LM of refine_pnp over the index set (gcr_refit with GCR_PNP)."""
    if neighborhood != 0:
        # py/gcransac_python.cpp:1093-1098: "Unknown neighborhood-graph identifier ... accepted values are 0 (FLANN-based)"
        return None, np.zeros(len(correspondences), dtype=bool)
    if sampler == 2:
        return None, np.zeros(len(correspondences), dtype=bool)            # :1113-1120 P-NAPSAC is not usable for PnP
    return _find(_cabi.PNP, correspondences, (0, 0, 0, 0), threshold, conf, spatial_coherence_weight, max_iters, min_iters,
                 sampler, 1, neighborhood_size, lo_number, False, seed, device, probabilities=probabilities,
                 sampler_variance=sampler_variance, scoring=scoring, pnp_bundle_adjustment=1 if bundle_adjustment else 0)


# ------------------------------------------------------------------------------------------------ batches ----
# P image pairs per call.  The reference has no such entry point (its counterpart is one problem per
# host core); here the pairs sit on the GPU together and advance in lock-step -- host/gcr_batch_engine.h -- and every
# pair returns what the single-pair function returns for it with seed seeds[p].
last_batch_statistics = None


def _find_batch(problem, correspondences, sizes, threshold, conf, spatial_coherence_weight, max_iters, min_iters, sampler,
                neighborhood, neighborhood_size, lo_number, eps_in_cell_size, seeds, device, seats, engines, intrinsics=None,
                use_degensac=False, f_bundle_adjustment=0, scoring=0, e_bundle_adjustment=0):
    global last_batch_statistics
    pts = np.ascontiguousarray(correspondences, dtype=np.float64)
    min_n = {_cabi.H: 4, _cabi.F: 7, _cabi.E: 5}[problem]
    if pts.ndim != 3 or pts.shape[2] != 4 or pts.shape[1] < min_n:
        raise ValueError("correspondences should be an array with dims [pairs,n,4], n>=%d" % min_n)
    if sampler not in (0, 1, 2):
        raise NotImplementedError("sampler=%d needs per-pair probabilities: use the single-pair function" % sampler)
    if neighborhood not in (0, 1):
        raise ValueError("neighborhood should be 0 (grid) or 1 (radius)")
    P, n = pts.shape[0], pts.shape[1]
    if seeds is None:
        seeds = np.arange(1, P + 1, dtype=np.uint64)
    seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
    if seeds.shape != (P,):
        raise ValueError("seeds should have one entry per pair")
    lib = _load_driver()
    s = _settings(lib, sizes, threshold, conf, spatial_coherence_weight, max_iters, min_iters, sampler, neighborhood,
                  neighborhood_size, lo_number, eps_in_cell_size, 0, use_degensac, f_bundle_adjustment, None, 0.1,
                  scoring, e_bundle_adjustment, 0)
    models = np.zeros((P, 3, 3))
    masks = np.zeros((P, n), dtype=np.uint8)
    counts = np.zeros(P, dtype=np.int32)
    stats = (RunStats * P)()
    if intrinsics is None:
        r = lib.gcr_find_model_batch(device, problem, pts.ctypes.data, n, P, C.byref(s), seeds.ctypes.data, seats, engines,
                                     models.ctypes.data, masks.ctypes.data, counts.ctypes.data, C.cast(stats, C.c_void_p))
    else:
        K1, K2 = intrinsics
        per_pair = 1 if K1.ndim == 3 else 0
        r = lib.gcr_find_essential_batch(device, pts.ctypes.data, n, P, K1.ctypes.data, K2.ctypes.data, per_pair, C.byref(s),
                                         seeds.ctypes.data, seats, engines, models.ctypes.data, masks.ctypes.data,
                                         counts.ctypes.data, C.cast(stats, C.c_void_p))
    if r < 0:
        raise _cabi.GcrError("gcr_find_model_batch failed (%d): %s" % (r, lib.gcr_driver_last_error().decode()))
    last_batch_statistics = stats
    return models, masks.astype(bool), counts


def findHomographyBatch(correspondences, h1, w1, h2, w2, threshold=1.0, conf=0.99, spatial_coherence_weight=0.975,
                        max_iters=10000, min_iters=50, use_sprt=True, min_inlier_ratio_for_sprt=0.00001, sampler=1, neighborhood=0,
                        neighborhood_size=4.0, use_space_partitioning=True, lo_number=50, seeds=None, device=0, scoring=0,
                        seats=0, engines=0):
    """findHomography for P pairs of n correspondences each ([P][n][4]): returns (H [P][3][3], masks bool [P][n],
    inlier counts [P]; count 0 = no model, where findHomography returns None).  seeds: one sample-stream seed per pair
    (default 1..P); pair p's result equals findHomography(correspondences[p], ..., seed=seeds[p])."""
    if use_space_partitioning and neighborhood != 0:
        neighborhood = 0                                                    # py/gcransac_python.cpp:4282-4286
    return _find_batch(_cabi.H, correspondences, (h1, w1, h2, w2), threshold, conf, spatial_coherence_weight, max_iters, min_iters,
                       sampler, neighborhood, neighborhood_size, lo_number, True, seeds, device, seats, engines, scoring=scoring)


def findFundamentalMatrixBatch(correspondences, h1, w1, h2, w2, threshold=1.0, conf=0.99, spatial_coherence_weight=0.0,
                               max_iters=10000, min_iters=50, use_sprt=True, min_inlier_ratio_for_sprt=0.00001, sampler=1,
                               neighborhood=1, neighborhood_size=20.0, lo_number=50, seeds=None, device=0, use_degensac=True,
                               bundle_adjustment=True, scoring=0, seats=0, engines=0):
    """findFundamentalMatrix (the reference's default pipeline: BA refits, DEGENSAC, post-run BA) for P pairs ([P][n][4])."""
    return _find_batch(_cabi.F, correspondences, (h1, w1, h2, w2), threshold, conf, spatial_coherence_weight, max_iters, min_iters,
                       sampler, neighborhood, neighborhood_size, lo_number, False, seeds, device, seats, engines,
                       use_degensac=use_degensac, f_bundle_adjustment=3 if bundle_adjustment else 0, scoring=scoring)


def findEssentialMatrixBatch(correspondences, K1, K2, h1, w1, h2, w2, threshold=1.0, conf=0.99, spatial_coherence_weight=0.1,
                             max_iters=10000, min_iters=50, use_sprt=True, min_inlier_ratio_for_sprt=0.01, sampler=1,
                             neighborhood=1, neighborhood_size=20.0, lo_number=50, seeds=None, device=0, scoring=0,
                             bundle_adjustment=True, seats=0, engines=0):
    """findEssentialMatrix for P pairs ([P][n][4] pixel correspondences); K1 / K2: 3 x 3 for all pairs or [P][3][3]."""
    K1 = np.ascontiguousarray(K1, dtype=np.float64)
    K2 = np.ascontiguousarray(K2, dtype=np.float64)
    if K1.shape != K2.shape or K1.shape[-2:] != (3, 3) or K1.ndim not in (2, 3):
        raise ValueError("K1 / K2 shape should be [3x3] or [pairs,3,3]")
    return _find_batch(_cabi.E, correspondences, (h1, w1, h2, w2), threshold, conf, spatial_coherence_weight, max_iters, min_iters,
                       sampler, neighborhood, neighborhood_size, lo_number, False, seeds, device, seats, engines,
                       intrinsics=(K1, K2), scoring=scoring, e_bundle_adjustment=1 if bundle_adjustment else 0)


def batch_launch_count():
    """Kernels launched so far by the batch engines of this process."""
    return int(_load_driver().gcr_batch_launch_count())


def release_batch_engines():
    """Frees the GPU contexts / buffers the batch calls keep between calls."""
    _load_driver().gcr_batch_engines_release()
