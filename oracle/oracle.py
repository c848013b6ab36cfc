"""ctypes bindings for the CPU oracle (TEST INFRASTRUCTURE ONLY).

`Oracle()` loads oracle/libgcr_oracle.so -- the self-contained FP64 restatement of the reference's hot path
(oracle/gcr_oracle.h cites the reference file:line each function follows).  `Ref()` loads
oracle/_ref/libgcr_ref.so -- the reference's own headers compiled verbatim through the container shims (only
buildable where /root/reference exists; the prebuilt .so travels to the GPU box).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
H, F, E, PNP = 0, 1, 2, 3
MSAC, MAGSAC = 0, 1

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")


def build_oracle(force=False):
    so = os.path.join(HERE, "libgcr_oracle.so")
    srcs = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cpp", ".h")) and not f.startswith("ref_")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["bash", os.path.join(HERE, "build_oracle.sh")], stdout=subprocess.DEVNULL)
    return so


def build_ref(force=False):
    """Build oracle/_ref/libgcr_ref.so when the reference tree is present; returns the path or None."""
    so = os.path.join(HERE, "_ref", "libgcr_ref.so")
    ref_dir = os.environ.get("GCR_REFERENCE_DIR", "/root/reference")
    if os.path.isdir(ref_dir):
        srcs = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cpp", ".h"))]
        if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            subprocess.check_call(["bash", os.path.join(HERE, "build_ref.sh")], stdout=subprocess.DEVNULL)
    return so if os.path.exists(so) else None


def build_ref_solvers(force=False):
    """Build oracle/_ref/libgcr_ref_solvers.so (the reference's solver / estimator headers compiled verbatim over
    oracle/shim_solvers) when the reference tree is present; returns the path or None."""
    so = os.path.join(HERE, "_ref", "libgcr_ref_solvers.so")
    ref_dir = os.environ.get("GCR_REFERENCE_DIR", "/root/reference")
    if os.path.isdir(ref_dir):
        srcs = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cpp", ".h"))]
        srcs += [os.path.join(HERE, "shim_solvers", "Eigen", "Eigen"), os.path.join(HERE, "shim_solvers", "unsupported", "Eigen", "Polynomials")]
        if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            subprocess.check_call(["bash", os.path.join(HERE, "build_ref_solvers.sh")], stdout=subprocess.DEVNULL)
    return so if os.path.exists(so) else None


def _as_pts(pts):
    return np.ascontiguousarray(pts, dtype=np.float64)


def _round(fn, problem, scoring, pts, seed, first, count, thr):
    pts = _as_pts(pts)
    s, idx, n = C.c_double(), C.c_uint64(), C.c_uint64()
    model = np.zeros(12 if problem == PNP else 9)
    fn(problem, scoring, pts, len(pts), seed, first, count, thr, C.byref(s), C.byref(idx), C.byref(n), model)
    return s.value, idx.value, n.value, model


class Oracle:
    def __init__(self):
        self.lib = lib = C.CDLL(build_oracle())
        lib.gcro_sqres.restype = C.c_double
        lib.gcro_sqres.argtypes = [C.c_int, _dp, _dp]
        lib.gcro_gamma_lut.restype = C.c_double
        lib.gcro_gamma_lut.argtypes = [C.c_int]
        for name in ("gcro_score_msac", "gcro_score_magsac"):
            fn = getattr(lib, name)
            fn.restype = None
            fn.argtypes = [C.c_int, _dp, C.c_int, _dp, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.c_void_p]
        lib.gcro_score_batch.restype = None
        lib.gcro_score_batch.argtypes = [C.c_int, C.c_int, _dp, C.c_int, _dp, C.c_int, C.c_double, _dp, _u64p]
        lib.gcro_residuals.restype = None
        lib.gcro_residuals.argtypes = [C.c_int, _dp, C.c_int, _dp, _dp]
        lib.gcro_philox4x32.restype = None
        lib.gcro_philox4x32.argtypes = [_u32p, _u32p, _u32p]
        lib.gcro_sample_uniform.restype = None
        lib.gcro_sample_uniform.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, _u32p]
        lib.gcro_prosac_growth.restype = None
        lib.gcro_prosac_growth.argtypes = [C.c_uint32, C.c_int, C.c_uint64, _u64p]
        lib.gcro_prosac_subset_size.restype = C.c_uint32
        lib.gcro_prosac_subset_size.argtypes = [_u64p, C.c_uint32, C.c_int, C.c_uint64]
        lib.gcro_importance_cdf.restype = None
        lib.gcro_importance_cdf.argtypes = [_dp, C.c_uint32, _dp]
        lib.gcro_sample_importance.restype = None
        lib.gcro_sample_importance.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, _dp, _u32p]
        lib.gcro_sample_prosac.restype = None
        lib.gcro_sample_prosac.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, _u64p, C.c_uint64, _u32p]
        lib.gcro_h_valid_sample.restype = C.c_int
        lib.gcro_h_valid_sample.argtypes = [_dp, _u32p]
        lib.gcro_solve_h4.restype = C.c_int
        lib.gcro_solve_h4.argtypes = [_dp, C.c_int, C.c_void_p, _dp]
        lib.gcro_h_valid_model.restype = C.c_int
        lib.gcro_h_valid_model.argtypes = [_dp]
        lib.gcro_normalize_points.restype = None
        lib.gcro_normalize_points.argtypes = [_dp, C.c_int, _u32p, C.c_int, _dp, _dp, _dp]
        lib.gcro_refine_f_lm.restype = C.c_int
        lib.gcro_refine_f_lm.argtypes = [_dp, C.c_int, C.c_void_p, _dp, C.c_int]
        for name in ("gcro_refit_h", "gcro_refit_f8", "gcro_refit_f_ba", "gcro_polish_f_ba"):
            fn = getattr(lib, name)
            fn.restype = C.c_int
            fn.argtypes = [_dp, _u32p, C.c_int, C.c_void_p, _dp]
        for name in ("gcro_solve_f7", "gcro_estimate_f7", "gcro_solve_e5_nister", "gcro_estimate_e5", "gcro_solve_p3p"):
            fn = getattr(lib, name)
            fn.restype = C.c_int
            fn.argtypes = [_dp, _u32p, _dp]
        lib.gcro_degensac_sample_homography.restype = C.c_int
        lib.gcro_degensac_sample_homography.argtypes = [_dp, _dp, _u32p, C.c_double, _dp]
        for name in ("gcro_solve_f_plane_parallax", "gcro_estimate_f_plane_parallax"):
            fn = getattr(lib, name)
            fn.restype = C.c_int
            fn.argtypes = [_dp, _u32p, _dp, _dp]
        lib.gcro_solve_e5_stewenius.restype = C.c_int
        lib.gcro_solve_e5_stewenius.argtypes = [_dp, _u32p, C.c_int, C.c_void_p, _dp]
        lib.gcro_refit_e.restype = C.c_int
        lib.gcro_refit_e.argtypes = [_dp, _u32p, C.c_int, C.c_void_p, _dp]
        lib.gcro_test_colpiv_q.restype = None
        lib.gcro_test_colpiv_q.argtypes = [_dp, C.c_int, _dp]
        lib.gcro_test_eigenvalues.restype = C.c_int
        lib.gcro_test_eigenvalues.argtypes = [_dp, C.c_int, _dp, _dp]
        lib.gcro_test_eigenvector.restype = None
        lib.gcro_test_eigenvector.argtypes = [_dp, C.c_int, C.c_double, _dp]
        lib.gcro_f_orientation_valid.restype = C.c_int
        lib.gcro_f_orientation_valid.argtypes = [_dp, _dp, _u32p, C.c_int]
        lib.gcro_f_valid_model_ratio.restype = C.c_int
        lib.gcro_f_valid_model_ratio.argtypes = [_dp, _dp, _u32p, C.c_int, C.c_int, C.c_double, C.c_double]
        lib.gcro_normalize_correspondences.restype = None
        lib.gcro_normalize_correspondences.argtypes = [_dp, C.c_int, _dp, _dp, _dp]
        lib.gcro_grid_build.restype = C.c_uint64
        lib.gcro_grid_build.argtypes = [_dp, C.c_int, C.c_int, _dp, C.c_uint64, _u64p, _u32p, _u32p, C.POINTER(C.c_uint32)]
        lib.gcro_iteration_number.restype = C.c_uint64
        lib.gcro_iteration_number.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_double]
        lib.gcro_round.restype = None
        lib.gcro_round.argtypes = [C.c_int, C.c_int, _dp, C.c_int, C.c_uint64, C.c_uint64, C.c_int, C.c_double,
                                   C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), _dp]

    def round(self, problem, scoring, pts, seed, first, count, thr):
        """Fixed-K pass (sample -> solve -> score -> best); returns (score, index, models_scored, model)."""
        return _round(self.lib.gcro_round, problem, scoring, pts, seed, first, count, thr)

    # ---- scoring ----
    def sqres(self, problem, pt, model):
        return self.lib.gcro_sqres(problem, _as_pts(pt), _as_pts(model).ravel())

    def residuals(self, problem, pts, model):
        pts = _as_pts(pts)
        out = np.empty(len(pts))
        self.lib.gcro_residuals(problem, pts, len(pts), _as_pts(model).ravel(), out)
        return out

    def score(self, problem, scoring, pts, model, thr, want_inliers=False):
        pts = _as_pts(pts)
        s, n = C.c_double(), C.c_uint64()
        idx = np.empty(len(pts), dtype=np.uint32) if want_inliers else None
        fn = self.lib.gcro_score_msac if scoring == MSAC else self.lib.gcro_score_magsac
        fn(problem, pts, len(pts), _as_pts(model).ravel(), thr, C.byref(s), C.byref(n),
           idx.ctypes.data if want_inliers else None)
        if want_inliers:
            return s.value, n.value, idx[: n.value].copy()
        return s.value, n.value

    def score_batch(self, problem, scoring, pts, models, thr):
        pts = _as_pts(pts)
        models = _as_pts(models).reshape(len(models), -1)
        scores = np.empty(len(models))
        ninl = np.empty(len(models), dtype=np.uint64)
        self.lib.gcro_score_batch(problem, scoring, pts, len(pts), models, len(models), thr, scores, ninl)
        return scores, ninl

    def gamma_lut(self, k):
        return self.lib.gcro_gamma_lut(int(k))

    # ---- samplers ----
    def philox(self, ctr, key):
        out = np.empty(4, dtype=np.uint32)
        self.lib.gcro_philox4x32(np.asarray(ctr, dtype=np.uint32), np.asarray(key, dtype=np.uint32), out)
        return out

    def sample_uniform(self, seed, first, count, n, m, stream=0):
        out = np.empty((count, m), dtype=np.uint32)
        for i in range(count):
            self.lib.gcro_sample_uniform(seed, first + i, stream, n, m, out[i])
        return out

    def prosac_growth(self, n, m, conv=100000):
        g = np.empty(n, dtype=np.uint64)
        self.lib.gcro_prosac_growth(n, m, conv, g)
        return g

    def prosac_subset_size(self, growth, n, m, k):
        return self.lib.gcro_prosac_subset_size(growth, n, m, k)

    def sample_prosac(self, seed, first_k, count, n, m, growth, conv=100000):
        out = np.empty((count, m), dtype=np.uint32)
        for i in range(count):
            self.lib.gcro_sample_prosac(seed, first_k + i, n, m, growth, conv, out[i])
        return out

    def importance_cdf(self, probabilities):
        p = np.ascontiguousarray(probabilities, dtype=np.float64)
        cdf = np.empty(len(p))
        self.lib.gcro_importance_cdf(p, len(p), cdf)
        return cdf

    def sample_importance(self, seed, first, count, m, cdf):
        out = np.empty((count, m), dtype=np.uint32)
        cdf = np.ascontiguousarray(cdf, dtype=np.float64)
        for i in range(count):
            self.lib.gcro_sample_importance(seed, first + i, len(cdf), m, cdf, out[i])
        return out

    # ---- solvers ----
    def h_valid_sample(self, pts, sample):
        return bool(self.lib.gcro_h_valid_sample(_as_pts(pts), np.asarray(sample, dtype=np.uint32)))

    def solve_h4(self, pts, sample):
        out = np.empty(9)
        s = np.ascontiguousarray(sample, dtype=np.uint32)
        pts = _as_pts(pts)
        ok = self.lib.gcro_solve_h4(pts, pts.shape[1], s.ctypes.data, out)
        return out.reshape(3, 3) if ok else None

    def h_valid_model(self, Hm):
        return bool(self.lib.gcro_h_valid_model(_as_pts(Hm).ravel()))

    def refit(self, problem, pts, sample, weights=None):
        out = np.empty(9)
        s = np.ascontiguousarray(sample, dtype=np.uint32)
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
        fn = self.lib.gcro_refit_h if problem == H else self.lib.gcro_refit_f8
        ok = fn(_as_pts(pts), s, len(s), None if w is None else w.ctypes.data, out)
        return out.reshape(3, 3) if ok else None

    def normalize_points(self, pts, sample):
        """normalizePoints (fundamental_estimator.h:637-734 / homography_estimator.h:213-310): the selected points in the
        Hartley frame and the two 3 x 3 transforms."""
        sample = np.ascontiguousarray(sample, dtype=np.uint32)
        out = np.zeros((len(sample), 4))
        T1, T2 = np.zeros(9), np.zeros(9)
        self.lib.gcro_normalize_points(_as_pts(pts), 4, sample, len(sample), out.reshape(-1), T1, T2)
        return out, T1.reshape(3, 3), T2.reshape(3, 3)

    def refine_f_lm(self, pts_in_frame, F, weights=None, max_iterations=25):
        """relative_pose/bundle.cpp:253-339 on points already in the working frame; returns (F, iterations)."""
        out = np.ascontiguousarray(F, dtype=np.float64).reshape(9).copy()
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
        it = self.lib.gcro_refine_f_lm(_as_pts(pts_in_frame), len(pts_in_frame), None if w is None else w.ctypes.data, out,
                                       max_iterations)
        return out.reshape(3, 3), it

    def refit_f_ba(self, pts, sample, weights=None, polish=False):
        """The reference's default non-minimal F fit (8-point + LM in the Hartley frame), or with polish=True the
        post-run bundle adjustment of py/gcransac_python.cpp:1851-1858 (pixel frame)."""
        out = np.zeros(9)
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
        fn = self.lib.gcro_polish_f_ba if polish else self.lib.gcro_refit_f_ba
        ok = fn(_as_pts(pts), np.ascontiguousarray(sample, dtype=np.uint32), len(sample), None if w is None else w.ctypes.data, out)
        return out.reshape(3, 3) if ok else None

    def _multi(self, fn, pts, sample, maxn, rows, cols):
        out = np.empty(maxn * rows * cols)
        cnt = fn(_as_pts(pts), np.ascontiguousarray(sample, dtype=np.uint32), out)
        return out[: cnt * rows * cols].reshape(cnt, rows, cols).copy()

    def solve_f7(self, pts, sample):
        return self._multi(self.lib.gcro_solve_f7, pts, sample, 3, 3, 3)

    def estimate_f7(self, pts, sample):
        return self._multi(self.lib.gcro_estimate_f7, pts, sample, 3, 3, 3)

    def degensac_sample_homography(self, F, pts, sample, sq_h_threshold=4.0):
        """fundamental_estimator.h:341-500: None, or the plane homography the H-degenerate part of the sample implies."""
        Hm = np.zeros(9)
        ok = self.lib.gcro_degensac_sample_homography(_as_pts(F).ravel(), _as_pts(pts),
                                                      np.ascontiguousarray(sample, dtype=np.uint32), sq_h_threshold, Hm)
        return Hm.reshape(3, 3) if ok else None

    def solve_f_plane_parallax(self, pts, sample, Hm, oriented=False):
        Fm = np.zeros(9)
        fn = self.lib.gcro_estimate_f_plane_parallax if oriented else self.lib.gcro_solve_f_plane_parallax
        ok = fn(_as_pts(pts), np.ascontiguousarray(sample, dtype=np.uint32), _as_pts(Hm).ravel(), Fm)
        return Fm.reshape(3, 3) if ok else None

    def solve_e5(self, pts, sample):
        return self._multi(self.lib.gcro_solve_e5_nister, pts, sample, 10, 3, 3)

    def estimate_e5(self, pts, sample):
        return self._multi(self.lib.gcro_estimate_e5, pts, sample, 10, 3, 3)

    def solve_e5_stewenius(self, pts, sample, weights=None):
        out = np.empty(90)
        s = np.ascontiguousarray(sample, dtype=np.uint32)
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
        cnt = self.lib.gcro_solve_e5_stewenius(_as_pts(pts), s, len(s), None if w is None else w.ctypes.data, out)
        return out[: cnt * 9].reshape(cnt, 3, 3).copy()

    def refit_e(self, pts, sample, weights=None):
        """EssentialMatrixEstimator::estimateModelNonminimal: normalised, E33 >= 0 models [cnt, 3, 3]."""
        out = np.empty(90)
        s = np.ascontiguousarray(sample, dtype=np.uint32)
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
        cnt = self.lib.gcro_refit_e(_as_pts(pts), s, len(s), None if w is None else w.ctypes.data, out)
        return out[: cnt * 9].reshape(cnt, 3, 3).copy()

    def refine_e_lm(self, pts, sample, E, weights=None):
        """motion_from_essential (cheirality on sample[0]) + the 5-dof LM of refine_relpose for every surviving pose:
        refined essential matrices [cnt <= 4, 3, 3].  weights: by position in `sample`."""
        out = np.empty(36)
        s = np.ascontiguousarray(sample, dtype=np.uint32)
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
        self.lib.gcro_refine_e_lm.restype = C.c_int
        self.lib.gcro_refine_e_lm.argtypes = [_dp, _u32p, C.c_int, C.c_void_p, _dp, _dp]
        cnt = self.lib.gcro_refine_e_lm(_as_pts(pts), s, len(s), None if w is None else w.ctypes.data,
                                        np.ascontiguousarray(E, dtype=np.float64).reshape(9), out)
        return out[: cnt * 9].reshape(cnt, 3, 3).copy()

    def polish_e_ba(self, pts, inliers, weights=None):
        """The post-run BA of findEssentialMatrix_ (EssentialMatrixBundleAdjustmentSolver::estimateModel on the inliers):
        models [cnt <= 40, 3, 3].  weights: by position in `inliers`."""
        out = np.empty(360)
        p = _as_pts(pts)
        s = np.ascontiguousarray(inliers, dtype=np.uint32)
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
        self.lib.gcro_polish_e_ba.restype = C.c_int
        self.lib.gcro_polish_e_ba.argtypes = [_dp, C.c_int, _u32p, C.c_int, C.c_void_p, _dp]
        cnt = self.lib.gcro_polish_e_ba(p, len(p), s, len(s), None if w is None else w.ctypes.data, out)
        return out[: cnt * 9].reshape(cnt, 3, 3).copy()

    def refit_pnp(self, pts5, sample, weights=None):
        """Non-minimal absolute pose: normalised DLT + refine_pnp's LM; P [3, 4] or None.  weights: by position."""
        out = np.empty(12)
        p = np.ascontiguousarray(pts5, dtype=np.float64)
        s = np.ascontiguousarray(sample, dtype=np.uint32)
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
        self.lib.gcro_refit_pnp.restype = C.c_int
        self.lib.gcro_refit_pnp.argtypes = [_dp, _u32p, C.c_int, C.c_void_p, _dp]
        cnt = self.lib.gcro_refit_pnp(p, s, len(s), None if w is None else w.ctypes.data, out)
        return out.reshape(3, 4).copy() if cnt else None

    def refine_pnp_lm(self, pts5, sample, P, weights=None):
        out = np.ascontiguousarray(P, dtype=np.float64).reshape(12).copy()
        p = np.ascontiguousarray(pts5, dtype=np.float64)
        s = np.ascontiguousarray(sample, dtype=np.uint32)
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
        self.lib.gcro_refine_pnp_lm.restype = C.c_int
        self.lib.gcro_refine_pnp_lm.argtypes = [_dp, _u32p, C.c_int, C.c_void_p, _dp]
        it = self.lib.gcro_refine_pnp_lm(p, s, len(s), None if w is None else w.ctypes.data, out)
        return out.reshape(3, 4), it

    def colpiv_q(self, A):
        A = np.ascontiguousarray(A, dtype=np.float64)
        Q = np.empty_like(A)
        self.lib.gcro_test_colpiv_q(A, A.shape[0], Q)
        return Q

    def eigenvalues(self, A):
        A = np.ascontiguousarray(A, dtype=np.float64)
        n = A.shape[0]
        wr, wi = np.empty(n), np.empty(n)
        ok = self.lib.gcro_test_eigenvalues(A, n, wr, wi)
        return bool(ok), wr + 1j * wi

    def eigenvector(self, A, lam):
        A = np.ascontiguousarray(A, dtype=np.float64)
        v = np.empty(A.shape[0])
        self.lib.gcro_test_eigenvector(A, A.shape[0], float(lam), v)
        return v

    def solve_p3p(self, pts5, sample):
        return self._multi(self.lib.gcro_solve_p3p, pts5, sample, 4, 3, 4)

    def normalize_correspondences(self, pts, K1, K2):
        pts = _as_pts(pts)
        out = np.empty_like(pts)
        self.lib.gcro_normalize_correspondences(pts, len(pts), _as_pts(K1), _as_pts(K2), out)
        return out

    # ---- grid / termination ----
    def grid_build(self, pts, cell_sizes, cells_per_axis):
        pts = _as_pts(pts)
        n, dims = pts.shape
        keys = np.empty(n, dtype=np.uint64)
        order = np.empty(n, dtype=np.uint32)
        offs = np.empty(n + 1, dtype=np.uint32)
        nc = C.c_uint32()
        nn = self.lib.gcro_grid_build(pts, n, dims, _as_pts(cell_sizes), cells_per_axis, keys, order, offs, C.byref(nc))
        return dict(keys=keys, sorted_idx=order, cell_offsets=offs[: nc.value + 1].copy(), n_cells=nc.value,
                    neighbor_number=nn)

    def iteration_number(self, inliers, n, m, log_probability):
        return self.lib.gcro_iteration_number(inliers, n, m, log_probability)


class Ref:
    """The reference's own headers compiled verbatim (oracle/_ref/libgcr_ref.so)."""

    SO_NAME = "libgcr_ref.so"

    def __init__(self):
        so = build_ref()
        if so is None:
            raise FileNotFoundError("oracle/_ref/libgcr_ref.so is not built and /root/reference is absent")
        so = os.path.join(os.path.dirname(so), self.SO_NAME)
        if not os.path.exists(so):
            raise FileNotFoundError("%s is not built" % so)
        self.lib = lib = C.CDLL(so)
        lib.ref_score.restype = None
        lib.ref_score.argtypes = [C.c_int, C.c_int, _dp, C.c_int, _dp, C.c_double, C.POINTER(C.c_double),
                                  C.POINTER(C.c_uint64), C.c_void_p]
        lib.ref_gamma_value.restype = C.c_double
        lib.ref_gamma_value.argtypes = [C.c_int]
        lib.ref_gamma_table_size.restype = C.c_int
        lib.ref_prosac.restype = None
        lib.ref_prosac.argtypes = [C.c_int, C.c_int, C.c_uint64, _u64p, C.c_int, _u32p, _u32p]
        lib.ref_grid.restype = C.c_uint64
        lib.ref_grid.argtypes = [_dp, C.c_int, _dp, C.c_uint64, _u64p, _u32p, _u32p, C.POINTER(C.c_uint32)]
        lib.ref_iteration_number.restype = C.c_uint64
        lib.ref_iteration_number.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_double]
        lib.ref_round.restype = None
        lib.ref_round.argtypes = [C.c_int, C.c_int, _dp, C.c_int, C.c_uint64, C.c_uint64, C.c_int, C.c_double,
                                  C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), _dp]

        lib.ref_run.restype = C.c_int
        lib.ref_run.argtypes = [C.c_int, C.c_int, _dp, C.c_int, _dp, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int,
                                C.c_int, C.c_int, C.c_uint64, _dp, np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS"),
                                _u64p, C.POINTER(C.c_double), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double]
        lib.ref_sample_importance.restype = None
        lib.ref_sample_importance.argtypes = [_dp, C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_int, _u32p]
        lib.ref_sample_adaptive_reordering.restype = None
        lib.ref_sample_adaptive_reordering.argtypes = [_dp, C.c_int, C.c_int, C.c_double, C.c_uint64, C.c_int, C.c_longlong, _u32p]
        lib.ref_set_image_sizes.restype = None
        lib.ref_set_image_sizes.argtypes = [C.c_double] * 4
        lib.ref_sample_pnapsac.restype = C.c_int
        lib.ref_sample_pnapsac.argtypes = [_dp, C.c_int, C.c_int, _dp, C.c_uint64, C.c_int, _u32p]
        lib.ref_set_sampler_variance.restype = None
        lib.ref_set_sampler_variance.argtypes = [C.c_double]
        lib.ref_set_point_probabilities.restype = None
        lib.ref_set_point_probabilities.argtypes = [_dp, C.c_int]
        lib.ref_radius_graph.restype = C.c_uint64
        lib.ref_radius_graph.argtypes = [_dp, C.c_int, C.c_double, _u32p, _u32p, C.c_uint64]

    def round(self, problem, scoring, pts, seed, first, count, thr):
        return _round(self.lib.ref_round, problem, scoring, pts, seed, first, count, thr)

    def run(self, problem, pts, cell_sizes, cells_per_axis, threshold, conf=0.99, spatial_weight=0.975, max_iters=10000,
            min_iters=50, lo_number=50, seed=1, scoring_kind=0, grid_pts=None, use_degensac=False, f_ba=0, sampler=0, radius=0.0,
            probabilities=None, sampler_variance=0.1, image_sizes=None):
        """GCRANSAC<>::run (GCRANSAC.h:202-627, verbatim) with the settings of py/gcransac_python.cpp:4498-4504 on the
        counter-based sample stream.  f_ba (F): bit 0 = the BA non-minimal solver (the reference's default estimator),
        bit 1 = the post-run polish of py/gcransac_python.cpp:1832-1887.  sampler: 0 = uniform, 1 = PROSAC (the points
        are taken to be ordered by quality), 3 = the reference's ImportanceSampler over `probabilities`, 4 = its
        AdaptiveReorderingSampler (probabilities / their maximum, `sampler_variance`), 2 = its ProgressiveNapsacSampler<4>
        (layers 16 / 8 / 4 / 2 over `image_sizes` = (w1, h1, w2, h2), on grid_pts when given).  radius > 0: the exact radius neighbourhood (neighborhood=1) instead of the grid.
        scoring_kind 0 = the reference's MSACScoringFunction, 1 = the CUDA adapter
        (RefCuda only).  Returns dict(model, mask, inliers, score, stats)."""
        pts = _as_pts(pts)
        model = np.zeros(12 if problem == PNP else 9)
        mask = np.zeros(len(pts), dtype=np.uint8)
        stats = np.zeros(10, dtype=np.uint64)
        score = C.c_double()
        if sampler == 2:
            self.lib.ref_set_image_sizes(*[float(v) for v in image_sizes])   # (w1, h1, w2, h2)
        if sampler in (3, 4):
            prob = np.ascontiguousarray(probabilities, dtype=np.float64)
            assert len(prob) == len(pts)
            self.lib.ref_set_point_probabilities(prob, len(prob))
            self.lib.ref_set_sampler_variance(sampler_variance)
        n_inl = self.lib.ref_run(problem, scoring_kind, pts, len(pts), _as_pts(cell_sizes), cells_per_axis, threshold,
                                 conf, spatial_weight, max_iters, min_iters, lo_number, seed, model, mask, stats,
                                 C.byref(score), None if grid_pts is None else _as_pts(grid_pts).ctypes.data,
                                 1 if use_degensac else 0, f_ba, sampler, radius)
        if n_inl < 0:
            raise ValueError("ref_run: unsupported problem / scoring kind (%d)" % n_inl)
        names = ("iterations", "graph_cuts", "local_optimizations", "neighbor_number", "accepted", "rejected",
                 "main_samples", "lo_samples", "degensac_runs", "degensac_updates")
        return dict(model=model.reshape(3, -1), mask=mask.astype(bool), inliers=n_inl, score=score.value,
                    stats=dict(zip(names, (int(v) for v in stats))))

    def score(self, problem, scoring, pts, model, thr, want_inliers=False):
        pts = _as_pts(pts)
        s, n = C.c_double(), C.c_uint64()
        idx = np.empty(len(pts), dtype=np.uint32) if want_inliers else None
        self.lib.ref_score(problem, scoring, pts, len(pts), _as_pts(model).ravel(), thr, C.byref(s), C.byref(n),
                           idx.ctypes.data if want_inliers else None)
        if want_inliers:
            return s.value, n.value, idx[: n.value].copy()
        return s.value, n.value

    def gamma_table(self):
        return np.array([self.lib.ref_gamma_value(k) for k in range(self.lib.ref_gamma_table_size())])

    def prosac(self, n, m, conv, calls):
        growth = np.empty(n, dtype=np.uint64)
        sizes = np.empty(calls, dtype=np.uint32)
        last = np.empty(calls, dtype=np.uint32)
        self.lib.ref_prosac(n, m, conv, growth, calls, sizes, last)
        return growth, sizes, last

    def sample_importance(self, probabilities, m, seed, first, count):
        """Samples [first, first + count) of the reference's own ImportanceSampler distribution object
        (samplers/importance_sampler.h:76-78, 134-135) on the counter-based word stream."""
        p = np.ascontiguousarray(probabilities, dtype=np.float64)
        out = np.empty((count, m), dtype=np.uint32)
        self.lib.ref_sample_importance(p, len(p), m, seed, first, count, out)
        return out

    def sample_pnapsac(self, pts, m, image_sizes, seed, count):
        """The first `count` samples of the reference's own ProgressiveNapsacSampler<4> (verbatim; layers 16 / 8 / 4 / 2,
        blending length 0.5 as the drivers construct it) on the counter-based word stream."""
        pts = _as_pts(pts)
        out = np.empty((count, m), dtype=np.uint32)
        rc = self.lib.ref_sample_pnapsac(pts, len(pts), m, np.ascontiguousarray(image_sizes, dtype=np.float64), seed, count, out)
        if rc != 0:
            raise ValueError("ref_sample_pnapsac failed (%d)" % rc)
        return out

    def sample_adaptive_reordering(self, probabilities, m, variance, seed, count, repeat_at=-1):
        """The first `count` samples of the reference's own AdaptiveReorderingSampler (samplers/
        adaptive_reordering_sampler.h, verbatim; rand() redirected to the counter-based stream) driven sample / update /
        sample / ... as GCRANSAC::run drives it."""
        p = np.ascontiguousarray(probabilities, dtype=np.float64)
        out = np.empty((count, m), dtype=np.uint32)
        self.lib.ref_sample_adaptive_reordering(p, len(p), m, variance, seed, count, repeat_at, out)
        return out

    def grid(self, pts, cell_sizes, cells_per_axis):
        pts = _as_pts(pts)
        n = len(pts)
        keys = np.empty(n, dtype=np.uint64)
        order = np.empty(n, dtype=np.uint32)
        offs = np.empty(n + 1, dtype=np.uint32)
        nc = C.c_uint32()
        nn = self.lib.ref_grid(pts, n, _as_pts(cell_sizes), cells_per_axis, keys, order, offs, C.byref(nc))
        return dict(keys=keys, sorted_idx=order, cell_offsets=offs[: nc.value + 1].copy(), n_cells=nc.value,
                    neighbor_number=nn)

    def iteration_number(self, inliers, n, m, log_probability):
        return self.lib.ref_iteration_number(inliers, n, m, log_probability)

    def radius_graph(self, pts, radius):
        """CSR (offsets [n + 1], neighbours) of the exact radius neighbourhood the verbatim driver runs on for radius > 0."""
        pts = _as_pts(pts)
        n = len(pts)
        offsets = np.empty(n + 1, dtype=np.uint32)
        total = self.lib.ref_radius_graph(pts, n, radius, offsets, np.empty(1, dtype=np.uint32), 0)
        nbrs = np.empty(max(int(total), 1), dtype=np.uint32)
        self.lib.ref_radius_graph(pts, n, radius, offsets, nbrs, total)
        return offsets, nbrs[:total]


class RefCuda(Ref):
    """The same verbatim reference driver with gcr::CudaMSACScoringFunction (graph-cut-ransac_b200/host/
    gcr_scoring_function.h) as its _ScoringFunction template argument; links libgcr_b200.so, needs a GPU to run."""
    SO_NAME = "libgcr_ref_cuda.so"
