"""ctypes view of the C ABI in include/gcr/gcr.h (libgcr_b200.so, sm_100a).

There is deliberately no fallback: if the shared library is missing or no CUDA device is present, loading /
context creation raises.  Nothing in here touches the CPU oracle.
"""
import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("GCR_LIB_PATH") or os.path.join(_PKG, "libgcr_b200.so")   # GCR_LIB_PATH: A/B builds of the same ABI

H, F, E, PNP = 0, 1, 2, 3
MSAC, MAGSAC = 0, 1
UNIFORM, PROSAC, IMPORTANCE = 0, 1, 2
WEIGHT_NONE, WEIGHT_POSITIONAL, WEIGHT_BY_POINT = 0, 1, 2
SCORE_EXACT_COUNTS = 1
TUNE_K3_CTAS_PER_SM, TUNE_ROUND_PREFETCH, TUNE_SMEM_CARVEOUT = 0, 1, 2

# every symbol include/gcr/gcr.h declares (tests check the library exports all of them)
BA_HARTLEY, BA_PIXEL = 1, 2

EXPORTS = [
    "gcr_ctx_create", "gcr_ctx_destroy", "gcr_ctx_set_stream", "gcr_sync", "gcr_last_error", "gcr_version",
    "gcr_device_info", "gcr_launch_count", "gcr_malloc", "gcr_free", "gcr_memcpy_h2d", "gcr_memcpy_d2h", "gcr_memset",
    "gcr_points_upload", "gcr_points_set_device", "gcr_points_get_norm", "gcr_sample", "gcr_set_point_probabilities", "gcr_sample_from_pool",
    "gcr_solve_minimal", "gcr_max_models", "gcr_sample_size", "gcr_model_len", "gcr_score", "gcr_score_exact",
    "gcr_score_host", "gcr_score_exact_host", "gcr_argmax", "gcr_key_score", "gcr_key_index", "gcr_refit",
    "gcr_grid_build", "gcr_round", "gcr_set_band_factor", "gcr_set_score_variant", "gcr_profile_enable",
    "gcr_profile_read", "gcr_counters_read", "gcr_round_results", "gcr_getscore_host", "gcr_max_refit_models",
    "gcr_solve_plane_parallax", "gcr_refit_f_ba", "gcr_set_score_trace",
    "gcr_memcpy_d2h_async", "gcr_radius_graph_count", "gcr_radius_graph_fill",
    "gcr_score_exact_pairs", "gcr_refit_pairs",
    "gcr_cache_stats", "gcr_round_winners", "gcr_comm_unique_id", "gcr_comm_init", "gcr_comm_destroy", "gcr_round_sharded",
    "gcr_winners_reduce", "gcr_ubench_fp32", "gcr_points_upload_f32", "gcr_set_tuning",
    "gcr_grid_edges", "gcr_csr_edges", "gcr_labeling_terms", "gcr_score_dropped_async",
    "gcr_refit_e_ba", "gcr_max_e_ba_models",
    "gcr_set_active_pairs", "gcr_score_exact_multi", "gcr_labeling_terms_multi", "gcr_sample_from_pools", "gcr_refit_multi",
    "gcr_host_alloc_pinned", "gcr_host_free_pinned", "gcr_memcpy_pinned_async", "gcr_gather_rows",
    "gcr_set_pair_streams", "gcr_points_replace_pairs",
]


class Winner(C.Structure):
    """gcr_winner (include/gcr/gcr.h): the 128-byte per-pair winner record of gcr_round_winners / gcr_round_sharded."""
    _fields_ = [("score", C.c_double), ("index", C.c_uint64), ("count", C.c_uint32), ("valid", C.c_uint32),
                ("model", C.c_double * 12), ("reserved", C.c_uint64)]


WINNER_DTYPE = np.dtype([("score", "<f8"), ("index", "<u8"), ("count", "<u4"), ("valid", "<u4"), ("model", "<f8", (12,)),
                         ("reserved", "<u8")])
assert C.sizeof(Winner) == 128 and WINNER_DTYPE.itemsize == 128

_lib = None


class GcrError(RuntimeError):
    pass


def load_library():
    """Load libgcr_b200.so; raises if it has not been built (python graph-cut-ransac_b200/build.py)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GcrError("%s is missing: build it with `python graph-cut-ransac_b200/build.py` "
                       "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, i32, u32, u64, f64 = C.c_void_p, C.c_int, C.c_uint32, C.c_uint64, C.c_double
    sig = {
        "gcr_ctx_create": (i32, [i32, C.POINTER(vp)]),
        "gcr_ctx_destroy": (i32, [vp]),
        "gcr_ctx_set_stream": (i32, [vp, vp]),
        "gcr_sync": (i32, [vp]),
        "gcr_last_error": (C.c_char_p, [vp]),
        "gcr_version": (C.c_char_p, []),
        "gcr_device_info": (i32, [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(C.c_size_t)]),
        "gcr_launch_count": (u64, [vp]),
        "gcr_malloc": (i32, [vp, C.c_size_t, C.POINTER(vp)]),
        "gcr_free": (i32, [vp, vp]),
        "gcr_memcpy_h2d": (i32, [vp, vp, vp, C.c_size_t]),
        "gcr_memcpy_d2h": (i32, [vp, vp, vp, C.c_size_t]),
        "gcr_memcpy_d2h_async": (i32, [vp, vp, vp, C.c_size_t]),
        "gcr_memset": (i32, [vp, vp, i32, C.c_size_t]),
        "gcr_points_upload": (i32, [vp, vp, i32, i32, i32]),
        "gcr_points_set_device": (i32, [vp, vp, i32, i32, i32]),
        "gcr_points_upload_f32": (i32, [vp, vp, i32, i32, i32]),
        "gcr_set_tuning": (i32, [vp, i32, i32]),
        "gcr_set_pair_streams": (i32, [vp, vp, vp]),
        "gcr_points_replace_pairs": (i32, [vp, vp, i32, vp]),
        "gcr_score_dropped_async": (i32, [vp, vp]),
        "gcr_refit_e_ba": (i32, [vp, i32, vp, u32, vp, vp, vp]),
        "gcr_max_e_ba_models": (i32, []),
        "gcr_grid_edges": (i32, [vp, vp, vp, u32, vp, vp, u64]),
        "gcr_csr_edges": (i32, [vp, vp, vp, vp, vp, u64]),
        "gcr_labeling_terms": (i32, [vp, i32, i32, vp, f64, f64, vp, u64, vp, vp, vp]),
        "gcr_points_get_norm": (i32, [vp, vp]),
        "gcr_sample": (i32, [vp, i32, i32, u64, i32, u64, vp]),
        "gcr_set_point_probabilities": (i32, [vp, vp, i32]),
        "gcr_sample_from_pool": (i32, [vp, vp, u32, i32, u64, i32, u64, u32, vp]),
        "gcr_solve_minimal": (i32, [vp, i32, vp, i32, vp, vp]),
        "gcr_solve_plane_parallax": (i32, [vp, vp, i32, vp, vp, vp]),
        "gcr_max_models": (i32, [i32]),
        "gcr_max_refit_models": (i32, [i32]),
        "gcr_sample_size": (i32, [i32]),
        "gcr_model_len": (i32, [i32]),
        "gcr_score": (i32, [vp, i32, i32, vp, vp, i32, f64, i32, vp, vp]),
        "gcr_score_exact": (i32, [vp, i32, i32, i32, vp, i32, f64, vp, vp, vp, vp]),
        "gcr_score_host": (i32, [vp, i32, i32, vp, vp, i32, f64, i32, vp, vp]),
        "gcr_score_exact_host": (i32, [vp, i32, i32, i32, vp, i32, f64, vp, vp, vp, vp]),
        "gcr_argmax": (i32, [vp, vp, i32, u32, vp]),
        "gcr_key_score": (C.c_float, [u64]),
        "gcr_key_index": (u32, [u64]),
        "gcr_refit": (i32, [vp, i32, i32, vp, vp, vp, i32, vp, i32, vp, vp]),
        "gcr_refit_f_ba": (i32, [vp, i32, vp, vp, vp, i32, vp, i32, i32, vp, vp]),
        "gcr_grid_build": (i32, [vp, i32, vp, u64, vp, vp, vp, C.POINTER(u32), C.POINTER(u64)]),
        "gcr_radius_graph_count": (i32, [vp, i32, f64, vp]),
        "gcr_radius_graph_fill": (i32, [vp, i32, f64, vp, vp]),
        "gcr_score_exact_pairs": (i32, [vp, i32, i32, vp, f64, vp, vp, vp]),
        "gcr_refit_pairs": (i32, [vp, i32, vp, vp, vp, vp]),
        "gcr_round": (i32, [vp, i32, i32, i32, u64, i32, u64, f64, i32, u32, vp, vp, C.POINTER(u64)]),
        "gcr_round_results": (i32, [vp, vp, vp]),
        "gcr_getscore_host": (i32, [vp, i32, i32, i32, vp, f64, f64, C.POINTER(f64), C.POINTER(u32), vp]),
        "gcr_cache_stats": (i32, [vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]),
        "gcr_round_winners": (i32, [vp, i32, i32, i32, u64, i32, u64, f64, i32, u64, i32, vp, C.POINTER(u64)]),
        "gcr_comm_unique_id": (i32, [vp]),
        "gcr_comm_init": (i32, [vp, vp, i32, i32]),
        "gcr_comm_destroy": (i32, [vp]),
        "gcr_round_sharded": (i32, [vp, i32, i32, i32, u64, u64, u64, f64, i32, i32, vp, C.POINTER(u64)]),
        "gcr_winners_reduce": (i32, [vp, vp, i32, vp]),
        "gcr_ubench_fp32": (i32, [vp, i32, f64, C.POINTER(f64 * 2), C.POINTER(f64)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    lib.gcr_set_band_factor.restype = i32
    lib.gcr_set_band_factor.argtypes = [vp, f64]
    lib.gcr_set_score_trace.restype = i32
    lib.gcr_set_score_trace.argtypes = [vp, vp]
    lib.gcr_set_score_variant.restype = i32
    lib.gcr_set_score_variant.argtypes = [vp, i32]
    lib.gcr_profile_enable.restype = i32
    lib.gcr_profile_enable.argtypes = [vp, i32]
    lib.gcr_profile_read.restype = i32
    lib.gcr_profile_read.argtypes = [vp, C.POINTER(f64), C.POINTER(u64), C.POINTER(u64)]
    lib.gcr_counters_read.restype = i32
    lib.gcr_counters_read.argtypes = [vp, C.POINTER(u64 * 8), i32]
    _lib = lib
    return lib


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return C.c_void_p(a.ctypes.data)


def comm_unique_id():
    """128-byte NCCL unique id for gcr_comm_init (create on rank 0, hand to every rank)."""
    buf = (C.c_char * 128)()
    st = load_library().gcr_comm_unique_id(C.byref(buf))
    if st != 0:
        raise GcrError("gcr_comm_unique_id failed with status %d (is libnccl.so.2 loadable?)" % st)
    return bytes(buf)


class Context:
    """One context per GPU per host thread (include/gcr/gcr.h)."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = C.c_void_p()
        st = self.lib.gcr_ctx_create(device, C.byref(h))
        if st != 0 or not h.value:
            raise GcrError("gcr_ctx_create(device=%d) failed with status %d: no CUDA device / driver "
                           "(this library has no CPU fallback)" % (device, st))
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.lib.gcr_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, st):
        if st != 0:
            raise GcrError("gcr status %d: %s" % (st, self.lib.gcr_last_error(self.h).decode()))

    # ---- plumbing ----
    def set_stream(self, cuda_stream_ptr):
        self.check(self.lib.gcr_ctx_set_stream(self.h, C.c_void_p(cuda_stream_ptr or 0)))

    def sync(self):
        self.check(self.lib.gcr_sync(self.h))

    def device_info(self):
        sm, khz, l2 = C.c_int(), C.c_int(), C.c_size_t()
        self.check(self.lib.gcr_device_info(self.h, C.byref(sm), C.byref(khz), C.byref(l2)))
        return dict(sm_count=sm.value, sm_clock_khz=khz.value, l2_bytes=l2.value)

    def launch_count(self):
        return int(self.lib.gcr_launch_count(self.h))

    def malloc(self, nbytes):
        p = C.c_void_p()
        self.check(self.lib.gcr_malloc(self.h, nbytes, C.byref(p)))
        return p.value

    def free(self, dptr):
        self.check(self.lib.gcr_free(self.h, C.c_void_p(dptr)))

    def to_device(self, arr):
        arr = np.ascontiguousarray(arr)
        d = self.malloc(arr.nbytes)
        self.check(self.lib.gcr_memcpy_h2d(self.h, C.c_void_p(d), _ptr(arr), arr.nbytes))
        return d

    def to_host(self, dptr, shape, dtype):
        out = np.empty(shape, dtype=dtype)
        self.check(self.lib.gcr_memcpy_d2h(self.h, _ptr(out), C.c_void_p(dptr), out.nbytes))
        return out

    def set_band_factor(self, f):
        self.check(self.lib.gcr_set_band_factor(self.h, f))

    def set_tuning(self, knob, value):
        self.check(self.lib.gcr_set_tuning(self.h, knob, value))

    def set_score_variant(self, v):
        self.check(self.lib.gcr_set_score_variant(self.h, v))

    def profile_enable(self, on):
        self.check(self.lib.gcr_profile_enable(self.h, 1 if on else 0))

    def profile_read(self):
        ms, n, ev = C.c_double(), C.c_uint64(), C.c_uint64()
        self.check(self.lib.gcr_profile_read(self.h, C.byref(ms), C.byref(n), C.byref(ev)))
        return ms.value, n.value, ev.value

    def counters(self, reset=False):
        out = (C.c_uint64 * 8)()
        self.check(self.lib.gcr_counters_read(self.h, C.byref(out), 1 if reset else 0))
        return list(out)

    def select_stats(self, reset=False):
        """(FP64 re-scores made by gcr_round's selection, largest relative FP32-vs-FP64 score error seen among them)."""
        c = self.counters(reset)
        return int(c[3]), float(np.array([c[4] & 0xFFFFFFFF], dtype=np.uint32).view(np.float32)[0])

    def ubench_fp32(self, mode=0, min_ms=200.0):
        """Measured FP32 FMA roofline of this device: (sustained TFLOP/s over >= min_ms, best single launch, ms run)."""
        tf, ms = (C.c_double * 2)(), C.c_double()
        self.check(self.lib.gcr_ubench_fp32(self.h, mode, min_ms, C.byref(tf), C.byref(ms)))
        return tf[0], tf[1], ms.value

    def cache_stats(self):
        b, u, n = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self.check(self.lib.gcr_cache_stats(self.h, C.byref(b), C.byref(u), C.byref(n)))
        return dict(blocks=b.value, in_use=u.value, bytes=n.value)

    # ---- winner records / hypothesis-sharded mode ----
    def round_winners_device(self, problem, scoring, sampler, first_index, count, seed, thr, flags, index_base, accumulate,
                             d_winners):
        ev = C.c_uint64()
        self.check(self.lib.gcr_round_winners(self.h, problem, scoring, sampler, first_index, count, seed, thr, flags,
                                              index_base, 1 if accumulate else 0, C.c_void_p(d_winners), C.byref(ev)))
        return ev.value

    def comm_init(self, unique_id, rank, world):
        buf = (C.c_char * 128).from_buffer_copy(bytes(unique_id))
        self.check(self.lib.gcr_comm_init(self.h, C.byref(buf), rank, world))

    def comm_destroy(self):
        self.check(self.lib.gcr_comm_destroy(self.h))

    def round_sharded_device(self, problem, scoring, sampler, first_index, total_samples, seed, thr, flags, chunk_samples,
                             d_winners):
        """gcr_round_sharded with a caller-owned device record array; returns this rank's slot evaluations."""
        ev = C.c_uint64()
        self.check(self.lib.gcr_round_sharded(self.h, problem, scoring, sampler, first_index, total_samples, seed, thr, flags,
                                              chunk_samples, C.c_void_p(d_winners), C.byref(ev)))
        return ev.value

    def round_sharded(self, problem, scoring, sampler, first_index, total_samples, seed, thr, flags=0, chunk_samples=65536):
        """Host-output convenience: the winner records as a numpy structured array [n_pairs] (WINNER_DTYPE)."""
        d = self.malloc(self.n_pairs * 128)
        try:
            ev = self.round_sharded_device(problem, scoring, sampler, first_index, total_samples, seed, thr, flags,
                                           chunk_samples, d)
            return self.to_host(d, (self.n_pairs,), WINNER_DTYPE), ev
        finally:
            self.free(d)

    def round_device(self, problem, scoring, sampler, first_index, count, seed, thr, flags, index_base, d_keys, d_best):
        """gcr_round with caller-owned device outputs (no host copies); returns the slot evaluations."""
        ev = C.c_uint64()
        self.check(self.lib.gcr_round(self.h, problem, scoring, sampler, first_index, count, seed, thr, flags,
                                      index_base, C.c_void_p(d_keys), C.c_void_p(d_best), C.byref(ev)))
        return ev.value

    # ---- points ----
    def points_upload(self, pts, n_pairs=1):
        pts = np.ascontiguousarray(pts, dtype=np.float64)
        cols = pts.shape[-1]
        n = pts.size // cols // n_pairs
        self._pts_keepalive = pts
        self.check(self.lib.gcr_points_upload(self.h, _ptr(pts), n, cols, n_pairs))
        self.n, self.cols, self.n_pairs = n, cols, n_pairs

    def points_upload_f32(self, pts, n_pairs=1):
        pts = np.ascontiguousarray(pts, dtype=np.float32)
        cols = pts.shape[-1]
        n = pts.size // cols // n_pairs
        self._pts_keepalive = pts
        self.check(self.lib.gcr_points_upload_f32(self.h, _ptr(pts), n, cols, n_pairs))
        self.n, self.cols, self.n_pairs = n, cols, n_pairs

    def points_set_device(self, dptr, n, cols, n_pairs=1):
        self.check(self.lib.gcr_points_set_device(self.h, C.c_void_p(dptr), n, cols, n_pairs))
        self.n, self.cols, self.n_pairs = n, cols, n_pairs

    def points_norm(self):
        out = np.empty((self.n_pairs, 8))
        self.check(self.lib.gcr_points_get_norm(self.h, _ptr(out)))
        return out

    # ---- K1 ----
    def sample(self, sampler, m, first_index, count, seed):
        d = self.malloc(self.n_pairs * count * m * 4)
        try:
            self.check(self.lib.gcr_sample(self.h, sampler, m, first_index, count, seed, C.c_void_p(d)))
            return self.to_host(d, (self.n_pairs, count, m), np.uint32)
        finally:
            self.free(d)

    def set_point_probabilities(self, probabilities):
        """The distribution GCR_IMPORTANCE samples draw from (samplers/importance_sampler.h:67-82)."""
        p = np.ascontiguousarray(probabilities, dtype=np.float64).ravel()
        self.check(self.lib.gcr_set_point_probabilities(self.h, _ptr(p), len(p)))

    def sample_from_pool(self, pool, m, first_index, count, seed, stream):
        pool = np.ascontiguousarray(pool, dtype=np.uint32)
        dp = self.to_device(pool)
        d = self.malloc(count * m * 4)
        try:
            self.check(self.lib.gcr_sample_from_pool(self.h, C.c_void_p(dp), len(pool), m, first_index, count, seed,
                                                     stream, C.c_void_p(d)))
            return self.to_host(d, (count, m), np.uint32)
        finally:
            self.free(d)
            self.free(dp)

    # ---- K2 ----
    def solve_minimal(self, problem, samples):
        samples = np.ascontiguousarray(samples, dtype=np.uint32).reshape(self.n_pairs, -1, self.lib.gcr_sample_size(problem))
        count = samples.shape[1]
        mm, ml = self.lib.gcr_max_models(problem), self.lib.gcr_model_len(problem)
        ds = self.to_device(samples)
        dm = self.malloc(self.n_pairs * count * mm * ml * 8)
        dv = self.malloc(self.n_pairs * count * mm)
        try:
            self.check(self.lib.gcr_solve_minimal(self.h, problem, C.c_void_p(ds), count, C.c_void_p(dm), C.c_void_p(dv)))
            return (self.to_host(dm, (self.n_pairs, count, mm, ml), np.float64),
                    self.to_host(dv, (self.n_pairs, count, mm), np.uint8))
        finally:
            for d in (ds, dm, dv):
                self.free(d)

    def solve_plane_parallax(self, samples, homography):
        """DEGENSAC's inner minimal solver: 2-point samples + the plane homography -> F (or invalid)."""
        samples = np.ascontiguousarray(samples, dtype=np.uint32).reshape(self.n_pairs, -1, 2)
        count = samples.shape[1]
        Hm = np.ascontiguousarray(homography, dtype=np.float64).reshape(9)
        ds = self.to_device(samples)
        dm = self.malloc(self.n_pairs * count * 72)
        dv = self.malloc(self.n_pairs * count)
        try:
            self.check(self.lib.gcr_solve_plane_parallax(self.h, C.c_void_p(ds), count, Hm.ctypes.data_as(C.c_void_p),
                                                         C.c_void_p(dm), C.c_void_p(dv)))
            return (self.to_host(dm, (self.n_pairs, count, 9), np.float64), self.to_host(dv, (self.n_pairs, count), np.uint8))
        finally:
            for d in (ds, dm, dv):
                self.free(d)

    # ---- K3 ----
    def score(self, problem, scoring, models, thr, valid=None, flags=0):
        ml = self.lib.gcr_model_len(problem)
        models = np.ascontiguousarray(models, dtype=np.float64).reshape(self.n_pairs, -1, ml)
        k = models.shape[1]
        scores = np.empty((self.n_pairs, k), dtype=np.float32)
        inl = np.empty((self.n_pairs, k), dtype=np.uint32)
        v = None if valid is None else np.ascontiguousarray(valid, dtype=np.uint8)
        self.check(self.lib.gcr_score_host(self.h, problem, scoring, _ptr(models), _ptr(v), k, thr, flags,
                                           _ptr(scores), _ptr(inl)))
        return scores, inl

    def score_exact(self, problem, scoring, models, thr, pair=0, want_inliers=True, want_residuals=False):
        ml = self.lib.gcr_model_len(problem)
        models = np.ascontiguousarray(models, dtype=np.float64).reshape(-1, ml)
        k = len(models)
        scores = np.empty(k)
        counts = np.empty(k, dtype=np.uint32)
        idx = np.empty((k, self.n), dtype=np.uint32) if want_inliers else None
        res = np.empty((k, self.n)) if want_residuals else None
        self.check(self.lib.gcr_score_exact_host(self.h, problem, scoring, pair, _ptr(models), k, thr, _ptr(scores),
                                                 _ptr(counts), _ptr(idx), _ptr(res)))
        out = [scores, counts]
        if want_inliers:
            out.append([idx[j, : counts[j]].copy() for j in range(k)])
        if want_residuals:
            out.append(res)
        return tuple(out)

    # ---- K4 ----
    def argmax(self, scores, index_base=0):
        scores = np.ascontiguousarray(scores, dtype=np.float32).reshape(self.n_pairs, -1)
        ds = self.to_device(scores)
        dk = self.malloc(self.n_pairs * 8)
        try:
            self.check(self.lib.gcr_argmax(self.h, C.c_void_p(ds), scores.shape[1], index_base, C.c_void_p(dk)))
            return self.to_host(dk, (self.n_pairs,), np.uint64)
        finally:
            self.free(ds)
            self.free(dk)

    def key_decode(self, key):
        return float(self.lib.gcr_key_score(int(key))), int(self.lib.gcr_key_index(int(key)))

    # ---- K6 ----
    def refit(self, problem, index_sets, weights=None, weight_mode=WEIGHT_NONE, pair=0, ba_frame=0):
        """ba_frame (F only): 0 = the linear fit; BA_HARTLEY / BA_PIXEL = + the LM bundle adjustment (gcr_refit_f_ba)."""
        sizes = np.array([len(s) for s in index_sets], dtype=np.uint32)
        offsets = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.uint32)
        flat = np.concatenate([np.asarray(s, dtype=np.uint32) for s in index_sets]).astype(np.uint32)
        count = len(index_sets)
        d_flat, d_off, d_sz = self.to_device(flat), self.to_device(offsets), self.to_device(sizes)
        d_w = None if weights is None else self.to_device(np.ascontiguousarray(weights, dtype=np.float64))
        R = self.lib.gcr_max_refit_models(problem)          # 1 (H, F, PnP) or 10 (E: Stewenius returns up to 10 models)
        ml = self.lib.gcr_model_len(problem)
        dm = self.malloc(count * R * ml * 8)
        dok = self.malloc(count * R)
        try:
            if ba_frame:
                self.check(self.lib.gcr_refit_f_ba(self.h, pair, C.c_void_p(d_flat), C.c_void_p(d_off), C.c_void_p(d_sz), count,
                                                   C.c_void_p(d_w) if d_w else None, weight_mode, ba_frame, C.c_void_p(dm),
                                                   C.c_void_p(dok)))
            else:
                self.check(self.lib.gcr_refit(self.h, problem, pair, C.c_void_p(d_flat), C.c_void_p(d_off), C.c_void_p(d_sz),
                                              count, C.c_void_p(d_w) if d_w else None, weight_mode, C.c_void_p(dm),
                                              C.c_void_p(dok)))
            if R == 1:
                return self.to_host(dm, (count, 3, ml // 3), np.float64), self.to_host(dok, (count,), np.uint8)
            return self.to_host(dm, (count, R, 3, 3), np.float64), self.to_host(dok, (count, R), np.uint8)
        finally:
            for d in (d_flat, d_off, d_sz, dm, dok):
                self.free(d)
            if d_w:
                self.free(d_w)

    def refit_e_ba(self, index_set, weights=None, pair=0):
        """gcr_refit_e_ba: the relative-pose bundle adjustment on one index set; (models [40, 3, 3], ok [40])."""
        idx = np.ascontiguousarray(index_set, dtype=np.uint32)
        d_idx = self.to_device(idx)
        d_w = None if weights is None else self.to_device(np.ascontiguousarray(weights, dtype=np.float64))
        R = self.lib.gcr_max_e_ba_models()
        dm, dok = self.malloc(R * 72), self.malloc(R)
        try:
            self.check(self.lib.gcr_refit_e_ba(self.h, pair, C.c_void_p(d_idx), len(idx), C.c_void_p(d_w) if d_w else None,
                                               C.c_void_p(dm), C.c_void_p(dok)))
            return self.to_host(dm, (R, 3, 3), np.float64), self.to_host(dok, (R,), np.uint8)
        finally:
            for d in (d_idx, dm, dok):
                self.free(d)
            if d_w:
                self.free(d_w)

    # ---- K7 ----
    def grid_build(self, cell_sizes, cells_per_axis, pair=0):
        cs = np.ascontiguousarray(cell_sizes, dtype=np.float64)
        n = self.n
        dk, ds, do = self.malloc(n * 8), self.malloc(n * 4), self.malloc((n + 1) * 4)
        nc, nn = C.c_uint32(), C.c_uint64()
        try:
            self.check(self.lib.gcr_grid_build(self.h, pair, _ptr(cs), cells_per_axis, C.c_void_p(dk), C.c_void_p(ds),
                                               C.c_void_p(do), C.byref(nc), C.byref(nn)))
            return dict(keys=self.to_host(dk, (n,), np.uint64), sorted_idx=self.to_host(ds, (n,), np.uint32),
                        cell_offsets=self.to_host(do, (n + 1,), np.uint32)[: nc.value + 1].copy(), n_cells=nc.value,
                        neighbor_number=nn.value)
        finally:
            for d in (dk, ds, do):
                self.free(d)

    def grid_edges(self, cell_sizes, cells_per_axis, pair=0):
        """labeling()'s edge list of the grid neighbourhood, built on the device: (edges [E, 2], edge_offsets [n + 1])."""
        cs = np.ascontiguousarray(cell_sizes, dtype=np.float64)
        n = self.n
        dk, ds, do, deo = self.malloc(n * 8), self.malloc(n * 4), self.malloc((n + 1) * 4), self.malloc((n + 1) * 4)
        nc, nn = C.c_uint32(), C.c_uint64()
        de = None
        try:
            self.check(self.lib.gcr_grid_build(self.h, pair, _ptr(cs), cells_per_axis, C.c_void_p(dk), C.c_void_p(ds),
                                               C.c_void_p(do), C.byref(nc), C.byref(nn)))
            de = self.malloc(max(nn.value, 1) * 8)
            self.check(self.lib.gcr_grid_edges(self.h, C.c_void_p(ds), C.c_void_p(do), nc.value, C.c_void_p(deo), C.c_void_p(de),
                                               nn.value))
            return self.to_host(de, (nn.value, 2), np.uint32), self.to_host(deo, (n + 1,), np.uint32)
        finally:
            for d in (dk, ds, do, deo, de):
                if d:
                    self.free(d)

    def csr_edges(self, offsets, neighbors):
        """labeling()'s edge list of a CSR neighbourhood (radius graph): the (i, j > i) entries, on the device."""
        offsets = np.ascontiguousarray(offsets, dtype=np.uint32)
        neighbors = np.ascontiguousarray(neighbors, dtype=np.uint32)
        E = len(neighbors) // 2
        do, dn = self.to_device(offsets), self.to_device(neighbors if len(neighbors) else np.zeros(1, np.uint32))
        deo, de = self.malloc((self.n + 1) * 4), self.malloc(max(E, 1) * 8)
        try:
            self.check(self.lib.gcr_csr_edges(self.h, C.c_void_p(do), C.c_void_p(dn), C.c_void_p(deo), C.c_void_p(de), E))
            return self.to_host(de, (E, 2), np.uint32), self.to_host(deo, (self.n + 1,), np.uint32)
        finally:
            for d in (do, dn, deo, de):
                self.free(d)

    def labeling_terms(self, problem, model, thr, lam, edges, pair=0):
        """(unary [n, 2], pairwise [E]) of labeling()'s energy for one model."""
        ml = self.lib.gcr_model_len(problem)
        edges = np.ascontiguousarray(edges, dtype=np.uint32).reshape(-1, 2)
        E = len(edges)
        dm = self.to_device(np.ascontiguousarray(model, dtype=np.float64).reshape(ml))
        de = self.to_device(edges if E else np.zeros((1, 2), np.uint32))
        du, dp, dd = self.malloc(self.n * 16), self.malloc(max(E, 1) * 8), self.malloc(self.n * 8)
        try:
            self.check(self.lib.gcr_labeling_terms(self.h, problem, pair, C.c_void_p(dm), thr, lam, C.c_void_p(de), E,
                                                   C.c_void_p(du), C.c_void_p(dp), C.c_void_p(dd)))
            return (self.to_host(du, (self.n, 2), np.float64), self.to_host(dp, (E,), np.float64),
                    self.to_host(dd, (self.n,), np.float64))
        finally:
            for d in (dm, de, du, dp, dd):
                self.free(d)

    def radius_graph(self, radius, pair=0):
        """The exact radius neighbourhood graph (neighborhood=1): CSR offsets [n + 1] and neighbours (ascending)."""
        dc = self.malloc(self.n * 4)
        try:
            self.check(self.lib.gcr_radius_graph_count(self.h, pair, radius, C.c_void_p(dc)))
            counts = self.to_host(dc, (self.n,), np.uint32)
            offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.uint32)
            total = int(offsets[-1])
            do, dn = self.to_device(offsets[:-1].copy()), self.malloc(max(total, 1) * 4)
            try:
                self.check(self.lib.gcr_radius_graph_fill(self.h, pair, radius, C.c_void_p(do), C.c_void_p(dn)))
                return offsets, self.to_host(dn, (total,), np.uint32)
            finally:
                self.free(do)
                self.free(dn)
        finally:
            self.free(dc)

    def final_refit_pairs(self, problem, d_models, thr, d_idx, d_counts, d_scores, d_refit, d_ok):
        """Device-resident tail of the batched configuration: inliers of every pair's model, then one refit per pair."""
        self.check(self.lib.gcr_score_exact_pairs(self.h, problem, MSAC, C.c_void_p(d_models), thr, C.c_void_p(d_scores),
                                                  C.c_void_p(d_counts), C.c_void_p(d_idx)))
        self.check(self.lib.gcr_refit_pairs(self.h, problem, C.c_void_p(d_idx), C.c_void_p(d_counts), C.c_void_p(d_refit),
                                            C.c_void_p(d_ok)))

    def round_results(self):
        """FP64 score and exact inlier count of each pair's winner of the last round."""
        sc = np.empty(self.n_pairs)
        cn = np.empty(self.n_pairs, dtype=np.uint32)
        self.check(self.lib.gcr_round_results(self.h, _ptr(sc), _ptr(cn)))
        return sc, cn

    # ---- fused round ----
    def round(self, problem, scoring, sampler, first_index, count, seed, thr, flags=0, index_base=0):
        ml = self.lib.gcr_model_len(problem)
        dk, dm = self.malloc(self.n_pairs * 8), self.malloc(self.n_pairs * ml * 8)
        ev = C.c_uint64()
        try:
            self.check(self.lib.gcr_round(self.h, problem, scoring, sampler, first_index, count, seed, thr, flags,
                                          index_base, C.c_void_p(dk), C.c_void_p(dm), C.byref(ev)))
            return (self.to_host(dk, (self.n_pairs,), np.uint64), self.to_host(dm, (self.n_pairs, ml), np.float64),
                    ev.value)
        finally:
            self.free(dk)
            self.free(dm)
