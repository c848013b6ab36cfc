"""The device solver headers compiled for the HOST (tests/host_harness/device_solvers_host.cpp: CUDA qualifiers defined
away; the ten lanes of a 5-point group = ten threads + a barrier) against the oracle, so the CPU-only tier checks the
kernels' arithmetic too: 7-point (folded full-pivot search) and the group-cooperative 5-point solver."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
from pygcransac import synth

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("host_harness") / "libdevice_solvers_host.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I",
                           os.path.join(ROOT, "graph-cut-ransac_b200", "csrc"),
                           os.path.join(HERE, "host_harness", "device_solvers_host.cpp"), "-o", so, "-lpthread"])
    return C.CDLL(so)


def _models(fn, *args, slots):
    out = np.zeros(slots * 9)
    valid = np.zeros(slots, dtype=np.uint8)
    n = fn(*[a.ctypes.data_as(C.c_void_p) for a in args], out.ctypes.data_as(C.c_void_p), valid.ctypes.data_as(C.c_void_p))
    return n, out.reshape(slots, 3, 3), valid.astype(bool)


def test_f7_device_code_matches_oracle(host, oracle):
    rng = np.random.default_rng(5)
    n_models = 0
    for seed, transform in ((1, lambda p: p), (2, np.round), (3, lambda p: (p - 512) / 1000.0)):
        pts = np.ascontiguousarray(transform(synth.epipolar_pair(400, 0.4, seed)[0]))
        for _ in range(400):
            s = np.ascontiguousarray(rng.choice(400, 7, replace=False), dtype=np.uint32)
            _, got, valid = _models(host.host_f7, pts, s, slots=3)
            want = oracle.estimate_f7(pts, s)                  # 7-pt + oriented constraint, ascending roots
            got = got[valid]
            assert len(got) == len(want), (seed, s)
            for g, w in zip(got, want):
                assert np.abs(g - w).max() <= 1e-9 * np.abs(w).max()   # same statements; libm on both sides here
            n_models += len(want)
    assert n_models > 200


def test_e5_group_device_code_matches_oracle(host, oracle):
    rng = np.random.default_rng(6)
    n_models = 0
    for seed, outliers in ((1, 0.0), (2, 0.5), (3, 0.85)):
        px, _, _, Kc, _ = synth.epipolar_pair(300, outliers, seed)
        pts = oracle.normalize_correspondences(px, Kc, Kc)
        for _ in range(150):
            s = rng.choice(300, 5, replace=False).astype(np.uint32)
            p5 = np.ascontiguousarray(pts[s])
            n_ok, got, valid = _models(host.host_e5_group, p5, slots=10)
            assert n_ok == valid.sum() >= 0
            want = oracle.estimate_e5(pts, s)                  # 5-pt (Nister) + oriented constraint
            got = got[valid]
            assert len(got) == len(want), (seed, s)
            for g, w in zip(got, want):
                assert np.array_equal(g, w)                    # bit-exact: same arithmetic order, only + - * / sqrt
            n_models += len(want)
    assert n_models > 300
    for p5 in (np.zeros((5, 4)), np.ones((5, 4))):             # rank-deficient samples: no models, no hang
        n_ok, got, valid = _models(host.host_e5_group, p5, slots=10)
        assert n_ok == 0 and not valid.any() and not got.any()


def _same_up_to_sign(a, b, tol):
    a, b = a / np.linalg.norm(a), b / np.linalg.norm(b)
    return min(np.abs(a - b).max(), np.abs(a + b).max()) <= tol


def test_stewenius_device_code_matches_oracle(host, oracle):
    """csrc/solve_e_stewenius.cuh (constraints, 10 x 10 elimination, action matrix, Hessenberg-QR eigenvalues, inverse-iteration
    eigenvectors: what k_refit_e's thread runs) on the HOST, fed the 4-D null-space basis of A^T A from the oracle's
    column-pivoted Householder QR, against the oracle's own estimateModelNonminimal on the same index sets."""
    rng = np.random.default_rng(11)
    host.host_stewenius_from_basis.restype = C.c_int
    checked = 0
    for seed, outliers in ((1, 0.0), (2, 0.3)):
        px, _, _, Kc, inl = synth.epipolar_pair(400, outliers, seed)
        pts = oracle.normalize_correspondences(px, Kc, Kc)
        pool = np.flatnonzero(inl)
        for size in (6, 12, 35, 120):
            for _ in range(6):
                s = np.sort(rng.choice(pool, size, replace=False)).astype(np.uint32)
                P = pts[s]
                A = np.stack([P[:, 0] * P[:, 2], P[:, 0] * P[:, 3], P[:, 0], P[:, 1] * P[:, 2], P[:, 1] * P[:, 3], P[:, 1],
                              P[:, 2], P[:, 3], np.ones(len(P))], axis=1)
                Q = oracle.colpiv_q(A.T @ A)
                N = np.ascontiguousarray(Q[:, 5:].T)                      # N[b][j] = Q[j][5 + b] (k_refit.cu colpiv_householder_q_last4_9)
                E = np.zeros(90)
                count = host.host_stewenius_from_basis(N.ctypes.data_as(C.c_void_p), E.ctypes.data_as(C.c_void_p))
                want = oracle.refit_e(pts, s)
                assert count == len(want), (seed, size, count, len(want))
                got = E[: 9 * count].reshape(count, 3, 3)
                for w in want:                                            # same model set (order by eigenvalue on both sides, but
                    assert any(_same_up_to_sign(g, w, 1e-7) for g in got)   # compared as sets: A^T A is summed in another order)
                checked += count
    assert checked > 100


def test_p3p_device_code_matches_oracle(host, oracle):
    """csrc/solve_p3p.cuh (Lambda-Twist) on the host against the oracle's restatement of estimators/solver_p3p.h."""
    rng = np.random.default_rng(12)
    host.host_p3p.restype = C.c_int
    n_models = 0
    for seed in (1, 2, 3):
        pts5 = np.ascontiguousarray(synth.pnp_scene(300, 0.3, seed)[0])
        for _ in range(200):
            s = np.ascontiguousarray(rng.choice(300, 3, replace=False), dtype=np.uint32)
            P = np.zeros(48)
            nm = host.host_p3p(pts5.ctypes.data_as(C.c_void_p), s.ctypes.data_as(C.c_void_p), P.ctypes.data_as(C.c_void_p))
            want = oracle.solve_p3p(pts5, s)
            got = P[: 12 * nm].reshape(nm, 3, 4)
            got = got[np.isfinite(got).all(axis=(1, 2))]
            assert len(got) == len(want), (seed, s, len(got), len(want))
            for g, w in zip(got, want):
                assert np.abs(g - w).max() <= 1e-9 * max(1.0, np.abs(w).max())
            n_models += len(want)
    assert n_models > 300
