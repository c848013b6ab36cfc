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
