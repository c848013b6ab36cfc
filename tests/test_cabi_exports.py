"""The C-ABI library loads and exports every symbol include/gcr/gcr.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest
from conftest import ROOT


def test_library_exports_every_declared_symbol():
    from pygcransac import _cabi
    lib = _cabi.load_library()
    header = open(os.path.join(ROOT, "include", "gcr", "gcr.h")).read()
    declared = set(re.findall(r"\b(gcr_[a-z0-9_]+)\s*\(", header))
    declared -= {"gcr_ctx", "gcr_status"}
    assert declared == set(_cabi.EXPORTS), declared ^ set(_cabi.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.gcr_version()
    assert lib.gcr_max_models(_cabi.F) == 3 and lib.gcr_sample_size(_cabi.E) == 5 and lib.gcr_model_len(_cabi.PNP) == 12


def test_key_roundtrip_is_host_callable():
    from pygcransac import _cabi
    lib = _cabi.load_library()
    import numpy as np
    for score, idx in ((0.0, 0), (12.5, 77), (-1.0, 5), (3.0e8, 2 ** 31)):
        bits = np.float32(score).view(np.uint32)
        ordered = (~bits & 0xFFFFFFFF) if bits & 0x80000000 else (bits | 0x80000000)
        key = (int(ordered) << 32) | (0xFFFFFFFF - idx)
        assert lib.gcr_key_score(key) == np.float32(score) and lib.gcr_key_index(key) == idx


def test_no_cpu_fallback_without_gpu():
    import torch
    from pygcransac import _cabi
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_cabi.GcrError):
        _cabi.Context(0)


def test_driver_library_exports_every_declared_symbol():
    """libgcr_driver.so (host/gcr_driver.h: the batched driver behind the Python and C++ entry points) exports what its
    header declares, and its settings structure is the one pygcransac mirrors (same size, same defaults)."""
    import ctypes as C
    import pygcransac
    path = os.path.join(ROOT, "graph-cut-ransac_b200", "host", "_ref", "libgcr_driver.so")
    if not os.path.exists(path):
        pytest.skip("host driver not built (needs the reference tree for its BK max-flow)")
    from pygcransac import _cabi
    _cabi.load_library()
    lib = C.CDLL(path)
    header = open(os.path.join(ROOT, "graph-cut-ransac_b200", "host", "gcr_driver.h")).read()
    header += open(os.path.join(ROOT, "graph-cut-ransac_b200", "host", "gcr_batch_engine.h")).read()   # the multi-pair calls
    declared = set(re.findall(r"\b(gcr_[a-z0-9_]+)\s*\(", header)) - {"gcr_ctx"}
    assert {"gcr_find_model", "gcr_find_essential", "gcr_run_settings_default", "gcr_find_model_batch", "gcr_find_essential_batch",
            "gcr_batch_engines_release", "gcr_driver_last_error"} <= declared
    for name in declared:
        assert hasattr(lib, name), name
    s = pygcransac.RunSettings()
    lib.gcr_run_settings_default.argtypes = [C.POINTER(pygcransac.RunSettings)]
    C.memset(C.byref(s), 0xFF, C.sizeof(s))               # every byte must be written by the default initialiser
    lib.gcr_run_settings_default(C.byref(s))
    assert (s.threshold, s.confidence, s.batch, s.sampler, s.scoring, s.sampler_variance) == (2.0, 0.95, 1024, 0, 0, 0.1)
    assert not s.point_probabilities and list(s.image_sizes) == [0.0] * 4 and s.neighborhood_radius == 20.0


def test_batch_engine_fibers_without_a_gpu():
    """The lock-step driver's host-side concurrency core (host/gcr_batch_engine_impl.h: fibers on their own stacks, a run
    queue served by a few OS threads, requests completed by a dispatcher thread, the PARKED / READY hand-over that lets a
    fiber resume on any thread exactly once) -- exercised with plain tokens instead of GPU requests."""
    import ctypes as C
    path = os.path.join(ROOT, "graph-cut-ransac_b200", "host", "_ref", "libgcr_driver.so")
    if not os.path.exists(path):
        pytest.skip("host driver not built (needs the reference tree for its BK max-flow)")
    from pygcransac import _cabi
    _cabi.load_library()
    lib = C.CDLL(path)
    lib.gcr_fiber_selftest.restype = C.c_longlong
    lib.gcr_fiber_selftest.argtypes = [C.c_int, C.c_int, C.c_int]
    for fibers, threads, rounds in ((1, 1, 10), (64, 4, 200), (300, 8, 50), (7, 16, 1000)):
        assert lib.gcr_fiber_selftest(fibers, threads, rounds) == fibers * rounds, (fibers, threads, rounds)
    assert lib.gcr_fiber_selftest(0, 1, 1) < 0


def test_reference_citations_point_into_the_reference_tree():
    """Every `file.h:line[-line]` citation of a reference file in this repository's sources and documents lies inside that
    file (tools/check_citations.py); only where the reference tree exists."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not os.path.isdir(os.path.join(os.environ.get("GCR_REFERENCE_DIR", "/root/reference"), "src", "pygcransac", "include")):
        import pytest
        pytest.skip("no reference tree")
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "check_citations.py")], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
