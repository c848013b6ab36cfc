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
