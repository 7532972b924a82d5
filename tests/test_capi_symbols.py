"""CPU tests of the C-ABI boundary: the library loads, exports every symbol include/roargraph_b200.h declares,
validates arguments, and fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def lib():
    from mysteryann_b200 import build, capi

    build.build()
    return capi.lib()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "roargraph_b200.h")).read()
    return sorted(set(re.findall(r"RG_API\s+[\w\s\*]+?\b(rg_\w+)\s*\(", src)))


def test_header_symbols_exported(lib):
    from mysteryann_b200 import capi

    names = declared_symbols()
    assert len(names) >= 13
    assert sorted(capi.SYMBOLS) == names
    for name in names:
        assert getattr(lib, name) is not None, name


def test_version_and_error_string(lib):
    assert b"sm_100a" in lib.rg_version_string()
    assert isinstance(lib.rg_last_error_string(), bytes)


def test_argument_validation_and_no_cpu_fallback(lib):
    from mysteryann_b200 import capi

    base = np.zeros((4, 8), np.float32)
    off = np.array([0, 1, 2, 3, 4], np.uint64)
    adj = np.array([1, 2, 3, 0], np.uint32)
    with pytest.raises(capi.RoarGraphError) as e:  # dim not a multiple of 8
        capi.Index(np.zeros((4, 6), np.float32), off, adj, 0)
    assert e.value.code == 1 and "multiple of 8" in str(e.value)
    with pytest.raises(capi.RoarGraphError) as e:  # entry point out of range
        capi.Index(base, off, adj, 9)
    assert e.value.code == 1
    if capi.device_count() == 0:
        with pytest.raises(capi.RoarGraphError) as e:
            capi.Index(base, off, adj, 0)
        assert e.value.code == capi.RG_ERR_NO_DEVICE and "no CPU fallback" in str(e.value)
