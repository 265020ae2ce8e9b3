"""The C-ABI library loads and exports every symbol include/nlpfilter.h declares (no compute)."""
import ctypes
import os
import re

import pytest

from nlp_filter_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "nlpfilter.h")).read()
    return sorted(set(re.findall(r"\b(nlpf_[a-z0-9_]+)\s*\(", src)))


def test_library_is_built_and_exports_header_symbols():
    if not os.path.exists(_capi.LIB_PATH):
        from nlp_filter_b200 import build
        build.build()
    L = ctypes.CDLL(_capi.LIB_PATH)
    names = _declared()
    assert len(names) >= 14
    for n in names:
        assert hasattr(L, n), "libnlpfilter.so does not export %s" % n
    assert set(_capi.EXPORTS) == set(names)


def test_version_and_loud_failure_without_gpu():
    L = _capi.lib()
    assert L.nlpf_version() == 100
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from nlp_filter_b200 import engine, shapes
    with pytest.raises(_capi.NlpfError):
        engine.WindowProblem(shapes.vdp_mhe())          # no CPU fallback: must raise, not compute
    with pytest.raises(_capi.NlpfError):
        engine.ls_solve([[[0.0, 0, 0]]], [[1.0]], [1])
    # the C entry point itself reports the missing device
    cdesc, keep = _capi.make_cdesc(shapes.vdp_mhe())
    h = ctypes.c_void_p()
    rc = L.nlpf_problem_create(ctypes.byref(cdesc), ctypes.byref(h))
    assert rc == -3 and b"no CUDA device" in L.nlpf_last_error()
