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


def test_ctypes_struct_mirrors_match_the_header_layout(tmp_path):
    """sizeof / offsetof of nlpf_desc and nlpf_opts as the C compiler sees them in include/nlpfilter.h against
    the ctypes mirrors in _capi.py (the binding INTEGRATION.md shows relies on exactly this layout)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no C compiler")
    fields_d = [f[0] for f in _capi.CDesc._fields_]
    fields_o = [f[0] for f in _capi.COpts._fields_]
    src = ['#include <stdio.h>', '#include <stddef.h>', '#include "nlpfilter.h"', 'int main(void) {',
           '  printf("desc %zu\\n", sizeof(nlpf_desc)); printf("opts %zu\\n", sizeof(nlpf_opts));']
    src += ['  printf("d.%s %%zu\\n", offsetof(nlpf_desc, %s));' % (f, f) for f in fields_d]
    src += ['  printf("o.%s %%zu\\n", offsetof(nlpf_opts, %s));' % (f, f) for f in fields_o]
    src += ['  return 0; }']
    c = tmp_path / "layout.c"
    c.write_text("\n".join(src))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    assert int(out["desc"]) == ctypes.sizeof(_capi.CDesc) and int(out["opts"]) == ctypes.sizeof(_capi.COpts)
    for f in fields_d:
        assert int(out["d." + f]) == getattr(_capi.CDesc, f).offset, f
    for f in fields_o:
        assert int(out["o." + f]) == getattr(_capi.COpts, f).offset, f
