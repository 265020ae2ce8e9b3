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


def test_damping_schedule_defaults_agree_between_header_and_oracle():
    """NLPF_LAM0 / NLPF_LAM_JUMP / NLPF_LAM_JUMP_INDEF / NLPF_LAM_DEC_FIRST in include/nlpfilter.h (what the library
    substitutes for nlpf_opts fields left at 0) are the constants oracle/nls.py uses; the ctypes mirror of nlpf_opts
    carries every field of the header's struct, in order."""
    import os
    import re
    from nlp_filter_b200 import _capi, desc as D
    from oracle import nls
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "nlpfilter.h")).read()
    get = lambda name: float(re.search(r"#define\s+%s\s+([0-9.eE+-]+)" % name, hdr).group(1))
    assert get("NLPF_LAM0") == nls.LAM0 and get("NLPF_LAM_JUMP") == nls.LAM_JUMP
    assert get("NLPF_LAM_JUMP_INDEF") == nls.LAM_JUMP_INDEF and get("NLPF_LAM_DEC_FIRST") == nls.LAM_DEC_FIRST
    assert nls.damping_schedule(D.SolveOpts()) == (nls.LAM0, nls.LAM_JUMP, nls.LAM_JUMP_INDEF, nls.LAM_DEC_FIRST)
    assert nls.damping_schedule(D.SolveOpts(lam0=1e-3, lam_jump=0.5, lam_jump_indef=2.0, lam_dec_first=7.0)) == (1e-3, 0.5, 2.0, 7.0)
    body = re.search(r"typedef struct nlpf_opts \{(.*?)\} nlpf_opts;", hdr, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if decl:
            names += [n.strip() for n in decl.split(None, 1)[1].split(",")]
    assert names == [f[0] for f in _capi.COpts._fields_], (names, [f[0] for f in _capi.COpts._fields_])
    o = D.SolveOpts(mu0_warm=1e-6)
    assert o.for_window(0) is o and o.for_window(3).mu0 == 1e-6 and o.for_window(3).mu0_warm == 1e-6
    assert D.SolveOpts().for_window(5).mu0 == 0.1
