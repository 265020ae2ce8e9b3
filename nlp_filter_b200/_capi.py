"""ctypes binding of libnlpfilter.so (include/nlpfilter.h).  PyTorch appears only as the owner
of device memory and streams: every call passes raw device pointers.

There is no CPU fallback: if the library is missing or no CUDA device is visible the calls
raise (`NlpfError`), they never route anywhere else.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NLPF_LIB") or os.path.join(_HERE, "libnlpfilter.so")

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int_p = ctypes.POINTER(ctypes.c_int)


class NlpfError(RuntimeError):
    pass


class CDesc(ctypes.Structure):
    _fields_ = [("K", ctypes.c_int), ("n", ctypes.c_int), ("m", ctypes.c_int), ("nv", ctypes.c_int),
                ("ng", ctypes.c_int), ("dyn_id", ctypes.c_int), ("loss", ctypes.c_int),
                ("has_prior", ctypes.c_int), ("nb", ctypes.c_int), ("E", ctypes.c_int), ("M", ctypes.c_int),
                ("Ldata", ctypes.c_int), ("T", ctypes.c_double), ("delta", ctypes.c_double),
                ("dyn_par", c_double_p), ("A", c_double_p), ("Bm", c_double_p), ("c0", c_double_p),
                ("cmap", c_int_p), ("ucol", c_int_p), ("Dt", c_double_p), ("cw", c_double_p),
                ("cq", c_double_p), ("qinv", c_double_p), ("pinv", c_double_p), ("snode", c_double_p),
                ("bidx", c_int_p), ("blo", c_double_p), ("bhi", c_double_p), ("Phi", c_double_p),
                ("rtype", c_int_p), ("repoch", c_int_p), ("ridx", c_int_p), ("roff", c_int_p)]


class COpts(ctypes.Structure):
    _fields_ = [("max_iter", ctypes.c_int), ("max_stage_iter", ctypes.c_int), ("tol_step", ctypes.c_double),
                ("lam0", ctypes.c_double), ("mu0", ctypes.c_double), ("mu_min", ctypes.c_double),
                ("bound_push", ctypes.c_double), ("tau_ftb", ctypes.c_double),
                ("lam_jump", ctypes.c_double), ("lam_jump_indef", ctypes.c_double), ("lam_dec_first", ctypes.c_double),
                ("mu0_warm", ctypes.c_double)]


EXPORTS = ("nlpf_version", "nlpf_last_error", "nlpf_device_count", "nlpf_problem_create",
           "nlpf_problem_destroy", "nlpf_problem_nz", "nlpf_assemble", "nlpf_factor_solve", "nlpf_solve",
           "nlpf_sample", "nlpf_ekf_run", "nlpf_ls_solve", "nlpf_ls_vel", "nlpf_fp64_probe", "nlpf_dmma_probe", "nlpf_blocktri_solve",
           "nlpf_problem_smem_bytes", "nlpf_ecef2enu", "nlpf_enu2lla", "nlpf_solve_chain",
           "nlpf_problem_has_chain_kernel", "nlpf_problem_profile", "nlpf_ms_gather", "nlpf_ms_blocks", "nlpf_ekf_run_general", "nlpf_defects", "nlpf_problem_set_user_dynamics")

_lib = None


def lib():
    """Load the shared library (once).  Raises NlpfError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NlpfError("libnlpfilter.so is not built (%s); run `python -c 'import __graft_entry__ as g; g.build()'`"
                        % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp = ctypes.c_void_p
    L.nlpf_version.restype = ctypes.c_int
    L.nlpf_last_error.restype = ctypes.c_char_p
    L.nlpf_device_count.restype = ctypes.c_int
    L.nlpf_problem_create.argtypes = [ctypes.POINTER(CDesc), ctypes.POINTER(vp)]
    L.nlpf_problem_destroy.argtypes = [vp]
    L.nlpf_problem_nz.argtypes = [vp]
    L.nlpf_problem_smem_bytes.argtypes = [vp]
    L.nlpf_assemble.argtypes = [vp, ctypes.c_int, vp, vp, vp, vp, vp, ctypes.c_double, vp, vp, vp, vp]
    L.nlpf_factor_solve.argtypes = [vp, ctypes.c_int, vp, vp, ctypes.c_double, vp, vp, vp]
    L.nlpf_solve.argtypes = [vp, ctypes.c_int, vp, vp, vp, vp, vp, ctypes.POINTER(COpts), vp, vp, vp, vp]
    L.nlpf_problem_set_user_dynamics.argtypes = [vp, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p]
    L.nlpf_defects.argtypes = [vp, ctypes.c_int, vp, vp, vp, vp]
    L.nlpf_sample.argtypes = [vp, ctypes.c_int, vp, ctypes.c_int, vp, vp, vp]
    L.nlpf_ekf_run.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, vp, vp, vp,
                               ctypes.c_double, ctypes.c_double, vp, vp, vp]
    L.nlpf_ekf_run_general.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, vp, vp, vp, vp,
                                       ctypes.c_double, vp, vp, ctypes.c_int, vp]
    L.nlpf_ls_solve.argtypes = [ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, vp, ctypes.c_int, ctypes.c_double, vp, vp]
    L.nlpf_ls_vel.argtypes = [ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, vp, vp, vp]
    L.nlpf_blocktri_solve.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, vp, ctypes.c_int, vp]
    L.nlpf_problem_has_chain_kernel.argtypes = [vp]
    L.nlpf_ms_gather.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, vp]
    L.nlpf_ms_blocks.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.nlpf_problem_profile.argtypes = [vp, ctypes.c_int, ctypes.POINTER(ctypes.c_ulonglong)]
    L.nlpf_solve_chain.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, vp, ctypes.POINTER(COpts), vp, vp, vp, vp, vp]
    L.nlpf_enu2lla.argtypes = [ctypes.c_longlong, vp, c_double_p, c_double_p, vp, vp, vp]
    L.nlpf_ecef2enu.argtypes = [ctypes.c_longlong, vp, c_double_p, c_double_p, ctypes.c_int, vp, vp]
    L.nlpf_fp64_probe.argtypes = [ctypes.c_int, vp, ctypes.POINTER(ctypes.c_double), vp]
    L.nlpf_dmma_probe.argtypes = [ctypes.c_int, vp, ctypes.POINTER(ctypes.c_double), vp]
    for name in EXPORTS:
        getattr(L, name)
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise NlpfError("libnlpfilter error %d: %s" % (rc, lib().nlpf_last_error().decode("utf-8", "replace")))


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def _ip(a):
    return a.ctypes.data_as(c_int_p)


def make_cdesc(d):
    """ProblemDesc (nlp_filter_b200.desc) -> (CDesc, keepalive list)."""
    d.finalize()
    c = CDesc()
    c.K, c.n, c.m, c.nv, c.ng = d.K, d.n, d.m, d.nv, d.ng
    c.dyn_id, c.loss, c.has_prior = d.dyn_id, d.loss, int(d.has_prior)
    c.nb, c.E, c.M, c.Ldata = d.nb, d.E, d.M, d.Ldata
    c.T, c.delta = d.T, d.delta
    keep = []

    def f8(a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        if a.size == 0:
            a = np.zeros(1)
        keep.append(a)
        return _dp(a)

    def i4(a):
        a = np.ascontiguousarray(a, dtype=np.int32)
        if a.size == 0:
            a = np.zeros(1, dtype=np.int32)
        keep.append(a)
        return _ip(a)

    c.dyn_par, c.A, c.Bm, c.c0 = f8(d.dyn_par), f8(d.A), f8(d.Bm[:, :max(d.m, 1)]), f8(d.c0)
    c.cmap, c.ucol = i4(d.cmap), i4(d.ucol)
    c.Dt, c.cw, c.cq = f8(d.Dt), f8(d.cw), f8(d.cq)
    c.qinv, c.pinv, c.snode = f8(d.qinv), f8(d.pinv), f8(d.snode)
    c.bidx, c.blo, c.bhi = i4(d.bidx), f8(d.blo), f8(d.bhi)
    c.Phi = f8(d.Phi)
    c.rtype, c.repoch, c.ridx, c.roff = i4(d.rtype), i4(d.repoch), i4(d.ridx), i4(d.roff)
    return c, keep


def make_copts(o):
    c = COpts()
    c.max_iter, c.max_stage_iter = o.max_iter, o.max_stage_iter
    c.tol_step, c.lam0, c.mu0, c.mu_min = o.tol_step, o.lam0, o.mu0, o.mu_min
    c.bound_push, c.tau_ftb = o.bound_push, o.tau_ftb
    c.lam_jump, c.lam_jump_indef, c.lam_dec_first = o.lam_jump, o.lam_jump_indef, o.lam_dec_first
    c.mu0_warm = o.mu0_warm
    return c
