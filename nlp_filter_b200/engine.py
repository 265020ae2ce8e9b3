"""Batched device entry points above the C ABI.

`WindowProblem` wraps one `nlpf_problem` handle (a problem class); its methods take fp64 CUDA
tensors (torch is used for device memory and streams only) and hand raw pointers to
libnlpfilter.so.  `solve_windows_host` is the public end-to-end call for a batch of windows
held in host (NumPy) buffers: pinned staging -> H2D -> whole-solve kernel -> D2H.

Additions to the reference API, never changes: the reference solves one window per
`NLP.solve()` call (nlp/nlp.py:76-83); these calls solve B independent windows of the same
problem class at once.
"""
import ctypes

import os

import numpy as np
import torch

from . import _capi
from .desc import SolveOpts


def _require_cuda():
    if not torch.cuda.is_available():
        raise _capi.NlpfError("no CUDA device: the estimator hot path has no CPU fallback")


def _ptr(t):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "device buffers must be contiguous CUDA tensors"
    return ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _want(name, t, shape):
    """Raise on a wrong dtype / device / element count before a raw pointer reaches the library (a wrong shape
    would otherwise be a silent out-of-bounds access on the device)."""
    if t is None:
        raise _capi.NlpfError("%s is required for this problem class (expected shape %s)" % (name, shape))
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()):
        raise _capi.NlpfError("%s must be a contiguous float64 CUDA tensor" % name)
    want = int(np.prod(shape))
    if t.numel() != want:
        raise _capi.NlpfError("%s has %d elements, expected %s = %d" % (name, t.numel(), "x".join(map(str, shape)), want))
    return t


def _f8(t, shape=None):
    if t is None:
        return None
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        t = torch.as_tensor(np.ascontiguousarray(t, dtype=np.float64)).cuda()
    t = t.to(torch.float64).contiguous()
    if shape is not None:
        t = t.reshape(shape)
    return t


class WindowProblem(object):
    """One problem class on the current CUDA device."""

    def __init__(self, desc):
        _require_cuda()
        self.desc = desc.finalize()
        self._lib = _capi.lib()
        cdesc, keep = _capi.make_cdesc(desc)
        h = ctypes.c_void_p()
        _capi.check(self._lib.nlpf_problem_create(ctypes.byref(cdesc), ctypes.byref(h)))
        self._h = h
        if desc.dyn_id == 6:                 # DYN_USER: compile the window kernels with the generated model (NVRTC)
            if not getattr(desc, "user_src", None):
                raise _capi.NlpfError("problem class with a user-defined model carries no generated source (desc.user_src)")
            here = os.path.dirname(os.path.abspath(__file__))
            _capi.check(self._lib.nlpf_problem_set_user_dynamics(h, desc.user_src.encode("utf-8"),
                                                                 os.path.join(here, "csrc").encode("utf-8"),
                                                                 os.path.join(os.path.dirname(here), "include").encode("utf-8")))
        self.nz = desc.nz
        self.smem_bytes = int(self._lib.nlpf_problem_smem_bytes(h))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.nlpf_problem_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- kernels (1)+(2) ---------------------------------------------------------------------
    def assemble(self, z, U=None, data=None, xprior=None, wshift=None, mu=0.0):
        d = self.desc
        z = _f8(z, (-1, d.nz))
        B = z.shape[0]
        U, data, xprior, wshift = _f8(U), _f8(data), _f8(xprior), _f8(wshift)
        self._check_inputs(B, U, data, xprior, wshift)
        H = torch.empty((B, d.nz, d.nz), dtype=torch.float64, device=z.device)
        g = torch.empty((B, d.nz), dtype=torch.float64, device=z.device)
        cost = torch.empty((B, 2), dtype=torch.float64, device=z.device)
        _capi.check(self._lib.nlpf_assemble(self._h, B, _ptr(z), _ptr(U), _ptr(data), _ptr(xprior), _ptr(wshift),
                                            float(mu), _ptr(H), _ptr(g), _ptr(cost), _stream()))
        return H, g, cost

    # -- kernel (3) --------------------------------------------------------------------------
    def factor_solve(self, H, g, lam=0.0):
        d = self.desc
        H = _f8(H, (-1, d.nz, d.nz))
        g = _f8(g, (-1, d.nz))
        B = H.shape[0]
        dz = torch.empty((B, d.nz), dtype=torch.float64, device=H.device)
        info = torch.empty((B,), dtype=torch.int32, device=H.device)
        _capi.check(self._lib.nlpf_factor_solve(self._h, B, _ptr(H), _ptr(g), float(lam), _ptr(dz), _ptr(info), _stream()))
        return dz, info

    # -- whole solve -------------------------------------------------------------------------
    def solve(self, z, U=None, data=None, xprior=None, wshift=None, opts=None, out=None):
        """z (B, nz) CUDA tensor: initial guess in, minimiser out (in place).
        Returns (z, cost, iters, status) device tensors; nothing is synchronised."""
        d = self.desc
        if z.dim() != 2 or z.shape[1] != d.nz:
            raise _capi.NlpfError("z must have shape (B, %d), got %s" % (d.nz, tuple(z.shape)))
        B = z.shape[0]
        _want("z", z, (B, d.nz))
        self._check_inputs(B, U, data, xprior, wshift)
        copts = _capi.make_copts(opts or SolveOpts())
        if out is None:
            cost = torch.empty((B,), dtype=torch.float64, device=z.device)
            iters = torch.empty((B,), dtype=torch.int32, device=z.device)
            status = torch.empty((B,), dtype=torch.int32, device=z.device)
        else:
            cost, iters, status = out
        _capi.check(self._lib.nlpf_solve(self._h, B, _ptr(z), _ptr(U), _ptr(data), _ptr(xprior), _ptr(wshift),
                                         ctypes.byref(copts), _ptr(cost), _ptr(iters), _ptr(status), _stream()))
        return z, cost, iters, status

    def _check_inputs(self, B, U, data, xprior, wshift, W=1):
        d = self.desc
        if d.m > 0 and any(c < 0 for c in np.asarray(d.ucol).reshape(-1)[:d.m]):   # parameter controls
            _want("U", U, (W, B, d.K, d.m) if W > 1 else (B, d.K, d.m))
        if d.Ldata > 0:
            _want("data", data, (W, B, d.Ldata) if W > 1 else (B, d.Ldata))
        if d.has_prior:
            _want("xprior", xprior, (B, d.n))
        if wshift is not None:
            _want("wshift", wshift, (B, d.K, d.n))

    def defects(self, z, U=None):
        """Collocation defects W (B, K, n) at z (nlp.py:222-235)."""
        d = self.desc
        z = _f8(z, (-1, d.nz))
        B = z.shape[0]
        W = torch.empty((B, d.K, d.n), dtype=torch.float64, device=z.device)
        _capi.check(self._lib.nlpf_defects(self._h, B, _ptr(z), _ptr(_f8(U)), _ptr(W), _stream()))
        return W

    def profile(self, enable=True):
        """Per-phase clock counters of the warp-per-window kernel (nlpf_problem_profile): returns what has been
        accumulated so far as a dict, then (re)starts or stops counting."""
        out = (ctypes.c_ulonglong * 5)()
        _capi.check(self._lib.nlpf_problem_profile(self._h, 1 if enable else 0, out))
        return dict(eval_cycles=int(out[0]), ldl_cycles=int(out[1]), evals=int(out[2]), factorisations=int(out[3]),
                    window_cycles=int(out[4]))

    def sample(self, z, PhiT):
        d = self.desc
        z = _f8(z, (-1, d.nz))
        PhiT = _f8(PhiT, (-1, d.K))
        B, nt = z.shape[0], PhiT.shape[0]
        out = torch.empty((B, nt, d.n), dtype=torch.float64, device=z.device)
        _capi.check(self._lib.nlpf_sample(self._h, B, _ptr(z), nt, _ptr(PhiT), _ptr(out), _stream()))
        return out


class HostBatch(object):
    """Pinned host staging + device buffers for repeated end-to-end solves of B windows.

    The batch is cut into `nchunk` contiguous chunks, each with its own stream and problem handle:
    chunk c+1's host-to-device copy overlaps chunk c's solve and chunk c-1's device-to-host copy
    (a GNSS window carries 25 KB of measurement records: 103 MB per 4096 windows, as long on PCIe
    as the solve itself).  Every chunk's inputs [z | U | data | xprior] sit back to back in ONE pinned
    buffer and one device buffer, its outputs [cost | iters | status] in another: one H2D and two D2H
    copies per chunk instead of eight small ones (the launch overhead of those was a tenth of a VDP step)."""

    def __init__(self, prob, B, nchunk=None):
        d = prob.desc
        self.prob, self.B = prob, B
        if nchunk is None:
            # measured on B200 (4096 windows, e2e windows/s by chunk count): van der Pol (2 KB per window,
            # solve-bound) 1: 3.07 M, 2: 3.61 M, 3: 3.74 M, 4: 3.57 M;  GNSS (25 KB of records per window,
            # PCIe-bound: 103 MB per step) 2: 1.41 M, 4: 1.31-1.54 M, 8: 1.55 M.  NLPF_E2E_CHUNKS overrides.
            env = os.environ.get("NLPF_E2E_CHUNKS")
            in_bytes = 8 * B * (d.nz + d.K * d.m + max(d.Ldata, 1) + d.n)
            nchunk = int(env) if env else (1 if B < 1024 else (8 if in_bytes >= (32 << 20) else 3))
        self.nchunk = max(1, min(int(nchunk), B))
        cuts = [(B * c) // self.nchunk for c in range(self.nchunk + 1)]
        self.chunks = [(cuts[c], cuts[c + 1]) for c in range(self.nchunk)]
        # every chunk has a handle of its own (scratch buffers of a handle are re-allocated when a later call on it needs
        # more -- a captured graph must not share them with the caller's handle)
        self.probs = [WindowProblem(d) for _ in range(self.nchunk)]
        self.streams = [torch.cuda.Stream() for _ in range(self.nchunk)] if self.nchunk > 1 else [None]
        mU = d.K * d.m if d.m > 0 else 0
        Ld = max(d.Ldata, 1)
        nx = d.n if d.has_prior else 0
        self._nz, self._mU, self._Ld, self._nx = d.nz, mU, Ld, nx
        self.h_in, self.d_in, self.h_meta, self.d_meta, self.h_zout = [], [], [], [], []
        for lo, hi in self.chunks:
            nb = hi - lo
            n_in = nb * (d.nz + mU + Ld + nx)
            self.h_in.append(torch.zeros((n_in,), dtype=torch.float64).pin_memory())
            self.d_in.append(torch.empty((n_in,), dtype=torch.float64, device="cuda"))
            self.h_meta.append(torch.zeros((2 * nb,), dtype=torch.float64).pin_memory())
            self.d_meta.append(torch.empty((2 * nb,), dtype=torch.float64, device="cuda"))
            self.h_zout.append(torch.zeros((nb, d.nz), dtype=torch.float64).pin_memory())
        self.h2d_bytes = sum(t.numel() * 8 for t in self.h_in)
        self.d2h_bytes = sum(t.numel() * 8 for t in self.h_meta) + sum(t.numel() * 8 for t in self.h_zout)

    def _views(self, buf, c):
        """(z, U, data, xprior) views of chunk c's packed input buffer (host or device)."""
        lo, hi = self.chunks[c]
        nb = hi - lo
        d = self.prob.desc
        o = 0
        z = buf[o:o + nb * self._nz].view(nb, self._nz); o += nb * self._nz
        U = buf[o:o + nb * self._mU].view(nb, d.K, d.m) if self._mU else None; o += nb * self._mU
        data = buf[o:o + nb * self._Ld].view(nb, self._Ld); o += nb * self._Ld
        xp = buf[o:o + nb * self._nx].view(nb, self._nx) if self._nx else None
        return z, U, data, xp

    def _meta(self, buf, c):
        lo, hi = self.chunks[c]
        nb = hi - lo
        ints = buf[nb:].view(torch.int32)
        return buf[:nb], ints[:nb], ints[nb:2 * nb]

    def load(self, z0, U=None, data=None, xprior=None):
        d = self.prob.desc
        z0 = torch.as_tensor(np.asarray(z0, dtype=np.float64).reshape(self.B, d.nz))
        U = None if not self._mU else torch.as_tensor(np.asarray(U, dtype=np.float64).reshape(self.B, d.K, d.m))
        data = None if data is None else torch.as_tensor(np.asarray(data, dtype=np.float64).reshape(self.B, self._Ld))
        xprior = None if not self._nx else torch.as_tensor(np.asarray(xprior, dtype=np.float64).reshape(self.B, d.n))
        for c, (lo, hi) in enumerate(self.chunks):
            hz, hU, hd, hx = self._views(self.h_in[c], c)
            hz.copy_(z0[lo:hi])
            if hU is not None:
                hU.copy_(U[lo:hi])
            if data is not None:
                hd.copy_(data[lo:hi])
            if hx is not None:
                hx.copy_(xprior[lo:hi])

    def _run_chunk(self, c, opts):
        self.d_in[c].copy_(self.h_in[c], non_blocking=True)
        z, U, data, xp = self._views(self.d_in[c], c)
        self.probs[c].solve(z, U, data, xp, None, opts, out=self._meta(self.d_meta[c], c))
        self.h_zout[c].copy_(z, non_blocking=True)
        self.h_meta[c].copy_(self.d_meta[c], non_blocking=True)

    def capture(self, opts=None):
        """Record one run() -- every chunk's copies and kernels on their streams -- into a CUDA graph; run() with the
        same options then replays it with a single launch.  Issuing a chunk from Python costs ~0.09 ms of host time
        (copies, memset, three kernels, events through ctypes / torch), which at 3-6 chunks is as long as the van der
        Pol solve itself and keeps the later chunks' copies from starting in time.  Call after one plain run() (first
        calls allocate scratch and set kernel attributes, which cannot be captured)."""
        self._graph, self._graph_key = None, None
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            self._issue(opts)                                   # (warm-up on the side stream, as torch's capture rules ask)
            side.synchronize()
            with torch.cuda.graph(g, stream=side):
                self._issue(opts)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._graph, self._graph_key = g, self._opts_key(opts)

    @staticmethod
    def _opts_key(opts):
        return None if opts is None else tuple(sorted(opts.__dict__.items()))

    def run(self, opts=None):
        """H2D of the staged inputs, solve, D2H of the estimates, chunk-pipelined.  Asynchronous with
        respect to the host: the current stream waits for every chunk; call
        torch.cuda.synchronize() (or wait on an event of the current stream) before reading results()."""
        if getattr(self, "_graph", None) is not None and self._graph_key == self._opts_key(opts):
            self._graph.replay()
            return
        self._issue(opts)

    def _issue(self, opts):
        if self.nchunk == 1:
            self._run_chunk(0, opts)
            return
        cur = torch.cuda.current_stream()
        start = torch.cuda.Event()
        start.record(cur)
        for c in range(self.nchunk):
            st = self.streams[c]
            st.wait_event(start)
            with torch.cuda.stream(st):
                self._run_chunk(c, opts)
                done = torch.cuda.Event()
                done.record(st)
            cur.wait_event(done)

    def results(self):
        """(z, cost, iters, status) NumPy arrays of the last run (synchronise first)."""
        z = np.concatenate([t.numpy() for t in self.h_zout], 0)
        metas = [self._meta(self.h_meta[c], c) for c in range(self.nchunk)]
        cost = np.concatenate([m[0].numpy() for m in metas])
        iters = np.concatenate([m[1].numpy() for m in metas])
        status = np.concatenate([m[2].numpy() for m in metas])
        return z, cost, iters, status

    def close(self):
        """Destroys the chunks' problem handles (and with them a captured graph's scratch: do not run() afterwards)."""
        self._graph = None
        for p in self.probs:
            p.close()
        self.probs = []

    def device_z(self):
        """(B, nz) contiguous device tensor of the estimates of the last run (for a collective)."""
        zs = [self._views(self.d_in[c], c)[0] for c in range(self.nchunk)]
        if self.nchunk == 1:
            return zs[0]
        if getattr(self, "_zcat", None) is None:           # allocated once: no allocation inside a timed step
            self._zcat = torch.empty((self.B, self._nz), dtype=torch.float64, device="cuda")
        for (lo, hi), zc in zip(self.chunks, zs):
            self._zcat[lo:hi].copy_(zc, non_blocking=True)
        return self._zcat


def solve_chain(prob, z, U_all, data_all, xprior, PhiDT, opts=None, keep=False):
    """Receding-horizon chains of B independent runs, all on the device (autonomous-car.py:232-288):
    for every window w: solve from the previous window's solution (warm start), then the next arrival
    cost is centred on this solution sampled at t = DT.  z (B, nz) and xprior (B, n) are updated in
    place; U_all (W, B, K, m), data_all (W, B, Ldata) device tensors.  Nothing is synchronised.
    Returns (iters_sum (B,), bad (B,) windows that did not converge[, Z (W, B, nz)])."""
    d = prob.desc
    W, B = data_all.shape[0], z.shape[0]
    _want("z", z, (B, d.nz))
    if not (xprior.is_cuda and xprior.dtype == torch.float64 and xprior.numel() == B * d.n):
        raise _capi.NlpfError("xprior must be a float64 CUDA tensor of shape (%d, %d)" % (B, d.n))
    if data_all.dim() != 3 or data_all.shape[1] != B or data_all.shape[2] != max(d.Ldata, 1):
        raise _capi.NlpfError("data_all must have shape (W, %d, %d), got %s" % (B, d.Ldata, tuple(data_all.shape)))
    if d.m > 0 and U_all is not None and tuple(U_all.shape) != (W, B, d.K, d.m):
        raise _capi.NlpfError("U_all must have shape (%d, %d, %d, %d), got %s" % (W, B, d.K, d.m, tuple(U_all.shape)))
    cost = torch.empty((B,), dtype=torch.float64, device=z.device)
    iters = torch.empty((B,), dtype=torch.int32, device=z.device)
    status = torch.empty((B,), dtype=torch.int32, device=z.device)
    it_sum = torch.zeros((B,), dtype=torch.int64, device=z.device)
    bad = torch.zeros((B,), dtype=torch.int64, device=z.device)
    Z = torch.empty((W, B, d.nz), dtype=torch.float64, device=z.device) if keep else None
    PhiDT = _f8(PhiDT, (-1, d.K))
    fused = (os.environ.get("NLPF_CHAIN_LOOP") is None and prob._lib.nlpf_problem_has_chain_kernel(prob._h) == 1
             and (d.m == 0 or U_all is not None) and z.is_contiguous() and data_all.is_contiguous()
             and xprior.is_contiguous() and (U_all is None or U_all.is_contiguous()))
    if fused:
        # one launch: every run's W windows back to back in one warp (no per-window launch boundary where the
        # whole batch waits for its slowest window)
        copts = _capi.make_copts(opts or SolveOpts())
        it32 = torch.empty((B,), dtype=torch.int32, device=z.device)
        bad32 = torch.empty((B,), dtype=torch.int32, device=z.device)
        _capi.check(prob._lib.nlpf_solve_chain(prob._h, B, W, _ptr(z), _ptr(U_all), _ptr(data_all), _ptr(xprior),
                                               _ptr(PhiDT), ctypes.byref(copts), _ptr(cost), _ptr(it32), _ptr(bad32),
                                               _ptr(Z), _stream()))
        it_sum, bad = it32.to(torch.int64), bad32.to(torch.int64)
        return (it_sum, bad, Z) if keep else (it_sum, bad)
    for w in range(W):
        prob.solve(z, None if U_all is None else U_all[w], data_all[w], xprior, None, (opts or SolveOpts()).for_window(w), out=(cost, iters, status))
        xprior.copy_(prob.sample(z, PhiDT)[:, 0, :])
        it_sum += iters
        bad += (status > 1)
        if keep:
            Z[w].copy_(z)
    return (it_sum, bad, Z) if keep else (it_sum, bad)


class HostChains(object):
    """Pinned host staging + device buffers for repeated end-to-end runs of B receding-horizon chains of W windows
    (solve_chain from host buffers).  The runs are cut into `nchunk` groups, each with its own stream, problem handle
    and run-major staging buffers (W, Bc, ...): group c+1's host-to-device copy overlaps group c's fused-chain launch
    (a vehicle chain of 200 windows carries 1.8 MB of records; 2048 of them are 3.8 GB, half as long on PCIe as the
    solve itself)."""

    def __init__(self, prob, B, W, nchunk=4):
        d = prob.desc
        self.prob, self.B, self.W = prob, B, W
        self.nchunk = max(1, min(int(nchunk), B))
        cuts = [(B * c) // self.nchunk for c in range(self.nchunk + 1)]
        self.chunks = [(cuts[c], cuts[c + 1]) for c in range(self.nchunk)]
        self.probs = [prob] + [WindowProblem(d) for _ in range(self.nchunk - 1)]
        self.streams = [torch.cuda.Stream() for _ in range(self.nchunk)]
        Ld = max(d.Ldata, 1)
        self.hU, self.hD, self.hX, self.dU, self.dD, self.dX, self.dZ, self.hZ, self.out = [], [], [], [], [], [], [], [], []
        for lo, hi in self.chunks:
            nb = hi - lo
            self.hU.append(torch.zeros((W, nb, d.K, max(d.m, 1)), dtype=torch.float64).pin_memory())
            self.hD.append(torch.zeros((W, nb, Ld), dtype=torch.float64).pin_memory())
            self.hX.append(torch.zeros((nb, d.n), dtype=torch.float64).pin_memory())
            self.dU.append(torch.empty((W, nb, d.K, max(d.m, 1)), dtype=torch.float64, device="cuda"))
            self.dD.append(torch.empty((W, nb, Ld), dtype=torch.float64, device="cuda"))
            self.dX.append(torch.empty((nb, d.n), dtype=torch.float64, device="cuda"))
            self.dZ.append(torch.empty((nb, d.nz), dtype=torch.float64, device="cuda"))
            self.hZ.append(torch.zeros((nb, d.nz), dtype=torch.float64).pin_memory())
            self.out.append(None)
        self.h2d_bytes = sum(t.numel() * 8 for t in self.hU + self.hD + self.hX)
        self.d2h_bytes = sum(t.numel() * 8 for t in self.hZ)

    def load(self, U_all, data_all, xhat0):
        """U_all (W, B, K, m), data_all (W, B, Ldata), xhat0 (B, n) NumPy arrays -> the groups' pinned buffers."""
        U_all, data_all, xhat0 = (torch.as_tensor(np.asarray(a, dtype=np.float64)) for a in (U_all, data_all, xhat0))
        for c, (lo, hi) in enumerate(self.chunks):
            self.hU[c].copy_(U_all[:, lo:hi].reshape(self.hU[c].shape))
            self.hD[c].copy_(data_all[:, lo:hi].reshape(self.hD[c].shape))
            self.hX[c].copy_(xhat0[lo:hi])

    def run(self, PhiDT, opts=None):
        """Every group: H2D of its inputs, z <- 0, one fused-chain launch, D2H of the final estimates.  Asynchronous;
        the current stream waits for all groups.  Results: results() after a synchronize."""
        cur = torch.cuda.current_stream()
        start = torch.cuda.Event()
        start.record(cur)
        for c in range(self.nchunk):
            st = self.streams[c]
            st.wait_event(start)
            with torch.cuda.stream(st):
                self.dU[c].copy_(self.hU[c], non_blocking=True)
                self.dD[c].copy_(self.hD[c], non_blocking=True)
                self.dX[c].copy_(self.hX[c], non_blocking=True)
                self.dZ[c].zero_()
                self.out[c] = solve_chain(self.probs[c], self.dZ[c], self.dU[c], self.dD[c], self.dX[c], PhiDT, opts)
                self.hZ[c].copy_(self.dZ[c], non_blocking=True)
                done = torch.cuda.Event()
                done.record(st)
            cur.wait_event(done)

    def results(self):
        """(z (B, nz) NumPy, iterations per run, windows not converged per run) of the last run (synchronise first)."""
        z = np.concatenate([t.numpy() for t in self.hZ], 0)
        it = np.concatenate([o[0].cpu().numpy() for o in self.out])
        bad = np.concatenate([o[1].cpu().numpy() for o in self.out])
        return z, it, bad

    def close(self):
        for p in self.probs[1:]:
            p.close()


def solve_windows_host(prob, z0, U=None, data=None, xprior=None, opts=None):
    """Solve B windows held in NumPy arrays; returns (z, J, iters, status) as NumPy arrays."""
    z0 = np.asarray(z0, dtype=np.float64).reshape(-1, prob.desc.nz)
    hb = HostBatch(prob, z0.shape[0])
    hb.load(z0, U, data, xprior)
    hb.run(opts)
    torch.cuda.synchronize()
    z, cost, iters, status = hb.results()
    out = z.copy(), cost.copy(), iters.copy(), status.copy()
    hb.close()
    return out


# ---------------------------------------------------------------------------------------------
# EKF and least squares
# ---------------------------------------------------------------------------------------------
EKF_GNSS5 = 0
EKF_CAR9 = 1


def ekf_run(model, mu, S, u, z, sat, nsat, Q, r, dt, par=None, want_traj=True, correct_only=False):
    """`steps` fused predict/update steps of B filters (utils/ekf.py:20-61).
    mu (B,n), S (B,n,n), u (B,steps,m), z (B,steps,smax), sat (B,steps,smax,3), nsat (B,steps) int32,
    Q (n,) diagonal or (n,n) full; r scalar variance or (smax,smax) full covariance.  correct_only: EKF.correct
    alone.  mu and S are updated in place (device tensors); returns the mean trajectory."""
    _require_cuda()
    L = _capi.lib()
    mu, S = _f8(mu), _f8(S)
    B, n = mu.shape
    u, z, sat = _f8(u), _f8(z), _f8(sat)
    steps, smax = z.shape[1], z.shape[2]
    if smax > 16:
        raise _capi.NlpfError("ekf_run: at most 16 satellite slots per step, got %d" % smax)
    if isinstance(nsat, np.ndarray) and nsat.size and (nsat.max() > smax or nsat.min() < 0):
        raise _capi.NlpfError("ekf_run: nsat outside [0, %d]" % smax)
    nsat = torch.as_tensor(nsat).to(device=mu.device, dtype=torch.int32).contiguous()
    for name, t, shape in (("S", S, (B, n, n)), ("z", z, (B, steps, smax)), ("sat", sat, (B, steps, smax, 3)),
                           ):
        _want(name, t, shape)
    if nsat.numel() != B * steps:
        raise _capi.NlpfError("ekf_run: nsat has %d entries, expected %d" % (nsat.numel(), B * steps))
    Q = _f8(Q)
    par_t = _f8(np.concatenate([np.asarray(par, float), np.zeros(16 - len(par))])) if par is not None else None
    traj = torch.empty((B, steps, n), dtype=torch.float64, device=mu.device) if want_traj else None
    r_full = np.ndim(r) == 2 if not isinstance(r, torch.Tensor) else r.dim() == 2
    if Q.dim() == 2 or r_full or correct_only:
        Qf = Q if Q.dim() == 2 else torch.diag(Q)
        Rf = _f8(r, (smax, smax)) if r_full else torch.eye(smax, dtype=torch.float64, device=mu.device) * float(r)
        _want("Q", Qf.contiguous(), (n, n))
        _capi.check(L.nlpf_ekf_run_general(int(model), B, steps, smax, _ptr(mu), _ptr(S), _ptr(u), _ptr(z), _ptr(sat), _ptr(nsat),
                                           _ptr(Qf.contiguous()), _ptr(Rf.contiguous()), float(dt), _ptr(par_t), _ptr(traj),
                                           1 if correct_only else 0, _stream()))
        return mu, S, traj
    _want("Q", Q, (n,))
    _capi.check(L.nlpf_ekf_run(int(model), B, steps, smax, _ptr(mu), _ptr(S), _ptr(u), _ptr(z), _ptr(sat), _ptr(nsat),
                               _ptr(Q), float(r), float(dt), _ptr(par_t), _ptr(traj), _stream()))
    return mu, S, traj


def ls_solve(sat, pr, nsat, x0=None, tcol=None, maxiter=100, tol=1e-7):
    """Iterative GNSS least squares of B epochs (utils/leastsquares.py:19-42, 66-94).
    sat (B,smax,3), pr (B,smax), nsat (B,) -> x (B, 4|5) = [x,y,z,b(,alpha)], iters (B,)."""
    _require_cuda()
    L = _capi.lib()
    sat, pr = _f8(sat), _f8(pr)
    B, smax = pr.shape
    if isinstance(nsat, np.ndarray) and nsat.size and (nsat.max() > smax or nsat.min() < 0):
        raise _capi.NlpfError("ls_solve: nsat outside [0, %d]" % smax)
    nsat = torch.as_tensor(nsat).to(device=sat.device, dtype=torch.int32).contiguous()
    _want("sat", sat, (B, smax, 3))
    if nsat.numel() != B:
        raise _capi.NlpfError("ls_solve: nsat has %d entries, expected %d" % (nsat.numel(), B))
    nu = 5 if tcol is not None else 4
    x = torch.zeros((B, nu), dtype=torch.float64, device=sat.device) if x0 is None else _f8(x0, (B, nu)).clone()
    tc = _f8(tcol) if tcol is not None else None
    iters = torch.empty((B,), dtype=torch.int32, device=sat.device)
    _capi.check(L.nlpf_ls_solve(B, smax, _ptr(sat), _ptr(pr), _ptr(nsat), _ptr(tc), _ptr(x), int(maxiter), float(tol),
                                _ptr(iters), _stream()))
    return x, iters


def ls_vel(sat, satvel, prr, nsat, xpos):
    """One-shot velocity / clock-drift solve (utils/leastsquares.py:45-63) -> (B,4) = [v, bdot]."""
    _require_cuda()
    L = _capi.lib()
    sat, satvel, prr, xpos = _f8(sat), _f8(satvel), _f8(prr), _f8(xpos)
    B, smax = prr.shape
    nsat = torch.as_tensor(nsat).to(device=sat.device, dtype=torch.int32).contiguous()
    out = torch.empty((B, 4), dtype=torch.float64, device=sat.device)
    _capi.check(L.nlpf_ls_vel(B, smax, _ptr(sat), _ptr(satvel), _ptr(prr), _ptr(nsat), _ptr(xpos), _ptr(out), _stream()))
    return out


def ecef2enu_batch(p_ecef, p_ref_ECEF, rotation_only=False):
    """ENU coordinates of many ECEF points about one reference point in a single launch
    (utils/utils.py:58-84 applied per point by the reference).  p_ecef (..., 3) array or CUDA tensor."""
    _require_cuda()
    from .utils.utils import _enu_rows
    L = _capi.lib()
    p = _f8(p_ecef)
    shape = p.shape
    p = p.reshape(-1, 3).contiguous()
    R = np.ascontiguousarray(_enu_rows(np.asarray(p_ref_ECEF, dtype=float)).reshape(-1))
    ref = np.ascontiguousarray(np.asarray(p_ref_ECEF, dtype=float).reshape(3))
    out = torch.empty_like(p)
    _capi.check(L.nlpf_ecef2enu(p.shape[0], _ptr(p), R.ctypes.data_as(_capi.c_double_p), ref.ctypes.data_as(_capi.c_double_p),
                                1 if rotation_only else 0, _ptr(out), _stream()))
    return out.reshape(shape)


def enu2lla_batch(p_enu, p_ref_ECEF, want_ecef=False):
    """[lat deg, lon deg, h m] (and optionally ECEF) of many local-ENU points in a single launch: the
    enu2ecef + ecef2lla chain the reference applies per sample (utils/utils.py:87-104, 4-37)."""
    _require_cuda()
    from .utils.utils import _enu_rows
    L = _capi.lib()
    p = _f8(p_enu)
    shape = p.shape
    p = p.reshape(-1, 3).contiguous()
    R = np.ascontiguousarray(_enu_rows(np.asarray(p_ref_ECEF, dtype=float)).reshape(-1))
    ref = np.ascontiguousarray(np.asarray(p_ref_ECEF, dtype=float).reshape(3))
    lla = torch.empty_like(p)
    ecef = torch.empty_like(p) if want_ecef else None
    _capi.check(L.nlpf_enu2lla(p.shape[0], _ptr(p), R.ctypes.data_as(_capi.c_double_p), ref.ctypes.data_as(_capi.c_double_p),
                               _ptr(ecef), _ptr(lla), _stream()))
    return (lla.reshape(shape), ecef.reshape(shape)) if want_ecef else lla.reshape(shape)


BLOCKTRI_RICCATI = 0
BLOCKTRI_CYCLIC = 1


def blocktri_solve(Dg, Lo, rhs, variant=BLOCKTRI_RICCATI):
    """Solve B SPD block-tridiagonal systems.  Dg, Lo (B, nblk, bs, bs),
    rhs (B, nblk, bs) -> x (B, nblk, bs), info (B,)."""
    _require_cuda()
    L = _capi.lib()
    Dg, Lo, rhs = _f8(Dg), _f8(Lo), _f8(rhs)
    B, nblk, bs = rhs.shape
    x = torch.empty_like(rhs)
    info = torch.empty((B,), dtype=torch.int32, device=rhs.device)
    _capi.check(L.nlpf_blocktri_solve(B, nblk, bs, _ptr(Dg), _ptr(Lo), _ptr(rhs), _ptr(x), _ptr(info), int(variant), _stream()))
    return x, info


def fp64_peak_tflops(iters=20000, reps=5, tensor=False):
    """Measured FP64 FMA throughput of the current device (roofline denominator); tensor=True measures the
    FP64 tensor-core path (mma.sync.m8n8k4.f64) instead of the vector FMA pipe."""
    _require_cuda()
    L = _capi.lib()
    probe = L.nlpf_dmma_probe if tensor else L.nlpf_fp64_probe
    if tensor:
        iters = max(1, iters // 8)
    sink = torch.zeros(4, dtype=torch.float64, device="cuda")
    flops = ctypes.c_double()
    best = 0.0
    for i in range(reps + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _capi.check(probe(int(iters), _ptr(sink), ctypes.byref(flops), _stream()))
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            best = max(best, flops.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best
