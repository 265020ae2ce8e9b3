"""Long horizons by multi-segment collocation (BASELINE.json's "horizon 200").

The reference is single-interval (nlp/nlp.py:189-200) and its monomial Lagrange basis breaks down for
N >~ 25 (nlp/collocation.py:87-111), so a 200-epoch horizon cannot be one reference window.  Here the
horizon [0, S*T_s] is S segments, each the reference's Chebyshev collocation on its own N_s + 1 nodes, with
neighbouring segments sharing their end node.  Every cost term of nlp/nlp.py:202-277 lives inside one segment:

    J(z) = sum_s J_seg(z_seg(s)) + arrival cost on node 0,     z_seg(s) = nodes s*N_s .. s*N_s + N_s

so kernels (1)+(2) run unchanged on B*S gathered sub-windows of the ordinary single-segment problem class
(`nlpf_assemble`), `nlpf_ms_blocks` overlap-adds the segment matrices into the block-tridiagonal normal
equations (block size N_s*n), and kernel (3) is `nlpf_blocktri_solve` -- block Cholesky (Riccati recursion)
or block cyclic reduction.  The damped Gauss-Newton iteration around them is the one of DESIGN.md section 3
(same acceptance rule, full evaluation at the trial point), batched over the windows: every window carries its
own damping, accepted / rejected / finished windows are masks on the device; one scalar is read back per
iteration to know when the batch is done.  With S = 1 this is the reference's single-interval problem and the
result equals the dense whole-solve kernel's (tests/test_gpu_multiseg.py).

torch is used for device memory and the element-wise bookkeeping of the iteration (plumbing); residuals,
Jacobians, J^T W J, the block assembly and the factorisation are the library's CUDA kernels.
"""
import ctypes

import numpy as np
import torch

from . import _capi, engine
from .desc import SolveOpts

RICCATI = engine.BLOCKTRI_RICCATI
CYCLIC = engine.BLOCKTRI_CYCLIC


class MultiSegmentProblem(object):
    """S segments of the single-segment problem class `seg_desc` (K = N_s + 1 nodes, horizon T_s, its own epochs
    and residual table -- the same for every segment; no arrival cost, bounds or global unknowns in `seg_desc`:
    the arrival cost of the long window is given here as `pinv`)."""

    def __init__(self, seg_desc, S, pinv=None):
        d = seg_desc.finalize()
        if d.has_prior or d.nb or d.ng or d.nv != d.n:
            raise NotImplementedError("segment class must be plain (arrival cost via pinv=, no bounds / globals / folded equalities)")
        self.d, self.S = d, int(S)
        self.n, self.Ns = d.n, d.K - 1
        self.bs = self.Ns * d.n
        self.kn = d.K * d.n
        self.nzt = self.S * self.bs + d.n
        self.nblk = self.S + 1
        self.prob = engine.WindowProblem(d)
        self._lib = self.prob._lib
        self.pinv = None if pinv is None else torch.as_tensor(np.asarray(pinv, dtype=np.float64).reshape(d.n)).cuda()

    def close(self):
        self.prob.close()

    # ---- kernels ------------------------------------------------------------------------------
    def _assemble(self, z, U, data, xprior, lam, out):
        """H, g, J of B long windows at z -> block-tridiagonal (Dg + lam diag, Lo), g, J in `out`."""
        B, S, d = z.shape[0], self.S, self.d
        st = engine._stream()
        p = engine._ptr
        _capi.check(self._lib.nlpf_ms_gather(B, S, self.kn, self.bs, p(z), p(out["zsub"]), st))
        _capi.check(self._lib.nlpf_assemble(self.prob._h, B * S, p(out["zsub"]), p(U), p(data), None, None, 0.0,
                                            p(out["Hs"]), p(out["gs"]), p(out["cs"]), st))
        self._blocks(z, xprior, lam, out)

    def _blocks(self, z, xprior, lam, out):
        B = z.shape[0]
        p = engine._ptr
        _capi.check(self._lib.nlpf_ms_blocks(B, self.S, self.n, self.bs, p(out["Hs"]), p(out["gs"]), p(out["cs"]),
                                             p(self.pinv), p(xprior) if self.pinv is not None else None, p(z), p(lam),
                                             p(out["Dg"]), p(out["Lo"]), p(out["g"]), p(out["J"]), engine._stream()))

    def _buffers(self, B):
        f = lambda *shape: torch.empty(shape, dtype=torch.float64, device="cuda")
        S, kn, bs = self.S, self.kn, self.bs
        return dict(zsub=f(B * S, kn), Hs=f(B * S, kn, kn), gs=f(B * S, kn), cs=f(B * S, 2), Dg=f(B, self.nblk, bs, bs),
                    Lo=f(B, self.nblk, bs, bs), g=f(B, self.nblk * bs), J=f(B))

    def normal_equations(self, z, U, data, xprior=None, lam=None):
        """(Dg, Lo, g, J) at z -- for tests and ncu."""
        z = engine._f8(z, (-1, self.nzt))
        B = z.shape[0]
        out = self._buffers(B)
        lam_t = None if lam is None else engine._f8(np.full(B, float(lam)))
        self._assemble(z, self._sub(U, B, self.d.K * max(self.d.m, 1)), self._sub(data, B, max(self.d.Ldata, 1)),
                       None if xprior is None else engine._f8(xprior, (B, self.n)), lam_t, out)
        return out["Dg"], out["Lo"], out["g"], out["J"]

    def _sub(self, a, B, width):
        if a is None:
            return None
        return engine._f8(a).reshape(B * self.S, width)

    # ---- whole solve ----------------------------------------------------------------------------
    def _state(self, B, has_xp, variant, o):
        """Static buffers (and, after the first sweep, the CUDA graph of one sweep) of a (batch size, options) combination:
        a sweep is 4 kernel launches and ~40 element-wise torch ops; issued one by one from Python it costs 4 ms of which
        1.3 ms are kernels, replayed as a graph it costs the kernels."""
        key = (B, has_xp, int(variant), tuple(sorted(o.__dict__.items())))
        st = getattr(self, "_states", {}).get(key)
        if st is None:
            f = lambda *shape: torch.empty(shape, dtype=torch.float64, device="cuda")
            d, S = self.d, self.S
            st = dict(z=f(B, self.nzt), zt=f(B, self.nzt), U=f(B * S, d.K * max(d.m, 1)), data=f(B * S, max(d.Ldata, 1)),
                      xp=f(B, self.n) if has_xp else None, cur=self._buffers(B), tri=self._buffers(B), lam=f(B),
                      iters=torch.zeros((B,), dtype=torch.int32, device="cuda"), status=torch.zeros((B,), dtype=torch.int32, device="cuda"),
                      active=torch.ones((B,), dtype=torch.bool, device="cuda"), info=torch.empty((B,), dtype=torch.int32, device="cuda"),
                      dz=f(B, self.nblk * self.bs), graph=None)
            if not hasattr(self, "_states"):
                self._states = {}
            self._states = {key: st}                    # one combination is kept (its buffers are tens of MB)
        return st

    def _sweep(self, st, o, variant):
        """One damped Gauss-Newton sweep over the whole batch on the static buffers (no host synchronisation, no
        data-dependent shapes: capturable)."""
        S = self.S
        p = engine._ptr
        z, zt, lam, cur, tri, active, iters, status, dz, info = (st[k] for k in ("z", "zt", "lam", "cur", "tri", "active", "iters", "status", "dz", "info"))
        _capi.check(self._lib.nlpf_blocktri_solve(z.shape[0], self.nblk, self.bs, p(cur["Dg"]), p(cur["Lo"]), p(cur["g"]), p(dz),
                                                  p(info), int(variant), engine._stream()))
        step = dz[:, :self.nzt]
        ok = (info == 0) & torch.isfinite(step).all(1) & active
        torch.sub(z, step, out=zt)
        self._assemble(zt, st["U"], st["data"], st["xp"], None, tri)
        acc = ok & torch.isfinite(tri["J"]) & (tri["J"] < cur["J"])
        rej = active & ~acc
        # accepted windows move: iterate, segment matrices and objective of the trial point become current
        m1, m2, m3 = acc[:, None], acc.repeat_interleave(S)[:, None], acc.repeat_interleave(S)[:, None, None]
        z.copy_(torch.where(m1, zt, z))
        cur["Hs"].copy_(torch.where(m3, tri["Hs"], cur["Hs"]))
        cur["gs"].copy_(torch.where(m2, tri["gs"], cur["gs"]))
        cur["cs"].copy_(torch.where(m2, tri["cs"], cur["cs"]))
        smax = step.abs().amax(1)
        zmax = z.abs().amax(1)
        small = smax < o.tol_step * (1.0 + zmax)
        # converged: an accepted step below the tolerance -- or a proposed step below it that the objective
        # cannot resolve any more (J equal to rounding: the rounding-floor case the fused kernels treat with
        # their first-order noise bound; nlpf_assemble does not report that bound, 1e-10 |J| stands in for it)
        floor = ok & ~acc & small & ((tri["J"] - cur["J"]) <= 1e-10 * cur["J"].abs())
        conv = (acc & small) | floor
        lam.copy_(torch.where(acc, torch.clamp(lam / 10.0, min=1e-12), torch.where(rej, torch.clamp(lam * 10.0, min=1e-4), lam)))
        iters += acc.to(torch.int32)
        stalled = rej & ~floor & (lam > 1e18)
        hit = active & ~conv & ~stalled & (iters >= o.max_iter)
        status.copy_(torch.where(conv, torch.zeros_like(status), torch.where(stalled, torch.full_like(status, 4), status)))   # 4: NLPF_ST_NOSTEP
        active &= ~(conv | stalled | hit)
        self._blocks(z, st["xp"], lam, cur)                  # normal equations of the current point with the new damping

    def solve(self, z, U, data, xprior=None, opts=None, variant=RICCATI, graph=True):
        """z (B, nzt) CUDA tensor in/out; U (B, S, K, m), data (B, S, Ldata_seg), xprior (B, n).
        Returns (z, J, iters, status) device tensors (status as in include/nlpfilter.h).  graph=True: from the second
        sweep on, a sweep is replayed as one CUDA graph (captured once per batch size / options)."""
        o = opts or SolveOpts()
        d = self.d
        z = engine._f8(z, (-1, self.nzt))
        B = z.shape[0]
        st = self._state(B, xprior is not None, variant, o)
        st["z"].copy_(z)
        st["U"].copy_(self._sub(U, B, d.K * max(d.m, 1)))
        st["data"].copy_(self._sub(data, B, max(d.Ldata, 1)))
        if xprior is not None:
            st["xp"].copy_(engine._f8(xprior, (B, self.n)))
        st["lam"].fill_(o.lam0 if o.lam0 > 0.0 else 1e-3)      # exact-Hessian class (nlpf_opts.lam0)
        st["iters"].zero_()
        st["status"].fill_(2)
        st["active"].fill_(True)
        self._assemble(st["z"], st["U"], st["data"], st["xp"], st["lam"], st["cur"])
        bad0 = ~torch.isfinite(st["cur"]["J"])
        st["status"].copy_(torch.where(bad0, torch.full_like(st["status"], 3), st["status"]))
        st["active"] &= ~bad0
        for sweep in range(4 * o.max_iter):
            if not bool(st["active"].any().item()):            # the one scalar read back per sweep
                break
            if graph and st["graph"] is None and sweep == 1:   # (the first sweep ran eagerly: kernel attributes, workspaces)
                torch.cuda.synchronize()
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                g = torch.cuda.CUDAGraph()
                with torch.cuda.stream(side):
                    with torch.cuda.graph(g, stream=side):
                        self._sweep(st, o, variant)
                torch.cuda.current_stream().wait_stream(side)
                st["graph"] = g                                 # (capturing does not execute: the replay below is sweep 1)
            if graph and st["graph"] is not None:
                st["graph"].replay()
            else:
                self._sweep(st, o, variant)
        z.copy_(st["z"])
        return z, st["cur"]["J"].clone(), st["iters"].clone(), st["status"].clone()

    def sample(self, z, t):
        """x(t) for t in [0, S*T_s]: the segment's Lagrange basis (nlp/collocation.py:113-121) -> (B, len(t), n)."""
        z = engine._f8(z, (-1, self.nzt)).cpu().numpy()
        t = np.asarray(t, dtype=float).reshape(-1)
        Ts = self.d.T
        seg = np.minimum((t / Ts).astype(int), self.S - 1)
        V = z.reshape(z.shape[0], -1, self.n)
        out = np.zeros((z.shape[0], t.shape[0], self.n))
        for i, (s, tt) in enumerate(zip(seg, t)):
            phi = self.d.cpm.basisMatrix([tt - s * Ts])[0]
            out[:, i, :] = np.einsum("k,bkn->bn", phi, V[:, s * self.Ns:s * self.Ns + self.d.K, :])
        return out
