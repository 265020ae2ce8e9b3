"""ORACLE (test infrastructure, never shipped): dense restatement of the multi-segment long-horizon problem
(nlp_filter_b200/multiseg.py).  The reference has no multi-segment collocation; what is restated is
the DEFINITION the product uses -- J(z) = sum_s J_seg(z_seg(s)) + arrival cost on node 0, every J_seg the reference's
single-interval objective (oracle/nls.py, nlp/nlp.py:202-277) -- assembled the opposite way: one dense
(nzt x nzt) matrix by overlap-add, solved with numpy.linalg, no block structure anywhere.  With S = 1 this IS the
reference's problem (pinned through oracle/nls.py).  Only tests/ and bench.py's CPU leg may import this."""
import numpy as np

from . import nls


def evaluate(seg, S, z, U, data, pinv=None, xprior=None):
    """J, g = grad(J/2), H = Gauss-Newton matrix of J/2 + exact second-order dynamics term (dense)."""
    n, Ns, K = seg.n, seg.K - 1, seg.K
    bs, kn = Ns * n, K * n
    nzt = S * bs + n
    J, g, H = 0.0, np.zeros(nzt), np.zeros((nzt, nzt))
    for s in range(S):
        zs = z[s * bs:s * bs + kn]
        win = nls.Window(seg, U=U[s], data=data[s])
        Js, _, gs, Hs = nls.evaluate(seg, win, zs)
        Hs = Hs + nls.second_order(seg, win, zs, Hs)
        J += Js
        g[s * bs:s * bs + kn] += gs
        H[s * bs:s * bs + kn, s * bs:s * bs + kn] += Hs
    if pinv is not None:
        d0 = z[:n] - xprior
        J += float(np.sum(pinv * d0 * d0))
        g[:n] += pinv * d0
        H[np.arange(n), np.arange(n)] += pinv
    return J, g, H


def solve(seg, S, z0, U, data, pinv=None, xprior=None, tol_step=1e-9, lam0=1e-3, max_iter=300):
    """Damped Gauss-Newton on the dense system (the decisions of oracle/nls.py: accept on decrease, lam / 10, else
    max(10 lam, 1e-4); stop on the step)."""
    z = np.array(z0, dtype=float)
    lam, iters = lam0, 0
    J, g, H = evaluate(seg, S, z, U, data, pinv, xprior)
    while iters < max_iter:
        accepted = False
        while lam <= 1e18:
            try:
                A = H + lam * np.diag(np.abs(np.diag(H)))      # |H_ii|: the second-order term can make a diagonal entry negative
                L = np.linalg.cholesky(A)
                dz = np.linalg.solve(L.T, np.linalg.solve(L, g))
            except np.linalg.LinAlgError:
                lam = max(10.0 * lam, 1e-4)
                continue
            zt = z - dz
            Jt, gt, Ht = evaluate(seg, S, zt, U, data, pinv, xprior)
            if np.isfinite(Jt) and Jt < J:
                accepted = True
                break
            lam = max(10.0 * lam, 1e-4)
        if not accepted:
            return z, dict(J=J, iters=iters, status=4)
        step = float(np.max(np.abs(dz)))
        z, J, g, H = zt, Jt, gt, Ht
        lam = max(lam / 10.0, 1e-12)
        iters += 1
        if step < tol_step * (1.0 + float(np.max(np.abs(z)))):
            return z, dict(J=J, iters=iters, status=0)
    return z, dict(J=J, iters=iters, status=2)


def dense_from_blocks(Dg, Lo, n_real):
    """Dense matrix of a block-tridiagonal system (Dg, Lo: (nblk, bs, bs)), cut to its first n_real rows/columns."""
    nblk, bs = Dg.shape[0], Dg.shape[1]
    A = np.zeros((nblk * bs, nblk * bs))
    for i in range(nblk):
        A[i * bs:(i + 1) * bs, i * bs:(i + 1) * bs] = Dg[i]
        if i > 0:
            A[i * bs:(i + 1) * bs, (i - 1) * bs:i * bs] = Lo[i]
            A[(i - 1) * bs:i * bs, i * bs:(i + 1) * bs] = Lo[i].T
    return A[:n_real, :n_real]
