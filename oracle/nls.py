"""ORACLE (test infrastructure, never shipped): CPU restatement of the reference's
estimation / optimal-control NLP and of the iteration that solves it.

What is restated
  * the objective and constraints that `nlp/nlp.py:202-286,304-317` (estimation) and
    `nlp/nlp.py:135-186` (optimal control) hand to CasADi/IPOPT, with the free process-noise
    variables W_k eliminated through their defining equality (nlp.py:222,235):
        W_k = (2/T) sum_j D_kj X_j - f(X_k, U_k)
    so the problem is a (robust) nonlinear least-squares problem in the node variables;
  * cost terms: dynamics cost (nlp.py:242-245 with cost_functions.py:20-31), measurement
    residual cost (nlp.py:247-277), arrival cost (nlp.py:279-286), stage cost of the OCPs
    (nlp.py:160-168); simple bounds (nlp.py:32-35,314-317) through a monotone log barrier.
  * the solver: damped Gauss-Newton (Levenberg-Marquardt on H + lam*diag(H)) with a
    fraction-to-the-boundary rule.  IPOPT itself (third party, version un-pinned, not in this
    image) is NOT restated; the iteration here only has to reach the same minimiser.

Parity status: PINNED for the vehicle shape by the reference's stored IPOPT outputs
`data/autonomous-car/filtering/nlp-l2.pkl` / `nlp-huber.pkl` (tests/golden/car_*.npz,
tests/test_oracle_car_golden.py); for the other shapes parity is UNPINNED (the reference has
no tests and no reproducible stored output for them) -- there this
restatement, validated on the vehicle shape, is the checker.

The Jacobian is formed densely and H = J^T W J is a plain matrix product: deliberately the
*opposite* organisation of the CUDA path (structured Kronecker assembly), so agreement of the
two is a meaningful check.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
import numpy as np

from nlp_filter_b200 import desc as D
from . import models

ST_CONVERGED = 0        # step below tolerance
ST_STALLED = 1          # no damping value produced a decrease (rounding-level stationarity)
ST_MAXITER = 2
ST_NUMERIC = 3          # non-finite cost at the starting point
ST_NOSTEP = 4           # no damping value produced a decrease at a NON-stationary point (failure)
# defaults of the damping schedule (include/nlpfilter.h, NLPF_LAM_*; nlpf_opts fields <= 0 select them)
LAM0 = 1e-12
LAM_JUMP = 1e-3         # a step that did not decrease the merit: lam <- max(10 lam, lam_jump)
LAM_JUMP_INDEF = 0.1    # a pivot that is not positive or a non-finite step: lam <- max(10 lam, lam_jump_indef)
LAM_DEC_FIRST = 1000.0  # accepted at the first damping value tried: lam / lam_dec_first (floor 1e-12); after rejections: lam / 10


def damping_schedule(opts):
    """(lam0, lam_jump, lam_jump_indef, lam_dec_first) with the defaults filled in."""
    pick = lambda v, dflt: v if v > 0.0 else dflt
    return (pick(opts.lam0, LAM0), pick(opts.lam_jump, LAM_JUMP), pick(opts.lam_jump_indef, LAM_JUMP_INDEF),
            pick(opts.lam_dec_first, LAM_DEC_FIRST))


STALL_GTOL = 1e-6       # stalled counts as stationary when max |g_i| / H_ii < STALL_GTOL (1 + |z|_inf)


class Window(object):
    """Per-window inputs (one estimation window / OCP instance)."""

    def __init__(self, desc, U=None, data=None, xprior=None, wshift=None):
        self.U = np.zeros((desc.K, max(desc.m, 1))) if U is None else np.asarray(U, float).reshape(desc.K, -1)
        self.data = np.zeros(desc.Ldata) if data is None else np.asarray(data, float).reshape(-1)
        self.xprior = np.zeros(desc.n) if xprior is None else np.asarray(xprior, float).reshape(-1)
        self.wshift = np.zeros((desc.K, desc.n)) if wshift is None else np.asarray(wshift, float).reshape(desc.K, desc.n)


def _node_quantities(desc, z, win):
    K, n, m, nv = desc.K, desc.n, desc.m, desc.nv
    V = z[:K * nv].reshape(K, nv)
    X = V[:, desc.cmap]
    if m > 0:
        U = np.array(win.U[:, :m], dtype=z.dtype)
        for i in range(m):
            if desc.ucol[i] >= 0:
                U[:, i] = V[:, desc.ucol[i]]
    else:
        U = np.zeros((K, 1), dtype=z.dtype)
    return V, X, U


def defects(desc, z, win):
    """W (K x n) and, row by row, its dense Jacobian (K*n x nz)."""
    K, n, m, nv, nz = desc.K, desc.n, desc.m, desc.nv, desc.nz
    V, X, U = _node_quantities(desc, z, win)
    W = desc.Dt @ X + win.wshift
    JW = np.zeros((K * n, nz), dtype=z.dtype)
    S = np.zeros((n, nv))
    S[np.arange(n), desc.cmap] = 1.0
    JW[:, :K * nv] = np.kron(desc.Dt, S)
    for k in range(K):
        f, F, Fu = models.dynamics(desc, X[k], U[k])
        W[k] -= f
        blk = JW[k * n:(k + 1) * n, k * nv:(k + 1) * nv]
        for i in range(n):
            blk[:, desc.cmap[i]] -= F[:, i]
        for i in range(m):
            if desc.ucol[i] >= 0:
                blk[:, desc.ucol[i]] -= Fu[:, i]
    return W, JW


def meas_residuals(desc, z, win, jac=True):
    """r_i = y_i - h_i, weights w_i and the dense Jacobian of r (M x nz); masked slots
    (weight 0: autonomous-car.py:260-263) are skipped, not multiplied by zero."""
    K, n, nv, ng, nz, M = desc.K, desc.n, desc.nv, desc.ng, desc.nz, desc.M
    r = np.zeros(M, dtype=z.dtype)
    wgt = np.zeros(M)
    Jr = np.zeros((M, nz), dtype=z.dtype) if jac else None
    if M == 0:
        return r, wgt, Jr
    V = z[:K * nv].reshape(K, nv)
    gl = z[K * nv:]
    Ve = desc.Phi @ V                              # interpolated node variables, nlp.py:264-269
    rec = lambda i, j: win.data[desc.roff[i] + j]
    cm = desc.cmap
    G = np.zeros((M, nv + ng), dtype=z.dtype)
    for rt in np.unique(desc.rtype):
        I = np.nonzero(desc.rtype == rt)[0]
        off = desc.roff[I]
        y = win.data[off]
        w = win.data[off + 1]
        live = w != 0.0
        I, off, y, w = I[live], off[live], y[live], w[live]
        if I.size == 0:
            continue
        idx = desc.ridx[I]
        x = Ve[desc.repoch[I]][:, cm]               # full state at the epoch of each residual
        rows = np.arange(I.size)
        xs = lambda c: x[rows, idx[:, c]]
        aux = lambda c: win.data[off + 2 + c]
        cols, grads = [], []
        if rt == D.RT_STATE:
            h = xs(0)
            cols = [cm[idx[:, 0]]]
            grads = [np.ones(I.size)]
        elif rt == D.RT_PR:
            d0, d1, d2 = xs(0) - aux(0), xs(1) - aux(1), xs(2) - aux(2)
            rho = np.sqrt(d0 ** 2 + d1 ** 2 + d2 ** 2)
            h = rho + xs(3)
            cols = [cm[idx[:, c]] for c in range(4)]
            grads = [d0 / rho, d1 / rho, d2 / rho, np.ones(I.size)]
        elif rt == D.RT_PRR:
            r0, r1, r2 = aux(0) - xs(0), aux(1) - xs(1), aux(2) - xs(2)
            rho = np.sqrt(r0 ** 2 + r1 ** 2 + r2 ** 2)
            e0, e1, e2 = r0 / rho, r1 / rho, r2 / rho
            v0, v1, v2 = aux(3) - xs(3), aux(4) - xs(4), aux(5) - xs(5)
            ev = e0 * v0 + e1 * v1 + e2 * v2
            h = ev + xs(6)
            cols = [cm[idx[:, c]] for c in range(7)]
            grads = [-(v0 - e0 * ev) / rho, -(v1 - e1 * ev) / rho, -(v2 - e2 * ev) / rho,
                     -e0, -e1, -e2, np.ones(I.size)]
        elif rt in (D.RT_RANGE, D.RT_RANGE_PT):
            dim = idx[:, 6]
            dd = []
            for c in range(3):
                on = dim > c
                if rt == D.RT_RANGE:
                    other = xs(3 + c)
                else:
                    gi = idx[:, 3 + c]
                    other = np.where(gi >= 0, gl[np.maximum(gi, 0)] if ng > 0 else 0.0, aux(c) if c < 3 else 0.0)
                dd.append(np.where(on, xs(c) - other, 0.0))
            h = np.sqrt(dd[0] ** 2 + dd[1] ** 2 + dd[2] ** 2 + .000001)
            for c in range(3):
                on = dim > c
                cols.append(cm[idx[:, c]])
                grads.append(np.where(on, dd[c] / h, 0.0))
                if rt == D.RT_RANGE:
                    cols.append(cm[idx[:, 3 + c]])
                    grads.append(np.where(on, -dd[c] / h, 0.0))
                else:
                    gi = idx[:, 3 + c]
                    cols.append(np.where(gi >= 0, nv + np.maximum(gi, 0), 0))
                    grads.append(np.where(on & (gi >= 0), -dd[c] / h, 0.0))
        elif rt in (D.RT_HEADING, D.RT_HEADING_PT):
            if rt == D.RT_HEADING:
                rx = xs(3) - xs(0) + .00001
                ry = xs(4) - xs(1)
            else:
                g0, g1 = idx[:, 3], idx[:, 4]
                p0 = np.where(g0 >= 0, gl[np.maximum(g0, 0)] if ng > 0 else 0.0, aux(0))
                p1 = np.where(g1 >= 0, gl[np.maximum(g1, 0)] if ng > 0 else 0.0, aux(1))
                rx = p0 - xs(0)
                ry = p1 - xs(1)
            d2 = rx * rx + ry * ry
            h = np.arctan2(rx, ry) if not np.iscomplexobj(rx) else -1j * np.log((ry + 1j * rx) / np.sqrt(d2))
            cols = [cm[idx[:, 0]], cm[idx[:, 1]]]
            grads = [-ry / d2, rx / d2]
            if rt == D.RT_HEADING:
                cols += [cm[idx[:, 3]], cm[idx[:, 4]]]
                grads += [ry / d2, -rx / d2]
            else:
                cols += [np.where(g0 >= 0, nv + np.maximum(g0, 0), 0), np.where(g1 >= 0, nv + np.maximum(g1, 0), 0)]
                grads += [np.where(g0 >= 0, ry / d2, 0.0), np.where(g1 >= 0, -rx / d2, 0.0)]
        else:
            raise ValueError("unknown residual type %r" % rt)
        r[I] = y - h
        wgt[I] = w
        if jac:
            for cc, gg in zip(cols, grads):
                np.add.at(G, (I, cc), gg)
    if jac:
        ph = desc.Phi[desc.repoch]                                     # M x K
        Jr[:, :K * nv] = -(ph[:, :, None] * G[:, None, :nv]).reshape(M, K * nv)
        if ng > 0:
            Jr[:, K * nv:] = -G[:, nv:]
    return r, wgt, Jr


def _huber(W, qinv, delta):
    """pseudo_huber_loss, cost_functions.py:25-31: value, half gradient, and the Gauss-Newton
    curvature weight of the residual sqrt(loss)*sign(W), i.e. (d sqrt(loss)/dW)^2 =
    q (s+1) / (2 s^2)  (equals q at W = 0, decays like q/(2s) for outliers)."""
    s = np.sqrt(1.0 + (W / delta) ** 2)
    val = 2.0 * qinv * delta ** 2 * (s - 1.0)
    return val, qinv * W / s, qinv * (s + 1.0) / (2.0 * s * s)


EPS = 2.220446049250313e-16


def second_order(desc, win, z, Hgn=None, info=None):
    """Exact second-order part of the dynamics-cost Hessian, sum_k sum_i lw_ki * (-hess f_i(X_k)),
    for the models whose kernels carry it (van der Pol, kinematic bicycle); zero otherwise.
    lw = weight * W is the half-gradient of the dynamics cost w.r.t. the defect.  Gauss-Newton
    alone converges only linearly on windows with large defects (up to 300 iterations on the
    4096-window van der Pol batch); with this term the tail disappears (max 18).
    Hgn (the Gauss-Newton matrix the term is added to): when given, every node's block is scaled back by
    t = min(1, min_i Hgn_ii / sum_j |S_ij|) so that the Gauss-Newton diagonal dominates its rows (Gershgorin) -- the
    kernels' safeguard against an indefinite H away from the solution; t = 1 near it, where lw is small.
    info (dict): info["clipped"] = some node had t < 1 (the solver then starts the iteration damped)."""
    K, n, nv = desc.K, desc.n, desc.nv
    S = np.zeros((desc.nz, desc.nz))
    if info is not None:
        info["clipped"] = False
    if desc.dyn_id not in (D.DYN_VDP, D.DYN_BICYCLE) or not np.any(desc.qinv != 0.0):
        return S
    W, _ = defects(desc, z, win)
    V, X, U = _node_quantities(desc, z, win)
    cm = desc.cmap
    for k in range(K):
        if desc.loss == D.LOSS_HUBER:
            _, hg, _ = _huber(W[k], desc.qinv, desc.delta)
            lw = desc.cw[k] * hg
        else:
            lw = desc.cw[k] * desc.qinv * W[k]
        o = k * nv
        if desc.dyn_id == D.DYN_VDP:        # f0 = (1 - x1^2) x0 - x1 + u:  hess f0 = [[0, -2 x1], [-2 x1, -2 x0]]
            x0, x1 = X[k]
            S[o + cm[0], o + cm[1]] += lw[0] * 2.0 * x1
            S[o + cm[1], o + cm[0]] += lw[0] * 2.0 * x1
            S[o + cm[1], o + cm[1]] += lw[0] * 2.0 * x0
        else:                               # f0 = v cos th, f1 = v sin th with th = x[2]
            v = models.BICYCLE_V_GAIN * U[k, 0] - models.BICYCLE_V_OFF
            th = X[k, 2]
            S[o + cm[2], o + cm[2]] += lw[0] * v * np.cos(th) + lw[1] * v * np.sin(th)
    if Hgn is not None:
        dgn = np.diag(Hgn)
        for k in range(K):
            o = k * nv
            r = np.sum(np.abs(S[o:o + nv, o:o + nv]), 1)
            nzr = r > 0.0
            if np.any(nzr):
                t = min(1.0, float(np.min(dgn[o:o + nv][nzr] / r[nzr])))
                S[o:o + nv, o:o + nv] *= t
                if info is not None and t < 1.0:
                    info["clipped"] = True
    return S


def evaluate(desc, win, z, mu=0.0, want_grad=True, with_noise=False):
    """Objective J (reference definition), merit = J/2 + barrier, and -- when want_grad --
    g = grad(merit), H = Gauss-Newton approximation of hess(merit).  with_noise appends a
    first-order bound on the rounding error of the merit (pseudoranges of 2e7 m minus predicted
    ranges lose ~9 digits, so the merit cannot resolve millimetre steps; the gradient can)."""
    K, n, nv, nz = desc.K, desc.n, desc.nv, desc.nz
    J = 0.0
    noise = 0.0
    g = np.zeros(nz)
    H = np.zeros((nz, nz))
    V = z[:K * nv].reshape(K, nv)
    # dynamics cost
    if np.any(desc.qinv != 0.0):
        W, JW = defects(desc, z, win)
        cwq = desc.cw[:, None] * np.ones((1, n))
        Vn, Xn, Un = _node_quantities(desc, z, win)
        mag = np.abs(desc.Dt) @ np.abs(Xn) + np.abs(desc.Dt @ Xn - W - win.wshift)      # sum|D x| + |f|
        if desc.loss == D.LOSS_HUBER:
            val, hg, hh = _huber(W, desc.qinv[None, :], desc.delta)
            J += float(np.sum(cwq * val))
            wg = (cwq * hg).reshape(-1)
            wh = (cwq * hh).reshape(-1)
            noise += float(np.sum(2.0 * np.abs(cwq * hg) * mag)) * EPS
        else:
            J += float(np.sum(cwq * desc.qinv[None, :] * W * W))
            wh = (cwq * desc.qinv[None, :]).reshape(-1)
            wg = wh * W.reshape(-1)
            noise += float(np.sum(2.0 * np.abs(cwq * desc.qinv[None, :] * W) * mag)) * EPS
        if want_grad:
            g += JW.T @ wg
            H += JW.T @ (wh[:, None] * JW)
    # measurement residual cost
    r, w, Jr = meas_residuals(desc, z, win, jac=want_grad)
    J += float(np.sum(w * r * r))
    if desc.M > 0:
        yy = win.data[desc.roff]
        noise += float(np.sum(2.0 * w * np.abs(r) * (2.0 * np.abs(yy) + np.abs(r)))) * EPS   # |y| + |h| <= 2|y| + |r|
    if want_grad and desc.M > 0:
        g += Jr.T @ (w * r)
        H += Jr.T @ (w[:, None] * Jr)
    # arrival cost on X_0
    if desc.has_prior:
        d0 = V[0, desc.cmap] - win.xprior
        J += float(np.sum(desc.pinv * d0 * d0))
        noise += float(np.sum(2.0 * desc.pinv * np.abs(d0) * (np.abs(V[0, desc.cmap]) + np.abs(win.xprior)))) * EPS
        if want_grad:
            for i in range(n):
                c = desc.cmap[i]
                g[c] += desc.pinv[i] * d0[i]
                H[c, c] += desc.pinv[i]
    # node (stage) cost  sum_k cq_k sum_i s_i V_ki^2
    if np.any(desc.snode != 0.0):
        wq = desc.cq[:, None] * desc.snode[None, :]
        J += float(np.sum(wq * V * V))
        if want_grad:
            g[:K * nv] += (wq * V).reshape(-1)
            H[np.arange(K * nv), np.arange(K * nv)] += wq.reshape(-1)
    merit = 0.5 * J
    # log barrier on the bounded node-variable components
    if mu > 0.0 and desc.nb > 0:
        for b in range(desc.nb):
            ii = np.arange(K) * nv + desc.bidx[b]
            if np.isfinite(desc.blo[b]):
                s = z[ii] - desc.blo[b]
                if np.any(s <= 0):
                    return (J, np.inf, g, H, noise) if with_noise else (J, np.inf, g, H)
                merit -= mu * float(np.sum(np.log(s)))
                g[ii] -= mu / s
                H[ii, ii] += mu / s ** 2
            if np.isfinite(desc.bhi[b]):
                s = desc.bhi[b] - z[ii]
                if np.any(s <= 0):
                    return (J, np.inf, g, H, noise) if with_noise else (J, np.inf, g, H)
                merit -= mu * float(np.sum(np.log(s)))
                g[ii] += mu / s
                H[ii, ii] += mu / s ** 2
    noise = 0.5 * noise + EPS * abs(merit)
    return (J, merit, g, H, noise) if with_noise else (J, merit, g, H)


def _max_step(desc, z, dz, tau):
    """Fraction-to-the-boundary step length."""
    K, nv = desc.K, desc.nv
    a = 1.0
    for b in range(desc.nb):
        ii = np.arange(K) * nv + desc.bidx[b]
        if np.isfinite(desc.blo[b]):
            neg = dz[ii] < 0
            if np.any(neg):
                a = min(a, tau * float(np.min((z[ii][neg] - desc.blo[b]) / (-dz[ii][neg]))))
        if np.isfinite(desc.bhi[b]):
            pos = dz[ii] > 0
            if np.any(pos):
                a = min(a, tau * float(np.min((desc.bhi[b] - z[ii][pos]) / dz[ii][pos])))
    return a


def push_inside(desc, z, push):
    K, nv = desc.K, desc.nv
    z = z.copy()
    for b in range(desc.nb):
        ii = np.arange(K) * nv + desc.bidx[b]
        lo = desc.blo[b] + push if np.isfinite(desc.blo[b]) else -np.inf
        hi = desc.bhi[b] - push if np.isfinite(desc.bhi[b]) else np.inf
        z[ii] = np.clip(z[ii], lo, hi)
    return z


def solve(desc, win, z0, opts=None):
    """Damped Gauss-Newton with monotone log barrier (DESIGN.md "Solver").  Returns
    (z, info) with info = dict(J, iters, status, lam, mu)."""
    opts = opts or D.SolveOpts()
    z = np.array(z0, dtype=float).reshape(-1)
    bounded = desc.nb > 0
    if bounded:
        z = push_inside(desc, z, opts.bound_push)
    mu = opts.mu0 if bounded else 0.0
    iters = 0
    status = ST_MAXITER
    lam0, lam_jump, lam_jump_indef, lam_dec_first = damping_schedule(opts)
    lam = lam0
    J, merit, g, H, nz0 = evaluate(desc, win, z, mu, with_noise=True)
    if not np.isfinite(merit):
        return z, dict(J=J, iters=0, status=ST_NUMERIC, lam=lam, mu=mu)
    while True:
        lam = lam0
        stage_steps = 0
        floor_steps = 0
        status = ST_MAXITER
        while iters < opts.max_iter and stage_steps < opts.max_stage_iter:
            if mu > opts.mu_min and stage_steps >= 1 and np.max(np.abs(g)) < 10.0 * mu:
                status = ST_CONVERGED
                break
            accepted = False
            floor = False
            small = False
            dg = np.diag(H).copy()                 # damping uses the (positive) Gauss-Newton diagonal
            dg = dg + 1e-14 * np.max(dg)
            so_info = {}
            Hn = H + second_order(desc, win, z, H, so_info)    # exact (safeguarded) Hessian where the model provides it
            if so_info["clipped"]:                 # the safeguard was needed: H is probably indefinite, no undamped try
                lam = max(lam, lam_jump_indef)
            lam_it0 = lam
            while lam <= 1e18:
                try:
                    L = np.linalg.cholesky(Hn + lam * np.diag(dg))
                    dz = -np.linalg.solve(L.T, np.linalg.solve(L, g))
                    ok = bool(np.all(np.isfinite(dz)))
                except np.linalg.LinAlgError:
                    ok = False
                if ok:
                    a = _max_step(desc, z, dz, opts.tau_ftb) if bounded else 1.0
                    zt = z + a * dz
                    # a first-try step below the tolerance ends the stage without an evaluation at the new point (the
                    # merit cannot resolve such a step); J stays the value at the previous iterate
                    if lam <= lam_it0 and a * float(np.max(np.abs(dz))) < opts.tol_step * (1.0 + float(np.max(np.abs(zt)))):
                        small = True
                        accepted = True
                        break
                    Jt, mt, _, _, nzt = evaluate(desc, win, zt, mu, want_grad=False, with_noise=True)
                    if np.isfinite(mt) and mt < merit:
                        accepted = True
                        break
                    # rounding floor of the merit: the predicted decrease is below what the merit
                    # can resolve, and the merit did not rise by more than its rounding bound
                    pred = -a * float(g @ dz)
                    nb = nz0 + nzt
                    if np.isfinite(mt) and (mt - merit) <= nb and pred <= 4.0 * nb:
                        floor = True      # accept: the Gauss-Newton step is still accurate (g is)
                        accepted = True
                        break
                lam = max(10.0 * lam, lam_jump if ok else lam_jump_indef)
            if not accepted:
                status = ST_STALLED
                with np.errstate(all="ignore"):
                    sg = np.max(np.abs(g) / (dg + 1e-300))
                if not (sg <= STALL_GTOL * (1.0 + float(np.max(np.abs(z))))):
                    status = ST_NOSTEP
                break
            step = a * float(np.max(np.abs(dz)))
            z = zt
            if small:
                iters += 1
                stage_steps += 1
                status = ST_CONVERGED
                break
            lam = max(lam / (lam_dec_first if lam <= lam_it0 else 10.0), 1e-12)
            iters += 1
            stage_steps += 1
            J, merit, g, H, nz0 = evaluate(desc, win, z, mu, with_noise=True)
            if step < opts.tol_step * (1.0 + float(np.max(np.abs(z)))):
                status = ST_CONVERGED
                break
            if floor:
                floor_steps += 1
                if floor_steps >= 5:
                    status = ST_CONVERGED
                    break
        if not bounded or mu <= opts.mu_min or iters >= opts.max_iter:
            break
        mu = max(opts.mu_min, min(0.2 * mu, mu ** 1.5))
        J, merit, g, H, nz0 = evaluate(desc, win, z, mu, with_noise=True)
    return z, dict(J=J, iters=iters, status=status, lam=lam, mu=mu)


def sample(desc, coll, z, t):
    """x(t) = sum_j phi_j(t) X_j (collocation.py:113-121) -> (len(t), n)."""
    K, nv = desc.K, desc.nv
    V = np.asarray(z)[:K * nv].reshape(K, nv)
    return coll.phi(t) @ V[:, desc.cmap]
