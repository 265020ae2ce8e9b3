"""The reference's problem-building API (nlp/nlp.py:8-317) over the B200 kernels.

Same classes, method names, argument meaning and return values as the reference, so its
scripts (single_integrator.py, van_der_pol.py, estimation_example.py, gnss_stationary.py,
rc-car.py, multi-receiver.py) run unmodified.  What differs is everything underneath:

  reference                                   here
  ------------------------------------------  ------------------------------------------------
  casadi.Opti variables / parameters          light handles (nlp/_sym.py), values kept in NumPy
  symbolic graph + AD (CasADi)                registered model ids with hand-derived Jacobians
  IPOPT + MUMPS per solve() (nlp.py:82)       one launch of the whole-solve kernel (nlpf_solve)
  free process-noise variables W_k            eliminated: W_k = (2/T) sum_j D_kj X_j - f(X_k,U_k)
  hard equalities of the OCP class            augmented Lagrangian around the same kernel
  simple bounds                               monotone log barrier inside the kernel

`build()` lowers the recorded problem to a flat `ProblemDesc` (nlp_filter_b200/desc.py);
`solve()` packs the current parameter values into the per-window buffers and calls the C ABI.
There is no CPU fallback: `solve()` raises if libnlpfilter.so or a CUDA device is missing.
Model callables must be the registered ones (nlp/dynamics.py, nlp/measurements.py,
nlp/cost_functions.py) or, for dynamics, any function affine in (x, u).
"""
import time

import numpy as np
from scipy.interpolate import interp1d

from . import collocation
from . import constraints  # noqa: F401  (the reference imports it here as well)
from . import cost_functions as _cf
from . import dynamics as _dyn
from . import measurements as _meas
from ._sym import Variable, Parameter, ParamElem, ParamPool, Elem, LinExpr
from .. import desc as D

_INF = float("inf")


class _Solution(object):
    """What `problem.sol` holds after a solve (stand-in for casadi OptiSol)."""

    def __init__(self, values):
        self._values = values          # var name -> ndarray

    def value(self, var):
        return self._values[var.name]

    def value_variables(self):
        return dict(self._values)


class _OptiStub(object):
    """`problem.opti` exists in the reference; scripts do not touch it.  Kept for attribute parity."""

    def __repr__(self):
        return "<nlp_filter_b200 problem store (no CasADi)>"


def _as_diag(M, n, what):
    M = np.asarray(M, dtype=float)
    if M.ndim == 0:
        return np.full(n, float(M))
    if M.ndim == 1:
        if M.shape[0] == 1 and n > 1:
            return np.full(n, float(M[0]))
        return M.reshape(n)
    if np.any(np.abs(M - np.diag(np.diag(M))) > 0):
        raise NotImplementedError("%s must be diagonal (the kernels weight each component separately)" % what)
    return np.diag(M).copy()


class NLP(object):
    """NLP class"""

    MAX_ITER = 3000

    def __init__(self, N):
        self.N = N
        self.var_names = []
        self.var_sizes = []
        self.w = {}
        self.opti = _OptiStub()
        self.sol = None
        self.solver = {}
        self._groups = {}             # group name -> [Variable]
        self._var_bounds = []         # (Variable, lb array, ub array) from addVariables
        self._eq = []                 # LinExpr == 0
        self._desc = None
        self._engine_problem = None
        self._built = False
        self.J = 0

    # ---- variables / parameters (nlp.py:18-56) ---------------------------------------------
    def addVariables(self, N_var, n_var, lb=None, ub=None, name='x'):
        X = []
        for i in range(N_var):
            v = Variable(name, len(self._groups.get(name, [])), n_var)
            self._groups.setdefault(name, []).append(v)
            self.var_names.append(v.name)
            self.var_sizes.append(n_var)
            self.w[v.name] = v
            X.append(v)
            if lb is not None or ub is not None:
                lo = np.full(n_var, -_INF) if lb is None else np.asarray(lb, dtype=float).reshape(n_var)
                hi = np.full(n_var, _INF) if ub is None else np.asarray(ub, dtype=float).reshape(n_var)
                self._var_bounds.append((v, lo, hi))
        self._built = False
        return X

    def addParameter(self, N_var, n_var, val=None):
        pool = self._param_pool()
        off0 = pool.alloc(N_var * n_var)                    # one contiguous block per call (vectorised uploads)
        P = []
        for i in range(N_var):
            p = Parameter(n_var, pool, off0 + i * n_var, (off0, N_var, n_var, i))
            P.append(p)
            if val is not None:
                self.setParameter(p, val)
        return P

    def _param_pool(self):
        if getattr(self, "_pool", None) is None:
            self._pool = ParamPool()
        return self._pool

    def addIneqConstraint(self, g, arguments, params=None):
        raise NotImplementedError("general inequality constraints are not lowered to the kernels; use the bound "
                                  "arguments of addVariables / addVarBounds")

    def addEqConstraint(self, h, arguments, params=None):
        e = h(arguments, params)
        self._eq.append(LinExpr.of(e))
        self._built = False

    def setParameter(self, p, val):
        if isinstance(p, (Parameter, ParamElem)):
            p.set(val)
        else:
            raise TypeError("setParameter expects a parameter returned by addParameter/addResidualCost")

    def setObjective(self):
        pass

    def build(self, verbose=True):
        self.setObjective()
        self._desc = self._compile()
        self._pack = None                     # gather plan of the data block: rebuilt lazily for the new residual table
        self._engine_problem = None
        self._built = True

    def initialGuess(self, x, x0):
        x.init = np.asarray(x0, dtype=float).reshape(-1).copy()

    # ---- solve / extraction (nlp.py:76-119) ------------------------------------------------
    def solve(self, warmstart=False):
        if not self._built:
            self.build()
        if warmstart and self.sol is not None:
            print('Warmstarting with previous solution')
            z0 = self._z_last.copy()
        else:
            z0 = self._initial_z()
        t0 = time.perf_counter()
        z, info = self._solve_device(z0)
        wall = time.perf_counter() - t0
        # 0 converged; 1 stalled AT a stationary point (the kernel checked the scaled gradient); 2 iteration limit;
        # 3 non-finite objective at the start; 4 no acceptable step at a non-stationary point -- IPOPT's names
        ok = info["status"] in (0, 1)
        names = {0: "Solve_Succeeded", 1: "Solve_Succeeded", 2: "Maximum_Iterations_Exceeded",
                 3: "Invalid_Number_Detected", 4: "Error_In_Step_Computation"}
        self.solver = {"t_wall_total": wall, "t_proc_total": wall, "iter_count": int(info["iters"]),
                       "return_status": names.get(int(info["status"]), "Internal_Error"),
                       "success": bool(ok), "objective": float(info["J"])}
        if not ok:
            raise RuntimeError("solver did not converge (status %d after %d iterations)" % (info["status"], info["iters"]))
        self._z_last = z
        self.sol = _Solution(self._unpack(z))

    def extractVariableValue(self, name, idx=0):
        var_name = name + '_' + str(idx)
        if var_name in self.w.keys():
            if self.sol is not None:
                return np.array(self.sol.value(self.w[var_name])).reshape(-1)
            else:
                print('Need to run solve() first')
                return None
        else:
            print('Could not find ' + name + '_' + str(idx) + '.')
            return None

    def extractSolution(self, name, t_array):
        X = []
        for i in range(self.N + 1):
            value = self.extractVariableValue(name, i)
            if value is None:
                return None
            X.append(value)
        T = len(t_array)
        sol = np.zeros((T, X[0].shape[0]))
        for t in range(T):
            sol[t, :] = self.CPM.evaluateSolution(t_array[t], X)
        return sol

    # ---- device plumbing -------------------------------------------------------------------
    def _problem_handle(self):
        if self._engine_problem is None:
            from .. import engine
            self._engine_problem = engine.WindowProblem(self._desc)
        return self._engine_problem

    def _run_kernel(self, z0, U, data, xprior, wshift, opts=None):
        from .. import engine
        import torch
        prob = self._problem_handle()
        dev = lambda a: None if a is None else torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64).reshape(1, -1)).cuda()
        z = dev(z0)
        if opts is None:            # a single problem may iterate as long as IPOPT would (its max_iter default is 3000)
            # a single problem is latency-, not throughput-bound: the tight step tolerance (the batch default is 1e-9)
            opts = D.SolveOpts(max_iter=self.MAX_ITER, max_stage_iter=self.MAX_ITER, tol_step=1e-11)
        _, cost, iters, status = prob.solve(z, dev(U), dev(data), dev(xprior), dev(wshift), opts)
        torch.cuda.synchronize()
        return z.cpu().numpy()[0], dict(J=float(cost.item()), iters=int(iters.item()), status=int(status.item()))


# =============================================================================================
class fixedTimeOptimalControlNLP(NLP):
    """Optimal control on one Chebyshev interval (nlp.py:122-186): hard collocation equalities,
    quadrature stage cost, end-point conditions, bounds."""

    RHO = 1.0e4            # augmented-Lagrangian weight of the equalities
    CON_TOL = 1.0e-10      # accepted equality violation
    MAX_OUTER = 25

    def __init__(self, N, T, n, m):
        super(fixedTimeOptimalControlNLP, self).__init__(N)
        self.T = T
        self.n = n
        self.m = m
        self.CPM = collocation.ChebyshevPseudospectralMethod(self.N, 0, T)
        self.J = 0
        self._dyn = None
        self._stage = None
        self._conds = []        # (Variable, values)

    def addDynamics(self, func, X, U, params=None):
        if len(X) != self.N + 1 or len(U) != self.N + 1:
            print('X and U must have N+1 points defined.')
        self._dyn = (func, list(X), list(U), params)
        self._built = False

    def addStageCost(self, func, X, U, params=None):
        self._stage = (func, list(X), list(U), params)
        self._built = False

    def addSingleConstraint(self, func, variables, params=None):
        self._eq.append(LinExpr.of(func(variables, params)))
        self._built = False

    def addInitialCondition(self, x0, values):
        self._conds.append((x0, np.asarray(values, dtype=float).reshape(-1)))
        self._built = False

    def addTerminalCondition(self, xT, values):
        self._conds.append((xT, np.asarray(values, dtype=float).reshape(-1)))
        self._built = False

    # ---- lowering --------------------------------------------------------------------------
    def _compile(self):
        if self._dyn is None:
            raise RuntimeError("addDynamics has not been called")
        func, X, U, params = self._dyn
        K, n, m = self.N + 1, self.n, self.m
        d = D.ProblemDesc(K, n, m, self.T)
        _lower_dynamics(d, func, n, m, params)
        d.resize_nv(n + m)
        d.cmap = np.arange(n, dtype=np.int32)
        d.ucol = (n + np.arange(max(m, 1))).astype(np.int32)
        d.Dt = (2.0 / self.T) * self.CPM.D
        d.cw = np.ones(K)                                   # equalities are not quadrature weighted
        d.cq = (self.T / 2.0) * self.CPM.w                  # nlp.py:167-168
        d.qinv = np.full(n, self.RHO)
        self._xgroup, self._ugroup = X[0].group, U[0].group
        if self._stage is not None:
            d.snode = _diag_quadratic(self._stage[0], n, m, self._stage[3])
        # end-point conditions -> RT_STATE residuals on unit basis rows (exact at the nodes)
        rt, ep, ix, rows = [], [], [], []
        self._cond_vals = []
        for var, vals in self._conds:
            if not isinstance(var, Variable) or var.group != self._xgroup:
                raise NotImplementedError("conditions are supported on state variables (X[0], X[N])")
            row = np.zeros(K)
            row[var.index] = 1.0
            rows.append(row)
            for i in range(n):
                rt.append(D.RT_STATE)
                ep.append(len(rows) - 1)
                ix.append([i] + [0] * (D.NIDX - 1))
                self._cond_vals.append(vals[i])
        d.E = len(rows)
        d.Phi = np.array(rows).reshape(d.E, K)
        d.set_residuals(rt, ep, np.array(ix, dtype=np.int32).reshape(-1, D.NIDX))
        if self._eq:
            raise NotImplementedError("addSingleConstraint / addEqConstraint are not lowered for the OCP class")
        # bounds (nlp.py:32-35)
        lo, hi = np.full(n + m, -_INF), np.full(n + m, _INF)
        for v, lb, ub in self._var_bounds:
            off = 0 if v.group == self._xgroup else n
            lo[off:off + v.n] = np.maximum(lo[off:off + v.n], lb)
            hi[off:off + v.n] = np.minimum(hi[off:off + v.n], ub)
        bcols = [c for c in range(n + m) if np.isfinite(lo[c]) or np.isfinite(hi[c])]
        d.bidx = np.array(bcols, dtype=np.int32)
        d.blo, d.bhi = lo[bcols], hi[bcols]
        return d.finalize()

    def _initial_z(self):
        K, n, m = self.N + 1, self.n, self.m
        z = np.zeros((K, n + m))
        for v in self._groups.get(self._xgroup, []):
            if v.init is not None:
                z[v.index, :n] = v.init
        for v in self._groups.get(self._ugroup, []):
            if v.init is not None:
                z[v.index, n:] = v.init
        return z.reshape(-1)

    def _defects(self, z):
        func, X, U, params = self._dyn
        K, n, m = self.N + 1, self.n, self.m
        V = z.reshape(K, n + m)
        W = (2.0 / self.T) * (self.CPM.D @ V[:, :n])
        for k in range(K):
            # u is always handed over as a 1-D array, as the reference's MX.sym('U', m) is (nlp.py:143-148) and as
            # _lower_dynamics probed the callable
            W[k] -= _call_dyn(func, V[k, :n], V[k, n:], m, params)
        return W

    def _solve_device(self, z0):
        """Method of multipliers around the whole-solve kernel: equalities g(z) = 0 enter as
        rho |g + lambda/rho|^2 (the kernel's dynamics-cost and residual slots), lambda += rho g."""
        d = self._desc
        K, n = d.K, d.n
        lamW = np.zeros((K, n))
        vals = np.array(self._cond_vals, dtype=float)
        lamC = np.zeros(vals.shape[0])
        z = z0
        iters = 0
        info = dict(J=0.0, iters=0, status=2)
        for outer in range(self.MAX_OUTER):
            data = np.zeros(d.Ldata)
            for r in range(d.M):
                data[d.roff[r]] = vals[r] - lamC[r] / self.RHO
                data[d.roff[r] + 1] = self.RHO
            z, info = self._run_kernel(z, None, data, None, lamW / self.RHO)
            iters += info["iters"]
            if info["status"] > 1:
                break
            W = self._defects(z)
            V = z.reshape(K, d.nv)
            c = np.array([V[int(np.argmax(d.Phi[d.repoch[r]])), d.ridx[r, 0]] - vals[r] for r in range(d.M)])
            viol = max(np.max(np.abs(W)), np.max(np.abs(c)) if c.size else 0.0)
            lamW += self.RHO * W
            lamC += self.RHO * c
            if viol < self.CON_TOL:
                break
        else:
            info["status"] = 2
        V = z.reshape(K, d.nv)
        info["J"] = float(np.sum(d.cq[:, None] * d.snode[None, :] * V * V))      # the reference's objective
        info["iters"] = iters
        self.J = info["J"]
        return z, info

    def solve_batch(self, x0_batch, xf_batch=None, opts=None):
        """Monte-Carlo over the end-point conditions (addition to the reference API; SURVEY.md section 8 row A8:
        "batching = Monte-Carlo over x0"): B copies of this optimal-control problem that differ only in the value
        of the FIRST condition added (addInitialCondition, nlp.py:170-174: x0_batch (B, n)) and optionally of the
        second (addTerminalCondition, nlp.py:176-180: xf_batch (B, n)), solved together -- every whole-solve
        launch, defect evaluation and multiplier update acts on all B problems.  Same method of multipliers as
        solve() on one problem.  Returns (X (B, K, n), U (B, K, m), J (B,), status (B,))."""
        import torch
        from .. import engine
        if not self._built:
            self.build()
        d = self._desc
        K, n, m = d.K, d.n, self.m
        prob = self._problem_handle()
        x0_batch = np.asarray(x0_batch, dtype=float).reshape(-1, n)
        B = x0_batch.shape[0]
        vals = np.tile(np.array(self._cond_vals, dtype=float), (B, 1))
        vals[:, :n] = x0_batch
        if xf_batch is not None:
            vals[:, n:2 * n] = np.asarray(xf_batch, dtype=float).reshape(B, n)
        dev = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).cuda()
        vals_d = dev(vals)
        node = torch.as_tensor(np.array([int(np.argmax(d.Phi[d.repoch[r]])) for r in range(d.M)])).cuda()
        comp = torch.as_tensor(d.ridx[:, 0].astype(np.int64)).cuda()
        roff = torch.as_tensor(d.roff.astype(np.int64)).cuda()
        z = dev(np.tile(self._initial_z(), (B, 1)))
        lamW = torch.zeros((B, K, n), dtype=torch.float64, device="cuda")
        lamC = torch.zeros((B, d.M), dtype=torch.float64, device="cuda")
        data = torch.zeros((B, max(d.Ldata, 1)), dtype=torch.float64, device="cuda")
        if opts is None:
            opts = D.SolveOpts(max_iter=self.MAX_ITER, max_stage_iter=self.MAX_ITER, tol_step=1e-11)
        status = torch.full((B,), 2, dtype=torch.int32, device="cuda")
        done = torch.zeros((B,), dtype=torch.bool, device="cuda")
        for outer in range(self.MAX_OUTER):
            data[:, roff] = vals_d - lamC / self.RHO
            data[:, roff + 1] = self.RHO
            _, cost, iters, st = prob.solve(z, None, data, None, (lamW / self.RHO).contiguous(), opts)
            W = prob.defects(z)
            V = z.view(B, K, d.nv)
            c = V[:, node, comp] - vals_d
            viol = torch.maximum(W.abs().amax((1, 2)), c.abs().amax(1) if d.M else torch.zeros_like(cost))
            failed = (st > 1) & ~done
            status[failed] = st[failed]
            upd = ~done & ~failed
            lamW += self.RHO * W * upd[:, None, None]        # finished problems keep their multipliers (and their solution)
            lamC += self.RHO * c * upd[:, None]
            ok = upd & (viol < self.CON_TOL)
            status[ok] = 0
            done |= ok | failed
            if bool(done.all().item()):
                break
        V = z.view(B, K, d.nv).cpu().numpy()
        J = np.sum(d.cq[None, :, None] * d.snode[None, None, :] * V * V, axis=(1, 2))      # the reference's objective
        return V[:, :, :n].copy(), V[:, :, n:].copy(), J, status.cpu().numpy()

    def _unpack(self, z):
        K, n, m = self.N + 1, self.n, self.m
        V = z.reshape(K, n + m)
        out = {}
        for v in self._groups.get(self._xgroup, []):
            out[v.name] = V[v.index, :n].copy()
        for v in self._groups.get(self._ugroup, []):
            out[v.name] = V[v.index, n:].copy()
        return out


# =============================================================================================
class fixedTimeOptimalEstimationNLP(NLP):
    """Moving-horizon estimation on one Chebyshev interval (nlp.py:189-317)."""

    def __init__(self, N, T, n, m):
        super(fixedTimeOptimalEstimationNLP, self).__init__(N)
        self.T = T
        self.n = n
        self.m = m
        self.CPM = collocation.ChebyshevPseudospectralMethod(self.N, 0, T)
        self.J = 0
        self._dyn = None
        self._dyncost = None
        self._res = []           # residual records
        self._prior = None
        self._bounds = []        # (idx, lb, ub) from addVarBounds

    # ---- model building (nlp.py:202-286) ---------------------------------------------------
    def addDynamics(self, func, X, t_array=None, u_array=None, params=None):
        if len(X) != self.N + 1:
            print('X must have N+1 points defined.')
        if self.m != 0:
            U = self.addParameter(self.N + 1, self.m)
        else:
            print('No control input being used for dynamics.')
            U = None
        W = self.addVariables(self.N + 1, self.n, name='w')
        self._dyn = (func, list(X), U, W, params)
        if u_array is not None:
            self.setControl(U, t_array, u_array)
        self._built = False
        return U, W

    def addDynamicsCost(self, cost_function, W, params=None):
        self._dyncost = (cost_function, params)
        self._built = False

    def addResidualCost(self, measurement_model, X, t_array, y_array, R, params=None):
        t_array = np.asarray(t_array, dtype=float)
        N_measurements = t_array.shape[0]
        p = y_array.shape[0] if y_array is not None else params["p"]
        Y = self.addParameter(N_measurements, p)
        for i in range(N_measurements):
            t = float(np.asarray(t_array[i]).reshape(-1)[0])
            self._res.append(dict(model=measurement_model, t=t, Y=Y[i], p=p, R=R, params=params))
        if y_array is not None:
            self.setMeasurement(Y, t_array, y_array)
        self._built = False
        return Y

    def addInitialCost(self, cost_function, X0, params=None, x0=None):
        x0_guess = self.addParameter(1, self.n)[0]
        self._prior = (cost_function, X0, params, x0_guess)
        if x0 is not None:
            self.setParameter(x0_guess, x0)
        self._built = False
        return x0_guess

    def initializeEstimate(self, X, t_array, xhat_array):
        t_nodes = self.CPM.tau2t(self.CPM.tau)
        interp = interp1d(t_array, xhat_array, fill_value="extrapolate")
        xhat_nodes = interp(t_nodes)
        for (i, x) in enumerate(X):
            self.initialGuess(x, xhat_nodes[:, i])

    def setControl(self, U, t_array, u_array):
        u_t = interp1d(t_array, u_array, fill_value="extrapolate")
        for k in range(self.N + 1):
            self.setParameter(U[k], u_t(self.CPM.tau2t(self.CPM.tau[k])))

    def setMeasurement(self, Y, t_array, y_array):
        """nlp/nlp.py:310-312.  The reference sets one parameter per column in a Python loop; here the parameters an
        addResidualCost / addParameter call returned are one block of the parameter pool, so the whole (p x N) matrix
        is uploaded with a single array assignment (any other list of parameters falls back to the per-column loop)."""
        y_array = np.asarray(y_array, dtype=float)
        n_col = len(t_array)
        if n_col and self._is_whole_block(Y, n_col) and y_array.shape == (Y[0].n, n_col):
            off0, N, n, _ = Y[0].block
            self._pool.buf[off0:off0 + N * n].reshape(N, n)[:] = y_array.T
            for y in Y:
                y.is_set = True
            return
        for i in range(n_col):
            self.setParameter(Y[i], y_array[:, i])

    def setMeasurements(self, Ys, y_arrays):
        """Vectorised upload of several measurement groups at once (addition to the reference API): Ys[g] is the list
        addResidualCost returned for group g, y_arrays[g] its (p x N_g) matrix."""
        for Y, y in zip(Ys, y_arrays):
            self.setMeasurement(Y, range(len(Y)), y)

    def _is_whole_block(self, Y, n_col):
        b = getattr(Y[0], "block", None)
        if b is None or b[1] != n_col or len(Y) != n_col or Y[0].pool is not getattr(self, "_pool", None):
            return False
        return all(getattr(y, "block", None) is not None and y.block[0] == b[0] and y.block[3] == i for i, y in enumerate(Y))

    def addVarBounds(self, X, idx, lb, ub):
        if len(X) != self.N + 1:
            raise NotImplementedError("addVarBounds is lowered for bounds on every node (got %d of %d)" % (len(X), self.N + 1))
        self._bounds.append((int(idx), float(lb), float(ub)))
        self._built = False

    # ---- lowering --------------------------------------------------------------------------
    def _compile(self):
        if self._dyn is None:
            raise RuntimeError("addDynamics has not been called")
        func, X, U, W, params = self._dyn
        K, n, m = self.N + 1, self.n, self.m
        d = D.ProblemDesc(K, n, m, self.T)
        _lower_dynamics(d, func, n, m, params)
        d.Dt = (2.0 / self.T) * self.CPM.D
        d.cw = (self.T / 2.0) * self.CPM.w                  # nlp.py:244-245
        d.cq = (self.T / 2.0) * self.CPM.w
        self._xgroup = X[0].group
        # global (non-node) unknowns, e.g. XA of multi-receiver.py:73
        self._globals = []
        ng = 0
        for gname, vs in self._groups.items():
            if gname in (self._xgroup, 'w'):
                continue
            for v in vs:
                self._globals.append((v, ng))
                ng += v.n
        d.ng = ng
        # linear equalities X_k[a] - X_k[b] == 0 at every node -> shared column
        cmap = np.arange(n, dtype=np.int32)
        if self._eq:
            pairs = {}
            for e in self._eq:
                tm = e.nonzero_terms()
                items = sorted(tm.items(), key=lambda kv: kv[0][1])
                if len(items) != 2 or e.const != 0.0 or items[0][1] + items[1][1] != 0.0 or \
                        items[0][0][0] is not items[1][0][0] or items[0][0][0].group != self._xgroup:
                    raise NotImplementedError("only equalities X_k[a] - X_k[b] == 0 on the node states are lowered")
                node = items[0][0][0].index
                pairs.setdefault((items[0][0][1], items[1][0][1]), set()).add(node)
            # union-find over the components: chained pairs (a,b), (b,c) end in one class whatever their order
            root = list(range(n))

            def find(i):
                while root[i] != i:
                    root[i] = root[root[i]]
                    i = root[i]
                return i
            for (a, b), nodes in pairs.items():
                if nodes != set(range(K)):
                    raise NotImplementedError("an equality must be imposed at every node to be folded")
                ra, rb = find(a), find(b)
                if ra != rb:
                    root[max(ra, rb)] = min(ra, rb)
            cmap = np.array([find(i) for i in range(n)], dtype=np.int32)
            used = sorted(set(cmap.tolist()))
            remap = {c: i for i, c in enumerate(used)}
            cmap = np.array([remap[c] for c in cmap], dtype=np.int32)
            d.resize_nv(len(used))
        d.cmap = cmap
        # dynamics cost (nlp.py:242-245)
        if self._dyncost is not None:
            cf, cp = self._dyncost
            kind = _cf.KERNEL_COSTS.get(cf)
            if kind is None:
                raise NotImplementedError("dynamics cost %r is not a registered cost function" % (cf,))
            if kind[0] == "l2":
                d.qinv = np.ones(n)
            else:
                d.qinv = _as_diag(cp["Q"], n, "the dynamics-cost weight Q")
            d.loss = kind[1]
            if kind[0] == "huber":
                d.delta = float(cp["delta"])
        # arrival cost (nlp.py:279-286)
        if self._prior is not None:
            cf, X0, cp, x0p = self._prior
            kind = _cf.KERNEL_COSTS.get(cf)
            if kind is None or kind[1] != D.LOSS_L2:
                raise NotImplementedError("the arrival cost must be l2_norm or weighted_l2_norm")
            if not isinstance(X0, Variable) or X0.group != self._xgroup or X0.index != 0:
                raise NotImplementedError("the arrival cost is lowered for X[0]")
            d.has_prior = 1
            d.pinv = np.ones(n) if kind[0] == "l2" else _as_diag(cp["Q"], n, "the arrival-cost weight")
        # bounds
        lo, hi = np.full(d.nv, -_INF), np.full(d.nv, _INF)
        for idx, lb, ub in self._bounds:
            lo[cmap[idx]] = max(lo[cmap[idx]], lb)
            hi[cmap[idx]] = min(hi[cmap[idx]], ub)
        for v, lb, ub in self._var_bounds:
            if v.group != self._xgroup:
                if np.any(np.isfinite(lb)) or np.any(np.isfinite(ub)):
                    raise NotImplementedError("bounds are lowered for the node states only")
                continue
            for i in range(n):
                lo[cmap[i]] = max(lo[cmap[i]], lb[i])
                hi[cmap[i]] = min(hi[cmap[i]], ub[i])
        bcols = [c for c in range(d.nv) if np.isfinite(lo[c]) or np.isfinite(hi[c])]
        d.bidx = np.array(bcols, dtype=np.int32)
        d.blo, d.bhi = lo[bcols], hi[bcols]
        # residual table (nlp.py:247-277)
        self._lower_residuals(d)
        return d.finalize()

    def _global_index(self, var):
        for v, off in self._globals:
            if v is var:
                return off
        raise NotImplementedError("variable %r is not a global unknown of this problem" % (var,))

    def _lower_residuals(self, d):
        times, rows = [], []
        rt, ep, ix = [], [], []
        self._rsrc = []              # per scalar residual: (y source, weight source, [aux sources])

        def epoch_of(t):
            for e, te in enumerate(times):
                if abs(te - t) <= 1e-12 * max(1.0, abs(t)):
                    return e
            times.append(t)
            rows.append(self.CPM.evaluateLagrangePolynomials(t))
            return len(times) - 1

        def src(x, i):
            if isinstance(x, Parameter):
                return (x, i)
            return float(np.asarray(x, dtype=float).reshape(-1)[i])

        for rec in self._res:
            model, params, p, Y, R = rec["model"], rec["params"] or {}, rec["p"], rec["Y"], rec["R"]
            kind = _meas.KERNEL_TYPES.get(model)
            if kind is None:
                raise NotImplementedError("measurement model %r is not registered (nlp/measurements.py)" % (model,))
            e = epoch_of(rec["t"])
            if isinstance(R, Parameter):
                if p != 1:
                    raise NotImplementedError("a parameter weight R is lowered for scalar measurements")
                wsrc = [(R, 0)]
            else:
                wsrc = list(_as_diag(R, p, "the measurement weight R"))
            none = [0] * D.NIDX
            if kind == "state":
                for i in range(p):
                    rt.append(D.RT_STATE); ep.append(e); ix.append([i] + none[1:])
                    self._rsrc.append(((Y, i), wsrc[i], []))
                continue
            if p != 1:
                raise NotImplementedError("%s is a scalar measurement model" % model.__name__)
            if kind in ("pr", "vpr"):
                idx = [0, 1, 8, 6] if kind == "vpr" else list(params.get("idx", [0, 1, 2, 3]))
                rt.append(D.RT_PR); ep.append(e); ix.append(idx + none[4:])
                self._rsrc.append(((Y, 0), wsrc[0], [src(params["sat_pos"], c) for c in range(3)]))
            elif kind == "prr":
                rt.append(D.RT_PRR); ep.append(e); ix.append([0, 1, 2, 4, 5, 6, 7, 0])
                self._rsrc.append(((Y, 0), wsrc[0], [src(params["sat_pos"], c) for c in range(3)] +
                                   [src(params["sat_vel"], c) for c in range(3)]))
            elif kind in ("range2", "range3"):
                dim = 2 if kind == "range2" else 3
                if "y" in params:
                    idx = list(params.get("idx", list(range(dim))))
                    tgt = params["y"]
                    if isinstance(tgt, Variable):
                        g0 = self._global_index(tgt)
                        gi = [g0 + c for c in range(dim)] + [-1] * (3 - dim)
                        aux = [0.0, 0.0, 0.0]
                    else:
                        gi = [-1, -1, -1]
                        aux = [src(tgt, c) for c in range(dim)] + [0.0] * (3 - dim)
                    rt.append(D.RT_RANGE_PT); ep.append(e)
                    ix.append(idx[:dim] + [0] * (3 - dim) + gi + [dim, 0])
                    self._rsrc.append(((Y, 0), wsrc[0], aux))
                else:
                    a, b = list(params["idxA"]), list(params["idxB"])
                    rt.append(D.RT_RANGE); ep.append(e)
                    ix.append(a[:dim] + [0] * (3 - dim) + b[:dim] + [0] * (3 - dim) + [dim, 0])
                    self._rsrc.append(((Y, 0), wsrc[0], []))
            elif kind == "heading":
                if "y" in params:
                    idx = list(params.get("idx", [0, 1]))
                    tgt = params["y"]
                    if isinstance(tgt, Variable):
                        g0 = self._global_index(tgt)
                        gi, aux = [g0, g0 + 1, -1], [0.0, 0.0, 0.0]
                    else:
                        gi, aux = [-1, -1, -1], [src(tgt, 0), src(tgt, 1), 0.0]
                    rt.append(D.RT_HEADING_PT); ep.append(e); ix.append(idx[:2] + [0] + gi + [2, 0])
                    self._rsrc.append(((Y, 0), wsrc[0], aux))
                else:
                    a, b = list(params["idxA"]), list(params["idxB"])
                    rt.append(D.RT_HEADING); ep.append(e); ix.append(a[:2] + [0] + b[:2] + [0, 2, 0])
                    self._rsrc.append(((Y, 0), wsrc[0], []))
        d.E = len(times)
        d.Phi = np.array(rows).reshape(d.E, d.K)
        d.set_residuals(rt, ep, np.array(ix, dtype=np.int32).reshape(-1, D.NIDX))

    # ---- per-solve buffers -----------------------------------------------------------------
    def _pack_plan(self):
        """Gather plan of the per-window data block, built once per build(): constants go into a template, every slot
        fed by a parameter becomes (destination index, pool offset).  Parameters that do not live in this problem's
        pool (created outside addParameter) keep the per-slot path."""
        d = self._desc
        const = np.zeros(d.Ldata)
        dst, src, slow = [], [], []
        pool = getattr(self, "_pool", None)
        for r, (ys, ws, aux) in enumerate(self._rsrc):
            o = int(d.roff[r])
            for c, sdesc in enumerate([ys, ws] + list(aux)):
                if isinstance(sdesc, tuple):
                    par, i = sdesc
                    if pool is not None and par.pool is pool:
                        dst.append(o + c); src.append(par.off + i)
                    else:
                        slow.append((o + c, par, i))
                else:
                    const[o + c] = sdesc
        self._pack = (const, np.asarray(dst, dtype=np.int64), np.asarray(src, dtype=np.int64), slow)

    def _pack_data(self):
        """Per-window data block [y, weight, aux...] per residual: one fancy-indexed gather from the parameter pool
        (the loop over every scalar residual this replaces took 1 ms per solve on the 612-residual GNSS script)."""
        if getattr(self, "_pack", None) is None:
            self._pack_plan()
        const, dst, src, slow = self._pack
        data = const.copy()
        if dst.size:
            data[dst] = self._pool.buf[src]
        for o, par, i in slow:
            data[o] = par.value[i]
        return data

    def _initial_z(self):
        d = self._desc
        V = np.zeros((d.K, d.nv))
        for v in self._groups.get(self._xgroup, []):
            if v.init is not None:
                for i in range(self.n):
                    V[v.index, d.cmap[i]] = v.init[i]
        g = np.zeros(d.ng)
        for v, off in self._globals:
            if v.init is not None:
                g[off:off + v.n] = v.init
        return np.concatenate([V.reshape(-1), g])

    def _solve_device(self, z0):
        d = self._desc
        U = None
        if self.m > 0:
            U = np.array([p.value for p in self._dyn[2]]).reshape(d.K, self.m)
        xprior = self._prior[3].value if self._prior is not None else None
        return self._run_kernel(z0, U, self._pack_data(), xprior, None)

    def _unpack(self, z):
        d = self._desc
        func, X, U, W, params = self._dyn
        V = z[:d.K * d.nv].reshape(d.K, d.nv)
        Xf = V[:, d.cmap]
        out = {}
        for v in self._groups.get(self._xgroup, []):
            out[v.name] = Xf[v.index].copy()
        for v, off in self._globals:
            out[v.name] = z[d.K * d.nv + off:d.K * d.nv + off + v.n].copy()
        # process-noise values from their defining equality (nlp.py:235)
        Dx = (2.0 / self.T) * (self.CPM.D @ Xf)
        for k, wv in enumerate(W):
            if self.m != 0:
                f = func(Xf[k], U[k].value, params)
            else:
                f = func(Xf[k], params)
            out[wv.name] = Dx[k] - np.asarray(f, dtype=float).reshape(-1)
        return out


# =============================================================================================
# helpers of the lowering
# =============================================================================================
def _call_dyn_raw(func, x, u, m, params):
    """The model callable with the reference's calling convention -- func(x, u, params), func(x, params) without
    controls (nlp/nlp.py:211-218) -- tolerating callables that take no params argument.  Returns whatever the callable
    returns (numbers, arrays, or traced expressions under the code generator)."""
    import inspect
    try:
        npar = len([q for q in inspect.signature(func).parameters.values()
                    if q.kind in (q.POSITIONAL_ONLY, q.POSITIONAL_OR_KEYWORD)])
    except (TypeError, ValueError):
        npar = 3
    if m == 0:
        return func(x, params) if npar >= 2 else func(x)
    return func(x, u, params) if npar >= 3 else func(x, u)


def _call_dyn(func, x, u, m, params):
    return np.asarray(_call_dyn_raw(func, x, u, m, params), dtype=float).reshape(-1)


def _lower_dynamics(d, func, n, m, params):
    """Select the kernel model of a dynamics callable (registry, else numeric affinity probe)."""
    reg = _dyn.KERNEL_MODELS.get(func)
    if reg is not None and reg[0] != D.DYN_AFFINE:
        dyn_id, rn, rm = reg
        if (rn, rm) != (n, m):
            raise ValueError("%s needs n=%d, m=%d (problem has n=%d, m=%d)" % (func.__name__, rn, rm, n, m))
        d.dyn_id = dyn_id
        if dyn_id in (D.DYN_VEHICLE, D.DYN_VEHICLE_GNSS):
            C = params["car_params"]
            d.dyn_par[:6] = [C["C_AF"], C["C_AR"], C["M"], C["D_F"], C["D_R"], C["I_Z"]]
        elif dyn_id == D.DYN_QUADCOPTER:
            I = np.asarray(params["I"], dtype=float)
            d.dyn_par[:4] = [params["m"], I[0, 0], I[1, 1], I[2, 2]]
        return
    # affine probe: f(x,u) = c0 + A x + B u, verified at random points
    rng = np.random.default_rng(12345)
    zx, zu = np.zeros(n), np.zeros(max(m, 1))
    uarg = lambda u: (u if m > 1 else (u if reg is not None and reg[1] == 1 and m == 1 else u))
    try:
        c0 = _call_dyn(func, zx, zu[:m] if m > 0 else None, m, params)
        if c0.shape[0] != n:
            raise ValueError
        A = np.zeros((n, n))
        B = np.zeros((n, max(m, 1)))
        for j in range(n):
            e = np.zeros(n); e[j] = 1.0
            A[:, j] = _call_dyn(func, e, zu[:m], m, params) - c0
        for j in range(m):
            e = np.zeros(m); e[j] = 1.0
            B[:, j] = _call_dyn(func, zx, e, m, params) - c0
        for _ in range(3):
            x, u = rng.normal(size=n), rng.normal(size=max(m, 1))
            f = _call_dyn(func, x, u[:m], m, params)
            if np.max(np.abs(f - (c0 + A @ x + B[:, :m] @ u[:m]))) > 1e-9 * (1 + np.max(np.abs(f))):
                raise ValueError
    except Exception:
        # neither registered nor affine: generate the CUDA model (value + symbolic Jacobian) from the callable itself;
        # csrc/user_model.cu compiles it into the window kernels with NVRTC when the problem is created on the device
        from . import _codegen
        try:
            src, info = _codegen.generate(func, n, m, params, _call_dyn_raw)
        except Exception as e:
            raise NotImplementedError("dynamics %r is neither a registered model (nlp/dynamics.py), nor affine in (x, u), nor "
                                      "traceable for code generation (%s); there is no CPU fallback" % (func, e))
        d.dyn_id = D.DYN_USER
        d.user_src = src
        d.user_func = (func, params, m)
        return
    d.dyn_id = D.DYN_AFFINE
    d.A, d.Bm, d.c0 = A, B, c0


def _diag_quadratic(func, n, m, params):
    """Stage cost l(x,u) = sum_i s_i v_i^2 (both reference stage costs, cost_functions.py:5-12)."""
    nv = n + m
    call = lambda v: float(np.asarray(func(v[:n], v[n:], params), dtype=float).reshape(-1)[0])
    if abs(call(np.zeros(nv))) > 0:
        raise NotImplementedError("the stage cost must vanish at the origin")
    s = np.array([call(np.eye(nv)[i]) for i in range(nv)])
    rng = np.random.default_rng(7)
    for _ in range(3):
        v = rng.normal(size=nv)
        if abs(call(v) - float(np.sum(s * v * v))) > 1e-10 * (1 + abs(call(v))):
            raise NotImplementedError("the stage cost must be a diagonal quadratic form of (x, u)")
    return s
