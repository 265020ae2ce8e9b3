"""CUDA code generation for model callables outside the kernel library.

The reference hands any Python callable to `casadi.Function` and lets CasADi differentiate it
(nlp/nlp.py:211-218).  Here the callable is run once on symbolic arguments (`Sym`: operator overloading, NumPy
ufuncs, and the casadi-like names of `_sym.py`), which records an expression graph for f(x, u); the graph is
differentiated symbolically with respect to every state and control component, common sub-expressions are shared
(hash-consing), and one CUDA device function is emitted -- model value, sparse Jacobian entries through `FAcc::add`,
control columns through `ucol` -- which `csrc/user_model.cu` compiles into the window kernels with NVRTC.

`params` of the callable are ordinary Python numbers / arrays: they are baked into the generated source as literals
(17 significant digits), so a problem class with other parameter values is another kernel.

Nothing here evaluates the model for the solver on the CPU: `evaluate()` exists to VERIFY the trace (value and
Jacobian against the callable itself and central differences) before any code is generated.
"""
import math

import numpy as np

_UNARY = {"sin": math.sin, "cos": math.cos, "tan": math.tan, "sqrt": math.sqrt, "exp": math.exp, "log": math.log,
          "neg": lambda a: -a, "abs": abs, "tanh": math.tanh, "asin": math.asin, "acos": math.acos, "atan": math.atan}
_BINARY = {"add": lambda a, b: a + b, "sub": lambda a, b: a - b, "mul": lambda a, b: a * b, "div": lambda a, b: a / b,
           "pow": lambda a, b: a ** b, "atan2": math.atan2}
_UFUNC = {np.sin: "sin", np.cos: "cos", np.tan: "tan", np.sqrt: "sqrt", np.exp: "exp", np.log: "log", np.negative: "neg",
          np.absolute: "abs", np.tanh: "tanh", np.arcsin: "asin", np.arccos: "acos", np.arctan: "atan",
          np.add: "add", np.subtract: "sub", np.multiply: "mul", np.true_divide: "div", np.power: "pow",
          np.arctan2: "atan2"}


class Graph(object):
    """Hash-consed expression nodes: (op, args) -> Sym, so equal sub-expressions are one node."""

    def __init__(self):
        self.table = {}
        self.nodes = []

    def node(self, op, args=(), val=None):
        key = (op, tuple(a.id for a in args), val)
        s = self.table.get(key)
        if s is None:
            s = Sym(self, op, tuple(args), val, len(self.nodes))
            self.table[key] = s
            self.nodes.append(s)
        return s

    def const(self, v):
        return self.node("const", (), float(v))

    def lift(self, v):
        if isinstance(v, Sym):
            return v
        if isinstance(v, np.ndarray) and v.ndim == 0:
            v = v.item()
        if isinstance(v, (int, float, np.floating, np.integer)):
            return self.const(v)
        raise TypeError("cannot use %r in a traced model expression" % (type(v),))

    # ---- simplifying constructors --------------------------------------------------------------
    def unary(self, op, a):
        a = self.lift(a)
        if a.op == "const":
            return self.const(_UNARY[op](a.val))
        if op == "neg" and a.op == "neg":
            return a.args[0]
        return self.node(op, (a,))

    def binary(self, op, a, b):
        a, b = self.lift(a), self.lift(b)
        ca, cb = a.op == "const", b.op == "const"
        if ca and cb:
            return self.const(_BINARY[op](a.val, b.val))
        if op == "add":
            if ca and a.val == 0.0:
                return b
            if cb and b.val == 0.0:
                return a
        elif op == "sub":
            if cb and b.val == 0.0:
                return a
            if ca and a.val == 0.0:
                return self.unary("neg", b)
        elif op == "mul":
            if (ca and a.val == 0.0) or (cb and b.val == 0.0):
                return self.const(0.0)
            if ca and a.val == 1.0:
                return b
            if cb and b.val == 1.0:
                return a
            if ca and a.val == -1.0:
                return self.unary("neg", b)
            if cb and b.val == -1.0:
                return self.unary("neg", a)
        elif op == "div":
            if ca and a.val == 0.0:
                return self.const(0.0)
            if cb and b.val == 1.0:
                return a
        elif op == "pow":
            if cb and b.val == 1.0:
                return a
            if cb and b.val == 0.0:
                return self.const(1.0)
            if cb and b.val == 2.0:
                return self.binary("mul", a, a)
        return self.node(op, (a, b))


class Sym(object):
    """One node of the expression graph; behaves like a scalar in arithmetic and NumPy ufuncs."""
    __array_priority__ = 1000

    def __init__(self, graph, op, args, val, idx):
        self.g, self.op, self.args, self.val, self.id = graph, op, args, val, idx

    def __add__(self, o): return self.g.binary("add", self, o)
    def __radd__(self, o): return self.g.binary("add", o, self)
    def __sub__(self, o): return self.g.binary("sub", self, o)
    def __rsub__(self, o): return self.g.binary("sub", o, self)
    def __mul__(self, o): return self.g.binary("mul", self, o)
    def __rmul__(self, o): return self.g.binary("mul", o, self)
    def __truediv__(self, o): return self.g.binary("div", self, o)
    def __rtruediv__(self, o): return self.g.binary("div", o, self)
    def __pow__(self, o): return self.g.binary("pow", self, o)
    def __rpow__(self, o): return self.g.binary("pow", o, self)
    def __neg__(self): return self.g.unary("neg", self)
    def __pos__(self): return self
    def __abs__(self): return self.g.unary("abs", self)

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        name = _UFUNC.get(ufunc)
        if method != "__call__" or name is None or kwargs:
            return NotImplemented
        if name in _UNARY:
            return self.g.unary(name, inputs[0])
        return self.g.binary(name, inputs[0], inputs[1])

    def __bool__(self):
        raise TypeError("a traced model cannot branch on its arguments (data-dependent `if`)")

    def __repr__(self):
        return "Sym(%s#%d)" % (self.op, self.id)


class SymVec(object):
    """x or u as seen by the callable: indexing gives Sym components, arithmetic acts component-wise."""

    def __init__(self, comps):
        self.c = list(comps)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return SymVec(self.c[i])
        return self.c[i]

    def __len__(self):
        return len(self.c)

    def __iter__(self):
        return iter(self.c)

    def _bin(self, o, f):
        if isinstance(o, SymVec):
            o = o.c
        if isinstance(o, (list, tuple, np.ndarray)):
            return SymVec([f(a, b) for a, b in zip(self.c, list(np.asarray(o, dtype=object).reshape(-1)))])
        return SymVec([f(a, o) for a in self.c])

    def __add__(self, o): return self._bin(o, lambda a, b: a + b)
    def __radd__(self, o): return self._bin(o, lambda a, b: b + a)
    def __sub__(self, o): return self._bin(o, lambda a, b: a - b)
    def __rsub__(self, o): return self._bin(o, lambda a, b: b - a)
    def __mul__(self, o): return self._bin(o, lambda a, b: a * b)
    def __rmul__(self, o): return self._bin(o, lambda a, b: b * a)
    def __truediv__(self, o): return self._bin(o, lambda a, b: a / b)
    def __neg__(self): return SymVec([-a for a in self.c])


def flatten(out):
    """Return value of a model callable -> flat list of Sym / numbers (vertcat of scalars and vectors)."""
    if isinstance(out, SymVec):
        return list(out.c)
    if isinstance(out, Sym):
        return [out]
    if isinstance(out, np.ndarray):
        res = []
        for v in out.reshape(-1):
            res.extend(flatten(v))
        return res
    if isinstance(out, (list, tuple)):
        res = []
        for v in out:
            res.extend(flatten(v))
        return res
    return [out]


# -------------------------------------------------------------------------------------------------
def trace(func, n, m, params, call):
    """Run the callable on symbolic x (n) and u (m).  `call(func, x, u, m, params)` is the caller's calling convention
    (with / without u and params).  Returns (graph, xs, us, f) with f a list of n Sym."""
    g = Graph()
    xs = [g.node("x", (), i) for i in range(n)]
    us = [g.node("u", (), j) for j in range(m)]
    out = flatten(call(func, SymVec(xs), SymVec(us), m, params))
    if len(out) != n:
        raise ValueError("the model returned %d components for %d states" % (len(out), n))
    return g, xs, us, [g.lift(v) for v in out]


def diff(g, e, var, memo):
    """d e / d var (var: an 'x' or 'u' leaf), symbolic, memoised per (node, var)."""
    key = (e.id, var.id)
    if key in memo:
        return memo[key]
    op, a = e.op, e.args
    if e is var:
        r = g.const(1.0)
    elif op in ("const", "x", "u"):
        r = g.const(0.0)
    elif op == "add":
        r = diff(g, a[0], var, memo) + diff(g, a[1], var, memo)
    elif op == "sub":
        r = diff(g, a[0], var, memo) - diff(g, a[1], var, memo)
    elif op == "mul":
        r = diff(g, a[0], var, memo) * a[1] + a[0] * diff(g, a[1], var, memo)
    elif op == "div":
        r = (diff(g, a[0], var, memo) - e * diff(g, a[1], var, memo)) / a[1]
    elif op == "neg":
        r = -diff(g, a[0], var, memo)
    elif op == "sin":
        r = g.unary("cos", a[0]) * diff(g, a[0], var, memo)
    elif op == "cos":
        r = -(g.unary("sin", a[0]) * diff(g, a[0], var, memo))
    elif op == "tan":
        r = (1.0 + e * e) * diff(g, a[0], var, memo)
    elif op == "tanh":
        r = (1.0 - e * e) * diff(g, a[0], var, memo)
    elif op == "sqrt":
        r = diff(g, a[0], var, memo) / (2.0 * e)
    elif op == "exp":
        r = e * diff(g, a[0], var, memo)
    elif op == "log":
        r = diff(g, a[0], var, memo) / a[0]
    elif op == "asin":
        r = diff(g, a[0], var, memo) / g.unary("sqrt", 1.0 - a[0] * a[0])
    elif op == "acos":
        r = -(diff(g, a[0], var, memo) / g.unary("sqrt", 1.0 - a[0] * a[0]))
    elif op == "atan":
        r = diff(g, a[0], var, memo) / (1.0 + a[0] * a[0])
    elif op == "atan2":       # atan2(y, x)
        y, x = a
        r = (x * diff(g, y, var, memo) - y * diff(g, x, var, memo)) / (x * x + y * y)
    elif op == "pow":
        base, ex = a
        if ex.op == "const":
            r = ex.val * g.binary("pow", base, ex.val - 1.0) * diff(g, base, var, memo)
        else:
            r = e * (diff(g, ex, var, memo) * g.unary("log", base) + ex * diff(g, base, var, memo) / base)
    elif op == "abs":
        r = (e / a[0]) * diff(g, a[0], var, memo)
    else:
        raise NotImplementedError("no derivative rule for %r" % op)
    memo[key] = r
    return r


def evaluate(nodes, xv, uv):
    """Numeric value of graph nodes at (xv, uv) -- verification of the trace only."""
    cache = {}

    def ev(e):
        if e.id in cache:
            return cache[e.id]
        if e.op == "const":
            v = e.val
        elif e.op == "x":
            v = float(xv[e.val])
        elif e.op == "u":
            v = float(uv[e.val])
        elif e.op in _UNARY:
            v = _UNARY[e.op](ev(e.args[0]))
        else:
            v = _BINARY[e.op](ev(e.args[0]), ev(e.args[1]))
        cache[e.id] = v
        return v
    return [ev(e) for e in nodes]


_CUDA_UNARY = {"sin": "sin(%s)", "cos": "cos(%s)", "tan": "tan(%s)", "sqrt": "sqrt(%s)", "exp": "exp(%s)", "log": "log(%s)",
               "neg": "(-%s)", "abs": "fabs(%s)", "tanh": "tanh(%s)", "asin": "asin(%s)", "acos": "acos(%s)", "atan": "atan(%s)"}
_CUDA_BINARY = {"add": "(%s + %s)", "sub": "(%s - %s)", "mul": "(%s * %s)", "div": "(%s / %s)", "pow": "pow(%s, %s)",
                "atan2": "atan2(%s, %s)"}


class _Emitter(object):
    def __init__(self):
        self.name = {}
        self.lines = []

    def ref(self, e):
        if e.id in self.name:
            return self.name[e.id]
        if e.op == "const":
            s = repr(float(e.val))
            if "e" not in s and "." not in s and "inf" not in s and "nan" not in s:
                s += ".0"
            if s in ("inf", "-inf", "nan"):
                raise ValueError("the traced model contains a non-finite constant")
            s = "(%s)" % s if s.startswith("-") else s
        elif e.op == "x":
            s = "x[%d]" % e.val
        elif e.op == "u":
            s = "u[%d]" % e.val
        else:
            if e.op in _CUDA_UNARY:
                rhs = _CUDA_UNARY[e.op] % self.ref(e.args[0])
            else:
                rhs = _CUDA_BINARY[e.op] % (self.ref(e.args[0]), self.ref(e.args[1]))
            s = "t%d" % e.id
            self.lines.append("const double %s = %s;" % (s, rhs))
        self.name[e.id] = s
        return s

    def take(self):
        out, self.lines = self.lines, []
        return out


def generate(func, n, m, params, call, rng=None, check_points=3):
    """Trace, verify, differentiate and emit.  Returns (cuda_source, info) or raises ValueError/NotImplementedError
    with the reason the callable cannot be lowered."""
    g, xs, us, f = trace(func, n, m, params, call)
    memo = {}
    dfx = [[diff(g, f[i], xs[j], memo) for j in range(n)] for i in range(n)]
    dfu = [[diff(g, f[i], us[j], memo) for j in range(m)] for i in range(n)]
    # ---- verification against the callable itself and central differences ----
    rng = rng or np.random.default_rng(2024)
    for _ in range(check_points):
        xv, uv = rng.normal(size=n), rng.normal(size=max(m, 1))
        ref = np.asarray(call(func, xv.copy(), uv[:m].copy(), m, params), dtype=float).reshape(-1)
        got = np.array(evaluate(f, xv, uv))
        if not np.all(np.isfinite(ref)):
            continue
        if np.max(np.abs(got - ref)) > 1e-10 * (1.0 + np.max(np.abs(ref))):
            raise ValueError("the traced expression does not reproduce the callable (data-dependent control flow?)")
        J = np.array([evaluate(row, xv, uv) for row in dfx]) if n else np.zeros((0, 0))
        for j in range(n):
            h = 1e-6 * (1.0 + abs(xv[j]))
            xp, xm = xv.copy(), xv.copy()
            xp[j] += h
            xm[j] -= h
            fd = (np.asarray(call(func, xp, uv[:m].copy(), m, params), dtype=float).reshape(-1)
                  - np.asarray(call(func, xm, uv[:m].copy(), m, params), dtype=float).reshape(-1)) / (2 * h)
            if np.max(np.abs(J[:, j] - fd)) > 1e-5 * (1.0 + np.max(np.abs(fd))):
                raise ValueError("symbolic Jacobian column %d disagrees with central differences" % j)
    # ---- emission ----
    em = _Emitter()
    body = []
    vals = [em.ref(e) for e in f]
    body += ["    " + l for l in em.take()]
    body += ["    f[%d] = %s;" % (i, v) for i, v in enumerate(vals)]
    jl = []
    nnz = 0
    for i in range(n):
        for j in range(n):
            e = dfx[i][j]
            if e.op == "const" and e.val == 0.0:
                continue
            r = em.ref(e)
            jl += ["        " + l for l in em.take()]
            jl.append("        F.add(%d, %d, %s);" % (i, j, r))
            nnz += 1
        for j in range(m):
            e = dfu[i][j]
            if e.op == "const" and e.val == 0.0:
                continue
            r = em.ref(e)
            jl += ["        " + l for l in em.take()]
            jl.append("        if (ucol[%d] >= 0) Fk[%d * nv + ucol[%d]] += %s;" % (j, i, j, r))
    src = ("// generated by nlp_filter_b200/nlp/_codegen.py from %s (n = %d, m = %d): value and symbolic Jacobian\n"
           "__device__ void nlpf_user_dyn(const double* par, const double* x, const double* u, double* f, bool jac,\n"
           "                              const FAcc& F, const int* ucol, double* Fk, int nv) {\n%s\n    if (jac) {\n%s\n    }\n}\n"
           % (getattr(func, "__name__", "a callable"), n, m, "\n".join(body), "\n".join(jl)))
    return src, dict(nodes=len(g.nodes), jac_nonzeros=nnz)
