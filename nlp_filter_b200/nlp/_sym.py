"""Stand-in for the handful of CasADi names the reference's model files import
(`from casadi import vertcat, sin, cos, tan, sqrt, dot, norm_2, atan2`, `casadi.mtimes`;
nlp/dynamics.py:2, nlp/measurements.py:2, nlp/cost_functions.py:2).

The model callables are dual-use in the reference: the same function is handed to CasADi
(symbolic) and to the numeric simulator (utils/simulate.py:6-8,46).  Here they are evaluated
numerically only -- the GPU kernels carry hand-derived Jacobians of the registered models
(`nlp_filter_b200.compile`), no automatic differentiation happens at run time -- so these
functions are thin NumPy wrappers.  Decision variables and parameters of a problem are the
light-weight handles below; they support exactly the indexing / linear arithmetic the scripts
use (`X[i][2]`, `x[0] - x[1]` in nlp/constraints.py:4-5).
"""
import numpy as np


def _num(x):
    return x.value if isinstance(x, Parameter) else x


def _symbolic(*args):
    """True when an argument is (or contains) a traced expression of the code generator (_codegen.Sym)."""
    from . import _codegen
    for a in args:
        if isinstance(a, (_codegen.Sym, _codegen.SymVec)):
            return True
        if isinstance(a, (list, tuple)) and any(isinstance(v, (_codegen.Sym, _codegen.SymVec)) for v in a):
            return True
        if isinstance(a, np.ndarray) and a.dtype == object:
            return True
    return False


def vertcat(*args):
    if _symbolic(*args):
        from . import _codegen
        return _codegen.flatten(list(args))
    parts = [np.atleast_1d(np.asarray(_num(a))).reshape(-1) for a in args]
    return np.concatenate(parts) if parts else np.zeros(0)


def sin(x):
    return np.sin(_num(x))


def cos(x):
    return np.cos(_num(x))


def tan(x):
    return np.tan(_num(x))


def sqrt(x):
    return np.sqrt(_num(x))


def exp(x):
    return np.exp(_num(x))


def log(x):
    return np.log(_num(x))


def atan2(y, x):
    return np.arctan2(_num(y), _num(x))


def dot(a, b):
    if _symbolic(a, b):
        from . import _codegen
        fa, fb = _codegen.flatten(a), _codegen.flatten(b)
        acc = 0.0
        for u, v in zip(fa, fb):
            acc = acc + u * v
        return acc
    return np.dot(np.asarray(_num(a)).reshape(-1), np.asarray(_num(b)).reshape(-1))


def norm_2(a):
    if _symbolic(a):
        return sqrt(dot(a, a))
    return np.linalg.norm(np.asarray(_num(a)).reshape(-1))


def mtimes(a, b):
    return np.matmul(_num(a), _num(b))


# ---------------------------------------------------------------------------------------------
# handles
# ---------------------------------------------------------------------------------------------
class LinExpr(object):
    """sum_i coeff_i * (variable, component) + const -- enough for nlp/constraints.py."""

    def __init__(self, terms=None, const=0.0):
        self.terms = dict(terms or {})
        self.const = float(const)

    @staticmethod
    def of(x):
        if isinstance(x, LinExpr):
            return x
        if isinstance(x, Elem):
            return LinExpr({(x.var, x.i): 1.0})
        if isinstance(x, (int, float, np.floating, np.integer)):
            return LinExpr({}, float(x))
        raise NotImplementedError("only linear expressions of decision variables are supported in constraints "
                                  "(got %r)" % (type(x),))

    def _combine(self, other, sign):
        o = LinExpr.of(other)
        t = dict(self.terms)
        for k, v in o.terms.items():
            t[k] = t.get(k, 0.0) + sign * v
        return LinExpr(t, self.const + sign * o.const)

    def __add__(self, o):
        return self._combine(o, 1.0)

    __radd__ = __add__

    def __sub__(self, o):
        return self._combine(o, -1.0)

    def __rsub__(self, o):
        return LinExpr.of(o)._combine(self, -1.0)

    def __neg__(self):
        return LinExpr({k: -v for k, v in self.terms.items()}, -self.const)

    def __mul__(self, c):
        if not isinstance(c, (int, float, np.floating, np.integer)):
            raise NotImplementedError("products of decision variables are not supported in constraints")
        return LinExpr({k: v * c for k, v in self.terms.items()}, self.const * c)

    __rmul__ = __mul__

    def nonzero_terms(self):
        return {k: v for k, v in self.terms.items() if v != 0.0}


class Elem(object):
    """One scalar component of a decision variable (e.g. X[i][2])."""

    def __init__(self, var, i):
        self.var, self.i = var, int(i)

    def __add__(self, o):
        return LinExpr.of(self) + o

    __radd__ = __add__

    def __sub__(self, o):
        return LinExpr.of(self) - o

    def __rsub__(self, o):
        return LinExpr.of(o) - self

    def __neg__(self):
        return -LinExpr.of(self)

    def __mul__(self, c):
        return LinExpr.of(self) * c

    __rmul__ = __mul__


class Variable(object):
    """A decision vector created by addVariables (nlp/nlp.py:18-36): group name, index, size."""

    def __init__(self, group, index, n):
        self.group, self.index, self.n = group, int(index), int(n)
        self.name = "%s_%d" % (group, index)
        self.init = None          # initial guess (initialGuess / initializeEstimate)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [Elem(self, k) for k in range(*i.indices(self.n))]
        if i < 0:
            i += self.n
        if not 0 <= i < self.n:
            raise IndexError(i)
        return Elem(self, i)

    def __len__(self):
        return self.n

    def __repr__(self):
        return "Variable(%s, n=%d)" % (self.name, self.n)


class ParamPool(object):
    """One flat float64 array behind every parameter of a problem (the problem-build
    path).  A parameter is a slice of it, so (a) a whole measurement matrix is uploaded with ONE array assignment
    (`setMeasurement` on the parameters an `addResidualCost` call returned) and (b) the per-window data block of the
    kernels is gathered with one fancy-indexing operation from precomputed offsets (`NLP._pack_data`) instead of a
    Python loop over every scalar residual."""

    def __init__(self):
        self.buf = np.zeros(256)
        self.used = 0

    def alloc(self, n):
        off = self.used
        self.used += int(n)
        if self.used > self.buf.shape[0]:
            nb = np.zeros(max(2 * self.buf.shape[0], self.used))
            nb[:off] = self.buf[:off]
            self.buf = nb
        return off


class Parameter(object):
    """A numeric parameter created by addParameter (nlp/nlp.py:38-47); set through setParameter.  Its value lives in
    the problem's ParamPool (`value` is a view of the pool); `block` = (first offset, count, size, position) of the
    addParameter call that created it."""

    def __init__(self, n, pool=None, off=None, block=None):
        self.n = int(n)
        self.pool = pool if pool is not None else ParamPool()
        self.off = self.pool.alloc(self.n) if off is None else int(off)
        self.block = block
        self.is_set = False

    @property
    def value(self):
        return self.pool.buf[self.off:self.off + self.n]

    def set(self, val):
        v = np.asarray(val, dtype=float).reshape(-1)
        if v.size == 1 and self.n > 1:
            v = np.full(self.n, float(v[0]))
        if v.size != self.n:
            raise ValueError("parameter of size %d cannot take a value of size %d" % (self.n, v.size))
        self.pool.buf[self.off:self.off + self.n] = v
        self.is_set = True

    def __getitem__(self, i):
        return ParamElem(self, i)

    def __len__(self):
        return self.n


class ParamElem(object):
    """Y[0] of a one-element measurement parameter (autonomous-car.py:259 via nlp.py:310-312)."""

    def __init__(self, par, i):
        self.par, self.i = par, int(i)

    def set(self, val):
        v = np.asarray(val, dtype=float).reshape(-1)
        self.par.pool.buf[self.par.off + self.i] = v[0]
        self.par.is_set = True
