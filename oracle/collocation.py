"""ORACLE (test infrastructure, never shipped): Chebyshev pseudospectral constants.

CPU restatement of the reference's `nlp/collocation.py:7-121`.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may import this package.

Parity status: PINNED.  `tests/golden/collocation_N*.npz` were produced by executing the
reference class itself in this container (`tests/golden/make_golden.py`, python-2 integer
division restored on `nlp/collocation.py:78,80`) and this restatement matches them bit for bit
(tests/test_oracle_collocation.py).

What is reproduced, quirks included:
  * nodes      tau_k = cos(k*pi/N), reversed to ascending order        (collocation.py:34-40)
  * D          classic CGL matrix on the DEscending nodes, negated, used on the ascending grid
               without permutation                                      (collocation.py:42-64)
  * w          the reference's *non* Clenshaw-Curtis weights: the "last term" and the mirror
               assignment sit inside the j loop, python-2 `//`          (collocation.py:66-85)
  * phi_j      monomial coefficients from repeated polynomial products, Horner evaluation
               (np.poly1d semantics => same rounding as the reference)  (collocation.py:87-111)
"""
import numpy as np


def cgl_nodes(N):
    """Ascending CGL nodes (collocation.py:34-40)."""
    desc = np.zeros(N + 1)
    for k in range(N + 1):
        desc[k] = np.cos(k * np.pi / N)
    return desc[::-1]


def diff_matrix(N, tau):
    """Differentiation matrix as the reference stores it (collocation.py:42-64)."""
    x = tau[::-1]
    c = np.ones(N + 1)
    c[0] = 2
    c[N] = 2
    D = np.zeros((N + 1, N + 1))
    for k in range(N + 1):
        for j in range(N + 1):
            if k == j:
                if k == 0:
                    D[k, j] = (2 * N ** 2 + 1) / 6.0
                elif k == N:
                    D[k, j] = -(2 * N ** 2 + 1) / 6.0
                else:
                    D[k, j] = -x[k] / (2 * (1 - x[k] ** 2))
            else:
                D[k, j] = (c[k] / c[j]) * (np.power(-1, j + k) / (x[k] - x[j]))
    return -D


def quad_weights(N):
    """The reference's quadrature weights, bug-for-bug (collocation.py:66-85)."""
    w = np.zeros(N + 1)
    if N % 2 == 0:
        w[0] = 1.0 / (N ** 2 - 1)
        a = 0
    else:
        w[0] = 1.0 / N ** 2
        a = 1
    w[N] = w[0]
    half = (N - a) // 2
    for s in range(1, half + 1):
        acc = 2.0 / N
        touched = False
        for j in range(1, half):
            acc += (4.0 / N) * (1.0 / (1.0 - 4.0 * j ** 2)) * np.cos(2 * np.pi * j * s / N)
            acc += (2.0 / N) * (1.0 / (1 - (N - a) ** 2)) * np.cos((N - a) * s * np.pi / N)
            touched = True
        w[s] = acc
        if touched:            # the mirror assignment lives inside the j loop in the reference
            w[N - s] = acc
    return w


def lagrange_coeffs(N, tau):
    """Monomial coefficients (highest power first) of phi_j (collocation.py:87-100).

    np.poly1d.__mul__ is np.convolve on the coefficient arrays, so building the products with
    np.convolve in the same order rounds identically."""
    rows = []
    for j in range(N + 1):
        cj = np.array([1.0])
        for k in range(N + 1):
            if k == j:
                continue
            s = 1 / (tau[j] - tau[k])
            cj = np.convolve(cj, np.array([1.0, -tau[k]]) * s)
        rows.append(cj)
    return rows


def horner(coeffs, x):
    """np.polyval / poly1d.__call__ recurrence: y = y*x + c."""
    y = np.zeros_like(np.asarray(x, dtype=float))
    for cval in coeffs:
        y = y * x + cval
    return y


class Collocation(object):
    """Constants of one (N, T) problem class, t0 = 0 (the only use: nlp.py:130,197)."""

    def __init__(self, N, T):
        self.N = int(N)
        self.T = float(T)
        self.tau = cgl_nodes(self.N)
        self.D = diff_matrix(self.N, self.tau)
        self.w = quad_weights(self.N)
        self.coeffs = lagrange_coeffs(self.N, self.tau)

    def t_nodes(self):
        return 0.5 * (self.T * self.tau + self.T)        # collocation.py:28 with t0 = 0

    def t2tau(self, t):
        return (2.0 * np.asarray(t, dtype=float) - self.T) / self.T   # collocation.py:32

    def phi(self, t):
        """Rows of Lagrange-basis values at times t -> array (len(t), N+1)."""
        tau = np.atleast_1d(self.t2tau(t)).astype(float).reshape(-1)
        out = np.zeros((tau.shape[0], self.N + 1))
        for j in range(self.N + 1):
            out[:, j] = horner(self.coeffs[j], tau)
        return out

    def Dt(self):
        return (2.0 / self.T) * self.D

    def c(self):
        return (self.T / 2.0) * self.w
