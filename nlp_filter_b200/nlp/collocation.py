"""Chebyshev pseudospectral constants behind the reference's class name and attributes
(`nlp/collocation.py:7-121`): `tau`, `D`, `w`, `phi`, `tau2t`, `t2tau`,
`evaluateLagrangePolynomials`, `evaluateSolution`.

Host-side precompute: the arrays built here are uploaded once per
problem class and live in the descriptor (`Dt`, `cw`, `Phi`).  Values are reproduced to the
bit, quirks included, because they are parity-relevant:
  * the quadrature weights are the reference's, not Clenshaw-Curtis -- its correction term and
    its mirror assignment are applied once per inner index (collocation.py:80-83, python-2
    integer division), e.g. sum(w) = 1.97333 for N = 5;
  * the Lagrange basis is kept as expanded monomial polynomials (np.poly1d) and evaluated with
    Horner's rule, which loses accuracy for N >~ 25 exactly as the reference does.
"""
import math

import numpy as np


class ChebyshevPseudospectralMethod():
    def __init__(self, N, t0, tf):
        self.N = N
        self.t0 = t0
        self.tf = tf
        self.buildNodes()
        self.buildDiffMatrix()
        self.buildQuadratureWeights()
        self.buildLagrangePolynomials()

    # -- interval maps (collocation.py:26-32) ------------------------------------------------
    def tau2t(self, tau):
        span = self.tf - self.t0
        return 0.5 * (span * tau + (self.tf + self.t0))

    def t2tau(self, t):
        span = self.tf - self.t0
        return (2.0 * t - span) / span

    # -- constants ---------------------------------------------------------------------------
    def buildNodes(self):
        """CGL points, ascending (collocation.py:34-40)."""
        N = self.N
        self.tau = np.array([np.cos((N - i) * np.pi / N) for i in range(N + 1)])

    def buildDiffMatrix(self):
        """xdot(tau_k) = sum_j D_kj x_j (collocation.py:42-64): the textbook matrix of the
        descending grid, sign-flipped, which is exact on the ascending grid by symmetry."""
        N = self.N
        x = self.tau[::-1]
        c = np.ones(N + 1)
        c[0] = c[N] = 2
        k = np.arange(N + 1)
        sign = np.power(-1, k[None, :] + k[:, None])
        diff = x[:, None] - x[None, :]
        np.fill_diagonal(diff, 1.0)
        M = (c[:, None] / c[None, :]) * (sign / diff)
        inner = np.arange(1, N)
        M[inner, inner] = -x[inner] / (2 * (1 - x[inner] ** 2))
        M[0, 0] = (2 * N ** 2 + 1) / 6.0
        M[N, N] = -(2 * N ** 2 + 1) / 6.0
        self.D = -M

    def buildQuadratureWeights(self):
        """Weights of the stage/dynamics cost (collocation.py:66-85), reproduced as written
        there rather than as Clenshaw-Curtis."""
        N = self.N
        odd = N % 2
        w = np.zeros(N + 1)
        w[0] = w[N] = 1.0 / N ** 2 if odd else 1.0 / (N ** 2 - 1)
        P = N - odd
        half = P // 2
        tail = 1.0 / (1 - P ** 2) if half > 1 else 0.0
        for s in range(1, half + 1):
            v = 2.0 / N
            for j in range(1, half):
                v += (4.0 / N) * (1.0 / (1.0 - 4.0 * j ** 2)) * np.cos(2 * np.pi * j * s / N)
                v += (2.0 / N) * tail * np.cos(P * s * np.pi / N)
            w[s] = v
            if half > 1:
                w[N - s] = v
        self.w = w

    def buildLagrangePolynomials(self):
        """phi_j as np.poly1d in tau (collocation.py:87-100)."""
        tau = self.tau
        self.phi = []
        for j in range(self.N + 1):
            coeff = np.array([1.0])
            for k in range(self.N + 1):
                if k != j:
                    scale = 1 / (tau[j] - tau[k])
                    coeff = np.convolve(coeff, scale * np.array([1.0, -tau[k]]))
            self.phi.append(np.poly1d(coeff))

    # -- evaluation --------------------------------------------------------------------------
    def evaluateLagrangePolynomials(self, t):
        """phi_j at time t -> (N+1,) (collocation.py:102-111).  t may be a float or a
        one-element array (scripts pass (1,)-shaped rows of a (1,1) t_array)."""
        tau = self.t2tau(t)
        if np.ndim(tau) > 0:
            tau = np.asarray(tau).reshape(-1)[0]
        return np.array([p(tau) for p in self.phi])

    def evaluateSolution(self, t, X):
        """x(t) = sum_j X_j phi_j(t) (collocation.py:113-121)."""
        phi = self.evaluateLagrangePolynomials(t)
        x = np.zeros(X[0].shape)
        for j in range(self.N + 1):
            x += X[j] * phi[j]
        return x

    def basisMatrix(self, t_array):
        """Rows of phi at several times -> (len(t_array), N+1).  (Addition to the reference API.)"""
        ts = np.asarray(t_array, dtype=float).reshape(-1)
        return np.array([self.evaluateLagrangePolynomials(tt) for tt in ts]).reshape(ts.shape[0], self.N + 1)
