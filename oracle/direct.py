"""ORACLE (test infrastructure, never shipped): the estimation objective written the way the
reference SCRIPTS build it -- one Python term per `addDynamicsCost` node and per `addResidualCost`
call -- from raw physical inputs (node states, controls, measurement lists), with the collocation
constants of `oracle/collocation.py` (pinned to the reference class by tests/golden/collocation_*).

Purpose: an INDEPENDENT check of the problem lowering.  `oracle/nls.py` and the CUDA kernels both
consume a `ProblemDesc` that product code (`nlp_filter_b200/shapes.py`, `desc.py`) built: `Dt`,
`cw`, `Phi`, the residual table (`rtype`, `repoch`, `ridx`), `cmap`, `qinv`, `pinv`, `A`, `Bm`.
An error there would cancel between the two.  Nothing in this file imports `nlp_filter_b200`;
tests/test_oracle_direct_lowering.py evaluates J both ways at random points.

Reference lines followed:
  dynamics defect   W_k = (2/T) sum_j D_kj X_j - f(X_k, U_k)          nlp/nlp.py:222-235
  dynamics cost     J += (T/2) w_k W_k^T Q^-1 W_k                      nlp/nlp.py:242-245, cost_functions.py:20-22
  residual cost     X_t = sum_j phi_j(t) X_j;  r = y - h(X_t);  J += r^T R r      nlp/nlp.py:247-277
  arrival cost      J += (X_0 - x0)^T P^-1 (X_0 - x0)                  nlp/nlp.py:279-286
  models            nlp/dynamics.py:61-115, nlp/measurements.py:4-70
Slow on purpose (Python loops over nodes and measurements): small cases only.
"""
import numpy as np

from .collocation import Collocation


# ---- models, restated from the reference files (plain NumPy) -------------------------------
def f_van_der_pol(x, u):                       # dynamics.py:61-66
    return np.array([(1 - x[1] ** 2) * x[0] - x[1] + u[0], x[0]])


def f_gnss_pos_and_bias(x, u):                 # dynamics.py:68-79: x = [x, y, z, b, bd]
    return np.array([u[0], u[1], u[2], x[4], 0.0])


def f_gnss_two_receiver(x, u):                 # dynamics.py:97-115: two copies, controls [vA, vB]
    return np.array([u[0], u[1], u[2], x[4], 0.0, u[3], u[4], u[5], x[9], 0.0])


def h_pseudorange(x, sat, idx=(0, 1, 2, 3)):   # measurements.py:56-70
    return np.sqrt((x[idx[0]] - sat[0]) ** 2 + (x[idx[1]] - sat[1]) ** 2 + (x[idx[2]] - sat[2]) ** 2) + x[idx[3]]


def h_range_3d(x, idxA, idxB):                 # measurements.py:46-53
    return np.sqrt((x[idxA[0]] - x[idxB[0]]) ** 2 + (x[idxA[1]] - x[idxB[1]]) ** 2 + (x[idxA[2]] - x[idxB[2]]) ** 2
                   + .000001)


def h_heading_2d(x, idxA, idxB):               # measurements.py:30-36
    r_y = x[idxB[1]] - x[idxA[1]]
    r_x = x[idxB[0]] - x[idxA[0]] + .00001
    return np.arctan2(r_x, r_y)


# ---- the objective, term by term -------------------------------------------------------------
def objective(N, T, f, Qinv_diag, X, U, residuals, prior=None):
    """X (N+1, n) node states, U (N+1, m) node controls; residuals: iterable of (t, y, R, h) with h a callable
    on the interpolated state and R the scalar weight; prior = (x0, Pinv_diag) or None."""
    cpm = Collocation(N, T)
    X = np.asarray(X, float)
    J = 0.0
    for k in range(N + 1):
        d_k = np.zeros(X.shape[1])
        for j in range(N + 1):
            d_k += cpm.D[k, j] * X[j]
        W = (2.0 / T) * d_k - f(X[k], U[k])
        J += (T / 2.0) * cpm.w[k] * float(np.sum(Qinv_diag * W * W))
    for (t, y, R, h) in residuals:
        phi_t = cpm.phi(np.array([t]))[0]
        X_t = np.zeros(X.shape[1])
        for j in range(N + 1):
            X_t += X[j] * phi_t[j]
        r = y - h(X_t)
        J += float(r * R * r)
    if prior is not None:
        x0, Pinv = prior
        d0 = X[0] - x0
        J += float(np.sum(Pinv * d0 * d0))
    return J


def vdp_mhe(X, U, t_meas, y, r_var=(0.01, 0.02), q=1e-2, N=20, T=10.0):
    """estimation_example.py:10-45 with dynamics.van_der_pol: full-state measurements y (E, 2) at t_meas,
    R = diag(1/r_var) (weighted through addResidualCost's R), Q = q I."""
    res = []
    for e, t in enumerate(t_meas):
        for i in range(2):
            res.append((t, y[e, i], 1.0 / r_var[i], (lambda x, i=i: x[i])))
    return objective(N, T, f_van_der_pol, np.array([1.0 / q, 1.0 / q]), X, U, res)


def gnss(X, U, t_meas, pr, sat, w, N=10, T=50.0):
    """gnss_stationary.py:100-128: one addResidualCost(measurements.pseudorange) per satellite per epoch;
    pr (E, S), sat (E, S, 3) ENU, w (E, S) = 1 / r_pr (0 = slot not used)."""
    Qinv = 1.0 / np.array([0.0001, 0.0001, 0.0001, 0.1, 0.001])          # gnss_stationary.py:18
    res = []
    for e, t in enumerate(t_meas):
        for s in range(pr.shape[1]):
            if w[e, s] != 0.0:
                res.append((t, pr[e, s], w[e, s], (lambda x, p=sat[e, s]: h_pseudorange(x, p))))
    return objective(N, T, f_gnss_pos_and_bias, Qinv, X, U, res)


def two_receiver(X, U, x0, prA, satA, wA, prB, satB, wB, distance, heading, N=10, T=5.0):
    """gnss-multi-receiver.py:42-131.  X (11, 10) must satisfy the script's equality X_k[2] == X_k[7]
    (constraints.py:4-5, imposed at every node) -- the caller builds X that way."""
    Q = np.array([.01, .01, .01, 0.01, 0.01, .01, .01, .01, 0.01, 0.01])
    P = 0.01 * np.array([1, 1, 1, 0.1, 0.1, 1, 1, 1, 0.1, 0.1])
    r_range, r_heading = 0.01, 0.1
    dt_range = dt_heading = 0.1
    res = []
    for t in np.linspace(0, T, int(np.floor(T / dt_range)) + 1):
        res.append((t, distance, dt_range * (1.0 / r_range), (lambda x: h_range_3d(x, [0, 1, 2], [5, 6, 7]))))
    for t in np.linspace(0, T, int(np.floor(T / dt_heading)) + 1):
        res.append((t, heading, dt_heading * (1.0 / r_heading), (lambda x: h_heading_2d(x, [0, 1], [5, 6]))))
    t_gnss = np.linspace(0, T, int(np.floor(T / 1.0)) + 1)
    for i, t in enumerate(t_gnss):
        for j in range(prA.shape[1]):
            if wA[i, j] != 0.0:
                res.append((t, prA[i, j], wA[i, j], (lambda x, p=satA[i, j]: h_pseudorange(x, p, (0, 1, 2, 3)))))
            if wB[i, j] != 0.0:
                res.append((t, prB[i, j], wB[i, j], (lambda x, p=satB[i, j]: h_pseudorange(x, p, (5, 6, 7, 8)))))
    assert np.max(np.abs(np.asarray(X)[:, 2] - np.asarray(X)[:, 7])) == 0.0
    return objective(N, T, f_gnss_two_receiver, 1.0 / Q, X, U, res, prior=(x0, 1.0 / P))
