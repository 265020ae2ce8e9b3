"""ORACLE (test infrastructure, never shipped): NumPy restatement of the reference's EKF and
GNSS least-squares path -- utils/ekf.py:20-61, utils/gnss.py:4-44,64-76,
utils/leastsquares.py:6-94 and the vehicle EKF model of autonomous-car.py:18-79.

Parity status: PINNED.  tests/golden/ekf_ls.npz holds outputs of the reference's own modules
(imported unmodified from /root/reference/utils by tests/golden/make_golden_ekf_ls.py) on the
stationary-receiver log, the simulated car data, the two-receiver logs and seeded synthetic
inputs; tests/test_oracle_ekf_ls.py checks this restatement against them (<= 1e-9 relative),
and the car filter additionally against the reference's stored data/autonomous-car/
filtering/ekf.pkl.

Quirks kept: the dynamics functions update the mean in place; the vehicle
Jacobian is evaluated at the propagated state; covariance update S_pred - K (H S_pred) with an
explicit inverse of the innovation matrix; LS uses the pseudo-inverse and stops on
|dx|_2 < 1e-7.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
import numpy as np


# ------------------------------------------------------------------------------------------
# measurement / dynamics models
# ------------------------------------------------------------------------------------------
def multi_pseudorange(x, sat_pos, cols=(0, 1, 2, 3)):
    """z_pred and H of gnss.multi_pseudorange (utils/gnss.py:4-44); `cols` are the state columns
    of x, y, z, b ((0,1,8,6) reproduces vehicle_sensors_model, autonomous-car.py:58-79)."""
    n = x.shape[0]
    p = np.array([x[cols[0]], x[cols[1]], x[cols[2]]])
    los = sat_pos - p[None, :]
    rng = np.sqrt(np.sum((p[None, :] - sat_pos) ** 2, axis=1))
    z = rng + x[cols[3]]
    H = np.zeros((sat_pos.shape[0], n))
    unit = -los / np.linalg.norm(los, axis=1)[:, None]
    H[:, cols[0]] = unit[:, 0]
    H[:, cols[1]] = unit[:, 1]
    H[:, cols[2]] = unit[:, 2]
    H[:, cols[3]] = 1.0
    return z, H


def dyn_gnss5(x, u, dt):
    """gnss.gnss_pos_and_bias (utils/gnss.py:64-76), in place."""
    x += dt * np.array([u[0], u[1], u[2], x[4], 0.0])
    G = np.eye(5)
    G[3, 4] = dt
    return x, G


def dyn_car9(x, u, dt, par):
    """discrete_vehicle_dynamics (autonomous-car.py:18-56) with the linear tyre model
    (utils/vehicle_sim.py:50-72); par = C_AF C_AR M D_F D_R I_Z.  In place; Jacobian at the
    propagated state."""
    C_AF, C_AR, M, D_F, D_R, I_Z = par[:6]
    a_r = (x[4] - D_R * x[5]) / x[3]
    a_f = (x[4] + D_F * x[5]) / x[3] - u[1]
    Fyr, Fyf = -C_AR * a_r, -C_AF * a_f
    xd = np.array([x[3] * np.cos(x[2]) - x[4] * np.sin(x[2]),
                   x[3] * np.sin(x[2]) + x[4] * np.cos(x[2]),
                   x[5],
                   (-Fyf * np.sin(u[1]) + u[0]) / M + x[5] * x[4],
                   (Fyf * np.cos(u[1]) + Fyr) / M - x[5] * x[3],
                   (D_F * Fyf * np.cos(u[1]) - D_R * Fyr) / I_Z,
                   x[7], 0.0, 0.0])
    x += dt * xd
    J = np.eye(9)
    fx = C_AF * (x[4] + D_F * x[5]) * (1. / x[3] ** 2)
    fy = -C_AF / x[3]
    fr = -C_AF * D_F / x[3]
    rx = C_AR * (x[4] - D_R * x[5]) * (1. / x[3] ** 2)
    ry = -C_AR / x[3]
    rr = C_AR * D_R / x[3]
    s2, c2, su, cu = np.sin(x[2]), np.cos(x[2]), np.sin(u[1]), np.cos(u[1])
    J[0, 2] += dt * (-x[3] * s2 - x[4] * c2)
    J[0, 3] += dt * c2
    J[0, 4] += -dt * s2
    J[1, 2] += dt * (x[3] * c2 - x[4] * s2)
    J[1, 3] += dt * s2
    J[1, 4] += dt * c2
    J[2, 5] += dt
    J[3, 3] += -(dt / M) * (su * fx)
    J[3, 4] += dt * (x[5] - (su * fy) / M)
    J[3, 5] += dt * (x[4] - (su * fr) / M)
    J[4, 3] += dt * ((cu * fx + rx) / M - x[5])
    J[4, 4] += (dt / M) * (cu * fy + ry)
    J[4, 5] += dt * ((cu * fr + rr) / M - x[3])
    J[5, 3] += (dt / I_Z) * (D_F * cu * fx - D_R * rx)
    J[5, 4] += (dt / I_Z) * (D_F * cu * fy - D_R * ry)
    J[5, 5] += (dt / I_Z) * (D_F * cu * fr - D_R * rr)
    J[6, 7] += dt
    return x, J


# ------------------------------------------------------------------------------------------
# EKF (utils/ekf.py:20-61)
# ------------------------------------------------------------------------------------------
def ekf_step(model, mu, S, u, z, sat_pos, Qdiag, rvar, dt, par=None, correct_only=False):
    """One EKF.update: predict, then correct when z is not None.  Returns (mu, S).  Qdiag may also be the full
    n x n process covariance, rvar the full measurement covariance (ekf.py:43,54 take matrices); correct_only
    is EKF.correct alone (ekf.py:46-61) applied to (mu, S) = (mu_pred, S_pred)."""
    mu = np.array(mu, dtype=float)
    cols = (0, 1, 2, 3) if model == 0 else (0, 1, 8, 6)
    if correct_only:
        S_pred = np.array(S, dtype=float)
    else:
        if model == 0:
            mu, G = dyn_gnss5(mu, u, dt)
        else:
            mu, G = dyn_car9(mu, u, dt, par)
        Qm = np.asarray(Qdiag, dtype=float)
        S_pred = np.matmul(G, np.matmul(S, G.T)) + (Qm if Qm.ndim == 2 else np.diag(Qm))
    if z is None or len(z) == 0:
        return mu, S_pred
    z_pred, H = multi_pseudorange(mu, sat_pos, cols)
    Rm = np.asarray(rvar, dtype=float)
    P = np.matmul(H, np.matmul(S_pred, H.T)) + (Rm if Rm.ndim == 2 else float(Rm) * np.eye(len(z)))
    K = np.matmul(S_pred, np.matmul(H.T, np.linalg.inv(P)))
    mu = mu + np.matmul(K, z - z_pred)
    S = S_pred - np.matmul(K, np.matmul(H, S_pred))
    return mu, S


def ekf_run(model, mu0, S0, u, z, sat, nsat, Qdiag, rvar, dt, par=None):
    """`steps` EKF steps of one filter.  u (steps,m), z (steps,smax), sat (steps,smax,3),
    nsat (steps,).  Returns (traj (steps,n), S)."""
    mu, S = np.array(mu0, float), np.array(S0, float)
    traj = np.zeros((u.shape[0], mu.shape[0]))
    for t in range(u.shape[0]):
        n = int(nsat[t])
        mu, S = ekf_step(model, mu, S, u[t], z[t, :n] if n > 0 else None, sat[t, :n] if n > 0 else None,
                         Qdiag, rvar, dt, par)
        traj[t] = mu
    return traj, S


# ------------------------------------------------------------------------------------------
# GNSS least squares (utils/leastsquares.py)
# ------------------------------------------------------------------------------------------
def geometry(sat_pos, x):
    """buildGeometryMatrix (leastsquares.py:6-16) with the bias column appended (:23)."""
    los = sat_pos - x[None, :]
    G = -los / np.linalg.norm(los, axis=1)[:, None]
    return np.hstack((G, np.ones((sat_pos.shape[0], 1))))


def iterative_ls(sat_pos, pr, x0=None, b0=0.0, tcol=None, alpha0=0.0, maxiter=100, tol=1e-7):
    """iterativeLeastSquares (leastsquares.py:19-42) or, with tcol, the multi-time-step variant
    (:66-94).  Returns (x(3), b[, alpha], iterations)."""
    x = np.zeros(3) if x0 is None else np.array(x0, float)
    b, alpha = float(b0), float(alpha0)
    it = 0
    for it in range(1, maxiter + 1):
        G = geometry(sat_pos, x)
        pred = np.linalg.norm(sat_pos - x[None, :], axis=1) + b
        if tcol is not None:
            G = np.hstack((G, tcol.reshape(-1, 1)))
            pred = pred + alpha * tcol
        dx = np.matmul(np.linalg.pinv(G), pr - pred)
        x = x + dx[0:3]
        b += dx[3]
        if tcol is not None:
            alpha += dx[4]
        if np.linalg.norm(dx) < tol:
            break
    if tcol is not None:
        return x, b, alpha, it
    return x, b, it


def ls_velocity(sat_pos, sat_vel, pr_rate, x):
    """iterativeLeastSquaresVel (leastsquares.py:45-63) -> (v(3), bdot)."""
    G = geometry(sat_pos, x)
    drho = pr_rate - np.sum(sat_vel * (-G[:, :3]), axis=1)
    sol = np.matmul(np.linalg.pinv(G), drho)
    return sol[:3], sol[3]
