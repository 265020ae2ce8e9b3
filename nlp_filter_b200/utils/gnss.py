"""EKF model functions under the reference's names (utils/gnss.py:4-76): value (and Jacobian
with jac=True) for one call.  `EKF` (utils/ekf.py) recognises these callables and runs the
fused CUDA filter step instead of calling them; they remain directly callable."""
import numpy as np


def pseudorange(x, params=None, jac=False):
    """|x[:3] - sat_pos| + x[3] for x = [x, y, z, b, bd]   (gnss.py:4-21)."""
    d = np.asarray(x[:3], dtype=float) - np.asarray(params["sat_pos"], dtype=float)
    rho = np.sqrt(d[0] ** 2 + d[1] ** 2 + d[2] ** 2)
    y = rho + x[3]
    if not jac:
        return y
    J = np.zeros(5)
    J[:3] = d / np.linalg.norm(d)
    J[3] = 1.0
    return y, J


def multi_pseudorange(x, params=None, jac=False):
    """Stack of pseudoranges to params["sat_pos"] (N x 3)   (gnss.py:24-44)."""
    sats = np.asarray(params["sat_pos"], dtype=float)
    N = sats.shape[0]
    y = np.zeros(N)
    J = np.zeros((N, 5))
    for i in range(N):
        y[i], J[i, :] = pseudorange(x, params={"sat_pos": sats[i, :]}, jac=True)
    return (y, J) if jac else y


def multi_pseudorange_and_bias(x, params=None, jac=False):
    """Pseudoranges plus a direct clock-bias pseudo-measurement (gnss.py:47-61; the reference's
    jac=False branch unpacks a single return value into two names and cannot run -- here both
    branches work)."""
    yp, Jp = multi_pseudorange(x, params={"sat_pos": params["sat_pos"]}, jac=True)
    y = np.append(yp, x[3])
    if not jac:
        return y
    J = np.vstack((Jp, np.zeros((1, 5))))
    return y, J


def gnss_pos_and_bias(x, u, params=None, jac=False):
    """x+ = x + dt [u, bd, 0], IN PLACE like the reference (gnss.py:64-76)."""
    dt = params["dt"]
    x += dt * np.array([u[0], u[1], u[2], x[4], 0.0])
    if not jac:
        return x
    J = np.eye(5)
    J[3, 4] = dt
    return x, J
