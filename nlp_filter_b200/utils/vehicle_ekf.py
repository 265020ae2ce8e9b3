"""The discrete vehicle EKF model that autonomous-car.py defines inline (autonomous-car.py:18-79),
provided as importable callables and registered with kernel model EKF_CAR9.  (Addition: the
reference keeps these two functions inside the script.)"""
import numpy as np

from . import ekf as _ekf
from . import gnss as _gnss
from . import vehicle_sim


def _numeric_step(x, u, params, jac):
    dt = params["dt"]
    C = params["car_params"]
    xd = vehicle_sim.vehicle_dynamics(x[:6], u, {"tire_model_func": vehicle_sim.linear_tire_model})
    x += dt * np.hstack((xd, np.array([x[7], 0.0, 0.0])))
    if not jac:
        return x
    J = np.eye(9)
    ivx = 1.0 / x[3]
    fx, fy, fr = C["C_AF"] * (x[4] + C["D_F"] * x[5]) * ivx ** 2, -C["C_AF"] * ivx, -C["C_AF"] * C["D_F"] * ivx
    rx, ry, rr = C["C_AR"] * (x[4] - C["D_R"] * x[5]) * ivx ** 2, -C["C_AR"] * ivx, C["C_AR"] * C["D_R"] * ivx
    s2, c2, su, cu = np.sin(x[2]), np.cos(x[2]), np.sin(u[1]), np.cos(u[1])
    M, IZ, DF, DR = C["M"], C["I_Z"], C["D_F"], C["D_R"]
    J[0, 2:5] += dt * np.array([-x[3] * s2 - x[4] * c2, c2, -s2])
    J[1, 2:5] += dt * np.array([x[3] * c2 - x[4] * s2, s2, c2])
    J[2, 5] += dt
    J[3, 3:6] += dt * np.array([-su * fx / M, x[5] - su * fy / M, x[4] - su * fr / M])
    J[4, 3:6] += dt * np.array([(cu * fx + rx) / M - x[5], (cu * fy + ry) / M, (cu * fr + rr) / M - x[3]])
    J[5, 3:6] += (dt / IZ) * np.array([DF * cu * fx - DR * rx, DF * cu * fy - DR * ry, DF * cu * fr - DR * rr])
    J[6, 7] += dt
    return x, J


def discrete_vehicle_dynamics(x, u, params=None, jac=False):
    """Explicit-Euler step of the 9-state vehicle + clock model, in place; the Jacobian is taken
    at the propagated state, as the reference does (autonomous-car.py:18-56)."""
    return _numeric_step(x, u, params, jac)


def vehicle_sensors_model(x, params=None, jac=False):
    """Pseudoranges of x = [px, py, psi, vx, vy, psid, b, bd, pz] (autonomous-car.py:58-79)."""
    xm = np.array([x[0], x[1], x[8], x[6], x[7]])
    if not jac:
        return _gnss.multi_pseudorange(xm, params=params, jac=False)
    y, J5 = _gnss.multi_pseudorange(xm, params=params, jac=True)
    J = np.zeros((y.shape[0], 9))
    J[:, 0], J[:, 1], J[:, 8], J[:, 6], J[:, 7] = J5[:, 0], J5[:, 1], J5[:, 2], J5[:, 3], J5[:, 4]
    return y, J


_ekf.register_model(discrete_vehicle_dynamics, vehicle_sensors_model, _ekf.EKF_CAR9)
