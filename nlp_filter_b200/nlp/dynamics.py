"""Continuous-time models xdot = f(x, u, params) under the reference's names and signatures
(nlp/dynamics.py:4-174).  Evaluated numerically (simulation, extraction of the process-noise
values); on the GPU each registered model is a hand-written device function with an analytic
Jacobian (csrc/nls_kernels.cu: dyn_eval) selected through `KERNEL_MODELS` below.  A callable
that is not registered is accepted when it is affine in (x, u) (detected numerically by
nlp_filter_b200.compile); anything else raises -- there is no CPU fallback.
"""
import numpy as np

from ._sym import vertcat, sin, cos, tan
from ..desc import (DYN_AFFINE, DYN_VDP, DYN_BICYCLE, DYN_VEHICLE, DYN_VEHICLE_GNSS, DYN_QUADCOPTER)

KERNEL_MODELS = {}      # callable -> (kernel dynamics id, n, m)


def _kernel(dyn_id, n, m):
    def deco(fn):
        KERNEL_MODELS[fn] = (dyn_id, n, m)
        return fn
    return deco


# ---- integrators (affine) ------------------------------------------------------------------
@_kernel(DYN_AFFINE, 1, 1)
def single_integrator(x, u, params=None):
    """x = [x], u = [v_x]; xdot = u   (dynamics.py:4-8; returns the scalar u[0] like the reference)."""
    return u[0]


@_kernel(DYN_AFFINE, 2, 2)
def single_integrator_2D(x, u, params=None):
    """x = [x, y], u = [v_x, v_y]; xdot = u   (dynamics.py:10-17)."""
    return vertcat(u[0], u[1])


@_kernel(DYN_AFFINE, 3, 3)
def single_integrator_3D(x, u, params=None):
    """x = [x, y, z], u = velocity   (dynamics.py:19-27)."""
    return vertcat(u[0], u[1], u[2])


@_kernel(DYN_AFFINE, 4, 2)
def double_integrator(x, u, params=None):
    """x = [x, y, xdot, ydot], u = acceleration   (dynamics.py:29-38)."""
    return vertcat(x[2], x[3], u[0], u[1])


@_kernel(DYN_QUADCOPTER, 12, 4)
def quadcopter(x, u, params=None):
    """x = [pos(3), phi, th, psi, vel(3), p, q, r], u = [T, Mx, My, Mz]; params m, I (3x3 principal)
    (dynamics.py:40-59)."""
    m, I = params["m"], params["I"]
    ph, th, ps = x[3], x[4], x[5]
    p, q, r = x[9], x[10], x[11]
    spin = q * sin(ph) + r * cos(ph)
    acc = u[0] / m
    return vertcat(
        x[6], x[7], x[8],
        p + spin * tan(th),
        q * cos(ph) - r * sin(ph),
        spin * (1.0 / cos(th)),
        acc * (sin(ph) * sin(ps) + cos(ph) * sin(th) * cos(ps)),
        acc * (cos(ph) * sin(th) * sin(ps) - sin(ph) * cos(ps)),
        acc * (cos(ph) * cos(th)) - 9.81,
        (1.0 / I[0, 0]) * (u[1] - (I[2, 2] - I[1, 1]) * q * r),
        (1.0 / I[1, 1]) * (u[2] - (I[0, 0] - I[2, 2]) * r * p),
        (1.0 / I[2, 2]) * (u[3] - (I[1, 1] - I[0, 0]) * p * q))


@_kernel(DYN_VDP, 2, 1)
def van_der_pol(x, u, params=None):
    """Van der Pol oscillator, x = [x0, x1]   (dynamics.py:61-66)."""
    return vertcat((1 - x[1] ** 2) * x[0] - x[1] + u, x[0])


# ---- GNSS receiver models (affine) ---------------------------------------------------------
@_kernel(DYN_AFFINE, 5, 3)
def gnss_pos_and_bias(x, u, params=None):
    """x = [x, y, z, b, bd]: position driven by u, clock bias b with constant drift bd
    (dynamics.py:68-79)."""
    return vertcat(u[0], u[1], u[2], x[4], 0.0)


def multi_receiver(x, params=None):
    """x = [pB(3), bB, vB(3), alphaB], no control   (dynamics.py:81-96)."""
    return vertcat(x[4], x[5], x[6], x[7], 0.0, 0.0, 0.0, 0.0)


KERNEL_MODELS[multi_receiver] = (DYN_AFFINE, 8, 0)


@_kernel(DYN_AFFINE, 10, 6)
def gnss_two_receiver(x, u, params=None):
    """x = [pA(3), bA, alphaA, pB(3), bB, alphaB], u = [vA(3), vB(3)]   (dynamics.py:98-115)."""
    return vertcat(u[0], u[1], u[2], x[4], 0.0, u[3], u[4], u[5], x[9], 0.0)


# ---- vehicles ------------------------------------------------------------------------------
@_kernel(DYN_BICYCLE, 6, 2)
def kinematic_bycicle_and_bias(x, u, params=None):
    """Kinematic bicycle + clock bias: x = [x, y, z, b, bd, th], u = [throttle, steer].  As in the
    reference the heading used by the kinematics is x[2] (dynamics.py:117-136)."""
    L = 0.28
    v = 8.72649116358 * u[0] - 0.856053299155
    delta = np.deg2rad(28) * u[1]
    return vertcat(v * np.cos(x[2]), v * np.sin(x[2]), 0.0, x[4], 0.0, (v / L) * np.tan(delta))


@_kernel(DYN_VEHICLE, 6, 2)
def vehicle_dynamics(x, u, params=None):
    """Single-track model with linear tyres, x = [px, py, psi, vx, vy, r], u = [F_xr, delta]
    (dynamics.py:148-164); params["car_params"] = utils/vehicle_sim.get_parameters()."""
    C = params["car_params"]
    eps = .001
    slip_r = (x[4] - C["D_R"] * x[5]) / (x[3] + eps)
    slip_f = (x[4] + C["D_F"] * x[5]) / (x[3] + eps) - u[1]
    F_yr = -C["C_AR"] * slip_r
    F_yf = -C["C_AF"] * slip_f
    return vertcat(
        x[3] * cos(x[2]) - x[4] * sin(x[2]),
        x[3] * sin(x[2]) + x[4] * cos(x[2]),
        x[5],
        (-F_yf * sin(u[1]) + u[0]) / C["M"] + x[5] * x[4],
        (F_yf * cos(u[1]) + F_yr) / C["M"] - x[5] * x[3],
        (C["D_F"] * F_yf * cos(u[1]) - C["D_R"] * F_yr) / C["I_Z"])


@_kernel(DYN_VEHICLE_GNSS, 9, 2)
def vehicle_dynamics_and_gnss(x, u, params=None):
    """[px, py, psi, vx, vy, psid, b, bd, pz]: vehicle + clock bias/drift + height
    (dynamics.py:166-174)."""
    return vertcat(vehicle_dynamics(x[:6], u, params), x[7], 0.0, 0.0)
