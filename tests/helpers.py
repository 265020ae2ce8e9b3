"""Shared builders of test problems and windows (used by CPU and GPU tests)."""
import os

import numpy as np
from scipy.interpolate import interp1d

from nlp_filter_b200 import desc as D
from nlp_filter_b200 import shapes, synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def car_golden():
    return np.load(os.path.join(GOLDEN, "car.npz"))


def car_window_inputs(d, g, t0, T=2.0):
    """U (K,2) and data block of the vehicle window starting at t0 (autonomous-car.py:232-263)."""
    tn = d.cpm.tau2t(d.cpm.tau)
    ti = np.where((g["traj_t"] >= t0) & (g["traj_t"] <= t0 + T))[0]
    gi = np.where((g["gnss_t"] >= t0) & (g["gnss_t"] <= t0 + T))[0]
    U = interp1d(g["traj_t"][ti] - t0, g["traj_u"][:, ti], fill_value="extrapolate")(tn).T
    data = np.zeros(d.Ldata)
    n_sat = g["sat_enu"].shape[0]
    w = 0.1 * (1.0 / float(g["r_pr"]))
    for e in range(d.E):
        for j in range(n_sat):
            o = d.roff[e * n_sat + j]
            data[o] = g["pr"][gi[e], j]
            data[o + 1] = w
            data[o + 2:o + 5] = g["sat_enu"][j]
    return np.ascontiguousarray(U), data


def car_xhat0(g):
    return np.hstack((g["traj_x0"], [float(g["b0"]), float(g["alpha"]), 0.0]))


def car_extract(d, z, ts):
    V = np.asarray(z).reshape(d.K, d.nv)
    xo = d.cpm.basisMatrix(ts) @ V
    return np.stack([xo[:, 0], xo[:, 1], xo[:, 8], xo[:, 6], xo[:, 2], xo[:, 3], xo[:, 4]], 1)


from nlp_filter_b200.synth import two_receiver_batch  # noqa: E402,F401  (kept importable from here)
