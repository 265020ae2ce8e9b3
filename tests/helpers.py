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


def two_receiver_batch(d, B, seed=5000):
    """Synthetic two-receiver windows (gnss-multi-receiver.py:42-216 shape): receivers 45.72 m apart,
    heading -44 deg, 10 satellite slots each with a few masked slots."""
    rng = np.random.default_rng(seed)
    K = d.K
    n_rh = 51
    n_pr_ep = 6
    n_sat = 10
    data = np.zeros((B, d.Ldata))
    pA = rng.normal(0, 20, (B, 3))
    hd = np.deg2rad(-44.0)
    dist = 45.72
    pB = pA + np.stack([dist * np.sin(hd) * np.ones(B), dist * np.cos(hd) * np.ones(B), np.zeros(B)], 1)
    bA, bB = rng.uniform(-500, 500, B), rng.uniform(-500, 500, B)
    aA, aB = rng.normal(0, 5, B), rng.normal(0, 5, B)
    el = np.arcsin(rng.uniform(np.sin(np.deg2rad(10)), 1.0, (B, n_sat)))
    az = rng.uniform(0, 2 * np.pi, (B, n_sat))
    rad = rng.uniform(2.0e7, 2.6e7, (B, n_sat))
    sat = np.stack([rad * np.cos(el) * np.sin(az), rad * np.cos(el) * np.cos(az), rad * np.sin(el)], 2)
    r = 0
    for e in range(n_rh):
        o = d.roff[r]; data[:, o] = dist; data[:, o + 1] = 0.1 * (1.0 / 0.01); r += 1
    for e in range(n_rh):
        o = d.roff[r]; data[:, o] = hd; data[:, o + 1] = 0.1 * (1.0 / 0.1); r += 1
    for e in range(n_pr_ep):
        t = float(e)
        for j in range(n_sat):
            for (p, b0, a, var) in ((pA, bA, aA, 10.0), (pB, bB, aB, 1.0)):
                o = d.roff[r]
                rho = np.sqrt(np.sum((p - sat[:, j, :]) ** 2, 1))
                data[:, o] = rho + b0 + a * t + rng.normal(0, np.sqrt(var), B)
                masked = rng.uniform(size=B) < 0.1
                data[:, o + 1] = np.where(masked, 0.0, 1.0 / var)
                data[:, o + 2:o + 5] = np.where(masked[:, None], 0.0, sat[:, j, :])
                data[:, o] = np.where(masked, 0.0, data[:, o])
                r += 1
    assert r == d.M
    xprior = np.concatenate([pA + rng.normal(0, 1, (B, 3)), bA[:, None], np.zeros((B, 1)),
                             pB + rng.normal(0, 1, (B, 3)), bB[:, None], np.zeros((B, 1))], 1)
    U = np.zeros((B, K, 6))
    z0 = np.zeros((B, K, d.nv))
    inv = np.array([0, 1, 2, 3, 4, 5, 6, 8, 9])      # folded columns -> state components
    z0[:] = xprior[:, None, inv]
    return dict(z0=z0.reshape(B, -1), U=U, data=data, xprior=xprior)
