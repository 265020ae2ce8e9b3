"""GNSS least-squares helpers under the reference's names (utils/leastsquares.py:6-169), with
the Gauss-Newton iterations running in the batched CUDA kernels (`nlpf_ls_solve`,
`nlpf_ls_vel`).  `runLeastSquares` solves all epochs of a log in ONE launch instead of the
reference's per-epoch Python loop.

Reference quirk kept: the default argument `x=np.zeros(3)` of
`iterativeLeastSquares` is a single array shared by all calls, so consecutive calls warm-start
from the previous solution and the returned position aliases that array.
"""
import numpy as np

from . import gnss    # noqa: F401  (the reference imports it)
from . import utils


def buildGeometryMatrix(sat_pos, x):
    """Rows -(x_sat - x)/|x_sat - x|   (leastsquares.py:6-16)."""
    los = np.asarray(sat_pos, dtype=float) - np.asarray(x, dtype=float)[None, :]
    return -los / np.linalg.norm(los, axis=1)[:, None]


def _device_ls(sat_pos, pr, x, b, tcol=None, alpha=0.0, maxiter=100):
    from .. import engine
    sat = np.asarray(sat_pos, dtype=float).reshape(1, -1, 3)
    ns = sat.shape[1]
    x0 = np.concatenate([np.asarray(x, dtype=float).reshape(3), [float(b)]] + ([[float(alpha)]] if tcol is not None else []))
    sol, it = engine.ls_solve(sat, np.asarray(pr, dtype=float).reshape(1, ns), np.array([ns], dtype=np.int32), x0=x0[None, :],
                              tcol=None if tcol is None else np.asarray(tcol, dtype=float).reshape(1, ns), maxiter=maxiter)
    return sol.cpu().numpy()[0], int(it.cpu().numpy()[0])


def iterativeLeastSquares(sat_pos, pr, x=np.zeros(3), b=0, maxiter=100):
    """Position and clock bias from one epoch of pseudoranges (leastsquares.py:19-42)."""
    sol, it = _device_ls(sat_pos, pr, x, b, maxiter=maxiter)
    x[:] = sol[:3]                      # in place: the (possibly default, shared) array is updated
    return x, sol[3]


def iterativeLeastSquaresVel(sat_pos, sat_vel, pr_rate, x):
    """Velocity and clock drift from pseudorange rates, one linear solve (leastsquares.py:45-63)."""
    from .. import engine
    sat = np.asarray(sat_pos, dtype=float).reshape(1, -1, 3)
    ns = sat.shape[1]
    out = engine.ls_vel(sat, np.asarray(sat_vel, dtype=float).reshape(1, ns, 3), np.asarray(pr_rate, dtype=float).reshape(1, ns),
                        np.array([ns], dtype=np.int32), np.asarray(x, dtype=float).reshape(1, 3)).cpu().numpy()[0]
    return out[:3], out[3]


def iterativeLeastSquares_multiTimeStep(t, sat_pos, pr, x=np.zeros(3), b0=0, alpha=0, maxiter=100):
    """Stationary receiver, clock bias b0 + alpha t, all epochs at once (leastsquares.py:66-94)."""
    sol, it = _device_ls(sat_pos, pr, x, b0, tcol=t, alpha=alpha, maxiter=maxiter)
    if it < maxiter:
        print('Multi time step least squares converged.')
    x[:] = sol[:3]
    return x, sol[3], sol[4]


_KEYS = ("bias", "bias_rate", "x_ECEF", "y_ECEF", "z_ECEF", "xd_ECEF", "yd_ECEF", "zd_ECEF", "x_ENU", "y_ENU", "z_ENU",
         "xd_ENU", "yd_ENU", "zd_ENU", "lat", "lon", "h")


def _pad(rows, width=None):
    T = len(rows)
    smax = max(np.asarray(r).shape[0] for r in rows)
    out = np.zeros((T, smax) if width is None else (T, smax, width))
    n = np.zeros(T, dtype=np.int32)
    for i, r in enumerate(rows):
        r = np.asarray(r, dtype=float)
        n[i] = r.shape[0]
        out[i, :n[i]] = r
    return out, n


def runLeastSquares(t, sat_pos, pr, sat_vel=None, pr_rate=None, p_ref_ECEF=None):
    """Per-epoch position (and velocity) solutions of a whole log (leastsquares.py:97-141).  All
    epochs are solved in one kernel launch, each started from the shared warm-start position the
    reference would have carried into the call."""
    from .. import engine
    T = t.shape[0]
    sol = {"t": t, "p_ref_ECEF": p_ref_ECEF}
    for k in _KEYS:
        sol[k] = np.zeros(T)
    sat, nsat = _pad(sat_pos, 3)
    prp, _ = _pad(pr)
    warm = iterativeLeastSquares.__defaults__[0]
    x0 = np.concatenate([np.tile(warm, (T, 1)), np.zeros((T, 1))], 1)
    x, it = engine.ls_solve(sat, prp, nsat, x0=x0)
    x = x.cpu().numpy()
    warm[:] = x[-1, :3]
    vel = None
    if sat_vel is not None:
        velp, _ = _pad(sat_vel, 3)
        prrp, _ = _pad(pr_rate)
        vel = engine.ls_vel(sat, velp, prrp, nsat, np.ascontiguousarray(x[:, :3])).cpu().numpy()
    for k in range(T):
        p = x[k, :3]
        lla = utils.ecef2lla(p)
        sol["x_ECEF"][k], sol["y_ECEF"][k], sol["z_ECEF"][k] = p
        sol["lat"][k], sol["lon"][k], sol["h"][k] = lla[0], lla[1], lla[2]
        sol["bias"][k] = x[k, 3]
        if p_ref_ECEF is not None:
            e = utils.ecef2enu(p, p_ref_ECEF)
            sol["x_ENU"][k], sol["y_ENU"][k], sol["z_ENU"][k] = e
        if vel is not None:
            v = vel[k, :3]
            sol["xd_ECEF"][k], sol["yd_ECEF"][k], sol["zd_ECEF"][k] = v
            sol["bias_rate"][k] = vel[k, 3]
            if p_ref_ECEF is not None:
                ve = utils.ecef2enu(v, p_ref_ECEF, rotation_only=True)
                sol["xd_ENU"][k], sol["yd_ENU"][k], sol["zd_ENU"][k] = ve
    return sol


def runBatchLeastSquares(t, sat_pos, pr, p_ref_ECEF=None):
    """Stationary receiver with linearly drifting clock over all epochs (leastsquares.py:144-169)."""
    sats = np.vstack([np.asarray(s, dtype=float).reshape(-1, 3) for s in sat_pos])
    prs = np.hstack([np.asarray(p, dtype=float) for p in pr])
    ts = np.hstack([[t[i]] * np.asarray(pr[i]).shape[0] for i in range(t.shape[0])])
    p_ECEF, b0, alpha = iterativeLeastSquares_multiTimeStep(ts, sats, prs)
    p_LLA = utils.ecef2lla(p_ECEF)
    sol = {"t": t, "p_ref_ECEF": p_ref_ECEF, "b0": b0, "alpha": alpha,
           "x_ECEF": p_ECEF[0], "y_ECEF": p_ECEF[1], "z_ECEF": p_ECEF[2],
           "lat": p_LLA[0], "lon": p_LLA[1], "h": p_LLA[2]}
    if p_ref_ECEF is not None:
        p_ENU = utils.ecef2enu(p_ECEF, p_ref_ECEF)
        sol["x_ENU"], sol["y_ENU"], sol["z_ENU"] = p_ENU[0], p_ENU[1], p_ENU[2]
    return sol
