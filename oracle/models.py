"""ORACLE (test infrastructure): continuous-time models and measurement functions with
hand-derived Jacobians, restating `nlp/dynamics.py` and `nlp/measurements.py` of the
reference in NumPy.  Each function cites the lines it follows.  Jacobians are checked against
complex-step differentiation of the value functions in tests/test_oracle_models.py.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
import numpy as np

from nlp_filter_b200 import desc as D

# car constants, utils/vehicle_sim.py:10-18, packed as dyn_par[0:6]
CAR_PAR = dict(C_AF=1.1441e5, C_AR=1.3388e5, M=2009.0, D_F=1.53, D_R=1.23, I_Z=2000.0)
CAR_PAR_ORDER = ("C_AF", "C_AR", "M", "D_F", "D_R", "I_Z")

BICYCLE_V_GAIN = 8.72649116358      # dynamics.py:127
BICYCLE_V_OFF = 0.856053299155
BICYCLE_L = 0.28                    # dynamics.py:126


def _vehicle6(x, u, par):
    """vehicle_dynamics, dynamics.py:148-164.  x = [px,py,psi,vx,vy,r], u = [F_xr, delta]."""
    C_AF, C_AR, M, D_F, D_R, I_Z = par[:6]
    px, py, psi, vx, vy, r = x
    Fx, dl = u[0], u[1]
    den = vx + 0.001
    Fyr = -C_AR * (vy - D_R * r) / den
    Fyf = -C_AF * ((vy + D_F * r) / den - dl)
    sp, cp = np.sin(psi), np.cos(psi)
    sd, cd = np.sin(dl), np.cos(dl)
    f = np.array([vx * cp - vy * sp,
                  vx * sp + vy * cp,
                  r,
                  (-Fyf * sd + Fx) / M + r * vy,
                  (Fyf * cd + Fyr) / M - r * vx,
                  (D_F * Fyf * cd - D_R * Fyr) / I_Z], dtype=x.dtype)
    # partials of the tyre forces
    a_r = C_AR * (vy - D_R * r) / den ** 2
    b_r = -C_AR / den
    c_r = C_AR * D_R / den
    a_f = C_AF * (vy + D_F * r) / den ** 2
    b_f = -C_AF / den
    c_f = -C_AF * D_F / den
    F = np.zeros((6, 6), dtype=x.dtype)
    F[0, 2] = -vx * sp - vy * cp
    F[0, 3] = cp
    F[0, 4] = -sp
    F[1, 2] = vx * cp - vy * sp
    F[1, 3] = sp
    F[1, 4] = cp
    F[2, 5] = 1.0
    F[3, 3] = -sd * a_f / M
    F[3, 4] = -sd * b_f / M + r
    F[3, 5] = -sd * c_f / M + vy
    F[4, 3] = (cd * a_f + a_r) / M - r
    F[4, 4] = (cd * b_f + b_r) / M
    F[4, 5] = (cd * c_f + c_r) / M - vx
    F[5, 3] = (D_F * cd * a_f - D_R * a_r) / I_Z
    F[5, 4] = (D_F * cd * b_f - D_R * b_r) / I_Z
    F[5, 5] = (D_F * cd * c_f - D_R * c_r) / I_Z
    return f, F


def _quadcopter(x, u, par):
    """quadcopter, dynamics.py:40-59; par = [m, Ixx, Iyy, Izz]."""
    m, Ix, Iy, Iz = par[:4]
    ph, th, ps = x[3], x[4], x[5]
    p, q, r = x[9], x[10], x[11]
    sph, cph = np.sin(ph), np.cos(ph)
    sth, cth, tth = np.sin(th), np.cos(th), np.tan(th)
    sps, cps = np.sin(ps), np.cos(ps)
    a = q * sph + r * cph
    tm = u[0] / m
    f = np.array([x[6], x[7], x[8],
                  p + a * tth,
                  q * cph - r * sph,
                  a / cth,
                  tm * (sph * sps + cph * sth * cps),
                  tm * (cph * sth * sps - sph * cps),
                  tm * (cph * cth) - 9.81,
                  (u[1] - (Iz - Iy) * q * r) / Ix,
                  (u[2] - (Ix - Iz) * r * p) / Iy,
                  (u[3] - (Iy - Ix) * p * q) / Iz], dtype=x.dtype)
    F = np.zeros((12, 12), dtype=x.dtype)
    F[0, 6] = F[1, 7] = F[2, 8] = 1.0
    da_dph = q * cph - r * sph
    F[3, 3] = da_dph * tth
    F[3, 4] = a / cth ** 2
    F[3, 9] = 1.0
    F[3, 10] = sph * tth
    F[3, 11] = cph * tth
    F[4, 3] = -q * sph - r * cph
    F[4, 10] = cph
    F[4, 11] = -sph
    F[5, 3] = da_dph / cth
    F[5, 4] = a * sth / cth ** 2
    F[5, 10] = sph / cth
    F[5, 11] = cph / cth
    F[6, 3] = tm * (cph * sps - sph * sth * cps)
    F[6, 4] = tm * (cph * cth * cps)
    F[6, 5] = tm * (sph * cps - cph * sth * sps)
    F[7, 3] = tm * (-sph * sth * sps - cph * cps)
    F[7, 4] = tm * (cph * cth * sps)
    F[7, 5] = tm * (cph * sth * cps + sph * sps)
    F[8, 3] = tm * (-sph * cth)
    F[8, 4] = tm * (-cph * sth)
    F[9, 10] = -(Iz - Iy) * r / Ix
    F[9, 11] = -(Iz - Iy) * q / Ix
    F[10, 9] = -(Ix - Iz) * r / Iy
    F[10, 11] = -(Ix - Iz) * p / Iy
    F[11, 9] = -(Iy - Ix) * q / Iz
    F[11, 10] = -(Iy - Ix) * p / Iz
    return f, F


def dynamics(desc, x, u):
    """f(x,u), df/dx (n x n), df/du (n x m) of the problem's dynamics model at one node."""
    n, m = desc.n, desc.m
    did = desc.dyn_id
    dt = np.result_type(x.dtype, u.dtype if m > 0 else x.dtype)
    Fu = np.zeros((n, max(m, 1)), dtype=dt)
    if did == getattr(D, "DYN_USER", 6):
        # a user's own callable (not in nlp/dynamics.py): the ORACLE evaluates the callable itself and differentiates
        # it numerically (central differences, step 1e-6 (1 + |v|): Jacobian good to ~1e-9) -- independent of the
        # symbolic differentiation and generated CUDA source the product path uses
        func, params, mm = desc.user_func

        def call(xx, uu):
            try:
                out = func(xx, uu, params) if mm > 0 else func(xx, params)
            except TypeError:
                out = func(xx, uu) if mm > 0 else func(xx)
            return np.asarray(out, dtype=float).reshape(-1)
        xr, ur = np.asarray(x, dtype=float), np.asarray(u, dtype=float)[:max(m, 1)]
        f = call(xr, ur[:m])
        F = np.zeros((n, n))
        for j in range(n):
            h = 1e-6 * (1.0 + abs(xr[j]))
            xp, xm = xr.copy(), xr.copy()
            xp[j] += h
            xm[j] -= h
            F[:, j] = (call(xp, ur[:m]) - call(xm, ur[:m])) / (2 * h)
        for j in range(m):
            h = 1e-6 * (1.0 + abs(ur[j]))
            up, um = ur.copy(), ur.copy()
            up[j] += h
            um[j] -= h
            Fu[:, j] = (call(xr, up[:m]) - call(xr, um[:m])) / (2 * h)
        return f, F, Fu
    if did == D.DYN_AFFINE:
        f = desc.A @ x + desc.c0
        if m > 0:
            f = f + desc.Bm[:, :m] @ u[:m]
            Fu[:, :m] = desc.Bm[:, :m]
        return f, desc.A.astype(dt), Fu
    if did == D.DYN_VDP:                                   # dynamics.py:61-66
        f = np.array([(1 - x[1] ** 2) * x[0] - x[1] + u[0], x[0]], dtype=dt)
        F = np.array([[1 - x[1] ** 2, -2 * x[0] * x[1] - 1], [1.0, 0.0]], dtype=dt)
        Fu[0, 0] = 1.0
        return f, F, Fu
    if did == D.DYN_BICYCLE:                               # dynamics.py:117-136
        v = BICYCLE_V_GAIN * u[0] - BICYCLE_V_OFF
        dl = np.deg2rad(28) * u[1]
        f = np.array([v * np.cos(x[2]), v * np.sin(x[2]), 0.0, x[4], 0.0,
                      (v / BICYCLE_L) * np.tan(dl)], dtype=dt)
        F = np.zeros((6, 6), dtype=dt)
        F[0, 2] = -v * np.sin(x[2])
        F[1, 2] = v * np.cos(x[2])
        F[3, 4] = 1.0
        return f, F, None
    if did == D.DYN_VEHICLE:
        f, F = _vehicle6(x.astype(dt), u, desc.dyn_par)
        return f, F, None
    if did == D.DYN_VEHICLE_GNSS:                          # dynamics.py:166-174
        f6, F6 = _vehicle6(x[:6].astype(dt), u, desc.dyn_par)
        f = np.concatenate([f6, np.array([x[7], 0.0, 0.0], dtype=dt)])
        F = np.zeros((9, 9), dtype=dt)
        F[:6, :6] = F6
        F[6, 7] = 1.0
        return f, F, None
    if did == D.DYN_QUADCOPTER:
        f, F = _quadcopter(x.astype(dt), u, desc.dyn_par)
        return f, F, None
    raise ValueError("unknown dynamics id %r" % did)


def measurement(rtype, idx, x, g, rec):
    """h and its sparse gradient for one scalar residual.

    x: state vector at the measurement time (length n); g: global unknowns; rec: the data record
    [y, weight, aux...].  Returns (h, [(state_comp, dh/dx)...], [(global_idx, dh/dg)...])."""
    if rtype == D.RT_STATE:                                 # measurements.py:4
        return x[idx[0]], [(idx[0], 1.0)], []
    if rtype == D.RT_PR:                                    # measurements.py:56-70, 81-87
        d = np.array([x[idx[0]] - rec[2], x[idx[1]] - rec[3], x[idx[2]] - rec[4]])
        rho = np.sqrt(d[0] ** 2 + d[1] ** 2 + d[2] ** 2)
        h = rho + x[idx[3]]
        return h, [(idx[0], d[0] / rho), (idx[1], d[1] / rho), (idx[2], d[2] / rho), (idx[3], 1.0)], []
    if rtype == D.RT_PRR:                                   # measurements.py:72-79
        r = np.array([rec[2] - x[idx[0]], rec[3] - x[idx[1]], rec[4] - x[idx[2]]])
        rho = np.sqrt(r[0] ** 2 + r[1] ** 2 + r[2] ** 2)
        e = r / rho
        dv = np.array([rec[5] - x[idx[3]], rec[6] - x[idx[4]], rec[7] - x[idx[5]]])
        h = dv @ e + x[idx[6]]
        dp = -(dv - e * (e @ dv)) / rho
        gx = [(idx[0], dp[0]), (idx[1], dp[1]), (idx[2], dp[2]),
              (idx[3], -e[0]), (idx[4], -e[1]), (idx[5], -e[2]), (idx[6], 1.0)]
        return h, gx, []
    if rtype == D.RT_RANGE:                                 # measurements.py:13-19, 46-53
        dim = idx[6]
        d = np.array([x[idx[i]] - x[idx[3 + i]] for i in range(dim)])
        h = np.sqrt(np.sum(d ** 2) + .000001)
        gx = [(idx[i], d[i] / h) for i in range(dim)] + [(idx[3 + i], -d[i] / h) for i in range(dim)]
        return h, gx, []
    if rtype == D.RT_RANGE_PT:                              # measurements.py:8-12, 40-45
        dim = idx[6]
        pt = [g[idx[3 + i]] if idx[3 + i] >= 0 else rec[2 + i] for i in range(dim)]
        d = np.array([x[idx[i]] - pt[i] for i in range(dim)])
        h = np.sqrt(np.sum(d ** 2) + .000001)
        gx = [(idx[i], d[i] / h) for i in range(dim)]
        gg = [(idx[3 + i], -d[i] / h) for i in range(dim) if idx[3 + i] >= 0]
        return h, gx, gg
    if rtype == D.RT_HEADING:                               # measurements.py:30-36
        ry = x[idx[4]] - x[idx[1]]
        rx = x[idx[3]] - x[idx[0]] + .00001
        d2 = rx * rx + ry * ry
        h = np.arctan2(rx, ry)
        gx = [(idx[0], -ry / d2), (idx[1], rx / d2), (idx[3], ry / d2), (idx[4], -rx / d2)]
        return h, gx, []
    if rtype == D.RT_HEADING_PT:                            # measurements.py:25-29
        pt = [g[idx[3 + i]] if idx[3 + i] >= 0 else rec[2 + i] for i in range(2)]
        ry = pt[1] - x[idx[1]]
        rx = pt[0] - x[idx[0]]
        d2 = rx * rx + ry * ry
        h = np.arctan2(rx, ry)
        gx = [(idx[0], -ry / d2), (idx[1], rx / d2)]
        gg = []
        if idx[3] >= 0:
            gg.append((idx[3], ry / d2))
        if idx[4] >= 0:
            gg.append((idx[4], -rx / d2))
        return h, gx, gg
    raise ValueError("unknown residual type %r" % rtype)
