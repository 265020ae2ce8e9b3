"""Seeded synthetic batches of the benchmark shapes: per-window input
buffers in exactly the layout the C ABI takes (z0, U, data, xprior as (B, ...) float64).

  vdp_mhe_batch   Van der Pol MHE windows: RK4 truth, noisy full-state measurements at 50 epochs
                  (estimation_example.py:10-45 pattern with dynamics.van_der_pol).
  gnss_batch      stationary-receiver pseudorange windows, 51 epochs x 12 satellites, ENU frame
                  (gnss_stationary.py:100-128).
Everything is NumPy on the host; no device code here.
"""
import numpy as np

from . import desc as D


def _interp_rows(t_src, y_src, t_dst):
    """Linear interpolation with linear extrapolation of (B, T) rows -> (B, len(t_dst))
    (what scipy interp1d(fill_value='extrapolate') does in nlp.py:297,306)."""
    t_src = np.asarray(t_src, float)
    idx = np.clip(np.searchsorted(t_src, t_dst, side="right") - 1, 0, len(t_src) - 2)
    t0, t1 = t_src[idx], t_src[idx + 1]
    w = (t_dst - t0) / (t1 - t0)
    return y_src[..., idx] * (1 - w) + y_src[..., idx + 1] * w


def vdp_rhs(x0, x1, u):
    return (1 - x1 ** 2) * x0 - x1 + u, x0        # nlp/dynamics.py:61-66


def vdp_mhe_batch(desc, B, seed=2000):
    rng = np.random.default_rng(seed)
    T = desc.T
    epochs = desc.epochs
    x0 = np.stack([rng.normal(0.0, 0.25, B), rng.normal(1.0, 0.25, B)], 1)
    phase = rng.uniform(0, 2 * np.pi, B)
    ufun = lambda t: 0.5 * np.sin(t + phase)
    dt = 0.01
    nstep = int(round(T / dt))
    xs = np.zeros((B, 2, nstep + 1))
    xs[:, :, 0] = x0
    a, b = x0[:, 0].copy(), x0[:, 1].copy()
    for i in range(nstep):
        t = i * dt
        k1 = vdp_rhs(a, b, ufun(t))
        k2 = vdp_rhs(a + 0.5 * dt * k1[0], b + 0.5 * dt * k1[1], ufun(t + 0.5 * dt))
        k3 = vdp_rhs(a + 0.5 * dt * k2[0], b + 0.5 * dt * k2[1], ufun(t + 0.5 * dt))
        k4 = vdp_rhs(a + dt * k3[0], b + dt * k3[1], ufun(t + dt))
        a = a + dt / 6 * (k1[0] + 2 * k2[0] + 2 * k3[0] + k4[0])
        b = b + dt / 6 * (k1[1] + 2 * k2[1] + 2 * k3[1] + k4[1])
        xs[:, 0, i + 1], xs[:, 1, i + 1] = a, b
    tt = np.arange(nstep + 1) * dt
    truth = np.stack([_interp_rows(tt, xs[:, 0, :], epochs), _interp_rows(tt, xs[:, 1, :], epochs)], 1)  # B,2,E
    rvar = 1.0 / desc.r_inv
    y = truth + rng.normal(size=truth.shape) * np.sqrt(rvar)[None, :, None]
    data = np.zeros((B, desc.Ldata))
    for e in range(desc.E):
        for i in range(2):
            o = desc.roff[e * 2 + i]
            data[:, o] = y[:, i, e]
            data[:, o + 1] = desc.r_inv[i]
    tn = desc.cpm.tau2t(desc.cpm.tau)
    u_epochs = 0.5 * np.sin(epochs[None, :] + phase[:, None])
    U = _interp_rows(epochs, u_epochs, tn)[:, :, None]                      # B,K,1
    z0 = np.stack([_interp_rows(epochs, y[:, 0, :], tn), _interp_rows(epochs, y[:, 1, :], tn)], 2).reshape(B, -1)
    return dict(z0=z0, U=np.ascontiguousarray(U), data=data, xprior=None, truth=truth, y=y)


def gnss_batch(desc, B, seed=3000, n_sat=12, r_pr=100.0):
    rng = np.random.default_rng(seed)
    E, K = desc.E, desc.K
    epochs = desc.epochs
    pos = rng.normal(0.0, 5.0, (B, 3))
    b0 = rng.uniform(-1e3, 1e3, B)
    alpha = rng.normal(-190.0, 20.0, B)
    # satellites on a shell above 10 deg elevation, slow linear motion
    el = np.arcsin(rng.uniform(np.sin(np.deg2rad(10)), 1.0, (B, n_sat)))
    az = rng.uniform(0, 2 * np.pi, (B, n_sat))
    rad = rng.uniform(2.0e7, 2.6e7, (B, n_sat))
    s0 = np.stack([rad * np.cos(el) * np.sin(az), rad * np.cos(el) * np.cos(az), rad * np.sin(el)], 2)  # B,S,3
    sv = rng.normal(0.0, 1.5e3, (B, n_sat, 3))
    sat = s0[:, None, :, :] + sv[:, None, :, :] * epochs[None, :, None, None]                           # B,E,S,3
    rho = np.sqrt(np.sum((pos[:, None, None, :] - sat) ** 2, -1))
    bias = b0[:, None] + alpha[:, None] * epochs[None, :]
    pr = rho + bias[:, :, None] + rng.normal(0.0, np.sqrt(r_pr), rho.shape)
    data = np.zeros((B, desc.Ldata))
    off = desc.roff.reshape(E, n_sat)
    for e in range(E):
        for j in range(n_sat):
            o = off[e, j]
            data[:, o] = pr[:, e, j]
            data[:, o + 1] = 1.0 / r_pr
            data[:, o + 2:o + 5] = sat[:, e, j, :]
    tn = desc.cpm.tau2t(desc.cpm.tau)
    # initial guess: origin, clock bias/drift from the mean pseudorange misfit at the origin
    mis = np.mean(pr - np.sqrt(np.sum(sat ** 2, -1)), 2)                      # B,E
    a_g = (mis[:, -1] - mis[:, 0]) / (epochs[-1] - epochs[0])
    z0 = np.zeros((B, K, 5))
    z0[:, :, 3] = mis[:, :1] + a_g[:, None] * tn[None, :]
    z0[:, :, 4] = a_g[:, None]
    U = np.zeros((B, K, 3))
    return dict(z0=z0.reshape(B, -1), U=U, data=data, xprior=None, pos=pos, b0=b0, alpha=alpha)


# ---------------------------------------------------------------------------------------------
# vehicle chains (autonomous-car.py:184-288 shape): B runs, each a chain of receding-horizon windows
# ---------------------------------------------------------------------------------------------
def _vehicle_rhs(x, u, par):
    """Single-track model with linear tyres (nlp/dynamics.py:148-164) + clock (166-174), vectorised
    over runs: x (B,9), u (B,2)."""
    C_AF, C_AR, M, D_F, D_R, I_Z = par[:6]
    psi, vx, vy, r = x[:, 2], x[:, 3], x[:, 4], x[:, 5]
    den = vx + 0.001
    Fyr = -C_AR * (vy - D_R * r) / den
    Fyf = -C_AF * ((vy + D_F * r) / den - u[:, 1])
    d = np.zeros_like(x)
    d[:, 0] = vx * np.cos(psi) - vy * np.sin(psi)
    d[:, 1] = vx * np.sin(psi) + vy * np.cos(psi)
    d[:, 2] = r
    d[:, 3] = (-Fyf * np.sin(u[:, 1]) + u[:, 0]) / M + r * vy
    d[:, 4] = (Fyf * np.cos(u[:, 1]) + Fyr) / M - r * vx
    d[:, 5] = (D_F * Fyf * np.cos(u[:, 1]) - D_R * Fyr) / I_Z
    d[:, 6] = x[:, 7]
    return d


def car_chain_batch(desc, B, n_windows, seed=4000, DT=1.0, alpha=200.0, r_pr=10.0, dt_gnss=0.1):
    """Synthetic vehicle runs: truth by RK4 (dt 0.01) of the model the estimator uses, smooth random
    throttle/steering, 11 satellites, pseudoranges at 10 Hz with clock drift `alpha` and variance
    `r_pr` (utils/vehicle_sim.py:119-121).  Returns per-window inputs for the whole chain:
    U (n_windows, B, K, 2), data (n_windows, B, Ldata), xhat0 (B, 9), truth at window starts."""
    rng = np.random.default_rng(seed)
    par = desc.dyn_par
    T = desc.T
    dt = 0.01
    t_end = (n_windows - 1) * DT + T
    nstep = int(round(t_end / dt))
    ph = rng.uniform(0, 2 * np.pi, (B, 2))
    amp = rng.uniform(0.5, 1.5, (B, 2))

    def ctrl(t):
        return np.stack([400.0 + 300.0 * amp[:, 0] * np.sin(0.3 * t + ph[:, 0]),
                         0.03 * amp[:, 1] * np.sin(0.5 * t + ph[:, 1])], 1)

    x = np.zeros((B, 9))
    x[:, 2] = rng.uniform(-1.0, 1.0, B)
    x[:, 3] = 10.0
    x[:, 7] = alpha
    xs = np.zeros((nstep + 1, B, 9))
    xs[0] = x
    for i in range(nstep):
        t = i * dt
        k1 = _vehicle_rhs(x, ctrl(t), par)
        k2 = _vehicle_rhs(x + 0.5 * dt * k1, ctrl(t + 0.5 * dt), par)
        k3 = _vehicle_rhs(x + 0.5 * dt * k2, ctrl(t + 0.5 * dt), par)
        k4 = _vehicle_rhs(x + dt * k3, ctrl(t + dt), par)
        x = x + dt / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
        xs[i + 1] = x
    n_sat = 11
    el = np.arcsin(rng.uniform(np.sin(np.deg2rad(15)), 1.0, n_sat))
    az = rng.uniform(0, 2 * np.pi, n_sat)
    rad = rng.uniform(2.0e7, 2.6e7, n_sat)
    sat = np.stack([rad * np.cos(el) * np.sin(az), rad * np.cos(el) * np.cos(az), rad * np.sin(el)], 1)
    tn = desc.cpm.tau2t(desc.cpm.tau)
    n_ep = desc.E
    K = desc.K
    U = np.zeros((n_windows, B, K, 2))
    data = np.zeros((n_windows, B, desc.Ldata))
    off = desc.roff.reshape(n_ep, n_sat)
    every = int(round(dt_gnss / dt))
    w_meas = dt_gnss * (1.0 / r_pr)
    for w in range(n_windows):
        t0 = w * DT
        for k in range(K):
            U[w, :, k, :] = ctrl(t0 + tn[k])
        i0 = int(round(t0 / dt))
        for e in range(n_ep):
            xe = xs[i0 + e * every]
            p = np.stack([xe[:, 0], xe[:, 1], xe[:, 8]], 1)
            rho = np.sqrt(np.sum((p[:, None, :] - sat[None, :, :]) ** 2, 2))
            pr = rho + xe[:, 6:7] + rng.normal(0.0, np.sqrt(r_pr), rho.shape)
            for j in range(n_sat):
                o = off[e, j]
                data[w, :, o] = pr[:, j]
                data[w, :, o + 1] = w_meas
                data[w, :, o + 2:o + 5] = sat[j]
    starts = xs[(np.arange(n_windows) * int(round(DT / dt)))]
    return dict(U=U, data=data, xhat0=xs[0].copy(), truth=starts, PhiDT=desc.cpm.basisMatrix([DT]))


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


def vdp_multiseg_batch(seg, S, B, seed=6000):
    """Long van der Pol windows for the multi-segment path: truth by RK4 (dt 0.005) over [0, S*Ts], noisy full-state
    measurements at every segment's epochs, controls at every segment's nodes, start point = measurements
    interpolated to the (shared) nodes.  Returns z0 (B, S*Ns*2 + 2), U (B, S, K, 1), data (B, S, Ldata_seg)."""
    rng = np.random.default_rng(seed)
    Ts, K, Ns = seg.T, seg.K, seg.K - 1
    T = S * Ts
    x0 = np.stack([rng.normal(0.0, 0.25, B), rng.normal(1.0, 0.25, B)], 1)
    phase = rng.uniform(0, 2 * np.pi, B)
    ufun = lambda t: 0.5 * np.sin(t + phase)
    dt = 0.005
    nstep = int(round(T / dt))
    xs = np.zeros((B, 2, nstep + 1))
    a, b = x0[:, 0].copy(), x0[:, 1].copy()
    xs[:, 0, 0], xs[:, 1, 0] = a, b
    for i in range(nstep):
        t = i * dt
        k1 = vdp_rhs(a, b, ufun(t))
        k2 = vdp_rhs(a + 0.5 * dt * k1[0], b + 0.5 * dt * k1[1], ufun(t + 0.5 * dt))
        k3 = vdp_rhs(a + 0.5 * dt * k2[0], b + 0.5 * dt * k2[1], ufun(t + 0.5 * dt))
        k4 = vdp_rhs(a + dt * k3[0], b + dt * k3[1], ufun(t + dt))
        a = a + dt / 6 * (k1[0] + 2 * k2[0] + 2 * k3[0] + k4[0])
        b = b + dt / 6 * (k1[1] + 2 * k2[1] + 2 * k3[1] + k4[1])
        xs[:, 0, i + 1], xs[:, 1, i + 1] = a, b
    tt = np.arange(nstep + 1) * dt
    t_ep = (np.arange(S)[:, None] * Ts + seg.epochs[None, :]).reshape(-1)              # global epoch times
    truth = np.stack([_interp_rows(tt, xs[:, 0, :], t_ep), _interp_rows(tt, xs[:, 1, :], t_ep)], 1)   # B, 2, S*E
    rvar = 1.0 / seg.r_inv
    y = truth + rng.normal(size=truth.shape) * np.sqrt(rvar)[None, :, None]
    E = seg.E
    data = np.zeros((B, S, seg.Ldata))
    for s in range(S):
        for e in range(E):
            for i in range(2):
                o = seg.roff[e * 2 + i]
                data[:, s, o] = y[:, i, s * E + e]
                data[:, s, o + 1] = seg.r_inv[i]
    tn_loc = seg.cpm.tau2t(seg.cpm.tau)
    U = np.zeros((B, S, K, 1))
    for s in range(S):
        U[:, s, :, 0] = 0.5 * np.sin((s * Ts + tn_loc)[None, :] + phase[:, None])
    t_nodes = np.concatenate([s * Ts + tn_loc[(0 if s == 0 else 1):] for s in range(S)])     # S*Ns + 1 shared nodes
    z0 = np.stack([_interp_rows(t_ep, y[:, 0, :], t_nodes), _interp_rows(t_ep, y[:, 1, :], t_nodes)], 2).reshape(B, -1)
    x_nodes = np.stack([_interp_rows(tt, xs[:, 0, :], t_nodes), _interp_rows(tt, xs[:, 1, :], t_nodes)], 2)
    return dict(z0=z0, U=U, data=data, truth_nodes=x_nodes, t_nodes=t_nodes, x0=x0)


def rc_batch(d, B, seed=7000, n_sat=12, r_pr=1.0):
    """RC-car windows (rc-car.py:85-114 shape: K=21, n=6, m=2, 41 epochs x 12 pseudoranges on state components
    [0, 1, 2, 3]): truth by RK4 of the model the estimator uses (nlp/dynamics.py:117-136, quirks included), smooth
    throttle / steering, start point = truth + noise (the script's zero start needs 100+ iterations)."""
    rng = np.random.default_rng(seed)
    T, E, K = d.T, d.E, d.K
    epochs = d.epochs
    ph = rng.uniform(0, 2 * np.pi, (B, 2))

    def ctrl(t):
        return np.stack([0.35 + 0.1 * np.sin(0.2 * t + ph[:, 0]), 0.3 * np.sin(0.15 * t + ph[:, 1])], 1)

    def rhs(x, u):
        v = 8.72649116358 * u[:, 0] - 0.856053299155
        dl = np.deg2rad(28) * u[:, 1]
        f = np.zeros_like(x)
        f[:, 0] = v * np.cos(x[:, 2])
        f[:, 1] = v * np.sin(x[:, 2])
        f[:, 3] = x[:, 4]
        f[:, 5] = (v / 0.28) * np.tan(dl)
        return f

    x = np.zeros((B, 6))
    x[:, 2] = rng.uniform(-1.0, 1.0, B)
    x[:, 3] = rng.uniform(-100, 100, B)
    x[:, 4] = rng.normal(0.0, 1.0, B)
    dt = 0.02
    nstep = int(round(T / dt))
    xs = np.zeros((nstep + 1, B, 6))
    xs[0] = x
    for i in range(nstep):
        t = i * dt
        k1 = rhs(x, ctrl(t))
        k2 = rhs(x + 0.5 * dt * k1, ctrl(t + 0.5 * dt))
        k3 = rhs(x + 0.5 * dt * k2, ctrl(t + 0.5 * dt))
        k4 = rhs(x + dt * k3, ctrl(t + dt))
        x = x + dt / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
        xs[i + 1] = x
    el = np.arcsin(rng.uniform(np.sin(np.deg2rad(15)), 1.0, (B, n_sat)))
    az = rng.uniform(0, 2 * np.pi, (B, n_sat))
    rad = rng.uniform(2.0e7, 2.6e7, (B, n_sat))
    sat = np.stack([rad * np.cos(el) * np.sin(az), rad * np.cos(el) * np.cos(az), rad * np.sin(el)], 2)
    data = np.zeros((B, d.Ldata))
    off = d.roff.reshape(E, n_sat)
    for e in range(E):
        xe = xs[int(round(epochs[e] / dt))]
        rho = np.sqrt(np.sum((xe[:, None, :3] - sat) ** 2, 2))
        pr = rho + xe[:, 3:4] + rng.normal(0.0, np.sqrt(r_pr), rho.shape)
        for j in range(n_sat):
            o = off[e, j]
            data[:, o] = pr[:, j]
            data[:, o + 1] = 1.0 / r_pr
            data[:, o + 2:o + 5] = sat[:, j]
    tn = d.cpm.tau2t(d.cpm.tau)
    U = np.stack([ctrl(t) for t in tn], 1)
    idx = np.clip(np.round(tn / dt).astype(int), 0, nstep)
    z0 = xs[idx].transpose(1, 0, 2) + rng.normal(0, 0.3, (B, K, 6)) * np.array([1, 1, 0.05, 1, 0.05, 0.05])
    return dict(z0=z0.reshape(B, -1), U=np.ascontiguousarray(U), data=data, xprior=None, truth=xs[idx].transpose(1, 0, 2))


def multi_receiver_batch(d, B, seed=8000):
    """multi-receiver.py:62-99 shape (K=16, n=8, m=0, three global unknowns XA; per epoch pseudorange + pseudorange rate
    per satellite and one planar range to XA): a receiver moving at constant velocity, start point = the truth's
    neighbourhood (the script's zero start is multi-modal)."""
    rng = np.random.default_rng(seed)
    n_sat = (d.M // d.E - 1) // 2
    pos = np.array([-545.0, -53.0, 75.0]) + rng.normal(0, 1, (B, 3))
    vel = rng.normal(0, 0.3, (B, 3))
    xa = pos[:, :2] + np.array([1.7, 1.7])
    el = np.arcsin(rng.uniform(np.sin(np.deg2rad(15)), 1.0, (B, n_sat)))
    az = rng.uniform(0, 2 * np.pi, (B, n_sat))
    rad = rng.uniform(2.0e7, 2.6e7, (B, n_sat))
    sat = np.stack([rad * np.cos(el) * np.sin(az), rad * np.cos(el) * np.cos(az), rad * np.sin(el)], 2)
    sv = rng.normal(0, 2e3, (B, n_sat, 3))
    data = np.zeros((B, d.Ldata))
    r = 0
    for e in range(d.E):
        t = d.epochs[e]
        p = pos + vel * t
        for j in range(n_sat):
            s = sat[:, j] + sv[:, j] * t
            los = s - p
            rho = np.linalg.norm(los, axis=1)
            o = d.roff[r]; data[:, o] = rho + 13.5 + 0.7 * t + rng.normal(0, 3, B); data[:, o + 1] = 0.01; data[:, o + 2:o + 5] = s; r += 1
            o = d.roff[r]; data[:, o] = np.sum((sv[:, j] - vel) * los / rho[:, None], 1) + 0.7 + rng.normal(0, 0.1, B)
            data[:, o + 1] = 10.0; data[:, o + 2:o + 5] = s; data[:, o + 5:o + 8] = sv[:, j]; r += 1
        o = d.roff[r]; data[:, o] = 2.4384; data[:, o + 1] = 100.0; r += 1
    assert r == d.M
    z0 = np.zeros((B, d.K, 8))
    z0[:, :, :3] = pos[:, None, :]
    z0[:, :, 3] = 13.5
    z0 = np.concatenate([z0.reshape(B, -1), xa, np.zeros((B, 1))], 1)
    return dict(z0=z0, U=None, data=data, xprior=None)
