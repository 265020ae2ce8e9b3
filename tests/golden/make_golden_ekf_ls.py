"""Part of make_golden.py: runs the reference's OWN utils/ekf.py, utils/gnss.py,
utils/leastsquares.py, utils/utils.py, utils/data.py (imported unmodified from
/root/reference/utils) and stores inputs + outputs as ekf_ls.npz.

  gs_*    gnss_stationary.py:15-98 replayed: per-epoch LS, batch LS, 5-state EKF on the
          stationary-receiver log (51 epochs, 11-12 satellites)
  car_*   autonomous-car.py:120-179 replayed: 9-state vehicle EKF on the simulated car data
          (first 3000 steps; model functions exec'd from the script's own lines 18-85)
  rx_*    two-receiver logs: runLeastSquares incl. the velocity solve (gnss-multi-receiver.py:33-34)
  syn_*   seeded synthetic epochs / filters through the same reference functions
"""
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def _pad(list_of_arrays, smax, width=None):
    T = len(list_of_arrays)
    shape = (T, smax) if width is None else (T, smax, width)
    out = np.zeros(shape)
    n = np.zeros(T, dtype=np.int32)
    for i, a in enumerate(list_of_arrays):
        a = np.asarray(a)
        n[i] = a.shape[0]
        out[i, :a.shape[0]] = a
    return out, n


def make_ekf_ls():
    sys.path.insert(0, os.path.join(REF, "utils"))
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    import utils as ru
    import gnss as rg
    import ekf as rekf
    import leastsquares as rls
    import data as rdata
    import vehicle_sim as rvs
    out = {}
    p_ref = ru.lla2ecef(np.array([37.4276, -122.1670, 0]))
    out["p_ref_ECEF"] = p_ref

    # ---------------- gnss_stationary.py:15-98 ----------------
    data = rdata.load_gnss_logs(os.path.join(REF, "data/gnss_stationary/gnss_log_2020_02_05_09_14_15"))
    T = 50
    Q = np.diag([0.0001, 0.0001, 0.0001, 0.1, 0.001])
    r_pr = 100
    u = np.zeros((3, T + 1))
    ls_ecef = np.zeros((T + 1, 4))
    ls_enu = np.zeros((T + 1, 3))
    for k in range(T + 1):
        # NB: the reference's default argument x=np.zeros(3) persists across calls (warm start)
        p, b = rls.iterativeLeastSquares(data["sat_pos"][k], data["pr"][k])
        ls_ecef[k, :3] = p
        ls_ecef[k, 3] = b
        ls_enu[k] = ru.ecef2enu(p, p_ref)
    sat_b = np.vstack([data["sat_pos"][k] for k in range(T + 1)])
    pr_b = np.hstack([data["pr"][k] for k in range(T + 1)])
    t_b = np.hstack([[float(k)] * data["pr"][k].shape[0] for k in range(T + 1)])
    pb, b0, alpha = rls.iterativeLeastSquares_multiTimeStep(t_b, sat_b, pr_b)
    out["gs_batch"] = np.array([pb[0], pb[1], pb[2], b0, alpha])
    out["gs_batch_sat"], out["gs_batch_pr"], out["gs_batch_t"] = sat_b, pr_b, t_b
    sat_pad, nsat = _pad([data["sat_pos"][k] for k in range(T + 1)], 12, 3)
    pr_pad, _ = _pad([data["pr"][k] for k in range(T + 1)], 12)
    out["gs_sat_ecef"], out["gs_pr"], out["gs_nsat"] = sat_pad, pr_pad, nsat
    out["gs_ls_ecef"], out["gs_ls_enu"] = ls_ecef, ls_enu
    bias_rate_guess = (ls_ecef[-1, 3] - ls_ecef[0, 3]) / T
    xhat0 = np.array([ls_enu[0, 0], ls_enu[0, 1], ls_enu[0, 2], ls_ecef[0, 3], bias_rate_guess])
    out["gs_ekf_x0"] = xhat0.copy()
    f = rekf.EKF(rg.gnss_pos_and_bias, rg.multi_pseudorange, xhat0, np.diag([1.0, 1, 1, 1, 1]))
    sat_enu = np.zeros((T + 1, 12, 3))
    mu_tr = np.zeros((T + 1, 5))
    for k in range(T + 1):
        sk = np.array([ru.ecef2enu(s, p_ref) for s in data["sat_pos"][k]])
        sat_enu[k, :sk.shape[0]] = sk
        R = np.diag(r_pr * np.ones(data["pr"][k].shape[0]))
        f.update(u[:, k], data["pr"][k], Q, R, dyn_func_params={"dt": 1}, meas_func_params={"sat_pos": sk})
        mu_tr[k] = f.mu
    out["gs_sat_enu"], out["gs_ekf_traj"], out["gs_ekf_S"] = sat_enu, mu_tr, f.S.copy()

    # ---------------- autonomous-car.py:18-85, 120-179 ----------------
    src = open(os.path.join(REF, "autonomous-car.py")).read().split("\n")
    ns = {"np": np, "gnss": rg, "vehicle_sim": rvs}
    exec(compile("\n".join(src[17:85]), "car_models", "exec"), ns)
    import pickle
    ld = lambda p: pickle.load(open(p, "rb"), encoding="latin1")
    sens = ld(os.path.join(REF, "data/autonomous-car/sim/sensor_data.pkl"))["gnss"]
    traj = ld(os.path.join(REF, "data/autonomous-car/sim/traj_data.pkl"))
    car_params = rvs.get_parameters()
    Q_EKF = .001 * np.diag([0.01, 0.01, 0.01, 100, 500, 500, .001, .001, .001])
    r_car = float(sens["R"])
    x0 = np.hstack((traj["x0"], np.array([sens["b0"], sens["alpha"], 0.0]))).astype(float)
    f = rekf.EKF(ns["discrete_vehicle_dynamics"], ns["vehicle_sensors_model"], x0.copy(), np.diag(np.ones(9)))
    NST = 3000
    sat_enu_car = np.array([ru.ecef2enu(s, p_ref) for s in sens["sat_pos"][0]])
    car_tr = np.zeros((NST, 9))
    car_ns = np.zeros(NST, dtype=np.int32)
    car_z = np.zeros((NST, 11))
    for k in range(NST):
        t = traj["t"][k]
        i0 = np.where(sens["t"] <= t + .00001)[0][-1]
        i1 = np.where(sens["t"] >= t - .00001)[0][0]
        if i0 == i1:
            sk, Rk, prk = sat_enu_car, np.diag(r_car * np.ones(11)), sens["pr"][i0]
            car_ns[k] = 11
            car_z[k] = prk
        else:
            sk = Rk = prk = None
        f.update(traj["u"][:, k], prk, Q_EKF, Rk, dyn_func_params={"dt": traj["dt"], "car_params": car_params},
                 meas_func_params={"sat_pos": sk})
        car_tr[k] = f.mu
    out["car_x0"], out["car_u"], out["car_nsat"], out["car_z"] = x0, traj["u"][:, :NST].T.copy(), car_ns, car_z
    out["car_sat_enu"], out["car_traj"], out["car_S"], out["car_Q"], out["car_r"] = sat_enu_car, car_tr, f.S.copy(), np.diag(Q_EKF), r_car
    out["car_par"] = np.array([car_params[k] for k in ("C_AF", "C_AR", "M", "D_F", "D_R", "I_Z")], dtype=float)

    # ---------------- two-receiver logs: LS + velocity ----------------
    for tag, pref in (("A", "rec1/rec1_gnss_log_50y_moving_"), ("B", "rec2/rec2_gnss_log_50y_moving_")):
        d = rdata.load_gnss_logs(os.path.join(REF, "data/gnss-multi-receiver", pref))
        sol = rls.runLeastSquares(np.asarray(d["t"], float), d["sat_pos"], d["pr"], d["sat_vel"], d["pr_rate"], p_ref)
        sat, nsat = _pad(d["sat_pos"], 12, 3)
        vel, _ = _pad(d["sat_vel"], 12, 3)
        pr, _ = _pad(d["pr"], 12)
        prr, _ = _pad(d["pr_rate"], 12)
        out["rx%s_sat" % tag], out["rx%s_vel" % tag], out["rx%s_pr" % tag], out["rx%s_prr" % tag], out["rx%s_nsat" % tag] = sat, vel, pr, prr, nsat
        out["rx%s_sol" % tag] = np.stack([sol[k] for k in ("x_ECEF", "y_ECEF", "z_ECEF", "bias", "xd_ECEF", "yd_ECEF",
                                                           "zd_ECEF", "bias_rate", "x_ENU", "y_ENU", "z_ENU",
                                                           "xd_ENU", "yd_ENU", "zd_ENU", "lat", "lon", "h")], 1)
        stored = ld(os.path.join(REF, "data/gnss-multi-receiver/filtering/leastsquares%s.pkl" % tag))
        out["rx%s_stored" % tag] = np.stack([np.asarray(stored[k]) for k in ("x_ENU", "y_ENU", "z_ENU", "bias", "xd_ENU",
                                                                                 "yd_ENU", "zd_ENU", "bias_rate")], 1)

    # ---------------- seeded synthetic epochs and filters ----------------
    rng = np.random.default_rng(4000)
    Bs, S = 64, 12
    rx = p_ref[None, :] + rng.normal(0, 50.0, (Bs, 3))
    dirs = rng.normal(size=(Bs, S, 3))
    up = p_ref / np.linalg.norm(p_ref)
    dirs += 1.5 * up[None, None, :]
    dirs /= np.linalg.norm(dirs, axis=2, keepdims=True)
    sat = rx[:, None, :] + dirs * rng.uniform(2.0e7, 2.5e7, (Bs, S, 1))
    bias = rng.uniform(-1e3, 1e3, Bs)
    pr = np.linalg.norm(sat - rx[:, None, :], axis=2) + bias[:, None] + rng.normal(0, 3.0, (Bs, S))
    nsat = rng.integers(5, S + 1, Bs).astype(np.int32)
    sol = np.zeros((Bs, 4))
    for b in range(Bs):
        p, bb = rls.iterativeLeastSquares(sat[b, :nsat[b]], pr[b, :nsat[b]], x=np.zeros(3), b=0)   # cold start each
        sol[b, :3], sol[b, 3] = p, bb
    out["syn_ls_sat"], out["syn_ls_pr"], out["syn_ls_nsat"], out["syn_ls_sol"] = sat, pr, nsat, sol
    # EKF: 16 filters x 40 steps, 5-state, ragged satellite counts, ENU frame
    Bf, steps = 16, 40
    sat_e = np.zeros((Bf, steps, S, 3))
    z_e = np.zeros((Bf, steps, S))
    ns_e = rng.integers(4, S + 1, (Bf, steps)).astype(np.int32)
    u_e = rng.normal(0, 0.1, (Bf, steps, 3))
    x0_e = np.zeros((Bf, 5))
    tr_e = np.zeros((Bf, steps, 5))
    S_e = np.zeros((Bf, 5, 5))
    for b in range(Bf):
        pos = rng.normal(0, 5.0, 3)
        b0, al = rng.uniform(-1e3, 1e3), rng.normal(-190, 20)
        el = np.arcsin(rng.uniform(np.sin(np.deg2rad(10)), 1.0, S))
        az = rng.uniform(0, 2 * np.pi, S)
        rad = rng.uniform(2.0e7, 2.6e7, S)
        s0 = np.stack([rad * np.cos(el) * np.sin(az), rad * np.cos(el) * np.cos(az), rad * np.sin(el)], 1)
        x0_e[b] = np.array([pos[0] + 3, pos[1] - 2, pos[2] + 1, b0 + 5, al + 1])
        f = rekf.EKF(rg.gnss_pos_and_bias, rg.multi_pseudorange, x0_e[b].copy(), np.diag([1.0, 1, 1, 1, 1]))
        for t in range(steps):
            pos = pos + u_e[b, t]
            st = s0 + rng.normal(0, 1e3, (S, 3)) * t
            n = ns_e[b, t]
            zz = np.linalg.norm(pos[None, :] - st, axis=1) + b0 + al * (t + 1) + rng.normal(0, 10.0, S)
            sat_e[b, t], z_e[b, t] = st, zz
            f.update(u_e[b, t], zz[:n], Q, np.diag(100.0 * np.ones(n)), dyn_func_params={"dt": 1.0},
                     meas_func_params={"sat_pos": st[:n]})
            tr_e[b, t] = f.mu
        S_e[b] = f.S
    out["syn_ekf_x0"], out["syn_ekf_u"], out["syn_ekf_z"], out["syn_ekf_sat"], out["syn_ekf_nsat"] = x0_e, u_e, z_e, sat_e, ns_e
    out["syn_ekf_traj"], out["syn_ekf_S"], out["syn_ekf_Q"] = tr_e, S_e, np.diag(Q)
    np.savez_compressed(os.path.join(HERE, "ekf_ls.npz"), **out)
    print("ekf_ls.npz written:", {k: np.shape(v) for k, v in out.items()})
    # quick self-report against the stored pickles
    car_stored = np.load(os.path.join(HERE, "car.npz"))["ekf"]
    m = np.stack([car_tr[:, 0], car_tr[:, 1], car_tr[:, 8], car_tr[:, 6], car_tr[:, 2], car_tr[:, 3], car_tr[:, 4]], 1)
    print("car EKF live-vs-stored (stored[k+1] = state after update k):", np.abs(m[:-1] - car_stored[1:NST]).max())


if __name__ == "__main__":
    make_ekf_ls()
