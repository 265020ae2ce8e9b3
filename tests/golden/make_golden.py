"""Generates the committed fixtures under tests/golden/ from the reference at /root/reference.

Run ONCE in the build container (the reference does not exist on the GPU box):
    python tests/golden/make_golden.py
Nothing in tests/, bench.py or smoke() reads /root/reference at run time; they read the .npz
files this script wrote.

What it captures
  collocation_N{3,5,10,15,20}.npz  tau, D, w, Phi produced by executing the reference class
        `nlp/collocation.py:7-121` itself (casadi/matplotlib/pdb stubbed: the class never
        touches them; python-2 integer division restored on lines 78 and 80).
  car.npz     inputs of autonomous-car.py (sim/sensor_data.pkl, sim/traj_data.pkl) and the
        reference's stored outputs filtering/{leastsquares,ekf,nlp-l2,nlp-huber}.pkl.
  ekf_ls.npz  outputs of the reference's own utils/ekf.py, utils/gnss.py,
        utils/leastsquares.py (imported unmodified from /root/reference/utils) on seeded
        synthetic inputs and on the gnss_stationary / two-receiver logs.
"""
import os
import pickle
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def _load_pkl(path):
    with open(path, "rb") as f:
        return pickle.load(f, encoding="latin1")


def reference_collocation_class():
    src = open(os.path.join(REF, "nlp/collocation.py")).read()
    lines = src.split("\n")
    assert "(self.N-a)/2 + 1" in lines[77] and "(self.N-a)/2)" in lines[79]
    lines[77] = lines[77].replace("(self.N-a)/2", "(self.N-a)//2")      # python-2 semantics
    lines[79] = lines[79].replace("(self.N-a)/2", "(self.N-a)//2")
    for name in ("casadi", "matplotlib", "matplotlib.pyplot", "pdb"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    ns = {}
    exec(compile("\n".join(lines), "ref_collocation", "exec"), ns)
    return ns["ChebyshevPseudospectralMethod"]


def make_collocation():
    cls = reference_collocation_class()
    for N, T in ((3, 1.0), (5, 2.0), (10, 50.0), (10, 5.0), (15, 63.0), (20, 10.0), (20, 40.0)):
        cpm = cls(N, 0, T)
        t = np.concatenate([np.linspace(0, T, 51), np.array([0.1234 * T, 0.987 * T])])
        Phi = np.array([cpm.evaluateLagrangePolynomials(tt) for tt in t])
        X = [np.array([np.sin(0.3 * k), np.cos(0.2 * k) * 100.0]) for k in range(N + 1)]
        xs = np.array([cpm.evaluateSolution(tt, X) for tt in t])
        np.savez_compressed(os.path.join(HERE, "collocation_N%d_T%g.npz" % (N, T)), N=N, T=T,
                            tau=cpm.tau, D=cpm.D, w=cpm.w, t=t, Phi=Phi, X=np.array(X), xs=xs,
                            t_nodes=cpm.tau2t(cpm.tau))
        print("collocation N=%d T=%g  sum(w)=%.6f" % (N, T, cpm.w.sum()))


def ref_utils():
    sys.path.insert(0, os.path.join(REF, "utils"))
    import utils as ru
    import gnss as rg
    import ekf as rekf
    import leastsquares as rls
    return ru, rg, rekf, rls


def make_car():
    ru, rg, rekf, rls = ref_utils()
    base = os.path.join(REF, "data/autonomous-car")
    sens = _load_pkl(base + "/sim/sensor_data.pkl")
    traj = _load_pkl(base + "/sim/traj_data.pkl")
    g = sens["gnss"]
    p_ref = ru.lla2ecef(np.array([37.4276, -122.1670, 0]))
    sat_ecef = np.array(g["sat_pos"])                       # (911, 11, 3)
    assert np.all(sat_ecef == sat_ecef[0])
    sat_enu = np.array([ru.ecef2enu(s, p_ref) for s in sat_ecef[0]])
    out = dict(p_ref_ECEF=p_ref, gnss_t=np.array(g["t"]), pr=np.array(g["pr"]), sat_ecef=sat_ecef[0],
               sat_enu=sat_enu, r_pr=float(g["R"]), alpha=float(g["alpha"]), b0=float(g["b0"]),
               traj_t=traj["t"], traj_u=traj["u"], traj_x=traj["x"], traj_x0=traj["x0"], dt=traj["dt"])
    for name in ("nlp-l2", "nlp-huber"):
        d = _load_pkl(base + "/filtering/%s.pkl" % name)
        key = name.replace("-", "_")
        out[key] = np.stack([np.array(d[k]) for k in ("x_ENU", "y_ENU", "z_ENU", "bias", "psi", "vx", "vy")], 1)
        out[key + "_t"] = np.array(d["t"])
        out[key + "_tsolve"] = np.array(d["t_solve"])
    d = _load_pkl(base + "/filtering/ekf.pkl")
    out["ekf"] = np.stack([np.array(d[k]) for k in ("x_ENU", "y_ENU", "z_ENU", "bias", "psi", "vx", "vy")], 1)
    d = _load_pkl(base + "/filtering/leastsquares.pkl")
    out["ls"] = np.stack([np.array(d[k]) for k in ("x_ECEF", "y_ECEF", "z_ECEF", "bias", "x_ENU", "y_ENU", "z_ENU",
                                                   "lat", "lon", "h")], 1)
    np.savez_compressed(os.path.join(HERE, "car.npz"), **out)
    print("car.npz written:", {k: np.shape(v) for k, v in out.items()})


if __name__ == "__main__":
    which = sys.argv[1:] or ["collocation", "car", "ekf_ls"]
    if "collocation" in which:
        make_collocation()
    if "car" in which:
        make_car()
    if "ekf_ls" in which:
        from make_golden_ekf_ls import make_ekf_ls
        make_ekf_ls()
