"""Direct builders of `ProblemDesc` for the reference's example problem classes.

The `nlp.nlp` API (nlp_filter_b200/nlp/nlp.py) lowers a script-built problem to the same
descriptors (tests/test_api_compile.py checks that); these builders exist so that the batched
benchmark and the kernel tests can create a problem class without replaying hundreds of
`addResidualCost` calls.  Collocation constants come from
`nlp_filter_b200.nlp.collocation.ChebyshevPseudospectralMethod` (reference-exact).

Shapes:  VDP-MHE  estimation_example.py:10-45 pattern with
dynamics.van_der_pol;  GNSS  gnss_stationary.py:100-128;  CAR  autonomous-car.py:184-220;
RC  rc-car.py:85-114;  2RX  gnss-multi-receiver.py:51-131;  MRX  multi-receiver.py:62-99.
"""
import numpy as np

from . import desc as D
from .nlp.collocation import ChebyshevPseudospectralMethod

CAR_PARAMS = (1.1441e5, 1.3388e5, 2009.0, 1.53, 1.23, 2000.0)   # C_AF C_AR M D_F D_R I_Z (utils/vehicle_sim.py:10-16)


def _base(N, T, n, m, epochs):
    cpm = ChebyshevPseudospectralMethod(N, 0, T)
    d = D.ProblemDesc(N + 1, n, m, T)
    d.Dt = (2.0 / T) * cpm.D
    d.cw = (T / 2.0) * cpm.w
    d.cq = (T / 2.0) * cpm.w
    epochs = np.asarray(epochs, dtype=float).reshape(-1)
    d.E = epochs.shape[0]
    d.Phi = np.array([cpm.evaluateLagrangePolynomials(t) for t in epochs]).reshape(d.E, N + 1)
    d.cpm = cpm
    d.epochs = epochs
    return d


def _idx(*vals):
    out = np.zeros(D.NIDX, dtype=np.int32)
    out[:len(vals)] = vals
    return out


def affine(d, A=None, B=None, c0=None):
    d.dyn_id = D.DYN_AFFINE
    if A is not None:
        d.A = np.asarray(A, float)
    if B is not None:
        d.Bm = np.asarray(B, float)
    if c0 is not None:
        d.c0 = np.asarray(c0, float)
    return d


def car(loss=D.LOSS_L2, delta=5.0, n_sat=11, N=5, T=2.0, dt_gnss=0.1):
    """Vehicle MHE window: K=6, n=9, m=2, 21 epochs x 11 pseudorange slots, arrival cost,
    bounds on psi and vx (autonomous-car.py:184-220, 293-318)."""
    n_ep = int(np.floor(T / dt_gnss)) + 1
    d = _base(N, T, 9, 2, np.linspace(0, T, n_ep))
    d.dyn_id = D.DYN_VEHICLE_GNSS
    d.dyn_par[:6] = CAR_PARAMS
    d.qinv = 1.0 / np.array([0.01, 0.01, 0.01, 100, 500, 500, .001, .001, .001])
    d.loss = loss
    d.delta = delta
    d.has_prior = 1
    d.pinv = np.ones(9)
    d.bidx = np.array([2, 3], dtype=np.int32)
    d.blo = np.array([-np.pi, 0.0])
    d.bhi = np.array([np.pi, np.inf])
    rt, ep, ix = [], [], []
    for e in range(n_ep):
        for j in range(n_sat):
            rt.append(D.RT_PR)
            ep.append(e)
            ix.append(_idx(0, 1, 8, 6))       # vehicle_pseudorange, measurements.py:81-87
    d.set_residuals(rt, ep, ix)
    return d.finalize()


def gnss(N=10, T=50.0, n_epochs=51, n_sat=12):
    """Stationary-receiver MHE: K=11, n=5, m=3, 51 epochs x <=12 pseudoranges
    (gnss_stationary.py:100-128)."""
    d = _base(N, T, 5, 3, np.linspace(0, T, n_epochs))
    A = np.zeros((5, 5))
    A[3, 4] = 1.0
    B = np.zeros((5, 3))
    B[0, 0] = B[1, 1] = B[2, 2] = 1.0
    affine(d, A, B)
    d.qinv = 1.0 / np.array([0.0001, 0.0001, 0.0001, 0.1, 0.001])
    rt, ep, ix = [], [], []
    for e in range(n_epochs):
        for j in range(n_sat):
            rt.append(D.RT_PR)
            ep.append(e)
            ix.append(_idx(0, 1, 2, 3))
    d.set_residuals(rt, ep, ix)
    return d.finalize()


def vdp_mhe(N=20, T=10.0, n_epochs=50, q=1e-2, r=(0.01, 0.02)):
    """Van der Pol MHE window: K=21, n=2, m=1, 50 full-state epochs (estimation_example.py:10-45
    with dynamics.van_der_pol; Q = 1e-2 I)."""
    d = _base(N, T, 2, 1, np.linspace(0, T, n_epochs))
    d.dyn_id = D.DYN_VDP
    d.qinv = np.array([1.0 / q, 1.0 / q])
    rt, ep, ix = [], [], []
    for e in range(n_epochs):
        for i in range(2):
            rt.append(D.RT_STATE)
            ep.append(e)
            ix.append(_idx(i))
    d.set_residuals(rt, ep, ix)
    d.r_inv = 1.0 / np.asarray(r, float)
    return d.finalize()


def rc(N=20, T=40.0, n_epochs=41, n_sat=12):
    """RC-car window: K=21, n=6, m=2, pseudoranges on [0,1,2,3] (rc-car.py:85-114)."""
    d = _base(N, T, 6, 2, np.linspace(0, T, n_epochs))
    d.dyn_id = D.DYN_BICYCLE
    d.qinv = 1.0 / np.array([1, 1, 0.001, .01, .01, 1])
    rt, ep, ix = [], [], []
    for e in range(n_epochs):
        for j in range(n_sat):
            rt.append(D.RT_PR)
            ep.append(e)
            ix.append(_idx(0, 1, 2, 3))
    d.set_residuals(rt, ep, ix)
    return d.finalize()


def two_receiver(N=10, T=5.0, n_sat=10, fold_equality=True):
    """Two-receiver window: K=11, n=10, m=6; per 1 Hz epoch 10 pseudorange slots per receiver,
    51 range + 51 heading residuals at 10 Hz, arrival cost, z_A == z_B at every node
    (gnss-multi-receiver.py:42-131)."""
    t_pr = np.linspace(0, T, int(np.floor(T / 1.0)) + 1)
    t_rh = np.linspace(0, T, int(np.floor(T / 0.1)) + 1)
    epochs = np.concatenate([t_rh, t_pr])                  # epoch rows: 51 fast ones, then 6 slow
    d = _base(N, T, 10, 6, epochs)
    A = np.zeros((10, 10))
    A[3, 4] = 1.0
    A[8, 9] = 1.0
    B = np.zeros((10, 6))
    for i, r in enumerate((0, 1, 2, 5, 6, 7)):
        B[r, i] = 1.0
    affine(d, A, B)
    d.qinv = 1.0 / np.array([.01, .01, .01, 0.01, 0.01, .01, .01, .01, 0.01, 0.01])
    d.has_prior = 1
    d.pinv = 1.0 / (0.01 * np.array([1, 1, 1, 0.1, 0.1, 1, 1, 1, 0.1, 0.1]))
    if fold_equality:                                       # X_k[7] := X_k[2]
        d.cmap = np.array([0, 1, 2, 3, 4, 5, 6, 2, 7, 8], dtype=np.int32)
        d.resize_nv(9)
    rt, ep, ix = [], [], []
    n_rh = t_rh.shape[0]
    for e in range(n_rh):                                   # range, measurements.py:46-53
        rt.append(D.RT_RANGE)
        ep.append(e)
        ix.append(_idx(0, 1, 2, 5, 6, 7, 3))
    for e in range(n_rh):                                   # heading, measurements.py:30-36
        rt.append(D.RT_HEADING)
        ep.append(e)
        ix.append(_idx(0, 1, 0, 5, 6, 0, 2))
    for e in range(t_pr.shape[0]):
        for j in range(n_sat):
            rt.append(D.RT_PR)
            ep.append(n_rh + e)
            ix.append(_idx(0, 1, 2, 3))
            rt.append(D.RT_PR)
            ep.append(n_rh + e)
            ix.append(_idx(5, 6, 7, 8))
    d.set_residuals(rt, ep, ix)
    return d.finalize()


def multi_receiver(N=15, T=63.0, n_epochs=64, n_sat=11):
    """multi-receiver.py:62-99 shape: K=16, n=8, m=0, three global unknowns XA; per epoch
    pseudorange + pseudorange-rate per satellite and one planar range to XA."""
    d = _base(N, T, 8, 0, np.linspace(0, T, n_epochs))
    A = np.zeros((8, 8))
    for i in range(4):
        A[i, 4 + i] = 1.0
    affine(d, A, np.zeros((8, 1)))
    d.ng = 3
    d.qinv = 1.0 / np.array([0.01, 0.01, 0.01, 0.01, 1., 1., 0.01, 0.01])
    rt, ep, ix = [], [], []
    for e in range(n_epochs):
        for j in range(n_sat):
            rt.append(D.RT_PR); ep.append(e); ix.append(_idx(0, 1, 2, 3))
            rt.append(D.RT_PRR); ep.append(e); ix.append(_idx(0, 1, 2, 4, 5, 6, 7))
        rt.append(D.RT_RANGE_PT); ep.append(e); ix.append(_idx(0, 1, 0, 0, 1, -1, 2))
    d.set_residuals(rt, ep, ix)
    return d.finalize()


def generic_model(dyn_id, n, m, K=7, T=3.0, n_epochs=9, extra=()):
    """Small problem around any kernel dynamics model with full-state measurements of every
    component (+ optional extra residual rows): used by the parity tests of the model library."""
    d = _base(K - 1, T, n, m, np.linspace(0, T, n_epochs))
    d.dyn_id = dyn_id
    d.dyn_par[:6] = CAR_PARAMS
    if dyn_id == D.DYN_QUADCOPTER:
        d.dyn_par[:4] = [1.3, 0.02, 0.03, 0.04]
    if dyn_id == D.DYN_AFFINE:
        rng = np.random.default_rng(99)
        d.A, d.Bm, d.c0 = rng.normal(size=(n, n)), rng.normal(size=(n, max(m, 1))), rng.normal(size=n)
    d.qinv = 1.0 + np.arange(n) * 0.5
    rt, ep, ix = [], [], []
    for e in range(n_epochs):
        for i in range(n):
            rt.append(D.RT_STATE); ep.append(e); ix.append(_idx(i))
        for (t, idx) in extra:
            rt.append(t); ep.append(e); ix.append(_idx(*idx))
    d.set_residuals(rt, ep, ix)
    return d.finalize()


def vdp_segment(Ns=5, Ts=1.0, n_epochs=5, q=1e-2, r=(0.01, 0.02)):
    """One segment of a long van der Pol window (multi-segment collocation): K = Ns + 1 nodes on
    [0, Ts], n_epochs full-state measurement epochs strictly inside the segment (so no epoch belongs to two
    segments).  40 such segments = a 200-epoch horizon (BASELINE.json configs[3]'s "horizon 200" read as one window)."""
    t = (np.arange(n_epochs) + 0.5) * (Ts / n_epochs)
    d = _base(Ns, Ts, 2, 1, t)
    d.dyn_id = D.DYN_VDP
    d.qinv = np.array([1.0 / q, 1.0 / q])
    rt, ep, ix = [], [], []
    for e in range(n_epochs):
        for i in range(2):
            rt.append(D.RT_STATE)
            ep.append(e)
            ix.append(_idx(i))
    d.set_residuals(rt, ep, ix)
    d.r_inv = 1.0 / np.asarray(r, float)
    return d.finalize()
