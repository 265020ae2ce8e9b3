"""Analytic Jacobians of the oracle's model library against complex-step differentiation of
the value functions, and a sympy check of two hand-derived formulas."""
import numpy as np
import pytest

from nlp_filter_b200 import desc as D
from nlp_filter_b200 import shapes
from oracle import models, nls


def _desc_for(dyn_id, n, m):
    d = D.ProblemDesc(4, n, m, 1.0)
    d.dyn_id = dyn_id
    d.dyn_par[:6] = shapes.CAR_PARAMS
    if dyn_id == D.DYN_QUADCOPTER:
        d.dyn_par[:4] = [1.3, 0.02, 0.03, 0.04]
    if dyn_id == D.DYN_AFFINE:
        rng = np.random.default_rng(0)
        d.A = rng.normal(size=(n, n))
        d.Bm = rng.normal(size=(n, max(m, 1)))
        d.c0 = rng.normal(size=n)
    return d


@pytest.mark.parametrize("dyn_id,n,m", [(D.DYN_AFFINE, 5, 3), (D.DYN_VDP, 2, 1), (D.DYN_BICYCLE, 6, 2),
                                        (D.DYN_VEHICLE, 6, 2), (D.DYN_VEHICLE_GNSS, 9, 2), (D.DYN_QUADCOPTER, 12, 4)])
def test_dynamics_jacobian_complex_step(dyn_id, n, m):
    d = _desc_for(dyn_id, n, m)
    rng = np.random.default_rng(1)
    for trial in range(5):
        x = rng.normal(size=n)
        if dyn_id in (D.DYN_VEHICLE, D.DYN_VEHICLE_GNSS):
            x[3] = 5.0 + rng.uniform()
        u = rng.normal(size=max(m, 1)) * 0.3
        f, F, Fu = models.dynamics(d, x, u)
        h = 1e-30
        for j in range(n):
            xc = x.astype(complex)
            xc[j] += 1j * h
            fc, _, _ = models.dynamics(d, xc, u.astype(complex))
            assert np.max(np.abs(np.imag(fc) / h - F[:, j])) <= 1e-11 * (1 + np.max(np.abs(F[:, j])))
        if Fu is not None and dyn_id in (D.DYN_AFFINE, D.DYN_VDP):
            for j in range(m):
                uc = u.astype(complex)
                uc[j] += 1j * h
                fc, _, _ = models.dynamics(d, x.astype(complex), uc)
                assert np.max(np.abs(np.imag(fc) / h - Fu[:, j])) <= 1e-12


def _idx(*v):
    out = np.zeros(D.NIDX, dtype=np.int64)
    out[:len(v)] = v
    return out


CASES = [
    (D.RT_STATE, _idx(2), 2),
    (D.RT_PR, _idx(0, 1, 8, 6), 5),
    (D.RT_PRR, _idx(0, 1, 2, 4, 5, 6, 7), 8),
    (D.RT_RANGE, _idx(0, 1, 2, 5, 6, 7, 3), 2),
    (D.RT_RANGE, _idx(0, 1, 0, 5, 6, 0, 2), 2),
    (D.RT_RANGE_PT, _idx(0, 1, 0, 0, 1, -1, 2), 5),
    (D.RT_RANGE_PT, _idx(0, 1, 2, -1, -1, -1, 3), 5),
    (D.RT_HEADING, _idx(0, 1, 0, 5, 6), 2),
    (D.RT_HEADING_PT, _idx(0, 1, 0, 1, 2), 5),
    (D.RT_HEADING_PT, _idx(0, 1, 0, -1, -1), 5),
]


@pytest.mark.parametrize("rt,idx,reclen", CASES)
def test_measurement_gradient_complex_step(rt, idx, reclen):
    rng = np.random.default_rng(2)
    n, ng = 10, 3
    for trial in range(4):
        x = rng.normal(size=n) * 10
        g = rng.normal(size=ng) * 10
        rec = rng.normal(size=reclen)
        if rt in (D.RT_PR, D.RT_PRR):
            rec[2:5] = rng.normal(size=3) * 2e7
        h0, gx, gg = models.measurement(rt, idx, x, g, rec)
        step = 1e-30
        full = np.zeros(n)
        for c, v in gx:
            full[c] += v
        fullg = np.zeros(ng)
        for c, v in gg:
            fullg[c] += v
        for j in range(n):
            xc = x.astype(complex)
            xc[j] += 1j * step
            if rt in (D.RT_HEADING, D.RT_HEADING_PT):
                continue        # arctan2 has no complex extension; covered by finite differences below
            hc, _, _ = models.measurement(rt, idx, xc, g.astype(complex), rec)
            assert abs(np.imag(hc) / step - full[j]) <= 1e-11 * (1 + abs(full[j]))
        for j in range(ng):
            if rt in (D.RT_HEADING, D.RT_HEADING_PT):
                continue
            gc = g.astype(complex)
            gc[j] += 1j * step
            hc, _, _ = models.measurement(rt, idx, x.astype(complex), gc, rec)
            assert abs(np.imag(hc) / step - fullg[j]) <= 1e-11 * (1 + abs(fullg[j]))
        if rt in (D.RT_HEADING, D.RT_HEADING_PT):
            e = 1e-6
            for j in range(n):
                xp, xm = x.copy(), x.copy()
                xp[j] += e
                xm[j] -= e
                fd = (models.measurement(rt, idx, xp, g, rec)[0] - models.measurement(rt, idx, xm, g, rec)[0]) / (2 * e)
                assert abs(fd - full[j]) <= 1e-7 * (1 + abs(full[j]))
            for j in range(ng):
                gp, gm = g.copy(), g.copy()
                gp[j] += e
                gm[j] -= e
                fd = (models.measurement(rt, idx, x, gp, rec)[0] - models.measurement(rt, idx, x, gm, rec)[0]) / (2 * e)
                assert abs(fd - fullg[j]) <= 1e-7 * (1 + abs(fullg[j]))


def test_sympy_van_der_pol_and_heading():
    sp = pytest.importorskip("sympy")
    x0, x1, u = sp.symbols("x0 x1 u")
    f = sp.Matrix([(1 - x1 ** 2) * x0 - x1 + u, x0])        # nlp/dynamics.py:61-66
    Jm = f.jacobian([x0, x1])
    d = _desc_for(D.DYN_VDP, 2, 1)
    pt = {x0: 0.3, x1: -1.2, u: 0.4}
    _, F, _ = models.dynamics(d, np.array([0.3, -1.2]), np.array([0.4]))
    assert np.allclose(np.array(Jm.subs(pt)).astype(float), F, atol=1e-14)
    xa, ya, xb, yb = sp.symbols("xa ya xb yb")
    h = sp.atan2(xb - xa + sp.Rational(1, 100000), yb - ya)     # nlp/measurements.py:30-36
    grads = [sp.diff(h, s) for s in (xa, ya, xb, yb)]
    vals = {xa: 1.0, ya: 2.0, xb: -3.0, yb: 0.5}
    x = np.zeros(10)
    x[[0, 1, 5, 6]] = [1.0, 2.0, -3.0, 0.5]
    _, gx, _ = models.measurement(D.RT_HEADING, _idx(0, 1, 0, 5, 6), x, np.zeros(0), np.zeros(2))
    assert np.allclose([float(gr.subs(vals)) for gr in grads], [v for _, v in gx], atol=1e-14)


def test_vectorised_residuals_match_scalar_definitions():
    """oracle/nls.meas_residuals (vectorised by type) against models.measurement row by row."""
    from tests import helpers
    d = shapes.two_receiver()
    bt = helpers.two_receiver_batch(d, 1, seed=9)
    rng = np.random.default_rng(3)
    z = bt["z0"][0] + rng.normal(size=d.nz)
    win = nls.Window(d, U=bt["U"][0], data=bt["data"][0], xprior=bt["xprior"][0])
    r, w, Jr = nls.meas_residuals(d, z, win)
    V = z.reshape(d.K, d.nv)
    Ve = d.Phi @ V
    for i in range(0, d.M, 7):
        rec = win.data[d.roff[i]:d.roff[i] + D.RT_RECLEN[int(d.rtype[i])]]
        if rec[1] == 0.0:
            assert r[i] == 0.0 and w[i] == 0.0 and not np.any(Jr[i])
            continue
        x = Ve[d.repoch[i]][d.cmap]
        h, gx, gg = models.measurement(int(d.rtype[i]), d.ridx[i], x, np.zeros(0), rec)
        assert abs(r[i] - (rec[0] - h)) <= 1e-9 * (1 + abs(rec[0]))
        row = np.zeros(d.nz)
        for c, v in gx:
            row[np.arange(d.K) * d.nv + d.cmap[c]] += -d.Phi[d.repoch[i]] * v
        assert np.max(np.abs(row - Jr[i])) <= 1e-12 * (1 + np.max(np.abs(row)))
