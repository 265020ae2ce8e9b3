"""Pins oracle/nls.py to the reference's stored IPOPT outputs for the vehicle MHE chain
(data/autonomous-car/filtering/nlp-l2.pkl, nlp-huber.pkl via tests/golden/car.npz)."""
import numpy as np
import pytest

from nlp_filter_b200 import desc as D
from nlp_filter_b200 import shapes
from oracle import nls
from tests import helpers


def _chain(loss, key, nwin):
    d = shapes.car(loss=loss)
    g = helpers.car_golden()
    gold = g[key]
    ts = np.linspace(0, 2.0, 10)
    PhiDT = d.cpm.basisMatrix([1.0])
    xhat0 = helpers.car_xhat0(g)
    z = np.zeros(d.nz)                       # CasADi's default initial guess (no initializeEstimate call)
    worst = np.zeros(7)
    Js = []
    for step in range(nwin):
        U, data = helpers.car_window_inputs(d, g, float(step))
        z, info = nls.solve(d, nls.Window(d, U=U, data=data, xprior=xhat0), z)
        assert info["status"] <= 1
        Js.append(info["J"])
        worst = np.maximum(worst, np.max(np.abs(helpers.car_extract(d, z, ts) - gold[step * 10:(step + 1) * 10]), 0))
        xhat0 = (PhiDT @ z.reshape(d.K, d.nv))[0]
    return worst, Js


def test_l2_chain_all_90_windows_within_a_fifth_of_a_millimetre_of_ipopt():
    """Every window the reference stored (90 chained solves, 10 samples x 7 components each).  Measured:
    worst 7.7e-5 over all components (x, y, z, clock bias [m], psi [rad], vx, vy [m/s])."""
    worst, Js = _chain(D.LOSS_L2, "nlp_l2", 90)
    assert np.all(worst < 2e-4), worst
    assert abs(Js[0] - 22.0979603177) < 1e-6


def test_huber_chain_all_90_windows_within_a_third_of_a_millimetre_of_ipopt():
    """Measured: positions / clock bias <= 9.3e-5 m, psi 2.4e-5 rad, vx 1.8e-4, vy 1.5e-4 m/s."""
    worst, Js = _chain(D.LOSS_HUBER, "nlp_huber", 90)
    assert np.all(worst[:5] < 2e-4) and np.all(worst[5:] < 3e-4), worst
    assert abs(Js[0] - 22.09346978) < 1e-5            # a different optimum from the L2 one (6 cm away)


def test_reference_weights_matter():
    """With true Clenshaw-Curtis weights the optimum moves by millimetres: the
    golden test above really discriminates the reference's weight formula."""
    d = shapes.car()
    g = helpers.car_golden()
    U, data = helpers.car_window_inputs(d, g, 0.0)
    z, _ = nls.solve(d, nls.Window(d, U=U, data=data, xprior=helpers.car_xhat0(g)), np.zeros(d.nz))
    N = d.K - 1
    th = np.pi * np.arange(N + 1) / N
    cc = np.zeros(N + 1)                      # Clenshaw-Curtis by moment fitting
    for k in range(N + 1):
        s = sum((2.0 / (1 - j * j)) * np.cos(j * th[k]) * (0.5 if j in (0, N) else 1.0) for j in range(0, N + 1, 2))
        cc[k] = (2.0 / N) * s * (0.5 if k in (0, N) else 1.0)
    assert abs(cc.sum() - 2.0) < 1e-12
    d2 = shapes.car()
    d2.cw = (d2.T / 2.0) * cc[::-1]
    z2, _ = nls.solve(d2, nls.Window(d2, U=U, data=data, xprior=helpers.car_xhat0(g)), np.zeros(d.nz))
    moved = np.max(np.abs(z2.reshape(6, 9)[:, :2] - z.reshape(6, 9)[:, :2]))
    assert moved > 3e-4
