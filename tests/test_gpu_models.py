"""Parity of every kernel dynamics model and residual type against the oracle: assembled H, g, J on
random points (kernels 1+2) and the converged minimiser of small problems (whole solve)."""
import numpy as np
import pytest

from nlp_filter_b200 import desc as D
from nlp_filter_b200 import shapes

pytestmark = pytest.mark.gpu

MODELS = [(D.DYN_AFFINE, 4, 2), (D.DYN_VDP, 2, 1), (D.DYN_BICYCLE, 6, 2), (D.DYN_VEHICLE, 6, 2),
          (D.DYN_VEHICLE_GNSS, 9, 2), (D.DYN_QUADCOPTER, 12, 4)]


def _random_window(d, rng, scale=1.0):
    B = 3
    z = rng.normal(size=(B, d.nz)) * scale
    if d.dyn_id in (D.DYN_VEHICLE, D.DYN_VEHICLE_GNSS):
        z.reshape(B, d.K, d.nv)[:, :, 3] = 6.0 + rng.uniform(size=(B, d.K))       # vx away from the tyre singularity
    U = rng.normal(size=(B, d.K, max(d.m, 1))) * 0.3
    data = np.zeros((B, d.Ldata))
    for r in range(d.M):
        o = d.roff[r]
        n = D.RT_RECLEN[int(d.rtype[r])]
        data[:, o] = rng.normal(size=B)
        data[:, o + 1] = rng.uniform(0.5, 2.0, B)
        if n > 2:
            data[:, o + 2:o + n] = rng.normal(size=(B, n - 2)) * (2e7 if d.rtype[r] in (D.RT_PR, D.RT_PRR) else 5.0)
        if d.rtype[r] == D.RT_PRR:
            data[:, o + 5:o + 8] = rng.normal(size=(B, 3)) * 3e3
    return z, U, data


def _compare(d, z, U, data, loss_tol=1e-10):
    from nlp_filter_b200 import engine
    from oracle import nls
    prob = engine.WindowProblem(d)
    H, g, cost = prob.assemble(z, U if d.m > 0 else None, data)
    H, g, cost = H.cpu().numpy(), g.cpu().numpy(), cost.cpu().numpy()
    for b in range(z.shape[0]):
        win = nls.Window(d, U=U[b] if d.m > 0 else None, data=data[b])
        J, merit, go, Ho = nls.evaluate(d, win, z[b])
        Ho = Ho + nls.second_order(d, win, z[b], Ho)
        assert abs(cost[b, 0] - J) <= loss_tol * abs(J)
        assert np.max(np.abs(g[b] - go)) <= loss_tol * np.max(np.abs(go))
        assert np.max(np.abs(H[b] - Ho)) <= loss_tol * np.max(np.abs(Ho))
    prob.close()


@pytest.mark.parametrize("dyn_id,n,m", MODELS)
@pytest.mark.parametrize("loss", [D.LOSS_L2, D.LOSS_HUBER])
def test_dynamics_models_assemble(dyn_id, n, m, loss):
    d = shapes.generic_model(dyn_id, n, m)
    d.loss = loss
    d.delta = 0.7
    rng = np.random.default_rng(10 * dyn_id + loss)
    z, U, data = _random_window(d, rng, scale=0.5)
    _compare(d, z, U, data)


def test_residual_types_assemble():
    """Every residual type in one problem: state, pseudorange, pseudorange rate, range / heading between
    components, range / heading to a data point."""
    extra = [(D.RT_PR, (0, 1, 2, 3)), (D.RT_PRR, (0, 1, 2, 4, 5, 6, 7)), (D.RT_RANGE, (0, 1, 2, 5, 6, 7, 3)),
             (D.RT_RANGE, (0, 1, 0, 5, 6, 0, 2)), (D.RT_HEADING, (0, 1, 0, 5, 6)),
             (D.RT_RANGE_PT, (0, 1, 2, -1, -1, -1, 3)), (D.RT_HEADING_PT, (5, 6, 0, -1, -1))]
    d = shapes.generic_model(D.DYN_AFFINE, 8, 0, extra=extra)
    rng = np.random.default_rng(42)
    z, U, data = _random_window(d, rng, scale=3.0)
    _compare(d, z, U, data)


def test_multi_receiver_shape_with_global_unknowns():
    """multi-receiver.py:62-99: pseudorange + rate + range to the global unknown XA; assemble parity and
    whole-solve parity from a start near the truth (the zero start is multi-modal)."""
    from nlp_filter_b200 import engine
    from oracle import nls
    d = shapes.multi_receiver(n_epochs=24, n_sat=8, T=23.0)
    rng = np.random.default_rng(8)
    B = 2
    pos = np.array([-545.0, -53.0, 75.0]) + rng.normal(0, 1, (B, 3))
    vel = rng.normal(0, 0.3, (B, 3))
    xa = pos[:, :2] + np.array([1.7, 1.7])
    el = np.arcsin(rng.uniform(np.sin(np.deg2rad(15)), 1.0, (B, 8)))
    az = rng.uniform(0, 2 * np.pi, (B, 8))
    rad = rng.uniform(2.0e7, 2.6e7, (B, 8))
    sat = np.stack([rad * np.cos(el) * np.sin(az), rad * np.cos(el) * np.cos(az), rad * np.sin(el)], 2)
    sv = rng.normal(0, 2e3, (B, 8, 3))
    data = np.zeros((B, d.Ldata))
    r = 0
    for e in range(d.E):
        t = d.epochs[e]
        p = pos + vel * t
        for j in range(8):
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
    zr = z0 + rng.normal(0, 0.5, z0.shape)
    _compare(d, zr, np.zeros((B, d.K, 1)), data, loss_tol=1e-9)
    prob = engine.WindowProblem(d)
    # This class has a nearly flat valley (multi-receiver.py needs 538 iterations; here ~200): both runs end
    # at the rounding floor of the objective (pseudoranges of 2e7 m squared: 4e-9 absolute on J = 49), and along
    # the valley a relative change of 1e-10 in J is a relative change of 1e-5 in z.  So the objectives must
    # agree to the rounding bound, the estimates to what that bound can resolve.
    z, J, iters, status = engine.solve_windows_host(prob, z0, None, data)
    assert np.all(status <= 1)
    for b in range(B):
        zo, info = nls.solve(d, nls.Window(d, data=data[b]), z0[b])
        assert abs(J[b] - info["J"]) <= 1e-9 * abs(info["J"])
        assert np.max(np.abs(z[b] - zo) / (1 + np.abs(zo))) < 1e-4
        assert z[b][-1] == 0.0                     # XA[2] never enters the cost and must not move
    prob.close()


@pytest.mark.parametrize("dyn_id,n,m", [(D.DYN_BICYCLE, 6, 2), (D.DYN_VEHICLE, 6, 2), (D.DYN_QUADCOPTER, 12, 4)])
def test_small_problem_solves_match_oracle(dyn_id, n, m):
    from nlp_filter_b200 import engine
    from oracle import nls
    d = shapes.generic_model(dyn_id, n, m)
    rng = np.random.default_rng(dyn_id)
    z, U, data = _random_window(d, rng, scale=0.2)
    if dyn_id == D.DYN_QUADCOPTER:                 # near-hover attitude measurements, loose process model
        d.qinv = d.qinv * 0.02
        for r in range(d.M):
            data[:, d.roff[r]] *= 0.15
    if dyn_id == D.DYN_VEHICLE:                    # measurements consistent with a forward-moving car
        for r in range(d.M):
            if d.ridx[r, 0] == 3:
                data[:, d.roff[r]] = 6.0
    prob = engine.WindowProblem(d)
    zs, J, iters, status = engine.solve_windows_host(prob, z, U, data)
    assert np.all(status <= 1)
    for b in range(z.shape[0]):
        zo, info = nls.solve(d, nls.Window(d, U=U[b], data=data[b]), z[b])
        assert abs(J[b] - info["J"]) <= 1e-8 * abs(info["J"])
        assert np.max(np.abs(zs[b] - zo) / (1 + np.abs(zo))) < 1e-6
    prob.close()
