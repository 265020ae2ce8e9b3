"""GPU parity tests of the NLS path (kernels 1+2, 3 and the whole solve) against the CPU oracle
and the reference's stored IPOPT outputs.  All calls go through the C ABI (ctypes)."""
import numpy as np
import pytest

from nlp_filter_b200 import desc as D
from nlp_filter_b200 import shapes, synth
from tests import helpers

pytestmark = pytest.mark.gpu


def _engine():
    from nlp_filter_b200 import engine
    return engine


def _oracle():
    from oracle import nls
    return nls


def _rel(a, b):
    return np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-300)


def _check_assemble(d, batch, zs, mu=0.0, tol=1e-11):
    eng, nls = _engine(), _oracle()
    prob = eng.WindowProblem(d)
    B = zs.shape[0]
    H, g, cost = prob.assemble(zs, batch.get("U"), batch["data"], batch.get("xprior"), None, mu)
    H, g, cost = H.cpu().numpy(), g.cpu().numpy(), cost.cpu().numpy()
    for b in range(B):
        win = nls.Window(d, U=None if batch.get("U") is None else batch["U"][b], data=batch["data"][b],
                         xprior=None if batch.get("xprior") is None else batch["xprior"][b])
        J, merit, go, Ho = nls.evaluate(d, win, zs[b], mu)
        Ho = Ho + nls.second_order(d, win, zs[b], Ho)      # the kernel adds the exact second-order dynamics term
        assert abs(cost[b, 0] - J) <= tol * abs(J), (b, cost[b, 0], J)
        assert abs(cost[b, 1] - merit) <= tol * abs(merit) + 1e-12
        assert _rel(g[b], go) < tol, (b, _rel(g[b], go))
        assert _rel(H[b], Ho) < tol, (b, _rel(H[b], Ho))
        assert np.max(np.abs(H[b] - H[b].T)) == 0.0
    prob.close()


def test_assemble_vdp():
    d = shapes.vdp_mhe()
    bt = synth.vdp_mhe_batch(d, 8, seed=11)
    rng = np.random.default_rng(1)
    zs = bt["z0"] + 0.05 * rng.normal(size=bt["z0"].shape)
    _check_assemble(d, bt, zs)


def test_assemble_gnss():
    d = shapes.gnss()
    bt = synth.gnss_batch(d, 6, seed=12)
    rng = np.random.default_rng(2)
    zs = bt["z0"] + rng.normal(size=bt["z0"].shape)
    _check_assemble(d, bt, zs, tol=1e-10)


@pytest.mark.parametrize("loss", [D.LOSS_L2, D.LOSS_HUBER])
def test_assemble_car_barrier(loss):
    d = shapes.car(loss=loss)
    g = helpers.car_golden()
    B = 4
    U, data = zip(*[helpers.car_window_inputs(d, g, float(t0)) for t0 in range(B)])
    rng = np.random.default_rng(3)
    zs = np.zeros((B, d.nz))
    V = zs.reshape(B, d.K, d.nv)
    V[:, :, 3] = 10.0
    V[:, :, 7] = 200.0
    zs += 0.1 * rng.normal(size=zs.shape)
    bt = dict(U=np.array(U), data=np.array(data), xprior=np.tile(helpers.car_xhat0(g), (B, 1)))
    _check_assemble(d, bt, zs, mu=0.1, tol=1e-10)
    _check_assemble(d, bt, zs, mu=0.0, tol=1e-10)


def test_assemble_two_receiver_masked_slots_and_folded_equality():
    d = shapes.two_receiver()
    bt = helpers.two_receiver_batch(d, 4)
    rng = np.random.default_rng(4)
    zs = bt["z0"] + 0.5 * rng.normal(size=bt["z0"].shape)
    _check_assemble(d, bt, zs, tol=1e-10)


def test_factor_solve_matches_lapack():
    eng = _engine()
    d = shapes.gnss()
    prob = eng.WindowProblem(d)
    rng = np.random.default_rng(5)
    B, nz = 16, d.nz
    A = rng.normal(size=(B, nz, nz + 5))
    H = A @ A.transpose(0, 2, 1) + 1e-3 * np.eye(nz)
    g = rng.normal(size=(B, nz))
    for lam in (0.0, 1e-3, 10.0):
        dz, info = prob.factor_solve(H, g, lam)
        dz, info = dz.cpu().numpy(), info.cpu().numpy()
        assert np.all(info == 0)
        for b in range(B):
            dg = np.diag(H[b]) + 1e-14 * np.max(np.diag(H[b]))
            ref = -np.linalg.solve(H[b] + lam * np.diag(dg), g[b])
            assert _rel(dz[b], ref) < 1e-9
    # an indefinite matrix must be reported, not factorised
    Hbad = H.copy()
    Hbad[3] = -Hbad[3]
    dz, info = prob.factor_solve(Hbad, g, 0.0)
    info = info.cpu().numpy()
    assert info[3] == 1 and info[2] == 0
    prob.close()


def _solve_both(d, bt, B, opts=None, tol=1e-6):
    eng, nls = _engine(), _oracle()
    prob = eng.WindowProblem(d)
    z, J, iters, status = eng.solve_windows_host(prob, bt["z0"][:B], None if bt.get("U") is None else bt["U"][:B],
                                                 bt["data"][:B], None if bt.get("xprior") is None else bt["xprior"][:B], opts)
    assert np.all(status <= 1), status
    worst = 0.0
    for b in range(B):
        win = nls.Window(d, U=None if bt.get("U") is None else bt["U"][b], data=bt["data"][b],
                         xprior=None if bt.get("xprior") is None else bt["xprior"][b])
        zo, info = nls.solve(d, win, bt["z0"][b], opts)
        assert info["status"] <= 1
        err = np.max(np.abs(z[b] - zo) / (1.0 + np.abs(zo)))
        worst = max(worst, err)
        assert err < tol, (b, err, info, iters[b])
        assert abs(J[b] - info["J"]) <= 1e-9 * abs(info["J"])
    prob.close()
    return worst


def test_solve_vdp_matches_oracle():
    d = shapes.vdp_mhe()
    bt = synth.vdp_mhe_batch(d, 24, seed=21)
    _solve_both(d, bt, 24)


def test_solve_vdp_epoch_dependent_and_missing_measurements():
    """Weights that change from epoch to epoch (and weight 0 = missing measurement) take the summed form of the
    measurement Hessian in the warp kernel; constant weights take its precomputed sum over the epochs."""
    d = shapes.vdp_mhe()
    bt = synth.vdp_mhe_batch(d, 12, seed=22)
    rng = np.random.default_rng(3)
    w = bt["data"].reshape(12, -1, 2)
    w[:, :, 1] *= rng.uniform(0.25, 4.0, size=w.shape[:2])
    w[:, :, 1][rng.uniform(size=w.shape[:2]) < 0.1] = 0.0
    w[0, :, 1] = bt["data"].reshape(12, -1, 2)[0, 0, 1]            # one window back to constant weights
    _solve_both(d, bt, 12)


def test_solve_gnss_matches_oracle_1mm():
    d = shapes.gnss()
    bt = synth.gnss_batch(d, 12, seed=22)
    eng, nls = _engine(), _oracle()
    prob = eng.WindowProblem(d)
    z, J, iters, status = eng.solve_windows_host(prob, bt["z0"], bt["U"], bt["data"])
    for b in range(12):
        zo, info = nls.solve(d, nls.Window(d, U=bt["U"][b], data=bt["data"][b]), bt["z0"][b])
        V, Vo = z[b].reshape(d.K, 5), zo.reshape(d.K, 5)
        assert np.max(np.abs(V[:, :3] - Vo[:, :3])) < 1e-3          # GNSS positions within 1 mm
        assert np.max(np.abs(V - Vo) / (1 + np.abs(Vo))) < 1e-6
    prob.close()


def test_solve_two_receiver_matches_oracle():
    d = shapes.two_receiver()
    bt = helpers.two_receiver_batch(d, 6)
    _solve_both(d, bt, 6)


@pytest.mark.parametrize("loss,key", [(D.LOSS_L2, "nlp_l2"), (D.LOSS_HUBER, "nlp_huber")])
def test_car_chain_reproduces_stored_ipopt_outputs(loss, key):
    """Receding-horizon chain of autonomous-car.py:232-288, all 90 windows the reference stored, against its
    IPOPT results: every component within 0.2 mm (L2) / 0.3 mm (pseudo-Huber); the CPU oracle's own distance
    to the pickles is 7.7e-5 / 1.8e-4 (tests/test_oracle_car_golden.py)."""
    eng = _engine()
    d = shapes.car(loss=loss)
    g = helpers.car_golden()
    prob = eng.WindowProblem(d)
    gold = g[key]
    ts = np.linspace(0, 2.0, 10)
    xhat0 = helpers.car_xhat0(g)
    z = np.zeros((1, d.nz))
    PhiDT = d.cpm.basisMatrix([1.0])
    worst = np.zeros(7)
    for step in range(90):
        U, data = helpers.car_window_inputs(d, g, float(step))
        z, J, iters, status = eng.solve_windows_host(prob, z, U[None], data[None], xhat0[None])
        assert status[0] <= 1, (step, status)
        est = helpers.car_extract(d, z[0], ts)
        worst = np.maximum(worst, np.max(np.abs(est - gold[step * 10:(step + 1) * 10]), 0))
        xhat0 = (PhiDT @ z[0].reshape(d.K, d.nv))[0]
    assert np.all(worst[:5] < 2e-4), worst          # x, y, z, clock bias [m], psi [rad]
    assert np.all(worst[5:] < (2e-4 if loss == D.LOSS_L2 else 3e-4)), worst      # vx, vy [m/s]
    prob.close()


def test_sample_matches_basis_product():
    eng = _engine()
    d = shapes.gnss()
    prob = eng.WindowProblem(d)
    rng = np.random.default_rng(7)
    z = rng.normal(size=(5, d.nz))
    ts = np.linspace(0, d.T, 17)
    Ph = d.cpm.basisMatrix(ts)
    out = prob.sample(z, Ph).cpu().numpy()
    ref = np.einsum("tk,bkn->btn", Ph, z.reshape(5, d.K, d.nv))
    assert np.max(np.abs(out - ref)) < 1e-12 * np.max(np.abs(ref))
    prob.close()


def test_empty_measurement_set_and_error_paths():
    eng = _engine()
    from nlp_filter_b200 import _capi
    d = shapes.vdp_mhe()
    d.set_residuals([], [], np.zeros((0, D.NIDX), dtype=np.int32))
    d.E = 0
    d.Phi = np.zeros((0, d.K))
    prob = eng.WindowProblem(d)          # dynamics cost only: still a valid problem
    bt = dict(z0=np.ones((2, d.nz)), U=np.zeros((2, d.K, 1)), data=np.zeros((2, 1)))
    _, _, c0 = prob.assemble(bt["z0"], bt["U"], bt["data"])
    z, J, iters, status = eng.solve_windows_host(prob, bt["z0"], bt["U"], bt["data"])
    assert np.all(np.isfinite(z)) and np.all(J < 1e-3 * c0.cpu().numpy()[:, 0])
    prob.close()
    bad = shapes.vdp_mhe()
    bad.resize_nv(30)          # nz = 630 exceeds the shared-memory-resident limit
    with pytest.raises(_capi.NlpfError):
        eng.WindowProblem(bad)


@pytest.mark.parametrize("shape", ["vdp", "gnss"])
def test_bad_window_is_reported_and_does_not_disturb_its_neighbours(shape):
    """A window with a non-finite measurement gets status NUMERIC (3); the other windows of the batch -- solved
    by neighbouring warps / CTAs of the same launch -- still match a run without it."""
    eng = _engine()
    if shape == "vdp":
        d = shapes.vdp_mhe()
        bt = synth.vdp_mhe_batch(d, 40, seed=31)
    else:
        d = shapes.gnss()
        bt = synth.gnss_batch(d, 40, seed=32)
    prob = eng.WindowProblem(d)
    ref = eng.solve_windows_host(prob, bt["z0"], bt.get("U"), bt["data"], bt.get("xprior"))
    data = bt["data"].copy()
    data[7, 0] = np.nan
    data[23, 0] = np.inf
    z, J, iters, status = eng.solve_windows_host(prob, bt["z0"], bt.get("U"), data, bt.get("xprior"))
    assert status[7] == 3 and status[23] == 3
    ok = np.ones(40, dtype=bool)
    ok[[7, 23]] = False
    assert np.all(status[ok] <= 1)
    assert np.array_equal(z[ok], ref[0][ok]) and np.array_equal(iters[ok], ref[2][ok])
    assert np.array_equal(z[7], bt["z0"].reshape(40, -1)[7])          # the start point is returned untouched
    prob.close()


@pytest.mark.parametrize("sched", [dict(), dict(lam0=1e-3, lam_jump=1e-30, lam_jump_indef=1e-30, lam_dec_first=10.0),
                                   dict(lam0=1e-6, lam_jump=1e-2, lam_jump_indef=1.0, lam_dec_first=100.0)])
def test_damping_schedule_options_reach_both_kernels_and_the_oracle(sched):
    """nlpf_opts.lam0 / lam_jump / lam_jump_indef / lam_dec_first (0 = the NLPF_LAM_* defaults; the second set is the
    round-1 schedule): the warp-per-window kernel (van der Pol batch, far starts: rejected steps and failed pivots
    happen) and the CTA-per-window kernel (RC-car batch) walk the same damping path as the oracle -- same iteration
    counts on nearly every window, same minimiser on all of them."""
    eng, nls = _engine(), _oracle()
    opts = D.SolveOpts(**sched)
    for d, bt, n_check in ((shapes.vdp_mhe(), synth.vdp_mhe_batch(shapes.vdp_mhe(), 64, seed=31), 24),
                           (shapes.rc(), synth.rc_batch(shapes.rc(), 6, seed=32), 3)):
        prob = eng.WindowProblem(d)
        z, J, it, st = eng.solve_windows_host(prob, bt["z0"], bt.get("U"), bt["data"], bt.get("xprior"), opts=opts)
        assert np.all(st <= 1)
        same = 0
        for b in range(n_check):
            win = nls.Window(d, U=None if bt.get("U") is None else bt["U"][b], data=bt["data"][b],
                             xprior=None if bt.get("xprior") is None else bt["xprior"][b])
            zo, info = nls.solve(d, win, bt["z0"][b], opts)
            assert info["status"] <= 1
            dz = (np.abs(z[b] - zo) / (1 + np.abs(zo))).reshape(d.K, d.nv)
            if d.n == 6:
                # RC-car model (nlp/dynamics.py:117-136): x5' = (v / 0.28) tan(delta) does not depend on the state and x5
                # is neither measured nor used by another component -- its level is a null direction of the problem
                # (only the regularisation fixes it); every other component is compared
                dz = np.delete(dz, 5, axis=1)
            assert np.max(dz) < 1e-6, (b, sched, np.max(dz, 0))
            same += int(it[b] == info["iters"])
        assert same >= n_check - max(1, n_check // 8), (same, n_check, sched)
        prob.close()
