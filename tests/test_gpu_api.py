"""The reference's Python API (nlp.nlp, utils.ekf, utils.leastsquares) driven the way its scripts
drive it, on committed fixtures, checked against the reference's stored outputs / the survey's
independently derived optima.  Everything below the API runs in libnlpfilter.so."""
import os

import numpy as np
import pytest

from tests import helpers
from tests.helpers import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _names():
    import nlp_filter_b200
    nlp_filter_b200.install_reference_names()


def test_single_integrator_ocp():
    """single_integrator.py:7-34 -> J* = 25.691769510223."""
    import nlp.nlp as nlp
    import nlp.dynamics as dynamics
    import nlp.cost_functions as cost_functions
    problem = nlp.fixedTimeOptimalControlNLP(10, 10., 2, 2)
    X = problem.addVariables(11, 2, name='x')
    U = problem.addVariables(11, 2, name='u')
    problem.addDynamics(dynamics.single_integrator_2D, X, U)
    problem.addStageCost(cost_functions.single_integrator, X, U)
    problem.addInitialCondition(X[0], [-3, 4])
    problem.build()
    problem.solve()
    t = np.linspace(0, 10., 20)
    x, u = problem.extractSolution('x', t), problem.extractSolution('u', t)
    assert abs(problem.solver["objective"] - 25.691769510223) < 1e-7
    assert np.allclose(x[[0, 1, 10, 19]], [[-3, 4], [-1.80564297, 2.40752396], [-0.00082742, 0.00110322],
                                           [-0.00247097, 0.00329462]], atol=2e-7)
    assert np.allclose(u[[0, 1, 10, 19]], [[3.11069858, -4.1475981], [1.68189997, -2.2425333], [0.01061655, -0.0141554],
                                           [-0.15455658, 0.20607544]], atol=2e-7)


def test_van_der_pol_ocp_with_bounds():
    """van_der_pol.py:10-38 -> J* = 2.8784425681, bounds inactive."""
    import nlp.nlp as nlp
    import nlp.dynamics as dynamics
    import nlp.cost_functions as cost_functions
    problem = nlp.fixedTimeOptimalControlNLP(20, 10., 2, 1)
    X = problem.addVariables(21, 2, [-np.inf, -0.25], [np.inf, np.inf], name='x')
    U = problem.addVariables(21, 1, [-1.0], [1.0], name='u')
    problem.addDynamics(dynamics.van_der_pol, X, U)
    problem.addStageCost(cost_functions.van_der_pol, X, U)
    problem.addInitialCondition(X[0], [0, 1])
    problem.build()
    problem.solve()
    t = np.linspace(0, 10., 20)
    x, u = problem.extractSolution('x', t), problem.extractSolution('u', t)
    assert abs(problem.solver["objective"] - 2.8784425681) < 2e-7
    assert np.allclose(x[[0, 5, 10, 19]], [[0, 1], [-0.159215, 0.043698], [0.017763, -0.015819], [-0.001766, -0.000131]],
                       atol=2e-6)
    assert np.allclose(u[[0, 5, 10, 19], 0], [-0.460142, 0.413541, -0.04135, -0.001701], atol=2e-6)


def test_gnss_stationary_flow():
    """gnss_stationary.py:39-141 on the stored log: LS chain, batch LS, EKF, MHE."""
    import nlp.nlp as nlp
    import nlp.dynamics as dynamics
    import nlp.cost_functions as cost_functions
    import nlp.measurements as measurements
    import utils.gnss as gnss
    import utils.leastsquares as ls
    import utils.utils as utils
    import utils.ekf as ekf
    G = np.load(os.path.join(GOLDEN, "ekf_ls.npz"))
    p_ref = utils.lla2ecef(np.array([37.4276, -122.1670, 0]))
    T = 50
    t = np.linspace(0, T, T + 1)
    sat_pos = [G["gs_sat_ecef"][k, :G["gs_nsat"][k]] for k in range(T + 1)]
    pr = [G["gs_pr"][k, :G["gs_nsat"][k]] for k in range(T + 1)]
    ls.iterativeLeastSquares.__defaults__[0][:] = 0.0
    LS = np.zeros((T + 1, 4))
    for k in range(T + 1):
        p, b = ls.iterativeLeastSquares(sat_pos[k], pr[k])
        LS[k, :3], LS[k, 3] = utils.ecef2enu(p, p_ref), b
    assert np.max(np.abs(LS[:, :3] - G["gs_ls_enu"])) < 1e-5 and np.max(np.abs(LS[:, 3] - G["gs_ls_ecef"][:, 3])) < 1e-5
    sol = ls.runBatchLeastSquares(t, sat_pos, pr, p_ref)
    assert np.allclose([sol["x_ENU"], sol["y_ENU"], sol["z_ENU"]], [649.8341849429796, 230.12118136513982, 5.293044609244814],
                       atol=1e-5)
    # EKF (gnss_stationary.py:76-98)
    Q = np.diag([0.0001, 0.0001, 0.0001, 0.1, 0.001])
    rate = (LS[-1, 3] - LS[0, 3]) / T
    xhat0 = np.array([LS[0, 0], LS[0, 1], LS[0, 2], LS[0, 3], rate])
    f = ekf.EKF(gnss.gnss_pos_and_bias, gnss.multi_pseudorange, xhat0, np.diag([1, 1, 1, 1, 1]))
    u = np.zeros((3, T + 1))
    EK = np.zeros((T + 1, 4))
    sat_enu = []
    for k in range(T + 1):
        EK[k] = f.mu[:4]
        sk = np.array([utils.ecef2enu(s, p_ref) for s in sat_pos[k]])
        sat_enu.append(sk)
        f.update(u[:, k], pr[k], Q, np.diag(100 * np.ones(pr[k].shape[0])), dyn_func_params={"dt": 1},
                 meas_func_params={"sat_pos": sk})
    assert np.max(np.abs(f.mu - G["gs_ekf_traj"][-1]) / (1 + np.abs(G["gs_ekf_traj"][-1]))) < 1e-8
    # MHE (gnss_stationary.py:100-141)
    problem = nlp.fixedTimeOptimalEstimationNLP(10, T, 5, 3)
    X = problem.addVariables(11, 5, name='x')
    xh = np.vstack((EK[:, 0], EK[:, 1], EK[:, 2], EK[:, 3], rate * np.ones(T + 1)))
    problem.initializeEstimate(X, t, xh)
    U, W = problem.addDynamics(dynamics.gnss_pos_and_bias, X, t, u, )
    problem.addDynamicsCost(cost_functions.weighted_l2_norm, W, {"Q": np.linalg.inv(Q)})
    for k in range(T + 1):
        for (i, prr) in enumerate(pr[k]):
            problem.addResidualCost(measurements.pseudorange, X, np.array([[float(k)]]), np.array([[prr]]),
                                    np.linalg.inv(np.diag([100])), {"sat_pos": sat_enu[k][i]})
    problem.build()
    problem.solve()
    problem.solve(warmstart=True)
    x_opt = problem.extractSolution('x', t)
    assert abs(problem.solver["objective"] - 890.56824012) < 1e-5
    assert np.allclose(x_opt[0], [649.847214, 230.104855, 5.375491, 7.418794, -193.682200], atol=1e-4)
    assert np.allclose(x_opt[25], [649.847906, 230.100276, 5.377270, -4836.340964, -193.619190], atol=1e-4)
    assert np.allclose(x_opt[50], [649.846741, 230.098330, 5.378006, -9672.679799, -193.477586], atol=1e-4)
    w0 = problem.extractVariableValue('w', 3)
    assert w0 is not None and w0.shape == (5,)
    assert problem.extractVariableValue('nope', 0) is None


def test_estimation_example_flow():
    """estimation_example.py:10-45 with a seeded noise draw; compared with the CPU oracle."""
    import nlp.nlp as nlp
    import nlp.dynamics as dynamics
    import nlp.cost_functions as cost_functions
    import nlp.measurements as measurements
    import utils.simulate as simulate
    from oracle import nls
    np.random.seed(3)
    T = 10.
    t = np.linspace(0, T, 50)
    u = np.vstack((np.sin(t), np.cos(t)))
    x = simulate.open_loop_sim(t, u, np.zeros(2), dynamics.single_integrator_2D)
    R = np.diag([0.01, 0.02])
    y = simulate.generate_measurements(x, measurements.full_state, R)
    problem = nlp.fixedTimeOptimalEstimationNLP(20, T, 2, 2)
    X = problem.addVariables(21, 2, name='x')
    _, W = problem.addDynamics(dynamics.single_integrator_2D, X, t, u)
    problem.addDynamicsCost(cost_functions.weighted_l2_norm, W, {"Q": np.linalg.inv(np.diag([0.0001, 0.0001]))})
    problem.addResidualCost(measurements.full_state, X, t, y, np.linalg.inv(R))
    problem.build()
    problem.solve()
    problem.solve(warmstart=True)
    xo = problem.extractSolution('x', t)
    assert np.max(np.abs(xo.T - x)) < 0.2
    d = problem._desc
    z, info = nls.solve(d, nls.Window(d, U=np.array([p.value for p in problem._dyn[2]]), data=problem._pack_data()),
                        np.zeros(d.nz))
    assert np.max(np.abs(z - problem._z_last) / (1 + np.abs(z))) < 1e-6


@pytest.mark.parametrize("huber,key", [(False, "nlp_l2"), (True, "nlp_huber")])
def test_autonomous_car_flow_against_stored_ipopt(huber, key):
    """autonomous-car.py:184-288 (and :293-388 for the pseudo-Huber loss) through the API:
    parameterised slots, masked satellites, bounds, arrival cost, warm-started chain."""
    import nlp.nlp as nlp
    import nlp.dynamics as dynamics
    import nlp.cost_functions as cost_functions
    import nlp.measurements as measurements
    import utils.utils as utils
    import utils.vehicle_sim as vehicle_sim
    g = helpers.car_golden()
    car_params = vehicle_sim.get_parameters()
    Q_NLP = np.diag([0.01, 0.01, 0.01, 100, 500, 500, .001, .001, .001])
    P_NLP = np.diag(1 * np.ones(9))
    T, N, n, m, N_sat, dt_gnss = 2, 5, 9, 2, 12, 0.1          # one slot more than satellites: last slot masked
    problem = nlp.fixedTimeOptimalEstimationNLP(N, T, n, m)
    X = problem.addVariables(N + 1, n, name='x')
    U, W = problem.addDynamics(dynamics.vehicle_dynamics_and_gnss, X, None, None, {"car_params": car_params})
    if huber:
        problem.addDynamicsCost(cost_functions.pseudo_huber_loss, W, {"Q": np.linalg.inv(Q_NLP), "delta": 5.0})
    else:
        problem.addDynamicsCost(cost_functions.weighted_l2_norm, W, {"Q": np.linalg.inv(Q_NLP)})
    problem.addVarBounds(X, 2, -np.pi, np.pi)
    problem.addVarBounds(X, 3, 0, np.inf)
    X0 = problem.addInitialCost(cost_functions.weighted_l2_norm, X[0], {"Q": np.linalg.inv(P_NLP)})
    N_gnss = int(np.floor(T / dt_gnss))
    t_gnss = np.linspace(0, T, N_gnss + 1)
    Y, R, sat_pos = [], [], []
    for i in range(N_gnss + 1):
        t_i = np.array([[t_gnss[i]]])
        Yi, Ri, Si = [], [], []
        for j in range(N_sat):
            s_ij = problem.addParameter(1, 3)[0]
            R_ij = problem.addParameter(1, 1)[0]
            Y_ij = problem.addResidualCost(measurements.vehicle_pseudorange, X, t_i, None, R_ij, {"p": 1, "sat_pos": s_ij})[0]
            Yi.append(Y_ij); Ri.append(R_ij); Si.append(s_ij)
        Y.append(Yi); R.append(Ri); sat_pos.append(Si)
    problem.build()
    xhat0 = helpers.car_xhat0(g)
    gold = g[key]
    worst = np.zeros(7)
    tsolve = []
    for step in range(12):
        t0 = float(step)
        ti = utils.get_time_indices(g["traj_t"], t0, t0 + T)
        gi = utils.get_time_indices(g["gnss_t"], t0, t0 + T)
        problem.setControl(U, g["traj_t"][ti] - t0, g["traj_u"][:, ti])
        problem.setParameter(X0, xhat0)
        for i in range(N_gnss + 1):
            t_i = np.array([[t_gnss[i]]])
            for j in range(N_sat):
                if j < 11:
                    problem.setParameter(R[i][j], dt_gnss * np.linalg.inv(np.diag([float(g["r_pr"])])))
                    problem.setParameter(sat_pos[i][j], g["sat_enu"][j])
                    problem.setMeasurement(Y[i][j], t_i, np.array([[g["pr"][gi[i], j]]]))
                else:
                    problem.setParameter(R[i][j], 0.0)
                    problem.setParameter(sat_pos[i][j], np.zeros(3))
                    problem.setMeasurement(Y[i][j], t_i, np.array([[0.0]]))
        problem.solve(warmstart=True)
        tt = np.linspace(0, T, 10)
        x_opt = problem.extractSolution('x', tt)
        xhat0 = problem.extractSolution('x', [1])
        est = np.stack([x_opt[:, 0], x_opt[:, 1], x_opt[:, 8], x_opt[:, 6], x_opt[:, 2], x_opt[:, 3], x_opt[:, 4]], 1)
        worst = np.maximum(worst, np.max(np.abs(est - gold[step * 10:(step + 1) * 10]), 0))
        tsolve.append(problem.solver["t_wall_total"])
    assert np.all(worst[:4] < 1e-3), worst
    assert worst[4] < 1e-3 and np.all(worst[5:] < 5e-3), worst
    assert len(tsolve) == 12 and all(ts > 0 for ts in tsolve)


def test_two_receiver_flow_with_equality_constraint():
    """gnss-multi-receiver.py:51-131 structure: range + heading + two receivers' pseudoranges,
    zA == zB through addEqConstraint; compared with the CPU oracle on the lowered problem."""
    import nlp.nlp as nlp
    import nlp.dynamics as dynamics
    import nlp.cost_functions as cost_functions
    import nlp.constraints as constraints
    import nlp.measurements as measurements
    from nlp_filter_b200 import shapes
    from oracle import nls
    Q = np.diag([.01, .01, .01, 0.01, 0.01, .01, .01, .01, 0.01, 0.01])
    P = 0.01 * np.diag([1, 1, 1, 0.1, 0.1, 1, 1, 1, 0.1, 0.1])
    T, N, n, m = 5, 10, 10, 6
    problem = nlp.fixedTimeOptimalEstimationNLP(N, T, n, m)
    X = problem.addVariables(N + 1, n, name='x')
    U, W = problem.addDynamics(dynamics.gnss_two_receiver, X, None, None)
    problem.addDynamicsCost(cost_functions.weighted_l2_norm, W, {"Q": np.linalg.inv(Q)})
    X0 = problem.addInitialCost(cost_functions.weighted_l2_norm, X[0], {"Q": np.linalg.inv(P)})
    t_range = np.linspace(0, T, 51)
    problem.addResidualCost(measurements.multi_receiver_range_3d, X, t_range, 45.72 * np.ones((1, 51)),
                            0.1 * np.array([1. / 0.01]), {"idxA": [0, 1, 2], "idxB": [5, 6, 7]})
    for i in range(N + 1):
        problem.addEqConstraint(constraints.equality_constaint, [X[i][2], X[i][7]])
    problem.addResidualCost(measurements.multi_receiver_heading_2d, X, t_range, np.deg2rad(-44) * np.ones((1, 51)),
                            0.1 * np.array([1. / 0.1]), {"idxA": [0, 1], "idxB": [5, 6]})
    dref = shapes.two_receiver()
    bt = helpers.two_receiver_batch(dref, 1, seed=77)
    slots = []
    for i in range(6):
        t_i = np.array([[float(i)]])
        for j in range(10):
            for idx in ([0, 1, 2, 3], [5, 6, 7, 8]):
                s = problem.addParameter(1, 3)[0]
                Rp = problem.addParameter(1, 1)[0]
                Yp = problem.addResidualCost(measurements.pseudorange, X, t_i, None, Rp, {"p": 1, "sat_pos": s, "idx": idx})[0]
                slots.append((Yp, Rp, s, t_i))
    problem.build()
    d = problem._desc
    assert d.nv == 9 and d.M == dref.M and list(d.cmap) == list(dref.cmap)
    base = 102
    for q, (Yp, Rp, s, t_i) in enumerate(slots):
        o = dref.roff[base + q]
        rec = bt["data"][0, o:o + 5]
        problem.setParameter(Rp, rec[1])
        problem.setParameter(s, rec[2:5])
        problem.setMeasurement(Yp, t_i, np.array([[rec[0]]]))
    problem.setControl(U, np.linspace(0, T, 6), np.zeros((6, 6)))
    problem.setParameter(X0, bt["xprior"][0])
    for k, xv in enumerate(X):
        problem.initialGuess(xv, bt["xprior"][0])
    problem.solve()
    win = nls.Window(dref, U=bt["U"][0], data=bt["data"][0], xprior=bt["xprior"][0])
    zo, info = nls.solve(dref, win, bt["z0"][0])
    xa = np.array([problem.extractVariableValue('x', k) for k in range(N + 1)])
    Vo = zo.reshape(dref.K, dref.nv)[:, dref.cmap]
    assert np.max(np.abs(xa - Vo) / (1 + np.abs(Vo))) < 1e-6
    assert np.max(np.abs(xa[:, 2] - xa[:, 7])) == 0.0            # the equality holds exactly
    assert abs(problem.solver["objective"] - info["J"]) < 1e-8 * info["J"]


def test_unregistered_models_fail_loudly():
    import nlp.nlp as nlp
    import nlp.cost_functions as cost_functions
    import utils.ekf as ekf
    from nlp_filter_b200._capi import NlpfError
    problem = nlp.fixedTimeOptimalEstimationNLP(5, 1.0, 2, 1)
    X = problem.addVariables(6, 2, name='x')
    # (a smooth unregistered model is now lowered by code generation -- tests/test_gpu_codegen.py; one that cannot be
    # traced still fails loudly instead of falling back to the CPU)
    problem.addDynamics(lambda x, u, params=None: np.array([np.sin(x[0]) + u[0], x[1] ** 2 if x[1] > 0 else 0.0]), X, None, None)
    with pytest.raises(NotImplementedError):
        problem.build()
    f = ekf.EKF(lambda *a, **k: None, lambda *a, **k: None, np.zeros(5), np.eye(5))
    with pytest.raises(NlpfError):
        f.update(np.zeros(3), None, np.eye(5), None, dyn_func_params={"dt": 1.0})


def test_batched_ocp_monte_carlo_over_x0_equals_single_solves():
    """Row A8: B copies of van_der_pol.py's optimal-control problem that differ in x0, solved together
    (solve_batch: one whole-solve launch, one defect evaluation and one multiplier update per outer iteration for
    all of them) == the reference-API solve of each one."""
    import nlp.nlp as nlp
    import nlp.dynamics as dynamics
    import nlp.cost_functions as cost_functions

    def make(x0):
        problem = nlp.fixedTimeOptimalControlNLP(20, 10., 2, 1)
        X = problem.addVariables(21, 2, [-np.inf, -0.25], [np.inf, np.inf], name='x')
        U = problem.addVariables(21, 1, [-1.0], [1.0], name='u')
        problem.addDynamics(dynamics.van_der_pol, X, U)
        problem.addStageCost(cost_functions.van_der_pol, X, U)
        problem.addInitialCondition(X[0], x0)
        problem.build()
        return problem
    rng = np.random.default_rng(8)
    x0s = np.array([0.0, 1.0])[None, :] + rng.normal(0, 0.2, (48, 2))
    x0s[0] = [0.0, 1.0]
    p = make([0.0, 1.0])
    Xb, Ub, Jb, st = p.solve_batch(x0s)
    assert np.all(st == 0), st
    assert abs(Jb[0] - 2.8784425681) < 1e-6                       # the script's own optimum
    assert np.max(np.abs(Xb[:, 0, :] - x0s)) < 1e-9               # the end-point condition holds
    for b in (0, 5, 47):
        q = make(x0s[b])
        q.solve()
        Xs = np.array([q.extractVariableValue('x', k) for k in range(21)])
        Us = np.array([q.extractVariableValue('u', k) for k in range(21)])
        assert np.max(np.abs(Xb[b] - Xs)) < 1e-7 and np.max(np.abs(Ub[b] - Us)) < 1e-7
        assert abs(Jb[b] - q.J) < 1e-8 * max(1.0, q.J)
