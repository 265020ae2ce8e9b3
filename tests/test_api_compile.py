"""CPU-only: problems recorded through the reference API (nlp.nlp) lower to the same flat
descriptors as the direct builders in nlp_filter_b200/shapes.py, and the API keeps the
reference's surface behaviour (names, return values, prints, errors)."""
import numpy as np
import pytest

from nlp_filter_b200 import desc as D
from nlp_filter_b200 import shapes


@pytest.fixture(autouse=True)
def _names():
    import nlp_filter_b200
    nlp_filter_b200.install_reference_names()


def _same(a, b, fields=("K", "n", "m", "nv", "ng", "dyn_id", "loss", "has_prior", "E", "M", "Ldata")):
    for f in fields:
        assert getattr(a, f) == getattr(b, f), f
    for f in ("Dt", "cw", "cq", "qinv", "pinv", "Phi", "A", "Bm", "c0", "blo", "bhi", "dyn_par"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    for f in ("cmap", "rtype", "repoch", "ridx", "roff", "bidx"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f


def test_car_problem_lowers_to_the_car_shape(capsys):
    import nlp.nlp as nlp
    import nlp.dynamics as dynamics
    import nlp.cost_functions as cost_functions
    import nlp.measurements as measurements
    import utils.vehicle_sim as vehicle_sim
    p = nlp.fixedTimeOptimalEstimationNLP(5, 2, 9, 2)
    X = p.addVariables(6, 9, name='x')
    U, W = p.addDynamics(dynamics.vehicle_dynamics_and_gnss, X, None, None, {"car_params": vehicle_sim.get_parameters()})
    Q = np.diag([0.01, 0.01, 0.01, 100, 500, 500, .001, .001, .001])
    p.addDynamicsCost(cost_functions.pseudo_huber_loss, W, {"Q": np.linalg.inv(Q), "delta": 5.0})
    p.addVarBounds(X, 2, -np.pi, np.pi)
    p.addVarBounds(X, 3, 0, np.inf)
    X0 = p.addInitialCost(cost_functions.weighted_l2_norm, X[0], {"Q": np.linalg.inv(np.eye(9))})
    for i in range(21):
        for j in range(11):
            s = p.addParameter(1, 3)[0]
            R = p.addParameter(1, 1)[0]
            Y = p.addResidualCost(measurements.vehicle_pseudorange, X, np.array([[np.linspace(0, 2, 21)[i]]]), None, R,
                                  {"p": 1, "sat_pos": s})
            assert len(Y) == 1
    p.build()
    _same(p._desc, shapes.car(loss=D.LOSS_HUBER, delta=5.0))
    assert len(U) == 6 and len(W) == 6 and p.var_names[:2] == ['x_0', 'x_1'] and p.var_sizes[0] == 9
    assert p.extractVariableValue('x', 0) is None
    assert 'Need to run solve() first' in capsys.readouterr().out


def test_gnss_problem_lowers_to_the_gnss_shape():
    import nlp.nlp as nlp
    import nlp.dynamics as dynamics
    import nlp.cost_functions as cost_functions
    import nlp.measurements as measurements
    T = 50
    t = np.linspace(0, T, T + 1)
    p = nlp.fixedTimeOptimalEstimationNLP(10, T, 5, 3)
    X = p.addVariables(11, 5, name='x')
    U, W = p.addDynamics(dynamics.gnss_pos_and_bias, X, t, np.zeros((3, T + 1)))
    p.addDynamicsCost(cost_functions.weighted_l2_norm, W, {"Q": np.linalg.inv(np.diag([0.0001, 0.0001, 0.0001, 0.1, 0.001]))})
    for k in range(T + 1):
        for i in range(12):
            p.addResidualCost(measurements.pseudorange, X, np.array([[float(k)]]), np.array([[2.1e7]]),
                              np.linalg.inv(np.diag([100])), {"sat_pos": np.array([1.0, 2.0, 2e7])})
    p.build()
    ref = shapes.gnss()
    _same(p._desc, ref)
    data = p._pack_data()
    assert data.shape == (ref.Ldata,) and data[0] == 2.1e7 and data[1] == 0.01 and data[4] == 2e7


def test_two_receiver_equalities_fold_into_a_shared_column():
    import nlp.nlp as nlp
    import nlp.dynamics as dynamics
    import nlp.constraints as constraints
    p = nlp.fixedTimeOptimalEstimationNLP(10, 5, 10, 6)
    X = p.addVariables(11, 10, name='x')
    p.addDynamics(dynamics.gnss_two_receiver, X, None, None)
    for i in range(11):
        p.addEqConstraint(constraints.equality_constaint, [X[i][2], X[i][7]])
    p.build()
    ref = shapes.two_receiver()
    assert list(p._desc.cmap) == list(ref.cmap) and p._desc.nv == 9
    assert np.array_equal(p._desc.A, ref.A) and np.array_equal(p._desc.Bm, ref.Bm)
    q = nlp.fixedTimeOptimalEstimationNLP(10, 5, 10, 6)
    Xq = q.addVariables(11, 10, name='x')
    q.addDynamics(dynamics.gnss_two_receiver, Xq, None, None)
    q.addEqConstraint(constraints.equality_constaint, [Xq[0][2], Xq[0][7]])      # not at every node
    with pytest.raises(NotImplementedError):
        q.build()


def test_ocp_lowering_and_affine_probe():
    import nlp.nlp as nlp
    import nlp.dynamics as dynamics
    import nlp.cost_functions as cost_functions
    p = nlp.fixedTimeOptimalControlNLP(20, 10., 2, 1)
    X = p.addVariables(21, 2, [-np.inf, -0.25], [np.inf, np.inf], name='x')
    U = p.addVariables(21, 1, [-1.0], [1.0], name='u')
    p.addDynamics(dynamics.van_der_pol, X, U)
    p.addStageCost(cost_functions.van_der_pol, X, U)
    p.addInitialCondition(X[0], [0, 1])
    p.build()
    d = p._desc
    assert d.dyn_id == D.DYN_VDP and d.nv == 3 and list(d.ucol) == [2] and list(d.bidx) == [1, 2]
    assert list(d.blo) == [-0.25, -1.0] and list(d.bhi) == [np.inf, 1.0] and np.array_equal(d.snode, [1, 1, 1])
    assert d.M == 2 and np.array_equal(d.Phi[0], np.eye(21)[0])
    # a user-defined affine model is accepted through the numeric probe ...
    q = nlp.fixedTimeOptimalEstimationNLP(5, 1.0, 2, 1)
    Xq = q.addVariables(6, 2, name='x')
    q.addDynamics(lambda x, u, params=None: np.array([x[1] + 2.0 * u[0], -3.0 * x[0] + 0.5]), Xq, None, None)
    q.build()
    assert q._desc.dyn_id == D.DYN_AFFINE and np.array_equal(q._desc.A, [[0, 1], [-3, 0]])
    assert np.array_equal(q._desc.Bm[:, 0], [2, 0]) and np.array_equal(q._desc.c0, [0, 0.5])
    # ... a non-affine unregistered one is lowered by code generation (CUDA source for NVRTC; round 2) ...
    r = nlp.fixedTimeOptimalEstimationNLP(5, 1.0, 2, 1)
    Xr = r.addVariables(6, 2, name='x')
    r.addDynamics(lambda x, u, params=None: np.array([np.sin(x[0]), u[0]]), Xr, None, None)
    r.build()
    assert r._desc.dyn_id == D.DYN_USER and "sin(x[0])" in r._desc.user_src and "cos(x[0])" in r._desc.user_src
    # ... and one that cannot be traced (data-dependent branch) is refused: no CPU fallback for arbitrary callables
    r2 = nlp.fixedTimeOptimalEstimationNLP(5, 1.0, 2, 1)
    Xr2 = r2.addVariables(6, 2, name='x')
    r2.addDynamics(lambda x, u, params=None: np.array([x[1] if x[0] > 0 else -x[1], u[0] * x[0]]), Xr2, None, None)
    with pytest.raises(NotImplementedError):
        r2.build()


def test_model_callables_are_dual_use_numeric():
    """The reference hands the same callables to its numeric simulator (utils/simulate.py:6-8,46)."""
    import nlp.dynamics as dynamics
    import nlp.measurements as measurements
    import nlp.cost_functions as cost_functions
    import utils.simulate as simulate
    t = np.linspace(0, 1.0, 11)
    x = simulate.open_loop_sim(t, np.vstack((np.ones(11), 2 * np.ones(11))), np.zeros(2), dynamics.single_integrator_2D)
    assert np.allclose(x[:, -1], [1.0, 2.0], atol=1e-6)
    assert np.allclose(dynamics.van_der_pol(np.array([0.5, -1.0]), 0.25), [(1 - 1.0) * 0.5 + 1.0 + 0.25, 0.5])
    assert np.isclose(measurements.pseudorange(np.array([1.0, 2, 2, 5, 0]), {"sat_pos": np.array([1.0, 2, 0])}), 7.0)
    assert np.isclose(measurements.multi_receiver_heading_2d(np.zeros(10), {"idxA": [0, 1], "idxB": [5, 6]}), np.arctan2(1e-5, 0.0))
    assert np.isclose(cost_functions.weighted_l2_norm(np.array([1.0, 2.0]), {"Q": np.diag([3.0, 0.5])}), 5.0)
    assert np.isclose(cost_functions.pseudo_huber_loss(np.array([0.0, 0.0]), {"Q": np.eye(2), "delta": 5.0}), 0.0)
    f = dynamics.vehicle_dynamics_and_gnss(np.array([0, 0, 0.1, 10.0, 0.2, 0.05, 1.0, 200.0, 0.0]), np.array([100.0, 0.01]),
                                           {"car_params": {"C_AF": 1.1441e5, "C_AR": 1.3388e5, "M": 2009, "D_F": 1.53,
                                                           "D_R": 1.23, "I_Z": 2000}})
    assert f.shape == (9,) and f[6] == 200.0
