"""SURVEY.md section 8(f) item 4 on the device: a model callable that is NOT in the kernel library goes through code
generation + NVRTC and the ordinary whole-solve / assemble entry points.  Checked against the oracle, which evaluates
the user's callable itself and differentiates it numerically (oracle/models.py) -- nothing shared with the symbolic
differentiation and the generated CUDA source."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _names():
    import nlp_filter_b200
    nlp_filter_b200.install_reference_names()


def _pendulum_cart(x, u, params=None):
    from casadi import vertcat, sin, cos, sqrt          # the stand-in published by install_reference_names()
    g_l, c = params["g_over_l"], params["damping"]
    return vertcat(x[1],
                   -g_l * sin(x[0]) - c * x[1] + u[0] * cos(x[0]),
                   x[3],
                   u[0] - 0.1 * x[3] * sqrt(1.0 + x[3] ** 2))


PAR = {"g_over_l": 9.81 / 0.7, "damping": 0.05}


def _problem(T=2.0, N=12, n_meas=25, seed=4):
    import nlp.nlp as nlp
    import nlp.cost_functions as cost_functions
    import nlp.measurements as measurements
    from scipy.integrate import solve_ivp
    rng = np.random.default_rng(seed)
    t = np.linspace(0, T, n_meas)
    ufun = lambda tt: 0.8 * np.sin(2.0 * tt)
    sol = solve_ivp(lambda tt, x: np.asarray(_pendulum_cart(x, [ufun(tt)], PAR), dtype=float), (0, T), [0.4, 0.0, 0.0, 0.1],
                    t_eval=t, rtol=1e-10, atol=1e-12)
    y = sol.y + rng.normal(size=sol.y.shape) * np.array([0.02, 0.05, 0.02, 0.05])[:, None]
    p = nlp.fixedTimeOptimalEstimationNLP(N, T, 4, 1)
    X = p.addVariables(N + 1, 4, name='x')
    p.initializeEstimate(X, t, y)
    U, W = p.addDynamics(_pendulum_cart, X, t, ufun(t)[None, :], PAR)
    p.addDynamicsCost(cost_functions.weighted_l2_norm, W, {"Q": np.linalg.inv(np.diag([1e-3, 1e-2, 1e-3, 1e-2]))})
    p.addResidualCost(measurements.full_state, X, t, y, np.linalg.inv(np.diag([0.02, 0.05, 0.02, 0.05]) ** 2))
    p.build()
    return p, X, t, sol.y


def test_user_defined_model_solves_like_the_oracle():
    from nlp_filter_b200 import desc as D
    from oracle import nls
    p, X, t, truth = _problem()
    d = p._desc
    assert d.dyn_id == D.DYN_USER
    p.solve()
    assert p.solver["success"]
    z = p._z_last
    Uv = np.array([q.value for q in p._dyn[2]]).reshape(d.K, 1)
    win = nls.Window(d, U=Uv, data=p._pack_data())
    zo, info = nls.solve(d, win, p._initial_z(), D.SolveOpts(tol_step=1e-11))
    assert info["status"] <= 1
    assert np.max(np.abs(z - zo) / (1 + np.abs(zo))) < 1e-6
    assert abs(p.solver["objective"] - info["J"]) < 1e-7 * info["J"]
    x_hat = p.extractSolution('x', t)
    assert np.sqrt(np.mean((x_hat.T - truth) ** 2)) < 0.05                 # and it is an estimate of the simulated truth
    # kernels (1)+(2) on the generated model: H, g, J against the oracle (its Jacobian is a central difference: 1e-6)
    prob = p._problem_handle()
    rng = np.random.default_rng(0)
    zs = z[None, :] + 0.05 * rng.normal(size=(3, d.nz))
    H, g, cost = prob.assemble(zs, np.tile(Uv, (3, 1, 1)), np.tile(win.data, (3, 1)))
    for b in range(3):
        Jo, mo, go, Ho = nls.evaluate(d, win, zs[b])
        assert abs(float(cost[b, 0]) - Jo) < 1e-11 * Jo
        assert np.max(np.abs(g[b].cpu().numpy() - go)) < 1e-6 * np.max(np.abs(go))
        assert np.max(np.abs(H[b].cpu().numpy() - Ho)) < 1e-6 * np.max(np.abs(Ho))


def test_unregistered_copy_of_a_library_model_gives_the_library_models_estimate():
    """van der Pol written by the user (so: traced, generated, NVRTC) against the hand-written VDP kernel model."""
    import nlp.nlp as nlp
    import nlp.dynamics as dynamics
    import nlp.cost_functions as cost_functions
    import nlp.measurements as measurements
    from nlp_filter_b200 import desc as D
    rng = np.random.default_rng(2)
    T, N = 5.0, 12
    t = np.linspace(0, T, 30)
    y = np.stack([np.cos(t), -np.sin(t)]) + 0.05 * rng.normal(size=(2, 30))
    u = 0.3 * np.sin(t)[None, :]
    out = []
    for func in (dynamics.van_der_pol, lambda x, u, params=None: [(1 - x[1] ** 2) * x[0] - x[1] + u[0], x[0]]):
        p = nlp.fixedTimeOptimalEstimationNLP(N, T, 2, 1)
        X = p.addVariables(N + 1, 2, name='x')
        p.initializeEstimate(X, t, y)
        U, W = p.addDynamics(func, X, t, u, None)
        p.addDynamicsCost(cost_functions.weighted_l2_norm, W, {"Q": np.linalg.inv(np.diag([1e-2, 1e-2]))})
        p.addResidualCost(measurements.full_state, X, t, y, np.linalg.inv(np.diag([0.05, 0.05]) ** 2))
        p.build()
        p.solve()
        out.append((p._desc.dyn_id, p._z_last.copy(), p.solver["objective"]))
    assert out[0][0] == D.DYN_VDP and out[1][0] == D.DYN_USER
    assert np.max(np.abs(out[0][1] - out[1][1])) < 1e-6 and abs(out[0][2] - out[1][2]) < 1e-8 * out[0][2]
