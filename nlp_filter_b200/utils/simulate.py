"""Open-loop simulation and measurement generation with the reference's names
(utils/simulate.py:6-47).  Host-side data generation, not on the hot path."""
import numpy as np
from scipy.integrate import ode
from scipy.interpolate import interp1d


def dynamics_wrapper(t, x, func, u, params=None):
    return func(x, u(t), params)


def open_loop_sim(t, u, x0, func, params=None):
    """Integrate xdot = func(x, u(t)) from x0 over the grid t (u: m x T, linearly interpolated)."""
    u_t = interp1d(t, u, fill_value="extrapolate")
    solver = ode(dynamics_wrapper)
    solver.set_initial_value(x0, t[0])
    solver.set_f_params(func, u_t, params)
    x = np.zeros((x0.shape[0], t.shape[0]))
    x[:, 0] = x0
    i = 0
    while solver.successful() and solver.t < t[-1]:
        i += 1
        x[:, i] = solver.integrate(solver.t + (t[i] - t[i - 1]))
    return x


def generate_measurements(x, measurement_model, sigma, params=None):
    """y_i = h(x_i) + N(0, sigma) for every column of x (n x T)."""
    T, p = x.shape[1], sigma.shape[0]
    y = np.zeros((p, T))
    for i in range(T):
        y[:, i] = measurement_model(x[:, i], params) + np.random.multivariate_normal(np.zeros(p), sigma)
    return y
