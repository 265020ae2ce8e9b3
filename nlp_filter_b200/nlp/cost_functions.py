"""Stage / noise cost functions under the reference's names (nlp/cost_functions.py:5-31).
`KERNEL_COSTS` tells the problem compiler which kernel loss each one selects."""
import numpy as np

from ._sym import mtimes, sqrt
from ..desc import LOSS_L2, LOSS_HUBER

KERNEL_COSTS = {}


def van_der_pol(x, u, params=None):
    """x0^2 + x1^2 + u^2   (cost_functions.py:5-7)."""
    return x[0] ** 2 + x[1] ** 2 + u ** 2


def single_integrator(x, u, params=None):
    """|x|^2 + |u|^2 for 2-D x, u   (cost_functions.py:10-12)."""
    return x[0] ** 2 + x[1] ** 2 + u[0] ** 2 + u[1] ** 2


def l2_norm(x, params=None):
    """||x||^2   (cost_functions.py:15-17)."""
    x = np.asarray(x, dtype=float).reshape(-1)
    return mtimes(x.T, x)


def weighted_l2_norm(x, params=None):
    """x^T Q x with Q = params["Q"]   (cost_functions.py:20-22)."""
    x = np.asarray(x, dtype=float).reshape(-1)
    return mtimes(x.T, mtimes(params["Q"], x))


def pseudo_huber_loss(x, params=None):
    """sum_i 2 Q_ii delta^2 (sqrt(1 + x_i^2/delta^2) - 1)   (cost_functions.py:25-31)."""
    Q, delta = params["Q"], params["delta"]
    return sum(2 * Q[i, i] * delta ** 2 * (sqrt(1 + (x[i]) ** 2 / delta ** 2) - 1.0) for i in range(Q.shape[0]))


KERNEL_COSTS.update({l2_norm: ("l2", LOSS_L2), weighted_l2_norm: ("wl2", LOSS_L2),
                     pseudo_huber_loss: ("huber", LOSS_HUBER)})
