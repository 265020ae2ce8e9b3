"""CPU-only: the oracle's solver against an INDEPENDENT minimiser (scipy BFGS / L-BFGS-B) of the same
objective.  oracle/nls.py restates both the reference's objective (pinned by the stored IPOPT outputs of the
vehicle example, tests/test_oracle_car_golden.py) and this repo's own Gauss-Newton / barrier iteration; for the
shapes without stored reference outputs (van der Pol, GNSS: "parity unpinned" in DESIGN.md) this test at least
rules out that the two implementations of the iteration agree on something that is not the minimiser:
  * the analytic gradient matches central differences of J,
  * a quasi-Newton method that shares no code with the iteration finds no lower J from the oracle's point
    (nor from the start point), and ends at the same estimate."""
import numpy as np
import pytest
from scipy.optimize import minimize

from nlp_filter_b200 import shapes, synth
from oracle import nls


def _funs(d, win):
    def fun(z):
        J, merit, g, H = nls.evaluate(d, win, z)
        return J, 2.0 * g                      # merit = J / 2 without bounds
    return fun


@pytest.mark.parametrize("shape", ["vdp", "gnss"])
def test_oracle_solution_is_the_minimiser_an_independent_method_finds(shape):
    if shape == "vdp":
        d = shapes.vdp_mhe()
        bt = synth.vdp_mhe_batch(d, 3, seed=51)
    else:
        d = shapes.gnss()
        bt = synth.gnss_batch(d, 1, seed=52)
    for b in range(bt["z0"].shape[0]):
        win = nls.Window(d, U=bt["U"][b], data=bt["data"][b])
        zo, info = nls.solve(d, win, bt["z0"][b])
        assert info["status"] <= 1
        fun = _funs(d, win)
        Jo, go = fun(zo)
        # (1) gradient vs central differences at a non-stationary point
        zt = bt["z0"][b].reshape(-1).copy()
        J0, g0 = fun(zt)
        rng = np.random.default_rng(b)
        for _ in range(6):
            v = rng.normal(size=zt.shape)
            v /= np.linalg.norm(v)
            h = 1e-5 * (1.0 + np.max(np.abs(zt)))
            fd = (fun(zt + h * v)[0] - fun(zt - h * v)[0]) / (2 * h)
            # (J itself is only good to ~1e-11 relative with 2e7 m pseudoranges: the difference quotient carries that)
            assert abs(fd - g0 @ v) <= 1e-6 * (abs(g0 @ v) + np.linalg.norm(g0) * 0.05 + 1e-9)
        # (2) stationarity of the oracle's point, relative to the gradient at the start
        assert np.linalg.norm(go) <= 1e-7 * (1.0 + np.linalg.norm(g0))
        # (3) an independent quasi-Newton run from the start point and from the oracle's point
        scale = 1.0 + np.abs(zo)
        for start in (zt, zo + 1e-3 * rng.normal(size=zo.shape) * scale):
            res = minimize(fun, start, jac=True, method="L-BFGS-B",
                           options=dict(maxiter=20000, maxfun=40000, ftol=1e-15, gtol=1e-10, maxcor=50))
            assert res.fun >= Jo * (1.0 - 1e-9) - 1e-12          # nothing below the oracle's optimum
            if res.fun <= Jo * (1.0 + 1e-9):                      # where the quasi-Newton run got all the way down
                assert np.max(np.abs(res.x - zo) / scale) < (1e-5 if shape == "vdp" else 1e-3)
