"""Pins the collocation constants (oracle restatement AND the product class) to outputs of the
reference's own class (tests/golden/collocation_*.npz, nlp/collocation.py:7-121)."""
import glob
import os

import numpy as np
import pytest

from nlp_filter_b200.nlp.collocation import ChebyshevPseudospectralMethod
from oracle.collocation import Collocation
from tests.helpers import GOLDEN

FILES = sorted(glob.glob(os.path.join(GOLDEN, "collocation_*.npz")))
ULP = 4 * np.finfo(float).eps


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_constants_match_reference(path):
    g = np.load(path)
    N, T = int(g["N"]), float(g["T"])
    o = Collocation(N, T)
    p = ChebyshevPseudospectralMethod(N, 0, T)
    for tau, D, w, Phi, tn in ((o.tau, o.D, o.w, o.phi(g["t"]), o.t_nodes()),
                               (p.tau, p.D, p.w, p.basisMatrix(g["t"]), p.tau2t(p.tau))):
        assert np.max(np.abs(tau - g["tau"])) <= ULP
        assert np.max(np.abs(D - g["D"]) / (1 + np.abs(g["D"]))) <= ULP
        assert np.max(np.abs(w - g["w"])) <= ULP
        assert np.max(np.abs(tn - g["t_nodes"])) <= ULP * T
        # the monomial-basis evaluation is ill conditioned for N = 20: compare relative to the
        # size of the intermediate Horner terms, not to the (cancelled) result
        assert np.max(np.abs(Phi - g["Phi"])) <= 1e-9 if N >= 20 else np.max(np.abs(Phi - g["Phi"])) <= 1e-13
    X = [g["X"][k] for k in range(N + 1)]
    xs = np.array([p.evaluateSolution(t, X) for t in g["t"]])
    assert np.max(np.abs(xs - g["xs"])) <= 1e-7 * np.max(np.abs(g["xs"]))


def test_reference_weight_quirks():
    """not Clenshaw-Curtis.  sum(w) != 2 and N = 3 has a zero weight."""
    assert abs(Collocation(5, 1.0).w.sum() - 1.9733333333333334) < 1e-12
    assert abs(Collocation(10, 1.0).w.sum() - 2.006060606060606) < 1e-9
    w3 = Collocation(3, 1.0).w
    assert w3[2] == 0.0 and abs(w3.sum() - 0.8888888888888888) < 1e-12
    assert np.allclose(Collocation(5, 1.0).w, [0.04, 0.339167, 0.6075, 0.6075, 0.339167, 0.04], atol=1e-6)


def test_diff_matrix_differentiates_on_ascending_grid():
    for N in (5, 10, 20):
        c = Collocation(N, 2.0)
        assert np.max(np.abs(c.D @ c.tau - 1.0)) < 1e-11
        assert np.max(np.abs(c.D @ np.ones(N + 1))) < 1e-11
        assert np.max(np.abs(c.D @ c.tau ** 3 - 3 * c.tau ** 2)) < 1e-10


def test_scalar_and_one_element_time_arguments():
    p = ChebyshevPseudospectralMethod(10, 0, 50.0)
    a = p.evaluateLagrangePolynomials(12.5)
    b = p.evaluateLagrangePolynomials(np.array([12.5]))      # rows of a (1,1) t_array, gnss_stationary.py:127
    assert a.shape == (11,) and np.array_equal(a, b)
    assert abs(a.sum() - 1.0) < 1e-10
