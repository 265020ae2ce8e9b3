"""oracle/multiseg.py with one segment is the reference's single-interval problem: same minimiser as oracle/nls.py
(which is pinned by the stored IPOPT outputs); and its dense H is the overlap-add it claims to be."""
import numpy as np

from nlp_filter_b200 import shapes, synth
from oracle import multiseg, nls


def test_one_segment_is_the_single_interval_problem():
    d = shapes.vdp_mhe()
    bt = synth.vdp_mhe_batch(d, 2, seed=5)
    for b in range(2):
        z1, i1 = nls.solve(d, nls.Window(d, U=bt["U"][b], data=bt["data"][b]), bt["z0"][b])
        z2, i2 = multiseg.solve(d, 1, bt["z0"][b], bt["U"][b][None], bt["data"][b][None])
        assert i1["status"] == 0 and i2["status"] == 0
        assert np.max(np.abs(z1 - z2)) < 1e-8 and abs(i1["J"] - i2["J"]) < 1e-9 * i1["J"]


def test_gradient_of_the_multi_segment_objective_by_finite_differences():
    seg = shapes.vdp_segment()
    S = 3
    bt = synth.vdp_multiseg_batch(seg, S, 1, seed=9)
    z = bt["z0"][0]
    J, g, H = multiseg.evaluate(seg, S, z, bt["U"][0], bt["data"][0], pinv=np.array([1.0, 2.0]), xprior=bt["x0"][0])
    for i in (0, 1, 9, 10, 11, 20, 31):
        e = np.zeros_like(z); e[i] = 1e-6
        Jp = multiseg.evaluate(seg, S, z + e, bt["U"][0], bt["data"][0], pinv=np.array([1.0, 2.0]), xprior=bt["x0"][0])[0]
        Jm = multiseg.evaluate(seg, S, z - e, bt["U"][0], bt["data"][0], pinv=np.array([1.0, 2.0]), xprior=bt["x0"][0])[0]
        assert abs((Jp - Jm) / 2e-6 - 2.0 * g[i]) < 1e-5 * (1 + abs(2.0 * g[i]))      # g is the gradient of J / 2
    assert np.allclose(H, H.T)
    assert np.all(H[:10, 22:] == 0.0)              # block s couples to block s + 1 only through one shared node
