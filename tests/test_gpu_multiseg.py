"""Row A9 end to end: multi-segment collocation lowering -> block-tridiagonal normal equations -> Riccati / cyclic
reduction inside the damped Gauss-Newton iteration (nlp_filter_b200/multiseg.py).
  * 1 segment = the reference's single-interval problem: same estimate as the dense whole-solve kernel;
  * S segments: the block-tridiagonal H, g, J equal the dense overlap-add of oracle/multiseg.py, the block solve equals
    the dense solve of that H, and the converged long-horizon estimate equals the oracle's dense iteration;
  * a 200-epoch horizon (40 segments x 5 epochs, BASELINE configs[3]'s "horizon 200" read as ONE window)."""
import numpy as np
import pytest

from nlp_filter_b200 import shapes, synth

pytestmark = pytest.mark.gpu


def _np(t):
    return t.cpu().numpy()


def test_one_segment_equals_the_dense_kernel():
    from nlp_filter_b200 import engine, multiseg
    d = shapes.vdp_mhe()
    B = 24
    bt = synth.vdp_mhe_batch(d, B, seed=77)
    prob = engine.WindowProblem(d)
    zd, Jd, itd, std = engine.solve_windows_host(prob, bt["z0"], bt["U"], bt["data"])
    ms = multiseg.MultiSegmentProblem(shapes.vdp_mhe(), 1)
    import torch
    z = torch.as_tensor(bt["z0"].copy()).cuda()
    for variant in (multiseg.RICCATI,):          # one block of 40 + the end node: bs = 40 > the cyclic-reduction limit
        zz, J, it, st = ms.solve(z.clone(), bt["U"][:, None], bt["data"][:, None], variant=variant)
        assert np.all(_np(st) == 0)
        assert np.max(np.abs(_np(zz) - zd) / (1 + np.abs(zd))) < 1e-6
        assert np.max(np.abs(_np(J) - Jd) / Jd) < 1e-9
    ms.close(); prob.close()


@pytest.mark.parametrize("S", [2, 8])
def test_block_tridiagonal_normal_equations_and_solve(S):
    from nlp_filter_b200 import engine, multiseg
    from oracle import multiseg as oms
    seg = shapes.vdp_segment()
    B = 5
    bt = synth.vdp_multiseg_batch(seg, S, B, seed=10 + S)
    ms = multiseg.MultiSegmentProblem(seg, S, pinv=np.array([2.0, 3.0]))
    z = bt["z0"] + 0.05 * np.random.default_rng(0).normal(size=bt["z0"].shape)
    Dg, Lo, g, J = [_np(t) for t in ms.normal_equations(z, bt["U"], bt["data"], xprior=bt["x0"])]
    nzt = ms.nzt
    for b in range(B):
        Jo, go, Ho = oms.evaluate(seg, S, z[b], bt["U"][b], bt["data"][b], pinv=np.array([2.0, 3.0]), xprior=bt["x0"][b])
        A = oms.dense_from_blocks(Dg[b], Lo[b], nzt)
        assert abs(J[b] - Jo) <= 1e-12 * abs(Jo)
        assert np.max(np.abs(g[b][:nzt] - go)) <= 1e-10 * np.max(np.abs(go))
        assert np.max(np.abs(A - Ho)) <= 1e-10 * np.max(np.abs(Ho))
    # damped system through both block solvers == dense solve
    lam = 0.1
    Dl, Ll, gl, _ = [t for t in ms.normal_equations(z, bt["U"], bt["data"], xprior=bt["x0"], lam=lam)]
    for variant in (0, 1):
        x, info = engine.blocktri_solve(Dl, Ll, gl.reshape(B, ms.nblk, ms.bs), variant)
        assert np.all(_np(info) == 0)
        for b in range(B):
            _, go, Ho = oms.evaluate(seg, S, z[b], bt["U"][b], bt["data"][b], pinv=np.array([2.0, 3.0]), xprior=bt["x0"][b])
            ref = np.linalg.solve(Ho + lam * np.diag(np.abs(np.diag(Ho))), go)
            assert np.max(np.abs(_np(x)[b].reshape(-1)[:nzt] - ref)) <= 1e-9 * (1 + np.max(np.abs(ref)))
    ms.close()


@pytest.mark.parametrize("variant", [0, 1])
def test_200_epoch_horizon_matches_the_dense_oracle(variant):
    from nlp_filter_b200 import multiseg
    from oracle import multiseg as oms
    import torch
    seg = shapes.vdp_segment()
    S, B = 40, 48                                  # 40 segments x 5 epochs = 200 epochs, 201 nodes, 402 unknowns per window
    bt = synth.vdp_multiseg_batch(seg, S, B, seed=6000)
    ms = multiseg.MultiSegmentProblem(seg, S)
    z = torch.as_tensor(bt["z0"].copy()).cuda()
    z_eager = torch.as_tensor(bt["z0"].copy()).cuda()
    z_eager, J_e, it_e, st_e = ms.solve(z_eager, bt["U"], bt["data"], variant=variant, graph=False)     # every sweep issued op by op
    z, J, it, st = ms.solve(z, bt["U"], bt["data"], variant=variant)                                     # sweeps 2.. replayed as one CUDA graph
    assert torch.equal(z, z_eager) and torch.equal(J, J_e) and torch.equal(it, it_e) and torch.equal(st, st_e)
    z2, J2, it2, st2 = ms.solve(torch.as_tensor(bt["z0"].copy()).cuda(), bt["U"], bt["data"], variant=variant)   # the cached graph, again
    assert torch.equal(z2, z) and torch.equal(it2, it)
    z, J, it, st = _np(z), _np(J), _np(it), _np(st)
    assert np.all(st == 0), np.bincount(st)
    rms = np.sqrt(np.mean((z.reshape(B, -1, 2) - bt["truth_nodes"]) ** 2))
    assert rms < 0.06                              # measurement noise is 0.1 / 0.14: the estimate is better than the data
    for b in (0, 17, 47):
        zo, info = oms.solve(seg, S, bt["z0"][b], bt["U"][b], bt["data"][b])
        assert info["status"] == 0
        assert np.max(np.abs(z[b] - zo) / (1 + np.abs(zo))) < 1e-6
        assert abs(J[b] - info["J"]) <= 1e-9 * info["J"]
    xs = ms.sample(z[:2], np.array([0.0, 3.3, 39.99]))
    assert np.allclose(xs[:, 0, :], z[:2, :2])
    ms.close()
