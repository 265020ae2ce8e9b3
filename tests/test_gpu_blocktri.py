"""Row A9: block-tridiagonal SPD solves (Riccati recursion and cyclic reduction) validated against
the dense kernel (3) (`nlpf_factor_solve`) on the same matrices and against LAPACK."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _system(B, nblk, bs, seed):
    rng = np.random.default_rng(seed)
    Lo = rng.normal(size=(B, nblk, bs, bs))
    Lo[:, 0] = 0.0
    Dg = np.zeros((B, nblk, bs, bs))
    for i in range(nblk):
        M = rng.normal(size=(B, bs, bs + 2))
        Dg[:, i] = M @ M.transpose(0, 2, 1) + 2.0 * bs * np.eye(bs)     # dominant: the whole matrix is SPD
    rhs = rng.normal(size=(B, nblk, bs))
    return Dg, Lo, rhs


def _dense(Dg, Lo):
    B, nblk, bs, _ = Dg.shape
    A = np.zeros((B, nblk * bs, nblk * bs))
    for i in range(nblk):
        A[:, i * bs:(i + 1) * bs, i * bs:(i + 1) * bs] = Dg[:, i]
        if i > 0:
            A[:, i * bs:(i + 1) * bs, (i - 1) * bs:i * bs] = Lo[:, i]
            A[:, (i - 1) * bs:i * bs, i * bs:(i + 1) * bs] = Lo[:, i].transpose(0, 2, 1)
    return A


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("nblk,bs", [(1, 4), (2, 3), (5, 6), (8, 10), (13, 16), (21, 10), (200, 10), (9, 30), (41, 32), (7, 45), (201, 45),
                                     (5, 48)])
def test_blocktri_matches_lapack(variant, nblk, bs):
    """Up to 200 segments; block sizes up to 48 through the Riccati recursion (a vehicle segment of 5 nodes x 9 states is
    45), up to 32 through cyclic reduction (its per-warp workspace)."""
    from nlp_filter_b200 import engine
    if variant == 1 and bs > 32:
        from nlp_filter_b200 import _capi
        with pytest.raises(_capi.NlpfError):
            engine.blocktri_solve(*_system(2, nblk, bs, seed=1), variant)
        return
    B = 6
    Dg, Lo, rhs = _system(B, nblk, bs, seed=100 + nblk)
    x, info = engine.blocktri_solve(Dg, Lo, rhs, variant)
    x, info = x.cpu().numpy(), info.cpu().numpy()
    assert np.all(info == 0)
    A = _dense(Dg, Lo)
    ref = np.linalg.solve(A, rhs.reshape(B, -1)[..., None])[..., 0].reshape(B, nblk, bs)
    assert np.max(np.abs(x - ref)) < 1e-10 * (1 + np.max(np.abs(ref)))


@pytest.mark.parametrize("variant", [0, 1])
def test_blocktri_matches_dense_kernel(variant):
    """Same matrices through kernel (3): the dense in-shared-memory LDL^T."""
    from nlp_filter_b200 import engine, shapes
    d = shapes.gnss()                      # nz = 55 = 11 blocks of 5
    prob = engine.WindowProblem(d)
    B, nblk, bs = 8, 11, 5
    Dg, Lo, rhs = _system(B, nblk, bs, seed=7)
    x, info = engine.blocktri_solve(Dg, Lo, rhs, variant)
    dz, info2 = prob.factor_solve(_dense(Dg, Lo), rhs.reshape(B, -1), 0.0)      # dz = -H^-1 g
    assert np.all(info.cpu().numpy() == 0) and np.all(info2.cpu().numpy() == 0)
    xd = -dz.cpu().numpy().reshape(B, nblk, bs)
    assert np.max(np.abs(x.cpu().numpy() - xd)) < 1e-10 * (1 + np.max(np.abs(xd)))
    prob.close()


@pytest.mark.parametrize("variant", [0, 1])
def test_blocktri_reports_indefinite(variant):
    from nlp_filter_b200 import engine
    Dg, Lo, rhs = _system(4, 6, 4, seed=3)
    Dg[2, 3] = -Dg[2, 3]
    x, info = engine.blocktri_solve(Dg, Lo, rhs, variant)
    info = info.cpu().numpy()
    assert info[2] == 1 and info[0] == 0 and info[1] == 0 and info[3] == 0
