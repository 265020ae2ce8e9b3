"""GPU parity of kernel (4) (fused EKF) and the LS kernels against the reference's own outputs
(tests/golden/ekf_ls.npz) and the CPU oracle.  Tolerance: 1e-9 relative (BASELINE north_star)."""
import os

import numpy as np
import pytest

from tests.helpers import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    return np.load(os.path.join(GOLDEN, "ekf_ls.npz"))


@pytest.fixture(scope="module", autouse=True)
def _names():
    import nlp_filter_b200
    nlp_filter_b200.install_reference_names()


def _eng():
    from nlp_filter_b200 import engine
    return engine


def _relerr(a, ref):
    return np.max(np.abs(a - ref) / (1 + np.abs(ref)))


def test_ekf_gnss5_stationary_log(G):
    eng = _eng()
    mu, S, traj = eng.ekf_run(eng.EKF_GNSS5, G["gs_ekf_x0"][None], np.eye(5)[None], np.zeros((1, 51, 3)),
                              G["gs_pr"][None], G["gs_sat_enu"][None], G["gs_nsat"][None],
                              np.array([0.0001, 0.0001, 0.0001, 0.1, 0.001]), 100.0, 1.0)
    assert _relerr(traj.cpu().numpy()[0], G["gs_ekf_traj"]) < 1e-9
    assert np.max(np.abs(S.cpu().numpy()[0] - G["gs_ekf_S"])) < 1e-9


def test_ekf_gnss5_synthetic_batch_ragged(G):
    eng = _eng()
    B = 16
    mu, S, traj = eng.ekf_run(eng.EKF_GNSS5, G["syn_ekf_x0"], np.tile(np.eye(5), (B, 1, 1)), G["syn_ekf_u"],
                              G["syn_ekf_z"], G["syn_ekf_sat"], G["syn_ekf_nsat"], G["syn_ekf_Q"], 100.0, 1.0)
    assert _relerr(traj.cpu().numpy(), G["syn_ekf_traj"]) < 1e-9
    assert np.max(np.abs(S.cpu().numpy() - G["syn_ekf_S"])) < 1e-9


def test_ekf_car9_with_skipped_updates(G):
    """100 Hz vehicle filter, GNSS every 10th step (autonomous-car.py:146-160); stable part only."""
    eng = _eng()
    NST = 1500
    sat = np.ascontiguousarray(np.broadcast_to(G["car_sat_enu"], (1, NST, 11, 3)))
    mu, S, traj = eng.ekf_run(eng.EKF_CAR9, G["car_x0"][None], np.eye(9)[None], G["car_u"][None, :NST],
                              G["car_z"][None, :NST], sat, G["car_nsat"][None, :NST], G["car_Q"], float(G["car_r"]),
                              0.01, par=G["car_par"])
    t = traj.cpu().numpy()[0]
    # rounding is amplified by this filter (it diverges at k = 5852 in the reference's own run)
    assert _relerr(t[:100], G["car_traj"][:100]) < 1e-9
    assert _relerr(t, G["car_traj"][:NST]) < 1e-7


def test_ekf_many_replicas_identical(G):
    eng = _eng()
    B = 300
    x0 = np.tile(G["syn_ekf_x0"][3], (B, 1))
    rep = lambda a: np.ascontiguousarray(np.broadcast_to(a[3], (B,) + a[3].shape))
    mu, S, traj = eng.ekf_run(eng.EKF_GNSS5, x0, np.tile(np.eye(5), (B, 1, 1)), rep(G["syn_ekf_u"]), rep(G["syn_ekf_z"]),
                              rep(G["syn_ekf_sat"]), rep(G["syn_ekf_nsat"]), G["syn_ekf_Q"], 100.0, 1.0)
    t = traj.cpu().numpy()
    assert np.all(t == t[0])
    assert _relerr(t[0], G["syn_ekf_traj"][3]) < 1e-9


def test_ls_synthetic_cold_start(G):
    eng = _eng()
    x, it = eng.ls_solve(G["syn_ls_sat"], G["syn_ls_pr"], G["syn_ls_nsat"])
    assert _relerr(x.cpu().numpy(), G["syn_ls_sol"]) < 1e-9
    assert np.all(it.cpu().numpy() <= 10)


def test_ls_stationary_log_independent_epochs(G):
    """Per-epoch LS from a cold start agrees with the reference's warm-started chain to the
    reference's own stopping tolerance (|dx| < 1e-7 m)."""
    eng = _eng()
    x, it = eng.ls_solve(G["gs_sat_ecef"], G["gs_pr"], G["gs_nsat"])
    x = x.cpu().numpy()
    assert np.max(np.abs(x - G["gs_ls_ecef"])) < 1e-5
    assert np.max(np.abs(x[:, :3] - G["gs_ls_ecef"][:, :3]) / np.abs(G["gs_ls_ecef"][:, :3])) < 1e-9


def test_ls_batch_multi_time_step(G):
    eng = _eng()
    n = G["gs_batch_pr"].shape[0]
    x, it = eng.ls_solve(G["gs_batch_sat"][None], G["gs_batch_pr"][None], np.array([n], dtype=np.int32),
                         tcol=G["gs_batch_t"][None])
    x = x.cpu().numpy()[0]
    ref = G["gs_batch"]
    assert np.max(np.abs(x[:3] - ref[:3])) < 1e-5 and abs(x[3] - ref[3]) < 1e-5 and abs(x[4] - ref[4]) < 1e-7


def test_ls_velocity_two_receiver_logs(G):
    eng = _eng()
    for tag in ("A", "B"):
        sol = G["rx%s_sol" % tag]
        out = eng.ls_vel(G["rx%s_sat" % tag], G["rx%s_vel" % tag], G["rx%s_prr" % tag], G["rx%s_nsat" % tag],
                         np.ascontiguousarray(sol[:, :3])).cpu().numpy()
        assert np.max(np.abs(out - sol[:, 4:8])) < 1e-9


def test_ecef2enu_batch_matches_reference_per_point(G):
    """one launch for all satellites of all epochs == the per-satellite call."""
    import utils.utils as utils
    eng = _eng()
    p_ref = G["p_ref_ECEF"]
    sats = G["gs_sat_ecef"].reshape(-1, 3)
    out = eng.ecef2enu_batch(sats, p_ref).cpu().numpy()
    rot = eng.ecef2enu_batch(sats, p_ref, rotation_only=True).cpu().numpy()
    for i in range(0, sats.shape[0], 37):
        ref = utils.ecef2enu(sats[i], p_ref)
        assert np.max(np.abs(out[i] - ref)) <= 1e-8 * (1 + np.max(np.abs(ref)))          # 2e7 m magnitudes: < 1 um
        assert np.max(np.abs(rot[i] - utils.ecef2enu(sats[i], p_ref, rotation_only=True))) <= 1e-8 * 3e7
    # stored ENU satellites of the golden file were produced by the reference's own ecef2enu
    k = 7
    n = int(G["gs_nsat"][k])
    assert np.max(np.abs(out.reshape(51, 12, 3)[k, :n] - G["gs_sat_enu"][k, :n])) < 1e-7
    assert np.array_equal(utils.ecef2enu_batch(sats[:5], p_ref), out[:5])


def test_enu2lla_batch_matches_reference_per_point(G):
    """estimates ENU -> ECEF -> geodetic in one launch == enu2ecef + ecef2lla per sample."""
    import utils.utils as utils
    eng = _eng()
    p_ref = G["p_ref_ECEF"]
    rng = np.random.default_rng(5)
    enu = rng.normal(size=(4096, 3)) * np.array([300.0, 300.0, 20.0])
    enu[7] = utils.ecef2enu(np.zeros(3), p_ref)                       # maps back to the ECEF origin (to rounding)
    lla, ecef = eng.enu2lla_batch(enu, p_ref, want_ecef=True)
    lla, ecef = lla.cpu().numpy(), ecef.cpu().numpy()
    for i in range(0, 4096, 41):
        e = utils.enu2ecef(enu[i], p_ref)
        ref = utils.ecef2lla(e)
        assert np.max(np.abs(ecef[i] - e)) <= 1e-8                          # metres (6.4e6 m magnitudes: a few ulp)
        assert abs(lla[i, 0] - ref[0]) <= 1e-12 and abs(lla[i, 1] - ref[1]) <= 1e-12   # degrees (1e-7 m)
        assert abs(lla[i, 2] - ref[2]) <= 1e-8
    assert np.all(np.isfinite(lla[np.arange(4096) != 7]))
    assert np.array_equal(utils.enu2lla_batch(enu[:9], p_ref), lla[:9])


def test_coordinate_kernels_match_values_stored_by_the_reference():
    """nlpf_ecef2enu / nlpf_enu2lla against numbers the REFERENCE produced (not this repo's host port):
    data/autonomous-car/filtering/leastsquares.pkl (tests/golden/car.npz["ls"]) stores, per epoch, the LS fix in
    ECEF, the same point through the reference's utils.ecef2enu, and lat/lon/h through its utils.ecef2lla
    (autonomous-car.py:141-160).  911 epochs."""
    eng = _eng()
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "car.npz"))
    ls, p_ref = g["ls"], g["p_ref_ECEF"]
    ecef, enu_ref, lla_ref = ls[:, :3], ls[:, 4:7], ls[:, 7:10]
    enu = eng.ecef2enu_batch(np.ascontiguousarray(ecef), p_ref).cpu().numpy()
    assert np.max(np.abs(enu - enu_ref)) < 1e-8                      # metres; the difference of 6.4e6 m numbers
    lla, back = eng.enu2lla_batch(np.ascontiguousarray(enu_ref), p_ref, want_ecef=True)
    lla, back = lla.cpu().numpy(), back.cpu().numpy()
    assert np.max(np.abs(back - ecef)) < 1e-8                        # ENU -> ECEF returns the stored ECEF fix
    assert np.max(np.abs(lla[:, :2] - lla_ref[:, :2])) < 1e-12       # degrees (1e-7 m on the ground)
    assert np.max(np.abs(lla[:, 2] - lla_ref[:, 2])) < 1e-8          # metres
    # every satellite of the gnss_stationary log: the reference's own ecef2enu output is in the golden file
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ekf_ls.npz"))
    out = eng.ecef2enu_batch(G["gs_sat_ecef"].reshape(-1, 3), G["p_ref_ECEF"]).cpu().numpy().reshape(51, 12, 3)
    for k in range(51):
        n = int(G["gs_nsat"][k])
        assert np.max(np.abs(out[k, :n] - G["gs_sat_enu"][k, :n])) < 1e-7      # 2.6e7 m magnitudes: 4e-15 relative
    # the car's satellites too (autonomous-car.py:152-154)
    out = eng.ecef2enu_batch(np.ascontiguousarray(g["sat_ecef"]), p_ref).cpu().numpy()
    assert np.max(np.abs(out - g["sat_enu"])) < 1e-7


def _rand_spd(rng, n, scale):
    M = rng.normal(size=(n, n))
    return scale * (M @ M.T + n * np.eye(n))


def test_ekf_full_covariances_and_correct_alone(G):
    """utils/ekf.py:43,54 take matrix covariances and :46-61 is callable on its own; the fused kernel takes a full Q
    (n x n), a full R (ns x ns) and a correction-only mode.  Checked against the restated update (oracle/ekf_ls.py)
    and through the reference's class interface."""
    from oracle import ekf_ls
    import utils.ekf as uekf
    import utils.gnss as ugnss
    eng = _eng()
    rng = np.random.default_rng(17)
    B, steps, S = 6, 12, 9
    x0 = G["syn_ekf_x0"][:B].copy()
    u = rng.normal(size=(B, steps, 3))
    z = G["syn_ekf_z"][:B, :steps, :S].copy()
    sat = G["syn_ekf_sat"][:B, :steps, :S].copy()
    nsat = rng.integers(4, S + 1, (B, steps)).astype(np.int32)
    Q, R = _rand_spd(rng, 5, 1e-3), _rand_spd(rng, S, 10.0)
    mu, Sd, traj = eng.ekf_run(eng.EKF_GNSS5, x0, np.tile(np.eye(5), (B, 1, 1)), u, z, sat, nsat, Q, R, 1.0)
    traj, Sd = traj.cpu().numpy(), Sd.cpu().numpy()
    for b in range(B):
        m, Sm = x0[b].copy(), np.eye(5)
        for t in range(steps):
            k = nsat[b, t]
            m, Sm = ekf_ls.ekf_step(0, m, Sm, u[b, t], z[b, t, :k], sat[b, t, :k], Q, R[:k, :k], 1.0)
            assert _relerr(traj[b, t], m) < 1e-9
        assert np.max(np.abs(Sd[b] - Sm)) < 1e-9 * (1 + np.max(np.abs(Sm)))
    # EKF.correct alone, and update() with matrix covariances, through the class interface
    f = uekf.EKF(ugnss.gnss_pos_and_bias, ugnss.multi_pseudorange, x0[0].copy(), np.eye(5))
    Rk = R[:7, :7]
    mu_cor, S_cor = f.correct(x0[0], 2.0 * np.eye(5), z[0, 0, :7], Rk, None, {"sat_pos": sat[0, 0, :7]})
    m_ref, S_ref = ekf_ls.ekf_step(0, x0[0], 2.0 * np.eye(5), None, z[0, 0, :7], sat[0, 0, :7], None, Rk, 0.0, correct_only=True)
    assert _relerr(x0[0] + mu_cor, m_ref) < 1e-9 and np.max(np.abs(2.0 * np.eye(5) + S_cor - S_ref)) < 1e-9
    assert np.array_equal(f.mu, x0[0]) and np.array_equal(f.S, np.eye(5))          # the filter itself is untouched
    f.update(u[0, 0], z[0, 0, :7], Q, Rk, dyn_func_params={"dt": 1.0}, meas_func_params={"sat_pos": sat[0, 0, :7]})
    m_ref, S_ref = ekf_ls.ekf_step(0, x0[0], np.eye(5), u[0, 0], z[0, 0, :7], sat[0, 0, :7], Q, Rk, 1.0)
    assert _relerr(f.mu, m_ref) < 1e-9 and np.max(np.abs(f.S - S_ref)) < 1e-9


def test_two_filters_per_warp_equal_one_filter_per_warp():
    """Above 2368 filters the kernel packs two filters into a warp (16 lanes each, their own synchronisation masks and
    measurement counts); every output element is still produced by the same sequence of operations, so the results of
    a large batch equal those of its filters run as a small batch bit for bit."""
    eng = _eng()
    rng = np.random.default_rng(23)
    Bf, steps, S = 3000, 10, 12
    pos = rng.normal(0, 5.0, (Bf, 3))
    b0, al = rng.uniform(-1e3, 1e3, Bf), rng.normal(-190, 20, Bf)
    dirs = rng.normal(size=(Bf, steps, S, 3)); dirs[..., 2] = np.abs(dirs[..., 2]) + 0.3
    sat = dirs / np.linalg.norm(dirs, axis=3, keepdims=True) * rng.uniform(2.0e7, 2.6e7, (Bf, steps, S, 1))
    t = np.arange(1, steps + 1)
    z = np.linalg.norm(pos[:, None, None, :] - sat, axis=3) + b0[:, None, None] + al[:, None, None] * t[None, :, None] + rng.normal(0, 10, (Bf, steps, S))
    nsat = rng.integers(0, S + 1, (Bf, steps)).astype(np.int32)
    nsat[nsat < 4] = 0
    x0 = np.concatenate([pos + rng.normal(0, 3, (Bf, 3)), b0[:, None], al[:, None]], 1)
    Q = np.array([0.0001, 0.0001, 0.0001, 0.1, 0.001])
    u = np.zeros((Bf, steps, 3))
    big = eng.ekf_run(eng.EKF_GNSS5, x0, np.tile(np.eye(5), (Bf, 1, 1)), u, z, sat, nsat, Q, 100.0, 1.0)
    sel = np.r_[0:60, 1500:1530, 2990:3000]
    small = eng.ekf_run(eng.EKF_GNSS5, x0[sel], np.tile(np.eye(5), (len(sel), 1, 1)), u[sel], z[sel], sat[sel], nsat[sel], Q, 100.0, 1.0)
    assert np.array_equal(big[2].cpu().numpy()[sel], small[2].cpu().numpy())
    assert np.array_equal(big[1].cpu().numpy()[sel], small[1].cpu().numpy())
    assert np.all(np.isfinite(big[2].cpu().numpy()))
