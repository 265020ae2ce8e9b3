"""Pins oracle/ekf_ls.py to outputs of the reference's own utils/ekf.py, utils/gnss.py and
utils/leastsquares.py (tests/golden/ekf_ls.npz) and to the reference's stored pickles."""
import os

import numpy as np
import pytest

from oracle import ekf_ls
from tests.helpers import GOLDEN


@pytest.fixture(scope="module")
def G():
    return np.load(os.path.join(GOLDEN, "ekf_ls.npz"))


def test_ls_stationary_log_warm_started_chain(G):
    """gnss_stationary.py:39-50: the reference's mutable default argument makes every epoch start
    from the previous epoch's solution; bias restarts from 0."""
    x = np.zeros(3)
    for k in range(51):
        n = int(G["gs_nsat"][k])
        x, b, it = ekf_ls.iterative_ls(G["gs_sat_ecef"][k, :n], G["gs_pr"][k, :n], x0=x, b0=0.0)
        ref = G["gs_ls_ecef"][k]
        assert np.max(np.abs(x - ref[:3])) < 1e-6 and abs(b - ref[3]) < 1e-6
        assert np.max(np.abs(x - ref[:3]) / np.abs(ref[:3])) < 1e-9


def test_ls_batch_multi_time_step(G):
    x, b0, alpha, it = ekf_ls.iterative_ls(G["gs_batch_sat"], G["gs_batch_pr"], tcol=G["gs_batch_t"])
    ref = G["gs_batch"]
    assert np.max(np.abs(x - ref[:3])) < 1e-5 and abs(b0 - ref[3]) < 1e-5 and abs(alpha - ref[4]) < 1e-7


def test_ls_synthetic_cold_start_and_ragged_counts(G):
    for b in range(64):
        n = int(G["syn_ls_nsat"][b])
        x, bias, it = ekf_ls.iterative_ls(G["syn_ls_sat"][b, :n], G["syn_ls_pr"][b, :n])
        ref = G["syn_ls_sol"][b]
        assert np.max(np.abs(np.append(x, bias) - ref) / (1 + np.abs(ref))) < 1e-9


def test_ls_velocity_two_receiver_logs(G):
    for tag in ("A", "B"):
        sol = G["rx%s_sol" % tag]
        for k in range(0, 107, 7):
            n = int(G["rx%s_nsat" % tag][k])
            v, bd = ekf_ls.ls_velocity(G["rx%s_sat" % tag][k, :n], G["rx%s_vel" % tag][k, :n],
                                       G["rx%s_prr" % tag][k, :n], sol[k, :3])
            assert np.max(np.abs(v - sol[k, 4:7])) < 1e-9 and abs(bd - sol[k, 7]) < 1e-9


def test_ekf_gnss5_stationary_log(G):
    n = G["gs_nsat"]
    traj, S = ekf_ls.ekf_run(0, G["gs_ekf_x0"], np.eye(5), np.zeros((51, 3)), G["gs_pr"], G["gs_sat_enu"], n,
                             np.array([0.0001, 0.0001, 0.0001, 0.1, 0.001]), 100.0, 1.0)
    ref = G["gs_ekf_traj"]
    assert np.max(np.abs(traj - ref) / (1 + np.abs(ref))) < 1e-9
    assert np.max(np.abs(S - G["gs_ekf_S"])) < 1e-9


def test_ekf_gnss5_synthetic_ragged(G):
    for b in range(16):
        traj, S = ekf_ls.ekf_run(0, G["syn_ekf_x0"][b], np.eye(5), G["syn_ekf_u"][b], G["syn_ekf_z"][b],
                                 G["syn_ekf_sat"][b], G["syn_ekf_nsat"][b], G["syn_ekf_Q"], 100.0, 1.0)
        ref = G["syn_ekf_traj"][b]
        assert np.max(np.abs(traj - ref) / (1 + np.abs(ref))) < 1e-9


def test_ekf_car9_against_live_reference_and_stored_pickle(G):
    NST = 1500
    sat = np.broadcast_to(G["car_sat_enu"], (NST, 11, 3))
    traj, S = ekf_ls.ekf_run(1, G["car_x0"], np.eye(9), G["car_u"][:NST], G["car_z"][:NST], sat, G["car_nsat"][:NST],
                             G["car_Q"], float(G["car_r"]), 0.01, G["car_par"])
    ref = G["car_traj"][:NST]
    err = np.abs(traj - ref) / (1 + np.abs(ref))
    # the vehicle filter amplifies rounding (it diverges at k = 5852 in the reference's own run,
    # SURVEY.md D.2): 1e-9 holds over the first 100 steps, 1e-7 over 1500
    assert np.max(err[:100]) < 1e-9 and np.max(err) < 1e-7
    stored = np.load(os.path.join(GOLDEN, "car.npz"))["ekf"]      # data/autonomous-car/filtering/ekf.pkl
    m = np.stack([traj[:, 0], traj[:, 1], traj[:, 8], traj[:, 6], traj[:, 2], traj[:, 3], traj[:, 4]], 1)
    assert np.max(np.abs(m[:-1] - stored[1:NST])) < 2e-8
