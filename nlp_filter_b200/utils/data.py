"""Log loaders with the reference's names (utils/data.py:9-89), python-3 safe.  Host I/O: not a
kernel."""
import pickle

import numpy as np
from scipy.io import loadmat

C = 299792458  # speed of light, m/s


def load_gnss_logs(prefix):
    """prefix + 'satposecef.mat' / 'ranges.mat' -> dict(t, sats, sat_pos, pr[, sat_vel, pr_rate]) with
    one array per epoch; pseudoranges corrected for ionosphere and satellite clock; invalid
    entries (NaN range or all-zero satellite position) dropped (data.py:9-75)."""
    sv = loadmat(prefix + 'satposecef.mat')["svPoss"]
    pos_all, iono, clk = sv[1:, :, :3], sv[1:, :, 3], sv[1:, :, 4]
    rng = loadmat(prefix + 'ranges.mat')["pseudoranges"]
    pos_only = rng.ndim == 2
    if pos_only:
        pr_all = rng[1:, :] + iono + C * clk
        sats = rng[0, :]
        print('Assuming data at 1 Hz with t0 = 0')
        times = range(pr_all.shape[0])
    else:
        pr_all = rng[1:, :, 0] + iono + C * clk
        rate_all, vel_all = rng[1:, :, 1], rng[1:, :, 2:5]
        if rng.shape[2] == 6:
            times = np.max(rng[1:, :, 5], axis=1)
        else:
            print('Assuming data at 1 Hz with t0 = 0')
            times = range(pr_all.shape[0])
        sats = rng[0, :, 0]
    valid = ~np.all(pos_all == 0.0, axis=2) & ~np.isnan(pr_all)
    data = {"t": times, "sats": sats, "sat_pos": [], "pr": []}
    if not pos_only:
        data["sat_vel"], data["pr_rate"] = [], []
    for k in range(pr_all.shape[0]):
        ok = valid[k]
        data["sat_pos"].append(pos_all[k, ok, :].reshape(-1, 3))
        data["pr"].append(pr_all[k, ok])
        if not pos_only:
            data["sat_vel"].append(vel_all[k, ok, :].reshape(-1, 3))
            data["pr_rate"].append(rate_all[k, ok])
    return data


def load_px4_logs(prefix):
    return load_obj(prefix)


def save_obj(obj, fname):
    with open(fname + '.pkl', 'wb') as f:
        pickle.dump(obj, f, pickle.HIGHEST_PROTOCOL)


def load_obj(fname):
    with open(fname + '.pkl', 'rb') as f:
        try:
            return pickle.load(f)
        except UnicodeDecodeError:          # python-2 pickles of the reference's data directory
            f.seek(0)
            return pickle.load(f, encoding='latin1')
