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


def load_gnss_logs_padded(prefix, smax=None):
    """Addition to the reference API: the same logs as `load_gnss_logs`, but as
    padded structure-of-arrays tensors with a validity count -- the layout the batched kernels read
    (`nlpf_ekf_run`, `nlpf_ls_solve`: sat[T][smax][3], z[T][smax], nsat[T]; window records with weight 0 for
    an empty slot) -- built with array operations instead of a Python loop over the epochs.

    Returns dict(t (T,), sats, nsat (T,) int32, sat_pos (T, smax, 3), pr (T, smax)[, sat_vel (T, smax, 3),
    pr_rate (T, smax)]): the valid satellites of every epoch first, in the log's order (exactly the rows of
    load_gnss_logs(prefix)["sat_pos"][k] / ["pr"][k]), zeros behind them."""
    sv = loadmat(prefix + 'satposecef.mat')["svPoss"]
    pos_all, iono, clk = sv[1:, :, :3], sv[1:, :, 3], sv[1:, :, 4]
    rng = loadmat(prefix + 'ranges.mat')["pseudoranges"]
    pos_only = rng.ndim == 2
    if pos_only:
        pr_all = rng[1:, :] + iono + C * clk
        sats = rng[0, :]
        times = np.arange(pr_all.shape[0], dtype=float)
        rate_all = vel_all = None
    else:
        pr_all = rng[1:, :, 0] + iono + C * clk
        rate_all, vel_all = rng[1:, :, 1], rng[1:, :, 2:5]
        times = np.max(rng[1:, :, 5], axis=1) if rng.shape[2] == 6 else np.arange(pr_all.shape[0], dtype=float)
        sats = rng[0, :, 0]
    valid = ~np.all(pos_all == 0.0, axis=2) & ~np.isnan(pr_all)
    T, S = valid.shape
    nsat = valid.sum(axis=1).astype(np.int32)
    if smax is None:
        smax = int(nsat.max()) if T else 0
    if T and int(nsat.max()) > smax:
        raise ValueError("an epoch has %d valid satellites, smax = %d" % (int(nsat.max()), smax))
    # stable partition of every row: valid entries first, log order kept
    order = np.argsort(~valid, axis=1, kind="stable")[:, :smax] if S >= smax else \
        np.pad(np.argsort(~valid, axis=1, kind="stable"), ((0, 0), (0, smax - S)))
    keep = (np.arange(smax)[None, :] < nsat[:, None])
    rows = np.arange(T)[:, None]

    def gather(a):
        out = a[rows, order]
        out = np.where(keep if out.ndim == 2 else keep[:, :, None], out, 0.0)
        return np.ascontiguousarray(out, dtype=np.float64)

    data = {"t": np.asarray(times, dtype=float), "sats": sats, "nsat": nsat, "sat_pos": gather(pos_all), "pr": gather(pr_all)}
    if not pos_only:
        data["sat_vel"], data["pr_rate"] = gather(vel_all), gather(rate_all)
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
