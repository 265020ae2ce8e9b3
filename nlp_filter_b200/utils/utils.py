"""Coordinate helpers with the reference's call signatures (utils/utils.py:4-110).

Host-side pass-through (SURVEY.md section 8 marks these "next", not a kernel): scripts call
them once per satellite per epoch.  WGS84, Bowring's iteration for the geodetic latitude with
the reference's stopping rule (<= 10 sweeps, |dlat| < 1e-6), so values agree with the
reference to rounding.
"""
import numpy as np

_WGS84_A = 6378137
_WGS84_F = 1 / 298.257223563


def ecef2lla(p_ECEF):
    """ECEF metres -> [lat deg, lon deg, h m]   (utils/utils.py:4-37).
    The zero vector maps to a tuple of three NaNs like the reference (utils.py:13-14)."""
    x, y, z = p_ECEF[0], p_ECEF[1], p_ECEF[2]
    if x == 0.0 and y == 0.0 and z == 0.0:
        return np.nan, np.nan, np.nan
    a, f = _WGS84_A, _WGS84_F
    e2 = 2 * f - f ** 2
    s = np.sqrt(x ** 2 + y ** 2)
    k = e2 * (1 - f) / (1 - e2)

    def bowring(beta):
        return np.arctan((z + a * np.sin(beta) ** 3 * k) / (s - e2 * a * np.cos(beta) ** 3))

    lat = bowring(np.arctan(z / ((1 - f) * s)))
    for _ in range(10):
        prev = lat
        lat = bowring(np.arctan((1 - f) * np.sin(lat) / (np.cos(lat))))
        if np.abs(prev - lat) < 1e-6:
            break
    rn = a / (np.sqrt(1 - e2 * np.sin(lat) ** 2))
    h = s * np.cos(lat) + (z + e2 * rn * np.sin(lat)) * np.sin(lat) - rn
    return np.array([np.rad2deg(lat), np.rad2deg(np.arctan2(y, x)), h])


def lla2ecef(p_LLA):
    """[lat deg, lon deg, h m] -> ECEF metres   (utils/utils.py:40-55)."""
    lat, lon, h = p_LLA[0], p_LLA[1], p_LLA[2]
    a = _WGS84_A
    finv = 298.257223563
    e2 = 2 * (1 / finv) - (1 / finv) ** 2
    sl, cl = np.sin(np.deg2rad(lat)), np.cos(np.deg2rad(lat))
    rn = a / (np.sqrt(1 - e2 * sl ** 2))
    return np.array([(rn + h) * cl * np.cos(np.deg2rad(lon)),
                     (rn + h) * cl * np.sin(np.deg2rad(lon)),
                     (rn * (1 - e2) + h) * sl])


def _enu_rows(p_ref_ECEF):
    """Rows of the ECEF->ENU rotation at the reference point (utils/utils.py:66-80)."""
    lla = ecef2lla(p_ref_ECEF)
    lat, lon = np.deg2rad(lla[0]), np.deg2rad(lla[1])
    sa, ca, so, co = np.sin(lat), np.cos(lat), np.sin(lon), np.cos(lon)
    R = np.zeros((3, 3))
    R[0, :] = [-so, co, 0.0]
    R[1, :] = [-sa * co, -sa * so, ca]
    R[2, :] = [ca * co, ca * so, sa]
    return R


def ecef2enu(p_ECEF, p_ref_ECEF, rotation_only=False):
    """ECEF -> local ENU at p_ref_ECEF; rotation_only for velocities (utils/utils.py:58-84)."""
    R = _enu_rows(p_ref_ECEF)
    if rotation_only:
        return np.matmul(R, p_ECEF)
    return np.matmul(R, p_ECEF - p_ref_ECEF)


def enu2ecef(p_ENU, p_ref_ECEF):
    """Local ENU -> ECEF   (utils/utils.py:87-104)."""
    R = _enu_rows(p_ref_ECEF)
    return np.matmul(R.T.copy(), p_ENU) + p_ref_ECEF


def get_time_indices(t, t0, tf):
    """Indices with t0 <= t[i] <= tf   (utils/utils.py:107-110)."""
    return np.where((t >= t0) & (t <= tf))[0]


def ecef2enu_batch(p_ECEF, p_ref_ECEF, rotation_only=False):
    """Many points at once on the device (addition to the reference API): `ecef2enu` applied to every
    row of p_ECEF (N, 3) about one reference point.  Returns a NumPy array."""
    from .. import engine
    return engine.ecef2enu_batch(np.asarray(p_ECEF, dtype=float), p_ref_ECEF, rotation_only).cpu().numpy()


def enu2lla_batch(p_ENU, p_ref_ECEF):
    """Many points at once on the device (addition to the reference API): `ecef2lla(enu2ecef(p, ref))` for
    every row of p_ENU (N, 3).  Returns a NumPy array of [lat deg, lon deg, h m] rows."""
    from .. import engine
    return engine.enu2lla_batch(np.asarray(p_ENU, dtype=float), p_ref_ECEF).cpu().numpy()
