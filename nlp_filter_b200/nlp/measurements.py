"""Measurement models h(x, params) under the reference's names (nlp/measurements.py:4-87).
Numeric evaluation here; the GPU residual types with analytic gradients are selected through
`KERNEL_TYPES` (csrc/nls_kernels.cu: meas_eval)."""
import numpy as np

from ._sym import vertcat, sin, cos, tan, sqrt, dot, norm_2, atan2, _num
from ..desc import RT_STATE, RT_PR, RT_PRR, RT_RANGE, RT_RANGE_PT, RT_HEADING, RT_HEADING_PT

_SOFT = .000001          # keeps sqrt differentiable at zero distance (measurements.py:13,19)

KERNEL_TYPES = {}


def _against_point(params, default_idx):
    idx = params["idx"] if "idx" in params else default_idx
    return idx, _num(params["y"])


def full_state(x, params=None):
    """y = x   (measurements.py:4-5)."""
    return x


def multi_receiver_range_2d(x, params=None):
    """Planar distance to params["y"] (data or a decision variable) or between the index sets
    idxA / idxB   (measurements.py:7-20)."""
    if "y" in params:
        idx, y = _against_point(params, [0, 1])
        return sqrt((x[idx[0]] - y[0]) ** 2 + (x[idx[1]] - y[1]) ** 2 + _SOFT)
    if "idxA" in params and "idxB" in params:
        a, b = params["idxA"], params["idxB"]
        return sqrt((x[a[0]] - x[b[0]]) ** 2 + (x[a[1]] - x[b[1]]) ** 2 + _SOFT)


def multi_receiver_heading_2d(x, params=None):
    """Heading of B seen from A, atan2(r_x, r_y): measured from the +y axis
    (measurements.py:22-37)."""
    if "y" in params:
        idx, y = _against_point(params, [0, 1])
        r_y = y[1] - x[idx[1]]
        r_x = y[0] - x[idx[0]]
    elif "idxA" in params and "idxB" in params:
        a, b = params["idxA"], params["idxB"]
        r_y = x[b[1]] - x[a[1]]
        r_x = x[b[0]] - x[a[0]] + .00001
    return atan2(r_x, r_y)


def multi_receiver_range_3d(x, params=None):
    """3-D version of multi_receiver_range_2d   (measurements.py:39-54)."""
    if "y" in params:
        idx, y = _against_point(params, [0, 1, 2])
        return sqrt((x[idx[0]] - y[0]) ** 2 + (x[idx[1]] - y[1]) ** 2 + (x[idx[2]] - y[2]) ** 2 + _SOFT)
    if "idxA" in params and "idxB" in params:
        a, b = params["idxA"], params["idxB"]
        return sqrt((x[a[0]] - x[b[0]]) ** 2 + (x[a[1]] - x[b[1]]) ** 2 + (x[a[2]] - x[b[2]]) ** 2 + _SOFT)


def pseudorange(x, params=None):
    """|p - sat_pos| + b with [x, y, z, b] at params["idx"] (default 0..3)   (measurements.py:56-70)."""
    idx = params["idx"] if "idx" in params else [0, 1, 2, 3]
    s = _num(params["sat_pos"])
    return sqrt((x[idx[0]] - s[0]) ** 2 + (x[idx[1]] - s[1]) ** 2 + (x[idx[2]] - s[2]) ** 2) + x[idx[3]]


def pseudorange_rate(x, params=None):
    """(v_sat - v) . LoS + clock drift, x = [p(3), b, v(3), alpha, ...]   (measurements.py:72-79)."""
    r = _num(params["sat_pos"]) - x[:3]
    los = r / norm_2(r)
    return dot(_num(params["sat_vel"]) - x[4:7], los) + x[7]


def vehicle_pseudorange(x, params=None):
    """pseudorange for x = [px, py, psi, vx, vy, psid, b, bd, pz]   (measurements.py:81-87)."""
    s = _num(params["sat_pos"])
    return sqrt((x[0] - s[0]) ** 2 + (x[1] - s[1]) ** 2 + (x[8] - s[2]) ** 2) + x[6]


KERNEL_TYPES.update({
    full_state: "state",
    pseudorange: "pr",
    vehicle_pseudorange: "vpr",
    pseudorange_rate: "prr",
    multi_receiver_range_2d: "range2",
    multi_receiver_range_3d: "range3",
    multi_receiver_heading_2d: "heading",
})
