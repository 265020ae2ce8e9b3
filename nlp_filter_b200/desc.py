"""Flat problem-class descriptor: what `nlpf_problem_create` (include/nlpfilter.h) consumes.

A built `nlp.nlp` problem (reference API: nlp/nlp.py:122-317) is lowered by
`nlp_filter_b200.compile` to one `ProblemDesc` (shared by every window of the batch) plus
per-window buffers (node controls, measurement records, arrival-cost centre, iterate).
The descriptor is pure data; it carries no compute.

Unknown vector of one window:  z = [V_0 .. V_{K-1}, globals],  V_k in R^nv,  nz = K*nv + ng.
State component i of node k is V_k[cmap[i]]  (cmap folds linear equalities such as
X_k[2] == X_k[7], gnss-multi-receiver.py:76-78, by sharing a column); control i is the
per-window parameter U[k, i] when ucol[i] < 0 (estimation, nlp.py:211-214) and the unknown
V_k[ucol[i]] otherwise (optimal control, nlp.py:135-158).
"""
import numpy as np

# dynamics model ids (nlp/dynamics.py) -- keep in sync with include/nlpfilter.h
DYN_AFFINE = 0          # f = A x + B u + c0: single/double integrator, gnss_pos_and_bias,
                        # multi_receiver, gnss_two_receiver (dynamics.py:4-38,68-115)
DYN_VDP = 1             # dynamics.py:61-66
DYN_BICYCLE = 2         # kinematic_bycicle_and_bias, dynamics.py:117-136
DYN_VEHICLE = 3         # vehicle_dynamics (n=6), dynamics.py:148-164
DYN_VEHICLE_GNSS = 4    # vehicle_dynamics_and_gnss (n=9), dynamics.py:166-174
DYN_QUADCOPTER = 5      # dynamics.py:40-59

# residual types (nlp/measurements.py)
RT_STATE = 0            # full_state component: h = x[i0]                    measurements.py:4
RT_PR = 1               # pseudorange / vehicle_pseudorange                  measurements.py:56-70,81-87
RT_PRR = 2              # pseudorange_rate                                   measurements.py:72-79
RT_RANGE = 3            # multi_receiver_range_2d/3d between idxA, idxB      measurements.py:13-19,46-53
RT_RANGE_PT = 4         # ... against params["y"] (data point or global unknown) measurements.py:8-12,40-45
RT_HEADING = 5          # multi_receiver_heading_2d between idxA, idxB       measurements.py:30-36
RT_HEADING_PT = 6       # ... against params["y"]                            measurements.py:25-29

# per-window data record length of each residual type: [y, weight, aux...]
RT_RECLEN = {RT_STATE: 2, RT_PR: 5, RT_PRR: 8, RT_RANGE: 2, RT_RANGE_PT: 5,
             RT_HEADING: 2, RT_HEADING_PT: 5}

LOSS_L2 = 0
LOSS_HUBER = 1

NIDX = 8                # index slots per residual
NDYNPAR = 16


class ProblemDesc(object):
    """See module docstring.  All arrays are C-contiguous float64 / int32."""

    def __init__(self, K, n, m, T):
        self.K = int(K)
        self.n = int(n)
        self.m = int(m)
        self.T = float(T)
        self.nv = self.n                     # node variables (after equality folding / +controls)
        self.ng = 0                          # global (non-node) unknowns
        self.dyn_id = DYN_AFFINE
        self.dyn_par = np.zeros(NDYNPAR)
        self.A = np.zeros((self.n, self.n))
        self.Bm = np.zeros((self.n, max(self.m, 1)))
        self.c0 = np.zeros(self.n)
        self.cmap = np.arange(self.n, dtype=np.int32)
        self.ucol = -np.ones(max(self.m, 1), dtype=np.int32)
        self.Dt = np.zeros((self.K, self.K))  # (2/T) D
        self.cw = np.zeros(self.K)            # weight of defect row k  ((T/2) w_k for MHE)
        self.cq = np.zeros(self.K)            # quadrature weight of the node cost
        self.qinv = np.zeros(self.n)          # diagonal of Q^-1 (zero => no dynamics cost)
        self.loss = LOSS_L2
        self.delta = 1.0
        self.has_prior = 0
        self.pinv = np.zeros(self.n)
        self.snode = np.zeros(self.nv)        # node quadratic cost diag (stage cost of OCPs)
        self.bidx = np.zeros(0, dtype=np.int32)
        self.blo = np.zeros(0)
        self.bhi = np.zeros(0)
        self.E = 0
        self.Phi = np.zeros((0, self.K))
        self.rtype = np.zeros(0, dtype=np.int32)
        self.repoch = np.zeros(0, dtype=np.int32)
        self.ridx = np.zeros((0, NIDX), dtype=np.int32)
        self.roff = np.zeros(0, dtype=np.int32)
        self.Ldata = 0

    # ---- derived sizes -------------------------------------------------------------------
    @property
    def nz(self):
        return self.K * self.nv + self.ng

    @property
    def M(self):
        return int(self.rtype.shape[0])

    @property
    def nb(self):
        return int(self.bidx.shape[0])

    @property
    def m_param(self):
        return int(np.sum(self.ucol[:self.m] < 0)) if self.m > 0 else 0

    def resize_nv(self, nv):
        self.nv = int(nv)
        self.snode = np.zeros(self.nv)

    def set_residuals(self, rtype, repoch, ridx):
        """Install the residual table and lay the per-window data records out back to back."""
        self.rtype = np.ascontiguousarray(rtype, dtype=np.int32).reshape(-1)
        self.repoch = np.ascontiguousarray(repoch, dtype=np.int32).reshape(-1)
        self.ridx = np.ascontiguousarray(ridx, dtype=np.int32).reshape(-1, NIDX)
        off = np.zeros(self.M, dtype=np.int32)
        pos = 0
        for i in range(self.M):
            off[i] = pos
            pos += RT_RECLEN[int(self.rtype[i])]
        self.roff = off
        self.Ldata = pos

    def finalize(self):
        """Normalise dtypes/shapes (idempotent)."""
        f8 = lambda a, shp: np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(shp))
        i4 = lambda a, shp: np.ascontiguousarray(np.asarray(a, dtype=np.int32).reshape(shp))
        mm = max(self.m, 1)
        self.dyn_par = f8(self.dyn_par, (NDYNPAR,))
        self.A = f8(self.A, (self.n, self.n))
        self.Bm = f8(self.Bm, (self.n, mm))
        self.c0 = f8(self.c0, (self.n,))
        self.cmap = i4(self.cmap, (self.n,))
        self.ucol = i4(self.ucol, (mm,))
        self.Dt = f8(self.Dt, (self.K, self.K))
        self.cw = f8(self.cw, (self.K,))
        self.cq = f8(self.cq, (self.K,))
        self.qinv = f8(self.qinv, (self.n,))
        self.pinv = f8(self.pinv, (self.n,))
        self.snode = f8(self.snode, (self.nv,))
        self.bidx = i4(self.bidx, (-1,))
        self.blo = f8(self.blo, (-1,))
        self.bhi = f8(self.bhi, (-1,))
        self.Phi = f8(self.Phi, (self.E, self.K))
        return self


class SolveOpts(object):
    """Options of the damped Gauss-Newton / log-barrier iteration (see DESIGN.md "Solver")."""

    def __init__(self, max_iter=300, tol_step=1e-9, lam0=1e-3, mu0=0.1, mu_min=1e-9,
                 bound_push=0.01, tau_ftb=0.99, max_stage_iter=300):
        self.max_iter = int(max_iter)
        self.tol_step = float(tol_step)
        self.lam0 = float(lam0)
        self.mu0 = float(mu0)
        self.mu_min = float(mu_min)
        self.bound_push = float(bound_push)
        self.tau_ftb = float(tau_ftb)
        self.max_stage_iter = int(max_stage_iter)
