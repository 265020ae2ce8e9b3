"""Extended Kalman filter with the reference's class interface (utils/ekf.py:4-61), running
each update as one fused predict/update launch of kernel (4) (`nlpf_ekf_run`).

The reference calls arbitrary Python model callables; the kernels implement registered model
pairs (below).  An unregistered pair raises -- there is no CPU fallback on this path.  For many
filters at once use `nlp_filter_b200.engine.ekf_run` (B filters x many steps per launch).

Quirks preserved: the mean is propagated in place; a step with z=None only
predicts (ekf.py:32-38); S <- S_pred - K (H S_pred) without symmetrisation (ekf.py:54-59).
"""
import numpy as np

from . import gnss as _gnss

EKF_GNSS5 = 0
EKF_CAR9 = 1

# (dyn_func, meas_func) -> kernel model id
_MODEL_PAIRS = {(_gnss.gnss_pos_and_bias, _gnss.multi_pseudorange): EKF_GNSS5}


def register_model(dyn_func, meas_func, model_id):
    """Declare that a (dynamics, measurement) callable pair is implemented by kernel model
    `model_id` (EKF_GNSS5: utils/gnss.py models; EKF_CAR9: the vehicle model of
    autonomous-car.py:18-79, see nlp_filter_b200.utils.vehicle_ekf)."""
    _MODEL_PAIRS[(dyn_func, meas_func)] = int(model_id)


class EKF(object):
    """Notation of Probabilistic Robotics: x_t, G_t = dyn_func(x_t-1, u_t); z_t, H_t = meas_func(x_t)."""

    def __init__(self, dyn_func, meas_func, mu0, S0):
        self.mu = mu0
        self.S = S0
        self.dynamics = dyn_func
        self.measurement = meas_func

    def _model(self, meas_func):
        key = (self.dynamics, meas_func if meas_func is not None else self.measurement)
        if key not in _MODEL_PAIRS:
            from .._capi import NlpfError
            raise NlpfError("EKF model pair %r is not registered with a CUDA kernel model "
                            "(utils.ekf.register_model); there is no CPU fallback" % (key,))
        return _MODEL_PAIRS[key]

    def update(self, u, z, Q, R, dyn_func_params=None, meas_func=None, meas_func_params=None):
        """One predict (+ correct when z is not None) step (ekf.py:20-38)."""
        from .. import engine
        model = self._model(meas_func)
        n = 5 if model == EKF_GNSS5 else 9
        mu = np.asarray(self.mu, dtype=float).reshape(1, n)
        S = np.asarray(self.S, dtype=float).reshape(1, n, n)
        Qd = np.asarray(Q, dtype=float)
        if Qd.ndim == 2 and not np.any(np.abs(Qd - np.diag(np.diag(Qd))) > 0):
            Qd = np.diag(Qd)                   # diagonal process covariance: the kernel's fast path (a full Q is passed as is)
        if z is not None:
            z = np.asarray(z, dtype=float).reshape(-1)
            ns = z.shape[0]
            Rm = np.asarray(R, dtype=float)
            r = float(Rm[0, 0]) if Rm.ndim == 2 else float(Rm.reshape(-1)[0])
            if Rm.ndim == 2 and np.any(np.abs(Rm - r * np.eye(ns)) > 0):
                r = Rm.reshape(ns, ns)         # general measurement covariance (ekf.py:54)
            elif Rm.ndim == 1 and Rm.size == ns and np.any(Rm != r):
                r = np.diag(Rm)
            sat = np.asarray(meas_func_params["sat_pos"], dtype=float).reshape(1, 1, ns, 3)
            zz = z.reshape(1, 1, ns)
        else:
            ns, r = 0, 1.0
            sat, zz = np.zeros((1, 1, 1, 3)), np.zeros((1, 1, 1))
        dt = float(dyn_func_params["dt"])
        par = None
        if model == EKF_CAR9:
            C = dyn_func_params["car_params"]
            par = [C["C_AF"], C["C_AR"], C["M"], C["D_F"], C["D_R"], C["I_Z"]]
        m = 3 if model == EKF_GNSS5 else 2
        uu = np.asarray(u, dtype=float).reshape(1, 1, m)
        mu_d, S_d, _ = engine.ekf_run(model, mu, S, uu, zz, sat, np.array([[ns]], dtype=np.int32), Qd, r, dt, par=par,
                                      want_traj=False)
        new_mu = mu_d.cpu().numpy()[0]
        if z is None and isinstance(self.mu, np.ndarray) and self.mu.shape == new_mu.shape \
                and self.mu.dtype == new_mu.dtype:
            self.mu[:] = new_mu            # prediction only: the reference's dynamics act on the mean in place
        else:
            self.mu = new_mu               # ekf.py:34 rebinds the mean to a fresh array
        self.S = S_d.cpu().numpy()[0]

    def predict(self, u, Q, dyn_func_params):
        """Prediction only (ekf.py:40-44): returns (mu_pred, S_pred) and, like the reference,
        leaves the propagated mean in self.mu."""
        S_before = self.S
        self.update(u, None, Q, None, dyn_func_params=dyn_func_params)
        S_pred = self.S
        self.S = S_before
        return self.mu, S_pred

    def correct(self, mu_pred, S_pred, z, R, meas_func, meas_func_params):
        """Measurement update alone (ekf.py:46-61): returns the corrections (mu_cor, S_cor) the reference returns,
        mu = mu_pred + mu_cor, S = S_pred + S_cor; the filter's own state is not touched."""
        from .. import engine
        model = self._model(meas_func)
        n = 5 if model == EKF_GNSS5 else 9
        z = np.asarray(z, dtype=float).reshape(-1)
        ns = z.shape[0]
        Rm = np.asarray(R, dtype=float)
        Rf = Rm.reshape(ns, ns) if Rm.ndim == 2 else (np.diag(Rm.reshape(-1)) if Rm.size == ns else float(Rm.reshape(-1)[0]) * np.eye(ns))
        mu0 = np.asarray(mu_pred, dtype=float).reshape(1, n)
        S0 = np.asarray(S_pred, dtype=float).reshape(1, n, n)
        sat = np.asarray(meas_func_params["sat_pos"], dtype=float).reshape(1, 1, ns, 3)
        par = [1.0] * 6 if model == EKF_CAR9 else None
        m = 3 if model == EKF_GNSS5 else 2
        mu_d, S_d, _ = engine.ekf_run(model, mu0.copy(), S0.copy(), np.zeros((1, 1, m)), z.reshape(1, 1, ns), sat,
                                      np.array([[ns]], dtype=np.int32), np.zeros(n), Rf, 0.0, par=par, want_traj=False,
                                      correct_only=True)
        return mu_d.cpu().numpy()[0] - mu0[0], S_d.cpu().numpy()[0] - S0[0]
