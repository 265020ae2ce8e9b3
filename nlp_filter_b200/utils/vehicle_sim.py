"""Vehicle constants, tyre models and sensor simulators with the reference's names
(utils/vehicle_sim.py:10-94).  Host-side; the constants feed the vehicle kernels."""
import numpy as np

C_AF = 1.1441e5  # N/rad
C_AR = 1.3388e5  # N/rad
MU = 0.75
M = 2009  # kg
D_F = 1.53  # m
D_R = 1.23  # m
I_Z = 2000  # kg m^2
H = 0.25  # m
G = 9.81  # m/s^2


def get_parameters():
    return {"C_AF": C_AF, "C_AR": C_AR, "MU": MU, "M": M, "D_F": D_F, "D_R": D_R, "I_Z": I_Z, "H": H, "G": G}


def tire_model(F_xr, delta, vx, vy, r):
    """Brush tyre model with friction-circle saturation -> [F_yr, F_yf] (vehicle_sim.py:25-47)."""
    Fz = [(M * G * D_F + H * F_xr) / (D_F + D_R), (M * G * D_R - H * F_xr) / (D_F + D_R)]
    slip = [np.arctan2(vy - D_R * r, vx), np.arctan2(vy + D_F * r, vx) - delta]
    stiff = [C_AR, C_AF]
    Fx = [F_xr, 0.0]
    F_y = np.zeros(2)
    for i in range(2):
        cap = np.sqrt(MU ** 2 * Fz[i] ** 2 - Fx[i] ** 2)
        ta = np.tan(slip[i])
        gamma = np.abs((stiff[i] * ta) / (3 * cap))
        if Fx[i] > MU * Fz[i]:
            F_y[i] = 0
        elif gamma < 1:
            F_y[i] = -stiff[i] * ta * (1 - gamma + (1. / 3) * gamma ** 2)
        else:
            F_y[i] = -cap * np.sign(ta)
    return F_y


def linear_tire_model(F_xr, delta, vx, vy, r):
    """F_y = -C_alpha * slip angle, small-angle slips -> [F_yr, F_yf] (vehicle_sim.py:50-58)."""
    return np.array([-C_AR * ((vy - D_R * r) / vx), -C_AF * ((vy + D_F * r) / vx - delta)])


def vehicle_dynamics(x, u, params=None):
    """x = [px, py, psi, vx, vy, r], u = [F_xr, delta]; params["tire_model_func"] (vehicle_sim.py:60-72)."""
    F_y = params["tire_model_func"](u[0], u[1], x[3], x[4], x[5])
    c, s = np.cos(x[2]), np.sin(x[2])
    return np.array([x[3] * c - x[4] * s,
                     x[3] * s + x[4] * c,
                     x[5],
                     (-F_y[1] * np.sin(u[1]) + u[0]) / M + x[5] * x[4],
                     (F_y[1] * np.cos(u[1]) + F_y[0]) / M - x[5] * x[3],
                     (D_F * F_y[1] * np.cos(u[1]) - D_R * F_y[0]) / I_Z])


def GNSS_measurement_sim(t, p_ECEF, sat_pos, R, alpha, b0):
    """One noisy pseudorange with clock bias b0 + alpha t (vehicle_sim.py:74-77)."""
    return np.linalg.norm(p_ECEF - sat_pos) + b0 + alpha * t + np.sqrt(R) * np.random.randn(1)


def compass_measurement_sim(heading, R):
    return heading + np.sqrt(R) * np.random.randn(1)


def gyro_measurement_sim(yaw_rate, R):
    return yaw_rate + np.sqrt(R) * np.random.randn(1)
