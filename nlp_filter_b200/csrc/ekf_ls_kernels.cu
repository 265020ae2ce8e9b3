// ekf_ls_kernels.cu -- kernel (4): fused EKF predict/update, and the GNSS iterative least squares.
//
// EKF: one warp per filter, `steps` consecutive predict/update steps per launch; the mean and
// covariance stay in shared memory for the whole trajectory, only u, z, satellite positions
// stream in and the mean streams out.  Replaces utils/ekf.py:20-61 with the model functions
// utils/gnss.py:4-44,64-76 (5-state GNSS) and autonomous-car.py:18-79 (9-state vehicle),
// quirks included: the dynamics act on the mean in place and the vehicle
// Jacobian is evaluated at the already-propagated state; the covariance update is the plain
// S_pred - K (H S_pred) form without symmetrisation (ekf.py:54-59).
//
// LS: Gauss-Newton on pseudoranges in ECEF (utils/leastsquares.py:19-94); the reference's
// pinv(G) dx (SVD) is replaced by the normal equations G^T G dx = G^T drho solved by Cholesky
// (G has full column rank whenever >= 4 (5) satellites are in view).
#include "common.cuh"
#include <math.h>
#include <stdlib.h>

#define EKF_NMAX 9
#define EKF_SMAX 16
#define EKF_WARPS 4

struct EkfSm {
    double mu[EKF_NMAX];
    double S[EKF_NMAX * EKF_NMAX];
    double Sp[EKF_NMAX * EKF_NMAX];
    double G[EKF_NMAX * EKF_NMAX];
    double T[EKF_NMAX * EKF_NMAX];
    double H[EKF_SMAX * EKF_NMAX];
    double HS[EKF_SMAX * EKF_NMAX];
    double Kt[EKF_SMAX * EKF_NMAX];      // K^T: row s, column i
    double P[EKF_SMAX * (EKF_SMAX + 1)];
    double v[EKF_SMAX];
};

// gnss.gnss_pos_and_bias, utils/gnss.py:64-76: x += dt*[u, x4, 0]; G = I, G[3][4] = dt
__device__ void ekf_dyn_gnss5(double* mu, const double* u, double dt, double* G) {
    for (int i = 0; i < 25; ++i) G[i] = 0.0;
    for (int i = 0; i < 5; ++i) G[i * 5 + i] = 1.0;
    G[3 * 5 + 4] = dt;
    const double b4 = mu[4];
    mu[0] += dt * u[0];
    mu[1] += dt * u[1];
    mu[2] += dt * u[2];
    mu[3] += dt * b4;
    mu[4] += dt * 0.0;
}

// discrete_vehicle_dynamics, autonomous-car.py:18-56 with the linear tyre model
// utils/vehicle_sim.py:50-58,60-72 (no epsilon), explicit Euler; the Jacobian uses the state
// AFTER the update (x is modified in place before J is formed).  par = C_AF C_AR M D_F D_R I_Z
__device__ void ekf_dyn_car9(double* x, const double* u, double dt, const double* par, double* J) {
    const double C_AF = par[0], C_AR = par[1], M = par[2], D_F = par[3], D_R = par[4], I_Z = par[5];
    {
        const double vx = x[3], vy = x[4], r = x[5];
        const double a_r = (vy - D_R * r) / vx;
        const double a_f = (vy + D_F * r) / vx - u[1];
        const double Fyr = -C_AR * a_r, Fyf = -C_AF * a_f;
        double sp, cp, sd, cd;
        sincos(x[2], &sp, &cp);
        sincos(u[1], &sd, &cd);
        double xd[9];
        xd[0] = vx * cp - vy * sp;
        xd[1] = vx * sp + vy * cp;
        xd[2] = r;
        xd[3] = (-Fyf * sd + u[0]) / M + r * vy;
        xd[4] = (Fyf * cd + Fyr) / M - r * vx;
        xd[5] = (D_F * Fyf * cd - D_R * Fyr) / I_Z;
        xd[6] = x[7];
        xd[7] = 0.0;
        xd[8] = 0.0;
        for (int i = 0; i < 9; ++i) x[i] += dt * xd[i];
    }
    for (int i = 0; i < 81; ++i) J[i] = 0.0;
    for (int i = 0; i < 9; ++i) J[i * 9 + i] = 1.0;
    const double ivx = 1.0 / x[3], ivx2 = 1.0 / (x[3] * x[3]);
    const double dFyf_dvx = C_AF * (x[4] + D_F * x[5]) * ivx2;
    const double dFyf_dvy = -C_AF * ivx;
    const double dFyf_dr = -C_AF * D_F * ivx;
    const double dFyr_dvx = C_AR * (x[4] - D_R * x[5]) * ivx2;
    const double dFyr_dvy = -C_AR * ivx;
    const double dFyr_dr = C_AR * D_R * ivx;
    double s2, c2, su, cu;
    sincos(x[2], &s2, &c2);
    sincos(u[1], &su, &cu);
    J[0 * 9 + 2] += dt * (-x[3] * s2 - x[4] * c2);
    J[0 * 9 + 3] += dt * c2;
    J[0 * 9 + 4] += -dt * s2;
    J[1 * 9 + 2] += dt * (x[3] * c2 - x[4] * s2);
    J[1 * 9 + 3] += dt * s2;
    J[1 * 9 + 4] += dt * c2;
    J[2 * 9 + 5] += dt;
    J[3 * 9 + 3] += -(dt / M) * (su * dFyf_dvx);
    J[3 * 9 + 4] += dt * (x[5] - (su * dFyf_dvy) / M);
    J[3 * 9 + 5] += dt * (x[4] - (su * dFyf_dr) / M);
    J[4 * 9 + 3] += dt * ((cu * dFyf_dvx + dFyr_dvx) / M - x[5]);
    J[4 * 9 + 4] += (dt / M) * (cu * dFyf_dvy + dFyr_dvy);
    J[4 * 9 + 5] += dt * ((cu * dFyf_dr + dFyr_dr) / M - x[3]);
    J[5 * 9 + 3] += (dt / I_Z) * (D_F * cu * dFyf_dvx - D_R * dFyr_dvx);
    J[5 * 9 + 4] += (dt / I_Z) * (D_F * cu * dFyf_dvy - D_R * dFyr_dvy);
    J[5 * 9 + 5] += (dt / I_Z) * (D_F * cu * dFyf_dr - D_R * dFyr_dr);
    J[6 * 9 + 7] += dt;
}

// MODEL is a template parameter: n and m are compile-time (index arithmetic of the small matrix products folds,
// loops unroll); the number of warps per CTA is the launch's blockDim (small batches: one filter per CTA so that
// 256 filters spread over all SMs instead of 64 of them).
// TILE = lanes per filter: 32 (one filter per warp: small batches, every filter gets a warp scheduler slot of its own)
// or 16 (two filters per warp: the Cholesky and the substitutions use at most 16 lanes -- EKF_SMAX rows -- so at
// machine-filling batch sizes half of every warp instruction was idle lanes).  A filter's lanes synchronise with
// the mask of their own half, so the two filters of a warp may take different numbers of measurements.
// Qfull / Rfull (may be null): full n x n process covariance and full smax x smax measurement covariance instead of
// the diagonal Q[n] and r I; correct_only: skip the prediction (S_pred = S, mean untouched) -- EKF.correct, ekf.py:46-61.
template <int MODEL, int TILE>
__global__ void __launch_bounds__(EKF_WARPS * 32)
ekf_kernel(int B, int steps, int smax, double* __restrict__ mu_io, double* __restrict__ S_io,
           const double* __restrict__ u, const double* __restrict__ z, const double* __restrict__ sat,
           const int* __restrict__ nsat, const double* __restrict__ Q, double rvar, double dt,
           const double* __restrict__ par, double* __restrict__ traj, const double* __restrict__ Qfull,
           const double* __restrict__ Rfull, int correct_only) {
    extern __shared__ double ekf_sm_raw[];
    const int lane = threadIdx.x & (TILE - 1), slot = threadIdx.x / TILE, wpc = blockDim.x / TILE;
    const unsigned tmask = (TILE == 32) ? 0xffffffffu : (0xffffu << (16 * ((threadIdx.x >> 4) & 1)));
    EkfSm& W = reinterpret_cast<EkfSm*>(ekf_sm_raw)[slot];
    constexpr int model = MODEL;
    constexpr int n = (MODEL == NLPF_EKF_GNSS5) ? 5 : 9;
    constexpr int m = (MODEL == NLPF_EKF_GNSS5) ? 3 : 2;
    // position / bias columns of the pseudorange model inside the state
    // GNSS5: x,y,z,b = 0,1,2,3 (utils/gnss.py:4-21); CAR9: x,y,z,b = 0,1,8,6 (autonomous-car.py:58-79)
    constexpr int ix = 0, iy = 1, iz = (MODEL == NLPF_EKF_GNSS5) ? 2 : 8, ib = (MODEL == NLPF_EKF_GNSS5) ? 3 : 6;
    double cpar[NLPF_NDYNPAR];
    for (int i = 0; i < NLPF_NDYNPAR; ++i) cpar[i] = par ? par[i] : 0.0;

    for (int f = blockIdx.x * wpc + slot; f < B; f += gridDim.x * wpc) {
        for (int i = lane; i < n; i += TILE) W.mu[i] = mu_io[(size_t)f * n + i];
        for (int i = lane; i < n * n; i += TILE) W.S[i] = S_io[(size_t)f * n * n + i];
        __syncwarp(tmask);
        for (int t = 0; t < steps; ++t) {
            const size_t ft = (size_t)f * steps + t;
            // ---- predict (ekf.py:40-44) ----
            if (lane == 0 && !correct_only) {
                double uu[3];
                for (int i = 0; i < m; ++i) uu[i] = u[ft * m + i];
                if (model == NLPF_EKF_GNSS5) ekf_dyn_gnss5(W.mu, uu, dt, W.G);
                else ekf_dyn_car9(W.mu, uu, dt, cpar, W.G);
            }
            __syncwarp(tmask);
            for (int e = lane; e < n * n; e += TILE) {      // T = S G^T
                const int i = e / n, j = e - i * n;
                double s = 0.0;
                for (int k = 0; k < n; ++k) s = fma(W.S[i * n + k], W.G[j * n + k], s);
                W.T[e] = s;
            }
            __syncwarp(tmask);
            for (int e = lane; e < n * n; e += TILE) {      // S_pred = G T + Q
                const int i = e / n, j = e - i * n;
                double s = 0.0;
                for (int k = 0; k < n; ++k) s = fma(W.G[i * n + k], W.T[k * n + j], s);
                W.Sp[e] = s + (Qfull ? Qfull[e] : ((i == j) ? Q[i] : 0.0));
            }
            if (correct_only) {
                __syncwarp(tmask);
                for (int e = lane; e < n * n; e += TILE) W.Sp[e] = W.S[e];
            }
            __syncwarp(tmask);
            const int ns = min(max(nsat[ft], 0), smax);      // counts come from device memory: never index past the slots
            if (ns > 0) {
                // ---- correct (ekf.py:46-61) ----
                for (int s = lane; s < ns; s += TILE) {
                    const double* sp = sat + (ft * smax + s) * 3;
                    const double d0 = W.mu[ix] - sp[0], d1 = W.mu[iy] - sp[1], d2 = W.mu[iz] - sp[2];
                    const double rho = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
                    for (int i = 0; i < n; ++i) W.H[s * n + i] = 0.0;
                    W.H[s * n + ix] = d0 / rho;       // -LoS/|LoS|, LoS = sat - x (gnss.py:16-18)
                    W.H[s * n + iy] = d1 / rho;
                    W.H[s * n + iz] = d2 / rho;
                    W.H[s * n + ib] = 1.0;
                    W.v[s] = z[ft * smax + s] - (rho + W.mu[ib]);
                }
                __syncwarp(tmask);
                for (int e = lane; e < ns * n; e += TILE) {          // HS = H S_pred
                    const int s = e / n, j = e - s * n;
                    double acc = 0.0;
                    for (int k = 0; k < n; ++k) acc = fma(W.H[s * n + k], W.Sp[k * n + j], acc);
                    W.HS[e] = acc;
                }
                __syncwarp(tmask);
                for (int e = lane; e < ns * ns; e += TILE) {         // P = (H S_pred) H^T + R  (ekf.py:54)
                    const int s = e / ns, s2 = e - s * ns;
                    double acc = 0.0;
                    for (int k = 0; k < n; ++k) acc = fma(W.HS[s * n + k], W.H[s2 * n + k], acc);
                    W.P[s * (EKF_SMAX + 1) + s2] = acc + (Rfull ? Rfull[s * smax + s2] : ((s == s2) ? rvar : 0.0));
                }
                __syncwarp(tmask);
                // Cholesky P = L L^T in place (lower), lane = row.  (A register-resident variant -- rows in registers, the
                // eliminated column by shuffles, loops unrolled to EKF_SMAX -- was measured at 0.65x of this one.)
                for (int j = 0; j < ns; ++j) {
                    double djj = sqrt(W.P[j * (EKF_SMAX + 1) + j]);
                    __syncwarp(tmask);
                    if (lane == j) W.P[j * (EKF_SMAX + 1) + j] = djj;
                    if (lane > j && lane < ns) W.P[lane * (EKF_SMAX + 1) + j] /= djj;
                    __syncwarp(tmask);
                    if (lane > j && lane < ns) {
                        const double lij = W.P[lane * (EKF_SMAX + 1) + j];
                        for (int k = j + 1; k <= lane; ++k)
                            W.P[lane * (EKF_SMAX + 1) + k] = fma(-lij, W.P[k * (EKF_SMAX + 1) + j], W.P[lane * (EKF_SMAX + 1) + k]);
                    }
                    __syncwarp(tmask);
                }
                // K^T = P^-1 (H S_pred): one right-hand-side column per lane (K = S_pred H^T P^-1).  (Column-oriented
                // substitutions with lane = row were measured at 0.74x: the pivot's n divisions then run one after the
                // other in a single lane instead of side by side in n lanes.)
                for (int c = lane; c < n; c += TILE) {
                    for (int s = 0; s < ns; ++s) {
                        double acc = W.HS[s * n + c];
                        for (int k = 0; k < s; ++k) acc = fma(-W.P[s * (EKF_SMAX + 1) + k], W.Kt[k * n + c], acc);
                        W.Kt[s * n + c] = acc / W.P[s * (EKF_SMAX + 1) + s];
                    }
                    for (int s = ns - 1; s >= 0; --s) {
                        double acc = W.Kt[s * n + c];
                        for (int k = s + 1; k < ns; ++k) acc = fma(-W.P[k * (EKF_SMAX + 1) + s], W.Kt[k * n + c], acc);
                        W.Kt[s * n + c] = acc / W.P[s * (EKF_SMAX + 1) + s];
                    }
                }
                __syncwarp(tmask);
                // Note: K uses HS^T = S_pred^T H^T; the reference forms S_pred H^T.  S_pred is symmetric up
                // to rounding, so the two differ at the 1e-16 level only.
                double mucor = 0.0;
                if (lane < n) {
                    for (int s = 0; s < ns; ++s) mucor = fma(W.Kt[s * n + lane], W.v[s], mucor);
                }
                for (int e = lane; e < n * n; e += TILE) {           // S = S_pred - K (H S_pred)
                    const int i = e / n, j = e - i * n;
                    double acc = 0.0;
                    for (int s = 0; s < ns; ++s) acc = fma(W.Kt[s * n + i], W.HS[s * n + j], acc);
                    W.S[e] = W.Sp[e] - acc;
                }
                __syncwarp(tmask);
                if (lane < n) W.mu[lane] += mucor;
            } else {
                for (int e = lane; e < n * n; e += TILE) W.S[e] = W.Sp[e];
            }
            __syncwarp(tmask);
            if (traj)
                for (int i = lane; i < n; i += TILE) traj[ft * n + i] = W.mu[i];
        }
        for (int i = lane; i < n; i += TILE) mu_io[(size_t)f * n + i] = W.mu[i];
        for (int i = lane; i < n * n; i += TILE) S_io[(size_t)f * n * n + i] = W.S[i];
        __syncwarp(tmask);
    }
}

static int ekf_launch(int model, int B, int steps, int smax, double* mu_inout, double* S_inout, const double* u,
                      const double* z, const double* sat, const int* nsat, const double* Q, double r, double dt,
                      const double* par, double* mu_traj_out, const double* Qfull, const double* Rfull, int correct_only,
                      void* stream) {
    if (model != NLPF_EKF_GNSS5 && model != NLPF_EKF_CAR9) { nlpf_set_error("unknown EKF model %d", model); return NLPF_E_ARG; }
    if (B < 1 || steps < 1 || smax < 1 || smax > EKF_SMAX) { nlpf_set_error("bad EKF sizes B=%d steps=%d smax=%d (smax<=%d)", B, steps, smax, EKF_SMAX); return NLPF_E_ARG; }
    if (!mu_inout || !S_inout || !u || !z || !sat || !nsat || (!Q && !Qfull)) { nlpf_set_error("null EKF buffer"); return NLPF_E_ARG; }
    if (model == NLPF_EKF_CAR9 && !par) { nlpf_set_error("vehicle EKF needs the car constants"); return NLPF_E_ARG; }
    static bool attr_done = false;
    if (!attr_done) {
        const int full = (int)(sizeof(EkfSm) * EKF_WARPS * 2);
        if (nlpf_check_cuda(cudaFuncSetAttribute(ekf_kernel<NLPF_EKF_GNSS5, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, full), "ekf smem attr")) return NLPF_E_CUDA;
        if (nlpf_check_cuda(cudaFuncSetAttribute(ekf_kernel<NLPF_EKF_CAR9, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, full), "ekf smem attr")) return NLPF_E_CUDA;
        if (nlpf_check_cuda(cudaFuncSetAttribute(ekf_kernel<NLPF_EKF_GNSS5, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, full), "ekf smem attr")) return NLPF_E_CUDA;
        if (nlpf_check_cuda(cudaFuncSetAttribute(ekf_kernel<NLPF_EKF_CAR9, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, full), "ekf smem attr")) return NLPF_E_CUDA;
        attr_done = true;
    }
    cudaStream_t st = (cudaStream_t)stream;
    static const char* force = getenv("NLPF_EKF_TILE");
    // two filters per warp once there are more filters than warp slots the machine keeps resident (148 SMs x 16 warps)
    const bool half = force ? (atoi(force) == 16) : (B > 148 * 16);
    if (!half) {
        // filters per CTA: 1 until every SM has a few CTAs, then up to EKF_WARPS
        int wpc = 1;
        while (wpc < EKF_WARPS && (B + wpc - 1) / wpc > 148 * 4) wpc *= 2;
        const size_t smem = sizeof(EkfSm) * wpc;
        const int blocks = (B + wpc - 1) / wpc;
        if (model == NLPF_EKF_GNSS5)
            ekf_kernel<NLPF_EKF_GNSS5, 32><<<blocks, wpc * 32, smem, st>>>(B, steps, smax, mu_inout, S_inout, u, z, sat, nsat, Q, r, dt, par, mu_traj_out, Qfull, Rfull, correct_only);
        else
            ekf_kernel<NLPF_EKF_CAR9, 32><<<blocks, wpc * 32, smem, st>>>(B, steps, smax, mu_inout, S_inout, u, z, sat, nsat, Q, r, dt, par, mu_traj_out, Qfull, Rfull, correct_only);
    } else {
        const int fpc = EKF_WARPS * 2;                    // filters per CTA
        const size_t smem = sizeof(EkfSm) * fpc;
        const int blocks = (B + fpc - 1) / fpc;
        if (model == NLPF_EKF_GNSS5)
            ekf_kernel<NLPF_EKF_GNSS5, 16><<<blocks, EKF_WARPS * 32, smem, st>>>(B, steps, smax, mu_inout, S_inout, u, z, sat, nsat, Q, r, dt, par, mu_traj_out, Qfull, Rfull, correct_only);
        else
            ekf_kernel<NLPF_EKF_CAR9, 16><<<blocks, EKF_WARPS * 32, smem, st>>>(B, steps, smax, mu_inout, S_inout, u, z, sat, nsat, Q, r, dt, par, mu_traj_out, Qfull, Rfull, correct_only);
    }
    return nlpf_check_cuda(cudaGetLastError(), "ekf_kernel launch");
}

extern "C" int nlpf_ekf_run(int model, int B, int steps, int smax, double* mu_inout, double* S_inout, const double* u,
                            const double* z, const double* sat, const int* nsat, const double* Q, double r, double dt,
                            const double* par, double* mu_traj_out, void* stream) {
    return ekf_launch(model, B, steps, smax, mu_inout, S_inout, u, z, sat, nsat, Q, r, dt, par, mu_traj_out, nullptr, nullptr, 0, stream);
}

extern "C" int nlpf_ekf_run_general(int model, int B, int steps, int smax, double* mu_inout, double* S_inout, const double* u,
                                    const double* z, const double* sat, const int* nsat, const double* Qfull, const double* Rfull,
                                    double dt, const double* par, double* mu_traj_out, int correct_only, void* stream) {
    if (!Qfull || !Rfull) { nlpf_set_error("nlpf_ekf_run_general needs the full Q (n x n) and R (smax x smax)"); return NLPF_E_ARG; }
    return ekf_launch(model, B, steps, smax, mu_inout, S_inout, u, z, sat, nsat, nullptr, 0.0, dt, par, mu_traj_out, Qfull, Rfull, correct_only, stream);
}

// ------------------------------------------------------------------------------------------
// GNSS iterative least squares: one warp per epoch / problem, lanes over satellites.
// ------------------------------------------------------------------------------------------
template <int NU>
__device__ bool chol_solve_small(double* N, double* b) {   // N packed lower NU(NU+1)/2, in place; b -> solution
    // (A reciprocal-square-root form -- one rsqrt per pivot, multiplications instead of the divisions -- ran 1.43x
    // faster, but its extra rounding keeps |dx| hovering around the reference's stopping tolerance at the fixed point:
    // warm restarts then need more than the 3 iterations tests/test_gpu_fullsize.py allows.  Divisions stay.)
    for (int j = 0; j < NU; ++j) {
        double d = N[j * (j + 1) / 2 + j];
        for (int k = 0; k < j; ++k) d -= N[j * (j + 1) / 2 + k] * N[j * (j + 1) / 2 + k];
        if (!(d > 0.0)) return false;
        d = sqrt(d);
        N[j * (j + 1) / 2 + j] = d;
        for (int i = j + 1; i < NU; ++i) {
            double s = N[i * (i + 1) / 2 + j];
            for (int k = 0; k < j; ++k) s -= N[i * (i + 1) / 2 + k] * N[j * (j + 1) / 2 + k];
            N[i * (i + 1) / 2 + j] = s / d;
        }
    }
    for (int i = 0; i < NU; ++i) {
        double s = b[i];
        for (int k = 0; k < i; ++k) s -= N[i * (i + 1) / 2 + k] * b[k];
        b[i] = s / N[i * (i + 1) / 2 + i];
    }
    for (int i = NU - 1; i >= 0; --i) {
        double s = b[i];
        for (int k = i + 1; k < NU; ++k) s -= N[k * (k + 1) / 2 + i] * b[k];
        b[i] = s / N[i * (i + 1) / 2 + i];
    }
    return true;
}

template <int NU>
__global__ void __launch_bounds__(128)
ls_kernel(int B, int smax, const double* __restrict__ sat, const double* __restrict__ pr, const int* __restrict__ nsat,
          const double* __restrict__ tcol, double* __restrict__ x_io, int maxiter, double tol, int* __restrict__ iters_out) {
    const int lane = threadIdx.x & 31;
    const int wglob = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    constexpr int NP = NU * (NU + 1) / 2;
    for (int e = wglob; e < B; e += nwarps) {
        double x[NU];
        for (int i = 0; i < NU; ++i) x[i] = x_io[(size_t)e * NU + i];
        const int ns = min(max(nsat[e], 0), smax);
        const double* se = sat + (size_t)e * smax * 3;
        const double* pe = pr + (size_t)e * smax;
        const double* te = tcol ? tcol + (size_t)e * smax : nullptr;
        int it = 0;
        for (; it < maxiter;) {
            double N[NP], rhs[NU];
            for (int i = 0; i < NP; ++i) N[i] = 0.0;
            for (int i = 0; i < NU; ++i) rhs[i] = 0.0;
            for (int s = lane; s < ns; s += 32) {
                const double d0 = x[0] - se[s * 3 + 0], d1 = x[1] - se[s * 3 + 1], d2 = x[2] - se[s * 3 + 2];
                const double rho = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
                double row[NU];
                row[0] = d0 / rho;      // -(sat - x)/|sat - x|, leastsquares.py:6-16
                row[1] = d1 / rho;
                row[2] = d2 / rho;
                row[3] = 1.0;
                double pred = rho + x[3];
                if (NU == 5) {
                    row[NU - 1] = te[s];
                    pred += x[NU - 1] * te[s];
                }
                const double dr = pe[s] - pred;
                int q = 0;
                for (int i = 0; i < NU; ++i) {
                    for (int j = 0; j <= i; ++j) N[q++] += row[i] * row[j];
                    rhs[i] += row[i] * dr;
                }
            }
            for (int i = 0; i < NP; ++i)
                for (int o = 16; o > 0; o >>= 1) N[i] += __shfl_xor_sync(0xffffffffu, N[i], o);
            for (int i = 0; i < NU; ++i)
                for (int o = 16; o > 0; o >>= 1) rhs[i] += __shfl_xor_sync(0xffffffffu, rhs[i], o);
            ++it;
            if (!chol_solve_small<NU>(N, rhs)) break;        // rank deficient geometry: leave x as is
            double nrm = 0.0;
            for (int i = 0; i < NU; ++i) { x[i] += rhs[i]; nrm += rhs[i] * rhs[i]; }
            if (sqrt(nrm) < tol) break;                      // leastsquares.py:39
        }
        if (lane == 0) {
            for (int i = 0; i < NU; ++i) x_io[(size_t)e * NU + i] = x[i];
            if (iters_out) iters_out[e] = it;
        }
    }
}

// Thread-per-epoch variant for batches that fill the machine.  The warp-per-epoch kernel above spends its time in the
// 14 x 5 shuffle reductions of every Gauss-Newton iteration (profiles/r1_ls_kernel_full.json: 12 of 32 lanes carry a
// satellite) and re-reads the epoch's records from L1/L2 each iteration; measured 139 M epochs/s at 16 384 epochs and
// only 155 M at a million (64 GB/s of HBM).  Here one thread owns an epoch: its <= SMAX satellite records are read ONCE
// (416 contiguous bytes) into registers, the normal equations are summed in the thread, no shuffles.  Same arithmetic
// per satellite and the same Cholesky; only the order of the sums over satellites differs (sequential instead of a
// butterfly), i.e. results agree to rounding.
template <int NU, int SMAX>
__global__ void __launch_bounds__(128)
ls_thread_kernel(int B, int smax, const double* __restrict__ sat, const double* __restrict__ pr, const int* __restrict__ nsat,
                 const double* __restrict__ tcol, double* __restrict__ x_io, int maxiter, double tol, int* __restrict__ iters_out) {
    constexpr int NP = NU * (NU + 1) / 2;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < (size_t)B; e += (size_t)gridDim.x * blockDim.x) {
        double x[NU];
#pragma unroll
        for (int i = 0; i < NU; ++i) x[i] = x_io[e * NU + i];
        const int ns = min(max(nsat[e], 0), smax);
        const double* se = sat + e * smax * 3;
        const double* pe = pr + e * smax;
        double sx[SMAX], sy[SMAX], sz[SMAX], pv[SMAX], tv[(NU == 5) ? SMAX : 1];
#pragma unroll
        for (int s = 0; s < SMAX; ++s) {
            const bool on = s < ns;
            sx[s] = on ? se[s * 3 + 0] : 0.0;
            sy[s] = on ? se[s * 3 + 1] : 0.0;
            sz[s] = on ? se[s * 3 + 2] : 0.0;
            pv[s] = on ? pe[s] : 0.0;
            if (NU == 5) tv[s] = on ? tcol[e * smax + s] : 0.0;
        }
        int it = 0;
        for (; it < maxiter;) {
            double N[NP], rhs[NU];
#pragma unroll
            for (int i = 0; i < NP; ++i) N[i] = 0.0;
#pragma unroll
            for (int i = 0; i < NU; ++i) rhs[i] = 0.0;
#pragma unroll
            for (int s = 0; s < SMAX; ++s) {
                if (s < ns) {
                    const double d0 = x[0] - sx[s], d1 = x[1] - sy[s], d2 = x[2] - sz[s];
                    const double rho = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
                    double row[NU];
                    row[0] = d0 / rho;      // -(sat - x)/|sat - x|, leastsquares.py:6-16
                    row[1] = d1 / rho;
                    row[2] = d2 / rho;
                    row[3] = 1.0;
                    double pred = rho + x[3];
                    if (NU == 5) {
                        row[NU - 1] = tv[s];
                        pred += x[NU - 1] * tv[s];
                    }
                    const double dr = pv[s] - pred;
                    int q = 0;
#pragma unroll
                    for (int i = 0; i < NU; ++i) {
#pragma unroll
                        for (int j = 0; j <= i; ++j) N[q++] += row[i] * row[j];
                        rhs[i] += row[i] * dr;
                    }
                }
            }
            ++it;
            if (!chol_solve_small<NU>(N, rhs)) break;        // rank deficient geometry: leave x as is
            double nrm = 0.0;
#pragma unroll
            for (int i = 0; i < NU; ++i) { x[i] += rhs[i]; nrm += rhs[i] * rhs[i]; }
            if (sqrt(nrm) < tol) break;                      // leastsquares.py:39
        }
#pragma unroll
        for (int i = 0; i < NU; ++i) x_io[e * NU + i] = x[i];
        if (iters_out) iters_out[e] = it;
    }
}

extern "C" int nlpf_ls_solve(int B, int smax, const double* sat, const double* pr, const int* nsat, const double* tcol,
                             double* x_inout, int maxiter, double tol, int* iters_out, void* stream) {
    if (B < 1 || smax < 1 || !sat || !pr || !nsat || !x_inout) { nlpf_set_error("bad LS argument"); return NLPF_E_ARG; }
    const int threads = 128;
    static const char* force = getenv("NLPF_LS_WARP");
    if (B >= 2048 && smax <= 16 && !force) {               // enough epochs for a thread each (16 384 epochs = 128 CTAs)
        const int tb = (B + threads - 1) / threads;
        cudaStream_t st = (cudaStream_t)stream;
        if (smax <= 12) {
            if (tcol) ls_thread_kernel<5, 12><<<tb, threads, 0, st>>>(B, smax, sat, pr, nsat, tcol, x_inout, maxiter, tol, iters_out);
            else ls_thread_kernel<4, 12><<<tb, threads, 0, st>>>(B, smax, sat, pr, nsat, nullptr, x_inout, maxiter, tol, iters_out);
        } else {
            if (tcol) ls_thread_kernel<5, 16><<<tb, threads, 0, st>>>(B, smax, sat, pr, nsat, tcol, x_inout, maxiter, tol, iters_out);
            else ls_thread_kernel<4, 16><<<tb, threads, 0, st>>>(B, smax, sat, pr, nsat, nullptr, x_inout, maxiter, tol, iters_out);
        }
        return nlpf_check_cuda(cudaGetLastError(), "ls_thread_kernel launch");
    }
    int blocks = (int)(((size_t)B * 32 + threads - 1) / threads);
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (tcol) ls_kernel<5><<<blocks, threads, 0, (cudaStream_t)stream>>>(B, smax, sat, pr, nsat, tcol, x_inout, maxiter, tol, iters_out);
    else ls_kernel<4><<<blocks, threads, 0, (cudaStream_t)stream>>>(B, smax, sat, pr, nsat, nullptr, x_inout, maxiter, tol, iters_out);
    return nlpf_check_cuda(cudaGetLastError(), "ls_kernel launch");
}

// one-shot velocity solve, leastsquares.py:45-63
__global__ void __launch_bounds__(128)
ls_vel_kernel(int B, int smax, const double* __restrict__ sat, const double* __restrict__ satvel,
              const double* __restrict__ prr, const int* __restrict__ nsat, const double* __restrict__ xpos,
              double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int wglob = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int e = wglob; e < B; e += nwarps) {
        const int ns = min(max(nsat[e], 0), smax);
        const double* se = sat + (size_t)e * smax * 3;
        const double* ve = satvel + (size_t)e * smax * 3;
        const double* pe = prr + (size_t)e * smax;
        const double x0 = xpos[(size_t)e * 3], x1 = xpos[(size_t)e * 3 + 1], x2 = xpos[(size_t)e * 3 + 2];
        double N[10], rhs[4];
        for (int i = 0; i < 10; ++i) N[i] = 0.0;
        for (int i = 0; i < 4; ++i) rhs[i] = 0.0;
        for (int s = lane; s < ns; s += 32) {
            const double d0 = x0 - se[s * 3], d1 = x1 - se[s * 3 + 1], d2 = x2 - se[s * 3 + 2];
            const double rho = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
            const double row[4] = {d0 / rho, d1 / rho, d2 / rho, 1.0};
            // drho = pr_rate - sat_vel . LoS_dir with LoS_dir = -G[k,:3]
            const double dr = pe[s] - (ve[s * 3] * (-row[0]) + ve[s * 3 + 1] * (-row[1]) + ve[s * 3 + 2] * (-row[2]));
            int q = 0;
            for (int i = 0; i < 4; ++i) {
                for (int j = 0; j <= i; ++j) N[q++] += row[i] * row[j];
                rhs[i] += row[i] * dr;
            }
        }
        for (int i = 0; i < 10; ++i)
            for (int o = 16; o > 0; o >>= 1) N[i] += __shfl_xor_sync(0xffffffffu, N[i], o);
        for (int i = 0; i < 4; ++i)
            for (int o = 16; o > 0; o >>= 1) rhs[i] += __shfl_xor_sync(0xffffffffu, rhs[i], o);
        if (!chol_solve_small<4>(N, rhs)) { for (int i = 0; i < 4; ++i) rhs[i] = nan(""); }
        if (lane == 0) for (int i = 0; i < 4; ++i) out[(size_t)e * 4 + i] = rhs[i];
    }
}

extern "C" int nlpf_ls_vel(int B, int smax, const double* sat, const double* satvel, const double* prr, const int* nsat,
                           const double* x, double* out, void* stream) {
    if (B < 1 || smax < 1 || !sat || !satvel || !prr || !nsat || !x || !out) { nlpf_set_error("bad LS-vel argument"); return NLPF_E_ARG; }
    const int threads = 128;
    int blocks = (int)(((size_t)B * 32 + threads - 1) / threads);
    if (blocks > 148 * 16) blocks = 148 * 16;
    ls_vel_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(B, smax, sat, satvel, prr, nsat, x, out);
    return nlpf_check_cuda(cudaGetLastError(), "ls_vel_kernel launch");
}

// ------------------------------------------------------------------------------------------
// FP64 FMA throughput probe (roofline denominator; MEASURED_PEAKS.json has no FP64 figure)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fp64_probe_kernel(int iters, double* sink) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1.0, a2 = a0 + 2.0, a3 = a0 + 3.0;
    double a4 = a0 + 4.0, a5 = a0 + 5.0, a6 = a0 + 6.0, a7 = a0 + 7.0;
    const double m = 0.999999999, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    const double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 123.456) sink[0] = s;
}

extern "C" int nlpf_fp64_probe(int iters, double* sink_dev, double* flops_out, void* stream) {
    const int blocks = 148 * 8, threads = 256;
    fp64_probe_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(iters, sink_dev);
    if (flops_out) *flops_out = 2.0 * 8.0 * (double)iters * blocks * threads;
    return nlpf_check_cuda(cudaGetLastError(), "fp64_probe launch");
}

// FP64 tensor-core (DMMA, mma.sync.m8n8k4.f64) throughput probe: answers BASELINE.json's "FP64 DMMA only if
// ncu shows the batched J^T J contraction benefits" with a number for this part.
__global__ void __launch_bounds__(256) dmma_probe_kernel(int iters, double* sink) {
    double c[8][2];
#pragma unroll
    for (int q = 0; q < 8; ++q) c[q][0] = c[q][1] = threadIdx.x * 1e-9 + q;
    const double a = 0.999999, b = 1e-3;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int q = 0; q < 8; ++q)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[q][0]), "+d"(c[q][1]) : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) s += c[q][0] + c[q][1];
    if (s == 123.456) sink[0] = s;
}

extern "C" int nlpf_dmma_probe(int iters, double* sink_dev, double* flops_out, void* stream) {
    const int blocks = 148 * 8, threads = 256;
    dmma_probe_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(iters, sink_dev);
    // one m8n8k4 per warp: 8*8*4 FMAs
    if (flops_out) *flops_out = 2.0 * 256.0 * 8.0 * (double)iters * blocks * (threads / 32);
    return nlpf_check_cuda(cudaGetLastError(), "dmma_probe launch");
}

// ------------------------------------------------------------------------------------------
// Batched ECEF -> ENU of satellite positions: the reference calls
// utils.ecef2enu once per satellite per epoch and re-derives the reference latitude by Bowring
// iteration every call (utils/utils.py:58-84; callers gnss_stationary.py:92-94,123).  Here the 3x3
// rotation of the reference point is formed once on the host and n points are rotated per launch:
//   out_i = R (p_i - p_ref)      (rotation_only: out_i = R p_i, for velocities)
// 48 B per point in + out: HBM-bound, one thread per point, grid-stride.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ecef2enu_kernel(size_t n, const double* __restrict__ p, double r00, double r01,
                                                        double r02, double r10, double r11, double r12, double r20,
                                                        double r21, double r22, double x0, double y0, double z0,
                                                        double* __restrict__ out) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double dx = p[3 * i] - x0, dy = p[3 * i + 1] - y0, dz = p[3 * i + 2] - z0;
        out[3 * i] = r00 * dx + r01 * dy + r02 * dz;
        out[3 * i + 1] = r10 * dx + r11 * dy + r12 * dz;
        out[3 * i + 2] = r20 * dx + r21 * dy + r22 * dz;
    }
}

extern "C" int nlpf_ecef2enu(long long n, const double* p_ecef, const double* R9_host, const double* ref3_host,
                             int rotation_only, double* out, void* stream) {
    if (n < 1 || !p_ecef || !R9_host || !ref3_host || !out) { nlpf_set_error("bad ecef2enu argument"); return NLPF_E_ARG; }
    const double* R = R9_host;
    const double x0 = rotation_only ? 0.0 : ref3_host[0], y0 = rotation_only ? 0.0 : ref3_host[1], z0 = rotation_only ? 0.0 : ref3_host[2];
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    ecef2enu_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((size_t)n, p_ecef, R[0], R[1], R[2], R[3], R[4], R[5], R[6], R[7],
                                                                  R[8], x0, y0, z0, out);
    return nlpf_check_cuda(cudaGetLastError(), "ecef2enu_kernel launch");
}

// Batched post-processing of the estimates: local ENU -> ECEF -> geodetic
// [lat deg, lon deg, h m], the per-sample chain enu2ecef + ecef2lla of autonomous-car.py:273-287
// (utils/utils.py:87-104, 4-37: WGS84, Bowring's iteration with the reference's stopping rule: at most 10
// sweeps, |dlat| < 1e-6).  One thread per point; 72 B per point in + out (+24 B for the optional ECEF copy).
__global__ void __launch_bounds__(256) enu2lla_kernel(size_t n, const double* __restrict__ p, double r00, double r01,
                                                       double r02, double r10, double r11, double r12, double r20,
                                                       double r21, double r22, double x0, double y0, double z0,
                                                       double* __restrict__ ecef_out, double* __restrict__ lla_out) {
    const double a = 6378137.0, f = 1.0 / 298.257223563;
    const double e2 = 2.0 * f - f * f;
    const double kk = e2 * (1.0 - f) / (1.0 - e2);
    const double r2d = 57.29577951308232;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double e = p[3 * i], nn = p[3 * i + 1], u = p[3 * i + 2];
        const double x = r00 * e + r10 * nn + r20 * u + x0;          // R^T p + ref
        const double y = r01 * e + r11 * nn + r21 * u + y0;
        const double z = r02 * e + r12 * nn + r22 * u + z0;
        if (ecef_out) { ecef_out[3 * i] = x; ecef_out[3 * i + 1] = y; ecef_out[3 * i + 2] = z; }
        if (x == 0.0 && y == 0.0 && z == 0.0) {                      // utils.py:13-14
            lla_out[3 * i] = lla_out[3 * i + 1] = lla_out[3 * i + 2] = nan("");
            continue;
        }
        const double s = sqrt(x * x + y * y);
        double sb, cb;
        sincos(atan(z / ((1.0 - f) * s)), &sb, &cb);
        double lat = atan((z + a * sb * sb * sb * kk) / (s - e2 * a * cb * cb * cb));
        for (int it = 0; it < 10; ++it) {
            const double prev = lat;
            double sl, cl;
            sincos(lat, &sl, &cl);
            sincos(atan((1.0 - f) * sl / cl), &sb, &cb);
            lat = atan((z + a * sb * sb * sb * kk) / (s - e2 * a * cb * cb * cb));
            if (fabs(prev - lat) < 1e-6) break;
        }
        double sl, cl;
        sincos(lat, &sl, &cl);
        const double rn = a / sqrt(1.0 - e2 * sl * sl);
        lla_out[3 * i] = lat * r2d;
        lla_out[3 * i + 1] = atan2(y, x) * r2d;
        lla_out[3 * i + 2] = s * cl + (z + e2 * rn * sl) * sl - rn;
    }
}

extern "C" int nlpf_enu2lla(long long n, const double* p_enu, const double* R9_host, const double* ref3_host,
                            double* ecef_out, double* lla_out, void* stream) {
    if (n < 1 || !p_enu || !R9_host || !ref3_host || !lla_out) { nlpf_set_error("bad enu2lla argument"); return NLPF_E_ARG; }
    const double* R = R9_host;
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    enu2lla_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((size_t)n, p_enu, R[0], R[1], R[2], R[3], R[4], R[5], R[6], R[7],
                                                                 R[8], ref3_host[0], ref3_host[1], ref3_host[2], ecef_out, lla_out);
    return nlpf_check_cuda(cudaGetLastError(), "enu2lla_kernel launch");
}
