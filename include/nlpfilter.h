/* nlpfilter.h -- C ABI of libnlpfilter.so: the B200 (sm_100a) replacement for the numerical
 * back end of kingdwd/nlp-filter's estimator hot path.
 *
 * The reference has no FFI of its own: its hot path runs inside CasADi/IPOPT/MUMPS (NLP) and
 * NumPy/LAPACK (EKF, least squares), reached from plain Python.  Each entry point below names
 * the reference interface whose arithmetic it takes over (paths relative to the reference
 * repository).  Conventions:
 *   - every function returns 0 on success and a negative NLPF_E_* code on failure; it never
 *     throws and never aborts.  nlpf_last_error() gives a thread-local message.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); calls only enqueue
 *     work, they do not synchronise.
 *   - all batch buffers are DEVICE pointers to contiguous fp64 / int32 arrays owned by the
 *     caller; descriptor arrays in nlpf_desc are HOST pointers, copied at create time.
 *   - the unknown vector of one window is z = [V_0 .. V_{K-1}, globals], V_k in R^nv,
 *     nz = K*nv + ng (see nlp_filter_b200/desc.py).
 */
#ifndef NLPFILTER_H
#define NLPFILTER_H

#ifdef __cplusplus
extern "C" {
#endif

#define NLPF_VERSION 100

/* error codes */
#define NLPF_OK 0
#define NLPF_E_ARG (-1)      /* bad argument / unsupported size */
#define NLPF_E_CUDA (-2)     /* CUDA runtime error (message in nlpf_last_error) */
#define NLPF_E_NOGPU (-3)    /* no CUDA device: there is no CPU fallback */

/* dynamics models, nlp/dynamics.py */
#define NLPF_DYN_AFFINE 0        /* f = A x + B u + c0: dynamics.py:4-38,68-115 */
#define NLPF_DYN_VDP 1           /* dynamics.py:61-66 */
#define NLPF_DYN_BICYCLE 2       /* dynamics.py:117-136 */
#define NLPF_DYN_VEHICLE 3       /* dynamics.py:148-164 */
#define NLPF_DYN_VEHICLE_GNSS 4  /* dynamics.py:166-174 */
#define NLPF_DYN_QUADCOPTER 5     /* dynamics.py:40-59 */
#define NLPF_DYN_USER 6           /* any callable the reference would hand to casadi.Function (nlp/nlp.py:211-218):
                                     f and df/d(x,u) are generated as CUDA source and compiled with NVRTC into a copy of
                                     the window kernels (nlpf_problem_set_user_dynamics) */

/* residual (measurement) types, nlp/measurements.py; per-window record = [y, weight, aux..] */
#define NLPF_RT_STATE 0       /* full_state component             measurements.py:4       rec 2 */
#define NLPF_RT_PR 1          /* pseudorange                      measurements.py:56-70   rec 5: +sat_pos[3] */
#define NLPF_RT_PRR 2         /* pseudorange_rate                 measurements.py:72-79   rec 8: +sat_pos[3], sat_vel[3] */
#define NLPF_RT_RANGE 3       /* range between two index sets     measurements.py:13-19   rec 2 */
#define NLPF_RT_RANGE_PT 4    /* range to a point / global var    measurements.py:8-12    rec 5: +point[3] */
#define NLPF_RT_HEADING 5     /* heading between two index sets   measurements.py:30-36   rec 2 */
#define NLPF_RT_HEADING_PT 6  /* heading to a point / global var  measurements.py:25-29   rec 5 */

#define NLPF_LOSS_L2 0        /* weighted_l2_norm,  cost_functions.py:20-22 */
#define NLPF_LOSS_HUBER 1     /* pseudo_huber_loss, cost_functions.py:25-31 */

#define NLPF_NIDX 8
#define NLPF_NDYNPAR 16

/* solve status per window */
#define NLPF_ST_CONVERGED 0
#define NLPF_ST_STALLED 1     /* no damping produced a decrease AND the point is stationary to rounding:
                                 max_i |g_i| / H_ii (the diagonally scaled gradient step) < 1e-6 (1 + |z|_inf) */
#define NLPF_ST_MAXITER 2
#define NLPF_ST_NUMERIC 3     /* non-finite objective at the starting point */
#define NLPF_ST_NOSTEP 4      /* no damping produced a decrease at a point that is NOT stationary (every
                                 factorisation failed or every step was non-finite): a failure, where
                                 opti.solve() of the reference would raise (nlp/nlp.py:82) */
#define NLPF_STALL_GTOL 1e-6
/* defaults of the damping schedule (nlpf_opts.lam0, lam_jump, lam_jump_indef, lam_dec_first <= 0 select them) */
#define NLPF_LAM0 1e-12
#define NLPF_LAM_JUMP 1e-3
#define NLPF_LAM_JUMP_INDEF 0.1
#define NLPF_LAM_DEC_FIRST 1000.0

/* Problem class = everything the reference fixes when a script has finished calling
 * addVariables/addDynamics/addDynamicsCost/addResidualCost/addInitialCost/addVarBounds/
 * addEqConstraint and build() (nlp/nlp.py:18-69,135-186,202-317).  Host pointers. */
typedef struct nlpf_desc {
    int K, n, m, nv, ng;          /* nodes N+1, states, controls, node variables, globals */
    int dyn_id, loss, has_prior;
    int nb, E, M, Ldata;          /* bounded components, epochs, residuals, doubles per window record block */
    double T, delta;
    const double* dyn_par;        /* [NLPF_NDYNPAR] model constants (car: C_AF C_AR M D_F D_R I_Z) */
    const double* A;              /* [n*n]   affine model */
    const double* Bm;             /* [n*max(m,1)] */
    const double* c0;             /* [n] */
    const int* cmap;              /* [n]  state component -> node-variable column */
    const int* ucol;              /* [max(m,1)] control -> node-variable column, or -1: per-window parameter */
    const double* Dt;             /* [K*K] (2/T) D, nlp/collocation.py:42-64 */
    const double* cw;             /* [K]  weight of defect row k ((T/2) w_k), nlp.py:244-245 */
    const double* cq;             /* [K]  quadrature weight of the node (stage) cost, nlp.py:167-168 */
    const double* qinv;           /* [n]  diagonal of Q^-1 */
    const double* pinv;           /* [n]  diagonal of P^-1 (arrival cost, nlp.py:279-286) */
    const double* snode;          /* [nv] diagonal node cost */
    const int* bidx;              /* [nb] bounded node-variable columns (nlp.py:314-317) */
    const double* blo;            /* [nb] may be -inf */
    const double* bhi;            /* [nb] may be +inf */
    const double* Phi;            /* [E*K] Lagrange basis rows at the measurement epochs, nlp.py:264-269 */
    const int* rtype;             /* [M] */
    const int* repoch;            /* [M] row of Phi */
    const int* ridx;              /* [M*NLPF_NIDX] state / global indices, meaning per type */
    const int* roff;              /* [M] offset of the residual's record in the window's data block */
} nlpf_desc;

typedef struct nlpf_opts {
    int max_iter;                 /* total accepted steps per window */
    int max_stage_iter;           /* accepted steps per barrier stage */
    double tol_step;              /* stop when |step|_inf < tol_step (1 + |z|_inf); such a step, when it is the first one tried in
                                     its iteration, is applied without evaluating the objective at the new point */
    double lam0;                  /* damping lam (H + lam D, D = diagonal of the Gauss-Newton part) of the first try: with the
                                     default 1e-12 a Newton step */
    double mu0, mu_min;           /* barrier schedule (only with bounds) */
    double bound_push;            /* initial distance from a bound */
    double tau_ftb;               /* fraction to the boundary */
    /* how the damping moves (cf. IPOPT's inertia correction: first perturbation 1e-4, growing on failure, shrinking after success) */
    double lam_jump;              /* a step that did not decrease the merit:           lam <- max(10 lam, lam_jump) */
    double lam_jump_indef;        /* a pivot that is not positive / a non-finite step: lam <- max(10 lam, lam_jump_indef) */
    double lam_dec_first;         /* accepted at the first value tried: lam <- max(lam / lam_dec_first, 1e-12); after rejections lam / 10 */
    double mu0_warm;              /* nlpf_solve_chain: barrier parameter at which windows 2.. of a run start (they are warm-started from
                                     the previous window; cf. IPOPT warm_start_init_point + mu_init); <= 0: mu0 for every window */
} nlpf_opts;

typedef struct nlpf_problem nlpf_problem;   /* opaque */

#ifndef __CUDACC_RTC__   /* (the window kernels are also compiled at run time with a generated model function: types only there) */
int nlpf_version(void);
const char* nlpf_last_error(void);
int nlpf_device_count(void);

/* replaces: problem construction up to NLP.build(), nlp/nlp.py:61-69 */
int nlpf_problem_create(const nlpf_desc* desc, nlpf_problem** out);
int nlpf_problem_destroy(nlpf_problem* p);
/* packed size of the lower triangle incl. the appended right-hand-side row, doubles */
int nlpf_problem_nz(const nlpf_problem* p);

/* Kernels (1)+(2): residuals, analytic Jacobians, J^T W J and J^T W r of B windows at z.
 * replaces: the CasADi AD evaluation IPOPT requests each iteration for the graph built by
 * nlp/nlp.py:225-235 (defects), :242-245, :264-273, :279-286.
 *   z[B*nz] U[B*K*m] data[B*Ldata] xprior[B*n] wshift[B*K*n] (wshift, U, xprior may be NULL)
 *   H_out[B*nz*nz] full symmetric row-major, g_out[B*nz], cost_out[B*2] = {J, merit} */
int nlpf_assemble(nlpf_problem* p, int B, const double* z, const double* U, const double* data,
                  const double* xprior, const double* wshift, double mu,
                  double* H_out, double* g_out, double* cost_out, void* stream);

/* Kernel (3): dz = -(H + lam (diag H + eps))^-1 g by an in-shared-memory LDL^T.
 * replaces: the MUMPS factorisation inside IPOPT (reached through nlp/nlp.py:82).
 *   H[B*nz*nz] (lower triangle read), g[B*nz]; dz_out[B*nz]; info_out[B] 0 ok / 1 not SPD */
int nlpf_factor_solve(nlpf_problem* p, int B, const double* H, const double* g, double lam,
                      double* dz_out, int* info_out, void* stream);

/* Whole solve of B independent windows, one CTA per window, iterating on chip.
 * replaces: NLP.solve(), nlp/nlp.py:76-83 (opti.solve() -> CasADi -> IPOPT -> MUMPS).
 *   z_inout[B*nz] initial guess in, minimiser out;
 *   cost_out[B] objective J at the last evaluated iterate (a final step below tol_step changes it by O(step^2) and is not
 *   re-evaluated); iters_out[B]; status_out[B] NLPF_ST_* */
int nlpf_solve(nlpf_problem* p, int B, double* z_inout, const double* U, const double* data,
               const double* xprior, const double* wshift, const nlpf_opts* opts,
               double* cost_out, int* iters_out, int* status_out, void* stream);

/* x(t) = sum_j phi_j(t) X_j for B windows at nt times.
 * replaces: NLP.extractSolution, nlp/nlp.py:99-119 (collocation.py:113-121).
 *   PhiT[nt*K] device; out[B*nt*n] */
int nlpf_sample(nlpf_problem* p, int B, const double* z, int nt, const double* PhiT,
                double* out, void* stream);

/* Collocation defects W_k = (2/T) sum_j D_kj X_j - f(X_k, U_k) of B problems: W_out[B*K*n].
 * replaces: the equality rows of fixedTimeOptimalControlNLP.addDynamics, nlp/nlp.py:150-158 (the batched
 * method-of-multipliers driver of the optimal-control class updates its multipliers with them). */
int nlpf_defects(nlpf_problem* p, int B, const double* z, const double* U, double* W_out, void* stream);

/* EKF models (utils/gnss.py, autonomous-car.py:18-79) */
#define NLPF_EKF_GNSS5 0      /* gnss.gnss_pos_and_bias + gnss.multi_pseudorange, n = 5 */
#define NLPF_EKF_CAR9 1       /* discrete_vehicle_dynamics + vehicle_sensors_model, n = 9 */

/* Kernel (4): `steps` fused predict/update steps of B independent filters.
 * replaces: EKF.update/predict/correct, utils/ekf.py:20-61, with the model functions above.
 *   mu_inout[B*n], S_inout[B*n*n];
 *   u[B*steps*m]   (m = 3 for GNSS5, 2 for CAR9)
 *   z[B*steps*smax], sat[B*steps*smax*3], nsat[B*steps] (0 = no measurement: predict only,
 *   ekf.py:32-38); Q[n] diagonal, r scalar pseudorange variance; dt; par[NLPF_NDYNPAR];
 *   mu_traj_out[B*steps*n] state after each step (may be NULL) */
int nlpf_ekf_run(int model, int B, int steps, int smax, double* mu_inout, double* S_inout,
                 const double* u, const double* z, const double* sat, const int* nsat,
                 const double* Q, double r, double dt, const double* par,
                 double* mu_traj_out, void* stream);
/* The same with a FULL process covariance Qfull[n*n] and a full measurement covariance Rfull[smax*smax] (its
 * leading nsat x nsat block is used at every step), as EKF.predict / EKF.correct accept (utils/ekf.py:43,54); and
 * correct_only != 0: EKF.correct alone (utils/ekf.py:46-61) -- (mu, S) in are (mu_pred, S_pred), no prediction. */
int nlpf_ekf_run_general(int model, int B, int steps, int smax, double* mu_inout, double* S_inout, const double* u,
                         const double* z, const double* sat, const int* nsat, const double* Qfull, const double* Rfull,
                         double dt, const double* par, double* mu_traj_out, int correct_only, void* stream);

/* Iterative GNSS least squares of B independent epochs (one solve per epoch).
 * replaces: iterativeLeastSquares, utils/leastsquares.py:19-42 (and, with tcol != NULL,
 * iterativeLeastSquares_multiTimeStep, :66-94, unknowns x,y,z,b0,alpha).
 *   sat[B*smax*3], pr[B*smax], nsat[B]; tcol[B*smax] or NULL; x_inout[B*nu] (nu = 4 or 5)
 *   warm start in, solution out; iters_out[B] */
int nlpf_ls_solve(int B, int smax, const double* sat, const double* pr, const int* nsat,
                  const double* tcol, double* x_inout, int maxiter, double tol,
                  int* iters_out, void* stream);

/* One-shot velocity solve, utils/leastsquares.py:45-63.
 *   sat[B*smax*3], satvel[B*smax*3], prr[B*smax], nsat[B], x[B*3] position; out[B*4] = v, bdot */
int nlpf_ls_vel(int B, int smax, const double* sat, const double* satvel, const double* prr,
                const int* nsat, const double* x, double* out, void* stream);

/* Row A9 (no reference counterpart; extension for multi-segment / long-horizon collocation):
 * solve B SPD block-tridiagonal systems with diagonal blocks Dg[B][nblk][bs*bs], sub-diagonal
 * blocks Lo[B][nblk][bs*bs] (Lo[.][0] ignored; A[i][i-1] = Lo[i], A[i-1][i] = Lo[i]^T), right-hand
 * sides rhs[B][nblk][bs].  variant 0 = block Cholesky / Riccati recursion, 1 = block cyclic
 * reduction.  x_out[B][nblk][bs]; info_out[B] 0 ok / 1 not positive definite.  bs <= 16. */
int nlpf_blocktri_solve(int B, int nblk, int bs, const double* Dg, const double* Lo, const double* rhs,
                        double* x_out, int* info_out, int variant, void* stream);

/* "Next" row: batched ECEF -> ENU of n points about one reference point.
 * replaces: utils.ecef2enu per satellite per epoch, utils/utils.py:58-84 (callers gnss_stationary.py:92-94,
 * autonomous-car.py:152-154).  p_ecef[n*3], out[n*3] device; R9_host = rows of the ECEF->ENU rotation,
 * ref3_host = reference point (HOST pointers); rotation_only != 0 rotates without subtracting (velocities). */
int nlpf_ecef2enu(long long n, const double* p_ecef, const double* R9_host, const double* ref3_host, int rotation_only,
                  double* out, void* stream);

/* "Next" row: batched post-processing of estimates, local ENU -> ECEF -> geodetic.
 * replaces: utils.enu2ecef + utils.ecef2lla per sample, utils/utils.py:87-104, 4-37 (autonomous-car.py:273-287).
 * p_enu[n*3] device; lla_out[n*3] = [lat deg, lon deg, h m]; ecef_out[n*3] or NULL; R9_host / ref3_host as above. */
int nlpf_enu2lla(long long n, const double* p_enu, const double* R9_host, const double* ref3_host, double* ecef_out,
                 double* lla_out, void* stream);

/* Receding-horizon chains of B independent runs in ONE launch (batch form of the window loop of
 * autonomous-car.py:232-288): for w = 0..W-1 solve window w of every run from the previous window's solution,
 * then centre the next arrival cost on that solution sampled at t = DT (phiDT[K], device).  U_all[W][B][K][m],
 * data_all[W][B][Ldata]; z[B][nz] and xprior[B][n] are updated in place; iters_sum_out[B] total iterations,
 * bad_out[B] number of windows that did not converge; Z_out[W][B][nz] every window's estimate, or NULL.
 * Only for problem classes with a warp-per-window kernel and an arrival cost (nlpf_problem_has_chain_kernel);
 * NLPF_E_ARG otherwise (loop nlpf_solve + nlpf_sample on the host instead). */
int nlpf_problem_has_chain_kernel(const nlpf_problem* p);
int nlpf_solve_chain(nlpf_problem* p, int B, int W, double* z, const double* U_all, const double* data_all, double* xprior,
                     const double* phiDT, const nlpf_opts* opts, double* cost_out, int* iters_sum_out, int* bad_out,
                     double* Z_out, void* stream);

/* Row A9, multi-segment collocation for long horizons (no reference counterpart: the reference is single-interval,
 * nlp/nlp.py:189-200, and its monomial Lagrange basis breaks down for N >~ 25, nlp/collocation.py:87-111).  The
 * horizon is S segments of the ordinary single-segment problem class that share their end nodes; unknowns are the
 * S*N_s + 1 nodes.  nlpf_assemble on B*S gathered sub-windows gives H_seg, g_seg, J_seg; nlpf_ms_blocks overlap-adds
 * them into the block-tridiagonal normal equations (bs = N_s*n, S+1 block rows, the last one padded with identity),
 * adds the arrival cost on node 0 (pinv[n], xprior[B][n]; may be NULL) and the damping lam[B] * diag(H) (NULL: none);
 * nlpf_blocktri_solve then solves them.  kn = (N_s+1)*n, step = N_s*n. */
int nlpf_ms_gather(int B, int S, int kn, int step, const double* z, double* zsub, void* stream);
int nlpf_ms_blocks(int B, int S, int n, int bs, const double* H_seg, const double* g_seg, const double* cost_seg,
                   const double* pinv, const double* xprior, const double* z, const double* lam, double* Dg, double* Lo,
                   double* g, double* J, void* stream);

/* Dynamics models outside the kernel library.
 * replaces: casadi.Function('f', [x, u], [func(x, u, params)]) + CasADi's automatic differentiation, nlp/nlp.py:211-218.
 * A problem class created with dyn_id = NLPF_DYN_USER (dyn_par[] = the model's numeric constants) gets its model here:
 * model_src is CUDA source of ONE device function named `nlpf_user_dyn`, returning void, with the parameters
 *     [const double* par, const double* x, const double* u, double* f, bool jac, const FAcc& F, const int* ucol,
 *      double* Fk, int nv]
 * (f = model value; when jac: F.add(i, j, df_i/dx_j), and Fk[i*nv + ucol[j]] += df_i/du_j where ucol[j] >= 0), as
 * nlp_filter_b200/nlp/_codegen.py generates it from a Python callable.  The window kernels are compiled with it at run
 * time (NVRTC, sm_100a; csrc_dir / include_dir = where nls_kernels.inl, common.cuh and this header live) and nlpf_solve /
 * nlpf_assemble use the result.  NLPF_E_ARG with the compiler log in nlpf_last_error() when the source does not compile. */
int nlpf_problem_set_user_dynamics(nlpf_problem* p, const char* model_src, const char* csrc_dir, const char* include_dir);

/* Measurement support (no reference counterpart; bench.py's roofline.eval_frac / ldl_frac): per-phase clock
 * counters of the warp-per-window kernel.  enable != 0 allocates and zeroes five device counters that every
 * following nlpf_solve / nlpf_solve_chain on this handle adds to; out5 (host, may be NULL) receives what has been
 * accumulated so far -- [warp-cycles inside the evaluation (kernels 1+2), warp-cycles inside factorisation + sweep
 * (kernel 3), evaluations, factorisations, warp-cycles spent on windows] -- and synchronises the device;
 * enable == 0 switches the counters off again.  Problem classes served by the CTA-per-window kernel count nothing. */
int nlpf_problem_profile(nlpf_problem* p, int enable, unsigned long long* out5);

/* shared memory (bytes) one window of this problem class needs in the whole-solve kernel */
int nlpf_problem_smem_bytes(const nlpf_problem* p);

/* FP64 FMA throughput probe used for the roofline denominator (bench.py): runs `iters`
 * dependent-chain-free DFMA per thread on a full grid, returns FLOP count in *flops_out. */
int nlpf_fp64_probe(int iters, double* sink_dev, double* flops_out, void* stream);
/* same for the FP64 tensor-core path (mma.sync.m8n8k4.f64) */
int nlpf_dmma_probe(int iters, double* sink_dev, double* flops_out, void* stream);
#endif /* !__CUDACC_RTC__ */

#ifdef __cplusplus
}
#endif
#endif /* NLPFILTER_H */
