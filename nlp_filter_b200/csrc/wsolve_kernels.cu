// wsolve_kernels.cu -- warp-per-window whole-solve kernel for SMALL estimation windows with a
// uniform measurement layout: the van der Pol MHE shape of BASELINE.json configs[1], the
// estimation_example.py pattern, the stationary-receiver GNSS shape (configs[2]) and the vehicle
// shape of autonomous-car.py (configs[3]: arrival cost, bounds through the log barrier, L2 or
// pseudo-Huber loss).
//
// Why a second kernel (profiles/r1_vdp_general_kernel_full.json, CTA-per-window kernel): ~27 k
// useful FMAs per iteration were buried under 40 k warp instructions -- runtime sizes and index
// tables, 128 threads of which <= 42 had work in most phases, ~70 block barriers per iteration, a
// 64-register cap.  Here every size is a template parameter (loops unroll, indices fold), one
// warp owns one window (only __syncwarp), windows are handed out through an atomic counter (no
// tail behind slow windows), and <= 16 warps/SM leave 128 registers per thread.  A receding-horizon
// run (W chained windows) stays in its warp from the first window to the last (WsChain).
//
// Same algorithm and decisions as solve_kernel (nls_kernels.inl) and oracle/nls.py: damped
// Gauss-Newton with the exact second-order dynamics term where the model has one, monotone log
// barrier for the bounds, LDL^T with the right-hand side appended, rounding-bound step
// acceptance.  Restrictions (checked on the host, anything else goes to the general kernel): no
// global unknowns, identity cmap, parameter controls, every epoch carries the same record set
// (NM state components, or NM pseudorange slots with one index quadruple).
#include "common.cuh"
#include <math.h>
#include <stdlib.h>

#ifndef WS_LIN_NEWTON
#define WS_LIN_NEWTON 1   // Newton steps after the 2^-20 reciprocal seed of the LDL pivots.  2 = full double precision (all
                          // pseudorange / vehicle classes).  The van der Pol and estimation_example classes (cond(H) ~ 1e4) use 1:
                          // pivots to 1e-12 relative perturb the STEP by that much, the iteration converges to the same
                          // point (the gradient decides, not the factorisation); +2.5 % on the van der Pol batch.
#endif
#ifndef WS_SOA
#define WS_SOA 0          // structure-of-arrays copy of the pseudorange records before the solve (ws_soa_kernel).  Measured:
                          // +0.6 % on the GNSS batch when a window took 8 iterations, -10 % now that it takes 2 (the copy
                          // of 103 MB is a fixed 0.04 ms per call): off; -DWS_SOA=1 builds it in
#endif
#ifndef WS_LINQ
#define WS_LINQ 1         // linear measurements with constant weights as a quadratic form (see ws_eval)
#endif
#ifndef WS_BACKSUB_REG_MAXNZ
#define WS_BACKSUB_REG_MAXNZ 48   // vehicle class (nz = 54): rolled sweep, +2 % (instruction fetch)
#endif
#ifndef WS_UNROLL_ROUNDS
#define WS_UNROLL_ROUNDS 1
#endif
#if WS_UNROLL_ROUNDS
#define WS_ROUND_PRAGMA _Pragma("unroll (C::RU)")
#else
#define WS_ROUND_PRAGMA _Pragma("unroll 1")
#endif
#ifndef WS_DMMA
#define WS_DMMA 1         // measurement Hessian of the pseudorange shapes on the FP64 tensor cores (mma.sync m8n8k4).
                          // Measured on B200: DMMA peak 36.7 TFLOP/s = the vector FMA peak (36.1) -- same pipe -- so the
                          // gain is issue slots only: GNSS batch 3.17-3.18 M windows/s with it, 3.12-3.13 M without
                          // (+1.6 %, two alternating runs each; neutral when first tried on a slower kernel); vehicle
                          // chains unchanged.  BASELINE.json: "DMMA only if the contraction benefits" -- it does, a little.
#endif

namespace {

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double wmax(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// row-padded packed lower triangle: rows have even length, so every row starts 16-byte aligned and pairs of
// entries at even columns load as double2
__device__ __forceinline__ constexpr int po(int r) { return 2 * ((r + 1) >> 1) * ((r + 2) >> 1); }
__device__ __forceinline__ constexpr int pk(int r, int c) { return po(r) + c; }
// 16-byte asynchronous global -> shared copy (LDGSTS): the constant part of H streams in while the node
// phases of the evaluation run.  CG: through L2 only (per-window cache, no reuse from L1).
template <bool CG>
__device__ __forceinline__ void cp_async16(double* dst_smem, const double* src) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(dst_smem);
    if (CG) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(src) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// reciprocal without the slow path of the IEEE division: MUFU seed (2^-20) + two Newton steps (pivots are
// checked for d > 0 and finiteness by the caller)
template <int NEWTON = 2>
__device__ __forceinline__ double frcp(double d) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    r = fma(fma(-d, r, 1.0), r, r);
    if (NEWTON >= 2) r = fma(fma(-d, r, 1.0), r, r);
    return r;
}

// ---- models (nlp/dynamics.py), compile-time sized; F is the N x N Jacobian, row-major ----
template <int DYN, int N, int M>
__device__ __forceinline__ void ws_dyn(const DevProb& P, const double* x, const double* u, double* f, double* F) {
#pragma unroll
    for (int i = 0; i < N * N; ++i) F[i] = 0.0;
    if constexpr (DYN == NLPF_DYN_VDP) {               // dynamics.py:61-66
        f[0] = (1.0 - x[1] * x[1]) * x[0] - x[1] + u[0];
        f[1] = x[0];
        F[0] = 1.0 - x[1] * x[1];
        F[1] = -2.0 * x[0] * x[1] - 1.0;
        F[N] = 1.0;
    } else if constexpr (DYN == NLPF_DYN_VEHICLE_GNSS) {   // dynamics.py:148-174; par = C_AF C_AR M D_F D_R I_Z
        const double C_AF = P.dyn_par[0], C_AR = P.dyn_par[1], Mv = P.dyn_par[2], D_F = P.dyn_par[3], D_R = P.dyn_par[4],
                     I_Z = P.dyn_par[5];
        const double psi = x[2], vx = x[3], vy = x[4], r = x[5];
        const double Fx = u[0], dl = u[1];
        const double den = vx + 0.001;
        const double Fyr = -C_AR * (vy - D_R * r) / den;
        const double Fyf = -C_AF * ((vy + D_F * r) / den - dl);
        double sp, cp, sd, cd;
        sincos(psi, &sp, &cp);
        sincos(dl, &sd, &cd);
        f[0] = vx * cp - vy * sp;
        f[1] = vx * sp + vy * cp;
        f[2] = r;
        f[3] = (-Fyf * sd + Fx) / Mv + r * vy;
        f[4] = (Fyf * cd + Fyr) / Mv - r * vx;
        f[5] = (D_F * Fyf * cd - D_R * Fyr) / I_Z;
        f[6] = x[7];
        f[7] = 0.0;
        f[8] = 0.0;
        const double den2 = den * den;
        const double a_r = C_AR * (vy - D_R * r) / den2, b_r = -C_AR / den, c_r = C_AR * D_R / den;
        const double a_f = C_AF * (vy + D_F * r) / den2, b_f = -C_AF / den, c_f = -C_AF * D_F / den;
        F[0 * N + 2] = -vx * sp - vy * cp;
        F[0 * N + 3] = cp;
        F[0 * N + 4] = -sp;
        F[1 * N + 2] = vx * cp - vy * sp;
        F[1 * N + 3] = sp;
        F[1 * N + 4] = cp;
        F[2 * N + 5] = 1.0;
        F[3 * N + 3] = -sd * a_f / Mv;
        F[3 * N + 4] = -sd * b_f / Mv + r;
        F[3 * N + 5] = -sd * c_f / Mv + vy;
        F[4 * N + 3] = (cd * a_f + a_r) / Mv - r;
        F[4 * N + 4] = (cd * b_f + b_r) / Mv;
        F[4 * N + 5] = (cd * c_f + c_r) / Mv - vx;
        F[5 * N + 3] = (D_F * cd * a_f - D_R * a_r) / I_Z;
        F[5 * N + 4] = (D_F * cd * b_f - D_R * b_r) / I_Z;
        F[5 * N + 5] = (D_F * cd * c_f - D_R * c_r) / I_Z;
        F[6 * N + 7] = 1.0;
    } else {                                           // affine: f = A x + B u + c0
#pragma unroll
        for (int i = 0; i < N; ++i) {
            double s = P.c0[i];
#pragma unroll
            for (int j = 0; j < N; ++j) {
                s = fma(P.A[i * N + j], x[j], s);
                F[i * N + j] = P.A[i * N + j];
            }
#pragma unroll
            for (int j = 0; j < M; ++j) s = fma(P.Bm[i * M + j], u[j], s);
            f[i] = s;
        }
    }
}

// structural sparsity of the model Jacobian F = df/dx (row r, column c), known at compile time.  The vehicle
// model has 17 non-zeros of 81; its Hessian terms only visit the entries they can change.
template <int DYN, int N>
struct FPat {
    static constexpr bool SPARSE = false;
    __host__ __device__ static constexpr bool nz(int, int) { return true; }
};
template <>
struct FPat<NLPF_DYN_VEHICLE_GNSS, 9> {
    static constexpr bool SPARSE = true;
    __host__ __device__ static constexpr bool nz(int r, int c) {       // the entries ws_dyn writes (dynamics.py:148-174)
        return (r == 0 || r == 1) ? (c >= 2 && c <= 4) : (r == 2) ? (c == 5) : (r >= 3 && r <= 5) ? (c >= 3 && c <= 5)
                                                                             : (r == 6) ? (c == 7) : false;
    }
    // pairs (a, b <= a) of columns that share a non-zero row: the only entries of F^T Lambda F; a | b << 4, a byte each
    static constexpr int NFTF = 10;
    static constexpr unsigned long long FTF_LO = 0x22ULL | (0x23ULL << 8) | (0x24ULL << 16) | (0x33ULL << 24) | (0x34ULL << 32) |
                                                 (0x44ULL << 40) | (0x35ULL << 48) | (0x45ULL << 56);
    static constexpr unsigned long long FTF_HI = 0x55ULL | (0x77ULL << 8);
};
constexpr bool fpat_check_vehicle() {               // FTF list == { (a,b): exists i with nz(i,a) && nz(i,b) }
    typedef FPat<NLPF_DYN_VEHICLE_GNSS, 9> P;
    int cnt = 0;
    for (int a = 0; a < 9; ++a)
        for (int b = 0; b <= a; ++b) {
            bool sh = false;
            for (int i = 0; i < 9; ++i) sh = sh || (P::nz(i, a) && P::nz(i, b));
            bool listed = false;
            for (int p = 0; p < P::NFTF; ++p) {
                const unsigned code = (unsigned)(((p < 8) ? (P::FTF_LO >> (8 * p)) : (P::FTF_HI >> (8 * (p - 8)))) & 255);
                if ((int)(code & 15) == a && (int)(code >> 4) == b) listed = true;
            }
            if (sh != listed) return false;
            cnt += sh ? 1 : 0;
        }
    return cnt == P::NFTF;
}
static_assert(fpat_check_vehicle(), "F^T Lambda F pair list does not match the Jacobian pattern");

// compile-time description of one problem class
template <int K_, int N_, int M_, int E_, int NM_, int DYN_, int MEAS_, int WARPS_, int NB_ = 0, bool PRIOR_ = false,
          int LOSS_ = NLPF_LOSS_L2, bool MMA_ = false, int MAXW_ = 16, bool AOS_ = false>
struct WsCfg {
    static constexpr bool MMA = MMA_;       // LDL^T with the matrix in registers as FP64 tensor-core tiles (ws_ldl_mma)
    static constexpr int MAXW = MAXW_;      // warps per SM the register budget allows (16: 128 registers, 12: 168)
    // longest-first hand-out by start cost (ws_key_kernel): measured +16 % on the van der Pol batch (iteration counts
    // 5..17, predicted by J(z0)), -5 % on the GNSS batch (cold start: J(z0) is the squared start offset and says nothing
    // about the 7..9 iterations) -- so only for the linear-measurement classes
    static constexpr bool LPT = (MEAS_ == 0);
    // rounds of 32 lanes of the evaluation loops unrolled (their dependent chains interleave): GNSS +1.8 %, van der Pol
    // neutral; not for the vehicle class, whose kernel is bound by instruction fetch (160 KB of SASS)
    static constexpr int RU = (DYN_ == NLPF_DYN_VEHICLE_GNSS) ? 1 : 4;
    // pseudorange records as structure-of-arrays [field][slot][epoch] per window (ws_soa_kernel makes the copy once per
    // call): the evaluation has one lane per epoch, so a warp's loads of one field of one slot are then consecutive
    // 8-byte words (2 x 128 B per request) instead of 32 words 480 B apart (the caller's record layout [y, w, sat xyz]
    // per residual, which every evaluation -- about ten per window -- re-read at 32 sectors per request)
    // Measured (B200, round 2, alternating runs): GNSS batch 3.10 -> 3.12 M windows/s (+0.6 %, copy included) -- the records
    // were already served from L1 / L2 after the first touch; vehicle chains (21 epochs: 21 of 32 lanes, and a 495 MB
    // copy per step) 539 k -> 529 k (-2 %).  So: on for classes with at least a warp of epochs, off otherwise.
    static constexpr bool SOA = (MEAS_ == 1) && (E_ >= 32) && (WS_SOA != 0) && !AOS_;      // AOS_: single-evaluation kernels (assemble)
    static constexpr int K = K_, N = N_, M = M_, E = E_;
    static constexpr int NM = NM_;          // records per epoch: measured components (MEAS 0) or satellite slots (MEAS 1)
    static constexpr int DYN = DYN_;
    static constexpr int MEAS = MEAS_;      // 0: RT_STATE of a fixed component set, 1: RT_PR with one index quadruple
    static constexpr int WARPS = WARPS_;    // windows (warps) per CTA
    static constexpr int NB = NB_;          // bounded node components (log barrier)
    static constexpr bool PRIOR = PRIOR_;   // arrival cost on X_0
    static constexpr int LOSS = LOSS_;
    static constexpr bool HCONST = (DYN_ == NLPF_DYN_AFFINE);   // dynamics part of H constant (in Hc)
    static_assert(!(HCONST && LOSS_ == NLPF_LOSS_HUBER), "affine + pseudo-Huber goes to the general kernel");
};

template <class C>
struct WsLayout {
    static constexpr int K = C::K, N = C::N, E = C::E, NM = C::NM;
    static constexpr int NZ = K * N;
    static constexpr int NP = po(NZ);                  // H rows 0..NZ-1 in the padded layout
    static constexpr int NR = E * NM;                  // residuals
    static constexpr int ev(int x) { return (x + 1) & ~1; }
    static constexpr int LDW = (NZ + 2 + 3) & ~3;      // LDL panel copy W^T[4][LDW]
    static constexpr int oH = 0;
    static constexpr int oZ = oH + po(NZ + 1);         // + the appended right-hand-side row
    static constexpr int oZT = oZ + ev(NZ);
    static constexpr int oG = oZT + ev(NZ);
    static constexpr int oDZ = oG + ev(NZ);
    static constexpr int oDS = oDZ + ev(NZ);
    // Vectors that are never live together share storage (shared memory per window is what limits residency):
    //  * dd (Gauss-Newton diagonal) = ds (diagonal of H) when H has no second-order or barrier part;
    //  * lw (weighted defects: written and read inside the evaluation) = dz (the step: written by the
    //    factorisation, consumed before the next evaluation);
    //  * fb (node dynamics: dead after the defect phase) = the head of Q (written by the measurement phase).
    static constexpr bool ALIAS_DS = C::HCONST && C::NB == 0;
    static constexpr int oDD = ALIAS_DS ? oDS : oDS + ev(NZ);
    static constexpr int oLW = oDZ;
    static constexpr int oWH = oDD + ev(NZ);
    static constexpr int oF = oWH + ev(C::LOSS == NLPF_LOSS_HUBER ? NZ : 0);
    static constexpr int NBLK = C::HCONST ? 0 : K * N * N;       // per-node Jacobians only for nonlinear models
    static constexpr int oQ = oF + ev(NBLK);
    static constexpr int oFB = oQ;
    static constexpr int NQ = (C::MEAS == 0) ? NR : E * 14;      // w*r per residual, or [G_e (10) | q_e (4)] per epoch
    static constexpr int NPB = C::MMA ? 32 * ((NZ + 8) / 8) : 4 * LDW;      // LDL side buffer: panel rows x 4 (ws_ldl_mma) or W^T
    static constexpr int NFQ = (ev(NBLK) + ev(NQ) > NPB) ? ev(NBLK) + ev(NQ) : NPB;   // F | Q, reused by the LDL
    static constexpr int oCLK = oF + NFQ;              // start clock of the current window (profiling only)
    // linear measurements with constant weights (WS_LINQ): b = Phi^T W y per unknown, [sum w y^2, flag], per window
    static constexpr int oBM = oCLK + 2;
    static constexpr int oC0 = oBM + ev(C::MEAS == 0 ? NZ : 0);
    static constexpr int PER_WARP = oC0 + (C::MEAS == 0 ? 2 : 0);
    static_assert(NQ >= NZ, "fb shares the head of Q");
    static constexpr int oDt = C::WARPS * PER_WARP;    // shared by the CTA
    static constexpr int NTILE = ((NZ + 4) / 4) * ((NZ + 4) / 2 + 1);   // bound on the tile table (row quads x column pairs)
    static constexpr int oTT = oDt + ev(K * K);        // tile table (uint16) and counts (int), as doubles
    static constexpr int oTC = oTT + ev((NTILE + 3) / 4);
    // class constants staged once per CTA: interpolation matrix Phi (E x K), cw_k * qinv_i, node-pair table
    static constexpr int oPhi = oTC + ev((NZ + 3) / 2);
    static constexpr int oCW = oPhi + ev(E * K);
    static constexpr int oPT = oCW + ev(NZ);
    static constexpr int oPP = oPT + ev((K * (K + 1) / 2 + 3) / 4);      // full K x K matrix sum_e phi_ej phi_el (linear measurements)
    static constexpr int TOTAL = oPP + ev(C::MEAS == 0 ? K * K : 0);
    // resident CTAs per SM: shared memory (227 KB, 1 KB reserved per CTA) or 16 warps, whichever is smaller;
    // the register cap follows from it (GNSS / vehicle shapes: 8 warps per SM, so up to 255 registers)
    static constexpr int CTAS_SMEM = (227 * 1024) / (TOTAL * 8 + 1024);
    static constexpr int CTAS = (CTAS_SMEM < C::MAXW / C::WARPS) ? (CTAS_SMEM < 1 ? 1 : CTAS_SMEM) : C::MAXW / C::WARPS;
};

// CTA-shared class constants (shared memory) and the measured components (registers)
struct WsConst {
    const double* Dt;        // K x K
    const double* Phi;       // E x K
    const double* cwq;       // cw_k * qinv_i
    const uint16_t* ptab;    // node pairs j | l << 8
    const double* PP;        // K x K: sum_e phi_ej phi_el (linear-measurement classes), or null
    int c[4];                // MEAS 0: component of slot s (s < NM <= 4);  MEAS 1: (c0, c1, c2 | cb)
};

// receding-horizon chain arguments of the kernel (W = 1, null pointers: plain batch)
struct WsChain {
    int W;                   // windows per run; U and data are [W][B][...]
    const double* phiDT;     // K: basis at t = DT (device)
    double* Zout;            // [W][B][NZ] every window's estimate, or null
};

// class constants of the CTA: D~, the tile tables of the table-driven LDL, Phi, cw_k * qinv_i, the node-pair table
// (the caller synchronises the CTA before the first use)
template <class C>
__device__ __forceinline__ WsConst ws_stage(const DevProb& P, double* smem) {
    typedef WsLayout<C> L;
    constexpr int K = C::K, N = C::N, E = C::E, NM = C::NM, NZ = L::NZ;
    double* Dt = smem + L::oDt;
    for (int i = threadIdx.x; i < K * K; i += C::WARPS * 32) Dt[i] = P.Dt[i];
    uint16_t* ttab = reinterpret_cast<uint16_t*>(smem + L::oTT);
    int* tcnt = reinterpret_cast<int*>(smem + L::oTC);
    if (!C::MMA) {
        for (int i = threadIdx.x; i < P.tile_cnt[NZ + 1]; i += C::WARPS * 32) ttab[i] = P.tile_tab[i];
        for (int i = threadIdx.x; i < NZ + 2; i += C::WARPS * 32) tcnt[i] = P.tile_cnt[i];
    }
    uint16_t* ptab_s = reinterpret_cast<uint16_t*>(smem + L::oPT);
    for (int i = threadIdx.x; i < E * K; i += C::WARPS * 32) smem[L::oPhi + i] = P.Phi[i];
    for (int i = threadIdx.x; i < NZ; i += C::WARPS * 32) smem[L::oCW + i] = P.cw[i / N] * P.qinv[i % N];
    for (int i = threadIdx.x; i < K * (K + 1) / 2; i += C::WARPS * 32) ptab_s[i] = P.ptab[i];
    if (C::MEAS == 0) {
        for (int i = threadIdx.x; i < K * (K + 1) / 2; i += C::WARPS * 32) {
            const int pe = P.ptab[i];
            const int j = pe & 255, l = pe >> 8;
            smem[L::oPP + j * K + l] = P.PtP[i];
            smem[L::oPP + l * K + j] = P.PtP[i];
        }
    }
    WsConst cs;
    cs.Dt = Dt;
    cs.Phi = smem + L::oPhi;
    cs.cwq = smem + L::oCW;
    cs.ptab = ptab_s;
    cs.PP = (C::MEAS == 0) ? smem + L::oPP : nullptr;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (C::MEAS == 0) cs.c[q] = (q < NM) ? P.act[P.rloc[q * NLPF_NIDX]] : -1;
        else cs.c[q] = P.act[P.run_loc[q]];
    }
    return cs;
}

// objective J, merit = J/2 + barrier, first-order rounding bound of the merit and (FULL) H (packed,
// shared memory), g.  merit = +inf flags an iterate outside the bounds or a non-finite objective.
template <class C, bool FULL>
__device__ __forceinline__ void ws_eval(const DevProb& P, double* sm, const WsConst& cs, const double* zc,
                                        const double* Ug, const double* dg, const double* xp, const double* h0, double mu,
                                        double& J_out, double& merit_out, double& noise_out, bool* clipped_out = nullptr) {
    typedef WsLayout<C> L;
    constexpr int K = C::K, N = C::N, M = C::M, E = C::E, NM = C::NM, DYN = C::DYN;
    constexpr int NZ = L::NZ;
    constexpr bool HUB = (C::LOSS == NLPF_LOSS_HUBER);
    const int lane = threadIdx.x & 31;
    double* Hp = sm + L::oH;
    double* g = sm + L::oG;
    double* lw = sm + L::oLW;
    double* fb = sm + L::oFB;
    double* wh = sm + L::oWH;
    double* Fm = sm + L::oF;
    double* qv = sm + L::oQ;
    double cost = 0.0, nse = 0.0, bar = 0.0;
    bool bad = false;
    const double* Dt = cs.Dt;
    auto whv = [&](int k, int i) -> double { return HUB ? wh[k * N + i] : cs.cwq[k * N + i]; };
    auto phi = [&](int e, int j) -> double { return cs.Phi[e * K + j]; };
    auto phiT = [&](int j, int e) -> double { return cs.Phi[e * K + j]; };

    if (FULL) {       // constant part (+ linear measurements), asynchronously; waited for before the first use of Hp
        for (int p = 2 * lane; p < L::NP; p += 64) cp_async16<C::MEAS == 0>(Hp + p, h0 + p);
    }
    // node dynamics
    for (int k = lane; k < K; k += 32) {
        double x[N], u[M > 0 ? M : 1], f[N], F[N * N];
#pragma unroll
        for (int i = 0; i < N; ++i) x[i] = zc[k * N + i];
#pragma unroll
        for (int i = 0; i < M; ++i) u[i] = Ug[k * M + i];
        ws_dyn<DYN, N, M>(P, x, u, f, F);
#pragma unroll
        for (int i = 0; i < N; ++i) fb[k * N + i] = f[i];
        if (FULL && !C::HCONST) {
#pragma unroll
            for (int i = 0; i < N * N; ++i) Fm[k * N * N + i] = F[i];
        }
    }
    __syncwarp();
    // defects W_ki = sum_j Dt_kj x_ji - f_ki, weights
    WS_ROUND_PRAGMA
    for (int t_r = 0; t_r < (NZ + 31) / 32; ++t_r) {      // rounds unrolled: their dependent chains interleave (WS_UNROLL_ROUNDS)
        const int t = lane + 32 * t_r;
        if (t >= NZ) continue;
        const int k = t / N, i = t - k * N;
        const double fki = fb[t];
        // (sums over nodes / epochs below are split over 2-4 accumulators: with 4 warps per scheduler the
        // kernel is bound by dependent-FMA latency, not by issue slots)
        double Wa = -fki, Wb = 0.0, mga = fabs(fki), mgb = 0.0;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const double dx = Dt[k * K + j] * zc[j * N + i];
            if (j & 1) { Wb += dx; mgb += fabs(dx); }
            else { Wa += dx; mga += fabs(dx); }
        }
        const double W = Wa + Wb, mag = mga + mgb;
        const double cq = cs.cwq[t];
        if (HUB) {                                          // pseudo_huber_loss, cost_functions.py:25-31
            const double wd = W / P.delta;
            const double s = sqrt(1.0 + wd * wd);
            cost += cq * 2.0 * P.delta * P.delta * (s - 1.0);
            lw[t] = cq * W / s;
            wh[t] = cq * (s + 1.0) / (2.0 * s * s);
        } else {
            cost = fma(cq * W, W, cost);
            lw[t] = cq * W;
        }
        nse += 2.0 * fabs(lw[t]) * mag;
    }
    __syncwarp();
    if (FULL) {
        // gradient of the dynamics cost: g_(j,a) = sum_k Dt_kj lw_ka - sum_i F_j[i][a] lw_ji
        WS_ROUND_PRAGMA
        for (int t_r = 0; t_r < (NZ + 31) / 32; ++t_r) {      // rounds unrolled: their dependent chains interleave (WS_UNROLL_ROUNDS)
            const int t = lane + 32 * t_r;
            if (t >= NZ) continue;
            const int j = t / N, a = t - j * N;
            double s = 0.0, s2 = 0.0;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                if (k & 1) s2 = fma(Dt[k * K + j], lw[k * N + a], s2);
                else s = fma(Dt[k * K + j], lw[k * N + a], s);
            }
            s += s2;
            if (C::HCONST) {
#pragma unroll
                for (int i = 0; i < N; ++i) s = fma(-P.A[i * N + a], lw[j * N + i], s);
            } else {
#pragma unroll
                for (int i = 0; i < N; ++i) s = fma(-Fm[j * N * N + i * N + a], lw[j * N + i], s);
            }
            g[t] = s;
        }
        cp_async_wait_all();
        __syncwarp();
        if (!C::HCONST) {
            // H_(j,a),(l,b) += -Dt_jl w_jb F_j[b][a] - Dt_lj w_la F_l[a][b]  (+ Huber: sum_k Dt_kj Dt_kl w_ka on a == b)
            for (int pr = lane; pr < K * (K + 1) / 2; pr += 32) {
                const int pe = cs.ptab[pr];
                const int j = pe & 255, l = pe >> 8;
                const double djl = Dt[j * K + l], dlj = Dt[l * K + j];
                if constexpr (FPat<DYN, N>::SPARSE) {
                    typedef FPat<DYN, N> FP;
#pragma unroll
                    for (int a = 0; a < N; ++a) {
                        double s1 = 0.0;
                        if (HUB) {
#pragma unroll
                            for (int k = 0; k < K; ++k) s1 = fma(Dt[k * K + j] * Dt[k * K + l], wh[k * N + a], s1);
                        }
                        const double wla = whv(l, a);
#pragma unroll
                        for (int b = 0; b < N; ++b) {
                            if (!(FP::nz(b, a) || FP::nz(a, b) || (HUB && a == b))) continue;     // folds after unrolling
                            if (j == l && b > a) continue;
                            double s = 0.0;
                            if (FP::nz(b, a)) s = fma(-djl * whv(j, b), Fm[j * N * N + b * N + a], s);
                            if (FP::nz(a, b)) s = fma(-dlj * wla, Fm[l * N * N + a * N + b], s);
                            if (HUB && a == b) s += s1;
                            Hp[pk(j * N + a, l * N + b)] += s;
                        }
                    }
                } else
#pragma unroll 3
                for (int a = 0; a < N; ++a) {
                    double s1 = 0.0;
                    if (HUB) {
#pragma unroll
                        for (int k = 0; k < K; ++k) s1 = fma(Dt[k * K + j] * Dt[k * K + l], wh[k * N + a], s1);
                    }
                    const double wla = whv(l, a);
                    if constexpr (N % 2 == 0) {
                        // column pairs as 16-byte read-modify-writes (row starts and l*N are even): neighbouring
                        // lanes then touch contiguous 16-byte pieces instead of 8 bytes every 16 (2-way conflicts).
                        // In a diagonal node block the partner of the diagonal entry is the row's padding slot.
#pragma unroll
                        for (int b = 0; b < N; b += 2) {
                            if (j == l && b > a) continue;
                            const double2 fl = *reinterpret_cast<const double2*>(Fm + l * N * N + a * N + b);
                            double s0 = -djl * whv(j, b) * Fm[j * N * N + b * N + a] - dlj * wla * fl.x;
                            double s1b = -djl * whv(j, b + 1) * Fm[j * N * N + (b + 1) * N + a] - dlj * wla * fl.y;
                            if (HUB && a == b) s0 += s1;
                            if (HUB && a == b + 1) s1b += s1;
                            double2* hp = reinterpret_cast<double2*>(Hp + pk(j * N + a, l * N + b));
                            double2 hv = *hp;
                            hv.x += s0;
                            hv.y += s1b;
                            *hp = hv;
                        }
                    } else {
#pragma unroll
                        for (int b = 0; b < N; ++b) {
                            if (j == l && b > a) continue;
                            double s = -djl * whv(j, b) * Fm[j * N * N + b * N + a] - dlj * wla * Fm[l * N * N + a * N + b];
                            if (HUB && a == b) s += s1;
                            Hp[pk(j * N + a, l * N + b)] += s;
                        }
                    }
                }
            }
            __syncwarp();
            // diagonal blocks: + F_k^T Lambda_k F_k
            if constexpr (FPat<DYN, N>::SPARSE) {
                typedef FPat<DYN, N> FP;
                for (int t = lane; t < K * FP::NFTF; t += 32) {
                    const int k = t / FP::NFTF, p = t - k * FP::NFTF;
                    const unsigned code = (unsigned)(((p < 8) ? (FP::FTF_LO >> (8 * p)) : (FP::FTF_HI >> (8 * (p - 8)))) & 255);
                    const int a = code & 15, b = code >> 4;
                    const double* Fk = Fm + k * N * N;
                    double s = 0.0;
#pragma unroll
                    for (int i = 0; i < N; ++i) s = fma(Fk[i * N + a] * whv(k, i), Fk[i * N + b], s);
                    Hp[pk(k * N + a, k * N + b)] += s;
                }
            } else
            for (int t = lane; t < K * N * N; t += 32) {
                const int k = t / (N * N), rem = t - k * N * N, a = rem / N, b = rem - a * N;
                if (b > a) continue;
                const double* Fk = Fm + k * N * N;
                double s = 0.0;
#pragma unroll
                for (int i = 0; i < N; ++i) s = fma(Fk[i * N + a] * whv(k, i), Fk[i * N + b], s);
                Hp[pk(k * N + a, k * N + b)] += s;
            }
        }
        __syncwarp();
    }
    // arrival cost (nlp.py:279-286); its constant Hessian block is part of Hc
    if (C::PRIOR) {
        for (int i = lane; i < N; i += 32) {
            const double d0 = zc[i] - xp[i];
            cost += P.pinv[i] * d0 * d0;
            nse += 2.0 * P.pinv[i] * fabs(d0) * (fabs(zc[i]) + fabs(xp[i]));
            if (FULL) g[i] += P.pinv[i] * d0;
        }
    }
    if (C::MEAS == 0 && WS_LINQ && sm[L::oC0 + 1] != 0.0) {
        // Linear measurements with the same weight at every epoch (the usual case; ws_hlin set the flag and b, c0 for this
        // window): the residual cost is a quadratic form whose matrix M = (sum_e phi_ej phi_el) (x) diag(w) is a class
        // constant, so   J_meas = z^T M z - 2 b^T z + c0,   grad/2 = M z - b   -- a K x K product per measured component
        // instead of the E x K interpolation and its transpose (13 % of the van der Pol kernel's samples, round 2).
        // Cancellation: |z^T M z| / J_meas is about 1e2 here, i.e. two digits of sixteen; the rounding bound carries it.
        const double* bm = sm + L::oBM;
        WS_ROUND_PRAGMA
        for (int t_r = 0; t_r < (NZ + 31) / 32; ++t_r) {
            const int t = lane + 32 * t_r;
            if (t >= NZ) continue;
            const int j = t / N, a = t - j * N;
            int s = -1;
#pragma unroll
            for (int q = 0; q < (NM < 4 ? NM : 4); ++q)
                if (cs.c[q] == a) s = q;
            if (s >= 0) {
                const double w = dg[2 * s + 1];
                double ma = 0.0, mb = 0.0, mc = 0.0;
#pragma unroll
                for (int l = 0; l < K; ++l) {
                    const double pp = cs.PP[j * K + l], zv = zc[l * N + a];
                    if (l % 3 == 0) ma = fma(pp, zv, ma);
                    else if (l % 3 == 1) mb = fma(pp, zv, mb);
                    else mc = fma(pp, zv, mc);
                }
                const double mz = w * ((ma + mb) + mc);
                const double zt_ = zc[t], bt_ = bm[t];
                cost += zt_ * (mz - 2.0 * bt_);
                nse += 2.0 * fabs(zt_) * (fabs(mz) + 2.0 * fabs(bt_));
                if (FULL) g[t] += mz - bt_;
            }
        }
        if (lane == 0) { cost += sm[L::oC0]; nse += 2.0 * sm[L::oC0]; }
    } else if (C::MEAS == 0) {
        // linear measurements: residual t = e*NM + s measures one state component at epoch e (record [y, w])
        WS_ROUND_PRAGMA
        for (int t_r = 0; t_r < (L::NR + 31) / 32; ++t_r) {      // rounds unrolled: their dependent chains interleave (WS_UNROLL_ROUNDS)
            const int t = lane + 32 * t_r;
            if (t >= L::NR) continue;
            const int e = t / NM;
            const int sl = t - e * NM;
            const int c = (sl == 0) ? cs.c[0] : (sl == 1) ? cs.c[1] : (sl == 2) ? cs.c[2] : cs.c[3];
            const double* rec = dg + 2 * t;                  // uniform layout: records [y, w] back to back (checked on the host)
            const double w = rec[1];
            double va = 0.0, vb = 0.0, vc = 0.0;
#pragma unroll
            for (int j = 0; j < K; ++j) {
                const double ph = phiT(j, e), zv = zc[j * N + c];
                if (j % 3 == 0) va = fma(ph, zv, va);
                else if (j % 3 == 1) vb = fma(ph, zv, vb);
                else vc = fma(ph, zv, vc);
            }
            const double v = (va + vb) + vc;
            const double r = (w != 0.0) ? rec[0] - v : 0.0;
            cost = fma(w * r, r, cost);
            nse += 2.0 * w * fabs(r) * (2.0 * fabs(rec[0]) + fabs(r));
            if (FULL) qv[t] = w * r;
        }
        if (FULL) {
            __syncwarp();
            WS_ROUND_PRAGMA
            for (int t_r = 0; t_r < (NZ + 31) / 32; ++t_r) {      // rounds unrolled: their dependent chains interleave (WS_UNROLL_ROUNDS)
                const int t = lane + 32 * t_r;
                if (t >= NZ) continue;
                const int j = t / N, a = t - j * N;
                int s = -1;                                // slot of component a in an epoch's record set, or -1
#pragma unroll
                for (int q = 0; q < (NM < 4 ? NM : 4); ++q)
                    if (cs.c[q] == a) s = q;
                if (s >= 0) {
                    double ac[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 3
                    for (int e = 0; e + 3 < E; e += 4) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) ac[q] = fma(phi(e + q, j), qv[(e + q) * NM + s], ac[q]);
                    }
#pragma unroll
                    for (int e = E & ~3; e < E; ++e) ac[e & 3] = fma(phi(e, j), qv[e * NM + s], ac[e & 3]);
                    g[t] -= (ac[0] + ac[1]) + (ac[2] + ac[3]);
                }
            }
        }
    } else {
        // pseudoranges (measurements.py:56-70, 81-87): one lane per epoch, NM satellite slots, components
        // (c0,c1,c2 | cb); G_e = sum w grad grad^T (10 unique) and q_e = sum w r grad in registers
        const int c0 = cs.c[0], c1 = cs.c[1], c2 = cs.c[2], cb = cs.c[3];
        for (int e = lane; e < E; e += 32) {
            double x0 = 0.0, x1 = 0.0, x2 = 0.0, xb = 0.0;
#pragma unroll
            for (int j = 0; j < K; ++j) {
                const double ph = phiT(j, e);
                x0 = fma(ph, zc[j * N + c0], x0);
                x1 = fma(ph, zc[j * N + c1], x1);
                x2 = fma(ph, zc[j * N + c2], x2);
                xb = fma(ph, zc[j * N + cb], xb);
            }
            double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
            double g00 = 0, g10 = 0, g11 = 0, g20 = 0, g21 = 0, g22 = 0, g30 = 0, g31 = 0, g32 = 0, g33 = 0;
            // records of the epoch's slots are fetched four at a time before any arithmetic (20 loads in
            // flight per lane: the slot-by-slot version spent a third of the kernel waiting on L2)
            // (not unrolled over the batches: with all 12 slots in straight-line code the vehicle kernel was 1.7 %
            // and the GNSS kernel 0.7 % slower -- instruction fetch, see the no_instructions stalls in profiles/)
#pragma unroll 1
            for (int s0 = 0; s0 < NM; s0 += 4) {
                double ry[4], rw[4], rx[4], rY[4], rz[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const bool in = s0 + q < NM;
                    if (C::SOA) {                                                // [field][slot][epoch]
                        const double* rec = dg + (in ? (s0 + q) * E + e : e);
                        ry[q] = rec[0];
                        rw[q] = in ? rec[NM * E] : 0.0;
                        rx[q] = rec[2 * NM * E];
                        rY[q] = rec[3 * NM * E];
                        rz[q] = rec[4 * NM * E];
                    } else {
                        const double* rec = dg + (in ? 5 * (e * NM + s0 + q) : 0);   // uniform layout: [y, w, sat xyz] (host-checked)
                        ry[q] = rec[0];
                        rw[q] = in ? rec[1] : 0.0;
                        rx[q] = rec[2];
                        rY[q] = rec[3];
                        rz[q] = rec[4];
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const double w = rw[q];
                    if (w == 0.0) continue;                     // masked slot: skipped, never 0 * NaN
                    const double d0 = x0 - rx[q], d1 = x1 - rY[q], d2 = x2 - rz[q];
                    // one reciprocal square root serves the range and the unit vector (rho = s2 * rsqrt(s2):
                    // <= 2 ulp, 1e-8 m at GNSS ranges) instead of sqrt + division
                    const double s2 = d0 * d0 + d1 * d1 + d2 * d2;
                    const double ir = rsqrt(s2);
                    const double rho = s2 * ir;
                    const double r = ry[q] - (rho + xb);
                    cost = fma(w * r, r, cost);
                    nse += 2.0 * w * fabs(r) * (2.0 * fabs(ry[q]) + fabs(r));
                    if (FULL) {
                        const double e0 = d0 * ir, e1 = d1 * ir, e2 = d2 * ir;
                        const double wr = w * r;
                        a0 = fma(wr, e0, a0); a1 = fma(wr, e1, a1); a2 = fma(wr, e2, a2); a3 += wr;
                        const double w0 = w * e0, w1 = w * e1, w2 = w * e2;
                        g00 = fma(w0, e0, g00); g10 = fma(w1, e0, g10); g11 = fma(w1, e1, g11);
                        g20 = fma(w2, e0, g20); g21 = fma(w2, e1, g21); g22 = fma(w2, e2, g22);
                        g30 += w0; g31 += w1; g32 += w2; g33 += w;
                    }
                }
            }
            if (FULL) {
                double2* Ge = reinterpret_cast<double2*>(qv + e * 14);     // 112-byte rows: 16-byte aligned
                Ge[0] = make_double2(g00, g10); Ge[1] = make_double2(g11, g20); Ge[2] = make_double2(g21, g22);
                Ge[3] = make_double2(g30, g31); Ge[4] = make_double2(g32, g33);
                Ge[5] = make_double2(a0, a1); Ge[6] = make_double2(a2, a3);
            }
        }
        if (FULL) {
            __syncwarp();
            // gradient: g_(j,c_a) -= sum_e phi_ej q_e[a]
            for (int t = lane; t < K * 4; t += 32) {
                const int j = t >> 2, a = t & 3;
                double ac[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 3
                for (int e = 0; e + 3 < E; e += 4) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) ac[q] = fma(phi(e + q, j), qv[(e + q) * 14 + 10 + a], ac[q]);
                }
#pragma unroll
                for (int e = E & ~3; e < E; ++e) ac[e & 3] = fma(phi(e, j), qv[e * 14 + 10 + a], ac[e & 3]);
                const int c = (a == 0) ? c0 : (a == 1) ? c1 : (a == 2) ? c2 : cb;
                g[j * N + c] -= (ac[0] + ac[1]) + (ac[2] + ac[3]);
            }
            constexpr int NPAIR = K * (K + 1) / 2;
#if WS_DMMA
            // H_(j,a),(l,b) += sum_e phi_ej phi_el G_e[a][b] as a small GEMM on the FP64 tensor cores:
            //   C[pair][q] = sum_e PP[pair][e] * G[e][q],  PP[pair][e] = phi_ej phi_el formed on the fly,
            // q = the 10 unique entries of the symmetric G_e (columns 10..15 of the second n-tile are padding).
            // mma.sync.m8n8k4.f64 fragments: A[row = lane/4][k = lane%4], B[k = lane%4][n = lane/4],
            // C[row = lane/4][2*(lane%4) + {0,1}].  Three m-tiles in flight (six independent accumulator chains).
            {
                constexpr int MT = (NPAIR + 7) / 8, KT = (E + 3) / 4;
                const int grp = lane >> 2, tig = lane & 3;
#pragma unroll 1
                for (int mt0 = 0; mt0 < MT; mt0 += 3) {
                    int jj[3], ll[3];
                    bool pon[3];
                    double acc[3][4];
#pragma unroll
                    for (int u = 0; u < 3; ++u) {
                        const int pp_ = (mt0 + u) * 8 + grp;
                        pon[u] = (mt0 + u < MT) && pp_ < NPAIR;
                        const int pe = cs.ptab[pon[u] ? pp_ : 0];
                        jj[u] = pe & 255;
                        ll[u] = pe >> 8;
                        acc[u][0] = acc[u][1] = acc[u][2] = acc[u][3] = 0.0;
                    }
#pragma unroll 2
                    for (int kt = 0; kt < KT; ++kt) {
                        const int e = kt * 4 + tig;
                        const bool eon = e < E;
                        const int ec = eon ? e : 0;
                        const double b0 = eon ? qv[ec * 14 + grp] : 0.0;
                        const double b1 = (eon && grp < 2) ? qv[ec * 14 + 8 + grp] : 0.0;
#pragma unroll
                        for (int u = 0; u < 3; ++u) {
                            const double av = (pon[u] && eon) ? phi(ec, jj[u]) * phi(ec, ll[u]) : 0.0;
                            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                         : "+d"(acc[u][0]), "+d"(acc[u][1]) : "d"(av), "d"(b0));
                            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                         : "+d"(acc[u][2]), "+d"(acc[u][3]) : "d"(av), "d"(b1));
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 3; ++u) {
                        if (!pon[u]) continue;
                        const int j = jj[u], l = ll[u];
#pragma unroll
                        for (int v = 0; v < 4; ++v) {
                            if (v >= 2 && tig != 0) continue;          // second n-tile: only columns 8, 9 are real
                            const int q = (v < 2) ? 2 * tig + v : 8 + (v - 2);
                            const int a_ = (q >= 6) ? 3 : (q >= 3) ? 2 : (q >= 1) ? 1 : 0;
                            const int b_ = q - ((a_ * (a_ + 1)) >> 1);
                            const int ca = (a_ == 0) ? c0 : (a_ == 1) ? c1 : (a_ == 2) ? c2 : cb;
                            const int cbb = (b_ == 0) ? c0 : (b_ == 1) ? c1 : (b_ == 2) ? c2 : cb;
                            const double hv = acc[u][v];
                            // (a,b) and, off the diagonal of G, its mirror (b,a); inside a diagonal node block only the
                            // entry of the lower triangle exists
                            const int r1 = j * N + ca, k1 = l * N + cbb;
                            const int r2 = j * N + cbb, k2 = l * N + ca;
                            if (k1 <= r1) Hp[pk(r1, k1)] += hv;
                            if (a_ != b_ && ((j != l && k2 <= r2) || (j == l && k1 > r1))) Hp[pk(r2, k2)] += hv;
                        }
                    }
                }
            }
#else
            // H_(j,a),(l,b) += sum_e phi_ej phi_el G_e[a][b]: two lanes share a node pair (epoch halves),
            // 10 unique values per pair (G_e symmetric), combined with one shuffle
            constexpr int EH = (E + 1) / 2;
            for (int it0 = 0; it0 < 2 * NPAIR; it0 += 32) {
                const int it = it0 + lane;
                const bool on = it < 2 * NPAIR;
                const int pe = cs.ptab[on ? (it >> 1) : 0];
                const int j = pe & 255, l = pe >> 8;
                const int eb = (it & 1) * EH, ee = (it & 1) ? E : EH;
                double h[10];
#pragma unroll
                for (int q = 0; q < 10; ++q) h[q] = 0.0;
                if (on) {
                    for (int e = eb; e < ee; ++e) {
                        const double pp = phi(e, j) * phi(e, l);
                        const double2* Ge = reinterpret_cast<const double2*>(qv + e * 14);
#pragma unroll
                        for (int q = 0; q < 5; ++q) {
                            const double2 gv = Ge[q];
                            h[2 * q] = fma(pp, gv.x, h[2 * q]);
                            h[2 * q + 1] = fma(pp, gv.y, h[2 * q + 1]);
                        }
                    }
                }
#pragma unroll
                for (int q = 0; q < 10; ++q) h[q] += __shfl_xor_sync(0xffffffffu, h[q], 1);
                if (on && !(it & 1)) {
                    const int cc[4] = {c0, c1, c2, cb};
                    int q = 0;
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int b = 0; b <= a; ++b, ++q) {
                            // (a,b) and, off the diagonal of G, its mirror (b,a); inside a diagonal node block only
                            // the entry of the lower triangle exists
                            const int r1 = j * N + cc[a], k1 = l * N + cc[b];
                            if (k1 <= r1) Hp[pk(r1, k1)] += h[q];
                            if (a != b) {
                                const int r2 = j * N + cc[b], k2 = l * N + cc[a];
                                if (k2 <= r2 && !(j == l)) Hp[pk(r2, k2)] += h[q];
                                if (j == l && k1 > r1) Hp[pk(r2, k2)] += h[q];
                            }
                        }
                }
            }
#endif
        }
    }
    // log barrier on the bounded components (nlp.py:314-317)
    if (C::NB > 0 && mu > 0.0) {
        if (FULL) __syncwarp();
        for (int t = lane; t < K * C::NB; t += 32) {
            const int k = t / (C::NB > 0 ? C::NB : 1), b = t - k * C::NB;
            const int idx = k * N + P.bidx[b];
            const double v = zc[idx];
            const double lo = P.blo[b], hi = P.bhi[b];
            if (lo > -1e300) {
                const double s = v - lo;
                if (!(s > 0.0)) bad = true;
                else {
                    bar -= mu * log(s);
                    if (FULL) { g[idx] -= mu / s; Hp[pk(idx, idx)] += mu / (s * s); }
                }
            }
            if (hi < 1e300) {
                const double s = hi - v;
                if (!(s > 0.0)) bad = true;
                else {
                    bar -= mu * log(s);
                    if (FULL) { g[idx] += mu / s; Hp[pk(idx, idx)] += mu / (s * s); }
                }
            }
        }
    }
    if (FULL) {
        __syncwarp();
        // Gauss-Newton diagonal for the damping, then the exact second-order dynamics term
        double* dd = sm + L::oDD;
        for (int i = lane; i < NZ; i += 32) dd[i] = Hp[pk(i, i)];
        if constexpr (DYN == NLPF_DYN_VDP) {
            __syncwarp();
            bool clipped = false;
            for (int k = lane; k < K; k += 32) {           // hess f0 = [[0, -2 x1], [-2 x1, -2 x0]]
                const double x0 = zc[k * N], x1 = zc[k * N + 1], l0 = lw[k * N];
                const double s01 = l0 * 2.0 * x1, s11 = l0 * 2.0 * x0;
                // safeguard: the node's block is scaled back until the Gauss-Newton diagonal dominates its rows
                // (Gershgorin; t = 1 near the solution, where lw is small) -- keeps H definite away from it
                const double t = fmin(1.0, fmin(Hp[pk(k * N, k * N)] / fabs(s01), Hp[pk(k * N + 1, k * N + 1)] / (fabs(s01) + fabs(s11))));
                Hp[pk(k * N + 1, k * N)] += t * s01;
                Hp[pk(k * N + 1, k * N + 1)] += t * s11;
                clipped = clipped || (t < 1.0);
            }
            // (the driver starts an iteration whose Hessian needed the safeguard at lam >= lam_jump_indef)
            if (clipped_out) *clipped_out = __any_sync(0xffffffffu, clipped);
        }
    }
    const double J = wsum(cost);
    double merit = 0.5 * J + ((C::NB > 0) ? wsum(bar) : 0.0);
    const double anybad = (C::NB > 0) ? wmax(bad ? 1.0 : 0.0) : 0.0;
    if (anybad > 0.0 || !isfinite(merit)) merit = INFINITY;
    J_out = J;
    merit_out = merit;
    noise_out = 2.220446049250313e-16 * (0.5 * wsum(nse) + fabs(merit));
    __syncwarp();
}

// constant H of a window with linear measurements: Hc + sum_e (phi_e phi_e^T) (x) diag(w_e), built in the window's
// shared-memory H and written to hl (NP doubles: an L2-resident cache the evaluation streams from)
template <class C>
__device__ __forceinline__ void ws_hlin(const DevProb& P, double* sm, const WsConst& cs, const double* dg, double* hl) {
    typedef WsLayout<C> L;
    constexpr int K = C::K, N = C::N, E = C::E, NM = C::NM;
    const int lane = threadIdx.x & 31;
    double* Hp = sm + L::oH;
    for (int p = lane; p < L::NP; p += 32) Hp[p] = P.Hcp[p];
    double* wts = sm + L::oQ;
    for (int t = lane; t < L::NR; t += 32) wts[t] = dg[2 * t + 1];
    // the usual case -- the same weight at every epoch (constant measurement covariance) -- needs no sum
    // over the epochs: sum_e phi_ej phi_el is a class constant (P.PtP, built on the host)
    bool same = true;
    for (int t = lane; t < L::NR; t += 32) same = same && (dg[2 * t + 1] == dg[2 * (t % NM) + 1]);
    same = __all_sync(0xffffffffu, same);
    __syncwarp();
    if (lane == 0) sm[L::oC0 + 1] = same ? 1.0 : 0.0;
    if (same) {
        // b_(j,a) = sum_e phi_ej w y_e,a and c0 = sum w y^2 of this window (ws_eval's quadratic form); a non-finite
        // measurement makes them non-finite, and with them the objective: the window is reported as before
        constexpr int NZ_ = K * N;
        double c0 = 0.0;
        for (int t = lane; t < L::NR; t += 32)
            if (dg[2 * t + 1] != 0.0) c0 = fma(dg[2 * t + 1] * dg[2 * t], dg[2 * t], c0);      // masked slots are skipped, never 0 * NaN
        c0 = wsum(c0);
        if (lane == 0) sm[L::oC0] = c0;
        for (int t = lane; t < NZ_; t += 32) {
            const int j = t / N, a = t - j * N;
            int s = -1;
#pragma unroll
            for (int q = 0; q < (NM < 4 ? NM : 4); ++q)
                if (cs.c[q] == a) s = q;
            double acc = 0.0;
            if (s >= 0 && dg[2 * s + 1] != 0.0) {
                const double w = dg[2 * s + 1];
                for (int e = 0; e < E; ++e) acc = fma(cs.Phi[e * K + j], w * dg[2 * (e * NM + s)], acc);
            }
            sm[L::oBM + t] = acc;
        }
        for (int pr = lane; pr < K * (K + 1) / 2; pr += 32) {
            const int pe = cs.ptab[pr];
            const int j = pe & 255, l = pe >> 8;
            const double pp = P.PtP[pr];
#pragma unroll
            for (int s = 0; s < NM; ++s) {
                const int c = cs.c[s < 4 ? s : 3];
                Hp[pk(j * N + c, l * N + c)] += pp * wts[s];
            }
        }
    } else
    for (int pr = lane; pr < K * (K + 1) / 2; pr += 32) {
        const int pe = cs.ptab[pr];
        const int j = pe & 255, l = pe >> 8;
        double acc[2][NM];
#pragma unroll
        for (int s = 0; s < NM; ++s) acc[0][s] = acc[1][s] = 0.0;
#pragma unroll 4
        for (int e = 0; e + 1 < E; e += 2) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const double pp = cs.Phi[(e + h) * K + j] * cs.Phi[(e + h) * K + l];
#pragma unroll
                for (int s = 0; s < NM; ++s) acc[h][s] = fma(pp, wts[(e + h) * NM + s], acc[h][s]);
            }
        }
        if (E & 1) {
            const double pp = cs.Phi[(E - 1) * K + j] * cs.Phi[(E - 1) * K + l];
#pragma unroll
            for (int s = 0; s < NM; ++s) acc[0][s] = fma(pp, wts[(E - 1) * NM + s], acc[0][s]);
        }
#pragma unroll
        for (int s = 0; s < NM; ++s) {
            const int c = cs.c[s < 4 ? s : 3];
            Hp[pk(j * N + c, l * N + c)] += acc[0][s] + acc[1][s];
        }
    }
    __syncwarp();
    for (int p = lane; p < L::NP; p += 32) hl[p] = Hp[p];
}

// nlpf_assemble for the classes of this file: kernels (1)+(2) alone -- one evaluation of H (Gauss-Newton part +
// exact second-order dynamics term), g, J per window, H written to HBM as the full symmetric nz x nz matrix.
// The same ws_eval the whole-solve kernel iterates on, so the assemble parity tests check it directly.
template <class C>
__global__ void __launch_bounds__(C::WARPS * 32, WsLayout<C>::CTAS)
ws_assemble_kernel(const __grid_constant__ DevProb P, int B, const double* __restrict__ z_in, const double* __restrict__ U,
                   const double* __restrict__ data, const double* __restrict__ xprior, double mu,
                   double* __restrict__ H_out, double* __restrict__ g_out, double* __restrict__ cost_out) {
    typedef WsLayout<C> L;
    constexpr int K = C::K, N = C::N, M = C::M, NZ = L::NZ;
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const WsConst cs = ws_stage<C>(P, smem);
    __syncthreads();
    double* sm = smem + warp * L::PER_WARP;
    double* zs = sm + L::oZ;
    double* Hp = sm + L::oH;
    for (int b = blockIdx.x * C::WARPS + warp; b < B; b += gridDim.x * C::WARPS) {
        for (int i = lane; i < NZ; i += 32) zs[i] = z_in[(size_t)b * NZ + i];
        const double* dg = data + (size_t)b * P.Ldata;
        double* Ho = H_out + (size_t)b * NZ * NZ;
        const double* h0 = P.Hcp;
        if (C::MEAS == 0) {               // the window's own output block doubles as the constant-H scratch (NP < NZ^2)
            __syncwarp();
            ws_hlin<C>(P, sm, cs, dg, Ho);
            h0 = Ho;
        }
        __syncwarp();
        double J, merit, nse;
        ws_eval<C, true>(P, sm, cs, zs, U + (size_t)b * K * (M > 0 ? M : 1), dg, C::PRIOR ? xprior + (size_t)b * N : nullptr,
                         h0, mu, J, merit, nse);
        __syncwarp();
        for (int t = lane; t < NZ * NZ; t += 32) {
            const int r = t / NZ, c = t - r * NZ;
            Ho[t] = (c <= r) ? Hp[pk(r, c)] : Hp[pk(c, r)];
        }
        for (int i = lane; i < NZ; i += 32) g_out[(size_t)b * NZ + i] = sm[L::oG + i];
        if (lane == 0) {
            cost_out[2 * b] = J;
            cost_out[2 * b + 1] = merit;
        }
        __syncwarp();
    }
}

// ---- longest-first hand-out -----------------------------------------------------------------------------------
// ncu of round 1 (profiles/r1_vdp_wsolve_kernel_full.json): smsp_cycles_active_avg / sm_cycles_elapsed_max = 0.65.
// Windows were handed out in index order; 4096 windows are 1.85 rounds of the resident warps, iteration counts run
// 5..17 (97 % within 5..9), and the launch ended on whichever warp drew one of the few long windows late.  The
// objective at the start point predicts those: on 640 oracle solves every window with >= 10 iterations is in the
// top third by J(z0) (Spearman 0.66 with the iteration count).  So: one cost-only evaluation per window (a few
// per cent of one iteration), windows ranked by it (counting rank, B^2 / 32 compares per warp), and the atomic
// counter hands them out in descending order -- the long ones start in the first round.
template <class C>
__global__ void __launch_bounds__(C::WARPS * 32, WsLayout<C>::CTAS)
ws_key_kernel(const __grid_constant__ DevProb P, int B, const double* __restrict__ z_in, const double* __restrict__ U,
              const double* __restrict__ data, const double* __restrict__ xprior, nlpf_opts opt, double* __restrict__ keys) {
    typedef WsLayout<C> L;
    constexpr int K = C::K, N = C::N, M = C::M, NZ = L::NZ;
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const WsConst cs = ws_stage<C>(P, smem);
    __syncthreads();
    double* sm = smem + warp * L::PER_WARP;
    double* zs = sm + L::oZ;
    for (int b = blockIdx.x * C::WARPS + warp; b < B; b += gridDim.x * C::WARPS) {
        for (int i = lane; i < NZ; i += 32) zs[i] = z_in[(size_t)b * NZ + i];
        if (C::MEAS == 0 && lane == 0) sm[L::oC0 + 1] = 0.0;      // no per-window quadratic form here: direct residuals
        __syncwarp();
        double J, merit, nse;
        ws_eval<C, false>(P, sm, cs, zs, U + (size_t)b * K * (M > 0 ? M : 1), data + (size_t)b * P.Ldata,
                          C::PRIOR ? xprior + (size_t)b * N : nullptr, nullptr, (C::NB > 0) ? opt.mu0 : 0.0, J, merit, nse);
        if (lane == 0) keys[b] = (J == J && J < 1e300) ? J : 1e300;      // non-finite start points first (they end at once)
        __syncwarp();
    }
}

// order[rank] = b with rank = number of windows with a larger key (ties: lower index first); one warp per window
__global__ void __launch_bounds__(256) ws_rank_kernel(int B, const double* __restrict__ keys, int* __restrict__ order) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (b >= B) return;
    const double kb = keys[b];
    int cnt = 0;
    for (int i = lane; i < B; i += 32) {
        const double k = keys[i];
        cnt += (k > kb || (k == kb && i < b)) ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) order[cnt] = b;
}

// backward sweep L^T x = y (y = the scaled appended row), one panel per step: x_B = L11^-T y_B as four
// independent dot products with the stored inverse, then y_i -= sum_c l_(j0+c),i x_c for the rows above
template <int NZ>
__device__ __forceinline__ void ws_backsub(double* Hp, double* xs) {
    const int lane = threadIdx.x & 31;
    double* y = Hp + pk(NZ, 0);
#pragma unroll 1
    for (int j0 = ((NZ - 1) >> 2) << 2; j0 >= 0; j0 -= 4) {
        const int pw = (NZ - j0 < 4) ? NZ - j0 : 4;
        const double2 y01 = *reinterpret_cast<const double2*>(y + j0);
        const double2 y23 = *reinterpret_cast<const double2*>(y + j0 + 2);
        const double m10 = Hp[pk(j0 + 1, j0)];
        const double2 m2 = *reinterpret_cast<const double2*>(Hp + pk(j0 + 2, j0));
        const double2 m3 = *reinterpret_cast<const double2*>(Hp + pk(j0 + 3, j0));
        const double m32 = Hp[pk(j0 + 3, j0 + 2)];
        double x[4];
        if (pw == 4) {
            x[3] = y23.y;
            x[2] = fma(m32, y23.y, y23.x);
            x[1] = fma(m3.y, y23.y, fma(m2.y, y23.x, y01.y));
            x[0] = fma(m3.x, y23.y, fma(m2.x, y23.x, fma(m10, y01.y, y01.x)));
        } else {
            x[3] = 0.0;
            x[2] = (pw > 2) ? y23.x : 0.0;
            x[1] = (pw > 1) ? ((pw > 2) ? fma(m2.y, y23.x, y01.y) : y01.y) : 0.0;
            x[0] = y01.x;
            if (pw > 1) x[0] = fma(m10, y01.y, x[0]);
            if (pw > 2) x[0] = fma(m2.x, y23.x, x[0]);
        }
        for (int i = lane; i < j0; i += 32) {
            double v = y[i];
            v = fma(-Hp[pk(j0, i)], x[0], v);
            if (pw > 1) v = fma(-Hp[pk(j0 + 1, i)], x[1], v);
            if (pw > 2) v = fma(-Hp[pk(j0 + 2, i)], x[2], v);
            if (pw > 3) v = fma(-Hp[pk(j0 + 3, i)], x[3], v);
            y[i] = v;
        }
        if (lane < pw) xs[j0 + lane] = (lane == 0) ? x[0] : (lane == 1) ? x[1] : (lane == 2) ? x[2] : x[3];
        __syncwarp();
    }
}

// Backward sweep with y in REGISTERS (lane i holds y_i and y_(i+32)): the version above goes through shared memory
// twice per panel (y and x written, re-read after a __syncwarp) -- 11 dependent round trips that were 14.6 % of the
// van der Pol kernel's samples (profiles/r2a_vdp_wsolve_stalls.txt: backsub, stalls wait + branch + short_sb).
// Here a panel step is: four y values by shuffle from their owner lanes -> x_B (three dependent FMAs, every lane
// the same) -> each lane's own y_i -= sum_c l_(j0+c),i x_c.  The l and m loads of every step are independent of the
// sweep, so with the loop unrolled they are in flight long before they are needed.
template <int NZ>
__device__ __forceinline__ void ws_backsub_reg(const double* Hp, double* xs) {
    static_assert(NZ <= 64, "two y registers per lane");
    const int lane = threadIdx.x & 31;
    const double* y = Hp + pk(NZ, 0);
    double yr0 = (lane < NZ) ? y[lane] : 0.0;
    double yr1 = (lane + 32 < NZ) ? y[lane + 32] : 0.0;
#pragma unroll
    for (int j0 = ((NZ - 1) >> 2) << 2; j0 >= 0; j0 -= 4) {
        const int pw = (NZ - j0 < 4) ? NZ - j0 : 4;
        // L11^-1 (strict lower part, stored in place by the factorisation) and the four panel rows of L: no
        // dependence on the sweep
        const double m10 = (pw > 1) ? Hp[pk(j0 + 1, j0)] : 0.0;
        const double2 m2 = (pw > 2) ? *reinterpret_cast<const double2*>(Hp + pk(j0 + 2, j0)) : make_double2(0.0, 0.0);
        const double2 m3 = (pw > 3) ? *reinterpret_cast<const double2*>(Hp + pk(j0 + 3, j0)) : make_double2(0.0, 0.0);
        const double m32 = (pw > 3) ? Hp[pk(j0 + 3, j0 + 2)] : 0.0;
        double la[4] = {0.0, 0.0, 0.0, 0.0}, lb[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (c < pw) {
                if (lane < j0) la[c] = Hp[pk(j0 + c, lane)];
                if (j0 > 32 && lane + 32 < j0) lb[c] = Hp[pk(j0 + c, lane + 32)];
            }
        }
        double yb[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int i = j0 + c;                      // compile-time after unrolling: register and owner lane are constants
            yb[c] = (c < pw) ? __shfl_sync(0xffffffffu, (i >= 32) ? yr1 : yr0, i & 31) : 0.0;
        }
        const double x3 = yb[3];
        const double x2 = fma(m32, yb[3], yb[2]);
        const double x1 = fma(m3.y, yb[3], fma(m2.y, yb[2], yb[1]));
        const double x0 = fma(m3.x, yb[3], fma(m2.x, yb[2], fma(m10, yb[1], yb[0])));
        yr0 -= fma(la[0], x0, la[1] * x1) + fma(la[2], x2, la[3] * x3);
        if (j0 > 32) yr1 -= fma(lb[0], x0, lb[1] * x1) + fma(lb[2], x2, lb[3] * x3);
        if (lane < pw) xs[j0 + lane] = (lane == 0) ? x0 : (lane == 1) ? x1 : (lane == 2) ? x2 : x3;
    }
    __syncwarp();
}

// LDL^T of the packed matrix with the right-hand side appended as row NZ, then the backward sweep.
// Right-looking, panels of 4 columns:
//   1. every lane factors the 4x4 diagonal block redundantly in registers (frcp pivots);
//   2. one lane per row below: w = a L11^-T by substitution; W^T goes to a side buffer, the SCALED
//      l = w / d is stored in place (so the backward sweep needs no pivots);
//   3. the trailing triangle is updated in 4-row x 2-column register tiles from a table (32 independent
//      FMAs per tile, all shared-memory traffic as 16-byte accesses; the per-entry loop this replaces
//      took 30 % of the kernel: 16 instructions and a load->4 FMA->store chain per entry).  Column PAIRS
//      per lane keep the tile and W^T accesses of neighbouring lanes contiguous: the 2-row x 4-column
//      shape had them 32 bytes apart, a 2-way bank conflict on every access;
//   4. the diagonal block is overwritten with the strict lower part of L11^-1, which turns the 4x4
//      triangular solve of the backward sweep into independent dot products.
template <int NZ, int LDW>
__device__ __forceinline__ bool ws_ldl(const uint16_t* __restrict__ ttab, const int* __restrict__ tcnt, double* Hp,
                                       double* xs, double* Wt) {
    const int lane = threadIdx.x & 31;
    bool ok = true;
#pragma unroll 1
    for (int j0 = 0; j0 < NZ; j0 += 4) {
        const int pw = (NZ - j0 < 4) ? NZ - j0 : 4;
        __syncwarp();
        double a[4][4], l[4][4], d[4], rd[4];
        {
            const double* r0 = Hp + pk(j0, j0);
            const double* r1 = Hp + pk(j0 + 1, j0);
            const double* r2 = Hp + pk(j0 + 2, j0);
            const double* r3 = Hp + pk(j0 + 3, j0);
            const double2 v1 = *reinterpret_cast<const double2*>(r1);
            const double2 v2 = *reinterpret_cast<const double2*>(r2);
            const double2 v3a = *reinterpret_cast<const double2*>(r3);
            const double2 v3b = *reinterpret_cast<const double2*>(r3 + 2);
            a[0][0] = r0[0];
            a[1][0] = v1.x; a[1][1] = v1.y;
            a[2][0] = v2.x; a[2][1] = v2.y; a[2][2] = r2[2];
            a[3][0] = v3a.x; a[3][1] = v3a.y; a[3][2] = v3b.x; a[3][3] = v3b.y;
            if (pw < 4) {
#pragma unroll
                for (int r = 1; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c <= r; ++c)
                        if (r >= pw) a[r][c] = (r == c) ? 1.0 : 0.0;
            }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll
            for (int r = c; r < 4; ++r) {
                double w = a[r][c];
#pragma unroll
                for (int q = 0; q < c; ++q) w = fma(-a[r][q], l[c][q], w);
                a[r][c] = w;
            }
            d[c] = a[c][c];
            rd[c] = frcp(d[c]);
#pragma unroll
            for (int r = c + 1; r < 4; ++r) l[r][c] = a[r][c] * rd[c];
        }
        // pivots must be positive and finite: 1/d computed by frcp is positive and finite exactly then (d <= 0, NaN,
        // +inf and subnormal d all give a non-positive or NaN reciprocal after the Newton steps); four compares instead of the
        // fmin/fmax tree, which was 6 % of the kernel's instructions
        if (!((rd[0] > 0.0) & (rd[1] > 0.0) & (rd[2] > 0.0) & (rd[3] > 0.0))) {
            ok = false;
            break;
        }
        const int j1 = j0 + pw;
        for (int i = j1 + lane; i <= NZ; i += 32) {
            double* Hi = Hp + pk(i, j0);
            const double2 u01 = *reinterpret_cast<const double2*>(Hi);
            const double2 u23 = *reinterpret_cast<const double2*>(Hi + 2);   // (pw < 4: past the row, unused)
            const double w0 = u01.x;
            const double w1 = fma(-w0, l[1][0], u01.y);
            const double w2 = fma(-w1, l[2][1], fma(-w0, l[2][0], u23.x));
            const double w3 = fma(-w2, l[3][2], fma(-w1, l[3][1], fma(-w0, l[3][0], u23.y)));
            Wt[i] = w0;
            Wt[LDW + i] = w1;
            Wt[2 * LDW + i] = w2;
            Wt[3 * LDW + i] = w3;
            if (pw == 4) {
                *reinterpret_cast<double2*>(Hi) = make_double2(w0 * rd[0], w1 * rd[1]);
                *reinterpret_cast<double2*>(Hi + 2) = make_double2(w2 * rd[2], w3 * rd[3]);
            } else {
                Hi[0] = w0 * rd[0];
                if (pw > 1) Hi[1] = w1 * rd[1];
                if (pw > 2) Hi[2] = w2 * rd[2];
            }
        }
        __syncwarp();
        if (lane == 31) {       // strict lower part of L11^-1 (rows/columns >= pw are identity)
            const double m10 = -l[1][0], m21 = -l[2][1], m32 = -l[3][2];
            const double m20 = fma(-l[2][1], m10, -l[2][0]);
            const double m31 = fma(-l[3][2], m21, -l[3][1]);
            const double m30 = fma(-l[3][2], m20, fma(-l[3][1], m10, -l[3][0]));
            if (pw > 1) Hp[pk(j0 + 1, j0)] = m10;
            if (pw > 2) { Hp[pk(j0 + 2, j0)] = m20; Hp[pk(j0 + 2, j0 + 1)] = m21; }
            if (pw > 3) { Hp[pk(j0 + 3, j0)] = m30; Hp[pk(j0 + 3, j0 + 1)] = m31; Hp[pk(j0 + 3, j0 + 2)] = m32; }
        }
        if (j1 >= NZ) continue;
        const int nitem = tcnt[NZ + 1 - j1];
        for (int it = lane; it < nitem; it += 32) {
            const int e = ttab[it];
            const int i = j1 + ((e & 255) << 2);          // first row of the quad (a multiple of 4)
            const int k0 = j1 + ((e >> 8) << 1);          // column pair
            double* R[4];
            R[0] = Hp + po(i);
            R[1] = R[0] + i + 2;                          // padded row lengths: i+2, i+2, i+4
            R[2] = R[1] + i + 2;
            R[3] = R[2] + i + 4;
            double2 x[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) x[r] = *reinterpret_cast<const double2*>(R[r] + k0);
            double2 wv[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) wv[c] = *reinterpret_cast<const double2*>(Wt + c * LDW + k0);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const double2 la = *reinterpret_cast<const double2*>(R[r] + j0);
                const double2 lb = *reinterpret_cast<const double2*>(R[r] + j0 + 2);
                x[r].x = fma(-la.x, wv[0].x, x[r].x);
                x[r].y = fma(-la.x, wv[0].y, x[r].y);
                x[r].x = fma(-la.y, wv[1].x, x[r].x);
                x[r].y = fma(-la.y, wv[1].y, x[r].y);
                x[r].x = fma(-lb.x, wv[2].x, x[r].x);
                x[r].y = fma(-lb.x, wv[2].y, x[r].y);
                x[r].x = fma(-lb.y, wv[3].x, x[r].x);
                x[r].y = fma(-lb.y, wv[3].y, x[r].y);
            }
#pragma unroll
            for (int r = 0; r < 4; ++r)
                if (k0 <= i + r && i + r <= NZ) *reinterpret_cast<double2*>(R[r] + k0) = x[r];
        }
    }
    __syncwarp();
    if (!ok) return false;
    ws_backsub<NZ>(Hp, xs);
    return true;
}

// ---- LDL^T with the matrix in REGISTERS as FP64 tensor-core accumulator tiles --------------------------------
// ncu of the table-driven version below (profiles/r1_vdp_wsolve_*): the kernel is bound by the shared-memory
// crossbar (132 M wavefronts in 0.75 ms = 0.93 per active cycle per SM), and the trailing update moves most of
// them: every 4 x 2 register tile reads and writes its C entries and reads 12 operands for 32 FMAs.  Here the
// whole (NZ+1) x (NZ+1) lower triangle lives in registers in the accumulator layout of mma.sync.m8n8k4.f64 --
// 8 x 8 tiles, lane (g = lane/4, t = lane%4) holds C[g][2t], C[g][2t+1] -- for the entire factorisation: a rank-4
// panel update of a tile is ONE instruction with one operand per lane for A and B, and C never moves.  Shared
// memory only carries the panel: its 4 columns go to a small buffer (rows x 4), the diagonal block is factored
// redundantly per lane as before, one lane per row below substitutes (w) and stores the scaled l = w / d into the
// packed matrix for the backward sweep, and the SAME buffer then feeds both fragments of the update
//     C[I][J] -= L_I D L_J^T :   A[g][k] = -w[8I+g][k] / d_k,   B[k][n] = w[8J+n][k]   (one 8-byte load per tile index).
// Rows NZ+1.. of the last tile row are padding (unit diagonal), row NZ is the appended right-hand side.
__device__ __forceinline__ void dmma884(double2& c, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c.x), "+d"(c.y) : "d"(a), "d"(b));
}

template <int NZ, int LDW, int RN>
__device__ __forceinline__ bool ws_ldl_mma(double* Hp, double* xs, double* PB) {
    constexpr int NB = (NZ + 8) / 8;                 // tile rows covering rows 0..NZ
    constexpr int NPANEL = (NZ + 3) / 4;
    // PB: 32 * NB doubles (WsLayout::NPB)
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    double2 c[NB * (NB + 1) / 2];
#pragma unroll
    for (int I = 0; I < NB; ++I)
#pragma unroll
        for (int J = 0; J <= I; ++J) {
            const int r = 8 * I + g, cc = 8 * J + 2 * t;
            double2 v = make_double2(0.0, 0.0);
            if (r <= NZ && cc <= r) {                // (always true off the diagonal tiles for r <= NZ)
                v = *reinterpret_cast<const double2*>(Hp + po(r) + cc);
                if (cc == r) v.y = 0.0;              // upper triangle / row padding
            }
            if (r >= NZ) {                           // unit pivots for the appended row and the padding rows
                if (cc == r) v.x = 1.0;
                if (cc + 1 == r) v.y = 1.0;
            }
            c[I * (I + 1) / 2 + J] = v;
        }
    bool ok = true;
#pragma unroll
    for (int p = 0; p < NPANEL; ++p) {
        const int j0 = 4 * p, J = p >> 1, h = p & 1;
        const int pw = (NZ - j0 < 4) ? NZ - j0 : 4;
        const int j1 = j0 + pw;
        double l[4][4], rd[4];
        if (ok) {
            __syncwarp();                            // the previous panel's fragment loads are done
            if ((t >> 1) == h) {                     // the lanes that hold columns j0..j0+3 of tile column J
#pragma unroll
                for (int I = J; I < NB; ++I)
                    *reinterpret_cast<double2*>(PB + (8 * I + g) * 4 + 2 * (t & 1)) = c[I * (I + 1) / 2 + J];
            }
            __syncwarp();
            double a[4][4], d[4];
            {
                const double* r0 = PB + j0 * 4;
                const double2 v1 = *reinterpret_cast<const double2*>(r0 + 4);
                const double2 v2a = *reinterpret_cast<const double2*>(r0 + 8);
                const double2 v2b = *reinterpret_cast<const double2*>(r0 + 10);
                const double2 v3a = *reinterpret_cast<const double2*>(r0 + 12);
                const double2 v3b = *reinterpret_cast<const double2*>(r0 + 14);
                a[0][0] = r0[0];
                a[1][0] = v1.x; a[1][1] = v1.y;
                a[2][0] = v2a.x; a[2][1] = v2a.y; a[2][2] = v2b.x;
                a[3][0] = v3a.x; a[3][1] = v3a.y; a[3][2] = v3b.x; a[3][3] = v3b.y;
                if (pw < 4) {
#pragma unroll
                    for (int r = 1; r < 4; ++r)
#pragma unroll
                        for (int q = 0; q <= r; ++q)
                            if (r >= pw) a[r][q] = (r == q) ? 1.0 : 0.0;
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
#pragma unroll
                for (int r = q; r < 4; ++r) {
                    double w = a[r][q];
#pragma unroll
                    for (int u = 0; u < q; ++u) w = fma(-a[r][u], l[q][u], w);
                    a[r][q] = w;
                }
                d[q] = a[q][q];
                rd[q] = frcp<RN>(d[q]);
#pragma unroll
                for (int r = q + 1; r < 4; ++r) l[r][q] = a[r][q] * rd[q];
            }
            if (!((rd[0] > 0.0) & (rd[1] > 0.0) & (rd[2] > 0.0) & (rd[3] > 0.0))) ok = false;   // see ws_ldl
        }
        if (ok) {
            for (int i = j1 + lane; i <= NZ; i += 32) {
                double* Pi = PB + i * 4;
                const double2 u01 = *reinterpret_cast<const double2*>(Pi);
                const double2 u23 = *reinterpret_cast<const double2*>(Pi + 2);
                const double w0 = u01.x;
                const double w1 = fma(-w0, l[1][0], u01.y);
                const double w2 = fma(-w1, l[2][1], fma(-w0, l[2][0], u23.x));
                const double w3 = fma(-w2, l[3][2], fma(-w1, l[3][1], fma(-w0, l[3][0], u23.y)));
                *reinterpret_cast<double2*>(Pi) = make_double2(w0, w1);
                *reinterpret_cast<double2*>(Pi + 2) = make_double2(w2, w3);
                double* Hi = Hp + pk(i, j0);
                if (pw == 4) {
                    *reinterpret_cast<double2*>(Hi) = make_double2(w0 * rd[0], w1 * rd[1]);
                    *reinterpret_cast<double2*>(Hi + 2) = make_double2(w2 * rd[2], w3 * rd[3]);
                } else {
                    Hi[0] = w0 * rd[0];
                    if (pw > 1) Hi[1] = w1 * rd[1];
                    if (pw > 2) Hi[2] = w2 * rd[2];
                }
            }
            if (lane == 31) {       // strict lower part of L11^-1 for the backward sweep (rows/columns >= pw are identity)
                const double m10 = -l[1][0], m21 = -l[2][1], m32 = -l[3][2];
                const double m20 = fma(-l[2][1], m10, -l[2][0]);
                const double m31 = fma(-l[3][2], m21, -l[3][1]);
                const double m30 = fma(-l[3][2], m20, fma(-l[3][1], m10, -l[3][0]));
                if (pw > 1) Hp[pk(j0 + 1, j0)] = m10;
                if (pw > 2) { Hp[pk(j0 + 2, j0)] = m20; Hp[pk(j0 + 2, j0 + 1)] = m21; }
                if (pw > 3) { Hp[pk(j0 + 3, j0)] = m30; Hp[pk(j0 + 3, j0 + 1)] = m31; Hp[pk(j0 + 3, j0 + 2)] = m32; }
            }
            if (j1 < NZ) {
                __syncwarp();
                const double nrd = -((t == 0) ? rd[0] : (t == 1) ? rd[1] : (t == 2) ? rd[2] : rd[3]);
                double fa[NB], fb[NB];
#pragma unroll
                for (int X = J + h; X < NB; ++X) {
                    double v = PB[(8 * X + g) * 4 + t];
                    if (h == 0 && X == J && g < 4) v = 0.0;       // the panel's own rows: nothing to update there
                    fb[X] = v;
                    fa[X] = v * nrd;
                }
#pragma unroll
                for (int I = J + h; I < NB; ++I)
#pragma unroll
                    for (int J2 = J + h; J2 <= I; ++J2) dmma884(c[I * (I + 1) / 2 + J2], fa[I], fb[J2]);
            }
        }
    }
    __syncwarp();
    if (!ok) return false;
    // (fully unrolled register sweep only where its code is small: the vehicle kernel is bound by instruction fetch --
    // ncu stall no_instruction 20-40 % in every phase at 160 KB of SASS -- so it keeps the rolled shared-memory sweep)
    if constexpr (NZ <= WS_BACKSUB_REG_MAXNZ) ws_backsub_reg<NZ>(Hp, xs);
    else ws_backsub<NZ>(Hp, xs);
    return true;
}

template <class C>
__global__ void __launch_bounds__(C::WARPS * 32, WsLayout<C>::CTAS)
wsolve_kernel(const __grid_constant__ DevProb P, int B, double* __restrict__ z_io, const double* __restrict__ U,
              const double* __restrict__ data, double* xprior, nlpf_opts opt,
              double* __restrict__ cost_out, int* __restrict__ iters_out, int* __restrict__ status_out,
              double* __restrict__ hlin_all, int* __restrict__ counter, const int* __restrict__ order,
              unsigned long long* __restrict__ prof, WsChain ch) {
    typedef WsLayout<C> L;
    constexpr int K = C::K, N = C::N, M = C::M, E = C::E, NM = C::NM;
    constexpr int NZ = L::NZ;
    constexpr bool BND = C::NB > 0;
    constexpr int NBD = BND ? C::NB : 1;
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint16_t* ttab = reinterpret_cast<uint16_t*>(smem + L::oTT);
    int* tcnt = reinterpret_cast<int*>(smem + L::oTC);
    const WsConst cs = ws_stage<C>(P, smem);
    __syncthreads();
    double* sm = smem + warp * L::PER_WARP;
    double* Hp = sm + L::oH;
    double* zs = sm + L::oZ;
    double* zt = sm + L::oZT;
    double* g = sm + L::oG;
    double* dz = sm + L::oDZ;
    double* ds = sm + L::oDS;
    double* dd = sm + L::oDD;
    // optional per-phase clock counters (nlpf_problem_profile): warp-cycles inside the evaluation and inside the
    // factorisation + sweep, their call counts, and the warp-cycles spent on windows in total
    // (added straight to the global counters by lane 0 when profiling is on: accumulators in registers cost the
    // van der Pol kernel spills under its 168-register cap)

    while (true) {
        int b = 0;
        if (lane == 0) b = atomicAdd(counter, 1);
        b = __shfl_sync(0xffffffffu, b, 0);
        if (b >= B) break;
        if (order) b = order[b];                    // longest-first (ws_key_kernel / ws_rank_kernel)
        if (prof && lane == 0) *reinterpret_cast<long long*>(sm + L::oCLK) = clock64();
        const double* xp = C::PRIOR ? xprior + (size_t)b * N : nullptr;
        for (int i = lane; i < NZ; i += 32) zs[i] = z_io[(size_t)b * NZ + i];
        // receding-horizon chain (autonomous-car.py:232-288): the W windows of run b back to back in this warp --
        // warm start from the previous solution, arrival cost re-centred on x(DT) -- so no run waits at a
        // per-window launch boundary for the slowest window of the batch.  W = 1: a plain batch of windows.
        int it_sum = 0, nbad = 0, it_last = 0, st_last = NLPF_ST_MAXITER;
        double J_last = 0.0;
#pragma unroll 1
        for (int w = 0; w < ch.W; ++w) {
        const double* Ug = U + ((size_t)w * B + b) * K * (M > 0 ? M : 1);
        const double* dg = data + ((size_t)w * B + b) * P.Ldata;
        const double* hlin = P.Hcp;
        if (C::MEAS == 0) {
            // constant H of this window: Hc + sum_e (phi_e phi_e^T) (x) diag(w_e), kept in an L2-resident cache
            double* hl = hlin_all + (size_t)b * L::NP;
            hlin = hl;
            ws_hlin<C>(P, sm, cs, dg, hl);
        }
        __syncwarp();
        if (BND) {   // start strictly inside the bounds (IPOPT bound_push)
            for (int t = lane; t < K * C::NB; t += 32) {
                const int k = t / NBD, bb = t - k * NBD;
                const int idx = k * N + P.bidx[bb];
                double v = zs[idx];
                if (P.blo[bb] > -1e300) v = fmax(v, P.blo[bb] + opt.bound_push);
                if (P.bhi[bb] < 1e300) v = fmin(v, P.bhi[bb] - opt.bound_push);
                zs[idx] = v;
            }
            __syncwarp();
        }
        // ---- solver driver (same decisions as solve_kernel in nls_kernels.inl and oracle/nls.py) ----
        // Written as a state machine around ONE inlined evaluation and ONE factorisation site: with the
        // evaluation inlined at its five logical call sites the vehicle kernel spent 23 % of its issue slots
        // waiting for instructions (ncu stall "no_instructions").
        //   EV_FIRST  evaluation at the start point            EV_STAGE  after a barrier-parameter update
        //   EV_TRIAL  at the trial point z - alpha dz          EV_RESTORE / EV_STALL  H, g back at z after a
        //                                                       rejected step (the factorisation overwrote H)
        enum { EV_FIRST, EV_STAGE, EV_TRIAL, EV_RESTORE, EV_STALL };
        enum { PC_STAGE_INIT, PC_ITER_TOP, PC_LAM_TOP, PC_SOLVE, PC_AFTER_TRIAL, PC_STALL_CHECK, PC_STAGE_END };
        double mu = BND ? ((w > 0 && opt.mu0_warm > 0.0) ? opt.mu0_warm : opt.mu0) : 0.0;   // windows 2.. of a chain are warm-started
        int iters = 0, status = NLPF_ST_MAXITER, stage_steps = 0, floor_steps = 0;
        double lam = opt.lam0, lam_it0 = opt.lam0, J = 0.0, merit = 0.0, nse0 = 0.0;
        double Jt = 0.0, mt = 0.0, nset = 0.0, dzmax = 0.0, alpha = 1.0, pred = 0.0, dm = 0.0;
        bool h_valid = true;
        bool clip_z = false, clip_t = false;     // the second-order term at z / at the trial point needed its safeguard
        int ev = EV_FIRST;
        while (true) {
            {
                double Je, me, ne;
                const long long pt0 = prof ? clock64() : 0;
                bool clip_e = false;
                ws_eval<C, true>(P, sm, cs, (ev == EV_TRIAL) ? zt : zs, Ug, dg, xp, hlin, mu, Je, me, ne, &clip_e);
                if (ev == EV_TRIAL) clip_t = clip_e; else clip_z = clip_e;
                if (prof && lane == 0) { atomicAdd(prof + 0, (unsigned long long)(clock64() - pt0)); atomicAdd(prof + 2, 1ULL); }
                if (ev == EV_TRIAL) { Jt = Je; mt = me; nset = ne; }
                else if (ev == EV_FIRST || ev == EV_STAGE) { J = Je; merit = me; nse0 = ne; }
            }
            if (ev == EV_FIRST && !isfinite(merit)) { status = NLPF_ST_NUMERIC; break; }
            int pc = (ev == EV_TRIAL) ? PC_AFTER_TRIAL : (ev == EV_RESTORE) ? PC_SOLVE : (ev == EV_STALL) ? PC_STALL_CHECK : PC_STAGE_INIT;
            bool need_eval = false, finished = false;
            while (!need_eval) {
                if (pc == PC_STAGE_INIT) {
                    lam = opt.lam0;
                    stage_steps = 0;
                    floor_steps = 0;
                    status = NLPF_ST_MAXITER;
                    pc = PC_ITER_TOP;
                } else if (pc == PC_ITER_TOP) {
                    if (!(iters < opt.max_iter && stage_steps < opt.max_stage_iter)) { pc = PC_STAGE_END; continue; }
                    double gm = 0.0;
                    dm = 0.0;
                    for (int i = lane; i < NZ; i += 32) {
                        gm = fmax(gm, fabs(g[i]));
                        ds[i] = Hp[pk(i, i)];
                        dm = fmax(dm, dd[i]);
                    }
                    dm = wmax(dm);
                    if (BND) {
                        gm = wmax(gm);
                        if (mu > opt.mu_min && stage_steps >= 1 && gm < 10.0 * mu) { status = NLPF_ST_CONVERGED; pc = PC_STAGE_END; continue; }
                    }
                    h_valid = true;
                    if (clip_z) lam = fmax(lam, opt.lam_jump_indef);    // H is probably indefinite: do not try it undamped
                    lam_it0 = lam;
                    pc = PC_LAM_TOP;
                } else if (pc == PC_LAM_TOP) {
                    if (!(lam <= 1e18)) {   // no acceptable step
                        status = NLPF_ST_STALLED;
                        if (!h_valid) { ev = EV_STALL; need_eval = true; continue; }   // H, g back at z (stationarity test, next stage)
                        pc = PC_STALL_CHECK;
                        continue;
                    }
                    if (!h_valid) { h_valid = true; ev = EV_RESTORE; need_eval = true; continue; }
                    pc = PC_SOLVE;
                } else if (pc == PC_SOLVE) {
                    for (int i = lane; i < NZ; i += 32) {
                        Hp[pk(i, i)] = ds[i] + lam * (dd[i] + 1e-14 * dm);
                        Hp[pk(NZ, i)] = g[i];
                    }
                    h_valid = false;
                    bool ok;
                    const long long pt1 = prof ? clock64() : 0;
                    if constexpr (C::MMA) ok = ws_ldl_mma<NZ, L::LDW, (C::MEAS == 0) ? WS_LIN_NEWTON : 2>(Hp, dz, sm + L::oF);
                    else ok = ws_ldl<NZ, L::LDW>(ttab, tcnt, Hp, dz, sm + L::oF);
                    if (prof && lane == 0) { atomicAdd(prof + 1, (unsigned long long)(clock64() - pt1)); atomicAdd(prof + 3, 1ULL); }
                    if (ok) {
                        double a = 1.0, dmx = 0.0, fin = 0.0, pr = 0.0;
                        for (int i = lane; i < NZ; i += 32) {
                            dmx = fmax(dmx, fabs(dz[i]));
                            if (!isfinite(dz[i])) fin = 1.0;
                            pr = fma(g[i], dz[i], pr);
                        }
                        if (BND) {   // fraction to the boundary
                            for (int t = lane; t < K * C::NB; t += 32) {
                                const int k = t / NBD, bb = t - k * NBD;
                                const int idx = k * N + P.bidx[bb];
                                const double sd = -dz[idx];
                                if (sd < 0.0 && P.blo[bb] > -1e300) a = fmin(a, opt.tau_ftb * (zs[idx] - P.blo[bb]) / (-sd));
                                if (sd > 0.0 && P.bhi[bb] < 1e300) a = fmin(a, opt.tau_ftb * (P.bhi[bb] - zs[idx]) / sd);
                            }
                            a = -wmax(-a);
                        }
                        dmx = wmax(dmx);
                        fin = wmax(fin);
                        pr = a * wsum(pr);
                        if (fin == 0.0) {
                            double ztm = 0.0;
                            for (int i = lane; i < NZ; i += 32) {
                                const double v = fma(-a, dz[i], zs[i]);
                                zt[i] = v;
                                ztm = fmax(ztm, fabs(v));
                            }
                            ztm = wmax(ztm);
                            __syncwarp();
                            // a first-try step below the tolerance ends the stage without an evaluation at the new point:
                            // the merit cannot resolve such a step (the rounding-bound rule would accept it), and H, g
                            // there would not be used.  J stays the value at the previous iterate (it differs by O(step^2)).
                            if (lam <= lam_it0 && a * dmx < opt.tol_step * (1.0 + ztm)) {
                                for (int i = lane; i < NZ; i += 32) zs[i] = zt[i];
                                __syncwarp();
                                ++iters;
                                ++stage_steps;
                                status = NLPF_ST_CONVERGED;
                                pc = PC_STAGE_END;
                                continue;
                            }
                            // FULL evaluation at the trial point: an accepted step (the usual case) needs H and g
                            // there anyway, so a cost-only pass would be paid on top (VDP +7 %, GNSS +20 %)
                            alpha = a;
                            dzmax = dmx;
                            pred = pr;
                            ev = EV_TRIAL;
                            need_eval = true;
                            continue;
                        }
                    }
                    lam = fmax(10.0 * lam, opt.lam_jump_indef);     // pivot not positive or step not finite
                    pc = PC_LAM_TOP;
                } else if (pc == PC_AFTER_TRIAL) {
                    bool accepted = false, at_floor = false;
                    if (isfinite(mt) && mt < merit) accepted = true;
                    else {
                        const double nb = nse0 + nset;
                        if (isfinite(mt) && (mt - merit) <= nb && pred <= 4.0 * nb) { accepted = true; at_floor = true; }
                    }
                    if (!accepted) { lam = fmax(10.0 * lam, opt.lam_jump); pc = PC_LAM_TOP; continue; }
                    const double step = alpha * dzmax;
                    double zm = 0.0;
                    for (int i = lane; i < NZ; i += 32) {
                        zs[i] = zt[i];
                        zm = fmax(zm, fabs(zt[i]));
                    }
                    zm = wmax(zm);
                    lam = fmax(lam / ((lam <= lam_it0) ? opt.lam_dec_first : 10.0), 1e-12);
                    ++iters;
                    ++stage_steps;
                    J = Jt;             // H, g, dd already belong to the accepted point
                    clip_z = clip_t;
                    merit = mt;
                    nse0 = nset;
                    __syncwarp();
                    pc = PC_ITER_TOP;
                    if (step < opt.tol_step * (1.0 + zm)) { status = NLPF_ST_CONVERGED; pc = PC_STAGE_END; }
                    else if (at_floor && ++floor_steps >= 5) { status = NLPF_ST_CONVERGED; pc = PC_STAGE_END; }
                } else if (pc == PC_STALL_CHECK) {
                    // "stalled" is a success only at a stationary point: the diagonally scaled gradient step must be
                    // below NLPF_STALL_GTOL (NaN fails the test); otherwise the window is reported as failed
                    double sg = 0.0, zm = 0.0;
                    bool nan_g = false;
                    for (int i = lane; i < NZ; i += 32) {
                        const double q = fabs(g[i]) / (dd[i] + 1e-14 * dm + 1e-300);
                        nan_g = nan_g || !(q == q);
                        sg = fmax(sg, q);
                        zm = fmax(zm, fabs(zs[i]));
                    }
                    sg = wmax(sg);
                    zm = wmax(zm);
                    nan_g = __any_sync(0xffffffffu, nan_g);
                    if (nan_g || !(sg <= NLPF_STALL_GTOL * (1.0 + zm))) status = NLPF_ST_NOSTEP;
                    pc = PC_STAGE_END;
                } else {   // PC_STAGE_END
                    if (!BND || mu <= opt.mu_min || iters >= opt.max_iter) { finished = true; break; }
                    mu = fmax(opt.mu_min, fmin(0.2 * mu, mu * sqrt(mu)));
                    ev = EV_STAGE;
                    need_eval = true;
                }
            }
            if (finished) break;
        }
        J_last = J;
        it_last = iters;
        st_last = status;
        it_sum += iters;
        nbad += (status > NLPF_ST_STALLED) ? 1 : 0;
        if (ch.Zout) {
            for (int i = lane; i < NZ; i += 32) ch.Zout[((size_t)w * B + b) * NZ + i] = zs[i];
        }
        if (ch.phiDT) {      // next arrival cost: x(DT) = sum_j phi_j(DT) X_j
            __syncwarp();
            for (int i = lane; i < N; i += 32) {
                double sx = 0.0;
#pragma unroll
                for (int j = 0; j < K; ++j) sx = fma(ch.phiDT[j], zs[j * N + i], sx);
                xprior[(size_t)b * N + i] = sx;
            }
            __syncwarp();
        }
        }   // windows of the chain
        for (int i = lane; i < NZ; i += 32) z_io[(size_t)b * NZ + i] = zs[i];
        if (lane == 0) {
            cost_out[b] = J_last;
            iters_out[b] = ch.phiDT ? it_sum : it_last;          // chain: total iterations, windows not converged
            status_out[b] = ch.phiDT ? nbad : st_last;
        }
        if (prof && lane == 0) atomicAdd(prof + 4, (unsigned long long)(clock64() - *reinterpret_cast<long long*>(sm + L::oCLK)));
        __syncwarp();
    }
}

// records of one window [e][s][5] -> [5][s][e]: one CTA per window, through shared memory so that both the read and the
// write are coalesced
__global__ void __launch_bounds__(256) ws_soa_kernel(size_t nwin, int E, int NM, int Ldata, const double* __restrict__ in,
                                                     double* __restrict__ out) {
    extern __shared__ double soa_sm[];
    const int n = E * NM * 5;
    for (size_t w = blockIdx.x; w < nwin; w += gridDim.x) {
        const double* src = in + w * (size_t)Ldata;
        double* dst = out + w * (size_t)Ldata;
        for (int i = threadIdx.x; i < n; i += blockDim.x) soa_sm[i] = src[i];
        __syncthreads();
        for (int o = threadIdx.x; o < n; o += blockDim.x) {
            const int e = o % E, s = (o / E) % NM, f = o / (E * NM);
            dst[o] = soa_sm[(e * NM + s) * 5 + f];
        }
        __syncthreads();
    }
}

template <class C>
int ws_make_soa(const nlpf_problem* p, size_t nwin, const double* data, double* soa, cudaStream_t st) {
    const int n = C::E * C::NM * 5;
    if (n != p->d.Ldata) { nlpf_set_error("record block of %d doubles does not match the class (%d)", p->d.Ldata, n); return NLPF_E_ARG; }
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(ws_soa_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        attr = true;
    }
    const size_t grid = nwin < (size_t)148 * 8 ? nwin : (size_t)148 * 8;
    ws_soa_kernel<<<(unsigned)grid, 256, sizeof(double) * n, st>>>(nwin, C::E, C::NM, p->d.Ldata, data, soa);
    return nlpf_check_cuda(cudaGetLastError(), "ws_soa_kernel launch");
}

template <class C>
int ws_launch(nlpf_problem* p, int B, double* z, const double* U, const double* data, double* xprior,
              const nlpf_opts* o, double* cost_out, int* iters_out, int* status_out, double* hlin, int* counter,
              WsChain ch, cudaStream_t st) {
    typedef WsLayout<C> L;
    auto kern = wsolve_kernel<C>;
    const size_t smem = sizeof(double) * L::TOTAL;
    static bool attr = false;
    if (!attr) {
        if (nlpf_check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "wsolve smem attr")) return NLPF_E_CUDA;
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        attr = true;
    }
    if (nlpf_check_cuda(cudaMemsetAsync(counter, 0, sizeof(int), st), "wsolve counter reset")) return NLPF_E_CUDA;
    if (C::SOA) {                   // hlin is the caller's scratch of W * B * Ldata doubles for the structure-of-arrays copy
        int rc = ws_make_soa<C>(p, (size_t)ch.W * B, data, hlin, st);
        if (rc) return rc;
        data = hlin;
    }
    int grid = 148 * L::CTAS;
    const int need = (B + C::WARPS - 1) / C::WARPS;
    // longest-first hand-out: only when some windows have to wait for a free warp (more windows than resident
    // warps), for plain batches (a chain's length is the same for every run), up to 16384 windows (the counting
    // rank is quadratic: 2 us at 4096).  NLPF_NO_LPT=1 keeps the index order.
    static const bool lpt_off = getenv("NLPF_NO_LPT") != nullptr;
    const int* order = nullptr;
    if (C::LPT && !lpt_off && ch.W == 1 && !ch.phiDT && B > grid * C::WARPS && B <= 16384 && p->order && p->keys) {
        auto kkern = ws_key_kernel<C>;
        static bool kattr = false;
        if (!kattr) {
            if (nlpf_check_cuda(cudaFuncSetAttribute(kkern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "ws_key smem attr")) return NLPF_E_CUDA;
            kattr = true;
        }
        kkern<<<grid, C::WARPS * 32, smem, st>>>(p->d, B, z, U, data, xprior, *o, p->keys);
        ws_rank_kernel<<<(B + 7) / 8, 256, 0, st>>>(B, p->keys, p->order);
        order = p->order;
    }
    if (grid > need) grid = need;
    kern<<<grid, C::WARPS * 32, smem, st>>>(p->d, B, z, U, data, xprior, *o, cost_out, iters_out, status_out, hlin, counter, order,
                                            p->prof, ch);
    return nlpf_check_cuda(cudaGetLastError(), "wsolve_kernel launch");
}

template <class C>
int ws_assemble_launch(nlpf_problem* p, int B, const double* z, const double* U, const double* data, const double* xprior,
                       double mu, double* H_out, double* g_out, double* cost_out, cudaStream_t st) {
    typedef WsLayout<C> L;
    auto kern = ws_assemble_kernel<C>;
    const size_t smem = sizeof(double) * L::TOTAL;
    static bool attr = false;
    if (!attr) {
        if (nlpf_check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "ws_assemble smem attr")) return NLPF_E_CUDA;
        attr = true;
    }
    int grid = 148 * L::CTAS;
    const int need = (B + C::WARPS - 1) / C::WARPS;
    if (grid > need) grid = need;
    if (C::SOA) {                   // the windows' output blocks (nz^2 > Ldata doubles each? no: use a pooled scratch)
        double* soa = nullptr;
        if (nlpf_check_cuda(cudaMallocAsync((void**)&soa, sizeof(double) * (size_t)B * p->d.Ldata, st), "cudaMallocAsync(record copy)")) return NLPF_E_CUDA;
        int rc = ws_make_soa<C>(p, (size_t)B, data, soa, st);
        if (!rc) {
            kern<<<grid, C::WARPS * 32, smem, st>>>(p->d, B, z, U, soa, xprior, mu, H_out, g_out, cost_out);
            rc = nlpf_check_cuda(cudaGetLastError(), "ws_assemble_kernel launch");
        }
        cudaFreeAsync(soa, st);
        return rc;
    }
    kern<<<grid, C::WARPS * 32, smem, st>>>(p->d, B, z, U, data, xprior, mu, H_out, g_out, cost_out);
    return nlpf_check_cuda(cudaGetLastError(), "ws_assemble_kernel launch");
}

// windows (warps) per CTA.  Pseudorange shapes are limited by shared memory (about 21 KB per window):
//  * GNSS: 2 CTAs x 5 = 10 windows per SM -- 4096 windows are 27.7 per SM, i.e. 3 rounds instead of the 4 of
//    8 windows per SM (2.41 M -> 2.61 M windows/s; 1 x 10 warps: 2.59 M).  Storing G_e in 32-epoch chunks to
//    get there without the aliased vectors cost more (-12 %) than it gave.
//  * vehicle: 4 CTAs x 2 = 8 windows per SM; 10 per SM caps the kernel at 168 registers (spills): -6 %.
#ifndef WS_GNSS_WARPS
#define WS_GNSS_WARPS 5
#endif
#ifndef WS_CAR_WARPS
#define WS_CAR_WARPS 2
#endif
#ifndef WS_LIN_WARPS
#define WS_LIN_WARPS 4     // register-resident LDL^T: 3 CTAs x 4 warps = 12 windows per SM at 168 registers
                           // (table version: 3 x 5 = 15 per SM at 128 registers; measured 5, 6, 7, 8 warps per CTA:
                           // 5.29, 5.22, 5.18, 5.12 M windows/s)
#endif
#ifndef WS_LIN_MMA
#define WS_LIN_MMA 1
#endif
#ifndef WS_LIN_MAXW
#define WS_LIN_MAXW 12
#endif
typedef WsCfg<21, 2, 1, 50, 2, NLPF_DYN_VDP, 0, WS_LIN_WARPS, 0, false, NLPF_LOSS_L2, WS_LIN_MMA != 0, WS_LIN_MAXW> CfgVdp;          // van der Pol MHE (BASELINE configs[1])
typedef WsCfg<21, 2, 2, 50, 2, NLPF_DYN_AFFINE, 0, WS_LIN_WARPS, 0, false, NLPF_LOSS_L2, WS_LIN_MMA != 0, WS_LIN_MAXW> CfgExample;   // estimation_example.py
#ifndef WS_GNSS_MMA
#define WS_GNSS_MMA 0
#endif
#ifndef WS_GNSS_MAXW
#define WS_GNSS_MAXW 16
#endif
#ifndef WS_CAR_MMA
#define WS_CAR_MMA 1
#endif
#ifndef WS_CAR_MAXW
#define WS_CAR_MAXW 8
#endif
typedef WsCfg<11, 5, 3, 51, 12, NLPF_DYN_AFFINE, 1, WS_GNSS_WARPS, 0, false, NLPF_LOSS_L2, WS_GNSS_MMA != 0, WS_GNSS_MAXW> CfgGnss;     // gnss_stationary.py MHE (BASELINE configs[2])
// nlpf_assemble evaluates once per window: the structure-of-arrays copy would cost more than it saves (measured 0.18 vs 0.14 ms)
typedef WsCfg<11, 5, 3, 51, 12, NLPF_DYN_AFFINE, 1, WS_GNSS_WARPS, 0, false, NLPF_LOSS_L2, WS_GNSS_MMA != 0, WS_GNSS_MAXW, true> CfgGnssAsm;
typedef WsCfg<6, 9, 2, 21, 11, NLPF_DYN_VEHICLE_GNSS, 1, WS_CAR_WARPS, 2, true, NLPF_LOSS_L2, WS_CAR_MMA != 0, WS_CAR_MAXW> CfgCar;         // autonomous-car.py:184-220
typedef WsCfg<6, 9, 2, 21, 11, NLPF_DYN_VEHICLE_GNSS, 1, WS_CAR_WARPS, 2, true, NLPF_LOSS_HUBER, WS_CAR_MMA != 0, WS_CAR_MAXW> CfgCarHuber;  // autonomous-car.py:293-318

}  // namespace

// instantiated shapes.  meas: 0 linear state measurements (NM components per epoch), 1 pseudoranges
// (NM satellite slots per epoch).  Returns the variant index or -1.
// 1 when the class of this variant wants a W * B * Ldata scratch for the structure-of-arrays record copy
extern "C" int nlpf_ws_wants_soa(int variant) {
    return (variant == 2 && CfgGnss::SOA) || (variant == 3 && CfgCar::SOA) || (variant == 4 && CfgCarHuber::SOA) ? 1 : 0;
}

extern "C" int nlpf_ws_variant(int K, int N, int M, int E, int NM, int dyn, int meas, int nb, int prior, int loss) {
    const bool plain = nb == 0 && !prior && loss == NLPF_LOSS_L2;
    if (plain && meas == 0 && K == 21 && N == 2 && M == 1 && E == 50 && NM == 2 && dyn == NLPF_DYN_VDP) return 0;
    if (plain && meas == 0 && K == 21 && N == 2 && M == 2 && E == 50 && NM == 2 && dyn == NLPF_DYN_AFFINE) return 1;
    if (plain && meas == 1 && K == 11 && N == 5 && M == 3 && E == 51 && NM == 12 && dyn == NLPF_DYN_AFFINE) return 2;
    if (meas == 1 && K == 6 && N == 9 && M == 2 && E == 21 && NM == 11 && dyn == NLPF_DYN_VEHICLE_GNSS && nb == 2 && prior)
        return loss == NLPF_LOSS_L2 ? 3 : 4;
    return -1;
}

extern "C" int nlpf_ws_assemble(int variant, nlpf_problem* p, int B, const double* z, const double* U, const double* data,
                                const double* xprior, double mu, double* H_out, double* g_out, double* cost_out, cudaStream_t st) {
    switch (variant) {
        case 0: return ws_assemble_launch<CfgVdp>(p, B, z, U, data, xprior, mu, H_out, g_out, cost_out, st);
        case 1: return ws_assemble_launch<CfgExample>(p, B, z, U, data, xprior, mu, H_out, g_out, cost_out, st);
        case 2: return ws_assemble_launch<CfgGnssAsm>(p, B, z, U, data, xprior, mu, H_out, g_out, cost_out, st);
        case 3: return ws_assemble_launch<CfgCar>(p, B, z, U, data, xprior, mu, H_out, g_out, cost_out, st);
        case 4: return ws_assemble_launch<CfgCarHuber>(p, B, z, U, data, xprior, mu, H_out, g_out, cost_out, st);
    }
    nlpf_set_error("unknown warp-solve variant %d", variant);
    return NLPF_E_ARG;
}

extern "C" int nlpf_ws_launch(int variant, nlpf_problem* p, int B, double* z, const double* U, const double* data,
                              double* xprior, const nlpf_opts* o, double* cost_out, int* iters_out, int* status_out,
                              double* hlin, int* counter, int W, const double* phiDT, double* Zout, cudaStream_t st) {
    WsChain ch;
    ch.W = W;
    ch.phiDT = phiDT;
    ch.Zout = Zout;
    switch (variant) {
        case 0: return ws_launch<CfgVdp>(p, B, z, U, data, xprior, o, cost_out, iters_out, status_out, hlin, counter, ch, st);
        case 1: return ws_launch<CfgExample>(p, B, z, U, data, xprior, o, cost_out, iters_out, status_out, hlin, counter, ch, st);
        case 2: return ws_launch<CfgGnss>(p, B, z, U, data, xprior, o, cost_out, iters_out, status_out, hlin, counter, ch, st);
        case 3: return ws_launch<CfgCar>(p, B, z, U, data, xprior, o, cost_out, iters_out, status_out, hlin, counter, ch, st);
        case 4: return ws_launch<CfgCarHuber>(p, B, z, U, data, xprior, o, cost_out, iters_out, status_out, hlin, counter, ch, st);
    }
    nlpf_set_error("unknown warp-solve variant %d", variant);
    return NLPF_E_ARG;
}
