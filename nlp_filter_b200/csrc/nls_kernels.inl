// nls_kernels.inl -- one CTA per estimation window: residual/Jacobian evaluation, structured
// assembly of the normal equations, in-shared-memory LDL^T and the damped Gauss-Newton /
// log-barrier iteration, all on chip (fp64).
//
// Replaces, for B windows at once, what the reference runs per window inside
// CasADi -> IPOPT -> MUMPS when a script calls NLP.solve() (nlp/nlp.py:76-83) on a problem built
// with nlp/nlp.py:202-286.  The mathematics:
//   W_k = sum_j Dt_kj X_j - f(X_k,U_k) (+ shift)           defects, nlp.py:225-235
//   J   = sum_k cw_k rho(W_k) + sum_i w_i (y_i - h_i(Phi_i V))^2 + (X_0-x0)^T P^-1 (X_0-x0)
//         + sum_k cq_k V_k^T S V_k
//   H   = Hc + sum_e (phi_e phi_e^T) (x) G_e + dynamics cross/diagonal blocks,   g = J^T W r
// H is dense (nz x nz) because the Chebyshev differentiation matrix and the Lagrange
// interpolation couple all nodes (collocation.py:42-64, nlp.py:266-269); it lives packed
// (lower triangle, row-major) in shared memory with the right-hand side appended as one more
// row, so the forward substitution happens inside the factorisation.
#include "common.cuh"
#ifndef __CUDACC_RTC__
#include <math.h>
#endif

// This file is compiled four times (nls_kernels_{gen,pr}.cu and their _big twins): NLPF_PR_RUNS = 1 adds the
// lane-group / register-accumulation path for pseudorange runs.  Keeping it out of the other
// instantiation matters: with the 64-register cap its 14 accumulators cost the van der Pol
// kernel 21 % (measured) although that problem class has no pseudoranges at all.
#ifndef NLPF_PR_RUNS
#error "include through nls_kernels_gen.cu / nls_kernels_pr.cu"
#endif
// NLPF_BIG = 1 (nls_kernels_*_big.cu): the same kernels for problem classes whose shared memory allows at most 5
// windows per SM anyway (two-receiver, RC, multi-receiver shapes): min 4 CTAs per SM, i.e. a 128-register cap
// instead of 64 (two-receiver batch: 85 k -> 97 k windows/s).
#ifndef NLPF_BIG
#define NLPF_BIG 0
#endif
#if NLPF_BIG
#undef NLPF_MIN_CTAS
#define NLPF_MIN_CTAS 4
#if NLPF_PR_RUNS
#define VN(name) name##_prb
#else
#define VN(name) name##_genb
#endif
#else
#if NLPF_PR_RUNS
#define VN(name) name##_pr
#else
#define VN(name) name##_gen
#endif
#endif

namespace VN(nlpf_ns) {      // both instantiations define the same helper names

// ------------------------------------------------------------------------------------------
// block-wide reductions; every thread returns the same value (same summation order)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ double block_sum(double v, double* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NLPF_NW; ++i) s += red[i];
    return s;
}
__device__ double block_max(double v, double* red) {
    v = warp_max(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = red[0];
#pragma unroll
    for (int i = 1; i < NLPF_NW; ++i) s = fmax(s, red[i]);
    return s;
}
__device__ double block_min(double v, double* red) {
    v = warp_min(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = red[0];
#pragma unroll
    for (int i = 1; i < NLPF_NW; ++i) s = fmin(s, red[i]);
    return s;
}

// Several reductions behind one pair of barriers (the per-value versions above cost 2 barriers
// each; an evaluation needs J, barrier, rounding bound and the out-of-bounds flag together).
// v[0..NS) are summed, v[NS..NS+NM) are maximised; results overwrite v in every thread.
template <int NS, int NM>
__device__ void block_reduce(double* v, double* red) {
#pragma unroll
    for (int k = 0; k < NS; ++k) v[k] = warp_sum(v[k]);
#pragma unroll
    for (int k = 0; k < NM; ++k) v[NS + k] = warp_max(v[NS + k]);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < NS + NM; ++k) red[(threadIdx.x >> 5) * (NS + NM) + k] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < NLPF_NW; ++i) s += red[i * (NS + NM) + k];
        v[k] = s;
    }
#pragma unroll
    for (int k = 0; k < NM; ++k) {
        double s = red[NS + k];
#pragma unroll
        for (int i = 1; i < NLPF_NW; ++i) s = fmax(s, red[i * (NS + NM) + NS + k]);
        v[NS + k] = s;
    }
}

// ------------------------------------------------------------------------------------------
// Kernel (1a): continuous-time models f(x,u) and dF/dv, nlp/dynamics.py.
// Fk is the n x nv Jacobian w.r.t. the node variables (state columns through cmap, unknown
// controls through ucol), row-major, zeroed by the caller when jac is set.
// ------------------------------------------------------------------------------------------
struct FAcc {
    double* Fk;
    const int* cmap;
    int nv;
    __device__ __forceinline__ void add(int row, int statecomp, double v) const {
        Fk[row * nv + cmap[statecomp]] += v;
    }
};

#ifdef NLPF_USER_DYN
// generated at run time from a user's Python callable (nlp_filter_b200/nlp/_codegen.py) and compiled with NVRTC
// together with this file (csrc/user_model.cu):
//   __device__ void nlpf_user_dyn(const double* par, const double* x, const double* u, double* f, bool jac,
//                                 const FAcc& F, const int* ucol, double* Fk, int nv);
#include "nlpf_user_dyn.inc"
#endif

__device__ void vehicle6(const double* par, const double* x, const double* u, double* f, bool jac,
                         const FAcc& F) {
    // vehicle_dynamics, dynamics.py:148-164; par = C_AF C_AR M D_F D_R I_Z
    const double C_AF = par[0], C_AR = par[1], M = par[2], D_F = par[3], D_R = par[4], I_Z = par[5];
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
    f[3] = (-Fyf * sd + Fx) / M + r * vy;
    f[4] = (Fyf * cd + Fyr) / M - r * vx;
    f[5] = (D_F * Fyf * cd - D_R * Fyr) / I_Z;
    if (!jac) return;
    const double den2 = den * den;
    const double a_r = C_AR * (vy - D_R * r) / den2, b_r = -C_AR / den, c_r = C_AR * D_R / den;
    const double a_f = C_AF * (vy + D_F * r) / den2, b_f = -C_AF / den, c_f = -C_AF * D_F / den;
    F.add(0, 2, -vx * sp - vy * cp);
    F.add(0, 3, cp);
    F.add(0, 4, -sp);
    F.add(1, 2, vx * cp - vy * sp);
    F.add(1, 3, sp);
    F.add(1, 4, cp);
    F.add(2, 5, 1.0);
    F.add(3, 3, -sd * a_f / M);
    F.add(3, 4, -sd * b_f / M + r);
    F.add(3, 5, -sd * c_f / M + vy);
    F.add(4, 3, (cd * a_f + a_r) / M - r);
    F.add(4, 4, (cd * b_f + b_r) / M);
    F.add(4, 5, (cd * c_f + c_r) / M - vx);
    F.add(5, 3, (D_F * cd * a_f - D_R * a_r) / I_Z);
    F.add(5, 4, (D_F * cd * b_f - D_R * b_r) / I_Z);
    F.add(5, 5, (D_F * cd * c_f - D_R * c_r) / I_Z);
}

__device__ void quadcopter12(const double* par, const double* x, const double* u, double* f, bool jac,
                             const FAcc& F) {
    // quadcopter, dynamics.py:40-59; par = m Ixx Iyy Izz
    const double m = par[0], Ix = par[1], Iy = par[2], Iz = par[3];
    const double p = x[9], q = x[10], r = x[11];
    double sph, cph, sth, cth, sps, cps;
    sincos(x[3], &sph, &cph);
    sincos(x[4], &sth, &cth);
    sincos(x[5], &sps, &cps);
    const double tth = sth / cth;
    const double a = q * sph + r * cph;
    const double tm = u[0] / m;
    f[0] = x[6];
    f[1] = x[7];
    f[2] = x[8];
    f[3] = p + a * tth;
    f[4] = q * cph - r * sph;
    f[5] = a / cth;
    f[6] = tm * (sph * sps + cph * sth * cps);
    f[7] = tm * (cph * sth * sps - sph * cps);
    f[8] = tm * (cph * cth) - 9.81;
    f[9] = (u[1] - (Iz - Iy) * q * r) / Ix;
    f[10] = (u[2] - (Ix - Iz) * r * p) / Iy;
    f[11] = (u[3] - (Iy - Ix) * p * q) / Iz;
    if (!jac) return;
    const double da = q * cph - r * sph;
    F.add(0, 6, 1.0);
    F.add(1, 7, 1.0);
    F.add(2, 8, 1.0);
    F.add(3, 3, da * tth);
    F.add(3, 4, a / (cth * cth));
    F.add(3, 9, 1.0);
    F.add(3, 10, sph * tth);
    F.add(3, 11, cph * tth);
    F.add(4, 3, -q * sph - r * cph);
    F.add(4, 10, cph);
    F.add(4, 11, -sph);
    F.add(5, 3, da / cth);
    F.add(5, 4, a * sth / (cth * cth));
    F.add(5, 10, sph / cth);
    F.add(5, 11, cph / cth);
    F.add(6, 3, tm * (cph * sps - sph * sth * cps));
    F.add(6, 4, tm * (cph * cth * cps));
    F.add(6, 5, tm * (sph * cps - cph * sth * sps));
    F.add(7, 3, tm * (-sph * sth * sps - cph * cps));
    F.add(7, 4, tm * (cph * cth * sps));
    F.add(7, 5, tm * (cph * sth * cps + sph * sps));
    F.add(8, 3, tm * (-sph * cth));
    F.add(8, 4, tm * (-cph * sth));
    F.add(9, 10, -(Iz - Iy) * r / Ix);
    F.add(9, 11, -(Iz - Iy) * q / Ix);
    F.add(10, 9, -(Ix - Iz) * r / Iy);
    F.add(10, 11, -(Ix - Iz) * p / Iy);
    F.add(11, 9, -(Iy - Ix) * q / Iz);
    F.add(11, 10, -(Iy - Ix) * p / Iz);
}

// f and (optionally) dF/dv of node k.  x, u are the node's full state / control values.
__device__ void dyn_eval(const DevProb& P, const double* x, const double* u, double* f, bool jac, double* Fk) {
    const int n = P.n, m = P.m, nv = P.nv;
    FAcc F{Fk, P.cmap, nv};
    switch (P.dyn_id) {
        case NLPF_DYN_AFFINE: {
            for (int i = 0; i < n; ++i) {
                double s = P.c0[i];
                for (int j = 0; j < n; ++j) s = fma(P.A[i * n + j], x[j], s);
                for (int j = 0; j < m; ++j) s = fma(P.Bm[i * m + j], u[j], s);
                f[i] = s;
            }
            if (jac)
                for (int i = 0; i < n * nv; ++i) Fk[i] = P.Fc[i];
            return;
        }
        case NLPF_DYN_VDP: {   // dynamics.py:61-66
            f[0] = (1.0 - x[1] * x[1]) * x[0] - x[1] + u[0];
            f[1] = x[0];
            if (jac) {
                F.add(0, 0, 1.0 - x[1] * x[1]);
                F.add(0, 1, -2.0 * x[0] * x[1] - 1.0);
                F.add(1, 0, 1.0);
                if (P.ucol[0] >= 0) Fk[0 * nv + P.ucol[0]] += 1.0;
            }
            return;
        }
        case NLPF_DYN_BICYCLE: {   // dynamics.py:117-136 (heading is x[2])
            const double v = 8.72649116358 * u[0] - 0.856053299155;
            const double dl = 0.48869219055841229 * u[1];   // deg2rad(28)
            double s2, c2;
            sincos(x[2], &s2, &c2);
            f[0] = v * c2;
            f[1] = v * s2;
            f[2] = 0.0;
            f[3] = x[4];
            f[4] = 0.0;
            f[5] = (v / 0.28) * tan(dl);
            if (jac) {
                F.add(0, 2, -v * s2);
                F.add(1, 2, v * c2);
                F.add(3, 4, 1.0);
            }
            return;
        }
        case NLPF_DYN_VEHICLE:
            vehicle6(P.dyn_par, x, u, f, jac, F);
            return;
        case NLPF_DYN_VEHICLE_GNSS:   // dynamics.py:166-174
            vehicle6(P.dyn_par, x, u, f, jac, F);
            f[6] = x[7];
            f[7] = 0.0;
            f[8] = 0.0;
            if (jac) F.add(6, 7, 1.0);
            return;
        case NLPF_DYN_QUADCOPTER:
            quadcopter12(P.dyn_par, x, u, f, jac, F);
            return;
#ifdef NLPF_USER_DYN
        case NLPF_DYN_USER:
            nlpf_user_dyn(P.dyn_par, x, u, f, jac, F, P.ucol, Fk, nv);
            return;
#endif
    }
}

// ------------------------------------------------------------------------------------------
// Kernel (1b): measurement models h and their sparse gradients, nlp/measurements.py.
// loc[] are local indices (< na: interpolated active component, >= na: global unknown).
// Returns the number of gradient entries written to (gl[], gv[]).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double lval(int l, const double* xe, const double* glob, int na) {
    return l < na ? xe[l] : glob[l - na];
}

__device__ int meas_eval(int rt, const int* loc, const double* rec, const double* xe, const double* glob,
                         int na, double& h, int* gl, double* gv) {
    switch (rt) {
        case NLPF_RT_STATE:
            h = lval(loc[0], xe, glob, na);
            gl[0] = loc[0];
            gv[0] = 1.0;
            return 1;
        case NLPF_RT_PR: {   // measurements.py:56-70, 81-87
            const double d0 = lval(loc[0], xe, glob, na) - rec[2];
            const double d1 = lval(loc[1], xe, glob, na) - rec[3];
            const double d2 = lval(loc[2], xe, glob, na) - rec[4];
            const double rho = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
            h = rho + lval(loc[3], xe, glob, na);
            const double ir = 1.0 / rho;
            gl[0] = loc[0]; gv[0] = d0 * ir;
            gl[1] = loc[1]; gv[1] = d1 * ir;
            gl[2] = loc[2]; gv[2] = d2 * ir;
            gl[3] = loc[3]; gv[3] = 1.0;
            return 4;
        }
        case NLPF_RT_PRR: {   // measurements.py:72-79
            const double r0 = rec[2] - lval(loc[0], xe, glob, na);
            const double r1 = rec[3] - lval(loc[1], xe, glob, na);
            const double r2 = rec[4] - lval(loc[2], xe, glob, na);
            const double rho = sqrt(r0 * r0 + r1 * r1 + r2 * r2);
            const double ir = 1.0 / rho;
            const double e0 = r0 * ir, e1 = r1 * ir, e2 = r2 * ir;
            const double v0 = rec[5] - lval(loc[3], xe, glob, na);
            const double v1 = rec[6] - lval(loc[4], xe, glob, na);
            const double v2 = rec[7] - lval(loc[5], xe, glob, na);
            const double ev = e0 * v0 + e1 * v1 + e2 * v2;
            h = ev + lval(loc[6], xe, glob, na);
            gl[0] = loc[0]; gv[0] = -(v0 - e0 * ev) * ir;
            gl[1] = loc[1]; gv[1] = -(v1 - e1 * ev) * ir;
            gl[2] = loc[2]; gv[2] = -(v2 - e2 * ev) * ir;
            gl[3] = loc[3]; gv[3] = -e0;
            gl[4] = loc[4]; gv[4] = -e1;
            gl[5] = loc[5]; gv[5] = -e2;
            gl[6] = loc[6]; gv[6] = 1.0;
            return 7;
        }
        case NLPF_RT_RANGE:        // measurements.py:13-19, 46-53: loc[0..2] vs loc[3..5], dim = loc[6]
        case NLPF_RT_RANGE_PT: {   // measurements.py:8-12, 40-45: loc[3..5] < 0 => point from the record
            const int dim = loc[6];
            double d[3] = {0.0, 0.0, 0.0};
            double s = .000001;
            for (int c = 0; c < dim; ++c) {
                const double other = (loc[3 + c] >= 0) ? lval(loc[3 + c], xe, glob, na) : rec[2 + c];
                d[c] = lval(loc[c], xe, glob, na) - other;
                s += d[c] * d[c];
            }
            h = sqrt(s);
            const double ih = 1.0 / h;
            int nn = 0;
            for (int c = 0; c < dim; ++c) {
                gl[nn] = loc[c]; gv[nn] = d[c] * ih; ++nn;
                if (loc[3 + c] >= 0) { gl[nn] = loc[3 + c]; gv[nn] = -d[c] * ih; ++nn; }
            }
            return nn;
        }
        case NLPF_RT_HEADING: {   // measurements.py:30-36 (atan2(r_x, r_y): angle from +y)
            const double ry = lval(loc[4], xe, glob, na) - lval(loc[1], xe, glob, na);
            const double rx = lval(loc[3], xe, glob, na) - lval(loc[0], xe, glob, na) + .00001;
            const double id2 = 1.0 / (rx * rx + ry * ry);
            h = atan2(rx, ry);
            gl[0] = loc[0]; gv[0] = -ry * id2;
            gl[1] = loc[1]; gv[1] = rx * id2;
            gl[2] = loc[3]; gv[2] = ry * id2;
            gl[3] = loc[4]; gv[3] = -rx * id2;
            return 4;
        }
        case NLPF_RT_HEADING_PT: {   // measurements.py:25-29
            const double p0 = (loc[3] >= 0) ? lval(loc[3], xe, glob, na) : rec[2];
            const double p1 = (loc[4] >= 0) ? lval(loc[4], xe, glob, na) : rec[3];
            const double ry = p1 - lval(loc[1], xe, glob, na);
            const double rx = p0 - lval(loc[0], xe, glob, na);
            const double id2 = 1.0 / (rx * rx + ry * ry);
            h = atan2(rx, ry);
            int nn = 0;
            gl[nn] = loc[0]; gv[nn] = -ry * id2; ++nn;
            gl[nn] = loc[1]; gv[nn] = rx * id2; ++nn;
            if (loc[3] >= 0) { gl[nn] = loc[3]; gv[nn] = ry * id2; ++nn; }
            if (loc[4] >= 0) { gl[nn] = loc[4]; gv[nn] = -rx * id2; ++nn; }
            return nn;
        }
    }
    h = 0.0;
    return 0;
}

// ------------------------------------------------------------------------------------------
// Per-window pointers
// ------------------------------------------------------------------------------------------
struct WinPtr {
    const double* U;       // K*m   or NULL
    const double* data;    // Ldata
    const double* xprior;  // n     or NULL
    const double* wshift;  // K*n   or NULL
    double* hlin;          // np    constant H of this window (Hc + linear measurements) or NULL
};

__device__ __forceinline__ int pk(int r, int c) { return ((r * (r + 1)) >> 1) + c; }

// ------------------------------------------------------------------------------------------
// Kernels (1)+(2): objective, merit and -- when FULL -- gradient g and Gauss-Newton matrix H
// (packed, in shared memory) of one window at the iterate zc (shared memory).
// Returns J and merit = J/2 + barrier (identical in every thread).  merit = +inf flags an
// iterate outside the bounds or a non-finite objective.
// ------------------------------------------------------------------------------------------
// hmode: 0 = assemble everything; 1 = all measurements are linear (RT_STATE) and their constant
// contribution to H is already folded into wp.hlin (built once per window by build_hlin): start
// from it and skip the O(E K^2 na^2) accumulation (ncu: 24 % of the VDP kernel's samples).
template <bool FULL>
__device__ void eval_window(const DevProb& P, double* sm, const double* zc, const WinPtr& wp, double mu,
                            double& J_out, double& merit_out, double& noise_out, int hmode = 0, bool* clipped_out = nullptr) {
    const int tid = threadIdx.x;
    const int K = P.K, n = P.n, m = P.m, nv = P.nv, ng = P.ng, nz = P.nz, na = P.na, nat = P.nat;
    double* Hp = sm + P.o_H;
    double* g = sm + P.o_g;
    double* fbuf = sm + P.o_f;     // K*n  f(X_k,U_k)
    double* lw = sm + P.o_lw;      // K*n  weight * W   (half gradient of the dynamics cost w.r.t. W)
    double* wh = sm + P.o_wh;      // K*n  curvature weight
    double* slw = sm + P.o_slw;    // K*nv S^T lw
    double* Fs = sm + P.o_F;       // K*n*nv
    double* GF = sm + P.o_GF;      // K*nv*nv  S^T Lambda_k F_k
    double* FF = sm + P.o_FF;      // K*nv*nv  F_k^T Lambda_k F_k
    double* whs = sm + P.o_whs;    // K*nv     S^T Lambda_k S (diagonal), Huber only
    double* Ve = sm + P.o_Ve;
    double* G = sm + P.o_G;
    double* q = sm + P.o_q;
    const double* Phi = P.const_smem ? sm + P.o_Phi : P.Phi;
    const double* Dt = P.const_smem ? sm + P.o_Dt : P.Dt;
    double* red = sm + P.o_red;
    const bool dyn_h = FULL && !P.hconst;            // H has a state-dependent dynamics part
    const bool need_F = dyn_h;                       // ... which needs dF/dv per node in shared memory

    double clipped = 0.0;   // 1: the second-order term of one of this thread's nodes needed its safeguard
    double cost = 0.0;      // thread-local share of J
    double bar = 0.0;       // thread-local share of the barrier
    double nse = 0.0;       // thread-local share of the first-order rounding bound of J (in units of eps)
    bool bad = false;

    // ---- phase A: H <- constant part, g <- 0 ----
    if (FULL) {
        const double* H0 = (hmode == 1) ? wp.hlin : P.Hc;
        for (int p = tid; p < P.np; p += NLPF_NT) Hp[p] = H0[p];
        for (int i = tid; i < nz; i += NLPF_NT) g[i] = 0.0;
    }
    // ---- phase B: node dynamics f (and dF/dv) ----
    if (P.has_dyncost) {
        // affine models whose H part is constant need no per-node pass: thread (k,i) of phase C forms
        // f_ki = c0_i + A_i. x_k + B_i. u_k itself (ncu: 8 % of the GNSS kernel waited here for 11 busy threads)
#if NLPF_PR_RUNS
        const bool affine_inline = P.fconst && !need_F;
#else
        const bool affine_inline = false;      // kept out of this instantiation: it cost the van der Pol kernel 10 % (register allocation)
#endif
        for (int k = affine_inline ? K : tid; k < K; k += NLPF_NT) {
            double x[NLPF_N_MAX], u[NLPF_M_MAX], f[NLPF_N_MAX];
            for (int i = 0; i < n; ++i) x[i] = zc[k * nv + P.cmap[i]];
            for (int i = 0; i < m; ++i)
                u[i] = (P.ucol[i] >= 0) ? zc[k * nv + P.ucol[i]] : (wp.U ? wp.U[k * m + i] : 0.0);
            double* Fk = Fs + k * n * nv;
            if (need_F)
                for (int i = 0; i < n * nv; ++i) Fk[i] = 0.0;
            dyn_eval(P, x, u, f, need_F, Fk);
            for (int i = 0; i < n; ++i) fbuf[k * n + i] = f[i];
        }
        if (!affine_inline) __syncthreads();
        // ---- phase C: defects W, weights ----
        for (int t = tid; t < K * n; t += NLPF_NT) {
            const int k = t / n, i = t - k * n;
            const int col = P.cmap[i];
            double fki;
            if (affine_inline) {
                fki = P.c0[i];
                for (int j = 0; j < n; ++j) fki = fma(P.A[i * n + j], zc[k * nv + P.cmap[j]], fki);
                for (int j = 0; j < m; ++j) {
                    const double uj = (P.ucol[j] >= 0) ? zc[k * nv + P.ucol[j]] : (wp.U ? wp.U[k * m + j] : 0.0);
                    fki = fma(P.Bm[i * m + j], uj, fki);
                }
            } else {
                fki = fbuf[t];
            }
            double W = -fki;
            double mag = fabs(fki);
            for (int j = 0; j < K; ++j) {
                const double dx = Dt[k * K + j] * zc[j * nv + col];
                W += dx;
                mag += fabs(dx);
            }
            if (wp.wshift) W += wp.wshift[t];
            const double cq = P.cw[k] * P.qinv[i];
            if (P.loss == NLPF_LOSS_HUBER) {
                const double wd = W / P.delta;
                const double s = sqrt(1.0 + wd * wd);
                cost += cq * 2.0 * P.delta * P.delta * (s - 1.0);
                lw[t] = cq * W / s;
                wh[t] = cq * (s + 1.0) / (2.0 * s * s);
            } else {
                cost += cq * W * W;
                lw[t] = cq * W;
                wh[t] = cq;
            }
            nse += 2.0 * fabs(lw[t]) * mag;
        }
        __syncthreads();
        if (FULL) {
            // ---- phase D: S^T lw, and the per-node blocks GF, FF (and S^T Lambda S) ----
            for (int t = tid; t < K * nv; t += NLPF_NT) {
                const int k = t / nv, a = t - k * nv;
                double s = 0.0, sw = 0.0;
                for (int i = 0; i < n; ++i)
                    if (P.cmap[i] == a) { s += lw[k * n + i]; sw += wh[k * n + i]; }
                slw[t] = s;
                whs[t] = sw;
            }
            if (dyn_h) {
                for (int t = tid; t < K * nv * nv; t += NLPF_NT) {
                    const int k = t / (nv * nv);
                    const int rem = t - k * nv * nv;
                    const int a = rem / nv, b = rem - a * nv;
                    const double* Fk = Fs + k * n * nv;
                    double sgf = 0.0, sff = 0.0;
                    for (int i = 0; i < n; ++i) {
                        const double w_i = wh[k * n + i];
                        const double fib = Fk[i * nv + b];
                        if (P.cmap[i] == a) sgf = fma(w_i, fib, sgf);          // GF[a][b] = sum_{cmap[i]=a} w F[i][b]
                        sff = fma(Fk[i * nv + a] * w_i, fib, sff);             // FF[a][b]
                    }
                    GF[t] = sgf;
                    FF[t] = sff;
                }
            }
            __syncthreads();
            // ---- phase E: gradient of the dynamics cost, arrival cost, node cost ----
            for (int t = tid; t < K * nv; t += NLPF_NT) {
                const int j = t / nv, a = t - j * nv;
                double s = 0.0;
                for (int k = 0; k < K; ++k) s = fma(Dt[k * K + j], slw[k * nv + a], s);
                const double* Fj = need_F ? (Fs + j * n * nv) : P.Fc;
                for (int i = 0; i < n; ++i) s = fma(-Fj[i * nv + a], lw[j * n + i], s);
                g[t] += s;
            }
            // state-dependent dynamics part of H
            if (dyn_h) {
                const bool hub = (P.loss == NLPF_LOSS_HUBER);
                for (int p = tid; p < P.np; p += NLPF_NT) {
                    const uint32_t e = P.tab[p];
                    const int j = e & 255, a = (e >> 8) & 255, l = (e >> 16) & 255, b = e >> 24;
                    if (j == 255) continue;                  // global rows carry no dynamics
                    double s = -Dt[j * K + l] * GF[(j * nv + b) * nv + a] - Dt[l * K + j] * GF[(l * nv + a) * nv + b];
                    if (j == l) s += FF[(j * nv + a) * nv + b];
                    if (hub && a == b) {
                        double s1 = 0.0;
                        for (int k = 0; k < K; ++k) s1 = fma(Dt[k * K + j] * Dt[k * K + l], whs[k * nv + a], s1);
                        s += s1;
                    }
                    Hp[p] += s;
                }
            }
        }
    }
    // arrival cost (nlp.py:279-286) and node cost (nlp.py:160-168): cost here, gradient below
    if (P.has_prior) {
        for (int i = tid; i < n; i += NLPF_NT) {
            const double xp = wp.xprior ? wp.xprior[i] : 0.0;
            const double d0 = zc[P.cmap[i]] - xp;
            cost += P.pinv[i] * d0 * d0;
            nse += 2.0 * P.pinv[i] * fabs(d0) * (fabs(zc[P.cmap[i]]) + fabs(xp));
        }
    }
    if (P.has_nodecost) {
        for (int t = tid; t < K * nv; t += NLPF_NT) {
            const int k = t / nv, a = t - k * nv;
            cost += P.cq[k] * P.snode[a] * zc[t] * zc[t];
        }
    }
    if (FULL) {
        __syncthreads();
        if (P.has_prior) {
            for (int a = tid; a < nv; a += NLPF_NT) {
                double s = 0.0;
                for (int i = 0; i < n; ++i)
                    if (P.cmap[i] == a) s += P.pinv[i] * (zc[a] - (wp.xprior ? wp.xprior[i] : 0.0));
                g[a] += s;
            }
        }
        if (P.has_nodecost) {
            __syncthreads();
            for (int t = tid; t < K * nv; t += NLPF_NT) {
                const int k = t / nv, a = t - k * nv;
                g[t] += P.cq[k] * P.snode[a] * zc[t];
            }
        }
    }
    // ---- phase G: measurement residuals, epoch chunk by epoch chunk ----
    const double* glob = zc + K * nv;
    for (int c0 = 0; c0 < P.E; c0 += P.ec) {
        const int ne = min(P.ec, P.E - c0);
        __syncthreads();
        for (int t = tid; t < ne * na; t += NLPF_NT) {
            const int el = t / na, a = t - el * na;
            const double* ph = Phi + (c0 + el) * K;
            const int col = P.act[a];
            double s = 0.0;
            for (int j = 0; j < K; ++j) s = fma(ph[j], zc[j * nv + col], s);
            Ve[t] = s;
        }
        const bool accH = FULL && hmode == 0;
        if (FULL) {
            if (accH)
                for (int t = tid; t < ne * nat * nat; t += NLPF_NT) G[t] = 0.0;
            for (int t = tid; t < ne * nat; t += NLPF_NT) q[t] = 0.0;
        }
        __syncthreads();
#if NLPF_PR_RUNS
        // Residuals of the chunk.  A group of LG lanes shares one epoch: pseudorange "runs" (consecutive
        // RT_PR residuals with the same index quadruple -- the bulk of every GNSS problem) are split
        // over the lanes, accumulated in registers (4 gradient + 10 outer-product values) and
        // combined with LG-wide shuffles; the lane that owns the epoch adds them to G_e / q_e.
        // (ncu, thread-per-epoch version: 24 % of the GNSS kernel's samples waited at the barrier
        // behind 42 busy threads of 128.)  Other residual types go through meas_eval on the owner lane.
        {
            const int LG = (ne <= NLPF_NT / 4) ? 4 : ((ne <= NLPF_NT / 2) ? 2 : 1);
            const int sub = tid & (LG - 1);
            const int ngrp = NLPF_NT / LG;
            for (int el0 = 0; el0 < ne; el0 += ngrp) {
                const int el = el0 + tid / LG;
                const bool live = el < ne;                      // whole groups are live or not (shuffles stay converged)
                const double* xe = Ve + (live ? el : 0) * na;
                double* Ge = G + (live ? el : 0) * nat * nat;
                double* qe = q + (live ? el : 0) * nat;
                const int e = c0 + (live ? el : 0);
                const int rbeg = live ? P.erun[e] : 0;
                const int nr = live ? P.erun[e + 1] - rbeg : 0;
                int nrmax = nr;                                 // same trip count for the whole warp: the shuffles
                for (int o = 16; o > 0; o >>= 1) nrmax = max(nrmax, __shfl_xor_sync(0xffffffffu, nrmax, o));   // below need it converged
                for (int kr = 0; kr < nrmax; ++kr) {
                    const bool act = kr < nr;
                    const int rn = rbeg + (act ? kr : 0);
                    const int i0 = act ? P.run_start[rn] : 0, i1 = act ? i0 + P.run_len[rn] : 0;
                    const int l0 = act ? P.run_loc[4 * rn] : 0, l1 = act ? P.run_loc[4 * rn + 1] : 0;
                    const int l2 = act ? P.run_loc[4 * rn + 2] : 0, l3 = act ? P.run_loc[4 * rn + 3] : 0;
                    const double x0 = lval(l0, xe, glob, na), x1 = lval(l1, xe, glob, na), x2 = lval(l2, xe, glob, na);
                    const double xb = lval(l3, xe, glob, na);
                    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;                       // sum w r grad
                    double g00 = 0, g10 = 0, g11 = 0, g20 = 0, g21 = 0, g22 = 0, g30 = 0, g31 = 0, g32 = 0, g33 = 0;
                    for (int i = i0 + sub; i < i1; i += LG) {
                        const double* rec = wp.data + P.roff[i];
                        const double w = rec[1];
                        if (w == 0.0) continue;                 // masked slot: skipped, never 0 * NaN
                        const double d0 = x0 - rec[2], d1 = x1 - rec[3], d2 = x2 - rec[4];
                        const double rho = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
                        const double r = rec[0] - (rho + xb);
                        cost += w * r * r;
                        nse += 2.0 * w * fabs(r) * (2.0 * fabs(rec[0]) + fabs(r));
                        if (FULL) {
                            const double ir = 1.0 / rho;
                            const double e0 = d0 * ir, e1 = d1 * ir, e2 = d2 * ir;
                            const double wr = w * r;
                            a0 = fma(wr, e0, a0); a1 = fma(wr, e1, a1); a2 = fma(wr, e2, a2); a3 += wr;
                            if (accH) {
                                const double w0 = w * e0, w1 = w * e1, w2 = w * e2;
                                g00 = fma(w0, e0, g00); g10 = fma(w1, e0, g10); g11 = fma(w1, e1, g11);
                                g20 = fma(w2, e0, g20); g21 = fma(w2, e1, g21); g22 = fma(w2, e2, g22);
                                g30 += w0; g31 += w1; g32 += w2; g33 += w;
                            }
                        }
                    }
                    if (FULL) {
                        for (int o = LG >> 1; o > 0; o >>= 1) {
                            a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o);
                            a2 += __shfl_xor_sync(0xffffffffu, a2, o); a3 += __shfl_xor_sync(0xffffffffu, a3, o);
                            if (accH) {
                                g00 += __shfl_xor_sync(0xffffffffu, g00, o); g10 += __shfl_xor_sync(0xffffffffu, g10, o);
                                g11 += __shfl_xor_sync(0xffffffffu, g11, o); g20 += __shfl_xor_sync(0xffffffffu, g20, o);
                                g21 += __shfl_xor_sync(0xffffffffu, g21, o); g22 += __shfl_xor_sync(0xffffffffu, g22, o);
                                g30 += __shfl_xor_sync(0xffffffffu, g30, o); g31 += __shfl_xor_sync(0xffffffffu, g31, o);
                                g32 += __shfl_xor_sync(0xffffffffu, g32, o); g33 += __shfl_xor_sync(0xffffffffu, g33, o);
                            }
                        }
                        if (sub == 0 && act) {
                            qe[l0] += a0; qe[l1] += a1; qe[l2] += a2; qe[l3] += a3;
                            if (accH) {
                                // symmetric 4x4 block scattered to (l_p, l_q); += so that coinciding indices add up
                                Ge[l0 * nat + l0] += g00;
                                Ge[l1 * nat + l0] += g10; Ge[l0 * nat + l1] += g10; Ge[l1 * nat + l1] += g11;
                                Ge[l2 * nat + l0] += g20; Ge[l0 * nat + l2] += g20; Ge[l2 * nat + l1] += g21; Ge[l1 * nat + l2] += g21;
                                Ge[l2 * nat + l2] += g22;
                                Ge[l3 * nat + l0] += g30; Ge[l0 * nat + l3] += g30; Ge[l3 * nat + l1] += g31; Ge[l1 * nat + l3] += g31;
                                Ge[l3 * nat + l2] += g32; Ge[l2 * nat + l3] += g32; Ge[l3 * nat + l3] += g33;
                            }
                        }
                    }
                }
                if (sub == 0 && live) {          // residuals of other types: owner lane, general path
                    const int i1 = P.estart[e + 1];
                    for (int i = P.gstart[e]; i < i1; ++i) {
                        const double* rec = wp.data + P.roff[i];
                        const double w = rec[1];
                        if (w == 0.0) continue;
                        int gl[8];
                        double gv[8];
                        double h;
                        const int nn = meas_eval(P.rtype[i], P.rloc + i * NLPF_NIDX, rec, xe, glob, na, h, gl, gv);
                        const double r = rec[0] - h;
                        cost += w * r * r;
                        nse += 2.0 * w * fabs(r) * (2.0 * fabs(rec[0]) + fabs(r));
                        if (FULL) {
                            for (int p = 0; p < nn; ++p) {
                                const double wg = w * gv[p];
                                qe[gl[p]] = fma(wg, r, qe[gl[p]]);
                                if (accH)
                                    for (int p2 = 0; p2 < nn; ++p2) Ge[gl[p] * nat + gl[p2]] = fma(wg, gv[p2], Ge[gl[p] * nat + gl[p2]]);
                            }
                        }
                    }
                }
            }
        }
#else
        for (int el = tid; el < ne; el += NLPF_NT) {          // thread per epoch, every residual through meas_eval
            const double* xe = Ve + el * na;
            double* Ge = G + el * nat * nat;
            double* qe = q + el * nat;
            const int i1 = P.estart[c0 + el + 1];
            for (int i = P.estart[c0 + el]; i < i1; ++i) {
                const double* rec = wp.data + P.roff[i];
                const double w = rec[1];
                if (w == 0.0) continue;                   // masked slot: skipped, never 0 * NaN
                int gl[8];
                double gv[8];
                double h;
                const int nn = meas_eval(P.rtype[i], P.rloc + i * NLPF_NIDX, rec, xe, glob, na, h, gl, gv);
                const double r = rec[0] - h;
                cost += w * r * r;
                nse += 2.0 * w * fabs(r) * (2.0 * fabs(rec[0]) + fabs(r));
                if (FULL) {
                    for (int p = 0; p < nn; ++p) {
                        const double wg = w * gv[p];
                        qe[gl[p]] = fma(wg, r, qe[gl[p]]);
                        if (accH)
                            for (int p2 = 0; p2 < nn; ++p2) Ge[gl[p] * nat + gl[p2]] = fma(wg, gv[p2], Ge[gl[p] * nat + gl[p2]]);
                    }
                }
            }
        }
#endif
        if (FULL) {
            __syncthreads();
            // node-node blocks: H[(j,a),(l,b)] += sum_e phi_ej phi_el G_e[a][b]
            const int nwork = accH ? P.npair * na : 0;
            if (NLPF_PR_RUNS && na == 4 && nat == 4) {
                // the pseudorange case (x, y, z, clock): exact trip counts, 16-byte loads of the G_e rows
                // (ncu: the predicated 8-wide version of this loop was 14 % of the GNSS kernel's instructions)
                for (int wk = tid; wk < nwork; wk += NLPF_NT) {
                    const int pr = wk >> 2, a = wk & 3;
                    const int pe = P.ptab[pr];
                    const int j = pe & 255, l = pe >> 8;
                    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
                    const double* ph = Phi + c0 * K;
                    const double2* Ga = reinterpret_cast<const double2*>(G + a * 4);
                    for (int el = 0; el < ne; ++el) {
                        const double pp = ph[j] * ph[l];
                        const double2 g01 = Ga[0], g23 = Ga[1];
                        a0 = fma(pp, g01.x, a0);
                        a1 = fma(pp, g01.y, a1);
                        a2 = fma(pp, g23.x, a2);
                        a3 = fma(pp, g23.y, a3);
                        ph += K;
                        Ga += 8;                     // next epoch: 16 doubles
                    }
                    const int r = j * nv + P.act[a];
                    const int cb = l * nv;
                    const int c0_ = cb + P.act[0], c1_ = cb + P.act[1], c2_ = cb + P.act[2], c3_ = cb + P.act[3];
                    if (c0_ <= r) Hp[pk(r, c0_)] += a0;
                    if (c1_ <= r) Hp[pk(r, c1_)] += a1;
                    if (c2_ <= r) Hp[pk(r, c2_)] += a2;
                    if (c3_ <= r) Hp[pk(r, c3_)] += a3;
                }
            } else {
                // work item = (node pair, row a, block of 4 columns b0..b0+3 <= a) of the LOWER triangle of G_e:
                // G_e is symmetric, so block (j,l) of H is too and entry (a,b) also serves (b,a).  (The version
                // that summed all na x na entries in predicated blocks of 8 was 37 % of the two-receiver kernel.)
                const int nsym = accH ? P.nsym : 0;
                // Threads of a warp share the block (consecutive wk = consecutive node pairs), so skipping the epochs
                // at which the block is structurally zero (P.gnz) is a warp-uniform branch.
                for (int wk = tid; wk < P.npair * nsym; wk += NLPF_NT) {
                    const int item = wk / P.npair;
                    const int pr = wk - item * P.npair;
                    const int se = P.sym_tab[item];
                    const uint8_t* nzp = P.gnz + (size_t)item * P.E + c0;
                    const int a = se & 255, b0 = se >> 8;
                    const int pe = P.ptab[pr];
                    const int j = pe & 255, l = pe >> 8;
                    double acc[4] = {0.0, 0.0, 0.0, 0.0};
                    const double* ph = Phi + c0 * K;
                    const double* Ga = G + a * nat + b0;
                    const int nb4 = (a - b0 + 1 < 4) ? a - b0 + 1 : 4;
                    if (nb4 == 4) {
                        for (int el = 0; el < ne; ++el, ph += K, Ga += nat * nat) {
                            if (!nzp[el]) continue;
                            const double pp = ph[j] * ph[l];
                            acc[0] = fma(pp, Ga[0], acc[0]);
                            acc[1] = fma(pp, Ga[1], acc[1]);
                            acc[2] = fma(pp, Ga[2], acc[2]);
                            acc[3] = fma(pp, Ga[3], acc[3]);
                        }
                    } else {
                        for (int el = 0; el < ne; ++el, ph += K, Ga += nat * nat) {
                            if (!nzp[el]) continue;
                            const double pp = ph[j] * ph[l];
                            acc[0] = fma(pp, Ga[0], acc[0]);
                            if (nb4 > 1) acc[1] = fma(pp, Ga[1], acc[1]);
                            if (nb4 > 2) acc[2] = fma(pp, Ga[2], acc[2]);
                        }
                    }
                    const int ca = P.act[a];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (q < nb4) {
                            const int cb = P.act[b0 + q];
                            const int r1 = j * nv + ca, k1 = l * nv + cb;       // (a,b)
                            const int r2 = j * nv + cb, k2 = l * nv + ca;       // its mirror (b,a)
                            if (k1 <= r1) Hp[pk(r1, k1)] += acc[q];
                            if (b0 + q != a && k2 <= r2 && !(j == l && k1 <= r1)) Hp[pk(r2, k2)] += acc[q];
                        }
                    }
                }
            }
            // gradient: g[(j,a)] -= sum_e phi_ej q_e[a]
            for (int t = tid; t < K * na; t += NLPF_NT) {
                const int j = t / na, a = t - j * na;
                double s = 0.0;
                for (int el = 0; el < ne; ++el) s = fma(Phi[(c0 + el) * K + j], q[el * nat + a], s);
                g[j * nv + P.act[a]] -= s;
            }
            // global unknowns: rows K*nv + gi
            if (ng > 0 && accH) {
                for (int t = tid; t < ng * K * na; t += NLPF_NT) {
                    const int gi = t / (K * na);
                    const int rem = t - gi * K * na;
                    const int l = rem / na, b = rem - l * na;
                    double s = 0.0;
                    for (int el = 0; el < ne; ++el) s = fma(Phi[(c0 + el) * K + l], G[(el * nat + na + gi) * nat + b], s);
                    Hp[pk(K * nv + gi, l * nv + P.act[b])] += s;
                }
                for (int t = tid; t < ng * ng; t += NLPF_NT) {
                    const int gi = t / ng, gj = t - gi * ng;
                    if (gj > gi) continue;
                    double s = 0.0;
                    for (int el = 0; el < ne; ++el) s += G[(el * nat + na + gi) * nat + na + gj];
                    Hp[pk(K * nv + gi, K * nv + gj)] += s;
                }
                for (int gi = tid; gi < ng; gi += NLPF_NT) {
                    double s = 0.0;
                    for (int el = 0; el < ne; ++el) s += q[el * nat + na + gi];
                    g[K * nv + gi] -= s;
                }
            }
        }
    }
    // ---- phase H: log barrier on bounded components ----
    if (mu > 0.0 && P.nb > 0) {
        if (FULL) __syncthreads();
        for (int t = tid; t < K * P.nb; t += NLPF_NT) {
            const int k = t / P.nb, b = t - k * P.nb;
            const int idx = k * nv + P.bidx[b];
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
    // ---- phase S: damping diagonal (Gauss-Newton diagonal, always positive) and the exact
    //      second-order part of the dynamics cost  sum_i lw_ki * (-hess f_i(X_k))  for the models
    //      that carry it (Gauss-Newton alone is only linearly convergent on large-defect windows) ----
    if (FULL) {
        __syncthreads();
        double* dd = sm + P.o_dd;
        for (int i = tid; i < nz; i += NLPF_NT) dd[i] = Hp[pk(i, i)];
        if (P.has_dyncost && (P.dyn_id == NLPF_DYN_VDP || P.dyn_id == NLPF_DYN_BICYCLE)) {
            __syncthreads();
            for (int k = tid; k < K; k += NLPF_NT) {
                const int o = k * nv;
                if (P.dyn_id == NLPF_DYN_VDP) {
                    // f0 = (1 - x1^2) x0 - x1 + u:  hess f0 = [[0, -2 x1], [-2 x1, -2 x0]]
                    const int c0 = P.cmap[0], c1 = P.cmap[1];
                    const double x0 = zc[o + c0], x1 = zc[o + c1];
                    const double l0 = lw[k * n + 0];
                    const int rhi = o + max(c0, c1), rlo = o + min(c0, c1);
                    const double s01 = l0 * 2.0 * x1, s11 = l0 * 2.0 * x0;
                    // safeguard: the node's block is scaled back until the Gauss-Newton diagonal dominates its rows
                    // (Gershgorin; t = 1 near the solution, where lw is small) -- keeps H definite away from it
                    const double t = fmin(1.0, fmin(Hp[pk(o + c0, o + c0)] / fabs(s01), Hp[pk(o + c1, o + c1)] / (fabs(s01) + fabs(s11))));
                    Hp[pk(rhi, rlo)] += t * s01;
                    Hp[pk(o + c1, o + c1)] += t * s11;
                    if (t < 1.0) clipped = 1.0;
                } else {
                    // f0 = v cos th, f1 = v sin th, th = x[2]
                    const int c2 = P.cmap[2];
                    const double u0 = (P.ucol[0] >= 0) ? zc[o + P.ucol[0]] : (wp.U ? wp.U[k * m + 0] : 0.0);
                    const double v = 8.72649116358 * u0 - 0.856053299155;
                    double s2, c2v;
                    sincos(zc[o + c2], &s2, &c2v);
                    const double s22 = lw[k * n + 0] * v * c2v + lw[k * n + 1] * v * s2;
                    const double t = fmin(1.0, Hp[pk(o + c2, o + c2)] / fabs(s22));      // same safeguard
                    Hp[pk(o + c2, o + c2)] += t * s22;
                    if (t < 1.0) clipped = 1.0;
                }
            }
        }
    }
    // ---- phase I: reductions ----
    double rv[5] = {cost, bar, nse, bad ? 1.0 : 0.0, clipped};
    block_reduce<3, 2>(rv, red);
    const double J = rv[0], B = rv[1], NS = rv[2], anybad = rv[3];
    if (clipped_out) *clipped_out = rv[4] > 0.0;
    double merit = 0.5 * J + B;
    if (anybad > 0.0 || !isfinite(merit)) merit = INFINITY;
    J_out = J;
    merit_out = merit;
    noise_out = 2.220446049250313e-16 * (0.5 * NS + fabs(merit));
    __syncthreads();
}

// Out-of-line copy of the full evaluation for the rarely taken call sites of the solve loop (first
// evaluation of a window, re-assembly after a rejected step, new barrier stage).  Inlining all five
// sites gave a 1.7 KB stack frame and five copies of the assembly code in the instruction stream.
__device__ __noinline__ void eval_window_cold(const DevProb& P, double* sm, const double* zc, const WinPtr& wp, double mu,
                                              double& J_out, double& merit_out, double& noise_out, int hmode, bool* clipped_out) {
    eval_window<true>(P, sm, zc, wp, mu, J_out, merit_out, noise_out, hmode, clipped_out);
}

// ------------------------------------------------------------------------------------------
// Kernel (3): LDL^T of the packed matrix with the right-hand side appended as row nz.
// On entry Hp holds the lower triangle of H (+ damping) and row nz holds g.  The elimination
// leaves the unscaled factor (column j still multiplied by d_j) and, in row nz, L^-1 g; the
// backward sweep then produces x = H^-1 g in xs.  Returns false when a pivot is not positive.
//
// Right-looking with panels of 4 columns (ncu of the column-at-a-time version: 33 % of the
// kernel's stall samples sat behind its per-column barrier with <= nz of 128 threads busy):
//   1. every thread factors the 4x4 diagonal block redundantly in registers (no barrier inside
//      the panel), 2. one thread per row forms the panel entries w_i = A_i,J L11^-T,
//   3. the trailing update A_ik -= sum_c (w_ic / d_c) w_kc is cut into (row, 8-column segment)
//      items from a host-built table so that all 128 threads share it evenly; w_i stays in
//      registers, 4 FMAs per 6 shared-memory accesses instead of 1 per 3.
// Two barriers per 4 columns; no square roots.
// ------------------------------------------------------------------------------------------
#if NLPF_BIG
// 128-register instantiation: the tiled variant of the warp-per-window kernel (wsolve_kernels.cu: ws_ldl) on the
// plain packed layout -- scaled factors l = w / d in place, the unscaled panel W^T in a side buffer (the
// evaluation's scratch, free during the factorisation), 4-row x 2-column register tiles for the trailing
// update (32 independent FMAs per tile instead of a load -> 4 FMA -> store chain per entry), and a backward
// sweep without divisions.  The 64-register instantiation keeps the per-entry version below (tiles spill there).
__device__ bool ldl_solve(const DevProb& P, int nz, double* Hp, double* xs) {
    const int tid = threadIdx.x;
    double* Wt = Hp + (P.o_f - P.o_H);
    const int LD = (nz + 5) & ~1;
    bool ok = true;
    for (int j0 = 0; j0 < nz; j0 += 4) {
        const int pw = min(4, nz - j0);
        __syncthreads();
        double a[4][4], l[4][4], d[4], rd[4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c <= r; ++c) a[r][c] = (r < pw) ? Hp[pk(j0 + r, j0 + c)] : (r == c ? 1.0 : 0.0);
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
            rd[c] = 1.0 / d[c];
#pragma unroll
            for (int r = c + 1; r < 4; ++r) l[r][c] = a[r][c] * rd[c];
        }
        bool good = true;
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (c < pw && (!(d[c] > 0.0) || !isfinite(d[c]))) good = false;
        if (!good) { ok = false; break; }
        const int j1 = j0 + pw;
        for (int i = j1 + tid; i <= nz; i += NLPF_NT) {
            double* Hi = Hp + pk(i, j0);
            double w[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (c < pw) {
                    double v = Hi[c];
#pragma unroll
                    for (int q = 0; q < c; ++q) v = fma(-w[q], l[c][q], v);
                    w[c] = v;
                    Hi[c] = v * rd[c];
                } else w[c] = 0.0;
                Wt[c * LD + i] = w[c];
            }
        }
        __syncthreads();
        if (tid == NLPF_NT - 1) {       // the block in the scaled convention: unit-lower l below the diagonal
#pragma unroll
            for (int r = 1; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < r; ++c)
                    if (r < pw) Hp[pk(j0 + r, j0 + c)] = l[r][c];
        }
        if (j1 >= nz) continue;
        const int nitem = P.tile_cnt[nz + 1 - j1];
        for (int it = tid; it < nitem; it += NLPF_NT) {
            const int e = P.tile_tab[it];
            const int i = j1 + ((e & 255) << 2);
            const int k0 = j1 + ((e >> 8) << 1);
            double* R[4];
            R[0] = Hp + pk(i, 0);
            R[1] = R[0] + i + 1;
            R[2] = R[1] + i + 2;
            R[3] = R[2] + i + 3;
            double x[4][2], wv[4][2];
#pragma unroll
            for (int r = 0; r < 4; ++r) { x[r][0] = R[r][k0]; x[r][1] = R[r][k0 + 1]; }
#pragma unroll
            for (int c = 0; c < 4; ++c) { wv[c][0] = Wt[c * LD + k0]; wv[c][1] = Wt[c * LD + k0 + 1]; }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const double lv = R[r][j0 + c];
                    x[r][0] = fma(-lv, wv[c][0], x[r][0]);
                    x[r][1] = fma(-lv, wv[c][1], x[r][1]);
                }
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                if (i + r <= nz) {
                    if (k0 <= i + r && k0 < nz) R[r][k0] = x[r][0];
                    if (k0 + 1 <= i + r && k0 + 1 < nz) R[r][k0 + 1] = x[r][1];
                }
            }
        }
    }
    __syncthreads();
    if (!ok) return false;
    // backward sweep by the first warp: y = scaled appended row; x_j = y_j, y_i -= l_ji x_j (i < j)
    double* y = Hp + pk(nz, 0);
    if (tid < 32) {
        for (int j = nz - 1; j >= 0; --j) {
            const double* Hj = Hp + pk(j, 0);
            const double xj = y[j];
            for (int i = tid; i < j; i += 32) y[i] = fma(-Hj[i], xj, y[i]);
            if (tid == 0) xs[j] = xj;
            __syncwarp();
        }
    }
    __syncthreads();
    return true;
}
#else
__device__ bool ldl_solve(const DevProb& P, int nz, double* Hp, double* xs) {
    const int tid = threadIdx.x;
    bool ok = true;
    for (int j0 = 0; j0 < nz; j0 += 4) {
        const int pw = min(4, nz - j0);
        __syncthreads();
        // ---- 1. diagonal block, redundantly per thread ----
        double a[4][4], l[4][4], d[4], rd[4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c <= r; ++c) a[r][c] = (r < pw) ? Hp[pk(j0 + r, j0 + c)] : (r == c ? 1.0 : 0.0);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            // w_rc = a_rc - sum_{q<c} w_rq l_cq  (stored back into a), d_c = w_cc
#pragma unroll
            for (int r = c; r < 4; ++r) {
                double w = a[r][c];
#pragma unroll
                for (int q = 0; q < c; ++q) w = fma(-a[r][q], l[c][q], w);
                a[r][c] = w;
            }
            d[c] = a[c][c];
            rd[c] = 1.0 / d[c];
#pragma unroll
            for (int r = c + 1; r < 4; ++r) l[r][c] = a[r][c] * rd[c];
        }
        bool good = true;
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (c < pw && (!(d[c] > 0.0) || !isfinite(d[c]))) good = false;
        if (!good) { ok = false; break; }           // uniform: every thread factored the same block
        const int j1 = j0 + pw;
        // ---- 2. panel rows: w_i = a_i,J L11^-T (unscaled storage), one thread per row ----
        for (int i = j1 + tid; i <= nz; i += NLPF_NT) {
            double* Hi = Hp + ((i * (i + 1)) >> 1) + j0;
            double w[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (c < pw) {
                    double v = Hi[c];
#pragma unroll
                    for (int q = 0; q < c; ++q) v = fma(-w[q], l[c][q], v);
                    w[c] = v;
                    Hi[c] = v;
                } else w[c] = 0.0;
            }
        }
        __syncthreads();
        if (tid == NLPF_NT - 1) {                    // the block itself, in the same storage convention; written
#pragma unroll                                       // only after every thread has read it (nobody reads it again
            for (int r = 0; r < 4; ++r)              // before the backward sweep)
#pragma unroll
                for (int c = 0; c <= r; ++c)
                    if (r < pw) Hp[pk(j0 + r, j0 + c)] = a[r][c];
        }
        // ---- 3. trailing update over (row, 8-column segment) items ----
        const int s = nz + 1 - j1;                   // rows j1 .. nz
        const int nitem = P.seg_cnt[s];
        for (int it = tid; it < nitem; it += NLPF_NT) {
            const int e = P.seg_tab[it];
            const int i = j1 + (e & 255);
            const int k0 = j1 + ((e >> 8) << 3);
            const int kend = min(min(k0 + 7, i), nz - 1);      // the appended row has no diagonal entry
            double* Hi = Hp + ((i * (i + 1)) >> 1);
            double li[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) li[c] = (c < pw) ? Hi[j0 + c] * rd[c] : 0.0;
            // (a two-columns-per-step version with the loads hoisted above the stores was 5-7 % SLOWER here: the
            // 64-register cap turns the extra live values into spills; the warp kernel has the tiled version)
            for (int k = k0; k <= kend; ++k) {
                const double* Hk = Hp + ((k * (k + 1)) >> 1) + j0;
                double v = Hi[k];
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (c < pw) v = fma(-li[c], Hk[c], v);
                Hi[k] = v;
            }
        }
    }
    __syncthreads();
    if (!ok) return false;
    // backward sweep by the first warp: u = row nz; x_j = u_j / d_j; u_i -= Hraw[j][i] x_j (i < j)
    double* u = Hp + ((nz * (nz + 1)) >> 1);
    if (tid < 32) {
        for (int j = nz - 1; j >= 0; --j) {
            const double* Hj = Hp + ((j * (j + 1)) >> 1);
            const double xj = u[j] / Hj[j];
            for (int i = tid; i < j; i += 32) u[i] = fma(-Hj[i], xj, u[i]);
            if (tid == 0) xs[j] = xj;
            __syncwarp();
        }
    }
    __syncthreads();
    return true;
}
#endif

// ------------------------------------------------------------------------------------------
// shared-memory staging common to the kernels
// ------------------------------------------------------------------------------------------
__device__ void stage_constants(const DevProb& P, double* sm) {
    if (!P.const_smem) return;
    for (int i = threadIdx.x; i < P.E * P.K; i += NLPF_NT) sm[P.o_Phi + i] = P.Phi[i];
    for (int i = threadIdx.x; i < P.K * P.K; i += NLPF_NT) sm[P.o_Dt + i] = P.Dt[i];
}

__device__ __forceinline__ WinPtr window_ptrs(const DevProb& P, int b, const double* U, const double* data,
                                              const double* xprior, const double* wshift) {
    WinPtr wp;
    wp.U = (U && P.m > 0) ? U + (size_t)b * P.K * P.m : nullptr;
    wp.data = data ? data + (size_t)b * P.Ldata : nullptr;
    wp.xprior = xprior ? xprior + (size_t)b * P.n : nullptr;
    wp.wshift = wshift ? wshift + (size_t)b * P.K * P.n : nullptr;
    wp.hlin = nullptr;
    return wp;
}

// ------------------------------------------------------------------------------------------
// nlpf_assemble: evaluation only (tests, ncu)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NLPF_NT)
VN(assemble_kernel)(DevProb P, int B, const double* __restrict__ z, const double* __restrict__ U,
                const double* __restrict__ data, const double* __restrict__ xprior,
                const double* __restrict__ wshift, double mu, double* __restrict__ H_out,
                double* __restrict__ g_out, double* __restrict__ cost_out) {
    extern __shared__ double sm[];
    stage_constants(P, sm);
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        __syncthreads();
        double* zs = sm + P.o_z;
        for (int i = threadIdx.x; i < P.nz; i += NLPF_NT) zs[i] = z[(size_t)b * P.nz + i];
        __syncthreads();
        WinPtr wp = window_ptrs(P, b, U, data, xprior, wshift);
        double J, merit, nse;
        eval_window<true>(P, sm, zs, wp, mu, J, merit, nse);
        const int nz = P.nz;
        double* Ho = H_out + (size_t)b * nz * nz;
        for (int t = threadIdx.x; t < nz * nz; t += NLPF_NT) {
            const int r = t / nz, c = t - r * nz;
            Ho[t] = (c <= r) ? sm[P.o_H + pk(r, c)] : sm[P.o_H + pk(c, r)];
        }
        for (int i = threadIdx.x; i < nz; i += NLPF_NT) g_out[(size_t)b * nz + i] = sm[P.o_g + i];
        if (threadIdx.x == 0) {
            cost_out[2 * b] = J;
            cost_out[2 * b + 1] = merit;
        }
    }
}

#if !NLPF_PR_RUNS && !NLPF_BIG
// ------------------------------------------------------------------------------------------
// nlpf_factor_solve: kernel (3) alone
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NLPF_NT)
factor_solve_kernel(DevProb P, int B, const double* __restrict__ H, const double* __restrict__ g, double lam,
                    double* __restrict__ dz_out, int* __restrict__ info_out) {
    extern __shared__ double sm[];
    const int nz = P.nz;
    double* Hp = sm + P.o_H;
    double* red = sm + P.o_red;
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        __syncthreads();
        const double* Hb = H + (size_t)b * nz * nz;
        double dmax = 0.0;
        for (int t = threadIdx.x; t < nz * nz; t += NLPF_NT) {
            const int r = t / nz, c = t - r * nz;
            if (c <= r) Hp[pk(r, c)] = Hb[t];
            if (c == r) dmax = fmax(dmax, Hb[t]);
        }
        dmax = block_max(dmax, red);
        for (int i = threadIdx.x; i < nz; i += NLPF_NT) {
            Hp[pk(i, i)] += lam * (Hp[pk(i, i)] + 1e-14 * dmax);
            Hp[pk(nz, i)] = g[(size_t)b * nz + i];
        }
        const bool ok = ldl_solve(P, nz, Hp, sm + P.o_dz);
        for (int i = threadIdx.x; i < nz; i += NLPF_NT) dz_out[(size_t)b * nz + i] = ok ? -sm[P.o_dz + i] : 0.0;
        if (threadIdx.x == 0) info_out[b] = ok ? 0 : 1;
    }
}

#endif

// Hc + sum_e (phi_e phi_e^T) (x) diag(w_e) of one window whose residuals are all RT_STATE: constant
// over the iterations, written once to the window's slot of the scratch buffer.
__device__ void build_hlin(const DevProb& P, double* sm, const WinPtr& wp) {
    const int tid = threadIdx.x;
    const int K = P.K, nv = P.nv, na = P.na;
    double* Hp = sm + P.o_H;
    const double* Phi = P.const_smem ? sm + P.o_Phi : P.Phi;
    double* wacc = sm + P.o_G;                   // E chunk x na weights
    for (int p = tid; p < P.np; p += NLPF_NT) Hp[p] = P.Hc[p];
    for (int c0 = 0; c0 < P.E; c0 += P.ec) {
        const int ne = min(P.ec, P.E - c0);
        __syncthreads();
        for (int t = tid; t < ne * na; t += NLPF_NT) wacc[t] = 0.0;
        __syncthreads();
        for (int el = tid; el < ne; el += NLPF_NT)
            for (int i = P.estart[c0 + el]; i < P.estart[c0 + el + 1]; ++i)
                wacc[el * na + P.rloc[i * NLPF_NIDX]] += wp.data[P.roff[i] + 1];
        __syncthreads();
        for (int wk = tid; wk < P.npair * na; wk += NLPF_NT) {
            const int pr = wk / na, a = wk - pr * na;
            const int pe = P.ptab[pr];
            const int j = pe & 255, l = pe >> 8;
            double acc = 0.0;
            for (int el = 0; el < ne; ++el) acc = fma(Phi[(c0 + el) * K + j] * Phi[(c0 + el) * K + l], wacc[el * na + a], acc);
            const int r = j * nv + P.act[a], c = l * nv + P.act[a];
            Hp[pk(r, c)] += acc;
        }
    }
    __syncthreads();
    for (int p = tid; p < P.np; p += NLPF_NT) wp.hlin[p] = Hp[p];
    __syncthreads();
}

// ------------------------------------------------------------------------------------------
// nlpf_solve: the whole damped Gauss-Newton / log-barrier iteration of one window per CTA.
// Same decisions, in the same order, as oracle/nls.py:solve (the CPU checker).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NLPF_NT, NLPF_MIN_CTAS)
VN(solve_kernel)(const __grid_constant__ DevProb P, int B, double* __restrict__ z_io, const double* __restrict__ U,
             const double* __restrict__ data, const double* __restrict__ xprior,
             const double* __restrict__ wshift, nlpf_opts opt, double* __restrict__ cost_out,
             int* __restrict__ iters_out, int* __restrict__ status_out, double* __restrict__ hlin) {
    extern __shared__ double sm[];
    const int tid = threadIdx.x;
    const int nz = P.nz, K = P.K, nv = P.nv;
    double* Hp = sm + P.o_H;
    double* zs = sm + P.o_z;
    double* zt = sm + P.o_zt;
    double* g = sm + P.o_g;
    double* ds = sm + P.o_ds;
    double* dz = sm + P.o_dz;
    double* dd = sm + P.o_dd;
    double* red = sm + P.o_red;
    stage_constants(P, sm);
    const bool bounded = P.nb > 0;

    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        __syncthreads();
        for (int i = tid; i < nz; i += NLPF_NT) zs[i] = z_io[(size_t)b * nz + i];
        __syncthreads();
        if (bounded) {   // start strictly inside the bounds (IPOPT bound_push)
            for (int t = tid; t < K * P.nb; t += NLPF_NT) {
                const int k = t / P.nb, bb = t - k * P.nb;
                const int idx = k * nv + P.bidx[bb];
                double v = zs[idx];
                if (P.blo[bb] > -1e300) v = fmax(v, P.blo[bb] + opt.bound_push);
                if (P.bhi[bb] < 1e300) v = fmin(v, P.bhi[bb] - opt.bound_push);
                zs[idx] = v;
            }
            __syncthreads();
        }
        WinPtr wp = window_ptrs(P, b, U, data, xprior, wshift);
        const int hmode = (hlin != nullptr && P.lin_meas) ? 1 : 0;
        if (hmode) {
            wp.hlin = hlin + (size_t)b * P.np;
            build_hlin(P, sm, wp);
        }
        double mu = bounded ? opt.mu0 : 0.0;
        int iters = 0;
        int status = NLPF_ST_MAXITER;
        double lam = opt.lam0;
        double J, merit, nse0;
        bool clip_z = false, clip_t = false;      // the second-order term at z / at the trial point needed its safeguard
        eval_window_cold(P, sm, zs, wp, mu, J, merit, nse0, hmode, &clip_z);
        if (!isfinite(merit)) {
            status = NLPF_ST_NUMERIC;
        } else {
            while (true) {   // barrier stages
                lam = opt.lam0;
                int stage_steps = 0, floor_steps = 0;
                status = NLPF_ST_MAXITER;
                while (iters < opt.max_iter && stage_steps < opt.max_stage_iter) {
                    // ||g||_inf and the damping diagonal
                    double gm = 0.0, dm = 0.0;
                    for (int i = tid; i < nz; i += NLPF_NT) {
                        gm = fmax(gm, fabs(g[i]));
                        ds[i] = Hp[pk(i, i)];
                        dm = fmax(dm, dd[i]);
                    }
                    {
                        double rv[2] = {gm, dm};
                        block_reduce<0, 2>(rv, red);
                        gm = rv[0];
                        dm = rv[1];
                    }
                    if (mu > opt.mu_min && stage_steps >= 1 && gm < 10.0 * mu) {
                        status = NLPF_ST_CONVERGED;
                        break;
                    }
                    bool accepted = false, at_floor = false;
                    bool h_valid = true;      // Hp still holds H(z)
                    double alpha = 1.0, Jt = J, mt = merit, dzmax = 0.0, nset = 0.0;
                    if (clip_z) lam = fmax(lam, opt.lam_jump_indef);     // H is probably indefinite: do not try it undamped
                    const double lam_it0 = lam;
                    while (lam <= 1e18) {
                        bool tried = false;   // a trial point was evaluated (else: pivot not positive / step not finite)
                        if (!h_valid) {       // the factorisation overwrote H: rebuild it at z
                            double Jd, md, nd;
                            eval_window_cold(P, sm, zs, wp, mu, Jd, md, nd, hmode, nullptr);
                            h_valid = true;
                        }
                        for (int i = tid; i < nz; i += NLPF_NT) {
                            Hp[pk(i, i)] = ds[i] + lam * (dd[i] + 1e-14 * dm);
                            Hp[pk(nz, i)] = g[i];
                        }
                        h_valid = false;
                        const bool ok = ldl_solve(P, nz, Hp, dz);
                        if (ok) {
                            // dz holds H^-1 g: the step is -dz.  fraction to the boundary:
                            double a = 1.0, dmx = 0.0, finite = 0.0;
                            for (int i = tid; i < nz; i += NLPF_NT) {
                                dmx = fmax(dmx, fabs(dz[i]));
                                if (!isfinite(dz[i])) finite = 1.0;
                            }
                            if (bounded) {
                                for (int t = tid; t < K * P.nb; t += NLPF_NT) {
                                    const int k = t / P.nb, bb = t - k * P.nb;
                                    const int idx = k * nv + P.bidx[bb];
                                    const double st = -dz[idx];
                                    if (st < 0.0 && P.blo[bb] > -1e300) a = fmin(a, opt.tau_ftb * (zs[idx] - P.blo[bb]) / (-st));
                                    if (st > 0.0 && P.bhi[bb] < 1e300) a = fmin(a, opt.tau_ftb * (P.bhi[bb] - zs[idx]) / st);
                                }
                            }
                            double pred = 0.0;                       // predicted decrease scale g^T H^-1 g
                            for (int i = tid; i < nz; i += NLPF_NT) pred = fma(g[i], dz[i], pred);
                            {
                                double rv[4] = {pred, -a, dmx, finite};      // min(a) = -max(-a)
                                block_reduce<1, 3>(rv, red);
                                pred = rv[0];
                                a = -rv[1];
                                dmx = rv[2];
                                finite = rv[3];
                            }
                            pred *= a;
                            if (finite == 0.0) {
                                double ztm = 0.0;
                                for (int i = tid; i < nz; i += NLPF_NT) {
                                    const double v = fma(-a, dz[i], zs[i]);
                                    zt[i] = v;
                                    ztm = fmax(ztm, fabs(v));
                                }
                                ztm = block_max(ztm, red);          // (ends with a barrier: zt is visible)
                                // a first-try step below the tolerance ends the stage without an evaluation at the new
                                // point: the merit cannot resolve such a step, and H, g there would not be used.  J stays
                                // the value at the previous iterate (it differs by O(step^2)).
                                if (lam <= lam_it0 && a * dmx < opt.tol_step * (1.0 + ztm)) {
                                    accepted = true;
                                    alpha = a;
                                    dzmax = dmx;
                                    nset = nse0;                    // (Jt, mt still hold J, merit)
                                    break;
                                }
                                // FULL evaluation at the trial point: an accepted step (the usual case) needs H and g
                                // there anyway; a rejected one rebuilds H at z below, as it had to before
                                eval_window<true>(P, sm, zt, wp, mu, Jt, mt, nset, hmode, &clip_t);
                                tried = true;
                                if (isfinite(mt) && mt < merit) {
                                    accepted = true;
                                    alpha = a;
                                    dzmax = dmx;
                                    break;
                                }
                                // rounding floor of the merit: no measurable decrease is predicted
                                // (the Gauss-Newton step is still accurate because g is): take it, count it
                                const double nb = nse0 + nset;
                                if (isfinite(mt) && (mt - merit) <= nb && pred <= 4.0 * nb) {
                                    at_floor = true;
                                    accepted = true;
                                    alpha = a;
                                    dzmax = dmx;
                                    break;
                                }
                            }
                        }
                        lam = fmax(10.0 * lam, tried ? opt.lam_jump : opt.lam_jump_indef);
                    }
                    if (!accepted) {
                        status = NLPF_ST_STALLED;
                        if (!h_valid) {   // leave H, g consistent with z for the stationarity test and a following stage
                            double Jd, md, nd;
                            eval_window_cold(P, sm, zs, wp, mu, Jd, md, nd, hmode, nullptr);
                        }
                        // "stalled" is a success only at a stationary point: the diagonally scaled gradient step must
                        // be below NLPF_STALL_GTOL (NaN fails the test); otherwise the window is reported as failed
                        double sg = 0.0, zmx = 0.0, nang = 0.0;
                        for (int i = tid; i < nz; i += NLPF_NT) {
                            const double q = fabs(g[i]) / (dd[i] + 1e-14 * dm + 1e-300);
                            if (!(q == q)) nang = 1.0;
                            sg = fmax(sg, q);
                            zmx = fmax(zmx, fabs(zs[i]));
                        }
                        {
                            double rv[3] = {sg, zmx, nang};
                            block_reduce<0, 3>(rv, red);
                            sg = rv[0];
                            zmx = rv[1];
                            nang = rv[2];
                        }
                        if (nang > 0.0 || !(sg <= NLPF_STALL_GTOL * (1.0 + zmx))) status = NLPF_ST_NOSTEP;
                        break;
                    }
                    const double step = alpha * dzmax;
                    double zm = 0.0;
                    for (int i = tid; i < nz; i += NLPF_NT) {
                        zs[i] = zt[i];
                        zm = fmax(zm, fabs(zt[i]));
                    }
                    zm = block_max(zm, red);
                    lam = fmax(lam / ((lam <= lam_it0) ? opt.lam_dec_first : 10.0), 1e-12);
                    ++iters;
                    ++stage_steps;
                    J = Jt;               // H, g, dd already belong to the accepted point
                    merit = mt;
                    nse0 = nset;
                    clip_z = clip_t;
                    if (step < opt.tol_step * (1.0 + zm)) {
                        status = NLPF_ST_CONVERGED;
                        break;
                    }
                    if (at_floor && ++floor_steps >= 5) {
                        status = NLPF_ST_CONVERGED;
                        break;
                    }
                }
                if (!bounded || mu <= opt.mu_min || iters >= opt.max_iter) break;
                mu = fmax(opt.mu_min, fmin(0.2 * mu, mu * sqrt(mu)));
                eval_window_cold(P, sm, zs, wp, mu, J, merit, nse0, hmode, &clip_z);
            }
        }
        for (int i = tid; i < nz; i += NLPF_NT) z_io[(size_t)b * nz + i] = zs[i];
        if (tid == 0) {
            cost_out[b] = J;
            iters_out[b] = iters;
            status_out[b] = status;
        }
    }
}

#if !NLPF_PR_RUNS && !NLPF_BIG
// ------------------------------------------------------------------------------------------
// nlpf_sample: x(t) = Phi(t) X, nlp.py:99-119
// ------------------------------------------------------------------------------------------
__global__ void sample_kernel(DevProb P, int B, const double* __restrict__ z, int nt,
                              const double* __restrict__ PhiT, double* __restrict__ out) {
    const size_t total = (size_t)B * nt * P.n;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int i = t % P.n;
        const int it = (t / P.n) % nt;
        const int b = t / ((size_t)P.n * nt);
        const double* zb = z + (size_t)b * P.nz;
        const int col = P.cmap[i];
        double s = 0.0;
        for (int j = 0; j < P.K; ++j) s = fma(PhiT[it * P.K + j], zb[j * P.nv + col], s);
        out[t] = s;
    }
}

// ------------------------------------------------------------------------------------------
// nlpf_defects: W_k = (2/T) sum_j D_kj X_j - f(X_k, U_k) (+ shift), nlp.py:222-235 -- the collocation equalities of the
// optimal-control class (nlp.py:150-158), whose method-of-multipliers driver needs them for B problems at once
// ------------------------------------------------------------------------------------------
__global__ void defects_kernel(DevProb P, int B, const double* __restrict__ z, const double* __restrict__ U,
                               double* __restrict__ W_out) {
    const int K = P.K, n = P.n, m = P.m, nv = P.nv;
    const size_t total = (size_t)B * K;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int k = t % K;
        const size_t b = t / K;
        const double* zb = z + b * P.nz;
        double x[NLPF_N_MAX], u[NLPF_M_MAX], f[NLPF_N_MAX];
        for (int i = 0; i < n; ++i) x[i] = zb[k * nv + P.cmap[i]];
        for (int i = 0; i < m; ++i) u[i] = (P.ucol[i] >= 0) ? zb[k * nv + P.ucol[i]] : (U ? U[(b * K + k) * m + i] : 0.0);
        dyn_eval(P, x, u, f, false, nullptr);
        for (int i = 0; i < n; ++i) {
            double w = -f[i];
            for (int j = 0; j < K; ++j) w = fma(P.Dt[k * K + j], zb[j * nv + P.cmap[i]], w);
            W_out[(b * K + k) * n + i] = w;
        }
    }
}

#endif

}  // namespace
using namespace VN(nlpf_ns);

#ifndef __CUDACC_RTC__
// ------------------------------------------------------------------------------------------
// launchers (called from capi.cu)
// ------------------------------------------------------------------------------------------
static int grid_for(int B) { return B < 1 ? 1 : B; }

extern "C" int VN(nlpf_launch_assemble)(nlpf_problem* p, int B, const double* z, const double* U, const double* data,
                                    const double* xprior, const double* wshift, double mu, double* H_out,
                                    double* g_out, double* cost_out, cudaStream_t st) {
    VN(assemble_kernel)<<<grid_for(B), NLPF_NT, p->smem_bytes, st>>>(p->d, B, z, U, data, xprior, wshift, mu, H_out, g_out, cost_out);
    return nlpf_check_cuda(cudaGetLastError(), "assemble_kernel launch");
}

#if !NLPF_PR_RUNS && !NLPF_BIG
extern "C" int nlpf_launch_factor_solve(nlpf_problem* p, int B, const double* H, const double* g, double lam,
                                        double* dz_out, int* info_out, cudaStream_t st) {
    factor_solve_kernel<<<grid_for(B), NLPF_NT, p->smem_bytes, st>>>(p->d, B, H, g, lam, dz_out, info_out);
    return nlpf_check_cuda(cudaGetLastError(), "factor_solve_kernel launch");
}

#endif

extern "C" int VN(nlpf_launch_solve)(nlpf_problem* p, int B, double* z, const double* U, const double* data,
                                 const double* xprior, const double* wshift, const nlpf_opts* o, double* cost_out,
                                 int* iters_out, int* status_out, double* hlin, cudaStream_t st) {
    VN(solve_kernel)<<<grid_for(B), NLPF_NT, p->smem_bytes, st>>>(p->d, B, z, U, data, xprior, wshift, *o, cost_out, iters_out, status_out, hlin);
    return nlpf_check_cuda(cudaGetLastError(), "solve_kernel launch");
}

#if !NLPF_PR_RUNS && !NLPF_BIG
extern "C" int nlpf_launch_sample(nlpf_problem* p, int B, const double* z, int nt, const double* PhiT, double* out,
                                  cudaStream_t st) {
    const size_t total = (size_t)B * nt * p->d.n;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    sample_kernel<<<blocks, 256, 0, st>>>(p->d, B, z, nt, PhiT, out);
    return nlpf_check_cuda(cudaGetLastError(), "sample_kernel launch");
}

#endif

#if !NLPF_PR_RUNS && !NLPF_BIG
extern "C" int nlpf_launch_defects(nlpf_problem* p, int B, const double* z, const double* U, double* W_out, cudaStream_t st) {
    const size_t total = (size_t)B * p->d.K;
    int blocks = (int)((total + 127) / 128);
    if (blocks > 148 * 16) blocks = 148 * 16;
    defects_kernel<<<blocks, 128, 0, st>>>(p->d, B, z, U, W_out);
    return nlpf_check_cuda(cudaGetLastError(), "defects_kernel launch");
}
#endif

extern "C" int VN(nlpf_set_smem_attrs)(size_t bytes) {
    cudaError_t e;
    e = cudaFuncSetAttribute(VN(assemble_kernel), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return nlpf_check_cuda(e, "smem attr assemble");
#if !NLPF_PR_RUNS && !NLPF_BIG
    e = cudaFuncSetAttribute(factor_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return nlpf_check_cuda(e, "smem attr factor");
#endif
    e = cudaFuncSetAttribute(VN(solve_kernel), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return nlpf_check_cuda(e, "smem attr solve");
    e = cudaFuncSetAttribute(VN(solve_kernel), cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return nlpf_check_cuda(e, "smem carveout solve");
    return 0;
}
#endif  // !__CUDACC_RTC__
