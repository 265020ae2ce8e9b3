// blocktri_kernels.cu -- SURVEY.md section 8 row A9: batched solve of SPD block-tridiagonal systems
//
//        [ D_0  L_1^T                 ] [x_0]   [r_0]
//        [ L_1  D_1   L_2^T           ] [x_1] = [r_1]        D_i, L_i : bs x bs,  nblk block rows
//        [      L_2   D_2   ...       ] [...]   [...]
//
// This structure has NO counterpart in the reference: its single-interval Chebyshev collocation
// gives a dense normal matrix (nlp/collocation.py:42-64).  It appears when a long
// horizon is cut into segments that share end nodes (multi-segment collocation, "horizon 200" of
// BASELINE.json), so the kernels here are an extension validated against the dense kernel (3).
//
//   variant 0  block Cholesky = Riccati recursion: S_0 = D_0, S_i = D_i - L_i S_{i-1}^-1 L_i^T,
//              sequential in the segments, one CTA per system, all 128 threads on each block op.
//   variant 1  block cyclic (odd-even) reduction: ceil(log2 nblk) levels; at stride s the rows with
//              odd index/s are eliminated into their even neighbours, one warp per kept row:
//                D_i -= L_i D_j^-1 L_i^T + L_j'^T D_j'^-1 L_j' ,  L_i <- -L_i D_j^-1 L_j ,  j = i-s, j' = i+s
//              then back-substitution level by level.
// Both leave the inputs untouched (they work in a per-call workspace) and report a non-positive
// pivot through info (1) instead of producing NaNs.
#include "common.cuh"
#include <cooperative_groups.h>
#include <math.h>

namespace cg = cooperative_groups;

#define BT_BS_MAX 16
#define BT_THREADS 128

// in-place Cholesky of the bs x bs lower triangle of A (row-major, leading dimension bs)
template <typename G>
__device__ bool grp_chol(const G& g, int bs, double* A) {
    bool ok = true;
    for (int j = 0; j < bs; ++j) {
        g.sync();
        const double d = A[j * bs + j];
        if (!(d > 0.0) || !isfinite(d)) { ok = false; break; }      // uniform within the group
        const double sd = sqrt(d);
        g.sync();
        for (int i = j + g.thread_rank(); i < bs; i += g.size()) A[i * bs + j] = (i == j) ? sd : A[i * bs + j] / sd;
        g.sync();
        const int rem = bs - (j + 1);
        for (int e = g.thread_rank(); e < rem * rem; e += g.size()) {
            const int i = j + 1 + e / rem, k = j + 1 + e % rem;
            if (k <= i) A[i * bs + k] -= A[i * bs + j] * A[k * bs + j];
        }
    }
    g.sync();
    return ok;
}

// X <- (C C^T)^-1 X for X (bs x ncol, leading dimension ldx): one thread per column
template <typename G>
__device__ void grp_spd_solve(const G& g, int bs, const double* C, double* X, int ncol, int ldx) {
    g.sync();
    for (int c = g.thread_rank(); c < ncol; c += g.size()) {
        for (int i = 0; i < bs; ++i) {
            double s = X[i * ldx + c];
            for (int k = 0; k < i; ++k) s -= C[i * bs + k] * X[k * ldx + c];
            X[i * ldx + c] = s / C[i * bs + i];
        }
        for (int i = bs - 1; i >= 0; --i) {
            double s = X[i * ldx + c];
            for (int k = i + 1; k < bs; ++k) s -= C[k * bs + i] * X[k * ldx + c];
            X[i * ldx + c] = s / C[i * bs + i];
        }
    }
    g.sync();
}

// workspace of one system: Dw, Lw, Rw (nblk x bs^2 each), rw, xw (nblk x bs each)
struct BtWork {
    double *D, *L, *R, *r, *x;
};
__device__ __forceinline__ BtWork bt_work(double* base, int nblk, int bs) {
    BtWork w;
    const size_t b2 = (size_t)nblk * bs * bs;
    w.D = base;
    w.L = base + b2;
    w.R = base + 2 * b2;
    w.r = base + 3 * b2;
    w.x = w.r + (size_t)nblk * bs;
    return w;
}

// ------------------------------------------------------------------------------------------
// variant 0: Riccati recursion (block Cholesky), one CTA per system
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BT_THREADS)
blocktri_riccati_kernel(int B, int nblk, int bs, const double* __restrict__ Dg, const double* __restrict__ Lo,
                        const double* __restrict__ rhs, double* __restrict__ x_out, int* __restrict__ info,
                        double* __restrict__ work, size_t work_stride) {
    __shared__ double C[BT_BS_MAX * BT_BS_MAX];                      // chol(S_{i-1})
    __shared__ double T[BT_BS_MAX * (BT_BS_MAX + 1)];                // S^-1 [L_i^T | y_{i-1}]
    __shared__ double Li[BT_BS_MAX * BT_BS_MAX];
    cg::thread_block g = cg::this_thread_block();
    const int tid = threadIdx.x, b2 = bs * bs, ld = bs + 1;
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        BtWork w = bt_work(work + (size_t)b * work_stride, nblk, bs);
        const double* Db = Dg + (size_t)b * nblk * b2;
        const double* Lb = Lo + (size_t)b * nblk * b2;
        const double* rb = rhs + (size_t)b * nblk * bs;
        bool ok = true;
        // forward: factor S_i, y_i = r_i - L_i S_{i-1}^-1 y_{i-1}; chol(S_i) kept in w.D, y in w.r
        for (int i = 0; i < nblk && ok; ++i) {
            g.sync();
            double* Si = w.D + (size_t)i * b2;
            for (int e = tid; e < b2; e += BT_THREADS) Si[e] = Db[(size_t)i * b2 + e];
            for (int e = tid; e < bs; e += BT_THREADS) w.r[(size_t)i * bs + e] = rb[(size_t)i * bs + e];
            if (i > 0) {
                for (int e = tid; e < b2; e += BT_THREADS) {
                    Li[e] = Lb[(size_t)i * b2 + e];
                    const int a = e / bs, c = e % bs;
                    T[c * ld + a] = Lb[(size_t)i * b2 + e];          // T[:, :bs] = L_i^T
                }
                for (int e = tid; e < bs; e += BT_THREADS) T[e * ld + bs] = w.r[(size_t)(i - 1) * bs + e];
                grp_spd_solve(g, bs, C, T, bs + 1, ld);                // T = S_{i-1}^-1 [L_i^T | y_{i-1}]
                for (int e = tid; e < b2; e += BT_THREADS) {
                    const int a = e / bs, c = e % bs;
                    double s = 0.0;
                    for (int k = 0; k < bs; ++k) s = fma(Li[a * bs + k], T[k * ld + c], s);
                    Si[e] -= s;
                }
                for (int a = tid; a < bs; a += BT_THREADS) {
                    double s = 0.0;
                    for (int k = 0; k < bs; ++k) s = fma(Li[a * bs + k], T[k * ld + bs], s);
                    w.r[(size_t)i * bs + a] -= s;
                }
            }
            g.sync();
            for (int e = tid; e < b2; e += BT_THREADS) C[e] = Si[e];
            ok = grp_chol(g, bs, C);
            for (int e = tid; e < b2; e += BT_THREADS) Si[e] = C[e];
        }
        g.sync();
        // backward: x_i = S_i^-1 (y_i - L_{i+1}^T x_{i+1})
        if (ok) {
            for (int i = nblk - 1; i >= 0; --i) {
                g.sync();
                for (int e = tid; e < b2; e += BT_THREADS) C[e] = w.D[(size_t)i * b2 + e];
                for (int a = tid; a < bs; a += BT_THREADS) {
                    double s = w.r[(size_t)i * bs + a];
                    if (i + 1 < nblk)
                        for (int k = 0; k < bs; ++k) s = fma(-Lb[(size_t)(i + 1) * b2 + k * bs + a], w.x[(size_t)(i + 1) * bs + k], s);
                    T[a * ld] = s;
                }
                grp_spd_solve(g, bs, C, T, 1, ld);
                for (int a = tid; a < bs; a += BT_THREADS) w.x[(size_t)i * bs + a] = T[a * ld];
            }
        }
        g.sync();
        for (int e = tid; e < nblk * bs; e += BT_THREADS) x_out[(size_t)b * nblk * bs + e] = ok ? w.x[e] : 0.0;
        if (tid == 0) info[b] = ok ? 0 : 1;
    }
}

// ------------------------------------------------------------------------------------------
// variant 1: block cyclic reduction, one CTA per system, one warp per kept block row and level
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BT_THREADS)
blocktri_cr_kernel(int B, int nblk, int bs, const double* __restrict__ Dg, const double* __restrict__ Lo,
                   const double* __restrict__ rhs, double* __restrict__ x_out, int* __restrict__ info,
                   double* __restrict__ work, size_t work_stride) {
    __shared__ double Tsm[BT_THREADS / 32][BT_BS_MAX * (2 * BT_BS_MAX + 1)];
    __shared__ double Nsm[BT_THREADS / 32][BT_BS_MAX * BT_BS_MAX];
    __shared__ int bad;
    cg::thread_block blk = cg::this_thread_block();
    cg::thread_block_tile<32> wp = cg::tiled_partition<32>(blk);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = BT_THREADS / 32;
    const int b2 = bs * bs, ld = 2 * bs + 1;
    double* T = Tsm[warp];
    double* Nw = Nsm[warp];
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        BtWork w = bt_work(work + (size_t)b * work_stride, nblk, bs);
        if (tid == 0) bad = 0;
        for (int e = tid; e < nblk * b2; e += BT_THREADS) {
            w.D[e] = Dg[(size_t)b * nblk * b2 + e];
            w.L[e] = Lo[(size_t)b * nblk * b2 + e];
        }
        for (int e = tid; e < nblk * bs; e += BT_THREADS) w.r[e] = rhs[(size_t)b * nblk * bs + e];
        blk.sync();
        int s = 1;
        for (; s < nblk; s <<= 1) {
            // phase A: Cholesky of the diagonal blocks eliminated at this level (index/s odd)
            for (int j = s + warp * 2 * s; j < nblk; j += nwarp * 2 * s)
                if (!grp_chol(wp, bs, w.D + (size_t)j * b2)) bad = 1;
            blk.sync();
            if (bad) break;
            // phase B: kept rows (index/s even) absorb their eliminated neighbours
            for (int i = warp * 2 * s; i < nblk; i += nwarp * 2 * s) {
                double* Di = w.D + (size_t)i * b2;
                double* Li = w.L + (size_t)i * b2;       // coupling i <- i-s at this level
                const int jl = i - s, jr = i + s;
                for (int e = lane; e < b2; e += 32) Nw[e] = 0.0;
                if (jl >= 0) {
                    const double* Cj = w.D + (size_t)jl * b2;
                    const double* Lj = w.L + (size_t)jl * b2;      // coupling jl <- jl-s
                    for (int e = lane; e < b2; e += 32) {
                        const int a = e / bs, c = e % bs;
                        T[c * ld + a] = Li[e];                                   // L_i^T
                        T[a * ld + bs + c] = (jl - s >= 0) ? Lj[e] : 0.0;          // L_j
                    }
                    for (int e = lane; e < bs; e += 32) T[e * ld + 2 * bs] = w.r[(size_t)jl * bs + e];
                    grp_spd_solve(wp, bs, Cj, T, 2 * bs + 1, ld);                 // D_j^-1 [L_i^T | L_j | r_j]
                    for (int e = lane; e < b2; e += 32) {
                        const int a = e / bs, c = e % bs;
                        double s1 = 0.0, s2 = 0.0;
                        for (int k = 0; k < bs; ++k) {
                            s1 = fma(Li[a * bs + k], T[k * ld + c], s1);
                            s2 = fma(Li[a * bs + k], T[k * ld + bs + c], s2);
                        }
                        Di[e] -= s1;
                        Nw[e] = -s2;                                             // new coupling i <- i-2s
                    }
                    for (int a = lane; a < bs; a += 32) {
                        double s3 = 0.0;
                        for (int k = 0; k < bs; ++k) s3 = fma(Li[a * bs + k], T[k * ld + 2 * bs], s3);
                        w.r[(size_t)i * bs + a] -= s3;
                    }
                    wp.sync();
                    // the eliminated row needs the coupling (i <- jl) of THIS level for its back-substitution
                    for (int e = lane; e < b2; e += 32) w.R[(size_t)jl * b2 + e] = Li[e];
                }
                if (jr < nblk) {
                    const double* Cj = w.D + (size_t)jr * b2;
                    const double* Mr = w.L + (size_t)jr * b2;      // coupling jr <- i :  A[jr,i] = M, A[i,jr] = M^T
                    for (int e = lane; e < b2; e += 32) T[(e / bs) * ld + (e % bs)] = Mr[e];
                    for (int e = lane; e < bs; e += 32) T[e * ld + bs] = w.r[(size_t)jr * bs + e];
                    grp_spd_solve(wp, bs, Cj, T, bs + 1, ld);                     // D_jr^-1 [M | r_jr]
                    for (int e = lane; e < b2; e += 32) {
                        const int a = e / bs, c = e % bs;
                        double s1 = 0.0;
                        for (int k = 0; k < bs; ++k) s1 = fma(Mr[k * bs + a], T[k * ld + c], s1);
                        Di[e] -= s1;
                    }
                    for (int a = lane; a < bs; a += 32) {
                        double s3 = 0.0;
                        for (int k = 0; k < bs; ++k) s3 = fma(Mr[k * bs + a], T[k * ld + bs], s3);
                        w.r[(size_t)i * bs + a] -= s3;
                    }
                }
                wp.sync();
                for (int e = lane; e < b2; e += 32) Li[e] = Nw[e];
                wp.sync();
            }
            blk.sync();
        }
        // root: only row 0 is left
        if (!bad && warp == 0) {
            if (!grp_chol(wp, bs, w.D)) bad = 1;
            else {
                for (int e = lane; e < bs; e += 32) T[e * ld] = w.r[e];
                grp_spd_solve(wp, bs, w.D, T, 1, ld);
                for (int e = lane; e < bs; e += 32) w.x[e] = T[e * ld];
            }
        }
        blk.sync();
        // back-substitution, coarsest level first: x_j = D_j^-1 (r_j - L_j x_{j-s} - R_j^T x_{j+s})
        if (!bad) {
            for (s >>= 1; s >= 1; s >>= 1) {
                for (int j = s + warp * 2 * s; j < nblk; j += nwarp * 2 * s) {
                    const double* Lj = w.L + (size_t)j * b2;
                    const double* Rj = w.R + (size_t)j * b2;       // coupling (j+s <- j) at the level j was eliminated
                    for (int a = lane; a < bs; a += 32) {
                        double v = w.r[(size_t)j * bs + a];
                        for (int k = 0; k < bs; ++k) v = fma(-Lj[a * bs + k], w.x[(size_t)(j - s) * bs + k], v);
                        if (j + s < nblk)
                            for (int k = 0; k < bs; ++k) v = fma(-Rj[k * bs + a], w.x[(size_t)(j + s) * bs + k], v);
                        T[a * ld] = v;
                    }
                    grp_spd_solve(wp, bs, w.D + (size_t)j * b2, T, 1, ld);
                    for (int a = lane; a < bs; a += 32) w.x[(size_t)j * bs + a] = T[a * ld];
                }
                blk.sync();
            }
        }
        blk.sync();
        for (int e = tid; e < nblk * bs; e += BT_THREADS) x_out[(size_t)b * nblk * bs + e] = bad ? 0.0 : w.x[e];
        if (tid == 0) info[b] = bad ? 1 : 0;
        blk.sync();
    }
}

extern "C" int nlpf_blocktri_solve(int B, int nblk, int bs, const double* Dg, const double* Lo, const double* rhs,
                                   double* x_out, int* info_out, int variant, void* stream) {
    if (B < 1 || nblk < 1 || bs < 1 || bs > BT_BS_MAX) { nlpf_set_error("bad block-tridiagonal sizes B=%d nblk=%d bs=%d (bs <= %d)", B, nblk, bs, BT_BS_MAX); return NLPF_E_ARG; }
    if (!Dg || !Lo || !rhs || !x_out || !info_out) { nlpf_set_error("null block-tridiagonal buffer"); return NLPF_E_ARG; }
    if (variant != 0 && variant != 1) { nlpf_set_error("variant must be 0 (Riccati) or 1 (cyclic reduction)"); return NLPF_E_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t stride = (size_t)nblk * (3 * bs * bs + 2 * bs);
    double* work = nullptr;
    if (nlpf_check_cuda(cudaMallocAsync((void**)&work, sizeof(double) * stride * B, st), "cudaMallocAsync(blocktri workspace)")) return NLPF_E_CUDA;
    const int grid = B < 148 * 8 ? B : 148 * 8;
    if (variant == 0) blocktri_riccati_kernel<<<grid, BT_THREADS, 0, st>>>(B, nblk, bs, Dg, Lo, rhs, x_out, info_out, work, stride);
    else blocktri_cr_kernel<<<grid, BT_THREADS, 0, st>>>(B, nblk, bs, Dg, Lo, rhs, x_out, info_out, work, stride);
    int rc = nlpf_check_cuda(cudaGetLastError(), "blocktri kernel launch");
    cudaFreeAsync(work, st);
    return rc;
}
