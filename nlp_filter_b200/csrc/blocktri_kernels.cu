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

#define BT_BS_MAX 48          // Riccati: three bs x bs blocks in (dynamic) shared memory = 55 KB at bs = 48
#define BT_BS_MAX_CR 32       // cyclic reduction: per warp bs x (2 bs + 1) + bs x bs; 4 warps = 98 KB at bs = 32
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
    extern __shared__ double bt_sm[];                                // sized by the launcher from bs (multi-segment blocks reach bs = 45)
    cg::thread_block g = cg::this_thread_block();
    const int tid = threadIdx.x, b2 = bs * bs, ld = bs + 1;
    double* C = bt_sm;                                               // chol(S_{i-1})          bs x bs
    double* T = C + b2;                                              // S^-1 [L_i^T | y_{i-1}]  bs x (bs + 1)
    double* Li = T + bs * ld;                                        //                         bs x bs
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
    extern __shared__ double bt_sm[];
    __shared__ int bad;
    cg::thread_block blk = cg::this_thread_block();
    cg::thread_block_tile<32> wp = cg::tiled_partition<32>(blk);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = BT_THREADS / 32;
    const int b2 = bs * bs, ld = 2 * bs + 1;
    double* T = bt_sm + (size_t)warp * (bs * ld + b2);               // per warp: bs x (2 bs + 1), then bs x bs
    double* Nw = T + bs * ld;
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
    const int bs_max = (variant == 1) ? BT_BS_MAX_CR : BT_BS_MAX;
    if (B < 1 || nblk < 1 || bs < 1 || bs > bs_max) { nlpf_set_error("bad block-tridiagonal sizes B=%d nblk=%d bs=%d (bs <= %d for this variant)", B, nblk, bs, bs_max); return NLPF_E_ARG; }
    if (!Dg || !Lo || !rhs || !x_out || !info_out) { nlpf_set_error("null block-tridiagonal buffer"); return NLPF_E_ARG; }
    if (variant != 0 && variant != 1) { nlpf_set_error("variant must be 0 (Riccati) or 1 (cyclic reduction)"); return NLPF_E_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t stride = (size_t)nblk * (3 * bs * bs + 2 * bs);
    double* work = nullptr;
    if (nlpf_check_cuda(cudaMallocAsync((void**)&work, sizeof(double) * stride * B, st), "cudaMallocAsync(blocktri workspace)")) return NLPF_E_CUDA;
    const int grid = B < 148 * 8 ? B : 148 * 8;
    const size_t sm_ric = sizeof(double) * (size_t)(2 * bs * bs + bs * (bs + 1));
    const size_t sm_cr = sizeof(double) * (size_t)(BT_THREADS / 32) * (bs * (2 * bs + 1) + bs * bs);
    static bool attr = false;
    if (!attr) {
        // keep the per-call workspace in the stream-ordered pool between calls: with the default release threshold (0)
        // every synchronisation hands the memory back to the driver and the next call maps it again -- milliseconds
        // per call, ten times the kernel itself on the 200-epoch workload
        int devid = 0;
        cudaMemPool_t pool;
        if (cudaGetDevice(&devid) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, devid) == cudaSuccess) {
            unsigned long long keep = ~0ULL;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaFuncSetAttribute(blocktri_riccati_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)(sizeof(double) * (2 * BT_BS_MAX * BT_BS_MAX + BT_BS_MAX * (BT_BS_MAX + 1))));
        cudaFuncSetAttribute(blocktri_cr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)(sizeof(double) * (BT_THREADS / 32) * (BT_BS_MAX_CR * (2 * BT_BS_MAX_CR + 1) + BT_BS_MAX_CR * BT_BS_MAX_CR)));
        attr = true;
    }
    if (variant == 0) blocktri_riccati_kernel<<<grid, BT_THREADS, sm_ric, st>>>(B, nblk, bs, Dg, Lo, rhs, x_out, info_out, work, stride);
    else blocktri_cr_kernel<<<grid, BT_THREADS, sm_cr, st>>>(B, nblk, bs, Dg, Lo, rhs, x_out, info_out, work, stride);
    int rc = nlpf_check_cuda(cudaGetLastError(), "blocktri kernel launch");
    cudaFreeAsync(work, st);
    return rc;
}

// ------------------------------------------------------------------------------------------
// Multi-segment collocation (row A9): a long horizon [0, S T_s] is cut into S segments, each the
// reference's single-interval Chebyshev collocation (nlp/collocation.py) on its own N_s + 1 nodes;
// neighbouring segments SHARE their end node (continuity by construction).  Unknowns: S N_s + 1
// nodes of n states.  Every cost term of nlp/nlp.py:202-277 lives inside one segment, so
//     J(z) = sum_s J_seg(z_seg(s)) + arrival cost on node 0          z_seg(s) = nodes s N_s .. s N_s + N_s
// and H, g are the overlap-add of the per-segment dense H_seg, g_seg -- which nlpf_assemble
// (kernels 1+2) computes for B * S "sub-windows" of the ordinary single-segment problem class.
// Ordering the unknowns segment by segment (block s = the first N_s nodes of segment s, block S = the
// last node alone) makes H block-tridiagonal with bs = N_s n: block s couples to block s + 1 only
// through the first node of block s + 1.
//
//   ms_gather_kernel   z[B][(S N_s + 1) n]  ->  zsub[B S][(N_s + 1) n]
//   ms_blocks_kernel   H_seg, g_seg, J_seg  ->  Dg, Lo [B][S+1][bs][bs], g [B][(S+1) bs], J[B], with the
//                      Levenberg-Marquardt damping lam_b * diag(H) already added to Dg (so a retry with a larger
//                      lam only re-reads H_seg) and the padding rows of the last block set to identity.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ms_gather_kernel(int B, int S, int kn, int step, int nzt, const double* __restrict__ z,
                                                        double* __restrict__ zsub) {
    const size_t total = (size_t)B * S * kn;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int i = t % kn;
        const int s = (t / kn) % S;
        const size_t b = t / ((size_t)kn * S);
        zsub[t] = z[b * nzt + (size_t)s * step + i];
    }
}

// one CTA per window; Hs[b][s] is the (kn x kn) row-major segment matrix, kn = bs + n
__global__ void __launch_bounds__(128) ms_blocks_kernel(int B, int S, int n, int bs, const double* __restrict__ Hs,
                                                        const double* __restrict__ gs, const double* __restrict__ cs,
                                                        const double* __restrict__ pinv, const double* __restrict__ xprior,
                                                        const double* __restrict__ z, const double* __restrict__ lam,
                                                        double* __restrict__ Dg, double* __restrict__ Lo, double* __restrict__ g,
                                                        double* __restrict__ J) {
    const int kn = bs + n, nblk = S + 1, b2 = bs * bs;
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        const double lm = lam ? lam[b] : 0.0;
        double* Db = Dg + (size_t)b * nblk * b2;
        double* Lb = Lo + (size_t)b * nblk * b2;
        double* gb = g + (size_t)b * nblk * bs;
        const int nzt = S * bs + n;
        for (int t = threadIdx.x; t < nblk * b2; t += blockDim.x) {
            const int blk = t / b2, e = t - blk * b2, r = e / bs, c = e - r * bs;
            double d = 0.0, l = 0.0;
            if (blk < S) d = Hs[((size_t)b * S + blk) * kn * kn + (size_t)r * kn + c];            // segment blk: its first N_s nodes
            if (blk > 0 && r < n && c < n)                                                        // last node of segment blk-1 = first node here
                d += Hs[((size_t)b * S + blk - 1) * kn * kn + (size_t)(bs + r) * kn + bs + c];
            if (blk == 0 && pinv && r == c && r < n) d += pinv[r];                                // arrival cost (nlp.py:279-286), half-gradient convention
            if (blk == S && (r >= n || c >= n)) d = (r == c) ? 1.0 : 0.0;                         // padding of the last block
            // damping lam |H_ii|: H carries the exact second-order dynamics term, so a diagonal entry can be negative far
            // from the solution; its magnitude keeps the damped matrix positive definite for large enough lam
            if (r == c && !(blk == S && r >= n)) d += lm * fabs(d);
            if (blk > 0 && r < n) l = Hs[((size_t)b * S + blk - 1) * kn * kn + (size_t)(bs + r) * kn + c];   // coupling blk <- blk-1
            Db[t] = d;
            Lb[t] = l;
        }
        for (int t = threadIdx.x; t < nblk * bs; t += blockDim.x) {
            const int blk = t / bs, r = t - blk * bs;
            double v = 0.0;
            if (blk < S) v = gs[((size_t)b * S + blk) * kn + r];
            if (blk > 0 && r < n) v += gs[((size_t)b * S + blk - 1) * kn + bs + r];
            if (blk == 0 && pinv && r < n) v += pinv[r] * (z[(size_t)b * nzt + r] - xprior[(size_t)b * n + r]);
            if (blk == S && r >= n) v = 0.0;
            gb[t] = v;
        }
        if (threadIdx.x < 32) {
            double s = 0.0;
            for (int k = threadIdx.x; k < S; k += 32) s += cs[2 * ((size_t)b * S + k)];
            if (pinv)
                for (int k = threadIdx.x; k < n; k += 32) {
                    const double d0 = z[(size_t)b * nzt + k] - xprior[(size_t)b * n + k];
                    s += pinv[k] * d0 * d0;
                }
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (threadIdx.x == 0) J[b] = s;
        }
    }
}

extern "C" int nlpf_ms_gather(int B, int S, int kn, int step, const double* z, double* zsub, void* stream) {
    if (B < 1 || S < 1 || kn < 1 || step < 1 || step >= kn || !z || !zsub) { nlpf_set_error("bad multi-segment gather argument"); return NLPF_E_ARG; }
    const size_t total = (size_t)B * S * kn;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    ms_gather_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(B, S, kn, step, S * step + (kn - step), z, zsub);
    return nlpf_check_cuda(cudaGetLastError(), "ms_gather launch");
}

extern "C" int nlpf_ms_blocks(int B, int S, int n, int bs, const double* H_seg, const double* g_seg, const double* cost_seg,
                              const double* pinv, const double* xprior, const double* z, const double* lam, double* Dg,
                              double* Lo, double* g, double* J, void* stream) {
    if (B < 1 || S < 1 || n < 1 || bs < n || bs % n != 0 || !H_seg || !g_seg || !cost_seg || !z || !Dg || !Lo || !g || !J || (pinv && !xprior)) {
        nlpf_set_error("bad multi-segment block-assembly argument");
        return NLPF_E_ARG;
    }
    ms_blocks_kernel<<<B < 148 * 8 ? B : 148 * 8, 128, 0, (cudaStream_t)stream>>>(B, S, n, bs, H_seg, g_seg, cost_seg, pinv, xprior, z, lam, Dg, Lo, g, J);
    return nlpf_check_cuda(cudaGetLastError(), "ms_blocks launch");
}
