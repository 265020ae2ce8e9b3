// capi.cu -- extern "C" entry points of libnlpfilter.so (declared in include/nlpfilter.h):
// argument checking, derivation of the device tables of a problem class, kernel launches.
// No torch types cross this boundary; buffers are raw device pointers.
#include "common.cuh"
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

static thread_local char g_err[512] = "";

void nlpf_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int nlpf_check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return NLPF_OK;
    nlpf_set_error("%s: %s", what, cudaGetErrorString(e));
    return NLPF_E_CUDA;
}

extern "C" int nlpf_launch_assemble_gen(nlpf_problem*, int, const double*, const double*, const double*, const double*,
                                        const double*, double, double*, double*, double*, cudaStream_t);
extern "C" int nlpf_launch_assemble_pr(nlpf_problem*, int, const double*, const double*, const double*, const double*,
                                       const double*, double, double*, double*, double*, cudaStream_t);
extern "C" int nlpf_launch_factor_solve(nlpf_problem*, int, const double*, const double*, double, double*, int*,
                                        cudaStream_t);
extern "C" int nlpf_launch_solve_gen(nlpf_problem*, int, double*, const double*, const double*, const double*,
                                     const double*, const nlpf_opts*, double*, int*, int*, double*, cudaStream_t);
extern "C" int nlpf_launch_solve_pr(nlpf_problem*, int, double*, const double*, const double*, const double*,
                                    const double*, const nlpf_opts*, double*, int*, int*, double*, cudaStream_t);
extern "C" int nlpf_launch_sample(nlpf_problem*, int, const double*, int, const double*, double*, cudaStream_t);
extern "C" int nlpf_launch_defects(nlpf_problem*, int, const double*, const double*, double*, cudaStream_t);
extern "C" void nlpf_user_release(nlpf_problem* p);
extern "C" int nlpf_user_launch_solve(nlpf_problem*, int, double*, const double*, const double*, const double*, const double*,
                                      const nlpf_opts*, double*, int*, int*, double*, cudaStream_t);
extern "C" int nlpf_user_launch_assemble(nlpf_problem*, int, const double*, const double*, const double*, const double*,
                                         const double*, double, double*, double*, double*, cudaStream_t);
extern "C" int nlpf_ws_variant(int K, int N, int M, int E, int NM, int dyn, int meas, int nb, int prior, int loss);
extern "C" int nlpf_ws_wants_soa(int variant);
extern "C" int nlpf_ws_assemble(int variant, nlpf_problem* p, int B, const double* z, const double* U, const double* data,
                                const double* xprior, double mu, double* H_out, double* g_out, double* cost_out, cudaStream_t st);
extern "C" int nlpf_ws_launch(int variant, nlpf_problem* p, int B, double* z, const double* U, const double* data,
                              double* xprior, const nlpf_opts* o, double* cost_out, int* iters_out, int* status_out,
                              double* hlin, int* counter, int W, const double* phiDT, double* Zout, cudaStream_t st);
extern "C" int nlpf_set_smem_attrs_gen(size_t bytes);
extern "C" int nlpf_set_smem_attrs_pr(size_t bytes);
// BIG instantiations (nls_kernels_*_big.cu): 128-register cap for classes with <= 5 windows per SM
extern "C" int nlpf_launch_assemble_genb(nlpf_problem*, int, const double*, const double*, const double*, const double*,
                                         const double*, double, double*, double*, double*, cudaStream_t);
extern "C" int nlpf_launch_assemble_prb(nlpf_problem*, int, const double*, const double*, const double*, const double*,
                                        const double*, double, double*, double*, double*, cudaStream_t);
extern "C" int nlpf_launch_solve_genb(nlpf_problem*, int, double*, const double*, const double*, const double*,
                                      const double*, const nlpf_opts*, double*, int*, int*, double*, cudaStream_t);
extern "C" int nlpf_launch_solve_prb(nlpf_problem*, int, double*, const double*, const double*, const double*,
                                     const double*, const nlpf_opts*, double*, int*, int*, double*, cudaStream_t);
extern "C" int nlpf_set_smem_attrs_genb(size_t bytes);
extern "C" int nlpf_set_smem_attrs_prb(size_t bytes);

extern "C" int nlpf_version(void) { return NLPF_VERSION; }
extern "C" const char* nlpf_last_error(void) { return g_err; }

extern "C" int nlpf_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

// ------------------------------------------------------------------------------------------
// blob builder: every table of a problem class goes into one device allocation
// ------------------------------------------------------------------------------------------
namespace {
struct Blob {
    std::vector<unsigned char> host;
    size_t add(const void* src, size_t bytes) {
        size_t off = (host.size() + 15) & ~size_t(15);
        host.resize(off + bytes);
        if (bytes) memcpy(host.data() + off, src, bytes);
        return off;
    }
    template <typename T>
    size_t addv(const std::vector<T>& v) { return add(v.data(), v.size() * sizeof(T)); }
};

// which index slots of a residual are state components / global indices
void slot_roles(int rt, const int* idx, int* is_state, int* is_glob) {
    for (int s = 0; s < NLPF_NIDX; ++s) is_state[s] = is_glob[s] = 0;
    switch (rt) {
        case NLPF_RT_STATE: is_state[0] = 1; break;
        case NLPF_RT_PR: for (int s = 0; s < 4; ++s) is_state[s] = 1; break;
        case NLPF_RT_PRR: for (int s = 0; s < 7; ++s) is_state[s] = 1; break;
        case NLPF_RT_RANGE: for (int c = 0; c < idx[6]; ++c) is_state[c] = is_state[3 + c] = 1; break;
        case NLPF_RT_RANGE_PT:
            for (int c = 0; c < idx[6]; ++c) { is_state[c] = 1; if (idx[3 + c] >= 0) is_glob[3 + c] = 1; }
            break;
        case NLPF_RT_HEADING: is_state[0] = is_state[1] = is_state[3] = is_state[4] = 1; break;
        case NLPF_RT_HEADING_PT:
            is_state[0] = is_state[1] = 1;
            if (idx[3] >= 0) is_glob[3] = 1;
            if (idx[4] >= 0) is_glob[4] = 1;
            break;
    }
}
int reclen(int rt) {
    switch (rt) {
        case NLPF_RT_STATE: case NLPF_RT_RANGE: case NLPF_RT_HEADING: return 2;
        case NLPF_RT_PR: case NLPF_RT_RANGE_PT: case NLPF_RT_HEADING_PT: return 5;
        case NLPF_RT_PRR: return 8;
    }
    return -1;
}
}  // namespace

extern "C" int nlpf_problem_create(const nlpf_desc* D, nlpf_problem** out) {
    if (!D || !out) { nlpf_set_error("null argument"); return NLPF_E_ARG; }
    *out = nullptr;
    if (nlpf_device_count() < 1) {
        nlpf_set_error("no CUDA device visible: libnlpfilter has no CPU fallback");
        return NLPF_E_NOGPU;
    }
    {   // per-call workspaces come from the stream-ordered pool (cudaMallocAsync): keep them cached between calls -- with the
        // default release threshold every synchronisation returns them to the driver and the next call maps them again
        static bool pool_done = false;
        if (!pool_done) {
            int devid = 0;
            cudaMemPool_t pool;
            if (cudaGetDevice(&devid) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, devid) == cudaSuccess) {
                unsigned long long keep = ~0ULL;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
            pool_done = true;
        }
    }
    const int K = D->K, n = D->n, m = D->m, nv = D->nv, ng = D->ng;
    const int mm = m > 0 ? m : 1;
    const int nz = K * nv + ng;
    if (K < 2 || K > NLPF_K_MAX || n < 1 || n > NLPF_N_MAX || m < 0 || m > NLPF_M_MAX || nv < 1 || nv > NLPF_NV_MAX ||
        ng < 0 || ng > NLPF_NG_MAX || nz > NLPF_NZ_MAX || D->E < 0 || D->M < 0 || D->nb < 0) {
        nlpf_set_error("unsupported problem size K=%d n=%d m=%d nv=%d ng=%d nz=%d (limits K<=%d n<=%d m<=%d nv<=%d ng<=%d nz<=%d)",
                       K, n, m, nv, ng, nz, NLPF_K_MAX, NLPF_N_MAX, NLPF_M_MAX, NLPF_NV_MAX, NLPF_NG_MAX, NLPF_NZ_MAX);
        return NLPF_E_ARG;
    }
    if (D->dyn_id < 0 || D->dyn_id > NLPF_DYN_USER) { nlpf_set_error("unknown dynamics id %d", D->dyn_id); return NLPF_E_ARG; }
    static const int dyn_n[] = {0, 2, 6, 6, 9, 12, 0};
    static const int dyn_m[] = {0, 1, 2, 2, 2, 4, 0};
    if (D->dyn_id != NLPF_DYN_AFFINE && D->dyn_id != NLPF_DYN_USER && (n != dyn_n[D->dyn_id] || m != dyn_m[D->dyn_id])) {
        nlpf_set_error("dynamics id %d needs n=%d m=%d, got n=%d m=%d", D->dyn_id, dyn_n[D->dyn_id], dyn_m[D->dyn_id], n, m);
        return NLPF_E_ARG;
    }
    for (int i = 0; i < n; ++i)
        if (D->cmap[i] < 0 || D->cmap[i] >= nv) { nlpf_set_error("cmap[%d]=%d out of range", i, D->cmap[i]); return NLPF_E_ARG; }
    for (int i = 0; i < m; ++i) {
        if (D->ucol[i] >= nv) { nlpf_set_error("ucol[%d]=%d out of range", i, D->ucol[i]); return NLPF_E_ARG; }
        if (D->ucol[i] >= 0 && D->dyn_id != NLPF_DYN_AFFINE && D->dyn_id != NLPF_DYN_VDP) {
            nlpf_set_error("unknown controls are only supported for the affine and van der Pol models");
            return NLPF_E_ARG;
        }
    }
    for (int b = 0; b < D->nb; ++b) {
        if (D->bidx[b] < 0 || D->bidx[b] >= nv) { nlpf_set_error("bidx[%d] out of range", b); return NLPF_E_ARG; }
        for (int b2 = 0; b2 < b; ++b2)
            if (D->bidx[b2] == D->bidx[b]) { nlpf_set_error("duplicate bounded component %d", D->bidx[b]); return NLPF_E_ARG; }
    }

    nlpf_problem* p = new nlpf_problem();
    memset(&p->d, 0, sizeof(DevProb));
    p->scratch = nullptr;
    p->scratch_doubles = 0;
    DevProb& P = p->d;
    P.K = K; P.n = n; P.m = m; P.nv = nv; P.ng = ng; P.nz = nz; P.np = nz * (nz + 1) / 2;
    P.dyn_id = D->dyn_id; P.loss = D->loss; P.has_prior = D->has_prior ? 1 : 0;
    P.nb = D->nb; P.E = D->E; P.M = D->M; P.Ldata = D->Ldata;
    P.delta = D->delta;
    P.npair = K * (K + 1) / 2;
    memcpy(P.dyn_par, D->dyn_par, sizeof(double) * NLPF_NDYNPAR);
    P.has_dyncost = 0;
    for (int i = 0; i < n; ++i) if (D->qinv[i] != 0.0) P.has_dyncost = 1;
    P.has_nodecost = 0;
    for (int a = 0; a < nv; ++a) if (D->snode[a] != 0.0) P.has_nodecost = 1;
    P.fconst = (D->dyn_id == NLPF_DYN_AFFINE) ? 1 : 0;
    P.hconst = (P.fconst && D->loss == NLPF_LOSS_L2) ? 1 : 0;
    if (!P.has_dyncost) P.hconst = 1;

    // ---- residual table: validate, sort by epoch, derive the active component set ----
    const int M = D->M, E = D->E;
    std::vector<int> order(M);
    for (int i = 0; i < M; ++i) order[i] = i;
    for (int i = 0; i < M; ++i) {
        const int rt = D->rtype[i];
        if (reclen(rt) < 0) { nlpf_set_error("unknown residual type %d", rt); delete p; return NLPF_E_ARG; }
        if (D->repoch[i] < 0 || D->repoch[i] >= E) { nlpf_set_error("residual %d: epoch out of range", i); delete p; return NLPF_E_ARG; }
        if (D->roff[i] < 0 || D->roff[i] + reclen(rt) > D->Ldata) { nlpf_set_error("residual %d: record outside the data block", i); delete p; return NLPF_E_ARG; }
        if ((rt == NLPF_RT_RANGE || rt == NLPF_RT_RANGE_PT) && (D->ridx[i * NLPF_NIDX + 6] < 1 || D->ridx[i * NLPF_NIDX + 6] > 3)) {
            nlpf_set_error("residual %d: range dimension must be 1..3", i); delete p; return NLPF_E_ARG;
        }
    }
    for (int i = 0; i < M; ++i) {          // state / global indices, before anything dereferences cmap[ridx[..]]
        int is_state[NLPF_NIDX], is_glob[NLPF_NIDX];
        const int* idx = D->ridx + i * NLPF_NIDX;
        slot_roles(D->rtype[i], idx, is_state, is_glob);
        for (int s = 0; s < NLPF_NIDX; ++s) {
            if (is_state[s] && (idx[s] < 0 || idx[s] >= n)) { nlpf_set_error("residual %d: state index %d out of range", i, idx[s]); delete p; return NLPF_E_ARG; }
            if (is_glob[s] && idx[s] >= ng) { nlpf_set_error("residual %d: global index out of range", i); delete p; return NLPF_E_ARG; }
        }
    }
    for (int i = 0; i < n; ++i)
        if (D->cmap[i] < 0 || D->cmap[i] >= nv) { nlpf_set_error("cmap[%d] = %d outside the %d node variables", i, D->cmap[i], nv); delete p; return NLPF_E_ARG; }
    // order: by epoch; inside an epoch pseudoranges first, grouped by their index quadruple (runs), then the rest
    auto key = [&](int i, int s) { return D->rtype[i] == NLPF_RT_PR ? D->cmap[D->ridx[i * NLPF_NIDX + s]] : 0; };
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        if (D->repoch[a] != D->repoch[b]) return D->repoch[a] < D->repoch[b];
        const int pa = D->rtype[a] == NLPF_RT_PR ? 0 : 1, pb = D->rtype[b] == NLPF_RT_PR ? 0 : 1;
        if (pa != pb) return pa < pb;
        if (pa == 1) return false;
        for (int s = 0; s < 4; ++s) if (key(a, s) != key(b, s)) return key(a, s) < key(b, s);
        return false;
    });
    std::vector<int> colact(nv, 0);
    for (int i = 0; i < M; ++i) {
        int is_state[NLPF_NIDX], is_glob[NLPF_NIDX];
        const int* idx = D->ridx + i * NLPF_NIDX;
        slot_roles(D->rtype[i], idx, is_state, is_glob);
        for (int s = 0; s < NLPF_NIDX; ++s) {
            if (is_state[s]) {
                if (idx[s] < 0 || idx[s] >= n) { nlpf_set_error("residual %d: state index %d out of range", i, idx[s]); delete p; return NLPF_E_ARG; }
                colact[D->cmap[idx[s]]] = 1;
            }
            if (is_glob[s] && idx[s] >= ng) { nlpf_set_error("residual %d: global index out of range", i); delete p; return NLPF_E_ARG; }
        }
    }
    std::vector<int> act, actinv(nv, -1);
    for (int a = 0; a < nv; ++a) if (colact[a]) { actinv[a] = (int)act.size(); act.push_back(a); }
    const int na = (int)act.size();
    if (na > NLPF_NV_MAX) { nlpf_set_error("measurements touch %d node components (limit %d)", na, NLPF_NV_MAX); delete p; return NLPF_E_ARG; }
    P.na = na; P.nat = na + ng;
    P.lin_meas = (M > 0) ? 1 : 0;
    for (int i = 0; i < M; ++i) if (D->rtype[i] != NLPF_RT_STATE) P.lin_meas = 0;
    std::vector<int> rtype_s(M), rloc_s((size_t)M * NLPF_NIDX, -1), roff_s(M), estart(E + 1, 0);
    for (int r = 0; r < M; ++r) {
        const int i = order[r];
        rtype_s[r] = D->rtype[i];
        roff_s[r] = D->roff[i];
        estart[D->repoch[i] + 1]++;
        int is_state[NLPF_NIDX], is_glob[NLPF_NIDX];
        const int* idx = D->ridx + i * NLPF_NIDX;
        slot_roles(D->rtype[i], idx, is_state, is_glob);
        for (int s = 0; s < NLPF_NIDX; ++s) {
            int v = idx[s];
            if (is_state[s]) v = actinv[D->cmap[idx[s]]];
            else if (is_glob[s]) v = na + idx[s];
            else if (s >= 3 && s <= 5 && (D->rtype[i] == NLPF_RT_RANGE_PT || D->rtype[i] == NLPF_RT_HEADING_PT)) v = -1;
            rloc_s[(size_t)r * NLPF_NIDX + s] = v;
        }
    }
    for (int e = 0; e < E; ++e) estart[e + 1] += estart[e];
    // pseudorange runs per epoch
    std::vector<int> gstart(E > 0 ? E : 1, 0), erun(E + 1, 0), run_start, run_len, run_loc;
    for (int e = 0; e < E; ++e) {
        erun[e] = (int)run_start.size();
        int r = estart[e];
        while (r < estart[e + 1] && rtype_s[r] == NLPF_RT_PR) {
            int r2 = r + 1;
            while (r2 < estart[e + 1] && rtype_s[r2] == NLPF_RT_PR &&
                   std::equal(rloc_s.begin() + (size_t)r * NLPF_NIDX, rloc_s.begin() + (size_t)r * NLPF_NIDX + 4,
                              rloc_s.begin() + (size_t)r2 * NLPF_NIDX)) ++r2;
            run_start.push_back(r); run_len.push_back(r2 - r);
            for (int s4 = 0; s4 < 4; ++s4) run_loc.push_back(rloc_s[(size_t)r * NLPF_NIDX + s4]);
            r = r2;
        }
        gstart[e] = r;
    }
    erun[E] = (int)run_start.size();
    p->use_pr = run_start.empty() ? 0 : 1;       // which instantiation of the window kernels serves this class
    if (run_start.empty()) { run_start.push_back(0); run_len.push_back(0); run_loc.assign(4, 0); }
    // epochs per chunk: keep the per-chunk scratch (Ve, G, q) near 8 KB
    {
        const int per = P.nat * P.nat + P.nat + (na > 0 ? na : 1);
        int ec = 1024 / (per > 0 ? per : 1);
        if (ec < 1) ec = 1;
        if (ec > NLPF_NT) ec = NLPF_NT;
        if (ec > E) ec = E > 0 ? E : 1;
        P.ec = ec;
    }

    // ---- constant part of H ----
    std::vector<double> Hc((size_t)P.np, 0.0);
    auto HC = [&](int r, int c) -> double& { return r >= c ? Hc[(size_t)r * (r + 1) / 2 + c] : Hc[(size_t)c * (c + 1) / 2 + r]; };
    std::vector<double> Fc((size_t)n * nv, 0.0);
    if (P.fconst) {
        for (int i = 0; i < n; ++i) {
            for (int j = 0; j < n; ++j) Fc[(size_t)i * nv + D->cmap[j]] += D->A[i * n + j];
            for (int j = 0; j < m; ++j) if (D->ucol[j] >= 0) Fc[(size_t)i * nv + D->ucol[j]] += D->Bm[i * mm + j];
        }
    }
    if (P.has_dyncost && D->loss == NLPF_LOSS_L2) {
        // rows of the defect Jacobian with the state-dependent block left out (or constant for affine models)
        const int R = K * n, C = K * nv;
        std::vector<double> Aj((size_t)R * C, 0.0);
        for (int k = 0; k < K; ++k)
            for (int i = 0; i < n; ++i) {
                double* row = &Aj[(size_t)(k * n + i) * C];
                for (int j = 0; j < K; ++j) row[j * nv + D->cmap[i]] += D->Dt[k * K + j];
                if (P.fconst) for (int a = 0; a < nv; ++a) row[k * nv + a] -= Fc[(size_t)i * nv + a];
            }
        for (int r = 0; r < C; ++r)
            for (int c = 0; c <= r; ++c) {
                double s = 0.0;
                for (int k = 0; k < K; ++k)
                    for (int i = 0; i < n; ++i) {
                        const double* row = &Aj[(size_t)(k * n + i) * C];
                        s += D->cw[k] * D->qinv[i] * row[r] * row[c];
                    }
                HC(r, c) += s;
            }
    }
    if (P.has_prior) for (int i = 0; i < n; ++i) HC(D->cmap[i], D->cmap[i]) += D->pinv[i];
    if (P.has_nodecost)
        for (int k = 0; k < K; ++k) for (int a = 0; a < nv; ++a) HC(k * nv + a, k * nv + a) += D->cq[k] * D->snode[a];

    // ---- index tables ----
    std::vector<uint32_t> tab((size_t)P.np);
    for (int r = 0; r < nz; ++r)
        for (int c = 0; c <= r; ++c) {
            uint32_t j, a, l, b;
            if (r < K * nv) { j = r / nv; a = r % nv; } else { j = 255; a = r - K * nv; }
            if (c < K * nv) { l = c / nv; b = c % nv; } else { l = 255; b = c - K * nv; }
            tab[(size_t)r * (r + 1) / 2 + c] = j | (a << 8) | (l << 16) | (b << 24);
        }
    std::vector<uint16_t> ptab((size_t)P.npair);
    {
        int q = 0;
        for (int j = 0; j < K; ++j) for (int l = 0; l <= j; ++l) ptab[q++] = (uint16_t)(j | (l << 8));
    }

    // transposed basis and the uniform linear-measurement layout (warp-per-window kernel)
    std::vector<double> PhiT((size_t)K * (E > 0 ? E : 1), 0.0);
    for (int e = 0; e < E; ++e) for (int j = 0; j < K; ++j) PhiT[(size_t)j * E + e] = D->Phi[(size_t)e * K + j];
    std::vector<int> mslot(n, -1);
    p->ws_variant = -1;
    p->counter = nullptr;
    p->order = nullptr;
    p->keys = nullptr;
    p->order_cap = 0;
    {
        bool ident = (nv == n) && (ng == 0);
        for (int i = 0; i < n; ++i) if (D->cmap[i] != i) ident = false;
        bool upar = true;
        for (int i = 0; i < m; ++i) if (D->ucol[i] >= 0) upar = false;
        const bool plain = ident && upar && !P.has_nodecost && P.has_dyncost && !getenv("NLPF_NO_WARP_KERNEL");
        if (E > 0 && M > 0 && M % E == 0) {
            const int NM = M / E;
            bool uniform = true;                    // every epoch carries the same record set, in the same order
            for (int e = 0; e < E && uniform; ++e) {
                if (estart[e + 1] - estart[e] != NM) uniform = false;
                for (int s2 = 0; s2 < NM && uniform; ++s2) {
                    if (rtype_s[e * NM + s2] != rtype_s[s2]) uniform = false;
                    for (int q = 0; q < 4 && uniform; ++q)
                        if (rloc_s[(size_t)(e * NM + s2) * NLPF_NIDX + q] != rloc_s[(size_t)s2 * NLPF_NIDX + q] &&
                            (q == 0 || rtype_s[s2] == NLPF_RT_PR)) uniform = false;
                }
            }
            // the warp kernel addresses records arithmetically: [y, w] / [y, w, sat xyz] back to back
            const int rlen = P.lin_meas ? 2 : 5;
            for (int r = 0; r < M && uniform; ++r) if (roff_s[r] != rlen * r) uniform = false;
            if (uniform && P.lin_meas && NM > 4) uniform = false;
            if (uniform && P.lin_meas) {
                for (int s2 = 0; s2 < NM; ++s2) mslot[act[rloc_s[(size_t)s2 * NLPF_NIDX]]] = s2;
                if (plain) p->ws_variant = nlpf_ws_variant(K, n, m, E, NM, D->dyn_id, 0, D->nb, P.has_prior, D->loss);
            } else if (uniform && plain) {
                bool allpr = true;                  // one pseudorange run per epoch with a single index quadruple
                for (int r = 0; r < M; ++r) if (rtype_s[r] != NLPF_RT_PR) allpr = false;
                for (int s2 = 1; s2 < NM && allpr; ++s2)
                    for (int q = 0; q < 4; ++q)
                        if (rloc_s[(size_t)s2 * NLPF_NIDX + q] != rloc_s[q]) allpr = false;
                if (allpr) p->ws_variant = nlpf_ws_variant(K, n, m, E, NM, D->dyn_id, 1, D->nb, P.has_prior, D->loss);
            }
        }
    }

    // LDL trailing-update work items: (row, 8-column segment) of the generic lower triangle with
    // nz+1 rows, row-major, so that the items of a trailing triangle with s rows are a prefix
    std::vector<uint16_t> seg_tab;
    std::vector<int> seg_cnt(nz + 2, 0);
    for (int di = 0; di <= nz; ++di) {
        seg_cnt[di] = (int)seg_tab.size();
        for (int dk0 = 0; dk0 <= di; dk0 += 8) seg_tab.push_back((uint16_t)(di | ((dk0 >> 3) << 8)));
    }
    seg_cnt[nz + 1] = (int)seg_tab.size();

    // warp-per-window kernel: padded packed layout of the constant part, 2x4 tiles of the generic triangle
    auto ppo = [](int r) { return 2 * ((r + 1) >> 1) * ((r + 2) >> 1); };
    std::vector<double> Hcp((size_t)ppo(nz), 0.0);
    for (int r = 0; r < nz; ++r)
        for (int c = 0; c <= r; ++c) Hcp[(size_t)ppo(r) + c] = Hc[(size_t)r * (r + 1) / 2 + c];
    std::vector<uint16_t> tile_tab;                // 4-row x 2-column tiles: row quad rq | column pair ct << 8
    std::vector<int> tile_cnt(nz + 2, 0);
    for (int s = 0; s <= nz + 1; ++s) {            // rows di < s; row quad rq covers di = 4rq .. 4rq+3
        if (s >= 1 && (s - 1) % 4 == 0) {          // a new row quad starts at di = s-1
            const int rq = (s - 1) / 4;
            for (int ct = 0; 2 * ct <= 4 * rq + 3; ++ct) tile_tab.push_back((uint16_t)(rq | (ct << 8)));
        }
        tile_cnt[s] = (int)tile_tab.size();
    }

    // sum over the epochs of phi_ej phi_el per node pair (order of ptab)
    std::vector<double> PtP(ptab.size(), 0.0);
    for (size_t pr = 0; pr < ptab.size(); ++pr) {
        const int j = ptab[pr] & 255, l = ptab[pr] >> 8;
        double acc = 0.0;
        for (int e = 0; e < E; ++e) acc += D->Phi[(size_t)e * K + j] * D->Phi[(size_t)e * K + l];
        PtP[pr] = acc;
    }
    // general measurement-Hessian accumulation: G_e is symmetric, so only b <= a is summed, in blocks of 4
    std::vector<uint16_t> sym_tab;
    for (int a = 0; a < na; ++a)
        for (int b0 = 0; b0 <= a; b0 += 4) sym_tab.push_back((uint16_t)(a | (b0 << 8)));
    P.nsym = (int)sym_tab.size();
    // ... and only at the epochs where a residual couples a with one of the block's columns (two-receiver shape:
    // 15 of the 45 entries at a range/heading epoch, 19 at a pseudorange epoch)
    std::vector<uint8_t> gnz((size_t)(P.nsym > 0 ? P.nsym : 1) * (E > 0 ? E : 1), 0);
    {
        std::vector<int> item_start(na + 1, 0);
        for (int a = 0; a < na; ++a) item_start[a + 1] = item_start[a] + a / 4 + 1;
        for (int e = 0; e < E; ++e)
            for (int r = estart[e]; r < estart[e + 1]; ++r) {
                int comps[NLPF_NIDX], nc = 0;
                for (int s2 = 0; s2 < NLPF_NIDX; ++s2) {
                    const int v = rloc_s[(size_t)r * NLPF_NIDX + s2];
                    if (v >= 0 && v < na) comps[nc++] = v;
                }
                for (int x = 0; x < nc; ++x)
                    for (int y = 0; y < nc; ++y)
                        if (comps[y] <= comps[x]) gnz[(size_t)(item_start[comps[x]] + comps[y] / 4) * E + e] = 1;
            }
    }

    // ---- blob ----
    Blob B;
    std::vector<double> zero1(1, 0.0);
    const size_t oA = B.add(D->A, sizeof(double) * n * n), oB = B.add(D->Bm, sizeof(double) * n * mm),
                 oc0 = B.add(D->c0, sizeof(double) * n), ocm = B.add(D->cmap, sizeof(int) * n),
                 ouc = B.add(D->ucol, sizeof(int) * mm), oDt = B.add(D->Dt, sizeof(double) * K * K),
                 ocw = B.add(D->cw, sizeof(double) * K), ocq = B.add(D->cq, sizeof(double) * K),
                 oqi = B.add(D->qinv, sizeof(double) * n), opi = B.add(D->pinv, sizeof(double) * n),
                 osn = B.add(D->snode, sizeof(double) * nv), obi = B.add(D->bidx, sizeof(int) * D->nb),
                 obl = B.add(D->blo, sizeof(double) * D->nb), obh = B.add(D->bhi, sizeof(double) * D->nb),
                 oPh = B.add(D->Phi, sizeof(double) * E * K), ort = B.addv(rtype_s), orl = B.addv(rloc_s),
                 oro = B.addv(roff_s), oes = B.addv(estart), oac = B.addv(act), oHc = B.addv(Hc), oFc = B.addv(Fc),
                 otb = B.addv(tab), opt = B.addv(ptab), osg = B.addv(seg_tab), osc = B.addv(seg_cnt),
                 oPT = B.addv(PhiT), oHp = B.addv(Hcp), oPP = B.addv(PtP), osy = B.addv(sym_tab), ogz = B.addv(gnz), ott = B.addv(tile_tab), otc = B.addv(tile_cnt), oms = B.addv(mslot), ogs = B.addv(gstart), oer = B.addv(erun), ors = B.addv(run_start), orn = B.addv(run_len), orc = B.addv(run_loc);
    B.add(zero1.data(), 16);
    p->blob_bytes = B.host.size();
    cudaGetDevice(&p->device);
    if (nlpf_check_cuda(cudaMalloc(&p->blob, p->blob_bytes), "cudaMalloc(problem tables)") != NLPF_OK) { delete p; return NLPF_E_CUDA; }
    if (nlpf_check_cuda(cudaMemcpy(p->blob, B.host.data(), p->blob_bytes, cudaMemcpyHostToDevice), "upload problem tables") != NLPF_OK) {
        cudaFree(p->blob); delete p; return NLPF_E_CUDA;
    }
    const unsigned char* base = (const unsigned char*)p->blob;
    P.A = (const double*)(base + oA); P.Bm = (const double*)(base + oB); P.c0 = (const double*)(base + oc0);
    P.cmap = (const int*)(base + ocm); P.ucol = (const int*)(base + ouc); P.Dt = (const double*)(base + oDt);
    P.cw = (const double*)(base + ocw); P.cq = (const double*)(base + ocq); P.qinv = (const double*)(base + oqi);
    P.pinv = (const double*)(base + opi); P.snode = (const double*)(base + osn); P.bidx = (const int*)(base + obi);
    P.blo = (const double*)(base + obl); P.bhi = (const double*)(base + obh); P.Phi = (const double*)(base + oPh);
    P.rtype = (const int*)(base + ort); P.rloc = (const int*)(base + orl); P.roff = (const int*)(base + oro);
    P.estart = (const int*)(base + oes); P.act = (const int*)(base + oac); P.Hc = (const double*)(base + oHc);
    P.Fc = (const double*)(base + oFc); P.tab = (const uint32_t*)(base + otb); P.ptab = (const uint16_t*)(base + opt);
    P.seg_tab = (const uint16_t*)(base + osg); P.seg_cnt = (const int*)(base + osc);
    P.sym_tab = (const uint16_t*)(base + osy); P.gnz = (const uint8_t*)(base + ogz);
    P.PtP = (const double*)(base + oPP);
    P.Hcp = (const double*)(base + oHp); P.tile_tab = (const uint16_t*)(base + ott); P.tile_cnt = (const int*)(base + otc);
    P.PhiT = (const double*)(base + oPT); P.mslot = (const int*)(base + oms);
    P.gstart = (const int*)(base + ogs); P.erun = (const int*)(base + oer); P.run_start = (const int*)(base + ors);
    P.run_len = (const int*)(base + orn); P.run_loc = (const int*)(base + orc);

    // ---- shared-memory layout (doubles) ----
    auto do_layout = [&]() {
        int o = 0;
        auto take = [&](int cnt) { int r = o; o += (cnt + 1) & ~1; return r; };
        P.o_H = take((nz + 1) * (nz + 2) / 2);
        P.o_z = take(nz); P.o_zt = take(nz); P.o_g = take(nz); P.o_ds = take(nz); P.o_dd = take(nz); P.o_dz = take(nz);
        P.o_f = take(K * n); P.o_lw = take(K * n); P.o_wh = take(K * n); P.o_slw = take(K * nv); P.o_whs = take(K * nv);
        const bool needF = !P.hconst;
        P.o_F = take(needF ? K * n * nv : 0);
        P.o_GF = take(needF ? K * nv * nv : 0);
        P.o_FF = take(needF ? K * nv * nv : 0);
        P.o_Ve = take(P.ec * (na > 0 ? na : 1));
        P.o_G = take(P.ec * P.nat * P.nat);
        P.o_q = take(P.ec * P.nat);
        // the evaluation scratch [o_f, here) doubles as the W^T side buffer of the tiled LDL^T (BIG instantiation)
        if (o - P.o_f < 4 * ((nz + 5) & ~1) + 8) take(4 * ((nz + 5) & ~1) + 8 - (o - P.o_f));
        {   // class constants staged in shared memory (default) or read through L1 (NLPF_CONST_SMEM=0: experiments)
            const char* ev = getenv("NLPF_CONST_SMEM");
            P.const_smem = ev ? atoi(ev) : 1;
        }
        P.o_Phi = take(P.const_smem ? E * K : 0);
        P.o_Dt = take(P.const_smem ? K * K : 0);
        P.o_red = take(32);
        P.smem_doubles = o;
        p->smem_bytes = (size_t)o * sizeof(double);
    };
    do_layout();
    // Classes that fit at most 5 windows per SM anyway (two-receiver, RC, multi-receiver): spend the shared memory
    // a 4-window residency leaves on longer epoch chunks -- fewer passes (barriers, read-modify-writes of H) over
    // the measurement phase.
    if ((227 * 1024) / (p->smem_bytes + 1024) <= 5 && E > 0) {
        const int ctas0 = (int)((227 * 1024) / (p->smem_bytes + 1024));
        const int keep = ctas0 < 4 ? ctas0 : 4;
        const int ec0 = P.ec;
        int best = ec0;
        for (int ec = ec0 + 1; ec <= E && ec <= NLPF_NT; ++ec) {
            P.ec = ec;
            do_layout();
            if ((int)((227 * 1024) / (p->smem_bytes + 1024)) >= keep && keep >= 1) best = ec; else break;
        }
        // whole chunks of equal size: the smallest chunk length with the same number of chunks
        const int nch = (E + best - 1) / best;
        P.ec = (E + nch - 1) / nch;
        do_layout();
    }
    if (p->smem_bytes > 227 * 1024) {
        nlpf_set_error("problem needs %zu bytes of shared memory per window (limit 232448)", p->smem_bytes);
        cudaFree(p->blob); delete p; return NLPF_E_ARG;
    }
    // shared memory allows at most 5 windows per SM: the 128-register instantiation loses no residency
    p->use_big = ((227 * 1024) / (p->smem_bytes + 1024) <= 5 && !getenv("NLPF_NO_BIG_KERNEL")) ? 1 : 0;
    if (nlpf_set_smem_attrs_gen(227 * 1024) != NLPF_OK || nlpf_set_smem_attrs_pr(227 * 1024) != NLPF_OK ||
        nlpf_set_smem_attrs_genb(227 * 1024) != NLPF_OK || nlpf_set_smem_attrs_prb(227 * 1024) != NLPF_OK) {
        cudaFree(p->blob); delete p; return NLPF_E_CUDA;
    }
    *out = p;
    return NLPF_OK;
}

extern "C" int nlpf_problem_destroy(nlpf_problem* p) {
    if (!p) return NLPF_OK;
    cudaFree(p->blob);
    if (p->scratch) cudaFree(p->scratch);
    if (p->counter) cudaFree(p->counter);
    if (p->order) cudaFree(p->order);
    if (p->prof) cudaFree(p->prof);
    nlpf_user_release(p);
    if (p->keys) cudaFree(p->keys);
    delete p;
    return NLPF_OK;
}

extern "C" int nlpf_problem_nz(const nlpf_problem* p) { return p ? p->d.nz : NLPF_E_ARG; }
extern "C" int nlpf_problem_smem_bytes(const nlpf_problem* p) { return p ? (int)p->smem_bytes : NLPF_E_ARG; }

static int check_batch(nlpf_problem* p, int B, const void* z, const void* data) {
    if (!p) { nlpf_set_error("null problem handle"); return NLPF_E_ARG; }
    if (B < 1) { nlpf_set_error("batch size must be >= 1"); return NLPF_E_ARG; }
    if (!z) { nlpf_set_error("null iterate buffer"); return NLPF_E_ARG; }
    if (p->d.M > 0 && !data) { nlpf_set_error("problem has %d residuals but the data buffer is null", p->d.M); return NLPF_E_ARG; }
    return NLPF_OK;
}

extern "C" int nlpf_assemble(nlpf_problem* p, int B, const double* z, const double* U, const double* data,
                             const double* xprior, const double* wshift, double mu, double* H_out, double* g_out,
                             double* cost_out, void* stream) {
    int rc = check_batch(p, B, z, data);
    if (rc) return rc;
    if (!H_out || !g_out || !cost_out) { nlpf_set_error("null output buffer"); return NLPF_E_ARG; }
    if (p->d.dyn_id == NLPF_DYN_USER)
        return nlpf_user_launch_assemble(p, B, z, U, data, xprior, wshift, mu, H_out, g_out, cost_out, (cudaStream_t)stream);
    if (p->ws_variant >= 0 && !wshift && (p->d.m == 0 || U) && (!p->d.has_prior || xprior))      // warp-per-window classes
        return nlpf_ws_assemble(p->ws_variant, p, B, z, U, data, xprior, mu, H_out, g_out, cost_out, (cudaStream_t)stream);
    auto launch = p->use_big ? (p->use_pr ? nlpf_launch_assemble_prb : nlpf_launch_assemble_genb)
                             : (p->use_pr ? nlpf_launch_assemble_pr : nlpf_launch_assemble_gen);
    return launch(p, B, z, U, data, xprior, wshift, mu, H_out, g_out, cost_out, (cudaStream_t)stream);
}

extern "C" int nlpf_factor_solve(nlpf_problem* p, int B, const double* H, const double* g, double lam, double* dz_out,
                                 int* info_out, void* stream) {
    if (!p || B < 1 || !H || !g || !dz_out || !info_out) { nlpf_set_error("bad argument"); return NLPF_E_ARG; }
    return nlpf_launch_factor_solve(p, B, H, g, lam, dz_out, info_out, (cudaStream_t)stream);
}

// damping-schedule fields <= 0 -> defaults (include/nlpfilter.h, NLPF_LAM_*; oracle/nls.py: damping_schedule)
static nlpf_opts resolve_opts(const nlpf_problem* p, const nlpf_opts* opts) {
    (void)p;
    nlpf_opts o = *opts;
    if (!(o.lam0 > 0.0)) o.lam0 = NLPF_LAM0;
    if (!(o.lam_jump > 0.0)) o.lam_jump = NLPF_LAM_JUMP;
    if (!(o.lam_jump_indef > 0.0)) o.lam_jump_indef = NLPF_LAM_JUMP_INDEF;
    if (!(o.lam_dec_first > 0.0)) o.lam_dec_first = NLPF_LAM_DEC_FIRST;
    return o;
}

// warp-per-window kernel (its own constant-H cache and work counter); W > 1 / phiDT: fused receding-horizon chains
static bool ws_eligible(const nlpf_problem* p, const double* U, const double* xprior, const double* wshift) {
    return p->ws_variant >= 0 && !wshift && (p->d.m == 0 || U) && (!p->d.has_prior || xprior);
}

static int ws_dispatch(nlpf_problem* p, int B, double* z, const double* U, const double* data, double* xprior,
                       const nlpf_opts* opts, double* cost_out, int* iters_out, int* status_out, int W,
                       const double* phiDT, double* Zout, cudaStream_t stream) {
    const int nzp = p->d.nz;
    size_t need = p->d.lin_meas ? (size_t)B * (2 * ((nzp + 1) >> 1) * ((nzp + 2) >> 1)) : 0;   // padded packed rows
    if (nlpf_ws_wants_soa(p->ws_variant)) need = (size_t)W * B * p->d.Ldata;                        // structure-of-arrays copy of the records
    if (need > p->scratch_doubles) {
        if (p->scratch) { cudaStreamSynchronize(stream); cudaFree(p->scratch); p->scratch = nullptr; p->scratch_doubles = 0; }
        if (nlpf_check_cuda(cudaMalloc((void**)&p->scratch, need * sizeof(double)), "cudaMalloc(constant-H cache)")) return NLPF_E_CUDA;
        p->scratch_doubles = need;
    }
    if (!p->counter && nlpf_check_cuda(cudaMalloc((void**)&p->counter, sizeof(int)), "cudaMalloc(work counter)")) return NLPF_E_CUDA;
    if (W == 1 && (size_t)B > p->order_cap && B <= 16384) {      // hand-out order of the batch (wsolve_kernels.cu: ws_key_kernel)
        if (p->order) { cudaStreamSynchronize(stream); cudaFree(p->order); cudaFree(p->keys); p->order = nullptr; p->keys = nullptr; p->order_cap = 0; }
        if (nlpf_check_cuda(cudaMalloc((void**)&p->order, (size_t)B * sizeof(int)), "cudaMalloc(hand-out order)")) return NLPF_E_CUDA;
        if (nlpf_check_cuda(cudaMalloc((void**)&p->keys, (size_t)B * sizeof(double)), "cudaMalloc(hand-out keys)")) return NLPF_E_CUDA;
        p->order_cap = (size_t)B;
    }
    return nlpf_ws_launch(p->ws_variant, p, B, z, U, data, xprior, opts, cost_out, iters_out, status_out, p->scratch,
                          p->counter, W, phiDT, Zout, stream);
}

// Per-phase clock counters of the warp-per-window kernel (bench.py: roofline.eval_frac / ldl_frac).  enable != 0
// allocates and zeroes five device counters that every following nlpf_solve / nlpf_solve_chain on this handle adds
// to; out (host, 5 values, may be null) receives [warp-cycles in the evaluation, warp-cycles in factorisation +
// sweep, evaluations, factorisations, warp-cycles on windows] accumulated so far (synchronises the device);
// enable == 0 switches the counters off again.
extern "C" int nlpf_problem_profile(nlpf_problem* p, int enable, unsigned long long* out) {
    if (!p) { nlpf_set_error("null problem handle"); return NLPF_E_ARG; }
    if (out) {
        for (int i = 0; i < 5; ++i) out[i] = 0;
        if (p->prof && nlpf_check_cuda(cudaMemcpy(out, p->prof, 5 * sizeof(unsigned long long), cudaMemcpyDeviceToHost), "profile read")) return NLPF_E_CUDA;
    }
    if (enable) {
        if (!p->prof && nlpf_check_cuda(cudaMalloc((void**)&p->prof, 5 * sizeof(unsigned long long)), "cudaMalloc(profile)")) return NLPF_E_CUDA;
        if (nlpf_check_cuda(cudaMemset(p->prof, 0, 5 * sizeof(unsigned long long)), "profile reset")) return NLPF_E_CUDA;
    } else if (p->prof) {
        cudaDeviceSynchronize();
        cudaFree(p->prof);
        p->prof = nullptr;
    }
    return NLPF_OK;
}

extern "C" int nlpf_problem_has_chain_kernel(const nlpf_problem* p) { return (p && p->ws_variant >= 0 && p->d.has_prior) ? 1 : 0; }

extern "C" int nlpf_solve_chain(nlpf_problem* p, int B, int W, double* z, const double* U_all, const double* data_all,
                                double* xprior, const double* phiDT, const nlpf_opts* opts, double* cost_out,
                                int* iters_sum_out, int* bad_out, double* Z_out, void* stream) {
    int rc = check_batch(p, B, z, data_all);
    if (rc) return rc;
    if (W < 1 || !xprior || !phiDT || !opts || !cost_out || !iters_sum_out || !bad_out) { nlpf_set_error("bad chain argument"); return NLPF_E_ARG; }
    if (!p->d.has_prior || !ws_eligible(p, U_all, xprior, nullptr)) {
        nlpf_set_error("no fused chain kernel for this problem class (loop nlpf_solve + nlpf_sample instead)");
        return NLPF_E_ARG;
    }
    const nlpf_opts ro = resolve_opts(p, opts);
    return ws_dispatch(p, B, z, U_all, data_all, xprior, &ro, cost_out, iters_sum_out, bad_out, W, phiDT, Z_out, (cudaStream_t)stream);
}

extern "C" int nlpf_solve(nlpf_problem* p, int B, double* z, const double* U, const double* data, const double* xprior,
                          const double* wshift, const nlpf_opts* opts, double* cost_out, int* iters_out,
                          int* status_out, void* stream) {
    int rc = check_batch(p, B, z, data);
    if (rc) return rc;
    if (!opts || !cost_out || !iters_out || !status_out) { nlpf_set_error("null opts/output buffer"); return NLPF_E_ARG; }
    const nlpf_opts ro = resolve_opts(p, opts);
    opts = &ro;
    if (ws_eligible(p, U, xprior, wshift))
        return ws_dispatch(p, B, z, U, data, const_cast<double*>(xprior), opts, cost_out, iters_out, status_out, 1, nullptr, nullptr,
                           (cudaStream_t)stream);
    double* hlin = nullptr;
    if (p->d.lin_meas) {        // per-window constant-H cache (not shared between concurrent calls on one handle)
        const size_t need = (size_t)B * p->d.np;
        if (need > p->scratch_doubles) {
            if (p->scratch) { cudaStreamSynchronize((cudaStream_t)stream); cudaFree(p->scratch); p->scratch = nullptr; p->scratch_doubles = 0; }
            if (nlpf_check_cuda(cudaMalloc((void**)&p->scratch, need * sizeof(double)), "cudaMalloc(constant-H cache)")) return NLPF_E_CUDA;
            p->scratch_doubles = need;
        }
        hlin = p->scratch;
    }
    if (p->d.dyn_id == NLPF_DYN_USER)
        return nlpf_user_launch_solve(p, B, z, U, data, xprior, wshift, opts, cost_out, iters_out, status_out, hlin, (cudaStream_t)stream);
    auto launch = p->use_big ? (p->use_pr ? nlpf_launch_solve_prb : nlpf_launch_solve_genb)
                             : (p->use_pr ? nlpf_launch_solve_pr : nlpf_launch_solve_gen);
    return launch(p, B, z, U, data, xprior, wshift, opts, cost_out, iters_out, status_out, hlin, (cudaStream_t)stream);
}

extern "C" int nlpf_defects(nlpf_problem* p, int B, const double* z, const double* U, double* W_out, void* stream) {
    if (!p || B < 1 || !z || !W_out) { nlpf_set_error("bad argument"); return NLPF_E_ARG; }
    return nlpf_launch_defects(p, B, z, U, W_out, (cudaStream_t)stream);
}

extern "C" int nlpf_sample(nlpf_problem* p, int B, const double* z, int nt, const double* PhiT, double* out, void* stream) {
    if (!p || B < 1 || nt < 1 || !z || !PhiT || !out) { nlpf_set_error("bad argument"); return NLPF_E_ARG; }
    return nlpf_launch_sample(p, B, z, nt, PhiT, out, (cudaStream_t)stream);
}
