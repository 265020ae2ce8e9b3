// common.cuh -- shared declarations of libnlpfilter.so (sm_100a, fp64 throughout).
#pragma once
#ifdef __CUDACC_RTC__
// run-time compilation of the window kernels with a generated model function (user_model.cu): no host headers
typedef unsigned char uint8_t;
typedef unsigned short uint16_t;
typedef unsigned int uint32_t;
typedef unsigned long long size_t_rtc_unused;
#ifndef INFINITY
#define INFINITY __longlong_as_double(0x7ff0000000000000LL)
#endif
#else
#include <cuda_runtime.h>
#include <stdint.h>
#endif
#include "nlpfilter.h"

#define NLPF_NT 128            // threads per window CTA (4 warps: one per SM sub-partition)
#define NLPF_NW (NLPF_NT / 32)
#ifndef NLPF_MIN_CTAS
#define NLPF_MIN_CTAS 8        // register cap 64/thread: 8 windows resident per SM (ncu: 128 regs limited it to 4)
#endif
#define NLPF_NA_MAX 8          // components b handled per work item of the H accumulation (register block)
#define NLPF_NG_MAX 8
#define NLPF_N_MAX 12
#define NLPF_M_MAX 6
#define NLPF_K_MAX 64
#define NLPF_NV_MAX 24
#define NLPF_NZ_MAX 224

// Device-side image of one problem class.  Passed to kernels by value.
struct DevProb {
    int K, n, m, nv, ng, nz, np;      // np = packed lower-triangle size nz(nz+1)/2
    int dyn_id, loss, has_prior, has_dyncost, has_nodecost;
    int const_smem;                   // 1: Phi and Dt staged in shared memory, 0: read through L1
    int lin_meas;                     // 1: every residual is RT_STATE: its H contribution is constant per window
    int hconst;                       // 1: the dynamics part of H is constant (L2 loss, affine model)
    int fconst;                       // 1: dF/dv is constant (affine model)
    int nb, E, M, Ldata;
    int na, nat;                      // active node components, na + ng
    int ec;                           // epochs per measurement chunk
    int npair;                        // K(K+1)/2
    double delta;
    double dyn_par[NLPF_NDYNPAR];
    const double* A;                  // n*n
    const double* Bm;                 // n*mm
    const double* c0;                 // n
    const int* cmap;                  // n
    const int* ucol;                  // mm
    const double* Dt;                 // K*K
    const double* cw;                 // K
    const double* cq;                 // K
    const double* qinv;               // n
    const double* pinv;               // n
    const double* snode;              // nv
    const int* bidx;                  // nb
    const double* blo;
    const double* bhi;
    const double* Phi;                // E*K
    const double* PhiT;               // K*E (transposed copy: coalesced when lanes run over epochs)
    const int* mslot;                 // n  slot of state component a in every epoch's record set, or -1 (uniform linear measurements)
    const int* rtype;                 // M (sorted by epoch)
    const int* rloc;                  // M*8 local indices: <na active comp, >=na global, -1 unused
    const int* roff;                  // M
    const int* estart;                // E+1
    const int* gstart;                // E    first residual of epoch e that is not part of a pseudorange run
    const int* erun;                  // E+1  runs of epoch e: [erun[e], erun[e+1])
    const int* run_start;             // nrun first residual of the run
    const int* run_len;               // nrun
    const int* run_loc;               // nrun*4 local indices x,y,z,b shared by the run
    const int* act;                   // na  -> node-variable column
    const double* Hc;                 // np  constant part of H (packed lower, row-major)
    const double* Fc;                 // n*nv constant dF/dv (affine models)
    const uint32_t* tab;              // np  packed index -> j | a<<8 | l<<16 | b<<24 (255 = global)
    const uint16_t* ptab;             // npair  pair index -> j | l<<8
    const uint16_t* seg_tab;          // LDL trailing-update items of the generic triangle: di | (dk0/8)<<8
    const int* seg_cnt;               // nz+2  number of items of a trailing triangle with s rows
    // warp-per-window kernel: H in the row-padded packed layout (row r starts at 2*((r+1)/2)*((r+2)/2), so every
    // row is 16-byte aligned), 4-row x 2-column trailing-update tiles rq | ct<<8 sorted by row quad
    const uint16_t* sym_tab;          // general H accumulation: (a | b0<<8) blocks of 4 entries of the lower triangle of G_e
    int nsym;
    const uint8_t* gnz;               // nsym*E  1: the block is structurally non-zero at that epoch
    const double* PtP;                // npair  sum_e phi_ej phi_el per node pair (constant-weight measurement Hessian)
    const double* Hcp;                // Hc in the padded layout
    const uint16_t* tile_tab;
    const int* tile_cnt;              // nz+2  tiles of a trailing triangle with s rows
    // shared-memory layout, offsets in doubles
    int o_H, o_z, o_zt, o_g, o_ds, o_dd, o_dz, o_f, o_lw, o_wh, o_slw, o_F, o_GF, o_FF, o_whs;
    int o_Ve, o_G, o_q, o_Phi, o_Dt, o_red, smem_doubles;
};

#ifndef __CUDACC_RTC__
struct nlpf_problem {
    DevProb d;
    void* blob;          // one device allocation holding every table
    double* scratch;     // B x np constant-H cache of linear-measurement problems (grown on demand)
    size_t scratch_doubles;
    size_t blob_bytes;
    size_t smem_bytes;
    int use_pr;          // 1: problem has pseudorange runs -> *_pr kernels
    int use_big;         // 1: <= 5 windows per SM by shared memory -> 128-register instantiations (*_genb / *_prb)
    int ws_variant;      // >= 0: small linear-measurement class served by the warp-per-window kernel
    int* counter;        // device work counter of the warp-per-window kernel
    int* order;          // longest-first hand-out of the warp-per-window kernel: window indices by descending start cost
    double* keys;        // ... and the start costs
    size_t order_cap;
    void *user_module, *user_solve, *user_assemble;   // NVRTC-compiled window kernels of a user-defined model (user_model.cu)
    unsigned long long* prof;   // nlpf_problem_profile: [eval cycles, LDL cycles, evaluations, factorisations, window cycles] or null
    int device;
};

void nlpf_set_error(const char* fmt, ...);
int nlpf_check_cuda(cudaError_t e, const char* what);
#endif  // !__CUDACC_RTC__
