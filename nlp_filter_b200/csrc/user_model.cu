// user_model.cu -- dynamics models that are NOT in the kernel library.
//
// The reference hands ANY Python callable to casadi.Function (nlp/nlp.py:211-218) and differentiates it
// automatically.  Here nlp_filter_b200/nlp/_codegen.py traces the callable symbolically, differentiates it and emits
// one CUDA device function
//     __device__ void nlpf_user_dyn(const double* par, const double* x, const double* u, double* f, bool jac,
//                                   const FAcc& F, const int* ucol, double* Fk, int nv);
// and this file compiles the window kernels (csrc/nls_kernels.inl, the same source the library was built from) with
// that function included -- NVRTC, sm_100a -- loads the cubin, and keeps the two kernel handles (whole solve, assemble)
// in the problem.  nlpf_solve / nlpf_assemble launch them through the driver API.  libnvrtc and libcuda are opened
// with dlopen on first use, so the library still loads (and exports every symbol) on a machine without a driver.
#include "common.cuh"
#include <cuda.h>
#include <dlfcn.h>
#include <nvrtc.h>
#include <stdio.h>
#include <string.h>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace {
struct Rtc {
    void *nvrtc = nullptr, *cuda = nullptr;
    nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*);
    nvrtcResult (*DestroyProgram)(nvrtcProgram*);
    nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*);
    nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*);
    nvrtcResult (*GetProgramLog)(nvrtcProgram, char*);
    nvrtcResult (*AddNameExpression)(nvrtcProgram, const char*);
    nvrtcResult (*GetLoweredName)(nvrtcProgram, const char*, const char**);
    nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*);
    nvrtcResult (*GetCUBIN)(nvrtcProgram, char*);
    CUresult (*ModuleLoadData)(CUmodule*, const void*);
    CUresult (*ModuleUnload)(CUmodule);
    CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*);
    CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int);
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**, void**);
    bool ok = false;
};

Rtc& rtc() {
    static Rtc r;
    static bool tried = false;
    if (tried) return r;
    tried = true;
    const char* nv[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"};
    for (const char* n : nv) if (!r.nvrtc) r.nvrtc = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    const char* cu[] = {"libcuda.so.1", "libcuda.so"};
    for (const char* n : cu) if (!r.cuda) r.cuda = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (!r.nvrtc || !r.cuda) return r;
    bool all = true;
#define NV(sym) all = all && ((*(void**)(&r.sym) = dlsym(r.nvrtc, "nvrtc" #sym)) != nullptr)
    NV(CreateProgram); NV(DestroyProgram); NV(CompileProgram); NV(GetProgramLogSize); NV(GetProgramLog);
    NV(AddNameExpression); NV(GetLoweredName); NV(GetCUBINSize); NV(GetCUBIN);
#undef NV
#define CU(sym, name) all = all && ((*(void**)(&r.sym) = dlsym(r.cuda, name)) != nullptr)
    CU(ModuleLoadData, "cuModuleLoadData"); CU(ModuleUnload, "cuModuleUnload"); CU(ModuleGetFunction, "cuModuleGetFunction");
    CU(FuncSetAttribute, "cuFuncSetAttribute"); CU(LaunchKernel, "cuLaunchKernel");
#undef CU
    r.ok = all;
    return r;
}
// compiled models of this process: (kernel variant + generated source) -> module.  A batch in chunks creates one
// problem handle per chunk (engine.HostBatch): they share one compilation.  Modules stay loaded for the process.
struct UserMod { CUmodule mod; CUfunction fs, fa; };
std::map<std::string, UserMod>& mod_cache() { static std::map<std::string, UserMod> c; return c; }
std::mutex& mod_mutex() { static std::mutex m; return m; }
}  // namespace

extern "C" int nlpf_problem_set_user_dynamics(nlpf_problem* p, const char* model_src, const char* csrc_dir, const char* include_dir) {
    if (!p || !model_src || !csrc_dir || !include_dir) { nlpf_set_error("bad argument"); return NLPF_E_ARG; }
    if (p->d.dyn_id != NLPF_DYN_USER) { nlpf_set_error("the problem class was not created with dyn_id NLPF_DYN_USER"); return NLPF_E_ARG; }
    Rtc& R = rtc();
    if (!R.ok) { nlpf_set_error("libnvrtc / libcuda could not be opened: no run-time compilation of user models on this machine"); return NLPF_E_CUDA; }
    cudaFree(0);                                   // make the runtime's primary context current for the driver calls below
    char head[256];
    snprintf(head, sizeof(head), "#define NLPF_USER_DYN 1\n#define NLPF_PR_RUNS %d\n#define NLPF_BIG %d\n#include \"nls_kernels.inl\"\n",
             p->use_pr ? 1 : 0, p->use_big ? 1 : 0);
    const std::string key = std::string(head) + model_src;
    {
        std::lock_guard<std::mutex> lk(mod_mutex());
        auto it = mod_cache().find(key);
        if (it != mod_cache().end()) {
            p->user_module = nullptr;              // owned by the cache
            p->user_solve = it->second.fs;
            p->user_assemble = it->second.fa;
            return NLPF_OK;
        }
    }
    const char* suffix = p->use_big ? (p->use_pr ? "prb" : "genb") : (p->use_pr ? "pr" : "gen");
    const std::string n_solve = std::string("nlpf_ns_") + suffix + "::solve_kernel_" + suffix;
    const std::string n_asm = std::string("nlpf_ns_") + suffix + "::assemble_kernel_" + suffix;
    const char* hn[] = {"nlpf_user_dyn.inc"};
    const char* hs[] = {model_src};
    nvrtcProgram prog;
    if (R.CreateProgram(&prog, head, "nlpf_user_kernels.cu", 1, hs, hn) != NVRTC_SUCCESS) { nlpf_set_error("nvrtcCreateProgram failed"); return NLPF_E_CUDA; }
    R.AddNameExpression(prog, n_solve.c_str());
    R.AddNameExpression(prog, n_asm.c_str());
    const std::string i1 = std::string("-I") + csrc_dir, i2 = std::string("-I") + include_dir;
    const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "--fmad=true", "-lineinfo", i1.c_str(), i2.c_str()};
    const nvrtcResult rc = R.CompileProgram(prog, 6, opts);
    if (rc != NVRTC_SUCCESS) {
        size_t ls = 0;
        R.GetProgramLogSize(prog, &ls);
        std::string log(ls + 1, '\0');
        R.GetProgramLog(prog, &log[0]);
        nlpf_set_error("NVRTC could not compile the generated model (%d): %.1500s", (int)rc, log.c_str());
        R.DestroyProgram(&prog);
        return NLPF_E_ARG;
    }
    const char *l_solve = nullptr, *l_asm = nullptr;
    size_t cs = 0;
    if (R.GetLoweredName(prog, n_solve.c_str(), &l_solve) != NVRTC_SUCCESS || R.GetLoweredName(prog, n_asm.c_str(), &l_asm) != NVRTC_SUCCESS ||
        R.GetCUBINSize(prog, &cs) != NVRTC_SUCCESS || cs == 0) {
        nlpf_set_error("NVRTC produced no cubin / kernel names");
        R.DestroyProgram(&prog);
        return NLPF_E_CUDA;
    }
    std::vector<char> cubin(cs);
    R.GetCUBIN(prog, cubin.data());
    CUmodule mod = nullptr;
    CUfunction fs = nullptr, fa = nullptr;
    CUresult cr = R.ModuleLoadData(&mod, cubin.data());
    if (cr == CUDA_SUCCESS) cr = R.ModuleGetFunction(&fs, mod, l_solve);
    if (cr == CUDA_SUCCESS) cr = R.ModuleGetFunction(&fa, mod, l_asm);
    if (cr == CUDA_SUCCESS) cr = R.FuncSetAttribute(fs, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, 227 * 1024);
    if (cr == CUDA_SUCCESS) cr = R.FuncSetAttribute(fa, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, 227 * 1024);
    R.DestroyProgram(&prog);
    if (cr != CUDA_SUCCESS) {
        nlpf_set_error("loading the compiled model failed (CUresult %d)", (int)cr);
        if (mod) R.ModuleUnload(mod);
        return NLPF_E_CUDA;
    }
    {
        std::lock_guard<std::mutex> lk(mod_mutex());
        mod_cache()[key] = UserMod{mod, fs, fa};
    }
    p->user_module = nullptr;                      // owned by the cache
    p->user_solve = fs;
    p->user_assemble = fa;
    return NLPF_OK;
}

extern "C" void nlpf_user_release(nlpf_problem* p) {
    if (p && p->user_module && rtc().ok) rtc().ModuleUnload((CUmodule)p->user_module);
    if (p) p->user_module = p->user_solve = p->user_assemble = nullptr;
}

extern "C" int nlpf_user_launch_solve(nlpf_problem* p, int B, double* z, const double* U, const double* data, const double* xprior,
                                      const double* wshift, const nlpf_opts* o, double* cost_out, int* iters_out, int* status_out,
                                      double* hlin, cudaStream_t st) {
    if (!p->user_solve) { nlpf_set_error("this problem class has a user-defined model that has not been compiled (nlpf_problem_set_user_dynamics)"); return NLPF_E_ARG; }
    nlpf_opts opt = *o;
    void* args[] = {&p->d, &B, &z, &U, &data, &xprior, &wshift, &opt, &cost_out, &iters_out, &status_out, &hlin};
    const CUresult cr = rtc().LaunchKernel((CUfunction)p->user_solve, (unsigned)(B < 1 ? 1 : B), 1, 1, NLPF_NT, 1, 1, (unsigned)p->smem_bytes,
                                           (CUstream)st, args, nullptr);
    if (cr != CUDA_SUCCESS) { nlpf_set_error("cuLaunchKernel(user solve) failed (CUresult %d)", (int)cr); return NLPF_E_CUDA; }
    return NLPF_OK;
}

extern "C" int nlpf_user_launch_assemble(nlpf_problem* p, int B, const double* z, const double* U, const double* data,
                                         const double* xprior, const double* wshift, double mu, double* H_out, double* g_out,
                                         double* cost_out, cudaStream_t st) {
    if (!p->user_assemble) { nlpf_set_error("this problem class has a user-defined model that has not been compiled (nlpf_problem_set_user_dynamics)"); return NLPF_E_ARG; }
    void* args[] = {&p->d, &B, &z, &U, &data, &xprior, &wshift, &mu, &H_out, &g_out, &cost_out};
    const CUresult cr = rtc().LaunchKernel((CUfunction)p->user_assemble, (unsigned)(B < 1 ? 1 : B), 1, 1, NLPF_NT, 1, 1, (unsigned)p->smem_bytes,
                                           (CUstream)st, args, nullptr);
    if (cr != CUDA_SUCCESS) { nlpf_set_error("cuLaunchKernel(user assemble) failed (CUresult %d)", (int)cr); return NLPF_E_CUDA; }
    return NLPF_OK;
}
