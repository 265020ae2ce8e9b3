// BIG instantiation (128-register cap, min 4 CTAs per SM) with the pseudorange-run path
#define NLPF_PR_RUNS 1
#define NLPF_BIG 1
#include "nls_kernels.inl"
