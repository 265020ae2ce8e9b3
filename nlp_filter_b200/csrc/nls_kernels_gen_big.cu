// BIG instantiation (128-register cap, min 4 CTAs per SM) of the generic window kernels
#define NLPF_PR_RUNS 0
#define NLPF_BIG 1
#include "nls_kernels.inl"
