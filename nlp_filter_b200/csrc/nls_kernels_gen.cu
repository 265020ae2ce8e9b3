// generic instantiation of the window kernels (no pseudorange-run path); also kernel (3) alone and nlpf_sample
#define NLPF_PR_RUNS 0
#include "nls_kernels.inl"
