// instantiation with the lane-group pseudorange-run path (GNSS, vehicle, RC, two-/multi-receiver problem classes)
#define NLPF_PR_RUNS 1
#include "nls_kernels.inl"
