// degnorm_b200 -- small-p fused baseline-selection kernels, P = 4 instantiation (see nmfoa_small.cuh).
#include "nmfoa_small.cuh"

int dn_launch_small4(const KArgs &a, const dn_plan *plan, cudaStream_t st) { return launch_small<4>(a, plan, st); }
