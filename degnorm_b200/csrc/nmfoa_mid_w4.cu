// degnorm_b200 -- mid-p fused baseline-selection kernel, 4 warps per CTA, two CTAs per SM (see nmfoa_mid.cuh).
#define MID_NW 4
#define MID_LAUNCHER dn_launch_mid4
#include "nmfoa_mid.cuh"
