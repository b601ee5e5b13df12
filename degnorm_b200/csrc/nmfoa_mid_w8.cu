// degnorm_b200 -- mid-p fused baseline-selection kernel, 8 warps per CTA, one CTA per SM (see nmfoa_mid.cuh).
#define MID_NW 8
#define MID_LAUNCHER dn_launch_mid8
#include "nmfoa_mid.cuh"
