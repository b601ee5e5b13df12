// degnorm_b200 -- mid-p fused baseline-selection kernel, warp-specialised: 8 Gram warps + 4 update warps per CTA, one
// CTA per SM (see nmfoa_mid.cuh, gram_mid_ws).
#define MID_NW 8
#define MID_NA 4
#define MID_LAUNCHER dn_launch_midws
#include "nmfoa_mid.cuh"
