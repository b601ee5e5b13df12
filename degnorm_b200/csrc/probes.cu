// degnorm_b200 -- measurement probes (not on the product path): the peaks the roofline of the fp64 kernels is
// quoted against, measured on the box the bench runs on ("FP64 peak is not in
// MEASURED_PEAKS: measure a dependent-FMA microbenchmark").
//
//   dn_probe_fp64 : every thread runs 16 independent DFMA chains; the kernel reports, per CTA, the SM clock cycles
//                   its loop took, so that bench.py can state DFMA / clock / SM as well as TFLOP/s (CUDA events).
//   dn_probe_lds  : 128-bit shared-memory loads with a chosen lane -> address pattern, cycles per warp-level load
//                   (= shared-memory wavefronts per request): what decides how the Gram tiles are laid over the lanes.
#include "common.cuh"

namespace {

constexpr int PROBE_CHAINS = 16;

__global__ void __launch_bounds__(256) probe_fp64_kernel(int iters, const double *seed, double *sink, long long *cycles) {
    double acc[PROBE_CHAINS];
    const double a = seed[0], b = seed[1];
#pragma unroll
    for (int k = 0; k < PROBE_CHAINS; ++k) acc[k] = (double)(threadIdx.x + k);
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int k = 0; k < PROBE_CHAINS; ++k) acc[k] = fma(acc[k], a, b);
    }
    __syncthreads();
    const long long t1 = clock64();
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < PROBE_CHAINS; ++k) s += acc[k];
    if (s == 123.456) sink[0] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

__global__ void __launch_bounds__(256) probe_lds_kernel(int pattern, int iters, double *sink, long long *cycles) {
    extern __shared__ double sh[];
    for (int e = threadIdx.x; e < 4096; e += blockDim.x) sh[e] = (double)e;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    int off;          // in 16-byte units
    switch (pattern) {
        case 0: off = lane; break;                       // 32 distinct, conflict free (512 B per request)
        case 1: off = 0; break;                          // all lanes one address
        case 2: off = lane >> 3; break;                  // each quarter-warp one address (4 distinct)
        case 3: off = lane & 7; break;                   // every quarter-warp reads the same 128 B
        case 4: off = lane & 3; break;                   // 4 distinct, interleaved inside each quarter
        case 5: off = (lane >> 2); break;                // 8 distinct, 4 neighbouring lanes share
        case 6: off = (lane & 15); break;                // 16 distinct (256 B), halves identical
        case 7: off = (lane >> 1); break;                // 16 distinct, lane pairs share
        case 8: off = (lane >> 4) * 8; break;            // 2 distinct, one per half-warp
        case 9: off = lane * 5; break;                   // stride 80 B (the P + 2 column stride of P = 8)
        default: {
            // patterns 10 .. 13: operand fetches of a 48 x 48 Gram triangle cut into 30 tiles of 8 rows x 6 columns,
            // tiles numbered row-major with rows of odd length padded so that aligned lane pairs share the row block
            int rb = 0, cb = 0, slot = 0;
            bool found = false;
            for (int r = 0; r < 6 && !found; ++r) {
                const int c_lo = (8 * r - 5 + 5) / 6;                  // first column block with 6 cb + 5 >= 8 r
                const int cnt = 8 - c_lo;
                for (int c = 0; c < cnt + (cnt & 1); ++c, ++slot)
                    if (slot == lane) { rb = r; cb = c_lo + (c < cnt ? c : cnt - 1); found = true; }
            }
            if (pattern == 10) off = 4 * rb;                                   // row operands, 64-byte blocks
            else if (pattern == 11) off = 4 * rb + ((rb >> 1) & 3);            // the same, fetch order rotated
            else if (pattern == 12) off = 3 * cb;                              // column operands, 48-byte blocks
            else off = 3 * lane;                                               // 32 distinct 48-byte blocks
            break;
        }
    }
    const unsigned base = (unsigned)__cvta_generic_to_shared(sh) + 16u * (unsigned)off;
    unsigned s0 = 0u;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            unsigned x, y, z, w;      // (one cheap integer op per load keeps the ALU far from being the limit)
            asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];\n"
                         : "=r"(x), "=r"(y), "=r"(z), "=r"(w)
                         : "r"(base + 2048u * (unsigned)k));
            s0 ^= x ^ y ^ z ^ w;
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (s0 == 0x12345u) sink[0] = 1.0;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

}  // namespace

extern "C" {

// Launches `ctas` CTAs of 256 threads, each thread 64 * iters DFMAs in 16 independent chains.  seed: 2 doubles
// (multiplier, addend; e.g. 1.0000001, 1e-9), sink: 1 double, cycles: `ctas` int64 (SM clocks of each CTA's loop).
// Returns the number of DFMAs issued in total (negative dn_status on error).
int64_t dn_probe_fp64(int32_t ctas, int32_t iters, const double *seed, double *sink, int64_t *cycles, void *stream) {
    if (ctas < 1 || iters < 1 || !seed || !sink || !cycles) return dn_fail(DN_ERR_INVALID, "bad probe argument%s");
    probe_fp64_kernel<<<ctas, 256, 0, (cudaStream_t)stream>>>(iters, seed, sink, (long long *)cycles);
    if (cudaGetLastError() != cudaSuccess) return dn_fail(DN_ERR_CUDA, "probe launch failed%s");
    return (int64_t)ctas * 256 * 4 * PROBE_CHAINS * iters;
}

// One CTA of 256 threads per launched block; every warp issues 8 * iters 128-bit shared loads with the lane -> address
// `pattern` (see the kernel).  Returns the warp-level loads issued per CTA.
int64_t dn_probe_lds(int32_t ctas, int32_t pattern, int32_t iters, double *sink, int64_t *cycles, void *stream) {
    if (ctas < 1 || iters < 1 || !sink || !cycles) return dn_fail(DN_ERR_INVALID, "bad probe argument%s");
    probe_lds_kernel<<<ctas, 256, 4096 * 8, (cudaStream_t)stream>>>(pattern, iters, sink, (long long *)cycles);
    if (cudaGetLastError() != cudaSuccess) return dn_fail(DN_ERR_CUDA, "probe launch failed%s");
    return (int64_t)8 * 8 * iters;
}

}  // extern "C"
