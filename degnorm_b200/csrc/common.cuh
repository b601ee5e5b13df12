// degnorm_b200 -- pieces shared by the kernel translation units (device primitives, kernel arguments).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>

#include "degnorm_b200.h"

// host-side error reporting (defined in abi.cu); message kept per thread for dn_last_error()
int dn_fail(int code, const char *fmt, const char *a = "", long long b = 0, long long c = 0);

#define DN_CUDA(call)                                                                                      \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess) return dn_fail(DN_ERR_CUDA, "%s (line %lld)", cudaGetErrorString(e_), __LINE__); \
    } while (0)

constexpr double EIG_TOL = 1.0e-14;
constexpr int EIG_FAST_STEPS = 64;     // power steps on G before the squaring fallback takes over
constexpr int EIG_MAX_SQUARINGS = 64;
constexpr double EIG_SUSPECT = 1.0e-3;  // eigenvector entry below this fraction of the largest, on a covered sample:
                                        // distrust the warm start (its overlap with a new top eigenvector may be lost)
constexpr int N_SMALL = 10;        // PP-sized shared vectors
constexpr int MODE_INIT = 0;       // ratio_svd on raw coverage (nmf.py:109-121)
constexpr int MODE_BS = 1;         // baseline_selection (nmf.py:189-372)

struct KArgs {
    const double *cov;
    const long long *off;
    const int *order;
    int n_work;
    int p, pp;
    const double *scale;
    const int *ds_start;
    int mode;
    int nmf_iter;
    double c;
    int bins, min_bins, min_hi, rate, skip, min_len;
    int flags;              // DN_FLAG_* (single-matrix helper methods)
    double *rho;
    unsigned char *ran;
    int *counters;
    double *kfac;
    double *e_first;
    double *est_rowsum;
    double *cov_rowsum;
    int resident_cols, ld_res;
    int ch, ldm, ks;
    int ms_doubles;         // scratch: M tile (tiled Gram) or reduction scratch (register Gram)
    int g_in_smem;
    double *ws;             // per-CTA slabs
    long long ws_stride;    // doubles per CTA slab
    long long ws_ld;        // row stride (columns) of the slab arrays
    int *queue;
    const double *row_max;  // n x p row maxima of the raw coverage (from the init pass) or NULL
    double *row_max_out;    // init pass: where to write them (or NULL)
    int eig_hint;           // small path: adaptive blind power steps on/off
    int eig_shared;         // small path, multi-warp CTAs: warp 0 solves alone instead of every warp redundantly
    double *est;            // fused estimates of the last outer iteration (or NULL): output buffer ...
    const long long *est_off;  // ... and the column offset of every gene's block in it (NULL: same as off)
    int nsets;              // tiled path: ceil(Gram tiles / threads); > 1 parks accumulators in global scratch
    long long gacc_doubles; // tiled path: size of that scratch per CTA
};

namespace {

// ---- warp / block primitives (fixed reduction trees: results are run-to-run deterministic) ------------------
__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}
__device__ __forceinline__ double warp_max(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = fmax(x, __shfl_xor_sync(0xffffffffu, x, o));
    return x;
}
template <int NT>
__device__ __forceinline__ double block_sum(double x, double *red) {
    x = warp_sum(x);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = x;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) s += red[w];
    __syncthreads();
    return s;
}
template <int NT>
__device__ __forceinline__ double block_max(double x, double *red) {
    x = warp_max(x);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = x;
    __syncthreads();
    double s = red[0];
#pragma unroll
    for (int w = 1; w < NT / 32; ++w) s = fmax(s, red[w]);
    __syncthreads();
    return s;
}
template <int NT>
__device__ __forceinline__ int block_sum_int(int x, int *red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = x;
    __syncthreads();
    int s = 0;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) s += red[w];
    __syncthreads();
    return s;
}

// Block version: any p <= NT, G anywhere (global for p > 64).  Called by all threads.
template <int NT>
__device__ int eig_block(const double *G, int pp, int p, double *v, double *red, bool cold, int max_steps, double tol,
                         bool bail, int *conv) {
    const int i = threadIdx.x;
    double vi = 0.0;
    if (cold) {
        double s = 0.0;
        if (i < p)
            for (int k = 0; k < p; ++k) s += G[(long long)k * pp + i];
        double n2 = block_sum<NT>(s * s, red);
        double inv = n2 > 0.0 ? 1.0 / sqrt(n2) : 0.0;
        vi = s * inv;
        if (i < pp) v[i] = vi;
        __syncthreads();
    } else {
        if (i < pp) vi = v[i];
    }
    int steps = 0;
    int ok = 0;
    double prev = 1.0e300;
    for (; steps < max_steps;) {
        double y = 0.0;
        if (i < p)
            for (int k = 0; k < p; ++k) y = fma(G[(long long)k * pp + i], v[k], y);
        ++steps;
        double n2 = block_sum<NT>(y * y, red);
        if (!(n2 > 0.0)) {
            if (i < pp) v[i] = 0.0;
            __syncthreads();
            ok = 1;
            break;
        }
        double w = y * (1.0 / sqrt(n2));
        double d = block_max<NT>(fabs(w - vi), red);
        vi = w;
        if (i < pp) v[i] = vi;
        __syncthreads();
        if (d <= tol) { ok = 1; break; }
        if (bail && steps >= 8 && d > 0.75 * prev) break;
        prev = d;
    }
    if (ok == 1 && bail) {
        const double m = (i < p && G[(long long)i * pp + i] > 0.0) ? vi : 1.0;
        const double vmin = -block_max<NT>(-m, red);
        const double vmax = block_max<NT>(vi, red);
        if (vmin < EIG_SUSPECT * vmax) ok = 2;
    }
    *conv = ok;
    return steps;
}

// Small-gap fallback (any p): repeated squaring B <- B.B / trace(B.B) starting from B = G / trace(G) squares the
// eigenvalue ratio each round, one power step with B per round tracks convergence, two steps with G polish.
// B0/B1 are per-CTA global scratch (pp*pp doubles each).  Called by all threads.  ARPACK (the reference) resolves
// such gaps exactly because its Krylov space spans all p dimensions; plain power iteration would need ~1/(1-r) steps.
template <int NT>
__device__ int eig_squaring(const double *G, int pp, int p, double *v, double *red, double *B0, double *B1, bool restart) {
    const int tid = threadIdx.x;
    const int nn = pp * pp;
    double tr = 0.0;
    for (int i = tid; i < p; i += NT) tr += G[(long long)i * pp + i];
    tr = block_sum<NT>(tr, red);
    if (!(tr > 0.0)) return 0;
    const double itr = 1.0 / tr;
    for (int e = tid; e < nn; e += NT) B0[e] = G[e] * itr;
    // Always restart from the uniform vector over the samples that have coverage: it has a positive overlap with
    // the (non-negative) top eigenvector, whereas a warm start may have lost it (see eig_warp).
    (void)restart;
    {
        double one = (tid < p && G[(long long)tid * pp + tid] > 0.0) ? 1.0 : 0.0;
        const double cntp = block_sum<NT>(one, red);
        if (tid < pp) v[tid] = cntp > 0.0 ? one / sqrt(cntp) : 0.0;
    }
    __syncthreads();
    double *B = B0, *Bn = B1;
    int rounds = 0;
    for (; rounds < EIG_MAX_SQUARINGS;) {
        double t = 0.0;
        for (int e = tid; e < nn; e += NT) {
            const int i = e / pp, j = e - i * pp;
            double s = 0.0;
            if (i < p && j < p)
                for (int k = 0; k < p; ++k) s = fma(B[i * pp + k], B[k * pp + j], s);
            Bn[e] = s;
            if (i == j) t += s;
        }
        t = block_sum<NT>(t, red);                 // also orders the Bn writes before the reads below
        ++rounds;
        if (!(t > 0.0)) break;
        const double it = 1.0 / t;
        for (int e = tid; e < nn; e += NT) Bn[e] *= it;
        __syncthreads();
        double *sw = B; B = Bn; Bn = sw;
        // one power step with the squared matrix
        double y = 0.0, vi = 0.0;
        if (tid < p) {
            vi = v[tid];
            for (int k = 0; k < p; ++k) y = fma(B[k * pp + tid], v[k], y);
        }
        const double n2 = block_sum<NT>(y * y, red);
        if (!(n2 > 0.0)) break;
        const double w = y * (1.0 / sqrt(n2));
        const double d = block_max<NT>(tid < p ? fabs(w - vi) : 0.0, red);
        if (tid < p) v[tid] = w;
        __syncthreads();
        // trace(B.B) with trace(B) = 1 reaches 1 exactly when B is numerically rank one
        if (1.0 - t <= 1.0e-15 && d <= EIG_TOL) break;
    }
    int conv;
    rounds += eig_block<NT>(G, pp, p, v, red, false, 2, 0.0, false, &conv);      // polish with G itself
    return rounds;
}
// |K| with entries < 1e-5 replaced by the smallest entry >= 1e-5 (nmf.py:329-330, 361-362).  dst may alias src.
__device__ void floor_abs(const double *src, double *dst, int p) {
    if (threadIdx.x == 0) {
        double mn = 1.0e300;
        for (int i = 0; i < p; ++i) {
            const double k = fabs(src[i]);
            if (k >= 1.0e-5 && k < mn) mn = k;
        }
        for (int i = 0; i < p; ++i) {
            const double k = fabs(src[i]);
            dst[i] = k < 1.0e-5 ? mn : k;     // mn stays 1e300 if no entry qualifies (the reference raises there)
        }
    }
    __syncthreads();
}

// numpy median of 1 - rho over p entries (nmf.py:257): > 1 ?
__device__ double median_one_minus(const double *rho, int p) {
    // small p: selection by rank counting, done by every thread identically (p <= 128)
    double lo = 0.0, hi = 0.0;
    const int k_lo = (p - 1) / 2, k_hi = p / 2;
    for (int i = 0; i < p; ++i) {
        const double ai = 1.0 - rho[i];
        int rank = 0;
        for (int j = 0; j < p; ++j) {
            const double aj = 1.0 - rho[j];
            rank += (aj < ai) || (aj == ai && j < i);
        }
        if (rank == k_lo) lo = ai;
        if (rank == k_hi) hi = ai;
    }
    return 0.5 * (lo + hi);
}

// Full-length estimate of one gene for the last outer iteration (nmf.py:217, 247, 333-337, 343-344, 350-351,
// 358-365).  Called by every thread that shares the gene: t0 = this thread's index among them, nt = how many.
// K = |K| floored as the reference does (the kfac output); E0 = E of the first fit (used when no column was filtered).
__device__ __forceinline__ void write_estimate(const double *F, int L, int p, const double *scale, int ex, int n0,
                                               const double *K, const double *E0, double *out, int t0, int nt) {
    if (ex == DN_EXIT_FEW_HICOV || ex == DN_EXIT_EMPTY_SAMPLE || ex == DN_EXIT_MEDIAN || ex < 0) {
        for (int i = 0; i < p; ++i) {
            const double s = scale[i];
            for (int j = t0; j < L; j += nt) out[(long long)i * L + j] = F[(long long)i * L + j] / s;
        }
    } else if (n0 < L) {
        for (int j = t0; j < L; j += nt) {
            double e = -1.0e300;
            for (int i = 0; i < p; ++i) e = fmax(e, (F[(long long)i * L + j] / scale[i]) / K[i]);
            for (int i = 0; i < p; ++i) {
                const double x = F[(long long)i * L + j] / scale[i];
                const double ke = K[i] * e;
                out[(long long)i * L + j] = ke < x ? x : ke;
            }
        }
    } else {
        for (int j = t0; j < L; j += nt) {
            if (ex == DN_EXIT_REFINED) {
                double e = -1.0e300;
                for (int i = 0; i < p; ++i) e = fmax(e, (F[(long long)i * L + j] / scale[i]) / K[i]);
                for (int i = 0; i < p; ++i) out[(long long)i * L + j] = K[i] * e;
            } else {
                const double e = __ldcg(E0 + j);
                for (int i = 0; i < p; ++i) {
                    double ke = K[i] * e;
                    if (ex != DN_EXIT_NO_SELECTION) {
                        const double x = F[(long long)i * L + j] / scale[i];
                        ke = ke < x ? x : ke;
                    }
                    out[(long long)i * L + j] = ke;
                }
            }
        }
    }
}

}  // namespace
