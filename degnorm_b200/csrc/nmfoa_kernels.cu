// degnorm_b200 -- fused NMF-OA / baseline-selection kernels for sm_100a (B200).
//
// One persistent CTA owns one gene at a time (genes are pulled longest-first from an atomic work queue) and
// runs the reference's whole per-gene flow for one outer DegNorm iteration without leaving the SM:
//
//   scale-on-load -> high-coverage filter (+ systematic down-sample) -> compaction of the kept columns into
//   shared memory (resident tier) or a per-CTA global slab (streamed tier) -> nmf() = 1 + nmf_iter passes of
//   {multiplier update, p x p Gram accumulate, top-eigenvector solve} -> DI -> bin-drop loop (<= bins-min_bins
//   more nmf() calls on the alive bins) -> envelope refine / fallbacks -> clipped DI row.
//
// What the passes restate (reference: /root/reference/degnorm/nmf.py, cited per function below):
//   rank_one_approx (nmf.py:55-64, scipy svds k=1) is replaced by: v = top eigenvector of the p x p Gram
//   matrix G = M M^T (M = x + lambda), found by warm-started power iteration ON G (p x p, in shared memory) to
//   |dv|_inf <= 1e-14.  Then K E = v (v^T M) exactly as the SVD gives, K = v*sigma, sigma^2 = sum_j (v^T M_j)^2.
//   Sign convention: M >= 0 so G >= 0 and the Perron vector is taken non-negative (the reference's K, E signs
//   are arbitrary and only K.E and |K| are used downstream).
//
// Gram accumulation is a register-tiled SYRK out of a shared-memory tile of M: thread (tile, kslice) owns a
// TR x TR block of G's upper triangle and a slice of the columns.  No tensor cores: rank-1, fp64.
//
// Tiers: a gene whose kept columns fit `resident_cols` keeps x and lambda in shared memory for the whole call
// sequence (HBM sees the raw coverage ~4 times per outer iteration); otherwise x and lambda live in a per-CTA
// global slab and every pass streams them (L2-resident when the slabs in flight fit the 126 MB L2).

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>

#include "degnorm_b200.h"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, const char *a = "", long long b = 0, long long c = 0) {
    snprintf(g_err, sizeof(g_err), fmt, a, b, c);
    return code;
}

#define DN_CUDA(call)                                                                                   \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) return fail(DN_ERR_CUDA, "%s (line %lld)", cudaGetErrorString(e_), __LINE__); \
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
};

// ---- shared-memory carve-up, shared by the host planner and the kernel --------------------------------------
struct Carve {
    long long small, red, binm, alive, ibuf, G, ms, xr, lr, resb, tb, total;   // offsets in doubles
};

__host__ __device__ inline Carve carve(int p, int pp, int g_in_smem, long long ms_doubles, int resident_cols, int ld_res) {
    Carve c;
    long long o = 0;
    c.small = o; o += (long long)N_SMALL * pp;
    c.red = o;   o += 64;
    c.binm = o;  o += DN_MAX_BINS;
    c.alive = o; o += DN_MAX_BINS / 2;
    c.ibuf = o;  o += 16;
    c.G = o;     o += g_in_smem ? (long long)pp * pp : 0;
    c.ms = o;    o += ms_doubles;
    c.xr = o;    o += resident_cols > 0 ? (long long)p * ld_res : 0;
    c.lr = o;    o += resident_cols > 0 ? (long long)p * ld_res : 0;
    c.resb = o;  o += resident_cols > 0 ? ld_res : 0;
    c.tb = o;    o += resident_cols > 0 ? ld_res : 0;
    c.total = o;
    return c;
}

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

// ---- per-gene state held in registers (uniform across the CTA) + shared pointers -----------------------------
struct Gene {
    // shared arrays
    double *v, *K, *K0, *rs0, *rsF, *rsC, *rsC0, *rho, *scale, *tmp, *red, *binm, *G, *ms;
    int *alive, *ibuf;
    // column storage of the current gene (shared or global)
    double *X, *Lm, *resb, *tb;
    long long ld;
    // current column set
    int n0;          // columns after the filters (width of F_start, nmf.py:237)
    int n_cur;       // columns of F_bin right now
    int cs;          // bin width ceil(n0/bins)
    int nb0;         // bins at the start
    int nalive;
    int eig_steps;
    int eig_fallbacks;
    double *B0;      // 2 * pp * pp doubles of global scratch for the small-gap eigen fallback
};

__device__ __forceinline__ int phys_col(const Gene &g, int vc) {
    if (g.nalive == g.nb0) return vc;
    int k = vc / g.cs;
    return g.alive[k] * g.cs + (vc - k * g.cs);
}

// ---- top eigenvector of G (p x p, symmetric, non-negative) by power iteration ---------------------------------
// Warp version: G in shared memory, p <= 64 (two rows per lane).  Called by warp 0 only.
__device__ int eig_warp(const double *G, int pp, int p, double *v, bool cold, int *conv) {
    const int lane = threadIdx.x & 31;
    const int r0 = lane, r1 = lane + 32;
    double v0 = 0.0, v1 = 0.0;
    if (cold) {
        // start from G.1 (row sums): positive for non-negative G, close to the Perron vector for near-rank-1 data
        double s0 = 0.0, s1 = 0.0;
        for (int k = 0; k < p; ++k) {
            if (r0 < p) s0 += G[k * pp + r0];
            if (r1 < p) s1 += G[k * pp + r1];
        }
        double n2 = warp_sum(s0 * s0 + s1 * s1);
        double inv = n2 > 0.0 ? 1.0 / sqrt(n2) : 0.0;
        v0 = s0 * inv;
        v1 = s1 * inv;
        __syncwarp();
        if (r0 < pp) v[r0] = v0;
        if (r1 < pp) v[r1] = v1;
        __syncwarp();
    } else {
        if (r0 < pp) v0 = v[r0];
        if (r1 < pp) v1 = v[r1];
    }
    int steps = 0;
    int ok = 0;
    double prev = 1.0e300;
    for (; steps < EIG_FAST_STEPS;) {
        double y0 = 0.0, y1 = 0.0;
        if (p <= 32) {
            if (r0 < p) {
                for (int k = 0; k < p; ++k) y0 = fma(G[k * pp + r0], v[k], y0);
            }
        } else {
            for (int k = 0; k < p; ++k) {
                double vk = v[k];
                if (r0 < p) y0 = fma(G[k * pp + r0], vk, y0);
                if (r1 < p) y1 = fma(G[k * pp + r1], vk, y1);
            }
        }
        ++steps;
        double n2 = warp_sum(y0 * y0 + y1 * y1);
        if (!(n2 > 0.0)) {            // all-zero matrix: the reference raises ArpackError here
            v0 = v1 = 0.0;
            __syncwarp();
            if (r0 < pp) v[r0] = 0.0;
            if (r1 < pp) v[r1] = 0.0;
            __syncwarp();
            ok = 1;
            break;
        }
        double inv = 1.0 / sqrt(n2);
        double w0 = y0 * inv, w1 = y1 * inv;
        double d = warp_max(fmax(fabs(w0 - v0), fabs(w1 - v1)));
        v0 = w0;
        v1 = w1;
        __syncwarp();
        if (r0 < pp) v[r0] = v0;
        if (r1 < pp) v[r1] = v1;
        __syncwarp();
        if (d <= EIG_TOL) { ok = 1; break; }
        if (steps >= 8 && d > 0.75 * prev) break;     // small spectral gap: let the squaring solver finish
        prev = d;
    }
    if (ok == 1) {
        // A warm start that has lost (underflowed) its component along the true top eigenvector can never regain
        // it: after an eigenvalue crossing between weakly coupled sample blocks power iteration would "converge"
        // to the wrong vector.  Entries ~0 on samples that do have coverage are the signature: re-solve robustly.
        const double m0 = (r0 < p && G[r0 * pp + r0] > 0.0) ? v0 : 1.0;
        const double m1 = (r1 < p && G[r1 * pp + r1] > 0.0) ? v1 : 1.0;
        const double vmin = -warp_max(-fmin(m0, m1));
        const double vmax = warp_max(fmax(v0, v1));
        if (vmin < EIG_SUSPECT * vmax) ok = 2;
    }
    if (lane == 0) *conv = ok;
    return steps;
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

// ---- one pass over the current columns: (optional multiplier update) + Gram accumulate -------------------------
// UPDATE=false: G = x x^T (first rank-one fit of nmf(), nmf.py:88).
// UPDATE=true : lambda <- max(0, lambda - c (K E - x)), M = x + lambda, G = M M^T (nmf.py:93-98),
//               with K E = v (v . M_old) per column.
template <int TR, int NT, bool UPDATE>
__device__ void gram_pass(const KArgs &a, Gene &g, int ti, int tj, int ks, bool tile_ok) {
    const int tid = threadIdx.x;
    const int p = a.p, pp = a.pp, ldm = a.ldm, CH = a.ch, KS = a.ks;
    double acc[TR][TR];
#pragma unroll
    for (int r = 0; r < TR; ++r)
#pragma unroll
        for (int q = 0; q < TR; ++q) acc[r][q] = 0.0;

    for (int base = 0; base < g.n_cur; base += CH) {
        const int ncol = min(CH, g.n_cur - base);
        // phase A: one thread per column
        if (tid < ncol) {
            const int pc = phys_col(g, base + tid);
            const double *xc = g.X + pc;
            if (!UPDATE) {
                for (int i = 0; i < p; ++i) g.ms[i * ldm + tid] = xc[i * g.ld];
            } else {
                double *lc = g.Lm + pc;
                double t = 0.0;
                for (int i = 0; i < p; ++i) t = fma(g.v[i], xc[i * g.ld] + lc[i * g.ld], t);
                for (int i = 0; i < p; ++i) {
                    const double x = xc[i * g.ld];
                    double l = lc[i * g.ld];
                    const double res = g.v[i] * t - x;        // est - x
                    l = l - a.c * res;
                    l = l < 0.0 ? 0.0 : l;
                    lc[i * g.ld] = l;
                    g.ms[i * ldm + tid] = x + l;
                }
            }
        }
        __syncthreads();
        // phase B: register-tiled SYRK out of the shared tile
        if (tile_ok) {
            const double *ma = g.ms + (ti * TR) * ldm;
            const double *mb = g.ms + (tj * TR) * ldm;
            for (int cidx = ks; cidx < ncol; cidx += KS) {
                double av[TR], bv[TR];
#pragma unroll
                for (int r = 0; r < TR; ++r) av[r] = ma[r * ldm + cidx];
#pragma unroll
                for (int r = 0; r < TR; ++r) bv[r] = mb[r * ldm + cidx];
#pragma unroll
                for (int r = 0; r < TR; ++r)
#pragma unroll
                    for (int q = 0; q < TR; ++q) acc[r][q] = fma(av[r], bv[q], acc[r][q]);
            }
        }
        __syncthreads();
    }
    // reduce the k-slices and mirror into the full square G (pp x pp)
    const int ntg = pp / TR;
    const int ntiles = ntg * (ntg + 1) / 2;
    if (KS == 1) {
        if (tile_ok) {
#pragma unroll
            for (int r = 0; r < TR; ++r)
#pragma unroll
                for (int q = 0; q < TR; ++q) {
                    const int i = ti * TR + r, j = tj * TR + q;
                    g.G[(long long)i * pp + j] = acc[r][q];
                    if (ti != tj) g.G[(long long)j * pp + i] = acc[r][q];
                }
        }
    } else {
        double *part = g.ms;      // the tile is free now (aliased)
        const int tile = tid - ks * ntiles;
        if (tile_ok) {
#pragma unroll
            for (int r = 0; r < TR; ++r)
#pragma unroll
                for (int q = 0; q < TR; ++q) part[((long long)ks * ntiles + tile) * (TR * TR) + r * TR + q] = acc[r][q];
        }
        __syncthreads();
        for (int e = tid; e < ntiles * TR * TR; e += NT) {
            double s = 0.0;
            for (int k = 0; k < KS; ++k) s += part[(long long)k * ntiles * TR * TR + e];
            const int tile_e = e / (TR * TR), rq = e - tile_e * (TR * TR);
            int t = tile_e, tii = 0;
            while (t >= ntg - tii) { t -= ntg - tii; ++tii; }
            const int tjj = tii + t;
            const int i = tii * TR + rq / TR, j = tjj * TR + rq % TR;
            g.G[(long long)i * pp + j] = s;
            if (tii != tjj) g.G[(long long)j * pp + i] = s;
        }
        if (pp > p) {             // the partials overwrote the tile's zero padding rows: restore them
            __syncthreads();
            for (int e = tid; e < (pp - p) * ldm; e += NT) g.ms[p * ldm + e] = 0.0;
        }
    }
    __syncthreads();
}

// ---- register-Gram pass for few samples (p <= P <= 12) ---------------------------------------------------------
// Each thread owns whole columns (c = tid, tid + NT, ...): it keeps the column's x and lambda in registers, does the
// multiplier update, and accumulates the column's contribution to all P(P+1)/2 Gram entries in registers -- no
// shared-memory operand traffic and exactly p(p+1)/2 FMAs per column.  The per-thread partial Grams are then
// reduced through a small shared scratch (16 entries x 32 lanes per round) and across warps.
// Scratch layout (g.ms): [NW][16*33] per-warp transpose buffers, then [NW][NE] per-warp sums.
constexpr int RED_ROUND = 16;
__host__ __device__ constexpr int reg_scratch_doubles(int P, int NT) {
    return (NT / 32) * (RED_ROUND * 33) + (NT / 32) * (P * (P + 1) / 2);
}

template <int P, int NT, bool UPDATE>
__device__ void gram_pass_reg(const KArgs &a, Gene &g) {
    constexpr int NE = P * (P + 1) / 2;
    constexpr int NW = NT / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int p = a.p;
    const long long ld = g.ld;
    double acc[NE];
#pragma unroll
    for (int e = 0; e < NE; ++e) acc[e] = 0.0;
    double vv[P];
#pragma unroll
    for (int i = 0; i < P; ++i) vv[i] = (UPDATE && i < p) ? g.v[i] : 0.0;
    const double c = a.c;
    for (int vc = tid; vc < g.n_cur; vc += NT) {
        const int pc = phys_col(g, vc);
        const double *xc = g.X + pc;
        double m[P];
        if (!UPDATE) {
#pragma unroll
            for (int i = 0; i < P; ++i) m[i] = i < p ? xc[i * ld] : 0.0;
        } else {
            double *lc = g.Lm + pc;
            double x[P], l[P];
#pragma unroll
            for (int i = 0; i < P; ++i) {
                x[i] = i < p ? xc[i * ld] : 0.0;
                l[i] = i < p ? lc[i * ld] : 0.0;
            }
            double t = 0.0;
#pragma unroll
            for (int i = 0; i < P; ++i) t = fma(vv[i], x[i] + l[i], t);
#pragma unroll
            for (int i = 0; i < P; ++i) {
                const double res = vv[i] * t - x[i];          // est - x
                double ln = l[i] - c * res;
                ln = ln < 0.0 ? 0.0 : ln;
                if (i < p) lc[i * ld] = ln;
                m[i] = x[i] + ln;
            }
        }
        int e = 0;
#pragma unroll
        for (int i = 0; i < P; ++i)
#pragma unroll
            for (int j = i; j < P; ++j) {
                acc[e] = fma(m[i], m[j], acc[e]);
                ++e;
            }
    }
    // warp-level transposed reduction, RED_ROUND entries per round; two lanes share one entry
    double *sc = g.ms + warp * (RED_ROUND * 33);
    double *part = g.ms + NW * (RED_ROUND * 33);
    const int k = lane >> 1, h = lane & 1;
#pragma unroll
    for (int r0 = 0; r0 < NE; r0 += RED_ROUND) {
#pragma unroll
        for (int q = 0; q < RED_ROUND; ++q)
            if (r0 + q < NE) sc[q * 33 + lane] = acc[r0 + q];
        __syncwarp();
        double s = 0.0;
        const double *row = sc + k * 33 + h * 16;
#pragma unroll
        for (int q = 0; q < 16; ++q) s += row[q];
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        if (h == 0 && r0 + k < NE) part[warp * NE + r0 + k] = s;
        __syncwarp();
    }
    __syncthreads();
    for (int e = tid; e < NE; e += NT) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) s += part[w * NE + e];
        int i = 0, t = e;
        while (t >= P - i) { t -= P - i; ++i; }
        const int j = i + t;
        g.G[i * P + j] = s;
        g.G[j * P + i] = s;
    }
    __syncthreads();
}

// Power iteration for P <= 16 with the matrix row in registers (G is P x P in shared memory, row stride P).
template <int P>
__device__ int eig_warp_small(const double *G, int p, double *v, bool cold, int *conv) {
    const int lane = threadIdx.x & 31;
    const bool row_ok = lane < P;
    double grow[P];
#pragma unroll
    for (int k = 0; k < P; ++k) grow[k] = row_ok ? G[lane * P + k] : 0.0;
    double vi = 0.0;
    if (cold) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < P; ++k) s += grow[k];
        double n2 = s * s;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, o);
        n2 = __shfl_sync(0xffffffffu, n2, 0);
        vi = n2 > 0.0 ? s * rsqrt(n2) : 0.0;
        __syncwarp();
        if (row_ok) v[lane] = vi;
        __syncwarp();
    } else {
        if (row_ok) vi = v[lane];
    }
    int steps = 0, ok = 0;
    double prev = 1.0e300;
    for (; steps < EIG_FAST_STEPS;) {
        double y = 0.0;
#pragma unroll
        for (int k = 0; k < P; ++k) y = fma(grow[k], v[k], y);
        ++steps;
        double n2 = y * y;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, o);
        n2 = __shfl_sync(0xffffffffu, n2, 0);
        if (!(n2 > 0.0)) {
            __syncwarp();
            if (row_ok) v[lane] = 0.0;
            __syncwarp();
            vi = 0.0;
            ok = 1;
            break;
        }
        const double w = y * rsqrt(n2);
        double d = fabs(w - vi);
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) d = fmax(d, __shfl_xor_sync(0xffffffffu, d, o));
        d = __shfl_sync(0xffffffffu, d, 0);
        vi = w;
        __syncwarp();
        if (row_ok) v[lane] = w;
        __syncwarp();
        if (d <= EIG_TOL) { ok = 1; break; }
        if (steps >= 8 && d > 0.75 * prev) break;
        prev = d;
    }
    if (ok == 1) {                                  // warm-start distrust rule, see eig_warp
        const double diag = row_ok ? G[lane * P + lane] : 0.0;
        double m = (lane < p && diag > 0.0) ? vi : 1.0;
        double mx = row_ok ? vi : 0.0;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            m = fmin(m, __shfl_xor_sync(0xffffffffu, m, o));
            mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        m = __shfl_sync(0xffffffffu, m, 0);
        mx = __shfl_sync(0xffffffffu, mx, 0);
        if (m < EIG_SUSPECT * mx) ok = 2;
    }
    if (lane == 0) *conv = ok;
    return steps;
}

template <int NT, int PREG>
__device__ __forceinline__ void eig_solve(const KArgs &a, Gene &g, bool cold) {
    int conv = 1;
    if (a.g_in_smem) {
        if (threadIdx.x < 32) {
            int s;
            if constexpr (PREG > 0) s = eig_warp_small<PREG>(g.G, a.p, g.v, cold, g.ibuf + 12);
            else s = eig_warp(g.G, a.pp, a.p, g.v, cold, g.ibuf + 12);
            g.eig_steps += s;
        }
        __syncthreads();
        conv = g.ibuf[12];
    } else {
        int s = eig_block<NT>(g.G, a.pp, a.p, g.v, g.red, cold, EIG_FAST_STEPS, 4.0 * EIG_TOL, true, &conv);
        g.eig_steps += s;
    }
    if (conv != 1) {                               // uniform across the CTA
        int s = eig_squaring<NT>(g.G, a.pp, a.p, g.v, g.red, g.B0, g.B0 + (long long)a.pp * a.pp, conv == 2);
        g.eig_steps += s;
        g.eig_fallbacks += 1;
    }
}

// ---- final pass of an nmf() call: everything the caller needs from K, E without materialising K.E ------------
// For the current columns, with t_j = v . (x_j + lambda_j):
//   sum_t  -> rs(K E)_i = v_i * sum_t      (nmf.py:247-254, 312-315)
//   sum_t2 -> sigma^2, K_i = v_i * sigma   (nmf.py:63-64)
//   res_j  = max_i ((KE_ij - x_ij)/(x_ij + 1))^2, KE clamped from below by x unless `first` (nmf.py:280-282, 318)
//   rsF_i  = sum_j x_ij ; rsC_i = sum_j max(KE_ij, x_ij)   (nmf.py:318-321, 343-345)
template <int NT>
__device__ void final_pass(const KArgs &a, Gene &g, bool first, bool have_lambda, bool want_res, double *e_first_g) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int p = a.p;
    double st = 0.0, st2 = 0.0;
    for (int vc = tid; vc < g.n_cur; vc += NT) {
        const int pc = phys_col(g, vc);
        const double *xc = g.X + pc;
        double t = 0.0;
        if (have_lambda) {
            const double *lc = g.Lm + pc;
            for (int i = 0; i < p; ++i) t = fma(g.v[i], xc[i * g.ld] + lc[i * g.ld], t);
        } else {
            for (int i = 0; i < p; ++i) t = fma(g.v[i], xc[i * g.ld], t);
        }
        g.tb[vc] = t;
        st += t;
        st2 = fma(t, t, st2);
        if (want_res) {
            double r = 0.0;
            for (int i = 0; i < p; ++i) {
                const double x = xc[i * g.ld];
                double ke = g.v[i] * t;
                if (!first) ke = ke < x ? x : ke;
                const double q = (ke - x) / (x + 1.0);
                r = fmax(r, q * q);
            }
            g.resb[vc] = r;
        }
    }
    const double sum_t = block_sum<NT>(st, g.red);
    const double sum_t2 = block_sum<NT>(st2, g.red);     // (the syncs inside also publish tb / resb)
    const double sigma = sqrt(sum_t2);
    if (e_first_g != nullptr) {                          // E of the first fit (only when no column was filtered)
        const double inv = sigma > 0.0 ? 1.0 / sigma : 0.0;
        for (int vc = tid; vc < g.n_cur; vc += NT) e_first_g[vc] = g.tb[vc] * inv;
    }
    // row-wise sums: one warp per sample
    for (int i = warp; i < p; i += NT / 32) {
        const double vi = g.v[i];
        double sF = 0.0, sC = 0.0;
        if (g.nalive == g.nb0) {
            const double *xr = g.X + (long long)i * g.ld;
            for (int vc = lane; vc < g.n_cur; vc += 32) {
                const double x = xr[vc];
                const double ke = vi * g.tb[vc];
                sF += x;
                sC += ke < x ? x : ke;
            }
        } else {
            for (int k = 0; k < g.nalive; ++k) {
                const int b = g.alive[k];
                const int lo = b * g.cs, w = min(g.cs, g.n0 - lo);
                const double *xr = g.X + (long long)i * g.ld + lo;
                const double *tr = g.tb + k * g.cs;
                for (int j = lane; j < w; j += 32) {
                    const double x = xr[j];
                    const double ke = vi * tr[j];
                    sF += x;
                    sC += ke < x ? x : ke;
                }
            }
        }
        sF = warp_sum(sF);
        sC = warp_sum(sC);
        if (lane == 0) {
            g.rsF[i] = sF;
            g.rsC[i] = sC;
            g.tmp[i] = vi * sum_t;        // rs(K E), unclamped
            g.K[i] = vi * sigma;          // K = u * s >= 0
        }
    }
    __syncthreads();
}

// nmf() on the current column set (nmf.py:78-107).  Leaves v, K, tmp=rs(KE), rsF, rsC, resb, tb.
template <int TR, int NT, int PREG>
__device__ void run_nmf(const KArgs &a, Gene &g, bool first, bool want_res, double *e_first_g, int ti, int tj, int ks,
                        bool tile_ok) {
    const int tid = threadIdx.x;
    const int T = a.nmf_iter;
    if (T > 0) {
        // lambda = 0 on the current columns (all of [0, n0): dead bins are never read)
        for (int i = 0; i < a.p; ++i)
            for (int j = tid; j < g.n0; j += NT) g.Lm[(long long)i * g.ld + j] = 0.0;
        __syncthreads();
    }
    if constexpr (PREG > 0) gram_pass_reg<PREG, NT, false>(a, g);
    else gram_pass<TR, NT, false>(a, g, ti, tj, ks, tile_ok);
    eig_solve<NT, PREG>(a, g, true);
    for (int it = 0; it < T; ++it) {
        if constexpr (PREG > 0) gram_pass_reg<PREG, NT, true>(a, g);
        else gram_pass<TR, NT, true>(a, g, ti, tj, ks, tile_ok);
        eig_solve<NT, PREG>(a, g, false);
    }
    final_pass<NT>(a, g, first, T > 0, want_res, e_first_g);
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

template <int TR, int NT, int PREG>
__global__ void __launch_bounds__(NT) nmfoa_kernel(const KArgs a) {
    extern __shared__ double smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int p = a.p, pp = a.pp;
    const Carve cv = carve(p, pp, a.g_in_smem, a.ms_doubles, a.resident_cols, a.ld_res);

    Gene g;
    double *sm = smem + cv.small;
    g.v = sm;            g.K = sm + pp;        g.K0 = sm + 2 * pp;   g.rs0 = sm + 3 * pp;  g.rsF = sm + 4 * pp;
    g.rsC = sm + 5 * pp; g.rsC0 = sm + 6 * pp; g.rho = sm + 7 * pp;  g.scale = sm + 8 * pp; g.tmp = sm + 9 * pp;
    g.red = smem + cv.red;
    g.binm = smem + cv.binm;
    g.alive = reinterpret_cast<int *>(smem + cv.alive);
    g.ibuf = reinterpret_cast<int *>(smem + cv.ibuf);
    g.ms = smem + cv.ms;
    double *slab = a.ws + (long long)blockIdx.x * a.ws_stride;
    g.B0 = slab;
    long long slab_o = 2ll * pp * pp;
    if (a.g_in_smem) {
        g.G = smem + cv.G;
    } else {
        g.G = slab + slab_o;
        slab_o += (long long)pp * pp;
    }
    g.eig_steps = 0;
    g.eig_fallbacks = 0;

    // Gram tile owned by this thread
    int ks = 0, ti = 0, tj = 0;
    bool tile_ok = false;
    if constexpr (PREG == 0) {
        const int ntg = pp / TR;
        const int ntiles = ntg * (ntg + 1) / 2;
        ks = tid / ntiles;
        tile_ok = ks < a.ks;
        int t = tid - ks * ntiles;
        while (t >= ntg - ti) { t -= ntg - ti; ++ti; }
        tj = ti + t;
    }
    // padded rows of the tile stay zero for the whole kernel
    for (int e = tid; e < a.ms_doubles; e += NT) g.ms[e] = 0.0;
    for (int e = tid; e < N_SMALL * pp; e += NT) sm[e] = 0.0;
    __syncthreads();

    for (;;) {
        if (tid == 0) g.ibuf[0] = atomicAdd(a.queue, 1);
        __syncthreads();
        const int w = g.ibuf[0];
        __syncthreads();
        if (w >= a.n_work) break;
        const int gid = a.order[w];
        const long long o0 = a.off[gid];
        const int L = (int)(a.off[gid + 1] - o0);
        const double *F = a.cov + (long long)p * o0;
        int *cnt = a.counters ? a.counters + (long long)gid * DN_NCOUNTERS : nullptr;
        g.eig_steps = 0;
        g.eig_fallbacks = 0;

        if (a.mode == MODE_INIT) {
            // ratio_svd on the raw matrix: all columns, no scaling, no multiplier updates
            g.X = const_cast<double *>(F);
            g.Lm = nullptr;
            g.ld = L;
            g.n0 = g.n_cur = L;
            g.cs = L; g.nb0 = 1; g.nalive = 1;
            // t buffer: resident if it fits, else slab
            const bool res_ok = L <= a.resident_cols;
            g.tb = res_ok ? smem + cv.tb : slab + slab_o;
            g.resb = nullptr;
            if (L >= 2) {
                run_nmf<TR, NT, PREG>(a, g, true, false, nullptr, ti, tj, ks, tile_ok);
                if (tid < p) {
                    a.est_rowsum[(long long)gid * p + tid] = g.rsC[tid];
                    a.cov_rowsum[(long long)gid * p + tid] = g.rsF[tid];
                }
            } else {
                // svds(k=1) is undefined for L < 2 (the reference raises); report est = cov
                if (tid < p) {
                    double s = 0.0;
                    for (int j = 0; j < L; ++j) s += F[(long long)tid * L + j];
                    a.est_rowsum[(long long)gid * p + tid] = s;
                    a.cov_rowsum[(long long)gid * p + tid] = s;
                }
            }
            if (cnt && tid == 0) {
                cnt[DN_CNT_EXIT] = 0; cnt[DN_CNT_N_HICOV] = L; cnt[DN_CNT_NMF_CALLS] = 1; cnt[DN_CNT_SUM_COLS] = L;
                cnt[DN_CNT_EIG_STEPS] = g.eig_steps; cnt[DN_CNT_DROPS_LO] = 0; cnt[DN_CNT_DROPS_HI] = 0;
                cnt[DN_CNT_RESIDENT] = (int)res_ok | (g.eig_fallbacks << 1);
            }
            __syncthreads();
            continue;
        }

        // ------------------------------------------------------------------ baseline_selection (nmf.py:189-372)
        if (tid < p) g.scale[tid] = a.scale[tid];
        __syncthreads();
        // (1) matrix max of the scaled coverage: max_j (F_ij / s_i) = (max_j F_ij) / s_i  (division is monotone)
        double tmax = -1.0e300;
        for (int i = 0; i < p; ++i) {
            const double *row = F + (long long)i * L;
            double m = -1.0e300;
            for (int j = tid; j < L; j += NT) m = fmax(m, row[j]);
            tmax = fmax(tmax, m / g.scale[i]);
        }
        const double gmax = block_max<NT>(tmax, g.red);
        const double thr = 0.1 * gmax;                                   // nmf.py:76
        // (2) count the kept columns: high coverage (strict >) and on the systematic sample (nmf.py:220-229)
        const int rate = a.rate;
        const int start = (rate > 1 && a.ds_start) ? a.ds_start[gid] : 0;
        const int ncand = start < L ? (L - start + rate - 1) / rate : 0;
        int mycount = 0;
        for (int k = tid; k < ncand; k += NT) {
            const long long col = start + (long long)k * rate;
            double cm = -1.0e300;
            for (int i = 0; i < p; ++i) cm = fmax(cm, F[(long long)i * L + col] / g.scale[i]);
            mycount += cm > thr;
        }
        const int n0 = block_sum_int<NT>(mycount, g.ibuf + 1);
        int exit_code = DN_EXIT_NONE;
        int ran = 0, nmf_calls = 0, sum_cols = 0;
        unsigned long long drops = 0ull;
        bool resident = false;
        bool k_is_refined = false;
        if (n0 < a.min_hi) {
            exit_code = DN_EXIT_FEW_HICOV;                               // nmf.py:232-233
        } else {
            // (3) compact the kept columns (scaled) into the working buffer
            resident = n0 <= a.resident_cols;
            if (resident) {
                g.X = smem + cv.xr; g.Lm = smem + cv.lr; g.resb = smem + cv.resb; g.tb = smem + cv.tb;
                g.ld = a.ld_res;
            } else {
                g.ld = a.ws_ld;
                g.X = slab + slab_o;
                g.Lm = g.X + (long long)p * g.ld;
                g.resb = g.Lm + (long long)p * g.ld;
                g.tb = g.resb + g.ld;
            }
            int running = 0;
            int *wcount = g.ibuf + 1;        // NT/32 ints
            for (int kb = 0; kb < ncand; kb += NT) {
                const int k = kb + tid;
                bool keep = false;
                long long col = 0;
                if (k < ncand) {
                    col = start + (long long)k * rate;
                    double cm = -1.0e300;
                    for (int i = 0; i < p; ++i) cm = fmax(cm, F[(long long)i * L + col] / g.scale[i]);
                    keep = cm > thr;
                }
                const unsigned bal = __ballot_sync(0xffffffffu, keep);
                if (lane == 0) wcount[warp] = __popc(bal);
                __syncthreads();
                int pre = running;
                int tot = 0;
                for (int q = 0; q < NT / 32; ++q) {
                    const int cq = wcount[q];
                    if (q < warp) pre += cq;
                    tot += cq;
                }
                if (keep) {
                    const int dst = pre + __popc(bal & ((1u << lane) - 1u));
                    for (int i = 0; i < p; ++i) g.X[(long long)i * g.ld + dst] = F[(long long)i * L + col] / g.scale[i];
                }
                running += tot;
                __syncthreads();
            }
            g.n0 = g.n_cur = n0;
            g.cs = n0; g.nb0 = 1; g.nalive = 1;          // no bins yet: identity column map
            // rs(F_start)
            for (int i = warp; i < p; i += NT / 32) {
                const double *xr = g.X + (long long)i * g.ld;
                double s = 0.0;
                for (int j = lane; j < n0; j += 32) s += xr[j];
                s = warp_sum(s);
                if (lane == 0) g.rs0[i] = s;
            }
            __syncthreads();
            bool any_empty = false;
            for (int i = 0; i < p; ++i) any_empty |= !(g.rs0[i] > 0.0);
            if (any_empty) {
                exit_code = DN_EXIT_EMPTY_SAMPLE;                        // nmf.py:241-242
            } else {
                const bool store_e = (a.e_first != nullptr) && (n0 == L);
                // (4) first fit (nmf.py:245-254)
                run_nmf<TR, NT, PREG>(a, g, true, true, store_e ? a.e_first + o0 : nullptr, ti, tj, ks, tile_ok);
                nmf_calls = 1; sum_cols = n0;
                if (tid < p) {
                    g.rho[tid] = 1.0 - g.rs0[tid] / (g.tmp[tid] + 1.0);
                    g.K0[tid] = g.K[tid];
                    g.rsC0[tid] = g.rsC[tid];
                }
                __syncthreads();
                if (median_one_minus(g.rho, p) > 1.0) {
                    exit_code = DN_EXIT_MEDIAN;                          // nmf.py:257-258
                } else {
                    double rmin = g.rho[0], rmax = g.rho[0];
                    for (int i = 1; i < p; ++i) { rmin = fmin(rmin, g.rho[i]); rmax = fmax(rmax, g.rho[i]); }
                    if (n0 >= a.min_len && rmin <= 0.2 && !a.skip) {     // nmf.py:265
                        g.cs = (n0 + a.bins - 1) / a.bins;               // utils.py:176-192
                        g.nb0 = (n0 + g.cs - 1) / g.cs;
                        g.nalive = g.nb0;
                        if (tid < g.nb0) g.alive[tid] = tid;
                        __syncthreads();
                        while (rmax > 0.1) {                             // nmf.py:273
                            ran = 1;
                            // mean squared-relative-residual per alive bin (nmf.py:280-283); one warp per bin
                            for (int k = warp; k < g.nalive; k += NT / 32) {
                                const int b = g.alive[k];
                                const int wdt = min(g.cs, g.n0 - b * g.cs);
                                const double *rr = g.resb + k * g.cs;
                                double s = 0.0;
                                for (int j = lane; j < wdt; j += 32) s += rr[j];
                                s = warp_sum(s);
                                if (lane == 0) g.binm[k] = s / (double)wdt;
                            }
                            __syncthreads();
                            int kd = 0;
                            double best = g.binm[0];
                            for (int k = 1; k < g.nalive; ++k)
                                if (g.binm[k] > best) { best = g.binm[k]; kd = k; }
                            if (best == 0.0) break;                      // nmf.py:286-287
                            const int bd = g.alive[kd];
                            const int wd = min(g.cs, g.n0 - bd * g.cs);
                            __syncthreads();
                            if (tid == 0)
                                for (int k = kd; k < g.nalive - 1; ++k) g.alive[k] = g.alive[k + 1];
                            __syncthreads();
                            g.nalive -= 1;
                            g.n_cur -= wd;
                            drops |= 1ull << bd;
                            if (g.n_cur < 2) break;                      // svds ValueError swallowed, nmf.py:306-310
                            run_nmf<TR, NT, PREG>(a, g, false, true, nullptr, ti, tj, ks, tile_ok);
                            nmf_calls += 1; sum_cols += g.n_cur;
                            double mn = g.tmp[0];
                            for (int i = 1; i < p; ++i) mn = fmin(mn, g.tmp[i]);
                            if (mn == 0.0) break;                        // nmf.py:315-316
                            __syncthreads();
                            if (tid < p) g.rho[tid] = 1.0 - g.rsF[tid] / (g.rsC[tid] + 1.0);   // nmf.py:318-321
                            __syncthreads();
                            rmax = g.rho[0];
                            for (int i = 1; i < p; ++i) rmax = fmax(rmax, g.rho[i]);
                            if (g.nalive <= a.min_bins || g.n_cur < a.min_len) break;        // nmf.py:323
                        }
                        __syncthreads();
                        bool fallback = true;
                        exit_code = DN_EXIT_FALLBACK;
                        if (rmax < 0.2) {                                // nmf.py:327-346
                            floor_abs(g.K, g.K, p);
                            double s = 0.0;
                            for (int j = tid; j < n0; j += NT) {
                                double e = -1.0e300;
                                for (int i = 0; i < p; ++i) e = fmax(e, g.X[(long long)i * g.ld + j] / g.K[i]);
                                s += e;
                            }
                            const double S = block_sum<NT>(s, g.red);
                            if (tid < p) g.rho[tid] = 1.0 - g.rs0[tid] / (g.K[tid] * S + 1.0);
                            __syncthreads();
                            rmax = g.rho[0];
                            for (int i = 1; i < p; ++i) rmax = fmax(rmax, g.rho[i]);
                            if (rmax > 0.9) {
                                exit_code = DN_EXIT_FALLBACK_HIGH;
                            } else {
                                exit_code = DN_EXIT_REFINED;
                                fallback = false;
                                k_is_refined = true;
                            }
                        }
                        if (fallback) {                                  // nmf.py:342-353
                            __syncthreads();
                            if (tid < p) g.rho[tid] = 1.0 - g.rs0[tid] / (g.rsC0[tid] + 1.0);
                            __syncthreads();
                        }
                    } else {
                        exit_code = DN_EXIT_NO_SELECTION;
                    }
                }
            }
        }
        // (5) outputs
        __syncthreads();
        const bool is_default = exit_code == DN_EXIT_FEW_HICOV || exit_code == DN_EXIT_EMPTY_SAMPLE ||
                                exit_code == DN_EXIT_MEDIAN;
        if (!is_default && !k_is_refined) {
            // K of the first fit; floored unless the estimate keeps the fit's own columns (n0 == L)
            if (n0 == L) {
                if (tid < p) g.K[tid] = g.K0[tid];
                __syncthreads();
            } else {
                floor_abs(g.K0, g.K, p);
            }
        }
        if (tid < p) {
            double r = is_default ? 0.0 : g.rho[tid];
            r = r > 0.9 ? 0.9 : r;                                       // nmf.py:398-399
            r = r < 0.0 ? 0.0 : r;
            a.rho[(long long)gid * p + tid] = r;
            if (a.kfac) a.kfac[(long long)gid * p + tid] = is_default ? 0.0 : g.K[tid];
        }
        if (tid == 0) {
            a.ran[gid] = (unsigned char)(is_default ? 0 : ran);
            if (cnt) {
                cnt[DN_CNT_EXIT] = exit_code; cnt[DN_CNT_N_HICOV] = n0; cnt[DN_CNT_NMF_CALLS] = nmf_calls;
                cnt[DN_CNT_SUM_COLS] = sum_cols; cnt[DN_CNT_EIG_STEPS] = g.eig_steps;
                cnt[DN_CNT_DROPS_LO] = (int)(drops & 0xffffffffull); cnt[DN_CNT_DROPS_HI] = (int)(drops >> 32);
                cnt[DN_CNT_RESIDENT] = (int)resident | (g.eig_fallbacks << 1);
            }
        }
        __syncthreads();
    }
}

// ---- estimates (last outer iteration only): nmf.py:217, 247, 333-337, 343-344, 350-351, 358-365 -----------------
__global__ void __launch_bounds__(256) estimates_kernel(const double *cov, const long long *off, const int *order,
                                                        int n_work, int p, const double *scale, const int *counters,
                                                        const double *kfac, const double *e_first, double *est) {
    for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
        const int gid = order[w];
        const long long o0 = off[gid];
        const int L = (int)(off[gid + 1] - o0);
        const double *F = cov + (long long)p * o0;
        double *out = est + (long long)p * o0;
        const int ex = counters[(long long)gid * DN_NCOUNTERS + DN_CNT_EXIT];
        const int n0 = counters[(long long)gid * DN_NCOUNTERS + DN_CNT_N_HICOV];
        const double *K = kfac + (long long)gid * p;
        if (ex == DN_EXIT_FEW_HICOV || ex == DN_EXIT_EMPTY_SAMPLE || ex == DN_EXIT_MEDIAN) {
            for (int i = 0; i < p; ++i) {
                const double s = scale[i];
                for (int j = threadIdx.x; j < L; j += blockDim.x) out[(long long)i * L + j] = F[(long long)i * L + j] / s;
            }
        } else if (n0 < L) {
            for (int j = threadIdx.x; j < L; j += blockDim.x) {
                double e = -1.0e300;
                for (int i = 0; i < p; ++i) e = fmax(e, (F[(long long)i * L + j] / scale[i]) / K[i]);
                for (int i = 0; i < p; ++i) {
                    const double x = F[(long long)i * L + j] / scale[i];
                    const double ke = K[i] * e;
                    out[(long long)i * L + j] = ke < x ? x : ke;
                }
            }
        } else {
            const double *E0 = e_first + o0;
            for (int j = threadIdx.x; j < L; j += blockDim.x) {
                if (ex == DN_EXIT_REFINED) {
                    double e = -1.0e300;
                    for (int i = 0; i < p; ++i) e = fmax(e, (F[(long long)i * L + j] / scale[i]) / K[i]);
                    for (int i = 0; i < p; ++i) out[(long long)i * L + j] = K[i] * e;
                } else {
                    const double e = E0[j];
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
}

// ---- n x p scalar updates --------------------------------------------------------------------------------------
constexpr int SLAB_ROWS = 64;

// mode 0: outer sums (nmf.py:575, 148-158).  mode 1: init sums (nmf.py:524-531).
__global__ void __launch_bounds__(256) sums_partial_kernel(int mode, const double *A, const double *B, const double *C,
                                                           int n, int p, double *rho0, double *partial) {
    __shared__ unsigned char flag[SLAB_ROWS];
    const int r0 = blockIdx.x * SLAB_ROWS;
    const int nr = min(SLAB_ROWS, n - r0);
    for (int r = threadIdx.x; r < nr; r += blockDim.x) {
        const long long o = (long long)(r0 + r) * p;
        double mx = -1.0e300;
        if (mode == 0) {
            for (int i = 0; i < p; ++i) mx = fmax(mx, B[o + i]);            // B = rho (clipped)
            flag[r] = mx == 0.0;                                             // non-baseline gene, nmf.py:155
        } else {
            for (int i = 0; i < p; ++i) {
                const double r0v = 1.0 - B[o + i] / (A[o + i] + 1.0);        // A = est_rowsum, B = cov_rowsum
                rho0[o + i] = r0v;
                mx = fmax(mx, r0v);
            }
            flag[r] = mx < 0.1;                                              // low-DI gene, nmf.py:528
        }
    }
    __syncthreads();
    double *out = partial + (long long)blockIdx.x * (3 * p + 1);
    for (int i = threadIdx.x; i < p; i += blockDim.x) {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        for (int r = 0; r < nr; ++r) {
            const long long o = (long long)(r0 + r) * p + i;
            if (mode == 0) {
                const double xw = A[o];                                      // A = x_weighted
                s0 += xw;
                if (flag[r]) s2 += xw; else s1 += xw / (1.0 - B[o]);
            } else {
                const double x = C[o];                                       // C = reads
                if (flag[r]) s0 += x;
                s1 += x;
            }
        }
        out[i] = s0; out[p + i] = s1; out[2 * p + i] = s2;
    }
    if (threadIdx.x == 0) {
        int c = 0;
        for (int r = 0; r < nr; ++r) c += flag[r];
        out[3 * p] = (double)c;
    }
}

__global__ void sums_final_kernel(const double *partial, int nblocks, int p, int nvec, double *sums) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nvec) {
        double s = 0.0;
        for (int b = 0; b < nblocks; ++b) s += partial[(long long)b * (3 * p + 1) + i];
        sums[i] = s;
    }
}

__device__ double median_of(const double *a, int p) {
    double lo = 0.0, hi = 0.0;
    const int k_lo = (p - 1) / 2, k_hi = p / 2;
    for (int i = 0; i < p; ++i) {
        int rank = 0;
        for (int j = 0; j < p; ++j) rank += (a[j] < a[i]) || (a[j] == a[i] && j < i);
        if (rank == k_lo) lo = a[i];
        if (rank == k_hi) hi = a[i];
    }
    return 0.5 * (lo + hi);
}

// nmf.py:148-158, 575-590
__global__ void __launch_bounds__(256) outer_apply_kernel(const double *sums, int n, int p, double *xw, double *rho,
                                                          double *x_adj, double *norm_factors, double *scale_factors) {
    extern __shared__ double sh[];
    double *avg = sh, *norm = sh + p, *col = sh + 2 * p;
    __shared__ unsigned char flag[SLAB_ROWS];
    for (int i = threadIdx.x; i < p; i += blockDim.x) {
        const double pre = sums[p + i] + sums[2 * p + i];          // colsum(x_adj) before the correction
        const double a = 1.0 - sums[i] / pre;                      // sample average DI
        avg[i] = a;
        col[i] = sums[p + i] + sums[2 * p + i] / (1.0 - a);        // colsum(x_adj) after the correction
    }
    __syncthreads();
    const double med = median_of(col, p);
    for (int i = threadIdx.x; i < p; i += blockDim.x) norm[i] = col[i] / med;
    __syncthreads();
    if (blockIdx.x == 0) {
        for (int i = threadIdx.x; i < p; i += blockDim.x) {
            norm_factors[i] = norm[i];
            scale_factors[i] *= norm[i];
        }
    }
    const int r0 = blockIdx.x * SLAB_ROWS;
    const int nr = min(SLAB_ROWS, n - r0);
    for (int r = threadIdx.x; r < nr; r += blockDim.x) {
        const long long o = (long long)(r0 + r) * p;
        double mx = -1.0e300;
        for (int i = 0; i < p; ++i) mx = fmax(mx, rho[o + i]);
        flag[r] = mx == 0.0;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < nr * p; e += blockDim.x) {
        const int r = e / p, i = e - r * p;
        const long long o = (long long)(r0 + r) * p + i;
        double rh = rho[o];
        if (flag[r]) { rh = avg[i]; rho[o] = rh; }
        const double w = xw[o];
        x_adj[o] = w / (1.0 - rh);
        xw[o] = w / norm[i];
    }
}

// nmf.py:529-535
__global__ void __launch_bounds__(256) init_apply_kernel(const double *sums, const double *reads, int n, int p,
                                                         double *xw, double *norm_factors, double *scale_factors) {
    extern __shared__ double sh[];
    double *cs = sh, *norm = sh + p;
    const bool any_low = sums[3 * p] > 0.0;
    for (int i = threadIdx.x; i < p; i += blockDim.x) cs[i] = any_low ? sums[i] : sums[p + i];
    __syncthreads();
    const double med = median_of(cs, p);
    for (int i = threadIdx.x; i < p; i += blockDim.x) norm[i] = cs[i] / med;
    __syncthreads();
    if (blockIdx.x == 0)
        for (int i = threadIdx.x; i < p; i += blockDim.x) { norm_factors[i] = norm[i]; scale_factors[i] = norm[i]; }
    const long long tot = (long long)n * p;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (long long)gridDim.x * blockDim.x)
        xw[e] = reads[e] / norm[e % p];
}

// ---- host side ------------------------------------------------------------------------------------------------
int tile_for_p(int p) { return p <= 64 ? 4 : 8; }

int check_params(const dn_params *prm) {
    if (!prm) return fail(DN_ERR_INVALID, "null params%s");
    if (prm->p < 2) return fail(DN_ERR_INVALID, "need at least 2 samples%s (p = %lld)", "", prm->p);
    if (prm->p > DN_MAX_SAMPLES) return fail(DN_ERR_UNSUPPORTED, "%sp = %lld exceeds DN_MAX_SAMPLES", "", prm->p);
    if (prm->bins < 1 || prm->bins > DN_MAX_BINS) return fail(DN_ERR_UNSUPPORTED, "%sbins = %lld outside [1, DN_MAX_BINS]", "", prm->bins);
    if (prm->downsample_rate < 1) return fail(DN_ERR_INVALID, "downsample_rate must be >= 1%s");
    if (prm->nmf_iter < 0) return fail(DN_ERR_INVALID, "nmf_iter must be >= 0%s");
    return DN_OK;
}

struct Derived { int tr, nt, preg, pp, ntiles, ch, ldm, ks, g_in_smem, ms_doubles; long long fixed_doubles; };

// Threads per CTA of the register-Gram path, by resident tier: small genes get one warp per gene (no block-wide
// barriers on the per-iteration critical path, many genes in flight per SM); larger tiers get more warps per gene
// so that the shared-memory-limited CTAs still fill the SM.
int threads_for_tier(int resident_cols) {
    if (resident_cols <= 0) return 256;
    if (resident_cols <= 128) return 32;
    if (resident_cols <= 256) return 64;
    if (resident_cols <= 512) return 128;
    return 256;
}

Derived derive(int p, int nt_reg) {
    Derived d;
    memset(&d, 0, sizeof(d));
    d.preg = p <= 4 ? 4 : (p <= 8 ? 8 : (p <= 12 ? 12 : 0));
    if (d.preg > 0) {
        d.tr = 0;
        d.nt = nt_reg;
        d.pp = d.preg;
        d.ntiles = 0;
        d.ch = 0; d.ldm = 0; d.ks = 1;
        d.g_in_smem = 1;
        d.ms_doubles = reg_scratch_doubles(d.preg, d.nt);
    } else {
        d.tr = tile_for_p(p);
        d.nt = 256;
        d.pp = (p + d.tr - 1) / d.tr * d.tr;
        const int ntg = d.pp / d.tr;
        d.ntiles = ntg * (ntg + 1) / 2;
        int ch = (48 * 1024 / 8) / d.pp / 32 * 32;
        if (ch < 32) ch = 32;
        if (ch > d.nt) ch = d.nt;
        d.ch = ch;
        d.ldm = ch + 1;
        int ks = d.nt / d.ntiles;
        const long long cap = (long long)d.pp * d.ldm / ((long long)d.ntiles * d.tr * d.tr);
        if (ks > cap) ks = (int)cap;
        if (ks < 1) ks = 1;
        d.ks = ks;
        d.g_in_smem = d.pp <= 64;
        d.ms_doubles = d.pp * d.ldm;
    }
    d.fixed_doubles = carve(p, d.pp, d.g_in_smem, d.ms_doubles, 0, 0).total;
    return d;
}

template <int TR, int NT, int PREG>
int launch(const KArgs &a, const dn_plan *plan, cudaStream_t st) {
    auto kern = nmfoa_kernel<TR, NT, PREG>;
    DN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, plan->smem_bytes));
    kern<<<plan->ctas, NT, plan->smem_bytes, st>>>(a);
    DN_CUDA(cudaGetLastError());
    return DN_OK;
}

template <int PREG>
int launch_reg(const KArgs &a, const dn_plan *plan, cudaStream_t st) {
    switch (plan->threads) {
        case 32: return launch<0, 32, PREG>(a, plan, st);
        case 64: return launch<0, 64, PREG>(a, plan, st);
        case 128: return launch<0, 128, PREG>(a, plan, st);
        case 256: return launch<0, 256, PREG>(a, plan, st);
    }
    return fail(DN_ERR_INVALID, "plan.threads must be 32, 64, 128 or 256%s");
}

int run_kernel(int mode, const double *cov, const int64_t *off, const int32_t *order, int32_t n_work,
               const dn_params *prm, const dn_plan *plan, const double *scale, const int32_t *ds_start, double *rho,
               uint8_t *ran, int32_t *counters, double *kfac, double *e_first, double *est_rowsum, double *cov_rowsum,
               void *workspace, int64_t workspace_bytes, void *stream) {
    int rc = check_params(prm);
    if (rc) return rc;
    if (!plan || !cov || !off || !order) return fail(DN_ERR_INVALID, "null pointer argument%s");
    if (n_work <= 0) return DN_OK;
    if (workspace_bytes < plan->ws_bytes || !workspace) return fail(DN_ERR_WORKSPACE, "workspace too small%s: need %lld, got %lld", "", plan->ws_bytes, workspace_bytes);
    const Derived d = derive(prm->p, plan->threads);
    if (plan->tile != d.tr || plan->threads != d.nt || plan->chunk_cols != d.ch)
        return fail(DN_ERR_INVALID, "plan does not match params (use dn_make_plan)%s");
    cudaStream_t st = (cudaStream_t)stream;
    KArgs a;
    memset(&a, 0, sizeof(a));
    a.cov = cov; a.off = (const long long *)off; a.order = order; a.n_work = n_work;
    a.p = prm->p; a.pp = d.pp; a.scale = scale; a.ds_start = ds_start; a.mode = mode;
    a.nmf_iter = mode == MODE_INIT ? 0 : prm->nmf_iter;
    a.c = prm->nmf_iter > 0 ? 1.0 / sqrt((double)prm->nmf_iter) : 0.0;       // nmf.py:91
    a.bins = prm->bins; a.min_bins = prm->min_bins; a.min_hi = prm->min_high_coverage;
    a.rate = mode == MODE_INIT ? 1 : prm->downsample_rate; a.skip = prm->skip_baseline_selection;
    a.min_len = prm->min_gene_len;
    a.rho = rho; a.ran = ran; a.counters = counters; a.kfac = kfac; a.e_first = e_first;
    a.est_rowsum = est_rowsum; a.cov_rowsum = cov_rowsum;
    a.resident_cols = plan->resident_cols;
    a.ld_res = plan->resident_cols;
    a.ch = d.ch; a.ldm = d.ldm; a.ks = d.ks; a.g_in_smem = d.g_in_smem; a.ms_doubles = d.ms_doubles;
    // workspace: [queue (256 B)] [per-CTA slabs]
    a.queue = (int *)workspace;
    a.ws = (double *)((char *)workspace + 256);
    a.ws_ld = plan->ws_cols;
    const long long g_d = (d.g_in_smem ? 2ll : 3ll) * d.pp * d.pp;
    const long long cols_d = mode == MODE_INIT ? plan->ws_cols : (2ll * prm->p + 2) * plan->ws_cols;
    a.ws_stride = (g_d + cols_d + 31) / 32 * 32;
    DN_CUDA(cudaMemsetAsync(a.queue, 0, 256, st));
    if (d.preg == 4) return launch_reg<4>(a, plan, st);
    if (d.preg == 8) return launch_reg<8>(a, plan, st);
    if (d.preg == 12) return launch_reg<12>(a, plan, st);
    if (d.tr == 4) return launch<4, 256, 0>(a, plan, st);
    return launch<8, 256, 0>(a, plan, st);
}

}  // namespace

extern "C" {

int dn_abi_version(void) { return DN_ABI_VERSION; }
const char *dn_last_error(void) { return g_err; }

int dn_device_info(int32_t *sm_count, int32_t *max_smem_optin, int32_t *cc) {
    int dev = 0;
    DN_CUDA(cudaGetDevice(&dev));
    int sm = 0, smem = 0, major = 0, minor = 0;
    DN_CUDA(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev));
    DN_CUDA(cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    DN_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    DN_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    if (sm_count) *sm_count = sm;
    if (max_smem_optin) *max_smem_optin = smem;
    if (cc) *cc = major * 10 + minor;
    return DN_OK;
}

int dn_make_plan(const dn_params *prm, int64_t max_cols, int32_t n_work, int32_t want_resident, int32_t for_init,
                 int32_t sm_count, int32_t max_smem_optin, dn_plan *plan) {
    int rc = check_params(prm);
    if (rc) return rc;
    if (!plan || max_cols < 1 || sm_count < 1 || max_smem_optin < 16 * 1024) return fail(DN_ERR_INVALID, "bad planning argument%s");
    // register-Gram path: the CTA size follows the resident tier that was asked for
    const int want_cols = want_resident < 0 ? (max_cols > 1 << 20 ? 1 << 20 : (int)max_cols)
                                             : (want_resident > max_cols ? (int)max_cols : want_resident);
    const Derived d = derive(prm->p, for_init ? 256 : threads_for_tier(want_cols));
    if (d.ntiles > d.nt) return fail(DN_ERR_UNSUPPORTED, "%sp = %lld needs more Gram tiles than threads", "", prm->p);
    memset(plan, 0, sizeof(*plan));
    plan->tile = d.tr;
    plan->threads = d.nt;
    plan->chunk_cols = d.ch;
    const long long fixed_b = d.fixed_doubles * 8;
    if (fixed_b > max_smem_optin) return fail(DN_ERR_UNSUPPORTED, "shared memory carve-up does not fit%s");
    const long long per_col = 8ll * (2 * prm->p + 2);
    long long fit = (max_smem_optin - fixed_b) / per_col;
    long long res = 0;
    if (want_resident != 0) {
        long long want = want_resident < 0 ? max_cols : (long long)want_resident;
        if (want > max_cols) want = max_cols;
        res = want < fit ? want : fit;
        res = res / 2 * 2;          // keep rows 16-byte aligned
        if (res < 2) res = 0;
    }
    plan->resident_cols = (int32_t)res;
    plan->smem_bytes = (int32_t)(fixed_b + res * per_col);
    plan->ws_cols = max_cols > res ? (max_cols + 7) / 8 * 8 : 0;
    // persistent CTAs: as many as the shared memory / thread budget of an SM allows
    int per_sm = (int)((228ll * 1024) / (plan->smem_bytes + 1024));
    if (per_sm > 2048 / d.nt) per_sm = 2048 / d.nt;
    if (per_sm > 32) per_sm = 32;
    if (d.preg == 12 && per_sm * d.nt > 256) per_sm = 256 / d.nt > 0 ? 256 / d.nt : 1;   // ~250 registers per thread
    if (per_sm < 1) per_sm = 1;
    long long ctas = (long long)sm_count * per_sm;
    if (ctas > n_work) ctas = n_work;
    if (ctas < 1) ctas = 1;
    plan->ctas = (int32_t)ctas;
    const long long g_d = (d.g_in_smem ? 2ll : 3ll) * d.pp * d.pp;
    const long long cols_d = for_init ? plan->ws_cols : (2ll * prm->p + 2) * plan->ws_cols;
    const long long stride = (g_d + cols_d + 31) / 32 * 32;
    plan->ws_bytes = 256 + ctas * stride * 8;
    return DN_OK;
}

int dn_init_ratio_svd(const double *cov, const int64_t *off, const int32_t *order, int32_t n_work, const dn_params *prm,
                      const dn_plan *plan, double *est_rowsum, double *cov_rowsum, int32_t *counters, void *workspace,
                      int64_t workspace_bytes, void *stream) {
    if (!est_rowsum || !cov_rowsum) return fail(DN_ERR_INVALID, "null output%s");
    return run_kernel(MODE_INIT, cov, off, order, n_work, prm, plan, nullptr, nullptr, nullptr, nullptr, counters,
                      nullptr, nullptr, est_rowsum, cov_rowsum, workspace, workspace_bytes, stream);
}

int dn_baseline_selection(const double *cov, const int64_t *off, const int32_t *order, int32_t n_work,
                          const dn_params *prm, const dn_plan *plan, const double *scale, const int32_t *ds_start,
                          double *rho, uint8_t *ran, int32_t *counters, double *kfac, double *e_first, void *workspace,
                          int64_t workspace_bytes, void *stream) {
    if (!scale || !rho || !ran) return fail(DN_ERR_INVALID, "null pointer argument%s");
    if (prm && prm->downsample_rate > 1 && !ds_start) return fail(DN_ERR_INVALID, "ds_start required when downsampling%s");
    return run_kernel(MODE_BS, cov, off, order, n_work, prm, plan, scale, ds_start, rho, ran, counters, kfac, e_first,
                      nullptr, nullptr, workspace, workspace_bytes, stream);
}

int dn_estimates(const double *cov, const int64_t *off, const int32_t *order, int32_t n_work, const dn_params *prm,
                 const double *scale, const int32_t *counters, const double *kfac, const double *e_first, double *est,
                 void *stream) {
    int rc = check_params(prm);
    if (rc) return rc;
    if (!cov || !off || !order || !scale || !counters || !kfac || !est) return fail(DN_ERR_INVALID, "null pointer argument%s");
    if (n_work <= 0) return DN_OK;
    int grid = n_work < 148 * 8 ? n_work : 148 * 8;
    estimates_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(cov, (const long long *)off, order, n_work, prm->p, scale,
                                                             counters, kfac, e_first, est);
    DN_CUDA(cudaGetLastError());
    return DN_OK;
}

int64_t dn_sums_workspace_bytes(int32_t n_genes, int32_t p) {
    const long long nb = (n_genes + SLAB_ROWS - 1) / SLAB_ROWS;
    return (nb > 0 ? nb : 1) * (3ll * p + 1) * 8;
}

static int sums_common(int mode, const double *A, const double *B, const double *C, int32_t n, int32_t p, double *rho0,
                       double *sums, void *workspace, int64_t workspace_bytes, void *stream) {
    if (n <= 0 || p <= 0 || !sums || !workspace) return fail(DN_ERR_INVALID, "bad argument%s");
    if (workspace_bytes < dn_sums_workspace_bytes(n, p)) return fail(DN_ERR_WORKSPACE, "workspace too small%s");
    const int nb = (n + SLAB_ROWS - 1) / SLAB_ROWS;
    cudaStream_t st = (cudaStream_t)stream;
    sums_partial_kernel<<<nb, 256, 0, st>>>(mode, A, B, C, n, p, rho0, (double *)workspace);
    DN_CUDA(cudaGetLastError());
    const int nvec = 3 * p + 1;
    sums_final_kernel<<<(nvec + 127) / 128, 128, 0, st>>>((const double *)workspace, nb, p, nvec, sums);
    DN_CUDA(cudaGetLastError());
    return DN_OK;
}

int dn_outer_sums(const double *x_weighted, const double *rho, int32_t n_genes, int32_t p, double *sums, void *workspace,
                  int64_t workspace_bytes, void *stream) {
    if (!x_weighted || !rho) return fail(DN_ERR_INVALID, "null pointer argument%s");
    return sums_common(0, x_weighted, rho, nullptr, n_genes, p, nullptr, sums, workspace, workspace_bytes, stream);
}

int dn_init_sums(const double *est_rowsum, const double *cov_rowsum, const double *reads, int32_t n_genes, int32_t p,
                 double *rho0, double *sums, void *workspace, int64_t workspace_bytes, void *stream) {
    if (!est_rowsum || !cov_rowsum || !reads || !rho0) return fail(DN_ERR_INVALID, "null pointer argument%s");
    return sums_common(1, est_rowsum, cov_rowsum, reads, n_genes, p, rho0, sums, workspace, workspace_bytes, stream);
}

int dn_outer_apply(const double *sums, int32_t n_genes, int32_t p, double *x_weighted, double *rho, double *x_adj,
                   double *norm_factors, double *scale_factors, void *stream) {
    if (!sums || !x_weighted || !rho || !x_adj || !norm_factors || !scale_factors || n_genes <= 0 || p <= 0)
        return fail(DN_ERR_INVALID, "bad argument%s");
    const int nb = (n_genes + SLAB_ROWS - 1) / SLAB_ROWS;
    outer_apply_kernel<<<nb, 256, 3 * p * sizeof(double), (cudaStream_t)stream>>>(sums, n_genes, p, x_weighted, rho, x_adj,
                                                                                   norm_factors, scale_factors);
    DN_CUDA(cudaGetLastError());
    return DN_OK;
}

int dn_init_apply(const double *sums, const double *reads, int32_t n_genes, int32_t p, double *x_weighted,
                  double *norm_factors, double *scale_factors, void *stream) {
    if (!sums || !reads || !x_weighted || !norm_factors || !scale_factors || n_genes <= 0 || p <= 0)
        return fail(DN_ERR_INVALID, "bad argument%s");
    long long tot = (long long)n_genes * p;
    int nb = (int)((tot + 255) / 256);
    if (nb > 148 * 8) nb = 148 * 8;
    init_apply_kernel<<<nb, 256, 2 * p * sizeof(double), (cudaStream_t)stream>>>(sums, reads, n_genes, p, x_weighted,
                                                                                  norm_factors, scale_factors);
    DN_CUDA(cudaGetLastError());
    return DN_OK;
}

}  // extern "C"
