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

#include "common.cuh"
#include "launch.h"

namespace {

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
    double *tp;      // NT partial dot products (phase A row slices)
    double *gacc;    // ntiles * TR * TR doubles of global scratch: Gram accumulators when tiles > threads
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

// TR consecutive rows of one column of the M tile.  Column-major tile (8 x 8 Gram tiles): TR / 2 128-bit loads.  A lane's
// operands for a column are two such runs (its tile's row block and column block), i.e. 2 TR doubles for TR^2 FMAs;
// with the earlier row-major tile every one of them was a separate 64-bit load and the row blocks of a warp's
// tiles fell on two bank groups (16-way conflicts).
template <int TR, bool CM>
__device__ __forceinline__ void load_rows(const double *blk, int cidx, int ldm, double (&out)[TR]) {
    if constexpr (CM) {
        const double2 *q = reinterpret_cast<const double2 *>(blk + cidx * ldm);
#pragma unroll
        for (int r = 0; r < TR / 2; ++r) { const double2 t = q[r]; out[2 * r] = t.x; out[2 * r + 1] = t.y; }
    } else {
#pragma unroll
        for (int r = 0; r < TR; ++r) out[r] = blk[r * ldm + cidx];
    }
}

// ---- one pass over the current columns: (optional multiplier update) + Gram accumulate -------------------------
// UPDATE=false: G = x x^T (first rank-one fit of nmf(), nmf.py:88).
// UPDATE=true : lambda <- max(0, lambda - c (K E - x)), M = x + lambda, G = M M^T (nmf.py:93-98),
//               with K E = v (v . M_old) per column.
template <int TR, int NT, bool UPDATE>
__device__ void gram_pass(const KArgs &a, Gene &g, int ti, int tj, int ks, bool tile_ok) {
    const int tid = threadIdx.x;
    const int p = a.p, pp = a.pp, ldm = a.ldm, CH = a.ch, KS = a.ks;
    constexpr bool CM = TR == 8;                  // column-major M tile (see load_rows); row-major for the 4 x 4 tiles
    auto mi = [&](int i, int c) { return CM ? c * ldm + i : i * ldm + c; };
    double acc[TR][TR];
#pragma unroll
    for (int r = 0; r < TR; ++r)
#pragma unroll
        for (int q = 0; q < TR; ++q) acc[r][q] = 0.0;

    for (int base = 0; base < g.n_cur; base += CH) {
        const int ncol = min(CH, g.n_cur - base);
        // phase A.  Small p (chunk as wide as the CTA): one thread per column.  Otherwise the CTA's NT / CH row
        // slices share a column (thread = column c, slice s; rows s, s + S, ...): the p loads of a column are then
        // spread over S threads instead of one dependent loop -- at p = 200 the chunk is 32 columns wide and a
        // single warp walking 2 x 200 strided rows per column left the pass latency-bound on global memory.
#ifdef TILED_ONE_THREAD_PER_COLUMN
        const int S = 1;                       // (A/B switch: the previous phase A)
#else
        const int S = NT / CH;
#endif
        if (S <= 1) {
        if (tid < ncol) {
            const int pc = phys_col(g, base + tid);
            const double *xc = g.X + pc;
            if (!UPDATE) {
                for (int i = 0; i < p; ++i) g.ms[mi(i, tid)] = xc[i * g.ld];
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
                    g.ms[mi(i, tid)] = x + l;
                }
            }
        }
        } else {
            const int c = tid % CH, sl = tid / CH;
            const bool act = c < ncol && sl < S;
            const int pc = act ? phys_col(g, base + c) : 0;
            const double *xc = g.X + pc;
            if (!UPDATE) {
                if (act) {
#pragma unroll 4
                    for (int i = sl; i < p; i += S) g.ms[mi(i, c)] = xc[(long long)i * g.ld];
                }
            } else {
                double *lc = g.Lm + pc;
                double tp0 = 0.0, tp1 = 0.0;
                if (act) {
                    int i = sl;
#pragma unroll 2
                    for (; i + S < p; i += 2 * S) {
                        tp0 = fma(g.v[i], xc[(long long)i * g.ld] + lc[(long long)i * g.ld], tp0);
                        tp1 = fma(g.v[i + S], xc[(long long)(i + S) * g.ld] + lc[(long long)(i + S) * g.ld], tp1);
                    }
                    if (i < p) tp0 = fma(g.v[i], xc[(long long)i * g.ld] + lc[(long long)i * g.ld], tp0);
                }
                // the slices' partial dot products meet in shared memory and are summed in slice order
                if (sl < S) g.tp[sl * CH + c] = tp0 + tp1;
                __syncthreads();
                double t = 0.0;
                for (int q = 0; q < S; ++q) t += g.tp[q * CH + c];
                __syncthreads();
                if (act) {
#pragma unroll 4
                    for (int i = sl; i < p; i += S) {
                        const double x = xc[(long long)i * g.ld];
                        double l = lc[(long long)i * g.ld];
                        const double res = g.v[i] * t - x;        // est - x
                        l = l - a.c * res;
                        l = l < 0.0 ? 0.0 : l;
                        lc[(long long)i * g.ld] = l;
                        g.ms[mi(i, c)] = x + l;
                    }
                }
            }
        }
        __syncthreads();
        if (a.nsets > 1) {
            // more Gram tiles than threads: every thread sweeps its tile of each set over the chunk, with the
            // accumulators of all tiles parked in the CTA's global scratch between chunks
            const int ntg_ = pp / TR;
            const int ntl_ = ntg_ * (ntg_ + 1) / 2;
            for (int set = 0; set < a.nsets; ++set) {
                const int tile = set * NT + tid;
                if (tile < ntl_) {
                    int t = tile, si = 0;
                    while (t >= ntg_ - si) { t -= ntg_ - si; ++si; }
                    const int sj = si + t;
                    double *ga = g.gacc + (long long)tile * (TR * TR);
                    if (base == 0) {
#pragma unroll
                        for (int r = 0; r < TR; ++r)
#pragma unroll
                            for (int q = 0; q < TR; ++q) acc[r][q] = 0.0;
                    } else {
#pragma unroll
                        for (int r = 0; r < TR; ++r)
#pragma unroll
                            for (int q = 0; q < TR; ++q) acc[r][q] = ga[r * TR + q];
                    }
                    const double *ma = g.ms + (CM ? si * TR : si * TR * ldm);
                    const double *mb = g.ms + (CM ? sj * TR : sj * TR * ldm);
                    for (int cidx = 0; cidx < ncol; ++cidx) {
                        double av[TR], bv[TR];
                        load_rows<TR, CM>(ma, cidx, ldm, av);
                        load_rows<TR, CM>(mb, cidx, ldm, bv);
#pragma unroll
                        for (int r = 0; r < TR; ++r)
#pragma unroll
                            for (int q = 0; q < TR; ++q) acc[r][q] = fma(av[r], bv[q], acc[r][q]);
                    }
#pragma unroll
                    for (int r = 0; r < TR; ++r)
#pragma unroll
                        for (int q = 0; q < TR; ++q) ga[r * TR + q] = acc[r][q];
                }
            }
        } else
        // phase B: register-tiled SYRK out of the shared tile
        if (tile_ok) {
            const double *ma = g.ms + (CM ? ti * TR : ti * TR * ldm);
            const double *mb = g.ms + (CM ? tj * TR : tj * TR * ldm);
            for (int cidx = ks; cidx < ncol; cidx += KS) {
                double av[TR], bv[TR];
                load_rows<TR, CM>(ma, cidx, ldm, av);
                load_rows<TR, CM>(mb, cidx, ldm, bv);
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
    if (a.nsets > 1) {
        __syncthreads();
        for (int e = tid; e < ntiles * TR * TR; e += NT) {
            const double sacc = g.gacc[e];
            const int tile_e = e / (TR * TR), rq = e - tile_e * (TR * TR);
            int t = tile_e, tii = 0;
            while (t >= ntg - tii) { t -= ntg - tii; ++tii; }
            const int tjj = tii + t;
            const int i = tii * TR + rq / TR, j = tjj * TR + rq % TR;
            g.G[(long long)i * pp + j] = sacc;
            if (tii != tjj) g.G[(long long)j * pp + i] = sacc;
        }
    } else if (KS == 1) {
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
            if (CM) {
                for (int e = tid; e < (pp - p) * CH; e += NT) g.ms[(e / (pp - p)) * ldm + p + e % (pp - p)] = 0.0;
            } else {
                for (int e = tid; e < (pp - p) * ldm; e += NT) g.ms[p * ldm + e] = 0.0;
            }
        }
    }
    __syncthreads();
}

template <int NT>
__device__ __forceinline__ void eig_solve(const KArgs &a, Gene &g, bool cold) {
    int conv = 1;
    if (a.g_in_smem) {
        if (threadIdx.x < 32) {
            int s;
            s = eig_warp(g.G, a.pp, a.p, g.v, cold, g.ibuf + 12);
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
template <int TR, int NT>
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
    gram_pass<TR, NT, false>(a, g, ti, tj, ks, tile_ok);
    eig_solve<NT>(a, g, true);
    for (int it = 0; it < T; ++it) {
        gram_pass<TR, NT, true>(a, g, ti, tj, ks, tile_ok);
        eig_solve<NT>(a, g, false);
    }
    final_pass<NT>(a, g, first, T > 0, want_res, e_first_g);
}

// (two CTAs per SM for the 4 x 4 tiles, which the init pass of every p uses: 128 registers, as before the row-slice
// phase A grew the unconstrained allocation to 246)
template <int TR, int NT>
__global__ void __launch_bounds__(NT, TR == 4 ? 2 : 1) nmfoa_kernel(const KArgs a) {
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
    g.tp = smem + cv.tp;
    double *slab = a.ws + (long long)blockIdx.x * a.ws_stride;
    g.B0 = slab;
    long long slab_o = 2ll * pp * pp;
    if (a.g_in_smem) {
        g.G = smem + cv.G;
    } else {
        g.G = slab + slab_o;
        slab_o += (long long)pp * pp;
    }
    g.gacc = slab + slab_o;
    if (a.nsets > 1) slab_o += a.gacc_doubles;
    g.eig_steps = 0;
    g.eig_fallbacks = 0;

    // Gram tile owned by this thread
    int ks = 0, ti = 0, tj = 0;
    bool tile_ok = false;
    {
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
                run_nmf<TR, NT>(a, g, true, false, nullptr, ti, tj, ks, tile_ok);
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
            if (a.row_max_out) {
                // row maxima of the raw coverage: max_j(F_ij / s_i) = (max_j F_ij) / s_i exactly, so the outer
                // iterations get the matrix maximum of the scaled coverage (nmf.py:76) without re-reading the gene
                for (int i = warp; i < p; i += NT / 32) {
                    const double *row = F + (long long)i * L;
                    double m = -1.0e300;
                    for (int j = lane; j < L; j += 32) m = fmax(m, row[j]);
                    m = warp_max(m);
                    if (lane == 0) a.row_max_out[(long long)gid * p + i] = m;
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
        if (a.row_max) {
            if (tid < p) tmax = a.row_max[(long long)gid * p + tid] / g.scale[tid];
        } else {
            for (int i = 0; i < p; ++i) {
                const double *row = F + (long long)i * L;
                double m = -1.0e300;
                for (int j = tid; j < L; j += NT) m = fmax(m, row[j]);
                tmax = fmax(tmax, m / g.scale[i]);
            }
        }
        const double gmax = block_max<NT>(tmax, g.red);
        const double thr = (a.flags & DN_FLAG_PLAIN_NMF) ? -1.0e300 : 0.1 * gmax;                                   // nmf.py:76
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
            if (any_empty && !(a.flags & DN_FLAG_PLAIN_NMF)) {
                exit_code = DN_EXIT_EMPTY_SAMPLE;                        // nmf.py:241-242
            } else {
                const bool store_e = (a.e_first != nullptr) && (n0 == L);
                // (4) first fit (nmf.py:245-254)
                run_nmf<TR, NT>(a, g, true, true, store_e ? a.e_first + o0 : nullptr, ti, tj, ks, tile_ok);
                nmf_calls = 1; sum_cols = n0;
                if (tid < p) {
                    g.rho[tid] = 1.0 - g.rs0[tid] / (g.tmp[tid] + 1.0);
                    g.K0[tid] = g.K[tid];
                    g.rsC0[tid] = g.rsC[tid];
                }
                __syncthreads();
                if (!(a.flags & DN_FLAG_PLAIN_NMF) && median_one_minus(g.rho, p) > 1.0) {
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
                            run_nmf<TR, NT>(a, g, false, true, nullptr, ti, tj, ks, tile_ok);
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
            if (!(a.flags & DN_FLAG_RAW_RHO)) r = r > 0.9 ? 0.9 : r;                                       // nmf.py:398-399
            if (!(a.flags & DN_FLAG_RAW_RHO)) r = r < 0.0 ? 0.0 : r;
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
        if (a.est) {
            write_estimate(F, L, p, g.scale, exit_code, n0, g.K, a.e_first ? a.e_first + o0 : nullptr,
                           a.est + (long long)p * (a.est_off ? a.est_off[gid] : o0), tid, NT);
            __syncthreads();
        }
    }
}

template <int TR, int NT>
int launch(const KArgs &a, const dn_plan *plan, cudaStream_t st) {
    auto kern = nmfoa_kernel<TR, NT>;
    DN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, plan->smem_bytes));
    kern<<<plan->ctas, NT, plan->smem_bytes, st>>>(a);
    DN_CUDA(cudaGetLastError());
    return DN_OK;
}

}  // namespace

int dn_launch_tiled(const KArgs &a, const dn_plan *plan, cudaStream_t st) {
    if (plan->tile == 4) return launch<4, 256>(a, plan, st);
    if (plan->tile == 8) return launch<8, 256>(a, plan, st);
    return dn_fail(DN_ERR_INVALID, "plan.tile must be 4 or 8 on the tiled path%s");
}
