// degnorm_b200 -- fused baseline-selection kernel for 49..208 samples, sm_100a ("wide" kernel; C5: p = 200).
//
// Same per-gene flow as the other kernels (reference: /root/reference/degnorm/nmf.py:189-372 around nmf.py:78-107).
// At p = 200 a column-pass needs p (p + 1) / 2 = 20,100 fused multiply-adds for the Gram matrix against 4,800
// algorithmic bytes: the pass is bound by the FP64 pipe (64 DFMA per clock and SM: ~314 clocks per column), HBM
// is second (~220 clocks at the measured copy bandwidth).  Design:
//
//   * one thread owns one 8 x 8 tile of the upper triangle of G for a WHOLE pass: 325 tiles at p = 200, held in
//     registers by 11 of the CTA's 12 warps (64 accumulators per thread); nothing is parked in memory between
//     chunks (the generic tiled kernel did, at ~5,000 clocks per column).  The register file of an SM (64 K words)
//     holds the triangle of up to 208 samples; wider cohorts stay on the generic kernel.
//   * x and M = x + lambda live in per-CTA global slabs, column-major with a column stride of pp + 2 doubles
//     (pp = samples rounded up to 8); the CTA walks its columns in 16-column chunks through a 3-stage ring filled by
//     TMA 1-D bulk copies (cp.async.bulk + one mbarrier per stage); the updated M goes back to the slab as ONE bulk
//     store per chunk.  Phase A (multiplier update: 16 lanes per column) is ~4 % of the work and simply precedes
//     phase B inside a chunk (two block barriers per chunk of 16 x 20,100 FMAs).
//   * phase B: per column a thread reads its 8 row operands and 8 column operands from the stage (8 LDS.128) and
//     issues 64 DFMA.  Shared memory is as loaded as the FP64 pipe here (4 wavefronts per 128-bit load of 32 distinct
//     addresses), so the tiles are dealt to the lanes row-major with every tile row padded to an even length: the two
//     lanes of an aligned pair then share their row operands, and a half-warp whose pairs read the same 16 bytes is
//     served in ONE wavefront (measured: tools/probe_peaks.py) -- 24 instead of 32 wavefronts per warp and column.
//     Row- and column-operand pairs are fetched in an order rotated by (tile row / 2) resp. (tile column / 2): the
//     lanes of a quarter-warp then hit distinct banks (tile rows / columns are 64 bytes apart).
//   * the p x p eigen-solve never materialises G: every thread multiplies its register tile (and its transpose) with
//     the current vector, partial products meet in shared memory in a fixed order.  G is written out (to the slab)
//     only for the rare small-gap fallback (repeated squaring, common.cuh).
//   * narrow cohorts (fewer tiles than threads) split the columns of a chunk over k-slices of threads; long genes
//     take a thread-block cluster (contiguous column slices, partial Grams summed through global slots in rank order),
//     exactly the scheme of the mid-p kernel.
//
// No tensor cores: fp64 FMA pipe only.
#include <cooperative_groups.h>
#include "common.cuh"
#include "launch.h"
#include "tma.cuh"
namespace cg = cooperative_groups;

namespace {

constexpr int WNT = WIDE_THREADS;         // 384 threads: 12 warps, three per SM sub-partition (168 registers each)
constexpr int WNW = WNT / 32;
constexpr int WCH = WIDE_CHUNK;           // columns per ring stage
constexpr int WLPC = 16;                  // lanes per column in the update / scan-type passes
constexpr int WCPR = WNT / WLPC;          // columns per CTA step in those passes (24)
constexpr int WRPL = (WIDE_MAX_PP + WLPC - 1) / WLPC;   // rows per lane there (13)

struct WGene {
    double *v, *K, *K0, *rs0, *rsF, *rsC, *rsC0, *rho, *scale, *tmp, *diag, *red, *binm, *part, *ring;
    int *alive, *ibuf, *lw;
    double *X, *M, *resb, *tb, *Gfull, *slots, *kslab;      // global (slab)
    long long slot_stride;
    int pp, cs_col, nb, ntiles, ks, ne;
    int n0, n_cur, n0g, n_curg, goff, cs, nb0, nalive;
    int crank, csize, xpar, eig_steps, eig_fallbacks;
    bool primed;
    unsigned long long *mbar;
    unsigned seq;                                  // ring chunks consumed so far (stage and mbarrier parity follow)
};

__device__ __forceinline__ int wlstart(const WGene &g, int k) {
    int s = 0;
    for (int q = 0; q < k; ++q) s += g.lw[g.alive[q]];
    return s;
}
__device__ __forceinline__ int tile_index(int nb, int ti, int tj) { return ti * nb - ti * (ti - 1) / 2 + (tj - ti); }

// All-to-all sum over the cluster of n doubles (vals: shared, published to the CTA) through the CTAs' global slots.
__device__ void wclu_allsum(WGene &g, const double *vals, int n, double *out) {
    const int tid = threadIdx.x;
    if (g.csize == 1) {
        __syncthreads();
        for (int k = tid; k < n; k += WNT) out[k] = vals[k];
        __syncthreads();
        return;
    }
    cg::cluster_group cl = cg::this_cluster();
    double *mine = g.slots + (long long)g.xpar * g.ne;
    for (int k = tid; k < n; k += WNT) mine[k] = vals[k];
    __threadfence();
    cl.sync();
    const double *first = g.slots - (long long)g.crank * g.slot_stride + (long long)g.xpar * g.ne;
    for (int k = tid; k < n; k += WNT) {
        double s = 0.0;
        for (int r = 0; r < g.csize; ++r) s += __ldcg(first + (long long)r * g.slot_stride + k);
        out[k] = s;
    }
    g.xpar ^= 1;
    __syncthreads();
}

// ---- one pass over this CTA's columns: (optional multiplier update) + Gram tile of M in `acc` ----------------------
// acc[r][q] = G[8 ti + rot8(r, ti)][8 tj + rot8(q, tj)] (the rotated fetch orders).
__device__ __forceinline__ int rot8(int q, int t) { return 2 * (((q >> 1) + ((t >> 1) & 3)) & 3) + (q & 1); }
struct WPass {
    double *ring, *M, *X, *v, *slots, *kslab;
    unsigned long long *mbar;
    long long slot_stride;
    double c;
    int n_cur, pp, cs_col, ti, tj, ks, nks, crank, csize, xpar, ne, primed, tile, ovl;
    unsigned seq;
};
struct WPassOut { unsigned seq; int xpar; };

template <bool UPDATE>
__device__ __forceinline__ WPassOut gram_wide(const WPass g, const bool prime_next, double (&acc)[8][8]) {
    const int tid = threadIdx.x, lane = tid & 31;
    const int n = g.n_cur, CS = g.cs_col;
    const int nchunk = (n + WCH - 1) / WCH;
    const int STG = 2 * WCH * CS;                          // doubles per ring stage (M then x)
    const unsigned CHB = (unsigned)(WCH * CS * 8);
    const unsigned seq0 = g.seq;
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[r][q] = 0.0;

    const int itid = g.ovl ? (WNW - 1) * 32 : 0;            // the thread that talks to the TMA unit
    auto issue = [&](int ch, bool with_x) {
        if (tid == itid && ch < nchunk) {
            const unsigned st = (seq0 + (unsigned)ch) % WIDE_RING;
            double *dst = g.ring + st * STG;
            fence_proxy_async_smem();
            mbar_expect_tx(g.mbar + st, with_x ? 2 * CHB : CHB);
            bulk_g2s(dst, g.M + (long long)ch * (WCH * CS), CHB, g.mbar + st);
            if (with_x) bulk_g2s(dst + WCH * CS, g.X + (long long)ch * (WCH * CS), CHB, g.mbar + st);
        }
    };
    if (!g.primed) {
        issue(0, UPDATE);
        issue(1, UPDATE);
    }
    const bool has_tile = g.tile >= 0;
    const int ao0 = 8 * g.ti + rot8(0, g.ti), ao1 = 8 * g.ti + rot8(2, g.ti), ao2 = 8 * g.ti + rot8(4, g.ti),
              ao3 = 8 * g.ti + rot8(6, g.ti);
    const int uo0 = 8 * g.tj + rot8(0, g.tj), uo1 = 8 * g.tj + rot8(2, g.tj), uo2 = 8 * g.tj + rot8(4, g.tj),
              uo3 = 8 * g.tj + rot8(6, g.tj);
    // phase A for one column of a chunk: 16 lanes (hl = 0 .. 15) share it, rows hl, hl + 16, ...
    auto phase_a = [&](double *sM, int ncol, int col, int hl) {
        const bool act = col < ncol;
        double *mc = sM + col * CS;
        const double *xc = sM + WCH * CS + col * CS;
        double tp = 0.0;
        if (act) {
            double t0 = 0.0, t1 = 0.0;                      // (two chains, four rows in flight)
            int r = hl;
#pragma unroll 1
            for (; r + 3 * WLPC < g.pp; r += 4 * WLPC) {
                const double m0 = mc[r], m1 = mc[r + WLPC], m2 = mc[r + 2 * WLPC], m3 = mc[r + 3 * WLPC];
                const double v0 = g.v[r], v1 = g.v[r + WLPC], v2 = g.v[r + 2 * WLPC], v3 = g.v[r + 3 * WLPC];
                t0 = fma(v0, m0, t0); t1 = fma(v1, m1, t1); t0 = fma(v2, m2, t0); t1 = fma(v3, m3, t1);
            }
            for (; r < g.pp; r += WLPC) t0 = fma(g.v[r], mc[r], t0);
            tp = t0 + t1;
        }
#pragma unroll
        for (int o = 1; o < WLPC; o <<= 1) tp += __shfl_xor_sync(0xffffffffu, tp, o);
        if (act) {
#pragma unroll 4
            for (int r = hl; r < g.pp; r += WLPC) {
                const double x = xc[r], m = mc[r];
                const double res = fma(g.v[r], tp, -x);
                const double w = fma(-g.c, res, m - x);
                mc[r] = fma(0.5, w + fabs(w), x);
            }
        }
    };
    // phase B: this thread's tile over the chunk's columns (its k-slice of them)
    auto phase_b = [&](const double *sM, int ncol) {
        if (has_tile) {
            const double *mc = sM + g.ks * CS;
            const int mstep = g.nks * CS;
#pragma unroll 1
            for (int cc = g.ks; cc < ncol; cc += g.nks, mc += mstep) {
                const double2 a0 = *reinterpret_cast<const double2 *>(mc + ao0);
                const double2 a1 = *reinterpret_cast<const double2 *>(mc + ao1);
                const double2 a2 = *reinterpret_cast<const double2 *>(mc + ao2);
                const double2 a3 = *reinterpret_cast<const double2 *>(mc + ao3);
                const double ar[8] = {a0.x, a0.y, a1.x, a1.y, a2.x, a2.y, a3.x, a3.y};
                {
                    const double2 u = *reinterpret_cast<const double2 *>(mc + uo0);
#pragma unroll
                    for (int r = 0; r < 8; ++r) { acc[r][0] = fma(ar[r], u.x, acc[r][0]); acc[r][1] = fma(ar[r], u.y, acc[r][1]); }
                }
                {
                    const double2 u = *reinterpret_cast<const double2 *>(mc + uo1);
#pragma unroll
                    for (int r = 0; r < 8; ++r) { acc[r][2] = fma(ar[r], u.x, acc[r][2]); acc[r][3] = fma(ar[r], u.y, acc[r][3]); }
                }
                {
                    const double2 u = *reinterpret_cast<const double2 *>(mc + uo2);
#pragma unroll
                    for (int r = 0; r < 8; ++r) { acc[r][4] = fma(ar[r], u.x, acc[r][4]); acc[r][5] = fma(ar[r], u.y, acc[r][5]); }
                }
                {
                    const double2 u = *reinterpret_cast<const double2 *>(mc + uo3);
#pragma unroll
                    for (int r = 0; r < 8; ++r) { acc[r][6] = fma(ar[r], u.x, acc[r][6]); acc[r][7] = fma(ar[r], u.y, acc[r][7]); }
                }
            }
        }
    };
    if (WIDE_OVERLAP && g.ovl) {
        // ---- overlapped schedule (WIDE_OVERLAP = 1; the twelfth warp owns no tile): it runs the TMA traffic and the
        // multiplier update of chunk ch + 1 WHILE the eleven Gram warps accumulate chunk ch -- one block barrier per
        // chunk instead of two.  Parity-tested, measured slower (see launch.h) and off by default.
        const bool updw = (tid >> 5) == WNW - 1;
        if (UPDATE && updw && nchunk > 0) {
            const unsigned k0 = seq0, st0 = k0 % WIDE_RING;
            mbar_wait(g.mbar + st0, (k0 / WIDE_RING) & 1u);
            for (int rd = 0; rd < WCH / 2; ++rd) phase_a(g.ring + st0 * STG, min(WCH, n), 2 * rd + (lane >> 4), lane & 15);
            fence_proxy_async_smem();
        }
#pragma unroll 1
        for (int ch = 0; ch < nchunk; ++ch) {
            const unsigned k = seq0 + (unsigned)ch, st = k % WIDE_RING;
            __syncthreads();                               // chunk ch - 1 accumulated everywhere, chunk ch updated
            if (updw) {
                if (lane == 0) {
                    if constexpr (UPDATE) {
                        if (ch >= 1) {
                            const unsigned sp = (k - 1) % WIDE_RING;
                            const int ncp = min(WCH, n - (ch - 1) * WCH);
                            bulk_s2g(g.M + (long long)(ch - 1) * (WCH * CS), g.ring + sp * STG, (unsigned)(ncp * CS * 8));
                        }
                        bulk_wait_read<0>();               // (this warp has a chunk's time to spare)
                    }
                    issue(ch + 2, UPDATE);                 // into the stage chunk ch - 1 has just left
                }
                __syncwarp();
                if (UPDATE && ch + 1 < nchunk) {
                    const unsigned k1 = k + 1, st1 = k1 % WIDE_RING;
                    mbar_wait(g.mbar + st1, (k1 / WIDE_RING) & 1u);
                    const int ncol1 = min(WCH, n - (ch + 1) * WCH);
                    for (int rd = 0; rd < WCH / 2; ++rd) phase_a(g.ring + st1 * STG, ncol1, 2 * rd + (lane >> 4), lane & 15);
                    fence_proxy_async_smem();              // the stage is read by the bulk store later
                }
            } else {
                if constexpr (!UPDATE) mbar_wait(g.mbar + st, (k / WIDE_RING) & 1u);
                phase_b(g.ring + st * STG, min(WCH, n - ch * WCH));
            }
        }
    } else {
    const int half = tid / WLPC, hl = tid % WLPC;           // phase A: column of the chunk / lane within the column
#pragma unroll 1
    for (int ch = 0; ch < nchunk; ++ch) {
        const unsigned k = seq0 + (unsigned)ch, st = k % WIDE_RING;
        double *sM = g.ring + st * STG;
        const int ncol = min(WCH, n - ch * WCH);
        __syncthreads();                                   // everyone is done with chunk ch - 1
        if constexpr (UPDATE) {
            // stage of chunk ch - 1 holds that chunk's final M -> one bulk store; the store of chunk ch - 2 has read
            // its stage by now -> that stage takes chunk ch + 1
            if (tid == itid) {
                if (ch >= 1) {
                    const unsigned sp = (k - 1) % WIDE_RING;
                    const int ncp = min(WCH, n - (ch - 1) * WCH);
                    bulk_s2g(g.M + (long long)(ch - 1) * (WCH * CS), g.ring + sp * STG, (unsigned)(ncp * CS * 8));
                }
                bulk_wait_read<1>();
            }
            if (ch + 1 >= 2) issue(ch + 1, true);
        } else {
            issue(ch + 2, false);
        }
        mbar_wait(g.mbar + st, (k / WIDE_RING) & 1u);       // chunk ch has landed
        if constexpr (UPDATE) {
            // phase A: 16 lanes per column (columns 0 .. 15 of the chunk: threads 0 .. 255)
            if (half < WCH) {
                phase_a(sM, ncol, half, hl);
                fence_proxy_async_smem();                  // the stage is read by the bulk store later
            }
            __syncthreads();                               // every thread reads every column in phase B
        }
        phase_b(sM, ncol);
    }
    }
    WPassOut out;
    out.seq = (seq0 + (unsigned)nchunk) % (2 * WIDE_RING);    // (stage, parity) of chunk k depend on k mod 2 RING only
    fence_proxy_async();
    __syncthreads();
    if constexpr (UPDATE) {
        if (tid == itid && nchunk > 0) {
            const unsigned sp = (seq0 + (unsigned)nchunk - 1u) % WIDE_RING;
            const int ncp = n - (nchunk - 1) * WCH;
            bulk_s2g(g.M + (long long)(nchunk - 1) * (WCH * CS), g.ring + sp * STG, (unsigned)(ncp * CS * 8));
            bulk_wait_all();                               // the slab holds the whole new M before anyone reads it
        }
        __syncthreads();
    }
    if (prime_next && tid == itid) {
        for (int q = 0; q < 2 && q < nchunk; ++q) {
            const unsigned st = (out.seq + (unsigned)q) % WIDE_RING;
            double *dst = g.ring + st * STG;
            fence_proxy_async_smem();
            mbar_expect_tx(g.mbar + st, 2 * CHB);
            bulk_g2s(dst, g.M + (long long)q * (WCH * CS), CHB, g.mbar + st);
            bulk_g2s(dst + WCH * CS, g.X + (long long)q * (WCH * CS), CHB, g.mbar + st);
        }
    }
    // ---- k-slices -> slice 0 (fixed order, through the slab), then the cluster sum (rank order, through the slots)
    out.xpar = g.xpar;
    if (g.nks > 1) {
        if (has_tile && g.ks > 0) {
            double *dst = g.kslab + ((long long)(g.ks - 1) * g.ne + (long long)g.tile * 64);
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int q = 0; q < 8; q += 2)
                    *reinterpret_cast<double2 *>(dst + r * 8 + q) = make_double2(acc[r][q], acc[r][q + 1]);
        }
        __syncthreads();
        if (has_tile && g.ks == 0) {
            for (int s = 1; s < g.nks; ++s) {
                const double *src = g.kslab + ((long long)(s - 1) * g.ne + (long long)g.tile * 64);
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int q = 0; q < 8; q += 2) {
                        const double2 t = __ldcg(reinterpret_cast<const double2 *>(src + r * 8 + q));
                        acc[r][q] += t.x;
                        acc[r][q + 1] += t.y;
                    }
            }
        }
        __syncthreads();
    }
    if (g.csize > 1) {
        cg::cluster_group cl = cg::this_cluster();
        if (has_tile && g.ks == 0) {
            double *dst = g.slots + (long long)g.xpar * g.ne + (long long)g.tile * 64;
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int q = 0; q < 8; q += 2)
                    *reinterpret_cast<double2 *>(dst + r * 8 + q) = make_double2(acc[r][q], acc[r][q + 1]);
        }
        __threadfence();
        cl.sync();
        if (has_tile && g.ks == 0) {
            const double *first = g.slots - (long long)g.crank * g.slot_stride + (long long)g.xpar * g.ne + (long long)g.tile * 64;
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    double s = 0.0;
                    for (int rk = 0; rk < g.csize; ++rk) s += __ldcg(first + (long long)rk * g.slot_stride + r * 8 + q);
                    acc[r][q] = s;
                }
        }
        out.xpar = g.xpar ^ 1;
        __syncthreads();
    }
    return out;
}

// ---- y = G v from the register tiles: part[tile][0..7] = tile . v[cols], part[tile][8..15] = tile^T . v[rows] -------
struct WTile { int tile, ti, tj, ks; bool owner; };          // owner: this thread holds the summed tile (k-slice 0)

__device__ __forceinline__ void tile_matvec(const WGene &g, const WTile &t, const double (&acc)[8][8]) {
    if (t.owner) {
        double vj[8], vi[8], y[8], z[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            vj[q] = g.v[8 * t.tj + rot8(q, t.tj)];
            vi[q] = g.v[8 * t.ti + rot8(q, t.ti)];
            y[q] = 0.0;
            z[q] = 0.0;
        }
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                y[r] = fma(acc[r][q], vj[q], y[r]);
                z[q] = fma(acc[r][q], vi[r], z[q]);
            }
        double *p0 = g.part + (long long)t.tile * 16;
#pragma unroll
        for (int r = 0; r < 8; ++r) p0[rot8(r, t.ti)] = y[r];
#pragma unroll
        for (int q = 0; q < 8; ++q) p0[8 + rot8(q, t.tj)] = (t.ti == t.tj) ? 0.0 : z[q];
    }
}
// row i of G v: the partial products of the tiles of block row / block column i / 8, in a fixed order
__device__ __forceinline__ double matvec_row(const WGene &g, int i) {
    const int b = i >> 3, r = i & 7;
    double s = 0.0;
    for (int ti = 0; ti < b; ++ti) s += g.part[(long long)tile_index(g.nb, ti, b) * 16 + 8 + r];
    for (int tj = b; tj < g.nb; ++tj) s += g.part[(long long)tile_index(g.nb, b, tj) * 16 + r];
    return s;
}

// Top eigenvector of G (held as register tiles) -> g.v.  Same rules as the other kernels: warm-started power iteration
// to |dv|_inf <= EIG_TOL, a converged warm start with a vanishing entry on a covered sample is distrusted, slow
// convergence (small spectral gap) and distrust hand over to repeated squaring on the materialised G (common.cuh).
__device__ void eig_wide(const KArgs &a, WGene &g, const WTile &t, const double (&acc)[8][8], bool cold) {
    const int tid = threadIdx.x;
    const int p = a.p, pp = g.pp;
    if (t.owner && t.ti == t.tj) {
        // (rows and columns of a diagonal tile are rotated alike: acc[r][r] is a diagonal entry of G)
#pragma unroll
        for (int r = 0; r < 8; ++r) g.diag[8 * t.ti + rot8(r, t.ti)] = acc[r][r];
    }
    if (cold) {
        if (tid < pp) g.v[tid] = tid < p ? 1.0 : 0.0;
    }
    __syncthreads();
    int steps = 0, ok = 0;
    double prev = 1.0e300;
    double vi = tid < pp ? g.v[tid] : 0.0;
    for (; steps < EIG_FAST_STEPS + 1;) {
        tile_matvec(g, t, acc);
        __syncthreads();
        const double y = tid < pp ? matvec_row(g, tid) : 0.0;
        ++steps;
        const double n2 = block_sum<WNT>(y * y, g.red);
        if (!(n2 > 0.0)) {
            if (tid < pp) g.v[tid] = 0.0;
            __syncthreads();
            ok = 1;
            break;
        }
        const double w = y * (1.0 / sqrt(n2));
        const double d = block_max<WNT>(tid < pp ? fabs(w - vi) : 0.0, g.red);
        vi = w;
        if (tid < pp) g.v[tid] = w;
        __syncthreads();
        if (d <= EIG_TOL) {
            const double m = (tid < p && g.diag[tid] > 0.0) ? vi : 1.0;
            const double vmin = -block_max<WNT>(-m, g.red);
            const double vmax = block_max<WNT>(tid < pp ? vi : 0.0, g.red);
            ok = vmin < EIG_SUSPECT * vmax ? 2 : 1;
            break;
        }
        if (steps >= 9 && d > 0.75 * prev) { ok = 3; break; }
        prev = d;
    }
    g.eig_steps += steps;
    if (ok != 1) {                                 // uniform across the CTA (and the cluster: same G everywhere)
        // materialise G (both triangles) in the slab for the squaring solver
        if (t.owner) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int i = 8 * t.ti + rot8(r, t.ti), j = 8 * t.tj + rot8(q, t.tj);
                    g.Gfull[(long long)i * pp + j] = acc[r][q];
                    g.Gfull[(long long)j * pp + i] = acc[r][q];
                }
        }
        __syncthreads();
        g.eig_steps += eig_squaring<WNT>(g.Gfull, pp, p, g.v, g.red, g.Gfull + (long long)pp * pp,
                                         g.Gfull + 2ll * pp * pp, ok == 2);
        g.eig_fallbacks += 1;
    }
}

// ---- row sums of the first n columns of a slab array (16 lanes per column, rows hl, hl + 16, ...) -> out[pp] -------
__device__ void rowsum_wide(WGene &g, const double *A, int n, double *out) {
    const int tid = threadIdx.x, grp = tid / WLPC, hl = tid % WLPC;
    const int CS = g.cs_col, pp = g.pp;
    double rs[WRPL];
#pragma unroll
    for (int k = 0; k < WRPL; ++k) rs[k] = 0.0;
    for (int col = grp; col < n; col += WCPR) {
        const double *xc = A + (long long)col * CS;
#pragma unroll
        for (int k = 0; k < WRPL; ++k) {
            const int r = hl + WLPC * k;
            if (r < pp) rs[k] += xc[r];
        }
    }
    double *scr = g.ring;                                   // the ring is idle outside the passes
    __syncthreads();
#pragma unroll
    for (int k = 0; k < WRPL; ++k) {
        const int r = hl + WLPC * k;
        if (r < pp) scr[grp * pp + r] = rs[k];
    }
    __syncthreads();
    if (tid < pp) {
        double s = 0.0;
        for (int q = 0; q < WCPR; ++q) s += scr[q * pp + tid];
        out[tid] = s;
    }
    __syncthreads();
}

// ---- final pass of an nmf() call: t_j = v . M_j, residuals, row sums (nmf.py:247-254, 280-283, 312-321) ------------
__device__ void final_pass_wide(const KArgs &a, WGene &g, bool first, bool want_res, double *e_first_g) {
    const int tid = threadIdx.x, grp = tid / WLPC, hl = tid % WLPC;
    const int n = g.n_cur, CS = g.cs_col, pp = g.pp;
    double vr[WRPL], sF[WRPL], sC[WRPL];
#pragma unroll
    for (int k = 0; k < WRPL; ++k) {
        const int r = hl + WLPC * k;
        vr[k] = r < pp ? g.v[r] : 0.0;
        sF[k] = 0.0;
        sC[k] = 0.0;
    }
    double st = 0.0, st2 = 0.0;
    const int nround = (n + WCPR - 1) / WCPR;
    for (int rd = 0; rd < nround; ++rd) {
        const int col = rd * WCPR + grp;
        double m[WRPL], x[WRPL];
        double tp = 0.0;
        if (col < n) {
            const double *mc = g.M + (long long)col * CS, *xc = g.X + (long long)col * CS;
#pragma unroll
            for (int k = 0; k < WRPL; ++k) {
                const int r = hl + WLPC * k;
                m[k] = r < pp ? __ldcg(mc + r) : 0.0;      // (M was written by the async proxy: bypass L1)
                x[k] = r < pp ? xc[r] : 0.0;
                tp = fma(vr[k], m[k], tp);
            }
        }
        double t = tp;
#pragma unroll
        for (int o = 1; o < WLPC; o <<= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        double r2 = 0.0;
        if (col < n) {
            double bn = 0.0, bd = 1.0;                      // largest |num| / den by cross-multiplication: one division
#pragma unroll
            for (int k = 0; k < WRPL; ++k) {
                const double ke = vr[k] * t;
                const double kc = ke < x[k] ? x[k] : ke;
                sF[k] += x[k];
                sC[k] += kc;
                if (want_res) {
                    const double num = fabs((first ? ke : kc) - x[k]), den = x[k] + 1.0;
                    if (num * bd > bn * den) { bn = num; bd = den; }
                }
            }
            const double qv = bn / bd;
            r2 = qv * qv;
        }
#pragma unroll
        for (int o = 1; o < WLPC; o <<= 1) r2 = fmax(r2, __shfl_xor_sync(0xffffffffu, r2, o));
        if (col < n && hl == 0) {
            g.tb[col] = t;
            if (want_res) g.resb[col] = r2;
            st += t;
            st2 = fma(t, t, st2);
        }
    }
    const double sum_t_l = block_sum<WNT>(st, g.red);
    const double sum_t2_l = block_sum<WNT>(st2, g.red);
    double *scr = g.ring;                                   // [group][2][pp], then the 2 pp + 2 totals
#pragma unroll
    for (int k = 0; k < WRPL; ++k) {
        const int r = hl + WLPC * k;
        if (r < pp) {
            scr[(grp * 2 + 0) * pp + r] = sF[k];
            scr[(grp * 2 + 1) * pp + r] = sC[k];
        }
    }
    __syncthreads();
    double *vals = scr + WCPR * 2 * pp;
    for (int e = tid; e < 2 * pp; e += WNT) {
        const int which = e / pp, i = e - which * pp;
        double s = 0.0;
        for (int q = 0; q < WCPR; ++q) s += scr[(q * 2 + which) * pp + i];
        vals[e] = s;
    }
    if (tid == 0) { vals[2 * pp] = sum_t_l; vals[2 * pp + 1] = sum_t2_l; }
    __syncthreads();
    wclu_allsum(g, vals, 2 * pp + 2, vals);
    const double sum_t = vals[2 * pp], sum_t2 = vals[2 * pp + 1];
    const double sigma = sqrt(sum_t2);
    if (e_first_g != nullptr) {
        const double inv = sigma > 0.0 ? 1.0 / sigma : 0.0;
        for (int col = tid; col < n; col += WNT) e_first_g[g.goff + col] = g.tb[col] * inv;
    }
    if (tid < pp) {
        const double vi = g.v[tid];
        g.rsF[tid] = vals[tid];
        g.rsC[tid] = vals[pp + tid];
        g.tmp[tid] = vi * sum_t;
        g.K[tid] = vi * sigma;
    }
    __syncthreads();
}

__device__ void run_nmf_wide(const KArgs &a, WGene &g, const WTile &t, bool first, bool want_res, double *e_first_g) {
    const int tid = threadIdx.x;
    {   // lambda = 0: M = x
        const double2 *src = reinterpret_cast<const double2 *>(g.X);
        double2 *dst = reinterpret_cast<double2 *>(g.M);
        const long long n2 = (long long)g.n_cur * (g.cs_col / 2);
        for (long long e = tid; e < n2; e += WNT) dst[e] = src[e];
    }
    fence_proxy_async();                  // the slab was written with ordinary stores; the TMA reads it next
    __syncthreads();
    const int T = a.nmf_iter;
    g.primed = false;
    WPass pa;
    pa.ring = g.ring; pa.M = g.M; pa.X = g.X; pa.v = g.v; pa.slots = g.slots; pa.kslab = g.kslab; pa.mbar = g.mbar;
    pa.slot_stride = g.slot_stride; pa.c = a.c; pa.n_cur = g.n_cur; pa.pp = g.pp; pa.cs_col = g.cs_col;
    pa.ti = t.ti; pa.tj = t.tj; pa.ks = t.ks; pa.nks = g.ks; pa.crank = g.crank; pa.csize = g.csize; pa.ne = g.ne;
    pa.tile = t.tile;
    pa.ovl = (WIDE_OVERLAP && wide_slots(g.nb) * g.ks <= (WNW - 1) * 32) ? 1 : 0;       // the twelfth warp owns no tile
    double acc[8][8];
    for (int it = -1; it < T; ++it) {
        pa.seq = g.seq;
        pa.xpar = g.xpar;
        pa.primed = g.primed ? 1 : 0;
        const bool prime_next = it + 1 < T;
        const WPassOut po = it < 0 ? gram_wide<false>(pa, prime_next, acc) : gram_wide<true>(pa, prime_next, acc);
        g.seq = po.seq;
        g.xpar = po.xpar;
        g.primed = prime_next;
        eig_wide(a, g, t, acc, it < 0);
    }
    final_pass_wide(a, g, first, want_res, e_first_g);
}

__global__ void __launch_bounds__(WNT, 1) nmfoa_wide_kernel(const KArgs a) {
    extern __shared__ double smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int p = a.p, pp = a.pp;
    const WideCarve cv = wide_carve(pp);
    WGene g;
    double *sm = smem + cv.small;
    g.v = sm;            g.K = sm + pp;        g.K0 = sm + 2 * pp;   g.rs0 = sm + 3 * pp;   g.rsF = sm + 4 * pp;
    g.rsC = sm + 5 * pp; g.rsC0 = sm + 6 * pp; g.rho = sm + 7 * pp;  g.scale = sm + 8 * pp; g.tmp = sm + 9 * pp;
    g.diag = sm + 10 * pp;
    g.red = smem + cv.red;
    g.binm = smem + cv.binm;
    g.alive = reinterpret_cast<int *>(smem + cv.alive);
    g.ibuf = reinterpret_cast<int *>(smem + cv.ibuf);
    g.lw = reinterpret_cast<int *>(smem + cv.lw);
    g.part = smem + cv.part;
    g.ring = smem + cv.ring;
    g.mbar = reinterpret_cast<unsigned long long *>(smem + cv.mbar);
    g.pp = pp;
    g.cs_col = pp + 2;
    g.nb = pp / 8;
    g.ntiles = g.nb * (g.nb + 1) / 2;
    g.ne = g.ntiles * 64;
    g.ks = wide_kslices(g.nb);
    g.seq = 0;
    if (tid == 0) {
        for (int q = 0; q < WIDE_RING; ++q) mbar_init(g.mbar + q, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    cg::cluster_group cl = cg::this_cluster();
    g.crank = (int)cl.block_rank();
    g.csize = (int)cl.num_blocks();
    g.xpar = 0;
    const long long wcols = a.ws_ld;                           // columns one CTA can hold
    double *slab = a.ws + (long long)blockIdx.x * a.ws_stride;
    g.slot_stride = a.ws_stride;
    g.Gfull = slab;
    g.slots = slab + 3ll * pp * pp;
    g.kslab = g.slots + 2ll * g.ne;
    g.X = g.kslab + (long long)(WIDE_MAX_KS - 1) * g.ne;
    g.M = g.X + (long long)g.cs_col * wcols;
    g.resb = g.M + (long long)g.cs_col * wcols;
    g.tb = g.resb + wcols;
    WTile t;
    {   // this thread's Gram tile.  Lane slots run row-major over the upper triangle, every tile row padded to an even
        // number of slots (the pad slot stays idle): aligned lane pairs then share the tile row.  k-slice = tid / slots.
        const int nslots = wide_slots(g.nb);
        t.ks = tid / nslots;
        t.tile = -1;
        t.ti = 0; t.tj = 0;
        if (t.ks < g.ks) {
            int rest = tid - t.ks * nslots;
            for (int ti = 0; ti < g.nb; ++ti) {
                const int len = g.nb - ti, padded = len + (len & 1);
                if (rest < padded) {
                    if (rest < len) { t.ti = ti; t.tj = ti + rest; t.tile = tile_index(g.nb, ti, t.tj); }
                    break;
                }
                rest -= padded;
            }
        }
        t.owner = t.tile >= 0 && t.ks == 0;
    }
    for (int e = tid; e < WIDE_NSMALL * pp; e += WNT) sm[e] = 0.0;
    g.eig_steps = 0;
    g.eig_fallbacks = 0;
    __syncthreads();
    cl.sync();
    const int CS = g.cs_col;

    for (;;) {
        if (g.crank == 0 && tid == 0) {
            const int tk = atomicAdd(a.queue, 1);
            for (int r = 0; r < g.csize; ++r) *cl.map_shared_rank(g.ibuf, r) = tk;
        }
        cl.sync();
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

        if (tid < pp) g.scale[tid] = tid < p ? a.scale[tid] : 1.0;
        __syncthreads();
        double tmax = -1.0e300;
        if (a.row_max) {
            if (tid < p) tmax = a.row_max[(long long)gid * p + tid] / g.scale[tid];
        } else {
            for (int i = 0; i < p; ++i) {
                const double *row = F + (long long)i * L;
                double m = -1.0e300;
                for (int j = tid; j < L; j += WNT) m = fmax(m, row[j]);
                tmax = fmax(tmax, m / g.scale[i]);
            }
        }
        const double gmax = block_max<WNT>(tmax, g.red);
        const double thr = (a.flags & DN_FLAG_PLAIN_NMF) ? -1.0e300 : 0.1 * gmax;      // nmf.py:76
        const int rate = a.rate;
        const int start = (rate > 1 && a.ds_start) ? a.ds_start[gid] : 0;
        const int ncand = start < L ? (L - start + rate - 1) / rate : 0;
        const int share = (ncand + g.csize - 1) / g.csize;
        const int k_lo = min(g.crank * share, ncand), k_hi = min(k_lo + share, ncand);
        int exit_code = DN_EXIT_NONE;
        int ran = 0, nmf_calls = 0, sum_cols = 0;
        unsigned long long drops = 0ull;
        bool k_is_refined = false;
        int n0 = 0;
        g.goff = 0;
        if (share > wcols) {
            exit_code = -1;
        } else {
            // keep + compact (one thread per candidate column; the scaled values go straight to the slab)
            int running = 0;
            int *wcount = g.ibuf + 1;
            for (int kb = k_lo; kb < k_hi; kb += WNT) {
                const int k = kb + tid;
                bool keep = false;
                long long col = 0;
                if (k < k_hi) {
                    col = start + (long long)k * rate;
                    double cm = -1.0e300;
                    for (int i = 0; i < p; ++i) cm = fmax(cm, F[(long long)i * L + col] / g.scale[i]);
                    keep = cm > thr;
                }
                const unsigned bal = __ballot_sync(0xffffffffu, keep);
                if (lane == 0) wcount[warp] = __popc(bal);
                __syncthreads();
                int pre = running, tot = 0;
#pragma unroll
                for (int q = 0; q < WNW; ++q) {
                    const int cq = wcount[q];
                    if (q < warp) pre += cq;
                    tot += cq;
                }
                if (keep) {
                    const int dst = pre + __popc(bal & ((1u << lane) - 1u));
                    double *xc = g.X + (long long)dst * CS;
                    for (int i = 0; i < CS; ++i) xc[i] = i < p ? F[(long long)i * L + col] / g.scale[i] : 0.0;
                }
                running += tot;
                __syncthreads();
            }
            g.n0 = g.n_cur = running;
            n0 = running;
            if (tid < g.csize) g.binm[tid] = tid == g.crank ? (double)running : 0.0;
            __syncthreads();
            wclu_allsum(g, g.binm, g.csize, g.binm);
            n0 = 0;
            for (int r = 0; r < g.csize; ++r) {
                if (r == g.crank) g.goff = n0;
                n0 += (int)g.binm[r];
            }
            __syncthreads();
        }
        g.n0g = g.n_curg = n0;
        if (exit_code == -1) {
        } else if (n0 < a.min_hi) {
            exit_code = DN_EXIT_FEW_HICOV;                               // nmf.py:232-233
        } else {
            g.cs = n0; g.nb0 = 1; g.nalive = 1;
            if (tid == 0) { g.alive[0] = 0; g.lw[0] = g.n0; }
            rowsum_wide(g, g.X, g.n0, g.rs0);
            wclu_allsum(g, g.rs0, pp, g.rs0);
            bool any_empty = false;
            for (int i = 0; i < p; ++i) any_empty |= !(g.rs0[i] > 0.0);
            if (any_empty && !(a.flags & DN_FLAG_PLAIN_NMF)) {
                exit_code = DN_EXIT_EMPTY_SAMPLE;                        // nmf.py:241-242
            } else {
                const bool store_e = (a.e_first != nullptr) && (n0 == L);
                bool first = true, in_loop = false;
                double rmax = 0.0;
                for (;;) {
                    run_nmf_wide(a, g, t, first, true, (first && store_e) ? a.e_first + o0 : nullptr);
                    nmf_calls += 1; sum_cols += g.n_curg;
                    if (first) {
                        if (tid < pp) {
                            g.rho[tid] = 1.0 - g.rs0[tid] / (g.tmp[tid] + 1.0);
                            g.K0[tid] = g.K[tid];
                            g.rsC0[tid] = g.rsC[tid];
                        }
                        __syncthreads();
                        if (!(a.flags & DN_FLAG_PLAIN_NMF) && median_one_minus(g.rho, p) > 1.0) { exit_code = DN_EXIT_MEDIAN; break; }
                        double rmin = g.rho[0];
                        rmax = g.rho[0];
                        for (int i = 1; i < p; ++i) { rmin = fmin(rmin, g.rho[i]); rmax = fmax(rmax, g.rho[i]); }
                        if (!(n0 >= a.min_len && rmin <= 0.2 && !a.skip)) { exit_code = DN_EXIT_NO_SELECTION; break; }   // nmf.py:265
                        g.cs = (n0 + a.bins - 1) / a.bins;               // utils.py:176-192
                        g.nb0 = (n0 + g.cs - 1) / g.cs;
                        g.nalive = g.nb0;
                        for (int b = tid; b < g.nb0; b += WNT) {
                            g.alive[b] = b;
                            const int lo = max(b * g.cs, g.goff), hi = min(min((b + 1) * g.cs, n0), g.goff + g.n0);
                            g.lw[b] = max(0, hi - lo);
                        }
                        __syncthreads();
                        in_loop = true;
                        first = false;
                    } else {
                        double mn = g.tmp[0];
                        for (int i = 1; i < p; ++i) mn = fmin(mn, g.tmp[i]);
                        if (mn == 0.0) break;                            // nmf.py:314
                        __syncthreads();
                        if (tid < pp) g.rho[tid] = 1.0 - g.rsF[tid] / (g.rsC[tid] + 1.0);
                        __syncthreads();
                        rmax = g.rho[0];
                        for (int i = 1; i < p; ++i) rmax = fmax(rmax, g.rho[i]);
                        if (g.nalive <= a.min_bins || g.n_curg < a.min_len) break;      // nmf.py:323
                    }
                    if (!(rmax > 0.1)) break;                            // nmf.py:273
                    ran = 1;
                    for (int k = warp; k < g.nalive; k += WNW) {
                        const int wl = g.lw[g.alive[k]];
                        const double *rr = g.resb + wlstart(g, k);
                        double s = 0.0;
                        for (int j = lane; j < wl; j += 32) s += rr[j];
                        s = warp_sum(s);
                        if (lane == 0) g.binm[k] = s;
                    }
                    __syncthreads();
                    wclu_allsum(g, g.binm, g.nalive, g.binm);
                    int kd = 0;
                    double best = -1.0;
                    for (int k = 0; k < g.nalive; ++k) {
                        const int b = g.alive[k];
                        const double mean = g.binm[k] / (double)min(g.cs, n0 - b * g.cs);
                        if (mean > best) { best = mean; kd = k; }
                    }
                    if (best == 0.0) break;                              // nmf.py:286-287
                    const int bd = g.alive[kd];
                    const int wd = min(g.cs, n0 - bd * g.cs);
                    {   // rotate this CTA's share of the dropped bin to the end of its current columns (M is scratch)
                        const int a0 = wlstart(g, kd);
                        const int wl = g.lw[bd];
                        const int tail = g.n_cur - a0 - wl;
                        const int h = CS / 2;
                        const double2 *xs = reinterpret_cast<const double2 *>(g.X + (long long)a0 * CS);
                        double2 *ms = reinterpret_cast<double2 *>(g.M + (long long)a0 * CS);
                        double2 *xd = reinterpret_cast<double2 *>(g.X + (long long)a0 * CS);
                        __syncthreads();
                        for (long long e = tid; e < (long long)(tail + wl) * h; e += WNT) ms[e] = xs[e];
                        __syncthreads();
                        for (long long e = tid; e < (long long)tail * h; e += WNT) xd[e] = ms[(long long)wl * h + e];
                        for (long long e = tid; e < (long long)wl * h; e += WNT) xd[(long long)tail * h + e] = ms[e];
                        if (tid == 0)
                            for (int k = kd; k < g.nalive - 1; ++k) g.alive[k] = g.alive[k + 1];
                        g.n_cur -= wl;
                    }
                    __syncthreads();
                    g.nalive -= 1;
                    g.n_curg -= wd;
                    drops |= 1ull << bd;
                    if (g.n_curg < 2) break;
                }
                if (in_loop) {
                    __syncthreads();
                    bool fallback = true;
                    exit_code = DN_EXIT_FALLBACK;
                    if (rmax < 0.2) {                                    // nmf.py:327
                        floor_abs(g.K, g.K, p);
                        double s = 0.0;
                        for (int j = tid; j < g.n0; j += WNT) {
                            const double *xc = g.X + (long long)j * CS;
                            double e = -1.0e300;
                            for (int i = 0; i < p; ++i) e = fmax(e, xc[i] / g.K[i]);
                            s += e;
                        }
                        double S = block_sum<WNT>(s, g.red);
                        if (tid == 0) g.binm[0] = S;
                        __syncthreads();
                        wclu_allsum(g, g.binm, 1, g.binm);
                        S = g.binm[0];
                        __syncthreads();
                        if (tid < pp) g.rho[tid] = 1.0 - g.rs0[tid] / (g.K[tid] * S + 1.0);
                        __syncthreads();
                        rmax = g.rho[0];
                        for (int i = 1; i < p; ++i) rmax = fmax(rmax, g.rho[i]);
                        if (rmax > 0.9) {
                            exit_code = DN_EXIT_FALLBACK_HIGH;           // nmf.py:342-346
                        } else {
                            exit_code = DN_EXIT_REFINED;
                            fallback = false;
                            k_is_refined = true;
                        }
                    }
                    if (fallback) {
                        __syncthreads();
                        if (tid < pp) g.rho[tid] = 1.0 - g.rs0[tid] / (g.rsC0[tid] + 1.0);
                        __syncthreads();
                    }
                }
            }
        }
        __syncthreads();
        const bool is_default = exit_code == DN_EXIT_FEW_HICOV || exit_code == DN_EXIT_EMPTY_SAMPLE ||
                                exit_code == DN_EXIT_MEDIAN || exit_code == -1;
        if (!is_default && !k_is_refined) {
            if (n0 == L) {
                if (tid < pp) g.K[tid] = g.K0[tid];
                __syncthreads();
            } else {
                floor_abs(g.K0, g.K, p);
            }
        }
        if (g.crank == 0) {
            if (tid < p) {
                double r = is_default ? 0.0 : g.rho[tid];
                if (!(a.flags & DN_FLAG_RAW_RHO)) r = r > 0.9 ? 0.9 : r;      // nmf.py:398-399
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
                    cnt[DN_CNT_RESIDENT] = (g.eig_fallbacks << 1);
                }
            }
        }
        __syncthreads();
        if (a.est) {
            write_estimate(F, L, p, g.scale, exit_code, n0, g.K, a.e_first ? a.e_first + o0 : nullptr,
                           a.est + (long long)p * (a.est_off ? a.est_off[gid] : o0), g.crank * WNT + tid, g.csize * WNT);
            __syncthreads();
        }
        cl.sync();
    }
}

}  // namespace

int dn_launch_wide(const KArgs &a, const dn_plan *plan, cudaStream_t st) {
    auto kern = nmfoa_wide_kernel;
    DN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, plan->smem_bytes));
    if (plan->cluster > 8) DN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    const int cl = plan->cluster > 0 ? plan->cluster : 1;
    cfg.gridDim = dim3(plan->ctas / cl * cl, 1, 1);
    cfg.blockDim = dim3(WNT, 1, 1);
    cfg.dynamicSmemBytes = plan->smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cl;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    DN_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
    return DN_OK;
}
