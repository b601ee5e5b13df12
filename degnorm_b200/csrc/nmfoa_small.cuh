// degnorm_b200 -- fused baseline-selection kernel for few samples (p <= 12), sm_100a.
//
// Same per-gene flow as nmfoa_tiled.cu (reference: /root/reference/degnorm/nmf.py:189-372 around nmf.py:78-107),
// re-laid-out for the regime where a gene has tens to a few hundred kept columns and the cost is instruction
// latency, not bytes:
//
//   * x and M = x + lambda live column-major with a padded column stride (P + 2 doubles), in shared memory
//     (resident tiers) or in a per-CTA global slab (streamed tier).  lambda itself is never stored:
//     lambda = M - x is recovered on load (exact when lambda = 0; otherwise within half an ulp of M).
//   * one inner iteration (nmf.py:93-98) = phase A, one lane per column: t = v.M_j, lambda update, new M_j;
//     phase B, lanes own 2 x TC tiles of the Gram matrix and sweep the columns their own warp just wrote
//     (operands come straight from shared memory: no cross-lane reduction of P(P+1)/2 partial sums);
//     a k-slice reduction over 2..8 lanes, a cross-warp sum for multi-warp CTAs; then the P x P eigen-solve.
//   * the eigen-solve keeps the matrix row and the whole vector in registers; steps exchange the new vector
//     through a warp-private shared buffer (one store, one __syncwarp, P/2 128-bit loads).  Warm solves run
//     `hint - 1` un-normalised steps blind (scaled by the last 1/lambda_1) before the first checked step; the hint
//     adapts to the step count the previous solve of the same nmf() call needed.  Every warp of a CTA solves
//     redundantly, so no barrier separates the solve from the next phase A.
//   * dropping a bin physically rotates its columns to the end of the buffer, so the current matrix is always the
//     contiguous range [0, n_cur): no per-column index mapping in the inner loop.
//
//   * long genes (CLU = true): a thread-block CLUSTER owns the gene.  Each CTA keeps a contiguous slice of the
//     kept columns in its own shared memory (or slab) and runs phases A/B on it; the only cross-CTA traffic in
//     the inner iteration is the 96-double partial Gram, written straight into every peer's shared memory
//     (DSMEM) and followed by ONE cluster barrier; every CTA then sums the slots in rank order and solves
//     redundantly, so all CTAs take bit-identical decisions.  Row sums, residual bin sums and column counts
//     use the same all-to-all exchange.
//
// No tensor cores: fp64 FMA pipe only.
#pragma once
#include <cooperative_groups.h>
#include "common.cuh"
#include "launch.h"
namespace cg = cooperative_groups;

namespace {

// Stopping test of the eigen-solve: the last step moved no entry of v by more than this.  A step contracts the
// error by lambda_2/lambda_1 (typically 0.05-0.2), so the vector returned is within ~1e-14 of the eigenvector.
constexpr double SMALL_EIG_TOL = 1.0e-13;

// streamed tier: Gram rounds of one 32-column block whose operand loads are issued back to back (tuning knob; the
// order of the sums does not depend on it)
#ifndef SMALL_STREAM_UNROLL
#define SMALL_STREAM_UNROLL 4
#endif
constexpr int STREAM_UNROLL = SMALL_STREAM_UNROLL;
#ifndef SMALL_RES_UNROLL      // the same for the shared-memory resident tiers
#define SMALL_RES_UNROLL 4
#endif
constexpr int RES_UNROLL = SMALL_RES_UNROLL;

template <int P> struct SmallCfg;
// TC: tile columns (tile = 2 rows x TC columns of G), NTILE tiles cover the upper triangle, NTP = tiles padded
// to a power of two, KS = k-slices (lanes = NTP * KS)
template <> struct SmallCfg<12> { static constexpr int TC = 3, NTILE = 16, NTP = 16, KS = 2; };
template <> struct SmallCfg<8> { static constexpr int TC = 2, NTILE = 10, NTP = 16, KS = 2; };
template <> struct SmallCfg<4> { static constexpr int TC = 2, NTILE = 3, NTP = 4, KS = 8; };

template <int NW>
__device__ __forceinline__ void bsync() {
    if constexpr (NW == 1) __syncwarp();
    else __syncthreads();
}

struct SGene {
    double *v, *K, *K0, *rs0, *rsF, *rsC, *rsC0, *rho, *scale, *tmp, *red, *binm, *G, *vx, *gpart;
    int *alive, *ibuf, *tab;
    double *X, *M, *resb, *tb;
    int n0, n_cur, cs, nb0, nalive;
    int eig_steps, eig_fallbacks, gpar;
    double *B0;
    // cluster state (CLU kernels; a lone CTA is rank 0 of 1): n0 / n_cur are this CTA's columns, n0g / n_curg the gene's
    double *stage;      // streamed tier: one 32-column stage of M per warp
    double *ring;       // streamed tier: per-warp cp.async ring of x / M blocks
    double *xbuf;       // exchange slots: 2 alternating sets of SMALL_CLMAX x SMALL_GPART doubles (peers write here)
    int *lw;            // this CTA's width of every original bin
    int crank, csize, xpar, n0g, n_curg, goff;
    bool primed;        // streamed tier: the ring already holds / awaits the first blocks of the next pass
    int tpack;      // this lane's Gram tile: r0 | offA << 4 | offB << 8 | u0 << 12 | u1 << 16 | u2 << 20; -1: none
};

template <int P>
__device__ __forceinline__ void tile_of(int t, int &r0, int &c0, bool &ok) {
    constexpr int TC = SmallCfg<P>::TC;
    int idx = 0;
    ok = false; r0 = 0; c0 = 0;
    for (int rb = 0; rb < P / 2; ++rb)
        for (int cb = 0; cb < P / TC; ++cb)
            if (cb * TC + TC - 1 >= 2 * rb) {
                if (idx == t) { r0 = 2 * rb; c0 = cb * TC; ok = true; }
                ++idx;
            }
}

// ---- cp.async (LDGSTS) helpers: 16-byte global -> shared copies that bypass registers and L1 --------------------
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

template <int P>
__device__ __forceinline__ void load_col(const double *base, double (&x)[P]) {
    const double2 *q = reinterpret_cast<const double2 *>(base);
#pragma unroll
    for (int i = 0; i < P / 2; ++i) {
        const double2 t = q[i];
        x[2 * i] = t.x;
        x[2 * i + 1] = t.y;
    }
}

// Column accessors.  Resident tiers (RES): shared memory, column-major, stride P+2 doubles.  Streamed tier: the
// CTA's global slab, "blocked row-major": 32-column blocks, inside a block row i holds its 32 columns contiguously,
// so a warp reading one column per lane issues fully coalesced 256-byte requests (no padding bytes either).
template <int P, bool RES>
__device__ __forceinline__ void ld_col(const double *base, int col, double (&x)[P]) {
    if constexpr (RES) {
        load_col<P>(base + col * (P + 2), x);
    } else {
        const double *q = base + (long long)(col >> 5) * (32 * P) + (col & 31);
#pragma unroll
        for (int i = 0; i < P; ++i) x[i] = q[i * 32];
    }
}
template <int P, bool RES>
__device__ __forceinline__ void st_col(double *base, int col, const double (&x)[P]) {
    if constexpr (RES) {
        double2 *q = reinterpret_cast<double2 *>(base + col * (P + 2));
#pragma unroll
        for (int i = 0; i < P / 2; ++i) q[i] = make_double2(x[2 * i], x[2 * i + 1]);
    } else {
        double *q = base + (long long)(col >> 5) * (32 * P) + (col & 31);
#pragma unroll
        for (int i = 0; i < P; ++i) q[i * 32] = x[i];
    }
}

template <int P>
__device__ __forceinline__ double dot_v(const double (&v)[P], const double (&m)[P]) {
    double t0 = 0.0, t1 = 0.0;
#pragma unroll
    for (int i = 0; i < P; i += 2) {
        t0 = fma(v[i], m[i], t0);
        t1 = fma(v[i + 1], m[i + 1], t1);
    }
    return t0 + t1;
}

// All-to-all sum over the cluster of n <= SMALL_GPART doubles (`vals`: local shared memory, already published
// to the CTA): every CTA writes its values into its slot in every peer (DSMEM), one cluster barrier, every CTA adds
// the slots in rank order -> bit-identical `out` everywhere.  Slot sets alternate, so one barrier per call is enough.
template <int NT>
__device__ __forceinline__ void clu_allsum(SGene &g, const double *vals, int n, double *out) {
    cg::cluster_group cl = cg::this_cluster();
    double *buf = g.xbuf + g.xpar * (SMALL_CLMAX * SMALL_GPART);
    g.xpar ^= 1;
    for (int e = threadIdx.x; e < n * g.csize; e += NT) {
        const int r = e / n, k = e - r * n;
        *cl.map_shared_rank(buf + g.crank * SMALL_GPART + k, r) = vals[k];
    }
    cl.sync();
    for (int k = threadIdx.x; k < n; k += NT) {
        double s = 0.0;
        for (int r = 0; r < g.csize; ++r) s += buf[r * SMALL_GPART + k];
        out[k] = s;
    }
    __syncthreads();
}

// local start of alive bin k in this CTA's current columns
__device__ __forceinline__ int lstart(const SGene &g, int k) {
    int s = 0;
    for (int q = 0; q < k; ++q) s += g.lw[g.alive[q]];
    return s;
}

// sums NV per-thread values over the CTA (fixed tree: warp shuffles, then warps in order) into out[0..NV)
template <int NV, int NW>
__device__ __forceinline__ void block_sum_vec(double (&x)[NV], double *part, double *out) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) x[k] = warp_sum(x[k]);
    if constexpr (NW == 1) {
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < NV; ++k) out[k] = x[k];
        }
        __syncwarp();
    } else {
        __syncthreads();          // `part` doubles as the Gram partial-sum buffer: every warp must be done reading it
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < NV; ++k) part[warp * NV + k] = x[k];
        }
        __syncthreads();
        for (int k = tid; k < NV; k += NW * 32) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < NW; ++w) s += part[w * NV + k];
            out[k] = s;
        }
        __syncthreads();
    }
}

// ---- one pass: (optional multiplier update of this warp's columns) + Gram of M -----------------------------------
// UPDATE=false: G = M M^T with M = x (first rank-one fit, nmf.py:88).
// UPDATE=true : lambda <- max(0, lambda - c (K E - x)), M = x + lambda, G = M M^T (nmf.py:93-98); K E = v (v.M_old).
template <int P, int NW, bool UPDATE, bool CLU, bool RES>
__device__ __forceinline__ void gram_small(const KArgs &a, SGene &g, const double (&v)[P], bool prime_next) {
    using Cfg = SmallCfg<P>;
    constexpr int TC = Cfg::TC, KS = Cfg::KS, NTP = Cfg::NTP, NTILE = Cfg::NTILE, CS = P + 2, NT = NW * 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = g.n_cur;
    const double c = a.c;
    double acc[2][TC];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int q = 0; q < TC; ++q) acc[r][q] = 0.0;

    // this warp's contiguous column range (whole 32-column blocks)
    const int per_warp = ((n + 31) / 32 + NW - 1) / NW * 32;
    const int c_lo = min(warp * per_warp, n), c_hi = min(c_lo + per_warp, n);
    const int oR = g.tpack & 15, oA = (g.tpack >> 4) & 15, oB = (g.tpack >> 8) & 15;
    const int ks = lane / NTP;
    if constexpr (!RES) {
        // Streamed tier: x and M live in the CTA's global slab (blocked row-major).  Each warp streams its 32-column
        // blocks through a private shared-memory ring filled by cp.async (SMALL_RING stages, so SMALL_RING - 1
        // blocks = 6 KB per array pair are in flight per warp while it computes): phase A reads the landed block,
        // writes the new M to the slab (coalesced) and to a column-major stage, phase B runs out of the stage.
        // Global traffic per column-iteration is exactly read x, read M, write M.
        constexpr int BLK = 32 * P;                              // doubles per block per array
        double *stage = g.stage + warp * (32 * CS);
        double *ring = g.ring + warp * (SMALL_RING * 2 * BLK);
        const int b_lo = c_lo >> 5, b_hi = c_hi > c_lo ? (c_hi + 31) >> 5 : b_lo;     // (c_lo is block-aligned unless the range is empty)
        // with_x: the consumer of the block is an UPDATE pass (needs x as well as M)
        auto issue = [&](int blk, bool with_x) {
            if (blk < b_hi) {
                double *dst = ring + (blk % SMALL_RING) * (2 * BLK);
                const double *srcM = g.M + (long long)blk * BLK;
#pragma unroll
                for (int c = 0; c < BLK / 64; ++c) cp_async16(dst + (c * 32 + lane) * 2, srcM + (c * 32 + lane) * 2);
                if (with_x) {
                    const double *srcX = g.X + (long long)blk * BLK;
#pragma unroll
                    for (int c = 0; c < BLK / 64; ++c)
                        cp_async16(dst + BLK + (c * 32 + lane) * 2, srcX + (c * 32 + lane) * 2);
                }
            }
            cp_async_commit();
        };
        // the first RING-1 blocks may already be on their way: the previous pass of this nmf() call primed them
        if (!g.primed) {
#pragma unroll
            for (int q = 0; q < SMALL_RING - 1; ++q) issue(b_lo + q, UPDATE);
        }
        for (int blk = b_lo; blk < b_hi; ++blk) {
            issue(blk + SMALL_RING - 1, UPDATE);
            cp_async_wait<SMALL_RING - 1>();
            __syncwarp();
            const double *rm = ring + (blk % SMALL_RING) * (2 * BLK) + lane;
            const int col = blk * 32 + lane;
            if (col < c_hi) {
                double m[P];
#pragma unroll
                for (int i = 0; i < P; ++i) m[i] = rm[i * 32];
                if constexpr (UPDATE) {
                    double x[P];
#pragma unroll
                    for (int i = 0; i < P; ++i) x[i] = rm[BLK + i * 32];
                    const double t = dot_v<P>(v, m);
#pragma unroll
                    for (int i = 0; i < P; ++i) {
                        const double res = fma(v[i], t, -x[i]);
                        const double w = fma(-c, res, m[i] - x[i]);
                        m[i] = fma(0.5, w + fabs(w), x[i]);
                    }
                    st_col<P, false>(g.M, col, m);
                }
                double2 *sq = reinterpret_cast<double2 *>(stage + lane * CS);
#pragma unroll
                for (int i = 0; i < P / 2; ++i) sq[i] = make_double2(m[2 * i], m[2 * i + 1]);
            }
            __syncwarp();
            if (g.tpack >= 0) {
                const int nb = min(32, c_hi - blk * 32);
                const double *mc = stage + ks * CS;
#pragma unroll STREAM_UNROLL
                for (int j = ks; j < nb; j += KS, mc += KS * CS) {
                    const double2 ar = *reinterpret_cast<const double2 *>(mc + oR);
                    const double2 ua = *reinterpret_cast<const double2 *>(mc + oA);
                    acc[0][0] = fma(ar.x, ua.x, acc[0][0]);
                    acc[0][1] = fma(ar.x, ua.y, acc[0][1]);
                    acc[1][0] = fma(ar.y, ua.x, acc[1][0]);
                    acc[1][1] = fma(ar.y, ua.y, acc[1][1]);
                    if constexpr (TC == 3) {
                        const double ub = mc[oB];
                        acc[0][2] = fma(ar.x, ub, acc[0][2]);
                        acc[1][2] = fma(ar.y, ub, acc[1][2]);
                    }
                }
            }
            __syncwarp();
        }
        // Prime the next pass: its first blocks were written early in this one (or are the untouched x = M of the
        // first fit), so they can travel while the Gram exchange and the eigen-solve run.
        g.primed = prime_next;
        if (prime_next) {
            __syncwarp();
#pragma unroll
            for (int q = 0; q < SMALL_RING - 1; ++q) issue(b_lo + q, true);
        }
    } else {
    if constexpr (UPDATE) {
        // phase A: one lane per column; columns are independent, two in flight per lane
#pragma unroll 2
        for (int col = c_lo + lane; col < c_hi; col += 32) {
            double x[P], m[P];
            load_col<P>(g.X + col * CS, x);
            load_col<P>(g.M + col * CS, m);
            const double t = dot_v<P>(v, m);
#pragma unroll
            for (int i = 0; i < P; ++i) {
                const double res = fma(v[i], t, -x[i]);        // est - x
                const double w = fma(-c, res, m[i] - x[i]);    // lambda - c (est - x)
                // x + max(0, w), branch- and select-free: w + |w| = 2 max(0, w) exactly, the fma rounds once
                m[i] = fma(0.5, w + fabs(w), x[i]);
            }
            double2 *mq = reinterpret_cast<double2 *>(g.M + col * CS);
#pragma unroll
            for (int i = 0; i < P / 2; ++i) mq[i] = make_double2(m[2 * i], m[2 * i + 1]);
        }
        __syncwarp();
    }
    if (g.tpack >= 0) {
        // phase B: this lane's Gram tile over its k-slice of the warp's columns
        const double *mc = g.M + (c_lo + ks) * CS;
#pragma unroll RES_UNROLL
        for (int col = c_lo + ks; col < c_hi; col += KS, mc += KS * CS) {
            const double2 ar = *reinterpret_cast<const double2 *>(mc + oR);
            const double2 ua = *reinterpret_cast<const double2 *>(mc + oA);
            acc[0][0] = fma(ar.x, ua.x, acc[0][0]);
            acc[0][1] = fma(ar.x, ua.y, acc[0][1]);
            acc[1][0] = fma(ar.y, ua.x, acc[1][0]);
            acc[1][1] = fma(ar.y, ua.y, acc[1][1]);
            if constexpr (TC == 3) {
                const double ub = mc[oB];
                acc[0][2] = fma(ar.x, ub, acc[0][2]);
                acc[1][2] = fma(ar.y, ub, acc[1][2]);
            }
        }
    }
    }
    // k-slice reduction (lane = ks * NTP + tile)
#pragma unroll
    for (int o = NTP; o < 32; o <<= 1)
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int q = 0; q < TC; ++q) acc[r][q] += __shfl_xor_sync(0xffffffffu, acc[r][q], o);

    if constexpr (NW == 1) {
        if (lane < NTP && g.tpack >= 0) {
            const int uu[3] = {(g.tpack >> 12) & 15, (g.tpack >> 16) & 15, (g.tpack >> 20) & 15};
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int q = 0; q < TC; ++q) {
                    const int i = (g.tpack & 15) + r, j = uu[q];
                    if (i <= j) {
                        g.G[i * P + j] = acc[r][q];
                        g.G[j * P + i] = acc[r][q];
                    }
                }
        }
        __syncwarp();
    } else {
        // one barrier per pass: partials go to one of two alternating buffers, then every warp sums all of them
        // (same order everywhere) into its own copy of G
        double *gp = g.gpart + (long long)g.gpar * (NW * SMALL_GPART);
        g.gpar ^= 1;
        if (lane < NTP && g.tpack >= 0) {
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int q = 0; q < TC; ++q) gp[warp * SMALL_GPART + lane * (2 * TC) + r * TC + q] = acc[r][q];
        }
        __syncthreads();
        const double *src = gp;
        int nsrc = NW;
        if constexpr (CLU) {
            // CTA sum of the warp partials, written straight into this CTA's slot in every peer; one cluster barrier
            cg::cluster_group cl = cg::this_cluster();
            double *buf = g.xbuf + g.xpar * (SMALL_CLMAX * SMALL_GPART);
            g.xpar ^= 1;
            for (int e = tid; e < NTILE * 2 * TC; e += NT) {
                double s = 0.0;
#pragma unroll
                for (int w = 0; w < NW; ++w) s += gp[w * SMALL_GPART + e];
                for (int r = 0; r < g.csize; ++r) *cl.map_shared_rank(buf + g.crank * SMALL_GPART + e, r) = s;
            }
            cl.sync();
            src = buf;
            nsrc = g.csize;
        }
        double *Gw = g.G + warp * (P * P);
        for (int e = lane; e < NTILE * 2 * TC; e += 32) {
            double s = 0.0;
            if constexpr (CLU) {
                for (int w = 0; w < nsrc; ++w) s += src[w * SMALL_GPART + e];
            } else {
#pragma unroll
                for (int w = 0; w < NW; ++w) s += gp[w * SMALL_GPART + e];
            }
            const int t = e / (2 * TC), rq = e - t * (2 * TC);
            const int r = rq / TC, q = rq - r * TC;
            const int i = g.tab[t * 4] + r, j = g.tab[t * 4 + 1 + q];
            if (i <= j) {
                Gw[i * P + j] = s;
                Gw[j * P + i] = s;
            }
        }
        __syncwarp();
    }
}

// ---- top eigenvector of G (P x P in shared memory), every warp redundantly ------------------------------------
// v: whole vector in registers (in: warm start unless cold; out: unit-norm eigenvector, or 0 for a zero matrix).
// inv_lam: 1 / lambda_1 from the last checked step; hint: steps the previous solve needed.
template <int P, int NW>
__device__ __forceinline__ int eig_small_core(const KArgs &a, SGene &g, double (&v)[P], bool cold, double &inv_lam, int &hint) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *vx = g.vx + warp * (2 * P);
    const int row = lane % P;
    const double *Gw = g.G + (NW > 1 ? warp * (P * P) : 0);
    double grow[P];
    load_col<P>(Gw + row * P, grow);
    int par = 0;
    // y = scale * G v, exchanged through the warp-private buffer (double-buffered: one __syncwarp per step)
    auto step = [&](double scale) {
        const double y = dot_v<P>(grow, v) * scale;
        if (lane < P) vx[par * P + lane] = y;
        __syncwarp();
        load_col<P>(vx + par * P, v);
        par ^= 1;
    };
    int steps = 0, ok = 0, checked = 0;
    if (cold) {
        // start from G.1 (row sums): positive for non-negative G, close to the Perron vector for near-rank-1 data
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < P; ++k) s += grow[k];
        if (lane < P) vx[par * P + lane] = s;
        __syncwarp();
        load_col<P>(vx + par * P, v);
        par ^= 1;
    } else if (a.eig_hint && hint > 1) {
        for (int b = 0; b < hint - 1; ++b) step(inv_lam);
        steps = hint - 1;
    }
    if (cold || steps > 0) {
        double n2 = 0.0;
#pragma unroll
        for (int k = 0; k < P; ++k) n2 = fma(v[k], v[k], n2);
        const double inv = n2 > 0.0 ? rsqrt(n2) : 0.0;
#pragma unroll
        for (int k = 0; k < P; ++k) v[k] *= inv;
    }
    double prev = 1.0e300, d = 0.0;
    for (; steps < EIG_FAST_STEPS;) {
        double old[P];
#pragma unroll
        for (int k = 0; k < P; ++k) old[k] = v[k];
        step(1.0);
        ++steps;
        ++checked;
        double n2 = 0.0;
#pragma unroll
        for (int k = 0; k < P; ++k) n2 = fma(v[k], v[k], n2);
        if (!(n2 > 0.0)) {            // all-zero matrix: the reference raises ArpackError here
#pragma unroll
            for (int k = 0; k < P; ++k) v[k] = 0.0;
            ok = 1;
            break;
        }
        const double inv = rsqrt(n2);
        // d = max_k |v_k - old_k|, tracked on the high words (|x| as an integer orders like |x|): ~2^-20 relative
        // resolution, which is plenty for a stopping test
        int hd = 0;
#pragma unroll
        for (int k = 0; k < P; ++k) {
            v[k] *= inv;
            hd = max(hd, __double2hiint(v[k] - old[k]) & 0x7fffffff);
        }
        d = __hiloint2double(hd, 0);
        inv_lam = inv;
        if (d < SMALL_EIG_TOL) { ok = 1; break; }
        if (checked >= 2 && steps >= 8 && d > 0.75 * prev) break;     // small spectral gap: squaring solver
        prev = d;
    }
    if (ok == 1) {
        // warm-start distrust rule (see nmfoa_tiled.cu eig_warp): an entry ~0 on a sample that has coverage.
        // v >= 0, so high words order like the values; a ratio below 1e-3 needs exponents >= 9 apart.
        int hmin = 0x7fffffff, hmax = 0;
#pragma unroll
        for (int k = 0; k < P; ++k) {
            const int h = __double2hiint(v[k]);
            hmax = max(hmax, h);
            hmin = min(hmin, k < a.p ? h : 0x7fffffff);
        }
        if (hmax - hmin >= (8 << 20)) {
            double vmin = 1.0e300, vmax = 0.0;
#pragma unroll
            for (int k = 0; k < P; ++k) {
                vmax = fmax(vmax, v[k]);
                if (k < a.p && Gw[k * P + k] > 0.0) vmin = fmin(vmin, v[k]);
            }
            if (vmin < EIG_SUSPECT * vmax) ok = 2;
        }
    }
    g.eig_steps += steps;
    if (ok == 1) {
        if (checked == 1 && d < 0.02 * SMALL_EIG_TOL) hint = steps > 1 ? steps - 1 : 1;
        else hint = steps;
    }
    return ok;
}

// Every warp solves redundantly (no barrier before the next phase A), or -- KArgs.eig_shared, multi-warp CTAs --
// warp 0 solves alone and hands v over through shared memory (one more barrier, but the other warps' issue slots
// go to the SM's other CTAs).  The small-gap / distrust fallback is block-wide either way.
template <int P, int NW>
__device__ __forceinline__ void eig_small(const KArgs &a, SGene &g, double (&v)[P], bool cold, double &inv_lam, int &hint) {
    constexpr int NT = NW * 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int ok;
    if (NW > 1 && a.eig_shared) {
        if (warp == 0) {
            ok = eig_small_core<P, NW>(a, g, v, cold, inv_lam, hint);
#pragma unroll
            for (int k = 0; k < P; ++k)
                if (lane == k) g.v[k] = v[k];
            if (lane == 0) g.ibuf[12] = ok;
        }
        __syncthreads();
        ok = g.ibuf[12];
        if (warp != 0) load_col<P>(g.v, v);
        __syncthreads();
    } else {
        ok = eig_small_core<P, NW>(a, g, v, cold, inv_lam, hint);
    }
    if (ok != 1) {                                 // uniform across the CTA (every warp sees the same verdict)
        if constexpr (NW > 1) __syncthreads();
        int s = eig_squaring<NT>(g.G, P, a.p, g.v, g.red, g.B0, g.B0 + P * P, ok == 2);
        g.eig_steps += s;
        g.eig_fallbacks += 1;
        load_col<P>(g.v, v);
        bsync<NW>();
        hint = 0;
    }
}

// ---- final pass of an nmf() call (see final_pass in nmfoa_tiled.cu for what each sum is) ----------------------
template <int P, int NW, bool CLU, bool RES>
__device__ void final_pass_small(const KArgs &a, SGene &g, const double (&v)[P], bool first, bool want_res,
                                 double *e_first_g) {
    constexpr int NT = NW * 32, NV = 2 + 2 * P;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = g.n_cur;
    double acc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] = 0.0;
    for (int col = tid; col < n; col += NT) {
        double x[P], m[P];
        ld_col<P, RES>(g.X, col, x);
        ld_col<P, RES>(g.M, col, m);
        const double t = dot_v<P>(v, m);
        g.tb[col] = t;
        acc[0] += t;
        acc[1] = fma(t, t, acc[1]);
        // max_i ((KE - x)/(x + 1))^2 (nmf.py:280-282): the largest |num|/den is found by cross-multiplication, so
        // the column costs one division instead of P
        double bn = 0.0, bd = 1.0;
#pragma unroll
        for (int i = 0; i < P; ++i) {
            const double ke = v[i] * t;
            const double kc = ke < x[i] ? x[i] : ke;
            acc[2 + i] += x[i];
            acc[2 + P + i] += kc;
            if (want_res) {
                const double num = fabs((first ? ke : kc) - x[i]), den = x[i] + 1.0;
                if (num * bd > bn * den) { bn = num; bd = den; }
            }
        }
        if (want_res) {
            const double q = bn / bd;
            g.resb[col] = q * q;
        }
    }
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < P; ++k)
            if (lane == k) g.v[k] = v[k];
    }
    block_sum_vec<NV, NW>(acc, g.gpart, g.red);       // (its barriers also publish tb / resb / v)
    if constexpr (CLU) clu_allsum<NT>(g, g.red, NV, g.red);
    const double sum_t = g.red[0], sum_t2 = g.red[1];
    const double sigma = sqrt(sum_t2);
    if (e_first_g != nullptr) {                          // E of the first fit (only when no column was filtered)
        const double inv = sigma > 0.0 ? 1.0 / sigma : 0.0;
        for (int col = tid; col < n; col += NT) e_first_g[g.goff + col] = g.tb[col] * inv;
    }
    if (tid < P) {
        const double vi = g.v[tid];
        g.rsF[tid] = g.red[2 + tid];
        g.rsC[tid] = g.red[2 + P + tid];
        g.tmp[tid] = vi * sum_t;          // rs(K E), unclamped
        g.K[tid] = vi * sigma;            // K = u * s >= 0
    }
    bsync<NW>();
}

// nmf() on the current columns [0, n_cur) (nmf.py:78-107).  Leaves v, K, tmp = rs(KE), rsF, rsC, resb, tb.
template <int P, int NW, bool CLU, bool RES>
__device__ void run_nmf_small(const KArgs &a, SGene &g, bool first, bool want_res, double *e_first_g) {
    constexpr int NT = NW * 32, CS = P + 2;
    const int tid = threadIdx.x;
    {   // lambda = 0: M = x
        const double2 *src = reinterpret_cast<const double2 *>(g.X);
        double2 *dst = reinterpret_cast<double2 *>(g.M);
        // (flat copy of the storage that holds the first n_cur columns, whole blocks in the streamed layout)
        const int n2 = RES ? g.n_cur * (CS / 2) : (g.n_cur + 31) / 32 * (32 * P / 2);
        for (int e = tid; e < n2; e += NT) dst[e] = src[e];
    }
    bsync<NW>();
    double v[P];
#pragma unroll
    for (int k = 0; k < P; ++k) v[k] = 0.0;
    double inv_lam = 1.0;
    int hint = 0;
    const int T = a.nmf_iter;
    g.primed = false;
    gram_small<P, NW, false, CLU, RES>(a, g, v, T > 0);
    eig_small<P, NW>(a, g, v, true, inv_lam, hint);
    for (int it = 0; it < T; ++it) {
        gram_small<P, NW, true, CLU, RES>(a, g, v, it + 1 < T);
        eig_small<P, NW>(a, g, v, false, inv_lam, hint);
    }
    final_pass_small<P, NW, CLU, RES>(a, g, v, first, want_res, e_first_g);
}

template <int P, int NW, bool RES, bool CLU>
__global__ void __launch_bounds__(NW * 32, (12 / NW) > 0 ? (12 / NW) : 1) nmfoa_small_kernel(const KArgs a) {
    extern __shared__ double smem[];
    using Cfg = SmallCfg<P>;
    constexpr int NT = NW * 32, TC = Cfg::TC;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int p = a.p;
    const SmallCarve cv = small_carve(P, NW, RES ? a.resident_cols : 0, CLU);

    SGene g;
    double *sm = smem + cv.small;
    g.v = sm;           g.K = sm + P;        g.K0 = sm + 2 * P;   g.rs0 = sm + 3 * P;   g.rsF = sm + 4 * P;
    g.rsC = sm + 5 * P; g.rsC0 = sm + 6 * P; g.rho = sm + 7 * P;  g.scale = sm + 8 * P; g.tmp = sm + 9 * P;
    g.red = smem + cv.red;
    g.binm = smem + cv.binm;
    g.alive = reinterpret_cast<int *>(smem + cv.alive);
    g.ibuf = reinterpret_cast<int *>(smem + cv.ibuf);
    g.G = smem + cv.G;
    g.vx = smem + cv.vx;
    g.gpart = smem + cv.gpart;
    g.tab = reinterpret_cast<int *>(smem + cv.tab);
    g.lw = reinterpret_cast<int *>(smem + cv.lw);
    g.xbuf = smem + cv.xbuf;
    g.stage = smem + cv.stage;
    g.ring = smem + cv.ring;
    g.crank = 0; g.csize = 1; g.xpar = 0;
    if constexpr (CLU) {
        cg::cluster_group cl = cg::this_cluster();
        g.crank = (int)cl.block_rank();
        g.csize = (int)cl.num_blocks();
    }
    double *slab = a.ws + (long long)blockIdx.x * a.ws_stride;
    g.B0 = slab;
    const int cap = RES ? a.resident_cols : (int)a.ws_ld;          // columns one CTA can hold
    if constexpr (RES) {
        g.X = smem + cv.X; g.M = smem + cv.M; g.resb = smem + cv.resb; g.tb = smem + cv.tb;
    } else {
        g.X = slab + 2 * P * P;
        const long long wcols = (a.ws_ld + 31) / 32 * 32;           // whole 32-column blocks (blocked row-major)
        g.M = g.X + (long long)P * wcols;
        g.resb = g.M + (long long)P * wcols;
        g.tb = g.resb + wcols;
    }
    {   // this lane's Gram tile; table of all tiles for the cross-warp sum
        const int t = lane % Cfg::NTP;
        int r0, c0, offA, offB, u0, u1, u2;
        bool ok;
        tile_of<P>(t, r0, c0, ok);
        if (TC == 3 && ((c0 / 3) & 1)) { offA = c0 + 1; offB = c0; u0 = c0 + 1; u1 = c0 + 2; u2 = c0; }
        else { offA = c0; offB = c0 + 2; u0 = c0; u1 = c0 + 1; u2 = c0 + 2; }
        if (TC == 2) { offB = 0; u2 = 0; }
        g.tpack = ok ? (r0 | offA << 4 | offB << 8 | u0 << 12 | u1 << 16 | u2 << 20) : -1;
        if (tid < Cfg::NTP) {
            g.tab[tid * 4] = r0; g.tab[tid * 4 + 1] = u0; g.tab[tid * 4 + 2] = u1; g.tab[tid * 4 + 3] = u2;
        }
    }
    for (int e = tid; e < N_SMALL * P; e += NT) sm[e] = 0.0;
    g.eig_steps = 0;
    g.eig_fallbacks = 0;
    g.gpar = 0;
    __syncthreads();
    if constexpr (CLU) cg::this_cluster().sync();        // peers' shared memory is live before anyone writes to it

    for (;;) {
        // one queue ticket per gene: a lone CTA takes it itself, a cluster's rank 0 hands it to every peer
        if constexpr (CLU) {
            cg::cluster_group cl = cg::this_cluster();
            if (g.crank == 0 && tid == 0) {
                const int t = atomicAdd(a.queue, 1);
                for (int r = 0; r < g.csize; ++r) *cl.map_shared_rank(g.ibuf, r) = t;
            }
            cl.sync();
        } else {
            if (tid == 0) g.ibuf[0] = atomicAdd(a.queue, 1);
            __syncthreads();
        }
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

        // ------------------------------------------------------------------ baseline_selection (nmf.py:189-372)
        if (tid < P) g.scale[tid] = tid < p ? a.scale[tid] : 1.0;
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
        // (2) keep the columns that are high coverage (strict >) and on the systematic sample (nmf.py:220-229),
        //     scaled, compacted in order into the working buffer.  In a cluster every CTA takes a contiguous
        //     share of the candidates.
        const int rate = a.rate;
        const int start = (rate > 1 && a.ds_start) ? a.ds_start[gid] : 0;
        const int ncand = start < L ? (L - start + rate - 1) / rate : 0;
        const int share = (ncand + g.csize - 1) / g.csize;
        const int k_lo = min(g.crank * share, ncand), k_hi = min(k_lo + share, ncand);
        int exit_code = DN_EXIT_NONE;
        int ran = 0, nmf_calls = 0, sum_cols = 0;
        unsigned long long drops = 0ull;
        bool k_is_refined = false;
        int n0 = 0;              // kept columns of the whole gene
        g.goff = 0;
        if (share > cap) {
            exit_code = -1;                    // planner error: the bucket's tier is too small for this gene
        } else {
            int running = 0;
            int *wcount = g.ibuf + 1;        // NW ints
            for (int kb = k_lo; kb < k_hi; kb += NT) {
                const int k = kb + tid;
                bool keep = false;
                double xv[P];
#pragma unroll
                for (int i = 0; i < P; ++i) xv[i] = 0.0;
                if (k < k_hi) {
                    const long long col = start + (long long)k * rate;
                    double cm = -1.0e300;
#pragma unroll
                    for (int i = 0; i < P; ++i)
                        if (i < p) {
                            xv[i] = F[(long long)i * L + col] / g.scale[i];
                            cm = fmax(cm, xv[i]);
                        }
                    keep = cm > thr;
                }
                const unsigned bal = __ballot_sync(0xffffffffu, keep);
                int pre = running, tot;
                if constexpr (NW == 1) {
                    tot = __popc(bal);
                } else {
                    if (lane == 0) wcount[warp] = __popc(bal);
                    __syncthreads();
                    tot = 0;
#pragma unroll
                    for (int q = 0; q < NW; ++q) {
                        const int cq = wcount[q];
                        if (q < warp) pre += cq;
                        tot += cq;
                    }
                }
                if (keep) {
                    const int dst = pre + __popc(bal & ((1u << lane) - 1u));
                    st_col<P, RES>(g.X, dst, xv);
                }
                running += tot;
                if constexpr (NW > 1) __syncthreads();
            }
            bsync<NW>();
            g.n0 = g.n_cur = running;
            n0 = running;
            if constexpr (CLU) {
                // column counts of every CTA -> gene total and this CTA's global offset
                if (tid < g.csize) g.binm[tid] = tid == g.crank ? (double)running : 0.0;
                __syncthreads();
                clu_allsum<NT>(g, g.binm, g.csize, g.binm);
                n0 = 0;
                for (int r = 0; r < g.csize; ++r) {
                    if (r == g.crank) g.goff = n0;
                    n0 += (int)g.binm[r];
                }
                __syncthreads();
            }
        }
        g.n0g = g.n_curg = n0;
        if (exit_code == -1) {
            // nothing: reported through the counters
        } else if (n0 < a.min_hi) {
            exit_code = DN_EXIT_FEW_HICOV;                               // nmf.py:232-233
        } else {
            g.cs = n0; g.nb0 = 1; g.nalive = 1;
            if (tid == 0) { g.alive[0] = 0; g.lw[0] = g.n0; }
            {   // rs(F_start)
                double rs[P];
#pragma unroll
                for (int i = 0; i < P; ++i) rs[i] = 0.0;
                for (int col = tid; col < g.n0; col += NT) {
                    double x[P];
                    ld_col<P, RES>(g.X, col, x);
#pragma unroll
                    for (int i = 0; i < P; ++i) rs[i] += x[i];
                }
                block_sum_vec<P, NW>(rs, g.gpart, g.rs0);
                if constexpr (CLU) clu_allsum<NT>(g, g.rs0, P, g.rs0);
            }
            bool any_empty = false;
            for (int i = 0; i < p; ++i) any_empty |= !(g.rs0[i] > 0.0);
            if (any_empty && !(a.flags & DN_FLAG_PLAIN_NMF)) {
                exit_code = DN_EXIT_EMPTY_SAMPLE;                        // nmf.py:241-242
            } else {
                const bool store_e = (a.e_first != nullptr) && (n0 == L);
                // (4) first fit (nmf.py:245-254), then the bin-drop loop (nmf.py:273-324); one nmf() call site
                bool first = true, in_loop = false;
                double rmax = 0.0;
                for (;;) {
                    run_nmf_small<P, NW, CLU, RES>(a, g, first, true, (first && store_e) ? a.e_first + o0 : nullptr);
                    nmf_calls += 1; sum_cols += g.n_curg;
                    if (first) {
                        if (tid < P) {
                            g.rho[tid] = 1.0 - g.rs0[tid] / (g.tmp[tid] + 1.0);
                            g.K0[tid] = g.K[tid];
                            g.rsC0[tid] = g.rsC[tid];
                        }
                        bsync<NW>();
                        if (!(a.flags & DN_FLAG_PLAIN_NMF) && median_one_minus(g.rho, p) > 1.0) {
                            exit_code = DN_EXIT_MEDIAN;                  // nmf.py:257-258
                            break;
                        }
                        double rmin = g.rho[0];
                        rmax = g.rho[0];
                        for (int i = 1; i < p; ++i) { rmin = fmin(rmin, g.rho[i]); rmax = fmax(rmax, g.rho[i]); }
                        if (!(n0 >= a.min_len && rmin <= 0.2 && !a.skip)) {      // nmf.py:265
                            exit_code = DN_EXIT_NO_SELECTION;
                            break;
                        }
                        g.cs = (n0 + a.bins - 1) / a.bins;               // utils.py:176-192
                        g.nb0 = (n0 + g.cs - 1) / g.cs;
                        g.nalive = g.nb0;
                        for (int b = tid; b < g.nb0; b += NT) {
                            g.alive[b] = b;
                            // this CTA's share of bin b: [b cs, (b+1) cs) cut to its global column range
                            const int lo = max(b * g.cs, g.goff), hi = min(min((b + 1) * g.cs, n0), g.goff + g.n0);
                            g.lw[b] = max(0, hi - lo);
                        }
                        bsync<NW>();
                        in_loop = true;
                        first = false;
                    } else {
                        double mn = g.tmp[0];
                        for (int i = 1; i < p; ++i) mn = fmin(mn, g.tmp[i]);
                        if (mn == 0.0) break;                            // nmf.py:315-316
                        bsync<NW>();
                        if (tid < P) g.rho[tid] = 1.0 - g.rsF[tid] / (g.rsC[tid] + 1.0);   // nmf.py:318-321
                        bsync<NW>();
                        rmax = g.rho[0];
                        for (int i = 1; i < p; ++i) rmax = fmax(rmax, g.rho[i]);
                        if (g.nalive <= a.min_bins || g.n_curg < a.min_len) break;       // nmf.py:323
                    }
                    if (!(rmax > 0.1)) break;                            // nmf.py:273
                    ran = 1;
                    // sum of the squared relative residuals per alive bin (nmf.py:280-283); one warp per bin.
                    // Alive bins are contiguous in every CTA: bin k starts at lstart(k) and is lw[alive[k]] wide here.
                    for (int k = warp; k < g.nalive; k += NW) {
                        const int wl = g.lw[g.alive[k]];
                        const double *rr = g.resb + lstart(g, k);
                        double s = 0.0;
                        for (int j = lane; j < wl; j += 32) s += rr[j];
                        s = warp_sum(s);
                        if (lane == 0) g.binm[k] = s;
                    }
                    bsync<NW>();
                    if constexpr (CLU) clu_allsum<NT>(g, g.binm, g.nalive, g.binm);
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
                    {   // rotate this CTA's part of the dropped bin to the end of its current columns (M is scratch)
                        const int a0 = lstart(g, kd);
                        const int wl = g.lw[bd];
                        const int tail = g.n_cur - a0 - wl;
                        bsync<NW>();
                        for (int c = tid; c < tail + wl; c += NT) {
                            double x[P];
                            ld_col<P, RES>(g.X, a0 + c, x);
                            st_col<P, RES>(g.M, a0 + c, x);
                        }
                        bsync<NW>();
                        for (int c = tid; c < tail + wl; c += NT) {
                            double x[P];
                            ld_col<P, RES>(g.M, a0 + (c < tail ? wl + c : c - tail), x);
                            st_col<P, RES>(g.X, a0 + c, x);
                        }
                        if (tid == 0)
                            for (int k = kd; k < g.nalive - 1; ++k) g.alive[k] = g.alive[k + 1];
                        g.n_cur -= wl;
                    }
                    bsync<NW>();
                    g.nalive -= 1;
                    g.n_curg -= wd;
                    drops |= 1ull << bd;
                    if (g.n_curg < 2) break;                             // svds ValueError swallowed, nmf.py:306-310
                }
                if (in_loop) {
                    bsync<NW>();
                    bool fallback = true;
                    exit_code = DN_EXIT_FALLBACK;
                    if (rmax < 0.2) {                                    // nmf.py:327-346
                        floor_abs(g.K, g.K, p);
                        double s = 0.0;
                        for (int j = tid; j < g.n0; j += NT) {
                            double x[P];
                            ld_col<P, RES>(g.X, j, x);
                            double e = -1.0e300;
#pragma unroll
                            for (int i = 0; i < P; ++i)
                                if (i < p) e = fmax(e, x[i] / g.K[i]);
                            s += e;
                        }
                        double S = block_sum<NT>(s, g.red);
                        if constexpr (CLU) {
                            if (tid == 0) g.binm[0] = S;
                            __syncthreads();
                            clu_allsum<NT>(g, g.binm, 1, g.binm);
                            S = g.binm[0];
                            __syncthreads();
                        }
                        if (tid < P) g.rho[tid] = 1.0 - g.rs0[tid] / (g.K[tid] * S + 1.0);
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
                    if (fallback) {                                      // nmf.py:342-353
                        __syncthreads();
                        if (tid < P) g.rho[tid] = 1.0 - g.rs0[tid] / (g.rsC0[tid] + 1.0);
                        __syncthreads();
                    }
                }
            }
        }
        // (5) outputs (rank 0 of a cluster; every CTA holds the same values)
        __syncthreads();
        const bool is_default = exit_code == DN_EXIT_FEW_HICOV || exit_code == DN_EXIT_EMPTY_SAMPLE ||
                                exit_code == DN_EXIT_MEDIAN || exit_code == -1;
        if (!is_default && !k_is_refined) {
            // K of the first fit; floored unless the estimate keeps the fit's own columns (n0 == L)
            if (n0 == L) {
                if (tid < P) g.K[tid] = g.K0[tid];
                __syncthreads();
            } else {
                floor_abs(g.K0, g.K, p);
            }
        }
        if (g.crank == 0) {
            if (tid < p) {
                double r = is_default ? 0.0 : g.rho[tid];
                if (!(a.flags & DN_FLAG_RAW_RHO)) r = r > 0.9 ? 0.9 : r;                                   // nmf.py:398-399
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
                    cnt[DN_CNT_RESIDENT] = (int)RES | (g.eig_fallbacks << 1);
                }
            }
        }
        __syncthreads();
        if (a.est) {
            // last outer iteration: the CTA(s) that own the gene write its full-length estimate (fused dn_estimates)
            write_estimate(F, L, p, g.scale, exit_code, n0, g.K, a.e_first ? a.e_first + o0 : nullptr,
                           a.est + (long long)p * (a.est_off ? a.est_off[gid] : o0), g.crank * NT + tid, g.csize * NT);
            __syncthreads();
        }
        if constexpr (CLU) cg::this_cluster().sync();    // nobody re-uses a peer's ticket slot before it was read
    }
}

template <int P, int NW, bool RES>
int launch_small_one(const KArgs &a, const dn_plan *plan, cudaStream_t st) {
    auto kern = nmfoa_small_kernel<P, NW, RES, false>;
    DN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, plan->smem_bytes));
    kern<<<plan->ctas, NW * 32, plan->smem_bytes, st>>>(a);
    DN_CUDA(cudaGetLastError());
    return DN_OK;
}

// cluster launch: plan->cluster CTAs per gene, grid = whole clusters
template <int P, bool RES>
int launch_small_cluster(const KArgs &a, const dn_plan *plan, cudaStream_t st) {
    auto kern = nmfoa_small_kernel<P, SMALL_CLU_WARPS, RES, true>;
    DN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, plan->smem_bytes));
    if (plan->cluster > 8) DN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(plan->ctas / plan->cluster * plan->cluster, 1, 1);
    cfg.blockDim = dim3(SMALL_CLU_WARPS * 32, 1, 1);
    cfg.dynamicSmemBytes = plan->smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = plan->cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    DN_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
    return DN_OK;
}

template <int P>
int launch_small(const KArgs &a, const dn_plan *plan, cudaStream_t st) {
    if (plan->cluster > 1) {
        if (plan->threads != SMALL_CLU_WARPS * 32) return dn_fail(DN_ERR_INVALID, "cluster plans use 256 threads%s");
        return plan->resident_cols > 0 ? launch_small_cluster<P, true>(a, plan, st)
                                       : launch_small_cluster<P, false>(a, plan, st);
    }
    if (plan->resident_cols > 0) {
        switch (plan->threads) {
            case 32: return launch_small_one<P, 1, true>(a, plan, st);
            case 64: return launch_small_one<P, 2, true>(a, plan, st);
            case 128: return launch_small_one<P, 4, true>(a, plan, st);
            case 256: return launch_small_one<P, 8, true>(a, plan, st);
            case 512: return launch_small_one<P, 16, true>(a, plan, st);
        }
        return dn_fail(DN_ERR_INVALID, "plan.threads must be 32..512 (power of two) on the small-p path%s");
    }
    if (plan->threads != 256) return dn_fail(DN_ERR_INVALID, "streamed small-p plans use 256 threads%s");
    return launch_small_one<P, 8, false>(a, plan, st);
}

}  // namespace
