// degnorm_b200 -- fused baseline-selection kernel for 13..48 samples, sm_100a ("mid-p" kernel).
//
// Same per-gene flow as the other two kernels (reference: /root/reference/degnorm/nmf.py:189-372 around
// nmf.py:78-107).  At p = 48 a column is 384 bytes and the Gram matrix has 1,176 entries per column: the pass is a
// memory stream (1,152 algorithmic bytes per column-iteration) with a SYRK riding on it (~1,400 FMAs per column),
// bound by HBM/L2 bandwidth with the FP64 pipe at roughly 40 %.  Design:
//
//   * one kernel, streamed: x and M = x + lambda live in per-CTA global slabs, column-major with a column stride
//     of P + 2 doubles (the shared-memory-friendly layout, so a chunk is one flat copy); small genes simply stay
//     in the 126 MB L2.  lambda = M - x is recovered on load, as in the small-p kernel.
//   * the CTA (8 warps) walks its columns in 64-column chunks through a 3-stage ring filled by TMA 1-D bulk copies
//     (cp.async.bulk + one mbarrier per stage).  Phase A: 4 lanes per column (12 rows each, two xor-shuffles for
//     t = v.M_j), new M written in place into the ring stage; the stage goes back to the slab as ONE bulk store
//     (shared -> global) issued after the next chunk barrier.  Phase B: warp w takes the 8 columns its own lanes updated; its
//     lanes own 30 tiles of 6 x 8 Gram entries that cover the upper triangle, operands straight from the stage
//     (7 LDS.128 per 48 FMAs), accumulators stay in registers for the whole pass.  One block barrier per chunk.
//   * per pass the 8 warps' partial Grams are summed by a fixed halving tree through shared memory; a cluster's CTAs
//     exchange their sums through per-CTA global slots (double-buffered, one cluster barrier per pass) and add them
//     in rank order, so every CTA holds the bit-identical G and solves redundantly (warp 0, power iteration on G in
//     shared memory with the same guards as the other kernels).
//   * long genes: a thread-block cluster (2..16 CTAs) per gene, contiguous column slices, exactly the scheme of the
//     small-p cluster kernels (bins stay contiguous per CTA, drops rotate locally).
//
// No tensor cores: fp64 FMA pipe only.
#pragma once
#include <cooperative_groups.h>
#include "common.cuh"
#include "launch.h"
#include "tma.cuh"
namespace cg = cooperative_groups;

namespace {

#ifndef MID_NW
#error "define MID_NW (warps per CTA: 4 or 8) and MID_LAUNCHER before including nmfoa_mid.cuh"
#endif
#ifndef MID_TMA_STORE
#define MID_TMA_STORE (MID_NW == 8)
#endif
// MID_TMA_STORE: the updated M of a chunk goes back to the slab as ONE bulk copy out of the ring stage (issued by
// one thread after the chunk barrier) instead of six 128-bit stores per thread in phase A.  The kernel is bound by
// its load/store pipe: on the 8-warp instantiation this took C3 from 0.52 to 0.59 of the HBM roofline; the 4-warp
// one (32-column chunks, loads only one chunk ahead of their use) measured no gain and keeps per-thread stores.
#ifndef MID_NA
#define MID_NA 0
#endif
#ifndef MID_B_PIPELINED
#define MID_B_PIPELINED 1      // phase B software-pipelined (measured +4 % on C3; 0: load two columns, then their FMAs)
#endif
// MID_NA > 0: warp-specialised instantiation.  Warps 0 .. MNW-1 ("B warps") only accumulate the Gram matrix, warps
// MNW .. MNW+MNA-1 ("A warps") only run the multiplier update one chunk ahead of them; nobody waits at a block barrier
// inside a pass (see gram_mid_ws).
constexpr int MNW = MID_NW;               // Gram warps per CTA (8: one CTA per SM; 4: two CTAs per SM)
constexpr int MNA = MID_NA;               // update warps (warp-specialised instantiation only)
constexpr int MP = MID_P;                 // padded samples
constexpr int MCS = MID_P + 2;            // column stride (doubles)
constexpr int MNT = (MNW + MNA) * 32;     // threads
constexpr int MNWT = MNW + MNA;           // warps
// Warp-specialised instantiation: 32-column chunks in a 6-stage ring (the same shared memory as 3 stages of 64): the
// stage of chunk k is refilled with chunk k + 6 once its consumers are done, so FOUR chunks (100 KB) are in flight per
// SM instead of one.  With one 51 KB chunk in flight per SM the pass was bound by memory-level parallelism (one chunk per
// HBM round trip), whatever the roles did.
constexpr int MCH = MNA > 0 ? MID_WS_CHUNK : mid_chunk(MNW);     // columns per chunk
constexpr int MRING = MNA > 0 ? MID_WS_RING : MID_RING;          // ring stages
constexpr int MCPR = MNT / 4;             // columns per CTA step where 4 lanes share a column (scan-type passes)
#ifndef WS_SETMAXNREG
#define WS_SETMAXNREG 2
#endif
constexpr int WS_LAG = MID_WS_LAG;    // bulk stores an update warp keeps in flight (< ring stages - 2)
constexpr int WS_REG_BASE = 168, WS_REG_B = 184, WS_REG_A = 136;   // see gram_mid_ws: 128 (168 - A) >= 256 (B - 168)
constexpr int IB_WCOUNT = 1;              // ibuf slots: [0] queue ticket, [1 .. 12] per-warp counts, [20] eigen flag,
constexpr int IB_EIG = 20;                //             [24 .. 29] consumers done with a ring stage
constexpr int IB_FREE = 24;
// Gram tiles: 30 tiles of 8 rows x 6 columns cover the upper triangle of 48 x 48 (row block rb, column block cb with
// 6 cb + 5 >= 8 rb).  They are dealt to the 32 lanes row-major with every tile row padded to an even length (two pad
// slots in all), so that the two lanes of an aligned pair share their ROW operands: a half-warp whose pairs read the
// same 16 bytes is served in one shared-memory wavefront instead of two (measured with tools/probe_peaks.py), which
// takes a column's operand fetch from 28 wavefronts (6 x 8 tiles, 7 loads of 4) to 20 (4 row loads of 2 + 3 column
// loads of 4).  Row pairs are fetched in an order rotated by (rb / 2) so that the row blocks of a half-warp (64 bytes
// apart) fall into different banks; the column blocks are 48 bytes apart and need no rotation.
constexpr int MNTILE = 30;
constexpr int MNE = MNTILE * 48;          // partial sums per warp / CTA
__device__ __forceinline__ int mrot8(int r, int rb) { return 2 * (((r >> 1) + ((rb >> 1) & 3)) & 3) + (r & 1); }

struct MGene {
    double *v, *K, *K0, *rs0, *rsF, *rsC, *rsC0, *rho, *scale, *tmp, *red, *binm, *G, *buf, *ring;
    int *alive, *ibuf, *lw, *tab;
    double *X, *M, *resb, *tb, *B0, *slots;      // global (slab)
    long long slot_stride;                        // doubles between consecutive CTAs' slabs
    int n0, n_cur, n0g, n_curg, goff, cs, nb0, nalive;
    int crank, csize, xpar, eig_steps, eig_fallbacks;
    bool primed;
    unsigned long long *mbar;                      // MID_RING mbarriers (one per ring stage)
    unsigned use0, use1, use2;                      // completed fills per stage (phase parity of its mbarrier)
    unsigned seq;                                   // warp-specialised instantiation: ring chunks consumed so far
};

__device__ __forceinline__ int mlstart(const MGene &g, int k) {
    int s = 0;
    for (int q = 0; q < k; ++q) s += g.lw[g.alive[q]];
    return s;
}

__device__ __forceinline__ void ld12(const double *p, double (&x)[12]) {
    const double2 *q = reinterpret_cast<const double2 *>(p);
#pragma unroll
    for (int i = 0; i < 6; ++i) { const double2 t = q[i]; x[2 * i] = t.x; x[2 * i + 1] = t.y; }
}
__device__ __forceinline__ void ld6(const double *p, double (&x)[6]) {
    const double2 *q = reinterpret_cast<const double2 *>(p);
#pragma unroll
    for (int i = 0; i < 3; ++i) { const double2 t = q[i]; x[2 * i] = t.x; x[2 * i + 1] = t.y; }
}
__device__ __forceinline__ void st6(double *p, const double (&x)[6]) {
    double2 *q = reinterpret_cast<double2 *>(p);
#pragma unroll
    for (int i = 0; i < 3; ++i) q[i] = make_double2(x[2 * i], x[2 * i + 1]);
}
__device__ __forceinline__ void st12(double *p, const double (&x)[12]) {
    double2 *q = reinterpret_cast<double2 *>(p);
#pragma unroll
    for (int i = 0; i < 6; ++i) q[i] = make_double2(x[2 * i], x[2 * i + 1]);
}

// All-to-all sum over the cluster of n doubles (vals: shared, published to the CTA) through the CTAs' global slots.
__device__ void mclu_allsum(MGene &g, const double *vals, int n, double *out) {
    const int tid = threadIdx.x;
    if (g.csize == 1) {
        __syncthreads();
        for (int k = tid; k < n; k += MNT) out[k] = vals[k];
        __syncthreads();
        return;
    }
    cg::cluster_group cl = cg::this_cluster();
    double *mine = g.slots + (long long)g.xpar * MNE;
    for (int k = tid; k < n; k += MNT) mine[k] = vals[k];
    __threadfence();
    cl.sync();
    const double *first = g.slots - (long long)g.crank * g.slot_stride + (long long)g.xpar * MNE;
    for (int k = tid; k < n; k += MNT) {
        double s = 0.0;
        for (int r = 0; r < g.csize; ++r) s += __ldcg(first + (long long)r * g.slot_stride + k);
        out[k] = s;
    }
    g.xpar ^= 1;
    __syncthreads();
}

// ---- end of a pass: the Gram warps' partial sums -> CTA sum -> cluster sum -> square G in shared memory -----------
struct FinArgs {
    double *buf, *slots, *G;
    const int *tab;
    long long slot_stride;
    int crank, csize, xpar;
};
__device__ __forceinline__ FinArgs fin_args(const MGene &g) {
    FinArgs f;
    f.buf = g.buf; f.slots = g.slots; f.G = g.G; f.tab = g.tab; f.slot_stride = g.slot_stride;
    f.crank = g.crank; f.csize = g.csize; f.xpar = g.xpar;
    return f;
}
// (returns the exchange-slot parity to use next)
__device__ __forceinline__ int gram_finish(FinArgs g, double (&acc)[8][6]) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // ---- the warps' partials -> CTA sum, in the order of a fixed halving tree: ((w0 + w4) + (w2 + w6)) + ((w1 + w5) +
    // (w3 + w7)).  Eight warps: the upper four park their tiles in four shared slots, the lower four add them and write
    // their sums back to the same slots (two block barriers); the last two levels of the tree are folded into the
    // reader below, which adds the four slots in tree order (the full tree took eight barriers per pass).
    const int tile = g.tab[4 * lane + 2];                       // dense tile id of this lane's slot (-1: pad slot)
    constexpr int NSLOT = MNW == 8 ? 4 : 1;
    if constexpr (MNW == 8) {
        double *slot_ = g.buf + (long long)(warp & 3) * MNE + (tile >= 0 ? tile : 0) * 48;
        if (warp >= 4 && warp < 8 && tile >= 0) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int q = 0; q < 6; q += 2)
                    *reinterpret_cast<double2 *>(slot_ + r * 6 + q) = make_double2(acc[r][q], acc[r][q + 1]);
        }
        __syncthreads();
        if (warp < 4 && tile >= 0) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int q = 0; q < 6; q += 2) {
                    const double2 t = *reinterpret_cast<const double2 *>(slot_ + r * 6 + q);
                    *reinterpret_cast<double2 *>(slot_ + r * 6 + q) = make_double2(acc[r][q] + t.x, acc[r][q + 1] + t.y);
                }
        }
        __syncthreads();
    } else {
        double *mine = g.buf + (long long)(warp & 3) * MNE + (tile >= 0 ? tile : 0) * 48;
        for (int half = MNW / 2; half >= 1; half >>= 1) {
            if (warp >= half && warp < 2 * half && tile >= 0) {
                double *dst = g.buf + (long long)(warp - half) * MNE + tile * 48;
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int q = 0; q < 6; q += 2)
                        *reinterpret_cast<double2 *>(dst + r * 6 + q) = make_double2(acc[r][q], acc[r][q + 1]);
            }
            __syncthreads();
            if (warp < half && tile >= 0) {
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int q = 0; q < 6; q += 2) {
                        const double2 t = *reinterpret_cast<const double2 *>(mine + r * 6 + q);
                        acc[r][q] += t.x;
                        acc[r][q + 1] += t.y;
                    }
            }
            __syncthreads();
        }
        if (warp == 0 && tile >= 0) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int q = 0; q < 6; q += 2)
                    *reinterpret_cast<double2 *>(g.buf + tile * 48 + r * 6 + q) = make_double2(acc[r][q], acc[r][q + 1]);
        }
        __syncthreads();
    }
    auto cta_sum = [&](int e) -> double {
        if constexpr (NSLOT == 4) return (g.buf[e] + g.buf[2 * MNE + e]) + (g.buf[MNE + e] + g.buf[3 * MNE + e]);
        else return g.buf[e];
    };
    // ---- cluster sum (rank order) and scatter into the square G
    bool local = true;
    if (g.csize > 1) {
        cg::cluster_group cl = cg::this_cluster();
        double *slot = g.slots + (long long)g.xpar * MNE;
        for (int e = tid; e < MNE; e += MNT) slot[e] = cta_sum(e);
        __threadfence();
        cl.sync();
        local = false;
    }
    const double *first = g.slots - (long long)g.crank * g.slot_stride + (long long)g.xpar * MNE;
    for (int e = tid; e < MNE; e += MNT) {
        double s;
        if (local) {
            s = cta_sum(e);
        } else {
            s = 0.0;
            for (int r = 0; r < g.csize; ++r) s += __ldcg(first + (long long)r * g.slot_stride + e);
        }
        const int t = e / 48, rq = e - t * 48;
        const int r0t = g.tab[128 + 2 * t], c0t = g.tab[128 + 2 * t + 1];     // (row, column) offset of dense tile t
        const int r = rq / 6, q = rq - 6 * r;
        const int i = r0t + mrot8(r, r0t >> 3), j = c0t + q;
        if (i <= j) {
            g.G[i * MP + j] = s;
            g.G[j * MP + i] = s;
        }
    }
    if (g.csize > 1) g.xpar ^= 1;
    __syncthreads();
    return g.xpar;
}

// ---- one pass over this CTA's columns: (optional multiplier update) + Gram of M -> G (shared, identical cluster-wide)
template <bool UPDATE>
// m_is_x: lambda is still zero, so this pass reads M from the x array (the first fit and the first update pass of an
// nmf() call: M = x is never copied; the slab's M is first written by the store-back of the first update pass).
__device__ void gram_mid(const KArgs &a, MGene &g, bool prime_next, bool m_is_x) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = g.n_cur;
    const int nchunk = (n + MCH - 1) / MCH;
    constexpr int STG = 2 * MCH * MCS;                     // doubles per ring stage (M then x)
    const double c = a.c;
    double acc[8][6];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int q = 0; q < 6; ++q) acc[r][q] = 0.0;
    double vq[12];
    const int q4 = tid & 3;
    if constexpr (UPDATE) ld12(g.v + 12 * q4, vq);
    // this lane's tile, read from the shared table (a load is not rematerialised inside the chunk loop)
    const int tr0 = *reinterpret_cast<volatile int *>(g.tab + 4 * lane + 2) >= 0
                        ? *reinterpret_cast<volatile int *>(g.tab + 4 * lane) : -1;
    const int tc0 = *reinterpret_cast<volatile int *>(g.tab + 4 * lane + 1);
    const int ao0 = tr0 + mrot8(0, tr0 >> 3), ao1 = tr0 + mrot8(2, tr0 >> 3), ao2 = tr0 + mrot8(4, tr0 >> 3),
              ao3 = tr0 + mrot8(6, tr0 >> 3);

    // One thread hands a chunk (25.6 KB of M, and of x for an update pass) to the TMA unit; the bytes land in the
    // ring stage and complete the stage's mbarrier.  Callers guarantee (by a __syncthreads) that nobody still uses
    // the stage and that the slab holds the data (ordinary stores of the previous pass): the proxy fence orders
    // those generic-proxy accesses before the async-proxy copy.
    const double *msrc = m_is_x ? g.X : g.M;
    auto issue = [&](int ch, bool with_x, const double *msrc_) {
        if (tid == 0 && ch < nchunk) {
            const int st = ch % MID_RING;
            double *dst = g.ring + st * STG;
            const unsigned bytes = MCH * MCS * 8;
            // (the slab's ordinary stores were fenced by their writers before the pass: only the stage matters here)
            // (the full fence is a MEMBAR.GPU on the issuing warp, which the other seven then wait for: -2 %)
            fence_proxy_async_smem();
            mbar_expect_tx(g.mbar + st, with_x ? 2 * bytes : bytes);
            bulk_g2s(dst, msrc_ + (long long)ch * (MCH * MCS), bytes, g.mbar + st);
            if (with_x) bulk_g2s(dst + MCH * MCS, g.X + (long long)ch * (MCH * MCS), bytes, g.mbar + st);
        }
    };
    auto wait_stage = [&](int ch) {
        const int st = ch % MID_RING;
        const unsigned par = (st == 0 ? g.use0 : (st == 1 ? g.use1 : g.use2)) & 1u;
        mbar_wait(g.mbar + st, par);
        if (st == 0) ++g.use0; else if (st == 1) ++g.use1; else ++g.use2;
    };
    if (!g.primed) {
#pragma unroll
        for (int q = 0; q < MID_RING - 1; ++q) issue(q, UPDATE, msrc);
    }
    constexpr bool TSTORE = (MID_TMA_STORE != 0) && UPDATE;
    // TSTORE schedule, at the barrier that opens chunk ch: stage (ch - 1) holds that chunk's final M -> bulk store;
    // the store of chunk ch - 2 (one chunk old) has finished reading its stage -> that stage takes chunk ch + 1.
    auto store_back = [&](int ch) {
        if (tid == 0 && ch >= 0 && ch < nchunk) {
            const int ncs = min(MCH, n - ch * MCH);
            bulk_s2g(g.M + (long long)ch * (MCH * MCS), g.ring + (ch % MID_RING) * STG, (unsigned)(ncs * MCS * 8));
        }
    };
    for (int ch = 0; ch < nchunk; ++ch) {
        if constexpr (TSTORE) fence_proxy_async_smem();    // this thread's phase-A stores of chunk ch - 1 -> async proxy
        __syncthreads();                                   // everyone is done with stage (ch - 1): refill it
        if constexpr (TSTORE) {
            store_back(ch - 1);
            if (tid == 0) bulk_wait_read<1>();
            if (ch + 1 >= MID_RING - 1) issue(ch + 1, UPDATE, msrc);
        } else {
            issue(ch + MID_RING - 1, UPDATE, msrc);
        }
        wait_stage(ch);                                    // chunk ch has landed
        double *sM = g.ring + (ch % MID_RING) * STG;
        const int ncol = min(MCH, n - ch * MCH);
        if constexpr (UPDATE) {
            // phase A: 4 lanes per column, 12 rows each
            const int cc = tid >> 2;
            const bool act = cc < ncol;
            double m[12], x[12];
            double tp = 0.0;
            if (act) {
                ld12(sM + cc * MCS + 12 * q4, m);
                ld12(sM + MCH * MCS + cc * MCS + 12 * q4, x);
                double t0 = 0.0, t1 = 0.0;
#pragma unroll
                for (int i = 0; i < 12; i += 2) { t0 = fma(vq[i], m[i], t0); t1 = fma(vq[i + 1], m[i + 1], t1); }
                tp = t0 + t1;
            }
            double t = tp;                                  // (shuffles outside the branch: whole warps take part)
            t += __shfl_xor_sync(0xffffffffu, t, 1);
            t += __shfl_xor_sync(0xffffffffu, t, 2);
            if (act) {
#pragma unroll
                for (int i = 0; i < 12; ++i) {
                    const double res = fma(vq[i], t, -x[i]);
                    const double w = fma(-c, res, m[i] - x[i]);
                    m[i] = fma(0.5, w + fabs(w), x[i]);
                }
                st12(sM + cc * MCS + 12 * q4, m);
                // (TSTORE: the stage is read by the bulk store after the next chunk barrier; the proxy fence that
                // has to stand between these stores and that copy is taken just before the barrier, when the stores
                // have long drained, instead of here, where every warp would sit out their latency: 4 % of the samples)
                if constexpr (!TSTORE) st12(g.M + ((long long)ch * MCH + cc) * MCS + 12 * q4, m);
            }
            __syncwarp();        // phase B of this warp only reads the 8 columns its own lanes just wrote
        }
        // phase B: warp w sweeps columns w, w + 8, ... of the chunk; lanes own 6 x 8 tiles
        if (tr0 >= 0) {
            // The four column pairs are fetched in an order rotated by (c0 / 16): in every one of the four load
            // instructions the lanes whose c0 differ by 16 or 32 doubles (same banks) then read different pairs, so
            // the loads are bank-conflict free; the accumulator columns are un-rotated when G is built.
            // Two columns per trip: 14 loads, then 96 FMAs (the shared-load latency is paid once per pair).
            // (columns 8 w .. 8 w + 7 of the chunk: the ones this warp updated in phase A, so no block barrier)
            int cc = warp * (MCH / MNW);
            const int cend = min(ncol, cc + MCH / MNW);
#define MID_LOADP(S, col)                                                                   \
    {                                                                                       \
        const double *mc_ = sM + (col) * MCS;                                               \
        S##a0 = *reinterpret_cast<const double2 *>(mc_ + ao0);                              \
        S##a1 = *reinterpret_cast<const double2 *>(mc_ + ao1);                              \
        S##a2 = *reinterpret_cast<const double2 *>(mc_ + ao2);                              \
        S##a3 = *reinterpret_cast<const double2 *>(mc_ + ao3);                              \
        S##u0 = *reinterpret_cast<const double2 *>(mc_ + tc0);                              \
        S##u1 = *reinterpret_cast<const double2 *>(mc_ + tc0 + 2);                          \
        S##u2 = *reinterpret_cast<const double2 *>(mc_ + tc0 + 4);                          \
    }
#define MID_FMAP(S)                                                                         \
    {                                                                                       \
        const double ar_[8] = {S##a0.x, S##a0.y, S##a1.x, S##a1.y, S##a2.x, S##a2.y, S##a3.x, S##a3.y}; \
        const double uc_[6] = {S##u0.x, S##u0.y, S##u1.x, S##u1.y, S##u2.x, S##u2.y};       \
        _Pragma("unroll") for (int r = 0; r < 8; ++r)                                       \
            _Pragma("unroll") for (int q = 0; q < 6; ++q) acc[r][q] = fma(ar_[r], uc_[q], acc[r][q]); \
    }
            double2 Aa0, Aa1, Aa2, Aa3, Au0, Au1, Au2, Ba0, Ba1, Ba2, Ba3, Bu0, Bu1, Bu2;
            // Two columns per trip: 14 loads, then 96 FMAs (the shared-load latency is paid once per pair).
            // (Keeping the next column's operands in flight behind this column's FMAs measured no faster: the
            // phase is bound by shared-memory wavefronts, 28 per column for the 6 x 8 tiles, not by load latency.)
#if MID_B_PIPELINED
            // software-pipelined: the operands of the next column are requested before the FMAs of the current one,
            // so that shared-memory fetch and FP64 issue overlap inside the warp (two operand sets, A and B)
            if (cc < cend) {
                MID_LOADP(A, cc);
#pragma unroll 1
                for (; cc + 2 < cend; cc += 2) {
                    MID_LOADP(B, cc + 1);
                    MID_FMAP(A);
                    MID_LOADP(A, cc + 2);
                    MID_FMAP(B);
                }
                if (cc + 1 < cend) {
                    MID_LOADP(B, cc + 1);
                    MID_FMAP(A);
                    MID_FMAP(B);
                } else {
                    MID_FMAP(A);
                }
            }
#else
#pragma unroll 1
            for (; cc + 1 < cend; cc += 2) {
                MID_LOADP(A, cc);
                MID_LOADP(B, cc + 1);
                MID_FMAP(A);
                MID_FMAP(B);
            }
            if (cc < cend) {
                MID_LOADP(A, cc);
                MID_FMAP(A);
            }
#endif
#undef MID_LOADP
#undef MID_FMAP
        }
    }
    fence_proxy_async();
    __syncthreads();
    if constexpr (TSTORE) {
        store_back(nchunk - 1);
        if (tid == 0) bulk_wait_all();                     // the slab holds the whole new M before anyone reads it
    }
    g.primed = prime_next;
    if (prime_next) {
#pragma unroll
        for (int q = 0; q < MID_RING - 1; ++q) issue(q, true, (!UPDATE && m_is_x) ? g.X : g.M);   // (the pass after the first fit)
    }
    g.xpar = gram_finish(fin_args(g), acc);
}

// ---- the same pass, warp-specialised (MID_NA > 0) ------------------------------------------------------------------
// Roles.  B warps (0 .. MNW-1) own the Gram accumulators and nothing else; A warps (MNW ..) run the multiplier update on
// the chunk that landed last, one chunk ahead of the B warps, and send the updated M back to the slab (one bulk store
// per A warp and chunk: its own 16 columns).  Hand-over per ring stage:
//     full[s]  (mbarrier, TMA transaction bytes) : chunk landed                        -> A warps (B warps on the first pass)
//     upd[s]   (mbarrier, every A thread arrives): M of the chunk is final in the stage -> B warps
//     free[s]  (shared counter)                  : B warps are done reading the stage and the A warps' bulk stores have
//                                                  read it; whoever arrives LAST requests chunk ch + RING into it
// so no thread ever waits for a stage to drain, and no block barrier is taken inside a pass: the update of chunk k + 1
// (latency-bound: shared loads -> dot product -> two shuffles -> 36 FMAs -> stores) overlaps the Gram FMAs of chunk k
// (throughput-bound on the FP64 pipe and the shared-memory wavefronts) on the same SM sub-partitions.
// Registers: twelve warps are three per SM sub-partition, whose 16,384 registers give every thread 168.  The pass
// takes what it needs by value (PassArgs) instead of through the per-gene state, whose three dozen pointers would
// otherwise stay live across it, and for its duration the update warps (one warpgroup) hand registers to the Gram
// warps (two warpgroups) with setmaxnreg.  The CTA's register pool is what the launch gave it (384 x 168), so the
// hand-over has to balance: 128 threads x (168 - WS_REG_A) >= 256 threads x (WS_REG_B - 168) -- a Gram warp that asks
// for more than the update warps released would wait for ever.  With it ptxas keeps the 48 accumulators and two
// columns' operands in registers through the column loop (without it the loop spills).
struct PassArgs {
    double *ring, *M, *X, *v;
    unsigned long long *mbar;
    int *ibuf;
    FinArgs fin;
    double c;
    int n_cur;
    unsigned seq;
    int primed;
};
struct PassOut { unsigned seq; int xpar; };
template <bool UPDATE>
__device__ __forceinline__ PassOut gram_mid_ws(const PassArgs g, const bool prime_next) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = g.n_cur;
    const int nchunk = (n + MCH - 1) / MCH;
    constexpr int STG = 2 * MCH * MCS;                     // doubles per ring stage (M then x)
    constexpr unsigned CHB = MCH * MCS * 8;                // bytes of one array's chunk
    unsigned long long *full = g.mbar, *upd = g.mbar + MRING;
    volatile int *freec = g.ibuf + IB_FREE;
    const bool is_b = warp < MNW;
    // Ring bookkeeping: chunks are numbered through the whole kernel (g.seq = chunks consumed so far, the same in
    // every thread); chunk number k lives in stage k % RING and is the (k / RING)-th fill of it, which gives the
    // phase parity of the stage's two mbarriers.
    const unsigned seq0 = g.seq;

    // request chunk ch of this pass (caller: the stage is free, the slab holds the data)
    auto issue = [&](int ch, bool with_x) {
        const unsigned st = (seq0 + (unsigned)ch) % MRING;
        double *dst = g.ring + st * STG;
        fence_proxy_async_smem();
        mbar_expect_tx(full + st, with_x ? 2 * CHB : CHB);
        bulk_g2s(dst, g.M + (long long)ch * (MCH * MCS), CHB, full + st);
        if (with_x) bulk_g2s(dst + MCH * MCS, g.X + (long long)ch * (MCH * MCS), CHB, full + st);
    };
    // one consumer (a B warp, or an A warp whose bulk store has read the stage) is done with chunk ch; called by lane 0
    auto release = [&](int ch, unsigned st) {
        if (ch + MRING < nchunk) {
            __threadfence_block();
            const int old = atomicAdd(const_cast<int *>(freec + st), 1);
            if (old == MNW + MNA - 1) {
                freec[st] = 0;
                __threadfence_block();
                issue(ch + MRING, UPDATE);
            }
        }
    };
    if (!g.primed) {
        if (tid == 0)
            for (int q = 0; q < MRING && q < nchunk; ++q) issue(q, UPDATE);
    }
    // (the accumulators are zeroed inside each role's branch -- at its start for the Gram warps, at its END for the
    // update warps -- so that they are not live through the update code, whose 36 doubles of column data would
    // otherwise be spilled: local memory has next to no L1 beside 224 KB of shared memory)
    double acc[8][6];

    if (is_b) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int q = 0; q < 6; ++q) acc[r][q] = 0.0;
        // -------------------------------------------------------------------------------------------- Gram warps
#if WS_SETMAXNREG >= 2
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(WS_REG_B));
#endif
        const int tr0 = g.fin.tab[4 * lane + 2] >= 0 ? g.fin.tab[4 * lane] : -1;
        const int tc0 = g.fin.tab[4 * lane + 1];
        const int ao0 = tr0 + mrot8(0, tr0 >> 3), ao1 = tr0 + mrot8(2, tr0 >> 3), ao2 = tr0 + mrot8(4, tr0 >> 3),
                  ao3 = tr0 + mrot8(6, tr0 >> 3);
        unsigned st = seq0 % MRING, par = (seq0 / MRING) & 1u;
#pragma unroll 1
        for (int ch = 0; ch < nchunk; ++ch) {
            mbar_wait(upd + st, par);
            const double *sM = g.ring + st * STG;
            const int ncol = min(MCH, n - ch * MCH);
            if (tr0 >= 0) {
                int cc = warp * (MCH / MNW);
                const int cend = min(ncol, cc + MCH / MNW);
#define MID_LOADP(S, col)                                                                   \
    {                                                                                       \
        const double *mc_ = sM + (col) * MCS;                                               \
        S##a0 = *reinterpret_cast<const double2 *>(mc_ + ao0);                              \
        S##a1 = *reinterpret_cast<const double2 *>(mc_ + ao1);                              \
        S##a2 = *reinterpret_cast<const double2 *>(mc_ + ao2);                              \
        S##a3 = *reinterpret_cast<const double2 *>(mc_ + ao3);                              \
        S##u0 = *reinterpret_cast<const double2 *>(mc_ + tc0);                              \
        S##u1 = *reinterpret_cast<const double2 *>(mc_ + tc0 + 2);                          \
        S##u2 = *reinterpret_cast<const double2 *>(mc_ + tc0 + 4);                          \
    }
#define MID_FMAP(S)                                                                         \
    {                                                                                       \
        const double ar_[8] = {S##a0.x, S##a0.y, S##a1.x, S##a1.y, S##a2.x, S##a2.y, S##a3.x, S##a3.y}; \
        const double uc_[6] = {S##u0.x, S##u0.y, S##u1.x, S##u1.y, S##u2.x, S##u2.y};       \
        _Pragma("unroll") for (int r = 0; r < 8; ++r)                                       \
            _Pragma("unroll") for (int q = 0; q < 6; ++q) acc[r][q] = fma(ar_[r], uc_[q], acc[r][q]); \
    }
                double2 Aa0, Aa1, Aa2, Aa3, Au0, Au1, Au2, Ba0, Ba1, Ba2, Ba3, Bu0, Bu1, Bu2;
#pragma unroll 1
                for (; cc + 1 < cend; cc += 2) {
                    MID_LOADP(A, cc);
                    MID_LOADP(B, cc + 1);
                    MID_FMAP(A);
                    MID_FMAP(B);
                }
                if (cc < cend) {
                    MID_LOADP(A, cc);
                    MID_FMAP(A);
                }
#undef MID_LOADP
#undef MID_FMAP
            }
            __syncwarp();
            if (lane == 0) release(ch, st);
            if (++st == MRING) { st = 0; par ^= 1u; }
        }
#if WS_SETMAXNREG >= 2
        // (the warps of a warpgroup synchronise between two setmaxnreg instructions, as PTX requires)
        asm volatile("bar.sync %0, 128;\n" ::"r"(1 + (warp >> 2)) : "memory");
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(WS_REG_BASE));
        asm volatile("bar.sync 4, %0;\n" ::"n"(MNT) : "memory");      // "surplus returned": see the update warps
#endif
    } else {
        // -------------------------------------------------------------------------------------------- update warps
#if WS_SETMAXNREG >= 2
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(WS_REG_A));
#endif
        const int aw = warp - MNW;                            // this warp's columns of a chunk: [CPA aw, CPA (aw + 1))
        constexpr int CPA = MCH / (MNA > 0 ? MNA : 1);
        // 8 lanes per column, 6 rows each: 18 doubles of column data per thread (12 rows per lane would not fit the
        // registers an update warp keeps during the pass, and local memory has next to no L1 here)
        const int q8 = lane & 7;
        const double c = g.c;
        double vq[6];
        if constexpr (UPDATE) ld6(g.v + 6 * q8, vq);
#pragma unroll 1
        for (int ch = 0; ch < nchunk; ++ch) {
            const unsigned k = seq0 + (unsigned)ch, st = k % MRING;
            mbar_wait(full + st, (k / MRING) & 1u);
            double *sM = g.ring + st * STG;
            const int ncol = min(MCH, n - ch * MCH);
            if constexpr (UPDATE) {
#pragma unroll 1
                for (int rd = 0; rd < CPA / 4; ++rd) {
                    const int cc = aw * CPA + rd * 4 + (lane >> 3);
                    const bool act = cc < ncol;
                    double m[6], x[6];
                    double tp = 0.0;
                    if (act) {
                        ld6(sM + cc * MCS + 6 * q8, m);
                        ld6(sM + MCH * MCS + cc * MCS + 6 * q8, x);
                        double t0 = 0.0, t1 = 0.0;
#pragma unroll
                        for (int i = 0; i < 6; i += 2) { t0 = fma(vq[i], m[i], t0); t1 = fma(vq[i + 1], m[i + 1], t1); }
                        tp = t0 + t1;
                    }
                    double t = tp;
                    t += __shfl_xor_sync(0xffffffffu, t, 1);
                    t += __shfl_xor_sync(0xffffffffu, t, 2);
                    t += __shfl_xor_sync(0xffffffffu, t, 4);
                    if (act) {
#pragma unroll
                        for (int i = 0; i < 6; ++i) {
                            const double res = fma(vq[i], t, -x[i]);
                            const double w = fma(-c, res, m[i] - x[i]);
                            m[i] = fma(0.5, w + fabs(w), x[i]);
                        }
                        st6(sM + cc * MCS + 6 * q8, m);
                    }
                }
                fence_proxy_async_smem();                     // the stage is read by this warp's bulk store below
            }
            mbar_arrive(upd + st);                            // (release: the Gram warps may read the stage)
            __syncwarp();
            if (lane == 0) {
                if constexpr (UPDATE) {
                    const int c0 = aw * CPA, nmine = min(CPA, ncol - c0);
                    if (nmine > 0)
                        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(
                                         g.M + ((long long)ch * MCH + c0) * MCS),
                                     "r"(smem_u32(sM + c0 * MCS)), "r"((unsigned)(nmine * MCS * 8))
                                     : "memory");
                    asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
                    // The stage of chunk ch - WS_LAG is released once the store issued WS_LAG chunks ago has read it:
                    // waiting for the store just issued would put a TMA round trip into every chunk of this warp
                    // (measured: the update warps then set the pace of the pass).
                    bulk_wait_read<WS_LAG>();
                    if (ch >= WS_LAG) release(ch - WS_LAG, (seq0 + (unsigned)(ch - WS_LAG)) % MRING);
                } else {
                    release(ch, st);
                }
            }
        }
        if constexpr (UPDATE) {
            if (lane == 0) bulk_wait_all();                   // the slab holds this warp's columns of the new M
        }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int q = 0; q < 6; ++q) acc[r][q] = 0.0;
#if WS_SETMAXNREG >= 2
        // The update warps take their registers back only after every Gram warp has returned its surplus (named
        // barrier 4, all threads).  Re-acquiring at the end of their own loop deadlocks on short passes: an update
        // warp that is done before the second Gram warpgroup has even asked takes the registers that group waits for.
        asm volatile("bar.sync 4, %0;\n" ::"n"(MNT) : "memory");
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(WS_REG_BASE));
#endif
    }
    PassOut out;
    out.seq = (seq0 + (unsigned)nchunk) % (2 * MRING);     // (stage, parity) of chunk k depend on k mod 2 RING only
    fence_proxy_async();
    __syncthreads();

    if (prime_next && tid == 0) {
        // the next pass is an update pass over the same columns: request its first chunks now (they travel while the
        // partial Grams are reduced and the eigen-solve runs)
        for (int q = 0; q < MRING && q < nchunk; ++q) {
            const unsigned st = (out.seq + (unsigned)q) % MRING;
            double *dst = g.ring + st * STG;
            fence_proxy_async_smem();
            mbar_expect_tx(full + st, 2 * CHB);
            bulk_g2s(dst, g.M + (long long)q * (MCH * MCS), CHB, full + st);
            bulk_g2s(dst + MCH * MCS, g.X + (long long)q * (MCH * MCS), CHB, full + st);
        }
    }
    out.xpar = gram_finish(g.fin, acc);
    return out;
}

// ---- top eigenvector of G (MP x MP, shared); same rules as eig_warp in nmfoa_tiled.cu -----------------------------
// The 48 x 48 mat-vec is spread over 48 x EKS threads (thread = row i, k-range of EKR), warp 0 combines the partial
// products, normalises and tests: two barriers per step instead of one warp doing 2 x 48 FMAs and as many shared
// loads per lane while seven warps wait.
constexpr int EKS = MNT / MP;                     // k-slices of the mat-vec (5 at 256 threads, 2 at 128)
constexpr int EKR = (MP + EKS - 1) / EKS;         // entries of v per slice (10 / 24)
__device__ int eig_mid_block(const double *G, int p, double *v, double *part, int *flag, bool cold) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i = tid % MP, ks = tid / MP;
    const int k_lo = ks * EKR;
    double gk[EKR];
#pragma unroll
    for (int j = 0; j < EKR; ++j) gk[j] = (ks < EKS && k_lo + j < MP) ? G[(k_lo + j) * MP + i] : 0.0;
    if (cold) {
        // G.1 (row sums) is the first iterate: start from the ones vector over the real samples
        if (tid < MP) v[tid] = tid < p ? 1.0 : 0.0;
    }
    __syncthreads();
    int steps = 0, ok = 0;
    double prev = 1.0e300;
    for (; steps < EIG_FAST_STEPS + 1;) {
        if (ks < EKS) {
            double y = 0.0;
#pragma unroll
            for (int j = 0; j < EKR; ++j) y = fma(gk[j], (k_lo + j < MP) ? v[k_lo + j] : 0.0, y);
            part[ks * MP + i] = y;
        }
        __syncthreads();
        ++steps;
        if (warp == 0) {
            const int r0 = lane, r1 = lane + 32;
            double y0 = 0.0, y1 = 0.0;
#pragma unroll
            for (int q = 0; q < EKS; ++q) {
                y0 += part[q * MP + r0];
                if (r1 < MP) y1 += part[q * MP + r1];
            }
            const double v0 = v[r0], v1 = r1 < MP ? v[r1] : 0.0;
            const double n2 = warp_sum(y0 * y0 + y1 * y1);
            int code = 0;
            if (!(n2 > 0.0)) {
                v[r0] = 0.0;
                if (r1 < MP) v[r1] = 0.0;
                code = 1;
            } else {
                const double inv = 1.0 / sqrt(n2);
                const double w0 = y0 * inv, w1 = y1 * inv;
                const double d = warp_max(fmax(fabs(w0 - v0), fabs(w1 - v1)));
                v[r0] = w0;
                if (r1 < MP) v[r1] = w1;
                if (d <= EIG_TOL) {
                    // warm-start distrust rule (see eig_warp in nmfoa_tiled.cu)
                    const double m0 = (r0 < p && G[r0 * MP + r0] > 0.0) ? w0 : 1.0;
                    const double m1 = (r1 < p && G[r1 * MP + r1] > 0.0) ? w1 : 1.0;
                    const double vmin = -warp_max(-fmin(m0, m1));
                    const double vmax = warp_max(fmax(w0, w1));
                    code = vmin < EIG_SUSPECT * vmax ? 2 : 1;
                } else if (steps >= 9 && d > 0.75 * prev) {
                    code = 3;                      // small spectral gap: let the squaring solver finish
                }
                prev = d;
            }
            if (lane == 0) *flag = code;
        }
        __syncthreads();
        ok = *flag;
        if (ok != 0) break;
    }
    __syncthreads();
    *flag = (ok == 1) ? 1 : (ok == 2 ? 2 : 0);
    return steps;
}

__device__ void eig_mid(const KArgs &a, MGene &g, bool cold) {
    const int s0 = eig_mid_block(g.G, a.p, g.v, g.buf, g.ibuf + IB_EIG, cold);
    g.eig_steps += s0;
    __syncthreads();
    const int conv = g.ibuf[IB_EIG];
    __syncthreads();
    if (conv != 1) {                               // uniform across the CTA (and the cluster: same G everywhere)
        const int s = eig_squaring<MNT>(g.G, MP, a.p, g.v, g.red, g.B0, g.B0 + MP * MP, conv == 2);
        g.eig_steps += s;
        g.eig_fallbacks += 1;
    }
}

// ---- final pass of an nmf() call: t_j, residuals, row sums (see final_pass in nmfoa_tiled.cu) --------------------
__device__ void final_pass_mid(const KArgs &a, MGene &g, bool first, bool want_res, double *e_first_g) {
    const int tid = threadIdx.x;
    const int n = g.n_cur;
    const int q4 = tid & 3;
    double vq[12];
    ld12(g.v + 12 * q4, vq);
    double st = 0.0, st2 = 0.0;
    double sF[12], sC[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) { sF[i] = 0.0; sC[i] = 0.0; }
    // 4 lanes per column again; whole warps iterate together (64 columns per CTA step)
    const int nround = (n + MCPR - 1) / MCPR;
    for (int rd = 0; rd < nround; ++rd) {
        const int col = rd * MCPR + (tid >> 2);
        double m[12], x[12];
        double tp = 0.0;
        if (col < n) {
            ld12cg(g.M + (long long)col * MCS + 12 * q4, m);
            ld12(g.X + (long long)col * MCS + 12 * q4, x);
#pragma unroll
            for (int i = 0; i < 12; ++i) tp = fma(vq[i], m[i], tp);
        }
        double t = tp;
        t += __shfl_xor_sync(0xffffffffu, t, 1);
        t += __shfl_xor_sync(0xffffffffu, t, 2);
        double r = 0.0;
        if (col < n) {
            double bn = 0.0, bd = 1.0;                      // largest |num| / den by cross-multiplication: one division
#pragma unroll
            for (int i = 0; i < 12; ++i) {
                const double ke = vq[i] * t;
                const double kc = ke < x[i] ? x[i] : ke;
                sF[i] += x[i];
                sC[i] += kc;
                if (want_res) {
                    const double num = fabs((first ? ke : kc) - x[i]), den = x[i] + 1.0;
                    if (num * bd > bn * den) { bn = num; bd = den; }
                }
            }
            const double qv = bn / bd;
            r = qv * qv;
        }
        r = fmax(r, __shfl_xor_sync(0xffffffffu, r, 1));
        r = fmax(r, __shfl_xor_sync(0xffffffffu, r, 2));
        if (col < n && q4 == 0) {
            g.tb[col] = t;
            if (want_res) g.resb[col] = r;
            st += t;
            st2 = fma(t, t, st2);
        }
    }
    // reductions: st, st2 over the CTA; sF / sC over the lanes with the same q4 (rows 12 q4 .. 12 q4 + 11)
    const double sum_t_l = block_sum<MNT>(st, g.red);
    const double sum_t2_l = block_sum<MNT>(st2, g.red);
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        // sum over lanes with equal (lane & 3): xor 4, 8, 16
        for (int o = 4; o < 32; o <<= 1) {
            sF[i] += __shfl_xor_sync(0xffffffffu, sF[i], o);
            sC[i] += __shfl_xor_sync(0xffffffffu, sC[i], o);
        }
    }
    const int lane = tid & 31, warp = tid >> 5;
    // buf as scratch: [warp][2][MP]
    if (lane < 4) {
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            g.buf[(warp * 2 + 0) * MP + 12 * lane + i] = sF[i];
            g.buf[(warp * 2 + 1) * MP + 12 * lane + i] = sC[i];
        }
    }
    __syncthreads();
    if (tid < 2 * MP) {
        const int which = tid / MP, i = tid - which * MP;
        double s = 0.0;
        for (int w = 0; w < MNWT; ++w) s += g.buf[(w * 2 + which) * MP + i];
        g.buf[MNWT * 2 * MP + tid] = s;             // [0, MP): rsF, [MP, 2MP): rsC
    }
    if (tid == 0) { g.buf[MNWT * 2 * MP + 2 * MP] = sum_t_l; g.buf[MNWT * 2 * MP + 2 * MP + 1] = sum_t2_l; }
    __syncthreads();
    double *vals = g.buf + MNWT * 2 * MP;           // 2 MP + 2 values
    mclu_allsum(g, vals, 2 * MP + 2, vals);
    const double sum_t = vals[2 * MP], sum_t2 = vals[2 * MP + 1];
    const double sigma = sqrt(sum_t2);
    if (e_first_g != nullptr) {
        const double inv = sigma > 0.0 ? 1.0 / sigma : 0.0;
        for (int col = tid; col < n; col += MNT) e_first_g[g.goff + col] = g.tb[col] * inv;
    }
    if (tid < MP) {
        const double vi = g.v[tid];
        g.rsF[tid] = vals[tid];
        g.rsC[tid] = vals[MP + tid];
        g.tmp[tid] = vi * sum_t;
        g.K[tid] = vi * sigma;
    }
    __syncthreads();
}

__device__ void run_nmf_mid(const KArgs &a, MGene &g, bool first, bool want_res, double *e_first_g) {
    const int tid = threadIdx.x;
    // The default instantiation never copies M = x: the first fit and the first update pass read the x array.
    constexpr bool NO_COPY = (MNA == 0) && (MID_TMA_STORE != 0);
    if (!NO_COPY || a.nmf_iter == 0) {   // lambda = 0: M = x
        const double2 *src = reinterpret_cast<const double2 *>(g.X);
        double2 *dst = reinterpret_cast<double2 *>(g.M);
        const long long n2 = (long long)g.n_cur * (MCS / 2);
        for (long long e = tid; e < n2; e += MNT) dst[e] = src[e];
    }
    fence_proxy_async();                  // the slab was written with ordinary stores; the TMA reads it next
    __syncthreads();
    const int T = a.nmf_iter;
    g.primed = false;
    if constexpr (MNA > 0) {
        PassArgs pa;
        pa.ring = g.ring; pa.M = g.M; pa.X = g.X; pa.v = g.v; pa.mbar = g.mbar; pa.ibuf = g.ibuf;
        pa.c = a.c; pa.n_cur = g.n_cur;
        for (int it = -1; it < T; ++it) {
            pa.fin = fin_args(g);
            pa.seq = g.seq;
            pa.primed = g.primed ? 1 : 0;
            const bool prime_next = it + 1 < T;
            const PassOut po = it < 0 ? gram_mid_ws<false>(pa, prime_next) : gram_mid_ws<true>(pa, prime_next);
            g.seq = po.seq;
            g.xpar = po.xpar;
            g.primed = prime_next;
            eig_mid(a, g, it < 0);
        }
    } else {
        const bool lazy_m = NO_COPY && T > 0;
        gram_mid<false>(a, g, T > 0, lazy_m);
        eig_mid(a, g, true);
        for (int it = 0; it < T; ++it) {
            gram_mid<true>(a, g, it + 1 < T, lazy_m && it == 0);
            eig_mid(a, g, false);
        }
    }
    final_pass_mid(a, g, first, want_res, e_first_g);
}

// (warp-specialised instantiation: 12 warps = three per SM sub-partition, whose 16,384 registers allow 168 per thread)
__global__ void __launch_bounds__(MNT, MNW <= 4 ? 2 : 1) nmfoa_mid_kernel(const KArgs a) {
    extern __shared__ double smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int p = a.p;
    const MidCarve cv = mid_carve(MNW);
    MGene g;
    double *sm = smem + cv.small;
    g.v = sm;            g.K = sm + MP;        g.K0 = sm + 2 * MP;   g.rs0 = sm + 3 * MP;   g.rsF = sm + 4 * MP;
    g.rsC = sm + 5 * MP; g.rsC0 = sm + 6 * MP; g.rho = sm + 7 * MP;  g.scale = sm + 8 * MP; g.tmp = sm + 9 * MP;
    g.red = smem + cv.red;
    g.binm = smem + cv.binm;
    g.alive = reinterpret_cast<int *>(smem + cv.alive);
    g.ibuf = reinterpret_cast<int *>(smem + cv.ibuf);
    g.lw = reinterpret_cast<int *>(smem + cv.lw);
    g.tab = reinterpret_cast<int *>(smem + cv.tab);
    g.G = smem + cv.G;
    g.buf = smem + cv.buf;
    g.ring = smem + cv.ring;
    g.mbar = reinterpret_cast<unsigned long long *>(smem + cv.mbar);
    g.use0 = g.use1 = g.use2 = 0;
    g.seq = 0;
    if (tid == 0) {
        for (int q = 0; q < MRING; ++q) mbar_init(g.mbar + q, 1);
        // warp-specialised instantiation: "updated" barriers (every update thread arrives) and the consumer counters
        for (int q = 0; q < MRING; ++q) mbar_init(g.mbar + MRING + q, MNA > 0 ? MNA * 32 : 1);
        for (int q = 0; q < MRING; ++q) g.ibuf[IB_FREE + q] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    cg::cluster_group cl = cg::this_cluster();
    g.crank = (int)cl.block_rank();
    g.csize = (int)cl.num_blocks();
    g.xpar = 0;
    const long long wcols = a.ws_ld;                           // columns one CTA can hold
    double *slab = a.ws + (long long)blockIdx.x * a.ws_stride;
    g.slot_stride = a.ws_stride;
    g.B0 = slab;
    g.slots = slab + 2 * MP * MP;
    g.X = g.slots + 2 * MNE;
    g.M = g.X + (long long)MCS * wcols;
    g.resb = g.M + (long long)MCS * wcols;
    g.tb = g.resb + wcols;
    if (tid == 0) {
        // lane slots (4 ints each: row offset, column offset, dense tile id or -1, -) and, from int 128 on, the
        // (row, column) offsets of the dense tiles
        int slot = 0, tile = 0;
        for (int rb = 0; rb < 6; ++rb) {
            const int c_lo = (8 * rb) / 6, cnt = 8 - c_lo;           // first column block with 6 cb + 5 >= 8 rb
            for (int c = 0; c < cnt + (cnt & 1); ++c, ++slot) {
                const bool real = c < cnt;
                const int cb = c_lo + (real ? c : cnt - 1);
                g.tab[4 * slot] = 8 * rb;
                g.tab[4 * slot + 1] = 6 * cb;
                g.tab[4 * slot + 2] = real ? tile : -1;
                g.tab[4 * slot + 3] = 0;
                if (real) { g.tab[128 + 2 * tile] = 8 * rb; g.tab[128 + 2 * tile + 1] = 6 * cb; ++tile; }
            }
        }
    }
    for (int e = tid; e < N_SMALL * MP; e += MNT) sm[e] = 0.0;
    for (int e = tid; e < MP * MP; e += MNT) g.G[e] = 0.0;
    g.eig_steps = 0;
    g.eig_fallbacks = 0;
    __syncthreads();
    cl.sync();

    for (;;) {
        if (g.crank == 0 && tid == 0) {
            const int t = atomicAdd(a.queue, 1);
            for (int r = 0; r < g.csize; ++r) *cl.map_shared_rank(g.ibuf, r) = t;
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

        if (tid < MP) g.scale[tid] = tid < p ? a.scale[tid] : 1.0;
        __syncthreads();
        double tmax = -1.0e300;
        if (a.row_max) {
            if (tid < p) tmax = a.row_max[(long long)gid * p + tid] / g.scale[tid];
        } else {
            for (int i = 0; i < p; ++i) {
                const double *row = F + (long long)i * L;
                double m = -1.0e300;
                for (int j = tid; j < L; j += MNT) m = fmax(m, row[j]);
                tmax = fmax(tmax, m / g.scale[i]);
            }
        }
        const double gmax = block_max<MNT>(tmax, g.red);
        const double thr = (a.flags & DN_FLAG_PLAIN_NMF) ? -1.0e300 : 0.1 * gmax;                                   // nmf.py:76
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
            // keep + compact (one thread per candidate column; the 48 scaled values go straight to the slab)
            int running = 0;
            int *wcount = g.ibuf + IB_WCOUNT;
            for (int kb = k_lo; kb < k_hi; kb += MNT) {
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
                for (int q = 0; q < MNWT; ++q) {
                    const int cq = wcount[q];
                    if (q < warp) pre += cq;
                    tot += cq;
                }
                if (keep) {
                    const int dst = pre + __popc(bal & ((1u << lane) - 1u));
                    double *xc = g.X + (long long)dst * MCS;
                    for (int i = 0; i < MCS; ++i) xc[i] = i < p ? F[(long long)i * L + col] / g.scale[i] : 0.0;
                }
                running += tot;
                __syncthreads();
            }
            g.n0 = g.n_cur = running;
            n0 = running;
            if (tid < g.csize) g.binm[tid] = tid == g.crank ? (double)running : 0.0;
            __syncthreads();
            mclu_allsum(g, g.binm, g.csize, g.binm);
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
            exit_code = DN_EXIT_FEW_HICOV;
        } else {
            g.cs = n0; g.nb0 = 1; g.nalive = 1;
            if (tid == 0) { g.alive[0] = 0; g.lw[0] = g.n0; }
            {   // rs(F_start): thread per row-quarter of a column, as in the final pass
                const int q4 = tid & 3;
                double rs[12];
#pragma unroll
                for (int i = 0; i < 12; ++i) rs[i] = 0.0;
                for (int col = tid >> 2; col < g.n0; col += MCPR) {
                    double x[12];
                    ld12(g.X + (long long)col * MCS + 12 * q4, x);
#pragma unroll
                    for (int i = 0; i < 12; ++i) rs[i] += x[i];
                }
#pragma unroll
                for (int i = 0; i < 12; ++i)
                    for (int o = 4; o < 32; o <<= 1) rs[i] += __shfl_xor_sync(0xffffffffu, rs[i], o);
                __syncthreads();
                if (lane < 4) {
#pragma unroll
                    for (int i = 0; i < 12; ++i) g.buf[warp * MP + 12 * lane + i] = rs[i];
                }
                __syncthreads();
                if (tid < MP) {
                    double s = 0.0;
                    for (int w2 = 0; w2 < MNWT; ++w2) s += g.buf[w2 * MP + tid];
                    g.rs0[tid] = s;
                }
                __syncthreads();
                mclu_allsum(g, g.rs0, MP, g.rs0);
            }
            bool any_empty = false;
            for (int i = 0; i < p; ++i) any_empty |= !(g.rs0[i] > 0.0);
            if (any_empty && !(a.flags & DN_FLAG_PLAIN_NMF)) {
                exit_code = DN_EXIT_EMPTY_SAMPLE;
            } else {
                const bool store_e = (a.e_first != nullptr) && (n0 == L);
                bool first = true, in_loop = false;
                double rmax = 0.0;
                for (;;) {
                    run_nmf_mid(a, g, first, true, (first && store_e) ? a.e_first + o0 : nullptr);
                    nmf_calls += 1; sum_cols += g.n_curg;
                    if (first) {
                        if (tid < MP) {
                            g.rho[tid] = 1.0 - g.rs0[tid] / (g.tmp[tid] + 1.0);
                            g.K0[tid] = g.K[tid];
                            g.rsC0[tid] = g.rsC[tid];
                        }
                        __syncthreads();
                        if (!(a.flags & DN_FLAG_PLAIN_NMF) && median_one_minus(g.rho, p) > 1.0) { exit_code = DN_EXIT_MEDIAN; break; }
                        double rmin = g.rho[0];
                        rmax = g.rho[0];
                        for (int i = 1; i < p; ++i) { rmin = fmin(rmin, g.rho[i]); rmax = fmax(rmax, g.rho[i]); }
                        if (!(n0 >= a.min_len && rmin <= 0.2 && !a.skip)) { exit_code = DN_EXIT_NO_SELECTION; break; }
                        g.cs = (n0 + a.bins - 1) / a.bins;
                        g.nb0 = (n0 + g.cs - 1) / g.cs;
                        g.nalive = g.nb0;
                        for (int b = tid; b < g.nb0; b += MNT) {
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
                        if (mn == 0.0) break;
                        __syncthreads();
                        if (tid < MP) g.rho[tid] = 1.0 - g.rsF[tid] / (g.rsC[tid] + 1.0);
                        __syncthreads();
                        rmax = g.rho[0];
                        for (int i = 1; i < p; ++i) rmax = fmax(rmax, g.rho[i]);
                        if (g.nalive <= a.min_bins || g.n_curg < a.min_len) break;
                    }
                    if (!(rmax > 0.1)) break;
                    ran = 1;
                    for (int k = warp; k < g.nalive; k += MNWT) {
                        const int wl = g.lw[g.alive[k]];
                        const double *rr = g.resb + mlstart(g, k);
                        double s = 0.0;
                        for (int j = lane; j < wl; j += 32) s += rr[j];
                        s = warp_sum(s);
                        if (lane == 0) g.binm[k] = s;
                    }
                    __syncthreads();
                    mclu_allsum(g, g.binm, g.nalive, g.binm);
                    int kd = 0;
                    double best = -1.0;
                    for (int k = 0; k < g.nalive; ++k) {
                        const int b = g.alive[k];
                        const double mean = g.binm[k] / (double)min(g.cs, n0 - b * g.cs);
                        if (mean > best) { best = mean; kd = k; }
                    }
                    if (best == 0.0) break;
                    const int bd = g.alive[kd];
                    const int wd = min(g.cs, n0 - bd * g.cs);
                    {   // rotate this CTA's share of the dropped bin to the end of its current columns (M is scratch)
                        const int a0 = mlstart(g, kd);
                        const int wl = g.lw[bd];
                        const int tail = g.n_cur - a0 - wl;
                        const int h = MCS / 2;
                        const double2 *xs = reinterpret_cast<const double2 *>(g.X + (long long)a0 * MCS);
                        double2 *ms = reinterpret_cast<double2 *>(g.M + (long long)a0 * MCS);
                        double2 *xd = reinterpret_cast<double2 *>(g.X + (long long)a0 * MCS);
                        __syncthreads();
                        for (long long e = tid; e < (long long)(tail + wl) * h; e += MNT) ms[e] = xs[e];
                        __syncthreads();
                        for (long long e = tid; e < (long long)tail * h; e += MNT) xd[e] = ms[(long long)wl * h + e];
                        for (long long e = tid; e < (long long)wl * h; e += MNT) xd[(long long)tail * h + e] = ms[e];
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
                    if (rmax < 0.2) {
                        floor_abs(g.K, g.K, p);
                        double s = 0.0;
                        for (int j = tid; j < g.n0; j += MNT) {
                            const double *xc = g.X + (long long)j * MCS;
                            double e = -1.0e300;
                            for (int i = 0; i < p; ++i) e = fmax(e, xc[i] / g.K[i]);
                            s += e;
                        }
                        double S = block_sum<MNT>(s, g.red);
                        if (tid == 0) g.binm[0] = S;
                        __syncthreads();
                        mclu_allsum(g, g.binm, 1, g.binm);
                        S = g.binm[0];
                        __syncthreads();
                        if (tid < MP) g.rho[tid] = 1.0 - g.rs0[tid] / (g.K[tid] * S + 1.0);
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
                    if (fallback) {
                        __syncthreads();
                        if (tid < MP) g.rho[tid] = 1.0 - g.rs0[tid] / (g.rsC0[tid] + 1.0);
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
                if (tid < MP) g.K[tid] = g.K0[tid];
                __syncthreads();
            } else {
                floor_abs(g.K0, g.K, p);
            }
        }
        if (g.crank == 0) {
            if (tid < p) {
                double r = is_default ? 0.0 : g.rho[tid];
                if (!(a.flags & DN_FLAG_RAW_RHO)) r = r > 0.9 ? 0.9 : r;
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
                           a.est + (long long)p * (a.est_off ? a.est_off[gid] : o0), g.crank * MNT + tid, g.csize * MNT);
            __syncthreads();
        }
        cl.sync();
    }
}

}  // namespace

int MID_LAUNCHER(const KArgs &a, const dn_plan *plan, cudaStream_t st) {
    auto kern = nmfoa_mid_kernel;
    DN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, plan->smem_bytes));
    if (plan->cluster > 8) DN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    const int cl = plan->cluster > 0 ? plan->cluster : 1;
    cfg.gridDim = dim3(plan->ctas / cl * cl, 1, 1);
    cfg.blockDim = dim3(MNT, 1, 1);
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
