// degnorm_b200 -- launch geometry shared by the host planner (abi.cu) and the kernel translation units.
#pragma once
#include "common.cuh"

// ---- tiled kernel (any p; nmfoa_tiled.cu): shared-memory carve-up ---------------------------------------------
struct Carve {
    long long small, red, binm, alive, ibuf, tp, G, ms, xr, lr, resb, tb, total;   // offsets in doubles
};

__host__ __device__ inline Carve carve(int p, int pp, int g_in_smem, long long ms_doubles, int resident_cols, int ld_res) {
    Carve c;
    long long o = 0;
    c.small = o; o += (long long)N_SMALL * pp;
    c.red = o;   o += 64;
    c.binm = o;  o += DN_MAX_BINS;
    c.alive = o; o += DN_MAX_BINS / 2;
    c.ibuf = o;  o += 16;
    // phase A's partial dot products (one per thread): their own 256 doubles when G lives in the global slab; when G
    // is in shared memory they borrow its first 256 entries (G is dead between the eigen-solve and the end of the next
    // Gram pass, which rewrites every entry; row slices exist only for pp > 24, so G holds at least 576 doubles)
    c.tp = o;    o += g_in_smem ? 0 : 256;
    c.G = o;     o += g_in_smem ? (long long)pp * pp : 0;
    c.ms = o;    o += ms_doubles;
    c.xr = o;    o += resident_cols > 0 ? (long long)p * ld_res : 0;
    c.lr = o;    o += resident_cols > 0 ? (long long)p * ld_res : 0;
    c.resb = o;  o += resident_cols > 0 ? ld_res : 0;
    c.tb = o;    o += resident_cols > 0 ? ld_res : 0;
    c.total = o;
    return c;
}

struct Derived { int tr, nt, pp, ntiles, ch, ldm, ks, g_in_smem, ms_doubles, nsets; long long fixed_doubles, gacc_doubles; };

inline int tile_for_p(int p) { return p <= 64 ? 4 : 8; }

inline Derived derive(int p) {
    Derived d;
    memset(&d, 0, sizeof(d));
    d.tr = tile_for_p(p);
    d.nt = 256;
    d.pp = (p + d.tr - 1) / d.tr * d.tr;
    const int ntg = d.pp / d.tr;
    d.ntiles = ntg * (ntg + 1) / 2;
    int ch = (48 * 1024 / 8) / d.pp / 32 * 32;
    if (ch < 32) ch = 32;
    if (ch > d.nt) ch = d.nt;
    d.ch = ch;
    // M tile: row-major with a padded row (ch + 1) for the 4 x 4 tiles of p <= 64; column-major with an even column
    // stride (128-bit operand loads) for the 8 x 8 tiles of p > 64
    d.ldm = d.tr == 8 ? d.pp + 2 : ch + 1;
    d.nsets = (d.ntiles + d.nt - 1) / d.nt;
    d.gacc_doubles = d.nsets > 1 ? ((long long)d.ntiles * d.tr * d.tr + 31) / 32 * 32 : 0;
    int ks = d.nt / d.ntiles;
    const long long msd = d.tr == 8 ? (long long)ch * d.ldm : (long long)d.pp * d.ldm;
    const long long cap = msd / ((long long)d.ntiles * d.tr * d.tr);
    if (ks > cap) ks = (int)cap;
    if (ks < 1) ks = 1;
    d.ks = ks;
    d.g_in_smem = d.pp <= 64;
    d.ms_doubles = (int)msd;
    d.fixed_doubles = carve(p, d.pp, d.g_in_smem, d.ms_doubles, 0, 0).total;
    return d;
}

// ---- small-p kernel (p <= 12; nmfoa_small.cuh): column-major x / M with a padded column stride ----------------
// P = p rounded up to 4, 8 or 12 (0: not on this path).  A column holds P doubles + 2 of padding so that a warp
// reading one column per lane with 128-bit loads touches every bank exactly once (stride = 4 mod 8 words).
__host__ __device__ inline int small_P(int p) { return p <= 4 ? 4 : (p <= 8 ? 8 : (p <= 12 ? 12 : 0)); }
__host__ __device__ inline int small_cs(int P) { return P + 2; }
constexpr int SMALL_GPART = 96;       // doubles per warp of partial Gram / reduction scratch
constexpr int SMALL_MAX_WARPS = 16;
constexpr int SMALL_CLMAX = 16;       // largest thread-block cluster per gene (non-portable size)
constexpr int SMALL_CLU_WARPS = 8;    // warps per CTA of the cluster kernels
constexpr int SMALL_RING = 3;         // streamed tier: cp.async ring stages per warp (RING - 1 blocks in flight)

struct SmallCarve {
    long long small, red, binm, alive, ibuf, G, vx, gpart, tab, lw, xbuf, stage, ring, X, M, resb, tb, total;   // offsets in doubles
};

__host__ __device__ inline SmallCarve small_carve(int P, int nw, int resident_cols, bool clu = false) {
    SmallCarve c;
    long long o = 0;
    c.small = o; o += (long long)N_SMALL * P;
    c.red = o;   o += 64;
    c.binm = o;  o += DN_MAX_BINS;
    c.alive = o; o += DN_MAX_BINS / 2;
    c.ibuf = o;  o += 16;
    c.G = o;     o += (long long)nw * P * P;                 // one copy per warp
    c.vx = o;    o += (long long)nw * 2 * P;
    c.gpart = o; o += (long long)(nw > 1 ? 2 : 1) * nw * SMALL_GPART;   // two alternating partial-sum buffers
    c.tab = o;   o += 32;                       // 16 tiles x 4 ints
    c.lw = o;    o += DN_MAX_BINS / 2;          // per-bin local widths (ints)
    c.xbuf = o;  o += clu ? 2ll * SMALL_CLMAX * SMALL_GPART : 0;   // cluster exchange slots
    const long long cs = small_cs(P);
    c.stage = o; o += resident_cols > 0 ? 0 : (long long)nw * 32 * cs;   // streamed tier: per-warp stage of M
    c.ring = o;  o += resident_cols > 0 ? 0 : (long long)nw * SMALL_RING * 2 * 32 * P;   // and cp.async ring
    c.X = o;     o += resident_cols > 0 ? cs * resident_cols : 0;
    c.M = o;     o += resident_cols > 0 ? cs * resident_cols : 0;
    c.resb = o;  o += resident_cols > 0 ? resident_cols : 0;
    c.tb = o;    o += resident_cols > 0 ? resident_cols : 0;
    c.total = o;
    return c;
}

// per-CTA global slab of the small path (doubles): eigen fallback scratch + streamed column storage
__host__ __device__ inline long long small_slab_doubles(int P, long long ws_cols) {
    // x and M blocked row-major (P doubles per column, whole 32-column blocks), residuals and t one double each
    const long long cols = (ws_cols + 31) / 32 * 32;
    const long long d = 2ll * P * P + (2ll * P + 2) * cols;
    return (d + 31) / 32 * 32;
}

// ---- mid-p kernel (13 <= p <= 48; nmfoa_mid.cuh): streamed, CTA-level TMA ring of (8 x warps)-column chunks ------
// Three instantiations: 8 warps (default: one CTA per SM, every warp updates and accumulates its own columns, one
// block barrier per 64-column chunk), 4 warps (two such CTAs per SM, 32-column chunks) and warp-specialised (8 Gram
// warps + 4 update warps, one CTA per SM, 32-column chunks in a 6-stage ring, no block barrier inside a pass; measured
// no faster than the 8-warp one: profiles/r02_mid_variants.md).
constexpr int MID_P = 48;             // samples padded to this
constexpr int MID_WARPS = 8;          // Gram warps per CTA
constexpr int MID_UPD_WARPS = 4;      // update warps of the warp-specialised instantiation
constexpr int MID_RING = 3;           // ring stages (RING - 1 chunks in flight)
#ifndef MID_WS_CHUNK_COLS             // (tuning builds: -DMID_WS_CHUNK_COLS=64 -DMID_WS_STAGES=3 -DMID_WS_LAG=1)
#define MID_WS_CHUNK_COLS 32
#define MID_WS_STAGES 6
#define MID_WS_LAG 3
#endif
constexpr int MID_WS_CHUNK = MID_WS_CHUNK_COLS;   // warp-specialised instantiation: columns per ring stage ...
constexpr int MID_WS_RING = MID_WS_STAGES;        // ... and stages (same bytes as 3 stages of 64 columns)
constexpr int MID_NE = 30 * 48;       // partial Gram sums (30 tiles of 6 x 8)
__host__ __device__ constexpr int mid_chunk(int nw) { return 8 * nw; }   // columns per ring stage

struct MidCarve { long long small, red, binm, alive, ibuf, lw, tab, mbar, G, buf, ring, total; };

__host__ __device__ inline MidCarve mid_carve(int nw) {
    MidCarve c;
    long long o = 0;
    c.small = o; o += (long long)N_SMALL * MID_P;
    c.red = o;   o += 64;
    c.binm = o;  o += DN_MAX_BINS;
    c.alive = o; o += DN_MAX_BINS / 2;
    c.ibuf = o;  o += 16;
    c.lw = o;    o += DN_MAX_BINS / 2;
    c.tab = o;   o += 96;                       // 32 lane slots x 4 ints + 30 tiles x 2 ints (nmfoa_mid.cuh)
    c.mbar = o;  o += 16;                       // "chunk landed" + "chunk updated" mbarriers (up to 6 + 6)
    const long long stage = 2ll * mid_chunk(nw) * (MID_P + 2);
    // 4-warp CTAs (two per SM): G lives in the ring's last stage, which is free between two passes (the ring is
    // primed with chunks 0 and 1 only) -- the eigen-solve is the only user of G
    c.G = nw <= 4 ? o + (nw / 2) * (long long)MID_NE + (MID_RING - 1) * stage : o;
    o += nw <= 4 ? 0 : (long long)MID_P * MID_P;
    c.buf = o;   o += (nw / 2) * (long long)MID_NE;
    c.ring = o;  o += (long long)MID_RING * stage;
    c.total = o;
    return c;
}

// per-CTA slab (doubles): eigen fallback scratch, two exchange slots, x, M, residuals, t
__host__ __device__ inline long long mid_slab_doubles(long long ws_cols) {
    const long long d = 2ll * MID_P * MID_P + 2ll * MID_NE + (2ll * (MID_P + 2) + 2) * ws_cols;
    return (d + 31) / 32 * 32;
}

// ---- wide kernel (49 <= p <= 208; nmfoa_wide.cu): one 8 x 8 Gram tile per thread, streamed through a TMA ring ---------
constexpr int WIDE_THREADS = 384;     // 12 warps: three per SM sub-partition
constexpr int WIDE_CHUNK = 16;        // columns per ring stage
constexpr int WIDE_RING = 3;
constexpr int WIDE_MAX_PP = 208;      // 26 x 27 / 2 = 351 tiles <= threads: the triangle of 208 samples fills the register file
constexpr int WIDE_MIN_P = 49;
constexpr int WIDE_MAX_KS = 8;        // k-slices of a chunk's columns when there are fewer tiles than threads
constexpr int WIDE_NSMALL = 11;       // pp-sized shared vectors
#ifndef WIDE_OVERLAP                  // 1: the twelfth warp runs the update of chunk ch + 1 while the others accumulate chunk ch
#define WIDE_OVERLAP 0                //    (one block barrier per chunk) -- measured SLOWER: one warp cannot update 16 columns of
#endif                                //    200 rows in the time eleven warps need for their tiles (C5 sample 13.0 -> 7.7 genes/s)
__host__ __device__ inline int wide_pp(int p) { return (p + 7) / 8 * 8; }
__host__ __device__ inline int wide_tiles(int pp) { return (pp / 8) * (pp / 8 + 1) / 2; }
// lane slots of one k-slice: the tiles row-major, every tile row padded to an even length (nmfoa_wide.cu)
__host__ __device__ inline int wide_slots(int nb) {
    int s = 0;
    for (int ti = 0; ti < nb; ++ti) s += (nb - ti) + ((nb - ti) & 1);
    return s;
}
__host__ __device__ inline int wide_kslices(int nb) {
    int ks = (WIDE_THREADS - (WIDE_OVERLAP ? 32 : 0)) / wide_slots(nb);   // (overlapped schedule: one warp kept free of tiles)
    return ks < 1 ? 1 : (ks > WIDE_MAX_KS ? WIDE_MAX_KS : ks);
}
struct WideCarve { long long small, red, binm, alive, ibuf, lw, mbar, part, ring, total; };
__host__ __device__ inline WideCarve wide_carve(int pp) {
    WideCarve c;
    long long o = 0;
    c.small = o; o += (long long)WIDE_NSMALL * pp;
    c.red = o;   o += 64;
    c.binm = o;  o += DN_MAX_BINS;
    c.alive = o; o += DN_MAX_BINS / 2;
    c.ibuf = o;  o += 16;
    c.lw = o;    o += DN_MAX_BINS / 2;
    c.mbar = o;  o += 4;
    c.part = o;  o += (long long)wide_tiles(pp) * 16;       // partial products of the register-tile mat-vec
    c.ring = o;  o += (long long)WIDE_RING * 2 * WIDE_CHUNK * (pp + 2);
    c.total = o;
    return c;
}
// per-CTA slab (doubles): G + two squaring matrices, two exchange slots, k-slice scratch, x, M, residuals, t
__host__ __device__ inline long long wide_slab_doubles(int pp, long long ws_cols) {
    const long long ne = (long long)wide_tiles(pp) * 64;
    const long long d = 3ll * pp * pp + 2 * ne + (WIDE_MAX_KS - 1) * ne + (2ll * (pp + 2) + 2) * ws_cols;
    return (d + 31) / 32 * 32;
}

// launchers (each defined in its own translation unit)
int dn_launch_wide(const KArgs &a, const dn_plan *plan, cudaStream_t st);
int dn_launch_mid8(const KArgs &a, const dn_plan *plan, cudaStream_t st);
int dn_launch_mid4(const KArgs &a, const dn_plan *plan, cudaStream_t st);
int dn_launch_midws(const KArgs &a, const dn_plan *plan, cudaStream_t st);
int dn_launch_tiled(const KArgs &a, const dn_plan *plan, cudaStream_t st);
int dn_launch_small4(const KArgs &a, const dn_plan *plan, cudaStream_t st);
int dn_launch_small8(const KArgs &a, const dn_plan *plan, cudaStream_t st);
int dn_launch_small12(const KArgs &a, const dn_plan *plan, cudaStream_t st);
