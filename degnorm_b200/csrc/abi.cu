// degnorm_b200 -- C ABI, host-side planning, and the small n x p / estimate kernels.
#include <stdlib.h>
#include "common.cuh"
#include "launch.h"

namespace {
thread_local char g_err[512] = "";
}  // namespace

int dn_fail(int code, const char *fmt, const char *a, long long b, long long c) {
    snprintf(g_err, sizeof(g_err), fmt, a, b, c);
    return code;
}

namespace {

// ---- estimates (last outer iteration only): nmf.py:217, 247, 333-337, 343-344, 350-351, 358-365 -----------------
__global__ void __launch_bounds__(256) estimates_kernel(const double *cov, const long long *off, const int *order,
                                                        int n_work, int p, const double *scale, const int *counters,
                                                        const double *kfac, const double *e_first,
                                                        const long long *est_off, double *est) {
    for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
        const int gid = order[w];
        const long long o0 = off[gid];
        const int L = (int)(off[gid + 1] - o0);
        const double *F = cov + (long long)p * o0;
        double *out = est + (long long)p * (est_off ? est_off[gid] : o0);
        const int ex = counters[(long long)gid * DN_NCOUNTERS + DN_CNT_EXIT];
        const int n0 = counters[(long long)gid * DN_NCOUNTERS + DN_CNT_N_HICOV];
        const double *K = kfac + (long long)gid * p;
        write_estimate(F, L, p, scale, ex, n0, K, e_first ? e_first + o0 : nullptr, out, threadIdx.x, blockDim.x);
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
#define fail dn_fail

int check_params(const dn_params *prm) {
    if (!prm) return fail(DN_ERR_INVALID, "null params%s");
    if (prm->p < 2) return fail(DN_ERR_INVALID, "need at least 2 samples%s (p = %lld)", "", prm->p);
    if (prm->p > DN_MAX_SAMPLES) return fail(DN_ERR_UNSUPPORTED, "%sp = %lld exceeds DN_MAX_SAMPLES", "", prm->p);
    if (prm->bins < 1 || prm->bins > DN_MAX_BINS) return fail(DN_ERR_UNSUPPORTED, "%sbins = %lld outside [1, DN_MAX_BINS]", "", prm->bins);
    if (prm->downsample_rate < 1) return fail(DN_ERR_INVALID, "downsample_rate must be >= 1%s");
    if (prm->nmf_iter < 0) return fail(DN_ERR_INVALID, "nmf_iter must be >= 0%s");
    return DN_OK;
}

// Warps per CTA of the small-p path by resident tier: a warp sweeps its own 32-column blocks, so small genes get
// one warp (no block-wide barrier anywhere in the inner iteration) and larger tiers enough warps for ~2-4 blocks each.
int warps_for_tier(long long cols) {
    if (cols <= 96) return 1;
    if (cols <= 192) return 2;
    if (cols <= 448) return 4;
    if (cols <= 1024) return 8;
    return 16;
}

int run_kernel(int mode, const double *cov, const int64_t *off, const int32_t *order, int32_t n_work,
               const dn_params *prm, const dn_plan *plan, const double *scale, const int32_t *ds_start,
               const double *row_max, double *row_max_out, double *rho,
               uint8_t *ran, int32_t *counters, double *kfac, double *e_first, double *est, const int64_t *est_off,
               double *est_rowsum, double *cov_rowsum,
               void *workspace, int64_t workspace_bytes, void *stream) {
    int rc = check_params(prm);
    if (rc) return rc;
    if (!plan || !cov || !off || !order) return fail(DN_ERR_INVALID, "null pointer argument%s");
    if (n_work <= 0) return DN_OK;
    if (workspace_bytes < plan->ws_bytes || !workspace) return fail(DN_ERR_WORKSPACE, "workspace too small%s: need %lld, got %lld", "", plan->ws_bytes, workspace_bytes);
    cudaStream_t st = (cudaStream_t)stream;
    KArgs a;
    memset(&a, 0, sizeof(a));
    a.cov = cov; a.off = (const long long *)off; a.order = order; a.n_work = n_work;
    a.p = prm->p; a.scale = scale; a.ds_start = ds_start; a.mode = mode;
    a.nmf_iter = mode == MODE_INIT ? 0 : prm->nmf_iter;
    a.c = prm->nmf_iter > 0 ? 1.0 / sqrt((double)prm->nmf_iter) : 0.0;       // nmf.py:91
    a.bins = prm->bins; a.min_bins = prm->min_bins; a.min_hi = prm->min_high_coverage;
    a.rate = mode == MODE_INIT ? 1 : prm->downsample_rate; a.skip = prm->skip_baseline_selection;
    a.min_len = prm->min_gene_len;
    a.flags = mode == MODE_INIT ? 0 : prm->flags;
    a.rho = rho; a.ran = ran; a.counters = counters; a.kfac = kfac; a.e_first = e_first;
    a.est_rowsum = est_rowsum; a.cov_rowsum = cov_rowsum;
    a.row_max = row_max; a.row_max_out = row_max_out;
    a.est = est; a.est_off = (const long long *)est_off;
    a.resident_cols = plan->resident_cols;
    a.ld_res = plan->resident_cols;
    // workspace: [queue (256 B)] [per-CTA slabs]
    a.queue = (int *)workspace;
    a.ws = (double *)((char *)workspace + 256);
    a.ws_ld = plan->ws_cols;
    DN_CUDA(cudaMemsetAsync(a.queue, 0, 256, st));
    if (plan->tile == 6) {
        // mid-p path (baseline selection only)
        if (mode != MODE_BS || prm->p <= 12 || prm->p > MID_P) return fail(DN_ERR_INVALID, "plan does not match params (use dn_make_plan)%s");
        a.pp = MID_P;
        a.ws_stride = mid_slab_doubles(plan->ws_cols);
        if (plan->threads == 128) return dn_launch_mid4(a, plan, st);
        if (plan->threads == 256) return dn_launch_mid8(a, plan, st);
        return dn_launch_midws(a, plan, st);
    }
    if (plan->tile == 8 && plan->threads == WIDE_THREADS) {
        // wide path (49..208 samples, baseline selection only)
        if (mode != MODE_BS || prm->p < WIDE_MIN_P || wide_pp(prm->p) > WIDE_MAX_PP) return fail(DN_ERR_INVALID, "plan does not match params (use dn_make_plan)%s");
        a.pp = wide_pp(prm->p);
        a.ws_stride = wide_slab_doubles(a.pp, plan->ws_cols);
        return dn_launch_wide(a, plan, st);
    }
    if (plan->tile == 0) {
        // small-p path (baseline selection only)
        const int P = small_P(prm->p);
        if (mode != MODE_BS || P == 0) return fail(DN_ERR_INVALID, "plan does not match params (use dn_make_plan)%s");
        a.pp = P;
        a.eig_hint = 1;
        {
            const char *e = getenv("DN_EIG_SHARED");       // tuning switch (default: redundant solves)
            a.eig_shared = e ? atoi(e) : 0;
        }
        a.ws_stride = small_slab_doubles(P, plan->ws_cols);
        if (P == 4) return dn_launch_small4(a, plan, st);
        if (P == 8) return dn_launch_small8(a, plan, st);
        return dn_launch_small12(a, plan, st);
    }
    const Derived d = derive(prm->p);
    if (plan->tile != d.tr || plan->threads != d.nt || plan->chunk_cols != d.ch)
        return fail(DN_ERR_INVALID, "plan does not match params (use dn_make_plan)%s");
    a.pp = d.pp;
    a.ch = d.ch; a.ldm = d.ldm; a.ks = d.ks; a.g_in_smem = d.g_in_smem; a.ms_doubles = d.ms_doubles;
    a.nsets = d.nsets; a.gacc_doubles = d.gacc_doubles;
    const long long g_d = (d.g_in_smem ? 2ll : 3ll) * d.pp * d.pp + d.gacc_doubles;
    const long long cols_d = mode == MODE_INIT ? plan->ws_cols : (2ll * prm->p + 2) * plan->ws_cols;
    a.ws_stride = (g_d + cols_d + 31) / 32 * 32;
    return dn_launch_tiled(a, plan, st);
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
                 int32_t warps, int32_t cluster, int32_t sm_count, int32_t max_smem_optin, dn_plan *plan) {
    int rc = check_params(prm);
    if (rc) return rc;
    if (!plan || max_cols < 1 || sm_count < 1 || max_smem_optin < 16 * 1024) return fail(DN_ERR_INVALID, "bad planning argument%s");
    memset(plan, 0, sizeof(*plan));
    plan->cluster = 1;
    if (!for_init && cluster >= 0 && prm->p > 12 && prm->p <= MID_P) {
        // ---- mid-p path (13..48 samples): streamed kernel, optional cluster per gene
        int cl = cluster > 1 ? cluster : 1;
        if (cl != 1 && cl != 2 && cl != 4 && cl != 8 && cl != 16) return fail(DN_ERR_INVALID, "cluster must be 1, 2, 4, 8 or 16%s");
        // default (warps = 0 or 8): one 8-warp CTA per SM, every warp updates and accumulates its own columns (measured
        // fastest: profiles/); warps = 4: two 4-warp CTAs per SM; warps = 12: the warp-specialised instantiation,
        // 8 Gram warps + 4 update warps
        const int nw = warps == 4 ? 4 : MID_WARPS;
        const int na = warps == 12 ? MID_UPD_WARPS : 0;
        const int chunk = na > 0 ? MID_WS_CHUNK : mid_chunk(nw);
        long long share = (max_cols + cl - 1) / cl;
        share = (share + chunk - 1) / chunk * chunk;
        plan->tile = 6;
        plan->threads = (nw + na) * 32;
        plan->cluster = cl;
        plan->resident_cols = 0;
        plan->smem_bytes = (int32_t)(mid_carve(nw).total * 8);
        plan->ws_cols = share;
        long long clusters = (long long)sm_count * (nw == 4 ? 2 : 1) / cl;
        if (clusters > n_work) clusters = n_work;
        if (clusters < 1) clusters = 1;
        plan->ctas = (int32_t)(clusters * cl);
        plan->ws_bytes = 256 + (long long)plan->ctas * mid_slab_doubles(share) * 8;
        return DN_OK;
    }
    if (!for_init && cluster >= 0 && prm->p >= WIDE_MIN_P && wide_pp(prm->p) <= WIDE_MAX_PP) {
        // ---- wide path (49..208 samples): one 8 x 8 Gram tile per thread, streamed; optional cluster per gene
        int cl = cluster > 1 ? cluster : 1;
        if (cl != 1 && cl != 2 && cl != 4 && cl != 8 && cl != 16) return fail(DN_ERR_INVALID, "cluster must be 1, 2, 4, 8 or 16%s");
        const int pp = wide_pp(prm->p);
        long long share = (max_cols + cl - 1) / cl;
        share = (share + WIDE_CHUNK - 1) / WIDE_CHUNK * WIDE_CHUNK;
        plan->tile = 8;
        plan->threads = WIDE_THREADS;
        plan->cluster = cl;
        plan->chunk_cols = WIDE_CHUNK;
        plan->resident_cols = 0;
        plan->smem_bytes = (int32_t)(wide_carve(pp).total * 8);
        if (plan->smem_bytes > max_smem_optin) return fail(DN_ERR_UNSUPPORTED, "shared memory carve-up does not fit%s");
        plan->ws_cols = share;
        long long clusters = (long long)sm_count / cl;
        if (clusters > n_work) clusters = n_work;
        if (clusters < 1) clusters = 1;
        plan->ctas = (int32_t)(clusters * cl);
        plan->ws_bytes = 256 + (long long)plan->ctas * wide_slab_doubles(pp, share) * 8;
        return DN_OK;
    }
    const int P = for_init ? 0 : small_P(prm->p);
    if (P > 0 && cluster > 1) {
        // ---- small-p path, one thread-block cluster per gene: every CTA holds ceil(max_cols / cluster) columns
        if (cluster != 2 && cluster != 4 && cluster != 8 && cluster != 16) return fail(DN_ERR_INVALID, "cluster must be 2, 4, 8 or 16%s");
        const long long per_col = 8ll * (2 * small_cs(P) + 2);
        long long share = (max_cols + cluster - 1) / cluster;
        share = (share + 7) / 8 * 8;
        const bool res = want_resident != 0 && small_carve(P, SMALL_CLU_WARPS, (int)share, true).total * 8 <= max_smem_optin;
        plan->tile = 0;
        plan->threads = SMALL_CLU_WARPS * 32;
        plan->cluster = cluster;
        plan->resident_cols = res ? (int32_t)share : 0;
        plan->smem_bytes = (int32_t)(small_carve(P, SMALL_CLU_WARPS, res ? (int)share : 0, true).total * 8);
        plan->ws_cols = res ? 0 : share;
        long long clusters = sm_count / cluster;                     // one CTA per SM (256 threads, up to 216 registers)
        if (clusters > n_work) clusters = n_work;
        if (clusters < 1) clusters = 1;
        plan->ctas = (int32_t)(clusters * cluster);
        plan->ws_bytes = 256 + (long long)plan->ctas * small_slab_doubles(P, plan->ws_cols) * 8;
        return DN_OK;
    }
    if (P > 0) {
        // ---- small-p path: a bucket is either wholly shared-memory resident or wholly streamed
        const long long per_col = 8ll * (2 * small_cs(P) + 2);
        long long want = want_resident < 0 ? max_cols : (want_resident > max_cols ? max_cols : (long long)want_resident);
        want = (want + 1) / 2 * 2;
        int nw = warps > 0 ? warps : warps_for_tier(want);
        if (nw != 1 && nw != 2 && nw != 4 && nw != 8 && nw != 16) return fail(DN_ERR_INVALID, "warps must be 1, 2, 4, 8 or 16%s");
        long long res = 0;
        if (want_resident != 0) {
            if (small_carve(P, nw, (int)want).total * 8 <= max_smem_optin) res = want;
        }
        if (res == 0) nw = 8;
        plan->tile = 0;
        plan->threads = nw * 32;
        plan->chunk_cols = 0;
        plan->resident_cols = (int32_t)res;
        plan->smem_bytes = (int32_t)(small_carve(P, nw, (int)res).total * 8);
        plan->ws_cols = res > 0 ? 0 : (max_cols + 7) / 8 * 8;
        int per_sm = (int)((228ll * 1024) / (plan->smem_bytes + 1024));
        if (per_sm > 12 / nw) per_sm = 12 / nw;                     // 168 registers per thread (__launch_bounds__)
        if (res == 0 && per_sm > 2) per_sm = 2;                      // streamed: keep the slabs in flight L2-sized
        if (per_sm < 1) per_sm = 1;
        long long ctas = (long long)sm_count * per_sm;
        if (ctas > n_work) ctas = n_work;
        if (ctas < 1) ctas = 1;
        plan->ctas = (int32_t)ctas;
        plan->ws_bytes = 256 + ctas * small_slab_doubles(P, plan->ws_cols) * 8;
        return DN_OK;
    }
    const Derived d = derive(prm->p);
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
    if (per_sm < 1) per_sm = 1;
    long long ctas = (long long)sm_count * per_sm;
    if (ctas > n_work) ctas = n_work;
    if (ctas < 1) ctas = 1;
    plan->ctas = (int32_t)ctas;
    const long long g_d = (d.g_in_smem ? 2ll : 3ll) * d.pp * d.pp + d.gacc_doubles;
    const long long cols_d = for_init ? plan->ws_cols : (2ll * prm->p + 2) * plan->ws_cols;
    const long long stride = (g_d + cols_d + 31) / 32 * 32;
    plan->ws_bytes = 256 + ctas * stride * 8;
    return DN_OK;
}

int dn_init_ratio_svd(const double *cov, const int64_t *off, const int32_t *order, int32_t n_work, const dn_params *prm,
                      const dn_plan *plan, double *est_rowsum, double *cov_rowsum, double *row_max, int32_t *counters,
                      void *workspace, int64_t workspace_bytes, void *stream) {
    if (!est_rowsum || !cov_rowsum) return fail(DN_ERR_INVALID, "null output%s");
    if (plan && plan->tile == 0) return fail(DN_ERR_INVALID, "the init pass needs a plan made with for_init = 1%s");
    return run_kernel(MODE_INIT, cov, off, order, n_work, prm, plan, nullptr, nullptr, nullptr, row_max, nullptr, nullptr,
                      counters, nullptr, nullptr, nullptr, nullptr, est_rowsum, cov_rowsum, workspace, workspace_bytes,
                      stream);
}

int dn_baseline_selection(const double *cov, const int64_t *off, const int32_t *order, int32_t n_work,
                          const dn_params *prm, const dn_plan *plan, const double *scale, const int32_t *ds_start,
                          const double *row_max, double *rho, uint8_t *ran, int32_t *counters, double *kfac,
                          double *e_first, double *est, const int64_t *est_off, void *workspace,
                          int64_t workspace_bytes, void *stream) {
    if (!scale || !rho || !ran) return fail(DN_ERR_INVALID, "null pointer argument%s");
    if (prm && prm->downsample_rate > 1 && !ds_start) return fail(DN_ERR_INVALID, "ds_start required when downsampling%s");
    if (est && !kfac) return fail(DN_ERR_INVALID, "fused estimates need kfac%s");
    return run_kernel(MODE_BS, cov, off, order, n_work, prm, plan, scale, ds_start, row_max, nullptr, rho, ran, counters,
                      kfac, e_first, est, est_off, nullptr, nullptr, workspace, workspace_bytes, stream);
}

int dn_estimates(const double *cov, const int64_t *off, const int32_t *order, int32_t n_work, const dn_params *prm,
                 const double *scale, const int32_t *counters, const double *kfac, const double *e_first,
                 const int64_t *est_off, double *est, void *stream) {
    int rc = check_params(prm);
    if (rc) return rc;
    if (!cov || !off || !order || !scale || !counters || !kfac || !est) return fail(DN_ERR_INVALID, "null pointer argument%s");
    if (n_work <= 0) return DN_OK;
    int grid = n_work < 148 * 8 ? n_work : 148 * 8;
    estimates_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(cov, (const long long *)off, order, n_work, prm->p, scale,
                                                             counters, kfac, e_first, (const long long *)est_off, est);
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
