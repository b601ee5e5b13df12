/*
 * degnorm_b200 -- C ABI of the B200-native NMF-OA engine (libdegnorm_b200.so).
 *
 * The reference (NUStatBioinfo/DegNorm v0.1.4) is pure Python and has no FFI layer; its boundary for this
 * path is the Python class GeneNMFOA (degnorm/nmf.py:10-601) and the free function run_gene_nmfoa_mpi
 * (degnorm/nmf_mpi.py:555-863).  This header is the thin C layer the drop-in Python class
 * (degnorm_b200/nmf.py) calls through ctypes; each entry point names the reference code it replaces.
 *
 * Conventions
 *   - plain C, no exceptions: every function returns 0 on success or a negative dn_status; a message is kept
 *     per thread and returned by dn_last_error().
 *   - all pointers are DEVICE pointers owned by the caller (e.g. torch tensors' data_ptr()) unless the name
 *     starts with h_; the library allocates nothing that outlives a call and never synchronises the
 *     stream (dn_device_info is the exception: it queries the device).
 *   - stream is a cudaStream_t passed as void*; NULL is the legacy default stream.
 *   - coverage layout ("ragged CSR buffer"): gene g occupies cov[p*off[g] .. p*off[g+1]) as a C-contiguous
 *     p x L_g block (sample-major rows), L_g = off[g+1]-off[g]; off has n_genes+1 int64 entries.
 *   - n x p matrices (rho, x_weighted, x_adj ...) are C-contiguous, row = gene id.
 *   - `order` is the work list: gene ids in the order CTAs pull them (longest first); outputs are always
 *     written at the gene id, so the caller's gene order is preserved whatever the work order.
 */
#ifndef DEGNORM_B200_H
#define DEGNORM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DN_ABI_VERSION 6
#define DN_MAX_BINS 64      /* baseline-selection bins held in the fused kernel (reference default: 20) */
#define DN_MAX_SAMPLES 256  /* p supported by the fused kernels (one thread per sample in the n x p steps) */
#define DN_NCOUNTERS 8      /* int32 counters per gene, see dn_counter */

typedef enum dn_status {
    DN_OK = 0,
    DN_ERR_INVALID = -1,    /* bad argument (maps to ValueError in the Python class) */
    DN_ERR_CUDA = -2,       /* CUDA runtime error; text in dn_last_error() */
    DN_ERR_UNSUPPORTED = -3,/* valid in the reference but outside this build's limits (p, bins) */
    DN_ERR_WORKSPACE = -4   /* workspace too small */
} dn_status;

/* exit / branch taken by baseline_selection for a gene (nmf.py line numbers) */
typedef enum dn_exit {
    DN_EXIT_NONE = 0,
    DN_EXIT_FEW_HICOV = 1,      /* nmf.py:232-233  fewer than min_high_coverage columns      */
    DN_EXIT_EMPTY_SAMPLE = 2,   /* nmf.py:241-242  a sample without coverage after filtering  */
    DN_EXIT_MEDIAN = 3,         /* nmf.py:257-258  median(1-rho) > 1                           */
    DN_EXIT_NO_SELECTION = 4,   /* nmf.py:265 false: first NMF-OA fit kept                    */
    DN_EXIT_REFINED = 5,        /* nmf.py:327-337  baseline found, envelope refined            */
    DN_EXIT_FALLBACK_HIGH = 6,  /* nmf.py:342-346  refined DI > 0.9, first fit restored        */
    DN_EXIT_FALLBACK = 7        /* nmf.py:349-353  no baseline, first fit restored             */
} dn_exit;
/* A NEGATIVE value in DN_CNT_EXIT (-1) is not a branch of the algorithm: the gene did not fit the plan of the launch it
 * was put in (its share of candidate columns exceeds plan.ws_cols -- a planner or ABI misuse).  Its outputs then hold
 * the default result (DI 0, flag 0); callers must treat it as an error (the Python class raises DegnormCudaError). */

/* per-gene int32 counters written by dn_baseline_selection (row = gene id, DN_NCOUNTERS columns) */
typedef enum dn_counter {
    DN_CNT_EXIT = 0,        /* dn_exit                                                     */
    DN_CNT_N_HICOV = 1,     /* columns factorised by the first nmf() call (after filters)  */
    DN_CNT_NMF_CALLS = 2,   /* nmf() calls made (1 + bins dropped, nmf.py:245,306)         */
    DN_CNT_SUM_COLS = 3,    /* sum of widths of all factorised matrices (algorithmic bytes)*/
    DN_CNT_EIG_STEPS = 4,   /* power-iteration steps spent in the p x p eigen-solves       */
    DN_CNT_DROPS_LO = 5,    /* bit b set: bin b was dropped (bins 0..31)                   */
    DN_CNT_DROPS_HI = 6,    /* bins 32..63                                                 */
    DN_CNT_RESIDENT = 7     /* bit 0: gene matrix held in shared memory (else streamed);
                               bits 1..: eigen-solves that needed the small-gap fallback    */
} dn_counter;

/* Normalised algorithm parameters: GeneNMFOA.__init__ (nmf.py:12-53) after abs/int/ceil. */
typedef struct dn_params {
    int32_t p;                  /* samples (rows of every coverage matrix)                       */
    int32_t nmf_iter;           /* --nmf-iter, multiplier updates per nmf() call (nmf.py:93)      */
    int32_t bins;               /* baseline-selection bins (nmf.py:269)                           */
    int32_t min_bins;           /* ceil(0.2*bins) (nmf.py:35)                                     */
    int32_t min_high_coverage;  /* nmf.py:34,52-53                                                */
    int32_t downsample_rate;    /* -d, systematic "take every" rate (nmf.py:36)                   */
    int32_t min_gene_len;       /* max(2, ceil(200*(1/rate))) evaluated as the reference does (nmf.py:261) */
    int32_t skip_baseline_selection; /* -s (nmf.py:265)                                          */
    int32_t flags;              /* 0 for GeneNMFOA.run; DN_FLAG_* for the single-matrix helper methods            */
} dn_params;

/* dn_params.flags: what the single-matrix methods of the class need from dn_baseline_selection.
 * DN_FLAG_PLAIN_NMF: the matrix is factorised as it is -- GeneNMFOA.nmf / rank_one_approx / ratio_svd
 *   (nmf.py:55-121): no high-coverage filter, no early exits (nmf.py:232-258), one nmf() call (set skip too);
 *   kfac receives K = u*s (>= 0 by convention), e_first receives E = vh.
 * DN_FLAG_RAW_RHO: rho is written as baseline_selection returns it (nmf.py:372), without the [0, 0.9] clip that
 *   par_apply_baseline_selection applies afterwards (nmf.py:398-399). */
#define DN_FLAG_PLAIN_NMF 1
#define DN_FLAG_RAW_RHO 2

/* Launch plan for one bucket of genes (all genes of one launch share a shared-memory carve-up). */
typedef struct dn_plan {
    int32_t tile;           /* 0: small-p kernel (p <= 12); 6: mid-p kernel (13..48); 4 or 8: tiled kernel's Gram tile edge */
    int32_t threads;        /* CTA size                                                         */
    int32_t ctas;           /* persistent CTAs to launch                                        */
    int32_t resident_cols;  /* columns of x and lambda held in shared memory (0: none)          */
    int32_t chunk_cols;     /* columns per Gram tile pass                                       */
    int32_t smem_bytes;     /* dynamic shared memory per CTA                                    */
    int32_t cluster;        /* CTAs per gene: 1, or a thread-block cluster of 2/4/8/16 (small-p path)  */
    int32_t reserved;
    int64_t ws_cols;        /* columns of per-CTA global workspace (0: none needed)             */
    int64_t ws_bytes;       /* total workspace bytes this launch needs (all CTAs + queue)       */
} dn_plan;

int dn_abi_version(void);
const char *dn_last_error(void);

/* Device facts the host planner needs: SM count, max opt-in shared memory per CTA, compute capability
 * (major*10+minor).  Fails with DN_ERR_CUDA when there is no usable CUDA device: there is no CPU path. */
int dn_device_info(int32_t *sm_count, int32_t *max_smem_optin, int32_t *cc);

/* Fill `plan` for a bucket whose largest gene has max_cols candidate columns (L for the init pass and
 * for downsample_rate 1, ceil(L/rate) otherwise).  want_resident: shared-memory column capacity wanted
 * (0 forces the streamed path, -1 = as many as fit).  for_init=1 sizes the workspace for dn_init_ratio_svd.
 * warps: warps per CTA on the small-p path (1, 2, 4, 8, 16; 0 = chosen from the tier); on the mid-p path 4 selects
 * the instantiation with two 4-warp CTAs per SM and 12 the warp-specialised one (8 Gram warps + 4 update warps);
 * anything else: 8 warps, one CTA per SM (the default); ignored elsewhere.
 * cluster: CTAs that share one gene on the small-p path (0/1: none; 2, 4, 8, 16: a thread-block cluster whose
 * CTAs each hold ceil(max_cols/cluster) columns and exchange the partial Gram through distributed shared memory).
 * For 13..48 samples the streamed mid-p kernel is planned (cluster 1..16); cluster = -1 asks for the generic tiled
 * kernel instead (the path of p > 48, kept selectable for cross-checks).
 * On the small-p path (p <= 12, for_init = 0) a bucket is wholly resident (max_cols fit in shared memory) or
 * wholly streamed; on the tiled path residency is decided per gene.
 * Pure host arithmetic (no device call): sm_count / max_smem come from dn_device_info. */
int dn_make_plan(const dn_params *prm, int64_t max_cols, int32_t n_work, int32_t want_resident, int32_t for_init,
                 int32_t warps, int32_t cluster, int32_t sm_count, int32_t max_smem_optin, dn_plan *plan);

/* Replaces run_ratio_svd_serial / ratio_svd over all genes + the two row sums run() takes of it
 * (nmf.py:109-140, 521-527): est_rowsum[g,i] = sum_j max(R1(F_g)_ij, F_g[i,j]), cov_rowsum[g,i] = sum_j F_g[i,j]
 * on the raw, unscaled, unfiltered matrices.  row_max (n x p or NULL): row_max[g,i] = max_j F_g[i,j], which
 * dn_baseline_selection can take instead of re-scanning every gene for the 0.1*max(F) threshold (nmf.py:76):
 * max_j(F_ij / s_i) = (max_j F_ij) / s_i exactly, because division by a positive scale is monotone. */
int dn_init_ratio_svd(const double *cov, const int64_t *off, const int32_t *order, int32_t n_work,
                      const dn_params *prm, const dn_plan *plan,
                      double *est_rowsum, double *cov_rowsum, double *row_max, int32_t *counters,
                      void *workspace, int64_t workspace_bytes, void *stream);

/* Replaces adjust_coverage_curves + par_apply_baseline_selection + baseline_selection + nmf +
 * rank_one_approx + get_high_coverage_idx + shift_bins + downsample_2d for one outer iteration
 * (nmf.py:55-107, 142-146, 160-406).  scale = current scale factors (p, device; coverage is divided by
 * them on load).  ds_start = systematic-sample offset per gene id (NULL when downsample_rate == 1).
 * row_max = row maxima from dn_init_ratio_svd (NULL: every gene is scanned for its maximum).
 * Outputs: rho (n x p, already clipped to [0, 0.9] as nmf.py:398-399), ran (n, uint8),
 * counters (n x DN_NCOUNTERS), kfac (n x p: |K| floored as nmf.py:361-362, input of dn_estimates),
 * e_first (sum_g L_g doubles or NULL: E of the first fit for genes where no column was filtered).
 * est (or NULL): when given (last outer iteration), every gene's full-length estimate is written by the CTA(s) that
 * just finished the gene, at column offset est_off[g] (NULL: off[g]) -- the fused form of dn_estimates, which lets the
 * caller start the device-to-host copy of a bucket's estimates as soon as the bucket's launch has finished. */
int dn_baseline_selection(const double *cov, const int64_t *off, const int32_t *order, int32_t n_work,
                          const dn_params *prm, const dn_plan *plan,
                          const double *scale, const int32_t *ds_start, const double *row_max,
                          double *rho, uint8_t *ran, int32_t *counters, double *kfac, double *e_first,
                          double *est, const int64_t *est_off,
                          void *workspace, int64_t workspace_bytes, void *stream);

/* Replaces the estimate assembly at nmf.py:217, 247, 333-365 for the last outer iteration: writes the
 * p x L_g estimate of every listed gene into est.  est_off (n_genes int64, or NULL): column offset of gene g's
 * block in est (block = est[p*est_off[g] ..]); NULL means the same ragged layout as cov.  A caller that wants the
 * device-to-host copy of the estimates to overlap the last iteration lays est out in work order and calls this
 * per bucket, on the bucket's stream, right after dn_baseline_selection. */
int dn_estimates(const double *cov, const int64_t *off, const int32_t *order, int32_t n_work,
                 const dn_params *prm, const double *scale, const int32_t *counters,
                 const double *kfac, const double *e_first, const int64_t *est_off, double *est, void *stream);

/* Outer update, part 1 (nmf.py:575, 148-158): per-sample sums over this rank's genes,
 * sums[0:p] = sum_g x_w, sums[p:2p] = sum over genes with a non-zero DI row of x_w/(1-rho),
 * sums[2p:3p] = sum over all-zero-DI genes of x_w.  With several GPUs the caller all-reduces `sums`. */
/* `sums` always has 3p+1 doubles. */
int dn_outer_sums(const double *x_weighted, const double *rho, int32_t n_genes, int32_t p,
                  double *sums, void *workspace, int64_t workspace_bytes, void *stream);

/* Outer update, part 2 (nmf.py:148-158, 575-590): from the (all-reduced) sums, corrects rho for genes that
 * had an all-zero DI row, writes x_adj, norm_factors, and updates x_weighted and scale_factors in place. */
int dn_outer_apply(const double *sums, int32_t n_genes, int32_t p, double *x_weighted, double *rho,
                   double *x_adj, double *norm_factors, double *scale_factors, void *stream);

/* Init scale (nmf.py:524-535): from est/cov row sums computes rho0, the low-DI flag per gene and
 * sums[0:p] = sum over low-DI genes of reads, sums[p:2p] = sum over all genes, sums[3p] = #low-DI genes
 * (as double).  With several GPUs the caller all-reduces `sums`; dn_init_apply then sets
 * norm_factors = scale_factors = count_sums/median and x_weighted = reads/norm_factors. */
int dn_init_sums(const double *est_rowsum, const double *cov_rowsum, const double *reads, int32_t n_genes,
                 int32_t p, double *rho0, double *sums, void *workspace, int64_t workspace_bytes, void *stream);
int dn_init_apply(const double *sums, const double *reads, int32_t n_genes, int32_t p, double *x_weighted,
                  double *norm_factors, double *scale_factors, void *stream);

/* workspace bytes dn_outer_sums / dn_init_sums need */
int64_t dn_sums_workspace_bytes(int32_t n_genes, int32_t p);

/* ---- measurement probes (not on the product path; csrc/probes.cu) ------------------------------------------------
 * SURVEY.md section 8(d) asks for an FP64 peak measured on the box: dn_probe_fp64 launches `ctas` CTAs of 256 threads,
 * each thread 64 * iters DFMAs in 16 independent chains, writes every CTA's SM clock cycles to cycles[ctas] and returns
 * the number of DFMAs issued.  dn_probe_lds does the same for 128-bit shared-memory loads with a lane -> address
 * pattern (0: 32 distinct conflict-free, 1: one address, 2: one address per quarter-warp, 3: every quarter-warp the same
 * 128 bytes, ...) and returns the warp-level loads per CTA.  seed: 2 doubles, sink: 1 double (device). */
int64_t dn_probe_fp64(int32_t ctas, int32_t iters, const double *seed, double *sink, int64_t *cycles, void *stream);
int64_t dn_probe_lds(int32_t ctas, int32_t pattern, int32_t iters, double *sink, int64_t *cycles, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DEGNORM_B200_H */
