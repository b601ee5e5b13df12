"""
Drop-in for the reference's degnorm/nmf.py: class GeneNMFOA with the same constructor, run(), save_results()
and public attributes (nmf.py:10-711), backed by the B200 CUDA engine.  The CLI callers
(degnorm/__main__.py:264-286) work unchanged against this class.

Differences from the reference that are deliberate and documented in DESIGN.md:
  * n_jobs is accepted and ignored (the GPU owns the gene-level parallelism; joblib threads are gone).
  * the rank-one step is a p x p Gram eigen-solve instead of scipy svds; K >= 0 by convention.
  * extras that do not exist in the reference: `device=`, `return_estimates=` keyword arguments and the
    `counters` / `timings` attributes.  return_estimates='lazy' makes run() return a LazyEstimates sequence: the
    p x L_g estimates stay un-materialised on the device and are computed and copied only for the genes that are
    indexed (the report and the plots of the CLI touch a handful of genes, __main__.py:292-309).
  * the estimates returned by run() are views into a pinned host buffer owned by this object; a second run()
    on the same object re-uses it (copy what must outlive the next run).
"""
import logging
import os
import pickle as pkl
import time
import warnings

import numpy as np
import torch

from . import _lib
from .engine import Params, ShardEngine, draw_offsets
from .packing import pack_coverage, pinned_buffer, unpack_estimates


class LazyEstimates(object):
    """Sequence of the p x L_g estimated coverage matrices of the last outer iteration (what GeneNMFOA.run
    returns, nmf.py:601), materialised on demand: indexing gene i (or fetch([i, j, ...])) runs the estimate kernel
    for just those genes and copies just their blocks to the host.  The coverage stays on the device for as long as
    this object lives.  Works wherever the reference indexes or iterates the list (save_results included)."""

    def __init__(self, engine, lengths, batch_columns=4_000_000):
        self._engine = engine
        self._lengths = np.asarray(lengths, dtype=np.int64)
        self._batch_columns = int(batch_columns)

    def __len__(self):
        return len(self._lengths)

    def fetch(self, gene_ids):
        """List of estimates for the given gene positions (one launch, one device-to-host copy)."""
        ids = [int(i) + (len(self) if int(i) < 0 else 0) for i in gene_ids]
        if any(i < 0 or i >= len(self) for i in ids):
            raise IndexError("gene index out of range")
        uniq = sorted(set(ids))
        p = self._engine.p
        est, o = self._engine.estimates(uniq)
        host = est.cpu().numpy()
        where = {g: k for k, g in enumerate(uniq)}
        return [host[p * o[where[g]]: p * o[where[g] + 1]].reshape(p, -1) for g in ids]

    def __getitem__(self, i):
        if isinstance(i, slice):
            return self.fetch(range(*i.indices(len(self))))
        return self.fetch([i])[0]

    def __iter__(self):
        # batches of ~batch_columns columns: one launch and one copy per batch
        lo, n = 0, len(self)
        while lo < n:
            hi, cols = lo, 0
            while hi < n and (hi == lo or cols + self._lengths[hi] <= self._batch_columns):
                cols += self._lengths[hi]
                hi += 1
            for m in self.fetch(range(lo, hi)):
                yield m
            lo = hi


def _single_matrix_engine(x, device, **kw):
    """A one-gene ShardEngine around x (p x L): the device path behind the single-matrix methods."""
    if not torch.cuda.is_available():
        raise RuntimeError("degnorm_b200 needs a CUDA device (B200); there is no CPU fallback")
    x = np.asarray(x, dtype=np.float64)
    if x.ndim != 2 or min(x.shape) < 2:
        # scipy's svds(k=1) refuses such input (nmf.py:63); the reference swallows this ValueError in its drop loop
        raise ValueError("`k` must be an integer satisfying `0 < k < min(A.shape)`.")
    dev = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
    p, L = x.shape
    with torch.cuda.device(dev):
        flat = torch.from_numpy(np.ascontiguousarray(x).ravel()).to(dev)
        eng = ShardEngine(Params(**kw), p, dev)
        eng.load(flat, np.array([0, L], dtype=np.int64), torch.ones((1, p), dtype=torch.float64, device=dev))
    return eng, x


def _factorise(x, nmf_iter, device):
    """K (p x 1), E (1 x L) of nmf() (nmf.py:78-107) -- of rank_one_approx (nmf.py:55-64) when nmf_iter = 0 -- on
    the matrix as it is, through the fused kernel (DN_FLAG_PLAIN_NMF).  Sign convention: K >= 0, E >= 0."""
    eng, x = _single_matrix_engine(x, device, degnorm_iter=1, nmf_iter=nmf_iter, min_high_coverage=2,
                                   skip_baseline_selection=True)
    out = eng.fit_once(flags=_lib.DN_FLAG_PLAIN_NMF)
    K = out["kfac"].cpu().numpy().reshape(-1, 1)
    E = out["e_first"].cpu().numpy().reshape(1, -1)
    return K, E, eng, out


class GeneNMFOA(object):

    def __init__(self, degnorm_iter=5, downsample_rate=1, min_high_coverage=50, nmf_iter=100, bins=20, n_jobs=1,
                 skip_baseline_selection=False, random_state=123, device=None, return_estimates=True):
        prm = Params(degnorm_iter=degnorm_iter, downsample_rate=downsample_rate,
                     min_high_coverage=min_high_coverage, nmf_iter=nmf_iter, bins=bins, n_jobs=n_jobs,
                     skip_baseline_selection=skip_baseline_selection, random_state=random_state)
        self._prm = prm
        # same attribute names as the reference (nmf.py:30-49)
        self.degnorm_iter = prm.degnorm_iter
        self.nmf_iter = prm.nmf_iter
        self.n_jobs = prm.n_jobs
        self.bins = prm.bins
        self.min_high_coverage = prm.min_high_coverage
        self.min_bins = prm.min_bins
        self.downsample_rate = prm.downsample_rate
        self.mem_splits = None
        self.x = None
        self.x_weighted = None
        self.x_adj = None
        self.p = None
        self.n_genes = None
        self.genes = None
        self.norm_factors = None
        self.scale_factors = None
        self.rho = None
        self.fitted = False
        self.ran_baseline_selection = None
        self.skip_baseline_selection = skip_baseline_selection
        self.random_state = random_state
        self.device = device
        self.return_estimates = return_estimates
        self.counters = None
        self.timings = {}
        self._host_cache = {}      # pinned staging buffers, re-used by later run() calls on this object
        self._group = None         # torch.distributed group: this object's genes are one shard of a larger run

    # ---- single-matrix methods of the reference (public by convention), on the device path -----------------------
    @staticmethod
    def rank_one_approx(x, device=None):
        """nmf.py:55-64: (K, E) = (u * s, vh) of the top singular triplet of x; here from the p x p Gram
        eigenvector, with K >= 0 and E >= 0 (the reference's signs are arbitrary, only K.dot(E) and |K| are used)."""
        K, E, _, _ = _factorise(x, 0, device)
        return K, E

    def nmf(self, x, factors=False):
        """nmf.py:78-107: NMF-OA of x (no final clamp, as in the reference); (K, E) if factors else K.dot(E)."""
        K, E, _, _ = _factorise(x, self.nmf_iter, self.device)
        return (K, E) if factors else K.dot(E)

    def ratio_svd(self, x):
        """nmf.py:109-121: rank-one product clamped from below by x."""
        eng, x = _single_matrix_engine(x, self.device, degnorm_iter=1, nmf_iter=0, min_high_coverage=2,
                                       skip_baseline_selection=True)
        out = eng.fit_once(flags=_lib.DN_FLAG_PLAIN_NMF, want_estimates=True, clamp_estimates=True)
        return out["est"].cpu().numpy().reshape(x.shape)

    def run_ratio_svd_serial(self, x):
        """nmf.py:123-124"""
        return list(map(self.ratio_svd, x))

    def baseline_selection(self, F):
        """nmf.py:189-372 for ONE gene (F is taken as it is, i.e. already scaled): (rho, estimate, ran) with rho as
        the reference returns it (unclipped).  self.p must be set (run() sets it; the reference needs it too).
        With down-sampling the start offset is drawn from the global numpy stream like nmf.py:420-422."""
        F = np.asarray(F, dtype=np.float64)
        ds = None
        if self.downsample_rate > 1:
            if self.downsample_rate >= F.shape[1]:
                raise ValueError('Cannot downsample at a rate < 1 / length(gene)')
            ds = np.array([np.random.choice(self.downsample_rate)], dtype=np.int32)
        eng, F = _single_matrix_engine(F, self.device, degnorm_iter=1, nmf_iter=self.nmf_iter, bins=self.bins,
                                       downsample_rate=self.downsample_rate, min_high_coverage=self.min_high_coverage,
                                       skip_baseline_selection=self.skip_baseline_selection)
        out = eng.fit_once(ds_row=ds, flags=_lib.DN_FLAG_RAW_RHO, want_estimates=True)
        if int(out["counters"][0, _lib.CNT_EXIT]) < 0:
            raise _lib.DegnormCudaError("gene does not fit the launch plan")
        return (out["rho"].cpu().numpy()[0], out["est"].cpu().numpy().reshape(F.shape), bool(out["ran"][0].item()))

    # ---- small host helpers the reference exposes as (static) methods -------------------------------------------
    @staticmethod
    def get_high_coverage_idx(x):
        """nmf.py:66-76"""
        return np.where(x.max(axis=0) > 0.1 * x.max())[0]

    @staticmethod
    def shift_bins(bins, dropped_bin):
        """nmf.py:160-187: re-index bins after `dropped_bin` has been deleted so they stay consecutive."""
        if dropped_bin == len(bins) or len(bins) == 1:
            return bins
        delta = bins[0][0] if dropped_bin == 0 else bins[dropped_bin][0] - bins[dropped_bin - 1][-1] - 1
        for k in range(dropped_bin, len(bins)):
            bins[k] = [idx - delta for idx in bins[k]]
        return bins

    @staticmethod
    def _systematic_sample(n, take_every):
        """nmf.py:408-426 (draws from the global legacy numpy stream, like the reference)."""
        if take_every >= n:
            return int(np.random.choice(n))
        start = np.random.choice(take_every)
        return np.arange(start, n, step=take_every, dtype=int)

    def downsample_2d(self, x, by_row=True):
        """nmf.py:428-453"""
        Li = x.shape[0 if by_row else 1]
        if self.downsample_rate == 1:
            return x, np.arange(0, Li)
        if self.downsample_rate >= Li:
            raise ValueError('Cannot downsample at a rate < 1 / length(gene)')
        idx = self._systematic_sample(Li, take_every=self.downsample_rate)
        return (x[idx, :], idx) if by_row else (x[:, idx], idx)

    def check_input(self, cov_mats):
        """nmf.py:455-481 (same checks, same messages)."""
        if self.x.shape[0] != self.n_genes:
            raise ValueError('Number of genes in read count matrix not equal to number of coverage matrices!')
        if not all(map(lambda z: z.ndim == 2, cov_mats)):
            raise ValueError('Not all coverage matrices are 2-d arrays!')
        li_vec = np.array([m.shape[1] for m in cov_mats])
        if np.sum(li_vec / self.p < 1) > 0:
            logging.warning('At least one coverage matrix is taller than it is wide.'
                            'Ensure that coverage matrices are shaped (p x L_i).')
        if self.downsample_rate > 1:
            if not np.min(li_vec) >= self.downsample_rate:
                raise ValueError('downsample_rate is too large; take-every size > at least one gene.')
        if not all(m.shape[0] == self.p for m in cov_mats):
            raise ValueError('Not all coverage matrices have the same number of samples (rows)!')

    def _check_resident(self, cov):
        """check_input (nmf.py:455-481) for coverage that is already packed on the device."""
        if self.x.shape[0] != self.n_genes:
            raise ValueError('Number of genes in read count matrix not equal to number of coverage matrices!')
        if np.sum(cov.lengths / self.p < 1) > 0:
            logging.warning('At least one coverage matrix is taller than it is wide.'
                            'Ensure that coverage matrices are shaped (p x L_i).')
        if self.downsample_rate > 1:
            if not np.min(cov.lengths) >= self.downsample_rate:
                raise ValueError('downsample_rate is too large; take-every size > at least one gene.')

    # ---- the hot path ---------------------------------------------------------------------------------------------
    def run(self, cov_dat, reads_dat):
        """GeneNMFOA.run (nmf.py:483-601): returns the list of estimated coverage matrices (p x L_g, cov_dat key
        order) of the last outer iteration and sets rho, x_adj, scale_factors, norm_factors, x_weighted,
        ran_baseline_selection, genes, p, n_genes, fitted."""
        if not torch.cuda.is_available():
            raise RuntimeError("degnorm_b200 needs a CUDA device (B200); there is no CPU fallback")
        t0 = time.perf_counter()
        from .gene_filter import DeviceCoverage
        resident = cov_dat if isinstance(cov_dat, DeviceCoverage) else None
        self.n_genes = len(cov_dat)
        self.genes = list(cov_dat.keys())
        self.x = np.copy(reads_dat)
        if resident is None:
            cov_mats = list(cov_dat.values())
            self.p = cov_mats[0].shape[0]
            gene_lengths = [m.shape[1] for m in cov_mats]
            nbytes = np.sum([m.nbytes for m in cov_mats])
        else:
            # coverage already packed and on the device (gene_filter.DeviceCoverage): no packing, no upload
            self.p = resident.p
            gene_lengths = resident.lengths.tolist()
            nbytes = 8.0 * resident.flat.numel()
        self.ran_baseline_selection = np.zeros(shape=[self.n_genes, self.degnorm_iter]).astype(bool)
        if resident is None:
            self.check_input(cov_mats)
        else:
            self._check_resident(resident)
        mem_splits = int(np.ceil(nbytes / 5e7))
        self.mem_splits = max(mem_splits, self.n_jobs)

        dev = torch.device(self.device if self.device is not None else "cuda:%d" % torch.cuda.current_device())
        with torch.cuda.device(dev):
            if resident is None:
                flat, offsets = pack_coverage(cov_mats, self.p, cache=self._host_cache)     # pinned host staging
            else:
                flat, offsets = resident.flat, resident.offsets
            t1 = time.perf_counter()
            cov_dev = flat.to(dev, non_blocking=True)
            reads_dev = torch.from_numpy(np.ascontiguousarray(self.x, dtype=np.float64)).to(dev)
            eng = ShardEngine(self._prm, self.p, dev, group=self._group)
            eng.load(cov_dev, offsets, reads_dev)
            ds = draw_offsets(self.n_genes, self._prm)        # also seeds the global numpy stream (nmf.py:556)
            est_host = None
            lazy = isinstance(self.return_estimates, str) and self.return_estimates == 'lazy'
            eager = bool(self.return_estimates) and not lazy
            if eager and self.n_genes > 0 and self.degnorm_iter > 0:
                est_host = pinned_buffer(flat.numel(), "est", self._host_cache)
            out = eng.run(ds, want_estimates=eager, est_host=est_host, keep_for_estimates=lazy)
            torch.cuda.synchronize(dev)
            t2 = time.perf_counter()
            self.rho = out["rho"].cpu().numpy()
            self.x_adj = out["x_adj"].cpu().numpy()
            self.x_weighted = out["x_weighted"].cpu().numpy()
            self.norm_factors = out["norm_factors"].cpu().numpy()
            self.scale_factors = out["scale_factors"].cpu().numpy()
            self.ran_baseline_selection = out["ran"].cpu().numpy().T.astype(bool)
            self.counters = out["counters"].cpu().numpy()
            eng.check_exit_codes(self.counters)
            self._log_run(out["scale_hist"].cpu().numpy())
            estimates = None
            if lazy and self.n_genes > 0 and self.degnorm_iter > 0:
                estimates = LazyEstimates(eng, gene_lengths)
            if eager and out["est"] is not None:
                if out["est_in_work_order"]:
                    # already on the host (copied bucket by bucket behind the last iteration): views in gene order
                    arr, eo, p_ = est_host.numpy(), eng.est_off, self.p
                    estimates = [arr[p_ * int(eo[g]): p_ * (int(eo[g]) + L)].reshape(p_, -1)
                                 for g, L in enumerate(gene_lengths)]
                else:
                    estimates = unpack_estimates(out["est"], offsets, self.p, cache=self._host_cache)
            t3 = time.perf_counter()
        self._engine = eng
        self.fitted = True
        self.timings = dict(pack_s=t1 - t0, device_s=t2 - t1, unpack_s=t3 - t2, total_s=t3 - t0,
                            launches=eng.launches)
        return estimates

    def _log_run(self, scale_hist):
        """The reference's logging.info lines (nmf.py:537-538, 571-572, 592-593), same text and order.  The whole
        run is queued on the device without a host synchronisation, so they are emitted once it has finished."""
        logging.info('Initial sequencing depth scale factors -- \n\t{0}'
                     .format(', '.join([str(x) for x in scale_hist[0]])))
        for i in range(self.degnorm_iter):
            if not self.skip_baseline_selection:
                logging.info('DegNorm iteration {0} -- {1} genes sent through baseline selection'
                             .format(i + 1, np.sum(self.ran_baseline_selection[:, i])))
            logging.info('DegNorm iteration {0} -- sequencing depth scale factors: \n\t{1}'
                         .format(i + 1, ', '.join([str(x) for x in scale_hist[i + 1]])))

    # ---- output writer (host side, same files and columns as nmf.py:603-711) --------------------------------------
    def save_results(self, estimates, gene_manifest_df, output_dir='.', sample_ids=None):
        from pandas import DataFrame, concat
        if not self.fitted:
            raise ValueError('Model not yet fit. NMF-OA has not been run.')
        if not os.path.isdir(output_dir):
            raise IOError('Directory {0} not found.'.format(output_dir))
        if not all([col in gene_manifest_df.columns.tolist() for col in ['chr', 'gene']]):
            raise ValueError('gene_manifest_df must have columns `chr` and `gene`.')
        if sample_ids:
            if len(sample_ids) != self.p:
                raise ValueError('Number of supplied sample IDs does not match number'
                                 'of samples used to fit GeneNMFOA object.')
        else:
            sample_ids = ['sample_{0}'.format(i + 1) for i in range(self.p)]
        gene_intersect = np.intersect1d(gene_manifest_df.gene.unique(), self.genes)
        if len(gene_intersect) < len(self.genes):
            warnings.warn('Gene manifest data does not encompass set of genes sent through DegNorm.')
        if len(gene_intersect) == 0:
            raise ValueError('No genes used in DegNorm were found in gene manifest dataframe!')
        gene_df = gene_manifest_df[gene_manifest_df.gene.isin(gene_intersect)]
        manifest_chroms = gene_df.chr.unique().tolist()
        first_chrom = gene_df.drop_duplicates('gene').set_index('gene').chr
        position = {g: i for i, g in reversed(list(enumerate(self.genes)))}
        chrom_genes = {chrom: list() for chrom in manifest_chroms}
        for gene in gene_intersect:
            chrom_genes[first_chrom[gene]].append(gene)
        chrom_gene_dict = dict()
        for chrom in manifest_chroms:
            # (a LazyEstimates materialises one chromosome per launch and copy; a list is simply indexed)
            where = [position[gene] for gene in chrom_genes[chrom]]
            mats = estimates.fetch(where) if isinstance(estimates, LazyEstimates) else [estimates[k] for k in where]
            chrom_gene_dict[chrom] = dict(zip(chrom_genes[chrom], mats))
        chrom_gene_dfs = list()
        for chrom in manifest_chroms:
            chrom_dir = os.path.join(output_dir, str(chrom))
            if not os.path.isdir(chrom_dir):
                os.makedirs(chrom_dir)
            with open(os.path.join(chrom_dir, 'estimated_coverage_matrices_{0}.pkl'.format(chrom)), 'wb') as f:
                pkl.dump(chrom_gene_dict[chrom], f)
            chrom_gene_dfs.append(DataFrame({'chr': chrom, 'gene': list(chrom_gene_dict[chrom].keys())}))
        chrom_gene_df = concat(chrom_gene_dfs)
        chrom_gene_df.set_index('gene', inplace=True)
        chrom_gene_df = chrom_gene_df.loc[self.genes]
        chrom_gene_df.reset_index(inplace=True)
        chrom_gene_df = chrom_gene_df[['chr', 'gene']]
        iter_names = ['iter_{0}'.format(i) for i in range(self.degnorm_iter)]
        for mat, cols, fname in ((self.rho, sample_ids, 'degradation_index_scores.csv'),
                                 (self.x_adj, sample_ids, 'adjusted_read_counts.csv'),
                                 (self.ran_baseline_selection, iter_names, 'ran_baseline_selection.csv')):
            df = concat([chrom_gene_df, DataFrame(mat, columns=cols)], axis=1)
            df = df[['chr', 'gene'] + cols]
            df.to_csv(os.path.join(output_dir, fname), index=False)
