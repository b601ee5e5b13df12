"""
The gene filter the reference CLI applies between loading the coverage and running NMF-OA
(degnorm/__main__.py:216-244, degnorm/__main_mpi__.py:374-376): a gene is dropped when its largest coverage value is
below --minimax-coverage or when it is not longer than the down-sampling rate; read counts and the gene table lose
the same rows; an empty result or a count mismatch raises the reference's ValueError.

The reference walks the dictionary and calls cov_mat.max() per gene.  Here the per-gene maxima come from ONE
segmented reduction over the packed ragged buffer (include/degnorm_b200.h layout; zero-copy when the loader produced
the matrices as views of one buffer, as coverage_merge / warm_start do) -- on the device when one is given, so that
a caller which keeps the coverage resident filters it where it lives.  SURVEY.md section 8 row f-3 (the filter half).
"""
from collections import OrderedDict

import numpy as np
import torch

from .packing import pack_coverage


def gene_max_coverage(flat, offsets, p):
    """Largest coverage value per gene of a packed buffer (torch tensor, host or device; gene g occupies
    [p*off[g], p*off[g+1]))."""
    lengths = torch.as_tensor(p * np.diff(np.asarray(offsets, dtype=np.int64)), device=flat.device)
    if lengths.numel() == 0:
        return torch.zeros(0, dtype=flat.dtype, device=flat.device)
    return torch.segment_reduce(flat, "max", lengths=lengths, unsafe=True)


def keep_mask(flat, offsets, p, minimax_coverage=0, downsample_rate=1):
    """Boolean numpy mask of the genes that stay (__main__.py:228-230: dropped if max < minimax_coverage or
    L <= downsample_rate)."""
    L = np.diff(np.asarray(offsets, dtype=np.int64))
    mx = gene_max_coverage(flat, offsets, p).cpu().numpy()
    return ~((mx < minimax_coverage) | (L <= downsample_rate))


def filter_genes(gene_cov_dict, genes_df, read_count_df, minimax_coverage=0, downsample_rate=1, device=None):
    """In place of the loop at __main__.py:221-244.  gene_cov_dict loses the dropped genes (in place, like the
    reference); returns (gene_cov_dict, genes_df, read_count_df) with the same rows dropped and the indices reset."""
    genes = genes_df.gene.tolist()
    mats = [gene_cov_dict[g] for g in genes]
    keep = np.ones(len(genes), dtype=bool)
    if mats:
        p = mats[0].shape[0]
        flat, offsets = pack_coverage(mats, p, pin=False)
        if device is not None:
            flat = flat.to(device)
        keep = keep_mask(flat, offsets, p, minimax_coverage, downsample_rate)
    delete_idx = np.flatnonzero(~keep).tolist()
    for i in delete_idx:
        del gene_cov_dict[genes[i]]
    if delete_idx:
        read_count_df = read_count_df.drop(delete_idx, axis=0).reset_index(drop=True)
        genes_df = genes_df.drop(delete_idx, axis=0).reset_index(drop=True)
    if (read_count_df.shape[0] == 0) or genes_df.empty or (len(gene_cov_dict) == 0):
        raise ValueError('No genes available to run through DegNorm!\n'
                         'Check that your requested genes are in genome annotation file.')
    if len(gene_cov_dict.keys()) != read_count_df.shape[0]:
        raise ValueError('Number of coverage matrices not equal to number of genes in read count DataFrame!')
    return gene_cov_dict, genes_df, read_count_df


class DeviceCoverage(object):
    """The coverage of a run, packed and resident where the kernels read it: one upload shared by the gene filter,
    GeneNMFOA.run (which accepts this object in place of the {gene: p x L_g} dictionary and then skips packing and
    the host-to-device copy) and the lazy estimates.

        cov = DeviceCoverage(gene_cov_dict, device="cuda:0")          # pack (zero-copy if contiguous) + one upload
        cov, genes_df, read_count_df = cov.filter(genes_df, read_count_df, minimax_coverage, downsample_rate)
        estimates = GeneNMFOA(...).run(cov, read_count_df[sample_ids].values)
    """

    def __init__(self, cov_dat=None, device=None, cache=None, _parts=None):
        if _parts is not None:
            self.genes, self.p, self.flat, self.offsets = _parts
        else:
            self.genes = list(cov_dat.keys())
            mats = list(cov_dat.values())
            if not mats:
                raise ValueError('No genes available to run through DegNorm!')
            if not all(m.ndim == 2 for m in mats):
                raise ValueError('Not all coverage matrices are 2-d arrays!')
            self.p = mats[0].shape[0]
            if not all(m.shape[0] == self.p for m in mats):
                raise ValueError('Not all coverage matrices have the same number of samples (rows)!')
            flat, self.offsets = pack_coverage(mats, self.p, pin=torch.cuda.is_available(), cache=cache)
            self.flat = flat.to(device) if device is not None else flat
        self.lengths = np.diff(self.offsets)

    def __len__(self):
        return len(self.genes)

    def keys(self):
        return list(self.genes)

    def keep_mask(self, minimax_coverage=0, downsample_rate=1):
        return keep_mask(self.flat, self.offsets, self.p, minimax_coverage, downsample_rate)

    def select(self, keep):
        """The genes with keep[g] true, re-packed on the device (one copy per run of consecutive kept genes)."""
        keep = np.asarray(keep, dtype=bool)
        if keep.all():
            return self
        idx = np.flatnonzero(keep)
        offsets = np.zeros(len(idx) + 1, dtype=np.int64)
        np.cumsum(self.lengths[idx], out=offsets[1:])
        flat = torch.empty(self.p * int(offsets[-1]), dtype=self.flat.dtype, device=self.flat.device)
        # runs of consecutive kept genes are contiguous in both buffers
        brk = np.flatnonzero(np.diff(idx) != 1) + 1
        for a, b in zip(np.concatenate(([0], brk)), np.concatenate((brk, [len(idx)]))):
            src_lo, src_hi = self.p * int(self.offsets[idx[a]]), self.p * int(self.offsets[idx[b - 1] + 1])
            dst_lo = self.p * int(offsets[a])
            flat[dst_lo:dst_lo + (src_hi - src_lo)] = self.flat[src_lo:src_hi]
        return DeviceCoverage(_parts=([self.genes[g] for g in idx], self.p, flat, offsets))

    def filter(self, genes_df, read_count_df, minimax_coverage=0, downsample_rate=1):
        """filter_genes for resident coverage (__main__.py:216-244): -> (DeviceCoverage, genes_df, read_count_df).
        The rows of genes_df must be this object's genes, in order."""
        if genes_df.gene.tolist() != self.genes:
            raise ValueError('genes_df does not list the genes of the coverage, in order')
        keep = self.keep_mask(minimax_coverage, downsample_rate)
        delete_idx = np.flatnonzero(~keep).tolist()
        if delete_idx:
            read_count_df = read_count_df.drop(delete_idx, axis=0).reset_index(drop=True)
            genes_df = genes_df.drop(delete_idx, axis=0).reset_index(drop=True)
        if (read_count_df.shape[0] == 0) or genes_df.empty or not keep.any():
            raise ValueError('No genes available to run through DegNorm!\n'
                             'Check that your requested genes are in genome annotation file.')
        cov = self.select(keep)
        if len(cov) != read_count_df.shape[0]:
            raise ValueError('Number of coverage matrices not equal to number of genes in read count DataFrame!')
        return cov, genes_df, read_count_df

    def to_dict(self):
        """Host copy as the reference's {gene: p x L_g} dictionary (views of one buffer)."""
        host = self.flat.cpu().numpy()
        return OrderedDict((g, host[self.p * int(self.offsets[k]): self.p * int(self.offsets[k + 1])].reshape(self.p, -1))
                           for k, g in enumerate(self.genes))
