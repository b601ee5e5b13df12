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
