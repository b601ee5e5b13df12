"""
Drop-in for the reference's degnorm/warm_start.py:load_from_previous (warm_start.py:10-106), the producer of the hot
path's input when DegNorm restarts from a previous run's output directory (--warm-start-dir): same arguments, same
files copied into the new output directory, same return dictionary (gene_cov_dict, read_count_df, genes_df,
sample_ids), same errors.

What is different underneath: the per-chromosome pickles are unpacked straight into ONE contiguous
float64 staging buffer (pinned when a CUDA device is present), gene after gene in the order the reference produces,
and the matrices handed back are C-contiguous p x L_g views into it.  GeneNMFOA.run recognises that layout
(packing._contiguous_view) and uploads the buffer as it is: no per-gene copy at run time, one host copy of the
coverage instead of two.
"""
import gc
import os
import pickle as pkl
import shutil
from collections import OrderedDict

import numpy as np

from .packing import pinned_buffer


def load_from_previous(degnorm_dir, new_dir, pack=True):
    from pandas import read_csv
    if not os.path.isdir(new_dir):
        raise IOError('new DegNorm output directory {0} not found.'.format(new_dir))
    exon_file = os.path.join(degnorm_dir, 'gene_exon_metadata.csv')
    read_count_file = os.path.join(degnorm_dir, 'read_counts.csv')
    # (a missing file raises FileNotFoundError, as in the reference)
    shutil.copy(exon_file, os.path.join(new_dir, 'gene_exon_metadata.csv'))
    shutil.copy(read_count_file, os.path.join(new_dir, 'read_counts.csv'))
    exon_df = read_csv(exon_file, low_memory=False)
    read_count_df = read_csv(read_count_file, low_memory=False)

    genes_df = exon_df[['chr', 'gene', 'gene_start', 'gene_end']].drop_duplicates().reset_index(drop=True)
    # genes present in both the annotation and the read counts
    keep = set(np.intersect1d(genes_df.gene, read_count_df.gene).tolist())
    genes_df = genes_df[genes_df.gene.isin(keep)]
    read_count_df = read_count_df[read_count_df.gene.isin(keep)]
    sample_ids = read_count_df.columns.tolist()[2:]

    loaded = OrderedDict()
    for chrom in genes_df.chr.unique().tolist():
        os.makedirs(os.path.join(new_dir, chrom))
        cov_file = os.path.join(degnorm_dir, chrom, 'coverage_matrices_{0}.pkl'.format(chrom))
        shutil.copy(cov_file, os.path.join(new_dir, chrom, 'coverage_matrices_{0}.pkl'.format(chrom)))
        with open(cov_file, 'rb') as f:
            cov_dat = pkl.load(f)
        for gene in cov_dat:
            if gene in keep:
                loaded[gene] = cov_dat[gene]
        del cov_dat
    gc.collect()

    gene_cov_dict = loaded
    if pack and len(loaded):
        p = next(iter(loaded.values())).shape[0]
        if all(m.ndim == 2 and m.shape[0] == p for m in loaded.values()):
            total = int(sum(m.shape[1] for m in loaded.values())) * p
            buf = pinned_buffer(total, "warm_start", {}).numpy()
            gene_cov_dict = OrderedDict()
            pos = 0
            for gene in list(loaded.keys()):
                m = loaded.pop(gene)                       # (released as soon as it is in the staging buffer)
                view = buf[pos:pos + m.size].reshape(m.shape)
                view[...] = m                               # any dtype / memory order -> C-contiguous float64
                gene_cov_dict[gene] = view
                pos += m.size
            gc.collect()

    genes = list(gene_cov_dict.keys())
    genes_df = genes_df.set_index('gene').loc[genes].reset_index(drop=False)
    read_count_df = read_count_df.set_index('gene').loc[genes].reset_index(drop=False)
    return dict(gene_cov_dict=gene_cov_dict, read_count_df=read_count_df, genes_df=genes_df, sample_ids=sample_ids)
