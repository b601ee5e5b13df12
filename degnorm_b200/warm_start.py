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

With `device=` the coverage goes to the GPU chromosome by chromosome WHILE the next pickle is being read: every
chromosome is packed into one of two pinned staging buffers and leaves on a copy stream as soon as it is packed,
so the host-to-device transfer hides behind the (much slower) unpickling.  gene_cov_dict is then a
gene_filter.DeviceCoverage -- the resident handle that the gene filter and GeneNMFOA.run accept in place of the
dictionary (no second packing pass, no upload at run time).
"""
import gc
import os
import pickle as pkl
import shutil
from collections import OrderedDict

import numpy as np

from .packing import pinned_buffer


def load_from_previous(degnorm_dir, new_dir, pack=True, device=None):
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

    if device is not None:
        gene_cov_dict = _stream_to_device(degnorm_dir, new_dir, genes_df.chr.unique().tolist(), keep, device)
        genes = gene_cov_dict.keys()
        genes_df = genes_df.set_index('gene').loc[genes].reset_index(drop=False)
        read_count_df = read_count_df.set_index('gene').loc[genes].reset_index(drop=False)
        return dict(gene_cov_dict=gene_cov_dict, read_count_df=read_count_df, genes_df=genes_df, sample_ids=sample_ids)

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


def _stream_to_device(degnorm_dir, new_dir, chroms, keep, device):
    """Chromosome pickles -> device, double-buffered: while chromosome k is read and packed on the host, chromosome
    k - 1 travels on a copy stream.  Returns a gene_filter.DeviceCoverage over all kept genes, in the reference's order
    (chromosomes as in the annotation, genes as in each pickle)."""
    import torch
    from .gene_filter import DeviceCoverage
    dev = torch.device(device)
    cache = {}
    copy_stream = torch.cuda.Stream(device=dev)
    done = [None, None]                    # per staging buffer: event of the last copy that read it
    parts, genes, lengths = [], [], []
    p = None
    for k, chrom in enumerate(chroms):
        os.makedirs(os.path.join(new_dir, chrom))
        cov_file = os.path.join(degnorm_dir, chrom, 'coverage_matrices_{0}.pkl'.format(chrom))
        shutil.copy(cov_file, os.path.join(new_dir, chrom, 'coverage_matrices_{0}.pkl'.format(chrom)))
        with open(cov_file, 'rb') as f:
            cov_dat = pkl.load(f)          # (the previous chromosome's copy runs meanwhile)
        mats = [(g, m) for g, m in cov_dat.items() if g in keep]
        del cov_dat
        if not mats:
            continue
        if p is None:
            p = mats[0][1].shape[0]
        if not all(m.ndim == 2 and m.shape[0] == p for _, m in mats):
            raise ValueError('Not all coverage matrices have the same number of samples (rows)!')
        total = int(sum(m.size for _, m in mats))
        slot = k & 1
        if done[slot] is not None:
            done[slot].synchronize()       # the staging buffer is free again
        stage = pinned_buffer(total, "stage%d" % slot, cache)
        buf = stage.numpy()
        pos = 0
        for g, m in mats:
            buf[pos:pos + m.size].reshape(m.shape)[...] = m        # any dtype / memory order -> C-contiguous float64
            pos += m.size
            genes.append(g)
            lengths.append(m.shape[1])
        del mats
        part = torch.empty(total, dtype=torch.float64, device=dev)
        with torch.cuda.stream(copy_stream):
            part.copy_(stage, non_blocking=True)
            done[slot] = torch.cuda.Event()
            done[slot].record(copy_stream)
        parts.append(part)
    if not parts:
        raise ValueError('No genes available to run through DegNorm!')
    copy_stream.synchronize()
    flat = parts[0] if len(parts) == 1 else torch.cat(parts)
    del parts
    offsets = np.zeros(len(lengths) + 1, dtype=np.int64)
    np.cumsum(np.asarray(lengths, dtype=np.int64), out=offsets[1:])
    return DeviceCoverage(_parts=(genes, p, flat, offsets))
