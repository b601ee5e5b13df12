"""
Drop-in for the reference's coverage merge (degnorm/reads_coverage_merge.py: merge_chrom_coverage :167-372,
merge_overlap_gene_coverage :93-164, merge_coverage :375-461), the step that turns the per-sample chromosome
coverage vectors (`chrom_coverage_<sample>_<chr>.npz`, sparse 1 x N) and the per-sample overlap-gene vectors
(`overlap_coverage_<sample>_<chr>.pkl`) into the {gene: p x L_g} dictionary GeneNMFOA.run consumes: same arguments,
same genes in the same order, same values, same per-chromosome pickles, same errors.

What is different underneath: the reference densifies a (span x p) matrix per gene group and cuts
one small array per gene out of it; here every gene's p x L_g block is written straight into ONE contiguous float64
staging buffer (pinned when a CUDA device is present), gene after gene in the reference's order, and the dictionary
holds C-contiguous views into it.  GeneNMFOA.run recognises that layout (packing._contiguous_view) and uploads the
buffer as it is -- the packed ragged buffer of include/degnorm_b200.h is built once, by the loader.
BAM parsing and the per-sample coverage files stay the reference's (reads.py); nothing here touches a GPU.
"""
import gc
import logging
import os
import pickle as pkl
from collections import OrderedDict

import numpy as np

from .packing import pinned_buffer

_SPAN = 32 << 20          # chromosome positions densified at a time per sample (256 MB of float64)


def _union_positions(starts, ends):
    """0-based chromosome positions covered by the 1-based inclusive exons (start, end): sorted union, as
    reads_coverage_merge.py:339-344 builds it."""
    if len(starts) == 1:
        return np.arange(starts[0] - 1, ends[0], dtype=np.int64)
    return np.unique(np.concatenate([np.arange(s - 1, e, dtype=np.int64) for s, e in zip(starts, ends)]))


def _plan_chrom(data_dir, sample_ids, chrom_exon_df, verbose):
    """Genes of one chromosome that come out of the chromosome coverage vectors: (chrom, [gene], [positions])
    in the reference's order (genes sorted by gene_end, :263-266), or no genes when no sample has a coverage file."""
    unique_chrom = chrom_exon_df.chr.unique()
    if len(unique_chrom) > 1:
        raise ValueError('chrom_exon_df contains exon data for more than one chromosome!')
    chrom = unique_chrom[0]
    npz_files = [os.path.join(data_dir, x, 'chrom_coverage_{0}_{1}.npz'.format(x, chrom)) for x in sample_ids]
    # (the reference draws a random order of the files from the global numpy stream to pick one that exists,
    # :233-237; the draw is kept so that the stream is left where the reference leaves it)
    has_chrom_coverage = False
    for f in np.random.choice(npz_files, size=len(npz_files), replace=False):
        if os.path.isfile(f):
            has_chrom_coverage = True
            break
    if not has_chrom_coverage:
        if verbose:
            logging.info('CHR {0} -- no chromosome coverage files available.'.format(chrom))
        return chrom, npz_files, [], []
    df = chrom_exon_df.sort_values('gene_end', axis=0)
    genes = df['gene'].unique().tolist()
    by_gene = {g: sub for g, sub in df.groupby('gene', sort=False)}
    pos = [_union_positions(by_gene[g].start.values, by_gene[g].end.values) for g in genes]
    return chrom, npz_files, genes, pos


def _load_overlap(data_dir, sample_ids, chrom):
    """Per-sample {gene: coverage vector} dictionaries of the chromosome's overlapping genes, or None when any
    sample's file is missing (the reference then returns an empty dictionary, :139-141)."""
    dicts = []
    for sample_id in sample_ids:
        cov_file = os.path.join(data_dir, sample_id, 'overlap_coverage_{0}_{1}.pkl'.format(sample_id, chrom))
        if not os.path.isfile(cov_file):
            return None
        with open(cov_file, 'rb') as f:
            dicts.append(pkl.load(f))
    return dicts


def _fill_overlap(views, dicts):
    """Rows of the overlapping genes, sample by sample (:146-157): a gene the first sample does not have is a
    KeyError as in the reference; a gene a later sample does not have keeps a zero row."""
    if not dicts:
        return
    for g in dicts[0]:
        views[g][...] = 0.0
    for i, d in enumerate(dicts):
        for g in d:
            if g not in dicts[0]:
                raise KeyError(g)
            views[g][i, :] = d[g]


def _fill_chrom(views, genes, pos, npz_files, chrom, verbose):
    """Coverage of `genes` (positions `pos`) from the samples' sparse chromosome vectors into the p x L_g views."""
    from scipy import sparse
    if not genes:
        return
    lo_g = np.array([q[0] for q in pos])
    hi_g = np.array([q[-1] + 1 for q in pos])
    # gene groups whose span stays below _SPAN positions (genes are sorted by gene_end)
    groups, a = [], 0
    while a < len(genes):
        b, lo, hi = a + 1, lo_g[a], hi_g[a]
        while b < len(genes) and max(hi, hi_g[b]) - min(lo, lo_g[b]) <= _SPAN:
            lo, hi = min(lo, lo_g[b]), max(hi, hi_g[b])
            b += 1
        groups.append((a, b, int(lo), int(hi)))
        a = b
    if verbose:
        logging.info('CHR {0} -- begin coverage matrix processing. \n'
                     'Using {1} gene splits for memory efficiency.'.format(chrom, len(groups)))
    for i, npz_file in enumerate(npz_files):
        if not os.path.isfile(npz_file):
            # no stored chromosome coverage for this sample (e.g. the chromosome was not read): zeros (:310-316)
            if verbose:
                logging.info('CHR {0} -- nonexistent chromosome coverage file {1} (imputing zeroes).'
                             .format(chrom, npz_file))
            for g in range(len(genes)):
                views[g][i, :] = 0.0
            continue
        sp = sparse.load_npz(npz_file).tocsr()
        for a, b, lo, hi in groups:
            if hi > sp.shape[1]:
                raise IndexError('gene positions beyond the end of the chromosome coverage vector {0}'.format(npz_file))
            dense = sp[:, lo:hi].toarray().ravel()
            for g in range(a, b):
                views[g][i, :] = dense[pos[g] - lo]
        del sp
    gc.collect()
    if verbose:
        logging.info('CHR {0} -- obtained {1} coverage matrices.'.format(chrom, len(genes)))


def merge_coverage(data_dir, sample_ids, exon_df, n_jobs=1, output_dir=None, verbose=True):
    """reads_coverage_merge.py:375-461.  Returns an OrderedDict {gene: p x L_g float64} for all genes in exon_df
    (per chromosome: the genes cut from the chromosome coverage vectors, sorted by gene_end, then the overlapping
    genes, whose matrices also replace same-named ones -- the dictionary merge at :431); the matrices are views of
    one staging buffer.  `n_jobs` is accepted and ignored."""
    p = len(sample_ids)
    chroms = exon_df.chr.unique()
    plans = []
    for chrom in chroms:
        chrom_df = exon_df[exon_df.chr == chrom]
        c, npz_files, genes, pos = _plan_chrom(data_dir, sample_ids, chrom_df, verbose)
        plans.append(dict(chrom=c, npz=npz_files, genes=genes, pos=pos))
    if verbose:
        logging.info('Joining overlapping genes\' coverage vectors into coverage matrices.')
    for plan in plans:
        dicts = _load_overlap(data_dir, sample_ids, plan["chrom"])
        plan["overlap"] = dicts
        ov_genes = list(dicts[0].keys()) if dicts else []
        # order and source of the chromosome's genes after {**chrom_dict, **overlap_dict}
        names = list(plan["genes"]) + [g for g in ov_genes if g not in set(plan["genes"])]
        ov = set(ov_genes)
        plan["names"] = names
        plan["from_overlap"] = [g in ov for g in names]
        where = {g: k for k, g in enumerate(plan["genes"])}
        plan["len"] = [len(dicts[0][g]) if g in ov else len(plan["pos"][where[g]]) for g in names]
    total = p * int(sum(sum(plan["len"]) for plan in plans))
    buf = pinned_buffer(total, "merge", {}).numpy()
    gene_cov_dict = OrderedDict()
    at = 0
    for plan in plans:
        views = []
        for L in plan["len"]:
            views.append(buf[at:at + p * L].reshape(p, L))
            at += p * L
        chrom_views = dict(zip(plan["names"], views))
        cut = [(g, q) for g, q, o in zip(plan["genes"], plan["pos"], plan["from_overlap"][:len(plan["genes"])]) if not o]
        _fill_chrom([chrom_views[g] for g, _ in cut], [g for g, _ in cut], [q for _, q in cut], plan["npz"],
                    plan["chrom"], verbose)
        _fill_overlap(chrom_views, plan["overlap"])
        plan["overlap"] = None
        for g in plan["names"]:
            gene_cov_dict[g] = chrom_views[g]
        if output_dir:
            save_dir = os.path.join(output_dir, str(plan["chrom"]))
            if not os.path.isdir(save_dir):
                os.makedirs(save_dir)
            chrom_cov_file = os.path.join(save_dir, 'coverage_matrices_{0}.pkl'.format(plan["chrom"]))
            if verbose:
                logging.info('CHR {0} -- saving coverage matrices to {1}'.format(plan["chrom"], chrom_cov_file))
            # (the reference's cut matrices are Fortran-ordered, its overlap matrices C-ordered: the pickles match)
            to_save = {g: (np.array(chrom_views[g]) if o else np.asfortranarray(chrom_views[g]))
                       for g, o in zip(plan["names"], plan["from_overlap"])}
            with open(chrom_cov_file, 'wb') as f:
                pkl.dump(to_save, f)
            del to_save
    gc.collect()
    return gene_cov_dict


def merge_chrom_coverage(data_dir, sample_ids, chrom_exon_df, verbose=True):
    """reads_coverage_merge.py:167-372: {gene: p x L_g float64} for the genes of ONE chromosome, cut out of the
    samples' chromosome coverage vectors (views of one staging buffer)."""
    p = len(sample_ids)
    chrom, npz_files, genes, pos = _plan_chrom(data_dir, sample_ids, chrom_exon_df, verbose)
    if not genes:
        return dict()
    total = p * int(sum(len(q) for q in pos))
    buf = pinned_buffer(total, "merge", {}).numpy()
    views, at = [], 0
    for q in pos:
        views.append(buf[at:at + p * len(q)].reshape(p, len(q)))
        at += p * len(q)
    _fill_chrom(views, genes, pos, npz_files, chrom, verbose)
    return dict(zip(genes, views))


def merge_overlap_gene_coverage(data_dir, sample_ids, chrom):
    """reads_coverage_merge.py:93-164: {gene: p x L_g float64} for the chromosome's overlapping genes."""
    dicts = _load_overlap(data_dir, sample_ids, chrom)
    if not dicts:
        return dict()
    p = len(sample_ids)
    genes = list(dicts[0].keys())
    total = p * int(sum(len(dicts[0][g]) for g in genes))
    buf = pinned_buffer(total, "merge", {}).numpy()
    out, at = dict(), 0
    for g in genes:
        L = len(dicts[0][g])
        out[g] = buf[at:at + p * L].reshape(p, L)
        at += p * L
    _fill_overlap(out, dicts)
    return out
