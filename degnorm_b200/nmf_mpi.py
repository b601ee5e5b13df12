"""
Drop-in for the reference's degnorm/nmf_mpi.py entry points (run_gene_nmfoa_mpi :555-863, save_results :448-552):
the same call, the same return value on rank 0, with every worker driving one B200 instead of a CPU node.

What changes underneath: rank 0 sends each worker its block of coverage matrices ONCE
(nmf_mpi.py:627); the per-iteration re-send of re-scaled matrices (:745-760) and the gathers of whole estimate
dictionaries (:690, :797) are gone -- scale factors are applied on load inside the kernel and the only
per-iteration exchange is an all-reduce of 3p+1 doubles.  n x p results are gathered to rank 0 at the end.

Down-sampling offsets: drawn once on rank 0 in the single-node order (np.random.seed(random_state); one draw per
gene per outer iteration, iteration-major) and sent with the shard, so a multi-worker run equals the single-node
run (the reference re-seeds every rank, nmf_mpi.py:731, and therefore does not -- SURVEY.md App. C-2).
"""
import logging
from collections import OrderedDict

import numpy as np

from .distributed import adapt, balanced_partition, partition_bounds
from .engine import Params, ShardEngine, draw_offsets
from .nmf import GeneNMFOA
from .packing import pack_coverage, unpack_estimates


def _check_input(x, cov_mats, n_genes, prm):
    """nmf_mpi.py:648-668 (rank 0 only, same messages)."""
    if x.shape[0] != n_genes:
        raise ValueError('Number of genes in read count matrix not equal to number of coverage matrices!')
    if not all(map(lambda z: z.ndim == 2, cov_mats)):
        raise ValueError('Not all coverage matrices are 2-d arrays!')
    p = cov_mats[0].shape[0]
    li_vec = np.array([m.shape[1] for m in cov_mats])
    if np.sum(li_vec / p < 1) > 0:
        logging.warning('At least one coverage matrix is taller than it is wide.'
                        'Ensure that coverage matrices are shaped (p x L_i).')
    if prm.downsample_rate > 1:
        if not np.min(li_vec) >= prm.downsample_rate:
            raise ValueError('downsample_rate is too large; take-every size > at least one gene.')
    return p


def run_gene_nmfoa_mpi(comm, cov_dat, reads_dat, degnorm_iter=5, downsample_rate=1, min_high_coverage=50,
                       nmf_iter=100, bins=20, n_jobs=1, skip_baseline_selection=False, random_state=123,
                       device=None, return_estimates=True, partition='balanced'):
    """Same contract as nmf_mpi.py:555-863.  `comm`: an mpi4py communicator (as degnorm_mpi passes), a
    torch.distributed group, or None.  Rank 0's cov_dat / reads_dat are authoritative.  Returns on rank 0
    {'estimates': OrderedDict gene -> p x L_g, 'rho', 'x_adj', 'ran_baseline_selection'}; None elsewhere.
    partition: 'balanced' (genes to workers by estimated work, results scattered back into cov_dat order) or
    'contiguous' (the reference's blocks, nmf_mpi.py:605-606)."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("degnorm_b200 needs a CUDA device (B200); there is no CPU fallback")
    c = adapt(comm)
    prm = Params(degnorm_iter=degnorm_iter, downsample_rate=downsample_rate, min_high_coverage=min_high_coverage,
                 nmf_iter=nmf_iter, bins=bins, n_jobs=n_jobs, skip_baseline_selection=skip_baseline_selection,
                 random_state=random_state)
    rank, size = c.rank, c.size
    if partition not in ('balanced', 'contiguous'):          # (checked on every rank: nobody is left waiting)
        raise ValueError("partition must be 'balanced' or 'contiguous'")
    # one GPU per worker; made current before any communication (NCCL object collectives stage through it)
    dev = torch.device(device if device is not None else "cuda:%d" % (rank % torch.cuda.device_count()))
    torch.cuda.set_device(dev)
    if rank == 0:
        try:
            genes = list(cov_dat.keys())
            n_genes = len(genes)
            x = np.ascontiguousarray(np.copy(reads_dat), dtype=np.float64)
            mats = list(cov_dat.values())
            p = _check_input(x, mats, n_genes, prm)
            ds = draw_offsets(n_genes, prm)                     # seeds the global numpy stream (nmf.py:556)
            if partition == 'contiguous':
                shards = [np.arange(lo, hi) for lo, hi in partition_bounds(n_genes, size)]
            else:
                li = np.array([m.shape[1] for m in mats], dtype=np.int64)
                shards = balanced_partition(p * ((li + prm.downsample_rate - 1) // prm.downsample_rate), size)
        except Exception as exc:
            # the workers are waiting for their shard: tell them, so that every rank raises instead of hanging
            for w in range(1, size):
                c.send_obj(dict(error="%s: %s" % (type(exc).__name__, exc)), dest=w, tag=333 + w)
            raise

        def shard(idx):
            return dict(p=p, mats=[mats[g] for g in idx], reads=x[idx], ds=None if ds is None else ds[:, idx])
        for w in range(size):
            logging.info('(%d/%d) -- %s will be responsible for %d genes.', rank + 1, size,
                         'host' if w == 0 else 'worker node %d' % w, len(shards[w]))
            if w > 0:
                c.send_obj(shard(shards[w]), dest=w, tag=333 + w)
        mine = shard(shards[0])
    else:
        mine = c.recv_obj(source=0, tag=333 + rank)
        if "error" in mine:
            raise ValueError("rank 0 rejected the input -- " + mine["error"])
    p = mine["p"]
    with torch.cuda.device(dev):
        flat, offsets = pack_coverage(mine["mats"], p) if len(mine["mats"]) else (torch.zeros(0, dtype=torch.float64),
                                                                                  np.zeros(1, dtype=np.int64))
        eng = ShardEngine(prm, p, dev, allreduce=c.allreduce_)
        eng.load(flat.to(dev), offsets, torch.from_numpy(np.ascontiguousarray(mine["reads"], dtype=np.float64)
                                                          .reshape(-1, p)).to(dev))
        out = eng.run(mine["ds"], want_estimates=return_estimates)
        torch.cuda.synchronize(dev)
        eng.check_exit_codes()
        part = dict(rho=out["rho"].cpu().numpy(), x_adj=out["x_adj"].cpu().numpy(),
                    ran=out["ran"].cpu().numpy().T.astype(bool), est=None)
        if return_estimates and out["est"] is not None:
            part["est"] = [np.array(m) for m in unpack_estimates(out["est"], offsets, p)]
    # star gather of the n x p results (and the last iteration's estimates) through rank 0 (:809-815)
    if rank > 0:
        c.send_obj(part, dest=0, tag=666 + rank)
        c.barrier()
        return None
    parts = [part] + [c.recv_obj(source=w, tag=666 + w) for w in range(1, size)]
    c.barrier()
    # scatter the workers' rows back into cov_dat order
    rho = np.zeros((n_genes, p))
    x_adj = np.zeros((n_genes, p))
    ran = np.zeros((n_genes, prm.degnorm_iter), dtype=bool)
    est_list = [None] * n_genes
    for idx, q in zip(shards, parts):
        if len(idx) == 0:
            continue
        rho[idx], x_adj[idx], ran[idx] = q["rho"], q["x_adj"], q["ran"]
        if return_estimates and q["est"] is not None:
            for g, m in zip(idx, q["est"]):
                est_list[g] = m
    estimates = None
    if return_estimates:
        estimates = OrderedDict(zip(genes, est_list))
    return {'estimates': estimates, 'rho': rho, 'x_adj': x_adj, 'ran_baseline_selection': ran}


def save_results(gene_manifest_df, estimates, rho, x_adj, ran_baseline_selection, sample_ids, output_dir):
    """nmf_mpi.py:448-552: same files and columns as GeneNMFOA.save_results, from the returned pieces."""
    m = GeneNMFOA(degnorm_iter=ran_baseline_selection.shape[1])
    m.genes = list(estimates.keys())
    m.rho, m.x_adj, m.ran_baseline_selection = rho, x_adj, ran_baseline_selection
    m.p = rho.shape[1]
    m.n_genes = rho.shape[0]
    m.fitted = True
    m.save_results(list(estimates.values()), gene_manifest_df, output_dir=output_dir, sample_ids=sample_ids)
