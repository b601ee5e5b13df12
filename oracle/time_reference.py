#!/usr/bin/env python
"""
Time the UNMODIFIED reference (imported read-only from /root/reference) on a length-stratified sample of a
BASELINE.json config, in the build container (the GPU box has no /root/reference), and the oracle port on the same
genes -- the CPU baseline BASELINE.md section 3 / SURVEY.md section 8(d) specify, plus the port/reference speed
ratio that `bench.py`'s cpu_baseline (kind "port") is read with.

    PYTHONDONTWRITEBYTECODE=1 python oracle/time_reference.py [--config c3] [--genes 208] [--workers 8]
                                                               [--subset 16] [--out profiles/r02_cpu_reference_c3.json]

Measured (all with nmf_iter=100, degnorm_iter=5, baseline selection on, BLAS threads pinned to 1):
  (a) GeneNMFOA(n_jobs=1).run            on a `--subset`-gene sub-sample (every k-th gene of the stratified list);
  (b) GeneNMFOA(n_jobs=workers).run      same genes (joblib threads; expected: no gain, the GIL binds);
  (c) the reference's MPI decomposition without MPI: a fork pool of `workers` processes over contiguous gene blocks
      (utils.split_into_chunks, nmf_mpi.py:603-629), every worker calling the reference's own
      nmf_mpi.par_apply_baseline_selection on its block, the n x p updates of nmf_mpi.py:821-838 on the parent,
      on all `--genes` genes;
  (d) the oracle port (oracle/nmfoa_oracle.py, scipy svds) with the same pool and blocks on the same genes.
"""
import argparse
import json
import os
import sys
import time
import warnings
from collections import OrderedDict

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
os.environ.setdefault("OMP_NUM_THREADS", "1")
warnings.filterwarnings("ignore")

import logging                                         # noqa: E402
logging.getLogger().setLevel(logging.ERROR)
import degnorm.nmf as ref_nmf                           # noqa: E402  (the real reference)
import degnorm.nmf_mpi as ref_mpi                       # noqa: E402
from degnorm.utils import split_into_chunks             # noqa: E402
from degnorm_b200.synth import CONFIGS, config_lengths, synth_numpy      # noqa: E402
from oracle import nmfoa_oracle as orc                  # noqa: E402

KW = dict(degnorm_iter=5, nmf_iter=100)
_W = {}


def stratified_lengths(config, n):
    """n gene lengths at evenly spaced quantiles of the config's (full-size) length distribution."""
    full = np.sort(config_lengths(config))
    q = (np.arange(n) + 0.5) / n
    return full[(q * len(full)).astype(int)]


def _init_worker(mats):
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    _W["mats"] = mats


def _ref_block(args):
    lo, hi, scale, kw = args
    dat = [(m.T / scale).T for m in _W["mats"][lo:hi]]
    _, rho, flags = ref_mpi.par_apply_baseline_selection(dat, n_jobs=1, mem_splits=1, **kw)
    return rho, flags


def _ref_init_block(args):
    lo, hi = args
    est = [ref_mpi.ratio_svd(m) for m in _W["mats"][lo:hi]]
    return np.vstack([e.sum(axis=1) for e in est]), np.vstack([m.sum(axis=1) for m in _W["mats"][lo:hi]])


def _port_block(args):
    lo, hi, scale, kw = args
    prm = orc.Params(rank1="svds", **kw)
    rows, flags = [], []
    for m in _W["mats"][lo:hi]:
        r_, _, f_ = orc.baseline_selection((m.T / scale).T, prm, 0, {})
        rows.append(r_)
        flags.append(f_)
    return np.clip(np.array(rows), 0.0, 0.9), np.array(flags)


def _port_init_block(args):
    lo, hi = args
    return (np.vstack([orc.ratio_svd(m, "svds").sum(axis=1) for m in _W["mats"][lo:hi]]),
            np.vstack([m.sum(axis=1) for m in _W["mats"][lo:hi]]))


def pooled_flow(mats, reads, workers, init_fn, block_fn, bs_kw):
    """The flow of nmf_mpi.run_gene_nmfoa_mpi with processes instead of ranks.  -> (seconds, rho, scale factors)"""
    import multiprocessing as mp
    n = len(mats)
    blocks = [(c[0], c[-1] + 1) for c in split_into_chunks(list(range(n)), n=workers)]
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(len(blocks), initializer=_init_worker, initargs=(mats,)) as pool:
        parts = pool.map(init_fn, blocks)
        est = np.vstack([a for a, _ in parts])
        cov = np.vstack([b for _, b in parts])
        rho = 1.0 - cov / (est + 1.0)
        low = rho.max(axis=1) < 0.1
        cs = reads[low].sum(axis=0) if low.any() else reads.sum(axis=0)
        norm = cs / np.median(cs)
        x_w, scale = reads / norm, norm.copy()
        for it in range(KW["degnorm_iter"]):
            parts = pool.map(block_fn, [(lo, hi, scale, bs_kw) for lo, hi in blocks])
            rho = np.vstack([a for a, _ in parts])
            x_adj = x_w / (1.0 - rho)
            nb = rho.max(axis=1) == 0
            if nb.any():
                avg = 1.0 - x_w.sum(axis=0) / x_adj.sum(axis=0)
                rho[nb] = avg
            x_adj = x_w / (1.0 - rho)
            norm = x_adj.sum(axis=0) / np.median(x_adj.sum(axis=0))
            x_w = x_w / norm
            scale = scale * norm
    return time.perf_counter() - t0, rho, scale


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c3")
    ap.add_argument("--genes", type=int, default=208)
    ap.add_argument("--workers", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--subset", type=int, default=16)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    p = cfg["p"]
    lengths = stratified_lengths(args.config, args.genes)
    lengths = np.random.default_rng(5).permutation(lengths)        # blocks should not be sorted by length
    mats, reads = synth_numpy(args.genes, p, cfg["seed"] + 501, lengths=lengths, jitter=1.0e-6)   # (no exact ties: DESIGN.md section 2)
    reads = np.maximum(reads, 1.0)
    out = dict(config=args.config, samples=p, genes=args.genes, workers=args.workers, cores=os.cpu_count(),
               lengths=dict(min=int(lengths.min()), median=float(np.median(lengths)), max=int(lengths.max()),
                            mean=float(lengths.mean())), kwargs=KW,
               versions=dict(numpy=np.__version__, scipy=__import__("scipy").__version__))
    # (c) reference, process pool
    ref_kw = dict(downsample_rate=cfg["downsample_rate"], min_high_coverage=2 if cfg["downsample_rate"] > 1 else 50,
                  nmf_iter=KW["nmf_iter"], bins=20, skip_baseline_selection=False)      # as nmf_mpi.py:777-785 passes them
    sec_c, rho_ref, scale_ref = pooled_flow(mats, reads, args.workers, _ref_init_block, _ref_block, ref_kw)
    out["reference_process_pool"] = dict(seconds=sec_c, genes_per_s=args.genes / sec_c,
                                         what="fork pool over contiguous gene blocks, every worker calls the unmodified "
                                              "nmf_mpi.par_apply_baseline_selection (n_jobs=1) on its block")
    print("(c) reference, %d processes: %.1f s  %.3f genes/s" % (args.workers, sec_c, args.genes / sec_c), flush=True)
    # (d) port, same pool
    port_kw = dict(downsample_rate=cfg["downsample_rate"], **KW)
    sec_d, rho_port, scale_port = pooled_flow(mats, reads, args.workers, _port_init_block, _port_block, port_kw)
    out["port_process_pool"] = dict(seconds=sec_d, genes_per_s=args.genes / sec_d,
                                    max_abs_rho_diff_vs_reference=float(np.abs(rho_port - rho_ref).max()),
                                    max_rel_scale_diff_vs_reference=float(np.abs(scale_port / scale_ref - 1.0).max()))
    out["port_over_reference_speed"] = sec_c / sec_d
    print("(d) port, %d processes: %.1f s  %.3f genes/s  (port/reference speed %.2f, max|d rho| %.1e)" % (
        args.workers, sec_d, args.genes / sec_d, sec_c / sec_d, out["port_process_pool"]["max_abs_rho_diff_vs_reference"]),
        flush=True)
    # (a), (b) the class on a sub-sample
    step = max(1, args.genes // args.subset)
    sub = list(range(0, args.genes, step))[:args.subset]
    cov = OrderedDict(("gene_%d" % g, mats[g]) for g in sub)
    for tag, nj in (("reference_n_jobs_1", 1), ("reference_n_jobs_%d_threads" % args.workers, args.workers)):
        model = ref_nmf.GeneNMFOA(n_jobs=nj, downsample_rate=cfg["downsample_rate"], **KW)
        t0 = time.perf_counter()
        model.run(cov, reads[sub].copy())
        sec = time.perf_counter() - t0
        out[tag] = dict(seconds=sec, genes=len(sub), genes_per_s=len(sub) / sec)
        print("%s: %d genes %.1f s  %.3f genes/s" % (tag, len(sub), sec, len(sub) / sec), flush=True)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
