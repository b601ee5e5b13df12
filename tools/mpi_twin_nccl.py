#!/usr/bin/env python
"""run_gene_nmfoa_mpi over NCCL, one rank per GPU (torchrun), checked against the single-GPU GeneNMFOA on rank 0.
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/mpi_twin_nccl.py"""
import os
import sys
from collections import OrderedDict

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from degnorm_b200 import GeneNMFOA, run_gene_nmfoa_mpi       # noqa: E402
from degnorm_b200.synth import synth_numpy                    # noqa: E402

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
kw = dict(degnorm_iter=3, nmf_iter=40, downsample_rate=4)
lengths = np.random.default_rng(7).integers(200, 5000, size=41)
mats, reads = synth_numpy(len(lengths), 12, 99, lengths=lengths, jitter=1e-6)
cov = OrderedDict(("g%d" % i, m) for i, m in enumerate(mats))
out = run_gene_nmfoa_mpi(dist.group.WORLD, cov if rank == 0 else OrderedDict(), reads, **kw)
if rank == 0:
    single = GeneNMFOA(**kw)
    est = single.run(cov, reads)
    assert np.array_equal(out["ran_baseline_selection"], single.ran_baseline_selection)
    d_rho = float(np.abs(out["rho"] - single.rho).max())
    d_adj = float((np.abs(out["x_adj"] - single.x_adj) / np.maximum(1.0, np.abs(single.x_adj))).max())
    d_est = max(float(np.abs(a - b).max()) for a, b in zip(out["estimates"].values(), est))
    assert d_rho < 1e-11 and d_adj < 1e-11 and d_est < 1e-8, (d_rho, d_adj, d_est)
    print("mpi twin over NCCL, %d ranks: max|dDI| %.2e, max rel d x_adj %.2e, max|d est| %.2e" % (
        dist.get_world_size(), d_rho, d_adj, d_est))
else:
    assert out is None
dist.destroy_process_group()
