"""Multi-worker path.  CPU part (gloo, world_size 2): the partition, the object send/recv, and the claim the
whole multi-GPU design rests on -- that an all-reduce of 3p+1 per-sample sums per outer iteration is all the
workers need to exchange (checked against the single-process oracle).  GPU part: run_gene_nmfoa_mpi on two
processes sharing one B200 (gloo) equals the single-process GeneNMFOA."""
import os
import socket
import sys
from collections import OrderedDict

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_partition_matches_reference_chunks():
    from degnorm_b200.distributed import partition_bounds
    # utils.split_into_chunks(list(range(n)), n=size): chunk size ceil(n/size) (utils.py:176-192)
    assert partition_bounds(10, 3) == [(0, 4), (4, 8), (8, 10)]
    assert partition_bounds(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    assert partition_bounds(9, 1) == [(0, 9)]
    assert partition_bounds(0, 2) == [(0, 0), (0, 0)]
    for n in (1, 7, 64, 1000):
        for size in (1, 2, 3, 8):
            b = partition_bounds(n, size)
            assert b[0][0] == 0 and b[-1][1] == n and all(x[1] == y[0] for x, y in zip(b[:-1], b[1:]))


# ---- sharded restatement of the n x p steps the CUDA library runs per rank (dn_init_sums / dn_init_apply /
#      dn_outer_sums / dn_outer_apply, include/degnorm_b200.h), with the oracle doing the per-gene work
def _sharded_oracle_run(comm, mats, reads, kw):
    import torch
    from oracle import nmfoa_oracle as orc
    prm = orc.Params(rank1="gram", **kw)
    p = reads.shape[1]
    n = len(mats)
    sums = np.zeros(3 * p + 1)
    if n:
        est = np.array([orc.ratio_svd(F, "gram").sum(axis=1) for F in mats])
        cov = np.array([F.sum(axis=1) for F in mats])
        low = (1.0 - cov / (est + 1.0)).max(axis=1) < 0.1
        sums[:p] = reads[low].sum(axis=0)
        sums[p:2 * p] = reads.sum(axis=0)
        sums[3 * p] = low.sum()
    t = torch.from_numpy(sums)
    comm.allreduce_(t)
    cs = sums[:p] if sums[3 * p] > 0 else sums[p:2 * p]
    norm = cs / np.median(cs)
    x_w, scale = reads / norm, norm.copy()
    rho = np.zeros((n, p))
    x_adj = np.zeros((n, p))
    for it in range(prm.degnorm_iter):
        rows = [orc.baseline_selection((F.T / scale).T, prm, None, {})[0] for F in mats]
        rho = np.clip(np.array(rows).reshape(n, p), 0.0, 0.9)
        nb = rho.max(axis=1) == 0 if n else np.zeros(0, dtype=bool)
        sums = np.zeros(3 * p + 1)
        sums[:p] = x_w.sum(axis=0)
        sums[p:2 * p] = (x_w[~nb] / (1.0 - rho[~nb])).sum(axis=0)
        sums[2 * p:3 * p] = x_w[nb].sum(axis=0)
        t = torch.from_numpy(sums)
        comm.allreduce_(t)
        avg = 1.0 - sums[:p] / (sums[p:2 * p] + sums[2 * p:3 * p])
        col = sums[p:2 * p] + sums[2 * p:3 * p] / (1.0 - avg)
        norm = col / np.median(col)
        rho[nb] = avg
        x_adj = x_w / (1.0 - rho)
        x_w = x_w / norm
        scale = scale * norm
    return rho, x_adj, scale


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from degnorm_b200.distributed import TorchComm, partition_bounds
        from degnorm_b200.synth import synth_numpy
        comm = TorchComm()
        assert (comm.rank, comm.size) == (rank, world)
        # object send / recv (the one-time shard scatter and the final gather use it)
        if rank == 0:
            comm.send_obj({"a": np.arange(3)}, dest=1)
        else:
            assert comm.recv_obj(source=0)["a"].tolist() == [0, 1, 2]
        kw = dict(degnorm_iter=2, nmf_iter=15)
        mats, reads = synth_numpy(5, 3, 11, lengths=np.array([260, 300, 280, 340, 250]), jitter=1e-6)
        lo, hi = partition_bounds(len(mats), world)[rank]
        rho, x_adj, scale = _sharded_oracle_run(comm, mats[lo:hi], reads[lo:hi], kw)
        q.put((rank, lo, hi, rho, x_adj, scale))
        comm.barrier()
    finally:
        dist.destroy_process_group()


def test_balanced_partition_covers_every_gene_once_and_levels_the_load():
    """genes to workers by estimated work instead of the reference's contiguous blocks."""
    from degnorm_b200.distributed import balanced_partition, partition_bounds
    rng = np.random.default_rng(3)
    for n, size in ((0, 3), (1, 4), (7, 2), (1000, 8), (5000, 3)):
        cost = np.sort(np.exp(rng.normal(8.0, 0.9, size=n)))[::-1].copy()       # annotation order: long genes first
        shards = balanced_partition(cost, size)
        assert len(shards) == size
        allg = np.concatenate(shards) if n else np.zeros(0, dtype=np.int64)
        assert sorted(allg.tolist()) == list(range(n))
        assert all((np.diff(s) > 0).all() for s in shards)                       # ascending: gene order is kept
        if n >= size:
            loads = np.array([cost[s].sum() for s in shards])
            assert loads.max() <= cost.sum() / size + cost.max() + 1e-9           # the greedy bound
            contiguous = np.array([cost[lo:hi].sum() for lo, hi in partition_bounds(n, size)])
            assert loads.max() <= contiguous.max() + 1e-9
    assert [s.tolist() for s in balanced_partition([5, 1, 4, 2], 2)] == [[0, 1], [2, 3]]


def test_gloo_two_workers_equal_one():
    import torch.multiprocessing as mp
    from degnorm_b200.synth import synth_numpy
    from oracle import nmfoa_oracle as orc
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    got = sorted([q.get(timeout=240) for _ in procs], key=lambda t: t[0])
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    kw = dict(degnorm_iter=2, nmf_iter=15)
    mats, reads = synth_numpy(5, 3, 11, lengths=np.array([260, 300, 280, 340, 250]), jitter=1e-6)
    ref = orc.run(mats, reads, orc.Params(rank1="gram", **kw), want_estimates=False)
    rho = np.vstack([g[3] for g in got])
    x_adj = np.vstack([g[4] for g in got])
    np.testing.assert_allclose(rho, ref["rho"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(x_adj, ref["x_adj"], rtol=1e-12, atol=1e-12)
    for g in got:
        np.testing.assert_allclose(g[5], ref["scale_factors"], rtol=1e-13)


def _subgroup_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from degnorm_b200.distributed import TorchComm
        grp = dist.new_group(ranks=[1, 2])                   # a group that does not start at global rank 0
        out = None
        if rank in (1, 2):
            c = TorchComm(grp)
            assert (c.rank, c.size) == (rank - 1, 2)
            if c.rank == 0:
                c.send_obj({"hello": 42}, dest=1, tag=333)       # group rank 1 = global rank 2
                got = c.recv_obj(source=1, tag=666)
            else:
                got = c.recv_obj(source=0, tag=333)
                c.send_obj("back", dest=0, tag=666)
            t = torch.tensor([float(rank)], dtype=torch.float64)
            c.allreduce_(t)
            c.barrier()
            out = (got, float(t.item()))
        q.put((rank, out))
    except Exception as exc:
        import traceback
        q.put((rank, "ERROR: %s\n%s" % (exc, traceback.format_exc())))
    finally:
        dist.destroy_process_group()


def test_torchcomm_speaks_group_ranks_in_a_subgroup():
    """TorchComm takes group ranks (as run_gene_nmfoa_mpi does); torch's object send/recv take GLOBAL ranks: in a
    sub-group that does not start at global rank 0 the two differ."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_subgroup_worker, args=(r, 3, port, q)) for r in range(3)]
    for pr in procs:
        pr.start()
    got = dict(q.get(timeout=240) for _ in procs)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert got[0] is None
    assert got[1] == ("back", 3.0) and got[2] == ({"hello": 42}, 3.0), got


# ------------------------------------------------------------------------------------------------ GPU
def _gpu_worker(rank, world, port, q, partition):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from degnorm_b200 import run_gene_nmfoa_mpi
        from degnorm_b200.synth import synth_numpy
        kw = dict(degnorm_iter=2, nmf_iter=30, downsample_rate=3)
        mats, reads = synth_numpy(7, 4, 5, lengths=np.array([300, 420, 700, 256, 512, 900, 333]), jitter=1e-6)
        cov = OrderedDict(("g%d" % i, m) for i, m in enumerate(mats))
        out = run_gene_nmfoa_mpi(dist.group.WORLD, cov if rank == 0 else OrderedDict(), reads, device="cuda:0",
                                 partition=partition, **kw)
        q.put((rank, None if out is None else {k: (v if k != "estimates" else list(v.values())) for k, v in out.items()}))
    except Exception as exc:                                   # surface the failure instead of a silent time-out
        import traceback
        q.put((rank, "ERROR: %s\n%s" % (exc, traceback.format_exc())))
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("partition", ["balanced", "contiguous"])
def test_two_processes_on_one_gpu_equal_single_process(partition):
    import torch.multiprocessing as mp
    from degnorm_b200 import GeneNMFOA, run_gene_nmfoa_mpi
    from degnorm_b200.synth import synth_numpy
    kw = dict(degnorm_iter=2, nmf_iter=30, downsample_rate=3)
    mats, reads = synth_numpy(7, 4, 5, lengths=np.array([300, 420, 700, 256, 512, 900, 333]), jitter=1e-6)
    cov = OrderedDict(("g%d" % i, m) for i, m in enumerate(mats))
    single = GeneNMFOA(**kw)
    est = single.run(cov, reads)
    est = [np.array(e) for e in est]
    solo = run_gene_nmfoa_mpi(None, cov, reads, **kw)                  # no communicator: one worker
    np.testing.assert_array_equal(solo["rho"], single.rho)
    np.testing.assert_array_equal(solo["ran_baseline_selection"], single.ran_baseline_selection)
    assert list(solo["estimates"].keys()) == list(cov.keys())

    class OneRankMPI(object):                     # the lowercase mpi4py object API degnorm_mpi passes (COMM_WORLD)
        rank, size = 0, 1

        def send(self, obj, dest, tag=0):
            raise AssertionError("nothing to send with one rank")

        def recv(self, source, tag=0):
            raise AssertionError("nothing to receive with one rank")

        def allreduce(self, x):
            return x

        def Barrier(self):
            pass
    via_mpi = run_gene_nmfoa_mpi(OneRankMPI(), cov, reads, **kw)
    np.testing.assert_array_equal(via_mpi["rho"], single.rho)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gpu_worker, args=(r, 2, port, q, partition)) for r in range(2)]
    for pr in procs:
        pr.start()
    got = dict(q.get(timeout=240) for _ in procs)
    for v in got.values():
        assert not isinstance(v, str), v
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert got[1] is None
    out = got[0]
    np.testing.assert_array_equal(out["ran_baseline_selection"], single.ran_baseline_selection)
    np.testing.assert_allclose(out["rho"], single.rho, rtol=0, atol=1e-12)
    np.testing.assert_allclose(out["x_adj"], single.x_adj, rtol=1e-12, atol=1e-12)
    for a, b in zip(out["estimates"], est):
        np.testing.assert_allclose(a, b, rtol=1e-10, atol=1e-10)


@pytest.mark.gpu
def test_mpi_twin_over_nccl_one_rank_per_gpu():
    """run_gene_nmfoa_mpi with a torch.distributed NCCL group, one rank per GPU (torchrun), against the single-GPU
    class on rank 0 (tools/mpi_twin_nccl.py asserts DI / adjusted counts / estimates / flags).  Needs two devices."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two devices")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(root, "tools", "mpi_twin_nccl.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "mpi twin over NCCL, 2 ranks" in r.stdout
