#!/usr/bin/env python
"""
bench.py -- genes/s of the full DegNorm NMF-OA path (degnorm_iter=5, nmf_iter=100, baseline selection on) on
synthetic coverage of BASELINE.json's shapes.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config c2] [--impl ours|reference]

A "step" is one complete GeneNMFOA flow (init ratio-SVD pass, 5 outer iterations of fused baseline selection +
scale-factor update, estimates of the last iteration) over one synthetic batch of genes:
  * `value`  : device-resident -- the packed coverage buffer is already in HBM when the timed region starts;
  * `e2e`    : the same flow through the drop-in GeneNMFOA.run() with HOST (pinned) numpy inputs, host->device and
               device->host copies (rho, x_adj, flags, estimates) inside the timed region;
  * `roofline`: fused baseline-selection kernel (the dominant launch group, one group per outer iteration):
               algorithmic bytes (raw-coverage scan + (24T+24) p L' per nmf() call, as if every pass
               were streamed) / CUDA-event time of the group on the launching stream / measured HBM peak;
  * `cpu_baseline`: the oracle port of the reference path (scipy svds rank-one step, exactly the reference's
               third-party call) on a bounded sample of the same workload, all host cores, reference-style gene
               decomposition (nmf_mpi.py:603-629: contiguous gene blocks per worker, rank 0 does the n x p updates).

With N > 1 (torchrun, one process per GPU) every rank owns its own batch of the same shape (weak scaling); the
only collective is the all-reduce of 3p+1 per-sample sums per outer iteration (NCCL).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from collections import OrderedDict

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "genes/sec full NMF-OA (5 outer iters)"
UNIT = "genes/s"
RUN_KW = dict(degnorm_iter=5, nmf_iter=100)


def host_memory_allows(nbytes):
    """True if `nbytes` of pinned host memory (all ranks of this node together) leave half of what is available."""
    avail = None
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        pass
    try:
        lim = open("/sys/fs/cgroup/memory.max").read().strip()
        if lim != "max":
            cur = int(open("/sys/fs/cgroup/memory.current").read().strip())
            avail = min(avail, int(lim) - cur) if avail is not None else int(lim) - cur
    except Exception:
        pass
    return avail is None or nbytes <= 0.5 * avail


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="c2")
    ap.add_argument("--genes", type=int, default=0, help="override the number of genes (debugging)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-genes", type=int, default=0, help="genes in the CPU sample (default: 2 per core)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--serial-buckets", action="store_true", help="tuning: run the tiers one after another")
    ap.add_argument("--force-cluster", type=int, default=0, help="tuning: every gene through clusters of this size")
    ap.add_argument("--force-streamed", action="store_true", help="tuning: every gene through the streamed tier")
    ap.add_argument("--cluster-min", type=int, default=-2, help="tuning: genes up to this many columns stay on one CTA (streamed)")
    ap.add_argument("--mid-clusters", default="", help="tuning: mid-p cluster table as size:max_cols,...")
    ap.add_argument("--mid-warps", type=int, default=0, help="tuning: warps per CTA of the mid-p kernel (4: two CTAs per SM)")
    ap.add_argument("--max-len", type=int, default=0, help="tuning: clip gene lengths (removes the long-gene tail)")
    ap.add_argument("--tiers", default="", help="small-p tiers as cols:warps,... (tuning)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ CPU reference arm
_W = {}


def _worker_init(mats, kw):
    os.environ["OMP_NUM_THREADS"] = "1"
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    from oracle import nmfoa_oracle as orc
    _W["mats"] = mats
    _W["prm"] = orc.Params(rank1="svds", **kw)
    _W["orc"] = orc


def _worker_init_pass(lohi):
    orc = _W["orc"]
    lo, hi = lohi
    est = [orc.ratio_svd(F, "svds").sum(axis=1) for F in _W["mats"][lo:hi]]
    cov = [F.sum(axis=1) for F in _W["mats"][lo:hi]]
    return np.array(est), np.array(cov)


def _worker_bs(args):
    orc = _W["orc"]
    lo, hi, scale, offs = args
    rows, flags = [], []
    for g in range(lo, hi):
        r_, _, f_ = orc.baseline_selection((_W["mats"][g].T / scale).T, _W["prm"], int(offs[g - lo]), {})
        rows.append(r_)
        flags.append(f_)
    return np.array(rows), np.array(flags)


def cpu_reference_run(mats, reads, kw, cores):
    """The reference's flow with its MPI decomposition (contiguous gene blocks per worker), fork-based.
    Returns seconds for the whole flow on these genes."""
    import multiprocessing as mp
    from oracle import nmfoa_oracle as orc
    prm = orc.Params(rank1="svds", **kw)
    n = len(mats)
    cores = max(1, min(cores, n))
    cs = -(-n // cores)
    blocks = [(lo, min(lo + cs, n)) for lo in range(0, n, cs)]
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(len(blocks), initializer=_worker_init, initargs=(mats, kw)) as pool:
        parts = pool.map(_worker_init_pass, blocks)
        est = np.vstack([a for a, _ in parts])
        cov = np.vstack([b for _, b in parts])
        rho0 = 1.0 - cov / (est + 1.0)
        low = rho0.max(axis=1) < 0.1
        cs_ = reads[low].sum(axis=0) if low.any() else reads.sum(axis=0)
        norm = cs_ / np.median(cs_)
        x_w, scale = reads / norm, norm.copy()
        offs = orc.draw_offsets(n, prm)
        for it in range(prm.degnorm_iter):
            parts = pool.map(_worker_bs, [(lo, hi, scale, offs[it, lo:hi]) for lo, hi in blocks])
            rho = np.clip(np.vstack([a for a, _ in parts]), 0.0, 0.9)
            _, norm, x_w, scale = orc.outer_update(x_w, rho, scale)
    return time.perf_counter() - t0


def cpu_sample(cfg, n_sample, seed):
    """A bounded sample of the workload: n_sample genes with lengths drawn like the config's."""
    from degnorm_b200.synth import synth_numpy, gene_lengths
    rng = np.random.default_rng(seed)
    lengths = gene_lengths(n_sample, rng, cfg["profile"])
    return synth_numpy(n_sample, cfg["p"], seed + 1, lengths=lengths)


def run_cpu_baseline(cfg, kw, n_sample, cores):
    mats, reads = cpu_sample(cfg, n_sample, cfg["seed"] + 77)
    reads = np.maximum(reads, 1.0)
    secs = cpu_reference_run(mats, reads, kw, cores)
    return dict(value=n_sample / secs, unit=UNIT, cores=cores, kind="port",
                sample="%d genes of %s (%d samples, downsample_rate %d), full flow, %.1f s wall; oracle port with "
                       "scipy svds (the reference's own third-party call), fork pool, contiguous gene blocks"
                       % (n_sample, cfg["name"], cfg["p"], kw.get("downsample_rate", 1), secs))


# ------------------------------------------------------------------------------------------------ clocks sampler
class Clocks(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = False
        self.sm, self.reasons, self.sm_max = [], set(), None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                self.sm.append(float(f[0]))
                self.sm_max = float(f[1])
                for nm, v in zip(names, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        return dict(sm_mhz=float(np.median(self.sm)) if self.sm else None, sm_max_mhz=self.sm_max,
                    reasons=sorted(self.reasons), samples=len(self.sm))


# ------------------------------------------------------------------------------------------------ main
def main():
    args = parse()
    # libraries (NCCL's version banner) write to fd 1: park the real stdout and hand fd 1 to stderr until the JSON line
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())
    from degnorm_b200.synth import CONFIGS, config_lengths
    cfg = dict(CONFIGS[args.config])
    cfg["name"] = args.config
    if args.genes:
        cfg["n_genes"] = args.genes
    kw = dict(RUN_KW)
    kw["downsample_rate"] = cfg["downsample_rate"]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1
    workload = "%s: %d genes x %d samples, log-normal lengths (%s), downsample_rate=%d, per GPU" % (
        args.config, cfg["n_genes"], cfg["p"], cfg["profile"], cfg["downsample_rate"])
    config = dict(workload=workload, degnorm_iter=kw["degnorm_iter"], nmf_iter=kw["nmf_iter"], baseline_selection=True,
                  downsample_rate=cfg["downsample_rate"], l2="inputs (GBs of coverage) larger than L2")

    if args.impl == "reference":
        if rank != 0:
            return
        n_sample = args.cpu_genes or 2 * cores
        vals = []
        for _ in range(args.warmup):
            pass                                    # a CPU run has nothing to warm; the pool forks per step
        for _ in range(max(1, args.steps)):
            cb = run_cpu_baseline(cfg, kw, n_sample, cores)
            vals.append(cb["value"])
        v = float(np.mean(vals))
        cb["value"] = v
        emit(dict(metric=METRIC, value=v, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                              ms_per_step=1000.0 * n_sample / v, higher_is_better=True, scaling="weak", vs_baseline=None,
                              dtype="f64", data="synthetic", config=config, impl="reference", cpu_baseline=cb,
                              e2e=dict(value=v, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0))
        return

    import torch
    import torch.distributed as dist
    from degnorm_b200 import GeneNMFOA
    from degnorm_b200.engine import Params, ShardEngine, draw_offsets
    from degnorm_b200.synth import synth_torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD

    # ---- synthetic batch, generated on the device (each rank its own seed)
    n, p = cfg["n_genes"], cfg["p"]
    lengths = config_lengths(args.config, n)
    if args.max_len:
        lengths = np.minimum(lengths, args.max_len)
    if world > 1:
        lengths = np.random.default_rng(cfg["seed"] + 1000 * rank).permutation(lengths)
    cov, off, reads = synth_torch(lengths, p, cfg["seed"] + 1000 * rank, dev)
    reads = torch.clamp(reads, min=1.0)
    prm = Params(**kw)
    ds = draw_offsets(n, prm)
    tiers = tuple(tuple(int(x) for x in t.split(":")) for t in args.tiers.split(",")) if args.tiers else None
    eng = ShardEngine(prm, p, dev, group=group, small_tiers=tiers, force_streamed=args.force_streamed)
    eng.force_cluster = args.force_cluster
    if args.mid_clusters:
        eng.mid_clusters = tuple(tuple(int(x) for x in t.split(":")) for t in args.mid_clusters.split(","))
    if args.mid_warps:
        eng.mid_warps = args.mid_warps
    if args.cluster_min >= -1:
        eng.cluster_min_cols = args.cluster_min
    eng.load(cov, off, reads)
    eng.record_events = True
    eng.serial_buckets = args.serial_buckets

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    clocks = Clocks(local)
    for _ in range(args.warmup):
        eng.run(ds, want_estimates=True)
    barrier()
    if rank == 0:
        clocks.start()
    phase = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        eng.run(ds, want_estimates=True)
        # (event times are read after the final synchronize; keep the handles of every step)
        phase.setdefault("_ev", []).append(list(eng.events))
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = eng.launches * args.steps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = world * n / (ms_per_step / 1000.0)

    # ---- roofline of the dominant kernel group (fused baseline selection), live CUDA-event timing
    bs_ms, all_ms = [], {}
    for evs in phase["_ev"]:
        for (n0, a), (n1, b) in zip(evs[:-1], evs[1:]):
            dt = a.elapsed_time(b)
            all_ms[n1.rstrip("0123456789")] = all_ms.get(n1.rstrip("0123456789"), 0.0) + dt
            if n1.startswith("bs"):
                bs_ms.append(dt)
    bs_bytes = eng.bs_bytes_per_iteration()
    total_bytes, parts = eng.algorithmic_bytes()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = float(np.mean(bs_bytes)) / (float(np.mean(bs_ms)) / 1000.0) / 1e9
    cnt = eng.out["counters"].cpu().numpy()
    # DRAM bytes of the launch group from the committed ncu pass of this very workload (profiles/), else null
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic_%s.json" % args.config)))
        if args.config == "c2" and n == 20000 and not args.tiers and not args.force_cluster and not args.force_streamed:
            traffic, traffic_src = float(tj["dram_bytes_per_launch_group"]), "profiles/r01_traffic_%s.json (ncu)" % args.config
    except Exception:
        pass
    kname = "nmfoa_small_kernel" if p <= 12 else ("nmfoa_mid_kernel" if p <= 48 else "nmfoa_kernel (tiled)")
    roofline = dict(bound="hbm", kernel=kname + " (fused baseline selection; one launch group per outer iteration)",
                    achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak, traffic=traffic,
                    traffic_source=traffic_src,
                    peak_source="MEASURED_PEAKS.json (measured copy bandwidth)" if peaks else "fallback 6650 GB/s",
                    algorithmic_bytes_per_launch=float(np.mean(bs_bytes)), ms_per_launch=float(np.mean(bs_ms)),
                    share_of_step=float(np.sum(bs_ms)) / ms,
                    note="algorithmic = as-if-streamed bytes (SURVEY 8d); genes resident in shared memory never touch HBM again inside an outer iteration, so real DRAM traffic is ~1 % of it and frac > 1 is possible",
                    resident_fraction=float((cnt[-1, :, 7] & 1).mean()),
                    nmf_calls_per_gene=float(cnt[:, :, 2].mean()), phases_ms_per_step={k: v / args.steps for k, v in all_ms.items()})
    solves = float((cnt[:, :, 2].astype(np.float64) * (kw["nmf_iter"] + 1)).sum())
    roofline["eig_steps_per_solve"] = float(cnt[:, :, 4].astype(np.float64).sum()) / max(solves, 1.0)
    roofline["eig_fallback_solves"] = int((cnt[:, :, 7] >> 1).sum())
    roofline["genes_with_fallbacks"] = int(((cnt[:, :, 7] >> 1) > 0).any(axis=0).sum())
    bk = eng.bucket_ms()
    cnt_last = cnt[-1]
    roofline["buckets"] = [dict(max_cols=bb.max_cols, genes=bb.n, threads=int(bb.plan.threads), ctas=int(bb.plan.ctas),
                                smem=int(bb.plan.smem_bytes), resident=int(bb.plan.resident_cols), cluster=int(bb.plan.cluster),
                                end_ms=[round(bk[it][k], 2) for it in sorted(bk)],
                                sum_cols=int(cnt_last[bb.order.cpu().numpy(), 3].sum()))
                           for k, bb in enumerate(eng.buckets)]
    clock_summary = None
    if rank == 0:
        clocks.stop_flag = True

    # ---- end to end through the drop-in class, host buffers
    e2e = None
    if not args.no_e2e and not host_memory_allows(2 * cov.numel() * 8 * world):
        # (every rank pins its coverage and its estimates: 13 GB per rank at C2; never drive the box out of memory)
        e2e = dict(value=None, unit=UNIT, skipped="pinned host buffers of all ranks would not fit in host memory")
    elif not args.no_e2e:
        host = torch.empty(cov.numel(), dtype=torch.float64).pin_memory()
        host.copy_(cov)
        torch.cuda.synchronize(dev)
        arr = host.numpy()
        reads_h = reads.cpu().numpy()
        cov_dict = OrderedDict(("g%d" % g, arr[p * int(off[g]):p * int(off[g + 1])].reshape(p, -1)) for g in range(n))
        del eng, cov
        torch.cuda.empty_cache()
        model = GeneNMFOA(device=dev, **kw)
        model._group = group
        ts = []
        for i in range(2 + max(1, args.steps)):
            barrier()
            t0 = time.perf_counter()
            est = model.run(cov_dict, reads_h)
            torch.cuda.synchronize(dev)
            dt = time.perf_counter() - t0
            if i >= 2:
                ts.append(dt)
        sec = float(np.mean(ts))
        if world > 1:
            t = torch.tensor([sec], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sec = float(t.item())
        h2d = arr.nbytes + reads_h.nbytes + off.nbytes + (ds.nbytes if ds is not None else 0) + 4 * n
        d2h = (model.rho.nbytes + model.x_adj.nbytes + model.x_weighted.nbytes + 2 * 8 * p + n * kw["degnorm_iter"]
               + model.counters.nbytes + (arr.nbytes if est is not None else 0))
        e2e = dict(value=world * n / sec, unit=UNIT, h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h),
                   seconds_per_step=sec, timings=model.timings, estimates_returned=est is not None)

    if rank == 0:
        clock_summary = clocks.summary()
        cb = None
        if not args.no_cpu and world == 1:            # (the CPU baseline is an N = 1 figure)
            cb = run_cpu_baseline(cfg, kw, args.cpu_genes or 2 * cores, cores)
        emit(dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                              ms_per_step=ms_per_step, higher_is_better=True, scaling="weak", vs_baseline=None,
                              dtype="f64", data="synthetic", config=config, e2e=e2e, gpu_launches=launches,
                              roofline=roofline, cpu_baseline=cb, clocks=clock_summary,
                              algorithmic_bytes_per_step=total_bytes, algorithmic_parts=parts))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
