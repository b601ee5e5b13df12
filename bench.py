#!/usr/bin/env python
"""
bench.py -- genes/s of the full DegNorm NMF-OA path (degnorm_iter=5, nmf_iter=100, baseline selection on) on
synthetic coverage of BASELINE.json's shapes.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config c3] [--genes G] [--impl ours|reference]

Default workload: **C3** (BASELINE.json configs[2], the north-star target: 60,000 genes x 48 samples, no
down-sampling) -- a fixed, seeded SAMPLE of G genes of that config (default 4,800; the full 60,000 need ~4 minutes per
step on one GPU, `--genes 60000` runs them) -- **sharded over the N ranks** (strong scaling: the same gene set at every
N, genes assigned by estimated work, one all-reduce of 3p+1 doubles per outer iteration over NCCL).

A "step" is one complete GeneNMFOA flow (init ratio-SVD pass, 5 outer iterations of fused baseline selection +
scale-factor update, estimates of the last iteration) over the gene set:
  * `value`  : device-resident -- every rank's packed coverage shard is already in HBM when the timed region starts;
  * `e2e`    : the same flow through the drop-in GeneNMFOA.run() with HOST numpy inputs (separately allocated
               matrices -> packed into pinned staging -> device), results (DI, adjusted counts, scale factors, flags,
               counters) read back, estimates materialised on demand (`return_estimates='lazy'`, what the CLI flow
               needs: it touches a handful of genes); `e2e.variants` adds the eager-estimates flows;
  * `roofline`: the fused baseline-selection kernel (one launch group per outer iteration): algorithmic bytes
               (raw-coverage scan + (24T+24) p L' per nmf() call) / CUDA-event time of the group on the
               launching stream / measured HBM copy bandwidth; for p > 48 the bound is the FP64 pipe and the figure
               is algorithmic DFMA flops against the DFMA peak measured live (`fp64_peak`);
  * `cpu_baseline`: the oracle port of the reference path (scipy svds rank-one step, exactly the reference's
               third-party call) on a bounded uniform sample of the same workload on the host cores.

With N > 1 (torchrun, one process per GPU) rank r owns shard r of the SAME gene set.
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
# genes of each config a default run takes (full size: pass --genes).  C3's 4,800 keep the driver's 25-step run of
# one GPU inside its time limit.
DEFAULT_GENES = dict(c1=1000, c2=20000, c3=4800, c4=160, c5=296)


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
    ap.add_argument("--config", default="c3")
    ap.add_argument("--genes", type=int, default=0, help="genes of the config to run (0: the default sample size)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: one gene set split over the ranks; weak: every rank its own gene set")
    ap.add_argument("--cpu-genes", type=int, default=0, help="genes in the CPU sample (default: 2 per core)")
    ap.add_argument("--cpu-full", action="store_true", help="CPU arm: run all outer iterations instead of one x 5")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-variants", action="store_true", help="e2e: only the headline flow")
    ap.add_argument("--variants", action="store_true",
                    help="e2e: also time the eager-estimates and zero-copy-input flows (default: only in short runs, steps <= 5)")
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


def _worker_init_pass(g):
    orc = _W["orc"]
    t0 = time.perf_counter()
    F = _W["mats"][g]
    est = orc.ratio_svd(F, "svds").sum(axis=1)
    cov = F.sum(axis=1)
    return g, est, cov, time.perf_counter() - t0


def _worker_bs(args):
    orc = _W["orc"]
    g, scale, off = args
    t0 = time.perf_counter()
    r_, _, f_ = orc.baseline_selection((_W["mats"][g].T / scale).T, _W["prm"], int(off), {})
    return g, r_, f_, time.perf_counter() - t0


def cpu_reference_run(mats, reads, kw, cores, full):
    """The reference's flow on these genes with a fork pool over genes (the MPI decomposition of nmf_mpi.py:603-629
    without MPI; genes are handed out longest first so that the cores stay busy).  full: all outer iterations;
    otherwise the init pass and ONE outer iteration are run and timed (every outer iteration repeats the same
    per-gene work, cf. the nmf() call counts of tests/golden/seed_p48.npz).  Returns
    (per-gene seconds of the init pass [n], per-gene seconds of an outer iteration [n], wall seconds, iterations run)."""
    import multiprocessing as mp
    from oracle import nmfoa_oracle as orc
    prm = orc.Params(rank1="svds", **kw)
    n = len(mats)
    cores = max(1, min(cores, n))
    order = list(np.argsort([-m.shape[1] for m in mats]))
    ctx = mp.get_context("fork")
    t_wall = time.perf_counter()
    with ctx.Pool(cores, initializer=_worker_init, initargs=(mats, kw)) as pool:
        p = mats[0].shape[0]
        est, cov, t_init = np.zeros((n, p)), np.zeros((n, p)), np.zeros(n)
        for g, e, c, dt in pool.imap_unordered(_worker_init_pass, order, chunksize=1):
            est[g], cov[g], t_init[g] = e, c, dt
        rho0 = 1.0 - cov / (est + 1.0)
        low = rho0.max(axis=1) < 0.1
        cs_ = reads[low].sum(axis=0) if low.any() else reads.sum(axis=0)
        norm = cs_ / np.median(cs_)
        x_w, scale = reads / norm, norm.copy()
        offs = orc.draw_offsets(n, prm)
        n_run = prm.degnorm_iter if full else 1
        t_iter = np.zeros(n)
        for it in range(n_run):
            rho = np.zeros((n, p))
            for g, r_, _, dt in pool.imap_unordered(_worker_bs, [(g, scale, offs[it, g]) for g in order], chunksize=1):
                rho[g] = r_
                t_iter[g] += dt
            rho = np.clip(rho, 0.0, 0.9)
            _, norm, x_w, scale = orc.outer_update(x_w, rho, scale)
        t_iter /= n_run
    return t_init, t_iter, time.perf_counter() - t_wall, n_run


def cpu_sample(cfg, n_sample, seed):
    """A bounded UNIFORM sample of the workload: n_sample genes with lengths drawn like the config's (long genes
    appear with their natural frequency)."""
    from degnorm_b200.synth import synth_numpy, gene_lengths
    rng = np.random.default_rng(seed)
    lengths = gene_lengths(n_sample, rng, cfg["profile"])
    return synth_numpy(n_sample, cfg["p"], seed + 1, lengths=lengths)


def run_cpu_baseline(cfg, kw, n_sample, cores, full=False, seed_shift=0):
    """genes/s of the host cores on this workload: cores * n / sum over the sample's genes of their single-core
    seconds (init pass + degnorm_iter outer iterations) -- i.e. perfect load balance over the cores, which is what the
    full-size job (thousands of genes per core) approaches; the wall time of a 2-genes-per-core sample is dominated by
    its longest gene and is reported beside it."""
    mats, reads = cpu_sample(cfg, n_sample, cfg["seed"] + 77 + 1000 * seed_shift)
    reads = np.maximum(reads, 1.0)
    t_init, t_iter, wall, n_run = cpu_reference_run(mats, reads, kw, cores, full)
    core_s = float((t_init + kw["degnorm_iter"] * t_iter).sum())
    # how the port compares with the UNMODIFIED reference on the same genes: measured in the build container, where
    # /root/reference is importable (oracle/time_reference.py -> profiles/r02_cpu_reference_<config>.json)
    ratio, ratio_src = None, None
    try:
        rj = json.load(open(os.path.join(ROOT, "profiles", "r02_cpu_reference_%s.json" % cfg["name"])))
        ratio, ratio_src = float(rj["port_over_reference_speed"]), "profiles/r02_cpu_reference_%s.json" % cfg["name"]
    except Exception:
        pass
    return dict(value=cores * n_sample / core_s, unit=UNIT, cores=cores, kind="port",
                port_over_reference_speed=ratio, port_over_reference_source=ratio_src,
                unmodified_reference_estimate=(cores * n_sample / core_s / ratio) if ratio else None,
                core_seconds_per_gene=core_s / n_sample, wall_seconds=wall, outer_iterations_timed=n_run,
                sample="%d genes drawn like %s (%d samples, downsample_rate %d), oracle port with scipy svds (the "
                       "reference's own third-party call), fork pool of %d single-threaded workers over genes; init "
                       "pass + %s timed per gene; value = cores x genes / sum of the genes' core-seconds of the full "
                       "flow (%.1f s wall for this sample)"
                       % (n_sample, cfg["name"], cfg["p"], kw.get("downsample_rate", 1), min(cores, n_sample),
                          "all %d outer iterations" % n_run if full else
                          "ONE outer iteration (x %d: every outer iteration repeats the same per-gene work)"
                          % kw["degnorm_iter"], wall))


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
    full_n = cfg["n_genes"]
    n_total = args.genes or DEFAULT_GENES.get(args.config, full_n)
    cfg["n_genes"] = n_total
    kw = dict(RUN_KW)
    kw["downsample_rate"] = cfg["downsample_rate"]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1
    strong = args.scaling == "strong"
    workload = "%s: %d %sgenes x %d samples, log-normal lengths (%s), downsample_rate=%d, %s" % (
        args.config, n_total, "" if n_total == full_n else "of the config's %d " % full_n, cfg["p"], cfg["profile"],
        cfg["downsample_rate"],
        "one gene set sharded over the ranks by estimated work" if strong else "per GPU")
    config = dict(workload=workload, genes=n_total, samples=cfg["p"], degnorm_iter=kw["degnorm_iter"],
                  nmf_iter=kw["nmf_iter"], baseline_selection=True, downsample_rate=cfg["downsample_rate"],
                  l2="inputs (GBs of coverage) larger than L2")

    if args.impl == "reference":
        if rank != 0:
            return
        n_sample = args.cpu_genes or 2 * cores
        vals, walls, cb = [], [], None
        for k in range(max(1, args.steps)):            # (a CPU run has nothing to warm; every step draws new genes)
            t_step = time.perf_counter()
            cb = run_cpu_baseline(cfg, kw, n_sample, cores, full=args.cpu_full, seed_shift=k)
            walls.append(time.perf_counter() - t_step)
            vals.append(cb["value"])
        v = float(np.mean(vals))
        cb["value"] = v
        if cb.get("port_over_reference_speed"):
            cb["unmodified_reference_estimate"] = v / cb["port_over_reference_speed"]
        cb["sample"] += "; %d steps, each a fresh sample" % len(vals)
        emit(dict(metric=METRIC, value=v, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                  ms_per_step=1000.0 * float(np.mean(walls)), higher_is_better=True,
                  scaling="strong" if strong else "weak",
                  vs_baseline=None, dtype="f64", data="synthetic", config=config, impl="reference", cpu_baseline=cb,
                  ms_per_step_note="wall time of one step = one bounded sample (synthesis + timed flow) on this box; "
                                   "the whole workload at `value` would take ms_full_workload",
                  ms_full_workload=1000.0 * n_total / v,
                  e2e=dict(value=v, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0))
        return

    import torch
    import torch.distributed as dist
    from degnorm_b200 import GeneNMFOA
    from degnorm_b200 import probes
    from degnorm_b200.distributed import balanced_partition
    from degnorm_b200.engine import Params, ShardEngine, draw_offsets
    from degnorm_b200.synth import synth_torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group, numa = None, None
    if world > 1:
        # one rank per GPU: stay on the GPU's NUMA node (pinned staging buffers are first-touched there), best effort
        from degnorm_b200.distributed import pin_to_gpu_numa_node
        numa = pin_to_gpu_numa_node(local)
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD

    # ---- synthetic gene set, generated on the device
    p = cfg["p"]
    prm = Params(**kw)
    lengths_all = config_lengths(args.config, n_total)
    if args.max_len:
        lengths_all = np.minimum(lengths_all, args.max_len)
    if strong:
        # every rank draws the SAME gene set from the config's seed and keeps its shard (genes to ranks by estimated
        # work, p x candidate columns: distributed.balanced_partition, the split run_gene_nmfoa_mpi makes)
        cand = (lengths_all + prm.downsample_rate - 1) // prm.downsample_rate
        shards = balanced_partition(p * cand, world)
        mine = shards[rank]
        if world == 1:
            cov, off, reads = synth_torch(lengths_all, p, cfg["seed"], dev)
        elif n_total > 20000:
            # full-size runs: every rank draws only ITS genes (same lengths as the one-GPU run, own random stream --
            # drawing all 58 GB on every rank would cost minutes of GPU time per rank for nothing)
            cov, off, reads = synth_torch(lengths_all[mine], p, cfg["seed"] + 7919 * (rank + 1), dev)
        else:
            cov_all, off_all, reads_all = synth_torch(lengths_all, p, cfg["seed"], dev)
            lengths = lengths_all[mine]
            off = np.zeros(len(mine) + 1, dtype=np.int64)
            np.cumsum(lengths, out=off[1:])
            cov = torch.empty(p * int(off[-1]), dtype=torch.float64, device=dev)
            for k, g in enumerate(mine):
                cov[p * int(off[k]):p * int(off[k + 1])].copy_(cov_all[p * int(off_all[g]):p * int(off_all[g + 1])])
            reads = reads_all[torch.as_tensor(mine, device=dev)].contiguous()
            del cov_all, reads_all
            torch.cuda.empty_cache()
        ds_all = draw_offsets(n_total, prm)
        ds = None if ds_all is None else np.ascontiguousarray(ds_all[:, mine])
        n = len(mine)
    else:
        lengths = lengths_all
        if world > 1:
            lengths = np.random.default_rng(cfg["seed"] + 1000 * rank).permutation(lengths)
        cov, off, reads = synth_torch(lengths, p, cfg["seed"] + 1000 * rank, dev)
        ds = draw_offsets(n_total, prm)
        n = n_total
    reads = torch.clamp(reads, min=1.0)
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

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

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
    ms_local = e0.elapsed_time(e1)
    launches = eng.launches * args.steps
    ms = max_over_ranks(ms_local)
    ms_per_step = ms / args.steps
    genes_all_ranks = n_total if strong else world * n_total
    value = genes_all_ranks / (ms_per_step / 1000.0)

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
    fp64 = probes.fp64_peak(dev) if rank == 0 else None
    # DRAM bytes of the launch group: from the committed ncu pass of this very command (profiles/), else null
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic_%s.json" % args.config)))
        if int(tj.get("genes", -1)) == n_total and world == 1 and not (args.tiers or args.force_cluster or args.force_streamed
                                                                      or args.mid_clusters or args.mid_warps or args.max_len):
            traffic, traffic_src = float(tj["dram_bytes_per_launch_group"]), "profiles/r02_traffic_%s.json (ncu, same command)" % args.config
    except Exception:
        pass
    resident_fraction = float((cnt[-1, :, 7] & 1).mean()) if n else 0.0
    sum_cols_iter = cnt[:, :, 3].astype(np.float64).sum(axis=1)
    if p <= 12:
        kname, bound = "nmfoa_small_kernel", ("on-chip (shared memory / issue latency): %.0f %% of the genes are shared-memory "
                                              "resident, so `frac` is an as-if-streamed figure, not a physical HBM fraction"
                                              % (100 * resident_fraction) if resident_fraction > 0.5 else "hbm")
    elif p <= 48:
        kname, bound = "nmfoa_mid_kernel", "hbm"
    else:
        kname, bound = "nmfoa_wide_kernel", "fp64"
    roofline = dict(bound=bound, kernel=kname + " (fused baseline selection; one launch group per outer iteration)",
                    achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak, traffic=traffic,
                    traffic_source=traffic_src,
                    peak_source="MEASURED_PEAKS.json (measured copy bandwidth)" if peaks else "fallback 6650 GB/s",
                    algorithmic_bytes_per_launch=float(np.mean(bs_bytes)), ms_per_launch=float(np.mean(bs_ms)),
                    share_of_step=float(np.sum(bs_ms)) / ms_local,
                    note="algorithmic bytes per unit: 24 p (x, lambda read + lambda written) per kept column per inner "
                         "iteration, (24T+24) p L' per nmf() call, 8 p L for the scan (SURVEY 8d); rank 0's launch group",
                    resident_fraction=resident_fraction,
                    nmf_calls_per_gene=float(cnt[:, :, 2].mean()) if n else 0.0,
                    phases_ms_per_step={k: v / args.steps for k, v in all_ms.items()})
    if fp64 is not None:
        # algorithmic DFMAs: the upper Gram triangle p (p + 1) / 2 per column-pass plus 4 p for the multiplier update
        # and the projection (~(p + 6) flops per element), T + 1 passes per nmf() call
        fl = 2.0 * (p * (p + 1) / 2.0 + 4.0 * p) * (kw["nmf_iter"] + 1) * float(np.mean(sum_cols_iter))
        tf = fl / (float(np.mean(bs_ms)) / 1000.0) / 1e12
        roofline["fp64"] = dict(achieved_tflops=tf, peak_tflops=fp64["tflops"], frac=tf / fp64["tflops"],
                                dfma_per_clk_per_sm=fp64["dfma_per_clk_per_sm"],
                                peak_source="dn_probe_fp64 (16 independent DFMA chains per thread), measured in this run")
        if bound == "fp64":
            roofline.update(achieved=tf, peak=fp64["tflops"], unit="TFLOP/s", frac=tf / fp64["tflops"],
                            hbm=dict(achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak))
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
    # per rank: genes, work, time inside the fused launches and time spent waiting for the slowest rank at the
    # all-reduce that ends every outer iteration (the `pre_bs` phase holds the wait)
    mine_stats = [float(n), float(sum_cols_iter.mean()), float(np.sum(bs_ms)) / args.steps,
                  all_ms.get("pre_bs", 0.0) / args.steps, ms_local / args.steps]
    if world > 1:
        t = torch.tensor(mine_stats, dtype=torch.float64, device=dev)
        gathered = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)
        rank_stats = [g.cpu().tolist() for g in gathered]
    else:
        rank_stats = [mine_stats]
    ranks = [dict(genes=int(r[0]), kept_columns_per_iteration=r[1], bs_ms_per_step=r[2], wait_ms_per_step=r[3],
                  ms_per_step=r[4]) for r in rank_stats]
    if rank == 0:
        clocks.stop_flag = True

    # ---- end to end through the drop-in class, host buffers
    e2e = None
    host_bytes = 2 * cov.numel() * 8 * (world if strong else world)
    if not args.no_e2e and not host_memory_allows(2 * host_bytes):
        e2e = dict(value=None, unit=UNIT, skipped="host copies of all ranks' coverage would not fit in host memory")
    elif not args.no_e2e:
        arr = cov.cpu().numpy()
        reads_h = reads.cpu().numpy()
        offs = [int(o) for o in off]
        # separately allocated matrices, as the reference's loaders hand them over (general pack path) ...
        cov_sep = OrderedDict(("g%d" % g, np.array(arr[p * offs[g]:p * offs[g + 1]].reshape(p, -1))) for g in range(n))
        del eng, cov
        torch.cuda.empty_cache()

        def time_flow(cov_dict, return_estimates, reps, warm):
            model = GeneNMFOA(device=dev, return_estimates=return_estimates, **kw)
            model._group = group
            ts, est = [], None
            for i in range(warm + reps):
                barrier()
                t0 = time.perf_counter()
                est = model.run(cov_dict, reads_h)
                torch.cuda.synchronize(dev)
                dt = time.perf_counter() - t0
                if i >= warm:
                    ts.append(dt)
            sec = max_over_ranks(float(np.mean(ts)))
            return sec, model, est

        sec, model, est = time_flow(cov_sep, "lazy", reps=2, warm=1)
        h2d = arr.nbytes + reads_h.nbytes + off.nbytes + (ds.nbytes if ds is not None else 0) + 4 * n
        d2h_small = (model.rho.nbytes + model.x_adj.nbytes + model.x_weighted.nbytes + 2 * 8 * p + n * kw["degnorm_iter"]
                     + model.counters.nbytes + 8 * p * (kw["degnorm_iter"] + 1))
        e2e = dict(value=genes_all_ranks / sec, unit=UNIT, h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h_small),
                   seconds_per_step=sec, timings=model.timings, estimates="lazy (materialised on demand)",
                   input="separately allocated p x L matrices (packed into pinned staging inside the timed region)",
                   flow="GeneNMFOA(return_estimates='lazy').run(cov_dict, reads)")
        del model, est
        if not args.no_variants and (args.variants or args.steps <= 5):
            # (long runs -- the driver's 25 steps -- keep to the headline flow: every variant costs two more full runs)
            variants = {}
            sec_e, model, est = time_flow(cov_sep, True, reps=1, warm=1)
            variants["eager_estimates"] = dict(value=genes_all_ranks / sec_e, seconds_per_step=sec_e,
                                               d2h_bytes_per_step=int(d2h_small + arr.nbytes), timings=model.timings,
                                               note="every gene's p x L estimate copied to the host (the reference's return value)")
            del model, est
            # ... and as back-to-back views of ONE pinned buffer (a loader that reads straight into staging): no repack
            host = torch.from_numpy(arr).pin_memory()
            arr_p = host.numpy()
            cov_view = OrderedDict(("g%d" % g, arr_p[p * offs[g]:p * offs[g + 1]].reshape(p, -1)) for g in range(n))
            sec_z, model, est = time_flow(cov_view, "lazy", reps=1, warm=1)
            variants["lazy_zero_copy_input"] = dict(value=genes_all_ranks / sec_z, seconds_per_step=sec_z, timings=model.timings,
                                                    note="matrices are views of one pinned staging buffer (no host repack)")
            e2e["variants"] = variants

    if rank == 0:
        clock_summary = clocks.summary()
        cb = None
        if not args.no_cpu and world == 1:            # (the CPU baseline is an N = 1 figure)
            cb = run_cpu_baseline(cfg, kw, args.cpu_genes or 2 * cores, cores, full=args.cpu_full)
        emit(dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                  ms_per_step=ms_per_step, higher_is_better=True, scaling="strong" if strong else "weak",
                  vs_baseline=None, dtype="f64", data="synthetic", config=config, e2e=e2e, gpu_launches=launches,
                  roofline=roofline, cpu_baseline=cb, clocks=clock_summary, fp64_peak=fp64, ranks=ranks,
                  algorithmic_bytes_per_step=total_bytes, algorithmic_parts=parts, numa=numa))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
