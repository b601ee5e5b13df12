"""
Device engine of the NMF-OA path: owns one GPU's shard of genes (ragged CSR coverage buffer + read counts)
and drives the CUDA library (include/degnorm_b200.h) through the whole GeneNMFOA.run flow
(reference: degnorm/nmf.py:483-601; distributed twin: degnorm/nmf_mpi.py:555-863).

PyTorch is used for device memory, streams and (with several GPUs) the NCCL all-reduce of the per-sample
sums; every arithmetic step of the path runs in the library's kernels.  No CPU fallback.
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from ._lib import DnParams, DnPlan, check

# resident tiers (columns) of the tiled kernel (p > 12)
RESIDENT_TIERS = (128, 256, 512, 1024, 2048, 4096, 8192, 16384)
# mid-p kernel (13 <= p <= 48): (cluster size, up to this many candidate columns)
MID_MAX_P = 48
MID_CLUSTERS = ((1, 16384), (2, 32768), (4, 65536), (8, 131072), (16, 1 << 31))   # measured sweep: profiles/r01_mid_cluster_sweep.txt
# wide kernel (49 <= p <= 208): one 8 x 8 Gram tile per thread; a CTA needs ~400 clocks per column-pass at p = 200, so
# genes get a cluster earlier than on the mid-p path
WIDE_MIN_P, WIDE_MAX_P = 49, 208
WIDE_CLUSTERS = ((1, 8192), (2, 16384), (4, 32768), (8, 65536), (16, 1 << 31))
# small-p kernel (p <= 12): (columns, warps per CTA); the column caps make whole numbers of CTAs fill an SM's 227 KB
SMALL_TIERS = ((36, 1), (64, 1), (96, 1), (154, 2), (204, 2), (284, 2), (420, 4), (856, 8))


class Params(object):
    """Constructor arguments normalised exactly as GeneNMFOA.__init__ does (nmf.py:30-53)."""

    def __init__(self, degnorm_iter=5, downsample_rate=1, min_high_coverage=50, nmf_iter=100, bins=20, n_jobs=1,
                 skip_baseline_selection=False, random_state=123):
        self.degnorm_iter = abs(int(degnorm_iter))
        self.nmf_iter = abs(int(nmf_iter))
        self.n_jobs = abs(int(n_jobs))
        self.bins = abs(int(bins))
        self.min_high_coverage = max(2, abs(int(min_high_coverage)))
        self.min_bins = int(math.ceil(self.bins * 0.2))
        self.downsample_rate = abs(int(downsample_rate))
        self.skip_baseline_selection = bool(skip_baseline_selection)
        self.random_state = random_state
        if self.downsample_rate > 1:
            self.min_high_coverage = 2
        if self.downsample_rate < 1:
            raise ValueError("downsample_rate must be >= 1")

    def to_c(self, p, flags=0):
        # nmf.py:261 evaluates max(2, ceil(200.0 * (1 / rate))) in floating point; do the same
        min_len = int(max(2, np.ceil(200.0 * (1 / self.downsample_rate))))
        return DnParams(int(p), self.nmf_iter, self.bins, self.min_bins, self.min_high_coverage,
                        self.downsample_rate, min_len, int(self.skip_baseline_selection), int(flags))


def draw_offsets(n_genes, prm):
    """Systematic-sample start offsets, one per gene per outer iteration, from the GLOBAL legacy numpy stream
    exactly as the reference consumes it: run() calls np.random.seed(random_state) (nmf.py:556) and every
    baseline_selection call draws np.random.choice(rate) (nmf.py:420-422), genes in order, iteration-major.
    The seeding of the global stream is a side effect the reference has; it is kept."""
    np.random.seed(prm.random_state)
    if prm.downsample_rate <= 1:
        return None
    # np.random.randint(0, r, size=k) consumes the legacy stream exactly like k successive np.random.choice(r)
    # calls (tests/test_host_logic.py checks it), so the draws are vectorised
    out = np.random.randint(0, prm.downsample_rate, size=prm.degnorm_iter * n_genes)
    return out.reshape(prm.degnorm_iter, n_genes).astype(np.int32)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class _Bucket(object):
    __slots__ = ("plan", "order", "n", "ws", "stream", "max_cols", "est_lo", "est_hi")


class ShardEngine(object):
    """One GPU's shard.  load() takes device tensors; run() leaves device tensors in self.out."""

    def __init__(self, prm, p, device, group=None, force_streamed=False, small_tiers=None, use_row_max=True,
                 allreduce=None):
        self.prm = prm
        self.p = int(p)
        self.device = torch.device(device)
        self.group = group
        self.allreduce = allreduce       # callable(tensor) summing in place over the workers (distributed.py)
        self.force_streamed = force_streamed
        self.small_tiers = small_tiers
        self.serial_buckets = False
        self.prioritise = True
        self.use_clusters = True
        self.cluster_min_cols = -1        # -1: default (4096 columns at P = 12); 0: clusters right above the tiers
        self.use_mid = True
        self.use_wide = True
        self.wide_clusters = None
        self.mid_clusters = None
        self.mid_warps = 0                # 0 / 8: 8 warps, one CTA per SM (default); 4: two 4-warp CTAs per SM;
                                          # 12: warp-specialised (8 Gram + 4 update warps), measured no faster
        self.force_cluster = 0
        self.clusters = (2, 4, 8, 16)
        self.stream_clusters = ((4, 65536), (8, 262144))       # (cluster size, up to this many candidate columns)
        self.bucket_starts = {}
        self.use_row_max = use_row_max
        self.cprm = prm.to_c(p)
        with torch.cuda.device(self.device):
            self.sm_count, self.max_smem, self.cc = _lib.device_info()
        self.lib = _lib.lib()
        self.launches = 0
        self.record_events = False       # bench.py: CUDA events around each phase, on the launching stream
        self.events = []

    # ---------------------------------------------------------------------------------------------------------
    def load(self, cov, offsets, reads):
        """cov: 1-D float64 device tensor (ragged CSR buffer); offsets: int64 numpy [n+1]; reads: n x p float64
        device tensor (this shard's rows)."""
        assert cov.dtype == torch.float64 and cov.is_cuda and cov.is_contiguous()
        self.cov = cov
        self.offsets_np = np.ascontiguousarray(offsets, dtype=np.int64)
        self.n = len(self.offsets_np) - 1
        self.lengths = np.diff(self.offsets_np)
        if self.n and int(self.lengths.max()) >= 2 ** 31 - 1:
            raise ValueError("gene longer than 2^31-2 positions")
        self.off_dev = torch.from_numpy(self.offsets_np).to(self.device)
        self.reads = reads.contiguous() if self.n > 0 else torch.zeros((1, self.p), dtype=torch.float64, device=self.device)
        self._plan()

    def _make_plan(self, max_cols, n_work, want_resident, for_init=False, warps=0, cluster=0):
        plan = DnPlan()
        check(self.lib.dn_make_plan(C.byref(self.cprm), int(max_cols), int(n_work), int(want_resident),
                                    int(for_init), int(warps), int(cluster), self.sm_count, self.max_smem,
                                    C.byref(plan)))
        return plan

    def _bucket(self, ids, cand, want_resident, for_init=False, warps=0, cluster=0):
        b = _Bucket()
        order = ids[np.argsort(-cand[ids], kind="stable")]
        b.n = len(order)
        b.max_cols = int(cand[ids].max())
        b.plan = self._make_plan(b.max_cols, b.n, want_resident, for_init, warps, cluster)
        b.order = torch.from_numpy(order.astype(np.int32)).to(self.device)
        b.ws = torch.empty(int(b.plan.ws_bytes), dtype=torch.uint8, device=self.device)
        b.stream = None
        return b

    def _plan(self):
        self._plan_buckets()
        # Buckets run concurrently on their own streams.  The genes that take longest (most columns) must start
        # first or they end up as the tail of the iteration: launch the largest-column bucket first and give it
        # the highest stream priority, so its CTAs are placed before the many small ones fill the SMs.
        self.buckets.sort(key=lambda b: -b.max_cols)
        lo, hi = torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else (0, -5)
        for k, b in enumerate(self.buckets):
            b.stream = torch.cuda.Stream(device=self.device, priority=max(hi, min(lo, hi + k)) if self.prioritise else 0)
        if self.init_bucket is not None:
            self.init_bucket.stream = torch.cuda.Stream(device=self.device)
        # estimate layout in work order (used when the estimates travel to the host bucket by bucket)
        self.est_off = np.zeros(max(self.n, 1), dtype=np.int64)
        pos = 0
        for b in self.buckets:
            ids = b.order.cpu().numpy()
            b.est_lo = pos
            self.est_off[ids] = pos + np.concatenate(([0], np.cumsum(self.lengths[ids])[:-1]))
            pos += int(self.lengths[ids].sum())
            b.est_hi = pos
        self.est_off_dev = torch.from_numpy(self.est_off).to(self.device)

    def _cluster_share_cap(self, cl):
        """Largest per-CTA column share a resident cluster plan accepts (pure host arithmetic)."""
        lo, hi = 8, 4096
        while lo < hi:
            mid = (lo + hi + 1) // 2 // 8 * 8
            if mid <= lo:
                break
            if self._make_plan(mid * cl, 1, -1, cluster=cl).resident_cols > 0:
                lo = mid
            else:
                hi = mid - 8
        return lo

    def _plan_buckets(self):
        n, L, r = self.n, self.lengths, self.prm.downsample_rate
        self.buckets = []
        self.init_bucket = None
        if n == 0:
            return
        # init pass (ratio_svd): every gene, all columns, coverage read in place
        self.init_bucket = self._bucket(np.arange(n), L, 0, for_init=True)
        # baseline selection: bucket by the number of candidate columns, ceil(L / rate)
        cand = (L + r - 1) // r
        if self.force_streamed or self.force_cluster:
            # testing aids: everything through the streamed tier and / or through clusters of a given size
            wide = WIDE_MIN_P <= self.p <= WIDE_MAX_P
            cl = self.force_cluster if (self.p <= MID_MAX_P or wide) else 0
            if (12 < self.p <= MID_MAX_P and not self.use_mid) or (wide and not self.use_wide):
                cl = -1
            self.buckets.append(self._bucket(np.arange(n), cand, 0 if self.force_streamed else -1,
                                             warps=self.mid_warps if 12 < self.p <= MID_MAX_P else 0, cluster=cl))
            return
        left = np.ones(n, dtype=bool)
        prev = 0
        if self.p <= 12:
            # the default tier table is written for P = 12 (240 bytes per column); fewer padded samples hold
            # proportionally more columns in the same shared memory
            P = 4 if self.p <= 4 else (8 if self.p <= 8 else 12)
            grow = 240.0 / (8 * (2 * (P + 2) + 2))
            tiers = self.small_tiers if self.small_tiers is not None else tuple(
                (int(t * grow) // 2 * 2, w) for t, w in SMALL_TIERS)
            if self.cluster_min_cols < 0:
                self.cluster_min_cols = int(4096 * 12 / P)
            for tier, warps in tiers:
                while tier > prev and self._make_plan(tier, 1, tier, warps=warps).resident_cols == 0:
                    tier -= 8                      # cap the tier at what fits beside this warp count's scratch
                sel = np.flatnonzero(left & (cand <= tier) & (cand > prev))
                if len(sel):
                    self.buckets.append(self._bucket(sel, cand, tier, warps=warps))
                    left[sel] = False
                prev = max(prev, tier)
            # genes just beyond the largest single-CTA tier may stay on one CTA, streamed from its slab (L2-resident)
            if self.cluster_min_cols > 0:
                sel = np.flatnonzero(left & (cand <= self.cluster_min_cols))
                if len(sel):
                    self.buckets.append(self._bucket(sel, cand, 0))
                    left[sel] = False
            # longer genes: one thread-block cluster per gene, columns split over its CTAs' shared memory
            for cl in (self.clusters if self.use_clusters else ()):
                cap = cl * self._cluster_share_cap(cl)
                sel = np.flatnonzero(left & (cand <= cap))
                if len(sel):
                    self.buckets.append(self._bucket(sel, cand, -1, cluster=cl))
                    left[sel] = False
            # beyond the largest cluster's shared memory: streamed from per-CTA slabs, still split over a cluster whose
            # size follows the gene (a CTA should stream at least ~16k columns per pass, or the per-iteration Gram
            # exchange, eigen-solve and pipeline refill dominate)
            if self.use_clusters:
                for cl, cap in self.stream_clusters:
                    sel = np.flatnonzero(left & (cand <= cap))
                    if len(sel):
                        self.buckets.append(self._bucket(sel, cand, 0, cluster=cl))
                        left[sel] = False
            rest = np.flatnonzero(left)
            if len(rest):
                self.buckets.append(self._bucket(rest, cand, 0, cluster=(self.clusters[-1] if self.use_clusters else 0)))
            return
        if self.p <= MID_MAX_P and self.use_mid:
            # mid-p kernel (13..48 samples): every gene streams from its slab; cluster size follows gene length
            for cl, cap in (self.mid_clusters or MID_CLUSTERS):
                sel = np.flatnonzero(left & (cand <= cap))
                if len(sel):
                    self.buckets.append(self._bucket(sel, cand, 0, warps=self.mid_warps, cluster=cl))
                    left[sel] = False
            return
        if WIDE_MIN_P <= self.p <= WIDE_MAX_P and self.use_wide:
            # wide kernel (49..208 samples): every gene streams from its slab; cluster size follows gene length
            for cl, cap in (self.wide_clusters or WIDE_CLUSTERS):
                sel = np.flatnonzero(left & (cand <= cap))
                if len(sel):
                    self.buckets.append(self._bucket(sel, cand, 0, cluster=cl))
                    left[sel] = False
            return
        tiled = -1 if self.p <= WIDE_MAX_P else 0         # (13..208 samples: ask the planner for the tiled kernel)
        for tier in RESIDENT_TIERS:
            plan = self._make_plan(tier, 1, tier, cluster=tiled)
            if plan.resident_cols < tier:
                break
            sel = np.flatnonzero(left & (cand <= tier) & (cand > prev))
            if len(sel):
                self.buckets.append(self._bucket(sel, cand, tier, cluster=tiled))
                left[sel] = False
            prev = tier
        rest = np.flatnonzero(left)
        if len(rest):
            # wholly streamed: a small shared-memory footprint lets several CTAs share an SM
            self.buckets.append(self._bucket(rest, cand, 0, cluster=tiled))

    # ---------------------------------------------------------------------------------------------------------
    def _allreduce(self, t):
        if self.allreduce is not None:
            self.allreduce(t)
        elif self.group is not None:
            torch.distributed.all_reduce(t, group=self.group)

    def run(self, ds_offsets=None, want_estimates=True, est_host=None, keep_for_estimates=False):
        """See _run; the engine's device is made current for the duration of the call (the library launches on the
        current device)."""
        with torch.cuda.device(self.device):
            return self._run(ds_offsets, want_estimates, est_host, keep_for_estimates)

    def estimates(self, gene_ids):
        """See _estimates; runs with the engine's device current."""
        with torch.cuda.device(self.device):
            return self._estimates(gene_ids)

    def _run(self, ds_offsets=None, want_estimates=True, est_host=None, keep_for_estimates=False):
        """ds_offsets: int32 numpy [degnorm_iter, n] (this shard's genes) or None.  Results -> self.out.
        est_host: pinned float64 host tensor of p * sum(L) elements: the estimates are then laid out in WORK order
        (bucket by bucket, self.est_off) and each bucket's block is copied to the host as soon as the bucket has
        finished its last outer iteration, overlapping the other buckets' compute.
        keep_for_estimates: no estimate is materialised, but what estimates() needs afterwards is kept (the E row of
        every gene's first fit: 1/p of the coverage bytes)."""
        prm, p, n, dev, lib = self.prm, self.p, self.n, self.device, self.lib
        f64 = dict(dtype=torch.float64, device=dev)
        main = torch.cuda.current_stream(dev)
        # the previous run's outputs go first: at C3's full size the estimates are as large as the coverage (74 GiB),
        # and two generations of them beside it do not fit
        self.out = None
        self._e_first = None
        n_iter = prm.degnorm_iter
        nn = max(n, 1)
        est_rs = torch.zeros((nn, p), **f64)
        cov_rs = torch.zeros((nn, p), **f64)
        rho = torch.zeros((nn, p), **f64)
        rho0 = torch.zeros((nn, p), **f64)
        x_w = torch.zeros((nn, p), **f64)
        x_adj = torch.zeros((nn, p), **f64)
        norm = torch.zeros(p, **f64)
        scale = torch.zeros(p, **f64)
        sums = torch.zeros(3 * p + 1, **f64)
        ran = torch.zeros((max(n_iter, 1), nn), dtype=torch.uint8, device=dev)
        counters = torch.zeros((max(n_iter, 1), nn, _lib.DN_NCOUNTERS), dtype=torch.int32, device=dev)
        init_counters = torch.zeros((nn, _lib.DN_NCOUNTERS), dtype=torch.int32, device=dev)
        sums_ws = torch.empty(int(lib.dn_sums_workspace_bytes(nn, p)), dtype=torch.uint8, device=dev)
        kfac = torch.zeros((nn, p), **f64)
        row_max = torch.zeros((nn, p), **f64) if self.use_row_max else None
        want_e_first = (want_estimates or keep_for_estimates) and prm.downsample_rate == 1 and n > 0
        e_first = torch.zeros(int(self.offsets_np[-1]), **f64) if want_e_first else None
        ds_dev = torch.from_numpy(np.ascontiguousarray(ds_offsets, dtype=np.int32)).to(dev) if ds_offsets is not None else None
        est = torch.empty_like(self.cov) if (want_estimates and n > 0 and n_iter > 0) else None
        overlap_est = est is not None and est_host is not None
        self.launches = 0
        self.events = []
        self.bucket_events = []

        def mark(name):
            if self.record_events:
                e = torch.cuda.Event(enable_timing=True)
                e.record(main)
                self.events.append((name, e))
        mark("start")

        # ---- init: ratio_svd row sums -> rho0, low-DI genes, norm factors (nmf.py:521-535)
        if n > 0:
            b = self.init_bucket
            check(lib.dn_init_ratio_svd(_ptr(self.cov), _ptr(self.off_dev), _ptr(b.order), b.n, C.byref(self.cprm),
                                        C.byref(b.plan), _ptr(est_rs), _ptr(cov_rs), _ptr(row_max), _ptr(init_counters), _ptr(b.ws),
                                        b.ws.numel(), C.c_void_p(main.cuda_stream)))
            check(lib.dn_init_sums(_ptr(est_rs), _ptr(cov_rs), _ptr(self.reads), n, p, _ptr(rho0), _ptr(sums),
                                   _ptr(sums_ws), sums_ws.numel(), C.c_void_p(main.cuda_stream)))
            self.launches += 3
        mark("init")
        self._allreduce(sums)
        check(lib.dn_init_apply(_ptr(sums), _ptr(self.reads), nn if n == 0 else n, p, _ptr(x_w), _ptr(norm), _ptr(scale),
                                C.c_void_p(main.cuda_stream)))
        self.launches += 1
        scale_used = scale.clone()
        # scale factors after the init pass and after every outer iteration (the reference logs them, nmf.py:537, 592)
        scale_hist = torch.zeros((n_iter + 1, p), **f64)
        scale_hist[0].copy_(scale)

        # ---- outer DegNorm iterations (nmf.py:560-596)
        for it in range(n_iter):
            last = it == n_iter - 1
            scale_used.copy_(scale)
            mark("pre_bs%d" % it)
            for k, b in enumerate(self.buckets):
                if self.serial_buckets:              # tuning aid: one bucket at a time, each timed on its own
                    if k > 0:
                        b.stream.wait_stream(self.buckets[k - 1].stream)
                    if self.record_events:
                        ev0 = torch.cuda.Event(enable_timing=True)
                        ev0.record(b.stream if k > 0 else main)
                        self.bucket_starts[(it, k)] = ev0
                b.stream.wait_stream(main)
                check(lib.dn_baseline_selection(
                    _ptr(self.cov), _ptr(self.off_dev), _ptr(b.order), b.n, C.byref(self.cprm), C.byref(b.plan),
                    _ptr(scale), _ptr(ds_dev[it]) if ds_dev is not None else C.c_void_p(0), _ptr(row_max),
                    _ptr(rho), _ptr(ran[it]), _ptr(counters[it]), _ptr(kfac),
                    _ptr(e_first) if (last and e_first is not None) else C.c_void_p(0),
                    _ptr(est) if (overlap_est and last) else C.c_void_p(0),
                    _ptr(self.est_off_dev) if (overlap_est and last) else C.c_void_p(0),
                    _ptr(b.ws), b.ws.numel(), C.c_void_p(b.stream.cuda_stream)))
                self.launches += 1
                if overlap_est and last:
                    # the kernel wrote this bucket's estimates itself (est != NULL): their trip to the host starts as
                    # soon as the bucket's launch has finished, while the other buckets still compute
                    lo, hi = p * int(b.est_lo), p * int(b.est_hi)
                    with torch.cuda.stream(b.stream):
                        est_host[lo:hi].copy_(est[lo:hi], non_blocking=True)
                if self.record_events:
                    ev = torch.cuda.Event(enable_timing=True)
                    ev.record(b.stream)
                    self.bucket_events.append((it, k, ev))
            for b in self.buckets:
                main.wait_stream(b.stream)
            mark("bs%d" % it)
            if n > 0:
                check(lib.dn_outer_sums(_ptr(x_w), _ptr(rho), n, p, _ptr(sums), _ptr(sums_ws), sums_ws.numel(),
                                        C.c_void_p(main.cuda_stream)))
                self.launches += 2
            else:
                sums.zero_()
            self._allreduce(sums)
            check(lib.dn_outer_apply(_ptr(sums), nn if n == 0 else n, p, _ptr(x_w), _ptr(rho), _ptr(x_adj), _ptr(norm),
                                     _ptr(scale), C.c_void_p(main.cuda_stream)))
            self.launches += 1
            scale_hist[it + 1].copy_(scale)

        if want_estimates and n > 0 and n_iter > 0 and not overlap_est:
            b = self.init_bucket
            check(lib.dn_estimates(_ptr(self.cov), _ptr(self.off_dev), _ptr(b.order), b.n, C.byref(self.cprm),
                                   _ptr(scale_used), _ptr(counters[n_iter - 1]), _ptr(kfac), _ptr(e_first),
                                   C.c_void_p(0), _ptr(est), C.c_void_p(main.cuda_stream)))
            self.launches += 1
        mark("end")
        self.out = dict(rho=rho[:n], rho0=rho0[:n], x_adj=x_adj[:n], x_weighted=x_w[:n], norm_factors=norm,
                        scale_factors=scale, ran=ran[:, :n], counters=counters[:, :n], init_counters=init_counters[:n],
                        est=est, kfac=kfac[:n], scale_used=scale_used, est_in_work_order=overlap_est,
                        scale_hist=scale_hist)
        self._e_first = e_first
        self._can_estimate = (want_estimates or keep_for_estimates) and n > 0 and n_iter > 0
        return self.out

    def _estimates(self, gene_ids):
        """Full-length estimates (nmf.py:217, 247, 333-365) of the listed genes of the last run(), materialised on
        demand: one dn_estimates launch over just those genes into a compact buffer.  Returns (device tensor,
        column offsets [len(ids) + 1]); gene k's block is est[p * o[k] : p * o[k + 1]] viewed as p x L."""
        if not getattr(self, "_can_estimate", False):
            raise ValueError("run(want_estimates=True) or run(keep_for_estimates=True) first")
        ids = np.asarray(gene_ids, dtype=np.int64).ravel()
        if len(ids) == 0:
            return torch.empty(0, dtype=torch.float64, device=self.device), np.zeros(1, dtype=np.int64)
        if ids.min() < 0 or ids.max() >= self.n or len(np.unique(ids)) != len(ids):
            raise ValueError("gene ids must be distinct and in [0, n_genes)")
        p, dev, o = self.p, self.device, self.out
        sub = np.zeros(len(ids) + 1, dtype=np.int64)
        np.cumsum(self.lengths[ids], out=sub[1:])
        est_off = np.zeros(self.n, dtype=np.int64)
        est_off[ids] = sub[:-1]
        est = torch.empty(p * int(sub[-1]), dtype=torch.float64, device=dev)
        order = torch.from_numpy(ids.astype(np.int32)).to(dev)
        est_off_dev = torch.from_numpy(est_off).to(dev)
        main = torch.cuda.current_stream(dev)
        last = self.prm.degnorm_iter - 1
        check(self.lib.dn_estimates(_ptr(self.cov), _ptr(self.off_dev), _ptr(order), len(ids), C.byref(self.cprm),
                                    _ptr(o["scale_used"]), _ptr(o["counters"][last]), _ptr(o["kfac"]),
                                    _ptr(self._e_first), _ptr(est_off_dev), _ptr(est), C.c_void_p(main.cuda_stream)))
        self.launches += 1
        return est, sub

    def check_exit_codes(self, counters_host=None):
        """A negative exit code means a gene did not fit the plan of the bucket it was put in (planner or ABI misuse):
        its row then holds the default result, which would silently distort the normalisation -- refuse instead."""
        cnt = counters_host if counters_host is not None else self.out["counters"].cpu().numpy()
        bad = np.argwhere(cnt[..., _lib.CNT_EXIT] < 0)
        if len(bad):
            raise _lib.DegnormCudaError("%d gene-iterations did not fit their launch plan (first: iteration %d, gene %d)"
                                        % (len(bad), int(bad[0][0]), int(bad[0][-1])))

    # ---------------------------------------------------------------------------------------------------------
    def fit_once(self, scale=None, ds_row=None, flags=0, want_estimates=False, clamp_estimates=False):
        """ONE pass of the fused kernel over this shard's genes outside the run() flow: what the single-matrix
        methods of the class need (GeneNMFOA.nmf / rank_one_approx / ratio_svd / baseline_selection,
        nmf.py:55-121, 189-372).  scale: p doubles (default ones: the matrices are taken as they are); flags:
        _lib.DN_FLAG_*.  Returns device tensors rho (n x p), ran (n), counters, kfac = K (n x p), e_first = E of the
        first fit (packed like the coverage, one row per gene) and, on request, the full-length estimates;
        clamp_estimates asks for max(K.E, x) (ratio_svd, nmf.py:118-119) where the fit was kept unclamped."""
        with torch.cuda.device(self.device):
            p, n, dev, lib = self.p, self.n, self.device, self.lib
            f64 = dict(dtype=torch.float64, device=dev)
            cprm = self.prm.to_c(p, flags)
            main = torch.cuda.current_stream(dev)
            scale_t = torch.ones(p, **f64) if scale is None else torch.as_tensor(scale, **f64).contiguous()
            rho = torch.zeros((n, p), **f64)
            kfac = torch.zeros((n, p), **f64)
            ran = torch.zeros(n, dtype=torch.uint8, device=dev)
            counters = torch.zeros((n, _lib.DN_NCOUNTERS), dtype=torch.int32, device=dev)
            e_first = torch.zeros(int(self.offsets_np[-1]), **f64) if self.prm.downsample_rate == 1 else None
            ds_dev = (torch.from_numpy(np.ascontiguousarray(ds_row, dtype=np.int32)).to(dev)
                      if ds_row is not None else None)
            # the matrix maximum behind the high-coverage threshold (nmf.py:76) is found by the kernel itself here
            for b in self.buckets:
                check(lib.dn_baseline_selection(
                    _ptr(self.cov), _ptr(self.off_dev), _ptr(b.order), b.n, C.byref(cprm), C.byref(b.plan),
                    _ptr(scale_t), _ptr(ds_dev), C.c_void_p(0), _ptr(rho), _ptr(ran), _ptr(counters), _ptr(kfac),
                    _ptr(e_first), C.c_void_p(0), C.c_void_p(0), _ptr(b.ws), b.ws.numel(),
                    C.c_void_p(main.cuda_stream)))
                self.launches += 1
            est = None
            if want_estimates:
                cnt = counters
                if clamp_estimates:
                    cnt = counters.clone()
                    kept = cnt[:, _lib.CNT_EXIT] == _lib.DN_EXIT_NO_SELECTION
                    cnt[kept, _lib.CNT_EXIT] = _lib.DN_EXIT_FALLBACK      # same K, E; estimate = max(K.E, x)
                est = torch.empty_like(self.cov)
                b = self.init_bucket
                check(lib.dn_estimates(_ptr(self.cov), _ptr(self.off_dev), _ptr(b.order), b.n, C.byref(cprm),
                                       _ptr(scale_t), _ptr(cnt), _ptr(kfac), _ptr(e_first), C.c_void_p(0), _ptr(est),
                                       C.c_void_p(main.cuda_stream)))
                self.launches += 1
            return dict(rho=rho, ran=ran, counters=counters, kfac=kfac, e_first=e_first, est=est)

    # ---------------------------------------------------------------------------------------------------------
    def phase_ms(self):
        """{phase: milliseconds} from the recorded events (call after a synchronize)."""
        out = {}
        for (n0, e0), (n1, e1) in zip(self.events[:-1], self.events[1:]):
            out[n1] = out.get(n1, 0.0) + e0.elapsed_time(e1)
        return out

    def bucket_ms(self):
        """Per outer iteration, per bucket: ms from the start of the baseline-selection phase to the end of that
        bucket's launch (buckets run concurrently on their own streams)."""
        starts = {int(n[6:]): ev for n, ev in self.events if n.startswith("pre_bs")}
        out = {}
        for it, k, ev in self.bucket_events:
            s = self.bucket_starts.get((it, k), starts[it])
            out.setdefault(it, {})[k] = s.elapsed_time(ev)
        return out

    def bs_bytes_per_iteration(self):
        """Algorithmic bytes of the fused baseline-selection launches of each outer iteration
        (scan of the raw coverage + every nmf() call as if streamed), SURVEY.md section 8(d)."""
        prm, p = self.prm, self.p
        cnt = self.out["counters"].to(torch.int64)
        sum_cols = cnt[:, :, _lib.CNT_SUM_COLS].sum(dim=1).cpu().numpy().astype(np.float64)
        return 8.0 * p * float(self.lengths.sum()) + (24.0 * prm.nmf_iter + 24.0) * p * sum_cols

    def algorithmic_bytes(self):
        """bytes the path would move if every pass were streamed from HBM,
        from the per-gene device counters.  Returns (total, per-part dict)."""
        prm, p = self.prm, self.p
        T = prm.nmf_iter
        Lsum = float(self.lengths.sum())
        cnt = self.out["counters"].to(torch.int64)
        sum_cols = float(cnt[:, :, _lib.CNT_SUM_COLS].sum().item())
        init = 16.0 * p * Lsum
        scan = 8.0 * p * Lsum * prm.degnorm_iter
        nmf = (24.0 * T + 24.0) * p * sum_cols
        est = 8.0 * p * Lsum if self.out["est"] is not None else 0.0
        return init + scan + nmf + est, dict(init=init, scan=scan, nmf=nmf, est=est, sum_cols=sum_cols)
