#!/usr/bin/env python
"""
Generate tests/golden/*.npz from the REAL reference (imported read-only from /root/reference).

Run in the build container only (the GPU box has no /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden.py [case ...]

The reference is run unmodified through GeneNMFOA(n_jobs=1).run(...).  Two observation hooks are
installed from the outside (the reference's files are never edited):
  * GeneNMFOA.nmf is wrapped to record the width of every matrix it factorises (-> nmf call lists);
  * for down-sampling cases `degnorm.nmf.svds` is wrapped so that ARPACK's random start vector comes
    from a private generator, not from the legacy global numpy stream that the reference's
    np.random.choice offsets are drawn from (modern scipy couples the two;
    the pinned scipy 0.19.1 does not).  Offsets are then exactly
    RandomState(random_state).choice(rate) per gene per outer iteration, in gene order.

Fixture contents: inputs (flattened coverage + lengths + reads + ctor kwargs) and outputs (rho, x_adj,
scale_factors, norm_factors, x_weighted, ran_baseline_selection, flattened estimates, nmf call widths).
"""
import os
import sys
import time
import logging
from collections import OrderedDict

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")

import warnings
warnings.filterwarnings("ignore")
import degnorm.nmf as ref_nmf                      # noqa: E402  (the real reference)
from degnorm_b200.synth import synth_numpy         # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
J = 1.0e-6          # tie-breaking jitter of the synthetic counts (synth_numpy docstring)
logging.getLogger().setLevel(logging.ERROR)


def flatten(mats):
    return np.concatenate([np.ascontiguousarray(m).ravel() for m in mats])


def run_reference(cov_mats, reads, kwargs, private_svds_rng=False):
    widths = []
    orig_nmf = ref_nmf.GeneNMFOA.nmf
    orig_svds = ref_nmf.svds

    def nmf_spy(self, x, factors=False):
        widths[-1].append(int(x.shape[1]))
        return orig_nmf(self, x, factors=factors)

    orig_bs = ref_nmf.GeneNMFOA.baseline_selection

    def bs_spy(self, F):
        widths.append([])
        return orig_bs(self, F)

    ref_nmf.GeneNMFOA.nmf = nmf_spy
    ref_nmf.GeneNMFOA.baseline_selection = bs_spy
    if private_svds_rng:
        gen = np.random.default_rng(7)
        ref_nmf.svds = lambda x, k=1: orig_svds(x, k=k, rng=gen)
    try:
        model = ref_nmf.GeneNMFOA(n_jobs=1, **kwargs)
        cov = OrderedDict(("gene_%d" % i, m) for i, m in enumerate(cov_mats))
        t0 = time.time()
        est = model.run(cov, reads.copy())
        dt = time.time() - t0
    finally:
        ref_nmf.GeneNMFOA.nmf = orig_nmf
        ref_nmf.GeneNMFOA.baseline_selection = orig_bs
        ref_nmf.svds = orig_svds
    n = len(cov_mats)
    n_iter = model.degnorm_iter
    assert len(widths) == n * n_iter
    # widths -> padded int matrix [iter, gene, 18] (-1 padded)
    w = -np.ones((n_iter, n, 18), dtype=np.int32)
    for k, lst in enumerate(widths):
        w[k // n, k % n, :len(lst)] = lst
    out = dict(rho=model.rho, x_adj=model.x_adj, scale_factors=model.scale_factors,
               norm_factors=model.norm_factors, x_weighted=model.x_weighted,
               ran=model.ran_baseline_selection, est_flat=flatten(est), nmf_widths=w,
               ref_seconds=np.float64(dt))
    return out


def save_case(name, cov_mats, reads, kwargs, **kw):
    out = run_reference(cov_mats, reads, kwargs, **kw)
    lengths = np.array([m.shape[1] for m in cov_mats], dtype=np.int64)
    layout_f = np.array([m.flags.f_contiguous and not m.flags.c_contiguous for m in cov_mats])
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, cov_flat=flatten(cov_mats), lengths=lengths, layout_f=layout_f, reads=reads,
                        p=np.int64(cov_mats[0].shape[0]),
                        kw_keys=np.array(list(kwargs.keys())), kw_vals=np.array([float(v) for v in kwargs.values()]),
                        **out)
    print("%-14s %3d genes  ref %.1fs  -> %s (%.0f KB)" % (name, len(cov_mats), out["ref_seconds"], path,
                                                         os.path.getsize(path) / 1024.0))


def input_digest(cov_mats, reads):
    import hashlib
    h = hashlib.sha256()
    for m in cov_mats:
        h.update(np.ascontiguousarray(m).tobytes())
    h.update(np.ascontiguousarray(reads).tobytes())
    return h.hexdigest()


def save_seeded_case(name, synth_kw, kwargs, **kw):
    """Large cases (p = 17 ... 200, genes of tens of thousands of columns): the fixture keeps the generator
    arguments, a digest of the generated inputs and the reference's n x p outputs (plus the row sums of its
    estimates); tests regenerate the inputs with synth_numpy (tests/conftest.py:load_seeded_case checks the digest)."""
    lengths = np.asarray(synth_kw["lengths"], dtype=np.int64)
    cov_mats, reads = synth_numpy(len(lengths), synth_kw["p"], synth_kw["seed"], lengths=lengths,
                                  fortran_every=synth_kw.get("fortran_every", 0), jitter=synth_kw.get("jitter", 0.0))
    out = run_reference(cov_mats, reads, kwargs, **kw)
    p = cov_mats[0].shape[0]
    ests, pos = [], 0
    for L in lengths:
        ests.append(out["est_flat"][pos:pos + p * int(L)].reshape(p, int(L)))
        pos += p * int(L)
    out["est_rowsum"] = np.array([e.sum(axis=1) for e in ests])
    out["est_max"] = np.array([e.max(axis=1) for e in ests])
    del out["est_flat"]
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, lengths=lengths, p=np.int64(p), seed=np.int64(synth_kw["seed"]),
                        fortran_every=np.int64(synth_kw.get("fortran_every", 0)),
                        jitter=np.float64(synth_kw.get("jitter", 0.0)), digest=np.array(input_digest(cov_mats, reads)),
                        kw_keys=np.array(list(kwargs.keys())), kw_vals=np.array([float(v) for v in kwargs.values()]),
                        **out)
    print("%-14s %3d genes  ref %.1fs  -> %s (%.0f KB)" % (name, len(cov_mats), out["ref_seconds"], path,
                                                         os.path.getsize(path) / 1024.0))


def case_kat():
    """SURVEY.md Appendix B.5 known-answer inputs, through the reference's own nmf() and ratio_svd()."""
    x = np.array([[10, 12, 14, 16, 18, 20, 22, 24], [5, 6, 7, 8, 9, 10, 11, 12], [2, 4, 9, 12, 14, 19, 22, 25]],
                 dtype=np.float64)
    m = ref_nmf.GeneNMFOA(nmf_iter=100)
    K, E = m.nmf(x, factors=True)
    est = K.dot(E)
    rs = est.sum(axis=1)
    rho = 1 - x.sum(axis=1) / (rs + 1)
    rsv = m.ratio_svd(x)
    rs2 = rsv.sum(axis=1)
    rho0 = 1 - x.sum(axis=1) / (rs2 + 1)
    np.savez_compressed(os.path.join(OUT, "kat.npz"), x=x, nmf_est=est, nmf_absK=np.abs(K).ravel(),
                        nmf_rowsum=rs, nmf_rho=rho, ratio_svd_est=rsv, ratio_svd_rowsum=rs2, ratio_svd_rho0=rho0)
    print("kat: rs", rs, "rho", rho, "rho0", rho0)


def lengths_uniform(n, lo, hi, seed):
    return np.random.default_rng(seed).integers(lo, hi, size=n)


def main(which):
    os.makedirs(OUT, exist_ok=True)
    cases = {
        "kat": case_kat,
        "run_p4": lambda: save_case(
            "run_p4", *synth_numpy(10, 4, 101, lengths=lengths_uniform(10, 150, 1300, 1), fortran_every=3, jitter=J),
            dict(degnorm_iter=3, nmf_iter=100)),
        "run_p4_ds": lambda: save_case(
            "run_p4_ds", *synth_numpy(8, 4, 102, lengths=lengths_uniform(8, 600, 4000, 2), jitter=J),
            dict(degnorm_iter=3, nmf_iter=100, downsample_rate=5), private_svds_rng=True),
        # integer counts: exact ties in the high-coverage test (see synth_numpy's docstring)
        "run_p4_ds_ties": lambda: save_case(
            "run_p4_ds_ties", *synth_numpy(8, 4, 102, lengths=lengths_uniform(8, 600, 4000, 2)),
            dict(degnorm_iter=3, nmf_iter=100, downsample_rate=5), private_svds_rng=True),
        "run_p12": lambda: save_case(
            "run_p12", *synth_numpy(6, 12, 103, lengths=lengths_uniform(6, 250, 1000, 3), jitter=J),
            dict(degnorm_iter=2, nmf_iter=100)),
        "run_skip": lambda: save_case(
            "run_skip", *synth_numpy(8, 4, 104, lengths=lengths_uniform(8, 150, 1500, 4), jitter=J),
            dict(degnorm_iter=2, nmf_iter=60, skip_baseline_selection=True)),
        "run_p3_bins": lambda: save_case(
            "run_p3_bins", *synth_numpy(6, 3, 105, lengths=lengths_uniform(6, 220, 900, 5), jitter=J),
            dict(degnorm_iter=2, nmf_iter=40, bins=10, min_high_coverage=30)),
        # ---- seeded cases at the reference's real settings (nmf_iter = 100, >= 2 outer iterations) for the kernels
        # of 13..48 samples (mid-p), > 48 samples, and the long-gene tiers (clusters, streamed slabs)
        "seed_p17": lambda: save_seeded_case(
            "seed_p17", dict(p=17, seed=201, lengths=lengths_uniform(6, 300, 2500, 11), jitter=J, fortran_every=3),
            dict(degnorm_iter=2, nmf_iter=100)),
        "seed_p48": lambda: save_seeded_case(
            "seed_p48", dict(p=48, seed=202, lengths=lengths_uniform(6, 300, 3000, 12), jitter=J, fortran_every=2),
            dict(degnorm_iter=3, nmf_iter=100)),
        "seed_p48_long": lambda: save_seeded_case(
            "seed_p48_long", dict(p=48, seed=203, lengths=np.array([20000, 40000, 900]), jitter=J),
            dict(degnorm_iter=2, nmf_iter=100)),
        "seed_p200": lambda: save_seeded_case(
            "seed_p200", dict(p=200, seed=204, lengths=np.array([260, 520, 900, 1500]), jitter=J),
            dict(degnorm_iter=2, nmf_iter=100)),
        "seed_p100": lambda: save_seeded_case(
            "seed_p100", dict(p=100, seed=206, lengths=np.array([300, 700, 1300]), jitter=J),
            dict(degnorm_iter=2, nmf_iter=100)),
        "seed_p12_long": lambda: save_seeded_case(
            "seed_p12_long", dict(p=12, seed=205, lengths=np.array([40000, 160000, 300000, 2000]), jitter=J),
            dict(degnorm_iter=2, nmf_iter=100)),
    }
    for name in (which or list(cases)):
        cases[name]()


if __name__ == "__main__":
    main(sys.argv[1:])
