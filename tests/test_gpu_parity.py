"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI
(libdegnorm_b200.so via ctypes); the numpy oracle and the golden fixtures (made by the real reference) are the
checkers.  Tolerances: DI (rho) and adjusted counts 1e-6 absolute as BASELINE.json's north_star states
(adjusted counts are O(1..1e4): 1e-6 relative + 1e-6 absolute), baseline-selection flags and nmf() call
sequences identical."""
from collections import OrderedDict

import numpy as np
import pytest

from conftest import load_case, RUN_CASES

pytestmark = pytest.mark.gpu

DI_TOL = 1e-6


def _gpu_run(mats, reads, **kwargs):
    from degnorm_b200 import GeneNMFOA
    m = GeneNMFOA(**kwargs)
    cov = OrderedDict(("g%d" % i, x) for i, x in enumerate(mats))
    est = m.run(cov, reads)
    return m, est


def _compare(m, est, ref, est_tol=1e-6):
    np.testing.assert_array_equal(m.ran_baseline_selection, ref["ran"])
    np.testing.assert_allclose(m.rho, ref["rho"], rtol=0, atol=DI_TOL)
    np.testing.assert_allclose(m.x_adj, ref["x_adj"], rtol=1e-6, atol=DI_TOL)
    np.testing.assert_allclose(m.scale_factors, ref["scale_factors"], rtol=1e-8, atol=0)
    np.testing.assert_allclose(m.norm_factors, ref["norm_factors"], rtol=1e-8, atol=0)
    np.testing.assert_allclose(m.x_weighted, ref["x_weighted"], rtol=1e-8, atol=1e-9)
    if est is not None:
        for a, b in zip(est, ref["estimates"]):
            assert a.shape == b.shape
            np.testing.assert_allclose(a, b, rtol=est_tol, atol=est_tol)


@pytest.mark.parametrize("case", RUN_CASES)
def test_golden_reference_cases(case):
    """Fixtures produced by the unmodified reference (oracle/gen_golden.py)."""
    mats, reads, kwargs, ref = load_case(case)
    m, est = _gpu_run(mats, reads, **kwargs)
    _compare(m, est, ref)
    # identical nmf() call sequence: number of calls and total factorised width per gene, every outer iteration
    w = ref["nmf_widths"]                      # [iter, gene, 18], -1 padded
    calls = (w >= 0).sum(axis=2)
    cols = np.where(w >= 0, w, 0).sum(axis=2)
    np.testing.assert_array_equal(m.counters[:, :, 2], calls)
    np.testing.assert_array_equal(m.counters[:, :, 3], cols)


def test_integer_count_ties_are_the_only_mismatches():
    """Integer counts make `max_i F_ij/s_i > 0.1*max(F/s)` (nmf.py:76) an exact tie for columns whose maximum is
    one tenth of the matrix maximum; the reference's own answer then depends on the last bit of its scale factors
    (ours differ from it by ~1e-15 relative).  Genes without such a flip must still match to 1e-6, and every
    flipped gene must really hold a tie."""
    from conftest import TIE_CASE
    mats, reads, kwargs, ref = load_case(TIE_CASE)
    m, est = _gpu_run(mats, reads, **kwargs)
    w = ref["nmf_widths"]
    n_hi_ref = np.where(w[:, :, 0] >= 0, w[:, :, 0], m.counters[:, :, 1])      # default-exit genes make no call
    flipped_at = [np.flatnonzero(m.counters[it, :, 1] != n_hi_ref[it]) for it in range(w.shape[0])]
    early = sum(len(f) for f in flipped_at[:-1])
    last = set(flipped_at[-1].tolist())
    scale = m._engine.out["scale_used"].cpu().numpy()
    for g in last:
        x = mats[g] / scale[:, None]
        thr = 0.1 * x.max()
        gap = np.abs(x.max(axis=0) - thr).min() / thr
        assert gap < 1e-12, "gene %d flipped without a tie (gap %.3e)" % (g, gap)
    keep = np.array([g for g in range(len(mats)) if g not in last])
    tol = DI_TOL if early == 0 else 1e-3
    np.testing.assert_allclose(m.rho[keep], ref["rho"][keep], rtol=0, atol=tol)
    np.testing.assert_array_equal(m.ran_baseline_selection[keep], ref["ran"][keep])
    assert len(last) + early <= 2


def _oracle(mats, reads, **kwargs):
    from oracle import nmfoa_oracle as orc
    prm = orc.Params(rank1="gram", **kwargs)
    return orc.run(mats, reads, prm)


@pytest.mark.parametrize("p,n_genes,kwargs", [
    (2, 6, dict(degnorm_iter=2, nmf_iter=30)),
    (5, 8, dict(degnorm_iter=2, nmf_iter=50)),
    (12, 6, dict(degnorm_iter=2, nmf_iter=40, downsample_rate=4)),
    (17, 5, dict(degnorm_iter=1, nmf_iter=30)),
    (48, 4, dict(degnorm_iter=1, nmf_iter=25)),
    (70, 3, dict(degnorm_iter=1, nmf_iter=10)),
    (130, 2, dict(degnorm_iter=1, nmf_iter=6)),
    (200, 2, dict(degnorm_iter=1, nmf_iter=5)),       # more Gram tiles than threads (multi-set path)
])
def test_against_oracle_across_sample_counts(p, n_genes, kwargs):
    from degnorm_b200.synth import synth_numpy
    rng = np.random.default_rng(p)
    lengths = rng.integers(150, 900, size=n_genes)
    mats, reads = synth_numpy(n_genes, p, 1000 + p, lengths=lengths, fortran_every=2)
    m, est = _gpu_run(mats, reads, **kwargs)
    ref = _oracle(mats, reads, **kwargs)
    _compare(m, est, ref)
    want_calls = np.array([t["nmf_calls"] for t in ref["traces"]])
    np.testing.assert_array_equal(m.counters[-1, :, 2], want_calls)


def test_mixed_lengths_through_the_default_planner_against_oracle():
    """A length mix that exercises the default planner end to end (resident tiers of every warp count, the
    single-CTA streamed bucket, down-sampling offsets): DI, adjusted counts, flags and call counts against the oracle."""
    from degnorm_b200.synth import synth_numpy
    rng = np.random.default_rng(2026)
    lengths = np.concatenate([rng.integers(60, 400, size=10), rng.integers(400, 1800, size=8),
                              rng.integers(1800, 7000, size=5), [9500, 30000]])
    kwargs = dict(degnorm_iter=2, nmf_iter=20, downsample_rate=2)
    mats, reads = synth_numpy(len(lengths), 12, 4242, lengths=lengths, jitter=1e-6)
    m, est = _gpu_run(mats, reads, **kwargs)
    ref = _oracle(mats, reads, **kwargs)
    _compare(m, est, ref)
    np.testing.assert_array_equal(m.counters[-1, :, 2], np.array([t["nmf_calls"] for t in ref["traces"]]))
    kinds = {(int(b.plan.threads), int(b.plan.resident_cols) > 0, int(b.plan.cluster)) for b in m._engine.buckets}
    assert any(not res for _, res, _ in kinds) and len({thr for thr, res, _ in kinds if res}) >= 3, kinds
    assert any(cl > 1 for _, _, cl in kinds), kinds          # the 4,750- and 15,000-column genes get clusters


def test_streamed_tier_equals_resident_tier():
    """Same genes through the shared-memory-resident path and the global-slab path: identical decisions,
    DI equal to rounding."""
    import torch
    from degnorm_b200.engine import Params, ShardEngine
    from degnorm_b200.packing import pack_coverage
    from degnorm_b200.synth import synth_numpy
    mats, reads = synth_numpy(8, 4, 77, lengths=np.array([300, 500, 800, 1200, 260, 640, 900, 410]))
    prm = Params(degnorm_iter=2, nmf_iter=60)
    flat, off = pack_coverage(mats, 4)
    outs = []
    for force in (False, True):
        eng = ShardEngine(prm, 4, "cuda:0", force_streamed=force)
        eng.load(flat.cuda(), off, torch.from_numpy(reads).cuda())
        o = eng.run(None, want_estimates=True)
        torch.cuda.synchronize()
        outs.append({k: v.cpu().numpy() for k, v in o.items() if torch.is_tensor(v)})
    a, b = outs
    assert (a["counters"][:, :, 7] & 1).max() == 1 and (b["counters"][:, :, 7] & 1).max() == 0
    np.testing.assert_array_equal(a["ran"], b["ran"])
    np.testing.assert_array_equal(a["counters"][:, :, :4], b["counters"][:, :, :4])
    np.testing.assert_allclose(a["rho"], b["rho"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(a["est"], b["est"], rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("p,cluster,streamed", [(4, 2, False), (12, 4, False), (12, 8, True), (7, 16, False)])
def test_cluster_path_equals_single_cta_path(p, cluster, streamed):
    """The same genes with one CTA per gene and with a thread-block cluster per gene (columns split over the CTAs'
    shared memory or slabs, partial Gram exchanged through distributed shared memory): identical decisions and
    call sequences, DI equal to rounding."""
    import torch
    from degnorm_b200.engine import Params, ShardEngine
    from degnorm_b200.packing import pack_coverage
    from degnorm_b200.synth import synth_numpy
    lengths = np.array([300, 520, 810, 1250, 260, 640, 2900, 410, 95, 1700])
    mats, reads = synth_numpy(len(lengths), p, 100 + p, lengths=lengths, jitter=1e-6)
    prm = Params(degnorm_iter=2, nmf_iter=40)
    flat, off = pack_coverage(mats, p)
    outs = []
    for cl in (0, cluster):
        eng = ShardEngine(prm, p, "cuda:0")
        eng.force_cluster = cl
        eng.force_streamed = streamed and cl > 0
        eng.load(flat.cuda(), off, torch.from_numpy(reads).cuda())
        if cl:
            assert all(int(b.plan.cluster) == cl for b in eng.buckets)
        o = eng.run(None, want_estimates=True)
        torch.cuda.synchronize()
        outs.append({k: v.cpu().numpy() for k, v in o.items() if torch.is_tensor(v)})
    a, b = outs
    np.testing.assert_array_equal(a["ran"], b["ran"])
    np.testing.assert_array_equal(a["counters"][:, :, :4], b["counters"][:, :, :4])
    np.testing.assert_array_equal(a["counters"][:, :, 5:7], b["counters"][:, :, 5:7])
    np.testing.assert_allclose(a["rho"], b["rho"], rtol=0, atol=1e-11)
    np.testing.assert_allclose(a["est"], b["est"], rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("p,cluster,warps", [(17, 1, 8), (48, 1, 8), (48, 4, 8), (30, 16, 8),
                                             (48, 1, 4), (17, 2, 4), (48, 4, 4), (30, 16, 4)])
def test_mid_kernel_equals_tiled_kernel(p, cluster, warps):
    """13..48 samples: the streamed mid-p kernel (8 warps per CTA, or two 4-warp CTAs per SM; optionally one
    cluster per gene) against the generic tiled kernel on the same genes: identical decisions and call sequences,
    DI equal to rounding."""
    import torch
    from degnorm_b200.engine import Params, ShardEngine
    from degnorm_b200.packing import pack_coverage
    from degnorm_b200.synth import synth_numpy
    lengths = np.array([300, 520, 1810, 150, 260, 2900, 410, 95])
    mats, reads = synth_numpy(len(lengths), p, 300 + p, lengths=lengths, jitter=1e-6)
    prm = Params(degnorm_iter=2, nmf_iter=25)
    flat, off = pack_coverage(mats, p)
    outs = []
    for use_mid in (False, True):
        eng = ShardEngine(prm, p, "cuda:0")
        eng.use_mid = use_mid
        eng.mid_warps = warps
        eng.force_cluster = cluster if (use_mid and cluster > 1) else 0
        eng.load(flat.cuda(), off, torch.from_numpy(reads).cuda())
        assert all((int(b.plan.tile) == 6) == use_mid for b in eng.buckets)
        assert not use_mid or all(int(b.plan.threads) == 32 * warps for b in eng.buckets)
        o = eng.run(None, want_estimates=True)
        torch.cuda.synchronize()
        outs.append({k: v.cpu().numpy() for k, v in o.items() if torch.is_tensor(v)})
    a, b = outs
    np.testing.assert_array_equal(a["ran"], b["ran"])
    np.testing.assert_array_equal(a["counters"][:, :, :4], b["counters"][:, :, :4])
    np.testing.assert_array_equal(a["counters"][:, :, 5:7], b["counters"][:, :, 5:7])
    np.testing.assert_allclose(a["rho"], b["rho"], rtol=0, atol=1e-11)
    np.testing.assert_allclose(a["est"], b["est"], rtol=1e-9, atol=1e-9)


def test_init_pass_matches_kat():
    """ratio_svd known-answer vector (produced by the reference)."""
    import os
    import torch
    from conftest import GOLDEN
    from degnorm_b200.engine import Params, ShardEngine
    from degnorm_b200.packing import pack_coverage
    d = np.load(os.path.join(GOLDEN, "kat.npz"))
    x = d["x"]
    flat, off = pack_coverage([x], 3)
    eng = ShardEngine(Params(degnorm_iter=0), 3, "cuda:0")
    eng.load(flat.cuda(), off, torch.ones((1, 3), dtype=torch.float64).cuda())
    o = eng.run(None, want_estimates=False)
    np.testing.assert_allclose(o["rho0"].cpu().numpy()[0], d["ratio_svd_rho0"], rtol=0, atol=1e-12)


def test_input_errors_match_reference():
    from degnorm_b200 import GeneNMFOA
    mats = [np.ones((3, 40)), np.ones((3, 50))]
    cov = OrderedDict((str(i), m) for i, m in enumerate(mats))
    with pytest.raises(ValueError):
        GeneNMFOA().run(cov, np.ones((3, 3)))                 # gene count mismatch (nmf.py:469-470)
    with pytest.raises(ValueError):
        GeneNMFOA(downsample_rate=45).run(cov, np.ones((2, 3)))   # take-every > a gene (nmf.py:479-481)
    with pytest.raises(ValueError):
        GeneNMFOA().save_results([], None)                    # not fitted (nmf.py:623-624)
