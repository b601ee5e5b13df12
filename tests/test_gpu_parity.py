"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI
(libdegnorm_b200.so via ctypes); the numpy oracle and the golden fixtures (made by the real reference) are the
checkers.  Tolerances: DI (rho) and adjusted counts 1e-6 absolute as BASELINE.json's north_star states
(adjusted counts are O(1..1e4): 1e-6 relative + 1e-6 absolute), baseline-selection flags and nmf() call
sequences identical."""
from collections import OrderedDict

import numpy as np
import pytest

from conftest import load_case, load_seeded_case, RUN_CASES, SEEDED_CASES, GOLDEN

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


@pytest.mark.parametrize("case", SEEDED_CASES)
def test_reference_fixtures_at_production_settings(case):
    """The kernels of 13..48 samples (mid-p; clusters of 2 and 4 on the 20,000- and 40,000-column genes), of more than
    48 samples, and the long-gene tiers of the small-p kernel (streamed clusters of 4, 8 and 16 on genes of 40,000 to
    300,000 columns; 149,579 kept columns through 17 nmf() calls) against the UNMODIFIED reference at its real
    settings: nmf_iter = 100, 2-3 outer iterations.  Fixtures: oracle/gen_golden.py (seeded cases)."""
    mats, reads, kwargs, ref = load_seeded_case(case)
    m, est = _gpu_run(mats, reads, **kwargs)
    _compare(m, None, ref)
    np.testing.assert_allclose(np.array([e.sum(axis=1) for e in est]), ref["est_rowsum"], rtol=1e-7)
    np.testing.assert_allclose(np.array([e.max(axis=1) for e in est]), ref["est_max"], rtol=1e-6)
    w = ref["nmf_widths"]
    np.testing.assert_array_equal(m.counters[:, :, 2], (w >= 0).sum(axis=2))
    np.testing.assert_array_equal(m.counters[:, :, 3], np.where(w >= 0, w, 0).sum(axis=2))
    clusters = sorted({int(b.plan.cluster) for b in m._engine.buckets})
    if case == "seed_p48_long":
        assert clusters == [1, 2, 4], clusters
    if case == "seed_p12_long":
        assert {4, 8, 16} <= set(clusters), clusters
        assert any(int(b.plan.cluster) > 1 and int(b.plan.resident_cols) == 0 for b in m._engine.buckets)


def test_single_matrix_methods_known_answers():
    """rank_one_approx / nmf / ratio_svd / baseline_selection as single-matrix methods on the device path
    (nmf.py:55-121, 189-372), against the reference's own outputs (tests/golden/kat.npz)."""
    import os
    from degnorm_b200 import GeneNMFOA
    d = np.load(os.path.join(GOLDEN, "kat.npz"))
    x = d["x"]
    m = GeneNMFOA(nmf_iter=100)
    K, E = m.nmf(x, factors=True)
    assert K.shape == (3, 1) and E.shape == (1, 8) and (K >= 0).all() and (E >= 0).all()
    np.testing.assert_allclose(K.ravel(), d["nmf_absK"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(K.dot(E), d["nmf_est"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(m.nmf(x), d["nmf_est"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(m.ratio_svd(x), d["ratio_svd_est"], rtol=0, atol=1e-9)
    K1, E1 = GeneNMFOA.rank_one_approx(x)
    u, s, vt = np.linalg.svd(x, full_matrices=False)
    np.testing.assert_allclose(K1.dot(E1), s[0] * np.outer(u[:, 0], vt[0]), rtol=0, atol=1e-9)
    np.testing.assert_allclose(np.linalg.norm(E1), 1.0, rtol=1e-12)
    with pytest.raises(ValueError):
        GeneNMFOA.rank_one_approx(x[:, :1])                     # svds(k=1) refuses min(shape) < 2 (nmf.py:63)
    # a zero row is legal input for nmf() (svds handles it)
    x0 = x.copy(); x0[1] = 0.0
    Kz, Ez = m.nmf(x0, factors=True)
    assert Kz[1, 0] == 0.0 and np.isfinite(Kz).all()


def test_single_gene_baseline_selection_method():
    """GeneNMFOA.baseline_selection(F) for one gene: unclipped rho, full-length estimate, flag -- against the oracle
    (which the golden fixtures pin), with and without down-sampling (offset drawn from the global numpy stream)."""
    from degnorm_b200 import GeneNMFOA
    from degnorm_b200.synth import synth_numpy
    from oracle import nmfoa_oracle as orc
    mats, _ = synth_numpy(3, 5, 31, lengths=np.array([700, 1500, 90]), jitter=1e-6)
    for rate in (1, 4):
        kw = dict(nmf_iter=60, downsample_rate=rate)
        m = GeneNMFOA(**kw)
        m.p = 5
        prm = orc.Params(rank1="gram", **kw)
        for F in mats:
            np.random.seed(5)
            rho, est, ran = m.baseline_selection(F)
            np.random.seed(5)
            start = int(np.random.choice(rate)) if rate > 1 else 0
            r_ref, e_ref, ran_ref = orc.baseline_selection(F, prm, start, {})
            assert ran == ran_ref
            np.testing.assert_allclose(rho, r_ref, rtol=0, atol=1e-8)
            np.testing.assert_allclose(est, e_ref, rtol=1e-7, atol=1e-7)


def test_run_logs_the_reference_lines(caplog):
    """nmf.py:537-538, 571-572, 592-593: same log lines, same order."""
    import logging
    from degnorm_b200.synth import synth_numpy
    mats, reads = synth_numpy(5, 4, 12, lengths=np.array([300, 420, 700, 256, 512]), jitter=1e-6)
    with caplog.at_level(logging.INFO):
        m, _ = _gpu_run(mats, reads, degnorm_iter=2, nmf_iter=20)
    msgs = [r.getMessage() for r in caplog.records]
    assert msgs[0].startswith("Initial sequencing depth scale factors -- \n\t")
    assert msgs[1] == "DegNorm iteration 1 -- %d genes sent through baseline selection" % m.ran_baseline_selection[:, 0].sum()
    assert msgs[2].startswith("DegNorm iteration 1 -- sequencing depth scale factors: \n\t")
    assert msgs[4].startswith("DegNorm iteration 2 -- sequencing depth scale factors: \n\t")
    assert msgs[4].endswith(", ".join(str(v) for v in m.scale_factors))


def test_non_current_device_with_lazy_estimates():
    """GeneNMFOA(device='cuda:1', return_estimates='lazy') while cuda:0 is current: the estimate launches that
    follow run() must go to the engine's device (skipped on a single-GPU box)."""
    import torch
    from degnorm_b200.synth import synth_numpy
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two devices")
    mats, reads = synth_numpy(6, 4, 9, lengths=np.array([300, 420, 700, 256, 512, 900]), jitter=1e-6)
    torch.cuda.set_device(0)
    lazy, est_lazy = _gpu_run(mats, reads, degnorm_iter=2, nmf_iter=30, device="cuda:1", return_estimates="lazy")
    eager, est = _gpu_run(mats, reads, degnorm_iter=2, nmf_iter=30, device="cuda:0")
    assert torch.cuda.current_device() == 0
    for k in range(len(mats)):
        np.testing.assert_array_equal(est_lazy[k], est[k])
    np.testing.assert_array_equal(lazy.rho, eager.rho)


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
    (17, 5, dict(degnorm_iter=2, nmf_iter=100)),
    (33, 4, dict(degnorm_iter=2, nmf_iter=100)),
    (48, 4, dict(degnorm_iter=2, nmf_iter=100)),
    (70, 3, dict(degnorm_iter=2, nmf_iter=100)),
    (130, 2, dict(degnorm_iter=2, nmf_iter=100)),
    (200, 2, dict(degnorm_iter=2, nmf_iter=100)),
    (256, 2, dict(degnorm_iter=1, nmf_iter=100)),
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


@pytest.mark.parametrize("p,cluster,warps", [(17, 1, 0), (48, 1, 0), (48, 2, 0), (48, 4, 0), (30, 16, 0),
                                             (48, 1, 12), (17, 2, 12), (48, 8, 12), (30, 16, 12),
                                             (48, 1, 4), (17, 2, 4), (30, 16, 4)])
def test_mid_kernel_equals_tiled_kernel(p, cluster, warps):
    """13..48 samples: the streamed mid-p kernel (warps = 0: the default, 8 warps that update and accumulate their own
    columns; 12: the warp-specialised instantiation, 8 Gram warps + 4 update warps; 4: two 4-warp CTAs per SM;
    optionally one cluster per gene) against the generic tiled kernel on the same genes: identical decisions and call
    sequences, DI equal to rounding."""
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
        assert not use_mid or all(int(b.plan.threads) == (256 if warps == 0 else 32 * warps) for b in eng.buckets)
        o = eng.run(None, want_estimates=True)
        torch.cuda.synchronize()
        outs.append({k: v.cpu().numpy() for k, v in o.items() if torch.is_tensor(v)})
    a, b = outs
    np.testing.assert_array_equal(a["ran"], b["ran"])
    np.testing.assert_array_equal(a["counters"][:, :, :4], b["counters"][:, :, :4])
    np.testing.assert_array_equal(a["counters"][:, :, 5:7], b["counters"][:, :, 5:7])
    np.testing.assert_allclose(a["rho"], b["rho"], rtol=0, atol=1e-11)
    np.testing.assert_allclose(a["est"], b["est"], rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("p,cluster", [(49, 1), (64, 1), (100, 1), (200, 1), (208, 1), (200, 2), (130, 4), (72, 16)])
def test_wide_kernel_equals_tiled_kernel(p, cluster):
    """49..208 samples: the wide kernel (one 8 x 8 Gram tile per thread held in registers for a whole pass, TMA ring,
    register-tile eigen mat-vec; k-slices for narrow cohorts; optionally one cluster per gene) against the generic
    tiled kernel on the same genes: identical decisions and call sequences, DI equal to rounding."""
    import torch
    from degnorm_b200.engine import Params, ShardEngine
    from degnorm_b200.packing import pack_coverage
    from degnorm_b200.synth import synth_numpy
    lengths = np.array([300, 520, 1210, 150, 95, 33])
    mats, reads = synth_numpy(len(lengths), p, 500 + p, lengths=lengths, jitter=1e-6)
    prm = Params(degnorm_iter=2, nmf_iter=12)
    flat, off = pack_coverage(mats, p)
    outs = []
    for use_wide in (False, True):
        eng = ShardEngine(prm, p, "cuda:0")
        eng.use_wide = use_wide
        eng.force_cluster = cluster if (use_wide and cluster > 1) else 0
        eng.load(flat.cuda(), off, torch.from_numpy(reads).cuda())
        assert all((int(b.plan.threads) == 384) == use_wide for b in eng.buckets)
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


def _engine_outputs(prm, p, flat, off, reads, ds):
    import torch
    from degnorm_b200.engine import ShardEngine
    eng = ShardEngine(prm, p, "cuda:0")
    eng.load(flat, off, reads)
    o = eng.run(ds, want_estimates=False)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy().copy() for k, v in o.items() if torch.is_tensor(v)}


def _equivariant(a, c, rows, per_iter, scale_want, lengths, tag, well, tol_clean=1e-9, tol_flipped=1e-5, max_flips=0):
    """Outputs `c` of a run on re-ordered inputs against the re-ordered outputs `a` of the original run, on the
    well-conditioned genes `well` (boolean, in c's gene order).

    Why only those ("ill-conditioned genes"): a gene with a handful of reads per sample has no
    dominant rank-one component (top Gram eigenvalues like 22.8, 20.3, 17.9), the 100 NMF-OA iterations then amplify
    rounding noise to O(0.1) in DI, and the exact test `min(rowsum(KE)) == 0` (nmf.py:314) on a numerically-zero
    factor entry is a coin toss -- the reference does not reproduce ITSELF on such genes between its two
    mathematically identical rank-one routes, nor under one ulp of scale-factor noise.  A re-ordering changes the
    order in which per-sample sums are added, i.e. exactly such noise.

    On the well-conditioned genes a decision (flags, nmf() call sequence, dropped bins) may change in at most
    `max_flips` gene-iterations (a flipped gene elsewhere moves every gene's scale factors by ~1e-6, enough to turn
    a decision that sits within 1e-6 of its threshold); all the others must agree to tol_clean when no gene at all
    flipped, else to tol_flipped."""
    import os
    decided = [0, 1, 2, 3, 5, 6]                            # exit, kept columns, nmf() calls, their widths, dropped bins
    same = (c["ran"] == per_iter(a["ran"])) & \
        (c["counters"][:, :, decided] == per_iter(a["counters"])[:, :, decided]).all(axis=2)
    flips = np.argwhere(~same)                              # (iteration, gene position in c)
    flipped = np.unique(flips[:, 1])
    well_flips = [(it, g) for it, g in flips if well[g]]
    keep = np.setdiff1d(np.flatnonzero(well), flipped)
    tol = tol_clean if len(flips) == 0 else tol_flipped
    worst = np.abs(c["rho"][keep] - rows(a["rho"])[keep]).max(axis=1)
    if os.path.isdir("gpurun_out") and (len(flips) or (worst > tol).any()):
        with open(os.path.join("gpurun_out", "equivariance_%s.txt" % tag), "w") as f:
            f.write("%d flips (%d on well-conditioned genes), %d well-conditioned genes beyond %.0e\n" % (
                len(flips), len(well_flips), int((worst > tol).sum()), tol))
            for it, g in flips:
                f.write("iter %d gene@%d L=%d well=%d counters here %s there %s rho diff %.3e\n" % (
                    it, g, lengths[g], well[g], c["counters"][it, g, :7].tolist(),
                    per_iter(a["counters"])[it, g, :7].tolist(), np.abs(c["rho"][g] - rows(a["rho"])[g]).max()))
            for k in keep[worst > tol]:
                f.write("gene@%d L=%d diff %.3e\n  rho here  %s\n  rho there %s\n" % (
                    k, lengths[k], np.abs(c["rho"][k] - rows(a["rho"])[k]).max(), np.round(c["rho"][k], 5).tolist(),
                    np.round(rows(a["rho"])[k], 5).tolist()))
    assert len(well_flips) <= max_flips, "%d gene-iterations of well-conditioned genes decided differently" % len(well_flips)
    assert len(flips) <= 2e-3 * same.size, "%d of %d gene-iterations decided differently" % (len(flips), same.size)
    np.testing.assert_allclose(c["rho"][keep], rows(a["rho"])[keep], rtol=0, atol=tol)
    np.testing.assert_allclose(c["x_adj"][keep], rows(a["x_adj"])[keep], rtol=tol, atol=tol)
    np.testing.assert_allclose(c["scale_factors"], scale_want, rtol=1e-12 if len(flips) == 0 else tol_flipped, atol=0)
    return flips


def test_full_size_c2_size_independent_properties_and_oracle_spot_check():
    """BASELINE.json configs[1] at its full size (20,000 genes x 12 samples, take-every 20, 5 x 100 iterations),
    which the oracle cannot finish: size-independent properties of the path instead (see _size_independent)."""
    _size_independent("c2", None, 24, 15, 10)


def test_c3_shaped_size_independent_properties_and_oracle_spot_check():
    """BASELINE.json configs[2], the north-star shape (48 samples, no down-sampling, 5 x 100 iterations; the mid-p
    kernel, single CTAs and clusters): 1,500 genes of it through the same size-independent properties and an oracle
    spot check on 12 length-stratified genes."""
    _size_independent("c3", 1500, 12, 7, 5)


def _size_independent(config, n_genes, n_oracle, min_well_posed, min_di_checked):
    """A workload the oracle cannot finish: size-independent properties of the path instead --
      (1) run-to-run determinism (bitwise);
      (2) identities of the outer update (nmf.py:575-590): x_adj (1 - rho) = x_weighted norm, scale = scale_used
          norm, median(norm) = 1, 0 <= rho <= 0.9, flags only where baseline selection ran;
      (3) gene-order equivariance: permuting the genes (with their down-sampling offsets) permutes the rows --
          exactly after one outer iteration (the initial scale factors are sums of integer counts, so every gene
          sees bit-identical inputs: no decision may change, DI to 1e-12), and up to the rare noise-decided genes
          (see _equivariant) after all five;
      (4) sample-order equivariance after one outer iteration: permuting the samples permutes the columns;
      (5) the oracle on a length-stratified sample of genes, fed the scale factors the GPU used in the last outer
          iteration: DI within 1e-6 and identical flags for every gene; identical nmf() call counts and factorised
          widths for every gene whose decisions are well-posed.
    Counts carry a 1e-6 relative jitter so that the high-coverage threshold has no exact ties."""
    import torch
    from degnorm_b200.engine import Params, draw_offsets
    from degnorm_b200.synth import CONFIGS, config_lengths, synth_torch
    from oracle import nmfoa_oracle as orc
    cfg = CONFIGS[config]
    n, p, rate = (n_genes or cfg["n_genes"]), cfg["p"], cfg["downsample_rate"]
    lengths = config_lengths(config, n)
    flat, off, reads = synth_torch(lengths, p, cfg["seed"], "cuda:0")
    gen = torch.Generator(device="cuda:0")
    gen.manual_seed(7)
    flat.mul_(1.0 + 1.0e-6 * torch.rand(flat.numel(), generator=gen, device="cuda:0", dtype=torch.float64))
    prm = Params(downsample_rate=rate)
    prm1 = Params(downsample_rate=rate, degnorm_iter=1)
    assert prm.degnorm_iter == 5 and prm.nmf_iter == 100
    ds = draw_offsets(n, prm)
    if ds is None:                                           # (no down-sampling: every start offset is 0)
        ds = np.zeros((prm.degnorm_iter, n), dtype=np.int32)
    ds_arg = (lambda d: d) if rate > 1 else (lambda d: None)
    a = _engine_outputs(prm, p, flat, off, reads, ds_arg(ds))
    a1 = _engine_outputs(prm1, p, flat, off, reads, ds_arg(ds[:1]))
    # well-conditioned genes: at least one count per sample and position on average (see _equivariant)
    csum = torch.cumsum(flat, 0)
    ends = torch.as_tensor(p * off[1:] - 1, device="cuda:0")
    tot = csum[ends].cpu().numpy()
    mean_cov = np.diff(np.concatenate(([0.0], tot))) / (p * lengths)
    well = mean_cov >= 1.0
    del csum
    assert 0.5 < well.mean() < 0.99

    # (1) determinism
    b = _engine_outputs(prm, p, flat, off, reads, ds_arg(ds))
    for k in ("rho", "x_adj", "x_weighted", "scale_factors", "ran", "counters"):
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)
    del b

    # (2) identities and ranges
    rho, x_adj, x_w, norm, scale = a["rho"], a["x_adj"], a["x_weighted"], a["norm_factors"], a["scale_factors"]
    assert rho.shape == (n, p) and rho.min() >= 0.0 and rho.max() <= 0.9
    np.testing.assert_allclose(x_adj * (1.0 - rho), x_w * norm[None, :], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(scale, a["scale_used"] * norm, rtol=1e-14, atol=0)
    assert abs(np.median(norm) - 1.0) < 1e-12
    exits = a["counters"][:, :, 0]
    default_exit = (exits >= 1) & (exits <= 3)
    assert not a["ran"][default_exit].any()
    assert (a["counters"][:, :, 1] <= (lengths[None, :] + rate - 1) // rate).all()
    assert a["ran"].any() and (exits == 5).any() and default_exit.any()          # the workload exercises the paths

    # (3) gene-order equivariance
    rng = np.random.default_rng(11)
    perm = rng.permutation(n)
    off_p = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lengths[perm], out=off_p[1:])
    flat_p = torch.empty_like(flat)
    for k, g in enumerate(perm):
        flat_p[p * off_p[k]: p * off_p[k + 1]] = flat[p * off[g]: p * off[g + 1]]
    reads_p = reads[torch.as_tensor(perm, device="cuda:0")].contiguous()
    ds_p = np.ascontiguousarray(ds[:, perm])
    c1 = _engine_outputs(prm1, p, flat_p, off_p, reads_p, ds_arg(ds_p[:1]))
    flips = _equivariant(a1, c1, lambda m: m[perm], lambda m: m[:, perm], a1["scale_factors"], lengths[perm], "genes_1iter",
                         np.ones(n, dtype=bool), tol_clean=1e-12)
    assert len(flips) == 0                                 # every gene, the ill-conditioned ones included
    c = _engine_outputs(prm, p, flat_p, off_p, reads_p, ds_arg(ds_p))
    del flat_p
    _equivariant(a, c, lambda m: m[perm], lambda m: m[:, perm], scale, lengths[perm], "genes", well[perm], max_flips=10)
    del c

    # (4) sample-order equivariance (one outer iteration)
    sp = rng.permutation(p)
    flat_s = torch.empty_like(flat)
    sp_dev = torch.as_tensor(sp, device="cuda:0")
    for g in range(n):
        L = int(lengths[g])
        flat_s[p * off[g]: p * off[g + 1]].view(p, L).copy_(flat[p * off[g]: p * off[g + 1]].view(p, L)[sp_dev])
    d1 = _engine_outputs(prm1, p, flat_s, off, reads[:, sp_dev].contiguous(), ds_arg(ds[:1]))
    del flat_s
    _equivariant(a1, d1, lambda m: m[:, sp], lambda m: m, a1["scale_factors"][sp], lengths, "samples_1iter", well,
                 max_flips=2)

    # (5) oracle spot check of the last outer iteration on a length-stratified sample
    order = np.argsort(lengths, kind="stable")
    pick = order[np.linspace(0, n - 1, n_oracle).astype(int)]
    if p > 12:
        pick = pick[lengths[pick] <= 12000]                  # (the numpy oracle needs minutes on longer 48-sample genes)
    oprm = orc.Params(rank1="gram", downsample_rate=rate)
    last = prm.degnorm_iter - 1
    di_checked = well_posed = 0
    for g in pick:
        L = int(lengths[g])
        F = flat[p * off[g]: p * off[g + 1]].view(p, L).cpu().numpy()
        tr = {}
        r_, _, f_ = orc.baseline_selection(F / a["scale_used"][:, None], oprm, int(ds[last, g]), tr)
        r_ = np.clip(r_, 0.0, 0.9)
        assert int(a["counters"][last, g, 1]) == tr["n_hi"], g
        if well[g] and tr.get("min_rel_K", 1.0) > 1e-9:
            # (ill-conditioned genes: see _equivariant; everywhere else the decisions and the DI must be the oracle's)
            well_posed += 1
            assert bool(a["ran"][last, g]) == bool(f_), g
            assert int(a["counters"][last, g, 2]) == tr["nmf_calls"], g
            assert int(a["counters"][last, g, 3]) == tr["sum_cols"], g
            if r_.max() > 0:                                   # (all-zero rows are replaced by the sample average)
                np.testing.assert_allclose(rho[g], r_, rtol=0, atol=DI_TOL, err_msg="gene %d" % g)
                di_checked += 1
    assert well_posed >= min_well_posed and di_checked >= min_di_checked, (well_posed, di_checked)


@pytest.mark.parametrize("case", ["run_p4", "run_p4_ds", "run_p12"])
def test_lazy_estimates_equal_eager_estimates_and_reference(case, tmp_path):
    """return_estimates='lazy' (SURVEY section 8 row f-2): estimates materialised on demand for the genes that are
    indexed -- bitwise the eager ones, equal to the reference fixture, and save_results writes identical files."""
    import filecmp
    import os
    import pandas as pd
    mats, reads, kwargs, ref = load_case(case)
    m0, est0 = _gpu_run(mats, reads, **kwargs)
    m1, est1 = _gpu_run(mats, reads, return_estimates='lazy', **kwargs)
    np.testing.assert_array_equal(m0.rho, m1.rho)
    assert len(est1) == len(est0) == len(mats)
    some = [len(mats) - 1, 0, 2, 0]
    for k, e in zip(some, est1.fetch(some)):
        np.testing.assert_array_equal(e, est0[k])
    np.testing.assert_array_equal(est1[-1], est0[-1])
    for a, b, c in zip(est1, est0, ref["estimates"]):
        np.testing.assert_array_equal(a, b)
        np.testing.assert_allclose(a, c, rtol=1e-6, atol=1e-6)
    with pytest.raises(IndexError):
        est1[len(mats)]
    manifest = pd.DataFrame({"chr": ["chr%d" % (1 + i % 3) for i in range(len(mats))], "gene": m0.genes})
    for tag, m, est in (("eager", m0, est0), ("lazy", m1, est1)):
        os.makedirs(str(tmp_path / tag))
        m.save_results(est, manifest, output_dir=str(tmp_path / tag))
    cmp = filecmp.dircmp(str(tmp_path / "eager"), str(tmp_path / "lazy"))
    assert not cmp.left_only and not cmp.right_only
    for sub in ["."] + sorted(cmp.common_dirs):
        files = sorted(os.listdir(str(tmp_path / "eager" / sub)))
        files = [f for f in files if os.path.isfile(str(tmp_path / "eager" / sub / f))]
        match, mismatch, errors = filecmp.cmpfiles(str(tmp_path / "eager" / sub), str(tmp_path / "lazy" / sub), files,
                                                   shallow=False)
        assert not mismatch and not errors and len(match) == len(files), (sub, mismatch, errors)


def test_gene_filter_on_the_device_equals_the_host_filter():
    """degnorm_b200.gene_filter: the per-gene maxima from one segmented reduction over the
    packed buffer on the GPU give the same kept set as the host pass, on a C1-shaped workload."""
    import torch
    from degnorm_b200.gene_filter import gene_max_coverage, keep_mask
    from degnorm_b200.synth import config_lengths, synth_torch
    lengths = config_lengths("c1")
    flat, off, _ = synth_torch(lengths, 4, 99, "cuda:0")
    host = flat.cpu()
    mx_dev = gene_max_coverage(flat, off, 4).cpu().numpy()
    want = np.array([host[4 * off[g]: 4 * off[g + 1]].max().item() for g in range(len(lengths))])
    np.testing.assert_array_equal(mx_dev, want)
    for minimax, rate in ((0, 1), (20, 1), (50, 300)):
        np.testing.assert_array_equal(keep_mask(flat, off, 4, minimax, rate), keep_mask(host, off, 4, minimax, rate))
        np.testing.assert_array_equal(keep_mask(flat, off, 4, minimax, rate), ~((want < minimax) | (lengths <= rate)))


def test_resident_coverage_handle_equals_the_dictionary_path():
    """gene_filter.DeviceCoverage: one upload shared by the gene filter and GeneNMFOA.run.  Same outputs as the
    dictionary path (bitwise: the kernels see the same packed buffer) and as the reference fixture; the filtered
    handle equals running on the filtered dictionary."""
    from collections import OrderedDict as OD
    import pandas as pd
    from degnorm_b200 import GeneNMFOA
    from degnorm_b200.gene_filter import DeviceCoverage, filter_genes
    mats, reads, kwargs, ref = load_case("run_p4")
    cov = OD(("g%d" % i, x) for i, x in enumerate(mats))
    m0 = GeneNMFOA(**kwargs)
    est0 = m0.run(cov, reads)
    dc = DeviceCoverage(cov, device="cuda:0")
    m1 = GeneNMFOA(**kwargs)
    est1 = m1.run(dc, reads)
    _compare(m1, est1, ref)
    np.testing.assert_array_equal(m1.rho, m0.rho)
    np.testing.assert_array_equal(m1.x_adj, m0.x_adj)
    for a, b in zip(est1, est0):
        np.testing.assert_array_equal(a, b)
    lazy = GeneNMFOA(return_estimates='lazy', **kwargs).run(dc, reads)
    np.testing.assert_array_equal(lazy[1], est0[1])
    # filter on the device, then run: equals the host filter followed by the dictionary path
    genes_df = pd.DataFrame({"chr": "chr1", "gene": list(cov.keys())})
    reads_df = pd.DataFrame(reads, columns=["s%d" % i for i in range(reads.shape[1])])
    reads_df.insert(0, "gene", list(cov.keys()))
    reads_df.insert(0, "chr", "chr1")
    thr = float(np.median([x.max() for x in mats]))
    dc2, g2, r2 = dc.filter(genes_df, reads_df, minimax_coverage=thr, downsample_rate=1)
    cov_h, g_h, r_h = filter_genes(OD(cov), genes_df, reads_df, minimax_coverage=thr, downsample_rate=1)
    assert 0 < len(dc2) < len(cov) and dc2.keys() == list(cov_h.keys()) and g2.equals(g_h) and r2.equals(r_h)
    cols = list(reads_df.columns[2:])
    ma = GeneNMFOA(**kwargs)
    ea = ma.run(dc2, r2[cols].values)
    mb = GeneNMFOA(**kwargs)
    eb = mb.run(cov_h, r_h[cols].values)
    np.testing.assert_array_equal(ma.rho, mb.rho)
    np.testing.assert_array_equal(ma.ran_baseline_selection, mb.ran_baseline_selection)
    for a, b in zip(ea, eb):
        np.testing.assert_array_equal(a, b)
    assert ma.genes == list(cov_h.keys())
