"""CPU tests of the host side: parameter normalisation, offsets, packing, the C ABI exports (no GPU calls)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def test_params_normalisation_matches_reference_ctor():
    # nmf.py:30-53
    from degnorm_b200.engine import Params
    p = Params(degnorm_iter=-3, downsample_rate=1, min_high_coverage=1, nmf_iter=7.9, bins=12)
    assert (p.degnorm_iter, p.nmf_iter, p.bins, p.min_high_coverage, p.min_bins) == (3, 7, 12, 2, 3)
    p = Params(downsample_rate=20, min_high_coverage=50)
    assert p.min_high_coverage == 2                      # forced when down-sampling (nmf.py:52-53)
    assert p.to_c(12).min_gene_len == 10                 # max(2, ceil(200 * (1/20)))  nmf.py:261
    assert Params(downsample_rate=3).to_c(4).min_gene_len == 67
    assert Params(downsample_rate=500).to_c(4).min_gene_len == 2


def test_offsets_follow_the_global_legacy_stream():
    from degnorm_b200.engine import Params, draw_offsets
    prm = Params(downsample_rate=20, degnorm_iter=3, random_state=123)
    off = draw_offsets(7, prm)
    np.random.seed(123)
    want = np.array([[np.random.choice(20) for _ in range(7)] for _ in range(3)])
    np.testing.assert_array_equal(off, want)
    assert off[0, :6].tolist() == [13, 2, 2, 6, 17, 19]
    # the reference's visible side effect: the global stream is seeded even without down-sampling (nmf.py:556)
    np.random.seed(5)
    assert draw_offsets(3, Params(random_state=99)) is None
    a = np.random.rand()
    np.random.seed(99)
    assert a == np.random.rand()


def test_packing_general_and_zero_copy():
    from degnorm_b200.packing import pack_coverage
    rng = np.random.default_rng(0)
    mats = [rng.random((3, L)) for L in (5, 9, 2)]
    mats[1] = np.asfortranarray(mats[1])
    flat, off = pack_coverage(mats, 3, pin=False)
    assert off.tolist() == [0, 5, 14, 16]
    for g, m in enumerate(mats):
        np.testing.assert_array_equal(flat.numpy()[3 * off[g]:3 * off[g + 1]].reshape(3, -1), m)
    # views of one buffer laid out back to back are used in place
    base = rng.random(3 * 16)
    views = [base[3 * off[g]:3 * off[g + 1]].reshape(3, -1) for g in range(3)]
    flat2, off2 = pack_coverage(views, 3, pin=False)
    assert flat2.numpy().__array_interface__["data"][0] == base.__array_interface__["data"][0]
    # ... but not when the order or layout differs
    flat3, _ = pack_coverage([views[1], views[0]], 3, pin=False)
    assert flat3.numpy().__array_interface__["data"][0] != base.__array_interface__["data"][0]
    np.testing.assert_array_equal(flat3.numpy()[:27].reshape(3, 9), views[1])


def test_library_exports_every_symbol_in_the_header():
    from degnorm_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "degnorm_b200.h")).read()
    declared = set(re.findall(r"\b(dn_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"dn_status", "dn_exit", "dn_counter", "dn_params", "dn_plan"}
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    lib = C.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert _lib.lib().dn_abi_version() == _lib.ABI_VERSION


def test_plan_is_pure_host_arithmetic():
    from degnorm_b200 import _lib
    from degnorm_b200.engine import Params
    lib = _lib.lib()
    plan = _lib.DnPlan()
    # small-p path (p <= 12): tile 0, warps from the tier, a bucket is wholly resident or wholly streamed
    prm = Params(downsample_rate=20).to_c(12)
    assert lib.dn_make_plan(C.byref(prm), 64, 100000, 64, 0, 0, 0, 148, 232448, C.byref(plan)) == 0
    assert plan.tile == 0 and plan.threads == 32 and plan.resident_cols == 64 and plan.ws_cols == 0
    assert plan.smem_bytes < 24 * 1024 and plan.ctas > 8 * 148
    assert lib.dn_make_plan(C.byref(prm), 400, 1000, 448, 0, 4, 0, 148, 232448, C.byref(plan)) == 0
    assert plan.threads == 128 and plan.resident_cols == 400 and 2 * (plan.smem_bytes + 1024) <= 228 * 1024
    assert lib.dn_make_plan(C.byref(prm), 100000, 1000, 0, 0, 0, 0, 148, 232448, C.byref(plan)) == 0
    assert plan.tile == 0 and plan.resident_cols == 0 and plan.ws_cols >= 100000 and plan.threads == 256
    assert lib.dn_make_plan(C.byref(prm), 64, 10, 64, 0, 3, 0, 148, 232448, C.byref(plan)) == _lib.DN_ERR_INVALID
    # the init pass and p > 12 use the tiled kernel
    assert lib.dn_make_plan(C.byref(prm), 5000, 1000, 0, 1, 0, 0, 148, 232448, C.byref(plan)) == 0
    assert plan.tile == 4 and plan.threads == 256
    # 13..48 samples: the streamed mid-p kernel (tile 6), one cluster per long gene; cluster = -1 asks for the tiled one
    prm48 = Params().to_c(48)
    assert lib.dn_make_plan(C.byref(prm48), 100000, 1000, 0, 0, 0, 16, 148, 232448, C.byref(plan)) == 0
    assert plan.tile == 6 and plan.cluster == 16 and plan.resident_cols == 0 and 16 * plan.ws_cols >= 100000
    assert plan.ws_cols % 64 == 0 and plan.ctas % 16 == 0 and plan.smem_bytes <= 232448
    assert lib.dn_make_plan(C.byref(prm48), 100000, 1000, -1, 0, 0, -1, 148, 232448, C.byref(plan)) == 0
    assert plan.tile == 4 and 0 < plan.resident_cols < 100000 and plan.ws_cols >= 100000
    # 49..208 samples: the wide kernel (one 8 x 8 Gram tile per thread, 384 threads, 16-column chunks); cluster = -1
    # asks for the generic tiled kernel, which also takes 209..256 samples
    prm64 = Params().to_c(64)
    assert lib.dn_make_plan(C.byref(prm64), 5000, 1000, 0, 0, 0, 0, 148, 232448, C.byref(plan)) == 0
    assert (plan.tile, plan.threads, plan.chunk_cols, plan.ctas) == (8, 384, 16, 148) and plan.ws_cols == 5008
    assert lib.dn_make_plan(C.byref(prm64), 5000, 1000, 0, 0, 0, -1, 148, 232448, C.byref(plan)) == 0 and plan.tile == 4
    bad = Params().to_c(1)
    assert lib.dn_make_plan(C.byref(bad), 128, 10, 0, 0, 0, 0, 148, 232448, C.byref(plan)) == _lib.DN_ERR_INVALID
    assert b"2 samples" in lib.dn_last_error()
    p200 = Params().to_c(200)
    assert lib.dn_make_plan(C.byref(p200), 3000, 100, 0, 0, 0, 4, 148, 232448, C.byref(plan)) == 0
    assert (plan.tile, plan.threads, plan.cluster, plan.ctas) == (8, 384, 4, 148) and plan.ws_cols == 752
    assert plan.resident_cols == 0 and plan.ws_bytes > 0 and plan.smem_bytes <= 232448
    p208, p209 = Params().to_c(208), Params().to_c(209)
    assert lib.dn_make_plan(C.byref(p208), 3000, 100, 0, 0, 0, 0, 148, 232448, C.byref(plan)) == 0
    assert plan.threads == 384 and plan.smem_bytes <= 232448
    assert lib.dn_make_plan(C.byref(p209), 3000, 100, 0, 0, 0, 0, 148, 232448, C.byref(plan)) == 0
    assert (plan.tile, plan.threads) == (8, 256)
    big = Params().to_c(5000)
    assert lib.dn_make_plan(C.byref(big), 128, 10, 0, 0, 0, 0, 148, 232448, C.byref(plan)) == _lib.DN_ERR_UNSUPPORTED


def test_no_cpu_fallback():
    """The product path refuses to run without a CUDA device instead of falling back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from collections import OrderedDict
    from degnorm_b200 import GeneNMFOA
    with pytest.raises(RuntimeError):
        GeneNMFOA().run(OrderedDict(a=np.ones((3, 60))), np.ones((1, 3)))


def test_product_never_imports_the_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, "degnorm_b200")):
        for f in files:
            if f.endswith(".py") or f.endswith(".cu") or f.endswith(".h"):
                txt = open(os.path.join(root, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f


def test_helper_statics_match_reference_semantics():
    from degnorm_b200 import GeneNMFOA
    assert GeneNMFOA.shift_bins([[0, 1], [4, 5], [6, 7]], 1) == [[0, 1], [2, 3], [4, 5]]     # nmf.py:160-187
    x = np.array([[0., 1., 10.], [0., 0.5, 2.]])
    assert GeneNMFOA.get_high_coverage_idx(x).tolist() == [2]                                   # strict >, nmf.py:76


def test_save_results_writes_the_reference_files(tmp_path):
    """nmf.py:603-711: degradation_index_scores.csv, adjusted_read_counts.csv, ran_baseline_selection.csv with columns
    chr, gene, <samples> / iter_k in cov_dat gene order, and one estimated_coverage_matrices_<chr>.pkl per chromosome
    holding {gene: p x L}.  (Host-side code: runs without a GPU on a hand-filled model.)"""
    import pickle
    import pandas as pd
    from degnorm_b200 import GeneNMFOA
    m = GeneNMFOA(degnorm_iter=2)
    with pytest.raises(ValueError):
        m.save_results([], pd.DataFrame({"chr": [], "gene": []}), output_dir=str(tmp_path))       # not fitted
    m.genes = ["gB", "gA", "gC"]
    m.p, m.n_genes, m.fitted = 2, 3, True
    m.rho = np.array([[0.1, 0.2], [0.0, 0.3], [0.5, 0.9]])
    m.x_adj = np.array([[10.0, 20.0], [1.0, 2.0], [5.0, 6.0]])
    m.ran_baseline_selection = np.array([[True, False], [False, False], [True, True]])
    est = [np.full((2, 3), 1.0), np.full((2, 4), 2.0), np.full((2, 5), 3.0)]
    manifest = pd.DataFrame({"chr": ["chr2", "chr1", "chr2", "chr9"], "gene": ["gA", "gB", "gC", "gZ"]})
    with pytest.raises(IOError):
        m.save_results(est, manifest, output_dir=str(tmp_path / "missing"))
    with pytest.raises(ValueError):
        m.save_results(est, manifest.rename(columns={"chr": "chrom"}), output_dir=str(tmp_path))
    with pytest.raises(ValueError):
        m.save_results(est, manifest, output_dir=str(tmp_path), sample_ids=["only_one"])
    m.save_results(est, manifest, output_dir=str(tmp_path), sample_ids=["s1", "s2"])
    di = pd.read_csv(tmp_path / "degradation_index_scores.csv")
    assert di.columns.tolist() == ["chr", "gene", "s1", "s2"] and di.gene.tolist() == ["gB", "gA", "gC"]
    assert di.chr.tolist() == ["chr1", "chr2", "chr2"]
    np.testing.assert_allclose(di[["s1", "s2"]].values, m.rho)
    adj = pd.read_csv(tmp_path / "adjusted_read_counts.csv")
    np.testing.assert_allclose(adj[["s1", "s2"]].values, m.x_adj)
    ran = pd.read_csv(tmp_path / "ran_baseline_selection.csv")
    assert ran.columns.tolist() == ["chr", "gene", "iter_0", "iter_1"] and ran.iter_0.tolist() == [True, False, True]
    with open(tmp_path / "chr2" / "estimated_coverage_matrices_chr2.pkl", "rb") as f:
        d2 = pickle.load(f)
    assert sorted(d2) == ["gA", "gC"] and d2["gA"].shape == (2, 4) and d2["gC"][0, 0] == 3.0
    with open(tmp_path / "chr1" / "estimated_coverage_matrices_chr1.pkl", "rb") as f:
        assert list(pickle.load(f)) == ["gB"]
    # default sample ids (nmf.py:639-640)
    m.save_results(est, manifest, output_dir=str(tmp_path))
    assert pd.read_csv(tmp_path / "degradation_index_scores.csv").columns.tolist() == ["chr", "gene", "sample_1", "sample_2"]


def test_lazy_estimates_sequence_protocol_without_a_gpu():
    """LazyEstimates over a stand-in engine: distinct sorted ids reach the engine once per
    request, results come back in the order asked (duplicates and negative indices included), slices and iteration
    batch their requests, out-of-range indices raise IndexError like a list."""
    import torch
    from degnorm_b200.nmf import LazyEstimates

    class FakeEngine(object):
        p = 3

        def __init__(self, lengths):
            self.lengths = np.asarray(lengths)
            self.calls = []

        def estimates(self, ids):
            ids = list(ids)
            assert ids == sorted(set(ids))
            self.calls.append(ids)
            o = np.zeros(len(ids) + 1, dtype=np.int64)
            np.cumsum(self.lengths[ids], out=o[1:])
            buf = np.concatenate([np.full(self.p * self.lengths[g], float(g)) for g in ids]) if ids else np.zeros(0)
            return torch.from_numpy(buf), o

    lengths = [5, 2, 7, 3, 4]
    eng = FakeEngine(lengths)
    est = LazyEstimates(eng, lengths, batch_columns=9)
    assert len(est) == 5
    got = est.fetch([3, 0, 3, -1])
    assert eng.calls == [[0, 3, 4]]
    assert [m.shape for m in got] == [(3, 3), (3, 5), (3, 3), (3, 4)]
    assert [float(m[0, 0]) for m in got] == [3.0, 0.0, 3.0, 4.0]
    assert est[2].shape == (3, 7) and float(est[2][1, 1]) == 2.0
    assert [float(m[0, 0]) for m in est[1:4]] == [1.0, 2.0, 3.0]
    eng.calls.clear()
    assert [float(m[0, 0]) for m in est] == [0.0, 1.0, 2.0, 3.0, 4.0]
    assert eng.calls == [[0, 1], [2], [3, 4]]                 # batches of at most 9 columns (a long gene goes alone)
    with pytest.raises(IndexError):
        est[5]
    with pytest.raises(IndexError):
        est.fetch([-6])


def _filter_case(seed=0):
    import pandas as pd
    from collections import OrderedDict
    rng = np.random.default_rng(seed)
    lengths = [40, 7, 120, 20, 300, 21, 64]
    peak = [30, 500, 4, 80, 9, 12, 10]
    buf = np.zeros(3 * sum(lengths))
    cov, at = OrderedDict(), 0
    for k, (L, m) in enumerate(zip(lengths, peak)):
        v = buf[at:at + 3 * L].reshape(3, L)
        at += 3 * L
        v[...] = rng.integers(0, m, size=(3, L))
        v[rng.integers(3), rng.integers(L)] = m
        cov["g%d" % k] = v
    genes_df = pd.DataFrame({"chr": "chr1", "gene": list(cov.keys()), "gene_start": 1, "gene_end": lengths})
    reads = pd.DataFrame({"chr": "chr1", "gene": list(cov.keys()), "s1": 1.0, "s2": 2.0, "s3": 3.0})
    return cov, genes_df, reads


def test_gene_filter_equals_the_reference_loop():
    """degnorm_b200.gene_filter against a literal restatement of __main__.py:221-244."""
    from collections import OrderedDict
    from degnorm_b200.gene_filter import filter_genes
    for minimax, rate in ((0, 1), (10, 1), (10, 20), (81, 6), (13, 21)):
        cov, genes_df, reads = _filter_case()
        want_cov = OrderedDict((g, m) for g, m in cov.items())
        delete_idx = []
        for i in range(genes_df.shape[0]):
            gene = genes_df.gene.iloc[i]
            cov_mat = want_cov[gene]
            if (cov_mat.max() < minimax) or (cov_mat.shape[1] <= rate):
                delete_idx.append(i)
                del want_cov[gene]
        want_genes = genes_df.drop(delete_idx, axis=0).reset_index(drop=True)
        want_reads = reads.drop(delete_idx, axis=0).reset_index(drop=True)
        got_cov, got_genes, got_reads = filter_genes(cov, genes_df, reads, minimax, rate)
        assert got_cov is cov and list(got_cov.keys()) == list(want_cov.keys())
        assert got_genes.equals(want_genes) and got_reads.equals(want_reads)
    cov, genes_df, reads = _filter_case()
    with pytest.raises(ValueError, match="No genes available"):
        filter_genes(cov, genes_df, reads, minimax_coverage=10 ** 6)
    cov, genes_df, reads = _filter_case()
    with pytest.raises(ValueError, match="Number of coverage matrices"):
        cov["extra"] = np.ones((3, 50))
        filter_genes(cov, genes_df, reads, minimax_coverage=0)


def test_device_coverage_handle_filters_like_filter_genes_on_host_tensors():
    """gene_filter.DeviceCoverage with host tensors (the same code runs on the device): packing, select (runs of
    kept genes), filter against filter_genes, round trip to the reference's dictionary."""
    from degnorm_b200.gene_filter import DeviceCoverage, filter_genes
    for minimax, rate in ((0, 1), (10, 1), (10, 20), (81, 6), (13, 21)):
        cov, genes_df, reads = _filter_case()
        dc = DeviceCoverage(cov)
        assert len(dc) == 7 and dc.keys() == list(cov.keys()) and dc.p == 3
        got, g_df, r_df = dc.filter(genes_df, reads, minimax, rate)
        want_cov, want_genes, want_reads = filter_genes(cov, genes_df, reads, minimax, rate)
        assert got.keys() == list(want_cov.keys())
        assert g_df.equals(want_genes) and r_df.equals(want_reads)
        back = got.to_dict()
        for g in want_cov:
            np.testing.assert_array_equal(back[g], want_cov[g])
        assert got.offsets[-1] == sum(m.shape[1] for m in want_cov.values())
    cov, genes_df, reads = _filter_case()
    dc = DeviceCoverage(cov)
    assert dc.select(np.ones(7, dtype=bool)) is dc
    with pytest.raises(ValueError, match="No genes available"):
        dc.filter(genes_df, reads, minimax_coverage=10 ** 6)
    with pytest.raises(ValueError, match="in order"):
        dc.filter(genes_df.iloc[::-1].reset_index(drop=True), reads)


def test_mid_kernel_plans_are_pure_host_arithmetic():
    """dn_make_plan for 13..48 samples (no device call): the default 8-warp instantiation and the warp-specialised
    one (8 Gram warps + 4 update warps, warps = 12) take one CTA per SM with a 153.6 KB ring, the 4-warp one fits two
    CTAs per SM (G parked in the free ring stage); workspace columns are whole chunks; clusters divide the CTA count."""
    import ctypes as C
    from degnorm_b200 import _lib
    from degnorm_b200.engine import Params
    lib = _lib.lib()
    prm = Params().to_c(48)
    sm, smem = 148, 232448

    def plan(max_cols, n_work, warps, cluster):
        pl = _lib.DnPlan()
        rc = lib.dn_make_plan(C.byref(prm), max_cols, n_work, 0, 0, warps, cluster, sm, smem, C.byref(pl))
        assert rc == 0, lib.dn_last_error()
        return pl
    pw = plan(5000, 10000, 12, 1)
    assert (pw.tile, pw.threads, pw.cluster, pw.ctas) == (6, 384, 1, 148)
    p8 = plan(5000, 10000, 0, 1)
    assert (p8.tile, p8.threads, p8.cluster, p8.ctas) == (6, 256, 1, 148)
    assert plan(5000, 10000, 8, 1).threads == 256
    assert pw.smem_bytes == p8.smem_bytes and pw.ws_cols == 5024 and pw.ws_cols % 32 == 0    # 32-column chunks, 6 stages
    assert p8.smem_bytes <= smem and 2 * (p8.smem_bytes + 1024) > smem          # one CTA per SM
    assert p8.ws_cols == 5056 and p8.ws_cols % 64 == 0
    p4 = plan(5000, 10000, 4, 1)
    assert (p4.tile, p4.threads, p4.ctas) == (6, 128, 296)
    assert 2 * (p4.smem_bytes + 1024) <= smem + 1024                             # two CTAs per SM
    assert p4.ws_cols == 5024 and p4.ws_cols % 32 == 0
    pc = plan(40000, 5, 0, 4)
    assert (pc.cluster, pc.ctas, pc.ws_cols) == (4, 20, 10048)
    assert plan(40000, 1000, 0, 16).ctas == 144                                   # 9 clusters of 16 on 148 SMs
    pl = _lib.DnPlan()
    assert lib.dn_make_plan(C.byref(prm), 5000, 10, 0, 0, 0, 3, sm, smem, C.byref(pl)) != 0      # cluster of 3


def test_numa_pinning_helper_is_best_effort(tmp_path):
    """distributed.pin_to_gpu_numa_node: cpulist parsing, and no exception (nor any change of affinity) when the
    device or the sysfs files are not there."""
    import os
    from degnorm_b200.distributed import parse_cpulist, pin_to_gpu_numa_node
    assert parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert parse_cpulist("5") == [5] and parse_cpulist("") == []
    before = os.sched_getaffinity(0)
    info = pin_to_gpu_numa_node(0, sysfs=str(tmp_path))
    assert info["pinned"] is False and os.sched_getaffinity(0) == before
