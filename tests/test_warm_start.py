"""load_from_previous drop-in (reference: degnorm/warm_start.py:10-106): same outputs and side effects on a fake
previous-run directory, and the coverage comes back as views of one contiguous buffer that pack_coverage takes
without copying."""
import os
import pickle

import numpy as np
import pandas as pd
import pytest


def _fake_run_dir(root):
    rng = np.random.default_rng(3)
    genes = {"chr1": ["gA", "gB", "gX"], "chr2": ["gC"]}
    exon_rows, cov = [], {}
    for chrom, gs in genes.items():
        os.makedirs(root / chrom)
        d = {}
        for k, g in enumerate(gs):
            L = 40 + 13 * k
            m = rng.poisson(5.0, size=(3, L)).astype(float)
            d[g] = np.asfortranarray(m) if k % 2 else m          # the merge step emits both layouts
            cov[g] = m
            exon_rows += [dict(chr=chrom, gene=g, gene_start=100 * k, gene_end=100 * k + L, start=100 * k, end=100 * k + 10),
                          dict(chr=chrom, gene=g, gene_start=100 * k, gene_end=100 * k + L, start=100 * k + 20, end=100 * k + L)]
        with open(root / chrom / ("coverage_matrices_%s.pkl" % chrom), "wb") as f:
            pickle.dump(d, f)
    pd.DataFrame(exon_rows).to_csv(root / "gene_exon_metadata.csv", index=False)
    # gX has no read counts, gZ no annotation: both must be dropped; row order differs from the coverage order
    pd.DataFrame({"chr": ["chr2", "chr1", "chr1", "chr9"], "gene": ["gC", "gB", "gA", "gZ"],
                  "s1": [7, 3, 1, 9], "s2": [8, 4, 2, 9], "s3": [9, 5, 3, 9]}).to_csv(root / "read_counts.csv", index=False)
    return cov


def test_load_from_previous_matches_reference_contract(tmp_path):
    from degnorm_b200.packing import pack_coverage
    from degnorm_b200.warm_start import load_from_previous
    old, new = tmp_path / "old", tmp_path / "new"
    os.makedirs(old)
    cov = _fake_run_dir(old)
    with pytest.raises(IOError):
        load_from_previous(str(old), str(new))                       # new directory must exist
    os.makedirs(new)
    out = load_from_previous(str(old), str(new))
    assert sorted(out) == ["gene_cov_dict", "genes_df", "read_count_df", "sample_ids"]
    assert out["sample_ids"] == ["s1", "s2", "s3"]
    genes = list(out["gene_cov_dict"].keys())
    assert genes == ["gA", "gB", "gC"]                               # chromosome by chromosome, pickle order, intersected
    assert out["read_count_df"].gene.tolist() == genes and out["genes_df"].gene.tolist() == genes
    assert out["read_count_df"].s1.tolist() == [1, 3, 7]
    assert out["genes_df"].columns.tolist() == ["gene", "chr", "gene_start", "gene_end"]
    for g in genes:
        np.testing.assert_array_equal(out["gene_cov_dict"][g], cov[g])
        assert out["gene_cov_dict"][g].flags.c_contiguous and out["gene_cov_dict"][g].dtype == np.float64
    # side effects: files copied into the new output directory
    assert os.path.isfile(new / "read_counts.csv") and os.path.isfile(new / "gene_exon_metadata.csv")
    assert os.path.isfile(new / "chr1" / "coverage_matrices_chr1.pkl") and os.path.isfile(new / "chr2" / "coverage_matrices_chr2.pkl")
    # the matrices are back-to-back views of one buffer: the packer uses it as it is
    mats = list(out["gene_cov_dict"].values())
    flat, off = pack_coverage(mats, 3, pin=False)
    assert flat.numpy().__array_interface__["data"][0] == mats[0].__array_interface__["data"][0]
    assert off.tolist() == [0, 40, 93, 133]
    # pack=False keeps the pickled arrays themselves (reference behaviour)
    os.makedirs(tmp_path / "new2")
    raw = load_from_previous(str(old), str(tmp_path / "new2"), pack=False)
    assert not raw["gene_cov_dict"]["gB"].flags.c_contiguous
    # a missing file raises FileNotFoundError like the reference
    os.remove(old / "read_counts.csv")
    os.makedirs(tmp_path / "new3")
    with pytest.raises(FileNotFoundError):
        load_from_previous(str(old), str(tmp_path / "new3"))


def test_same_as_the_reference_loader_when_it_is_importable(tmp_path):
    """Only where /root/reference exists (the build container): the unmodified reference loader on the same directory."""
    import sys
    ref_root = "/root/reference"
    if not os.path.isdir(os.path.join(ref_root, "degnorm")):
        pytest.skip("reference checkout not present")
    sys.dont_write_bytecode = True
    sys.path.insert(0, ref_root)
    try:
        from degnorm.warm_start import load_from_previous as ref_load
    except Exception as exc:                                          # pragma: no cover
        pytest.skip("reference not importable here: %s" % exc)
    finally:
        sys.path.remove(ref_root)
    from degnorm_b200.warm_start import load_from_previous
    old = tmp_path / "old"
    os.makedirs(old)
    _fake_run_dir(old)
    os.makedirs(tmp_path / "a")
    os.makedirs(tmp_path / "b")
    ours = load_from_previous(str(old), str(tmp_path / "a"))
    ref = ref_load(str(old), str(tmp_path / "b"))
    assert list(ours["gene_cov_dict"].keys()) == list(ref["gene_cov_dict"].keys())
    for g in ref["gene_cov_dict"]:
        np.testing.assert_array_equal(ours["gene_cov_dict"][g], ref["gene_cov_dict"][g])
    assert ours["sample_ids"] == ref["sample_ids"]
    pd.testing.assert_frame_equal(ours["read_count_df"], ref["read_count_df"])
    pd.testing.assert_frame_equal(ours["genes_df"], ref["genes_df"])


@pytest.mark.gpu
def test_streamed_to_device_equals_host_path_and_runs(tmp_path):
    """load_from_previous(device=...): chromosome-by-chromosome upload overlapped with reading the next pickle; the
    resident coverage equals the host path's buffer bit for bit, in the same gene order, and GeneNMFOA.run takes it."""
    import torch
    from degnorm_b200 import GeneNMFOA
    from degnorm_b200.packing import pack_coverage
    from degnorm_b200.warm_start import load_from_previous
    old = tmp_path / "old"
    os.makedirs(old)
    _fake_run_dir(old)
    os.makedirs(tmp_path / "a")
    os.makedirs(tmp_path / "b")
    host = load_from_previous(str(old), str(tmp_path / "a"))
    dev = load_from_previous(str(old), str(tmp_path / "b"), device="cuda:0")
    dc = dev["gene_cov_dict"]
    assert dc.keys() == list(host["gene_cov_dict"].keys()) and dc.p == 3
    flat, off = pack_coverage(list(host["gene_cov_dict"].values()), 3, pin=False)
    assert dc.offsets.tolist() == off.tolist() and dc.flat.is_cuda
    np.testing.assert_array_equal(dc.flat.cpu().numpy(), flat.numpy())
    assert dev["read_count_df"].equals(host["read_count_df"]) and dev["genes_df"].equals(host["genes_df"])
    assert os.path.isfile(tmp_path / "b" / "chr2" / "coverage_matrices_chr2.pkl")
    kw = dict(degnorm_iter=1, nmf_iter=5, min_high_coverage=5)
    reads = host["read_count_df"][host["sample_ids"]].values.astype(float)
    m_host, m_dev = GeneNMFOA(**kw), GeneNMFOA(**kw)
    m_host.run(host["gene_cov_dict"], reads)
    m_dev.run(dc, reads)
    np.testing.assert_array_equal(m_host.rho, m_dev.rho)
    np.testing.assert_array_equal(m_host.x_adj, m_dev.x_adj)
