"""degnorm_b200.coverage_merge against a brute-force statement of what the merge
computes and, when /root/reference is present (the build container; never the GPU box), against the unmodified
reference functions (reads_coverage_merge.py:93-164, 167-372, 375-461): same genes in the same order, same values,
equal per-chromosome pickles; the matrices come back as views of ONE contiguous buffer that packing recognises."""
import os
import pickle as pkl
import sys

import numpy as np
import pytest

REF = "/root/reference"


def _make_dataset(tmp_path, seed=0):
    import pandas as pd
    from scipy import sparse
    rng = np.random.default_rng(seed)
    data_dir = str(tmp_path / "data")
    samples = ["s1", "s2", "s3"]
    n_pos = {"chr1": 6000, "chr2": 4000, "chrM": 500}
    dense = {}
    for s in samples:
        os.makedirs(os.path.join(data_dir, s))
        for chrom, n in n_pos.items():
            if (s, chrom) == ("s2", "chr2") or chrom == "chrM":
                continue                                    # a sample without this chromosome; a chromosome nobody read
            v = rng.poisson(2.0, size=n) * (rng.random(n) < 0.5)
            dense[s, chrom] = v
            sparse.save_npz(os.path.join(data_dir, s, "chrom_coverage_%s_%s.npz" % (s, chrom)),
                            sparse.csr_matrix(v.reshape(1, -1)))
    rows = [("chr1", "A", 10, 60), ("chr1", "A", 50, 120), ("chr1", "B", 300, 400), ("chr1", "C", 1000, 1150),
            ("chr1", "C", 1100, 1300), ("chr1", "D", 900, 1250), ("chr1", "E", 5000, 5999),
            ("chr2", "F", 1, 80), ("chr2", "G", 3000, 3500), ("chr2", "G", 3600, 3700), ("chrM", "H", 5, 50)]
    exon = pd.DataFrame(rows, columns=["chr", "gene", "start", "end"])
    exon["gene_start"] = exon.groupby("gene").start.transform("min")
    exon["gene_end"] = exon.groupby("gene").end.transform("max")
    # overlapping genes of chr1 (C and D overlap: their vectors come from the overlap pickles, with their own lengths)
    for k, s in enumerate(samples):
        ov = {"C": rng.poisson(3.0, size=260).astype(float), "D": rng.poisson(1.0, size=330).astype(float)}
        if k == 2:
            del ov["D"]                                     # a later sample without the gene: zero row
        with open(os.path.join(data_dir, s, "overlap_coverage_%s_chr1.pkl" % s), "wb") as f:
            pkl.dump(ov, f)
    return data_dir, samples, exon, dense


def _brute_force(samples, exon, dense, chrom):
    sub = exon[exon.chr == chrom].sort_values("gene_end")
    out = {}
    for g in sub.gene.unique():
        e = sub[sub.gene == g]
        pos = sorted(set(q for s0, e0 in zip(e.start, e.end) for q in range(s0 - 1, e0)))
        out[g] = np.array([[float(dense[s, chrom][q]) if (s, chrom) in dense else 0.0 for q in pos] for s in samples])
    return out


def test_merge_against_brute_force_and_zero_copy_layout(tmp_path):
    from degnorm_b200.coverage_merge import merge_chrom_coverage, merge_coverage, merge_overlap_gene_coverage
    from degnorm_b200.packing import pack_coverage
    data_dir, samples, exon, dense = _make_dataset(tmp_path)
    one = merge_chrom_coverage(data_dir, samples, exon[exon.chr == "chr1"], verbose=False)
    want = _brute_force(samples, exon, dense, "chr1")
    assert list(one.keys()) == list(want.keys()) == ["A", "B", "D", "C", "E"]
    for g in want:
        assert one[g].dtype == np.float64 and one[g].flags.c_contiguous
        np.testing.assert_array_equal(one[g], want[g])
    assert merge_chrom_coverage(data_dir, samples, exon[exon.chr == "chrM"], verbose=False) == {}
    with pytest.raises(ValueError):
        merge_chrom_coverage(data_dir, samples, exon, verbose=False)          # more than one chromosome
    ov = merge_overlap_gene_coverage(data_dir, samples, "chr1")
    assert list(ov.keys()) == ["C", "D"] and ov["C"].shape == (3, 260) and ov["D"].shape == (3, 330)
    assert not ov["D"][2].any() and ov["D"][0].any()
    assert merge_overlap_gene_coverage(data_dir, samples, "chr2") == {}

    out_dir = str(tmp_path / "out")
    os.makedirs(out_dir)
    allg = merge_coverage(data_dir, samples, exon, output_dir=out_dir, verbose=False)
    assert list(allg.keys()) == ["A", "B", "D", "C", "E", "F", "G"]           # chrM contributes nothing
    for g in ("A", "B", "E"):
        np.testing.assert_array_equal(allg[g], want[g])
    np.testing.assert_array_equal(allg["C"], ov["C"])                        # overlap vectors replace the cut ones
    np.testing.assert_array_equal(allg["D"], ov["D"])
    want2 = _brute_force(samples, exon, dense, "chr2")
    np.testing.assert_array_equal(allg["G"], want2["G"])
    assert not allg["G"][1].any()                                            # s2 has no chr2 file: zeros
    # one contiguous staging buffer: the packer takes it as it is (no copy)
    mats = list(allg.values())
    flat, off = pack_coverage(mats, 3, pin=False)
    assert flat.numpy().__array_interface__["data"][0] == mats[0].__array_interface__["data"][0]
    assert off[-1] == sum(m.shape[1] for m in mats)
    with open(os.path.join(out_dir, "chr1", "coverage_matrices_chr1.pkl"), "rb") as f:
        saved = pkl.load(f)
    assert list(saved.keys()) == ["A", "B", "D", "C", "E"]
    np.testing.assert_array_equal(saved["A"], want["A"])
    assert saved["A"].flags.f_contiguous and saved["C"].flags.c_contiguous    # the reference's memory orders


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "degnorm")), reason="reference tree not present")
def test_merge_equals_the_unmodified_reference(tmp_path):
    from degnorm_b200 import coverage_merge as ours
    data_dir, samples, exon, _ = _make_dataset(tmp_path, seed=5)
    if not hasattr(np, "float_"):
        np.float_ = np.float64                      # the reference predates numpy 2 (reads_coverage_merge.py:151, 353)
    sys.path.insert(0, REF)
    try:
        import degnorm.reads_coverage_merge as rcm
    finally:
        sys.path.remove(REF)
    out_a, out_b = str(tmp_path / "ref"), str(tmp_path / "ours")
    os.makedirs(out_a)
    os.makedirs(out_b)
    np.random.seed(1)
    ref = rcm.merge_coverage(data_dir, samples, exon, n_jobs=1, output_dir=out_a, verbose=False)
    state_ref = np.random.get_state()[2]
    np.random.seed(1)
    got = ours.merge_coverage(data_dir, samples, exon, n_jobs=1, output_dir=out_b, verbose=False)
    assert np.random.get_state()[2] == state_ref              # the global numpy stream is left where the reference leaves it
    assert list(got.keys()) == list(ref.keys())
    for g in ref:
        assert got[g].shape == ref[g].shape and got[g].dtype == ref[g].dtype
        np.testing.assert_array_equal(got[g], ref[g])
    for chrom in ("chr1", "chr2"):
        with open(os.path.join(out_a, chrom, "coverage_matrices_%s.pkl" % chrom), "rb") as f:
            a = f.read()
        with open(os.path.join(out_b, chrom, "coverage_matrices_%s.pkl" % chrom), "rb") as f:
            b = f.read()
        assert a == b, chrom                                  # byte-identical pickles
    for chrom in ("chr1", "chr2"):
        sub = exon[exon.chr == chrom]
        np.random.seed(2)
        r1 = rcm.merge_chrom_coverage(data_dir, samples, sub, verbose=False)
        np.random.seed(2)
        g1 = ours.merge_chrom_coverage(data_dir, samples, sub, verbose=False)
        assert list(r1.keys()) == list(g1.keys())
        for g in r1:
            np.testing.assert_array_equal(g1[g], r1[g])
        r2 = rcm.merge_overlap_gene_coverage(data_dir, samples, chrom)
        g2 = ours.merge_overlap_gene_coverage(data_dir, samples, chrom)
        assert list(r2.keys()) == list(g2.keys())
        for g in r2:
            np.testing.assert_array_equal(g2[g], r2[g])


def test_merge_in_several_position_groups(tmp_path, monkeypatch):
    """The chromosome vectors are densified one gene group at a time (coverage_merge._SPAN positions): a span small
    enough to force a group per gene or two gives the same matrices."""
    from degnorm_b200 import coverage_merge as cm
    data_dir, samples, exon, dense = _make_dataset(tmp_path, seed=3)
    monkeypatch.setattr(cm, "_SPAN", 400)
    got = cm.merge_chrom_coverage(data_dir, samples, exon[exon.chr == "chr1"], verbose=False)
    want = _brute_force(samples, exon, dense, "chr1")
    assert list(got.keys()) == list(want.keys())
    for g in want:
        np.testing.assert_array_equal(got[g], want[g])
    # a gene that runs past the end of the stored chromosome vector is an IndexError, as in the reference
    bad = exon[exon.chr == "chr2"].copy()
    bad.loc[bad.gene == "G", "end"] = 10 ** 6
    bad["gene_end"] = bad.groupby("gene").end.transform("max")
    with pytest.raises(IndexError):
        cm.merge_chrom_coverage(data_dir, samples, bad, verbose=False)
