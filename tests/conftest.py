import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_case(name):
    """Golden fixture -> (cov_mats, reads, kwargs, outputs dict).  Fixtures come from the real reference
    (oracle/gen_golden.py)."""
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    p = int(d["p"])
    mats = []
    pos = 0
    for L, f in zip(d["lengths"], d["layout_f"]):
        L = int(L)
        m = d["cov_flat"][pos:pos + p * L].reshape(p, L).copy()
        if f:
            m = np.asfortranarray(m)
        mats.append(m)
        pos += p * L
    kwargs = {}
    for k, v in zip(d["kw_keys"], d["kw_vals"]):
        k = str(k)
        kwargs[k] = bool(v) if k == "skip_baseline_selection" else int(v)
    ests = []
    pos = 0
    for L in d["lengths"]:
        L = int(L)
        ests.append(d["est_flat"][pos:pos + p * L].reshape(p, L))
        pos += p * L
    out = {k: d[k] for k in ("rho", "x_adj", "scale_factors", "norm_factors", "x_weighted", "ran", "nmf_widths")}
    out["estimates"] = ests
    return mats, d["reads"].copy(), kwargs, out


def load_seeded_case(name):
    """Seeded golden fixture (oracle/gen_golden.py:save_seeded_case): the inputs are regenerated with synth_numpy
    from the stored generator arguments and checked against the stored digest; the outputs are the unmodified
    reference's.  -> (cov_mats, reads, kwargs, outputs dict)."""
    import hashlib
    from degnorm_b200.synth import synth_numpy
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    lengths = d["lengths"]
    mats, reads = synth_numpy(len(lengths), int(d["p"]), int(d["seed"]), lengths=lengths,
                              fortran_every=int(d["fortran_every"]), jitter=float(d["jitter"]))
    h = hashlib.sha256()
    for m in mats:
        h.update(np.ascontiguousarray(m).tobytes())
    h.update(np.ascontiguousarray(reads).tobytes())
    if h.hexdigest() != str(d["digest"]):
        pytest.skip("synth_numpy does not reproduce the fixture's inputs on this numpy build")
    kwargs = {}
    for k, v in zip(d["kw_keys"], d["kw_vals"]):
        k = str(k)
        kwargs[k] = bool(v) if k == "skip_baseline_selection" else int(v)
    out = {k: d[k] for k in ("rho", "x_adj", "scale_factors", "norm_factors", "x_weighted", "ran", "nmf_widths",
                             "est_rowsum", "est_max")}
    return mats, reads, kwargs, out


SEEDED_CASES = ["seed_p17", "seed_p48", "seed_p48_long", "seed_p100", "seed_p200", "seed_p12_long"]
RUN_CASES = ["run_p4", "run_p4_ds", "run_p12", "run_skip", "run_p3_bins"]
TIE_CASE = "run_p4_ds_ties"      # integer counts: exact ties in the high-coverage test (DESIGN.md "ties")


@pytest.fixture(scope="session")
def golden_loader():
    return load_case
