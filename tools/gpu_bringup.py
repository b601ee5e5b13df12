#!/usr/bin/env python
"""Bring-up / debugging aid: runs the golden cases on the GPU and prints a per-gene comparison with the
fixtures (made by the real reference).  Usage: python tools/gpu_bringup.py [case ...]"""
import os
import sys
import time
from collections import OrderedDict

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_case, RUN_CASES          # noqa: E402
from degnorm_b200 import GeneNMFOA                 # noqa: E402
from degnorm_b200._lib import EXIT_NAMES           # noqa: E402

np.set_printoptions(linewidth=200, precision=6, suppress=False)
for case in (sys.argv[1:] or RUN_CASES):
    mats, reads, kwargs, ref = load_case(case)
    m = GeneNMFOA(**kwargs)
    t0 = time.time()
    est = m.run(OrderedDict(("g%d" % i, x) for i, x in enumerate(mats)), reads)
    dt = time.time() - t0
    w = ref["nmf_widths"]
    calls = (w >= 0).sum(axis=2)
    cols = np.where(w >= 0, w, 0).sum(axis=2)
    print("== %s  kwargs=%s  %.2fs  timings=%s" % (case, kwargs, dt, m.timings))
    print("  max|d rho| %.3e   max|d x_adj| %.3e   max rel d scale %.3e   ran equal: %s" % (
        np.abs(m.rho - ref["rho"]).max(), np.abs(m.x_adj - ref["x_adj"]).max(),
        np.abs(m.scale_factors / ref["scale_factors"] - 1).max(), np.array_equal(m.ran_baseline_selection, ref["ran"])))
    print("  est max abs diff per gene:", ["%.2e" % np.abs(a - b).max() for a, b in zip(est, ref["estimates"])])
    for it in range(m.counters.shape[0]):
        c = m.counters[it]
        print("  iter %d: exits %s" % (it, [EXIT_NAMES[int(e)] for e in c[:, 0]]))
        print("          n_hi %s calls gpu %s ref %s  cols gpu %s ref %s  eig_steps/solve %s resident %s" % (
            c[:, 1].tolist(), c[:, 2].tolist(), calls[it].tolist(), c[:, 3].tolist(), cols[it].tolist(),
            np.round(c[:, 4] / np.maximum(1, c[:, 2] * (m.nmf_iter + 1)), 1).tolist(), (c[:, 7] & 1).tolist()))
        print("          eig fallbacks %s" % (c[:, 7] >> 1).tolist())
    bad = np.argwhere(np.abs(m.rho - ref["rho"]) > 1e-6)
    if len(bad):
        print("  MISMATCH rows:", sorted(set(bad[:, 0].tolist())))
        for g in sorted(set(bad[:, 0].tolist()))[:4]:
            print("   gene", g, "gpu", m.rho[g], "\n        ref", ref["rho"][g])
