#!/usr/bin/env python
"""Debug aid: one gene of a golden case, GPU vs numpy oracle, sweeping nmf_iter."""
import os, sys
from collections import OrderedDict
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_case
from degnorm_b200 import GeneNMFOA
from oracle import nmfoa_oracle as orc
np.set_printoptions(linewidth=200, precision=10)
case, g = sys.argv[1], int(sys.argv[2])
mats, reads, kwargs, ref = load_case(case)
sel = [g] if len(sys.argv) < 4 else [int(a) for a in sys.argv[2:]]
mats = [mats[i] for i in sel]; reads = reads[sel]
for T in (0, 1, 2, 3, 5, 10, 20, 40, 70, 100):
    for it in (1, 2):
        kw = dict(kwargs); kw["nmf_iter"] = T; kw["degnorm_iter"] = it
        m = GeneNMFOA(**kw); m.run(OrderedDict((str(i), x) for i, x in enumerate(mats)), reads)
        o = orc.run(mats, reads, orc.Params(rank1="gram", **kw))
        c = m.counters[-1]
        print("T=%3d iters=%d  max|drho|=%.3e  gpu exit %s calls %s n_hi %s fb %s | orc exit %s calls %s n_hi %s" % (
            T, it, np.abs(m.rho - o["rho"]).max(), c[:, 0].tolist(), c[:, 2].tolist(), c[:, 1].tolist(), (c[:, 7] >> 1).tolist(),
            [t["exit"] for t in o["traces"]], [t["nmf_calls"] for t in o["traces"]], [t["n_hi"] for t in o["traces"]]))
        if np.abs(m.rho - o["rho"]).max() > 1e-6:
            print("   gpu rho", m.rho, "\n   orc rho", o["rho"], "\n   scale gpu", m.scale_factors, "orc", o["scale_factors"])
