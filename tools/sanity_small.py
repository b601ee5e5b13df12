#!/usr/bin/env python
"""Tiny run through every kernel family (for compute-sanitizer): small-p resident / cluster / streamed, mid-p with
and without clusters, tiled."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from degnorm_b200.engine import Params, ShardEngine          # noqa: E402
from degnorm_b200.packing import pack_coverage               # noqa: E402
from degnorm_b200.synth import synth_numpy                   # noqa: E402

cases = [(4, dict(), "small resident"), (12, dict(force_cluster=4), "small cluster resident"),
         (12, dict(force_cluster=2, force_streamed=True), "small cluster streamed"),
         (12, dict(force_streamed=True), "small streamed"), (48, dict(), "mid"), (20, dict(force_cluster=4), "mid cluster"),
         (70, dict(), "tiled")]
lengths = np.array([300, 520, 95, 1250, 260, 700])
for p, opt, name in cases:
    mats, reads = synth_numpy(len(lengths), p, 5 + p, lengths=lengths, jitter=1e-6)
    flat, off = pack_coverage(mats, p)
    eng = ShardEngine(Params(degnorm_iter=1, nmf_iter=6), p, "cuda:0", force_streamed=opt.get("force_streamed", False))
    eng.force_cluster = opt.get("force_cluster", 0)
    eng.load(flat.cuda(), off, torch.from_numpy(reads).cuda())
    out = eng.run(None, want_estimates=True)
    torch.cuda.synchronize()
    print(name, "ok", float(out["rho"].sum()))
