#!/usr/bin/env python
"""Generator of tests/golden/ill_conditioned.npz (GPU box; DESIGN.md section 2, "ill-conditioned genes").

Runs the full-size synthetic C2 workload (20,000 genes x 12 samples, take-every 20) twice -- genes in the caller's
order and in a permuted order, which changes nothing but the order in which per-sample sums are added -- and dumps
  * a gene whose bin-drop loop ends at a different nmf() call in the two runs after the SECOND outer iteration
    (scale factors 1e-15 apart; a factor entry is numerically zero, nmf.py:314), and
  * a starved gene whose DI differs by O(0.1) between the two runs after the FIFTH (scale factors 1e-6 apart),
with the coverage, the down-sampling offset and both scale-factor vectors.  No oracle is involved here; the CPU test
tests/test_oracle_golden.py::test_reference_algorithm_is_ill_conditioned_on_starved_genes reads the file.

usage (GPU box):  python tools/dump_flipped_genes.py [out.npz]     (default gpurun_out/ill_conditioned.npz)
"""
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from degnorm_b200.engine import Params, ShardEngine, draw_offsets
from degnorm_b200.synth import CONFIGS, config_lengths, synth_torch

ZERO_FACTOR_GENE, STARVED_GENE = 15437, 18385         # gene ids in the caller's order (found by the equivariance test)

cfg = CONFIGS["c2"]
n, p, rate = cfg["n_genes"], cfg["p"], cfg["downsample_rate"]
lengths = config_lengths("c2")
flat, off, reads = synth_torch(lengths, p, cfg["seed"], "cuda:0")
gen = torch.Generator(device="cuda:0")
gen.manual_seed(7)
flat.mul_(1.0 + 1.0e-6 * torch.rand(flat.numel(), generator=gen, device="cuda:0", dtype=torch.float64))
rng = np.random.default_rng(11)
perm = rng.permutation(n)
off_p = np.zeros(n + 1, dtype=np.int64)
np.cumsum(lengths[perm], out=off_p[1:])
flat_p = torch.empty_like(flat)
for k, g in enumerate(perm):
    flat_p[p * off_p[k]: p * off_p[k + 1]] = flat[p * off[g]: p * off[g + 1]]
reads_p = reads[torch.as_tensor(perm, device="cuda:0")].contiguous()


def run(prm, flat, off, reads, ds):
    eng = ShardEngine(prm, p, "cuda:0")
    eng.load(flat, off, reads)
    o = eng.run(ds, want_estimates=False)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy().copy() for k, v in o.items() if torch.is_tensor(v)}


def gene(g):
    return flat[p * off[g]: p * off[g + 1]].view(p, int(lengths[g])).cpu().numpy()


out = {}
for n_iter, tag, g in ((2, "zero_factor", ZERO_FACTOR_GENE), (5, "starved", STARVED_GENE)):
    prm = Params(downsample_rate=rate, degnorm_iter=n_iter)
    ds = draw_offsets(n, prm)
    a = run(prm, flat, off, reads, ds)
    c = run(prm, flat_p, off_p, reads_p, np.ascontiguousarray(ds[:, perm]))
    k = int(np.flatnonzero(perm == g)[0])
    last = n_iter - 1
    print(tag, "gene", g, "L", lengths[g], "scale rel diff %.2e" % np.abs(a["scale_used"] / c["scale_used"] - 1).max(),
          "counters", a["counters"][last][g][:7].tolist(), c["counters"][last][k][:7].tolist(),
          "max |dDI| %.3e" % np.abs(a["rho"][g] - c["rho"][k]).max())
    out["F_" + tag] = gene(g)
    out["ds_" + tag] = ds[last, g]
    sa, sc = ("scale1_a", "scale1_c") if tag == "zero_factor" else ("scale_a", "scale_c")
    out[sa], out[sc] = a["scale_used"], c["scale_used"]
path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/ill_conditioned.npz"
np.savez_compressed(path, **out)
print("wrote", path)
