"""Diagnostic (GPU): runs the full-size C2 workload twice, genes in two different orders, and dumps the coverage of
genes whose results differ to gpurun_out/flips2.npz (source of tests/golden/ill_conditioned.npz)."""
import numpy as np, torch, sys
sys.path.insert(0, '.')
from degnorm_b200.engine import Params, ShardEngine, draw_offsets
from degnorm_b200.synth import CONFIGS, config_lengths, synth_torch
cfg = CONFIGS["c2"]; n, p, rate = cfg["n_genes"], cfg["p"], cfg["downsample_rate"]
lengths = config_lengths("c2")
flat, off, reads = synth_torch(lengths, p, cfg["seed"], "cuda:0")
gen = torch.Generator(device="cuda:0"); gen.manual_seed(7)
flat.mul_(1.0 + 1.0e-6 * torch.rand(flat.numel(), generator=gen, device="cuda:0", dtype=torch.float64))
prm = Params(downsample_rate=rate)
ds = draw_offsets(n, prm)
def run(flat, off, reads, ds):
    eng = ShardEngine(prm, p, "cuda:0"); eng.load(flat, off, reads)
    o = eng.run(ds, want_estimates=False); torch.cuda.synchronize()
    return {k: v.cpu().numpy().copy() for k, v in o.items() if torch.is_tensor(v)}, eng
a, eng = run(flat, off, reads, ds)
rng = np.random.default_rng(11); perm = rng.permutation(n)
off_p = np.zeros(n + 1, dtype=np.int64); np.cumsum(lengths[perm], out=off_p[1:])
flat_p = torch.empty_like(flat)
for k, g in enumerate(perm):
    flat_p[p * off_p[k]: p * off_p[k + 1]] = flat[p * off[g]: p * off[g + 1]]
c, _ = run(flat_p, off_p, reads[torch.as_tensor(perm, device="cuda:0")].contiguous(), np.ascontiguousarray(ds[:, perm]))
print("scale_used rel diff (last iteration):", np.abs(a["scale_used"]/c["scale_used"]-1).max())
out = {}
for k in (1131, 2580, 601, 4672):
    g = perm[k]
    F = flat[p * off[g]: p * off[g + 1]].view(p, int(lengths[g])).cpu().numpy()
    out["F_%d" % g] = F; out["ds_%d" % g] = ds[4, g]
    out["cnt_a_%d" % g] = a["counters"][4][g]; out["cnt_c_%d" % g] = c["counters"][4][k]
    out["rho_a_%d" % g] = a["rho"][g]; out["rho_c_%d" % g] = c["rho"][k]
    print("gene", g, "L", lengths[g], a["counters"][4][g][:7].tolist(), c["counters"][4][k][:7].tolist())
out["scale_a"] = a["scale_used"]; out["scale_c"] = c["scale_used"]
np.savez_compressed("gpurun_out/flips2.npz", **out)
