"""Tuning aid: the same genes through the generic tiled kernel and through the mid-p / wide kernel (optionally one
cluster per gene): per-gene DI difference and counters.  usage: compare_kernels.py P CLUSTER DEGNORM_ITER NMF_ITER"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from degnorm_b200.engine import Params, ShardEngine
from degnorm_b200.packing import pack_coverage
from degnorm_b200.synth import synth_numpy
p, cluster, n_it, T = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
lengths = np.array([300, 520, 1210, 150, 95, 33])
mats, reads = synth_numpy(len(lengths), p, 500 + p, lengths=lengths, jitter=1e-6)
prm = Params(degnorm_iter=n_it, nmf_iter=T)
flat, off = pack_coverage(mats, p)
outs = []
for new in (False, True):
    eng = ShardEngine(prm, p, "cuda:0")
    if p > 48: eng.use_wide = new
    else: eng.use_mid = new
    eng.force_cluster = cluster if (new and cluster > 1) else 0
    eng.load(flat.cuda(), off, torch.from_numpy(reads).cuda())
    o = eng.run(None, want_estimates=False)
    torch.cuda.synchronize()
    outs.append({k: v.cpu().numpy() for k, v in o.items() if torch.is_tensor(v)})
a, b = outs
print("p", p, "cluster", cluster, "per-gene max|d rho|", np.abs(a["rho"] - b["rho"]).max(axis=1))
print("counters old", a["counters"][0, :, :4].tolist()); print("counters new", b["counters"][0, :, :4].tolist())
