"""
Synthetic coverage generator for parity tests and benchmarks.

There is no network and the reference ships no coverage data (its test BAMs are missing), so every
workload is synthetic: log-normal gene lengths, a smooth multi-bump coverage envelope per gene,
per-sample depth, a 3'-biased degradation ramp per gene x sample, Poisson counts.

`CONFIGS` names the BASELINE.json workloads.  `synth_numpy` is the specification (used for tests and
golden fixtures); `synth_torch` draws from the same family with torch ops so that bench-scale inputs
(GBs) can be produced on the GPU in a second -- the two are not bit-identical and nothing relies on
that.
"""
import math

import numpy as np

BASE_SEED = 20261018

# name -> (n_genes, p, downsample_rate, length profile)
CONFIGS = {
    "c1": dict(n_genes=1000, p=4, downsample_rate=1, profile="pc", seed=BASE_SEED + 1),
    "c2": dict(n_genes=20000, p=12, downsample_rate=20, profile="pc", seed=BASE_SEED + 2),
    "c3": dict(n_genes=60000, p=48, downsample_rate=1, profile="pc", seed=BASE_SEED + 3),
    "c4": dict(n_genes=5000, p=12, downsample_rate=1, profile="long", seed=BASE_SEED + 4),
    "c5": dict(n_genes=20000, p=200, downsample_rate=1, profile="pc", seed=BASE_SEED + 5),
}

PROFILES = {
    # mu, sigma of log-length, clip range
    "pc": (math.log(2500.0), 0.8, 200, 100000),
    "long": (math.log(60000.0), 1.3, 10000, 2000000),
}


def gene_lengths(n_genes, rng, profile="pc", lmin=None, lmax=None):
    mu, sigma, lo, hi = PROFILES[profile]
    lo = lo if lmin is None else lmin
    hi = hi if lmax is None else lmax
    L = np.rint(np.exp(rng.normal(mu, sigma, size=n_genes))).astype(np.int64)
    return np.clip(L, lo, hi)


def synth_numpy(n_genes, p, seed, profile="pc", lmin=None, lmax=None, fortran_every=0, lengths=None, jitter=0.0):
    """
    Returns (cov_mats, reads): list of n_genes float64 p x L_g arrays and an n_genes x p float64 count
    matrix.  fortran_every=k makes every k-th gene Fortran-contiguous (the reference's merge step emits
    both layouts -- reads_coverage_merge.py:331,353 vs :155-159).
    jitter > 0 multiplies every count by (1 + jitter*u), u ~ U(0,1): integer counts make the reference's
    high-coverage test `max_i F_ij > 0.1*max(F)` (nmf.py:76) an EXACT tie whenever a column maximum is one tenth
    of the matrix maximum, and the outcome then hangs on the last bit of the scale factors; parity fixtures use
    jittered counts so that the reference's answer is well defined ("ties").
    """
    rng = np.random.default_rng(seed)
    if lengths is None:
        lengths = gene_lengths(n_genes, rng, profile, lmin, lmax)
    depth = rng.uniform(0.5, 2.0, size=p)
    severity = rng.uniform(0.0, 0.6, size=p)
    cov_mats = []
    reads = np.zeros((n_genes, p))
    for g in range(n_genes):
        L = int(lengths[g])
        j = np.arange(L, dtype=np.float64)
        amp = math.exp(rng.normal(math.log(30.0), 1.2))
        if rng.random() < 0.05:
            amp *= 1.0e-3                                  # starved genes: "< 50 high-coverage positions"
        nb = int(rng.integers(1, 5))
        env = np.full(L, 0.15)
        for _ in range(nb):
            c = rng.uniform(0.0, L)
            s = rng.uniform(0.03, 0.3) * L
            w = rng.uniform(0.3, 1.0)
            env += w * np.exp(-0.5 * ((j - c) / s) ** 2)
        env *= amp
        abundance = depth * np.exp(rng.normal(0.0, 0.25, size=p))
        if rng.random() < 0.30:
            delta = np.zeros(p)                             # intact gene: the max(rho) <= 0.1 path
        else:
            delta = severity * rng.beta(2.0, 3.0, size=p)
        slope = np.minimum(2.0 * delta, 0.98)
        ramp = 1.0 - slope[:, None] * (1.0 - j[None, :] / L)
        F = rng.poisson(abundance[:, None] * env[None, :] * ramp).astype(np.float64)
        reads[g] = np.rint(F.sum(axis=1) / 100.0)
        if jitter > 0.0:
            F = F * (1.0 + jitter * rng.random(F.shape))
        if fortran_every and g % fortran_every == fortran_every - 1:
            F = np.asfortranarray(F)
        cov_mats.append(F)
    return cov_mats, reads


def synth_torch(lengths, p, seed, device, dtype=None, out=None):
    """
    Same family as synth_numpy, generated with torch on `device` straight into one packed buffer:
    gene g occupies out[p*off[g] : p*off[g+1]] as a C-contiguous p x L_g block.
    Returns (packed fp64 tensor, offsets int64 numpy [n+1], reads float64 tensor n x p).
    """
    import torch
    dtype = dtype or torch.float64
    lengths = np.asarray(lengths, dtype=np.int64)
    n = len(lengths)
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lengths, out=off[1:])
    total = int(off[-1])
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed))
    if out is None:
        out = torch.empty(total * p, dtype=dtype, device=device)
    reads = torch.empty((n, p), dtype=torch.float64, device=device)
    depth = 0.5 + 1.5 * torch.rand(p, generator=gen, device=device, dtype=torch.float64)
    severity = 0.6 * torch.rand(p, generator=gen, device=device, dtype=torch.float64)
    # per-gene scalars, drawn in bulk
    amp = torch.exp(math.log(30.0) + 1.2 * torch.randn(n, generator=gen, device=device, dtype=torch.float64))
    starved = torch.rand(n, generator=gen, device=device) < 0.05
    amp = torch.where(starved, amp * 1.0e-3, amp)
    nb = torch.randint(1, 5, (n,), generator=gen, device=device)
    centre = torch.rand((n, 4), generator=gen, device=device, dtype=torch.float64)
    width = 0.03 + 0.27 * torch.rand((n, 4), generator=gen, device=device, dtype=torch.float64)
    weight = 0.3 + 0.7 * torch.rand((n, 4), generator=gen, device=device, dtype=torch.float64)
    weight = weight * (torch.arange(4, device=device)[None, :] < nb[:, None])
    abundance = depth[None, :] * torch.exp(0.25 * torch.randn((n, p), generator=gen, device=device,
                                                               dtype=torch.float64))
    intact = torch.rand(n, generator=gen, device=device) < 0.30
    beta = torch.distributions.Beta(torch.tensor(2.0, device=device, dtype=torch.float64),
                                    torch.tensor(3.0, device=device, dtype=torch.float64))
    torch.manual_seed(int(seed))
    delta = severity[None, :] * beta.sample((n, p))
    delta = torch.where(intact[:, None], torch.zeros_like(delta), delta)
    slope = torch.clamp(2.0 * delta, max=0.98)
    # genes are generated in length-sorted groups so each group is one dense [g, p, Lmax] tensor op
    order = np.argsort(lengths, kind="stable")
    budget = 32 * 1024 * 1024                               # elements per group
    i = 0
    while i < n:
        Lmax_i = int(lengths[order[i]])
        k = i
        while k < n and (k - i + 1) * int(lengths[order[k]]) * p <= budget:
            k += 1
        k = max(k, i + 1)
        ids = order[i:k]
        Lmax = int(lengths[ids[-1]])
        tid = torch.as_tensor(ids, device=device)
        Lg = torch.as_tensor(lengths[ids], device=device, dtype=torch.float64)
        j = torch.arange(Lmax, device=device, dtype=torch.float64)[None, :]              # 1 x Lmax
        u = j / Lg[:, None]                                                                # g x Lmax
        env = torch.full_like(u, 0.15)
        for b in range(4):
            z = (u - centre[tid, b][:, None]) / width[tid, b][:, None]
            env = env + weight[tid, b][:, None] * torch.exp(-0.5 * z * z)
        env = env * amp[tid][:, None]
        ramp = 1.0 - slope[tid][:, :, None] * (1.0 - u[:, None, :])                        # g x p x Lmax
        lam = abundance[tid][:, :, None] * env[:, None, :] * ramp
        F = torch.poisson(lam.to(torch.float32), generator=gen).to(dtype)
        valid = (j < Lg[:, None])[:, None, :]
        F = F * valid
        reads[tid] = torch.round(F.sum(dim=2, dtype=torch.float64) / 100.0)
        for q, g in enumerate(ids):
            L = int(lengths[g])
            out[p * int(off[g]): p * int(off[g + 1])].view(p, L).copy_(F[q, :, :L])
        i = k
        del F, lam, ramp, env, u
    return out, off, reads


def config_lengths(name, n_genes=None):
    cfg = CONFIGS[name]
    rng = np.random.default_rng(cfg["seed"])
    return gene_lengths(n_genes or cfg["n_genes"], rng, cfg["profile"])
