"""
CPU oracle for the DegNorm NMF-OA hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product path (degnorm_b200/) never does: it fails loudly when the CUDA extension is
missing.

This is a numpy restatement of the algorithm in the reference's degnorm/nmf.py (v0.1.4).  It is not a
copy: bins are kept as an alive-mask over fixed column ranges instead of the
reference's np.delete / shift_bins bookkeeping, the rank-one step is pluggable, and every routine can emit
a trace that the CUDA path is compared against.  Each function cites the reference lines it restates.

Rank-one step (`rank1=`):
  * "svds"  -- scipy.sparse.linalg.svds(x, k=1), the third-party call the reference makes
               (nmf.py:63; pinned scipy==0.19.1 in config/requirements.txt:3, scipy 1.18.1 installed
               here).  A fixed v0 is passed so the legacy global numpy RNG is left alone (SURVEY.md
               Appendix C-1); results do not depend on v0.
  * "gram"  -- top eigenvector of the p x p Gram matrix via numpy.linalg.eigh;
               ~50x faster, agrees with "svds" to ~1e-15.  Used for the larger parity cases.

Parity pin: tests/golden/*.npz were produced by the *real* reference imported from /root/reference by
oracle/gen_golden.py; tests/test_oracle_golden.py checks this module against them.
"""
import math

import numpy as np

__all__ = ["Params", "rank_one", "nmf", "ratio_svd", "high_coverage_idx", "bin_bounds",
           "baseline_selection", "draw_offsets", "run", "full_length_estimate"]


class Params(object):
    """Normalised algorithm parameters (nmf.py:30-53)."""

    def __init__(self, degnorm_iter=5, downsample_rate=1, min_high_coverage=50, nmf_iter=100, bins=20,
                 skip_baseline_selection=False, random_state=123, rank1="svds"):
        self.degnorm_iter = abs(int(degnorm_iter))
        self.nmf_iter = abs(int(nmf_iter))
        self.bins = abs(int(bins))
        self.min_high_coverage = max(2, abs(int(min_high_coverage)))
        self.min_bins = math.ceil(self.bins * 0.2)
        self.downsample_rate = abs(int(downsample_rate))
        self.skip_baseline_selection = bool(skip_baseline_selection)
        self.random_state = random_state
        self.rank1 = rank1
        if self.downsample_rate > 1:          # nmf.py:52-53
            self.min_high_coverage = 2


# --------------------------------------------------------------------------------------------------
# rank-one approximation, NMF-OA, ratio-SVD
# --------------------------------------------------------------------------------------------------
def rank_one(x, method="svds"):
    """Top singular triplet as (K = u*s  [p x 1],  E = vh  [1 x L]).  nmf.py:55-64."""
    if min(x.shape) < 2:
        # svds(k=1) demands k < min(shape); the reference swallows this ValueError at nmf.py:306-310.
        raise ValueError("rank_one needs min(shape) >= 2, got %r" % (x.shape,))
    if method == "svds":
        from scipy.sparse.linalg import svds
        u, s, vh = svds(x, k=1, v0=np.ones(min(x.shape)))
        return u * s, vh
    if method == "gram":
        g = x @ x.T
        _, vecs = np.linalg.eigh(g)
        v = vecs[:, -1]
        if v.sum() < 0:
            v = -v
        proj = v @ x                               # sigma * vh
        sigma = math.sqrt(float(proj @ proj))
        if sigma == 0.0:
            raise ArithmeticError("all-zero matrix (the reference raises ArpackError here)")
        return (v * sigma).reshape(-1, 1), (proj / sigma).reshape(1, -1)
    raise ValueError("unknown rank-one method %r" % (method,))


def nmf(x, nmf_iter, method="svds"):
    """NMF-OA factors of x after nmf_iter multiplier updates (no final clamp).  nmf.py:78-107."""
    K, E = rank_one(x, method)
    est = K @ E
    lam = np.zeros_like(x)
    c = 1.0 / math.sqrt(nmf_iter) if nmf_iter > 0 else 0.0
    for _ in range(nmf_iter):
        lam = lam - c * (est - x)
        np.maximum(lam, 0.0, out=lam)
        K, E = rank_one(x + lam, method)
        est = K @ E
    return K, E


def ratio_svd(x, method="svds"):
    """One rank-one fit clamped from below by x.  nmf.py:109-121."""
    K, E = rank_one(x, method)
    return np.maximum(K @ E, x)


def high_coverage_idx(F):
    """Columns whose sample-wise max exceeds 10 % of the matrix max (strict).  nmf.py:66-76."""
    return np.flatnonzero(F.max(axis=0) > 0.1 * F.max())


def bin_bounds(n, n_chunks):
    """[lo, hi) ranges of utils.split_into_chunks(range(n), n_chunks).  utils.py:176-192."""
    cs = int(math.ceil(n / n_chunks))
    out = []
    i = 0
    while i * cs < n:
        out.append((i * cs, min(i * cs + cs, n)))
        i += 1
    return out


def _floored_abs(K):
    """|K| with entries < 1e-5 replaced by the smallest entry >= 1e-5.  nmf.py:329-330, 361-362."""
    K = np.abs(K).copy()
    big = K[K >= 1.0e-5]
    if big.size == 0:
        raise ArithmeticError("no K entry >= 1e-5 (np.min of an empty array in the reference)")
    K[K < 1.0e-5] = big.min()
    return K


def _di(cov_rowsum, est_rowsum):
    return 1.0 - cov_rowsum / (est_rowsum + 1.0)


def full_length_estimate(F, K):
    """Envelope back-out over all L columns.  nmf.py:358-365."""
    K = _floored_abs(K)
    E = (F / K).max(axis=0, keepdims=True)
    return np.maximum(K @ E, F)


# exit / branch codes recorded in traces (shared vocabulary with the CUDA counters)
EXIT_FEW_HICOV = 1      # nmf.py:232-233
EXIT_EMPTY_SAMPLE = 2   # nmf.py:241-242
EXIT_MEDIAN = 3         # nmf.py:257-258
EXIT_NO_SELECTION = 4   # nmf.py:265 false: unclamped first fit
EXIT_REFINED = 5        # nmf.py:327-337
EXIT_FALLBACK_HIGH = 6  # nmf.py:342-346
EXIT_FALLBACK = 7       # nmf.py:349-353


def baseline_selection(F, prm, ds_start=None, trace=None):
    """
    Baseline selection for one (already scale-adjusted) p x L coverage matrix.  nmf.py:189-372.

    ds_start: systematic-sample offset when prm.downsample_rate > 1 (the draw itself is made by the caller,
    one per gene per outer iteration -- nmf.py:420-422).
    Returns (rho [p], estimate [p x L], ran_baseline_selection).
    """
    p, L = F.shape
    tr = trace if trace is not None else {}
    tr.update(exit=0, n_hi=0, nmf_calls=0, sum_cols=0, drops=[])
    rho_default = np.zeros(p)

    hi = high_coverage_idx(F)
    if prm.downsample_rate > 1:
        if prm.downsample_rate >= L:                                   # nmf.py:443-444
            raise ValueError("Cannot downsample at a rate < 1 / length(gene)")
        keep = np.arange(ds_start, L, prm.downsample_rate)
        hi = np.intersect1d(keep, hi)
    n = hi.size
    tr["n_hi"] = int(n)
    if n < prm.min_high_coverage:
        tr["exit"] = EXIT_FEW_HICOV
        return rho_default, F, False

    F0 = np.ascontiguousarray(F[:, hi])
    rs0 = F0.sum(axis=1)
    if np.sum(rs0 > 0) < p:
        tr["exit"] = EXIT_EMPTY_SAMPLE
        return rho_default, F, False

    def fit(cols):
        tr["nmf_calls"] += 1
        tr["sum_cols"] += int(cols.shape[1])
        K_, E_ = nmf(cols, prm.nmf_iter, prm.rank1)
        # smallest factor entry relative to the largest, over all fits: a numerically-zero entry (~1e-15) makes the
        # reference's exact test `min(rowsum(KE)) == 0` (nmf.py:314) a coin toss decided by rounding noise
        k_ = np.abs(np.asarray(K_)).ravel()
        tr["min_rel_K"] = min(tr.get("min_rel_K", 1.0), float(k_.min() / k_.max()) if k_.max() > 0 else 0.0)
        return K_, E_

    K, E = fit(F0)
    K0, E0 = K.copy(), E.copy()
    KE = K @ E
    estimate = KE.copy()
    rho = _di(rs0, KE.sum(axis=1))
    if np.nanmedian(1.0 - rho) > 1:
        tr["exit"] = EXIT_MEDIAN
        return rho_default, F, False

    ran = False
    min_len = max(2, math.ceil(200.0 * (1.0 / prm.downsample_rate)))
    if n >= min_len and np.nanmin(rho) <= 0.2 and not prm.skip_baseline_selection:
        bounds = bin_bounds(n, prm.bins)
        alive = np.ones(len(bounds), dtype=bool)

        def alive_cols():
            m = np.zeros(n, dtype=bool)
            for b, (lo, hi_) in enumerate(bounds):
                if alive[b]:
                    m[lo:hi_] = True
            return m

        Fb = F0
        while np.nanmax(rho) > 0.1:
            ran = True
            rel = (KE - Fb) / (Fb + 1.0)
            res = np.max(rel * rel, axis=0)                 # per-column max squared relative residual
            # bin means over the *current* (compacted) column set
            mask = alive_cols()
            pos = np.cumsum(mask) - 1                       # original column -> compact column
            ss = []
            ids = []
            for b, (lo, hi_) in enumerate(bounds):
                if alive[b]:
                    ss.append(np.nanmean(res[pos[lo]:pos[hi_ - 1] + 1]))
                    ids.append(b)
            ss = np.array(ss)
            if np.nanmax(ss) == 0:
                break
            d = ids[int(np.nanargmax(ss))]
            alive[d] = False
            tr["drops"].append(int(d))
            Fb = np.ascontiguousarray(F0[:, alive_cols()])
            n_cur = Fb.shape[1]
            try:
                K, E = fit(Fb)
            except ValueError:
                break
            KE = K @ E
            if np.min(KE.sum(axis=1)) == 0:
                break
            KE = np.maximum(KE, Fb)
            rho = _di(Fb.sum(axis=1), KE.sum(axis=1))
            if alive.sum() <= prm.min_bins or n_cur < min_len:
                break

        if np.nanmax(rho) < 0.2:
            K = _floored_abs(K)
            E = (F0 / K).max(axis=0, keepdims=True)
            estimate = K @ E
            rho = _di(rs0, estimate.sum(axis=1))
            tr["exit"] = EXIT_REFINED
            if np.nanmax(rho) > 0.9:
                K, E = K0, E0
                estimate = np.maximum(K @ E, F0)
                rho = _di(rs0, estimate.sum(axis=1))
                tr["exit"] = EXIT_FALLBACK_HIGH
        else:
            K, E = K0, E0
            estimate = np.maximum(K @ E, F0)
            rho = _di(rs0, estimate.sum(axis=1))
            tr["exit"] = EXIT_FALLBACK
    else:
        tr["exit"] = EXIT_NO_SELECTION

    if estimate.shape[1] < L:
        estimate = full_length_estimate(F, K)
    tr["K"] = np.abs(np.asarray(K)).ravel().copy()
    return rho, estimate, ran


# --------------------------------------------------------------------------------------------------
# whole run
# --------------------------------------------------------------------------------------------------
def draw_offsets(n_genes, prm):
    """
    Canonical systematic-sample offsets: after np.random.seed(random_state), one
    np.random.choice(rate) per gene per outer iteration, iteration-major, in gene order (nmf.py:420-422,
    556).  A private legacy RandomState reproduces that stream without touching the global one.
    """
    if prm.downsample_rate <= 1:
        return np.zeros((prm.degnorm_iter, n_genes), dtype=np.int64)
    rs = np.random.RandomState(prm.random_state)
    out = np.empty((prm.degnorm_iter, n_genes), dtype=np.int64)
    for it in range(prm.degnorm_iter):
        for g in range(n_genes):
            out[it, g] = rs.choice(prm.downsample_rate)
    return out


def init_scale(cov_mats, reads, method="svds"):
    """Initial DI scores and normalisation factors.  nmf.py:521-535."""
    est_sums = np.vstack([ratio_svd(F, method).sum(axis=1) for F in cov_mats])
    cov_sums = np.vstack([F.sum(axis=1) for F in cov_mats])
    rho0 = 1.0 - cov_sums / (est_sums + 1.0)
    low = rho0.max(axis=1) < 0.1
    count_sums = reads[low, :].sum(axis=0) if np.any(low) else reads.sum(axis=0)
    norm = count_sums / np.median(count_sums)
    return rho0, norm


def outer_update(x_weighted, rho, scale_factors):
    """Read-count / scale-factor update after one outer iteration.  nmf.py:148-158, 575-590.

    rho is the clipped DI matrix; it is corrected in place for genes whose row max is 0."""
    x_adj = x_weighted / (1.0 - rho)
    nb = rho.max(axis=1) == 0
    if nb.sum() > 0:
        rho[nb, :] = 1.0 - x_weighted.sum(axis=0) / x_adj.sum(axis=0)
    x_adj = x_weighted / (1.0 - rho)
    colsum = x_adj.sum(axis=0)
    norm = colsum / np.median(colsum)
    return x_adj, norm, x_weighted / norm, scale_factors * norm


def run(cov_mats, reads, prm, want_estimates=True, offsets=None, progress=None):
    """GeneNMFOA.run restated (nmf.py:483-601) for a list of p x L_g float64 matrices and an n x p count
    matrix.  Returns a dict with rho, x_adj, scale_factors, norm_factors, x_weighted, ran, estimates,
    traces (last outer iteration) and rho0."""
    n = len(cov_mats)
    x = np.array(reads, dtype=np.float64)
    p = cov_mats[0].shape[0]
    if x.shape[0] != n:
        raise ValueError("Number of genes in read count matrix not equal to number of coverage matrices!")
    if prm.downsample_rate > 1 and min(F.shape[1] for F in cov_mats) < prm.downsample_rate:
        raise ValueError("downsample_rate is too large; take-every size > at least one gene.")
    rho0, norm = init_scale(cov_mats, x, prm.rank1)
    x_w = x / norm
    scale = norm.copy()
    ran = np.zeros((n, prm.degnorm_iter), dtype=bool)
    if offsets is None:
        offsets = draw_offsets(n, prm)
    rho = rho0
    x_adj = None
    estimates = None
    traces = None
    scale_used = scale.copy()
    for it in range(prm.degnorm_iter):
        scale_used = scale.copy()
        rows, flags, ests, trs = [], [], [], []
        for g, F in enumerate(cov_mats):
            tr = {}
            r_, e_, f_ = baseline_selection((F.T / scale).T, prm, int(offsets[it, g]), tr)
            rows.append(r_)
            flags.append(f_)
            trs.append(tr)
            if want_estimates and it == prm.degnorm_iter - 1:
                ests.append(e_)
            if progress is not None:
                progress(it, g)
        rho = np.clip(np.vstack(rows), 0.0, 0.9)               # nmf.py:397-399
        ran[:, it] = flags
        x_adj, norm, x_w_new, scale = outer_update(x_w, rho, scale)
        x_w = x_w_new
        estimates, traces = ests, trs
    return dict(rho=rho, x_adj=x_adj, scale_factors=scale, norm_factors=norm, x_weighted=x_w, ran=ran,
                estimates=estimates, traces=traces, rho0=rho0, p=p, scale_used_last=scale_used)
