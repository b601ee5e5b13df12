"""Host side of the ragged CSR coverage buffer: {gene: p x L_g float64} -> one flat (pinned) buffer in which
gene g occupies [p*off[g], p*off[g+1]) as a C-contiguous p x L_g block (include/degnorm_b200.h), and back.
Inputs may be C- or Fortran-contiguous (the reference's merge step emits both, reads_coverage_merge.py:155-159,
331, 353); they are never modified.

Two paths:
  * zero-copy: the matrices already are back-to-back C-contiguous float64 views of one host buffer (what a
    loader that reads straight into a staging buffer produces, and what bench.py builds): the buffer is used as is.
  * general: matrices are copied into a cached pinned staging buffer by a small thread pool (numpy copies release
    the GIL)."""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

def pinned_buffer(numel, tag, cache):
    """A pinned float64 host tensor of at least `numel` elements, cached in `cache` (a dict owned by the caller;
    pinned allocation is slow, so one GeneNMFOA object re-uses its staging buffers from run to run)."""
    t = cache.get(tag)
    if t is None or t.numel() < numel:
        cache[tag] = None
        t = torch.empty(max(int(numel), 1), dtype=torch.float64)
        if torch.cuda.is_available():
            t = t.pin_memory()
        cache[tag] = t
    return t[:numel]


def _root(a):
    while isinstance(a.base, np.ndarray):
        a = a.base
    return a


def _contiguous_view(cov_mats, p, offsets):
    """The flat ndarray the matrices are back-to-back views of, or None."""
    m0 = cov_mats[0]
    root = _root(m0)
    if root.dtype != np.float64 or not root.flags.c_contiguous:
        return None
    base = m0.__array_interface__["data"][0]
    for g, m in enumerate(cov_mats):
        if m.dtype != np.float64 or not m.flags.c_contiguous or m.shape[0] != p:
            return None
        if m.__array_interface__["data"][0] != base + 8 * p * int(offsets[g]):
            return None
        if _root(m) is not root:
            return None
    start = (base - root.__array_interface__["data"][0]) // 8
    total = p * int(offsets[-1])
    flat = root.reshape(-1)
    if start < 0 or start + total > flat.size:
        return None
    return flat[start:start + total]


def pack_coverage(cov_mats, p, pin=True, threads=None, cache=None):
    """-> (flat float64 host tensor, int64 offsets [n+1])."""
    n = len(cov_mats)
    lengths = np.fromiter((m.shape[1] for m in cov_mats), dtype=np.int64, count=n)
    offsets = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lengths, out=offsets[1:])
    total = int(offsets[-1]) * p
    if n > 0 and total > 0:
        view = _contiguous_view(cov_mats, p, offsets)
        if view is not None:
            return torch.from_numpy(view), offsets
    flat = pinned_buffer(total, "cov", {} if cache is None else cache) if pin else torch.empty(total, dtype=torch.float64)
    dst = flat.numpy()

    def copy_range(lo, hi):
        for g in range(lo, hi):
            a, b = p * int(offsets[g]), p * int(offsets[g + 1])
            dst[a:b].reshape(p, -1)[...] = cov_mats[g]          # handles any strides / dtype

    threads = threads or min(16, os.cpu_count() or 1)
    if total < (1 << 22) or threads == 1:
        copy_range(0, n)
    else:
        # split by bytes, not by gene count
        cuts = np.searchsorted(offsets, np.linspace(0, offsets[-1], 4 * threads + 1)[1:-1])
        bounds = [0] + sorted(set(int(c) for c in cuts)) + [n]
        with ThreadPoolExecutor(max_workers=threads) as ex:
            list(ex.map(lambda ab: copy_range(*ab), zip(bounds[:-1], bounds[1:])))
    return flat, offsets


def unpack_estimates(est_dev, offsets, p, cache=None):
    """Device ragged buffer -> list of p x L_g numpy arrays (views into one pinned host copy).  With a `cache` the
    host copy is re-used by the next call with the same cache: np.copy() anything that must outlive it."""
    host = pinned_buffer(est_dev.numel(), "est", {} if cache is None else cache)
    host.copy_(est_dev)
    if est_dev.is_cuda:
        torch.cuda.synchronize(est_dev.device)
    arr = host.numpy()
    return [arr[p * int(offsets[g]): p * int(offsets[g + 1])].reshape(p, -1) for g in range(len(offsets) - 1)]
