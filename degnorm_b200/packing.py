"""Host side of the ragged CSR coverage buffer: {gene: p x L_g float64} -> one flat pinned buffer in which
gene g occupies [p*off[g], p*off[g+1]) as a C-contiguous p x L_g block (include/degnorm_b200.h), and back.
Inputs may be C- or Fortran-contiguous (the reference's merge step emits both, reads_coverage_merge.py:155-159,
331, 353); they are never modified."""
import numpy as np
import torch


def pack_coverage(cov_mats, p, pin=True):
    lengths = np.fromiter((m.shape[1] for m in cov_mats), dtype=np.int64, count=len(cov_mats))
    offsets = np.zeros(len(cov_mats) + 1, dtype=np.int64)
    np.cumsum(lengths, out=offsets[1:])
    total = int(offsets[-1]) * p
    flat = torch.empty(total, dtype=torch.float64)
    if pin and torch.cuda.is_available() and total > 0:
        flat = flat.pin_memory()
    dst = flat.numpy()
    for g, m in enumerate(cov_mats):
        a, b = p * int(offsets[g]), p * int(offsets[g + 1])
        dst[a:b].reshape(p, -1)[...] = m            # handles any strides / dtype
    return flat, offsets


def unpack_estimates(est_dev, offsets, p):
    """Device ragged buffer -> list of p x L_g numpy arrays (views into one host copy)."""
    host = torch.empty(est_dev.shape, dtype=est_dev.dtype, pin_memory=torch.cuda.is_available() and est_dev.numel() > 0)
    host.copy_(est_dev)
    if est_dev.is_cuda:
        torch.cuda.synchronize(est_dev.device)
    arr = host.numpy()
    return [arr[p * int(offsets[g]): p * int(offsets[g + 1])].reshape(p, -1) for g in range(len(offsets) - 1)]
