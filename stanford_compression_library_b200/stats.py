"""Statistics that feed the coders: byte histograms on the device and frequency-table normalisation.

`DataBlock.get_counts` / `get_empirical_distribution` (scl/core/data_block.py:37-94) are the
reference's way to derive a model from data; `histogram_blocks` does it for a whole batch in one
launch, and `normalize_frequencies` turns counts into a `Frequencies` whose total is a power of
two, which tANS requires (tANS.py:42-44) and the rANS fast paths prefer.
"""
import ctypes

import numpy as np
import torch

from . import _cabi
from .core.prob_dist import Frequencies
from .device import _ptr, _stream, require_cuda


def histogram_blocks(data: torch.Tensor, sizes=None, per_block: bool = True, total: bool = True):
    """data: uint8 [B, N] on the GPU.  Returns (counts uint32-as-int32 [B, 256] or None, totals int64 [256] or None)."""
    require_cuda()
    assert data.is_cuda and data.dtype == torch.uint8 and data.dim() == 2
    data = data.contiguous()
    B, N = data.shape
    if sizes is not None:
        sizes = sizes.to(device=data.device, dtype=torch.int32).contiguous()
    counts = torch.empty((B, 256), dtype=torch.int32, device=data.device) if per_block else None
    tot = torch.zeros(256, dtype=torch.int64, device=data.device) if total else None
    with torch.cuda.device(data.device):
        rc = _cabi.lib().scl_histogram_blocks(_ptr(data), data.stride(0), _ptr(sizes), N, B, _ptr(counts), _ptr(tot), _stream())
    _cabi.check(rc, "scl_histogram_blocks")
    return counts, tot


def normalize_frequencies(counts, total_freq: int = 4096, keep_zeros: bool = False) -> Frequencies:
    """Integer table with sum == total_freq from raw counts (index = byte value).

    f_b = max(1, floor(c_b * total_freq / sum c)) for every symbol that occurs (all 256 if
    keep_zeros), then the surplus / deficit is taken from / given to the most frequent symbols --
    the SURVEY.md 8(d) quantiser generalised so that it always lands exactly on total_freq.
    Keys are byte values in ascending order (dict order defines the cumulative table)."""
    c = np.asarray(counts.cpu() if isinstance(counts, torch.Tensor) else counts, dtype=np.int64).reshape(-1)
    present = np.ones_like(c, dtype=bool) if keep_zeros else c > 0
    n_present = int(present.sum())
    if n_present == 0:
        raise ValueError("no symbols")
    if n_present > total_freq:
        raise ValueError("total_freq is smaller than the alphabet")
    tot = int(c.sum())
    f = np.zeros_like(c)
    if tot == 0:
        f[present] = 1
    else:
        f[present] = np.maximum(1, (c[present] * total_freq) // tot)
    diff = total_freq - int(f.sum())
    order = np.argsort(-c, kind="stable")
    order = order[present[order]]
    i = 0
    while diff != 0:
        j = order[i % len(order)]
        if diff > 0:
            f[j] += diff if i == 0 else 1
            diff = total_freq - int(f.sum())
        elif f[j] > 1:
            take = min(f[j] - 1, -diff)
            f[j] -= take
            diff += take
        i += 1
    assert int(f.sum()) == total_freq and (f[present] >= 1).all()
    return Frequencies({int(b): int(f[b]) for b in range(len(c)) if present[b]})


def empirical_frequencies(data: torch.Tensor, sizes=None, total_freq=None) -> Frequencies:
    """Frequencies of a whole batch: raw counts (total_freq=None) or normalised to total_freq."""
    _, tot = histogram_blocks(data, sizes=sizes, per_block=False, total=True)
    tot = tot.cpu().numpy()
    if total_freq is None:
        return Frequencies({int(b): int(tot[b]) for b in range(256) if tot[b] > 0})
    return normalize_frequencies(tot, total_freq)
