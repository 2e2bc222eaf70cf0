"""Synthetic inputs of the BASELINE.json configs: the Zipf-1.0 quantiser and
on-device i.i.d. sampling.  Used by bench.py, the smoke test and the GPU parity tests."""
import numpy as np
import torch

from .core.prob_dist import Frequencies


def zipf_freq_list(n_sym: int = 256, M: int = 4096, s: float = 1.0):
    """f_b = max(1, floor(p_b * M)), p_b ~ 1/(b+1)^s, deficit added to f_0."""
    p = 1.0 / np.arange(1, n_sym + 1, dtype=np.float64) ** s
    p /= p.sum()
    f = np.maximum(1, np.floor(p * M).astype(np.int64))
    f[0] += M - f.sum()
    assert f.sum() == M and f.min() >= 1
    return [int(x) for x in f]


def zipf_frequencies(n_sym: int = 256, M: int = 4096, s: float = 1.0) -> Frequencies:
    return Frequencies({b: f for b, f in enumerate(zipf_freq_list(n_sym, M, s))})


def zipf_probabilities(n_sym: int = 256, s: float = 1.0):
    """The source distribution of the benchmark data: p_b ~ 1/(b+1)^s (not the quantised table)."""
    p = 1.0 / np.arange(1, n_sym + 1, dtype=np.float64) ** s
    return (p / p.sum()).tolist()


def sample_blocks(weights, n_blocks: int, block_len: int, seed: int, device, chunk_blocks: int = 16384) -> torch.Tensor:
    """uint8 [n_blocks, block_len] of i.i.d. draws with P(b) ~ weights[b] by inverse CDF, generated on `device`."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    f = torch.tensor(weights, dtype=torch.float64, device=device)
    cdf = torch.cumsum(f / f.sum(), 0).to(torch.float32)
    cdf[-1] = 2.0  # guard against u == 1.0 rounding
    out = torch.empty((n_blocks, block_len), dtype=torch.uint8, device=device)
    for b0 in range(0, n_blocks, chunk_blocks):
        b1 = min(n_blocks, b0 + chunk_blocks)
        u = torch.rand((b1 - b0, block_len), generator=g, device=device, dtype=torch.float32)
        out[b0:b1] = torch.searchsorted(cdf, u, right=True).clamp_(max=len(weights) - 1).to(torch.uint8)
    return out


def sample_stream_blocks(weights, lo: int, hi: int, block_len: int, device, chunk_blocks: int = 16384, base_seed: int = 0) -> torch.Tensor:
    """Blocks [lo, hi) of ONE global synthetic stream: chunk k of `chunk_blocks` blocks is drawn from its own
    generator (seed base_seed + k), so a rank's shard holds the same bytes whatever the number of ranks the
    stream is split over (strong scaling: the 8 GiB of BASELINE configs[4] at 1, 2, 4 or 8 GPUs)."""
    f = torch.tensor(weights, dtype=torch.float64, device=device)
    cdf = torch.cumsum(f / f.sum(), 0).to(torch.float32)
    cdf[-1] = 2.0
    out = torch.empty((hi - lo, block_len), dtype=torch.uint8, device=device)
    g = torch.Generator(device=device)
    for k in range(lo // chunk_blocks, (hi + chunk_blocks - 1) // chunk_blocks):
        g.manual_seed(base_seed + k)
        u = torch.rand((chunk_blocks, block_len), generator=g, device=device, dtype=torch.float32)
        sym = torch.searchsorted(cdf, u, right=True).clamp_(max=len(weights) - 1).to(torch.uint8)
        c0 = k * chunk_blocks
        a, b = max(lo, c0), min(hi, c0 + chunk_blocks)
        out[a - lo : b - lo] = sym[a - c0 : b - c0]
    return out
