"""Multi-GPU plumbing for the block-sharded path.

Blocks are independent, so the path shards by contiguous block ranges with NO payload exchange:
rank r owns blocks [r*B/n, (r+1)*B/n).  The only collectives are tiny: a broadcast of the
frequency table from rank 0 (so every rank builds identical device tables) and an optional
all-gather of per-rank compressed sizes (global offsets of a concatenated stream).  Works with
the `nccl` backend (device tensors, NVLink/NVSwitch) and with `gloo` (CPU tests).
"""
import numpy as np
import torch
import torch.distributed as dist

from .core.prob_dist import Frequencies


def shard_range(n_blocks: int, rank: int, world_size: int):
    """Contiguous block range [lo, hi) owned by `rank`; sizes differ by at most one block."""
    base, rem = divmod(int(n_blocks), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _comm_device():
    if dist.is_initialized() and dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def broadcast_frequencies(freqs, src: int = 0) -> Frequencies:
    """Broadcast an integer-keyed Frequencies (keys 0..255) from rank `src`.

    Non-source ranks may pass None.  Wire format: int64[1 + 2*256] = [n, keys..., counts...]."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return freqs
    dev = _comm_device()
    wire = torch.zeros(1 + 512, dtype=torch.int64, device=dev)
    if dist.get_rank() == src:
        keys = list(freqs.freq_dict)
        assert len(keys) <= 256 and all(isinstance(k, (int, np.integer)) for k in keys), "only integer-keyed tables travel"
        wire[0] = len(keys)
        wire[1 : 1 + len(keys)] = torch.tensor([int(k) for k in keys], dtype=torch.int64)
        wire[257 : 257 + len(keys)] = torch.tensor([int(freqs.freq_dict[k]) for k in keys], dtype=torch.int64)
    dist.broadcast(wire, src=src)
    w = wire.cpu().tolist()
    n = w[0]
    return Frequencies({int(k): int(f) for k, f in zip(w[1 : 1 + n], w[257 : 257 + n])})


def gather_compressed_sizes(local_bytes: int):
    """All-gather of per-rank compressed byte counts -> (list per rank, this rank's global byte offset)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [int(local_bytes)], 0
    dev = _comm_device()
    mine = torch.tensor([int(local_bytes)], dtype=torch.int64, device=dev)
    out = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(dist.get_world_size())]
    dist.all_gather(out, mine)
    sizes = [int(t.item()) for t in out]
    return sizes, sum(sizes[: dist.get_rank()])
