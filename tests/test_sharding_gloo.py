"""CPU tier: the N>1 host logic (block-range sharding, frequency-table broadcast, size all-gather)
with world_size=2 over gloo.  No GPU: the ranks exchange the table, derive identical parameters and
each "encodes" its shard with the C oracle so that the concatenation order can be checked."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist

    from oracle import scl_oracle as so
    from stanford_compression_library_b200.compressors.rANS import rANSParams
    from stanford_compression_library_b200.sharding import broadcast_frequencies, gather_compressed_sizes, shard_range
    from stanford_compression_library_b200.workloads import zipf_frequencies

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        freqs = broadcast_frequencies(zipf_frequencies() if rank == 0 else None)
        params = rANSParams(freqs, NUM_BITS_OUT=8, RANGE_FACTOR=1 << 12)
        n_blocks = 37
        lo, hi = shard_range(n_blocks, rank, world)
        rng = np.random.default_rng(0)  # every rank sees the same global stream, encodes only its shard
        data = rng.integers(0, 256, size=(n_blocks, 64)).astype(np.uint8)
        oracle = so.Oracle.rans([int(f) for f in freqs.freq_list], NUM_BITS_OUT=8, RANGE_FACTOR=1 << 12)
        out, bits, st = oracle.encode_batch(data[lo:hi], out_stride=256)
        local_bytes = int(((bits + 7) // 8).sum())
        sizes, my_off = gather_compressed_sizes(local_bytes)
        q.put((rank, list(freqs.freq_dict.items())[:4], int(params.NUM_STATE_BITS), (lo, hi), local_bytes, sizes, my_off, [int(b) for b in bits]))
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions_exactly():
    from stanford_compression_library_b200.sharding import shard_range

    for n in (0, 1, 7, 8, 37, 2097152):
        for w in (1, 2, 3, 8):
            parts = [shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in parts) - min(h - l for l, h in parts) <= 1
    assert shard_range(2097152, 3, 8) == (786432, 1048576)  # BASELINE cfg5: 262144 blocks per GPU


@pytest.mark.timeout(180)
def test_broadcast_and_sharding_world2_gloo():
    from oracle import scl_oracle as so

    so.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=150) for _ in procs)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    r0, r1 = res
    assert r0[1] == r1[1] and r0[2] == r1[2] == 32  # identical table and derived parameters on both ranks
    assert r0[3] == (0, 19) and r1[3] == (19, 37)
    assert r0[5] == r1[5] == [r0[4], r1[4]]  # all-gathered sizes agree
    assert r0[6] == 0 and r1[6] == r0[4]  # global byte offsets of the concatenated stream
    # concatenation in rank order == single-process encode of the whole stream
    from stanford_compression_library_b200.workloads import zipf_freq_list

    rng = np.random.default_rng(0)
    data = rng.integers(0, 256, size=(37, 64)).astype(np.uint8)
    _, bits, _ = so.Oracle.rans(zipf_freq_list(), NUM_BITS_OUT=8, RANGE_FACTOR=1 << 12).encode_batch(data, out_stride=256)
    assert r0[7] + r1[7] == [int(b) for b in bits]
