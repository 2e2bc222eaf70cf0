"""Ragged fast encoder and lane-per-block histogram under compute-sanitizer (small shapes, results checked):
    compute-sanitizer --tool memcheck python tools/sanitize_target2.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stanford_compression_library_b200.compressors.rANS import rANSDecoder, rANSEncoder, rANSParams  # noqa: E402
from stanford_compression_library_b200.compressors.tANS import tANSDecoder, tANSEncoder, tANSParams  # noqa: E402
from stanford_compression_library_b200.stats import histogram_blocks  # noqa: E402
from stanford_compression_library_b200.workloads import sample_blocks, zipf_frequencies, zipf_probabilities  # noqa: E402

torch.cuda.set_device(0)
fr = zipf_frequencies()
g = torch.Generator(device="cuda:0")
g.manual_seed(7)
for name in ("rans", "tans"):
    prm = rANSParams(fr) if name == "rans" else tANSParams(fr, RANGE_FACTOR=1)
    enc, dec = (rANSEncoder(prm), rANSDecoder(prm)) if name == "rans" else (tANSEncoder(prm), tANSDecoder(prm))
    B, N = 148 * 3 * 32 + 45, 320
    data = sample_blocks(zipf_probabilities(), B, N, seed=3, device="cuda:0")
    sizes = torch.randint(0, N + 1, (B,), generator=g, device="cuda:0", dtype=torch.int32)
    sizes[:5] = torch.tensor([0, 1, 63, 64, N], dtype=torch.int32, device="cuda:0")
    e = enc.encode_blocks(data, sizes=sizes).check()
    enc.device_coder().debug_path(1)
    e1 = enc.encode_blocks(data, sizes=sizes).check()
    enc.device_coder().debug_path(0)
    assert torch.equal(e.bit_len, e1.bit_len) and torch.equal(e.pack().buf[: e1.total_bytes()], e1.pack().buf[: e1.total_bytes()])
    for framed in (False, True):
        p = enc.encode_blocks_packed(data, sizes=sizes, framed=framed).check()
        d = dec.decode_blocks(p, N).check()
        assert torch.equal(d.sizes, sizes)
B, N = 148 * 7 * 32 + 99, 160
data = sample_blocks(zipf_probabilities(), B, N, seed=4, device="cuda:0")
sizes = torch.randint(0, N + 1, (B,), generator=g, device="cuda:0", dtype=torch.int32)
for sz in (None, sizes):
    n = torch.full((B,), N, device="cuda:0", dtype=torch.int64) if sz is None else sz.to(torch.int64)
    mask = torch.arange(N, device="cuda:0")[None, :] < n[:, None]
    keys = (torch.arange(B, device="cuda:0", dtype=torch.int64)[:, None] * 256 + data.to(torch.int64))[mask]
    ref = torch.bincount(keys, minlength=B * 256).reshape(B, 256)
    counts, tot = histogram_blocks(data, sizes=sz)
    assert torch.equal(counts.to(torch.int64), ref) and torch.equal(tot, ref.sum(0))
    _, t2 = histogram_blocks(data, sizes=sz, per_block=False)
    assert torch.equal(t2, ref.sum(0))
torch.cuda.synchronize()
print("sanitize_target2: ok")
