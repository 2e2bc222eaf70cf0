"""Ragged batches (per-block sizes): the second-generation encoder against the first-generation kernels.
    python tools/measure_ragged.py [--blocks 262144]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stanford_compression_library_b200.compressors.rANS import rANSDecoder, rANSEncoder, rANSParams  # noqa: E402
from stanford_compression_library_b200.workloads import sample_blocks, zipf_frequencies, zipf_probabilities  # noqa: E402


def timeit(fn, iters=5):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=262144)
    a = ap.parse_args()
    torch.cuda.set_device(0)
    B, N = a.blocks, 4096
    prm = rANSParams(zipf_frequencies())
    enc, dec = rANSEncoder(prm), rANSDecoder(prm)
    data = sample_blocks(zipf_probabilities(), B, N, seed=0, device="cuda:0")
    g = torch.Generator(device="cuda:0")
    g.manual_seed(1)
    sizes = torch.randint(N // 2, N + 1, (B,), generator=g, device="cuda:0", dtype=torch.int32)
    raw = int(sizes.sum())
    out = {"blocks": B, "row_bytes": N, "sizes": "uniform in [N/2, N]", "raw_bytes": raw}
    for name, mode in (("second_generation", 0), ("first_generation", 1)):
        enc.device_coder().debug_path(mode)
        e = enc.encode_blocks(data, sizes=sizes).check()
        t = timeit(lambda: enc.encode_blocks(data, sizes=sizes, reuse=e))
        out[name + "_slots_ms"] = t
        out[name + "_slots_GBps"] = raw / t / 1e6
        p = enc.encode_blocks_packed(data, sizes=sizes).check()
        t = timeit(lambda: enc.encode_blocks_packed(data, sizes=sizes, reuse=p))
        out[name + "_packed_ms"] = t
        out[name + "_packed_GBps"] = raw / t / 1e6
    enc.device_coder().debug_path(0)
    d = dec.decode_blocks(p, N).check()
    assert torch.equal(d.sizes, sizes)
    t = timeit(lambda: dec.decode_blocks(p, N, reuse=d))
    out["decode_ms"] = t
    out["decode_GBps"] = raw / t / 1e6
    print(json.dumps(out))


if __name__ == "__main__":
    main()
