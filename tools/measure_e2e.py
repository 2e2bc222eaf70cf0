"""End-to-end legs of HostCodecPipeline timed separately, for several chunk sizes / ring depths.
python tools/measure_e2e.py [--blocks 262144]   -> one JSON line per setting"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stanford_compression_library_b200.compressors.rANS import rANSDecoder, rANSEncoder, rANSParams  # noqa: E402
from stanford_compression_library_b200.pipeline import HostCodecPipeline  # noqa: E402
from stanford_compression_library_b200.workloads import sample_blocks, zipf_frequencies, zipf_probabilities  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=262144)
    ap.add_argument("--iters", type=int, default=4)
    args = ap.parse_args()
    B, N = args.blocks, 4096
    torch.cuda.set_device(0)
    params = rANSParams(zipf_frequencies())
    enc, dec = rANSEncoder(params), rANSDecoder(params)
    data = sample_blocks(zipf_probabilities(), B, N, seed=0, device="cuda:0")
    host_in = torch.empty((B, N), dtype=torch.uint8, pin_memory=True)
    host_in.copy_(data)
    host_out = torch.empty((B, N), dtype=torch.uint8, pin_memory=True)
    host_c = torch.empty(B * N, dtype=torch.uint8, pin_memory=True)
    del data
    # kernel time of one chunk-sized launch (what the pipeline's compute stream runs per chunk)
    for nb in (8192, 16384, 32768, 65536):
        d = torch.empty((nb, N), dtype=torch.uint8, device="cuda:0").copy_(host_in[:nb])
        p = enc.encode_blocks_packed(d)
        o = dec.decode_blocks(p, N)
        ts = []
        for fn in (lambda: enc.encode_blocks_packed(d, reuse=p), lambda: dec.decode_blocks(p, N, reuse=o)):
            fn()
            torch.cuda.synchronize()
            best = 1e9
            for _ in range(10):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            ts.append(best)
        print(json.dumps(dict(chunk_blocks=nb, encode_packed_kernel_ms=ts[0], decode_kernel_ms=ts[1], h2d_ms_at_55GBps=nb * N / 55e6)), flush=True)
        del d, p, o
    for chunk, depth in ((65536, 3), (32768, 3), (16384, 3), (16384, 4), (8192, 4)):
        pipe = HostCodecPipeline(enc, dec, N, B, chunk_blocks=chunk, depth=depth)
        total, lens = pipe.encode(host_in, host_c)
        pipe.decode(host_c, lens, host_out)
        torch.cuda.synchronize()
        assert torch.equal(host_out, host_in)
        te = td = 1e9
        for _ in range(args.iters):
            t0 = time.perf_counter()
            total, lens = pipe.encode(host_in, host_c)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            pipe.decode(host_c, lens, host_out)
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            te, td = min(te, t1 - t0), min(td, t2 - t1)
        raw = B * N
        print(json.dumps(dict(chunk=chunk, depth=depth, encode_ms=te * 1e3, decode_ms=td * 1e3, roundtrip_GBps=raw / (te + td) / 1e9,
                              encode_leg_h2d_GBps=raw / te / 1e9, decode_leg_d2h_GBps=raw / td / 1e9, coded_bytes=total)), flush=True)
        del pipe
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
