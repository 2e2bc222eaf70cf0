"""Does one launch over 2M blocks cost more per block than the same work in launches of 2 rounds each?
(bench at 8 GiB: decode 10.1 ms vs 8 x 0.886 ms.)  Diagnostic.  python tools/measure_chunking.py [--blocks N]"""
import argparse
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stanford_compression_library_b200 import _cabi  # noqa: E402
from stanford_compression_library_b200.compressors.rANS import rANSDecoder, rANSEncoder, rANSParams  # noqa: E402
from stanford_compression_library_b200.device import DecodedBlocks, EncodedBlocks  # noqa: E402
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
    ap.add_argument("--blocks", type=int, default=2097152)
    a = ap.parse_args()
    torch.cuda.set_device(0)
    B, N = a.blocks, 4096
    prm = rANSParams(zipf_frequencies())
    enc, dec = rANSEncoder(prm), rANSDecoder(prm)
    data = sample_blocks(zipf_probabilities(), B, N, seed=0, device="cuda:0")
    e = enc.encode_blocks(data).check()
    d = dec.decode_blocks(e, N).check()
    assert torch.equal(d.symbols[:, :N], data)
    stride = e.out_stride
    abs_off = e.bit_offset.clone()  # the chunked encodes below rewrite e.bit_offset relative to each chunk
    out = {"blocks": B}
    for chunk in (B, 1048576, 524288, 262144, 131072):
        if chunk > B:
            continue

        def enc_chunks():
            for lo in range(0, B, chunk):
                hi = min(B, lo + chunk)
                view = EncodedBlocks(e.buf[lo * stride : hi * stride + 16], e.bit_offset[lo:hi], e.bit_len[lo:hi], e.status[lo:hi], stride)
                enc.encode_blocks(data[lo:hi], reuse=view)

        def dec_chunks():
            for lo in range(0, B, chunk):
                hi = min(B, lo + chunk)
                # offsets are relative to the sub-buffer handed to the call
                view = EncodedBlocks(e.buf[lo * stride : hi * stride + 16], rel_off[lo:hi], e.bit_len[lo:hi], None, stride)
                dv = DecodedBlocks(d.symbols[lo:hi], d.sizes[lo:hi], d.bits_consumed[lo:hi], d.status[lo:hi])
                dec.decode_blocks(view, N, reuse=dv)

        te = timeit(enc_chunks)
        e.check()
        rel_off = abs_off - (torch.arange(B, device="cuda:0", dtype=torch.int64) // chunk) * (chunk * stride * 8)
        d.symbols.zero_()
        td = timeit(dec_chunks)
        assert torch.equal(d.symbols[:, :N], data)
        out["launches_of_%d" % chunk] = {"encode_slots_ms": te, "decode_ms": td, "encode_ms_per_GiB": te / (B * N / 2**30), "decode_ms_per_GiB": td / (B * N / 2**30)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
