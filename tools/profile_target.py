"""One launch of each second-generation rANS kernel at a given batch size, for ncu:
    ncu --set full --clock-control none --import-source on -k regex:"fast_|pack_v2|range_" -o out python tools/profile_target.py --blocks 262144 --with-pack
Launch order: encode (slots), encode (packed, fused), decode (from the packed stream).  --repeat N repeats the triple."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stanford_compression_library_b200.compressors.rANS import rANSDecoder, rANSEncoder, rANSParams  # noqa: E402
from stanford_compression_library_b200.compressors.tANS import tANSDecoder, tANSEncoder, tANSParams  # noqa: E402
from stanford_compression_library_b200.workloads import sample_blocks, zipf_frequencies, zipf_probabilities  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=262144)
    ap.add_argument("--block-len", type=int, default=4096)
    ap.add_argument("--coder", default="rans", choices=["rans", "rans_nbo8", "tans", "range"])
    ap.add_argument("--with-pack", action="store_true", help="also run the standalone compaction / framing kernels on the slot output")
    ap.add_argument("--repeat", type=int, default=1)
    ap.add_argument("--py-chunks", type=int, default=1)
    a = ap.parse_args()
    torch.cuda.set_device(0)
    fr = zipf_frequencies()
    if a.coder == "tans":
        prm = tANSParams(fr, RANGE_FACTOR=1)
        enc, dec = tANSEncoder(prm), tANSDecoder(prm)
    elif a.coder == "range":
        from stanford_compression_library_b200.compressors.range_coder import RangeCoderParams, RangeDecoder, RangeEncoder

        prm = RangeCoderParams()
        enc, dec = RangeEncoder(prm, fr), RangeDecoder(prm, fr)
    else:
        prm = rANSParams(fr) if a.coder == "rans" else rANSParams(fr, NUM_BITS_OUT=8, RANGE_FACTOR=1 << 12)
        enc, dec = rANSEncoder(prm), rANSDecoder(prm)
    B, N = a.blocks, a.block_len
    data = sample_blocks(zipf_probabilities(), B, N, seed=0, device="cuda:0")
    if a.py_chunks > 1:  # the same batch as `py_chunks` separate calls on views (tools/measure_chunking.py under ncu)
        from stanford_compression_library_b200.device import DecodedBlocks, EncodedBlocks

        e = enc.encode_blocks(data)
        d = dec.decode_blocks(e, N)
        torch.cuda.synchronize()
        abs_off, stride, chunk = e.bit_offset.clone(), e.out_stride, B // a.py_chunks
        for lo in range(0, B, chunk):
            hi = min(B, lo + chunk)
            view = EncodedBlocks(e.buf[lo * stride : hi * stride + 16], e.bit_offset[lo:hi], e.bit_len[lo:hi], e.status[lo:hi], stride)
            enc.encode_blocks(data[lo:hi], reuse=view)
        rel = abs_off - (torch.arange(B, device="cuda:0", dtype=torch.int64) // chunk) * (chunk * stride * 8)
        d.symbols.zero_()
        for lo in range(0, B, chunk):
            hi = min(B, lo + chunk)
            view = EncodedBlocks(e.buf[lo * stride : hi * stride + 16], rel[lo:hi], e.bit_len[lo:hi], None, stride)
            dec.decode_blocks(view, N, reuse=DecodedBlocks(d.symbols[lo:hi], d.sizes[lo:hi], d.bits_consumed[lo:hi], d.status[lo:hi]))
        torch.cuda.synchronize()
        assert torch.equal(d.symbols[:, :N], data)
        print("ok (python-level chunks)", B, a.py_chunks)
        return
    e = p = d = None
    for _ in range(a.repeat):
        e = enc.encode_blocks(data, reuse=e)
        p = enc.encode_blocks_packed(data, capacity=B * N, reuse=p)
        d = dec.decode_blocks(p, N, reuse=d)
    if a.with_pack:
        e.pack()
        e.frame()
    torch.cuda.synchronize()
    e.check(), p.check(), d.check()
    assert torch.equal(d.symbols[:, :N], data)
    print("ok", B, N, a.coder, int(p.byte_offset[-1]))


if __name__ == "__main__":
    main()
