"""One launch of each second-generation rANS kernel at a given batch size, for ncu:
    ncu --set full --clock-control none --import-source on -k regex:fast_ -o out python tools/profile_target.py --blocks 262144
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
    ap.add_argument("--coder", default="rans", choices=["rans", "rans_nbo8", "tans"])
    ap.add_argument("--repeat", type=int, default=1)
    a = ap.parse_args()
    torch.cuda.set_device(0)
    fr = zipf_frequencies()
    if a.coder == "tans":
        prm = tANSParams(fr, RANGE_FACTOR=1)
        enc, dec = tANSEncoder(prm), tANSDecoder(prm)
    else:
        prm = rANSParams(fr) if a.coder == "rans" else rANSParams(fr, NUM_BITS_OUT=8, RANGE_FACTOR=1 << 12)
        enc, dec = rANSEncoder(prm), rANSDecoder(prm)
    B, N = a.blocks, a.block_len
    data = sample_blocks(zipf_probabilities(), B, N, seed=0, device="cuda:0")
    e = p = d = None
    for _ in range(a.repeat):
        e = enc.encode_blocks(data, reuse=e)
        p = enc.encode_blocks_packed(data, capacity=B * N, reuse=p)
        d = dec.decode_blocks(p, N, reuse=d)
    torch.cuda.synchronize()
    e.check(), p.check(), d.check()
    assert torch.equal(d.symbols[:, :N], data)
    print("ok", B, N, a.coder, int(p.byte_offset[-1]))


if __name__ == "__main__":
    main()
