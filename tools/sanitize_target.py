"""A small fused packed encode + decode of every copy-pool variant, for compute-sanitizer:
    compute-sanitizer --tool memcheck  python tools/sanitize_target.py
    compute-sanitizer --tool racecheck python tools/sanitize_target.py --small
Checks the output against encode + pack as the tests do (so a sanitizer run is also a parity run)."""
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
    ap.add_argument("--small", action="store_true")
    a = ap.parse_args()
    torch.cuda.set_device(0)
    fr = zipf_frequencies()
    shapes = [(148 * 4 * 32 + 77, 40), (3000, 1300), (700, 5400)] if not a.small else [(148 * 2 * 32 + 5, 24), (600, 2600)]
    for name in ("rans", "tans"):
        prm = rANSParams(fr) if name == "rans" else tANSParams(fr, RANGE_FACTOR=1)
        enc, dec = (rANSEncoder(prm), rANSDecoder(prm)) if name == "rans" else (tANSEncoder(prm), tANSDecoder(prm))
        for mode in (0, 32, 64, 128):
            enc.device_coder().debug_path(mode)
            for framed in (False, True):
                for B, N in shapes:
                    data = sample_blocks(zipf_probabilities(), B, N, seed=B, device="cuda:0")
                    data[::3] &= 0x0F
                    e = enc.encode_blocks(data).check()
                    p = enc.encode_blocks_packed(data, framed=framed).check()
                    total = int(p.byte_offset[-1])
                    want = e.frame()[0] if framed else e.pack().buf
                    assert torch.equal(p.buf[:total], want[:total]), (name, mode, framed, B, N)
                    d = dec.decode_blocks(p, N).check()
                    assert torch.equal(d.symbols[:, :N], data)
            enc.device_coder().debug_path(0)
    torch.cuda.synchronize()
    print("sanitize_target: ok")


if __name__ == "__main__":
    main()
