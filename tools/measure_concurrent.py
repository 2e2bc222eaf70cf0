"""Small batches (BASELINE configs[1]: 65 536 blocks = one half-filled round of the persistent grid, 14 warps per SM,
latency-bound) on TWO streams: does the encode of batch k+1 share the SMs with the decode of batch k?
Each kernel is one CTA per SM with ~110 KB of shared memory at this size, so both fit on an SM at once.
    python tools/measure_concurrent.py [--blocks 65536]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stanford_compression_library_b200.compressors.rANS import rANSDecoder, rANSEncoder, rANSParams  # noqa: E402
from stanford_compression_library_b200.workloads import sample_blocks, zipf_frequencies, zipf_probabilities  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=65536)
    ap.add_argument("--pairs", type=int, default=8)
    ap.add_argument("--mode", type=int, default=0, help="scl_coder_debug_path of the encoder (32 = no staging rings: less shared memory)")
    a = ap.parse_args()
    torch.cuda.set_device(0)
    B, N = a.blocks, 4096
    prm = rANSParams(zipf_frequencies())
    enc, dec = rANSEncoder(prm), rANSDecoder(prm)
    enc.device_coder().debug_path(a.mode)
    A = sample_blocks(zipf_probabilities(), B, N, seed=1, device="cuda:0")
    Bt = sample_blocks(zipf_probabilities(), B, N, seed=2, device="cuda:0")
    pa = enc.encode_blocks_packed(A).check()
    pb = enc.encode_blocks_packed(Bt).check()
    da = dec.decode_blocks(pa, N).check()
    assert torch.equal(da.symbols[:, :N], A)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def sequential():
        for _ in range(a.pairs):
            enc.encode_blocks_packed(Bt, reuse=pb)
            dec.decode_blocks(pa, N, reuse=da)

    def concurrent():
        s1.wait_stream(torch.cuda.current_stream())
        s2.wait_stream(torch.cuda.current_stream())
        for _ in range(a.pairs):
            with torch.cuda.stream(s1):
                enc.encode_blocks_packed(Bt, reuse=pb)
            with torch.cuda.stream(s2):
                dec.decode_blocks(pa, N, reuse=da)
        torch.cuda.current_stream().wait_stream(s1)
        torch.cuda.current_stream().wait_stream(s2)

    out = {"blocks": B, "encoder_mode": a.mode, "pairs_per_measurement": a.pairs}
    for name, fn in (("sequential_one_stream", sequential), ("two_streams", concurrent)):
        fn()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out[name + "_ms_per_pair"] = best / a.pairs
        out[name + "_GBps_raw_bytes_coded_plus_decoded"] = 2 * B * N / (best / a.pairs) / 1e6
    pb.check()
    assert torch.equal(da.symbols[:, :N], A)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
