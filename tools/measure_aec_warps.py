"""Warps per CTA of the 8-bit-counter arithmetic coder (scl_coder_debug_path bits 16-20): residency experiment.
    python tools/measure_aec_warps.py [--blocks 262144 1048576]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stanford_compression_library_b200 import Frequencies  # noqa: E402
from stanford_compression_library_b200.compressors.arithmetic_coding import AECParams, ArithmeticDecoder, ArithmeticEncoder  # noqa: E402
from stanford_compression_library_b200.compressors.probability_models import AdaptiveIIDFreqModel  # noqa: E402
from stanford_compression_library_b200.workloads import sample_blocks, zipf_probabilities  # noqa: E402


def timeit(fn, iters=3):
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
    ap.add_argument("--blocks", type=int, nargs="+", default=[262144, 1048576])
    a = ap.parse_args()
    torch.cuda.set_device(0)
    N = 1024
    prm = AECParams()
    uni = Frequencies({b: 1 for b in range(256)})
    enc = ArithmeticEncoder(prm, AdaptiveIIDFreqModel(uni, prm.MAX_ALLOWED_TOTAL_FREQ))
    dec = ArithmeticDecoder(prm, AdaptiveIIDFreqModel(uni, prm.MAX_ALLOWED_TOTAL_FREQ))
    for B in a.blocks:
        data = sample_blocks(zipf_probabilities(), B, N, seed=0, device="cuda:0")
        e = enc.encode_blocks(data).check()
        d = dec.decode_blocks(e, N).check()
        ref = e.buf.clone()
        for w in (4, 6, 8, 11, 16, 20, 24):
            enc.device_coder().debug_path(w << 16)
            dec.device_coder().debug_path(w << 16)
            te = timeit(lambda: enc.encode_blocks(data, reuse=e))
            assert torch.equal(e.buf, ref)
            td = timeit(lambda: dec.decode_blocks(e, N, reuse=d))
            assert torch.equal(d.symbols[:, :N], data)
            print(json.dumps({"blocks": B, "warps_per_cta": w, "encode_ms": te, "decode_ms": td, "encode_GBps": B * N / te / 1e6, "decode_GBps": B * N / td / 1e6}), flush=True)
        enc.device_coder().debug_path(0)
        dec.device_coder().debug_path(0)


if __name__ == "__main__":
    main()
