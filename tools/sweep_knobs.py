"""Experiment knobs of the second-generation rANS kernels (scl_coder_debug_path high bits).  The CTA round barrier and
the L1 prefetch this tool was written for measured no effect (profiles/r2f_sweep.jsonl) and are gone from the kernels;
what is left are the variants of the packed encoder's copy pool (staging rings on / off, piece size, copy warps).  Diagnostic; one JSON line per (blocks, mode).
    python tools/sweep_knobs.py [--blocks 262144 2097152]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stanford_compression_library_b200.compressors.rANS import rANSDecoder, rANSEncoder, rANSParams  # noqa: E402
from stanford_compression_library_b200.workloads import sample_blocks, zipf_frequencies, zipf_probabilities  # noqa: E402

NO_RING, PIECE_512, PIECE_1024 = 32, 64, 128  # scl_coder_debug_path bits; bits 12-15 = copy warps


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
    ap.add_argument("--blocks", type=int, nargs="+", default=[262144, 2097152])
    a = ap.parse_args()
    torch.cuda.set_device(0)
    N = 4096
    prm = rANSParams(zipf_frequencies())
    enc, dec = rANSEncoder(prm), rANSDecoder(prm)
    for B in a.blocks:
        data = sample_blocks(zipf_probabilities(), B, N, seed=0, device="cuda:0")
        e = enc.encode_blocks(data).check()
        p = enc.encode_blocks_packed(data, capacity=B * N).check()
        d = dec.decode_blocks(p, N).check()
        ref = p.buf[: int(p.byte_offset[-1])].clone()
        gib = B * N / 2**30
        for mode in (0,):
            enc.device_coder().debug_path(mode)
            dec.device_coder().debug_path(mode)
            te = timeit(lambda: enc.encode_blocks(data, reuse=e))
            td = timeit(lambda: dec.decode_blocks(p, N, reuse=d))
            assert torch.equal(d.symbols[:, :N], data)
            print(json.dumps({"blocks": B, "mode": mode, "encode_slots_ms_per_GiB": te / gib, "decode_ms_per_GiB": td / gib}), flush=True)
        for mode in (0, 2 << 12, 3 << 12, 5 << 12):
            enc.device_coder().debug_path(mode)
            tp = timeit(lambda: enc.encode_blocks_packed(data, capacity=B * N, reuse=p))
            p.check()
            assert torch.equal(p.buf[: ref.numel()], ref)
            print(json.dumps({"blocks": B, "mode": mode, "encode_packed_ms_per_GiB": tp / gib}), flush=True)
        enc.device_coder().debug_path(0)
        dec.device_coder().debug_path(0)
        del data, e, p, d, ref
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
