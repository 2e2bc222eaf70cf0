"""Cost of the contiguous output: slot encode, slot encode + scan + pack kernel, and the fused packed encode;
decode from slots and from the packed buffer.  Informative (bench.py is the contract).
    python tools/measure_packed.py [--blocks 262144] [--block-len 4096]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stanford_compression_library_b200.compressors.rANS import rANSDecoder, rANSEncoder, rANSParams  # noqa: E402
from stanford_compression_library_b200.compressors.tANS import tANSDecoder, tANSEncoder, tANSParams  # noqa: E402
from stanford_compression_library_b200.workloads import sample_blocks, zipf_frequencies, zipf_probabilities  # noqa: E402

PEAK = 6548.5
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timeit(fn, iters=7):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    t = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
    return t[0], t[len(t) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=262144)
    ap.add_argument("--block-len", type=int, default=4096)
    a = ap.parse_args()
    torch.cuda.set_device(0)
    fr = zipf_frequencies()
    data = sample_blocks(zipf_probabilities(), a.blocks, a.block_len, seed=0, device="cuda:0")
    B, N = data.shape
    raw = B * N
    for name, mk in (("rans_default", lambda: rANSParams(fr)), ("rans_nbo8_rf4096", lambda: rANSParams(fr, NUM_BITS_OUT=8, RANGE_FACTOR=1 << 12)),
                     ("tans_L4096", lambda: tANSParams(fr, RANGE_FACTOR=1))):
        prm = mk()
        enc, dec = (tANSEncoder(prm), tANSDecoder(prm)) if name.startswith("tans") else (rANSEncoder(prm), rANSDecoder(prm))
        e = enc.encode_blocks(data).check()
        p = enc.encode_blocks_packed(data).check()
        want = e.pack()
        C = int(p.byte_offset[-1])
        assert C == e.total_bytes() and torch.equal(p.buf[:C], want.buf[:C]), "fused packed output differs from encode + pack"
        d = dec.decode_blocks(e, N).check()
        d2 = dec.decode_blocks(p, N).check()
        assert torch.equal(d.symbols[:, :N], data) and torch.equal(d2.symbols[:, :N], data)
        del want
        t_enc = timeit(lambda: enc.encode_blocks(data, reuse=e))
        t_pack = timeit(lambda: e.pack())
        t_fused = timeit(lambda: enc.encode_blocks_packed(data, reuse=p))
        t_dec_slots = timeit(lambda: dec.decode_blocks(e, N, reuse=d))
        t_dec_packed = timeit(lambda: dec.decode_blocks(p, N, reuse=d2))
        alg = raw + C
        print(json.dumps({
            "coder": name, "blocks": B, "block_len": N, "coded_bytes": C,
            "encode_slots_ms": t_enc, "pack_incl_scan_and_alloc_ms": t_pack, "encode_packed_fused_ms": t_fused,
            "decode_from_slots_ms": t_dec_slots, "decode_from_packed_ms": t_dec_packed,
            "frac_encode_slots": alg / t_enc[0] / 1e6 / PEAK, "frac_encode_packed_fused": alg / t_fused[0] / 1e6 / PEAK,
            "frac_encode_then_pack": alg / (t_enc[0] + t_pack[0]) / 1e6 / PEAK,
            "frac_decode_slots": alg / t_dec_slots[0] / 1e6 / PEAK, "frac_decode_packed": alg / t_dec_packed[0] / 1e6 / PEAK,
            "note": "[best, median] ms; fractions = (raw + coded bytes) / best / %.1f GB/s" % PEAK}))
        del e, p, d, d2
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
