"""Throughput of every coder kernel on the BASELINE config shapes (informative; bench.py is the
contract).  python tools/measure_coders.py [--blocks-scale 1.0]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stanford_compression_library_b200 import Frequencies  # noqa: E402
from stanford_compression_library_b200.compressors.arithmetic_coding import AECParams, ArithmeticDecoder, ArithmeticEncoder  # noqa: E402
from stanford_compression_library_b200.compressors.probability_models import AdaptiveIIDFreqModel, AdaptiveOrderKFreqModel  # noqa: E402
from stanford_compression_library_b200.compressors.rANS import rANSDecoder, rANSEncoder, rANSParams  # noqa: E402
from stanford_compression_library_b200.compressors.range_coder import RangeCoderParams, RangeDecoder, RangeEncoder  # noqa: E402
from stanford_compression_library_b200.compressors.tANS import tANSDecoder, tANSEncoder, tANSParams  # noqa: E402
from stanford_compression_library_b200.workloads import sample_blocks, zipf_frequencies, zipf_probabilities  # noqa: E402


def timeit(fn, iters=5):
    fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    return min(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))


def run(name, enc, dec, data):
    B, N = data.shape
    e = enc.encode_blocks(data).check()
    d = dec.decode_blocks(e, N).check()
    assert torch.equal(d.symbols[:, :N], data)
    te = timeit(lambda: enc.encode_blocks(data, reuse=e))
    td = timeit(lambda: dec.decode_blocks(e, N, reuse=d))
    raw, C = B * N, e.total_bytes()
    res = dict(coder=name, blocks=B, block_len=N, bits_per_symbol=8 * C / raw, encode_ms=te, decode_ms=td, encode_GBps=raw / te / 1e6, decode_GBps=raw / td / 1e6,
               encode_roofline_frac=(raw + C) / te / 1e6 / 6548.5, decode_roofline_frac=(raw + C) / td / 1e6 / 6548.5)
    print(json.dumps(res))
    return res


def main():
    torch.cuda.set_device(0)
    fr = zipf_frequencies()
    p = zipf_probabilities()
    d4k = sample_blocks(p, 65536, 4096, seed=0, device="cuda:0")
    d1k = sample_blocks(p, 262144, 1024, seed=1, device="cuda:0")
    from stanford_compression_library_b200.stats import histogram_blocks
    big = sample_blocks(p, 262144, 4096, seed=2, device="cuda:0")
    th = timeit(lambda: histogram_blocks(big, per_block=False))
    tb = timeit(lambda: histogram_blocks(big, total=False))
    print(json.dumps(dict(coder="histogram 262144 x 4 KiB", total_only_ms=th, total_only_GBps=big.numel() / th / 1e6, roofline_frac=big.numel() / th / 1e6 / 6548.5,
                          per_block_ms=tb, per_block_GBps=big.numel() / tb / 1e6)))
    del big
    run("rANS default (cfg2)", rANSEncoder(rANSParams(fr)), rANSDecoder(rANSParams(fr)), d4k)
    # pack / frame kernels on the cfg2 batch's encoder output: bytes moved = 2 x coded bytes
    e = rANSEncoder(rANSParams(fr)).encode_blocks(d4k).check()
    C = e.total_bytes()
    tp = timeit(lambda: e.pack())
    tf = timeit(lambda: e.frame())
    print(json.dumps(dict(coder="pack / frame of 65536 rANS streams (incl. torch cumsum + allocation)", coded_bytes=C, pack_ms=tp, frame_ms=tf,
                          pack_GBps_read_plus_write=2 * C / tp / 1e6, frame_GBps_read_plus_write=2 * C / tf / 1e6)))
    del e
    run("rANS nbo8 rf4096 (cfg2)", rANSEncoder(rANSParams(fr, NUM_BITS_OUT=8, RANGE_FACTOR=1 << 12)), rANSDecoder(rANSParams(fr, NUM_BITS_OUT=8, RANGE_FACTOR=1 << 12)), d4k)
    tp = tANSParams(fr, RANGE_FACTOR=1)
    run("tANS RF=1 L=4096 (cfg3)", tANSEncoder(tp), tANSDecoder(tp), d4k)
    rp = RangeCoderParams()
    run("range coder (cfg2 shape)", RangeEncoder(rp, fr), RangeDecoder(rp, fr), d4k)
    big = sample_blocks(p, 262144, 4096, seed=2, device="cuda:0")
    run("range coder (262144 x 4 KiB, the bench.py shape)", RangeEncoder(rp, fr), RangeDecoder(rp, fr), big)
    run("tANS RF=1 L=4096 (262144 x 4 KiB, the bench.py shape)", tANSEncoder(tp), tANSDecoder(tp), big)
    del big
    torch.cuda.empty_cache()
    ap = AECParams()
    uni = Frequencies({b: 1 for b in range(256)})
    run("arithmetic adaptive order-0, 262144 x 1 KiB (cfg4 shape / 4)", ArithmeticEncoder(ap, AdaptiveIIDFreqModel(uni, ap.MAX_ALLOWED_TOTAL_FREQ)),
        ArithmeticDecoder(ap, AdaptiveIIDFreqModel(uni, ap.MAX_ALLOWED_TOTAL_FREQ)), d1k)
    # order-k context model on a sticky 4-symbol source
    g = torch.Generator(device="cuda:0").manual_seed(3)
    fresh = torch.randint(0, 4, (262144, 1024), generator=g, device="cuda:0", dtype=torch.uint8)
    hold = torch.rand((262144, 1024), generator=g, device="cuda:0") < 0.7
    run_idx = torch.cummax(torch.where(hold, 0, torch.arange(1024, device="cuda:0")[None, :].expand(262144, -1)), dim=1).values
    sticky = torch.gather(fresh, 1, run_idx)
    for k in (0, 2):
        run("arithmetic order-%d context model, 4 symbols, 262144 x 1 KiB" % k, ArithmeticEncoder(ap, AdaptiveOrderKFreqModel([0, 1, 2, 3], k, ap.MAX_ALLOWED_TOTAL_FREQ)),
            ArithmeticDecoder(ap, AdaptiveOrderKFreqModel([0, 1, 2, 3], k, ap.MAX_ALLOWED_TOTAL_FREQ)), sticky)


if __name__ == "__main__":
    main()
