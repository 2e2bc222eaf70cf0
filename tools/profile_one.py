"""Run one coder a few times on a fixed shape (for ncu captures): python tools/profile_one.py aec|range|tans|rans [blocks] [len]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.measure_coders import *  # noqa: F401,F403,E402

kind = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
N = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
torch.cuda.set_device(0)
fr, p = zipf_frequencies(), zipf_probabilities()
data = sample_blocks(p, B, N, seed=0, device="cuda:0")
if kind == "aec":
    ap = AECParams()
    uni = Frequencies({b: 1 for b in range(256)})
    enc = ArithmeticEncoder(ap, AdaptiveIIDFreqModel(uni, ap.MAX_ALLOWED_TOTAL_FREQ))
    dec = ArithmeticDecoder(ap, AdaptiveIIDFreqModel(uni, ap.MAX_ALLOWED_TOTAL_FREQ))
elif kind == "range":
    enc, dec = RangeEncoder(RangeCoderParams(), fr), RangeDecoder(RangeCoderParams(), fr)
elif kind == "tans":
    tp = tANSParams(fr, RANGE_FACTOR=1)
    enc, dec = tANSEncoder(tp), tANSDecoder(tp)
else:
    enc, dec = rANSEncoder(rANSParams(fr)), rANSDecoder(rANSParams(fr))
for _ in range(2):
    e = enc.encode_blocks(data).check()
    d = dec.decode_blocks(e, N).check()
torch.cuda.synchronize()
assert torch.equal(d.symbols[:, :N], data)
print("ok", kind, B, N)
