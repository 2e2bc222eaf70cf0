"""One encode + decode of the adaptive arithmetic coder (cfg4 shape / 4) for ncu:
    ncu --set full --clock-control none --import-source on -k regex:aec -o out python tools/profile_aec.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stanford_compression_library_b200 import Frequencies  # noqa: E402
from stanford_compression_library_b200.compressors.arithmetic_coding import AECParams, ArithmeticDecoder, ArithmeticEncoder  # noqa: E402
from stanford_compression_library_b200.compressors.probability_models import AdaptiveIIDFreqModel  # noqa: E402
from stanford_compression_library_b200.workloads import sample_blocks, zipf_probabilities  # noqa: E402

torch.cuda.set_device(0)
B, N = 262144, 1024
prm = AECParams()
uni = Frequencies({b: 1 for b in range(256)})
enc = ArithmeticEncoder(prm, AdaptiveIIDFreqModel(uni, prm.MAX_ALLOWED_TOTAL_FREQ))
dec = ArithmeticDecoder(prm, AdaptiveIIDFreqModel(uni, prm.MAX_ALLOWED_TOTAL_FREQ))
data = sample_blocks(zipf_probabilities(), B, N, seed=0, device="cuda:0")
e = enc.encode_blocks(data).check()
d = dec.decode_blocks(e, N).check()
assert torch.equal(d.symbols[:, :N], data)
print("ok")
