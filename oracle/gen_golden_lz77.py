"""Generate tests/golden/lz77_rans_v1.npz: the reference's LZ77 stream container with its Huffman
stage swapped for the reference's own rANS coder.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python oracle/gen_golden_lz77.py

Everything is produced by UNMODIFIED reference code: `EliasDeltaUintEncoder`
(elias_delta_uint_coder.py:43-72), `rANSEncoder` (rANS.py:186-210), `LogScaleBinnedIntegerEncoder`
(lz77.py:213-266), `LZ77StreamsEncoder` (lz77.py:300-358) and the `LZ77Encoder` parser
(lz77.py:525-603) that supplies realistic sequences.  The only substitution is the one the
product makes: the name `EmpiricalIntHuffmanEncoder` inside scl.compressors.lz77 is bound to a
class with the same container layout (lz77.py:160-165) whose value stream comes from
`rANSEncoder(rANSParams(Frequencies(counts)))` instead of `HuffmanEncoder`.
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_loader import import_reference  # noqa: E402

import_reference()
import scl.compressors.lz77 as ref_lz77  # noqa: E402
from scl.compressors.elias_delta_uint_coder import EliasDeltaUintEncoder  # noqa: E402
from scl.compressors.rANS import rANSEncoder, rANSParams  # noqa: E402
from scl.core.data_block import DataBlock  # noqa: E402
from scl.core.data_encoder_decoder import DataEncoder  # noqa: E402
from scl.core.prob_dist import Frequencies  # noqa: E402
from scl.utils.bitarray_utils import uint_to_bitarray  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "lz77_rans_v1.npz")
HDR = ref_lz77.ENCODED_BLOCK_SIZE_HEADER_BITS


class RefEmpiricalIntRansEncoder(DataEncoder):
    """lz77.py:140-168 with `HuffmanEncoder(prob_dist_sorted)` replaced by the reference rANS coder on
    `Frequencies` of the same sorted counts."""

    def __init__(self, alphabet_size):
        self.alphabet_size = alphabet_size

    def encode_block(self, data_block):
        vals = data_block.data_list
        assert all(0 <= v < self.alphabet_size for v in vals)
        counts = DataBlock(vals).get_counts()
        if len(counts) == 0:
            return uint_to_bitarray(0, HDR)
        freqs = Frequencies({i: counts[i] for i in sorted(counts)})
        values_encoding = rANSEncoder(rANSParams(freqs)).encode_block(DataBlock(vals))
        counts_list = [counts.get(i, 0) for i in range(self.alphabet_size)]
        counts_encoding = EliasDeltaUintEncoder().encode_block(DataBlock(counts_list))
        return uint_to_bitarray(len(counts_encoding), HDR) + counts_encoding + uint_to_bitarray(len(values_encoding), HDR) + values_encoding


ref_lz77.EmpiricalIntHuffmanEncoder = RefEmpiricalIntRansEncoder  # the one substitution

cases, arrays = [], {}


def add(kind, note, expected, **inputs):
    i = len(cases)
    meta = {"id": i, "kind": kind, "note": note, "nbits": len(expected)}
    for k, v in inputs.items():
        if isinstance(v, (int, str)):
            meta[k] = v
        else:
            arrays["c%d_%s" % (i, k)] = np.asarray(v, dtype=np.int64)
    arrays["c%d_enc" % i] = np.frombuffer(expected.tobytes(), dtype=np.uint8)
    cases.append(meta)


rng = np.random.default_rng(0)

# 1. Elias delta: the reference's own vectors (elias_delta_uint_coder.py:153-166) + a random block
vals = [0, 1, 3, 4, 5, 100]
add("elias", "elias_delta_uint_coder.py:153-166", EliasDeltaUintEncoder().encode_block(DataBlock(vals)), vals=vals)
vals = [int(v) for v in rng.integers(0, 5000, size=300)]
add("elias", "random < 5000", EliasDeltaUintEncoder().encode_block(DataBlock(vals)), vals=vals)

# 2. empirical rANS container
for n, alpha, note in ((0, 256, "empty"), (1, 256, "single value"), (700, 256, "bytes, skewed"), (500, 48, "bins alphabet"), (64, 5, "tiny alphabet")):
    p = 1.0 / np.arange(1, alpha + 1)
    vals = [int(v) for v in rng.choice(alpha, size=n, p=p / p.sum())]
    add("empirical", note, RefEmpiricalIntRansEncoder(alpha).encode_block(DataBlock(vals)), vals=vals, alphabet_size=alpha)

# 3. log-scale binned integers (reference class, lz77.py:232-266)
for offset, hi, n in ((0, 1000, 200), (16, 70000, 300), (4, 40, 100)):
    vals = [int(v) for v in rng.integers(0, hi, size=n)]
    add("logbin", "offset %d values < %d" % (offset, hi), ref_lz77.LogScaleBinnedIntegerEncoder(offset=offset).encode_block(DataBlock(vals)), vals=vals, offset=offset)

# 4. whole LZ77 blocks: sequences and literals from the reference parser
texts = [
    b"abracadabra abracadabra abracadabra, said the rANS coder to the range coder; " * 6,
    bytes(rng.integers(0, 4, size=1500, dtype=np.uint8)),
    b"A" * 400 + bytes(rng.integers(0, 256, size=300, dtype=np.uint8)) + b"A" * 400,
]
for t in texts:
    seqs, literals = ref_lz77.LZ77Encoder().lz77_parse_and_generate_sequences(DataBlock(list(t)))
    expected = ref_lz77.LZ77StreamsEncoder().encode_block(seqs, literals)
    add("lz77_block", "parser output for %d input bytes: %d sequences, %d literals" % (len(t), len(seqs), len(literals)), expected,
        literal_counts=[s.literal_count for s in seqs], match_lengths=[s.match_length for s in seqs],
        match_offsets=[s.match_offset for s in seqs], literals=list(literals))

arrays["meta"] = np.frombuffer(json.dumps({"cases": cases}).encode(), dtype=np.uint8)
np.savez_compressed(OUT, **arrays)
print("wrote", OUT, "with", len(cases), "cases;", os.path.getsize(OUT), "bytes")
