"""rANS as the entropy stage of LZ77 streams (second half).

The reference's LZ77 codec entropy-codes four integer streams per block -- literal counts, match
lengths, match offsets (log-scale binned) and the literal bytes -- with an *empirical Huffman*
coder (scl/compressors/lz77.py:122-210 `EmpiricalIntHuffmanEncoder/Decoder`, :213-297
`LogScaleBinnedIntegerEncoder/Decoder`, :300-358 / :361-445 `LZ77StreamsEncoder/Decoder`).
This module keeps those classes' constructors, stream layout and bit accounting and swaps the
Huffman stage for this backend's GPU rANS coder:

    [len(counts_encoding) : 32][Elias-delta(counts[0..alphabet_size))]
    [len(values_encoding) : 32][rANSEncoder(rANSParams(Frequencies(counts))).encode_block(values)]

`Frequencies` holds the raw counts of the symbols that occur, in ascending symbol order (what the
reference feeds its Huffman tree, lz77.py:146-148), so each piece of the stream is something the
unmodified reference can produce: tests/golden/lz77_rans_v1.npz is composed from the reference's
own `EliasDeltaUintEncoder`, `rANSEncoder` and LZ77 parser (oracle/gen_golden_lz77.py).

The LZ77 parser itself (match finding) is out of scope; `LZ77Sequence`
is only the record type the streams coder consumes.
"""
import math
from dataclasses import dataclass
from typing import List

import numpy as np

from ..core.data_block import DataBlock
from ..core.data_encoder_decoder import DataDecoder, DataEncoder
from ..core.prob_dist import Frequencies
from ..utils.bitarray_utils import BitArray, bitarray_to_uint, uint_to_bitarray
from .rANS import rANSDecoder, rANSEncoder, rANSParams

ENCODED_BLOCK_SIZE_HEADER_BITS = 32  # lz77.py:108-110


@dataclass
class LZ77Sequence:
    """lz77.py:115-125: copy `literal_count` literals, then `match_length` bytes from `match_offset` back."""

    literal_count: int = 0
    match_length: int = 0
    match_offset: int = 0


# ---- Elias delta (scl/compressors/elias_delta_uint_coder.py:43-121), host side: it codes <= 288 counts ----
class EliasDeltaUintEncoder(DataEncoder):
    def encode_symbol(self, x: int) -> BitArray:
        assert isinstance(x, (int, np.integer)) and x >= 0
        y = bin(int(x) + 1)[2:]            # binary of Y = X + 1
        m = bin(len(y))[2:]                # binary of M = N + 1, N = len(y) - 1
        return BitArray("0" * (len(m) - 1) + m + y[1:])

    def encode_block(self, data_block: DataBlock) -> BitArray:
        return BitArray("".join(self.encode_symbol(s).to01() for s in data_block.data_list))


class EliasDeltaUintDecoder(DataDecoder):
    def decode_symbol(self, encoded_bitarray: BitArray):
        s = encoded_bitarray.to01() if isinstance(encoded_bitarray, BitArray) else encoded_bitarray
        return self._decode_at(s, 0)

    @staticmethod
    def _decode_at(s: str, pos: int):
        one = s.find("1", pos)
        if one < 0:
            raise IndexError("bitarray index out of range")  # what the reference's bit-by-bit scan raises
        l = one - pos
        m = int(s[one : one + l + 1], 2)
        n = m - 1
        p = one + l + 1
        y = int("1" + s[p : p + n], 2) if n else 1
        return y - 1, (p + n) - pos

    def decode_block(self, bitarray: BitArray):
        s = bitarray.to01()
        out, pos = [], 0
        while pos < len(s):
            x, used = self._decode_at(s, pos)
            out.append(x)
            pos += used
        return DataBlock(out), pos


# ---- empirical rANS in the reference's EmpiricalIntHuffman container (lz77.py:122-210) --------------------
class EmpiricalIntRansEncoder(DataEncoder):
    """Values in [0, alphabet_size), alphabet_size <= 256.  `rans_kwargs` are passed to rANSParams."""

    def __init__(self, alphabet_size, **rans_kwargs):
        assert alphabet_size <= 256, "this backend's tables hold at most 256 symbols"
        self.alphabet_size = alphabet_size
        self.rans_kwargs = rans_kwargs

    def encode_block(self, data_block: DataBlock) -> BitArray:
        vals = np.asarray(list(data_block.data_list), dtype=np.int64)
        assert vals.size == 0 or (vals.min() >= 0 and vals.max() < self.alphabet_size)  # lz77.py:141
        if vals.size == 0:
            return uint_to_bitarray(0, ENCODED_BLOCK_SIZE_HEADER_BITS)  # lz77.py:166-168
        counts = np.bincount(vals, minlength=self.alphabet_size)
        freqs = Frequencies({int(i): int(counts[i]) for i in np.nonzero(counts)[0]})  # ascending symbol order (lz77.py:146-148)
        values_encoding = rANSEncoder(rANSParams(freqs, **self.rans_kwargs)).encode_block(DataBlock(vals.tolist()))
        counts_encoding = EliasDeltaUintEncoder().encode_block(DataBlock([int(c) for c in counts]))
        return (uint_to_bitarray(len(counts_encoding), ENCODED_BLOCK_SIZE_HEADER_BITS) + counts_encoding
                + uint_to_bitarray(len(values_encoding), ENCODED_BLOCK_SIZE_HEADER_BITS) + values_encoding)


class EmpiricalIntRansDecoder(DataDecoder):
    def __init__(self, alphabet_size, **rans_kwargs):
        assert alphabet_size <= 256
        self.alphabet_size = alphabet_size
        self.rans_kwargs = rans_kwargs

    def decode_block(self, encoded_bitarray: BitArray):
        H = ENCODED_BLOCK_SIZE_HEADER_BITS
        used = 0
        counts_size = bitarray_to_uint(encoded_bitarray[:H])
        used += H
        if counts_size == 0:
            return DataBlock([]), used  # lz77.py:186-187
        counts, n = EliasDeltaUintDecoder().decode_block(encoded_bitarray[used : used + counts_size])
        assert counts_size == n  # lz77.py:192
        used += counts_size
        counts = counts.data_list
        freqs = Frequencies({i: counts[i] for i in range(self.alphabet_size) if counts[i] > 0})
        values_size = bitarray_to_uint(encoded_bitarray[used : used + H])
        used += H
        decoded, n = rANSDecoder(rANSParams(freqs, **self.rans_kwargs)).decode_block(encoded_bitarray[used : used + values_size])
        assert values_size == n  # lz77.py:207
        used += values_size
        return decoded, used


# ---- log-scale binning (lz77.py:213-297) ---------------------------------------------------------------------
class LogScaleBinnedIntegerEncoder(DataEncoder):
    def __init__(self, offset=0, max_num_bins=32, **rans_kwargs):
        self.offset = offset
        self.max_num_bins = max_num_bins + self.offset
        self.empirical_encoder = EmpiricalIntRansEncoder(alphabet_size=self.max_num_bins, **rans_kwargs)

    def encode_block(self, data_block: DataBlock) -> BitArray:
        bins, pieces = [], []
        for val in data_block.data_list:
            val = int(val)
            assert val >= 0
            if val < self.offset:
                bins.append(val)
                continue
            v1 = val - self.offset + 1
            nb = int(math.log2(v1))  # the reference's float expression, on purpose (lz77.py:248)
            if nb >= self.max_num_bins:
                raise ValueError("Value %d is too large to be encoded with %d bins" % (val - self.offset, self.max_num_bins))
            bins.append(nb + self.offset)
            if nb:
                pieces.append(bin((1 << nb) | (v1 - (1 << nb)))[3:])  # residual v1 - 2^nb in exactly nb bits (lz77.py:255-263)
        bins_encoding = self.empirical_encoder.encode_block(DataBlock(bins))
        return bins_encoding + BitArray("".join(pieces))


class LogScaleBinnedIntegerDecoder(DataDecoder):
    def __init__(self, offset=0, max_num_bins=32, **rans_kwargs):
        self.offset = offset
        self.max_num_bins = max_num_bins + self.offset
        self.empirical_decoder = EmpiricalIntRansDecoder(alphabet_size=self.max_num_bins, **rans_kwargs)

    def decode_block(self, encoded_bitarray: BitArray):
        bins, used = self.empirical_decoder.decode_block(encoded_bitarray)
        s = encoded_bitarray[used:].to01()
        pos = 0
        out = []
        for b in bins.data_list:
            if b < self.offset:
                out.append(b)
                continue
            nb = b - self.offset
            if nb and pos + nb > len(s):
                raise ValueError("non-empty bitarray expected")  # bitarray_to_uint of a short slice in the reference
            residual = int(s[pos : pos + nb], 2) if nb else 0
            pos += nb
            out.append(self.offset + (1 << nb) + residual - 1)
        return DataBlock(out), used + pos


# ---- the four streams of one LZ77 block (lz77.py:300-445) ---------------------------------------------------------
class LZ77StreamsEncoder(DataEncoder):
    def __init__(self, log_scale_binned_coder_offset=16, **rans_kwargs):
        self.log_scale_binned_coder_offset = log_scale_binned_coder_offset
        self.rans_kwargs = rans_kwargs

    def encode_lz77_sequences(self, lz77_sequences: List[LZ77Sequence]) -> BitArray:
        coder = LogScaleBinnedIntegerEncoder(offset=self.log_scale_binned_coder_offset, **self.rans_kwargs)
        out = BitArray()
        out += coder.encode_block(DataBlock([s.literal_count for s in lz77_sequences]))
        out += coder.encode_block(DataBlock([s.match_length for s in lz77_sequences]))
        out += coder.encode_block(DataBlock([s.match_offset for s in lz77_sequences]))
        return out

    def encode_literals(self, literals: List) -> BitArray:
        return EmpiricalIntRansEncoder(alphabet_size=256, **self.rans_kwargs).encode_block(DataBlock(list(literals)))

    def encode_block(self, lz77_sequences: List[LZ77Sequence], literals: List) -> BitArray:
        return self.encode_lz77_sequences(lz77_sequences) + self.encode_literals(literals)


class LZ77StreamsDecoder(DataDecoder):
    def __init__(self, log_scale_binned_coder_offset=16, **rans_kwargs):
        self.log_scale_binned_coder_offset = log_scale_binned_coder_offset
        self.rans_kwargs = rans_kwargs

    def decode_lz77_sequences(self, encoded_bitarray: BitArray):
        coder = LogScaleBinnedIntegerDecoder(offset=self.log_scale_binned_coder_offset, **self.rans_kwargs)
        used = 0
        streams = []
        for _ in range(3):  # literal counts, match lengths, match offsets
            vals, n = coder.decode_block(encoded_bitarray[used:])
            streams.append(vals.data_list)
            used += n
        return [LZ77Sequence(a, b, c) for a, b, c in zip(*streams)], used

    def decode_literals(self, encoded_bitarray: BitArray):
        literals, n = EmpiricalIntRansDecoder(alphabet_size=256, **self.rans_kwargs).decode_block(encoded_bitarray)
        return literals.data_list, n

    def decode_block(self, encoded_bitarray: BitArray):
        seqs, n1 = self.decode_lz77_sequences(encoded_bitarray)
        literals, n2 = self.decode_literals(encoded_bitarray[n1:])
        return (seqs, literals), n1 + n2
