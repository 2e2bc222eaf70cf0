"""Carry-less ("Russian") byte range coder on the GPU behind the reference's RangeCoderParams /
RangeEncoder / RangeDecoder API (scl/compressors/range_coder.py:55-317).

Stream layout: [size : DATA_BLOCK_SIZE_BITS][bytes released by normalize ...][PRECISION/8 flush bytes].
"""
from dataclasses import dataclass

from .. import _cabi
from ..core.data_block import DataBlock
from ..core.data_encoder_decoder import DataDecoder, DataEncoder
from ..core.prob_dist import Frequencies
from ..utils.bitarray_utils import BitArray
from ._gpu_base import GpuCoderBase


@dataclass
class RangeCoderParams:
    DATA_BLOCK_SIZE_BITS: int = 32
    PRECISION: int = 32

    def __post_init__(self):
        assert self.PRECISION % 8 == 0
        self.TOP = 1 << (self.PRECISION - 8)
        self.BOTTOM = 1 << (self.PRECISION - 16)
        self.MASK = (1 << self.PRECISION) - 1


class _RangeCoder(GpuCoderBase):
    def __init__(self, params: RangeCoderParams, freqs: Frequencies):
        self.params = params
        self.freqs = freqs
        # same constructor checks as range_coder.py:84-85
        assert min(self.freqs.freq_dict.values()) > 0
        assert self.freqs.total_freq <= self.params.BOTTOM

    def _freqs(self):
        return self.freqs

    def _make_cabi_params(self):
        return _cabi.SclParams(coder=_cabi.CODER_RANGE, data_block_size_bits=int(self.params.DATA_BLOCK_SIZE_BITS), num_bits_out=0,
                               range_factor=0, num_state_bits=0, precision=int(self.params.PRECISION), model=0, max_allowed_total_freq=0)


class RangeEncoder(_RangeCoder, DataEncoder):
    def encode_block(self, data_block: DataBlock) -> BitArray:
        return self._encode_one(data_block)


class RangeDecoder(_RangeCoder, DataDecoder):
    def decode_block(self, encoded_bitarray: BitArray):
        return self._decode_one(encoded_bitarray)
