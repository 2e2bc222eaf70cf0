"""Streaming rANS on the GPU behind the reference's rANSParams / rANSEncoder / rANSDecoder API
(scl/compressors/rANS.py:78-297).

`encode_block` / `decode_block` produce and consume exactly the reference's bitstream:
    [size : DATA_BLOCK_SIZE_BITS][final state : NUM_STATE_BITS][renormalisation chunks, last symbol first]
The per-symbol loops run in CUDA (csrc/scl_lane.cuh: rans32_* fast path for 32-bit states,
rans64_* for every other parameter set), one warp lane per DataBlock; `encode_blocks` /
`decode_blocks` are the batched entry points.
"""
from dataclasses import dataclass

from .. import _cabi
from ..core.data_block import DataBlock
from ..core.data_encoder_decoder import DataDecoder, DataEncoder
from ..core.prob_dist import Frequencies
from ..utils.bitarray_utils import BitArray, get_bit_width
from ._gpu_base import GpuCoderBase


@dataclass
class rANSParams:
    """Same fields, defaults and derived values as the reference (rANS.py:78-120)."""

    freqs: Frequencies
    DATA_BLOCK_SIZE_BITS: int = 32
    NUM_BITS_OUT: int = 1
    RANGE_FACTOR: int = 1 << 16

    def __post_init__(self):
        self.M = self.freqs.total_freq
        self.L = self.RANGE_FACTOR * self.M
        self.H = self.L * (1 << self.NUM_BITS_OUT) - 1
        self.min_shrunk_state = {s: self.RANGE_FACTOR * f for s, f in self.freqs.freq_dict.items()}
        self.max_shrunk_state = {s: self.RANGE_FACTOR * f * (1 << self.NUM_BITS_OUT) - 1 for s, f in self.freqs.freq_dict.items()}
        self.INITIAL_STATE = self.L
        self.NUM_STATE_BITS = get_bit_width(self.H)  # the reference's float formula, on purpose
        self.BITS_OUT_MASK = 1 << self.NUM_BITS_OUT

    def _cabi(self, coder=_cabi.CODER_RANS) -> _cabi.SclParams:
        return _cabi.SclParams(coder=coder, data_block_size_bits=int(self.DATA_BLOCK_SIZE_BITS), num_bits_out=int(self.NUM_BITS_OUT),
                               range_factor=int(self.RANGE_FACTOR), num_state_bits=int(self.NUM_STATE_BITS), precision=0, model=0,
                               max_allowed_total_freq=0)


class _RansCoder(GpuCoderBase):
    _CODER = _cabi.CODER_RANS

    def __init__(self, rans_params: rANSParams):
        self.params = rans_params

    def _freqs(self):
        return self.params.freqs

    def _make_cabi_params(self):
        return self.params._cabi(self._CODER)


class rANSEncoder(_RansCoder, DataEncoder):
    """rANSEncoder.encode_block (rANS.py:186-210) on the device."""

    def encode_block(self, data_block: DataBlock) -> BitArray:
        return self._encode_one(data_block)


class rANSDecoder(_RansCoder, DataDecoder):
    """rANSDecoder.decode_block (rANS.py:270-297) on the device; tolerates trailing bits and
    returns the exact number of bits consumed."""

    def decode_block(self, encoded_bitarray: BitArray):
        return self._decode_one(encoded_bitarray)
