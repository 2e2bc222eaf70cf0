"""Finite-precision arithmetic coder on the GPU behind the reference's AECParams /
ArithmeticEncoder / ArithmeticDecoder API (scl/compressors/arithmetic_coding.py:20-287).

Bit-exact with the reference, including its strict `high < HALF` / `low > HALF` tests, the
break-before-renormalise on the last symbol and the trailing-bit accounting of the decoder.
As in the reference the frequency model is stateful across `encode_block` calls; the batched
`encode_blocks` / `decode_blocks` give every block a fresh copy of the model's current table.
"""
from dataclasses import dataclass

import torch

from .. import _cabi
from ..core.data_block import DataBlock
from ..core.data_encoder_decoder import DataDecoder, DataEncoder
from ..utils.bitarray_utils import BitArray
from ._gpu_base import GpuCoderBase
from .probability_models import FreqModelBase


@dataclass
class AECParams:
    DATA_BLOCK_SIZE_BITS: int = 32
    PRECISION: int = 32

    def __post_init__(self):
        self.FULL = 1 << self.PRECISION
        self.HALF = 1 << (self.PRECISION - 1)
        self.QTR = 1 << (self.PRECISION - 2)
        self.MAX_ALLOWED_TOTAL_FREQ = self.QTR
        self.MAX_BLOCK_SIZE = 1 << self.DATA_BLOCK_SIZE_BITS


class _AecCoder(GpuCoderBase):
    def __init__(self, params: AECParams, freq_model: FreqModelBase):
        self.params = params
        self.freq_model = freq_model  # updated in place by every encode_block / decode_block

    def _freqs(self):
        return self.freq_model.freqs_current

    def _make_cabi_params(self):
        if self.freq_model.CABI_MODEL is None:
            raise NotImplementedError("frequency model %s has no device implementation" % type(self.freq_model).__name__)
        return _cabi.SclParams(coder=_cabi.CODER_AEC, data_block_size_bits=int(self.params.DATA_BLOCK_SIZE_BITS), num_bits_out=0,
                               range_factor=0, num_state_bits=0, precision=int(self.params.PRECISION), model=self.freq_model.CABI_MODEL,
                               max_allowed_total_freq=int(self.freq_model.max_allowed_total_freq))

    def _model_tensor(self):
        dev = self.device_coder()
        counts = [int(f) for f in self.freq_model.freqs_current.freq_list]
        return torch.tensor([counts], dtype=torch.int64, device=dev.device)

    def _writeback(self, model_t):
        self.freq_model._set_counts(model_t[0].tolist())


class ArithmeticEncoder(_AecCoder, DataEncoder):
    def encode_block(self, data_block: DataBlock) -> BitArray:
        # arithmetic_coding.py:85: `assert size < (1 << MAX_BLOCK_SIZE)` can never fail for a
        # representable size (and costs ~0.75 s in the reference); the real limit is the header:
        if data_block.size >> self.params.DATA_BLOCK_SIZE_BITS:
            raise OverflowError("data_block.size does not fit DATA_BLOCK_SIZE_BITS")
        m = self._model_tensor()
        out = self._encode_one(data_block, model=m)
        self._writeback(m)
        return out


class ArithmeticDecoder(_AecCoder, DataDecoder):
    def decode_block(self, encoded_bitarray: BitArray):
        m = self._model_tensor()
        out = self._decode_one(encoded_bitarray, model=m)
        self._writeback(m)
        return out
