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

    def _cache_key(self):
        m = self.freq_model
        if m.CABI_MODEL == _cabi.MODEL_ORDER_K:  # the handle only depends on the alphabet and k
            return ("order_k", tuple(m.alphabet), m.k, int(m.max_allowed_total_freq))
        return super()._cache_key()

    def _make_cabi_params(self):
        if self.freq_model.CABI_MODEL is None:
            raise NotImplementedError("frequency model %s has no device implementation" % type(self.freq_model).__name__)
        return _cabi.SclParams(coder=_cabi.CODER_AEC, data_block_size_bits=int(self.params.DATA_BLOCK_SIZE_BITS), num_bits_out=0,
                               range_factor=0, num_state_bits=0, precision=int(self.params.PRECISION), model=self.freq_model.CABI_MODEL,
                               model_order=int(getattr(self.freq_model, "k", 0)) if self.freq_model.CABI_MODEL == _cabi.MODEL_ORDER_K else 0,
                               max_allowed_total_freq=int(self.freq_model.max_allowed_total_freq))

    def _blocks_independent(self) -> bool:
        # the reference carries the (mutated) model from one encode_block call to the next (arithmetic_coding.py:118,
        # data_encoder_decoder.py:23-27): only a model that never changes makes the blocks of a stream independent
        from .probability_models import FixedFreqModel

        return isinstance(self.freq_model, FixedFreqModel)

    def _max_symbols_for_bits(self, nbits: int) -> int:
        # an adaptive model can drive the cost of a repeated symbol towards zero bits: no bound from the
        # stream length, only the flat allocation cap
        from .probability_models import FixedFreqModel

        if isinstance(self.freq_model, FixedFreqModel):
            return super()._max_symbols_for_bits(nbits)
        return 1 << 28

    def _model_tensor(self, n_blocks=1):
        dev = self.device_coder()
        return torch.tensor([self.freq_model._to_table()], dtype=torch.int64, device=dev.device).repeat(n_blocks, 1)

    def _writeback(self, model_t):
        self.freq_model._from_table(model_t[0].tolist())

    # batched API: every block starts from a copy of the model's CURRENT state.  For the fixed / IID
    # models that is the creation-time table of the handle (re-created when the table changes); the
    # order-k table is passed explicitly.
    def encode_blocks(self, data, sizes=None, reuse=None):
        if self.freq_model.CABI_MODEL != _cabi.MODEL_ORDER_K:
            return super().encode_blocks(data, sizes=sizes, reuse=reuse)
        return self.device_coder().encode_blocks(data, sizes=sizes, model=self._batch_model(int(data.shape[0])), reuse=reuse)

    def encode_blocks_packed(self, data, sizes=None, framed=False, capacity=None, reuse=None):
        model = self._batch_model(int(data.shape[0])) if self.freq_model.CABI_MODEL == _cabi.MODEL_ORDER_K else None
        return self.device_coder().encode_blocks_packed(data, sizes=sizes, model=model, framed=framed, capacity=capacity, reuse=reuse)

    def decode_blocks(self, enc, max_block_len, out=None, reuse=None):
        if self.freq_model.CABI_MODEL != _cabi.MODEL_ORDER_K:
            return super().decode_blocks(enc, max_block_len, out=out, reuse=reuse)
        return self.device_coder().decode_blocks(enc, max_block_len, model=self._batch_model(enc.n_blocks), out=out, reuse=reuse)

    def _batch_model(self, n_blocks):
        m = self.freq_model
        in_hbm = len(m.alphabet) ** m.k * (len(m.alphabet) + 1) > 1600  # csrc kAecCtxMaxWords: the table is the lanes' working storage
        table = m._to_table()
        if not in_hbm and table[-1] == 0 and all(v == 1 for v in table[:-1]):
            return None  # untouched model: the kernels start from all ones / context 0 themselves
        return self._model_tensor(n_blocks)


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
