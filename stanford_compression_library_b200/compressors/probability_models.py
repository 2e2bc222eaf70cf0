"""Frequency models for the arithmetic coder (scl/compressors/probability_models.py).

The model objects live on the host and keep the reference's semantics -- in particular the
table is MUTATED as symbols are coded and is never reset between blocks
(probability_models.py:39-44, data_encoder_decoder.py:23-27).  On the device each block's model
is a 256-counter Fenwick tree in shared memory (csrc/scl_lane.cuh aec_model_update); after a
single-block call the final counts are copied back into `freqs_current`.
"""
import abc
import copy

import numpy as np

from .. import _cabi
from ..core.prob_dist import Frequencies


class FreqModelBase(abc.ABC):
    CABI_MODEL = None

    def __init__(self, freqs_initial: Frequencies, max_allowed_total_freq):
        self.freqs_current = copy.deepcopy(freqs_initial)
        self.max_allowed_total_freq = max_allowed_total_freq

    @abc.abstractmethod
    def update_model(self, s):
        raise NotImplementedError

    def _set_counts(self, counts):
        for k, v in zip(list(self.freqs_current.freq_dict), counts):
            self.freqs_current.freq_dict[k] = int(v)


class FixedFreqModel(FreqModelBase):
    CABI_MODEL = _cabi.MODEL_FIXED

    def update_model(self, s):
        pass


class AdaptiveIIDFreqModel(FreqModelBase):
    CABI_MODEL = _cabi.MODEL_ADAPTIVE_IID

    def update_model(self, s):
        # host-side single-symbol update with the reference's rule (probability_models.py:86-92);
        # the coders do not call this -- the device applies the same rule per symbol
        fd = self.freqs_current.freq_dict
        fd[s] += 1
        if self.freqs_current.total_freq >= self.max_allowed_total_freq:
            for k, f in fd.items():
                fd[k] = max(f // 2, 1)


class AdaptiveOrderKFreqModel(FreqModelBase):
    """Order-k context model (probability_models.py:95-160).  k = 0 is exactly the adaptive IID
    model started from all-ones (arithmetic_coding.py:449-463) and runs on the device; k > 0 is
    listed as "next" in SURVEY.md 8(f) and is not implemented by this backend."""

    CABI_MODEL = _cabi.MODEL_ADAPTIVE_IID

    def __init__(self, alphabet, k: int, max_allowed_total_freq: int):
        assert k >= 0
        if k > 0:
            raise NotImplementedError("order-k (k > 0) context models are not implemented on the device yet (SURVEY.md 8f)")
        self.k = k
        self.alphabet = alphabet
        super().__init__(Frequencies({a: 1 for a in alphabet}), max_allowed_total_freq)

    def update_model(self, s):
        AdaptiveIIDFreqModel.update_model(self, s)
