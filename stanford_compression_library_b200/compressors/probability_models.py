"""Frequency models for the arithmetic coder (scl/compressors/probability_models.py).

The model objects live on the host and keep the reference's semantics -- in particular the
table is MUTATED as symbols are coded and is never reset between blocks
(probability_models.py:39-44, data_encoder_decoder.py:23-27).  On the device each block's model
lives in shared memory (csrc/scl_aec.cuh; first-generation Fenwick tree in csrc/scl_lane.cuh);
after a single-block call the final table is copied back into the model object.
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

    # ---- device table: freqs_current in alphabet order (include/scl_b200.h scl_coder_model_words) ----
    def _to_table(self):
        return [int(f) for f in self.freqs_current.freq_list]

    def _from_table(self, table):
        for k, v in zip(list(self.freqs_current.freq_dict), table):
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
    """Order-k context model (probability_models.py:95-168): counts of (k+1)-tuples, all ones at
    the start, the past k symbols (as alphabet indices, initially all 0) select the row used for the
    next symbol.  Same attributes as the reference (`freqs_kplus1_tuple`, `past_k`, `freqs_current`).
    On the device the table of one block lives in shared memory while len(alphabet)^k * (len(alphabet) + 1)
    <= 1600 words (csrc/scl_aec.cuh AecCtxPolicy); larger tables -- a byte alphabet at k = 1 is 65 536 counters --
    stay in HBM, one copy per block being coded, with only the row totals in shared memory (AecCtxGlobalPolicy;
    len(alphabet)^k <= 512 contexts).

    As in the reference there is no halving that works: when one count reaches
    max_allowed_total_freq the reference's `np.max(count // 2, 1)` raises; here the coder raises
    AssertionError (status SCL_ST_TOTAL_FREQ)."""

    CABI_MODEL = _cabi.MODEL_ORDER_K

    def __init__(self, alphabet, k: int, max_allowed_total_freq: int):
        assert k >= 0
        self.k = k
        self.alphabet = list(alphabet)
        self.alphabet_to_idx = {a: i for i, a in enumerate(self.alphabet)}
        self.freqs_kplus1_tuple = np.ones([len(self.alphabet)] * (k + 1), dtype=int)
        self.max_allowed_total_freq = max_allowed_total_freq
        self.past_k = [0] * k

    @property
    def freqs_current(self):
        row = self.freqs_kplus1_tuple[tuple(self.past_k)] if self.k > 0 else self.freqs_kplus1_tuple
        return Frequencies(dict(zip(self.alphabet, (int(f) for f in np.ravel(row)))))

    def update_model(self, s):
        # host-side single-symbol update (:137-168); the coders apply the same rule on the device
        cur = (*self.past_k, self.alphabet_to_idx[s])
        self.freqs_kplus1_tuple[cur] += 1
        if self.k > 0:
            self.past_k = self.past_k[1:] + [self.alphabet_to_idx[s]]
        if self.freqs_kplus1_tuple[cur] >= self.max_allowed_total_freq:
            raise AssertionError("a (k+1)-tuple count reached max_allowed_total_freq (the reference's halving raises here)")

    # ---- device table: [counts row-major][context index] (include/scl_b200.h scl_coder_model_words) ----
    def _context_index(self):
        ctx = 0
        for d in self.past_k:
            ctx = ctx * len(self.alphabet) + int(d)
        return ctx

    def _to_table(self):
        return [int(v) for v in np.ravel(self.freqs_kplus1_tuple)] + [self._context_index()]

    def _from_table(self, table):
        n = len(self.alphabet)
        self.freqs_kplus1_tuple = np.array(table[:-1], dtype=int).reshape([n] * (self.k + 1))
        ctx, digits = int(table[-1]), []
        for _ in range(self.k):
            digits.append(ctx % n)
            ctx //= n
        self.past_k = digits[::-1]
