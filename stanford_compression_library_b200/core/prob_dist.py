"""ProbabilityDist / Frequencies: the reference's distribution types (scl/core/prob_dist.py).

`Frequencies` is a boundary type of the hot path: the ORDER of its keys (dict insertion order,
prob_dist.py:169-171) defines the cumulative table (prob_dist.py:193-205) and therefore the
bitstream.  `to_arrays()` is the hand-off to the device: alphabet bytes + counts in that order.
"""
import numpy as np


class ProbabilityDist:
    """Ordered {symbol: probability} (scl/core/prob_dist.py:6-90)."""

    def __init__(self, prob_dict=None):
        self._validate_prob_dist(prob_dict)
        self.prob_dict = prob_dict

    def __repr__(self):
        return "ProbabilityDist(%r" % (self.prob_dict,)

    @property
    def size(self):
        return len(self.prob_dict)

    @property
    def alphabet(self):
        return list(self.prob_dict)

    @property
    def prob_list(self):
        return list(self.prob_dict.values())

    @classmethod
    def get_sorted_prob_dist(cls, prob_dict, descending=False):
        return cls(dict(sorted(prob_dict.items(), key=lambda kv: kv[1], reverse=descending)))

    @classmethod
    def normalize_prob_dict(cls, prob_dict):
        total = sum(prob_dict.values())
        return cls({k: v / total for k, v in prob_dict.items()})

    @property
    def cumulative_prob_dict(self):
        out, acc = {}, 0
        for k, p in self.prob_dict.items():
            out[k] = acc
            acc += p
        return out

    @property
    def entropy(self):
        return sum(-p * np.log2(p) for p in self.prob_dict.values())

    def probability(self, symbol):
        return self.prob_dict[symbol]

    def neg_log_probability(self, symbol):
        return -np.log2(self.probability(symbol))

    @staticmethod
    def _validate_prob_dist(prob_dict):
        # same checks and error types as prob_dist.py:77-90
        total = 0
        for p in prob_dict.values():
            assert p >= 1e-6, "probabilities negative or too small cause stability issues"
            total += p
        if abs(total - 1.0) > 1e-8:
            raise ValueError("probabilities do not sum to 1")


def get_avg_neg_log_prob(prob_dist: ProbabilityDist, data_block) -> float:
    """Average -log2 p(s) over the block (scl/core/prob_dist.py:143-158)."""
    data = data_block.data_list
    return sum(prob_dist.neg_log_probability(s) for s in data) / len(data)


class Frequencies:
    """Ordered {symbol: integer count} (scl/core/prob_dist.py:161-229)."""

    def __init__(self, freq_dict=None):
        self.freq_dict = freq_dict

    def __repr__(self):
        return "Frequencies(%r" % (self.freq_dict,)

    @property
    def size(self):
        return len(self.freq_dict)

    @property
    def alphabet(self):
        return list(self.freq_dict)

    @property
    def freq_list(self):
        return list(self.freq_dict.values())

    @property
    def total_freq(self) -> int:
        return np.sum(self.freq_list)  # numpy integer, like the reference (prob_dist.py:188-191)

    @property
    def cumulative_freq_dict(self) -> dict:
        out, acc = {}, 0
        for k, f in self.freq_dict.items():
            out[k] = acc
            acc += f
        return out

    def frequency(self, symbol):
        return self.freq_dict[symbol]

    def get_prob_dist(self) -> ProbabilityDist:
        total = self.total_freq
        return ProbabilityDist({k: f / total for k, f in self.freq_dict.items()})

    @staticmethod
    def _validate_freq_dist(freq_dict):
        for f in freq_dict.values():
            assert f > 0, "frequency cannot be negative or 0"
            assert isinstance(f, int)

    # ---- device hand-off --------------------------------------------------------------------
    def byte_alphabet(self):
        """(alphabet_bytes, is_native): the byte value each key is coded as on the device.

        If every key is an integer in 0..255 the keys ARE the byte values (so `uint8` tensors of
        raw data can be fed to `encode_blocks` directly); otherwise key i is coded as byte i and
        the single-block API translates symbols <-> indices on the host.
        """
        keys = self.alphabet
        if len(keys) > 256:
            raise NotImplementedError("this backend codes at most 256 distinct symbols per table")
        native = all(isinstance(k, (int, np.integer)) and not isinstance(k, bool) and 0 <= int(k) <= 255 for k in keys)
        if native:
            return np.array([int(k) for k in keys], dtype=np.uint8), True
        return np.arange(len(keys), dtype=np.uint8), False

    def to_arrays(self):
        """(alphabet uint8[n], freq uint64[n]) in dict order."""
        alpha, _ = self.byte_alphabet()
        return alpha, np.array([int(f) for f in self.freq_dict.values()], dtype=np.uint64)
