"""DataBlock: the unit the coders work on (scl/core/data_block.py:5-106).

`data_list` may be any sequence -- a list, a numpy array (the reference itself passes one,
arithmetic_coding.py:397-402) or a `torch.uint8` tensor, which is the natural form for this
backend.
"""
from collections import Counter

from .prob_dist import ProbabilityDist


class DataBlock:
    def __init__(self, data_list):
        self.data_list = data_list

    @property
    def size(self):
        return len(self.data_list)

    def _as_python_list(self):
        d = self.data_list
        return d.tolist() if hasattr(d, "tolist") else list(d)

    def get_alphabet(self):
        alphabet = set()
        for d in self._as_python_list():  # grown element by element like data_block.py:30-35
            alphabet.add(d)
        return alphabet

    def get_counts(self, order=0):
        """{symbol: count}.  Keys come out in the reference's order -- the iteration order of the alphabet
        SET (data_block.py:57-64), not first occurrence: key order defines the cumulative table of a
        `Frequencies` built from these counts, hence the bitstream."""
        if order != 0:
            raise NotImplementedError("[order != 0] counts not implemented")
        counts = Counter(self._as_python_list())
        return {a: counts[a] for a in self.get_alphabet()}

    def get_empirical_distribution(self, order=0) -> ProbabilityDist:
        if order != 0:
            raise NotImplementedError("[order != 0] empirical counts not implemented")
        n = self.size
        return ProbabilityDist({s: c / n for s, c in self.get_counts().items()})

    def get_entropy(self, order=0):
        if order != 0:
            raise NotImplementedError("[order != 0] Entropy computation not implemented")
        return self.get_empirical_distribution().entropy
