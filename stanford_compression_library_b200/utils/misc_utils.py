"""Small host helpers with the reference's names (scl/utils/misc_utils.py:6-11)."""
import functools

import numpy as np

cache = functools.lru_cache(maxsize=None)


def is_power_of_two(x) -> bool:
    # Same float test as the reference (misc_utils.py:9-11) so that tANSParams accepts and
    # rejects exactly the same totals.
    return float(np.log2(x)).is_integer()
