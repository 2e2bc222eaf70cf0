"""Stand-in for `bitarray.util` (see package docstring): ba2int, int2ba, urandom.

Call sites in the reference: scl/utils/bitarray_utils.py:2,34,38,42.
"""
import os

from . import bitarray


def ba2int(a, signed=False):
    assert not signed
    if len(a) == 0:
        raise ValueError("non-empty bitarray expected")
    v = 0
    for bit in a:
        v = (v << 1) | bit
    return v


def int2ba(i, length=None, endian="big", signed=False):
    assert not signed and endian == "big"
    if not isinstance(i, int):
        raise TypeError("int expected")
    if i < 0:
        raise OverflowError("unsigned integer not positive")
    if length is None:
        s = bin(i)[2:]
    else:
        if length <= 0:
            raise ValueError("length must be > 0")
        if i >= (1 << length):
            raise OverflowError("unsigned integer not in range(0, %d), got %d" % (1 << length, i))
        s = bin(i)[2:].rjust(length, "0")
    return bitarray(s)


def urandom(n, endian="big"):
    raw = os.urandom((n + 7) // 8)
    a = bitarray()
    a.frombytes(raw)
    return a[:n]
