"""Host-side bit container and integer<->bits helpers.

The reference aliases the third-party `bitarray.bitarray` C extension as `BitArray`
(scl/utils/bitarray_utils.py:25) and uses it purely as a container.  This backend keeps coded
streams on the device as byte buffers + bit lengths; `BitArray` here is the host-side view
returned by `encode_block` / accepted by `decode_block`, implemented on numpy with the subset of
the bitarray API the reference's callers and tests rely on (construction from a '01' string,
len, iteration, indexing, slicing, +, +=, ==, extend, frombytes, tobytes).  Bit order is
big-endian (MSB first) like bitarray's default.
"""
import os

import numpy as np


class BitArray:
    __slots__ = ("_b",)

    def __init__(self, init=None):
        if init is None:
            self._b = np.zeros(0, dtype=np.uint8)
        elif isinstance(init, BitArray):
            self._b = init._b.copy()
        elif isinstance(init, str):
            raw = np.frombuffer(init.encode("ascii"), dtype=np.uint8)
            raw = raw[~np.isin(raw, (32, 95, 9, 10, 13))]  # bitarray ignores whitespace and '_'
            if raw.size and not np.isin(raw, (48, 49)).all():
                raise ValueError("expected '0' or '1' (or whitespace)")
            self._b = (raw - 48).astype(np.uint8)
        elif isinstance(init, (int, np.integer)):
            self._b = np.zeros(int(init), dtype=np.uint8)
        else:
            self._b = (np.asarray(list(init)) != 0).astype(np.uint8)

    # ---- construction from device output --------------------------------------------------
    @classmethod
    def from_packed(cls, packed, nbits: int, bit_offset: int = 0) -> "BitArray":
        """Bits [bit_offset, bit_offset + nbits) of an MSB-first packed byte buffer."""
        packed = np.frombuffer(packed, dtype=np.uint8) if isinstance(packed, (bytes, bytearray)) else np.asarray(packed, dtype=np.uint8)
        first, last = bit_offset >> 3, (bit_offset + nbits + 7) >> 3
        bits = np.unpackbits(packed[first:last])
        start = bit_offset - 8 * first
        out = cls()
        out._b = bits[start : start + nbits].copy()
        if out._b.size != nbits:
            raise ValueError("packed buffer too short for %d bits at offset %d" % (nbits, bit_offset))
        return out

    # ---- container protocol ----------------------------------------------------------------
    def __len__(self):
        return int(self._b.size)

    def __iter__(self):
        return iter(self._b.tolist())

    def __getitem__(self, key):
        if isinstance(key, slice):
            out = BitArray()
            out._b = self._b[key].copy()
            return out
        return int(self._b[key])

    def __setitem__(self, key, value):
        if isinstance(value, BitArray):
            self._b[key] = value._b
        else:
            self._b[key] = 1 if value else 0

    def __eq__(self, other):
        if not isinstance(other, BitArray):
            return NotImplemented
        return self._b.size == other._b.size and bool((self._b == other._b).all())

    __hash__ = None

    def __add__(self, other):
        out = BitArray()
        out._b = np.concatenate([self._b, _as_bits(other)])
        return out

    def __iadd__(self, other):
        self._b = np.concatenate([self._b, _as_bits(other)])
        return self

    def __repr__(self):
        return "BitArray('%s')" % self.to01()

    def __copy__(self):
        return BitArray(self)

    def __deepcopy__(self, memo):
        return BitArray(self)

    # ---- bitarray methods the reference calls ---------------------------------------------
    def copy(self):
        return BitArray(self)

    def append(self, bit):
        self._b = np.append(self._b, np.uint8(1 if bit else 0))

    def extend(self, other):
        self._b = np.concatenate([self._b, _as_bits(other)])

    def to01(self) -> str:
        return (self._b + 48).astype(np.uint8).tobytes().decode("ascii")

    def tolist(self):
        return self._b.tolist()

    def frombytes(self, data):
        self._b = np.concatenate([self._b, np.unpackbits(np.frombuffer(bytes(data), dtype=np.uint8))])

    def tobytes(self) -> bytes:
        return np.packbits(self._b).tobytes()  # right-padded with zero bits

    def to_packed(self) -> np.ndarray:
        return np.packbits(self._b)

    def count(self, value=1):
        ones = int(self._b.sum())
        return ones if value else int(self._b.size) - ones


def _as_bits(other) -> np.ndarray:
    if isinstance(other, BitArray):
        return other._b
    return BitArray(other)._b


def get_bit_width(x) -> int:
    """Minimum number of bits for the unsigned integer x.

    Deliberately the reference's float formula, ceil(log2(x + 1)) (scl/utils/bitarray_utils.py:8-20):
    NUM_STATE_BITS is part of the bitstream format, so it must be derived identically -- including
    where float64 rounding makes it differ from int.bit_length() (x >= 2^49).
    """
    assert x >= 0
    if x == 0:
        return 1
    return int(np.ceil(np.log2(x + 1)))


def uint_to_bitarray(x: int, bit_width=None) -> BitArray:
    """Unsigned integer -> MSB-first bits (scl/utils/bitarray_utils.py:28-34; bitarray.util.int2ba)."""
    assert isinstance(x, (int, np.integer))
    x = int(x)
    if x < 0:
        raise OverflowError("unsigned integer not positive")
    if bit_width is None:
        return BitArray(bin(x)[2:])
    if bit_width <= 0:
        raise ValueError("length must be > 0")
    if x >> bit_width:
        raise OverflowError("unsigned integer not in range(0, %d), got %d" % (1 << bit_width, x))
    return BitArray(bin(x)[2:].rjust(bit_width, "0"))


def bitarray_to_uint(bit_array: BitArray) -> int:
    """MSB-first bits -> unsigned integer (scl/utils/bitarray_utils.py:37-38; bitarray.util.ba2int)."""
    if len(bit_array) == 0:
        raise ValueError("non-empty bitarray expected")
    return int(bit_array.to01(), 2)


def get_random_bitarray(size) -> BitArray:
    raw = np.frombuffer(os.urandom((size + 7) // 8), dtype=np.uint8)
    return BitArray.from_packed(raw, size)
