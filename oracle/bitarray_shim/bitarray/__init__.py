"""Pure-Python stand-in for the third-party `bitarray` package.

TEST INFRASTRUCTURE ONLY (lives under oracle/).  The reference library
(/root/reference, scl/utils/bitarray_utils.py:1-2,25) uses `bitarray.bitarray`
purely as a bit *container*; the package is not installed in this image and
there is no network.  This module provides exactly the surface the reference
touches so that the UNMODIFIED reference can be imported
to validate the C restatement and to generate the golden vectors committed
under tests/golden/.  It is deliberately the dumbest possible implementation
(a Python list of 0/1 ints, big-endian bit order) so that it is obviously
correct; it is never imported by the product package.
"""


class bitarray:
    __slots__ = ("_b",)

    def __init__(self, init=None, endian="big"):
        assert endian == "big"
        if init is None:
            self._b = []
        elif isinstance(init, str):
            bits = []
            for ch in init:
                if ch == "0":
                    bits.append(0)
                elif ch == "1":
                    bits.append(1)
                elif ch in " _\n\t\r\v":
                    continue
                else:
                    raise ValueError("expected '0' or '1' (or whitespace), got %r" % ch)
            self._b = bits
        elif isinstance(init, bitarray):
            self._b = list(init._b)
        elif isinstance(init, int):
            self._b = [0] * init
        else:
            self._b = [1 if x else 0 for x in init]

    # -- container protocol -------------------------------------------------
    def __len__(self):
        return len(self._b)

    def __iter__(self):
        return iter(self._b)

    def __getitem__(self, key):
        if isinstance(key, slice):
            out = bitarray()
            out._b = self._b[key]
            return out
        return self._b[key]

    def __setitem__(self, key, value):
        if isinstance(key, slice):
            if isinstance(value, bitarray):
                self._b[key] = value._b
            else:
                n = len(self._b[key])
                self._b[key] = [1 if value else 0] * n
        else:
            self._b[key] = 1 if value else 0

    def __eq__(self, other):
        if not isinstance(other, bitarray):
            return NotImplemented
        return self._b == other._b

    def __ne__(self, other):
        r = self.__eq__(other)
        return r if r is NotImplemented else not r

    __hash__ = None

    def __add__(self, other):
        out = bitarray()
        out._b = self._b + _bits_of(other)
        return out

    def __iadd__(self, other):
        self._b += _bits_of(other)
        return self

    def __repr__(self):
        return "bitarray('%s')" % self.to01()

    def __bool__(self):
        return len(self._b) > 0

    def __copy__(self):
        return bitarray(self)

    def __deepcopy__(self, memo):
        return bitarray(self)

    # -- methods used by the reference --------------------------------------
    def copy(self):
        return bitarray(self)

    def append(self, bit):
        self._b.append(1 if bit else 0)

    def extend(self, other):
        self._b += _bits_of(other)

    def to01(self):
        return "".join("1" if b else "0" for b in self._b)

    def tolist(self):
        return list(self._b)

    def frombytes(self, data):
        for byte in bytes(data):
            for i in range(7, -1, -1):
                self._b.append((byte >> i) & 1)

    def tobytes(self):
        b = self._b
        out = bytearray((len(b) + 7) // 8)
        for i, bit in enumerate(b):
            if bit:
                out[i >> 3] |= 0x80 >> (i & 7)
        return bytes(out)

    def count(self, value=1):
        return self._b.count(1 if value else 0)

    def endian(self):
        return "big"


def _bits_of(other):
    if isinstance(other, bitarray):
        return other._b
    if isinstance(other, str):
        return bitarray(other)._b
    return [1 if x else 0 for x in other]


__version__ = "0.0-shim"
