"""ctypes front-end of oracle/scl_oracle.c (the C restatement) -- TEST INFRASTRUCTURE ONLY.

Symbols are alphabet *indices* (dict order of the Frequencies); streams are numpy uint8
arrays packed MSB-first plus a bit length, i.e. exactly `BitArray.tobytes()` + `len()`.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libscl_oracle.so")

CODER_RANS, CODER_TANS, CODER_RANGE, CODER_AEC = 0, 1, 2, 3
MODEL_FIXED, MODEL_ADAPTIVE_IID, MODEL_ORDER_K = 0, 1, 2

STATUS = {
    0: "ok",
    1: "bad symbol (KeyError)",
    2: "end state != INITIAL_STATE (AssertionError)",
    3: "overflow (OverflowError / buffer)",
    4: "truncated stream (ValueError)",
    5: "bad parameters",
    6: "total_freq >= MAX_ALLOWED_TOTAL_FREQ (AssertionError)",
}


def build(force=False):
    src = os.path.join(_HERE, "scl_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "_ref/libscl_oracle.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        u8p, u32p, u64p, i32p = (ctypes.POINTER(t) for t in (ctypes.c_uint8, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int32))
        L.scl_oracle_create.restype = ctypes.c_void_p
        L.scl_oracle_create.argtypes = [ctypes.c_int, u64p, ctypes.c_uint32] + [ctypes.c_uint64] * 4 + [ctypes.POINTER(ctypes.c_int)]
        L.scl_oracle_destroy.argtypes = [ctypes.c_void_p]
        L.scl_oracle_encode_block.argtypes = [ctypes.c_void_p, u8p, ctypes.c_uint64, u64p, u8p, ctypes.c_uint64, u64p]
        L.scl_oracle_decode_block.argtypes = [ctypes.c_void_p, u8p, ctypes.c_uint64, ctypes.c_uint64, u64p, u8p, ctypes.c_uint64, u64p, u64p]
        L.scl_oracle_encode_batch.argtypes = [ctypes.c_void_p, u8p, ctypes.c_uint64, u32p, ctypes.c_uint32, ctypes.c_uint64, u8p, ctypes.c_uint64, u64p, i32p, ctypes.c_int]
        L.scl_oracle_decode_batch.argtypes = [ctypes.c_void_p, u8p, u64p, u64p, ctypes.c_uint64, u8p, ctypes.c_uint64, u32p, u64p, i32p, ctypes.c_int]
        L.scl_oracle_tans_tables.argtypes = [ctypes.c_void_p, u64p, u64p, u32p, u64p, u32p, u64p]
        L.scl_oracle_max_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t)) if a is not None else None


def ref_get_bit_width(x) -> int:
    """The reference's float formula (scl/utils/bitarray_utils.py:8-20), quirk included."""
    assert x >= 0
    if x == 0:
        return 1
    return int(np.ceil(np.log2(x + 1)))


class OracleError(Exception):
    def __init__(self, code):
        super().__init__(STATUS.get(code, "status %d" % code))
        self.code = code


class Oracle:
    """One coder configuration.  See scl_oracle.c `oracle_cfg` for the meaning of p0..p3."""

    def __init__(self, coder, freqs, p0, p1, p2=0, p3=0):
        self.coder = coder
        self.freq = np.ascontiguousarray(np.asarray(freqs, dtype=np.uint64))
        self.n_sym = int(self.freq.size)
        err = ctypes.c_int(0)
        self.h = lib().scl_oracle_create(coder, _p(self.freq, ctypes.c_uint64), self.n_sym, int(p0), int(p1), int(p2), int(p3), ctypes.byref(err))
        if not self.h:
            raise OracleError(err.value)

    # -- constructors mirroring the reference's parameter classes ----------------------
    @classmethod
    def rans(cls, freqs, DATA_BLOCK_SIZE_BITS=32, NUM_BITS_OUT=1, RANGE_FACTOR=1 << 16, tans=False):
        M = int(np.sum(np.asarray(freqs, dtype=np.uint64)))
        H = RANGE_FACTOR * M * (1 << NUM_BITS_OUT) - 1
        return cls(CODER_TANS if tans else CODER_RANS, freqs, DATA_BLOCK_SIZE_BITS, NUM_BITS_OUT, RANGE_FACTOR, ref_get_bit_width(H))

    @classmethod
    def tans(cls, freqs, **kw):
        return cls.rans(freqs, tans=True, **kw)

    @classmethod
    def range_coder(cls, freqs, DATA_BLOCK_SIZE_BITS=32, PRECISION=32):
        return cls(CODER_RANGE, freqs, DATA_BLOCK_SIZE_BITS, PRECISION)

    @classmethod
    def aec(cls, freqs_initial, DATA_BLOCK_SIZE_BITS=32, PRECISION=32, model=MODEL_ADAPTIVE_IID, max_allowed_total_freq=None, k=0):
        """model=MODEL_ORDER_K: `freqs_initial` only gives the alphabet size; the model table passed to
        encode_block/decode_block is [n_sym**k * n_sym counts][context index] (uint64)."""
        if max_allowed_total_freq is None:
            max_allowed_total_freq = 1 << (PRECISION - 2)
        return cls(CODER_AEC, freqs_initial, DATA_BLOCK_SIZE_BITS, PRECISION, model | (k << 8), max_allowed_total_freq)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().scl_oracle_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # -- single block --------------------------------------------------------------------
    def encode_block(self, sym, model_freq=None, cap_bytes=None):
        sym = np.ascontiguousarray(np.asarray(sym, dtype=np.uint8))
        n = int(sym.size)
        cap = cap_bytes or (n * 8 + 64)
        out = np.zeros(cap, dtype=np.uint8)
        nbits = ctypes.c_uint64(0)
        rc = lib().scl_oracle_encode_block(self.h, _p(sym, ctypes.c_uint8), n, _p(model_freq, ctypes.c_uint64), _p(out, ctypes.c_uint8), cap, ctypes.byref(nbits))
        if rc:
            raise OracleError(rc)
        nb = int(nbits.value)
        return out[: (nb + 7) // 8].copy(), nb

    def decode_block(self, packed, nbits, bit_offset=0, model_freq=None, cap=None):
        packed = np.ascontiguousarray(np.asarray(packed, dtype=np.uint8))
        cap = cap or max(16, int(nbits) * 8)
        out = np.zeros(cap, dtype=np.uint8)
        n = ctypes.c_uint64(0)
        used = ctypes.c_uint64(0)
        rc = lib().scl_oracle_decode_block(self.h, _p(packed, ctypes.c_uint8), int(bit_offset), int(nbits), _p(model_freq, ctypes.c_uint64), _p(out, ctypes.c_uint8), cap, ctypes.byref(n), ctypes.byref(used))
        if rc:
            raise OracleError(rc)
        return out[: int(n.value)].copy(), int(used.value)

    # -- batches (OpenMP) ----------------------------------------------------------------
    def encode_batch(self, sym2d, sizes=None, out_stride=None, n_threads=0):
        sym2d = np.ascontiguousarray(sym2d, dtype=np.uint8)
        nb, blen = sym2d.shape
        out_stride = out_stride or (blen * 2 + 32)
        out = np.zeros((nb, out_stride), dtype=np.uint8)
        bits = np.zeros(nb, dtype=np.uint64)
        status = np.zeros(nb, dtype=np.int32)
        sz = None if sizes is None else np.ascontiguousarray(sizes, dtype=np.uint32)
        lib().scl_oracle_encode_batch(self.h, _p(sym2d, ctypes.c_uint8), blen, _p(sz, ctypes.c_uint32), blen, nb, _p(out, ctypes.c_uint8), out_stride, _p(bits, ctypes.c_uint64), _p(status, ctypes.c_int32), n_threads)
        return out, bits, status

    def decode_batch(self, buf, bit_offsets, bit_lens, out_stride, n_threads=0):
        buf = np.ascontiguousarray(buf, dtype=np.uint8).reshape(-1)
        bit_offsets = np.ascontiguousarray(bit_offsets, dtype=np.uint64)
        bit_lens = np.ascontiguousarray(bit_lens, dtype=np.uint64)
        nb = bit_offsets.size
        out = np.zeros((nb, out_stride), dtype=np.uint8)
        sizes = np.zeros(nb, dtype=np.uint32)
        used = np.zeros(nb, dtype=np.uint64)
        status = np.zeros(nb, dtype=np.int32)
        lib().scl_oracle_decode_batch(self.h, _p(buf, ctypes.c_uint8), _p(bit_offsets, ctypes.c_uint64), _p(bit_lens, ctypes.c_uint64), nb, _p(out, ctypes.c_uint8), out_stride, _p(sizes, ctypes.c_uint32), _p(used, ctypes.c_uint64), _p(status, ctypes.c_int32), n_threads)
        return out, sizes, used, status

    def tans_tables_for(self, L):
        enc = np.zeros(L, dtype=np.uint64)
        row = np.zeros(self.n_sym, dtype=np.uint64)
        nb = np.zeros(self.n_sym, dtype=np.uint32)
        th = np.zeros(self.n_sym, dtype=np.uint64)
        ds = np.zeros(L, dtype=np.uint32)
        dx = np.zeros(L, dtype=np.uint64)
        rc = lib().scl_oracle_tans_tables(self.h, _p(enc, ctypes.c_uint64), _p(row, ctypes.c_uint64), _p(nb, ctypes.c_uint32), _p(th, ctypes.c_uint64), _p(ds, ctypes.c_uint32), _p(dx, ctypes.c_uint64))
        if rc:
            raise OracleError(rc)
        return dict(enc_table=enc, enc_row=row, nbits_base=nb, thresh=th, dec_sym=ds, dec_shrunk=dx)


def max_threads():
    return int(lib().scl_oracle_max_threads())


def bits_to_str(packed, nbits):
    """'0101...' string of a packed stream (for comparison with the reference's KAT literals)."""
    b = np.unpackbits(np.asarray(packed, dtype=np.uint8))[:nbits]
    return "".join("1" if x else "0" for x in b)
