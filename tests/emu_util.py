"""ctypes driver for tests/host_emu/libscl_emu.so (CPU lane emulation; test infrastructure)."""
import ctypes
import os
import subprocess

import numpy as np

from stanford_compression_library_b200._cabi import SclParams

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "host_emu", "emu.cpp")
_SO = os.path.join(_HERE, "host_emu", "libscl_emu.so")
_CSRC = os.path.join(os.path.dirname(_HERE), "stanford_compression_library_b200", "csrc")
_DEPS = [_SRC] + [os.path.join(_CSRC, f) for f in ("scl_lane.cuh", "scl_fast.cuh", "scl_aec.cuh", "scl_defs.h", "scl_tables.hpp")]


def build():
    if os.path.exists(_SO) and all(os.path.getmtime(d) <= os.path.getmtime(_SO) for d in _DEPS):
        return _SO
    subprocess.check_call(["/usr/bin/g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                           "-fsanitize=undefined", "-fno-sanitize-recover=undefined", "-static-libubsan", "-o", _SO, _SRC])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        vp, u32, u64 = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint64
        L.emu_create.argtypes = [ctypes.POINTER(SclParams), vp, vp, u32, ctypes.POINTER(vp)]
        L.emu_destroy.argtypes = [vp]
        L.emu_path.argtypes = [vp, ctypes.c_int]
        L.emu_force_generic.argtypes = [vp]
        L.emu_max_encoded_bytes.restype = u64
        L.emu_max_encoded_bytes.argtypes = [vp, u64]
        L.emu_tans_tables.argtypes = [vp, vp, vp, u64]
        L.emu_encode_blocks.argtypes = [vp, vp, u64, vp, u32, u64, vp, u64, vp, vp, vp, vp]
        L.emu_decode_blocks.argtypes = [vp, vp, u64, vp, vp, u64, vp, u64, vp, vp, vp, vp]
        L.emu_set_aec2.argtypes = [vp, ctypes.c_int]
        L.emu_aec_model8_ok.argtypes = [vp, u64]
        L.emu_aec_renorm_counts.argtypes = [u32, u64, u64, vp, vp, vp, vp]
        L.emu_v2_eligible.argtypes = [vp]
        L.emu_encode_blocks_v2.argtypes = [vp, vp, u64, u32, u64, vp, u64, vp, vp, vp]
        L.emu_decode_blocks_v2.argtypes = [vp, vp, u64, vp, vp, u64, vp, u64, vp, vp, vp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def aligned_zeros(n, dtype=np.uint8, align=32):
    raw = np.zeros(n * np.dtype(dtype).itemsize + align, dtype=np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off : off + n * np.dtype(dtype).itemsize].view(dtype)


class EmuCoder:
    def __init__(self, params: SclParams, alphabet, freq):
        alphabet = None if alphabet is None else np.ascontiguousarray(alphabet, dtype=np.uint8)
        freq = np.ascontiguousarray(freq, dtype=np.uint64)
        self.n_sym = freq.size
        h = ctypes.c_void_p()
        rc = lib().emu_create(ctypes.byref(params), _p(alphabet), _p(freq), freq.size, ctypes.byref(h))
        if rc == 3:  # SCL_E_UNSUPPORTED
            raise NotImplementedError("emu_create: outside the backend's limits")
        if rc:
            raise ValueError("emu_create rc=%d" % rc)
        self.h = h

    def __del__(self):
        if getattr(self, "h", None):
            lib().emu_destroy(self.h)
            self.h = None

    def path(self, decode):
        return lib().emu_path(self.h, int(decode))

    def force_generic(self):
        lib().emu_force_generic(self.h)

    def tans_tables(self, L):
        enc = np.zeros(L, dtype=np.uint32)
        dec = np.zeros(L, dtype=np.uint32)
        assert lib().emu_tans_tables(self.h, _p(enc), _p(dec), L) == 0
        return enc, dec

    def set_aec2(self, on=True):
        """1 / True = second-generation arithmetic lanes (16-bit counters), 2 = 8-bit counters where eligible"""
        lib().emu_set_aec2(self.h, int(on))

    def aec_model8_ok(self, block_len):
        return bool(lib().emu_aec_model8_ok(self.h, int(block_len)))

    def v2_eligible(self):
        return bool(lib().emu_v2_eligible(self.h))

    def encode_v2(self, sym2d, out_stride=None):
        sym2d = np.ascontiguousarray(sym2d, dtype=np.uint8)
        B, N = sym2d.shape
        stride_in = max(16, (N + 15) // 16 * 16)
        symbuf = aligned_zeros(B * stride_in)
        symbuf.reshape(B, stride_in)[:, :N] = sym2d
        stride = out_stride or int(lib().emu_max_encoded_bytes(self.h, N))
        out = aligned_zeros(B * stride + 32)
        off = np.zeros(B, dtype=np.uint64)
        ln = np.zeros(B, dtype=np.uint64)
        st = np.zeros(B, dtype=np.uint32)
        rc = lib().emu_encode_blocks_v2(self.h, _p(symbuf), stride_in, N, B, _p(out), stride, _p(off), _p(ln), _p(st))
        assert rc == 0, rc
        return out, off, ln, st

    def decode_v2(self, buf, bit_off, bit_len, max_len):
        B = len(bit_off)
        stride = max(32, (max_len + 31) // 32 * 32)
        bufa = aligned_zeros(buf.size)
        bufa[:] = buf
        sym = aligned_zeros(B * stride).reshape(B, stride)
        sizes = np.zeros(B, dtype=np.uint32)
        used = np.zeros(B, dtype=np.uint64)
        st = np.zeros(B, dtype=np.uint32)
        bo = np.ascontiguousarray(bit_off, dtype=np.uint64)
        bl = None if bit_len is None else np.ascontiguousarray(bit_len, dtype=np.uint64)
        rc = lib().emu_decode_blocks_v2(self.h, _p(bufa), bufa.size, _p(bo), _p(bl), B, _p(sym), stride, _p(sizes), _p(used), _p(st))
        assert rc == 0, rc
        return sym, sizes, used, st

    def encode(self, sym2d, sizes=None, model=None, out_stride=None):
        sym2d = np.ascontiguousarray(sym2d, dtype=np.uint8)
        B, N = sym2d.shape
        # kernels read rows with 16-byte vector loads when aligned: keep the test rows aligned too
        stride_in = max(16, (N + 15) // 16 * 16)
        symbuf = aligned_zeros(B * stride_in)
        symbuf.reshape(B, stride_in)[:, :N] = sym2d
        stride = out_stride or int(lib().emu_max_encoded_bytes(self.h, N))
        out = aligned_zeros(B * stride + 16)
        off = np.zeros(B, dtype=np.uint64)
        ln = np.zeros(B, dtype=np.uint64)
        st = np.zeros(B, dtype=np.uint32)
        sz = None if sizes is None else np.ascontiguousarray(sizes, dtype=np.uint32)
        rc = lib().emu_encode_blocks(self.h, _p(symbuf), stride_in, _p(sz), N, B, _p(out), stride, _p(off), _p(ln), _p(model), _p(st))
        assert rc == 0
        return out, off, ln, st

    def decode(self, buf, bit_off, bit_len, max_len, model=None):
        B = len(bit_off)
        stride = max(16, (max_len + 15) // 16 * 16)
        bufa = aligned_zeros(buf.size)
        bufa[:] = buf
        sym = aligned_zeros(B * stride).reshape(B, stride)
        sizes = np.zeros(B, dtype=np.uint32)
        used = np.zeros(B, dtype=np.uint64)
        st = np.zeros(B, dtype=np.uint32)
        bo = np.ascontiguousarray(bit_off, dtype=np.uint64)
        bl = None if bit_len is None else np.ascontiguousarray(bit_len, dtype=np.uint64)
        rc = lib().emu_decode_blocks(self.h, _p(bufa), bufa.size, _p(bo), _p(bl), B, _p(sym), stride, _p(sizes), _p(used), _p(model), _p(st))
        assert rc == 0
        return sym, sizes, used, st


def extract_bits(buf, off, n):
    """bits [off, off+n) of an MSB-first packed buffer, re-packed left-aligned"""
    first, last = int(off) >> 3, (int(off) + int(n) + 7) >> 3
    bits = np.unpackbits(buf[first:last])[int(off) - 8 * first :][: int(n)]
    return np.packbits(bits)


def params_from_case(c):
    from stanford_compression_library_b200 import _cabi

    p = c["params"]
    if c["coder"] in ("rans", "tans"):
        return SclParams(coder=_cabi.CODER_TANS if c["coder"] == "tans" else _cabi.CODER_RANS, data_block_size_bits=p["DATA_BLOCK_SIZE_BITS"],
                         num_bits_out=p["NUM_BITS_OUT"], range_factor=p["RANGE_FACTOR"], num_state_bits=p["NUM_STATE_BITS"], precision=0, model=0,
                         max_allowed_total_freq=0)
    if c["coder"] == "range":
        return SclParams(coder=_cabi.CODER_RANGE, data_block_size_bits=p["DATA_BLOCK_SIZE_BITS"], num_bits_out=0, range_factor=0, num_state_bits=0,
                         precision=p["PRECISION"], model=0, max_allowed_total_freq=0)
    kind = {"adaptive_iid": _cabi.MODEL_ADAPTIVE_IID, "order_k": _cabi.MODEL_ORDER_K}.get(c["model"]["kind"], _cabi.MODEL_FIXED)
    return SclParams(coder=_cabi.CODER_AEC, data_block_size_bits=p["DATA_BLOCK_SIZE_BITS"], num_bits_out=0, range_factor=0, num_state_bits=0,
                     precision=p["PRECISION"], model=kind, model_order=c["model"].get("k", 0), max_allowed_total_freq=c["model"]["max_total"])
