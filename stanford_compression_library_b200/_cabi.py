"""ctypes binding of the C-ABI shared library (include/scl_b200.h).

The library is built in-tree (`python -m stanford_compression_library_b200.build` or
`__graft_entry__.build()`) as csrc/libscl_b200.so.  There is NO CPU fallback: if the library is
missing, or no CUDA device is present, every coder call raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libscl_b200.so")

# codes from include/scl_b200.h
E_OK, E_INVALID, E_CUDA, E_UNSUPPORTED = 0, 1, 2, 3
ST_OK, ST_BAD_SYMBOL, ST_STATE_MISMATCH, ST_OVERFLOW, ST_TRUNCATED, ST_TOTAL_FREQ, ST_EMPTY_BLOCK = 0, 1, 2, 3, 4, 6, 7
CODER_RANS, CODER_TANS, CODER_RANGE, CODER_AEC = 0, 1, 2, 3
MODEL_FIXED, MODEL_ADAPTIVE_IID, MODEL_ORDER_K = 0, 1, 2


class SclParams(ctypes.Structure):
    _fields_ = [
        ("coder", ctypes.c_int32),
        ("data_block_size_bits", ctypes.c_uint32),
        ("num_bits_out", ctypes.c_uint32),
        ("range_factor", ctypes.c_uint64),
        ("num_state_bits", ctypes.c_uint32),
        ("precision", ctypes.c_uint32),
        ("model", ctypes.c_int32),
        ("model_order", ctypes.c_uint32),
        ("max_allowed_total_freq", ctypes.c_uint64),
    ]


class BackendUnavailable(RuntimeError):
    pass


_lib = None


def lib():
    """The loaded C-ABI library; raises BackendUnavailable if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BackendUnavailable(
            "CUDA backend library not found at %s -- build it with `python -m stanford_compression_library_b200.build` "
            "(there is no CPU fallback)" % LIB_PATH
        )
    L = ctypes.CDLL(LIB_PATH)
    vp, u32, u64, i32 = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int
    L.scl_coder_create.restype = i32
    L.scl_coder_create.argtypes = [ctypes.POINTER(SclParams), vp, vp, u32, vp, ctypes.POINTER(vp)]
    L.scl_coder_destroy.restype = None
    L.scl_coder_destroy.argtypes = [vp]
    L.scl_coder_max_encoded_bytes.restype = u64
    L.scl_coder_max_encoded_bytes.argtypes = [vp, u64]
    L.scl_coder_model_words.restype = u64
    L.scl_coder_model_words.argtypes = [vp]
    L.scl_coder_path.restype = i32
    L.scl_coder_path.argtypes = [vp, i32]
    L.scl_encode_blocks.restype = i32
    L.scl_encode_blocks.argtypes = [vp, vp, u64, vp, u32, u64, vp, u64, vp, vp, vp, vp, vp]
    L.scl_decode_blocks.restype = i32
    L.scl_decode_blocks.argtypes = [vp, vp, u64, vp, vp, u64, vp, u64, vp, vp, vp, vp, vp]
    L.scl_pack_blocks.restype = i32
    L.scl_pack_blocks.argtypes = [vp, vp, vp, u64, vp, vp, u32, vp]
    L.scl_frame_blocks.restype = i32
    L.scl_frame_blocks.argtypes = [vp, vp, vp, u64, vp, vp, u32, vp]
    L.scl_packed_offsets.restype = i32
    L.scl_packed_offsets.argtypes = [vp, vp, u64, u32, vp, vp, vp]
    L.scl_encode_packed_workspace_bytes.restype = u64
    L.scl_encode_packed_workspace_bytes.argtypes = [vp, u64]
    L.scl_encode_blocks_packed.restype = i32
    L.scl_encode_blocks_packed.argtypes = [vp, vp, u64, vp, u32, u64, vp, u64, vp, u64, u32, vp, vp, vp, vp, vp, vp, u64, vp]
    L.scl_tans_tables_to_host.restype = i32
    L.scl_tans_tables_to_host.argtypes = [vp, vp, vp, u64, vp]
    L.scl_histogram_blocks.restype = i32
    L.scl_histogram_blocks.argtypes = [vp, u64, vp, u32, u64, vp, vp, vp]
    L.scl_coder_debug_path.restype = None
    L.scl_coder_debug_path.argtypes = [vp, i32]
    L.scl_debug_copy_only.restype = i32
    L.scl_debug_copy_only.argtypes = [vp, u64, vp, u64, vp, u64, u32, vp, vp, vp, vp, u32, u32, vp]
    L.scl_coder_debug_trace.restype = None
    L.scl_coder_debug_trace.argtypes = [vp, vp, u64]
    L.scl_last_cuda_error.restype = ctypes.c_char_p
    L.scl_version.restype = ctypes.c_char_p
    _lib = L
    return L


EXPORTS = [
    "scl_coder_create", "scl_coder_destroy", "scl_coder_max_encoded_bytes", "scl_coder_model_words", "scl_coder_path", "scl_encode_blocks",
    "scl_decode_blocks", "scl_packed_offsets", "scl_encode_packed_workspace_bytes", "scl_encode_blocks_packed", "scl_pack_blocks", "scl_frame_blocks", "scl_tans_tables_to_host", "scl_histogram_blocks", "scl_coder_debug_path", "scl_coder_debug_trace", "scl_debug_copy_only", "scl_last_cuda_error", "scl_version",
]


def check(rc, what):
    if rc == E_OK:
        return
    if rc == E_CUDA:
        raise RuntimeError("%s: CUDA error: %s" % (what, lib().scl_last_cuda_error().decode()))
    if rc == E_UNSUPPORTED:
        raise NotImplementedError("%s: parameters valid for the reference but outside this backend's limits (see DESIGN.md)" % what)
    raise ValueError("%s: invalid arguments (code %d)" % (what, rc))
