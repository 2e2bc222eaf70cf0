"""Device-side objects: a coder handle (tables in HBM) and the batched encoded/decoded forms.

PyTorch is used only for device memory, streams and (in sharding.py) torch.distributed; all
coding work is done by the CUDA kernels behind the C-ABI (include/scl_b200.h).
"""
from dataclasses import dataclass
import ctypes

import numpy as np
import torch

from . import _cabi
from .utils.bitarray_utils import BitArray

_STATUS_EXC = {
    _cabi.ST_BAD_SYMBOL: (KeyError, "symbol not in the Frequencies alphabet"),
    _cabi.ST_STATE_MISMATCH: (AssertionError, "rANS/tANS end state != INITIAL_STATE (corrupt stream)"),
    _cabi.ST_OVERFLOW: (OverflowError, "block size does not fit DATA_BLOCK_SIZE_BITS, or the output slot is too small"),
    _cabi.ST_TRUNCATED: (ValueError, "encoded stream ended inside a block"),
    _cabi.ST_TOTAL_FREQ: (AssertionError, "the frequency total is too large (>= MAX_ALLOWED_TOTAL_FREQ)"),
    _cabi.ST_EMPTY_BLOCK: (ValueError, "arithmetic decoder: a block of size 0 cannot be decoded (the reference does not terminate on it)"),
}


def require_cuda():
    _cabi.lib()  # raises BackendUnavailable if the extension is not built
    if not torch.cuda.is_available():
        raise _cabi.BackendUnavailable("no CUDA device: this backend has no CPU fallback")


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def raise_for_status(status: torch.Tensor):
    """Turn per-block status words into the exception the reference would have raised."""
    bad = torch.nonzero(status)
    if bad.numel() == 0:
        return
    b = int(bad[0])
    code = int(status[b])
    exc, msg = _STATUS_EXC.get(code, (RuntimeError, "status %d" % code))
    raise exc("block %d: %s" % (b, msg))


@dataclass
class EncodedBlocks:
    """Output of `encode_blocks`: per-block bit streams inside one device byte buffer.

    Block b is bits [bit_offset[b], bit_offset[b] + bit_len[b]) of `buf` (MSB-first).
    `byte_offset` (int64 [B + 1], packed / framed layouts only): where block b's record starts in `buf`;
    the last entry is the total number of bytes.
    """

    buf: torch.Tensor  # uint8 [n_bytes], device
    bit_offset: torch.Tensor  # int64 [B]
    bit_len: torch.Tensor  # int64 [B]
    status: torch.Tensor  # int32 [B]
    out_stride: int = 0
    byte_offset: torch.Tensor = None
    framed: bool = False
    _scratch: object = None  # buffers a reusing call writes again (encode_blocks_packed)

    @property
    def n_blocks(self):
        return int(self.bit_len.numel())

    def check(self):
        if self.status is not None:
            raise_for_status(self.status)
        return self

    def total_bytes(self) -> int:
        """Sum over blocks of ceil(bit_len / 8): the `C` of the roofline accounting."""
        if self.byte_offset is not None and not self.framed:
            return int(self.byte_offset[-1])
        return int(((self.bit_len + 7) // 8).sum())

    def block(self, b: int) -> BitArray:
        off, n = int(self.bit_offset[b]), int(self.bit_len[b])
        first, last = off >> 3, (off + n + 7) >> 3
        host = self.buf[first:last].cpu().numpy()
        return BitArray.from_packed(host, n, off - 8 * first)

    def packed_offsets(self, framed: bool = False):
        """(byte offsets int64 [B + 1], bit offsets int64 [B]) of the packed / framed layout: a device scan
        (scl_packed_offsets), no host round trip."""
        B = self.n_blocks
        dev = self.buf.device
        byte_off = torch.empty(B + 1, dtype=torch.int64, device=dev)
        bit_off = torch.empty(B, dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            rc = _cabi.lib().scl_packed_offsets(_ptr(self.bit_len), None, B, 1 if framed else 0, _ptr(byte_off), _ptr(bit_off), _stream())
        _cabi.check(rc, "scl_packed_offsets")
        return byte_off, bit_off

    def _repack(self, framed: bool, bytewise: bool):
        byte_off, bit_off = self.packed_offsets(framed)
        total = int(byte_off[-1])
        dst = torch.empty(total + 16, dtype=torch.uint8, device=self.buf.device)  # the kernel writes every byte below `total`
        dst[total:].zero_()
        fn = _cabi.lib().scl_frame_blocks if framed else _cabi.lib().scl_pack_blocks
        with torch.cuda.device(self.buf.device):
            rc = fn(_ptr(self.buf), _ptr(self.bit_offset), _ptr(self.bit_len), self.n_blocks, _ptr(dst), _ptr(byte_off), 1 if bytewise else 0, _stream())
        _cabi.check(rc, "scl_frame_blocks" if framed else "scl_pack_blocks")
        return dst, byte_off, bit_off, total

    def pack(self, bytewise: bool = False) -> "EncodedBlocks":
        """Contiguous, byte-aligned, left-aligned streams == concatenated BitArray.tobytes().
        (`encode_blocks_packed` produces this form directly, in the encode launch itself.)
        bytewise=True takes the first-generation copy kernel (tests)."""
        dst, byte_off, bit_off, _ = self._repack(False, bytewise)
        return EncodedBlocks(dst, bit_off, self.bit_len, self.status, 0, byte_off, False)

    def frame(self, bytewise: bool = False):
        """Bytes of the reference's EncodedBlockWriter file format (encoded_stream.py:150-175).

        Returns (uint8 device tensor, int64 byte offsets [B+1])."""
        if self.framed:
            return self.buf[: int(self.byte_offset[-1])], self.byte_offset
        dst, byte_off, _, total = self._repack(True, bytewise)
        return dst[:total], byte_off

    @classmethod
    def from_bitarrays(cls, blocks, device="cuda"):
        """Host BitArrays -> device buffer (each block byte-aligned, 16 B of slack at the end)."""
        packed = [np.frombuffer(b.tobytes(), dtype=np.uint8) for b in blocks]
        lens = np.array([len(b) for b in blocks], dtype=np.int64)
        nbytes = np.array([p.size for p in packed], dtype=np.int64)
        offs = np.concatenate([[0], np.cumsum(nbytes)[:-1]]).astype(np.int64) if len(blocks) else np.zeros(0, dtype=np.int64)
        host = np.concatenate(packed + [np.zeros(16, dtype=np.uint8)]) if packed else np.zeros(16, dtype=np.uint8)
        return cls(torch.from_numpy(host).to(device), torch.from_numpy(offs * 8).to(device), torch.from_numpy(lens).to(device),
                   torch.zeros(len(blocks), dtype=torch.int32, device=device), 0)


@dataclass
class DecodedBlocks:
    symbols: torch.Tensor  # uint8 [B, stride]
    sizes: torch.Tensor  # int32 [B]
    bits_consumed: torch.Tensor  # int64 [B]
    status: torch.Tensor  # int32 [B]

    def check(self):
        raise_for_status(self.status)
        return self


class DeviceCoder:
    """Owns one `scl_coder` handle (parameters + lookup tables resident in HBM)."""

    def __init__(self, params: _cabi.SclParams, alphabet: np.ndarray, freq: np.ndarray, device=None):
        require_cuda()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.n_sym = int(freq.size)
        alphabet = np.ascontiguousarray(alphabet, dtype=np.uint8)
        freq = np.ascontiguousarray(freq, dtype=np.uint64)
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            rc = _cabi.lib().scl_coder_create(ctypes.byref(params), alphabet.ctypes.data_as(ctypes.c_void_p), freq.ctypes.data_as(ctypes.c_void_p),
                                              self.n_sym, _stream(), ctypes.byref(h))
        _cabi.check(rc, "scl_coder_create")
        self._h = h
        self.params = params

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _cabi.lib().scl_coder_destroy(h)
            except Exception:
                pass

    def max_encoded_bytes(self, block_len: int) -> int:
        return int(_cabi.lib().scl_coder_max_encoded_bytes(self._h, int(block_len)))

    def path(self, decode: bool) -> str:
        return "fast32" if _cabi.lib().scl_coder_path(self._h, 1 if decode else 0) == 0 else "generic64"

    def _to_device(self, t, dtype):
        if not isinstance(t, torch.Tensor):
            t = torch.as_tensor(np.ascontiguousarray(t))
        if t.dtype != dtype:
            t = t.to(dtype)
        return t.to(self.device, non_blocking=True).contiguous()

    def encode_blocks(self, data, sizes=None, model=None, out_stride=None, reuse: EncodedBlocks = None) -> EncodedBlocks:
        """data: uint8 [B, N] (device, or host -- copied).  sizes: optional int32 [B] (ragged).
        `reuse`: an EncodedBlocks from an earlier call of the same shape whose buffers are overwritten
        (no allocation in the call)."""
        data = self._to_device(data, torch.uint8)
        if data.dim() == 1:
            data = data[None, :]
        B, N = data.shape
        if sizes is not None:
            sizes = self._to_device(sizes, torch.int32)
        if reuse is not None:
            stride, buf, bit_off, bit_len, status = reuse.out_stride, reuse.buf, reuse.bit_offset, reuse.bit_len, reuse.status
            assert buf.numel() >= B * stride and bit_len.numel() == B
        else:
            stride = out_stride or self.max_encoded_bytes(N)
            buf = torch.empty(B * stride + 16, dtype=torch.uint8, device=self.device)
            bit_off = torch.empty(B, dtype=torch.int64, device=self.device)
            bit_len = torch.empty(B, dtype=torch.int64, device=self.device)
            status = torch.empty(B, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            rc = _cabi.lib().scl_encode_blocks(self._h, _ptr(data) if N else _ptr(buf), data.stride(0) if N else 0, _ptr(sizes), N, B, _ptr(buf), stride,
                                               _ptr(bit_off), _ptr(bit_len), _ptr(model), _ptr(status), _stream())
        _cabi.check(rc, "scl_encode_blocks")
        return reuse if reuse is not None else EncodedBlocks(buf, bit_off, bit_len, status, stride)

    def encode_blocks_packed(self, data, sizes=None, model=None, framed: bool = False, capacity: int = None, reuse: EncodedBlocks = None) -> EncodedBlocks:
        """encode_blocks + the contiguous output of the reference's writer in ONE call (scl_encode_blocks_packed):
        `buf` is b"".join(encode_block(b).tobytes()) -- or, framed=True, the bytes of the EncodedBlockWriter file
        (encoded_stream.py:150-175) -- `byte_offset[B]` its length, `bit_offset` what decode_blocks wants.
        capacity: bytes to reserve for `buf` (default: the worst case, B * max_encoded_bytes (+ 5 framed));
        `reuse`: the EncodedBlocks of an earlier call of the same shape (no allocation in the call)."""
        data = self._to_device(data, torch.uint8)
        if data.dim() == 1:
            data = data[None, :]
        B, N = data.shape
        if sizes is not None:
            sizes = self._to_device(sizes, torch.int32)
        lib = _cabi.lib()
        if reuse is not None:
            assert reuse._scratch is not None and reuse.bit_len.numel() == B and reuse.framed == bool(framed)
            scratch, stride, ws = reuse._scratch
            dst, byte_off, bit_off, bit_len, status = reuse.buf, reuse.byte_offset, reuse.bit_offset, reuse.bit_len, reuse.status
        else:
            stride = self.max_encoded_bytes(N)
            scratch = torch.empty(B * stride + 16, dtype=torch.uint8, device=self.device)
            cap = int(capacity) if capacity is not None else B * (stride + (5 if framed else 0))
            dst = torch.empty(cap + 16, dtype=torch.uint8, device=self.device)
            ws = torch.empty(max(1, int(lib.scl_encode_packed_workspace_bytes(self._h, B))), dtype=torch.uint8, device=self.device)
            byte_off = torch.empty(B + 1, dtype=torch.int64, device=self.device)
            bit_off = torch.empty(B, dtype=torch.int64, device=self.device)
            bit_len = torch.empty(B, dtype=torch.int64, device=self.device)
            status = torch.empty(B, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            rc = lib.scl_encode_blocks_packed(self._h, _ptr(data) if N else _ptr(scratch), data.stride(0) if N else 0, _ptr(sizes), N, B, _ptr(scratch), stride,
                                              _ptr(dst), dst.numel() - 16, 1 if framed else 0, _ptr(byte_off), _ptr(bit_off), _ptr(bit_len), _ptr(model),
                                              _ptr(status), _ptr(ws), ws.numel(), _stream())
        _cabi.check(rc, "scl_encode_blocks_packed")
        return reuse if reuse is not None else EncodedBlocks(dst, bit_off, bit_len, status, 0, byte_off, bool(framed), (scratch, stride, ws))

    def debug_path(self, mode: int):
        """Test hook (scl_coder_debug_path): 1 = first-generation kernels, 2 = v2 decode with sector stores,
        3 / 4 = v2 decode always / never pipe-balanced, 0 = default.  Per handle."""
        _cabi.lib().scl_coder_debug_path(self._h, int(mode))

    def debug_trace(self, trace: torch.Tensor = None):
        """Diagnostic hook (scl_coder_debug_trace): an int64 device tensor of >= SMs * 32 * 40 zeros receives the fused
        packed encoder's per-warp timestamps; None switches it off.  Keep the tensor alive while it is set."""
        self._trace = trace
        _cabi.lib().scl_coder_debug_trace(self._h, _ptr(trace), 0 if trace is None else trace.numel())

    def decode_blocks(self, enc: EncodedBlocks, max_block_len: int, model=None, out=None, reuse: DecodedBlocks = None) -> DecodedBlocks:
        B = enc.n_blocks
        if reuse is not None:
            out, sizes, used, status = reuse.symbols, reuse.sizes, reuse.bits_consumed, reuse.status
            stride = out.stride(0)
        else:
            if out is None:
                stride = max((int(max_block_len) + 31) // 32 * 32, 32)
                out = torch.empty((B, stride), dtype=torch.uint8, device=self.device)
            else:
                # a caller-supplied destination is used as it is (the row stride is also the capacity the
                # decoder checks the header's size against), never silently re-strided
                if not (isinstance(out, torch.Tensor) and out.dtype == torch.uint8 and out.device == self.device and out.dim() == 2):
                    raise ValueError("out must be a 2-D uint8 tensor on %s" % self.device)
                if out.shape[0] < B or not out.is_contiguous() or out.shape[1] < 1:
                    raise ValueError("out must be contiguous with at least %d rows" % B)
                stride = out.shape[1]
            sizes = torch.empty(B, dtype=torch.int32, device=self.device)
            used = torch.empty(B, dtype=torch.int64, device=self.device)
            status = torch.empty(B, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            rc = _cabi.lib().scl_decode_blocks(self._h, _ptr(enc.buf), enc.buf.numel(), _ptr(enc.bit_offset), _ptr(enc.bit_len), B, _ptr(out), stride,
                                               _ptr(sizes), _ptr(used), _ptr(model), _ptr(status), _stream())
        _cabi.check(rc, "scl_decode_blocks")
        return reuse if reuse is not None else DecodedBlocks(out, sizes, used, status)

    def tans_tables(self, L: int):
        enc = np.zeros(L, dtype=np.uint32)
        dec = np.zeros(L, dtype=np.uint32)
        with torch.cuda.device(self.device):
            rc = _cabi.lib().scl_tans_tables_to_host(self._h, enc.ctypes.data_as(ctypes.c_void_p), dec.ctypes.data_as(ctypes.c_void_p), L, _stream())
        _cabi.check(rc, "scl_tans_tables_to_host")
        return enc, dec
