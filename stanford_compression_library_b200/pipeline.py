"""Host-buffer codec pipeline: the end-to-end form of the batched API.

`encode_blocks` / `decode_blocks` work on device tensors.  When the data lives in HOST memory the
PCIe copies dominate (a B200 codes ~1 TB/s, a PCIe 5 x16 link moves ~55 GB/s per direction), so the
useful thing to do is overlap them: the batch is cut into chunks of blocks and each chunk runs
H2D -> kernel(s) -> D2H on one of two CUDA streams, so that one chunk's upload, another's coding and
a third's download are in flight together (the two copy directions use separate DMA engines).

Outputs are the packed form: all streams concatenated, each starting on a byte boundary
(== b"".join(encode_block(b).tobytes())), plus the per-block bit lengths.
"""
import torch

from .device import EncodedBlocks


class HostCodecPipeline:
    def __init__(self, encoder, decoder, block_len: int, n_blocks: int, chunk_blocks: int = 32768, device=None):
        self.enc, self.dec = encoder, decoder
        self.N, self.B = int(block_len), int(n_blocks)
        self.chunk = min(int(chunk_blocks), self.B)
        self.dev = encoder.device_coder().device if device is None else torch.device(device)
        self.n_chunks = (self.B + self.chunk - 1) // self.chunk
        self.streams = [torch.cuda.Stream(device=self.dev) for _ in range(2)]
        dc = encoder.device_coder()
        self.stride = dc.max_encoded_bytes(self.N)
        # per-stream device staging (allocated once: nothing is allocated inside encode()/decode())
        self._d_raw = [torch.empty((self.chunk, self.N), dtype=torch.uint8, device=self.dev) for _ in range(2)]
        self._d_enc = [None, None]
        self._d_dec = [None, None]
        self._d_packed = [torch.empty(self.chunk * self.stride + 64, dtype=torch.uint8, device=self.dev) for _ in range(2)]
        self._d_offs = [torch.empty(self.chunk, dtype=torch.int64, device=self.dev) for _ in range(2)]
        self._h_len = torch.empty(self.B, dtype=torch.int64, pin_memory=True)
        self._h_status = torch.empty(self.B, dtype=torch.int32, pin_memory=True)
        self._h_sizes = torch.empty(self.B, dtype=torch.int32, pin_memory=True)
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def max_packed_bytes(self) -> int:
        return self.B * self.stride

    # ------------------------------------------------------------------------------------------
    def encode(self, host_raw: torch.Tensor, host_packed: torch.Tensor):
        """host_raw: pinned uint8 [B, N]; host_packed: pinned uint8 [>= total coded bytes].
        Returns (total_bytes, bit_len[B] on the host)."""
        from . import _cabi
        from .device import _ptr

        lib = _cabi.lib()
        for st in self.streams:
            st.wait_stream(torch.cuda.current_stream(self.dev))
        total = 0
        pending = None  # (chunk index, stream slot, packed view length event)
        self.h2d_bytes = self.d2h_bytes = 0

        def finish(k, slot, lo, hi, ev):
            nonlocal total
            ev.synchronize()  # bit lengths of chunk k are on the host: we now know how many bytes to fetch
            nbytes = (self._h_len[lo:hi] + 7) // 8
            n = int(nbytes.sum())
            with torch.cuda.stream(self.streams[slot]):
                host_packed[total : total + n].copy_(self._d_packed[slot][:n], non_blocking=True)
            self.d2h_bytes += n + (hi - lo) * 12
            total += n

        for k in range(self.n_chunks):
            slot = k & 1
            lo, hi = k * self.chunk, min(self.B, (k + 1) * self.chunk)
            nb = hi - lo
            st = self.streams[slot]
            with torch.cuda.stream(st):
                d_raw = self._d_raw[slot][:nb]
                d_raw.copy_(host_raw[lo:hi], non_blocking=True)
                e = self.enc.device_coder().encode_blocks(d_raw, reuse=self._d_enc[slot] if self._d_enc[slot] is not None and self._d_enc[slot].n_blocks == nb else None)
                self._d_enc[slot] = e
                nbytes = (e.bit_len + 7) // 8
                offs = torch.cumsum(nbytes, 0) - nbytes
                rc = lib.scl_pack_blocks(_ptr(e.buf), _ptr(e.bit_offset), _ptr(e.bit_len), nb, _ptr(self._d_packed[slot]), _ptr(offs),
                                         torch.cuda.current_stream().cuda_stream)
                _cabi.check(rc, "scl_pack_blocks")
                self._h_len[lo:hi].copy_(e.bit_len, non_blocking=True)
                self._h_status[lo:hi].copy_(e.status, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
            self.h2d_bytes += nb * self.N
            if pending is not None:
                finish(*pending)
            pending = (k, slot, lo, hi, ev)
        finish(*pending)
        for st in self.streams:
            st.synchronize()
        if int(self._h_status.abs().sum()) != 0:
            from .device import raise_for_status

            raise_for_status(self._h_status)
        return total, self._h_len

    # ------------------------------------------------------------------------------------------
    def decode(self, host_packed: torch.Tensor, host_bit_len: torch.Tensor, host_out: torch.Tensor):
        """Inverse of encode(): host_packed/bit_len as produced above -> host_out pinned uint8 [B, N]."""
        nbytes_all = (host_bit_len + 7) // 8
        ends = torch.cumsum(nbytes_all, 0)
        self.h2d_bytes = self.d2h_bytes = 0
        for st in self.streams:
            st.wait_stream(torch.cuda.current_stream(self.dev))
        for k in range(self.n_chunks):
            slot = k & 1
            lo, hi = k * self.chunk, min(self.B, (k + 1) * self.chunk)
            nb = hi - lo
            b0 = int(ends[lo - 1]) if lo else 0
            b1 = int(ends[hi - 1])
            st = self.streams[slot]
            with torch.cuda.stream(st):
                d_c = self._d_packed[slot]
                d_c[: b1 - b0].copy_(host_packed[b0:b1], non_blocking=True)
                lens = self._d_offs[slot][:nb]
                lens.copy_(host_bit_len[lo:hi], non_blocking=True)
                nbytes = (lens + 7) // 8
                offs = (torch.cumsum(nbytes, 0) - nbytes) * 8
                enc = EncodedBlocks(d_c, offs, lens, None, 0)
                reuse = self._d_dec[slot] if self._d_dec[slot] is not None and self._d_dec[slot].sizes.numel() == nb else None
                d = self.dec.device_coder().decode_blocks(enc, self.N, reuse=reuse)
                self._d_dec[slot] = d
                host_out[lo:hi].copy_(d.symbols[:, : self.N], non_blocking=True)
                self._h_status[lo:hi].copy_(d.status, non_blocking=True)
                self._h_sizes[lo:hi].copy_(d.sizes, non_blocking=True)
            self.h2d_bytes += (b1 - b0) + nb * 8
            self.d2h_bytes += nb * self.N + nb * 8
        for st in self.streams:
            st.synchronize()
        if int(self._h_status.abs().sum()) != 0:
            from .device import raise_for_status

            raise_for_status(self._h_status)
        return host_out
