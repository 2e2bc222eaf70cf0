"""Host-buffer codec pipeline: the end-to-end form of the batched API.

`encode_blocks` / `decode_blocks` work on device tensors.  When the data lives in HOST memory the
PCIe copies dominate (a B200 codes ~1 TB/s, a PCIe 5 x16 link moves ~55 GB/s per direction, ~50 GB/s
each when both directions run: tools/measure_pcie.py), so the only useful thing to do is keep the
link busy.  The batch is cut into chunks of blocks; three CUDA streams -- uploads, kernels,
downloads -- each run their own phase of every chunk in order, tied together by events, over a ring
of `depth` staging slots.  An upload therefore never queues behind a download (it did when a chunk's
three phases shared one stream), and the busier direction of each leg runs back to back.

Outputs are the packed form: all streams concatenated, each starting on a byte boundary
(== b"".join(encode_block(b).tobytes())), plus the per-block bit lengths.
"""
import torch

from .device import EncodedBlocks


class HostCodecPipeline:
    def __init__(self, encoder, decoder, block_len: int, n_blocks: int, chunk_blocks: int = 16384, device=None, depth: int = 3):
        self.enc, self.dec = encoder, decoder
        self.N, self.B = int(block_len), int(n_blocks)
        self.chunk = max(1, min(int(chunk_blocks), self.B))
        self.dev = encoder.device_coder().device if device is None else torch.device(device)
        self.n_chunks = (self.B + self.chunk - 1) // self.chunk
        self.depth = D = max(2, int(depth))
        self.s_h2d = torch.cuda.Stream(device=self.dev)
        self.s_comp = torch.cuda.Stream(device=self.dev)
        self.s_d2h = torch.cuda.Stream(device=self.dev)
        dc = encoder.device_coder()
        self.stride = dc.max_encoded_bytes(self.N)
        # per-slot device staging (allocated once: nothing is allocated inside encode()/decode())
        self._d_raw = [torch.empty((self.chunk, self.N), dtype=torch.uint8, device=self.dev) for _ in range(D)]
        self._d_enc = [None] * D
        self._d_dec = [None] * D
        self._d_packed = [torch.empty(self.chunk * self.stride + 64, dtype=torch.uint8, device=self.dev) for _ in range(D)]
        self._d_lens = [torch.empty(self.chunk, dtype=torch.int64, device=self.dev) for _ in range(D)]
        self._d_byte_off = [torch.empty(self.chunk + 1, dtype=torch.int64, device=self.dev) for _ in range(D)]
        self._d_bit_off = [torch.empty(self.chunk, dtype=torch.int64, device=self.dev) for _ in range(D)]
        self._h_len = torch.empty(self.B, dtype=torch.int64, pin_memory=True)
        self._h_status = torch.empty(self.B, dtype=torch.int32, pin_memory=True)
        self._h_sizes = torch.empty(self.B, dtype=torch.int32, pin_memory=True)
        self._h_chunk_bytes = torch.empty(self.n_chunks, dtype=torch.int64, pin_memory=True)
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def max_packed_bytes(self) -> int:
        return self.B * self.stride

    def _begin(self):
        cur = torch.cuda.current_stream(self.dev)
        for st in (self.s_h2d, self.s_comp, self.s_d2h):
            st.wait_stream(cur)
        self.h2d_bytes = self.d2h_bytes = 0

    def _end(self):
        for st in (self.s_h2d, self.s_comp, self.s_d2h):
            st.synchronize()
        if int(self._h_status.abs().sum()) != 0:
            from .device import raise_for_status

            raise_for_status(self._h_status)

    # ------------------------------------------------------------------------------------------
    def encode(self, host_raw: torch.Tensor, host_packed: torch.Tensor):
        """host_raw: pinned uint8 [B, N]; host_packed: pinned uint8 [>= total coded bytes].
        Returns (total_bytes, bit_len[B] on the host)."""
        from . import _cabi
        from .device import _ptr

        lib = _cabi.lib()
        dc = self.enc.device_coder()
        D = self.depth
        self._begin()
        total = 0
        comp_done = [None] * D  # kernels of the chunk that last used slot s (they read d_raw[s], write d_packed[s])
        d2h_done = [None] * D   # download of the chunk that last used slot s (it reads d_packed[s])
        pending = None

        def download(slot, k, lo, hi, ev):
            # the chunk's packed size is on the host once its kernel is done: only then is the number of
            # bytes to fetch known (one 8-byte word per chunk, written by the encode kernel itself)
            nonlocal total
            ev.synchronize()
            n = int(self._h_chunk_bytes[k])
            with torch.cuda.stream(self.s_d2h):
                host_packed[total : total + n].copy_(self._d_enc[slot].buf[:n], non_blocking=True)
                d2h_done[slot] = torch.cuda.Event()
                d2h_done[slot].record()
            self.d2h_bytes += n + (hi - lo) * 12 + 8
            total += n

        for k in range(self.n_chunks):
            slot = k % D
            lo, hi = k * self.chunk, min(self.B, (k + 1) * self.chunk)
            nb = hi - lo
            d_raw = self._d_raw[slot][:nb]
            with torch.cuda.stream(self.s_h2d):
                if comp_done[slot] is not None:
                    self.s_h2d.wait_event(comp_done[slot])
                d_raw.copy_(host_raw[lo:hi], non_blocking=True)
                up = torch.cuda.Event()
                up.record()
            with torch.cuda.stream(self.s_comp):
                self.s_comp.wait_event(up)
                if d2h_done[slot] is not None:
                    self.s_comp.wait_event(d2h_done[slot])
                # one launch: symbols in, this chunk's contiguous stream + offsets + total out
                old = self._d_enc[slot]
                e = dc.encode_blocks_packed(d_raw, reuse=old if old is not None and old.n_blocks == nb else None)
                self._d_enc[slot] = e
                self._h_len[lo:hi].copy_(e.bit_len, non_blocking=True)
                self._h_status[lo:hi].copy_(e.status, non_blocking=True)
                self._h_chunk_bytes[k : k + 1].copy_(e.byte_offset[nb:], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
                comp_done[slot] = ev
            self.h2d_bytes += nb * self.N
            if pending is not None:
                download(*pending)
            pending = (slot, k, lo, hi, ev)
        if pending is not None:
            download(*pending)
        self._end()
        return total, self._h_len

    # ------------------------------------------------------------------------------------------
    def decode(self, host_packed: torch.Tensor, host_bit_len: torch.Tensor, host_out: torch.Tensor):
        """Inverse of encode(): host_packed/bit_len as produced above -> host_out pinned uint8 [B, N]."""
        from . import _cabi
        from .device import _ptr

        lib = _cabi.lib()
        nbytes_all = (host_bit_len + 7) >> 3  # host tensors: where each chunk's bytes lie in host_packed
        ends = torch.cumsum(nbytes_all, 0)
        bounds = [0] + [int(ends[min(self.B, (k + 1) * self.chunk) - 1]) for k in range(self.n_chunks)]
        dc = self.dec.device_coder()
        D = self.depth
        self._begin()
        comp_done = [None] * D  # kernels of the chunk that last used slot s (they read d_packed[s] / d_lens[s])
        d2h_done = [None] * D   # download of the chunk that last used slot s (it reads the decoded symbols)
        for k in range(self.n_chunks):
            slot = k % D
            lo, hi = k * self.chunk, min(self.B, (k + 1) * self.chunk)
            nb = hi - lo
            b0, b1 = bounds[k], bounds[k + 1]
            d_c = self._d_packed[slot]
            lens = self._d_lens[slot][:nb]
            with torch.cuda.stream(self.s_h2d):
                if comp_done[slot] is not None:
                    self.s_h2d.wait_event(comp_done[slot])
                d_c[: b1 - b0].copy_(host_packed[b0:b1], non_blocking=True)
                lens.copy_(host_bit_len[lo:hi], non_blocking=True)
                up = torch.cuda.Event()
                up.record()
            with torch.cuda.stream(self.s_comp):
                self.s_comp.wait_event(up)
                if d2h_done[slot] is not None:
                    self.s_comp.wait_event(d2h_done[slot])
                offs = self._d_bit_off[slot][:nb]
                rc = lib.scl_packed_offsets(_ptr(lens), None, nb, 0, _ptr(self._d_byte_off[slot]), _ptr(offs), torch.cuda.current_stream().cuda_stream)
                _cabi.check(rc, "scl_packed_offsets")
                enc = EncodedBlocks(d_c, offs, lens, None, 0)
                old = self._d_dec[slot]
                d = dc.decode_blocks(enc, self.N, reuse=old if old is not None and old.sizes.numel() == nb else None)
                self._d_dec[slot] = d
                done = torch.cuda.Event()
                done.record()
                comp_done[slot] = done
            with torch.cuda.stream(self.s_d2h):
                self.s_d2h.wait_event(done)
                host_out[lo:hi].copy_(d.symbols[:, : self.N], non_blocking=True)
                self._h_status[lo:hi].copy_(d.status, non_blocking=True)
                self._h_sizes[lo:hi].copy_(d.sizes, non_blocking=True)
                d2h_done[slot] = torch.cuda.Event()
                d2h_done[slot].record()
            self.h2d_bytes += (b1 - b0) + nb * 8
            self.d2h_bytes += nb * self.N + nb * 8
        self._end()
        return host_out
