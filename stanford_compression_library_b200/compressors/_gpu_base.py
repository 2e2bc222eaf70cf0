"""Shared plumbing of the four GPU coders: symbol <-> byte translation for the single-block
reference API and the batched tensor API on top of a DeviceCoder handle."""
import numpy as np
import torch

from ..core.data_block import DataBlock
from ..core.prob_dist import Frequencies
from ..device import DecodedBlocks, DeviceCoder, EncodedBlocks
from ..utils.bitarray_utils import BitArray


class GpuCoderBase:
    """Mixin.  Subclasses provide `_freqs()` (the Frequencies defining the alphabet) and
    `_make_cabi_params()`; the DeviceCoder is created lazily on first use."""

    _dev = None
    _dev_key = None

    def _freqs(self) -> Frequencies:
        raise NotImplementedError

    def _make_cabi_params(self):
        raise NotImplementedError

    # ---- handle --------------------------------------------------------------------------
    def _cache_key(self):
        f = self._freqs()
        return (tuple(f.alphabet), tuple(int(x) for x in f.freq_list))

    def device_coder(self) -> DeviceCoder:
        key = self._cache_key()
        if self._dev is None or self._dev_key != key:
            alpha, freq = self._freqs().to_arrays()
            self._dev = DeviceCoder(self._make_cabi_params(), alpha, freq)
            self._dev_key = key
            self._alpha_bytes = alpha
            self._sym2byte = {s: int(b) for s, b in zip(self._freqs().alphabet, alpha)}
            self._byte2sym = {int(b): s for s, b in zip(self._freqs().alphabet, alpha)}
            self._native = self._freqs().byte_alphabet()[1]
        return self._dev

    # ---- single-block translation ----------------------------------------------------------
    def _block_to_tensor(self, data_block: DataBlock) -> torch.Tensor:
        dev = self.device_coder()
        d = data_block.data_list
        if isinstance(d, torch.Tensor) and d.dtype == torch.uint8 and self._native:
            return d.reshape(1, -1).to(dev.device)
        seq = d.tolist() if hasattr(d, "tolist") else list(d)
        try:
            arr = np.fromiter((self._sym2byte[s] for s in seq), dtype=np.uint8, count=len(seq))
        except KeyError as e:  # same error the reference raises from freq_dict[s]
            raise KeyError(e.args[0]) from None
        return torch.from_numpy(arr).reshape(1, -1).to(dev.device)

    def _row_to_block(self, row: np.ndarray) -> DataBlock:
        if self._native:
            return DataBlock(row.tolist())
        return DataBlock([self._byte2sym[int(b)] for b in row])

    # ---- batched API ---------------------------------------------------------------------
    def encode_blocks(self, data, sizes=None, reuse: EncodedBlocks = None) -> EncodedBlocks:
        """Encode B independent blocks: data uint8 [B, N] (device or host tensor / ndarray)."""
        return self.device_coder().encode_blocks(data, sizes=sizes, reuse=reuse)

    def encode_blocks_packed(self, data, sizes=None, framed: bool = False, capacity: int = None, reuse: EncodedBlocks = None) -> EncodedBlocks:
        """encode_blocks with the writer's contiguous output produced by the same launch:
        `buf[: byte_offset[-1]]` == b"".join(encode_block(b).tobytes()) (or the framed file bytes)."""
        return self.device_coder().encode_blocks_packed(data, sizes=sizes, framed=framed, capacity=capacity, reuse=reuse)

    def decode_blocks(self, enc: EncodedBlocks, max_block_len: int, out=None, reuse: DecodedBlocks = None) -> DecodedBlocks:
        """Decode B independent streams into uint8 [B, >=max_block_len]."""
        return self.device_coder().decode_blocks(enc, max_block_len, out=out, reuse=reuse)

    # ---- the reference's stream-level entry points, batched -----------------------------------
    # DataEncoder.encode / DataDecoder.decode (data_encoder_decoder.py:56-69, 131-144) call encode_block /
    # decode_block once per block; here `blocks_per_batch` blocks are pulled from the stream and coded by ONE
    # launch (the framing of EncodedBlockWriter included), which is what a device needs.  Same bytes, same
    # asserts.  Coders whose blocks are not independent (an adaptive model carried from block to block, as the
    # reference's arithmetic coder does) keep the per-block loop: `_blocks_independent()`.
    blocks_per_batch = 4096

    def _blocks_independent(self) -> bool:
        return True

    def _blocks_to_array(self, blocks):
        """list of DataBlocks -> (uint8 [n, max_len] zero-padded, sizes int32 [n]); KeyError for unknown symbols"""
        self.device_coder()
        n = len(blocks)
        sizes = np.fromiter((b.size for b in blocks), dtype=np.int32, count=n)
        out = np.zeros((n, max(1, int(sizes.max()) if n else 1)), dtype=np.uint8)
        for i, b in enumerate(blocks):
            d = b.data_list
            if self._native and isinstance(d, np.ndarray) and d.dtype == np.uint8:
                out[i, : d.size] = d
                continue
            seq = d.tolist() if hasattr(d, "tolist") else d
            try:
                out[i, : len(seq)] = np.fromiter((self._sym2byte[x] for x in seq), dtype=np.uint8, count=len(seq))
            except KeyError as e:
                raise KeyError(e.args[0]) from None
        return out, sizes

    def encode(self, data_stream, block_size: int, encode_writer):
        if self.blocks_per_batch <= 1 or not self._blocks_independent() or not hasattr(encode_writer, "write_encoded_blocks"):
            return super().encode(data_stream, block_size, encode_writer)
        while True:
            blocks = data_stream.get_block_batch(block_size, self.blocks_per_batch)
            if not blocks:
                break
            data, sizes = self._blocks_to_array(blocks)
            ragged = bool((sizes != data.shape[1]).any())
            enc = self.encode_blocks_packed(torch.from_numpy(data), sizes=torch.from_numpy(sizes) if ragged else None, framed=True).check()
            encode_writer.write_encoded_blocks(enc)

    def _peek_sizes(self, enc: EncodedBlocks):
        """block sizes from the stream headers of a batch (host side, vectorised): needed to size the output"""
        nbits = self._size_bits()
        host = enc.buf.cpu().numpy()
        offs = enc.bit_offset.cpu().numpy().astype(np.int64)
        lens = enc.bit_len.cpu().numpy().astype(np.int64)
        if nbits > 32:
            raise NotImplementedError("DATA_BLOCK_SIZE_BITS > 32 in the batched stream decoder")
        if (lens < nbits).any():
            raise ValueError("non-empty bitarray expected")
        first = offs >> 3
        idx = first[:, None] + np.arange(5)[None, :]
        b = host[np.minimum(idx, host.size - 1)].astype(np.uint64)
        word = (b[:, 0] << np.uint64(32)) | (b[:, 1] << np.uint64(24)) | (b[:, 2] << np.uint64(16)) | (b[:, 3] << np.uint64(8)) | b[:, 4]
        sh = (np.uint64(40) - np.uint64(nbits) - (offs & 7).astype(np.uint64))
        return ((word >> sh) & np.uint64((1 << nbits) - 1)).astype(np.int64)

    def decode(self, encode_reader, output_stream):
        if self.blocks_per_batch <= 1 or not self._blocks_independent() or not hasattr(encode_reader, "get_encoded_blocks"):
            return super().decode(encode_reader, output_stream)
        dev = self.device_coder().device
        while True:
            enc = encode_reader.get_encoded_blocks(device=dev, max_blocks=self.blocks_per_batch)
            if enc is None or enc.n_blocks == 0:
                break
            sizes = self._peek_sizes(enc)
            limit = self._max_symbols_for_bits(int(enc.bit_len.max()))
            if int(sizes.max()) > limit:
                raise ValueError("a block header declares %d symbols, more than its stream can hold (corrupt file?)" % int(sizes.max()))
            dec = self.decode_blocks(enc, max(1, int(sizes.max()))).check()
            if not bool((dec.bits_consumed == enc.bit_len).all()):
                raise AssertionError("num_bits_consumed != len(encoded_block)")  # data_encoder_decoder.py:141
            sym = dec.symbols.cpu().numpy()
            got = dec.sizes.cpu().numpy()
            for b in range(sym.shape[0]):
                output_stream.write_block(self._row_to_block(sym[b, : got[b]]))

    # ---- reference single-block API --------------------------------------------------------
    def _encode_one(self, data_block: DataBlock, model=None) -> BitArray:
        t = self._block_to_tensor(data_block)
        enc = self.device_coder().encode_blocks(t, model=model)
        enc.check()
        return enc.block(0)

    def _decode_one(self, bitarray: BitArray, model=None):
        dev = self.device_coder()
        enc = EncodedBlocks.from_bitarrays([bitarray], device=dev.device)
        # The block size comes from the stream's own header, so a corrupt stream could ask for a ~4 GiB
        # output before any validity check.  Bound it by what a stream of this length can hold
        # (`_max_symbols_for_bits`: from the cheapest symbol of a static table; adaptive models only get
        # the flat cap).
        size_cap = self._peek_size(bitarray)
        limit = self._max_symbols_for_bits(len(bitarray))
        if size_cap > limit:
            raise ValueError("block header declares %d symbols, more than a %d-bit stream can hold (corrupt stream?)" % (size_cap, len(bitarray)))
        dec = dev.decode_blocks(enc, size_cap, model=model)
        dec.check()
        n = int(dec.sizes[0])
        row = dec.symbols[0, :n].cpu().numpy()
        return self._row_to_block(row), int(dec.bits_consumed[0])

    def _peek_size(self, bitarray: BitArray) -> int:
        nbits = self._size_bits()
        head = bitarray[:nbits]
        if len(head) == 0:
            raise ValueError("non-empty bitarray expected")
        return int(head.to01(), 2)

    def _size_bits(self) -> int:
        return int(self.params.DATA_BLOCK_SIZE_BITS)

    def _max_symbols_for_bits(self, nbits: int) -> int:
        """Upper bound on the symbols a stream of `nbits` bits can decode to.  A symbol of probability p
        costs about -log2 p bits; with a static table whose most probable symbol has f_max of M that is
        log2(M / f_max) bits, which can be arbitrarily close to (or exactly) zero -- then there is no
        bound from the length and only a flat allocation cap applies."""
        f = self._freqs()
        fl = [int(x) for x in f.freq_list]
        total, fmax = sum(fl), max(fl)
        flat_cap = 1 << 28  # symbols we are willing to allocate for on the say-so of one header
        if fmax >= total:
            return flat_cap
        import math

        per_sym = math.log2(total / fmax)
        return int(min(flat_cap, 64 + (nbits + 64) / per_sym))


def encode_uint8_file(encoder, input_file_path: str, encoded_file_path: str, block_size: int = 10000, blocks_per_batch: int = 65536):
    """Batched counterpart of `DataEncoder.encode_file` for byte files: reads the input in batches of
    `blocks_per_batch` blocks, encodes each batch in one launch and writes the reference's framed
    format (scl/core/encoded_stream.py:150-175).  The output file is byte-identical to what
    `encoder.encode(Uint8FileDataStream(...), block_size, EncodedBlockWriter(...))` produces."""
    import torch

    from ..core.data_stream import Uint8FileDataStream
    from ..core.encoded_stream import EncodedBlockWriter

    with Uint8FileDataStream(input_file_path, "rb") as fds, EncodedBlockWriter(encoded_file_path) as writer:
        while True:
            got = fds.get_blocks(block_size, blocks_per_batch)
            if got is None:
                break
            data, sizes = got
            ragged = int(sizes[-1]) != block_size
            enc = encoder.encode_blocks(torch.from_numpy(data), sizes=torch.from_numpy(sizes) if ragged else None).check()
            writer.write_encoded_blocks(enc)


def decode_uint8_file(decoder, encoded_file_path: str, output_file_path: str, block_size: int = 10000, blocks_per_batch: int = 65536):
    """Batched counterpart of `DataDecoder.decode_file` for byte files written by the reference's
    `EncodedBlockWriter` (or by `encode_uint8_file`): the file is read and decoded `blocks_per_batch` records at
    a time, so device and host memory stay bounded whatever the file size."""
    from ..core.encoded_stream import EncodedBlockReader

    with EncodedBlockReader(encoded_file_path) as reader, open(output_file_path, "wb") as out:
        while True:
            enc = reader.get_encoded_blocks(device=decoder.device_coder().device, max_blocks=blocks_per_batch)
            if enc is None or enc.n_blocks == 0:
                return
            dec = decoder.decode_blocks(enc, block_size).check()
            if not bool((dec.bits_consumed == enc.bit_len).all()):
                raise AssertionError("num_bits_consumed != len(encoded_block)")  # data_encoder_decoder.py:141
            sym = dec.symbols.cpu().numpy()
            sizes = dec.sizes.cpu().numpy()
            if (sizes == sym.shape[1]).all():
                out.write(sym.tobytes())
            else:
                for b in range(sym.shape[0]):
                    out.write(sym[b, : sizes[b]].tobytes())
