"""Block framing of encoded files, byte-identical to the reference (scl/core/encoded_stream.py).

On disk every encoded block is
    [u32 big-endian payload size in bytes][3-bit pad count][pad zero bits][the block's bits]
(`Padder.add_byte_padding` :22-46, `HeaderHandler.add_header` :93-103, `EncodedBlockWriter.write_block`
:150-175).  The host classes below implement that format for single BitArrays; for batches the
same bytes are produced on the device by `EncodedBlocks.frame()` (csrc `pack_kernel<FRAMED>`),
see `write_encoded_blocks`.
"""
import numpy as np

from ..utils.bitarray_utils import BitArray, bitarray_to_uint, uint_to_bitarray


class Padder:
    NUM_PAD_BITS = 3

    @classmethod
    def add_byte_padding(cls, payload_bitarray: BitArray) -> BitArray:
        assert isinstance(payload_bitarray, BitArray)
        num_pad = (8 - (len(payload_bitarray) + cls.NUM_PAD_BITS) % 8) % 8
        return uint_to_bitarray(num_pad, bit_width=cls.NUM_PAD_BITS) + BitArray("0" * num_pad) + payload_bitarray

    @classmethod
    def remove_byte_padding(cls, payload_pad_bitarray: BitArray) -> BitArray:
        assert isinstance(payload_pad_bitarray, BitArray)
        num_pad = bitarray_to_uint(payload_pad_bitarray[: cls.NUM_PAD_BITS])
        return payload_pad_bitarray[cls.NUM_PAD_BITS + num_pad :]


class HeaderHandler:
    NUM_HEADER_BYTES = 4
    NUM_HEADER_BITS = NUM_HEADER_BYTES * 8
    MAX_PAYLOAD_SIZE = 1 << NUM_HEADER_BITS

    @classmethod
    def add_header(cls, payload_bitarray: BitArray) -> BitArray:
        assert len(payload_bitarray) % 8 == 0
        arr_size = len(payload_bitarray) // 8
        assert arr_size < cls.MAX_PAYLOAD_SIZE
        return uint_to_bitarray(arr_size, bit_width=cls.NUM_HEADER_BITS) + payload_bitarray

    @classmethod
    def get_payload_size(cls, header_bytes: bytes) -> int:
        assert isinstance(header_bytes, bytes) and len(header_bytes) == cls.NUM_HEADER_BYTES
        return int.from_bytes(header_bytes, "big")


class EncodedBlockWriter:
    def __init__(self, file_path: str):
        self.file_path = file_path

    def __enter__(self):
        self.file_writer = open(self.file_path, "wb")
        return self

    def __exit__(self, exc_type, exc_value, exc_traceback):
        self.file_writer.close()

    def write_block(self, encoded_block: BitArray):
        assert isinstance(encoded_block, BitArray)
        self.file_writer.write(HeaderHandler.add_header(Padder.add_byte_padding(encoded_block)).tobytes())

    def write_encoded_blocks(self, encoded):
        """Write a whole batch (`EncodedBlocks` from `encode_blocks`) in one go: the framing is
        done by the device kernel, the host only copies the finished bytes."""
        framed, _ = encoded.frame()
        self.file_writer.write(framed.cpu().numpy().tobytes())


class EncodedBlockReader:
    def __init__(self, file_path: str):
        self.file_path = file_path

    def __enter__(self):
        self.file_reader = open(self.file_path, "rb")
        return self

    def __exit__(self, exc_type, exc_value, exc_traceback):
        self.file_reader.close()

    def get_block(self):
        header_bytes = self.file_reader.read(HeaderHandler.NUM_HEADER_BYTES)
        if len(header_bytes) == 0:
            return None
        assert len(header_bytes) == HeaderHandler.NUM_HEADER_BYTES
        payload_size = HeaderHandler.get_payload_size(header_bytes)
        payload_bytes = self.file_reader.read(payload_size)
        assert len(payload_bytes) == payload_size
        padded = BitArray()
        padded.frombytes(payload_bytes)
        return Padder.remove_byte_padding(padded)

    def get_encoded_blocks(self, device="cuda", max_blocks: int = None):
        """Read up to `max_blocks` of the remaining blocks (default: all of them) and hand them to the device as one
        `EncodedBlocks` (bit offsets point straight into the file image: no per-block host work beyond the header
        walk).  Returns None at the end of the file."""
        import torch

        from ..device import EncodedBlocks

        if max_blocks is None:
            raw = np.frombuffer(self.file_reader.read(), dtype=np.uint8)
            if raw.size == 0:
                return None
        else:
            # walk the headers of the next `max_blocks` records without reading more of the file than they span
            start = self.file_reader.tell()
            pos, n = start, 0
            while n < max_blocks:
                self.file_reader.seek(pos)
                hb = self.file_reader.read(HeaderHandler.NUM_HEADER_BYTES)
                if len(hb) == 0:
                    break
                assert len(hb) == HeaderHandler.NUM_HEADER_BYTES, "truncated encoded file"
                pos += HeaderHandler.NUM_HEADER_BYTES + HeaderHandler.get_payload_size(hb)
                n += 1
            if n == 0:
                return None
            self.file_reader.seek(start)
            raw = np.frombuffer(self.file_reader.read(pos - start), dtype=np.uint8)
            assert raw.size == pos - start, "truncated encoded file"
        offs, lens, pos = [], [], 0
        while pos < raw.size:
            size = int.from_bytes(raw[pos : pos + 4].tobytes(), "big")
            first = int(raw[pos + 4]) if size else 0
            num_pad = first >> 5
            offs.append((pos + 4) * 8 + 3 + num_pad)
            lens.append(size * 8 - 3 - num_pad)
            pos += 4 + size
        assert pos == raw.size, "truncated encoded file"
        buf = torch.from_numpy(np.concatenate([raw, np.zeros(64, dtype=np.uint8)])).to(device)
        return EncodedBlocks(buf, torch.tensor(offs, dtype=torch.int64, device=device), torch.tensor(lens, dtype=torch.int64, device=device), None, 0)
