"""Data streams with the reference's interface (scl/core/data_stream.py:10-258) -- bulk versions.

The reference builds every block symbol by symbol (`get_block` calls `get_symbol` in a Python
loop, data_stream.py:48-63, a self-described TODO).  Same classes and semantics here, but
`get_block` / `write_block` move whole blocks at once, and `Uint8FileDataStream.get_blocks`
returns MANY blocks as one `uint8` array -- the shape the batched GPU coders take.
"""
import abc

import numpy as np

from .data_block import DataBlock


class DataStream(abc.ABC):
    @abc.abstractmethod
    def seek(self, pos: int):
        pass

    @abc.abstractmethod
    def get_symbol(self):
        """next symbol, or None at the end of the stream"""

    def get_block(self, block_size: int):
        """a DataBlock of up to `block_size` symbols, or None when the stream is exhausted"""
        data = []
        for _ in range(block_size):
            s = self.get_symbol()
            if s is None:
                break
            data.append(s)
        return DataBlock(data) if data else None

    def get_block_batch(self, block_size: int, max_blocks: int):
        """Up to `max_blocks` consecutive blocks as a list of DataBlocks ([] when the stream is exhausted): what the
        batched `DataEncoder.encode` pulls per launch.  Streams override it with a bulk read where they can."""
        out = []
        while len(out) < max_blocks:
            b = self.get_block(block_size)
            if b is None:
                break
            out.append(b)
        return out

    @abc.abstractmethod
    def write_symbol(self, s):
        pass

    def write_block(self, data_block: DataBlock):
        for s in data_block.data_list:
            self.write_symbol(s)

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc_value, exc_traceback):
        pass


class ListDataStream(DataStream):
    """wrapper around a list of symbols (data_stream.py:92-160).

    Deliberate deviation: the reference's `write_symbol` (data_stream.py:151-160) never advances
    `current_ind`, so writing [1,2,3] to an empty ListDataStream leaves [3]; here writes advance
    the position (the evident intent, and what `DataDecoder.decode` needs from an output stream)."""

    def __init__(self, input_list):
        assert isinstance(input_list, list)
        self.input_list = input_list
        self.current_ind = 0

    def seek(self, pos: int):
        assert pos <= len(self.input_list)
        self.current_ind = pos

    def get_symbol(self):
        if self.current_ind >= len(self.input_list):
            return None
        s = self.input_list[self.current_ind]
        self.current_ind += 1
        return s

    def get_block(self, block_size: int):  # bulk slice instead of a per-symbol loop
        data = self.input_list[self.current_ind : self.current_ind + block_size]
        self.current_ind += len(data)
        return DataBlock(data) if data else None

    def get_block_batch(self, block_size: int, max_blocks: int):
        data = self.input_list[self.current_ind : self.current_ind + block_size * max_blocks]
        self.current_ind += len(data)
        return [DataBlock(data[i : i + block_size]) for i in range(0, len(data), block_size)]

    def write_symbol(self, s):
        assert self.current_ind <= len(self.input_list)
        if self.current_ind < len(self.input_list):
            self.input_list[self.current_ind] = s
        else:
            self.input_list.append(s)
        self.current_ind += 1

    def write_block(self, data_block: DataBlock):
        data = list(data_block.data_list)
        self.input_list[self.current_ind : self.current_ind + len(data)] = data
        self.current_ind += len(data)


class FileDataStream(DataStream):
    """file-backed stream; opens on __enter__, closes on __exit__ (data_stream.py:140-186)"""

    def __init__(self, file_path: str, permissions="r"):
        self.file_path = file_path
        self.permissions = permissions

    def __enter__(self):
        self.file_obj = open(self.file_path, self.permissions)
        return self

    def __exit__(self, exc_type, exc_value, exc_traceback):
        self.file_obj.close()

    def seek(self, pos: int):
        self.file_obj.seek(pos)


class TextFileDataStream(FileDataStream):
    """characters of a text file (data_stream.py:189-211)"""

    def get_symbol(self):
        s = self.file_obj.read(1)
        return s if s else None

    def get_block(self, block_size: int):
        s = self.file_obj.read(block_size)
        return DataBlock(list(s)) if s else None

    def get_block_batch(self, block_size: int, max_blocks: int):
        s = self.file_obj.read(block_size * max_blocks)
        return [DataBlock(list(s[i : i + block_size])) for i in range(0, len(s), block_size)]

    def write_symbol(self, s):
        self.file_obj.write(s)

    def write_block(self, data_block: DataBlock):
        self.file_obj.write("".join(data_block.data_list))


class Uint8FileDataStream(FileDataStream):
    """bytes of a binary file as ints 0..255 (data_stream.py:214-235); open with "rb" / "wb" """

    def get_symbol(self):
        s = self.file_obj.read(1)
        if not s:
            return None
        return s[0]

    def get_block(self, block_size: int):
        s = self.file_obj.read(block_size)
        return DataBlock(list(s)) if s else None

    def get_block_batch(self, block_size: int, max_blocks: int):
        got = self.get_blocks(block_size, max_blocks)
        if got is None:
            return []
        data, sizes = got
        return [DataBlock(data[b, : sizes[b]]) for b in range(data.shape[0])]  # numpy rows: no per-symbol Python objects

    def get_blocks(self, block_size: int, max_blocks: int):
        """Up to `max_blocks` blocks at once: (uint8 [n, block_size] zero-padded, sizes int32 [n]) or None."""
        raw = self.file_obj.read(block_size * max_blocks)
        if not raw:
            return None
        a = np.frombuffer(raw, dtype=np.uint8)
        n = (a.size + block_size - 1) // block_size
        out = np.zeros((n, block_size), dtype=np.uint8)
        out.reshape(-1)[: a.size] = a
        sizes = np.full(n, block_size, dtype=np.int32)
        sizes[-1] = a.size - (n - 1) * block_size
        return out, sizes

    def write_symbol(self, s):
        assert 0 <= s <= 255
        self.file_obj.write(bytes([s]))

    def write_block(self, data_block: DataBlock):
        d = data_block.data_list
        self.file_obj.write(bytes(d.tolist() if hasattr(d, "tolist") else d))
