"""DataEncoder / DataDecoder: the reference's codec interface, kept verbatim as the drop-in
boundary (scl/core/data_encoder_decoder.py:15-160).

Subclasses implement `encode_block(DataBlock) -> BitArray` and
`decode_block(BitArray) -> (DataBlock, num_bits_consumed)`; the stream-level loops are the
reference's.  The GPU coders additionally expose `encode_blocks` / `decode_blocks`, the batched
form the hardware wants (thousands of independent blocks per launch).
"""
import abc

from ..utils.bitarray_utils import BitArray
from .data_block import DataBlock


class DataEncoder(abc.ABC):
    def reset(self):
        """clear state persisted across encode_block calls (no-op by default, as in the reference)"""

    def encode_block(self, data_block: DataBlock) -> BitArray:
        raise NotImplementedError

    def encode(self, data_stream, block_size: int, encode_writer):
        # data_encoder_decoder.py:56-69
        while True:
            data_block = data_stream.get_block(block_size)
            if data_block is None:
                break
            output = self.encode_block(data_block)
            assert isinstance(output, BitArray)
            encode_writer.write_block(output)

    def encode_file(self, input_file_path: str, encoded_file_path: str, block_size: int = 10000):
        from .data_stream import TextFileDataStream
        from .encoded_stream import EncodedBlockWriter

        with TextFileDataStream(input_file_path, "r") as fds:
            with EncodedBlockWriter(encoded_file_path) as writer:
                self.encode(fds, block_size=block_size, encode_writer=writer)


class DataDecoder(abc.ABC):
    def reset(self):
        """clear state persisted across decode_block calls (no-op by default, as in the reference)"""

    def decode_block(self, bitarray: BitArray):
        raise NotImplementedError

    def decode(self, encode_reader, output_stream):
        # data_encoder_decoder.py:131-144
        while True:
            encoded_block = encode_reader.get_block()
            if encoded_block is None:
                break
            output_block, num_bits_consumed = self.decode_block(encoded_block)
            assert num_bits_consumed == len(encoded_block)
            output_stream.write_block(output_block)

    def decode_file(self, encoded_file_path: str, output_file_path: str):
        from .data_stream import TextFileDataStream
        from .encoded_stream import EncodedBlockReader

        with EncodedBlockReader(encoded_file_path) as reader:
            with TextFileDataStream(output_file_path, "w") as fds:
                self.decode(reader, fds)
