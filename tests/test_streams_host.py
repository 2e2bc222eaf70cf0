"""CPU tier: host stream / framing classes against the reference's semantics and byte format
(scl/core/data_stream.py, scl/core/encoded_stream.py)."""
import os

import numpy as np
import pytest

from oracle.ref_loader import reference_available
from stanford_compression_library_b200.core.data_block import DataBlock
from stanford_compression_library_b200.core.data_stream import ListDataStream, TextFileDataStream, Uint8FileDataStream
from stanford_compression_library_b200.core.encoded_stream import EncodedBlockReader, EncodedBlockWriter, HeaderHandler, Padder
from stanford_compression_library_b200.utils.bitarray_utils import BitArray


def test_list_data_stream_blocks_and_writes():
    # mirrors data_stream.py:264-290
    ds = ListDataStream(list(range(10)))
    for i in range(3):
        assert ds.get_block(3).data_list == [3 * i, 3 * i + 1, 3 * i + 2]
    assert ds.get_block(3).data_list == [9]
    assert ds.get_block(3) is None
    ds.seek(0)
    ds.write_block(DataBlock([7, 7]))
    assert ds.input_list[:3] == [7, 7, 2]
    out = ListDataStream([])
    out.write_block(DataBlock([1, 2, 3]))
    out.write_symbol(4)
    assert out.input_list == [1, 2, 3, 4]


def test_file_streams_roundtrip(tmp_path):
    p = tmp_path / "t.txt"
    with TextFileDataStream(str(p), "w") as f:
        f.write_block(DataBlock(list("hello world")))
    with TextFileDataStream(str(p), "r") as f:
        assert f.get_block(4).data_list == list("hell")
        assert f.get_symbol() == "o"
        f.seek(0)
        assert "".join(f.get_block(100).data_list) == "hello world"
        assert f.get_block(1) is None
    q = tmp_path / "t.bin"
    with Uint8FileDataStream(str(q), "wb") as f:
        f.write_block(DataBlock([0, 255, 7]))
        f.write_symbol(9)
    with Uint8FileDataStream(str(q), "rb") as f:
        assert f.get_block(3).data_list == [0, 255, 7] and f.get_symbol() == 9 and f.get_symbol() is None
    with Uint8FileDataStream(str(q), "rb") as f:
        data, sizes = f.get_blocks(3, 10)
        assert data.tolist() == [[0, 255, 7], [9, 0, 0]] and sizes.tolist() == [3, 1]
        assert f.get_blocks(3, 10) is None


def test_padder_and_header_format():
    # encoded_stream.py:61-75, 115-131 + the exact bit layout
    for payload in (BitArray("10110"), BitArray("1" * 23), BitArray(""), BitArray("1" * 13)):
        padded = Padder.add_byte_padding(payload)
        assert len(padded) % 8 == 0 and Padder.remove_byte_padding(padded) == payload
    assert Padder.add_byte_padding(BitArray("10110")).to01() == "000" + "10110"
    assert Padder.add_byte_padding(BitArray("1" * 16)).to01() == "101" + "00000" + "1" * 16  # byte-aligned payload -> 0xA0 prefix
    framed = HeaderHandler.add_header(Padder.add_byte_padding(BitArray("1" * 23)))
    assert framed.tobytes()[:4] == (4).to_bytes(4, "big") and HeaderHandler.get_payload_size(framed.tobytes()[:4]) == 4


def test_block_writer_reader_roundtrip(tmp_path):
    blocks = [BitArray("1" * 7), BitArray("0101" * 5), BitArray(""), BitArray("1")]
    p = str(tmp_path / "e.bin")
    with EncodedBlockWriter(p) as w:
        for b in blocks:
            w.write_block(b)
    with EncodedBlockReader(p) as r:
        got = []
        while True:
            b = r.get_block()
            if b is None:
                break
            got.append(b)
    assert got == blocks


@pytest.mark.skipif(not reference_available(), reason="/root/reference not present")
def test_framing_bytes_equal_reference(tmp_path):
    from oracle.ref_loader import import_reference

    import_reference()
    from scl.core.encoded_stream import EncodedBlockWriter as RefWriter
    from scl.utils.bitarray_utils import BitArray as RefBitArray

    rng = np.random.default_rng(0)
    strings = ["".join("1" if b else "0" for b in rng.integers(0, 2, size=n)) for n in (0, 1, 4, 5, 8, 13, 61, 64, 1000)]
    a, b = str(tmp_path / "ours.bin"), str(tmp_path / "ref.bin")
    with EncodedBlockWriter(a) as w:
        for s in strings:
            w.write_block(BitArray(s))
    with RefWriter(b) as w:
        for s in strings:
            w.write_block(RefBitArray(s))
    assert open(a, "rb").read() == open(b, "rb").read()
