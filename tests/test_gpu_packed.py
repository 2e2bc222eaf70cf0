"""GPU tier (-m gpu): the contiguous-output encode (scl_encode_blocks_packed) -- the fused kernel for the
second-generation rANS / tANS path and the encode + scan + copy sequence for every other kernel family --
against the slot encoder + pack()/frame() (whose bytes the other GPU tests pin to BitArray.tobytes(), to the
reference's EncodedBlockWriter format and to the oracle), against the oracle directly, and through the decoders.
Reference behaviour matched: scl/core/encoded_stream.py:150-175 (write_block), rANS.py:199-208."""
import numpy as np
import pytest
import torch

from oracle import scl_oracle as so

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.cuda.set_device(0)
    from stanford_compression_library_b200 import _cabi

    _cabi.lib()
    yield


def _codec(name):
    from stanford_compression_library_b200 import Frequencies
    from stanford_compression_library_b200.compressors.arithmetic_coding import AECParams, ArithmeticDecoder, ArithmeticEncoder
    from stanford_compression_library_b200.compressors.probability_models import AdaptiveIIDFreqModel
    from stanford_compression_library_b200.compressors.rANS import rANSDecoder, rANSEncoder, rANSParams
    from stanford_compression_library_b200.compressors.range_coder import RangeCoderParams, RangeDecoder, RangeEncoder
    from stanford_compression_library_b200.compressors.tANS import tANSDecoder, tANSEncoder, tANSParams
    from stanford_compression_library_b200.workloads import zipf_frequencies

    fr = zipf_frequencies()
    if name == "rans_default":
        p = rANSParams(fr)
        return rANSEncoder(p), rANSDecoder(p)
    if name == "rans_nbo8":
        p = rANSParams(fr, NUM_BITS_OUT=8, RANGE_FACTOR=1 << 12)
        return rANSEncoder(p), rANSDecoder(p)
    if name == "rans_generic64":  # 64-bit state: generic kernels -> the encode + scan + copy sequence
        p = rANSParams(fr, NUM_BITS_OUT=8, RANGE_FACTOR=1 << 16)
        return rANSEncoder(p), rANSDecoder(p)
    if name == "tans":
        p = tANSParams(fr, RANGE_FACTOR=1)
        return tANSEncoder(p), tANSDecoder(p)
    if name == "range":
        p = RangeCoderParams()
        return RangeEncoder(p, fr), RangeDecoder(p, fr)
    ap = AECParams()
    uni = Frequencies({b: 1 for b in range(256)})
    return (ArithmeticEncoder(ap, AdaptiveIIDFreqModel(uni, ap.MAX_ALLOWED_TOTAL_FREQ)),
            ArithmeticDecoder(ap, AdaptiveIIDFreqModel(uni, ap.MAX_ALLOWED_TOTAL_FREQ)))


def _check_against_slots(enc, dec, data, framed):
    B, N = data.shape
    e = enc.encode_blocks(data).check()
    p = enc.encode_blocks_packed(data, framed=framed).check()
    torch.cuda.synchronize()
    assert torch.equal(p.bit_len, e.bit_len)
    if framed:
        want, woffs = e.frame()
        total = int(p.byte_offset[-1])
        assert total == want.numel() and torch.equal(p.byte_offset, woffs)
        assert torch.equal(p.buf[:total], want)
    else:
        want = e.pack()
        total = int(p.byte_offset[-1])
        assert total == e.total_bytes() and torch.equal(p.byte_offset, want.byte_offset)
        assert torch.equal(p.bit_offset, want.bit_offset)
        assert torch.equal(p.buf[:total], want.buf[:total])
    d = dec.decode_blocks(p, N).check()  # straight from the packed / framed buffer
    d0 = dec.decode_blocks(e, N).check()  # (the arithmetic decoder's bits-consumed quirk for a block of the first symbol only: compare like with like)
    assert torch.equal(d.symbols[:, :N], data) and torch.equal(d.bits_consumed, d0.bits_consumed)
    return p


@pytest.mark.parametrize("framed", [False, True], ids=["packed", "framed"])
@pytest.mark.parametrize("name", ["rans_default", "rans_nbo8", "tans", "rans_generic64", "range", "aec"])
@pytest.mark.parametrize("shape", [(1, 4096), (31, 200), (33, 64), (1000, 1030), (4737, 320)], ids=lambda s: "%dx%d" % s)
def test_packed_encode_equals_slots_plus_pack(name, shape, framed):
    from stanford_compression_library_b200.workloads import sample_blocks, zipf_probabilities

    B, N = shape
    if name == "aec":
        N = min(N, 512)
    enc, dec = _codec(name)
    data = sample_blocks(zipf_probabilities(), B, N, seed=B + N, device="cuda:0")
    data[0, :] = 255  # rarest symbol: the longest possible stream next to ordinary ones
    if B > 2:
        data[B // 2, :] = 0  # and the shortest
    _check_against_slots(enc, dec, data, framed)


def test_packed_encode_sixteen_rounds_short_blocks():
    """Many rounds of the persistent grid with very little coding per round (64-symbol blocks): the copy pool falls
    behind, so the coding warps meet the back-pressure (a round cannot be entered before the round two back is
    resolved) -- the regime of the 8 GiB single-GPU bench.  Bytes against the slot path + pack()."""
    from stanford_compression_library_b200.compressors.rANS import rANSDecoder, rANSEncoder, rANSParams
    from stanford_compression_library_b200.workloads import sample_blocks, zipf_frequencies, zipf_probabilities

    B, N = 148 * 28 * 32 * 16 + 77, 64
    params = rANSParams(zipf_frequencies())
    enc, dec = rANSEncoder(params), rANSDecoder(params)
    data = sample_blocks(zipf_probabilities(), B, N, seed=4, device="cuda:0")
    for _ in range(3):
        _check_against_slots(enc, dec, data, False)


@pytest.mark.parametrize("kw", [{}, dict(NUM_BITS_OUT=8, RANGE_FACTOR=1 << 12)], ids=["default", "nbo8_rf12"])
def test_packed_encode_many_rounds_vs_oracle(kw):
    """More tasks than one round of the persistent grid holds (the look-back then crosses rounds), a block
    count that is not a multiple of 32, reuse of the output object; bytes compared with the ORACLE's own
    concatenation for a strided sample and with the slot path for everything."""
    from stanford_compression_library_b200.compressors.rANS import rANSDecoder, rANSEncoder, rANSParams
    from stanford_compression_library_b200.workloads import sample_blocks, zipf_freq_list, zipf_frequencies, zipf_probabilities

    B, N = 148 * 28 * 32 * 2 + 32 * 57 + 13, 128
    params = rANSParams(zipf_frequencies(), **kw)
    enc, dec = rANSEncoder(params), rANSDecoder(params)
    data = sample_blocks(zipf_probabilities(), B, N, seed=3, device="cuda:0")
    p = _check_against_slots(enc, dec, data, False)
    # second call into the same object: identical bytes (workspace reset, no stale look-back state)
    snap = p.buf[: int(p.byte_offset[-1])].clone()
    p2 = enc.encode_blocks_packed(data, reuse=p).check()
    assert p2 is p and torch.equal(p.buf[: int(p.byte_offset[-1])], snap)
    oracle = so.Oracle.rans(zipf_freq_list(), **kw)
    host = data.cpu().numpy()
    offs = p.byte_offset.cpu().numpy()
    buf = p.buf.cpu().numpy()
    for b in list(range(0, B, 9973)) + [B - 1]:
        ref_bytes, ref_bits = oracle.encode_block(host[b])
        assert int(p.bit_len[b]) == ref_bits
        assert buf[offs[b] : offs[b + 1]].tobytes() == ref_bytes.tobytes()


def test_packed_encode_destination_too_small():
    """A record that would end past dst is dropped and flagged (OverflowError), earlier records are intact."""
    from stanford_compression_library_b200.compressors.rANS import rANSEncoder, rANSParams
    from stanford_compression_library_b200.workloads import sample_blocks, zipf_frequencies, zipf_probabilities

    B, N = 64, 1024
    enc = rANSEncoder(rANSParams(zipf_frequencies()))
    data = sample_blocks(zipf_probabilities(), B, N, seed=8, device="cuda:0")
    full = enc.encode_blocks_packed(data).check()
    offs = full.byte_offset.cpu().numpy()
    cap = int(offs[40]) + 5  # room for 40 records and a bit
    small = enc.encode_blocks_packed(data, capacity=cap)
    st = small.status.cpu().numpy()
    assert (st[:40] == 0).all() and (st[40:] == 3).all()
    assert torch.equal(small.buf[: int(offs[40])], full.buf[: int(offs[40])])
    with pytest.raises(OverflowError):
        small.check()


def test_packed_bad_symbol_takes_no_room():
    from stanford_compression_library_b200 import Frequencies
    from stanford_compression_library_b200.compressors.rANS import rANSEncoder, rANSParams

    fr = Frequencies({b: 16 for b in range(0, 256, 2)})  # odd byte values are not in the alphabet
    enc = rANSEncoder(rANSParams(fr))
    data = (torch.randint(0, 128, (96, 256), device="cuda:0", dtype=torch.int32) * 2).to(torch.uint8)
    data[17, 100] = 3
    p = enc.encode_blocks_packed(data)
    st = p.status.cpu().numpy()
    assert st[17] == 1 and (np.delete(st, 17) == 0).all()
    offs = p.byte_offset.cpu().numpy()
    assert offs[18] == offs[17]
    with pytest.raises(KeyError):
        p.check()
    good = enc.encode_blocks(data).pack()
    for b in (16, 18, 95):
        n = (int(p.bit_len[b]) + 7) // 8
        go = int(good.byte_offset[b])
        assert torch.equal(p.buf[offs[b] : offs[b] + n], good.buf[go : go + n])


def test_packed_offsets_scan():
    from stanford_compression_library_b200.device import EncodedBlocks

    rng = np.random.default_rng(5)
    for B in (1, 2047, 2048, 2049, 70001):
        lens = rng.integers(0, 40000, size=B).astype(np.int64)
        e = EncodedBlocks(torch.zeros(16, dtype=torch.uint8, device="cuda:0"), None, torch.from_numpy(lens).cuda(), None, 0)
        for framed in (False, True):
            byte_off, bit_off = e.packed_offsets(framed)
            sz = 4 + (lens + 3 + 7) // 8 if framed else (lens + 7) // 8
            want = np.concatenate([[0], np.cumsum(sz)])
            assert byte_off.cpu().numpy().tolist() == want.tolist()
            lead = (32 + 3 + (8 - (lens + 3) % 8) % 8) if framed else 0
            assert bit_off.cpu().numpy().tolist() == (8 * want[:-1] + lead).tolist()


@pytest.mark.parametrize("kw", [{}, dict(NUM_BITS_OUT=8, RANGE_FACTOR=1 << 12)], ids=["default", "nbo8_rf12"])
def test_rans_cfg5_shard_shape_packed_vs_oracle(kw):
    """BASELINE configs[4]'s per-GPU shard at full size, 262 144 blocks x 4 KiB, exactly as bench.py runs it: the fused
    packed encode, and the decode launch in the form the library picks by itself at this size (>= 24 tasks per SM: the
    pipe-balanced `BAL` kernel, un-forced).  A strided 64-block sample (first and last block included) is bit-compared
    with the oracle; every block round-trips with exact bit accounting; slot encode + pack() gives the same bytes.
    Reference behaviour matched: rANS.py:186-210, 270-297."""
    from stanford_compression_library_b200.compressors.rANS import rANSDecoder, rANSEncoder, rANSParams
    from stanford_compression_library_b200.workloads import sample_stream_blocks, zipf_freq_list, zipf_frequencies, zipf_probabilities

    B, N = 262144, 4096
    params = rANSParams(zipf_frequencies(), **kw)
    enc, dec = rANSEncoder(params), rANSDecoder(params)
    data = sample_stream_blocks(zipf_probabilities(), 3 * B, 4 * B, N, "cuda:0")  # the 4th shard of the 8 GiB stream
    p = enc.encode_blocks_packed(data, capacity=B * N).check()
    d = dec.decode_blocks(p, N).check()
    assert torch.equal(d.symbols[:, :N], data) and torch.equal(d.bits_consumed, p.bit_len)
    assert int(d.sizes.min()) == N == int(d.sizes.max())
    total = int(p.byte_offset[-1])
    assert total == p.total_bytes() == int(((p.bit_len + 7) // 8).sum())
    oracle = so.Oracle.rans(zipf_freq_list(), **kw)
    idx = sorted({int(round(i * (B - 1) / 63)) for i in range(64)})
    host = data[torch.tensor(idx, device="cuda:0")].cpu().numpy()
    offs = p.byte_offset.cpu().numpy()
    for j, b in enumerate(idx):
        ref_bytes, ref_bits = oracle.encode_block(host[j])
        assert int(p.bit_len[b]) == ref_bits
        assert p.buf[offs[b] : offs[b + 1]].cpu().numpy().tobytes() == ref_bytes.tobytes(), "block %d differs from the oracle" % b
    del d
    e = enc.encode_blocks(data).check()
    want = e.pack()
    assert torch.equal(want.buf[:total], p.buf[:total]) and torch.equal(want.byte_offset, p.byte_offset)


@pytest.mark.parametrize("name", ["rans_default", "tans", "range"])
def test_stream_level_encode_decode_are_batched_and_byte_identical(name, tmp_path):
    """DataEncoder.encode / DataDecoder.decode (data_encoder_decoder.py:56-69, 131-144) on a byte file that spans
    several batches on BOTH legs: the batched loops (one fused, framed launch per `blocks_per_batch` blocks) write
    exactly the bytes of the reference's block-at-a-time loop, and read them back; so do the bounded
    encode_uint8_file / decode_uint8_file helpers."""
    from stanford_compression_library_b200.compressors._gpu_base import decode_uint8_file, encode_uint8_file
    from stanford_compression_library_b200.core.data_stream import Uint8FileDataStream
    from stanford_compression_library_b200.core.encoded_stream import EncodedBlockReader, EncodedBlockWriter
    from stanford_compression_library_b200.workloads import zipf_probabilities

    rng = np.random.default_rng(17)
    raw = rng.choice(256, size=53 * 1000 + 123, p=np.array(zipf_probabilities())).astype(np.uint8).tobytes()
    src = tmp_path / "in.bin"
    src.write_bytes(raw)
    enc, dec = _codec(name)
    outs = {}
    for tag, per_batch in (("loop", 1), ("batched", 7)):
        enc.blocks_per_batch = per_batch
        path = str(tmp_path / ("enc_%s.bin" % tag))
        with Uint8FileDataStream(str(src), "rb") as fds, EncodedBlockWriter(path) as w:
            enc.encode(fds, block_size=1000, encode_writer=w)
        outs[tag] = open(path, "rb").read()
    assert outs["loop"] == outs["batched"] and len(outs["loop"]) > 0
    path2 = str(tmp_path / "enc_file_api.bin")
    encode_uint8_file(enc, str(src), path2, block_size=1000, blocks_per_batch=11)
    assert open(path2, "rb").read() == outs["loop"]
    for tag, per_batch in (("loop", 1), ("batched", 5)):
        dec.blocks_per_batch = per_batch
        back = str(tmp_path / ("dec_%s.bin" % tag))
        with EncodedBlockReader(path2) as r, Uint8FileDataStream(back, "wb") as ods:
            dec.decode(r, ods)
        assert open(back, "rb").read() == raw
    back = str(tmp_path / "dec_file_api.bin")
    decode_uint8_file(dec, path2, back, block_size=1000, blocks_per_batch=9)
    assert open(back, "rb").read() == raw


def test_stream_level_encode_text_symbols_batched(tmp_path):
    """the batched loop with a non-byte alphabet (characters): symbols are mapped to bytes per batch"""
    from stanford_compression_library_b200 import Frequencies
    from stanford_compression_library_b200.compressors.rANS import rANSDecoder, rANSEncoder, rANSParams
    from stanford_compression_library_b200.core.data_stream import TextFileDataStream
    from stanford_compression_library_b200.core.encoded_stream import EncodedBlockReader, EncodedBlockWriter

    rng = np.random.default_rng(3)
    text = "".join(rng.choice(list("abcde \n"), size=4321, p=[0.3, 0.2, 0.15, 0.1, 0.1, 0.1, 0.05]))
    src = tmp_path / "t.txt"
    src.write_text(text)
    params = rANSParams(Frequencies({"a": 30, "b": 20, "c": 15, "d": 10, "e": 10, " ": 10, "\n": 5}))
    enc, dec = rANSEncoder(params), rANSDecoder(params)
    outs = {}
    for tag, per_batch in (("loop", 1), ("batched", 4)):
        enc.blocks_per_batch = per_batch
        path = str(tmp_path / ("t_%s.bin" % tag))
        with TextFileDataStream(str(src), "r") as fds, EncodedBlockWriter(path) as w:
            enc.encode(fds, block_size=500, encode_writer=w)
        outs[tag] = open(path, "rb").read()
    assert outs["loop"] == outs["batched"]
    dec.blocks_per_batch = 3
    back = str(tmp_path / "t_back.txt")
    with EncodedBlockReader(str(tmp_path / "t_batched.bin")) as r, TextFileDataStream(back, "w") as ods:
        dec.decode(r, ods)
    assert open(back).read() == text
    with pytest.raises(KeyError):  # a character outside the alphabet: the reference's freq_dict[s] KeyError
        (tmp_path / "bad.txt").write_text("abcz")
        with TextFileDataStream(str(tmp_path / "bad.txt"), "r") as fds, EncodedBlockWriter(str(tmp_path / "bad.bin")) as w:
            enc.encode(fds, block_size=500, encode_writer=w)


@pytest.mark.parametrize("framed", [False, True], ids=["packed", "framed"])
def test_split_launches_give_the_same_streams(framed):
    """Batches of more than 2^30 blocks are coded by several launches of whole rounds (the fused encoder's
    running prefix continues across them).  With the test hook that lowers the span to 64 MiB the same batch is coded
    in one launch and in four: identical bytes, offsets and decoded symbols, both for the slot and the packed form."""
    from stanford_compression_library_b200.compressors.rANS import rANSDecoder, rANSEncoder, rANSParams
    from stanford_compression_library_b200.workloads import sample_blocks, zipf_frequencies, zipf_probabilities

    B, N = 148 * 28 * 32 * 3 + 1234, 256
    params = rANSParams(zipf_frequencies())
    enc, dec = rANSEncoder(params), rANSDecoder(params)
    data = sample_blocks(zipf_probabilities(), B, N, seed=21, device="cuda:0")
    e1 = enc.encode_blocks(data).check()
    p1 = enc.encode_blocks_packed(data, framed=framed).check()
    d1 = dec.decode_blocks(p1, N).check()
    try:
        enc.device_coder().debug_path(256)
        dec.device_coder().debug_path(256)
        e2 = enc.encode_blocks(data).check()
        p2 = enc.encode_blocks_packed(data, framed=framed).check()
        d2 = dec.decode_blocks(p2, N).check()
        d3 = dec.decode_blocks(e2, N).check()
    finally:
        enc.device_coder().debug_path(0)
        dec.device_coder().debug_path(0)
    total = int(p1.byte_offset[-1])
    assert torch.equal(e1.bit_len, e2.bit_len) and torch.equal(e1.bit_offset, e2.bit_offset)
    assert torch.equal(e1.pack().buf[: e1.total_bytes()], e2.pack().buf[: e1.total_bytes()])
    assert torch.equal(p1.byte_offset, p2.byte_offset) and torch.equal(p1.bit_offset, p2.bit_offset) and torch.equal(p1.buf[:total], p2.buf[:total])
    for d in (d1, d2, d3):
        assert torch.equal(d.symbols[:, :N], data) and torch.equal(d.bits_consumed, e1.bit_len)


@pytest.mark.parametrize("framed", [False, True], ids=["packed", "framed"])
@pytest.mark.parametrize("mode", [0, 32, 64, 128, 5 << 12, (8 << 12) | 64, (1 << 12) | 128], ids=["default", "no_ring", "piece512", "piece1024", "copy5", "copy8_piece512", "copy1_piece1024"])
@pytest.mark.parametrize("name", ["rans_default", "tans"])
def test_copy_pool_variants_give_the_same_bytes(name, mode, framed):
    """The fused encoder's copy pool has two forms of the same copy -- through registers (coding warps that ran out of
    symbols) and through a shared-memory ring filled by bulk async copies, in pieces of 512 / 1024 / 2048 bytes (the
    dedicated copy warps) -- and which one moves a given stream depends on timing.  Every variant, forced through the
    per-handle hook, must produce the bytes of encode + pack: streams shorter than one 16-byte chunk, streams of one
    piece, of several pieces, and of exactly a whole number of pieces' worth of chunks all occur here."""
    from stanford_compression_library_b200.workloads import sample_blocks, zipf_probabilities

    enc, dec = _codec(name)
    try:
        enc.device_coder().debug_path(mode)
        for (B, N), seed in (((148 * 28 * 32 + 777, 8), 1), ((20000, 150), 2), ((9000, 700), 3), ((4737 * 2, 2600), 4), ((4200, 5400), 5)):
            data = sample_blocks(zipf_probabilities(), B, N, seed=seed, device="cuda:0")
            # rows of very different compressibility: every bit / word / granule alignment of the source occurs
            data[::3] &= 0x0F
            data[1::7] = 0
            _check_against_slots(enc, dec, data, framed)
    finally:
        enc.device_coder().debug_path(0)


def test_rans_cfg5_whole_stream_on_one_gpu_packed_vs_oracle():
    """BASELINE configs[4] at N = 1: all 2 097 152 blocks x 4 KiB (8 GiB) in ONE fused launch (16 rounds of the
    persistent grid: the copy pool lags rounds behind the coder and the whole CTA finishes the rest) and one decode
    launch.  Every block round-trips with exact bit accounting; the record offsets are the running sum of the record
    sizes; a strided 48-block sample (first and last block included) is bit-compared with the oracle.
    Skipped on devices with less than 48 GB of free memory.  Reference behaviour matched: rANS.py:186-210, 270-297,
    encoded_stream.py:150-175 (concatenation order = block order)."""
    from stanford_compression_library_b200.compressors.rANS import rANSDecoder, rANSEncoder, rANSParams
    from stanford_compression_library_b200.workloads import sample_stream_blocks, zipf_freq_list, zipf_frequencies, zipf_probabilities

    free, _ = torch.cuda.mem_get_info()
    if free < 48 << 30:
        pytest.skip("needs 48 GB of free device memory")
    B, N = 2097152, 4096
    params = rANSParams(zipf_frequencies())
    enc, dec = rANSEncoder(params), rANSDecoder(params)
    data = sample_stream_blocks(zipf_probabilities(), 0, B, N, "cuda:0")
    p = enc.encode_blocks_packed(data, capacity=B * N).check()
    nbytes = (p.bit_len + 7) // 8
    total = int(p.byte_offset[-1])
    assert total == int(nbytes.sum())
    assert torch.equal(p.byte_offset[1:], torch.cumsum(nbytes, 0).to(p.byte_offset.dtype)) and int(p.byte_offset[0]) == 0
    assert torch.equal(p.bit_offset, 8 * p.byte_offset[:-1])
    d = dec.decode_blocks(p, N).check()
    assert torch.equal(d.symbols[:, :N], data) and torch.equal(d.bits_consumed, p.bit_len)
    del d
    oracle = so.Oracle.rans(zipf_freq_list())
    idx = sorted({int(round(i * (B - 1) / 47)) for i in range(48)})
    host = data[torch.tensor(idx, device="cuda:0")].cpu().numpy()
    offs = p.byte_offset.cpu().numpy()
    for j, b in enumerate(idx):
        ref_bytes, ref_bits = oracle.encode_block(host[j])
        assert int(p.bit_len[b]) == ref_bits
        assert p.buf[offs[b] : offs[b + 1]].cpu().numpy().tobytes() == ref_bytes.tobytes(), "block %d differs from the oracle" % b
    del p, data
    torch.cuda.empty_cache()


@pytest.mark.parametrize("name", ["rans_default", "rans_nbo8", "tans"])
def test_ragged_batches_take_the_fast_encoder_and_match_the_oracle(name):
    """Blocks of different sizes in one batch (`sizes`): the second-generation encoder runs as many tiles per warp as the
    warp's longest block needs and every lane stops at its own end.  Sizes 0, 1, 63, 64, 65, a full row and random ones in
    between; slot output, fused packed and fused framed output are compared with the first-generation kernels on every
    block (bytes and bit lengths) and with the oracle on a sample; the decoder returns every block at its own size."""
    from stanford_compression_library_b200.workloads import sample_blocks, zipf_freq_list, zipf_probabilities

    enc, dec = _codec(name)
    B, N = 148 * 28 * 32 + 1000, 448
    data = sample_blocks(zipf_probabilities(), B, N, seed=33, device="cuda:0")
    g = torch.Generator(device="cuda:0")
    g.manual_seed(5)
    sizes = torch.randint(0, N + 1, (B,), generator=g, device="cuda:0", dtype=torch.int32)
    sizes[:6] = torch.tensor([0, 1, 63, 64, 65, N], dtype=torch.int32, device="cuda:0")
    sizes[-1] = 0
    sizes[5000:5032] = N  # one warp of full rows (the full-tile path) among ragged ones
    e = enc.encode_blocks(data, sizes=sizes).check()
    try:
        enc.device_coder().debug_path(1)  # first-generation kernels
        e1 = enc.encode_blocks(data, sizes=sizes).check()
    finally:
        enc.device_coder().debug_path(0)
    assert torch.equal(e.bit_len, e1.bit_len)
    want = e1.pack()
    got = e.pack()
    total = e1.total_bytes()
    assert torch.equal(got.buf[:total], want.buf[:total]) and torch.equal(got.byte_offset, want.byte_offset)
    for framed in (False, True):
        p = enc.encode_blocks_packed(data, sizes=sizes, framed=framed).check()
        assert torch.equal(p.bit_len, e1.bit_len)
        if framed:
            fw, foffs = e1.frame()
            assert torch.equal(p.byte_offset, foffs) and torch.equal(p.buf[: fw.numel()], fw)
        else:
            assert torch.equal(p.byte_offset, want.byte_offset) and torch.equal(p.buf[:total], want.buf[:total])
        d = dec.decode_blocks(p, N).check()
        assert torch.equal(d.sizes, sizes)
        mask = torch.arange(N, device="cuda:0")[None, :] < sizes[:, None]
        assert torch.equal(d.symbols[:, :N][mask], data[mask])
    kind = "tans" if name == "tans" else "rans"
    kw = dict(NUM_BITS_OUT=8, RANGE_FACTOR=1 << 12) if name == "rans_nbo8" else (dict(RANGE_FACTOR=1) if name == "tans" else {})
    oracle = getattr(so.Oracle, kind)(zipf_freq_list(), **kw)
    offs = want.byte_offset.cpu().numpy()
    host, hs = data.cpu().numpy(), sizes.cpu().numpy()
    for b in list(range(8)) + [5000, 5031, B - 2, B - 1] + list(range(100, B, B // 23)):
        ref_bytes, ref_bits = oracle.encode_block(host[b, : hs[b]])
        assert int(e.bit_len[b]) == ref_bits
        assert got.buf[offs[b] : offs[b + 1]].cpu().numpy().tobytes() == ref_bytes.tobytes(), "block %d (size %d) differs from the oracle" % (b, hs[b])


def test_ragged_size_beyond_the_row_is_reported_not_read():
    """A `sizes` entry larger than the row (a caller error) must not make the fast encoder read past the row: the block
    codes nothing and carries the overflow status; its neighbours are untouched."""
    from stanford_compression_library_b200.workloads import sample_blocks, zipf_probabilities

    enc, dec = _codec("rans_default")
    B, N = 4000, 256
    data = sample_blocks(zipf_probabilities(), B, N, seed=9, device="cuda:0")
    sizes = torch.full((B,), N, dtype=torch.int32, device="cuda:0")
    good = enc.encode_blocks(data, sizes=sizes).check()
    sizes[1234] = N + 1
    e = enc.encode_blocks(data, sizes=sizes)
    st = e.status.cpu().numpy()
    assert st[1234] != 0 and (np.delete(st, 1234) == 0).all()
    keep = torch.ones(B, dtype=torch.bool, device="cuda:0")
    keep[1234] = False
    assert torch.equal(e.bit_len[keep], good.bit_len[keep])
    with pytest.raises(Exception):
        e.check()
