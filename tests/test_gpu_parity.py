"""GPU tier (-m gpu): the CUDA kernels, called through the C-ABI / the drop-in classes, against
 (a) the golden vectors generated from the unmodified reference,
 (b) the C oracle on seeded random batches, and
 (c) size-independent properties at the BASELINE.json sizes (round trip, exact bit accounting).
Bit-exact everywhere: this path is pure integer arithmetic."""
import copy

import numpy as np
import pytest
import torch

from oracle import scl_oracle as so
from tests.golden_util import case_id, expected_final_model, load_golden, with_garbage

pytestmark = pytest.mark.gpu

CASES = load_golden()


@pytest.fixture(scope="module", autouse=True)
def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.cuda.set_device(0)
    from stanford_compression_library_b200 import _cabi

    _cabi.lib()  # must be the built CUDA library, loudly
    yield


def _force(mode, *coders):
    """kernel-selection test hook, per handle (scl_coder_debug_path): 1 = first-generation kernels, 2 = v2 decode with
    sector stores, 3 / 4 = v2 decode always / never pipe-balanced, 0 = default"""
    for c in coders:
        c.device_coder().debug_path(mode)


def _F(freqs):
    from stanford_compression_library_b200 import Frequencies

    return Frequencies({i: int(f) for i, f in enumerate(freqs)})


def make_codec(c):
    """(encoder, decoder) drop-in objects for a golden case."""
    from stanford_compression_library_b200.compressors.arithmetic_coding import AECParams, ArithmeticDecoder, ArithmeticEncoder
    from stanford_compression_library_b200.compressors.probability_models import AdaptiveIIDFreqModel, AdaptiveOrderKFreqModel, FixedFreqModel
    from stanford_compression_library_b200.compressors.rANS import rANSDecoder, rANSEncoder, rANSParams
    from stanford_compression_library_b200.compressors.range_coder import RangeCoderParams, RangeDecoder, RangeEncoder
    from stanford_compression_library_b200.compressors.tANS import tANSDecoder, tANSEncoder, tANSParams

    p = c["params"]
    if c["coder"] in ("rans", "tans"):
        P, E, D = (tANSParams, tANSEncoder, tANSDecoder) if c["coder"] == "tans" else (rANSParams, rANSEncoder, rANSDecoder)
        params = P(_F(c["freqs"]), DATA_BLOCK_SIZE_BITS=p["DATA_BLOCK_SIZE_BITS"], NUM_BITS_OUT=p["NUM_BITS_OUT"], RANGE_FACTOR=p["RANGE_FACTOR"])
        assert params.NUM_STATE_BITS == p["NUM_STATE_BITS"]
        return E(params), D(params)
    if c["coder"] == "range":
        params = RangeCoderParams(**p)
        return RangeEncoder(params, _F(c["freqs"])), RangeDecoder(params, _F(c["freqs"]))
    params = AECParams(**p)
    if c["model"]["kind"] == "order_k":
        m = AdaptiveOrderKFreqModel(list(range(len(c["freqs"]))), c["model"]["k"], c["model"]["max_total"])
        return ArithmeticEncoder(params, m), ArithmeticDecoder(params, copy.deepcopy(m))
    cls = AdaptiveIIDFreqModel if c["model"]["kind"] == "adaptive_iid" else FixedFreqModel
    m = cls(_F(c["freqs"]), c["model"]["max_total"])
    return ArithmeticEncoder(params, m), ArithmeticDecoder(params, copy.deepcopy(m))


@pytest.mark.parametrize("c", CASES, ids=case_id)
def test_dropin_classes_match_golden(c):
    from stanford_compression_library_b200 import BitArray, DataBlock

    enc, dec = make_codec(c)
    ba = enc.encode_block(DataBlock(c["data"].tolist()))
    assert isinstance(ba, BitArray)
    assert len(ba) == c["nbits"]
    assert ba.tobytes() == c["enc"].tobytes()
    if c["coder"] == "aec":
        assert enc.freq_model._to_table() == expected_final_model(c)
    packed, total = with_garbage(c["enc"], c["nbits"], c["garbage"])
    block, used = dec.decode_block(BitArray.from_packed(packed, total))
    assert list(block.data_list) == c["data"].tolist()
    assert used == c["consumed"] == c["nbits"]
    if c["coder"] == "aec":
        assert dec.freq_model._to_table() == expected_final_model(c)


def test_kat_string_symbols():
    # rANS.py:303-360 / tANS.py:340-415 with the reference's own symbols "A","B","C"
    from stanford_compression_library_b200 import BitArray, DataBlock, Frequencies
    from stanford_compression_library_b200.compressors.rANS import rANSDecoder, rANSEncoder, rANSParams
    from stanford_compression_library_b200.compressors.tANS import tANSDecoder, tANSEncoder, tANSParams

    freq = Frequencies({"A": 3, "B": 3, "C": 2})
    data = DataBlock(["A", "C", "B"])
    expected = BitArray("00011" + "1011" + "10" + "01" + "0")
    for P, E, D in ((rANSParams, rANSEncoder, rANSDecoder), (tANSParams, tANSEncoder, tANSDecoder)):
        params = P(freq, DATA_BLOCK_SIZE_BITS=5, NUM_BITS_OUT=1, RANGE_FACTOR=1)
        assert params.INITIAL_STATE == 8 and params.NUM_STATE_BITS == 4
        got = E(params).encode_block(data)
        assert got == expected
        block, used = D(params).decode_block(got + BitArray("1101"))
        assert block.data_list == ["A", "C", "B"] and used == 14
    # tANS lookup tables built on the device (tANS.py:285-337)
    enc = tANSEncoder(tANSParams(freq, DATA_BLOCK_SIZE_BITS=5, NUM_BITS_OUT=1, RANGE_FACTOR=1))
    assert enc.base_encode_step_table == {("A", 3): 8, ("A", 4): 9, ("A", 5): 10, ("B", 3): 11, ("B", 4): 12, ("B", 5): 13, ("C", 2): 14, ("C", 3): 15}
    assert enc.shrink_state_num_out_bits_base_table == {"A": 1, "B": 1, "C": 2}
    assert enc.shrink_state_thresh_table == {"A": 12, "B": 12, "C": 16}
    dec = tANSDecoder(tANSParams(freq, DATA_BLOCK_SIZE_BITS=5, NUM_BITS_OUT=1, RANGE_FACTOR=1))
    assert dec.base_decode_step_table == {8: ("A", 3), 9: ("A", 4), 10: ("A", 5), 11: ("B", 3), 12: ("B", 4), 13: ("B", 5), 14: ("C", 2), 15: ("C", 3)}
    assert dec.expand_state_num_bits_table == {2: 2, 3: 2, 4: 1, 5: 1}
    with pytest.raises(KeyError):
        rANSEncoder(rANSParams(freq)).encode_block(DataBlock(["A", "Z"]))


def _seeded_sizes(lo, hi, n, seed=7):
    return torch.from_numpy(np.random.default_rng(seed).integers(lo, hi, size=n).astype(np.int32)).cuda()


def _compare_batch_with_oracle(coder_enc, coder_dec, oracle, data, sizes=None, sample=None, consumed_equals_length=True):
    """encode on the GPU, compare every (sampled) block's bits with the oracle, decode, compare."""
    B, N = data.shape
    e = coder_enc.encode_blocks(data, sizes=sizes).check()
    host = data.cpu().numpy()
    hs = None if sizes is None else sizes.cpu().numpy().astype(np.uint32)
    idx = range(B) if sample is None else sample
    sub = np.ascontiguousarray(host[list(idx)])
    subsz = None if hs is None else np.ascontiguousarray(hs[list(idx)])
    ref_out, ref_bits, ref_st = oracle.encode_batch(sub, sizes=subsz, out_stride=e.out_stride + 64)
    assert (ref_st == 0).all()
    off = e.bit_offset.cpu().numpy()
    ln = e.bit_len.cpu().numpy()
    buf = e.buf.cpu().numpy()
    for j, b in enumerate(idx):
        assert int(ln[b]) == int(ref_bits[j]), "block %d: bit length %d != oracle %d" % (b, ln[b], ref_bits[j])
        first, last = int(off[b]) >> 3, (int(off[b]) + int(ln[b]) + 7) >> 3
        bits = np.unpackbits(buf[first:last])[int(off[b]) - 8 * first :][: int(ln[b])]
        assert np.packbits(bits).tobytes() == ref_out[j, : (int(ln[b]) + 7) // 8].tobytes(), "block %d: bits differ from the oracle" % b
    d = coder_dec.decode_blocks(e, N).check()
    want_sizes = torch.full((B,), N, dtype=torch.int32, device=data.device) if sizes is None else sizes.to(torch.int32)
    assert torch.equal(d.sizes, want_sizes)
    if consumed_equals_length:
        assert torch.equal(d.bits_consumed, e.bit_len)
    else:  # arithmetic coder: the reference's own count can be len - 1 (see test_aec_kernel_generations_agree)
        offs = np.arange(len(ref_bits), dtype=np.uint64) * np.uint64((e.out_stride + 64) * 8)
        _, _, ref_used, ref_st2 = oracle.decode_batch(ref_out, offs, ref_bits, N)
        assert (ref_st2 == 0).all()
        assert np.array_equal(ref_used.astype(np.int64), d.bits_consumed.cpu().numpy()[list(idx)])
    if sizes is None:
        assert torch.equal(d.symbols[:, :N], data)
    else:
        mask = torch.arange(N, device=data.device)[None, :] < sizes[:, None]
        assert torch.equal(d.symbols[:, :N][mask], data[mask])
    return e, d


@pytest.mark.parametrize("kw", [{}, dict(NUM_BITS_OUT=8, RANGE_FACTOR=1 << 12), dict(NUM_BITS_OUT=2, RANGE_FACTOR=1 << 10), dict(NUM_BITS_OUT=16, RANGE_FACTOR=1 << 20)],
                         ids=["default", "nbo8_rf12", "nbo2_rf10", "nbo16_rf20_generic64"])
def test_rans_batch_vs_oracle_zipf(kw):
    from stanford_compression_library_b200.compressors.rANS import rANSDecoder, rANSEncoder, rANSParams
    from stanford_compression_library_b200.workloads import sample_blocks, zipf_freq_list, zipf_frequencies

    fl = zipf_freq_list()
    params = rANSParams(zipf_frequencies(), **kw)
    B, N = (1024, 4096) if "RANGE_FACTOR" not in kw or kw["RANGE_FACTOR"] != 1 << 20 else (256, 1024)
    data = sample_blocks(fl, B, N, seed=1, device="cuda:0")
    sizes = torch.from_numpy(np.random.default_rng(1).integers(0, N + 1, size=B).astype(np.int32)).cuda()
    oracle = so.Oracle.rans(fl, **kw)
    enc, dec = rANSEncoder(params), rANSDecoder(params)
    _compare_batch_with_oracle(enc, dec, oracle, data)
    _compare_batch_with_oracle(enc, dec, oracle, data, sizes=sizes)


def test_rans_non_power_of_two_total_and_uniform_cfg1():
    # BASELINE cfg1: 4 KiB uniform bytes; table {b:16} and the counts+1 table (M = 4352)
    from stanford_compression_library_b200.compressors.rANS import rANSDecoder, rANSEncoder, rANSParams

    rng = np.random.default_rng(0)
    u = rng.integers(0, 256, size=(64, 4096)).astype(np.uint8)
    data = torch.from_numpy(u).cuda()
    for fl in ([16] * 256, (np.bincount(u[0], minlength=256) + 1).tolist()):
        for kw in ({}, dict(NUM_BITS_OUT=8, RANGE_FACTOR=1 << 12)):
            params = rANSParams(_F(fl), **kw)
            _compare_batch_with_oracle(rANSEncoder(params), rANSDecoder(params), so.Oracle.rans(fl, **kw), data)


def test_tans_batch_vs_oracle_and_equals_rans():
    from stanford_compression_library_b200.compressors.rANS import rANSEncoder, rANSParams
    from stanford_compression_library_b200.compressors.tANS import tANSDecoder, tANSEncoder, tANSParams
    from stanford_compression_library_b200.workloads import sample_blocks, zipf_freq_list, zipf_frequencies

    fl = zipf_freq_list()
    data = sample_blocks(fl, 512, 4096, seed=2, device="cuda:0")
    for rf in (1, 4, 1 << 8):  # 4096-, 16384- (shared memory) and 2^20-state (global memory) tables
        params = tANSParams(zipf_frequencies(), RANGE_FACTOR=rf)
        e, _ = _compare_batch_with_oracle(tANSEncoder(params), tANSDecoder(params), so.Oracle.tans(fl, RANGE_FACTOR=rf), data, sample=range(0, 512, 7))
        r = rANSEncoder(rANSParams(zipf_frequencies(), RANGE_FACTOR=rf)).encode_blocks(data).check()
        assert torch.equal(r.bit_len, e.bit_len)
        assert torch.equal(r.pack().buf, e.pack().buf)  # identical bitstreams


def test_range_batch_vs_oracle():
    from stanford_compression_library_b200.compressors.range_coder import RangeCoderParams, RangeDecoder, RangeEncoder
    from stanford_compression_library_b200.workloads import sample_blocks, zipf_freq_list, zipf_frequencies

    fl = zipf_freq_list()
    data = sample_blocks(fl, 512, 4096, seed=3, device="cuda:0")
    sizes = _seeded_sizes(0, 4097, 512)
    params = RangeCoderParams()
    enc, dec = RangeEncoder(params, zipf_frequencies()), RangeDecoder(params, zipf_frequencies())
    _compare_batch_with_oracle(enc, dec, so.Oracle.range_coder(fl), data, sizes=sizes, sample=range(0, 512, 5))
    # extreme skew edge cases of range_coder.py:351-367
    skew = [1, 1, 65534]
    d2 = torch.from_numpy(np.stack([np.tile([0, 1, 2], 200), np.zeros(600, dtype=np.int64), np.full(600, 2)]).astype(np.uint8)).cuda()
    enc, dec = RangeEncoder(params, _F(skew)), RangeDecoder(params, _F(skew))
    _compare_batch_with_oracle(enc, dec, so.Oracle.range_coder(skew), d2)


@pytest.mark.parametrize("name,shape", [("zipf", (1000, 4100)), ("zipf", (77, 333)), ("two", (300, 1000)), ("skew", (200, 2048)), ("T16", (130, 777)),
                                        ("flat", (64, 4096))])
def test_range_kernel_generations_agree(name, shape):
    """Second-generation range-coder kernels (TMA tiles, sector rings, voted normalisation) against the
    first-generation ones and the oracle: identical streams, each decodes the other's output, for block
    counts that are not a multiple of 32 (padding lanes vote too) and lengths off the 64-byte tile."""
    from stanford_compression_library_b200 import _cabi
    from stanford_compression_library_b200.compressors.range_coder import RangeCoderParams, RangeDecoder, RangeEncoder
    from stanford_compression_library_b200.workloads import zipf_freq_list

    tables = {"zipf": zipf_freq_list(), "skew": [4096 - 255] + [1] * 255, "two": [4095, 1], "T16": [5, 3, 7, 1], "flat": [16] * 256}
    fl = tables[name]
    B, N = shape
    rng = np.random.default_rng(31)
    p = np.asarray(fl, dtype=np.float64) / sum(fl)
    host = rng.choice(len(fl), size=(B, N), p=p).astype(np.uint8)
    host[0, :] = len(fl) - 1                      # rarest symbol everywhere
    host[1, :] = rng.integers(0, len(fl), N)      # far from the table's distribution: many multi-byte releases
    data = torch.from_numpy(host).cuda()
    params = RangeCoderParams()
    enc, dec = RangeEncoder(params, _F(fl)), RangeDecoder(params, _F(fl))
    try:
        _force(1, enc, dec)
        e1 = enc.encode_blocks(data).check()
        p1 = e1.pack(bytewise=True)
        _force(0, enc, dec)
        e2 = enc.encode_blocks(data).check()
        p2 = e2.pack()
        assert torch.equal(e1.bit_len, e2.bit_len) and torch.equal(e1.bit_offset, e2.bit_offset)
        assert torch.equal(p1.buf, p2.buf)
        d2 = dec.decode_blocks(e1, N).check()      # v2 decoder on v1 output (slots)
        d3 = dec.decode_blocks(p2, N).check()      # v2 decoder on packed streams (byte-granular offsets)
        _force(1, enc, dec)
        d1 = dec.decode_blocks(e2, N).check()      # v1 decoder on v2 output
    finally:
        _force(0, enc, dec)
    for d in (d1, d2, d3):
        assert torch.equal(d.symbols[:, :N], data) and torch.equal(d.bits_consumed, e1.bit_len)
        assert int(d.sizes.min()) == N == int(d.sizes.max())
    oracle = so.Oracle.range_coder(fl)
    for b in (0, 1, 2, B - 1):
        ref_bytes, ref_bits = oracle.encode_block(host[b])
        got = e2.block(b)
        assert len(got) == ref_bits and got.tobytes() == ref_bytes.tobytes()


def test_range_v2_decoder_ragged_sizes_and_alphabet_check():
    """v2 decoder on a batch whose blocks differ in size (per-lane control flow instead of the vote), and the
    v2 encoder's bad-symbol status when a byte value is not in the alphabet."""
    from stanford_compression_library_b200.compressors.range_coder import RangeCoderParams, RangeDecoder, RangeEncoder

    fl = [5, 3, 7, 1]
    rng = np.random.default_rng(32)
    B, N = 100, 500
    host = rng.integers(0, 4, size=(B, N)).astype(np.uint8)
    sizes = rng.integers(0, N + 1, size=B).astype(np.int32)
    params = RangeCoderParams()
    enc, dec = RangeEncoder(params, _F(fl)), RangeDecoder(params, _F(fl))
    e = enc.encode_blocks(torch.from_numpy(host).cuda(), sizes=torch.from_numpy(sizes).cuda()).check()  # ragged: first-generation encoder
    d = dec.decode_blocks(e, N).check()
    assert torch.equal(d.sizes.cpu(), torch.from_numpy(sizes))
    out = d.symbols.cpu().numpy()
    for b in range(B):
        assert (out[b, : sizes[b]] == host[b, : sizes[b]]).all()
    host[3, 17] = 9  # not in the alphabet
    e = enc.encode_blocks(torch.from_numpy(host).cuda())
    with pytest.raises(KeyError):
        e.check()
    assert int((e.status != 0).sum()) == 1 and int(e.status[3]) != 0


def test_aec_batch_vs_oracle_cfg4_shape():
    from stanford_compression_library_b200.compressors.arithmetic_coding import AECParams, ArithmeticDecoder, ArithmeticEncoder
    from stanford_compression_library_b200.compressors.probability_models import AdaptiveIIDFreqModel
    from stanford_compression_library_b200.workloads import sample_blocks, zipf_freq_list

    fl = zipf_freq_list()
    data = sample_blocks(fl, 1024, 1024, seed=4, device="cuda:0")
    sizes = _seeded_sizes(1, 1025, 1024)
    params = AECParams()
    uni = [1] * 256
    enc = ArithmeticEncoder(params, AdaptiveIIDFreqModel(_F(uni), params.MAX_ALLOWED_TOTAL_FREQ))
    dec = ArithmeticDecoder(params, AdaptiveIIDFreqModel(_F(uni), params.MAX_ALLOWED_TOTAL_FREQ))
    oracle = so.Oracle.aec(uni)
    _compare_batch_with_oracle(enc, dec, oracle, data, sample=range(0, 1024, 9), consumed_equals_length=False)
    _compare_batch_with_oracle(enc, dec, oracle, data, sizes=sizes, sample=range(0, 1024, 11), consumed_equals_length=False)
    assert enc.freq_model.freqs_current.freq_list == uni  # batched calls do not mutate the host model


def test_aec_model_persists_across_encode_block_calls():
    # the reference never resets the model between blocks (data_encoder_decoder.py:23-27)
    from stanford_compression_library_b200 import DataBlock
    from stanford_compression_library_b200.compressors.arithmetic_coding import AECParams, ArithmeticDecoder, ArithmeticEncoder
    from stanford_compression_library_b200.compressors.probability_models import AdaptiveIIDFreqModel

    rng = np.random.default_rng(7)
    blocks = [rng.integers(0, 4, size=200).astype(np.uint8) for _ in range(3)]
    params = AECParams()
    enc = ArithmeticEncoder(params, AdaptiveIIDFreqModel(_F([1, 1, 1, 1]), params.MAX_ALLOWED_TOTAL_FREQ))
    dec = ArithmeticDecoder(params, AdaptiveIIDFreqModel(_F([1, 1, 1, 1]), params.MAX_ALLOWED_TOTAL_FREQ))
    oracle = so.Oracle.aec([1, 1, 1, 1])
    mf = np.array([1, 1, 1, 1], dtype=np.uint64)
    for blk in blocks:
        ref_bytes, ref_bits = oracle.encode_block(blk, model_freq=mf)
        ba = enc.encode_block(DataBlock(blk.tolist()))
        assert len(ba) == ref_bits and ba.tobytes() == ref_bytes.tobytes()
        assert [int(x) for x in enc.freq_model.freqs_current.freq_list] == mf.tolist()
        out, used = dec.decode_block(ba)
        assert list(out.data_list) == blk.tolist() and used == ref_bits


def _markov2(n, seed=0):
    """the source of the reference's order-k test (arithmetic_coding.py:388-402), restated"""
    rng = np.random.default_rng(seed)
    bits = rng.choice(2, size=n - 2)
    x = np.zeros(n, dtype=np.uint8)
    x[0], x[1] = rng.choice(3), rng.choice(3)
    for i in range(2, n):
        x[i] = (x[i - 1] + x[i - 2] + bits[i - 2]) % 3
    return x


def test_aec_order_k_reference_test_restated():
    """scl/compressors/arithmetic_coding.py:405-463: lossless + expected bitrate for k = 0..3 on a
    2nd-order Markov source, order 0 == the adaptive IID model exactly; every stream also against the oracle."""
    from stanford_compression_library_b200 import DataBlock
    from stanford_compression_library_b200.compressors.arithmetic_coding import AECParams, ArithmeticDecoder, ArithmeticEncoder
    from stanford_compression_library_b200.compressors.probability_models import AdaptiveIIDFreqModel, AdaptiveOrderKFreqModel

    n = 10000
    x = _markov2(n)
    blk = DataBlock(x.tolist())
    streams = {}
    for k, expected in ((0, np.log2(3)), (1, np.log2(3)), (2, 1.0), (3, 1.0)):
        params = AECParams()
        m = AdaptiveOrderKFreqModel([0, 1, 2], k, params.MAX_ALLOWED_TOTAL_FREQ)
        enc, dec = ArithmeticEncoder(params, m), ArithmeticDecoder(params, copy.deepcopy(m))
        ba = enc.encode_block(blk)
        out, used = dec.decode_block(ba)
        assert list(out.data_list) == x.tolist() and used in (len(ba), len(ba) - 1)
        assert abs(len(ba) / n - expected) < 0.1, (k, len(ba) / n)
        oracle = so.Oracle.aec([1, 1, 1], model=so.MODEL_ORDER_K, k=k)
        table = np.array([1] * 3 ** (k + 1) + [0], dtype=np.uint64)
        ref, ref_bits = oracle.encode_block(x, model_freq=table)
        assert len(ba) == ref_bits and ba.tobytes() == ref.tobytes()
        assert enc.freq_model._to_table() == table.tolist() == dec.freq_model._to_table()
        assert enc.freq_model.freqs_kplus1_tuple.shape == (3,) * (k + 1) and len(enc.freq_model.past_k) == k
        streams[k] = ba
    params = AECParams()
    iid = ArithmeticEncoder(params, AdaptiveIIDFreqModel(_F([1, 1, 1]), params.MAX_ALLOWED_TOTAL_FREQ))
    assert iid.encode_block(blk) == streams[0]


def test_aec_order_k_batched_and_persistent():
    from stanford_compression_library_b200.compressors.arithmetic_coding import AECParams, ArithmeticDecoder, ArithmeticEncoder
    from stanford_compression_library_b200.compressors.probability_models import AdaptiveOrderKFreqModel

    rng = np.random.default_rng(11)
    for n_sym, k, P in ((2, 5, 32), (4, 2, 32), (6, 2, 20), (16, 1, 32), (3, 5, 32)):
        B, N = 256, 700
        host = rng.integers(0, n_sym, size=(B, N)).astype(np.uint8)
        keep = rng.random((B, N)) < 0.7
        for j in range(1, N):
            host[:, j] = np.where(keep[:, j], host[:, j - 1], host[:, j])
        data = torch.from_numpy(host).cuda()
        sizes = _seeded_sizes(1, N + 1, B)
        params = AECParams(PRECISION=P)
        m = AdaptiveOrderKFreqModel(list(range(n_sym)), k, params.MAX_ALLOWED_TOTAL_FREQ)
        enc, dec = ArithmeticEncoder(params, m), ArithmeticDecoder(params, copy.deepcopy(m))
        oracle = so.Oracle.aec([1] * n_sym, PRECISION=P, model=so.MODEL_ORDER_K, k=k)
        _compare_batch_with_oracle(enc, dec, oracle, data, sample=range(0, B, 7), consumed_equals_length=False)
        _compare_batch_with_oracle(enc, dec, oracle, data, sizes=sizes, sample=range(0, B, 5), consumed_equals_length=False)
        assert enc.freq_model._to_table() == [1] * n_sym ** (k + 1) + [0]  # batched calls leave the host model alone
        # after a single-block call the model has moved on; batched calls then start every block from THAT state
        from stanford_compression_library_b200 import DataBlock

        ba = enc.encode_block(DataBlock(host[0, :100].tolist()))
        dec.decode_block(ba)
        table = enc.freq_model._to_table()
        assert table == dec.freq_model._to_table() and table != [1] * n_sym ** (k + 1) + [0]
        e = enc.encode_blocks(data[:8]).check()
        d = dec.decode_blocks(e, N).check()
        assert torch.equal(d.symbols[:, :N], data[:8])
        for b in range(8):
            ref, ref_bits = oracle.encode_block(host[b], model_freq=np.array(table, dtype=np.uint64))
            assert int(e.bit_len[b]) == ref_bits and e.block(b).tobytes() == ref.tobytes()
        assert enc.freq_model._to_table() == table


def test_aec_order_k_count_limit_and_table_limit():
    from stanford_compression_library_b200 import DataBlock
    from stanford_compression_library_b200.compressors.arithmetic_coding import AECParams, ArithmeticEncoder
    from stanford_compression_library_b200.compressors.probability_models import AdaptiveOrderKFreqModel

    params = AECParams()
    enc = ArithmeticEncoder(params, AdaptiveOrderKFreqModel([0, 1], 1, 20))
    with pytest.raises(AssertionError):  # a count reaches 20: the reference's update_model raises (probability_models.py:164-168)
        enc.encode_block(DataBlock([0] * 40))
    enc = ArithmeticEncoder(params, AdaptiveOrderKFreqModel(list(range(40)), 1, params.MAX_ALLOWED_TOTAL_FREQ))
    enc.encode_block(DataBlock([0, 1, 2]))  # 40 * 41 words do not fit shared memory: the table stays in HBM
    enc = ArithmeticEncoder(params, AdaptiveOrderKFreqModel(list(range(23)), 2, params.MAX_ALLOWED_TOTAL_FREQ))
    with pytest.raises(NotImplementedError):  # 529 contexts: more than the 512 rows of totals kept in shared memory
        enc.encode_block(DataBlock([0, 1, 2]))


def _orderk_large_cases():
    import json
    import os

    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "orderk_large_v1.npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    out = []
    for c in meta["cases"]:
        c = dict(c)
        c["data"], c["enc"], c["final"] = z["c%d_data" % c["id"]], z["c%d_enc" % c["id"]], z["c%d_final" % c["id"]]
        out.append(c)
    return out


@pytest.mark.parametrize("c", _orderk_large_cases(), ids=lambda c: "%d-%s" % (c["id"], c["note"][:34].replace(" ", "_")))
def test_aec_order_k_large_tables_match_reference_golden(c):
    """AdaptiveOrderKFreqModel over a BYTE alphabet at k = 1 (and other tables too large for shared memory), on the
    device, against vectors from the unmodified reference (oracle/gen_golden_orderk_large.py; probability_models.py:95-160):
    bits, num_bits_consumed on stream + garbage, the model object's final counts and context -- through the drop-in
    classes; then a batch of blocks, each from its own copy of the model, against the oracle."""
    from stanford_compression_library_b200 import BitArray, DataBlock
    from stanford_compression_library_b200.compressors.arithmetic_coding import AECParams, ArithmeticDecoder, ArithmeticEncoder
    from stanford_compression_library_b200.compressors.probability_models import AdaptiveOrderKFreqModel

    n_sym, k, p = len(c["freqs"]), c["model"]["k"], c["params"]
    params = AECParams(DATA_BLOCK_SIZE_BITS=p["DATA_BLOCK_SIZE_BITS"], PRECISION=p["PRECISION"])
    enc = ArithmeticEncoder(params, AdaptiveOrderKFreqModel(list(range(n_sym)), k, c["model"]["max_total"]))
    dec = ArithmeticDecoder(params, AdaptiveOrderKFreqModel(list(range(n_sym)), k, c["model"]["max_total"]))
    want_final = c["final"].astype(np.int64).tolist() + [c["model"]["final_ctx"]]
    ba = enc.encode_block(DataBlock(c["data"].tolist()))
    assert len(ba) == c["nbits"] and ba.tobytes() == c["enc"].tobytes()
    assert enc.freq_model._to_table() == want_final
    packed, total = with_garbage(c["enc"], c["nbits"], c["garbage"])
    block, used = dec.decode_block(BitArray.from_packed(packed, total))
    assert block.data_list == c["data"].tolist() and used == c["consumed"]
    assert dec.freq_model._to_table() == want_final
    # batched: B blocks, each starting from a copy of a fresh model
    B, N = 40, min(300, max(8, c["n"]))
    rng = np.random.default_rng(c["id"])
    host = rng.integers(0, n_sym, size=(B, N)).astype(np.uint8)
    host[:, 1::2] = host[:, 0::2][:, : host[:, 1::2].shape[1]]  # repeated symbols: contexts matter
    enc2 = ArithmeticEncoder(params, AdaptiveOrderKFreqModel(list(range(n_sym)), k, c["model"]["max_total"]))
    dec2 = ArithmeticDecoder(params, AdaptiveOrderKFreqModel(list(range(n_sym)), k, c["model"]["max_total"]))
    data = torch.from_numpy(host).cuda()
    e = enc2.encode_blocks(data).check()
    d = dec2.decode_blocks(e, N).check()
    assert torch.equal(d.symbols[:, :N], data)
    oracle = so.Oracle.aec([1] * n_sym, DATA_BLOCK_SIZE_BITS=p["DATA_BLOCK_SIZE_BITS"], PRECISION=p["PRECISION"], model=so.MODEL_ORDER_K, k=k,
                           max_allowed_total_freq=c["model"]["max_total"])
    fresh = np.array([1] * (n_sym ** (k + 1)) + [0], dtype=np.uint64)
    for b in (0, 1, B - 1):
        ref, ref_bits = oracle.encode_block(host[b], model_freq=fresh.copy())
        assert int(e.bit_len[b]) == ref_bits and e.block(b).tobytes() == ref.tobytes()


@pytest.mark.parametrize("coder", ["rans_default", "rans_nbo8", "range", "aec"])
def test_pack_and_frame_kernel_generations_agree(coder):
    """pack / frame, second generation (warp per block, 16-byte chunks by funnel shift) against the byte-wise
    first generation: identical bytes for bit-misaligned (61-bit rANS header), byte-aligned (NUM_BITS_OUT=8,
    range coder) and forward bit-granular (arithmetic) streams of ragged lengths, empty blocks included."""
    from stanford_compression_library_b200 import _cabi
    from stanford_compression_library_b200.compressors.arithmetic_coding import AECParams, ArithmeticEncoder
    from stanford_compression_library_b200.compressors.probability_models import AdaptiveIIDFreqModel
    from stanford_compression_library_b200.compressors.rANS import rANSEncoder, rANSParams
    from stanford_compression_library_b200.compressors.range_coder import RangeCoderParams, RangeEncoder
    from stanford_compression_library_b200.workloads import sample_blocks, zipf_freq_list, zipf_frequencies

    fl = zipf_freq_list()
    B, N = 700, 4096
    data = sample_blocks(fl, B, N, seed=41, device="cuda:0")
    sizes = _seeded_sizes(0, N + 1, B)
    sizes[:4] = torch.tensor([0, 1, 15, N], dtype=sizes.dtype)
    if coder == "rans_default":
        enc = rANSEncoder(rANSParams(zipf_frequencies()))
    elif coder == "rans_nbo8":
        enc = rANSEncoder(rANSParams(zipf_frequencies(), NUM_BITS_OUT=8, RANGE_FACTOR=1 << 12))
    elif coder == "range":
        enc = RangeEncoder(RangeCoderParams(), zipf_frequencies())
    else:
        ap = AECParams()
        enc = ArithmeticEncoder(ap, AdaptiveIIDFreqModel(_F([1] * 256), ap.MAX_ALLOWED_TOTAL_FREQ))
        data, sizes = data[:, :1024].contiguous(), torch.clamp(sizes, max=1024)
    e = enc.encode_blocks(data, sizes=sizes).check()
    p1, (f1, o1) = e.pack(bytewise=True), e.frame(bytewise=True)
    p2, (f2, o2) = e.pack(), e.frame()
    pp1 = p2.pack()  # packing an already packed (byte-aligned, contiguous) buffer is the identity
    assert torch.equal(p1.buf, p2.buf) and torch.equal(p1.bit_offset, p2.bit_offset)
    assert torch.equal(f1, f2) and torch.equal(o1, o2)
    assert torch.equal(pp1.buf, p2.buf)
    for b in (0, 1, 2, 3, 4, B - 1):  # and against the host's BitArray.tobytes()
        off, nb = int(p2.bit_offset[b]) // 8, (int(e.bit_len[b]) + 7) // 8
        assert p2.buf[off : off + nb].cpu().numpy().tobytes() == e.block(b).tobytes()


def test_pack_and_frame_kernels():
    from stanford_compression_library_b200.compressors.rANS import rANSDecoder, rANSEncoder, rANSParams
    from stanford_compression_library_b200.workloads import sample_blocks, zipf_freq_list, zipf_frequencies

    fl = zipf_freq_list()
    params = rANSParams(zipf_frequencies())  # 61-bit header: streams are not byte aligned
    data = sample_blocks(fl, 300, 512, seed=5, device="cuda:0")
    sizes = _seeded_sizes(0, 513, 300)
    enc = rANSEncoder(params)
    e = enc.encode_blocks(data, sizes=sizes).check()
    packed = e.pack()
    framed, foffs = e.frame()
    hb, ho, hl = packed.buf.cpu().numpy(), (packed.bit_offset // 8).cpu().numpy(), e.bit_len.cpu().numpy()
    fb, fo = framed.cpu().numpy(), foffs.cpu().numpy()
    for b in range(300):
        ba = e.block(b)
        nb = (len(ba) + 7) // 8
        assert hb[ho[b] : ho[b] + nb].tobytes() == ba.tobytes()
        # reference framing (encoded_stream.py:22-46, 93-103): header + padded payload
        n = len(ba)
        num_pad = (8 - (n + 3) % 8) % 8
        bits = np.concatenate([[(num_pad >> 2) & 1, (num_pad >> 1) & 1, num_pad & 1], np.zeros(num_pad, dtype=np.uint8), np.array(ba.tolist(), dtype=np.uint8)]).astype(np.uint8)
        payload = np.packbits(bits).tobytes()
        want = len(payload).to_bytes(4, "big") + payload
        assert fb[fo[b] : fo[b + 1]].tobytes() == want
    # decoding straight from the packed (left-aligned) form gives the same result
    d = rANSDecoder(params).decode_blocks(packed, 512).check()
    assert torch.equal(d.sizes, sizes) and torch.equal(d.bits_consumed, e.bit_len)


@pytest.mark.parametrize("kw", [{}, dict(NUM_BITS_OUT=8, RANGE_FACTOR=1 << 12)], ids=["default", "nbo8_rf12"])
def test_rans_full_size_roundtrip_cfg2(kw):
    """BASELINE cfg2 at full size: 65536 blocks x 4 KiB.  Properties: exact round trip, bits
    consumed == bits produced for every block, end state accepted, compressed size close to the
    table's cross-entropy; plus a strided sample of blocks bit-compared with the oracle."""
    from stanford_compression_library_b200.compressors.rANS import rANSDecoder, rANSEncoder, rANSParams
    from stanford_compression_library_b200.workloads import sample_blocks, zipf_freq_list, zipf_frequencies, zipf_probabilities

    fl = zipf_freq_list()
    params = rANSParams(zipf_frequencies(), **kw)
    B, N = 65536, 4096
    data = sample_blocks(zipf_probabilities(), B, N, seed=6, device="cuda:0")  # true Zipf-1.0 source
    enc, dec = rANSEncoder(params), rANSDecoder(params)
    e, d = _compare_batch_with_oracle(enc, dec, so.Oracle.rans(fl, **kw), data, sample=range(0, B, 2048))
    q = np.array(fl) / 4096.0
    xent = float(-(np.array(zipf_probabilities()) * np.log2(q)).sum())  # 6.2296 bits/symbol
    bits_per_sym = float(e.bit_len.sum()) / (B * N)
    assert abs(bits_per_sym - xent) < 0.05, (bits_per_sym, xent)


def test_try_lossless_compression_harness_with_trailing_bits():
    # mirrors scl/utils/test_utils.py:73-108 + rANS.py:363-401 on the drop-in classes
    from stanford_compression_library_b200 import Frequencies
    from stanford_compression_library_b200.compressors.rANS import rANSDecoder, rANSEncoder, rANSParams
    from stanford_compression_library_b200.utils.test_utils import get_random_data_block, try_lossless_compression

    freqs_list = [
        Frequencies({"A": 1, "B": 1, "C": 2}),
        Frequencies({"A": 12, "B": 34, "C": 1, "D": 45}),
        Frequencies({"A": 34, "B": 35, "C": 546, "D": 1, "E": 13, "F": 245}),
        Frequencies({"A": 5, "B": 5, "C": 5, "D": 5, "E": 5, "F": 5}),
        Frequencies({"A": 1, "B": 3}),
    ]
    params_list = [
        rANSParams(freqs_list[0]),
        rANSParams(freqs_list[1]),
        rANSParams(freqs_list[2], NUM_BITS_OUT=8),
        rANSParams(freqs_list[3], RANGE_FACTOR=1 << 12),
        rANSParams(freqs_list[4], RANGE_FACTOR=1 << 4),
    ]
    for freq, params in zip(freqs_list, params_list):
        block = get_random_data_block(freq.get_prob_dist(), 10000, seed=0)
        ok, nbits, _ = try_lossless_compression(block, rANSEncoder(params), rANSDecoder(params), add_extra_bits_to_encoder_output=True)
        assert ok


@pytest.mark.parametrize("kw", [{}, dict(NUM_BITS_OUT=8, RANGE_FACTOR=1 << 12)], ids=["default", "nbo8_rf12"])
@pytest.mark.parametrize("shape", [(33, 64), (1000, 100), (4097, 4096), (96, 8200)], ids=lambda s: "%dx%d" % s)
def test_rans_kernel_generations_agree(kw, shape):
    """The TMA/ring kernels (v2, default for uniform batches) and the first-generation kernels must
    produce identical streams and decode each other's output, for block counts that are not a
    multiple of 32 and block lengths that are not a multiple of the 64-byte tile."""
    from stanford_compression_library_b200 import _cabi
    from stanford_compression_library_b200.compressors.rANS import rANSDecoder, rANSEncoder, rANSParams
    from stanford_compression_library_b200.workloads import sample_blocks, zipf_freq_list, zipf_frequencies, zipf_probabilities

    B, N = shape
    params = rANSParams(zipf_frequencies(), **kw)
    enc, dec = rANSEncoder(params), rANSDecoder(params)
    data = sample_blocks(zipf_probabilities(), B, N, seed=11, device="cuda:0")
    data[0, :] = 255  # rarest symbol: worst-case bits per symbol
    try:
        _force(1, enc, dec)
        e1 = enc.encode_blocks(data).check()
        p1 = e1.pack(bytewise=True)
        _force(0, enc, dec)
        e2 = enc.encode_blocks(data).check()
        p2 = e2.pack()
        assert torch.equal(e1.bit_len, e2.bit_len) and torch.equal(e1.bit_offset, e2.bit_offset)
        assert torch.equal(p1.buf, p2.buf)
        d2 = dec.decode_blocks(e1, N).check()  # v2 decoder (TMA tile stores) on v1 output
        _force(2, enc, dec)
        d3 = dec.decode_blocks(e1, N).check()  # v2 decoder with per-lane sector stores
        _force(3, enc, dec)
        d4 = dec.decode_blocks(e1, N).check()  # v2 decoder, pipe-balanced instruction selection (large-batch form)
        _force(4, enc, dec)
        d5 = dec.decode_blocks(e1, N).check()  # v2 decoder, plain form
        _force(1, enc, dec)
        d1 = dec.decode_blocks(e2, N).check()  # v1 decoder on v2 output
    finally:
        _force(0, enc, dec)
    for d in (d1, d2, d3, d4, d5):
        assert torch.equal(d.symbols[:, :N], data) and torch.equal(d.bits_consumed, e1.bit_len)
        assert int(d.sizes.min()) == N == int(d.sizes.max())
    # and both against the oracle on a few blocks
    oracle = so.Oracle.rans(zipf_freq_list(), **kw)
    host = data.cpu().numpy()
    for b in (0, 1, B - 1):
        ref_bytes, ref_bits = oracle.encode_block(host[b])
        got = e2.block(b)
        assert len(got) == ref_bits and got.tobytes() == ref_bytes.tobytes()


@pytest.mark.parametrize("rf", [1, 4], ids=["L4096", "L16384"])
def test_tans_kernel_generations_agree(rf):
    from stanford_compression_library_b200 import _cabi
    from stanford_compression_library_b200.compressors.tANS import tANSDecoder, tANSEncoder, tANSParams
    from stanford_compression_library_b200.workloads import sample_blocks, zipf_freq_list, zipf_frequencies, zipf_probabilities

    B, N = 1000, 4100
    params = tANSParams(zipf_frequencies(), RANGE_FACTOR=rf)
    enc, dec = tANSEncoder(params), tANSDecoder(params)
    data = sample_blocks(zipf_probabilities(), B, N, seed=12, device="cuda:0")
    data[0, :] = 255
    try:
        _force(1, enc, dec)
        e1 = enc.encode_blocks(data).check()
        p1 = e1.pack(bytewise=True)
        _force(0, enc, dec)
        e2 = enc.encode_blocks(data).check()
        assert torch.equal(e1.bit_len, e2.bit_len) and torch.equal(p1.buf, e2.pack().buf)
        d2 = dec.decode_blocks(e1, N).check()
        _force(3, enc, dec)
        d3 = dec.decode_blocks(e1, N).check()  # pipe-balanced form
        _force(4, enc, dec)
        d4 = dec.decode_blocks(e1, N).check()  # plain form
        _force(1, enc, dec)
        d1 = dec.decode_blocks(e2, N).check()
    finally:
        _force(0, enc, dec)
    for d in (d1, d2, d3, d4):
        assert torch.equal(d.symbols[:, :N], data) and torch.equal(d.bits_consumed, e1.bit_len)
    oracle = so.Oracle.tans(zipf_freq_list(), RANGE_FACTOR=rf)
    host = data.cpu().numpy()
    for b in (0, 1, B - 1):
        ref_bytes, ref_bits = oracle.encode_block(host[b])
        got = e2.block(b)
        assert len(got) == ref_bits and got.tobytes() == ref_bytes.tobytes()


def test_host_pipeline_roundtrip_matches_direct_api():
    """HostCodecPipeline (pinned host buffers, chunked over two streams) must produce exactly the
    packed bytes of the direct device API, for a block count that is not a multiple of the chunk."""
    from stanford_compression_library_b200.compressors.rANS import rANSDecoder, rANSEncoder, rANSParams
    from stanford_compression_library_b200.pipeline import HostCodecPipeline
    from stanford_compression_library_b200.workloads import sample_blocks, zipf_frequencies, zipf_probabilities

    B, N = 5000, 1024
    params = rANSParams(zipf_frequencies())
    enc, dec = rANSEncoder(params), rANSDecoder(params)
    data = sample_blocks(zipf_probabilities(), B, N, seed=21, device="cuda:0")
    host_in = torch.empty((B, N), dtype=torch.uint8, pin_memory=True)
    host_in.copy_(data)
    pipe = HostCodecPipeline(enc, dec, N, B, chunk_blocks=1536)
    host_c = torch.empty(pipe.max_packed_bytes(), dtype=torch.uint8, pin_memory=True)
    host_out = torch.zeros((B, N), dtype=torch.uint8, pin_memory=True)
    for _ in range(2):  # second pass reuses every staging buffer
        total, lens = pipe.encode(host_in, host_c)
        direct = enc.encode_blocks(data).check()
        packed = direct.pack()
        assert torch.equal(lens, direct.bit_len.cpu())
        assert total == direct.total_bytes()
        assert torch.equal(host_c[:total], packed.buf[:total].cpu())
        host_out.zero_()
        pipe.decode(host_c, lens, host_out)
        assert torch.equal(host_out, host_in)


def test_file_level_roundtrip_and_reference_file_format(tmp_path):
    """§8(f) rows 1+3: batched file encode writes the reference's EncodedBlockWriter format
    (device framing kernel) byte-for-byte equal to framing each block's BitArray on the host, the
    reference-style reader/decoder reads it back, and the reference API encode_file/decode_file works."""
    from stanford_compression_library_b200 import Frequencies
    from stanford_compression_library_b200.compressors._gpu_base import decode_uint8_file, encode_uint8_file
    from stanford_compression_library_b200.compressors.rANS import rANSDecoder, rANSEncoder, rANSParams
    from stanford_compression_library_b200.core.data_stream import Uint8FileDataStream
    from stanford_compression_library_b200.core.encoded_stream import EncodedBlockReader, EncodedBlockWriter
    from stanford_compression_library_b200.workloads import zipf_frequencies, zipf_probabilities

    rng = np.random.default_rng(9)
    raw = rng.choice(256, size=100_000 + 37, p=np.array(zipf_probabilities())).astype(np.uint8)
    src, enc_a, enc_b, back = (str(tmp_path / n) for n in ("in.bin", "a.scl", "b.scl", "out.bin"))
    open(src, "wb").write(raw.tobytes())
    params = rANSParams(zipf_frequencies())  # 61-bit header: every block needs bit-granular framing
    enc, dec = rANSEncoder(params), rANSDecoder(params)
    encode_uint8_file(enc, src, enc_a, block_size=1000, blocks_per_batch=64)  # 101 blocks, last one ragged, 2 batches
    # (a) reference-style path: one block at a time through encode_block + the host framing classes
    with Uint8FileDataStream(src, "rb") as fds, EncodedBlockWriter(enc_b) as w:
        enc.encode(fds, block_size=1000, encode_writer=w)
    assert open(enc_a, "rb").read() == open(enc_b, "rb").read()
    # (b) batched decode of the file
    decode_uint8_file(dec, enc_a, back, block_size=1000)
    assert open(back, "rb").read() == raw.tobytes()
    # (c) block-at-a-time decode through the reference loop (asserts bits consumed == block length)
    with EncodedBlockReader(enc_a) as r, Uint8FileDataStream(back, "wb") as out:
        dec.decode(r, out)
    assert open(back, "rb").read() == raw.tobytes()
    # (d) the reference's text-file helpers
    text = "abracadabra " * 200
    tsrc, tenc, tback = (str(tmp_path / n) for n in ("t.txt", "t.scl", "t.out"))
    open(tsrc, "w").write(text)
    counts = {ch: text.count(ch) for ch in sorted(set(text))}
    tparams = rANSParams(Frequencies(counts))
    rANSEncoder(tparams).encode_file(tsrc, tenc, block_size=500)
    rANSDecoder(tparams).decode_file(tenc, tback)
    assert open(tback).read() == text


def test_histogram_blocks_vs_numpy_and_end_to_end_model():
    """§8(f) row 2: device histogram == DataBlock.get_counts semantics (numpy restatement), and a table
    built from the data itself (normalised to 2^12) codes that data bit-exactly like the oracle."""
    from stanford_compression_library_b200 import DataBlock
    from stanford_compression_library_b200.compressors.rANS import rANSDecoder, rANSEncoder, rANSParams
    from stanford_compression_library_b200.stats import empirical_frequencies, histogram_blocks
    from stanford_compression_library_b200.workloads import sample_blocks, zipf_probabilities

    B, N = 777, 4096 + 48
    data = sample_blocks(zipf_probabilities(), B, N, seed=31, device="cuda:0")
    data[:, ::7] = 0  # extra contention on one bin
    sizes = _seeded_sizes(0, N + 1, B)
    host = data.cpu().numpy()
    for sz in (None, sizes):
        counts, tot = histogram_blocks(data, sizes=sz)
        hs = np.full(B, N) if sz is None else sz.cpu().numpy()
        ref = np.stack([np.bincount(host[b, : hs[b]], minlength=256) for b in range(B)])
        assert np.array_equal(counts.cpu().numpy(), ref)
        assert np.array_equal(tot.cpu().numpy(), ref.sum(0))
    # same numbers as the reference-style host method on one block
    blk = DataBlock(host[3].tolist())
    gc = blk.get_counts()
    assert all(gc.get(v, 0) == int(ref_v) for v, ref_v in enumerate(np.bincount(host[3], minlength=256)))
    # model from data -> code the data
    freqs = empirical_frequencies(data, total_freq=4096)
    assert int(freqs.total_freq) == 4096
    params = rANSParams(freqs)
    enc, dec = rANSEncoder(params), rANSDecoder(params)
    e = enc.encode_blocks(data).check()
    d = dec.decode_blocks(e, N).check()
    assert torch.equal(d.symbols[:, :N], data)
    alpha = list(freqs.freq_dict)
    oracle = so.Oracle.rans([freqs.freq_dict[a] for a in alpha])
    idx_of = np.full(256, 255, dtype=np.uint8)
    idx_of[alpha] = np.arange(len(alpha), dtype=np.uint8)
    for b in (0, B - 1):
        ref_bytes, ref_bits = oracle.encode_block(idx_of[host[b]])
        got = e.block(b)
        assert len(got) == ref_bits and got.tobytes() == ref_bytes.tobytes()


def test_histogram_lane_per_block_form_vs_torch():
    """The histogram's lane-per-block form (batches of at least one block per lane of the grid, 32-byte aligned rows):
    per-block counts and grid totals, uniform and ragged sizes, totals alone (running counts, no per-block output) --
    against torch.bincount on every block."""
    from stanford_compression_library_b200.stats import histogram_blocks
    from stanford_compression_library_b200.workloads import sample_blocks, zipf_probabilities

    B, N = 148 * 7 * 32 * 2 + 777, 608  # N % 32 == 0 (aligned rows), the tail loop runs for ragged sizes
    data = sample_blocks(zipf_probabilities(), B, N, seed=41, device="cuda:0")
    data[:, ::5] = 0
    g = torch.Generator(device="cuda:0")
    g.manual_seed(3)
    sizes = torch.randint(0, N + 1, (B,), generator=g, device="cuda:0", dtype=torch.int32)
    sizes[:4] = torch.tensor([0, 1, 31, N], dtype=torch.int32, device="cuda:0")
    for sz in (None, sizes):
        n = torch.full((B,), N, device="cuda:0", dtype=torch.int64) if sz is None else sz.to(torch.int64)
        mask = torch.arange(N, device="cuda:0")[None, :] < n[:, None]
        keys = (torch.arange(B, device="cuda:0", dtype=torch.int64)[:, None] * 256 + data.to(torch.int64))[mask]
        ref = torch.bincount(keys, minlength=B * 256).reshape(B, 256)
        counts, tot = histogram_blocks(data, sizes=sz)
        assert torch.equal(counts.to(torch.int64), ref) and torch.equal(tot, ref.sum(0))
        _, tot_only = histogram_blocks(data, sizes=sz, per_block=False)
        assert torch.equal(tot_only, ref.sum(0))
        counts_only, none = histogram_blocks(data, sizes=sz, total=False)
        assert none is None and torch.equal(counts_only.to(torch.int64), ref)


def test_aec_kernel_generations_agree():
    """arithmetic coder: the closed-form / dp2a kernels (default for batches) vs the loop-literal
    first-generation kernels, plus the oracle on sampled blocks; ragged sizes; PRECISION 32 and 16."""
    from stanford_compression_library_b200 import _cabi
    from stanford_compression_library_b200.compressors.arithmetic_coding import AECParams, ArithmeticDecoder, ArithmeticEncoder
    from stanford_compression_library_b200.compressors.probability_models import AdaptiveIIDFreqModel, FixedFreqModel
    from stanford_compression_library_b200.workloads import sample_blocks, zipf_freq_list, zipf_probabilities

    B, N = 700, 1024
    data = sample_blocks(zipf_probabilities(), B, N, seed=41, device="cuda:0")
    data[0, :] = 255
    data[1, :] = 0
    sizes = torch.from_numpy(np.random.default_rng(42).integers(1, N + 1, size=B).astype(np.int32)).cuda()
    sizes[2:40] = torch.arange(1, 39, dtype=torch.int32, device="cuda:0")  # many very short blocks: where the quirk lives
    host = data.cpu().numpy()
    hs = sizes.cpu().numpy()
    for P, mk in ((32, lambda p: AdaptiveIIDFreqModel(_F([1] * 256), p.MAX_ALLOWED_TOTAL_FREQ)), (16, lambda p: AdaptiveIIDFreqModel(_F([1] * 256), p.MAX_ALLOWED_TOTAL_FREQ)),
                  (32, lambda p: AdaptiveIIDFreqModel(_F([1] * 256), 700)), (32, lambda p: FixedFreqModel(_F(zipf_freq_list()), p.MAX_ALLOWED_TOTAL_FREQ))):
        params = AECParams(PRECISION=P)
        enc, dec = ArithmeticEncoder(params, mk(params)), ArithmeticDecoder(params, mk(params))
        try:
            _force(1, enc, dec)
            e1 = enc.encode_blocks(data, sizes=sizes).check()
            p1 = e1.pack(bytewise=True)
            _force(0, enc, dec)
            e2 = enc.encode_blocks(data, sizes=sizes).check()
            assert torch.equal(e1.bit_len, e2.bit_len) and torch.equal(p1.buf, e2.pack().buf)
            d2 = dec.decode_blocks(e1, N).check()
            _force(1, enc, dec)
            d1 = dec.decode_blocks(e2, N).check()
        finally:
            _force(0, enc, dec)
        mask = torch.arange(N, device="cuda:0")[None, :] < sizes[:, None]
        for d in (d1, d2):
            assert torch.equal(d.sizes, sizes)
            assert torch.equal(d.symbols[:, :N][mask], data[mask])
        # Both generations must report the same num_bits_consumed as the oracle for EVERY block.  It is
        # usually the stream length, but the reference's trailing-bit accounting
        # (arithmetic_coding.py:277-282) returns len - 1 on some blocks (always when the block holds only
        # the first alphabet symbol, ~2.5 % of short random blocks; tests/test_oracle_golden.py pins this
        # against the live reference), so the oracle -- not the length -- is the expectation.
        assert torch.equal(d1.bits_consumed, d2.bits_consumed)
        m = enc.freq_model
        kind = so.MODEL_FIXED if isinstance(m, FixedFreqModel) else so.MODEL_ADAPTIVE_IID
        oracle = so.Oracle.aec([int(f) for f in m.freqs_current.freq_list], PRECISION=P, model=kind, max_allowed_total_freq=int(m.max_allowed_total_freq))
        used = d2.bits_consumed.cpu().numpy()
        ref_out, ref_bits, ref_st = oracle.encode_batch(host, sizes=hs.astype(np.uint32), out_stride=e2.out_stride + 64)
        assert (ref_st == 0).all() and np.array_equal(ref_bits.astype(np.int64), e2.bit_len.cpu().numpy())
        _, _, ref_used, ref_st2 = oracle.decode_batch(ref_out, np.arange(B, dtype=np.uint64) * np.uint64((e2.out_stride + 64) * 8), ref_bits, N)
        assert (ref_st2 == 0).all() and np.array_equal(ref_used.astype(np.int64), used)
        for b in (0, 1, 2, B - 1):
            got = e2.block(b)
            assert got.tobytes() == ref_out[b, : (int(ref_bits[b]) + 7) // 8].tobytes()


def test_tans_full_size_roundtrip_cfg3():
    """BASELINE cfg3 at full size: tANS, 4096-state tables (RANGE_FACTOR=1), 65536 blocks x 4 KiB."""
    from stanford_compression_library_b200.compressors.rANS import rANSEncoder, rANSParams
    from stanford_compression_library_b200.compressors.tANS import tANSDecoder, tANSEncoder, tANSParams
    from stanford_compression_library_b200.workloads import sample_blocks, zipf_freq_list, zipf_frequencies, zipf_probabilities

    B, N = 65536, 4096
    params = tANSParams(zipf_frequencies(), RANGE_FACTOR=1)
    data = sample_blocks(zipf_probabilities(), B, N, seed=61, device="cuda:0")
    enc, dec = tANSEncoder(params), tANSDecoder(params)
    e, d = _compare_batch_with_oracle(enc, dec, so.Oracle.tans(zipf_freq_list(), RANGE_FACTOR=1), data, sample=range(0, B, 4096))
    r = rANSEncoder(rANSParams(zipf_frequencies(), RANGE_FACTOR=1)).encode_blocks(data).check()
    assert torch.equal(r.bit_len, e.bit_len) and torch.equal(r.pack().buf, e.pack().buf)  # tANS == rANS bit for bit


def test_aec_full_size_roundtrip_cfg4():
    """BASELINE cfg4 at full size: adaptive order-0 arithmetic coder, 1 048 576 blocks x 1 KiB, fresh
    uniform model per block.  Exact round trip for every block, oracle bits + consumed on a sample."""
    from stanford_compression_library_b200.compressors.arithmetic_coding import AECParams, ArithmeticDecoder, ArithmeticEncoder
    from stanford_compression_library_b200.compressors.probability_models import AdaptiveIIDFreqModel
    from stanford_compression_library_b200.workloads import sample_blocks, zipf_probabilities

    B, N = 1048576, 1024
    params = AECParams()
    uni = [1] * 256
    data = sample_blocks(zipf_probabilities(), B, N, seed=62, device="cuda:0")
    enc = ArithmeticEncoder(params, AdaptiveIIDFreqModel(_F(uni), params.MAX_ALLOWED_TOTAL_FREQ))
    dec = ArithmeticDecoder(params, AdaptiveIIDFreqModel(_F(uni), params.MAX_ALLOWED_TOTAL_FREQ))
    e, d = _compare_batch_with_oracle(enc, dec, so.Oracle.aec(uni), data, sample=range(0, B, 65536), consumed_equals_length=False)
    bits_per_sym = float(e.bit_len.sum()) / (B * N)
    assert 6.2 < bits_per_sym < 7.0  # Zipf-1.0 entropy 6.22 b/sym + the adaptive model's learning cost over 1 KiB


def test_aec_8bit_counter_model_agrees_with_16bit_and_oracle():
    """The arithmetic coder's 8-bit-counter model (AecModel8: half the shared memory, counters past 255 escape to a
    per-lane list) against the 16-bit kernels (debug path 5) and the oracle: cfg4-shaped blocks (1 KiB from 256 ones),
    rows that push one / two counters far past 255, a small max_allowed_total_freq that fires the halving rule with
    escaped counters present, ragged sizes; and the eligibility bound (blocks of 1280 symbols take the 16-bit kernels)."""
    from stanford_compression_library_b200 import Frequencies
    from stanford_compression_library_b200.compressors.arithmetic_coding import AECParams, ArithmeticDecoder, ArithmeticEncoder
    from stanford_compression_library_b200.compressors.probability_models import AdaptiveIIDFreqModel
    from stanford_compression_library_b200.workloads import sample_blocks, zipf_probabilities

    B, N = 4100, 1024
    data = sample_blocks(zipf_probabilities(), B, N, seed=77, device="cuda:0")
    data[0, :] = 7                       # one counter reaches 1025
    data[1, ::2] = 3
    data[1, 1::2] = 250                  # two counters reach 513
    data[2, :300] = 0
    data[3, :] = torch.arange(N, device="cuda:0") % 256  # flat
    sizes = _seeded_sizes(1, N + 1, B)
    sizes[:4] = N
    uni = Frequencies({b: 1 for b in range(256)})
    host, hs = data.cpu().numpy(), sizes.cpu().numpy()
    for max_total in (None, 700):
        ap = AECParams()
        mt = ap.MAX_ALLOWED_TOTAL_FREQ if max_total is None else max_total
        enc = ArithmeticEncoder(ap, AdaptiveIIDFreqModel(uni, mt))
        dec = ArithmeticDecoder(ap, AdaptiveIIDFreqModel(uni, mt))
        e8 = enc.encode_blocks(data, sizes=sizes).check()
        d8 = dec.decode_blocks(e8, N).check()
        try:
            _force(5, enc, dec)
            e16 = enc.encode_blocks(data, sizes=sizes).check()
            d16 = dec.decode_blocks(e8, N).check()
        finally:
            _force(0, enc, dec)
        assert torch.equal(e8.bit_len, e16.bit_len) and torch.equal(e8.pack().buf, e16.pack().buf)
        assert torch.equal(d8.bits_consumed, d16.bits_consumed) and torch.equal(d8.sizes, sizes) and torch.equal(d16.sizes, sizes)
        mask = torch.arange(N, device="cuda:0")[None, :] < sizes[:, None]
        assert torch.equal(d8.symbols[:, :N][mask], data[mask]) and torch.equal(d16.symbols[:, :N][mask], data[mask])
        oracle = so.Oracle.aec([1] * 256, max_allowed_total_freq=mt)
        for b in (0, 1, 2, 3, 4, 5, B - 1):
            ref, ref_bits = oracle.encode_block(host[b, : hs[b]])
            assert int(e8.bit_len[b]) == ref_bits and e8.block(b).tobytes() == ref.tobytes(), (max_total, b)
    # past the bound the library takes the 16-bit kernels by itself: same answers as the oracle
    long_blocks = sample_blocks(zipf_probabilities(), 64, 1280, seed=78, device="cuda:0")
    long_blocks[0, :] = 9
    enc = ArithmeticEncoder(AECParams(), AdaptiveIIDFreqModel(uni, AECParams().MAX_ALLOWED_TOTAL_FREQ))
    e = enc.encode_blocks(long_blocks).check()
    oracle = so.Oracle.aec([1] * 256)
    for b in (0, 1, 63):
        ref, ref_bits = oracle.encode_block(long_blocks[b].cpu().numpy())
        assert int(e.bit_len[b]) == ref_bits and e.block(b).tobytes() == ref.tobytes()
