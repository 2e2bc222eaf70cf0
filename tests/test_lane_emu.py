"""CPU tier: the product's per-lane recurrences (csrc/scl_lane.cuh) and host table builders
(csrc/scl_tables.hpp), compiled for the host by tests/host_emu, against the golden vectors from
the unmodified reference and against the C oracle on random inputs.  The same comparisons run
against the real CUDA kernels in tests/test_gpu_parity.py (-m gpu)."""
import numpy as np
import pytest

from oracle import scl_oracle as so
from tests.emu_util import EmuCoder, extract_bits, params_from_case
from tests.golden_util import case_id, expected_final_model, fresh_model_table, load_golden, with_garbage

CASES = load_golden()


def _roundtrip_case(c, force_generic=False):
    coder = EmuCoder(params_from_case(c), None, c["freqs"])
    if force_generic:
        coder.force_generic()
    n = c["n"]
    model = fresh_model_table(c)[None] if c["coder"] == "aec" else None
    out, off, ln, st = coder.encode(c["data"].reshape(1, -1), model=model)
    assert st[0] == 0
    assert int(ln[0]) == c["nbits"]
    assert extract_bits(out, off[0], ln[0]).tobytes() == c["enc"].tobytes()
    if c["coder"] == "aec":
        assert model[0].tolist() == expected_final_model(c)
    # decode stream + garbage placed at an odd bit offset
    packed, total = with_garbage(c["enc"], c["nbits"], c["garbage"])
    lead = 5
    bits = np.concatenate([np.ones(lead, dtype=np.uint8), np.unpackbits(packed)[:total]])
    buf = np.concatenate([np.packbits(bits), np.zeros(16, dtype=np.uint8)])
    model = fresh_model_table(c)[None] if c["coder"] == "aec" else None
    if c["coder"] == "aec" and n == 0:
        return
    sym, sizes, used, st = coder.decode(buf, [lead], [total], max(n, 1), model=model)
    assert st[0] == 0
    assert int(sizes[0]) == n
    assert sym[0, :n].tolist() == c["data"].tolist()
    assert int(used[0]) == c["consumed"]
    return coder


@pytest.mark.parametrize("c", CASES, ids=case_id)
def test_emu_matches_golden(c):
    _roundtrip_case(c)


@pytest.mark.parametrize("c", [c for c in CASES if c["coder"] == "rans"], ids=case_id)
def test_emu_generic_rans_path_matches_golden(c):
    _roundtrip_case(c, force_generic=True)


def test_fast_path_selection():
    by_note = {c["note"]: c for c in CASES}
    z = by_note["cfg2 zipf default params"]
    coder = EmuCoder(params_from_case(z), None, z["freqs"])
    assert coder.path(False) == 0 and coder.path(True) == 0  # both fast for the benchmark table
    z8 = by_note["cfg2 zipf NBO=8 RF=2^12"]
    coder = EmuCoder(params_from_case(z8), None, z8["freqs"])
    assert coder.path(False) == 0 and coder.path(True) == 0
    big = by_note["H ~ 2^46: 64-bit state"]
    coder = EmuCoder(params_from_case(big), None, big["freqs"])
    assert coder.path(False) == 1 and coder.path(True) == 1
    npo2 = by_note["cfg1 counts+1 (M not a power of two) default params"]
    coder = EmuCoder(params_from_case(npo2), None, npo2["freqs"])
    assert coder.path(False) == 0 and coder.path(True) == 1  # magic-number division still applies to encode


def test_tans_tables_kat():
    c = next(c for c in CASES if c["coder"] == "tans" and c["note"].startswith("KAT"))
    coder = EmuCoder(params_from_case(c), None, c["freqs"])
    enc, dec = coder.tans_tables(8)
    assert enc.tolist() == [8, 9, 10, 11, 12, 13, 14, 15]  # tANS.py:297-306
    assert [(int(e) & 0xFF, int(e) >> 8) for e in dec] == [(0, 3), (0, 4), (0, 5), (1, 3), (1, 4), (1, 5), (2, 2), (2, 3)]  # :322-331


def _random_freqs(rng, n_sym, total=None):
    f = rng.integers(1, 50, size=n_sym).astype(np.int64)
    if total is not None:  # normalise to an exact total, keep >= 1
        f = np.maximum(1, np.floor(f / f.sum() * (total - n_sym)).astype(np.int64) + 1)
        f[0] += total - f.sum()
        assert f.min() >= 1 and f.sum() == total
    return [int(x) for x in f]


@pytest.mark.parametrize("seed", range(8))
def test_emu_vs_oracle_random_rans_tans(seed):
    from stanford_compression_library_b200 import _cabi
    from stanford_compression_library_b200._cabi import SclParams

    rng = np.random.default_rng(seed)
    n_sym = int(rng.choice([2, 3, 17, 256]))
    pow2 = bool(rng.integers(0, 2))
    freqs = _random_freqs(rng, n_sym, total=int(rng.choice([256, 1024, 4096])) if pow2 else None)
    nbo = int(rng.choice([1, 1, 2, 4, 8]))
    rf = int(rng.choice([1, 2, 1 << 4, 1 << 8, 1 << 12, 1 << 16]))
    M = sum(freqs)
    H = rf * M * (1 << nbo) - 1
    nsb = so.ref_get_bit_width(H)
    B, N = 9, int(rng.integers(1, 300))
    p = np.array(freqs, dtype=np.float64)
    sym = rng.choice(n_sym, size=(B, N), p=p / p.sum()).astype(np.uint8)
    sizes = rng.integers(0, N + 1, size=B).astype(np.uint32)
    oracle = so.Oracle.rans(freqs, NUM_BITS_OUT=nbo, RANGE_FACTOR=rf)
    coders = [("rans", _cabi.CODER_RANS)]
    if pow2 and nbo == 1 and rf * M <= (1 << 20):
        coders.append(("tans", _cabi.CODER_TANS))
    for name, kind in coders:
        prm = SclParams(coder=kind, data_block_size_bits=32, num_bits_out=nbo, range_factor=rf, num_state_bits=nsb, precision=0, model=0,
                        max_allowed_total_freq=0)
        coder = EmuCoder(prm, None, freqs)
        out, off, ln, st = coder.encode(sym, sizes=sizes)
        assert (st == 0).all()
        for b in range(B):
            enc, nb = oracle.encode_block(sym[b, : sizes[b]])
            assert nb == ln[b], (name, b)
            assert extract_bits(out, off[b], ln[b]).tobytes() == enc.tobytes(), (name, b)
        dsym, dsz, used, st = coder.decode(out, off, ln, N)
        assert (st == 0).all() and (dsz == sizes).all() and (used == ln).all()
        for b in range(B):
            assert dsym[b, : sizes[b]].tolist() == sym[b, : sizes[b]].tolist()


@pytest.mark.parametrize("seed", range(4))
def test_emu_vs_oracle_random_range_aec(seed):
    from stanford_compression_library_b200 import _cabi
    from stanford_compression_library_b200._cabi import SclParams

    rng = np.random.default_rng(100 + seed)
    n_sym = int(rng.choice([2, 5, 64, 256]))
    freqs = _random_freqs(rng, n_sym)
    B, N = 6, int(rng.integers(1, 400))
    p = np.array(freqs, dtype=np.float64)
    sym = rng.choice(n_sym, size=(B, N), p=p / p.sum()).astype(np.uint8)
    sizes = rng.integers(1, N + 1, size=B).astype(np.uint32)
    # range coder
    oracle = so.Oracle.range_coder(freqs)
    prm = SclParams(coder=_cabi.CODER_RANGE, data_block_size_bits=32, num_bits_out=0, range_factor=0, num_state_bits=0, precision=32, model=0,
                    max_allowed_total_freq=0)
    coder = EmuCoder(prm, None, freqs)
    out, off, ln, st = coder.encode(sym, sizes=sizes)
    assert (st == 0).all()
    for b in range(B):
        enc, nb = oracle.encode_block(sym[b, : sizes[b]])
        assert nb == ln[b] and extract_bits(out, off[b], ln[b]).tobytes() == enc.tobytes()
    dsym, dsz, used, st = coder.decode(out, off, ln, N)
    assert (st == 0).all() and (dsz == sizes).all() and (used == ln).all()
    for b in range(B):
        assert dsym[b, : sizes[b]].tolist() == sym[b, : sizes[b]].tolist()
    # arithmetic coder, adaptive from uniform, fresh model per block; PRECISION 32 and 16
    for P, max_total in ((32, 1 << 30), (16, 1 << 14), (32, 700)):
        uni = [1] * n_sym
        oracle = so.Oracle.aec(uni, PRECISION=P, max_allowed_total_freq=max_total)
        prm = SclParams(coder=_cabi.CODER_AEC, data_block_size_bits=32, num_bits_out=0, range_factor=0, num_state_bits=0, precision=P,
                        model=_cabi.MODEL_ADAPTIVE_IID, max_allowed_total_freq=max_total)
        coder = EmuCoder(prm, None, uni)
        out, off, ln, st = coder.encode(sym, sizes=sizes)
        assert (st == 0).all(), st
        for b in range(B):
            enc, nb = oracle.encode_block(sym[b, : sizes[b]])
            assert nb == ln[b], (P, b)
            assert extract_bits(out, off[b], ln[b]).tobytes() == enc.tobytes()
        dsym, dsz, used, st = coder.decode(out, off, ln, N)
        assert (st == 0).all() and (dsz == sizes).all() and (used == ln).all()
        for b in range(B):
            assert dsym[b, : sizes[b]].tolist() == sym[b, : sizes[b]].tolist()


def test_emu_alphabet_mapping_and_bad_symbol():
    from stanford_compression_library_b200 import _cabi
    from stanford_compression_library_b200._cabi import SclParams

    freqs = [3, 3, 2]
    alphabet = np.array([65, 66, 67], dtype=np.uint8)  # 'A','B','C' coded as their byte values
    prm = SclParams(coder=_cabi.CODER_RANS, data_block_size_bits=5, num_bits_out=1, range_factor=1, num_state_bits=4, precision=0, model=0,
                    max_allowed_total_freq=0)
    coder = EmuCoder(prm, alphabet, freqs)
    out, off, ln, st = coder.encode(np.array([[65, 67, 66]], dtype=np.uint8))
    assert st[0] == 0 and so.bits_to_str(extract_bits(out, off[0], ln[0]), int(ln[0])) == "00011101110010"
    sym, sizes, used, st = coder.decode(out, off, ln, 3)
    assert sym[0, :3].tolist() == [65, 67, 66] and used[0] == 14
    out, off, ln, st = coder.encode(np.array([[65, 68, 66]], dtype=np.uint8))
    assert st[0] == _cabi.ST_BAD_SYMBOL


def test_emu_status_words():
    from stanford_compression_library_b200 import _cabi
    from stanford_compression_library_b200._cabi import SclParams

    freqs = [1, 1, 2]
    prm = SclParams(coder=_cabi.CODER_RANS, data_block_size_bits=32, num_bits_out=1, range_factor=1 << 16, num_state_bits=19, precision=0,
                    model=0, max_allowed_total_freq=0)
    coder = EmuCoder(prm, None, freqs)
    sym = np.array([[0, 1, 2, 2, 1, 0, 2, 2]], dtype=np.uint8)
    out, off, ln, st = coder.encode(sym)
    corrupt = out.copy()
    corrupt[int(off[0]) // 8 + 5] ^= 0x20  # flip a state bit
    _, _, _, st = coder.decode(corrupt, off, ln, 8)
    assert st[0] == _cabi.ST_STATE_MISMATCH
    # output slot too small -> overflow status, no out-of-bounds write (UBSan/ASan-clean)
    out, off, ln, st = coder.encode(np.zeros((1, 64), dtype=np.uint8) + 1, out_stride=16)
    assert st[0] == _cabi.ST_OVERFLOW
    # size does not fit DATA_BLOCK_SIZE_BITS
    prm2 = SclParams(coder=_cabi.CODER_RANS, data_block_size_bits=2, num_bits_out=1, range_factor=1 << 16, num_state_bits=19, precision=0,
                     model=0, max_allowed_total_freq=0)
    out, off, ln, st = EmuCoder(prm2, None, freqs).encode(sym)
    assert st[0] == _cabi.ST_OVERFLOW


# ---- second-generation lane structs (csrc/scl_fast.cuh: funnel-shift packing, sector ring I/O) ----
@pytest.mark.parametrize("c", [c for c in CASES if c["coder"] in ("rans", "tans")], ids=case_id)
def test_emu_v2_matches_golden(c):
    coder = EmuCoder(params_from_case(c), None, c["freqs"])
    if not coder.v2_eligible():
        pytest.skip("parameter set not eligible for the v2 fast path")
    n = c["n"]
    out, off, ln, st = coder.encode_v2(c["data"].reshape(1, -1))
    assert st[0] == 0 and int(ln[0]) == c["nbits"]
    assert extract_bits(out, off[0], ln[0]).tobytes() == c["enc"].tobytes()
    packed, total = with_garbage(c["enc"], c["nbits"], c["garbage"])
    for lead in (0, 5, 250, 261):
        bits = np.concatenate([np.ones(lead, dtype=np.uint8), np.unpackbits(packed)[:total]])
        buf = np.concatenate([np.packbits(bits), np.zeros(7, dtype=np.uint8)])
        sym, sizes, used, st = coder.decode_v2(buf, [lead], [total], max(n, 1))
        assert st[0] == 0 and int(sizes[0]) == n and int(used[0]) == c["consumed"]
        assert sym[0, :n].tolist() == c["data"].tolist()


@pytest.mark.parametrize("kw", [dict(nbo=1, rf=1 << 16), dict(nbo=8, rf=1 << 12), dict(nbo=1, rf=1 << 4)], ids=["default", "nbo8_rf12", "nbo1_rf4"])
def test_emu_v2_vs_oracle_zipf_lengths(kw):
    from stanford_compression_library_b200 import _cabi
    from stanford_compression_library_b200._cabi import SclParams
    from stanford_compression_library_b200.workloads import zipf_freq_list, zipf_probabilities

    fl = zipf_freq_list()
    rng = np.random.default_rng(3)
    nsb = so.ref_get_bit_width(kw["rf"] * 4096 * (1 << kw["nbo"]) - 1)
    prm = SclParams(coder=_cabi.CODER_RANS, data_block_size_bits=32, num_bits_out=kw["nbo"], range_factor=kw["rf"], num_state_bits=nsb,
                    precision=0, model=0, max_allowed_total_freq=0)
    coder = EmuCoder(prm, None, fl)
    assert coder.v2_eligible()
    oracle = so.Oracle.rans(fl, NUM_BITS_OUT=kw["nbo"], RANGE_FACTOR=kw["rf"])
    for N in (0, 1, 15, 16, 17, 31, 32, 33, 64, 100, 511, 1024, 4096):
        sym = rng.choice(256, size=(5, N), p=np.array(zipf_probabilities())).astype(np.uint8)
        if N:
            sym[0, :] = 255  # rarest symbol everywhere: the most bits per symbol the table can produce
        out, off, ln, st = coder.encode_v2(sym)
        assert (st == 0).all()
        for b in range(5):
            enc, nb = oracle.encode_block(sym[b])
            assert nb == ln[b], (N, b)
            assert extract_bits(out, off[b], ln[b]).tobytes() == enc.tobytes(), (N, b)
        dsym, dsz, used, st = coder.decode_v2(out, off, ln, max(N, 1))
        assert (st == 0).all() and (dsz == N).all() and (used == ln).all()
        assert (dsym[:, :N] == sym).all()


@pytest.mark.parametrize("rf", [1, 4])
def test_emu_tans_v2_vs_oracle_zipf_lengths(rf):
    from stanford_compression_library_b200 import _cabi
    from stanford_compression_library_b200._cabi import SclParams
    from stanford_compression_library_b200.workloads import zipf_freq_list, zipf_probabilities

    fl = zipf_freq_list()
    rng = np.random.default_rng(4)
    nsb = so.ref_get_bit_width(rf * 4096 * 2 - 1)
    prm = SclParams(coder=_cabi.CODER_TANS, data_block_size_bits=32, num_bits_out=1, range_factor=rf, num_state_bits=nsb, precision=0, model=0,
                    max_allowed_total_freq=0)
    coder = EmuCoder(prm, None, fl)
    assert coder.v2_eligible()
    oracle = so.Oracle.tans(fl, RANGE_FACTOR=rf)
    for N in (0, 1, 31, 32, 33, 100, 1024, 4096):
        sym = rng.choice(256, size=(4, N), p=np.array(zipf_probabilities())).astype(np.uint8)
        if N:
            sym[0, :] = 255
        out, off, ln, st = coder.encode_v2(sym)
        assert (st == 0).all()
        for b in range(4):
            enc, nb = oracle.encode_block(sym[b])
            assert nb == ln[b] and extract_bits(out, off[b], ln[b]).tobytes() == enc.tobytes(), (N, b)
        dsym, dsz, used, st = coder.decode_v2(out, off, ln, max(N, 1))
        assert (st == 0).all() and (dsz == N).all() and (used == ln).all() and (dsym[:, :N] == sym).all()


# ---- second-generation range-coder lanes (csrc/scl_range.cuh) -----------------------------------
@pytest.mark.parametrize("c", [c for c in CASES if c["coder"] == "range"], ids=case_id)
def test_emu_range_v2_matches_golden(c):
    coder = EmuCoder(params_from_case(c), None, c["freqs"])
    if not coder.v2_eligible():
        pytest.skip("parameter set not eligible for the range-coder v2 lanes")
    n = c["n"]
    out, off, ln, st = coder.encode_v2(c["data"].reshape(1, -1))
    assert st[0] == 0 and int(ln[0]) == c["nbits"]
    assert extract_bits(out, off[0], ln[0]).tobytes() == c["enc"].tobytes()
    packed, total = with_garbage(c["enc"], c["nbits"], c["garbage"])
    for lead in (0, 8, 248, 264):
        bits = np.concatenate([np.ones(lead, dtype=np.uint8), np.unpackbits(packed)[:total]])
        buf = np.concatenate([np.packbits(bits), np.zeros(7, dtype=np.uint8)])
        sym, sizes, used, st = coder.decode_v2(buf, [lead], [total], max(n, 1))
        assert st[0] == 0 and int(sizes[0]) == n and int(used[0]) == c["consumed"]
        assert sym[0, :n].tolist() == c["data"].tolist()


def _range_tables():
    from stanford_compression_library_b200.workloads import zipf_freq_list

    skew = [4096 - 255] + [1] * 255          # one dominant symbol: long settled runs, rare symbols release 2+ bytes
    two = [4095, 1]                           # the reference's hardest shape: underflow chains
    small_t = [5, 3, 7, 1]                    # T = 16, the smallest total the v2 lanes take
    flat = [16] * 256
    return {"zipf": zipf_freq_list(), "skew": skew, "two": two, "T16": small_t, "flat": flat}


@pytest.mark.parametrize("name", ["zipf", "skew", "two", "T16", "flat"])
def test_emu_range_v2_vs_oracle(name):
    from stanford_compression_library_b200 import _cabi
    from stanford_compression_library_b200._cabi import SclParams

    fl = _range_tables()[name]
    n_sym = len(fl)
    rng = np.random.default_rng(7)
    prm = SclParams(coder=_cabi.CODER_RANGE, data_block_size_bits=32, num_bits_out=0, range_factor=0, num_state_bits=0, precision=32, model=0,
                    max_allowed_total_freq=0)
    coder = EmuCoder(prm, None, fl)
    assert coder.v2_eligible()
    oracle = so.Oracle.range_coder(fl)
    p = np.asarray(fl, dtype=np.float64) / sum(fl)
    for N in (0, 1, 2, 15, 16, 17, 31, 32, 33, 63, 64, 100, 511, 1024, 4096):
        sym = rng.choice(n_sym, size=(6, N), p=p).astype(np.uint8)
        if N:
            sym[0, :] = n_sym - 1                     # rarest symbol everywhere
            sym[1, :] = 0                             # most frequent symbol everywhere
            sym[2, :] = rng.integers(0, n_sym, N)     # uniform draws: far from the table's distribution
        out, off, ln, st = coder.encode_v2(sym)
        assert (st == 0).all(), (N, st)
        for b in range(6):
            enc, nb = oracle.encode_block(sym[b])
            assert nb == ln[b], (N, b, nb, ln[b])
            assert extract_bits(out, off[b], ln[b]).tobytes() == enc.tobytes(), (N, b)
        dsym, dsz, used, st = coder.decode_v2(out, off, ln, max(N, 1))
        assert (st == 0).all() and (dsz == N).all() and (used == ln).all(), (N, st, dsz, used, ln)
        assert (dsym[:, :N] == sym).all()
        # trailing garbage and the oracle's decoder agree on a bit-shifted copy as well
        b = 3
        bits = np.concatenate([np.zeros(40, dtype=np.uint8), np.unpackbits(extract_bits(out, off[b], ln[b]))[: int(ln[b])], rng.integers(0, 2, 77).astype(np.uint8)])
        buf = np.concatenate([np.packbits(bits), np.zeros(8, dtype=np.uint8)])
        dsym, dsz, used, st = coder.decode_v2(buf, [40], [int(ln[b]) + 77], max(N, 1))
        assert st[0] == 0 and dsz[0] == N and used[0] == ln[b] and (dsym[0, :N] == sym[b]).all()


@pytest.mark.parametrize("seed", range(6))
def test_emu_range_v2_random_tables_vs_oracle(seed):
    """random power-of-two totals (16 .. 4096), random alphabet sizes and skews, data drawn uniformly (not from the
    table): the v2 lanes against the oracle, bit for bit, and back"""
    from stanford_compression_library_b200 import _cabi
    from stanford_compression_library_b200._cabi import SclParams

    rng = np.random.default_rng(100 + seed)
    T = 1 << int(rng.integers(4, 13))
    n_sym = int(rng.integers(2, min(256, T) + 1))
    # random composition of T into n_sym positive parts, skewed
    w = rng.random(n_sym) ** int(rng.integers(1, 6))
    f = np.maximum(1, np.floor(w / w.sum() * (T - n_sym)).astype(np.int64) + 1)
    while f.sum() > T:
        f[np.argmax(f)] -= 1
    f[np.argmin(f)] += T - f.sum()
    if f.max() > 4095:  # outside the v2 lanes' table format: split the surplus
        pytest.skip("f > 4095")
    fl = [int(x) for x in f]
    assert sum(fl) == T and min(fl) >= 1
    prm = SclParams(coder=_cabi.CODER_RANGE, data_block_size_bits=32, num_bits_out=0, range_factor=0, num_state_bits=0, precision=32, model=0,
                    max_allowed_total_freq=0)
    coder = EmuCoder(prm, None, fl)
    assert coder.v2_eligible()
    oracle = so.Oracle.range_coder(fl)
    for N in (1, 33, 257, 1500):
        sym = rng.integers(0, n_sym, size=(4, N)).astype(np.uint8)
        sym[0, :] = int(np.argmin(f))
        out, off, ln, st = coder.encode_v2(sym)
        assert (st == 0).all()
        for b in range(4):
            enc, nb = oracle.encode_block(sym[b])
            assert nb == ln[b] and extract_bits(out, off[b], ln[b]).tobytes() == enc.tobytes(), (T, n_sym, N, b)
        dsym, dsz, used, st = coder.decode_v2(out, off, ln, N)
        assert (st == 0).all() and (dsz == N).all() and (used == ln).all() and (dsym[:, :N] == sym).all()


def test_emu_range_v2_corrupt_stream_matches_v1_lanes():
    """garbage input: the v2 lanes must make the v1 (reference-literal) lanes' choices (last-symbol
    fallbacks of searchsorted) for as long as the stream lasts"""
    from stanford_compression_library_b200 import _cabi
    from stanford_compression_library_b200._cabi import SclParams
    from stanford_compression_library_b200.workloads import zipf_freq_list

    fl = zipf_freq_list()
    prm = SclParams(coder=_cabi.CODER_RANGE, data_block_size_bits=32, num_bits_out=0, range_factor=0, num_state_bits=0, precision=32, model=0,
                    max_allowed_total_freq=0)
    coder = EmuCoder(prm, None, fl)
    rng = np.random.default_rng(9)
    for trial in range(6):
        N = 200
        payload = rng.integers(0, 256, 600).astype(np.uint8)
        buf = np.concatenate([np.array([0, 0, 0, N], dtype=np.uint8), payload, np.zeros(40, dtype=np.uint8)])
        total = 8 * (4 + 600)
        s1, z1, u1, st1 = coder.decode(buf, [0], [total], N)
        s2, z2, u2, st2 = coder.decode_v2(buf, [0], [total], N)
        assert st1[0] == st2[0] and u1[0] == u2[0] and z1[0] == z2[0]
        assert (s1[0, :N] == s2[0, :N]).all()


# ---- second-generation arithmetic-coder lanes (csrc/scl_aec.cuh) -------------------------------
def _renorm_literal(P, low, high):
    """the reference's two loops (arithmetic_coding.py:126-150), literally"""
    HALF, QTR = 1 << (P - 1), 1 << (P - 2)
    n = m = 0
    while high < HALF or low > HALF:
        if high < HALF:
            low, high = low << 1, high << 1
        else:
            low, high = (low - HALF) << 1, (high - HALF) << 1
        n += 1
    while low > QTR and high < 3 * QTR:
        low, high = (low - QTR) << 1, (high - QTR) << 1
        m += 1
    return n, m, low, high


@pytest.mark.parametrize("P", [32, 16, 8, 5])
def test_aec_closed_form_renormalisation_matches_literal_loops(P):
    import ctypes

    from tests.emu_util import lib

    rng = np.random.default_rng(P)
    FULL = 1 << P
    cases = []
    for _ in range(20000):
        a, b = sorted(int(x) for x in rng.integers(0, FULL + 1, size=2))
        if a == b:
            continue
        cases.append((a, b))
    # boundary-heavy cases: powers of two, all-ones patterns, tiny ranges
    specials = sorted({v for k in range(P + 1) for v in ((1 << k) - 1, 1 << k, (1 << k) + 1, FULL - (1 << k), FULL - (1 << k) - 1, FULL - (1 << k) + 1) if 0 <= v <= FULL})
    for a in specials:
        for b in specials:
            if a < b and a < FULL:
                cases.append((a, b))
    if P <= 8:
        cases = [(a, b) for a in range(FULL) for b in range(a + 1, FULL + 1)]  # exhaustive
    n_ = ctypes.c_uint32()
    m_ = ctypes.c_uint32()
    lo_ = ctypes.c_uint64()
    hi_ = ctypes.c_uint64()
    for low, high in cases:
        n, m, lo2, hi2 = _renorm_literal(P, low, high)
        lib().emu_aec_renorm_counts(P, low, high, ctypes.byref(n_), ctypes.byref(m_), ctypes.byref(lo_), ctypes.byref(hi_))
        assert (n_.value, m_.value, lo_.value, hi_.value) == (n, m, lo2, hi2), (P, low, high)


@pytest.mark.parametrize("c", [c for c in CASES if c["coder"] == "aec"], ids=case_id)
def test_emu_aec2_matches_golden(c):
    coder = EmuCoder(params_from_case(c), None, c["freqs"])
    coder.set_aec2(True)
    n = c["n"]
    model = fresh_model_table(c)[None].copy()
    out, off, ln, st = coder.encode(c["data"].reshape(1, -1), model=model)
    assert st[0] == 0 and int(ln[0]) == c["nbits"]
    assert extract_bits(out, off[0], ln[0]).tobytes() == c["enc"].tobytes()
    assert model[0].tolist() == expected_final_model(c)
    if n == 0:
        return
    packed, total = with_garbage(c["enc"], c["nbits"], c["garbage"])
    lead = 3
    bits = np.concatenate([np.ones(lead, dtype=np.uint8), np.unpackbits(packed)[:total]])
    buf = np.concatenate([np.packbits(bits), np.zeros(16, dtype=np.uint8)])
    model = fresh_model_table(c)[None].copy()
    sym, sizes, used, st = coder.decode(buf, [lead], [total], n, model=model)
    assert st[0] == 0 and int(sizes[0]) == n and int(used[0]) == c["consumed"]
    assert sym[0, :n].tolist() == c["data"].tolist()
    assert model[0].tolist() == expected_final_model(c)


@pytest.mark.parametrize("seed", range(6))
def test_emu_aec2_vs_oracle_random(seed):
    from stanford_compression_library_b200 import _cabi
    from stanford_compression_library_b200._cabi import SclParams

    rng = np.random.default_rng(500 + seed)
    n_sym = int(rng.choice([1, 2, 5, 17, 64, 256]))
    init = [int(x) for x in rng.integers(1, 30, size=n_sym)] if seed % 2 else [1] * n_sym
    B, N = 5, int(rng.integers(1, 600))
    skew = rng.dirichlet(np.ones(n_sym) * 0.3)
    sym = rng.choice(n_sym, size=(B, N), p=skew).astype(np.uint8)
    sizes = rng.integers(1, N + 1, size=B).astype(np.uint32)
    for P, max_total, model in ((32, 1 << 30, _cabi.MODEL_ADAPTIVE_IID), (16, 1 << 14, _cabi.MODEL_ADAPTIVE_IID), (32, sum(init) + 40, _cabi.MODEL_ADAPTIVE_IID),
                                (32, 1 << 30, _cabi.MODEL_FIXED), (12, 1 << 10, _cabi.MODEL_ADAPTIVE_IID)):
        if sum(init) >= (1 << (P - 2)) or max_total <= sum(init) and model == _cabi.MODEL_ADAPTIVE_IID and max_total < n_sym:
            continue
        oracle = so.Oracle.aec(init, PRECISION=P, max_allowed_total_freq=max_total, model=so.MODEL_ADAPTIVE_IID if model == _cabi.MODEL_ADAPTIVE_IID else so.MODEL_FIXED)
        prm = SclParams(coder=_cabi.CODER_AEC, data_block_size_bits=32, num_bits_out=0, range_factor=0, num_state_bits=0, precision=P, model=model,
                        max_allowed_total_freq=max_total)
        coder = EmuCoder(prm, None, init)
        coder.set_aec2(True)
        out, off, ln, st = coder.encode(sym, sizes=sizes)
        ok = st == 0
        for b in range(B):
            try:
                enc, nb = oracle.encode_block(sym[b, : sizes[b]])
            except so.OracleError as e:
                assert st[b] == e.code, (P, max_total, b, st[b], e.code)
                continue
            assert st[b] == 0 and nb == ln[b], (P, max_total, b)
            assert extract_bits(out, off[b], ln[b]).tobytes() == enc.tobytes(), (P, max_total, b)
        dsym, dsz, used, st2 = coder.decode(out, off, ln, N)
        for b in range(B):
            if ok[b]:
                assert st2[b] == 0 and dsz[b] == sizes[b] and used[b] == ln[b], (P, max_total, b)
                assert dsym[b, : sizes[b]].tolist() == sym[b, : sizes[b]].tolist()


# ---- order-k context model (csrc/scl_aec.cuh AecCtxPolicy) ------------------------------------
@pytest.mark.parametrize("n_sym,k", [(1, 2), (2, 0), (2, 1), (2, 7), (3, 3), (4, 2), (7, 1), (16, 1), (39, 1), (256, 0),
                                     (40, 1), (16, 2), (2, 9), (256, 1)])  # second row: tables kept in HBM (AecCtxGlobalPolicy)
def test_emu_order_k_vs_oracle_random(n_sym, k):
    """two consecutive batches through the same model tables (the reference's model object is never
    reset between encode_block calls), every stream and every final table against the oracle"""
    from stanford_compression_library_b200 import _cabi
    from stanford_compression_library_b200._cabi import SclParams

    rng = np.random.default_rng(900 + n_sym * 10 + k)
    B, N = 3, int(rng.integers(20, 500))
    words = n_sym ** (k + 1) + 1
    for P, max_total in ((32, 1 << 30), (14, 1 << 12), (32, 24)):
        oracle = so.Oracle.aec([1] * n_sym, PRECISION=P, model=so.MODEL_ORDER_K, k=k, max_allowed_total_freq=max_total)
        prm = SclParams(coder=_cabi.CODER_AEC, data_block_size_bits=32, num_bits_out=0, range_factor=0, num_state_bits=0, precision=P,
                        model=_cabi.MODEL_ORDER_K, model_order=k, max_allowed_total_freq=max_total)
        coder = EmuCoder(prm, None, [1] * n_sym)
        fresh = np.array([1] * (words - 1) + [0], dtype=np.uint64)
        m_enc, m_dec = np.tile(fresh, (B, 1)), np.tile(fresh, (B, 1))
        m_ref_e, m_ref_d = np.tile(fresh, (B, 1)), np.tile(fresh, (B, 1))
        alive = np.ones(B, dtype=bool)
        for rnd in range(2):
            # a sticky source so that contexts matter: repeat the previous symbol with probability 0.6
            sym = rng.integers(0, n_sym, size=(B, N)).astype(np.uint8)
            keep = rng.random((B, N)) < 0.6
            for j in range(1, N):
                sym[:, j] = np.where(keep[:, j], sym[:, j - 1], sym[:, j])
            sizes = rng.integers(1, N + 1, size=B).astype(np.uint32)
            out, off, ln, st = coder.encode(sym, sizes=sizes, model=m_enc)
            dsym, dsz, used, st2 = coder.decode(out, off, ln, N, model=m_dec)
            for b in range(B):
                if not alive[b]:
                    continue
                try:
                    enc, nb = oracle.encode_block(sym[b, : sizes[b]], model_freq=m_ref_e[b])
                except so.OracleError as e:
                    assert st[b] == e.code == _cabi.ST_TOTAL_FREQ, (P, max_total, b, st[b], e.code)
                    alive[b] = False  # the reference has raised: its model object is in no defined state
                    continue
                assert st[b] == 0 and nb == ln[b], (P, max_total, rnd, b)
                assert extract_bits(out, off[b], ln[b]).tobytes() == enc.tobytes()
                assert m_enc[b].tolist() == m_ref_e[b].tolist()
                ref_sym, ref_used = oracle.decode_block(enc, nb, model_freq=m_ref_d[b], cap=N)
                assert st2[b] == 0 and dsz[b] == sizes[b] and used[b] == ref_used
                assert dsym[b, : sizes[b]].tolist() == sym[b, : sizes[b]].tolist() == ref_sym.tolist()
                assert m_dec[b].tolist() == m_ref_d[b].tolist()
        if max_total == 24 and n_sym <= 4 and k <= 2 and N > 200:
            assert not alive.all()  # the small limit is meant to trip here


def test_emu_order_k_table_size_limit():
    from stanford_compression_library_b200 import _cabi
    from stanford_compression_library_b200._cabi import SclParams

    def make(n_sym, k):
        prm = SclParams(coder=_cabi.CODER_AEC, data_block_size_bits=32, num_bits_out=0, range_factor=0, num_state_bits=0, precision=32,
                        model=_cabi.MODEL_ORDER_K, model_order=k, max_allowed_total_freq=1 << 30)
        return EmuCoder(prm, None, [1] * n_sym)

    make(3, 5)  # 243 * 4 = 972 words: shared memory
    make(39, 1)  # 39 * 40 = 1560 words: shared memory
    for n_sym, k in ((40, 1), (256, 1), (2, 9), (16, 2), (22, 2)):  # table in HBM, <= 512 rows of totals in shared memory
        make(n_sym, k)
    for n_sym, k in ((3, 6), (2, 10), (256, 2), (23, 2)):  # more than 512 contexts
        with pytest.raises(NotImplementedError):
            make(n_sym, k)


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
def test_orderk_large_tables_match_reference_golden(c):
    """AdaptiveOrderKFreqModel with tables too large for shared memory (a BYTE alphabet at k = 1 among them), vectors
    from the unmodified reference (oracle/gen_golden_orderk_large.py): the oracle and the product's lane code
    (AecCtxGlobalPolicy, working in place on the model table) reproduce the reference's bits, its final count table
    and context, and its num_bits_consumed on stream + garbage."""
    from stanford_compression_library_b200 import _cabi
    from stanford_compression_library_b200._cabi import SclParams

    n_sym, k, p = len(c["freqs"]), c["model"]["k"], c["params"]
    want_final = c["final"].astype(np.uint64).tolist() + [c["model"]["final_ctx"]]
    fresh = np.array([1] * (n_sym ** (k + 1)) + [0], dtype=np.uint64)
    oracle = so.Oracle.aec([1] * n_sym, DATA_BLOCK_SIZE_BITS=p["DATA_BLOCK_SIZE_BITS"], PRECISION=p["PRECISION"], model=so.MODEL_ORDER_K, k=k,
                           max_allowed_total_freq=c["model"]["max_total"])
    m = fresh.copy()
    ref_bytes, ref_bits = oracle.encode_block(c["data"], model_freq=m)
    assert ref_bits == c["nbits"] and ref_bytes.tobytes() == c["enc"].tobytes()
    assert m.tolist() == want_final
    prm = SclParams(coder=_cabi.CODER_AEC, data_block_size_bits=p["DATA_BLOCK_SIZE_BITS"], num_bits_out=0, range_factor=0, num_state_bits=0,
                    precision=p["PRECISION"], model=_cabi.MODEL_ORDER_K, model_order=k, max_allowed_total_freq=c["model"]["max_total"])
    coder = EmuCoder(prm, None, [1] * n_sym)
    m = fresh.copy()[None]
    out, off, ln, st = coder.encode(c["data"].reshape(1, -1), model=m)
    assert st[0] == 0 and int(ln[0]) == c["nbits"]
    assert extract_bits(out, off[0], ln[0]).tobytes() == c["enc"].tobytes()
    assert m[0].tolist() == want_final
    packed, total = with_garbage(c["enc"], c["nbits"], c["garbage"])
    bits = np.concatenate([np.ones(3, dtype=np.uint8), np.unpackbits(packed)[:total]])
    buf = np.concatenate([np.packbits(bits), np.zeros(16, dtype=np.uint8)])
    m = fresh.copy()[None]
    sym, sizes, used, st = coder.decode(buf, [3], [total], c["n"], model=m)
    assert st[0] == 0 and int(sizes[0]) == c["n"] and sym[0, : c["n"]].tolist() == c["data"].tolist()
    assert int(used[0]) == c["consumed"] and m[0].tolist() == want_final


def test_generic_rans_decode_is_bounded_on_a_zero_state():
    """A malformed stream whose state field is 0 (or decodes to 0) can never reach L by shifting in the
    zero bits the reader returns past the end: the reference raises ValueError from bitarray_to_uint on
    an empty slice (rANS.py:256); the generic lane must stop with TRUNCATED instead of spinning."""
    from stanford_compression_library_b200._cabi import CODER_RANS, SclParams

    freq = np.array([3, 3, 2], dtype=np.uint64)
    p = SclParams(coder=CODER_RANS, data_block_size_bits=32, num_bits_out=1, range_factor=1 << 16, num_state_bits=so.ref_get_bit_width((8 << 16) * 2 - 1))
    coder = EmuCoder(p, None, freq)
    coder.force_generic()
    buf = np.zeros(64, dtype=np.uint8)
    buf[3] = 1  # 32-bit size header = 1, every other bit 0 => state 0
    sym, sizes, used, st = coder.decode(buf, [0], [8 * 48], 8)
    assert st[0] == 4  # SCL_ST_TRUNCATED
    # and with no bit_len given the bound is the end of the buffer
    sym, sizes, used, st = coder.decode(buf, [0], None, 8)
    assert st[0] == 4


# ---- arithmetic coder, 8-bit counters (csrc/scl_aec.cuh AecModel8 / AecIid8Policy) --------------------------------
@pytest.mark.parametrize("c", [c for c in CASES if c["coder"] == "aec" and c["model"]["kind"] != "order_k"], ids=case_id)
def test_emu_aec_model8_matches_golden(c):
    """the 8-bit-counter model on every golden case it is eligible for (else the call falls back to 16 bits): same
    bits and bits consumed as the reference"""
    coder = EmuCoder(params_from_case(c), None, c["freqs"])
    coder.set_aec2(2)
    n = c["n"]
    out, off, ln, st = coder.encode(c["data"].reshape(1, -1))  # no model table: every block from the creation-time counts
    assert st[0] == 0 and int(ln[0]) == c["nbits"]
    assert extract_bits(out, off[0], ln[0]).tobytes() == c["enc"].tobytes()
    if n == 0:
        return
    packed, total = with_garbage(c["enc"], c["nbits"], c["garbage"])
    bits = np.concatenate([np.ones(3, dtype=np.uint8), np.unpackbits(packed)[:total]])
    buf = np.concatenate([np.packbits(bits), np.zeros(16, dtype=np.uint8)])
    sym, sizes, used, st = coder.decode(buf, [3], [total], n)
    assert st[0] == 0 and int(sizes[0]) == n and int(used[0]) == c["consumed"]
    assert sym[0, :n].tolist() == c["data"].tolist()


@pytest.mark.parametrize("seed", range(8))
def test_emu_aec_model8_vs_oracle_random(seed):
    """random alphabets / initial counts / block lengths up to the eligibility bound, skewed sources that push one or
    several counters past 255 (the `big` list), small max_allowed_total_freq values that fire the halving rule with
    big symbols present, the fixed model; every stream against the oracle, every block decoded back"""
    from stanford_compression_library_b200 import _cabi
    from stanford_compression_library_b200._cabi import SclParams

    rng = np.random.default_rng(7000 + seed)
    n_sym = int(rng.choice([1, 2, 3, 17, 100, 256]))
    init = [int(x) for x in rng.integers(1, 4, size=n_sym)] if seed % 2 else [1] * n_sym
    B = 6
    N = int(min(1270, 5 * 256 - sum(init) - 1)) if seed % 3 else int(rng.integers(1, 700))
    N = max(N, 1)
    # rows: very skewed (one symbol nearly always), two heavy symbols, flat
    sym = np.zeros((B, N), dtype=np.uint8)
    for b in range(B):
        k = b % 3
        if k == 0:
            p = np.full(n_sym, 0.02 / max(1, n_sym - 1))
            p[rng.integers(0, n_sym)] = 0.98 if n_sym > 1 else 1.0
        elif k == 1 and n_sym >= 2:
            p = np.full(n_sym, 0.04 / max(1, n_sym - 2)) if n_sym > 2 else np.zeros(n_sym)
            i, j = rng.choice(n_sym, size=2, replace=False)
            p[i], p[j] = 0.5, 0.46 if n_sym > 2 else 0.5
        else:
            p = np.ones(n_sym)
        sym[b] = rng.choice(n_sym, size=N, p=p / p.sum())
    sizes = np.array([N, N, N, max(1, N // 2), 1, N], dtype=np.uint32)
    for P, max_total, model in ((32, 1 << 30, _cabi.MODEL_ADAPTIVE_IID), (32, sum(init) + 300, _cabi.MODEL_ADAPTIVE_IID), (16, 1 << 14, _cabi.MODEL_ADAPTIVE_IID),
                                (32, 700, _cabi.MODEL_ADAPTIVE_IID), (32, 1 << 30, _cabi.MODEL_FIXED)):
        if sum(init) >= (1 << (P - 2)) or (model == _cabi.MODEL_ADAPTIVE_IID and max_total <= sum(init)):
            continue
        oracle = so.Oracle.aec(init, PRECISION=P, max_allowed_total_freq=max_total, model=so.MODEL_ADAPTIVE_IID if model == _cabi.MODEL_ADAPTIVE_IID else so.MODEL_FIXED)
        prm = SclParams(coder=_cabi.CODER_AEC, data_block_size_bits=32, num_bits_out=0, range_factor=0, num_state_bits=0, precision=P, model=model,
                        max_allowed_total_freq=max_total)
        coder = EmuCoder(prm, None, init)
        assert coder.aec_model8_ok(N)
        coder.set_aec2(2)
        out, off, ln, st = coder.encode(sym, sizes=sizes)
        ok = st == 0
        for b in range(B):
            try:
                enc, nb = oracle.encode_block(sym[b, : sizes[b]])
            except so.OracleError as e:
                assert st[b] == e.code, (P, max_total, b, st[b], e.code)
                continue
            assert st[b] == 0 and nb == ln[b], (P, max_total, b)
            assert extract_bits(out, off[b], ln[b]).tobytes() == enc.tobytes(), (P, max_total, b)
        dsym, dsz, used, st2 = coder.decode(out, off, ln, N)
        for b in range(B):
            if ok[b]:
                assert st2[b] == 0 and dsz[b] == sizes[b], (P, max_total, b)
                assert dsym[b, : sizes[b]].tolist() == sym[b, : sizes[b]].tolist()


def test_emu_aec_model8_eligibility():
    from stanford_compression_library_b200 import _cabi
    from stanford_compression_library_b200._cabi import SclParams

    def mk(init, model=_cabi.MODEL_ADAPTIVE_IID):
        prm = SclParams(coder=_cabi.CODER_AEC, data_block_size_bits=32, num_bits_out=0, range_factor=0, num_state_bits=0, precision=32, model=model,
                        max_allowed_total_freq=1 << 30)
        return EmuCoder(prm, None, init)

    c = mk([1] * 256)
    assert c.aec_model8_ok(1024) and c.aec_model8_ok(1279) and not c.aec_model8_ok(1280)  # (256 + n) // 256 <= 5
    assert not mk([1] * 255 + [256]).aec_model8_ok(10)  # an initial count does not fit a byte
    assert mk([255] * 256, _cabi.MODEL_FIXED).aec_model8_ok(1 << 20)  # a fixed model never grows
    assert not mk([255] * 256).aec_model8_ok(1)
