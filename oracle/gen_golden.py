"""Generate tests/golden/golden_v1.npz from the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python oracle/gen_golden.py

Every case records: coder, parameters, the Frequencies table in dict order, the data as
alphabet indices, and what the reference produced -- `encode_block(...).tobytes()`, its bit
length, and `decode_block` of (stream + seeded garbage bits) -> (symbols, num_bits_consumed).
The grids are the reference's own test grids (rANS.py:363-401, tANS.py:418-452,
arithmetic_coding.py:297-381, range_coder.py:320-374) plus the BASELINE.json configs at
fixture-friendly sizes.  Deterministic: fixed seeds, no wall-clock input.
"""
import copy
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_loader import import_reference  # noqa: E402

import_reference()
from scl.compressors.arithmetic_coding import AECParams, ArithmeticDecoder, ArithmeticEncoder  # noqa: E402
from scl.compressors.probability_models import AdaptiveIIDFreqModel, AdaptiveOrderKFreqModel, FixedFreqModel  # noqa: E402
from scl.compressors.rANS import rANSDecoder, rANSEncoder, rANSParams  # noqa: E402
from scl.compressors.range_coder import RangeCoderParams, RangeDecoder, RangeEncoder  # noqa: E402
from scl.compressors.tANS import tANSDecoder, tANSEncoder, tANSParams  # noqa: E402
from scl.core.data_block import DataBlock  # noqa: E402
from scl.core.prob_dist import Frequencies  # noqa: E402
from scl.utils.bitarray_utils import BitArray  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "golden_v1.npz")


def zipf_freqs(n_sym=256, M=4096, s=1.0):
    """SURVEY.md 8(d) quantiser: f_b = max(1, floor(p_b*M)), deficit added to f_0."""
    p = 1.0 / np.arange(1, n_sym + 1, dtype=np.float64) ** s
    p /= p.sum()
    f = np.maximum(1, np.floor(p * M).astype(np.int64))
    f[0] += M - f.sum()
    assert f.sum() == M and f.min() >= 1
    return [int(x) for x in f]


def draw(freqs, n, seed):
    f = np.asarray(freqs, dtype=np.float64)
    rng = np.random.default_rng(seed)
    return rng.choice(len(freqs), size=n, p=f / f.sum()).astype(np.uint8)


def garbage(seed, max_bits=100):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(0, max_bits))
    return "".join("1" if b else "0" for b in rng.integers(0, 2, size=n))


cases = []
arrays = {}


def record(coder, params, freqs, data, enc_bits, dec_syms, dec_consumed, extra_garbage, note="", model=None, blocks=None):
    i = len(cases)
    meta = dict(id=i, coder=coder, params=params, freqs=[int(x) for x in freqs], n=int(len(data)), nbits=int(len(enc_bits)),
                consumed=int(dec_consumed), garbage=extra_garbage, note=note)
    if model is not None:
        meta["model"] = model
    cases.append(meta)
    arrays["c%d_data" % i] = np.asarray(data, dtype=np.uint8)
    arrays["c%d_enc" % i] = np.frombuffer(enc_bits.tobytes(), dtype=np.uint8).copy()
    assert list(dec_syms) == [int(x) for x in data], "reference round trip failed?!"


def F(freqs):
    return Frequencies({i: int(f) for i, f in enumerate(freqs)})


def run_rans(freqs, data, note="", tans=False, **kw):
    P, E, D = (tANSParams, tANSEncoder, tANSDecoder) if tans else (rANSParams, rANSEncoder, rANSDecoder)
    params = P(F(freqs), **kw)
    enc = E(params).encode_block(DataBlock([int(x) for x in data]))
    g = garbage(len(cases))
    dec, used = D(params).decode_block(enc + BitArray(g))
    assert used == len(enc)
    pd = dict(DATA_BLOCK_SIZE_BITS=params.DATA_BLOCK_SIZE_BITS, NUM_BITS_OUT=params.NUM_BITS_OUT, RANGE_FACTOR=params.RANGE_FACTOR,
              NUM_STATE_BITS=int(params.NUM_STATE_BITS))
    record("tans" if tans else "rans", pd, freqs, data, enc, dec.data_list, used, g, note)


def run_range(freqs, data, note="", **kw):
    params = RangeCoderParams(**kw)
    fr = F(freqs)
    enc = RangeEncoder(params, fr).encode_block(DataBlock([int(x) for x in data]))
    g = garbage(len(cases))
    dec, used = RangeDecoder(params, fr).decode_block(enc + BitArray(g))
    assert used == len(enc)
    record("range", dict(DATA_BLOCK_SIZE_BITS=params.DATA_BLOCK_SIZE_BITS, PRECISION=params.PRECISION), freqs, data, enc, dec.data_list, used, g, note)


def run_aec(freqs_initial, data, model="adaptive_iid", note="", max_total=None, **kw):
    params = AECParams(**kw)
    mt = params.MAX_ALLOWED_TOTAL_FREQ if max_total is None else max_total
    cls = AdaptiveIIDFreqModel if model == "adaptive_iid" else FixedFreqModel
    m_enc = cls(F(freqs_initial), mt)
    m_dec = copy.deepcopy(m_enc)
    enc = ArithmeticEncoder(params, m_enc).encode_block(DataBlock([int(x) for x in data]))
    g = garbage(len(cases))
    dec, used = ArithmeticDecoder(params, m_dec).decode_block(enc + BitArray(g))
    assert used == len(enc), (used, len(enc))
    final = [int(m_enc.freqs_current.freq_dict[i]) for i in range(len(freqs_initial))]
    record("aec", dict(DATA_BLOCK_SIZE_BITS=params.DATA_BLOCK_SIZE_BITS, PRECISION=params.PRECISION), freqs_initial, data, enc,
           dec.data_list, used, g, note, model=dict(kind=model, max_total=int(mt), final_freqs=final))


def run_aec_order_k(n_alpha, k, data, note="", **kw):
    """AdaptiveOrderKFreqModel (probability_models.py:95-160) over alphabet 0..n_alpha-1."""
    params = AECParams(**kw)
    m_enc = AdaptiveOrderKFreqModel(list(range(n_alpha)), k, params.MAX_ALLOWED_TOTAL_FREQ)
    m_dec = copy.deepcopy(m_enc)
    enc = ArithmeticEncoder(params, m_enc).encode_block(DataBlock([int(x) for x in data]))
    g = garbage(len(cases))
    dec, used = ArithmeticDecoder(params, m_dec).decode_block(enc + BitArray(g))
    final = [int(x) for x in np.ravel(m_enc.freqs_kplus1_tuple)]
    ctx = 0
    for p in m_enc.past_k:
        ctx = ctx * n_alpha + int(p)
    record("aec", dict(DATA_BLOCK_SIZE_BITS=params.DATA_BLOCK_SIZE_BITS, PRECISION=params.PRECISION), [1] * n_alpha, data, enc, dec.data_list, used, g, note,
           model=dict(kind="order_k", k=k, max_total=int(params.MAX_ALLOWED_TOTAL_FREQ), final_freqs=final, final_ctx=ctx))


def markov2(n, seed):
    """the reference's 2nd-order Markov test source (arithmetic_coding.py:384-402)"""
    rng = np.random.default_rng(seed)
    bits = rng.choice(2, size=n - 2)
    x = np.zeros(n, dtype=np.uint8)
    x[0], x[1] = rng.choice(3), rng.choice(3)
    for i in range(2, n):
        x[i] = (x[i - 1] + x[i - 2] + bits[i - 2]) % 3
    return x


def main():
    # ---- the reference's literal known-answer vector (rANS.py:303-360 / tANS.py:340-415) ----
    run_rans([3, 3, 2], [0, 2, 1], note="KAT rANS.py:303-360 expects 00011 1011 10 01 0", DATA_BLOCK_SIZE_BITS=5, NUM_BITS_OUT=1, RANGE_FACTOR=1)
    assert arrays["c0_enc"].tobytes() == BitArray("00011101110010").tobytes()
    run_rans([3, 3, 2], [0, 2, 1], tans=True, note="KAT tANS.py:340-415", DATA_BLOCK_SIZE_BITS=5, NUM_BITS_OUT=1, RANGE_FACTOR=1)

    # ---- rANS grid (rANS.py:363-379) ----
    grid = [
        ([1, 1, 2], {}),
        ([12, 34, 1, 45], {}),
        ([34, 35, 546, 1, 13, 245], dict(NUM_BITS_OUT=8)),
        ([5, 5, 5, 5, 5, 5], dict(RANGE_FACTOR=1 << 12)),
        ([1, 3], dict(RANGE_FACTOR=1 << 4)),
    ]
    for k, (fr, kw) in enumerate(grid):
        run_rans(fr, draw(fr, 400, seed=k), note="rANS.py:363-379 grid", **kw)
    # extra parameter corners: empty, single symbol, NBO in {2,3,8}, RF=1, 64-bit state
    run_rans([1, 1, 2], [], note="empty block")
    run_rans([7], np.zeros(50, dtype=np.uint8), note="single-symbol alphabet")
    run_rans([1, 1, 2], draw([1, 1, 2], 200, 11), note="NBO=2", NUM_BITS_OUT=2, RANGE_FACTOR=1)
    run_rans([12, 34, 1, 45], draw([12, 34, 1, 45], 200, 12), note="NBO=3 RF=7", NUM_BITS_OUT=3, RANGE_FACTOR=7)
    run_rans([34, 35, 546, 1, 13, 245], draw([34, 35, 546, 1, 13, 245], 300, 13), note="H ~ 2^46: 64-bit state", NUM_BITS_OUT=16, RANGE_FACTOR=1 << 20)
    run_rans([1, 65535], draw([1, 65535], 300, 14), note="extreme skew", NUM_BITS_OUT=8, RANGE_FACTOR=1 << 8)

    # ---- BASELINE cfg1: 4 KiB uniform-random bytes, 256 symbols ----
    u = np.random.default_rng(0).integers(0, 256, 4096).astype(np.uint8)
    run_rans([16] * 256, u, note="cfg1 uniform table default params")
    run_rans([16] * 256, u, note="cfg1 uniform table NBO=8 RF=2^12", NUM_BITS_OUT=8, RANGE_FACTOR=1 << 12)
    cnt = (np.bincount(u, minlength=256) + 1).tolist()
    run_rans(cnt, u[:1024], note="cfg1 counts+1 (M not a power of two) default params")
    run_rans(cnt, u[:1024], note="cfg1 counts+1 NBO=8 RF=2^12", NUM_BITS_OUT=8, RANGE_FACTOR=1 << 12)

    # ---- BASELINE cfg2/cfg3 table: Zipf-1.0 over 256 symbols, M=4096 ----
    zf = zipf_freqs()
    z = draw(zf, 2048, seed=0)
    run_rans(zf, z, note="cfg2 zipf default params")
    run_rans(zf, z, note="cfg2 zipf NBO=8 RF=2^12", NUM_BITS_OUT=8, RANGE_FACTOR=1 << 12)
    run_rans(zf, z[:777], note="cfg2 zipf ragged length", NUM_BITS_OUT=8, RANGE_FACTOR=1 << 12)
    run_rans(zf, z[:1024], tans=True, note="cfg3 zipf tANS RF=1", RANGE_FACTOR=1)
    run_rans(zf, z[:512], tans=True, note="cfg3 zipf tANS RF=4", RANGE_FACTOR=4)

    # ---- tANS grid (tANS.py:418-433) ----
    tgrid = [([1, 1, 2], dict(RANGE_FACTOR=1)), ([1, 3], dict(RANGE_FACTOR=1 << 4)), ([3, 4, 9], dict(RANGE_FACTOR=1 << 8))]
    for k, (fr, kw) in enumerate(tgrid):
        run_rans(fr, draw(fr, 400, seed=20 + k), tans=True, note="tANS.py:418-433 grid (RF reduced for the last entry)", **kw)
    run_rans([1, 1, 2], [], tans=True, note="tANS empty block", RANGE_FACTOR=1)

    # ---- range coder (range_coder.py:320-374) ----
    rgrid = [[1, 1, 2], [12, 34, 1, 45], [34, 35, 546, 1, 13, 245], [1, 65534]]
    for k, fr in enumerate(rgrid):
        run_range(fr, draw(fr, 500, seed=30 + k), note="range_coder.py:333-349 grid")
    run_range([1, 65535], np.array([0, 1] * 300, dtype=np.uint8), note="edge: A,C alternating (:351-355)")
    run_range([1, 1, 65534], np.array([0, 1, 2] * 200, dtype=np.uint8), note="edge (:356-359)")
    run_range([1, 1, 65534], np.zeros(400, dtype=np.uint8), note="edge all-A (:360-363)")
    run_range([1, 1, 65534], np.full(400, 2, dtype=np.uint8), note="edge all-C (:364-367)")
    d = draw([12, 34, 1, 45], 50, seed=0)
    for l in (0, 1, 2, 3, 7, 49):
        run_range([12, 34, 1, 45], d[:l], note="lengths (:369-374)")
    run_range(zf, z[:1500], note="zipf 256 symbols")
    run_range([12, 34, 1, 45], draw([12, 34, 1, 45], 300, 39), note="DBSB=12", DATA_BLOCK_SIZE_BITS=12)

    # ---- arithmetic coder (arithmetic_coding.py:297-381) ----
    agrid = [
        ([1, 1, 2], {}),
        ([12, 34, 1, 45], {}),
        ([34, 35, 546, 1, 13, 245], dict(DATA_BLOCK_SIZE_BITS=12)),
        ([5, 5, 5, 5, 5, 5], dict(DATA_BLOCK_SIZE_BITS=12, PRECISION=16)),
    ]
    for k, (fr, kw) in enumerate(agrid):
        data = draw(fr, 300, seed=40 + k)
        run_aec(fr, data, note="test_arithmetic_coding: adaptive model initialised with the data freqs", **kw)
        run_aec([1] * len(fr), data, note="test_adaptive_arithmetic_coding: uniform init", **kw)
        run_aec(fr, data, model="fixed", note="FixedFreqModel", **kw)
    run_aec([1, 1, 2], draw([1, 1, 2], 400, 50), note="halving rule fires (max_total=64)", max_total=64)
    run_aec([1] * 256, z[:1024], note="cfg4: 1 KiB zipf block, uniform init over 256 symbols")
    run_aec([1] * 256, z[:1], note="single symbol block")
    run_aec([1] * 256, [], note="AEC empty block: encode only (reference decoder does not terminate)") if False else None

    # ---- order-k adaptive context model (arithmetic_coding.py:404-447, probability_models.py:95-160) ----
    mk = markov2(400, seed=0)
    for k in (0, 1, 2, 3):
        run_aec_order_k(3, k, mk, note="order-%d model on the reference's 2nd-order Markov source" % k, DATA_BLOCK_SIZE_BITS=12)
    run_aec_order_k(2, 4, draw([3, 1], 300, 70), note="binary alphabet, order 4", DATA_BLOCK_SIZE_BITS=12)
    run_aec_order_k(4, 2, draw([5, 1, 1, 3], 300, 71), note="4 symbols, order 2", DATA_BLOCK_SIZE_BITS=12, PRECISION=16)
    run_aec_order_k(5, 3, draw([4, 3, 2, 1, 1], 600, 72), note="5 symbols, order 3 (750 counters)", DATA_BLOCK_SIZE_BITS=12)

    meta = json.dumps(dict(version=1, reference_commit="5e9a699db81d7452cdf4f34b5b7023bac39f5dd5", cases=cases))
    np.savez_compressed(OUT, meta=np.frombuffer(meta.encode(), dtype=np.uint8), **arrays)
    print("wrote %s: %d cases, %d bytes" % (OUT, len(cases), os.path.getsize(OUT)))


if __name__ == "__main__":
    main()
