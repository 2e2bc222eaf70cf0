"""rANS as the entropy stage of LZ77 streams: the container classes against
tests/golden/lz77_rans_v1.npz, which oracle/gen_golden_lz77.py composed from the unmodified
reference (Elias delta + rANS + LogScaleBinnedIntegerEncoder + LZ77StreamsEncoder + the LZ77 parser).

CPU tier: the host logic (Elias delta, counts container, log-scale binning, stream concatenation)
with the C oracle standing in for the rANS stage.  GPU tier: the shipped classes, rANS on the device.
"""
import json
import os

import numpy as np
import pytest

from oracle import scl_oracle as so
from stanford_compression_library_b200 import BitArray, DataBlock
from stanford_compression_library_b200.compressors import lz77_entropy as lz

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lz77_rans_v1.npz")


def load_cases():
    z = np.load(GOLDEN)
    meta = json.loads(bytes(z["meta"]).decode())
    out = []
    for c in meta["cases"]:
        c = dict(c)
        for k in z.files:
            pre = "c%d_" % c["id"]
            if k.startswith(pre):
                c[k[len(pre):]] = z[k]
        c["expected"] = BitArray.from_packed(c["enc"], c["nbits"])
        out.append(c)
    return out


CASES = load_cases()
ids = ["%02d-%s-%s" % (c["id"], c["kind"], c["note"][:30].replace(" ", "_")) for c in CASES]


class _OracleRansEncoder:
    """oracle-backed stand-in for the device rANSEncoder (CPU tier only)"""

    def __init__(self, params):
        self.p = params

    def encode_block(self, data_block):
        f = self.p.freqs
        idx = {s: i for i, s in enumerate(f.alphabet)}
        sym = np.array([idx[s] for s in data_block.data_list], dtype=np.uint8)
        packed, nbits = so.Oracle.rans([int(x) for x in f.freq_list]).encode_block(sym)
        return BitArray.from_packed(packed, nbits)


class _OracleRansDecoder:
    def __init__(self, params):
        self.p = params

    def decode_block(self, bitarray):
        f = self.p.freqs
        sym, used = so.Oracle.rans([int(x) for x in f.freq_list]).decode_block(bitarray.to_packed(), len(bitarray))
        return DataBlock([f.alphabet[int(i)] for i in sym]), used


def _run_case(c):
    """encode with the module's classes, compare with the reference-composed stream, decode back (with trailing garbage)"""
    garbage = BitArray("1011001")
    if c["kind"] == "elias":
        vals = c["vals"].tolist()
        got = lz.EliasDeltaUintEncoder().encode_block(DataBlock(vals))
        assert got == c["expected"]
        dec, used = lz.EliasDeltaUintDecoder().decode_block(got)
        assert dec.data_list == vals and used == len(got)
    elif c["kind"] == "empirical":
        vals = c["vals"].tolist() if "vals" in c else []
        got = lz.EmpiricalIntRansEncoder(c["alphabet_size"]).encode_block(DataBlock(vals))
        assert got == c["expected"]
        dec, used = lz.EmpiricalIntRansDecoder(c["alphabet_size"]).decode_block(got + garbage)
        assert list(dec.data_list) == vals and used == len(got)
    elif c["kind"] == "logbin":
        vals = c["vals"].tolist()
        got = lz.LogScaleBinnedIntegerEncoder(offset=c["offset"]).encode_block(DataBlock(vals))
        assert got == c["expected"]
        dec, used = lz.LogScaleBinnedIntegerDecoder(offset=c["offset"]).decode_block(got + garbage)
        assert list(dec.data_list) == vals and used == len(got)
    else:
        seqs = [lz.LZ77Sequence(int(a), int(b), int(o)) for a, b, o in zip(c["literal_counts"], c["match_lengths"], c["match_offsets"])]
        literals = c["literals"].tolist()
        got = lz.LZ77StreamsEncoder().encode_block(seqs, literals)
        assert got == c["expected"]
        (dseqs, dlits), used = lz.LZ77StreamsDecoder().decode_block(got + garbage)
        assert dseqs == seqs and list(dlits) == literals and used == len(got)


@pytest.mark.parametrize("c", CASES, ids=ids)
def test_host_logic_with_oracle_rans_stage(c, monkeypatch):
    so.build()
    monkeypatch.setattr(lz, "rANSEncoder", _OracleRansEncoder)
    monkeypatch.setattr(lz, "rANSDecoder", _OracleRansDecoder)
    _run_case(c)


def test_log_binning_rejects_too_large_values():
    with pytest.raises(ValueError):
        lz.LogScaleBinnedIntegerEncoder(offset=0, max_num_bins=4).encode_block(DataBlock([100]))


@pytest.mark.gpu
@pytest.mark.parametrize("c", CASES, ids=ids)
def test_device_rans_stage_matches_reference_composition(c):
    _run_case(c)


@pytest.mark.parametrize("seed", range(5))
def test_streams_round_trip_random_sequences_with_oracle_stage(seed, monkeypatch):
    """random LZ77 sequences (small and huge offsets, zero counts, empty literal lists) through the streams coder
    and back, the oracle standing in for the rANS stage"""
    so.build()
    monkeypatch.setattr(lz, "rANSEncoder", _OracleRansEncoder)
    monkeypatch.setattr(lz, "rANSDecoder", _OracleRansDecoder)
    rng = np.random.default_rng(seed)
    n = int(rng.integers(0, 60))
    seqs = [lz.LZ77Sequence(int(rng.integers(0, 40)), int(rng.integers(0, 300)), int(rng.integers(0, 1 << int(rng.integers(1, 31))))) for _ in range(n)]
    literals = [int(x) for x in rng.integers(0, 256, size=int(rng.integers(0, 200)))]
    bits = lz.LZ77StreamsEncoder().encode_block(seqs, literals)
    (dseqs, dlits), used = lz.LZ77StreamsDecoder().decode_block(bits + BitArray("110"))
    assert dseqs == seqs and list(dlits) == literals and used == len(bits)
