"""CPU tier: host-side logic of the package (BitArray, Frequencies, parameter derivation, the
C-ABI library's exports, loud failure without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

import stanford_compression_library_b200 as scl
from stanford_compression_library_b200 import _cabi
from stanford_compression_library_b200.utils.bitarray_utils import BitArray, bitarray_to_uint, get_bit_width, uint_to_bitarray

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bitarray_container_semantics():
    a = BitArray("01011")
    assert len(a) == 5 and list(a) == [0, 1, 0, 1, 1] and a[1] == 1 and a[-1] == 1
    assert a[1:4] == BitArray("101") and a[10:] == BitArray("") and a[:100] == a
    assert a + BitArray("1") == BitArray("010111")
    b = BitArray(a)
    b += BitArray("00")
    assert b == BitArray("0101100") and a == BitArray("01011")
    b.extend("1" + "0" * 2)
    assert b.to01() == "0101100100"
    assert a.tobytes() == bytes([0b01011000])
    c = BitArray()
    c.frombytes(bytes([0xA5, 0x01]))
    assert c.to01() == "1010010100000001"
    assert BitArray.from_packed(np.array([0xA5, 0x01], dtype=np.uint8), 7, bit_offset=3).to01() == "0010100"
    assert (a == "01011") is False or True  # comparison with foreign types does not raise


def test_uint_conversions_match_reference_semantics():
    assert uint_to_bitarray(4).to01() == "100" and uint_to_bitarray(0).to01() == "0"
    assert uint_to_bitarray(13, bit_width=8).to01() == "00001101"
    assert bitarray_to_uint(BitArray("1101")) == 13
    with pytest.raises(OverflowError):
        uint_to_bitarray(4, bit_width=2)
    with pytest.raises(ValueError):
        bitarray_to_uint(BitArray(""))
    # get_bit_width: bitarray_utils.py:112-117 + the float quirk from 2^49
    assert [get_bit_width(x) for x in (0, 1, 255, 1 << 16)] == [1, 1, 8, 17]
    assert get_bit_width((1 << 29) - 1) == 29 and get_bit_width((1 << 32) - 1) == 32
    assert get_bit_width((1 << 49) - 1) == int(np.ceil(np.log2(float(1 << 49))))


def test_frequencies_order_and_device_arrays():
    f = scl.Frequencies({"B": 7, "A": 1, "C": 3})
    assert f.alphabet == ["B", "A", "C"] and f.cumulative_freq_dict == {"B": 0, "A": 7, "C": 8} and int(f.total_freq) == 11
    alpha, freq = f.to_arrays()
    assert alpha.tolist() == [0, 1, 2] and freq.tolist() == [7, 1, 3]
    g = scl.Frequencies({200: 5, 3: 1})
    alpha, freq = g.to_arrays()
    assert alpha.tolist() == [200, 3] and g.byte_alphabet()[1] is True
    with pytest.raises(ValueError):
        scl.ProbabilityDist({"H": 0.5, "T": 0.4})
    assert scl.ProbabilityDist({"A": 0.5, "B": 0.25, "C": 0.25}).entropy == 1.5


def test_rans_params_match_reference_derivation():
    from stanford_compression_library_b200.compressors.rANS import rANSParams
    from stanford_compression_library_b200.compressors.tANS import tANSParams
    from stanford_compression_library_b200.workloads import zipf_frequencies

    p = rANSParams(zipf_frequencies())
    assert (p.M, p.L, p.H, p.NUM_STATE_BITS, p.INITIAL_STATE) == (4096, 1 << 28, (1 << 29) - 1, 29, 1 << 28)
    p = rANSParams(zipf_frequencies(), NUM_BITS_OUT=8, RANGE_FACTOR=1 << 12)
    assert (p.L, p.H, p.NUM_STATE_BITS) == (1 << 24, (1 << 32) - 1, 32)
    p = rANSParams(scl.Frequencies({"A": 3, "B": 3, "C": 2}), DATA_BLOCK_SIZE_BITS=5, NUM_BITS_OUT=1, RANGE_FACTOR=1)
    assert p.min_shrunk_state == {"A": 3, "B": 3, "C": 2} and p.max_shrunk_state == {"A": 5, "B": 5, "C": 3}
    with pytest.raises(AssertionError):
        tANSParams(scl.Frequencies({"A": 3, "B": 4}))  # M not a power of two (tANS.py:42-44)
    with pytest.raises(AssertionError):
        tANSParams(scl.Frequencies({"A": 1, "B": 3}), NUM_BITS_OUT=2)


def test_golden_params_agree(golden):
    from stanford_compression_library_b200.compressors.rANS import rANSParams

    for c in golden:
        if c["coder"] in ("rans", "tans"):
            p = c["params"]
            fr = scl.Frequencies({i: f for i, f in enumerate(c["freqs"])})
            mine = rANSParams(fr, DATA_BLOCK_SIZE_BITS=p["DATA_BLOCK_SIZE_BITS"], NUM_BITS_OUT=p["NUM_BITS_OUT"], RANGE_FACTOR=p["RANGE_FACTOR"])
            assert mine.NUM_STATE_BITS == p["NUM_STATE_BITS"]


def test_cabi_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "scl_b200.h")).read()
    declared = set(re.findall(r"\b(scl_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"scl_coder", "scl_params"}
    assert declared == set(_cabi.EXPORTS), declared ^ set(_cabi.EXPORTS)
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in _cabi.lib().scl_version()
    assert ctypes.sizeof(_cabi.SclParams) == 48  # must match struct scl_params


def test_no_cpu_fallback_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from stanford_compression_library_b200.compressors.rANS import rANSEncoder, rANSParams

    enc = rANSEncoder(rANSParams(scl.Frequencies({"A": 1, "B": 3})))
    with pytest.raises(_cabi.BackendUnavailable):
        enc.encode_block(scl.DataBlock(["A", "B"]))


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "stanford_compression_library_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                src = open(os.path.join(dirpath, fn)).read()
                assert "import oracle" not in src and "from oracle" not in src and "scl_oracle" not in src, fn


def test_normalize_frequencies_properties():
    from stanford_compression_library_b200.stats import normalize_frequencies

    rng = np.random.default_rng(0)
    for trial in range(50):
        n = int(rng.integers(1, 257))
        c = np.zeros(256, dtype=np.int64)
        idx = rng.choice(256, size=n, replace=False)
        c[idx] = rng.integers(1, 10 ** int(rng.integers(1, 7)), size=n)
        M = int(rng.choice([256, 1024, 4096, 1 << 16]))
        if n > M:
            continue
        f = normalize_frequencies(c, M)
        assert sum(f.freq_dict.values()) == M and min(f.freq_dict.values()) >= 1
        assert list(f.freq_dict) == sorted(int(i) for i in idx)  # ascending byte order, only symbols that occur
    # the benchmark table: Zipf-1.0 expected counts reproduce SURVEY 8(d)'s quantiser exactly
    from stanford_compression_library_b200.workloads import zipf_freq_list, zipf_probabilities

    p = np.array(zipf_probabilities())
    f = normalize_frequencies(np.round(p * 1e9).astype(np.int64), 4096)
    assert [f.freq_dict[b] for b in range(256)] == zipf_freq_list()


def test_bench_reference_arm_prints_exactly_one_json_line():
    """bench.py's contract on stdout: ONE JSON line (library chatter goes to stderr), carrying the
    tier's keys for the reference arm; under torchrun only rank 0 prints."""
    import json
    import subprocess
    import sys

    env = dict(os.environ)
    env.pop("RANK", None)
    env.pop("WORLD_SIZE", None)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "rans_encode_plus_decode_throughput" and d["unit"] == "MB/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # a non-zero rank of a multi-process launch does no work and prints nothing
    env.update(RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
