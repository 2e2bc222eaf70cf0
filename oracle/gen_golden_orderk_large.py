"""Generate tests/golden/orderk_large_v1.npz from the UNMODIFIED reference: AdaptiveOrderKFreqModel cases whose
count table does not fit shared memory on the device (probability_models.py:95-160; SURVEY.md 8f rank 4 for
BYTE alphabets).  TEST INFRASTRUCTURE ONLY; run in the build container (needs /root/reference):

    python oracle/gen_golden_orderk_large.py

Same record layout as oracle/gen_golden.py (whose helpers it reuses); the final count tables are stored as
arrays (`c<i>_final`), not in the JSON meta.
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import gen_golden as gg  # noqa: E402  (imports the reference through oracle/ref_loader.py)

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "orderk_large_v1.npz")


def text_like(n, seed):
    """bytes with first-order structure: a random walk over a small set of byte values mixed with Zipf noise"""
    rng = np.random.default_rng(seed)
    p = 1.0 / np.arange(1, 257)
    p /= p.sum()
    out = np.zeros(n, dtype=np.uint8)
    cur = 65
    for i in range(n):
        if rng.random() < 0.6:
            cur = (cur + int(rng.integers(-2, 3))) % 256
        else:
            cur = int(rng.choice(256, p=p))
        out[i] = cur
    return out


def main():
    gg.run_aec_order_k(256, 1, text_like(700, 80), note="byte alphabet, order 1 (65 536 counters: table in HBM)")
    gg.run_aec_order_k(256, 1, np.zeros(40, dtype=np.uint8), note="byte alphabet, order 1, first symbol only (bits-consumed quirk)", DATA_BLOCK_SIZE_BITS=16)
    gg.run_aec_order_k(256, 0, gg.draw(gg.zipf_freqs(), 500, 81), note="byte alphabet, order 0 (fits shared memory; same path as small alphabets)")
    gg.run_aec_order_k(41, 1, gg.draw(list(range(1, 42)), 500, 82), note="41 symbols, order 1 (1 722 words: just past the shared-memory limit)", PRECISION=24)
    gg.run_aec_order_k(16, 2, gg.draw([9, 7, 5, 5, 4, 3, 3, 2, 2, 2, 1, 1, 1, 1, 1, 1], 900, 83), note="16 symbols, order 2 (256 rows)")
    arrays = dict(gg.arrays)
    for c in gg.cases:
        arrays["c%d_final" % c["id"]] = np.asarray(c["model"].pop("final_freqs"), dtype=np.uint32)
    meta = json.dumps(dict(cases=gg.cases, generator="oracle/gen_golden_orderk_large.py"))
    np.savez_compressed(OUT, meta=np.frombuffer(meta.encode(), dtype=np.uint8), **arrays)
    print("wrote", OUT, len(gg.cases), "cases", os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
