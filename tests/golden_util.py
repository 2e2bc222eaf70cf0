"""Loader for tests/golden/golden_v1.npz (written by oracle/gen_golden.py from the unmodified reference)."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz")


def load_golden():
    z = np.load(GOLDEN)
    meta = json.loads(bytes(z["meta"]).decode())
    cases = []
    for c in meta["cases"]:
        c = dict(c)
        c["data"] = z["c%d_data" % c["id"]]
        c["enc"] = z["c%d_enc" % c["id"]]
        cases.append(c)
    return cases


def with_garbage(enc, nbits, garbage):
    """packed stream + '0101' garbage string -> (packed, total_bits)"""
    bits = np.unpackbits(np.asarray(enc, dtype=np.uint8))[:nbits]
    g = np.array([1 if ch == "1" else 0 for ch in garbage], dtype=np.uint8)
    allb = np.concatenate([bits, g])
    return np.packbits(allb), int(allb.size)


def case_id(c):
    return "%02d-%s-%s" % (c["id"], c["coder"], c["note"][:40].replace(" ", "_"))


def fresh_model_table(c):
    """the model table handed to the oracle: IID/fixed -> the initial freqs; order-k -> ones + context 0"""
    if c["coder"] != "aec":
        return None
    if c["model"]["kind"] == "order_k":
        n = len(c["freqs"]) ** (c["model"]["k"] + 1)
        return np.array([1] * n + [0], dtype=np.uint64)
    return np.array(c["freqs"], dtype=np.uint64)


def expected_final_model(c):
    m = c["model"]
    return m["final_freqs"] + [m["final_ctx"]] if m["kind"] == "order_k" else m["final_freqs"]
