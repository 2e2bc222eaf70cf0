"""Timeline of the fused packed encoder (scl_coder_debug_trace): when the coding warps finish each round, how far
behind the copy pool runs, how long the tail is.  Diagnostic; prints one JSON object.
    python tools/trace_packed.py [--blocks 262144] [--block-len 4096]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stanford_compression_library_b200.compressors.rANS import rANSEncoder, rANSParams  # noqa: E402
from stanford_compression_library_b200.workloads import sample_blocks, zipf_frequencies, zipf_probabilities  # noqa: E402

WORDS = 40


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=262144)
    ap.add_argument("--block-len", type=int, default=4096)
    ap.add_argument("--mode", type=int, default=0, help="scl_coder_debug_path value (copy pool variants)")
    a = ap.parse_args()
    torch.cuda.set_device(0)
    enc = rANSEncoder(rANSParams(zipf_frequencies()))
    data = sample_blocks(zipf_probabilities(), a.blocks, a.block_len, seed=0, device="cuda:0")
    p = enc.encode_blocks_packed(data).check()
    n_sm = torch.cuda.get_device_properties(0).multi_processor_count
    trace = torch.zeros(n_sm * 32 * WORDS, dtype=torch.int64, device="cuda:0")
    dc = enc.device_coder()
    dc.debug_path(a.mode)
    for _ in range(2):
        enc.encode_blocks_packed(data, reuse=p)
    torch.cuda.synchronize()
    trace.zero_()
    dc.debug_trace(trace)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    enc.encode_blocks_packed(data, reuse=p)
    ev1.record()
    torch.cuda.synchronize()
    dc.debug_trace(None)
    t = trace.cpu().numpy().reshape(n_sm, 32, WORDS).astype(np.float64)
    start = t[:, :, 0]
    t0 = start[start > 0].min()
    us = lambda x: (x - t0) / 1e3  # noqa: E731
    out = {"mode": a.mode, "blocks": a.blocks, "block_len": a.block_len, "kernel_ms_cuda_events": ev0.elapsed_time(ev1)}
    rounds = []
    for r in range(19):
        e = t[:, :, 1 + r]
        e = e[e > 0]
        if e.size == 0:
            break
        rounds.append({"round": r, "warps": int(e.size), "end_us_min": us(e.min()), "end_us_median": us(np.median(e)), "end_us_max": us(e.max())})
    out["coding_rounds"] = rounds
    tasks = t[:, :, 20]
    is_copy_warp = (t[:, :, 1] == 0) & (tasks > 0)  # never finished a coding round: a dedicated copy warp
    helper = (t[:, :, 1] > 0) & (tasks > 0)
    for name, m in (("copy_warps", is_copy_warp), ("coding_warps_helping", helper)):
        if m.any():
            busy, wait = t[:, :, 23][m] / 1e3, t[:, :, 24][m] / 1e3
            out[name] = {"warps": int(m.sum()), "tasks_total": int(tasks[m].sum()), "tasks_per_warp_median": float(np.median(tasks[m])),
                         "busy_us_median": float(np.median(busy)), "busy_us_max": float(busy.max()), "waiting_for_resolution_us_median": float(np.median(wait)),
                         "us_per_task_median": float(np.median(busy / tasks[m])), "ring_wait_cycles_per_task_median": float(np.median(t[:, :, 25][m] / tasks[m])),
                         "first_copy_start_us_median": us(np.median(t[:, :, 21][m])), "last_copy_end_us_median": us(np.median(t[:, :, 22][m])),
                         "last_copy_end_us_max": us(t[:, :, 22][m].max())}
    last_code = max(r["end_us_max"] for r in rounds) if rounds else 0.0
    out["tail_us_after_last_coding_warp"] = us(t[:, :, 22].max()) - last_code
    print(json.dumps(out))


if __name__ == "__main__":
    main()
