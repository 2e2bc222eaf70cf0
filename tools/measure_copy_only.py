"""The fused encoder's copy stage alone (scl_debug_copy_only): N warps per SM re-copy the streams of a finished
packed encode with no coder beside them.  If 4 warps per SM alone are as slow per task as the 4 copy warps inside the
fused kernel, the pool is bound by memory latency; if they are much faster, by the coding warps' issue slots.
    python tools/measure_copy_only.py [--blocks 262144]"""
import argparse
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stanford_compression_library_b200 import _cabi  # noqa: E402
from stanford_compression_library_b200.compressors.rANS import rANSEncoder, rANSParams  # noqa: E402
from stanford_compression_library_b200.workloads import sample_blocks, zipf_frequencies, zipf_probabilities  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=262144)
    a = ap.parse_args()
    torch.cuda.set_device(0)
    B, N = a.blocks, 4096
    enc = rANSEncoder(rANSParams(zipf_frequencies()))
    data = sample_blocks(zipf_probabilities(), B, N, seed=0, device="cuda:0")
    p = enc.encode_blocks_packed(data).check()
    total = int(p.byte_offset[-1])
    ref = p.buf[:total].clone()
    scratch, stride, ws = p._scratch
    lib = _cabi.lib()
    ptr = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    h = enc.device_coder()._h
    n_sm = torch.cuda.get_device_properties(0).multi_processor_count
    for stages, warps in [(st, w) for st in (0, 0x104, 0x202, 0x203, 0x204) for w in (1, 4, 8, 16) if w * ((st & 255) * (2140 if st >> 9 else 1120 if st >> 8 else 608) + 512) <= 200 * 1024]:
        p.buf.zero_()

        def run():
            rc = lib.scl_debug_copy_only(h, B, ptr(scratch), stride, ptr(p.buf), p.buf.numel() - 16, 0, ptr(p.byte_offset), ptr(p.bit_offset), ptr(p.bit_len),
                                         ptr(p.status), warps, stages, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
            assert rc == 0

        run()
        torch.cuda.synchronize()
        assert torch.equal(p.buf[:total], ref), "copy-only output differs (stages %d, warps %d)" % (stages, warps)
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        tasks_per_warp = (B / 32) / (n_sm * warps)
        print(json.dumps({"ring_stages": stages & 255, "piece_bytes": 2048 if stages >> 9 else 1024 if stages >> 8 else 512, "warps_per_sm": warps, "ms": best, "us_per_task_per_warp": best * 1e3 / tasks_per_warp, "GBps_read_plus_write": 2 * total / best / 1e6,
                          "GBps_per_warp": total / best / 1e6 / (n_sm * warps)}), flush=True)


if __name__ == "__main__":
    main()
