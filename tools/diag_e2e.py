"""Why does the e2e leg move ~42 GB/s when plain copies move 50-55?  Separates the suspects:
copy-only replay of the pipeline's transfer pattern, the same with kernels running beside it, and
the NUMA placement of the process.  python tools/diag_e2e.py  -> JSON lines"""
import json
import os
import subprocess
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stanford_compression_library_b200.compressors.rANS import rANSDecoder, rANSEncoder, rANSParams  # noqa: E402
from stanford_compression_library_b200.pipeline import HostCodecPipeline  # noqa: E402
from stanford_compression_library_b200.workloads import sample_blocks, zipf_frequencies, zipf_probabilities  # noqa: E402


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=20).stdout.strip()
    except Exception as e:  # noqa: BLE001
        return "ERR %s" % e


def main():
    print(json.dumps({"topo": sh("nvidia-smi topo -m | head -6"), "numa": sh("numactl --hardware 2>/dev/null | head -8 || lscpu | grep -i numa"),
                      "lscpu_numa": sh("lscpu | grep -i numa"), "affinity": len(os.sched_getaffinity(0)),
                      "gpu_numa": sh("cat /sys/bus/pci/devices/$(nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader -i 0 | tr A-Z a-z | sed 's/^0000//')/numa_node 2>/dev/null")}), flush=True)
    B, N = 262144, 4096
    torch.cuda.set_device(0)
    dev = torch.device("cuda:0")
    params = rANSParams(zipf_frequencies())
    enc, dec = rANSEncoder(params), rANSDecoder(params)
    data = sample_blocks(zipf_probabilities(), B, N, seed=0, device=dev)
    host_in = torch.empty((B, N), dtype=torch.uint8, pin_memory=True)
    host_in.copy_(data)
    host_out = torch.empty((B, N), dtype=torch.uint8, pin_memory=True)
    host_c = torch.empty(B * N, dtype=torch.uint8, pin_memory=True)
    d_a = torch.empty((B, N), dtype=torch.uint8, device=dev)
    d_b = torch.empty(B * N, dtype=torch.uint8, device=dev)
    s1, s2, s3 = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
    e = enc.encode_blocks(data)
    C = e.total_bytes()
    chunk = 16384
    nch = B // chunk
    cb = C // nch

    def wall(fn, iters=4):
        fn()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(iters):
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        return best

    def h2d_only():
        with torch.cuda.stream(s1):
            for k in range(nch):
                d_a[k * chunk : (k + 1) * chunk].copy_(host_in[k * chunk : (k + 1) * chunk], non_blocking=True)

    def d2h_only():
        with torch.cuda.stream(s2):
            for k in range(nch):
                host_out[k * chunk : (k + 1) * chunk].copy_(d_a[k * chunk : (k + 1) * chunk], non_blocking=True)

    def enc_pattern():  # H2D raw + D2H coded, no kernels
        for k in range(nch):
            with torch.cuda.stream(s1):
                d_a[k * chunk : (k + 1) * chunk].copy_(host_in[k * chunk : (k + 1) * chunk], non_blocking=True)
            with torch.cuda.stream(s2):
                host_c[k * cb : (k + 1) * cb].copy_(d_b[k * cb : (k + 1) * cb], non_blocking=True)

    def enc_pattern_with_kernels():
        for k in range(nch):
            with torch.cuda.stream(s1):
                d_a[k * chunk : (k + 1) * chunk].copy_(host_in[k * chunk : (k + 1) * chunk], non_blocking=True)
            with torch.cuda.stream(s2):
                host_c[k * cb : (k + 1) * cb].copy_(d_b[k * cb : (k + 1) * cb], non_blocking=True)
            with torch.cuda.stream(s3):
                enc.encode_blocks(data[k * chunk : (k + 1) * chunk], reuse=None)

    raw = B * N
    res = {}
    res["h2d_only_GBps"] = raw / wall(h2d_only) / 1e9
    res["d2h_only_GBps"] = raw / wall(d2h_only) / 1e9
    t = wall(enc_pattern)
    res["enc_pattern_copy_only_ms"] = t * 1e3
    res["enc_pattern_h2d_GBps"] = raw / t / 1e9
    t = wall(enc_pattern_with_kernels)
    res["enc_pattern_with_kernels_ms"] = t * 1e3
    pipe = HostCodecPipeline(enc, dec, N, B, chunk_blocks=chunk, depth=3)
    t = wall(lambda: pipe.encode(host_in, host_c))
    res["pipeline_encode_ms"] = t * 1e3
    total, lens = pipe.encode(host_in, host_c)
    t = wall(lambda: pipe.decode(host_c, lens, host_out))
    res["pipeline_decode_ms"] = t * 1e3
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
