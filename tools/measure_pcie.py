"""Host<->device copy bandwidth from pinned memory on this box: the ceiling of bench.py's `e2e` leg.
python tools/measure_pcie.py [--mib 1024]   -> one JSON line"""
import argparse
import json

import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mib", type=int, default=1024)
    args = ap.parse_args()
    n = args.mib << 20
    dev = torch.device("cuda", 0)
    h_a = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_b = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_a.fill_(3)
    d_a = torch.empty(n, dtype=torch.uint8, device=dev)
    d_b = torch.full((n,), 5, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def timed(fn, iters=5):
        fn()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(iters):
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            s1.wait_stream(torch.cuda.current_stream())
            s2.wait_stream(torch.cuda.current_stream())
            fn()
            torch.cuda.current_stream().wait_stream(s1)
            torch.cuda.current_stream().wait_stream(s2)
            t1.record()
            torch.cuda.synchronize()
            best = min(best, t0.elapsed_time(t1))
        return best

    def h2d():
        with torch.cuda.stream(s1):
            d_a.copy_(h_a, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h_b.copy_(d_b, non_blocking=True)

    def both():
        h2d()
        d2h()

    def chunked(chunk_mib):
        c = chunk_mib << 20

        def fn():
            for lo in range(0, n, c):
                with torch.cuda.stream(s1):
                    d_a[lo : lo + c].copy_(h_a[lo : lo + c], non_blocking=True)
                with torch.cuda.stream(s2):
                    h_b[lo : lo + c].copy_(d_b[lo : lo + c], non_blocking=True)

        return fn

    res = {
        "bytes": n,
        "h2d_GBps": n / timed(h2d) / 1e6,
        "d2h_GBps": n / timed(d2h) / 1e6,
        "both_directions_each_GBps": n / timed(both) / 1e6,
        "both_directions_128MiB_chunks_each_GBps": n / timed(chunked(128)) / 1e6,
        "both_directions_16MiB_chunks_each_GBps": n / timed(chunked(16)) / 1e6,
    }
    print(json.dumps(res))


if __name__ == "__main__":
    main()
