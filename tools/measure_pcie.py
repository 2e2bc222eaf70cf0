"""Host<->device copy bandwidth from pinned memory on this box: the ceiling of bench.py's `e2e` leg.

    python tools/measure_pcie.py [--mib 1024]                                    one GPU   -> one JSON line
    python -m torch.distributed.run --nproc-per-node N tools/measure_pcie.py     N GPUs AT THE SAME TIME: per-rank and
                                                                                 aggregate numbers (rank 0 prints)
With several ranks every measurement starts at a barrier, so the ranks' copies compete for the host's memory
bandwidth and the PCIe root complexes exactly as bench.py's e2e legs do at --gpus N.  --bind pins each rank to
the CPUs of its GPU's NUMA node before the pinned buffers are allocated (first touch places them there)."""
import argparse
import json
import os

import torch


def numa_of_gpu(index):
    try:
        bus = torch.cuda.get_device_properties(index).pci_bus_id
        dom = torch.cuda.get_device_properties(index).pci_domain_id
        dev = torch.cuda.get_device_properties(index).pci_device_id
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (dom, bus, dev)
        return int(open(path).read().strip())
    except Exception:
        return None


def cpus_of_node(node):
    try:
        txt = open("/sys/devices/system/node/node%d/cpulist" % node).read().strip()
        out = set()
        for part in txt.split(","):
            a, _, b = part.partition("-")
            out.update(range(int(a), int(b or a) + 1))
        return out
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mib", type=int, default=1024)
    ap.add_argument("--bind", action="store_true")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("NCCL_DEBUG", "WARN")  # keep NCCL's version banner off stdout
        dist.init_process_group("nccl", device_id=dev)
    node = numa_of_gpu(local)
    bound = None
    if args.bind and node is not None and node >= 0:
        cpus = cpus_of_node(node)
        if cpus:
            os.sched_setaffinity(0, cpus & os.sched_getaffinity(0) or cpus)
            bound = sorted(os.sched_getaffinity(0))
    n = args.mib << 20
    h_a = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_b = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_a.fill_(3)
    h_b.fill_(1)
    d_a = torch.empty(n, dtype=torch.uint8, device=dev)
    d_b = torch.full((n,), 5, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def timed(fn, iters=5):
        fn()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(iters):
            if dist:
                dist.barrier()
            torch.cuda.synchronize()
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

    mine = {
        "h2d_GBps": n / timed(h2d) / 1e6,
        "d2h_GBps": n / timed(d2h) / 1e6,
        "both_directions_each_GBps": n / timed(both) / 1e6,
        "both_directions_64MiB_chunks_each_GBps": n / timed(chunked(64)) / 1e6,
    }
    info = {"gpu_numa_node": node, "bound_cpus": bound, "affinity": len(os.sched_getaffinity(0))}
    if world > 1:
        keys = sorted(mine)
        t = torch.tensor([mine[k] for k in keys], dtype=torch.float64, device=dev)
        allv = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allv, t)
        infos = [None] * world
        dist.all_gather_object(infos, info)
        if rank == 0:
            per_rank = [{k: float(v[i]) for i, k in enumerate(keys)} for v in allv]
            agg = {k + "_sum": sum(r[k] for r in per_rank) for k in keys}
            agg.update({k + "_min": min(r[k] for r in per_rank) for k in keys})
            print(json.dumps({"ranks": world, "bytes_per_rank": n, "bind": args.bind, "aggregate": agg, "per_rank": per_rank, "rank_info": infos,
                              "host_cpus": os.cpu_count(), "numa_nodes": len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")]) if os.path.isdir("/sys/devices/system/node") else None}))
        dist.destroy_process_group()
    else:
        mine.update({"bytes": n, "bind": args.bind, **info, "host_cpus": os.cpu_count()})
        print(json.dumps(mine))


if __name__ == "__main__":
    main()
