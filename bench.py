#!/usr/bin/env python
"""bench.py -- the BASELINE.json metric: rANS encode+decode MB/s (uncompressed bytes) on 8 GiB of Zipf-1.0
bytes, blocks of 4 KiB, sharded by block range over N GPUs; % of the HBM roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scaling strong|weak]

A "step" = one encode pass (symbols in -> the CONTIGUOUS reference stream out: encode_blocks_packed) + one
decode pass (packed stream in -> symbols out) over this rank's shard.  Default --scaling strong: the global
stream is BASELINE configs[4], 2 097 152 blocks x 4 KiB = 8 GiB, at every N (N=1 holds all of it);
--scaling weak fixes 262 144 blocks (1 GiB) per GPU as round 1 did.  configs[1] (65 536 blocks on one GPU) is
measured in the same run at N=1 and reported under "also".  Every rank bit-compares a strided sample of its
GPU streams with the CPU oracle outside the timed region ("parity").  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BLOCK_LEN = 4096
BLOCKS_PER_GPU = 262144       # --scaling weak
GLOBAL_BLOCKS = 2097152       # --scaling strong: 8 GiB (BASELINE configs[4])
PARITY_BLOCKS = 256
METRIC = "rans_encode_plus_decode_throughput"
UNIT = "MB/s"
# headline parameter set = the reference's defaults, rANSParams(freqs) (rANS.py:88-95);
# the 32-bit-state / byte-renormalisation set is measured beside it
VARIANTS = {
    "default": {},
    "nbo8_rf4096": dict(NUM_BITS_OUT=8, RANGE_FACTOR=1 << 12),
}
HEADLINE = "default"


_REAL_STDOUT = None


_T0 = time.time()


def progress(msg):
    """stage markers on stderr (stdout carries only the JSON line)"""
    if int(os.environ.get("RANK", "0")) == 0:
        sys.stderr.write("[bench +%.1fs] %s\n" % (time.time() - _T0, msg))
        sys.stderr.flush()


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def host_threads():
    """Host threads the CPU arm may use: the cores this process is allowed on (torchrun exports
    OMP_NUM_THREADS=1, so the OpenMP default is not a usable answer)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(int(r[1]) for r in self.rows if len(r) >= 8 and r[1].isdigit())
        mx = [int(r[2]) for r in self.rows if len(r) >= 8 and r[2].isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(fl, kw, data_host, n_threads, label):
    """The oracle (C restatement of the reference's loops, oracle/scl_oracle.c) on the host cores:
    encode + decode of a bounded sample of the same workload.  A reported baseline, not a target."""
    import numpy as np

    from oracle import scl_oracle as so

    oracle = so.Oracle.rans(fl, **kw)
    B, N = data_host.shape
    stride = 2 * N + 64
    t0 = time.perf_counter()
    out, bits, st = oracle.encode_batch(data_host, out_stride=stride, n_threads=n_threads)
    t1 = time.perf_counter()
    offs = np.arange(B, dtype=np.uint64) * np.uint64(stride * 8)
    dec, sz, used, st2 = oracle.decode_batch(out, offs, bits, N, n_threads=n_threads)
    t2 = time.perf_counter()
    assert (st == 0).all() and (st2 == 0).all() and (dec == data_host).all()
    nbytes = B * N
    return {
        "value": nbytes / (t2 - t0) / 1e6, "unit": UNIT, "cores": n_threads, "kind": "port",
        "sample": "%d blocks x %d B of the same Zipf batch, %s; encode %.1f MB/s, decode %.1f MB/s" % (B, N, label, nbytes / (t1 - t0) / 1e6, nbytes / (t2 - t1) / 1e6),
        "note": "C restatement of the reference's per-symbol loops (OpenMP across blocks); the reference itself is pure Python, "
                "~10^4 times slower: it is timed beside this on a few blocks per core (`python_reference`)",
    }


def bench_other_configs(data, freqs, peak, dev):
    """tANS on configs[2] and the arithmetic coder on configs[3]: encode / decode MB/s of raw bytes, CUDA events,
    best of 5, round trip checked."""
    import torch

    from stanford_compression_library_b200 import Frequencies
    from stanford_compression_library_b200.compressors.arithmetic_coding import AECParams, ArithmeticDecoder, ArithmeticEncoder
    from stanford_compression_library_b200.compressors.probability_models import AdaptiveIIDFreqModel
    from stanford_compression_library_b200.compressors.tANS import tANSDecoder, tANSEncoder, tANSParams

    def measure(enc, dec, d):
        nB, n = d.shape
        e = enc.encode_blocks(d)
        r = dec.decode_blocks(e, n)
        e.check(), r.check()
        assert torch.equal(r.symbols[:, :n], d), "round trip failed"
        C = e.total_bytes()
        best = [1e30, 1e30]
        for _ in range(5):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record()
            enc.encode_blocks(d, reuse=e)
            ev[1].record()
            dec.decode_blocks(e, n, reuse=r)
            ev[2].record()
            torch.cuda.synchronize()
            best = [min(best[0], ev[0].elapsed_time(ev[1])), min(best[1], ev[1].elapsed_time(ev[2]))]
        raw = nB * n
        return {"blocks": nB, "block_len": n, "encode_ms": best[0], "decode_ms": best[1], "encode_MBps": raw / best[0] / 1e3, "decode_MBps": raw / best[1] / 1e3,
                "bits_per_symbol": 8.0 * C / raw, "roofline_encode_frac": (raw + C) / best[0] / 1e6 / peak, "roofline_decode_frac": (raw + C) / best[1] / 1e6 / peak}

    out = {}
    tp = tANSParams(freqs, RANGE_FACTOR=1)  # 256-symbol alphabet, L = M = 4096 states (SURVEY.md 8a on configs[2])
    out["cfg3_tans_65536_blocks_L4096"] = measure(tANSEncoder(tp), tANSDecoder(tp), data[:65536])
    ap = AECParams()
    uni = Frequencies({b: 1 for b in range(256)})
    d4 = data.reshape(-1, 1024)[: 1 << 20]  # the same Zipf bytes as 1 KiB blocks
    out["cfg4_arithmetic_adaptive_order0_%d_blocks_x_1KiB" % d4.shape[0]] = measure(
        ArithmeticEncoder(ap, AdaptiveIIDFreqModel(uni, ap.MAX_ALLOWED_TOTAL_FREQ)), ArithmeticDecoder(ap, AdaptiveIIDFreqModel(uni, ap.MAX_ALLOWED_TOTAL_FREQ)), d4)
    return out


def python_reference_leg(fl, kw, rows, cores, blocks_per_worker=2):
    """The UNMODIFIED reference coder (scl/compressors/rANS.py:123-297, shim-backed BitArray) on the host cores:
    one worker per core, `blocks_per_worker` blocks of the same batch each (oracle/ref_python_bench.py).
    Returns (result dict for the JSON line, [(bytes, nbits)] per block) or (dict with "unavailable", None)."""
    try:
        from oracle import ref_loader, ref_python_bench

        if not ref_loader.reference_available():
            return {"kind": "reference-python", "unavailable": "no reference tree (neither /root/reference nor the staged oracle/_ref/pyref copy)"}, None
        r = ref_python_bench.run(fl, kw, rows, cores, blocks_per_worker)
    except Exception as e:  # a reported baseline must not take the bench down
        return {"kind": "reference-python", "unavailable": "%s: %s" % (type(e).__name__, e)}, None
    nbytes = r["n_blocks"] * len(rows[0])
    cpu_s = r["enc_cpu_s"] + r["dec_cpu_s"]
    return {
        "kind": "reference-python", "value": nbytes / r["wall_s"] / 1e6, "unit": UNIT, "cores": r["workers"],
        "per_core_MBps": nbytes / cpu_s / 1e6, "encode_MBps_per_core": nbytes / r["enc_cpu_s"] / 1e6, "decode_MBps_per_core": nbytes / r["dec_cpu_s"] / 1e6,
        "sample": "%d blocks x %d B of the same Zipf batch (%d per worker), encode + decode, wall %.1f s" % (r["n_blocks"], len(rows[0]), blocks_per_worker, r["wall_s"]),
        "note": "the reference's own rANSEncoder / rANSDecoder, unmodified, imported from %s with the pure-Python stand-in for the missing "
                "`bitarray` wheel (oracle/bitarray_shim): the arithmetic is the reference's, the bit-container overhead is the shim's" % ref_loader.REFERENCE_ROOT,
    }, r["streams"]


def sample_indices(n_blocks, want):
    """A strided sample of block indices that always holds the first and the last block."""
    want = max(2, min(int(want), int(n_blocks)))
    idx = sorted({int(round(i * (n_blocks - 1) / (want - 1))) for i in range(want)}) if n_blocks > 1 else [0]
    return idx


def gather_streams(e, idx):
    """[(bytes, nbits)] of the packed streams of blocks `idx` (device EncodedBlocks with byte_offset) -- host side, test leg."""
    import torch

    ii = torch.tensor(idx, dtype=torch.int64, device=e.buf.device)
    lo = e.byte_offset[ii].cpu().tolist()
    nb = e.bit_len[ii].cpu().tolist()
    out = []
    for a, n in zip(lo, nb):
        out.append((e.buf[a : a + (n + 7) // 8].cpu().numpy().tobytes(), int(n)))
    return out


def parity_check(fl, kw, data, e, want):
    """GPU bits vs CPU oracle bits on a strided sample of this rank's blocks (outside the timed region).
    Returns (ok, blocks_checked, first mismatching block or -1)."""
    import torch

    from oracle import scl_oracle as so

    idx = sample_indices(data.shape[0], want)
    host = data[torch.tensor(idx, dtype=torch.int64, device=data.device)].cpu().numpy()
    got = gather_streams(e, idx)
    oracle = so.Oracle.rans(fl, **kw)
    for j, b in enumerate(idx):
        ref_bytes, ref_bits = oracle.encode_block(host[j])
        if got[j][1] != ref_bits or got[j][0] != ref_bytes.tobytes():
            return False, j, b
        dec, used = oracle.decode_block(ref_bytes, ref_bits)  # and the oracle reads its own stream back
        if used != ref_bits or not (dec == host[j]).all():
            return False, j, b
    return True, len(idx), -1


def run_reference(args, rank, world):
    """--impl reference: the CPU implementation of the path on all host threads, same metric / config, bounded
    sample per step.  The timed arm is the C restatement (oracle/_ref/libscl_oracle.so): the reference is pure
    Python, ~10^4 times slower, and is timed beside it on a few blocks per core (`python_reference`)."""
    if rank != 0:
        return
    import numpy as np

    from oracle import scl_oracle as so
    from stanford_compression_library_b200.workloads import zipf_freq_list, zipf_probabilities

    so.build()
    fl = zipf_freq_list()
    threads = host_threads()
    B = min(BLOCKS_PER_GPU, 1024 * threads)
    rng = np.random.default_rng(0)
    data = rng.choice(256, size=(B, BLOCK_LEN), p=np.array(zipf_probabilities())).astype(np.uint8)
    kw = VARIANTS[HEADLINE]
    res = None
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        res = cpu_baseline(fl, kw, data, threads, HEADLINE)
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    value = B * BLOCK_LEN * len(times) / total / 1e6
    res["value"] = value
    if not args.no_python_reference:
        py, streams = python_reference_leg(fl, kw, [r.tolist() for r in data[: 2 * threads]], threads)
        if streams is not None:  # the port is pinned to the reference here as well
            oracle = so.Oracle.rans(fl, **kw)
            py["port_agrees"] = all(oracle.encode_block(data[j])[0].tobytes() == sb and oracle.encode_block(data[j])[1] == nb for j, (sb, nb) in enumerate(streams))
        res["python_reference"] = py
    total_blocks = GLOBAL_BLOCKS if args.scaling == "strong" else BLOCKS_PER_GPU * world
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": workload_text(total_blocks, world, args.scaling) + "; CPU arm: %d-block x %d B sample of it per step" % (B, BLOCK_LEN),
                   "params": HEADLINE, "block_len": BLOCK_LEN},
        "cpu_baseline": res,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_text(total_blocks, world, scaling):
    return ("rANS encode+decode of %d blocks x %d B Zipf-1.0 bytes = %.2f GiB (%s scaling: %d blocks per GPU on %d GPU(s)%s), 256-symbol static "
            "Frequencies (M=4096), reference-default rANSParams (NUM_BITS_OUT=1, RANGE_FACTOR=2^16)"
            % (total_blocks, BLOCK_LEN, total_blocks * BLOCK_LEN / 2**30, scaling, total_blocks // world, world,
               "; BASELINE configs[4]" if total_blocks == GLOBAL_BLOCKS else ""))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--global-blocks", type=int, default=GLOBAL_BLOCKS, help="--scaling strong: blocks of the whole stream")
    ap.add_argument("--blocks-per-gpu", type=int, default=BLOCKS_PER_GPU, help="--scaling weak: blocks per GPU")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--e2e-chunk", type=int, default=16384)
    ap.add_argument("--e2e-depth", type=int, default=3)
    ap.add_argument("--parity-blocks", type=int, default=PARITY_BLOCKS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-python-reference", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line: keep a private handle to it and point fd 1 at stderr for the rest of
    # the run, so that library chatter (NCCL's version banner under NCCL_DEBUG, torch warnings) cannot get in
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist

    from stanford_compression_library_b200 import build as scl_build
    from stanford_compression_library_b200.compressors.rANS import rANSDecoder, rANSEncoder, rANSParams
    from stanford_compression_library_b200.sharding import broadcast_frequencies, gather_compressed_sizes, shard_range
    from stanford_compression_library_b200.workloads import sample_stream_blocks, zipf_frequencies, zipf_probabilities

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")  # keep NCCL's version banner off stdout: one JSON line only
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        scl_build.build()
    if world > 1:
        dist.barrier()

    # frequency table: built on rank 0, broadcast over NCCL (the only collective on this path)
    freqs = broadcast_frequencies(zipf_frequencies() if rank == 0 else None)
    fl = [int(f) for f in freqs.freq_list]
    N = BLOCK_LEN
    total_blocks = args.global_blocks if args.scaling == "strong" else args.blocks_per_gpu * world
    lo, hi = shard_range(total_blocks, rank, world)  # this rank's contiguous block range of the global stream
    B = hi - lo
    progress("generating %d blocks on the device" % B)
    data = sample_stream_blocks(zipf_probabilities(), lo, hi, N, dev)  # the same global bytes at every N
    torch.cuda.synchronize()
    progress("data ready")
    peak, peak_src = measured_peak_gbs()
    sampler = ClockSampler(local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def bench_variant(name, data, steps, warmup, sample_clocks=False, packed=True, parity=0):
        """packed=True: the step the metric names -- symbols in -> contiguous stream out (one fused launch), packed stream
        in -> symbols out.  packed=False: round 1's form (streams left in their fixed-stride slots), reported beside it."""
        params = rANSParams(freqs, **VARIANTS[name])
        enc, dec = rANSEncoder(params), rANSDecoder(params)
        nB = data.shape[0]
        do_enc = (lambda reuse: enc.encode_blocks_packed(data, capacity=nB * N + (1 << 20), reuse=reuse)) if packed else (lambda reuse: enc.encode_blocks(data, reuse=reuse))
        progress("variant %s, %d blocks, %s: first pass" % (name, nB, "packed" if packed else "slots"))
        e = do_enc(None)
        d = dec.decode_blocks(e, N)
        e.check(), d.check()
        assert torch.equal(d.symbols[:, :N], data) and torch.equal(d.bits_consumed, e.bit_len), "round trip failed"
        C = e.total_bytes()
        par = None
        if parity:  # GPU bits == CPU oracle bits, on this rank's own shard, before anything is timed
            ok, checked, bad = parity_check(fl, VARIANTS[name], data, e, parity)
            par = (ok, checked, bad)
            progress("parity vs oracle: %s (%d blocks)" % (ok, checked))
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
        for _ in range(warmup):
            do_enc(e)
            dec.decode_blocks(e, N, reuse=d)
        barrier()
        if sample_clocks:
            sampler.start()
        for i in range(steps):
            ev[i][0].record()
            do_enc(e)
            ev[i][1].record()
            dec.decode_blocks(e, N, reuse=d)
            ev[i][2].record()
        barrier()
        assert torch.equal(d.symbols[:, :N], data), "round trip failed after the timed loop"
        total_ms = ev[0][0].elapsed_time(ev[-1][2])
        enc_ms = sorted(ev[i][0].elapsed_time(ev[i][1]) for i in range(steps))
        dec_ms = sorted(ev[i][1].elapsed_time(ev[i][2]) for i in range(steps))
        t = torch.tensor([total_ms, sum(enc_ms) / steps, sum(dec_ms) / steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)  # device time, max over ranks
        total_ms, enc_avg, dec_avg = t.tolist()
        raw = nB * N
        r = dict(name=name, C=C, raw=raw, total_ms=total_ms, enc_ms=enc_avg, dec_ms=dec_avg, enc_best=enc_ms[0], dec_best=dec_ms[0], steps=steps, packed=packed,
                 parity=par, paths=(enc.device_coder().path(False), dec.device_coder().path(True)))
        del e, d
        torch.cuda.empty_cache()
        return r, enc, dec

    def summarize(r, n_gpus):
        raw, C = r["raw"], r["C"]
        return {
            "value": n_gpus * raw * r["steps"] / (r["total_ms"] * 1e-3) / 1e6,
            "ms_per_step": r["total_ms"] / r["steps"],
            "encode_MBps": n_gpus * raw / (r["enc_ms"] * 1e-3) / 1e6,
            "decode_MBps": n_gpus * raw / (r["dec_ms"] * 1e-3) / 1e6,
            "encode_ms": r["enc_ms"], "decode_ms": r["dec_ms"], "blocks_per_gpu": raw // N,
            "output": "contiguous stream (fused packed encode)" if r["packed"] else "fixed-stride slots (no compaction)",
            "compressed_bytes_per_gpu": C, "bits_per_symbol": 8.0 * C / raw, "kernel_paths": r["paths"],
            "roofline_encode_frac": (raw + C) / (r["enc_ms"] * 1e-3) / 1e9 / peak,
            "roofline_decode_frac": (raw + C) / (r["dec_ms"] * 1e-3) / 1e9 / peak,
            "hbm_read_only_frac": {"encode": raw / (r["enc_ms"] * 1e-3) / 1e9 / peak, "decode": C / (r["dec_ms"] * 1e-3) / 1e9 / peak},
        }

    # clocks / throttle reasons are sampled over every timed loop of this run (headline + variants)
    head, enc, dec = bench_variant(HEADLINE, data, args.steps, args.warmup, sample_clocks=True, parity=args.parity_blocks)
    others = {k: bench_variant(k, data, max(5, args.steps // 2), 3, parity=min(64, args.parity_blocks))[0] for k in VARIANTS if k != HEADLINE}
    slots = bench_variant(HEADLINE, data, max(5, args.steps // 2), 3, packed=False)[0]
    if args.steps < 100 and B * N <= (2 << 30):  # short runs on small shards: keep the GPU under load a little longer so nvidia-smi gets samples
        bench_variant(HEADLINE, data, 100, 0)
    head["clocks"] = sampler.stop()
    hs = summarize(head, world)

    # ---- parity: every rank compared a strided sample of ITS blocks with the oracle; all ranks must agree --------
    ok, checked, bad = head["parity"] if head["parity"] else (True, 0, -1)
    for r in others.values():
        if r["parity"]:
            ok = ok and r["parity"][0]
    pt = torch.tensor([1 if ok else 0, checked, 1], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(pt, op=dist.ReduceOp.SUM)
    parity = {"ok": int(pt[0]) == int(pt[2]), "blocks_checked": int(pt[1]), "ranks": int(pt[2]), "per_rank_sample": checked,
              "what": "packed GPU stream bytes + bit_len of a strided sample of every rank's shard (first and last block included) == CPU oracle "
                      "(oracle/scl_oracle.c, pinned to the reference's golden vectors); all variants; outside the timed region"}
    assert parity["ok"], "GPU bitstream differs from the CPU oracle (rank %d, block %d of its shard)" % (rank, bad)

    # ---- roofline of the dominant kernel (the longer of the two launches of a step) ----------
    dom = "decode" if head["dec_ms"] >= head["enc_ms"] else "encode"
    dom_ms = max(head["dec_ms"], head["enc_ms"])
    alg_bytes = head["raw"] + head["C"]  # N + C per launch: enc reads N writes C, dec reads C writes N
    achieved = alg_bytes / (dom_ms * 1e-3) / 1e9
    traffic = None
    kernel_name = "fast_%s_v2_kernel" % dom
    tp = os.path.join(ROOT, "profiles", "traffic.json")  # dram bytes per launch from the committed ncu --set full capture
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp)).get("rans32_%s_kernel%s" % (dom, "_packed" if dom == "encode" else ""), {})
            per_block = tj.get("dram_bytes_per_block")
            traffic = per_block * B if per_block else None
            kernel_name = tj.get("kernel", kernel_name)
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                "avg_launch_ms": dom_ms, "encode_frac": hs["roofline_encode_frac"], "decode_frac": hs["roofline_decode_frac"],
                "hbm_read_only_frac": hs["hbm_read_only_frac"],
                "note": "algorithmic bytes = raw + coded per launch; the fused encoder also moves every coded byte through its scratch slot "
                        "(LIFO streams: write, read back, write packed), which shows in `traffic`, not in `achieved`"}

    # ---- end to end through the public API with HOST buffers ----------------------------------
    # HostCodecPipeline: pinned host raw -> (H2D, fused packed encode, D2H) -> pinned host coded bytes + bit lengths,
    # then host coded -> (H2D, decode, D2H) -> pinned host raw; chunks flow through an upload, a kernel and a
    # download stream over a ring of staging slots.  Every byte crosses PCIe inside the timed region.
    from stanford_compression_library_b200.pipeline import HostCodecPipeline

    progress("e2e: pinning host buffers")
    host_in = torch.empty((B, N), dtype=torch.uint8, pin_memory=True)
    host_in.copy_(data)
    host_out = torch.empty((B, N), dtype=torch.uint8, pin_memory=True)
    pipe = HostCodecPipeline(enc, dec, N, B, chunk_blocks=args.e2e_chunk, depth=args.e2e_depth)
    host_c = torch.empty(head["C"] + (1 << 20), dtype=torch.uint8, pin_memory=True)
    e2e_steps = max(1, min(args.e2e_steps, args.steps))
    h2d = d2h = 0

    def e2e_step():
        nonlocal h2d, d2h
        total, lens = pipe.encode(host_in, host_c)
        h2d, d2h = pipe.h2d_bytes, pipe.d2h_bytes
        pipe.decode(host_c, lens, host_out)
        h2d += pipe.h2d_bytes
        d2h += pipe.d2h_bytes
        return total

    progress("e2e: first step")
    total_c = e2e_step()
    barrier()
    progress("e2e: timed steps")
    assert total_c == head["C"] and torch.equal(host_out, host_in), "e2e round trip failed"
    host_out.zero_()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(e2e_steps):
        e2e_step()
    t1.record()
    barrier()
    assert torch.equal(host_out, host_in), "e2e round trip failed"
    tt = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e = {"value": world * B * N * e2e_steps / (tt.item() * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "steps": e2e_steps, "note": "HostCodecPipeline: pinned host buffers, %d-block chunks over upload / kernel / download streams, %d staging slots; raw and "
                                       "coded bytes cross PCIe inside the timed region (per GPU)" % (pipe.chunk, pipe.depth)}
    del pipe, host_out, host_c
    torch.cuda.empty_cache()

    progress("e2e done")
    # total compressed size of the global stream (all-gather of 8 x u64)
    sizes, my_off = gather_compressed_sizes(head["C"])

    also = {"slots_" + HEADLINE: summarize(slots, world)}
    for k, r in others.items():
        also[k] = summarize(r, world)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # the per-GPU shard of the 8-GPU run (262 144 blocks = 1 GiB) and BASELINE configs[1] (65 536 blocks) in the same run
        for nb_cfg, tag in ((262144, "shard_262144_blocks_"), (65536, "cfg2_65536_blocks_")):
            if nb_cfg < B:
                for k in VARIANTS:
                    also[tag + k] = summarize(bench_variant(k, data[:nb_cfg], max(5, args.steps // 2), 3)[0], 1)
        # BASELINE configs[2] (tANS, same 65 536 x 4 KiB batch) and configs[3] (arithmetic coder, adaptive order-0 model,
        # 1 048 576 x 1 KiB) through the same drop-in classes: informative lines, the headline stays rANS
        also.update(bench_other_configs(data[:262144], freqs, peak, dev))
        from oracle import scl_oracle as so

        so.build()
        threads = host_threads()
        nb = min(B, 1024 * threads)
        progress("cpu baseline (C port, %d threads)" % threads)
        cpu = cpu_baseline(fl, VARIANTS[HEADLINE], data[:nb].cpu().numpy(), threads, HEADLINE)
        if not args.no_python_reference:
            # the reference's own Python coder on the same batch, its bits compared with the GPU's
            progress("python reference leg")
            n_py = min(B, 2 * threads)
            rows = data[:n_py].cpu().numpy()
            py, streams = python_reference_leg(fl, VARIANTS[HEADLINE], [r.tolist() for r in rows], threads)
            if streams is not None:
                e = enc.encode_blocks_packed(data[:n_py]).check()
                got = gather_streams(e, list(range(len(streams))))
                py["gpu_bits_equal"] = all(g == s for g, s in zip(got, streams))
                py["blocks_compared"] = len(streams)
                assert py["gpu_bits_equal"], "GPU bitstream differs from the Python reference's"
            cpu["python_reference"] = py

    progress("done")
    if rank == 0:
        line = {
            "metric": METRIC, "value": hs["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": hs["ms_per_step"], "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {
                "workload": workload_text(total_blocks, world, args.scaling),
                "step": "encode: symbols in -> contiguous reference stream + offsets out (one fused launch); decode: that stream in -> symbols out",
                "params": HEADLINE, "blocks_per_gpu": B, "global_blocks": total_blocks, "block_len": N,
                "parallelism": "block-range sharding, %d rank(s), no payload collective" % world,
                "l2": "inputs (%.2f GiB raw, ~%.2f GiB coded per GPU) exceed the 126 MB L2; no explicit flush" % (B * N / 2**30, head["C"] / 2**30),
            },
            "encode_MBps": hs["encode_MBps"], "decode_MBps": hs["decode_MBps"], "bits_per_symbol": hs["bits_per_symbol"],
            "kernel_paths": hs["kernel_paths"], "roofline": roofline, "parity": parity, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": 2 * args.steps,
            "clocks": head["clocks"], "also": also, "global_compressed_bytes": int(sum(sizes)),
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
