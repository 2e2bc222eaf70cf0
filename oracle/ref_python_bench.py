"""The reference's own pure-Python rANS coder, timed on the host cores -- BENCH INFRASTRUCTURE ONLY.

bench.py's `cpu_baseline.python_reference` leg (north_star: "the reference's pure-Python path timed on the
GPU box's own host cores in the same run"): a multiprocessing.Pool with one worker
per core, each worker encodes and decodes K blocks of the SAME batch the GPU coded through the UNMODIFIED
`scl.compressors.rANS.rANSEncoder / rANSDecoder` (/root/reference/scl/compressors/rANS.py:123-297, imported
by oracle/ref_loader.py -- from the staged copy on the GPU box) on the shim-backed BitArray (the real
`bitarray` wheel is not in the image, so container overhead differs from upstream; the arithmetic is the
reference's).  Returns the coded bytes so that the caller can compare them with the GPU's.
"""
import os
import sys
import time

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(job):
    freq_list, kw, rows = job
    if _ROOT not in sys.path:
        sys.path.insert(0, _ROOT)
    from oracle import ref_loader

    scl = ref_loader.import_reference()
    from scl.compressors.rANS import rANSDecoder, rANSEncoder, rANSParams
    from scl.core.data_block import DataBlock
    from scl.core.prob_dist import Frequencies

    params = rANSParams(Frequencies({i: int(f) for i, f in enumerate(freq_list)}), **kw)
    enc, dec = rANSEncoder(params), rANSDecoder(params)
    out = []
    t_enc = t_dec = 0.0
    for row in rows:
        block = DataBlock(list(row))
        t0 = time.perf_counter()
        ba = enc.encode_block(block)
        t1 = time.perf_counter()
        decoded, used = dec.decode_block(ba)
        t2 = time.perf_counter()
        assert decoded.data_list == block.data_list and used == len(ba)
        t_enc += t1 - t0
        t_dec += t2 - t1
        out.append((ba.tobytes(), len(ba)))
    return out, t_enc, t_dec


def run(freq_list, kw, blocks, n_workers, blocks_per_worker):
    """blocks: list of lists of ints (rows of the batch).  Encodes + decodes n_workers * blocks_per_worker of
    them.  Returns dict(streams=[(bytes, nbits)...] in input order, wall_s, enc_cpu_s, dec_cpu_s, n_blocks)."""
    import multiprocessing as mp

    n = min(len(blocks), n_workers * blocks_per_worker)
    jobs = [(list(freq_list), dict(kw), blocks[i : i + blocks_per_worker]) for i in range(0, n, blocks_per_worker)]
    ctx = mp.get_context("spawn")  # the parent holds a CUDA context: never fork it
    with ctx.Pool(min(n_workers, len(jobs))) as pool:
        pool.map(_warm, range(min(n_workers, len(jobs))))  # interpreter start-up and imports are not coder time
        t0 = time.perf_counter()
        res = pool.map(_worker, jobs, chunksize=1)
        wall = time.perf_counter() - t0
    streams = [s for r in res for s in r[0]]
    return dict(streams=streams, wall_s=wall, enc_cpu_s=sum(r[1] for r in res), dec_cpu_s=sum(r[2] for r in res), n_blocks=n,
                workers=min(n_workers, len(jobs)))


def _warm(_):
    if _ROOT not in sys.path:
        sys.path.insert(0, _ROOT)
    from oracle import ref_loader

    ref_loader.import_reference()
    import scl.compressors.rANS  # noqa: F401
    import numpy  # noqa: F401

    return os.getpid()
