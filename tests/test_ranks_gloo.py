"""world_size-2 gloo tests (CPU) of the N>1 plumbing: order-preserving read shards, max/sum reductions,
ordered gather -- the only things ranks exchange on this path (there is no data-path collective).  The
per-read work in the test is the CPU oracle standing in for the device stages."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

import helpers as H
from sigfish_b200 import build as B
from sigfish_b200 import ranks, synth


def test_shard_range_is_a_partition():
    for n in (0, 1, 7, 8, 100, 4097):
        for w in (1, 2, 3, 8):
            parts = [ranks.shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [e - b for b, e in parts]
            assert max(sizes) - min(sizes) <= 1


def test_host_shard_ranges_equal_weight():
    """sf_shard_ranges: contiguous ranges of about the same total weight (the pipeline passes the estimated device work
    of every read: q x reference columns + 640 cell-times per sample)"""
    B.build_all()
    L = C.CDLL(B.LIB_HOST)
    L.sf_shard_ranges.argtypes = [C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]
    rng = np.random.default_rng(0)
    for n, g in ((0, 2), (1, 4), (5, 2), (48, 8), (512, 3), (3, 8)):
        lens = rng.integers(1, 9000, size=max(n, 1)).astype(np.int64)
        begin = np.zeros(g + 1, dtype=np.int32)
        L.sf_shard_ranges(n, lens.ctypes.data_as(C.c_void_p), g, begin.ctypes.data_as(C.c_void_p))
        assert begin[0] == 0 and begin[g] == n and np.all(np.diff(begin) >= 0)
        if n >= 8 * g:
            tot = lens[:n].sum()
            per = [lens[begin[i]:begin[i + 1]].sum() for i in range(g)]
            assert max(per) <= tot / g + 9000


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    sys.path.insert(0, os.path.join(H.ROOT, "tests"))
    R = ranks.Ranks(backend="gloo")
    k = 6
    mean, _ = synth.make_model(k)
    rng = np.random.default_rng(12)
    seqs = [synth.random_sequence(4000, rng)]
    sigs, _ = synth.simulate_reads(seqs, k, mean, 9, seed=5, bases_per_read=400)  # same on every rank
    b, e = ranks.shard_range(len(sigs), R.rank, R.world)
    ref = H.OracleRef(seqs, mean, k, 0, 250)
    rows = []
    for i in range(b, e):
        h = H.orc_map(ref, sigs[i], 8192.0, 10.0, 1402.882, 0, 250, 50)
        rows.append((i, h.rid, h.strand.decode(), h.pos_st, h.pos_end, float(h.score)))
    R.barrier()
    R.host_barrier()  # the CPU-side barrier bench.py uses around its from-files stage
    mx = R.max([10.0 + R.rank, 3.0])
    sm = R.sum([float(e - b), 250.0 * (e - b)])
    allrows = R.gather_ordered(rows)
    if R.rank == 0:
        q.put((mx, sm, allrows))
    R.barrier()
    R.close()


def test_world_size_2_gloo_sharded_job_matches_single_rank():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    mx, sm, rows = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert mx == [11.0, 3.0]
    assert sm == [9.0, 2250.0]
    assert [r[0] for r in rows] == list(range(9))  # input order survives the gather
    # the same job on one rank
    k = 6
    mean, _ = synth.make_model(k)
    rng = np.random.default_rng(12)
    seqs = [synth.random_sequence(4000, rng)]
    sigs, _ = synth.simulate_reads(seqs, k, mean, 9, seed=5, bases_per_read=400)
    ref = H.OracleRef(seqs, mean, k, 0, 250)
    for i, row in enumerate(rows):
        h = H.orc_map(ref, sigs[i], 8192.0, 10.0, 1402.882, 0, 250, 50)
        assert row == (i, h.rid, h.strand.decode(), h.pos_st, h.pos_end, float(h.score))
    assert ranks.job_throughput(2.0e12, 250.0) == pytest.approx(8000.0)
