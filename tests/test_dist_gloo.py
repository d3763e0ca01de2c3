"""world_size-2 gloo test of the N>1 plumbing (runs on CPU): chunk sharding, gather of the
variable-length candidate records, histogram / FAS all-reduce, CCX row-block gather."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from detex_b200 import parallel
from detex_b200._lib import CAND_DTYPE


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        nchunks, S = 7, 3
        lo, hi = parallel.shard_range(nchunks, rank, world)
        # every rank "detects" on its own chunks: deterministic fake candidates per chunk
        recs = []
        for c in range(lo, hi):
            rng = np.random.default_rng(100 + c)
            k = int(rng.integers(0, 5))
            r = np.zeros(k, dtype=CAND_DTYPE)
            r["row"] = c * S + rng.integers(0, S, size=k)
            r["t"] = rng.integers(0, 1000, size=k)
            r["ds"] = rng.uniform(0.3, 1, size=k)
            recs.append(r)
        local = np.concatenate(recs) if recs else np.zeros(0, dtype=CAND_DTYPE)
        allc = parallel.gather_records(local)
        hist = np.zeros((S, 400), dtype=np.int64)
        for c in range(lo, hi):
            hist[c % S, c] += c + 1
        hist = parallel.allreduce_sum(hist)
        fasst = parallel.allreduce_sum(np.full((S, 5), float(rank + 1)))
        N = 9
        blocks = parallel.ccx_row_blocks(N, world)
        b0, b1 = blocks[rank]
        blk = np.arange(b0, b1, dtype=np.float64)[:, None] * np.ones((1, N))
        full = parallel.gather_row_blocks(blk, blocks, N)
        if rank == 0:
            q.put((allc, hist, fasst, full))
    finally:
        dist.destroy_process_group()


def test_two_rank_gather_equals_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    allc, hist, fasst, full = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process expectation
    S = 3
    exp = []
    for c in range(7):
        rng = np.random.default_rng(100 + c)
        k = int(rng.integers(0, 5))
        r = np.zeros(k, dtype=CAND_DTYPE)
        r["row"] = c * S + rng.integers(0, S, size=k)
        r["t"] = rng.integers(0, 1000, size=k)
        r["ds"] = rng.uniform(0.3, 1, size=k)
        exp.append(r)
    exp = np.concatenate(exp)
    assert np.array_equal(allc, exp)
    eh = np.zeros((S, 400), dtype=np.int64)
    for c in range(7):
        eh[c % S, c] += c + 1
    assert np.array_equal(hist, eh)
    assert np.all(fasst == 3.0)
    assert np.array_equal(full[:, 0], np.arange(8, dtype=np.float64))


def _detect_worker(rank, world, port, q):
    """One rank of a chunk-sharded detection run: the host path of bench.py / parallel.py with the
    oracle-backed engine stand-in in place of the GPU (test infrastructure, tests/oracle_engine.py)."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from oracle_engine import OracleEngine
    from detex_b200 import synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        chunks, bases, _ = synth.detection_case(91, 5, 2500, 100, 3, [1, 3, 2], planted=2)
        thr = [0.3, 0.3, 0.3]
        S = len(bases)
        lo, hi = parallel.shard_range(len(chunks), rank, world)
        eng = OracleEngine()
        eng.set_bases(0, bases, 3, thresholds=thr)
        eng.load_chunks(chunks[lo:hi])
        eng.detect_run(0, lta_window=50, want_fas=True)
        c = eng.candidates()
        c["row"] += lo * S                               # local chunk index -> global (bench.py)
        mx, _ = eng.rowstats()
        hist, fas = eng.hist(0), eng.fas(0)
        if world > 1:
            c = parallel.gather_records(c)
            hist = parallel.allreduce_sum(hist)
            fas = parallel.allreduce_sum(fas)
            best = parallel.allreduce_max(mx.max(axis=0).astype(np.float64))
        else:
            best = mx.max(axis=0).astype(np.float64)
        if rank == 0:
            order = np.lexsort((c["t"], c["row"]))
            q.put((c[order], hist, fas, best))
    finally:
        if world > 1:
            dist.destroy_process_group()


def test_sharded_detection_equals_single_process():
    ctx = mp.get_context("spawn")
    out = {}
    for world in (1, 2):
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_detect_worker, args=(r, world, port, q)) for r in range(world)]
        for p in procs:
            p.start()
        out[world] = q.get(timeout=300)
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
    (c1, h1, f1, b1), (c2, h2, f2, b2) = out[1], out[2]
    assert len(c1) > 0 and np.array_equal(c1, c2)        # same triggers, bit for bit
    assert np.array_equal(h1, h2) and h1.sum() == 5 * 3 * 2401
    assert np.allclose(f1, f2, rtol=1e-12) and np.array_equal(b1, b2)


def _ccx_worker(rank, world, port, q):
    """One rank of the dealt-row CCX (parallel.ccx_sharded) with the oracle-backed engine stand-in."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from oracle_engine import OracleEngine
    from detex_b200 import synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        X = synth.event_families(77, 3, 6, 40, 3, max_shift=8)        # 18 events, n = 120
        res = parallel.ccx_sharded(OracleEngine(), X, 3)
        only_root = parallel.ccx_sharded(OracleEngine(), X, 3, root=0)   # gather to the rank that clusters
        assert (only_root is None) == (rank != 0)
        if rank == 0:
            assert all(np.array_equal(a, b) for a, b in zip(res, only_root))
        # no collective at all: every rank writes its rows into one shared host matrix (parallel.CcxHostBuffer)
        buf = parallel.CcxHostBuffer(OracleEngine(), len(X))
        shared = parallel.ccx_sharded(OracleEngine(), X, 3, host=buf)
        assert all(np.array_equal(a, b) for a, b in zip(res, shared))
        assert shared[0] is buf.cc and buf.lag.dtype == np.int32
        dist.barrier()
        buf.close()
        if rank == world - 1:
            q.put(res)
    finally:
        dist.destroy_process_group()


def test_dealt_rows_are_balanced_and_complete():
    for N in (9, 50, 600, 4096, 16384):
        for world in (2, 4, 8):
            deal = parallel.ccx_deal_rows(N, world)
            assert sorted(np.concatenate(deal).tolist()) == list(range(N - 1))
            assert all(np.all(np.diff(r) > 0) for r in deal if len(r) > 1)
            rows = [len(r) for r in deal]
            pairs = [int((N - 1 - r.astype(np.int64)).sum()) for r in deal]
            assert max(rows) - min(rows) <= 2
            assert max(pairs) - min(pairs) <= 2 * N
            slots, nmax = parallel.ccx_slot_rows(N, world)
            assert len(slots) == world * nmax and sorted(slots[slots >= 0].tolist()) == list(range(N - 1))


def test_eight_rank_dealt_ccx_equals_single_process():
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from detex_b200 import synth
    from oracle import detex_oracle as orc
    world = 8
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ccx_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    cc, lag, sub = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    X = synth.event_families(77, 3, 6, 40, 3, max_shift=8)
    rcc, rlag, rsub = orc.make_cclags(X, 3)
    iu = np.triu_indices(len(X), 1)
    assert np.array_equal(cc, rcc[iu[0], iu[1] - 1])
    assert np.array_equal(lag.astype(float), rlag[iu[0], iu[1] - 1])
    assert np.array_equal(sub, rsub[iu[0], iu[1] - 1])


def _long_worker(rank, world, port, q):
    """One rank of a time-segment-sharded detection on a long array (SSDetex.run_long_array)."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from oracle_engine import OracleEngine
    from detex_b200 import detect, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        chunks, bases, _ = synth.detection_case(95, 1, 7000, 100, 3, [2, 3], planted=3)
        names = ["SS0", "SS1"]
        det = detect.SSDetex(dict(zip(names, bases)), {n: 0.3 for n in names}, {n: [0.0, 1.0] for n in names}, 3,
                             engine=OracleEngine(), triggerLTATime=0.5)
        df, mx = det.run_long_array(chunks[0], 100.0, 50.0, seg_lags=1000, batch=2,
                                    shard=(rank, world) if world > 1 else None)
        if rank == 0:
            q.put((df.STMP.values, df.DS.values, df.DS_STALTA.values, list(df.Name), mx,
                   {k: v.copy() for k, v in det.histdic.items()}))
    finally:
        if world > 1:
            dist.destroy_process_group()


def test_time_segment_sharding_two_ranks_equals_one():
    ctx = mp.get_context("spawn")
    out = {}
    for world in (1, 2):
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_long_worker, args=(r, world, port, q)) for r in range(world)]
        for p in procs:
            p.start()
        out[world] = q.get(timeout=300)
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
    a, b = out[1], out[2]
    assert len(a[0]) > 0 and np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[3] == b[3]
    assert np.array_equal(a[2], b[2]) and a[4] == b[4]
    assert all(np.array_equal(a[5][k], b[5][k]) for k in a[5])


def _hostbuf_fail_worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from oracle_engine import OracleEngine
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        outcome = "created"
        try:
            parallel.CcxHostBuffer(OracleEngine(), 50, shm_dir="/nonexistent/shm")
        except RuntimeError:
            outcome = "refused"
        dist.barrier()                      # nobody is stuck in a collective of the failed constructor
        q.put((rank, outcome))
    finally:
        dist.destroy_process_group()


def test_host_buffer_is_refused_by_every_rank_when_the_segment_cannot_be_made():
    """A tmpfs that cannot hold the matrix (here: a directory that does not exist) makes EVERY rank raise
    RuntimeError -- callers fall back to the NCCL gather (bench.py does) -- instead of one rank raising and the
    others waiting in a collective."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_hostbuf_fail_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == [(0, "refused"), (1, "refused")]
