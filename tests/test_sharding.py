"""Multi-GPU plumbing on CPU: the path shards by independent pairs with no data-path collective, so the
only distributed pieces are the shard assignment and the max-over-ranks / sum-over-ranks reduction of the timing line.  Both
are exercised here with a world_size-2 gloo group."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w = importlib.import_module("workloads")
    full = w.config2(2000, seed=5)
    mine = full.shard(rank, world)
    t = torch.tensor([10.0 + rank, 20.0 - rank], dtype=torch.float64)
    c = torch.tensor([float(mine.cells), float(mine.npairs)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    ids = np.arange(rank, full.npairs, world)
    ok = (np.array_equal(mine.read_len, full.read_len[ids]) and np.array_equal(mine.masklen, full.masklen[ids])
          and np.array_equal(mine.reads[:50], full.reads[full.read_off[ids[0]]:full.read_off[ids[0]] + 50][:50]))
    q.put((rank, ok, t.tolist(), c.tolist(), full.cells, full.npairs))
    dist.destroy_process_group()


def test_world_size_2_gloo_sharding_and_reduction():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, t, c, cells, npairs in out:
        assert ok
        assert t == [11.0, 20.0]                       # max over ranks
        assert c == [float(cells), float(npairs)]      # shards partition the batch exactly


def test_generators_are_deterministic_and_shaped():
    w = importlib.import_module("workloads")
    a, b = w.config2(500, seed=9), w.config2(500, seed=9)
    assert np.array_equal(a.reads, b.reads) and np.array_equal(a.refs, b.refs)
    assert a.read_len.min() >= 150 and a.read_len.max() <= 300 and (a.ref_len == 1000).all() and a.flag == 1
    c1 = w.config1(100)
    assert (c1.read_len == 250).all() and (c1.ref_len == 500).all() and (c1.masklen == 125).all() and c1.cells == 100 * 250 * 500
    f = w.fuzz_pairs(50, 3)
    assert f.gapO > f.gapE and f.reads.max() <= 4 and f.masklen.min() >= 15


def test_mixed_stream_plan_is_deterministic_balanced_and_chunking_independent():
    """BASELINE configs[4] as a stream (workloads.MixedStream): every rank derives the same plan, the plan covers every chunk once, the
    ranks' cost shares are balanced, and a pair's bases depend only on (seed, global index) -- not on how the stream was chunked."""
    import importlib
    import numpy as np
    w = importlib.import_module("workloads")
    a = w.MixedStream(60_000, seed=15, world=4)
    b = w.MixedStream(60_000, seed=15, world=4)
    assert (a.bounds == b.bounds).all() and a.plan(4) == b.plan(4)
    plan = a.plan(4)
    assert sorted(k for r in plan for k in r) == list(range(a.nchunks))
    share = np.array([a.chunk_cost[r].sum() for r in plan]) / a.chunk_cost.sum()
    assert share.max() - share.min() < 0.06, share
    assert int(a.chunk_cells.sum()) == a.total_cells and (np.diff(a.rl) >= 0).all()          # length-sorted
    c = w.MixedStream(60_000, seed=15, chunk_cost=3.0e10)                                     # other chunking, same stream
    k = a.nchunks // 2
    lo, hi = int(a.bounds[k]), int(a.bounds[k + 1])
    ba = a.chunk(k, threads=2)
    kc = int(np.searchsorted(c.bounds, lo, side="right")) - 1
    bc = c.chunk(kc, threads=3)
    off = lo - int(c.bounds[kc])
    n = min(hi - lo, bc.npairs - off, 50)
    for i in range(n):
        assert (ba.reads[ba.read_off[i]:ba.read_off[i + 1]] == bc.reads[bc.read_off[off + i]:bc.read_off[off + i + 1]]).all()
        assert (ba.refs[ba.ref_off[i]:ba.ref_off[i + 1]] == bc.refs[bc.ref_off[off + i]:bc.ref_off[off + i + 1]]).all()
