"""mpn_pool_*: one batch sharded over several devices == the same batch on one engine, pair by pair.
With one visible GPU the pool is built from two engines on that device, which exercises the same queue / range / gather code."""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

w = importlib.import_module("workloads")
B = importlib.import_module("megapath-nano_b200.batch")


def ndev():
    import torch
    return torch.cuda.device_count()


def mixed_batch(seed, flag):
    """short reads of every strip class, a few long / clamped pairs, in shuffled order"""
    parts = [w.make_pairs(3000, (20, 700), 1.3, err=0.03, seed=seed, flag=flag), w.make_pairs(2500, (150, 300), 1000, err=0.02, seed=seed + 1, flag=flag),
             w.make_pairs(40, (1500, 9000), 1.2, err=0.06, seed=seed + 2, flag=flag, chunk=8)]
    reads = np.concatenate([p.reads for p in parts]); refs = np.concatenate([p.refs for p in parts])
    rl = np.concatenate([p.read_len for p in parts]); fl = np.concatenate([p.ref_len for p in parts])
    ro = np.zeros(len(rl) + 1, np.int64); np.cumsum(rl, out=ro[1:])
    fo = np.zeros(len(fl) + 1, np.int64); np.cumsum(fl, out=fo[1:])
    b = w.PairBatch(reads, ro, refs, fo, np.maximum(rl // 2, 15).astype(np.int32), flag=flag, name="mixed")
    return b.subset(np.random.default_rng(seed).permutation(b.npairs))


@pytest.mark.parametrize("flag", [0, 1])
def test_pool_equals_single_engine(flag):
    b = mixed_batch(900 + flag, flag)
    eng = B.Engine(0)
    rec, cig = eng.align(b)
    want, wantc = B.as_table(rec, cig, 256)
    eng.close()
    devs = list(range(min(ndev(), 8))) if ndev() >= 2 else [0, 0]
    for devices in ([0], devs):
        pool = B.Pool(devices)
        prec, pcig = pool.align(b)
        got, gotc = B.as_table(prec, pcig, 256)
        shares = pool.last_shares()
        pool.close()
        assert (got == want).all() and (gotc == wantc).all(), (devices, np.nonzero((got != want).any(axis=1))[0][:5])
        assert sum(s["pairs"] for s in shares) == b.npairs and sum(s["cells"] for s in shares) == b.cells
        if len(devices) > 1:
            assert sum(s["pairs"] > 0 for s in shares) >= 2, shares    # the queue was shared (a small batch has fewer ranges than an 8-GPU box has devices)


def test_pool_spans_form():
    """pairs that share sequences (realigner regions): one range per device, arena uploaded once per device"""
    rng = np.random.default_rng(5)
    haps = [rng.integers(0, 4, size=int(n), dtype=np.int8) for n in rng.integers(300, 900, size=12)]
    reads = [h[o:o + 150].copy() for h in haps for o in rng.integers(0, len(h) - 150, size=40)]
    for r in reads:
        r[rng.integers(0, 150, size=3)] = rng.integers(0, 4, size=3)
    arena = np.concatenate(haps + reads)
    starts = np.concatenate([[0], np.cumsum([len(x) for x in haps + reads])])
    rd_start, rd_len, rf_start, rf_len = [], [], [], []
    for k in range(len(reads)):
        for h in (k // 40, (k // 40 + 1) % len(haps)):
            rd_start.append(starts[len(haps) + k]); rd_len.append(150); rf_start.append(starts[h]); rf_len.append(len(haps[h]))
    proto = w.PairBatch(np.zeros(0, np.int8), np.zeros(1, np.int64), np.zeros(0, np.int8), np.zeros(1, np.int64), np.zeros(0, np.int32), flag=0x0f)
    mask = np.array(rd_len, np.int32)
    eng = B.Engine(0)
    rec, cig = eng.align_spans(proto, arena, rd_start, rd_len, rf_start, rf_len, mask)
    want, wantc = B.as_table(rec, cig, 64)
    eng.close()
    pool = B.Pool(list(range(ndev())) if ndev() >= 2 else [0, 0])
    prec, pcig = pool.align_spans(proto, arena, rd_start, rd_len, rf_start, rf_len, mask)
    got, gotc = B.as_table(prec, pcig, 64)
    pool.close()
    assert (got == want).all() and (gotc == wantc).all()
