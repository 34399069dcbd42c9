"""mpn_pool on every visible device: one config-2 batch and one mixed-length batch through B.Pool against one engine.
    python tests/harness/pool_bench.py [pairs]      (on a multi-GPU box: gpurun --gpus N)"""
import importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
w = importlib.import_module("workloads")
B = importlib.import_module("megapath-nano_b200.batch")
pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
ndev = torch.cuda.device_count()
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
out = {"devices": ndev}
for name, b in (("config2", w.config2(pairs, seed=1000)), ("mixed", w.MixedStream(60_000, seed=15, chunk_cost=1e18, max_pairs=60_000, flag=1).chunk(0, threads=16))):
    hb = type("HostBatch", (), {})()
    for k in ("mat", "n", "gapO", "gapE", "flag", "filters", "filterd", "score_size", "name"):
        setattr(hb, k, getattr(b, k))
    hb.npairs = b.npairs
    hb.reads, hb.read_off, hb.refs, hb.ref_off, hb.masklen = pin(b.reads), pin(b.read_off), pin(b.refs), pin(b.ref_off), pin(b.masklen)
    cap = int(b.read_len.sum() + b.ref_len.sum()) // 2 + 64 * b.npairs
    eng = B.Engine(0)
    rec, cig = eng.align(hb, cigar_cap=cap)
    t0 = time.perf_counter(); rec, cig = eng.align(hb, cigar_cap=cap, out=rec, cig=cig); t1 = time.perf_counter() - t0
    want = B.as_table(rec[::97], cig, 512)
    eng.close()
    res = {"pairs": b.npairs, "cells": b.cells, "one_engine_s": t1, "one_engine_gcups": b.cells / t1 / 1e9}
    for nd in sorted({1, 2, 4, 8, ndev}):
        if nd > ndev:
            continue
        pool = B.Pool(nd)
        prec, pcig = pool.align(hb, cigar_cap=cap)
        t0 = time.perf_counter(); prec, pcig = pool.align(hb, cigar_cap=cap, out=prec, cig=pcig); tp = time.perf_counter() - t0
        got = B.as_table(prec[::97], pcig, 512)
        res[f"pool{nd}"] = {"s": tp, "gcups": b.cells / tp / 1e9, "identical_sample": bool((got[0] == want[0]).all() and (got[1] == want[1]).all()), "shares": pool.last_shares()}
        pool.close()
    out[name] = res
print(json.dumps(out))
