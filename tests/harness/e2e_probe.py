"""End-to-end time of mpn_align_batch_packed4 on BASELINE configs[1] under different pipeline settings (environment switches of
engine.cu are read once per process, so every variant runs in a child process on the same saved workload).
usage: python tests/harness/e2e_probe.py [npairs] > gpurun_out/e2e_probe.json"""
import importlib
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
VARIANTS = [
    ("default", {}),
    ("naux5", {"MPN_NAUX": "5"}),
    ("naux8", {"MPN_NAUX": "8"}),
    ("naux2", {"MPN_NAUX": "2"}),
    ("depth6_naux5", {"MPN_PIPE_DEPTH": "6", "MPN_NAUX": "5"}),
    ("chunk256k_naux5", {"MPN_CHUNK_PAIRS": "262144", "MPN_NAUX": "5"}),
    ("no_pipe_fork", {"MPN_NO_PIPE_FORK": "1"}),
]


def child(path):
    import torch
    B = importlib.import_module("megapath-nano_b200.batch")
    w = importlib.import_module("workloads")
    z = np.load(path)
    b = w.PairBatch(z["reads"], z["read_off"], z["refs"], z["ref_off"], z["masklen"], flag=1)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    hb = type("HostBatch", (), {})()
    for k in ("mat", "n", "gapO", "gapE", "flag", "filters", "filterd", "score_size", "name"):
        setattr(hb, k, getattr(b, k))
    hb.npairs = b.npairs
    hb.reads, hb.read_off, hb.refs, hb.ref_off, hb.masklen = None, pin(b.read_off), None, pin(b.ref_off), pin(b.masklen)
    r4, f4 = pin(B.pack4(b.reads)), pin(B.pack4(b.refs))
    cigar_cap = b.npairs * 24 + len(b.reads) // 4 + 4096
    out = torch.zeros(b.npairs * B.RESULT_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    cig = torch.zeros(cigar_cap, dtype=torch.int32).pin_memory()
    eng = B.Engine(0)
    for _ in range(2):
        eng.align_packed4(hb, r4, f4, cigar_cap=cigar_cap, out=out, cig=cig)
    ts = []
    for _ in range(6):
        t0 = time.perf_counter()
        eng.align_packed4(hb, r4, f4, cigar_cap=cigar_cap, out=out, cig=cig)
        ts.append(1e3 * (time.perf_counter() - t0))
    recs = np.frombuffer(out.numpy(), dtype=B.RESULT_DTYPE)
    print(json.dumps({"ms": ts, "median_ms": float(np.median(ts)), "cells": int(b.cells), "checksum": int(recs["score1"].astype(np.int64).sum() + recs["ref_begin1"].astype(np.int64).sum())}))


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        return child(sys.argv[2])
    npairs = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    w = importlib.import_module("workloads")
    b = w.config2(npairs)
    path = "/dev/shm/e2e_probe.npz"
    np.savez(path, reads=b.reads, read_off=b.read_off, refs=b.refs, ref_off=b.ref_off, masklen=b.masklen)
    res = {}
    for name, env in VARIANTS:
        e = dict(os.environ); e.update(env)
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", path], env=e, capture_output=True, text=True)
        line = [l for l in p.stdout.splitlines() if l.startswith("{")]
        res[name] = json.loads(line[-1]) if line else {"error": p.stderr[-2000:]}
        if name == "timing":
            res[name]["stderr_tail"] = [l for l in p.stderr.splitlines() if "mpn_ssw" in l][-8:]
        if "median_ms" in res[name]:
            res[name]["gcups"] = res[name]["cells"] / res[name]["median_ms"] / 1e6
    print(json.dumps(res, indent=1))
    os.remove(path)


if __name__ == "__main__":
    main()
