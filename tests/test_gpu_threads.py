"""The per-pair ABI is called from several threads at once (the reference's ssw.c is re-entrant): the
replacement serialises on one process-wide engine and must return the same records as single-threaded calls."""
import ctypes as ct
import importlib
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

w = importlib.import_module("workloads")
pyssw = importlib.import_module("megapath-nano_b200.pyssw")


def test_concurrent_per_pair_calls():
    rng = np.random.default_rng(5)
    ref = "".join("ACGT"[i] for i in rng.integers(0, 4, size=500))
    queries = []
    for k in range(40):
        s = int(rng.integers(0, 250)); q = list(ref[s:s + 200])
        q[50:50] = list("ACG"[:k % 3]); del q[120:120 + k % 4]
        queries.append("".join(q))
    a = pyssw.SSW(); a.set_reference_sequence(ref)
    want = [a.align(q) for q in queries]
    got = [None] * len(queries)

    def work(lo, hi):
        b = pyssw.SSW(); b.set_reference_sequence(ref)
        for k in range(lo, hi):
            got[k] = b.align(queries[k])

    ts = [threading.Thread(target=work, args=(k * 10, k * 10 + 10)) for k in range(4)]
    [t.start() for t in ts]; [t.join() for t in ts]
    assert got == want
