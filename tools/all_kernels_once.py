"""Small mixed workload in which every kernel of the engine runs at least once (short / long / clamped reads, all traceback kernels,
one realigner region): a quick target for a debugger or a profiler capture.  python tools/all_kernels_once.py"""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
w = importlib.import_module("workloads")
B = importlib.import_module("megapath-nano_b200.batch")
R = importlib.import_module("megapath-nano_b200.realigner")
eng = B.Engine(0)
n = 0
for b in (w.config2(300, seed=3), w.fuzz_pairs(120, 5, flag=1), w.fuzz_pairs(80, 6, alphabet=2, flag=0x0f),
          w.make_pairs(6, (2500, 3500), 1.2, err=0.08, seed=9, flag=1, chunk=8), w.make_pairs(3, 2600, 3000, err=0.05, seed=10, flag=0, chunk=8)):
    rec, cig = eng.align(b, cigar_cap=int(b.read_len.sum() + b.ref_len.sum()))
    n += b.npairs
    assert (rec["status"] == 0).all()
rg = w.config3(1, seed=3, max_reads=60, max_haps=4)[0]
R.realign_reads(rg)
print("all_kernels_once: ran", n, "pairs + 1 region")
