import importlib, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
w = importlib.import_module("megapath-nano_b200.workloads")
B = importlib.import_module("megapath-nano_b200.batch")
b = w.make_pairs(400000, (150, 300), 1000, err=0.02, seed=5, flag=1, mask="half", n_frac=5e-5)
import numpy as np
ro = b.read_off
hasn = np.add.reduceat((b.reads == 4).astype(np.int32), ro[:-1]) > 0
eng = B.Engine(0); eng.set_profile(True); h = eng.upload(b)
for it in range(4):
    eng.run(h); ph = eng.phase_ms()
print("reads with N: %.2f %%" % (100 * hasn.mean()), {k: round(v, 2) for k, v in ph.items()}, "total", round(sum(ph.values()), 2), "GCUPS", round(b.cells / sum(ph.values()) / 1e6))
