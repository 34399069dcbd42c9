"""Cost of reads that contain N (code 4) in a config-2 shaped batch: python tools/n_reads_bench.py [pairs]
Three arms per N density: no N at all, N with the reference's matrix (constant N column: the packed kernel's N mode), N with a matrix
whose N column varies (mat[4][4] changed: such reads are flagged and redone by the int32 kernel)."""
import importlib, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
w = importlib.import_module("workloads")
B = importlib.import_module("megapath-nano_b200.batch")
pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
eng = B.Engine(0); eng.set_profile(True)
for n_frac in (0.0, 5e-5, 4e-4):
    b = w.make_pairs(pairs, (150, 300), 1000, err=0.02, seed=5, flag=1, mask="half", n_frac=n_frac)
    hasn = np.add.reduceat((b.reads == 4).astype(np.int32), b.read_off[:-1]) > 0
    for arm in ("constant N column", "varying N column"):
        if arm.startswith("varying"):
            if n_frac == 0.0:
                continue
            b.mat = b.mat.copy(); b.mat[24] = -5
        h = eng.upload(b)
        best = None
        for it in range(5):
            eng.run(h); ph = eng.phase_ms()
            if it >= 2 and (best is None or sum(ph.values()) < sum(best.values())):
                best = ph
        eng.free(h)
        print("reads with N: %5.2f %%  %-18s" % (100 * hasn.mean(), arm), {k: round(v, 2) for k, v in best.items()}, "total", round(sum(best.values()), 2),
              "GCUPS", round(b.cells / sum(best.values()) / 1e6), flush=True)
