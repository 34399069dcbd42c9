"""Kernel-phase timing of one resident batch: python tools/phase_bench.py [pairs] [config] -- prints forward/finish/reverse/trace ms
(CUDA events inside the engine) and forward-pass GCUPS.  MPN_SSW_LIB selects a kernel-variant build for A/B runs."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
w = importlib.import_module("workloads")
B = importlib.import_module("megapath-nano_b200.batch")
pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 300000
cfg = int(sys.argv[2]) if len(sys.argv) > 2 else 2
flag = int(sys.argv[3]) if len(sys.argv) > 3 else None
b = {1: w.config1, 2: w.config2, 4: w.config4, 5: w.config5}[cfg](pairs, seed=1000)
if flag is not None:
    b.flag = flag
eng = B.Engine(0)
eng.set_profile(True)
h = eng.upload(b)
best = None
for it in range(5 if cfg < 4 else 3):
    eng.run(h)
    ph = eng.phase_ms()
    if it >= (2 if cfg < 4 else 1) and (best is None or ph["forward"] < best["forward"]):
        best = ph
tot = sum(best.values())
print(f"{os.environ.get('MPN_SSW_LIB', 'default'):28s} pairs {pairs} fwd {best['forward']:.3f} fin {best['finish']:.3f} rev {best['reverse']:.3f} tr {best['trace']:.3f} total {tot:.3f} ms"
      f"  fwd {b.cells / best['forward'] / 1e6:.0f} GCUPS  all {b.cells / tot / 1e6:.0f} GCUPS", flush=True)
eng.free(h)
