"""Ad-hoc GPU parity run: python tests/harness/gpu_check.py [nseeds] -- compares the CUDA engine with the compiled reference (oracle/_ref)."""
import importlib, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
w = importlib.import_module("workloads")
B = importlib.import_module("megapath-nano_b200.batch")
from oracle import oracle

def compare(b, eng, label, threads=16, cap=256):
    impl = "ref" if oracle.have_ref() else "port"
    r, c, _ = oracle.run_batch(b.reads, b.read_off, b.refs, b.ref_off, b.masklen, b.mat, b.n, gapO=b.gapO, gapE=b.gapE, flag=b.flag,
                               filters=b.filters, filterd=b.filterd, score_size=b.score_size, threads=threads, impl=impl, cigar_cap=cap)
    t0 = time.time()
    rec, cig = eng.align(b)
    dt = time.time() - t0
    g, gc = B.as_table(rec, cig, cap)
    neq = (r != g).any(axis=1) | (c != gc).any(axis=1)
    nb = int(neq.sum())
    print(f"{label}: pairs {b.npairs} mismatches {nb} gpu_e2e {dt*1e3:.1f} ms  ({b.cells/dt/1e9:.1f} GCUPS e2e)", flush=True)
    if nb:
        for i in np.nonzero(neq)[0][:6]:
            print("   pair", i, "rl", b.read_len[i], "fl", b.ref_len[i], "mask", b.masklen[i], "\n     ref", r[i], c[i][:8], "\n     gpu", g[i], gc[i][:8], "status", rec["status"][i])
    return nb

if __name__ == "__main__":
    nseeds = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    eng = B.Engine()
    bad = 0
    b = w.config1(2000); bad += compare(b, eng, "config1 x2000 flag0")
    b = w.config1(2000); b.flag = 8; bad += compare(b, eng, "config1 x2000 flag8")
    b = w.config2(3000); bad += compare(b, eng, "config2 x3000 flag1")
    for seed in range(nseeds):
        alpha = 2 if seed % 3 == 0 else 4
        flag = [1, 0, 8, 0x0f, 2, 4][seed % 6]
        b = w.fuzz_pairs(300, seed, alphabet=alpha, flag=flag)
        b.filters = 100 if flag == 2 else 0
        b.filterd = 30 if flag == 4 else 32767
        bad += compare(b, eng, f"fuzz seed {seed} alpha {alpha} flag {flag} gap {b.gapO}/{b.gapE}")
    print("stats", eng.stats())
    print("TOTAL MISMATCHES", bad)
    sys.exit(1 if bad else 0)
