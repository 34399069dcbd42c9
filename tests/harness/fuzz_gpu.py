"""Large differential run on the GPU box: seeded adversarial batches (workloads.fuzz_pairs + long-read mixes) through the C ABI against
the compiled reference ssw.c.  python tests/harness/fuzz_gpu.py [seeds] [pairs_per_seed]  -> one JSON line (mismatches must be 0)."""
import importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
w = importlib.import_module("workloads")
B = importlib.import_module("megapath-nano_b200.batch")
from oracle import oracle
nseeds = int(sys.argv[1]) if len(sys.argv) > 1 else 40
per = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
eng = B.Engine(0)
impl = "ref" if oracle.have_ref() else "port"
tot = bad = 0
t0 = time.time()
detail = []
for s in range(nseeds):
    flag = [1, 0, 8, 0x0f, 2, 4, 3, 5][s % 8]
    if s % 5 == 4:      # long / mixed-length batch: multi-strip kernel, clamp, warp traceback
        b = w.make_pairs(max(8, per // 40), (900, 6000), 1.15, err=0.06, seed=7000 + s, flag=1 if s % 2 else 0, chunk=16)
        b.mat = w.dna_matrix(int(3 + s % 3), 5)
    else:
        b = w.fuzz_pairs(per, 9000 + s, alphabet=2 if s % 3 == 0 else 4, flag=flag, max_read=700 if s % 4 else 1500, max_ref=500 if s % 4 else 1800)
        b.filters = 90 if flag in (2, 3) else 0
        b.filterd = 35 if flag in (4, 5) else 32767
    cap = 2048
    rec, cig = eng.align(b, cigar_cap=int(b.read_len.sum() + b.ref_len.sum() + 16 * b.npairs))
    g, gc = B.as_table(rec, cig, cap)
    r, c, _ = oracle.run_batch(b.reads, b.read_off, b.refs, b.ref_off, b.masklen, b.mat, b.n, gapO=b.gapO, gapE=b.gapE, flag=b.flag, filters=b.filters,
                               filterd=b.filterd, score_size=b.score_size, threads=os.cpu_count() or 8, impl=impl, cigar_cap=cap)
    nb = int(((r != g).any(axis=1) | (c != gc).any(axis=1)).sum())
    tot += b.npairs; bad += nb
    if nb:
        detail.append({"seed": s, "bad": nb})
print(json.dumps({"seeds": nseeds, "pairs": tot, "mismatches": bad, "oracle": impl, "seconds": round(time.time() - t0, 1), "detail": detail[:10]}))
