"""Phase split (forward / finish / reverse / trace) of the Smith-Waterman batch of ONE realigner region: the H haplotype-vs-reference
pairs plus reads x haplotypes, flag 0x0f, as realign_region.cpp submits them.  python tools/region_phase.py [seed]"""
import importlib, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
w = importlib.import_module("workloads")
B = importlib.import_module("megapath-nano_b200.batch")
rg = w.config3(3, seed=int(sys.argv[1]) if len(sys.argv) > 1 else 13)[2]
enc = lambda s: np.frombuffer(s.encode(), dtype=np.uint8)
tr = np.full(256, 4, np.int8)
for ch, v in zip("ACGT", range(4)):
    tr[ord(ch)] = v
reads, refs = [], []
for h in rg.haplotypes:
    reads.append(tr[enc(h)]); refs.append(tr[enc(rg.reference)])
for r in rg.reads[::2]:
    for h in rg.haplotypes:
        reads.append(tr[enc(r)]); refs.append(tr[enc(h)])
ro = np.concatenate([[0], np.cumsum([len(x) for x in reads])]).astype(np.int64)
fo = np.concatenate([[0], np.cumsum([len(x) for x in refs])]).astype(np.int64)
b = w.PairBatch(np.concatenate(reads), ro, np.concatenate(refs), fo, np.diff(ro).astype(np.int32), flag=0x0f)
eng = B.Engine(0); eng.set_profile(True)
h = eng.upload(b)
for it in range(4):
    eng.run(h); ph = eng.phase_ms()
print(f"region: {len(rg.haplotypes)} haplotypes of {len(rg.haplotypes[0])} bp, {b.npairs} pairs, phases ms:", {k: round(v, 3) for k, v in ph.items()}, "total", round(sum(ph.values()), 3))
