"""Differential run of the realigner's fast pass: GPU kernel (mpn_fastpass) against the host k-mer index (the reference's algorithm) on seeded
adversarial + config-3 regions.  python tools/fastpass_fuzz.py [seeds] [lib]  -> one JSON line (differences must be 0).
With lib = oracle/_hosttest/realigner_hosttest.so the scalar stand-in of the kernel's formulation is checked instead (no GPU needed)."""
import importlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
w = importlib.import_module("workloads")
R = importlib.import_module("megapath-nano_b200.realigner")
nseeds = int(sys.argv[1]) if len(sys.argv) > 1 else 40
lib = sys.argv[2] if len(sys.argv) > 2 else None
t0 = time.time()
haps = places = placed = dropped = diff = 0
kernel_ms = 0.0
for seed in range(3000, 3000 + nseeds):
    regs = w.fastpass_adversarial(12, seed=seed) + w.config3(3, seed=seed, max_reads=200, max_haps=10, n_frac=0.003 if seed % 2 else 0.0)
    s0, p0, ms = R.fastpass_only(regs, 0, lib)
    s1, p1, _ = R.fastpass_only(regs, 1, lib)
    kernel_ms += ms
    haps += len(s0); places += len(p0) // 2
    placed += sum(1 for x in p1[1::2] if x >= 0); dropped += sum(1 for x in s1 if x == 0)
    diff += sum(a != b for a, b in zip(s0, s1)) + sum(a != b for a, b in zip(p0, p1))
print(json.dumps(dict(seeds=nseeds, regions=15 * nseeds, haplotypes=haps, haplotype_read_pairs=places, placed=placed, haplotypes_with_score_0=dropped,
                      differences=diff, kernel_ms_total=round(kernel_ms, 2), seconds=round(time.time() - t0, 1), against="host k-mer index")))
