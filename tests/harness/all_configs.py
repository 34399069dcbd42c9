"""One JSON line per BASELINE config (1-5): GPU GCUPS (kernel phases with inputs resident + the public call with host buffers), the compiled
reference ssw.c on all host cores over a bounded sample of the SAME pairs, and the parity count of that sample (mismatching pairs / sample).
    python tests/harness/all_configs.py            # on the GPU box; needs oracle/_ref (travels with gpurun) for the reference column
Sizes: config 1 and 3 at full size, config 2 at 1 M pairs, config 4 at 1024 pairs (flag 0 and flag 1), config 5 on a 20 000-pair sample of its
length distribution (the full 10 M pairs are ~4.5e14 cells: generated per chunk in a production run, not held in memory)."""
import dataclasses, importlib, json, os, subprocess, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
w = importlib.import_module("workloads")
B = importlib.import_module("megapath-nano_b200.batch")
R = importlib.import_module("megapath-nano_b200.realigner")
from oracle import oracle
ncores = os.cpu_count() or 8
eng = B.Engine(0)
eng.set_profile(True)


def sw_config(name, b, sample, cap=64):
    t0 = time.perf_counter(); rec, cig = eng.align(b, cigar_cap=int(b.read_len.sum() + b.ref_len.sum()) // 2 + 64 * b.npairs); eng.align(b, cigar_cap=len(cig), out=rec, cig=cig)
    h = eng.upload(b)
    best = None
    for it in range(3):
        eng.run(h); ph = eng.phase_ms()
        if best is None or sum(ph.values()) < sum(best.values()):
            best = ph
    eng.free(h)
    t0 = time.perf_counter(); eng.align(b, cigar_cap=len(cig), out=rec, cig=cig); e2e_s = time.perf_counter() - t0
    idx = np.linspace(0, b.npairs - 1, min(sample, b.npairs)).astype(np.int64)
    sb = b.subset(idx)
    r, c, secs = oracle.run_batch(sb.reads, sb.read_off, sb.refs, sb.ref_off, sb.masklen, sb.mat, sb.n, gapO=sb.gapO, gapE=sb.gapE, flag=sb.flag, filters=sb.filters,
                                  filterd=sb.filterd, score_size=sb.score_size, threads=ncores, impl="ref" if oracle.have_ref() else "port", cigar_cap=cap)
    g, gc = B.as_table(rec[idx], cig, cap)
    bad = int(((r != g).any(axis=1) | (c != gc).any(axis=1)).sum())
    kern_s = sum(best.values()) * 1e-3
    print(json.dumps({"config": name, "pairs": b.npairs, "cells": b.cells, "flag": b.flag, "gpu_kernel_gcups": b.cells / kern_s / 1e9, "gpu_phase_ms": {k: round(v, 3) for k, v in best.items()},
                      "gpu_e2e_gcups": b.cells / e2e_s / 1e9, "reference_cpu_gcups": sb.cells / secs / 1e9, "reference_cores": ncores,
                      "parity_sample_pairs": int(len(idx)), "parity_mismatches": bad}), flush=True)


sw_config("1: 10k x (250bp vs 500bp), flag 0", w.config1(), 10_000)
sw_config("2: 1M x (150-300bp vs 1kb), flag 1", w.config2(1_000_000, seed=1000), 40_000)
# config 3: the realigner end to end
regions = w.config3(200, seed=13)
R.realign_reads(regions[0])
R.realign_regions(regions)                          # first big batch grows the device pool and the pinned staging
t0 = time.perf_counter(); got = R.realign_regions_packed(regions); t_b = time.perf_counter() - t0
st = R.last_stats()
ref_path = os.path.join(ROOT, "oracle", "_ref", "realigner_ref")
line = {"config": "3: realigner, 200 amplicon regions", "reads": sum(len(r.reads) for r in regions), "ssw_pairs": st["pairs"], "ssw_cells": st["cells"], "gpu_batched_s": t_b,
        "gpu_reads_per_s": sum(len(r.reads) for r in regions) / t_b}
if os.path.exists(ref_path):
    t0 = time.perf_counter()
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ref_realigner_runner.py"), ref_path], input=json.dumps([dataclasses.asdict(r) for r in regions]).encode(), capture_output=True, check=True)
    t_r = time.perf_counter() - t0
    want = [(a[0], a[1]) for a in json.loads(p.stdout)]
    line.update({"reference_cpu_s": t_r, "reference_cores": 1, "parity_reads": line["reads"],
                 "parity_mismatches": sum(1 for (gp, gc), (wp, wc) in zip(got, want) for i in range(len(wp)) if gp[i] != wp[i] or gc[i] != wc[i])})
print(json.dumps(line), flush=True)
sw_config("4: 1024 x (10kb vs 12kb), match 4 (int16 clamp), flag 0", w.config4(1024, seed=14, flag=0), 48, cap=4096)
sw_config("4: 1024 x (10kb vs 12kb), match 4 (int16 clamp), flag 1", w.config4(1024, seed=15, flag=1), 48, cap=4096)
sw_config("5: 20k-pair sample of the mixed-length set (100bp-20kb log-uniform, target 1.2x), flag 0", w.config5(20_000, seed=15, flag=0), 400)
