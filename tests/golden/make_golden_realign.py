#!/usr/bin/env python
"""Regenerates tests/golden/realign_golden.json.gz by running the UNMODIFIED reference realigner (oracle/_ref/realigner_ref,
compiled by oracle/Makefile from ssw_cpp.cpp + ssw.c + realigner.cpp) on seeded regions of BASELINE configs[2]
(workloads.config3) plus hand-made edge regions.  Run in the build container:

    python tests/golden/make_golden_realign.py

Each entry = the arguments of realign_reads (realigner.cpp:856) and what it returned: new position + CIGAR per read."""
import dataclasses
import gzip
import importlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402

w = importlib.import_module("workloads")


def edge_regions():
    """reads <= 32 bases (never indexed), reads with N, lower-case bases, a read longer than every haplotype, a read that is
    pure noise, duplicate haplotypes, a single haplotype equal to the reference, deletion- and insertion-only haplotypes"""
    import numpy as np
    rng = np.random.default_rng(5)
    dna = lambda n: "".join("ACGT"[i] for i in rng.integers(0, 4, size=n))
    pre, cen, suf = dna(40), dna(300), dna(40)
    ref = pre + cen + suf
    hap_del = pre + cen[:140] + cen[152:] + suf
    hap_ins = pre + cen[:200] + "ACGTTGCA" + cen[200:] + suf
    hap_snv = pre + cen[:100] + ("A" if cen[100] != "A" else "C") + cen[101:] + suf
    out = []
    reads = [hap_del[100:250], hap_del[120:300], hap_ins[150:330], hap_snv[60:260], ref[10:42], ref[5:30], dna(120),
             hap_del[90:240].lower(), hap_ins[160:300][:70] + "N" + hap_ins[160:300][71:], ref + "ACGT", hap_del[150:200] + dna(60)]
    pos = [1000 + k for k in range(len(reads))]
    cig = [f"{len(r)}M" for r in reads]
    out.append(w.RegionWorkload(ref, [ref, hap_del, hap_ins, hap_snv], reads, pos, cig, 1000, 40, 40))
    out.append(w.RegionWorkload(ref, [hap_del, hap_del, ref], reads[:4], pos[:4], cig[:4], 77, 40, 40))
    out.append(w.RegionWorkload(ref, [ref], reads[:6], pos[:6], cig[:6], 5, 40, 40))
    out.append(w.RegionWorkload(ref, [hap_ins], [hap_ins[k:k + 150] for k in range(0, 230, 10)], list(range(23)), ["150M"] * 23, 0, 0, 0))
    out.append(w.RegionWorkload(ref, [hap_snv, hap_ins, hap_del, ref], [], [], [], 9, 40, 40))
    return out


def run_reference(regions):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ref_realigner_runner.py"), oracle.realigner_ref_path()],
                       input=json.dumps([dataclasses.asdict(r) for r in regions]).encode(), capture_output=True, check=True)
    return json.loads(p.stdout)


def main():
    oracle.build()
    regions = edge_regions() + w.config3(10, seed=301, max_reads=120, max_haps=6) + w.config3(4, seed=302, max_reads=80, max_haps=12, n_frac=0.004)
    res = run_reference(regions)
    doc = [{"region": dataclasses.asdict(r), "positions": a[0], "cigars": a[1]} for r, a in zip(regions, res)]
    path = os.path.join(ROOT, "tests", "golden", "realign_golden.json.gz")
    with gzip.open(path, "wt", compresslevel=9) as f:
        json.dump(doc, f)
    nreads = sum(len(d["positions"]) for d in doc)
    moved = sum(1 for d in doc for c, c0 in zip(d["cigars"], d["region"]["cigars"]) if c != c0)
    print(f"wrote {path}: {len(doc)} regions, {nreads} reads, {moved} realigned, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
