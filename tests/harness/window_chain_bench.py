"""BASELINE configs[2] one step earlier (SURVEY.md section 8f N4 -> N3 -> hot path): windows -> haplotype assembly on the host threads
(realign/debruijn_graph) -> realigner inputs -> every Smith-Waterman pair of all regions in one GPU batch (realign/realigner).
python tests/harness/window_chain_bench.py [windows] [max_reads] [oracle_sample]  -> two JSON lines (assembler alone, then the whole chain).  The assembler's CPU baseline is the Python
restatement oracle/dbg_oracle.py on a sample of the windows (the reference's own assembler needs Boost and cannot be built here)."""
import importlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
w = importlib.import_module("workloads")
D = importlib.import_module("megapath-nano_b200.debruijn")
R = importlib.import_module("megapath-nano_b200.realigner")
nwin = int(sys.argv[1]) if len(sys.argv) > 1 else 200
max_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 300
sample = int(sys.argv[3]) if len(sys.argv) > 3 else 8
wins = w.config3_windows(nwin, seed=51, max_reads=max_reads)
nreads = sum(len(x.reads) for x in wins)
inputs = [(x.chrom[x.win_start:x.win_end], x.reads, x.low_quality) for x in wins]
D.consensus_windows(inputs[:4])
t0 = time.perf_counter(); cons = D.consensus_windows(inputs); t_asm = time.perf_counter() - t0
t0 = time.perf_counter(); regions, kept = D.regions_from_windows(wins); t_chain = time.perf_counter() - t0
from oracle import dbg_oracle
t0 = time.perf_counter()
same = sum(dbg_oracle.get_consensus(r, ",".join(rd), ",".join(lq))[0] == c for (r, rd, lq), c in zip(inputs[:sample], cons[:sample]))
t_or = (time.perf_counter() - t0) * nwin / max(sample, 1)
out = dict(windows=nwin, reads=nreads, haplotypes=sum(len(c) for c in cons), assemble_s=round(t_asm, 4), windows_per_s=round(nwin / t_asm),
           python_restatement_s_extrapolated=round(t_or, 2), restatement_sample=sample, restatement_equal=same, regions=len(regions), chain_host_s=round(t_chain, 4))
print(json.dumps(out), flush=True)      # the realigner aborts the process without a GPU (no CPU fallback): assembler numbers are out by then
try:
    R.realign_regions_packed(regions[:2])
    t0 = time.perf_counter(); res = R.realign_regions_packed(regions); t_re = time.perf_counter() - t0
    st = R.last_stats()
    out.update(realign_s=round(t_re, 4), sw_pairs=st["pairs"], sw_cells=st["cells"], fast_pass_s=round(st["fast_pass_s"], 4), gpu_s=round(st["gpu_s"], 4),
               compose_s=round(st["compose_s"], 4), reads_per_s_whole_chain=round(sum(len(r.reads) for r in regions) / (t_chain + t_re)))
except Exception as ex:                   # no GPU here: assembler numbers only
    out["realign"] = f"not run ({type(ex).__name__})"
print(json.dumps(out))
