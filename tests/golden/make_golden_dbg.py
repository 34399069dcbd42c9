"""Golden vectors of the window assembler: inputs + the haplotype lists of the REFERENCE ITSELF -- its unmodified
debruijn_graph.cpp compiled over oracle/boost_shim into oracle/_ref/debruijn_graph_ref (oracle/Makefile) -- plus the k of the restatement
oracle/dbg_oracle.py (the reference does not report k).  python tests/golden/make_golden_dbg.py -> tests/golden/dbg_golden.json.gz"""
import gzip, importlib, json, os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
w = importlib.import_module("workloads")
D = importlib.import_module("megapath-nano_b200.debruijn")
from oracle import dbg_oracle, oracle

REF = oracle.dbg_ref_path()
assert REF, "oracle/_ref/debruijn_graph_ref not built (make -C oracle ref)"
wins = w.dbg_windows(14, seed=7001, max_reads=60) + w.dbg_windows(6, seed=7002, max_reads=40, repeat_frac=1.0) + [
    ("ACGTACGTAC", ["ACGTACGTACGT"], [""]),
    ("A" * 120, ["A" * 90], [""]),
    (w.dbg_windows(1, seed=7003, max_reads=30)[0][0], [""], [""]),
] + [(r, rd, lq) for r, rd, lq, _ in w.dbg_cap_windows(seed=41)]
doc = []
for ref, reads, lowq in wins:
    haps = D.get_consensus(ref, reads, lowq, lib_path=REF)           # the reference object, through the reference's own ctypes call sequence
    _, k = dbg_oracle.get_consensus(ref, ",".join(reads), ",".join(lowq))
    doc.append(dict(ref=ref, reads=reads, lowq=lowq, haplotypes=haps, k=k))
with gzip.open(os.path.join(HERE, "dbg_golden.json.gz"), "wt") as f:
    json.dump(doc, f)
print(len(doc), "windows,", sum(len(d["haplotypes"]) for d in doc), "haplotypes,", os.path.getsize(os.path.join(HERE, "dbg_golden.json.gz")), "bytes")
