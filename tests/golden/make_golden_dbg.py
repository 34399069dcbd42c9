"""Golden vectors of the window assembler: inputs + the haplotype lists / k of the string-keyed restatement
oracle/dbg_oracle.py.  PARITY UNPINNED against the reference itself (its debruijn_graph.cpp needs Boost.Graph, not available here): these
fixtures freeze the restated behaviour so that neither implementation drifts.  python tests/golden/make_golden_dbg.py -> tests/golden/dbg_golden.json.gz"""
import gzip, importlib, json, os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
w = importlib.import_module("workloads")
from oracle import dbg_oracle

wins = w.dbg_windows(14, seed=7001, max_reads=60) + w.dbg_windows(6, seed=7002, max_reads=40, repeat_frac=1.0) + [
    ("ACGTACGTAC", ["ACGTACGTACGT"], [""]),
    ("A" * 120, ["A" * 90], [""]),
    (w.dbg_windows(1, seed=7003, max_reads=30)[0][0], [""], [""]),
]
doc = []
for ref, reads, lowq in wins:
    haps, k = dbg_oracle.get_consensus(ref, ",".join(reads), ",".join(lowq))
    doc.append(dict(ref=ref, reads=reads, lowq=lowq, haplotypes=haps, k=k))
with gzip.open(os.path.join(HERE, "dbg_golden.json.gz"), "wt") as f:
    json.dump(doc, f)
print(len(doc), "windows,", sum(len(d["haplotypes"]) for d in doc), "haplotypes,", os.path.getsize(os.path.join(HERE, "dbg_golden.json.gz")), "bytes")
