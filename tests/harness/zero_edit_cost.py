"""What the ZERO-EDIT path costs in the reference's own process model (SURVEY.md section 8b "Threading"; bin/realignment/realignment.sh:50-60
starts one process per candidate position, pyssw.py:137-147 issues one ssw_align per read): a fresh Python process imports the reference's
unmodified pyssw.py (baseline/_ref), loads libssw.so -- the product or the compiled reference ssw.c -- and aligns N ~400 bp queries against
a 601 bp reference, one call per query.  Prints, per library, the wall time of the whole process and of the align loop alone.
    python tests/harness/zero_edit_cost.py [queries per process] [processes]"""
import json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle
from test_gpu_reference_callers import queries_601, PRODUCT_LIBSSW

CHILD = r"""
import importlib.util, json, sys, time
t0 = time.perf_counter()
spec = importlib.util.spec_from_file_location("reference_pyssw", sys.argv[1]); m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
job = json.load(sys.stdin)
a = m.SSW(lib_path=sys.argv[2]); a.set_reference_sequence(job["ref"])
t1 = time.perf_counter()
first = a.align(job["qs"][0])
t2 = time.perf_counter()
res = [a.align(q) for q in job["qs"][1:]]
t3 = time.perf_counter()
print(json.dumps({"import_s": t1 - t0, "first_call_s": t2 - t1, "loop_s": t3 - t2, "calls": len(job["qs"]) - 1, "check": sum(r[0] for r in res)}))
"""
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
procs = int(sys.argv[2]) if len(sys.argv) > 2 else 3
d = oracle.build_ref_callers() or oracle.callers_dir()
ref, qs = queries_601(7, n)
job = json.dumps({"ref": ref, "qs": qs}).encode()
out = {"queries_per_process": n + 1}
for name, lib in (("reference_ssw_c", oracle.ref_path()), ("product_gpu", PRODUCT_LIBSSW)):
    runs = []
    for _ in range(procs):
        t0 = time.perf_counter()
        p = subprocess.run([sys.executable, "-c", CHILD, os.path.join(d, "pyssw.py"), lib], input=job, capture_output=True, check=True)
        r = json.loads(p.stdout); r["process_s"] = time.perf_counter() - t0
        runs.append(r)
    best = min(runs, key=lambda r: r["process_s"])
    out[name] = {**best, "us_per_call": 1e6 * best["loop_s"] / best["calls"]}
assert out["reference_ssw_c"]["check"] == out["product_gpu"]["check"]
out["note"] = "process_s includes interpreter start; first_call_s on the product includes CUDA context + engine creation"
print(json.dumps(out))
