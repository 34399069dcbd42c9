"""BASELINE configs[2] (realigner amplicon workload) timing: python tests/harness/realign_bench.py [regions] [seed]
Per-region realign_reads latency and whole-set mpn_realign_regions throughput on the GPU, the compiled reference realigner
(oracle/_ref/realigner_ref, clean subprocess, one core) beside it."""
import dataclasses, importlib, json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
w = importlib.import_module("workloads")
R = importlib.import_module("megapath-nano_b200.realigner")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 13
regions = w.config3(n, seed=seed)
reads = sum(len(r.reads) for r in regions)
R.realign_reads(regions[0])                      # context + engine creation
t0 = time.perf_counter()
lat = []
single = []
for rg in regions:
    t = time.perf_counter(); single.append(R.realign_reads(rg)); lat.append(time.perf_counter() - t)
t_single = time.perf_counter() - t0
t0 = time.perf_counter(); many = R.realign_regions(regions); t_cold = time.perf_counter() - t0        # first big batch: device pool + pinned staging grow
t0 = time.perf_counter(); many = R.realign_regions(regions); t_many = time.perf_counter() - t0
st = R.last_stats()
t0 = time.perf_counter(); packed = R.realign_regions_packed(regions); t_packed = time.perf_counter() - t0
assert many == single and packed == many
lat.sort()
out = {"regions": n, "reads": reads, "ssw_pairs": st["pairs"], "ssw_cells": st["cells"],
       "per_region_ms": {"p50": 1e3 * lat[len(lat) // 2], "p95": 1e3 * lat[int(len(lat) * 0.95)], "sum_s": t_single},
       "batched_s": t_many, "batched_first_call_s": t_cold, "batched_packed_s": t_packed, "batched_packed_reads_per_s": reads / t_packed, "batched_split_s": {k: st[k] for k in ("fast_pass_s", "gpu_s", "compose_s")},
       "batched_gcups_ssw_only": st["cells"] / max(st["gpu_s"], 1e-9) / 1e9, "batched_reads_per_s": reads / t_many}
import ctypes
L = R.load(); L.mpn_realign_last_fastpass_kernel_ms.restype = ctypes.c_double
out["fastpass_kernel_ms"] = L.mpn_realign_last_fastpass_kernel_ms()
# the fast pass alone, both placements of the step: GPU kernel (copies and result marshalling included) vs host k-mer index on all cores
t0 = time.perf_counter(); s0, p0, kms = R.fastpass_only(regions, 0); out["fastpass_only_gpu_s"] = time.perf_counter() - t0
t0 = time.perf_counter(); s1, p1, _ = R.fastpass_only(regions, 1); out["fastpass_only_host_s"] = time.perf_counter() - t0
out["fastpass_identical"] = (s0 == s1 and p0 == p1)
out["fastpass_hap_read_pairs"] = len(p0) // 2
ref = os.path.join(ROOT, "oracle", "_ref", "realigner_ref")
if os.path.exists(ref) and "noref" not in sys.argv:
    t0 = time.perf_counter()
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ref_realigner_runner.py"), ref], input=json.dumps([dataclasses.asdict(r) for r in regions]).encode(), capture_output=True, check=True)
    t_ref = time.perf_counter() - t0
    want = [(a[0], a[1]) for a in json.loads(p.stdout)]
    out["reference_cpu_s"] = t_ref
    out["reference_reads_per_s"] = reads / t_ref
    out["identical_to_reference"] = want == many
print(json.dumps(out))
