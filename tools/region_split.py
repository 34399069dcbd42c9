import importlib, sys, time
sys.path.insert(0, "/root/repo")
w = importlib.import_module("workloads")
R = importlib.import_module("megapath-nano_b200.realigner")
regions = w.config3(60, seed=13)
R.realign_reads(regions[0])
tot = dict(wall=0, fast=0, gpu=0, comp=0, pairs=0)
for rg in regions:
    t = time.perf_counter(); R.realign_reads(rg); tot["wall"] += time.perf_counter() - t
    st = R.last_stats(); tot["fast"] += st["fast_pass_s"]; tot["gpu"] += st["gpu_s"]; tot["comp"] += st["compose_s"]; tot["pairs"] += st["pairs"]
n = len(regions)
print({k: (round(1e3 * v / n, 3) if k != "pairs" else v / n) for k, v in tot.items()})
