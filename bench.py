#!/usr/bin/env python
"""bench.py -- GCUPS of the batched Smith-Waterman hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--pairs P] [--config 1|2] [--impl ours|reference]
  N > 1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one pass of the hot path (forward score kernels -> second-best epilogue -> reverse score kernels -> banded
traceback + CIGAR) over one batch of BASELINE configs[1]: P pairs (default 1,000,000) of U{150..300} bp reads against 1 kb
haplotypes, flag = 1, maskLen = readLen/2, +4/-6, gapO 8 / gapE 2.  Every rank runs its own P pairs
(weak scaling, no collective: pairs are independent); GCUPS counts forward-matrix cells only.

  value : all ranks' cells / max-over-ranks device time of K steps, inputs already resident in HBM (CUDA events on the launch stream)
  e2e   : same through the public API call Engine.align() with pinned HOST buffers: H2D copies, scheduling, kernels, D2H inside the timed region
  roofline : integer-ALU / DPX issue roofline of the dominant (forward score) kernel -- 4.5 alu-pipe instructions per 2 cells at
             64 lanes/clk/SM (profiles/r01_ubench_cell_pipes.md) -> 28.44 cells/clk/SM -- plus the HBM figure showing it is non-binding
  cpu_baseline : the reference ssw.c (oracle/_ref) on all host cores over a bounded sample of the same workload

--impl reference times the reference's own CPU implementation (oracle/_ref/libssw_ref.so, else the scalar port) instead.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "megapath-nano_b200"

ALU_LANES_PER_CLK_SM = 64.0          # measured: tools/ubench_dpx.cu, profiles/r01_ubench_dpx.json
ALU_INSTR_PER_PACKED_CELL = 4.5      # PRMT + VIMNMX3.RELU + 2 VIADDMNMX.RELU + 0.5 VIMNMX3 (SASS of sw_strip16_kernel)
BYTES_PER_CELL_ALGO = None           # computed per workload: (readLen + refLen) in + 4 B/column record + ~64 B out per pair


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--pairs", type=int, default=1_000_000)
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--cpu-sample", type=int, default=0, help="pairs in the CPU baseline sample (0 = auto, ~15 s)")
    return ap.parse_args()


def make_workload(w, cfg, pairs, seed):
    if cfg == 1:
        return w.config1(pairs, seed=seed)
    return w.config2(pairs, seed=seed)


class ClockSampler(threading.Thread):
    """samples nvidia-smi clocks / throttle reasons of one GPU while the timed region runs"""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.rows = []
        self.stop_flag = threading.Event()
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag.is_set():
                    break
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def finish(self):
        self.stop_flag.set()
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        sm.sort()
        busy = [x for x in sm if x > 0.5 * (mx[0] if mx else 1)] or sm
        return {"sm_mhz": busy[len(busy) // 2] if busy else None, "sm_max_mhz": mx[0] if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_run(b, threads, flag):
    from oracle import oracle
    impl = "ref" if oracle.have_ref() else "port"
    _, _, secs = oracle.run_batch(b.reads, b.read_off, b.refs, b.ref_off, b.masklen, b.mat, b.n, gapO=b.gapO, gapE=b.gapE, flag=flag,
                                  filters=b.filters, filterd=b.filterd, score_size=b.score_size, threads=threads, impl=impl, cigar_cap=64)
    return secs, ("reference" if impl == "ref" else "port")


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    w = importlib.import_module("workloads")
    ncores = os.cpu_count() or 1
    wl_name = {1: "configs[0]: 250bp reads x 500bp targets, flag 0", 2: "configs[1]: U{150..300}bp reads x 1kb haplotypes, flag 1 (begin + banded traceback + CIGAR)"}[args.config]

    # ------------------------------------------------------------------ reference arm: ssw.c on the host cores ------------
    if args.impl == "reference":
        if rank != 0:
            return
        # bounded sample per step: ~ncores * 2500 pairs of the same workload (about 1-2 s of ssw.c per step)
        sample = args.cpu_sample or max(2000, ncores * 2500)
        b = make_workload(w, args.config, sample, seed=100)
        for _ in range(args.warmup):
            cpu_reference_run(b.subset(range(min(sample, ncores * 64))), ncores, b.flag)
        t = 0.0
        kind = "reference"
        for _ in range(args.steps):
            s, kind = cpu_reference_run(b, ncores, b.flag)
            t += s
        gcups = b.cells * args.steps / t / 1e9
        line = {"impl": "reference", "metric": "GCUPS", "value": gcups, "unit": "GCUPS", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "s16/u8 (SSE2)",
                "data": "synthetic", "config": {"workload": wl_name, "pairs_per_step": sample, "scoring": "+4/-6 gapO8 gapE2", "note": "bounded sample of the GPU arm's workload"},
                "cpu_baseline": {"value": gcups, "unit": "GCUPS", "cores": ncores, "kind": kind, "sample": f"{sample} pairs per step x {args.steps} steps, one pair per thread"},
                "e2e": {"value": gcups, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ our arm ----------------------------------------------
    import numpy as np
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    B = importlib.import_module(PKG + ".batch")
    eng = B.Engine(local_rank)
    stream = torch.cuda.Stream(device=local_rank)       # a non-default stream: its handle is what the C ABI launches on
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)

    b = make_workload(w, args.config, args.pairs, seed=1000 + rank)
    # pinned host copies of the inputs (what a caller hands to the public API)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    hb = type("HostBatch", (), {})()
    for k in ("mat", "n", "gapO", "gapE", "flag", "filters", "filterd", "score_size", "name"):
        setattr(hb, k, getattr(b, k))
    hb.npairs = b.npairs
    hb.reads, hb.read_off, hb.refs, hb.ref_off, hb.masklen = pin(b.reads), pin(b.read_off), pin(b.refs), pin(b.ref_off), pin(b.masklen)
    cigar_cap = b.npairs * 24 + len(b.reads) // 4 + 4096
    out = torch.zeros(b.npairs * B.RESULT_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    cig = torch.zeros(cigar_cap, dtype=torch.int32).pin_memory()
    cells = b.cells

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- kernel-only: inputs resident in HBM
    h = eng.upload(hb)
    eng.set_profile(True)
    launches0 = eng.stats()["launches"]
    for _ in range(args.warmup):
        eng.run(h)
    torch.cuda.synchronize()
    launches_per_step = (eng.stats()["launches"] - launches0) // max(args.warmup, 1)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    phase = {"forward": 0.0, "finish": 0.0, "reverse": 0.0, "trace": 0.0}
    e0.record(stream)
    for _ in range(args.steps):
        eng.run(h)
    e1.record(stream)
    barrier()
    dev_ms = e0.elapsed_time(e1)
    # per-phase split of one more (untimed) step
    eng.run(h)
    torch.cuda.synchronize()
    ph = eng.phase_ms() or phase
    rec, _ = eng.fetch(h, b.npairs, cigar_cap, out=out, cig=cig)
    eng.free(h)

    # ---- end to end through the public call, pinned host buffers, copies inside the timed region
    for _ in range(max(1, min(args.warmup, 2))):
        eng.align(hb, cigar_cap=cigar_cap, out=out, cig=cig)
    barrier()
    t0 = time.perf_counter()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record(stream)
    for _ in range(args.steps):
        eng.align(hb, cigar_cap=cigar_cap, out=out, cig=cig)
    e3.record(stream)
    barrier()
    e2e_ms = max(e2.elapsed_time(e3), 1e3 * (time.perf_counter() - t0))
    clocks = sampler.finish() if sampler else None
    recs = np.frombuffer(out.numpy(), dtype=B.RESULT_DTYPE)
    assert int((recs["status"] != 0).sum()) == 0 and int(recs["score1"].min()) > 0, "e2e results look wrong"
    h2d = int(len(b.reads) + len(b.refs) + b.npairs * (4 + 40) + 25)
    d2h = int(b.npairs * (32 + 24) + int(recs["cigar_len"].sum()) * 4)

    # ---- max over ranks
    t = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device="cuda")
    c = torch.tensor([float(cells)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    dev_ms, e2e_ms = float(t[0]), float(t[1])
    total_cells = float(c[0])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = total_cells * args.steps / (dev_ms * 1e-3) / 1e9
    e2e = total_cells * args.steps / (e2e_ms * 1e-3) / 1e9
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    sm_max = float(peaks.get("sm_max_mhz", 1965.0))
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    nsm = torch.cuda.get_device_properties(local_rank).multi_processor_count
    peak_gcups = nsm * (ALU_LANES_PER_CLK_SM * 2.0 / ALU_INSTR_PER_PACKED_CELL) * sm_max * 1e6 / 1e9
    fwd_ms = ph["forward"]
    fwd_gcups = cells / (fwd_ms * 1e-3) / 1e9 if fwd_ms > 0 else None
    algo_bytes = float(len(b.reads) + len(b.refs) + 4 * len(b.refs) + 16 * b.npairs)      # sequences in + one 4 B record per column + ends out
    roof = {"bound": "alu", "kernel": "sw_strip16_kernel (forward score pass)", "achieved": fwd_gcups, "peak": peak_gcups, "unit": "GCUPS",
            "frac": (fwd_gcups / peak_gcups) if fwd_gcups else None,
            "peak_basis": f"{nsm} SMs x 64 alu lanes/clk/SM (measured, profiles/r01_ubench_dpx.json) x 2 cells / 4.5 alu instr x {sm_max:.0f} MHz max SM clock (MEASURED_PEAKS.json)",
            "launch_ms": fwd_ms, "phase_ms": ph,
            "hbm": {"achieved_gbs": algo_bytes / (fwd_ms * 1e-3) / 1e9 if fwd_ms > 0 else None, "peak_gbs": hbm_peak,
                    "frac": (algo_bytes / (fwd_ms * 1e-3) / 1e9 / hbm_peak) if fwd_ms > 0 else None, "algorithmic_bytes_per_launch": algo_bytes},
            "traffic": None}
    # DRAM bytes of the same forward pass from one `ncu --set full` capture (profiles/r01e_strip16_traffic.json, sum over the length bins
    # of one pass = one "launch" of the dominant kernel); only quoted when the capture was taken on this workload size
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r01e_strip16_traffic.json")))
        if args.config == 2 and b.npairs == 1_000_000:
            roof["traffic"] = tr["forward_dram_bytes"]
            roof["traffic_note"] = "dram__bytes_read.sum + dram__bytes_write.sum over the forward launches of one step; algorithmic bytes %.3g" % algo_bytes
    except Exception:
        pass
    if clocks and clocks.get("sm_mhz"):
        roof["frac_at_observed_clock"] = fwd_gcups / (peak_gcups * clocks["sm_mhz"] / sm_max) if fwd_gcups else None

    # ---- CPU baseline: reference ssw.c, one pair per thread on all host cores, bounded sample (N=1 only)
    cpu = None
    if world == 1:
        try:
            sample = args.cpu_sample or max(4000, ncores * 6000)
            sample = min(sample, b.npairs)
            sb = b.subset(np.arange(sample))
            secs, kind = cpu_reference_run(sb, ncores, b.flag)
            cpu = {"value": sb.cells / secs / 1e9, "unit": "GCUPS", "cores": ncores, "kind": kind,
                   "sample": f"first {sample} pairs of the same batch, one pair per thread, {secs:.2f} s"}
        except Exception as ex:      # the checker is optional for the product, never for the number
            cpu = {"value": None, "unit": "GCUPS", "cores": ncores, "kind": "unavailable", "sample": str(ex)[:200]}

    line = {"metric": "GCUPS", "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "s16x2 (int32 traceback)",
            "data": "synthetic",
            "config": {"workload": wl_name, "pairs_per_gpu": b.npairs, "cells_per_gpu": cells, "scoring": "+4/-6, N=-6, gapO 8, gapE 2, score_size 2",
                       "sharding": f"{world} x independent shards, no collective", "l2": "inputs + column records (>5 GB per step) exceed the 126 MB L2; no flush needed"},
            "e2e": {"value": e2e, "unit": "GCUPS", "ms_per_step": e2e_ms / args.steps, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches_per_step * args.steps),
            "roofline": roof, "cpu_baseline": cpu, "clocks": clocks}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
