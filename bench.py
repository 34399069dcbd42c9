#!/usr/bin/env python
"""bench.py -- GCUPS of the batched Smith-Waterman hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config 1|2|3|4|5] [--pairs P] [--flag F] [--impl ours|reference]
  N > 1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Default workload (the one BASELINE.json's metric is quoted on, configs[1]): one "step" = one pass of the hot path (forward score kernels
-> second-best epilogue -> reverse score kernels -> banded traceback + CIGAR) over P pairs (default 1,000,000) of U{150..300} bp reads
against 1 kb haplotypes, flag = 1, maskLen = readLen/2, +4/-6, gapO 8 / gapE 2.  Every rank runs its own P pairs
(weak scaling, no collective: pairs are independent).  GCUPS counts forward-matrix cells only.

  value    : all ranks' cells / max-over-ranks device time of K steps, inputs already resident in HBM (CUDA events on the launch stream)
  e2e      : the same through the public call with pinned HOST buffers: H2D copies, scheduling, kernels, D2H inside the timed region.  Three input formats are
             timed: 2 bits per base (mpn_align_batch_packed2, the headline), 4 bits (mpn_align_batch_packed4), one int8 code per base (mpn_align_batch)
  roofline : integer-ALU / DPX issue roofline of the dominant (forward score) kernel -- 4.5 alu-pipe instructions per 2 cells at
             64 lanes/clk/SM (profiles/r01_ubench_cell_pipes.md) -> 28.44 cells/clk/SM -- plus the HBM figure showing it is non-binding
  parity   : a stratified sample (>= 100 k pairs at full size) of the TIMED batch's records and CIGARs diffed against the compiled, unmodified
             reference ssw.c (oracle/_ref) after the timed region
  cpu_baseline : that reference on all host cores over a bounded sample of the same workload (N = 1 only)
  configs  : (N = 1, default run) bounded-size runs of the other BASELINE configs -- 1, 3, 4 (flag 0 and 1), 5 -- each with kernel / e2e GCUPS
             and its parity count, so every figure in DESIGN.md is reproducible by this command

--config 1|3|4|5 makes that config the main line.  --config 5 is BASELINE configs[4] at size: 10 M mixed-length pairs as a stream of
length-sorted chunks dealt to the ranks by cost (strong scaling: the total is fixed, every rank derives the same plan and generates only
its own chunks; results are reduced to a checksum and per-rank parity counts).
--impl reference times the reference's own CPU implementation (oracle/_ref/libssw_ref.so, unmodified ssw.c, one pair per thread on all
host cores) on a bounded sample of the same config.
"""
import argparse
import importlib
import json
import os
import queue
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "megapath-nano_b200"

ALU_LANES_PER_CLK_SM = 64.0          # measured: tools/ubench_dpx.cu, profiles/r01_ubench_dpx.json
ALU_INSTR_PER_PACKED_CELL = 4.5      # PRMT + VIMNMX3.RELU + 2 VIADDMNMX.RELU + 0.5 VIMNMX3: the DP recurrence itself (SASS of sw_strip16_kernel)
WL = {1: "configs[0]: 10k pairs, 250bp reads x 500bp targets, flag 0 (score + end positions)",
      2: "configs[1]: U{150..300}bp reads x 1kb haplotypes, flag 1 (begin + banded traceback + CIGAR)",
      3: "configs[2]: realigner amplicon workload, per-region haplotypes x overlapping reads (mpn_realign_regions_packed)",
      4: "configs[3]: ONT-scale, 10kb reads (8% error) x 12kb windows (int16 clamp path)",
      5: "configs[4]: mixed-length pairs 100bp-20kb (log-uniform), target 1.2x read, length-sorted chunks dealt by cost"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--pairs", type=int, default=0, help="pairs per GPU (configs 1, 2, 4), regions (config 3), total pairs (config 5); 0 = the BASELINE size")
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--flag", type=int, default=None, help="ssw_align flag override (config 4 / 5: 0 or 1)")
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--cpu-sample", type=int, default=0, help="pairs in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-configs", action="store_true", help="skip the bounded runs of the other configs on the default line")
    return ap.parse_args()


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


# ------------------------------------------------------------------------------------------------ the reference on the host cores
def cpu_reference_run(b, threads, cap=64):
    """(table, cigars, seconds, kind): the compiled unmodified ssw.c when oracle/_ref is present (always on the GPU box: it travels), else the
    scalar restatement -- the one place besides tests/ and smoke() that may execute oracle/."""
    from oracle import oracle
    impl = "ref" if oracle.have_ref() else "port"
    r, c, secs = oracle.run_batch(b.reads, b.read_off, b.refs, b.ref_off, b.masklen, b.mat, b.n, gapO=b.gapO, gapE=b.gapE, flag=b.flag,
                                  filters=b.filters, filterd=b.filterd, score_size=b.score_size, threads=threads, impl=impl, cigar_cap=cap)
    return r, c, secs, ("reference" if impl == "ref" else "port")


def parity_of_sample(B, b, rec, cig, idx, threads, cap=64):
    """diff the GPU records of pairs idx against the reference run on the same pairs -> (dict for the JSON line, reference seconds, sample cells)"""
    sb = b.subset(idx)
    r, c, secs, kind = cpu_reference_run(sb, threads, cap)
    g, gc = B.as_table(rec[idx], cig, cap)
    bad = int(((r != g).any(axis=1) | (c != gc).any(axis=1)).sum())
    return {"pairs": int(len(idx)), "mismatches": bad, "against": "oracle/_ref/libssw_ref.so (unmodified ssw.c)" if kind == "reference" else "oracle/ssw_oracle.c (restatement)",
            "fields": "score1 score2 ref_begin1 ref_end1 read_begin1 read_end1 ref_end2 cigarLen + CIGAR words"}, secs, sb.cells


def make_pair_workload(w, cfg, pairs, seed, flag=None, threads=8):
    import numpy as np
    if cfg == 1:
        b = w.config1(pairs or 10_000, seed=seed)
    elif cfg == 2:
        b = w.config2(pairs or 1_000_000, seed=seed)
    elif cfg == 4:
        n = pairs or 1024
        b = w.gen_pairs_fast(np.full(n, 10_000), np.full(n, 12_000), seed, err=0.08, threads=threads, flag=1 if flag is None else flag,
                             name=f"config4: 10kb x 12kb flag{1 if flag is None else flag}")
    else:
        raise SystemExit(f"config {cfg} is not a pair workload")
    if flag is not None:
        b.flag = flag
    return b


# ------------------------------------------------------------------------------------------------ our arm, pair workloads (configs 1, 2, 4)
def run_pairs(args, env, b, steps, warmup, sample_pairs, cap=64, want_clocks=True):
    """kernel-only + e2e + parity for one PairBatch on this rank's engine; returns a dict of local measurements"""
    import numpy as np
    import torch
    B, eng, stream = env["B"], env["eng"], env["stream"]
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    hb = type("HostBatch", (), {})()
    for k in ("mat", "n", "gapO", "gapE", "flag", "filters", "filterd", "score_size", "name"):
        setattr(hb, k, getattr(b, k))
    hb.npairs = b.npairs
    hb.reads, hb.read_off, hb.refs, hb.ref_off, hb.masklen = pin(b.reads), pin(b.read_off), pin(b.refs), pin(b.ref_off), pin(b.masklen)
    cigar_cap = b.npairs * 24 + len(b.reads) // 4 + 4096 if b.read_len.max() <= 2000 else int(b.read_len.sum() + b.ref_len.sum()) // 2 + 64 * b.npairs
    out = torch.zeros(b.npairs * B.RESULT_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    cig = torch.zeros(max(cigar_cap, 1), dtype=torch.int32).pin_memory()

    # ---- kernel-only: inputs resident in HBM
    h = eng.upload(hb)
    launches0 = eng.stats()["launches"]
    for _ in range(warmup):
        eng.run(h)
    torch.cuda.synchronize()
    launches_per_step = (eng.stats()["launches"] - launches0) // max(warmup, 1)
    eng.set_profile(True)                        # phase events of the timed steps (averaged below): they are the roofline's launch durations
    sampler = ClockSampler(env["local_rank"]) if (env["rank"] == 0 and want_clocks) else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    env["barrier"]()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        eng.run(h)
    e1.record(stream)
    env["barrier"]()
    dev_ms = e0.elapsed_time(e1)
    ph, ph_runs = eng.phase_ms_mean()            # per-phase split: mean over the timed steps (CUDA events on the launch stream inside the engine)
    ph = ph or {"forward": 0.0, "finish": 0.0, "reverse": 0.0, "trace": 0.0}
    eng.set_profile(False)
    eng.fetch(h, b.npairs, cigar_cap, out=out, cig=cig)
    eng.free(h)

    # ---- end to end through the public call, pinned host buffers, copies inside the timed region
    for _ in range(max(1, min(warmup, 2))):
        eng.align(hb, cigar_cap=cigar_cap, out=out, cig=cig)
    env["barrier"]()
    t0 = time.perf_counter()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record(stream)
    for _ in range(steps):
        eng.align(hb, cigar_cap=cigar_cap, out=out, cig=cig)
    e3.record(stream)
    env["barrier"]()
    e2e_ms = max(e2.elapsed_time(e3), 1e3 * (time.perf_counter() - t0))
    # ---- the same through the nibble-packed entry point (mpn_align_batch_packed4): half the host->device bytes
    e2e4_ms = None
    if b.n <= 16:
        r4, f4 = pin(B.pack4(b.reads)), pin(B.pack4(b.refs))
        for _ in range(max(1, min(warmup, 2))):
            eng.align_packed4(hb, r4, f4, cigar_cap=cigar_cap, out=out, cig=cig)
        env["barrier"]()
        t0 = time.perf_counter()
        e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e4.record(stream)
        for _ in range(steps):
            eng.align_packed4(hb, r4, f4, cigar_cap=cigar_cap, out=out, cig=cig)
        e5.record(stream)
        env["barrier"]()
        e2e4_ms = max(e4.elapsed_time(e5), 1e3 * (time.perf_counter() - t0))
    # ---- and through the 2-bit entry point (mpn_align_batch_packed2): a quarter of the bytes, N codes as an exception list.  The timed batch's
    #      results (checked against the reference below) are the ones of this last path.
    e2e2_ms, h2d2 = None, None
    if b.n <= 5:
        (r2a, rx), (f2a, fx) = B.pack2(b.reads), B.pack2(b.refs)
        r2, f2 = pin(r2a), pin(f2a)
        rxp, fxp = (pin(rx) if len(rx) else rx), (pin(fx) if len(fx) else fx)
        for _ in range(max(1, min(warmup, 2))):
            eng.align_packed2(hb, r2, rxp, f2, fxp, cigar_cap=cigar_cap, out=out, cig=cig)
        env["barrier"]()
        t0 = time.perf_counter()
        e6, e7 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e6.record(stream)
        for _ in range(steps):
            eng.align_packed2(hb, r2, rxp, f2, fxp, cigar_cap=cigar_cap, out=out, cig=cig)
        e7.record(stream)
        env["barrier"]()
        e2e2_ms = max(e6.elapsed_time(e7), 1e3 * (time.perf_counter() - t0))
        h2d2 = int(len(r2a) + len(f2a) + 8 * (len(rx) + len(fx)) + b.npairs * (4 + 48) + 25)
    clocks = sampler.finish() if sampler else None
    recs = np.frombuffer(out.numpy(), dtype=B.RESULT_DTYPE)
    cigs = np.frombuffer(cig.numpy(), dtype=np.uint32)
    assert int((recs["status"] != 0).sum()) == 0, "e2e results carry NULL records"

    # ---- parity of the timed batch itself: stratified sample (every k-th pair) against the reference
    n_s = min(sample_pairs, b.npairs)
    idx = np.unique(np.linspace(0, b.npairs - 1, n_s).astype(np.int64))
    par, ref_secs, ref_cells = parity_of_sample(B, b, recs, cigs, idx, env["threads"], cap)
    h2d = int(len(b.reads) + len(b.refs) + b.npairs * (4 + 48) + 25)
    d2h = int(b.npairs * (32 + (24 if b.flag else 0)) + int(recs["cigar_len"].sum()) * 4)
    return {"cells": b.cells, "dev_ms": dev_ms, "e2e_ms": e2e_ms, "e2e4_ms": e2e4_ms or e2e_ms, "h2d4": h2d - (len(b.reads) + len(b.refs)) // 2, "e2e2_ms": e2e2_ms or e2e4_ms or e2e_ms, "h2d2": h2d2 or h2d, "phase_ms": ph, "launches_per_step": int(launches_per_step), "clocks": clocks, "parity": par,
            "cpu": {"secs": ref_secs, "cells": ref_cells, "pairs": int(len(idx))}, "h2d": h2d, "d2h": d2h,
            "algo_bytes": float(len(b.reads) + len(b.refs) + 4 * len(b.refs) + 16 * b.npairs)}


# ------------------------------------------------------------------------------------------------ our arm, config 3 (realigner)
def run_config3(args, env, nregions, steps, warmup, ref_regions=6):
    import dataclasses
    w = env["w"]
    R = importlib.import_module(PKG + ".realigner")
    regions = w.config3(nregions, seed=13)
    reads = sum(len(r.reads) for r in regions)
    R.realign_reads(regions[0])
    for _ in range(max(1, min(warmup, 2))):
        R.realign_regions_packed(regions)
    t0 = time.perf_counter()
    for _ in range(steps):
        got = R.realign_regions_packed(regions)
    secs = (time.perf_counter() - t0) / steps
    st = R.last_stats()
    out = {"regions": nregions, "reads": reads, "ssw_pairs": int(st["pairs"]), "ssw_cells": int(st["cells"]), "s_per_pass": secs, "reads_per_s": reads / secs,
           "e2e_gcups": st["cells"] / secs / 1e9, "gpu_gcups_ssw_only": st["cells"] / max(st["gpu_s"], 1e-9) / 1e9,
           "split_s": {k: st[k] for k in ("fast_pass_s", "gpu_s", "compose_s")}}
    # parity + CPU baseline: the compiled reference realigner (one core, clean subprocess) on the first regions
    ref = os.path.join(ROOT, "oracle", "_ref", "realigner_ref")
    if os.path.exists(ref):
        sub = regions[:ref_regions]
        t0 = time.perf_counter()
        p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ref_realigner_runner.py"), ref], input=json.dumps([dataclasses.asdict(r) for r in sub]).encode(), capture_output=True, check=True)
        t_ref = time.perf_counter() - t0
        want = [(a[0], a[1]) for a in json.loads(p.stdout)]
        bad = sum(1 for (gp, gc), (wp, wc) in zip(got, want) for i in range(len(wp)) if gp[i] != wp[i] or gc[i] != wc[i])
        nr = sum(len(r.reads) for r in sub)
        out["parity"] = {"reads": nr, "mismatches": bad, "against": "oracle/_ref/realigner_ref (unmodified realigner.cpp + ssw_cpp.cpp + ssw.c)"}
        out["cpu_baseline"] = {"reads_per_s": nr / t_ref, "cores": 1, "kind": "reference", "sample": f"first {len(sub)} regions ({nr} reads), {t_ref:.2f} s"}
    return out


# ------------------------------------------------------------------------------------------------ our arm, config 5 (stream of length-sorted chunks)
def run_config5(args, env, npairs, steps, warmup, flag, sample_every=1000):
    """one step = one pass over the whole stream.  Per chunk: generate on host threads (pure function of seed and pair index, so every rank
    generates only its own chunks), upload, timed run (CUDA events), fetch.  value: cells / sum of the timed runs; e2e: cells / wall time of
    the pass (generation runs ahead on a producer thread; copies, scheduling and record conversion are inside)."""
    import numpy as np
    import torch
    w, B, eng, stream = env["w"], env["B"], env["eng"], env["stream"]
    rank, world = env["rank"], env["world"]
    st = w.MixedStream(npairs, seed=15, flag=flag, world=world)
    mine = st.plan(world)[rank]
    gen_threads = max(2, env["ncores"] // world - 1)

    # chunks are generated straight into pinned host buffers (three sets: one being consumed, two queued ahead)
    nbuf = 3
    rb_max = max(st.chunk_bytes(k)[0] for k in mine) if mine else 1
    fb_max = max(st.chunk_bytes(k)[1] for k in mine) if mine else 1
    pin_r = [torch.empty(rb_max, dtype=torch.int8).pin_memory() for _ in range(nbuf)]
    pin_f = [torch.empty(fb_max, dtype=torch.int8).pin_memory() for _ in range(nbuf)]
    free_sets = queue.Queue()
    for i in range(nbuf):
        free_sets.put(i)

    def producer(ids, q):
        for k in ids:
            i = free_sets.get()
            q.put((k, st.chunk(k, threads=gen_threads, reads_out=pin_r[i].numpy(), refs_out=pin_f[i].numpy()), i))
        q.put(None)

    # warm-up: small chunks of the same stream (a full pass is minutes on one GPU; stated in the line)
    small = sorted(mine, key=lambda k: st.chunk_cost[k])[:max(warmup, 0)]
    for k in small:
        b = st.chunk(k, threads=gen_threads)
        eng.align(b)
    torch.cuda.synchronize()
    env["barrier"]()
    dev_ms = 0.0
    checksum = np.zeros(4, dtype=np.int64)
    sample = []                                       # every sample_every-th pair of the stream: (global index, its bases, its record, its CIGAR words)
    launches0 = eng.stats()["launches"]
    h2d = d2h = 0
    t_wall = time.perf_counter()
    for _ in range(steps):
        q = queue.Queue(maxsize=nbuf)
        th = threading.Thread(target=producer, args=(mine, q), daemon=True)
        th.start()
        while True:
            item = q.get()
            if item is None:
                break
            k, b, buf = item
            cap = int(b.read_len.sum() + b.ref_len.sum()) // 2 + 64 * b.npairs if flag else 1
            h = eng.upload(b)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); eng.run(h); e1.record(stream)
            rec, cig = eng.fetch(h, b.npairs, cap)
            eng.free(h)
            dev_ms += e0.elapsed_time(e1)
            h2d += len(b.reads) + len(b.refs) + 52 * b.npairs
            d2h += (32 + (24 if flag else 0)) * b.npairs + 4 * int(rec["cigar_len"].sum())
            assert int((rec["status"] != 0).sum()) == 0
            checksum += np.array([rec["score1"].astype(np.int64).sum(), rec["ref_end1"].astype(np.int64).sum(), rec["read_end1"].astype(np.int64).sum(), b.npairs])
            a = int(st.bounds[k])
            for i in range((-a) % sample_every, b.npairs, sample_every):
                if len(sample) < 4000:
                    o, n = int(rec["cigar_off"][i]), int(rec["cigar_len"][i])
                    sample.append((a + i, b.subset([i]), rec[i:i + 1].copy(), cig[o:o + n].copy()))
            free_sets.put(buf)
        th.join()
    wall_ms = 1e3 * (time.perf_counter() - t_wall)
    launches = eng.stats()["launches"] - launches0
    # parity: the sampled pairs against the reference (the slowest part on the host: long pairs at ~1 GCUPS per core), on this rank's share of the cores
    bad, ref_secs, ref_cells = 0, 0.0, 0
    lock = threading.Lock()

    def check(items):
        nonlocal bad, ref_secs, ref_cells
        for gi, sb, rec1, cw in items:
            capw = max(64, len(cw) + 8)
            r, c, secs, kind = cpu_reference_run(sb, 1, capw)
            one = rec1.copy(); one["cigar_off"] = 0
            g, gc = B.as_table(one, cw if len(cw) else np.zeros(1, np.uint32), capw)
            with lock:
                ref_secs += secs; ref_cells += sb.cells
                bad += int((r != g).any() or (c != gc).any())
    nthr = max(1, env["threads"])
    ths = [threading.Thread(target=check, args=(sample[t::nthr],)) for t in range(nthr)]
    [t.start() for t in ths]; [t.join() for t in ths]
    cells = int(sum(st.chunk_cells[k] for k in mine))
    return {"cells": cells * steps, "total_cells": st.total_cells, "dev_ms": dev_ms, "wall_ms": wall_ms, "chunks": len(mine), "nchunks": st.nchunks, "launches": int(launches),
            "checksum": checksum, "parity_pairs": len(sample), "parity_bad": bad, "ref_secs": ref_secs, "ref_cells": ref_cells, "h2d": h2d // max(steps, 1), "d2h": d2h // max(steps, 1),
            "warmup_chunks": len(small), "ref_threads": nthr}


# ------------------------------------------------------------------------------------------------ reference arm
def reference_arm(args, w, ncores):
    import numpy as np
    steps = args.steps or 5
    cfg = args.config
    line = {"impl": "reference", "metric": "GCUPS", "unit": "GCUPS", "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "higher_is_better": True,
            "scaling": "strong" if cfg == 5 else "weak", "vs_baseline": None, "dtype": "s16/u8 (SSE2)", "data": "synthetic"}
    if cfg == 3:
        import dataclasses
        regions = w.config3(args.pairs or 6, seed=13)
        ref = os.path.join(ROOT, "oracle", "_ref", "realigner_ref")
        if not os.path.exists(ref):
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/realigner_ref not built"}))
            return
        t0 = time.perf_counter()
        for _ in range(steps):
            subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ref_realigner_runner.py"), ref], input=json.dumps([dataclasses.asdict(r) for r in regions]).encode(), capture_output=True, check=True)
        t = time.perf_counter() - t0
        reads = sum(len(r.reads) for r in regions)
        line.update({"metric": "reads/s (realigner end to end)", "unit": "reads/s", "value": reads * steps / t, "ms_per_step": 1e3 * t / steps,
                     "config": {"workload": WL[3], "regions_per_step": len(regions), "note": "bounded sample: the reference realigner takes ~0.3 s per region on one core (its process model is one process per position)"},
                     "cpu_baseline": {"value": reads * steps / t, "unit": "reads/s", "cores": 1, "kind": "reference", "sample": f"{len(regions)} regions per step"},
                     "e2e": {"value": reads * steps / t, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        print(json.dumps(line))
        return
    if cfg == 5:
        st = w.MixedStream(args.pairs or 10_000_000, seed=15, flag=args.flag or 0)
        every = max(1, st.npairs // (args.cpu_sample or max(400, ncores * 40)))
        idx = np.arange(every // 2, st.npairs, every)
        b = w.gen_pairs_fast(st.rl[idx], st.fl[idx], 15, err=0.05, threads=ncores, flag=args.flag or 0)   # same length distribution, stratified 1/every
        sample_note = f"stratified 1/{every} subsample of the length-sorted stream ({len(idx)} pairs per step), extrapolated: GCUPS is intensive"
    else:
        full = {1: 10_000, 2: 1_000_000, 4: 1024}[cfg]
        sample = args.cpu_sample or {1: 10_000, 2: max(2000, ncores * 2500), 4: max(16, ncores * 2)}[cfg]
        b = make_pair_workload(w, cfg, min(sample, args.pairs or full), seed=100, flag=args.flag, threads=ncores)
        sample_note = f"{b.npairs} pairs per step of the GPU arm's workload ({args.pairs or full} pairs per GPU): GCUPS is intensive, the figure extrapolates"
    for _ in range(min(args.warmup, 1)):
        cpu_reference_run(b.subset(range(min(b.npairs, ncores * 8))), ncores)
    t, kind = 0.0, "reference"
    for _ in range(steps):
        _, _, s, kind = cpu_reference_run(b, ncores)
        t += s
    gcups = b.cells * steps / t / 1e9
    line.update({"value": gcups, "ms_per_step": 1e3 * t / steps,
                 "config": {"workload": WL[cfg], "pairs_per_step": b.npairs, "flag": int(b.flag), "scoring": "+4/-6 gapO8 gapE2", "note": sample_note},
                 "cpu_baseline": {"value": gcups, "unit": "GCUPS", "cores": ncores, "kind": kind, "sample": sample_note + ", one pair per thread"},
                 "e2e": {"value": gcups, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(line))


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    w = importlib.import_module("workloads")
    ncores = os.cpu_count() or 1
    if args.impl == "reference":
        if rank == 0:
            reference_arm(args, w, ncores)
        return

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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    env = {"B": B, "eng": eng, "stream": stream, "rank": rank, "local_rank": local_rank, "world": world, "barrier": barrier, "w": w, "ncores": ncores,
           "threads": max(1, ncores // world)}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    sm_max = float(peaks.get("sm_max_mhz", 1965.0))
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)"
    nsm = torch.cuda.get_device_properties(local_rank).multi_processor_count
    peak_gcups = nsm * (ALU_LANES_PER_CLK_SM * 2.0 / ALU_INSTR_PER_PACKED_CELL) * sm_max * 1e6 / 1e9

    def reduce_max_sum(maxes, sums):
        t = torch.tensor(list(maxes), dtype=torch.float64, device="cuda")
        c = torch.tensor(list(sums), dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(c, op=dist.ReduceOp.SUM)
        return [float(x) for x in t], [float(x) for x in c]

    cfg = args.config
    line = {"metric": "GCUPS", "unit": "GCUPS", "n_gpus": world, "warmup": args.warmup, "higher_is_better": True, "vs_baseline": None,
            "dtype": "s16x2 (int32 traceback)", "data": "synthetic"}

    if cfg in (1, 2, 4):
        steps = args.steps or (5 if cfg != 1 else 200)
        b = make_pair_workload(w, cfg, args.pairs, seed=1000 + rank, flag=args.flag, threads=env["threads"])
        sample = args.cpu_sample or {1: 10_000, 2: 100_000, 4: 24}[cfg]
        m = run_pairs(args, env, b, steps, args.warmup, sample, cap=64 if cfg != 4 else 8192)
        (dev_ms, e2e_ms, e2e4_ms, e2e2_ms), (cells, bad, npar) = reduce_max_sum([m["dev_ms"], m["e2e_ms"], m["e2e4_ms"], m["e2e2_ms"]], [m["cells"], m["parity"]["mismatches"], m["parity"]["pairs"]])
        if rank == 0:
            value = cells * steps / (dev_ms * 1e-3) / 1e9
            e2e = cells * steps / (e2e_ms * 1e-3) / 1e9
            e2e4 = cells * steps / (e2e4_ms * 1e-3) / 1e9
            e2e2 = cells * steps / (e2e2_ms * 1e-3) / 1e9
            ph = m["phase_ms"]
            fwd_ms = ph["forward"]
            fwd_gcups = m["cells"] / (fwd_ms * 1e-3) / 1e9 if fwd_ms > 0 else None
            kernel = "sw_strip16_kernel (forward score pass)" if cfg != 4 else "sw_long16_kernel (forward score pass, int16 clamp)"
            roof = {"bound": "alu", "kernel": kernel, "achieved": fwd_gcups, "peak": peak_gcups, "unit": "GCUPS", "frac": (fwd_gcups / peak_gcups) if fwd_gcups else None,
                    "peak_basis": f"{nsm} SMs x 64 alu lanes/clk/SM (measured, profiles/r01_ubench_dpx.json) x 2 cells / 4.5 alu instr x {sm_max:.0f} MHz max SM clock ({peak_src})",
                    "launch_ms": fwd_ms, "phase_ms": ph, "whole_step_frac": value / world / peak_gcups,
                    "hbm": {"achieved_gbs": m["algo_bytes"] / (fwd_ms * 1e-3) / 1e9 if fwd_ms > 0 else None, "peak_gbs": hbm_peak,
                            "frac": (m["algo_bytes"] / (fwd_ms * 1e-3) / 1e9 / hbm_peak) if fwd_ms > 0 else None, "algorithmic_bytes_per_launch": m["algo_bytes"]},
                    "traffic": None, "traffic_note": "not measured by this command (ncu is never run inside the bench); profiles/ holds the ncu capture of the same build"}
            clocks = m["clocks"]
            if clocks and clocks.get("sm_mhz") and fwd_gcups:
                roof["frac_at_observed_clock"] = fwd_gcups / (peak_gcups * clocks["sm_mhz"] / sm_max)
            cpu = None
            if world == 1:
                cpu = {"value": m["cpu"]["cells"] / m["cpu"]["secs"] / 1e9, "unit": "GCUPS", "cores": env["threads"], "kind": "reference" if "libssw_ref" in m["parity"]["against"] else "port",
                       "sample": f"{m['cpu']['pairs']} pairs (every k-th) of the timed batch, one pair per thread, {m['cpu']['secs']:.2f} s; GCUPS is intensive, so the figure extrapolates to the batch"}
            par = dict(m["parity"]); par["pairs"] = int(npar); par["mismatches"] = int(bad); par["note"] = "sample of the batch the e2e timing ran on, summed over ranks"
            line.update({"value": value, "steps": steps, "ms_per_step": dev_ms / steps, "scaling": "weak",
                         "config": {"workload": WL[cfg], "pairs_per_gpu": b.npairs, "cells_per_gpu": m["cells"], "flag": int(b.flag), "scoring": "+4/-6, N=-6, gapO 8, gapE 2, score_size 2",
                                    "sharding": f"{world} x independent shards, no collective",
                                    "l2": "inputs + column records (>5 GB per step) exceed the 126 MB L2; no flush needed" if cfg == 2 else "batch re-run back to back (see steps); column records written every step"},
                         "e2e": {"value": e2e2, "unit": "GCUPS", "ms_per_step": e2e2_ms / steps, "h2d_bytes_per_step": m["h2d2"], "d2h_bytes_per_step": m["d2h"],
                                 "api": "mpn_align_batch_packed2 (Engine.align_packed2): pinned host buffers, bases packed 4 per byte + exception list for N, offsets / maskLen / records as int arrays",
                                 "packed4_input": {"value": e2e4, "ms_per_step": e2e4_ms / steps, "h2d_bytes_per_step": m["h2d4"], "api": "mpn_align_batch_packed4 (Engine.align_packed4): bases nibble-packed (2 per byte)"},
                                 "int8_input": {"value": e2e, "ms_per_step": e2e_ms / steps, "h2d_bytes_per_step": m["h2d"], "api": "mpn_align_batch (Engine.align): one int8 code per base, the reference's own sequence format"}},
                         "gpu_launches": int(m["launches_per_step"] * steps), "roofline": roof, "parity": par, "cpu_baseline": cpu, "clocks": clocks})

    elif cfg == 3:
        steps = args.steps or 3
        m = run_config3(args, env, args.pairs or 200, steps, args.warmup, ref_regions=8)
        if rank == 0:
            line.update({"metric": "reads/s (realigner end to end)", "unit": "reads/s", "value": m["reads_per_s"], "steps": steps, "ms_per_step": 1e3 * m["s_per_pass"], "scaling": "weak",
                         "config": {"workload": WL[3], "regions": m["regions"], "reads": m["reads"], "ssw_pairs": m["ssw_pairs"], "ssw_cells": m["ssw_cells"]},
                         "e2e": {"value": m["reads_per_s"], "unit": "reads/s", "gcups": m["e2e_gcups"], "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                                 "note": "the realigner call IS host-in / host-out: value and e2e coincide; split_s gives fast pass / SW batch / CIGAR algebra"},
                         "detail": m, "parity": m.get("parity"), "cpu_baseline": m.get("cpu_baseline"), "gpu_launches": None})

    elif cfg == 5:
        steps = args.steps or 1
        npairs = args.pairs or 10_000_000
        flag = args.flag or 0
        m = run_config5(args, env, npairs, steps, args.warmup, flag)
        (dev_ms, wall_ms), sums = reduce_max_sum([m["dev_ms"], m["wall_ms"]], [m["cells"], m["parity_pairs"], m["parity_bad"], m["launches"], m["ref_secs"], m["ref_cells"], m["h2d"], m["d2h"]] + [float(x) for x in m["checksum"]])
        per_rank = torch.tensor([m["dev_ms"], m["wall_ms"], float(m["cells"]), float(m["chunks"])], dtype=torch.float64, device="cuda")
        allr = [torch.zeros_like(per_rank) for _ in range(world)]
        if world > 1:
            dist.all_gather(allr, per_rank)
        else:
            allr = [per_rank]
        if rank == 0:
            cells = sums[0]
            line.update({"value": cells / (dev_ms * 1e-3) / 1e9, "steps": steps, "ms_per_step": dev_ms / steps, "scaling": "strong",
                         "config": {"workload": WL[5], "pairs_total": npairs, "cells_total": m["total_cells"], "flag": flag, "chunks": m["nchunks"], "scoring": "+4/-6, N=-6, gapO 8, gapE 2, score_size 2",
                                    "sharding": f"length-sorted chunks dealt to {world} rank(s), heaviest first to the least loaded; no collective, checksum + parity counts reduced on rank 0",
                                    "warmup_note": f"warm-up = {m['warmup_chunks']} small chunks of the same stream per rank (a full pass is minutes on one GPU)",
                                    "l2": "every chunk is new data (hundreds of MB of bases + column records), far above the 126 MB L2"},
                         "e2e": {"value": cells / (wall_ms * 1e-3) / 1e9, "unit": "GCUPS", "ms_per_step": wall_ms / steps, "h2d_bytes_per_step": int(sums[6]), "d2h_bytes_per_step": int(sums[7]),
                                 "note": "wall time of the pass: host generation runs ahead on a producer thread; upload (H2D), scheduling, kernels, fetch (D2H) and record conversion are inside"},
                         "gpu_launches": int(sums[3]),
                         "per_rank": [{"dev_ms": float(t[0]), "wall_ms": float(t[1]), "cells": float(t[2]), "chunks": int(t[3])} for t in allr],
                         "checksum": {"score1": int(sums[8]), "ref_end1": int(sums[9]), "read_end1": int(sums[10]), "pairs": int(sums[11])},
                         "parity": {"pairs": int(sums[1]), "mismatches": int(sums[2]), "against": "oracle/_ref/libssw_ref.so (unmodified ssw.c)", "note": "every 1000th pair of the length-sorted stream (at most 4000 per rank)"},
                         "roofline": {"bound": "alu", "kernel": "sw_long16_kernel + sw_strip16_kernel (all phases)", "achieved": cells / (dev_ms * 1e-3) / 1e9 / world, "peak": peak_gcups, "unit": "GCUPS",
                                      "frac": cells / (dev_ms * 1e-3) / 1e9 / world / peak_gcups, "traffic": None},
                         "cpu_baseline": {"value": sums[5] / sums[4] / 1e9 * m["ref_threads"] if sums[4] > 0 else None, "unit": "GCUPS", "cores": m["ref_threads"], "kind": "reference",
                                          "sample": f"the {int(sums[1])} parity pairs, one pair per thread ({sums[4]:.1f} core-seconds)"} if world == 1 else None})
    else:
        raise SystemExit("unknown --config")

    # ---- bounded runs of the other configs on the default line (N = 1)
    if rank == 0 and world == 1 and cfg == 2 and not args.no_configs and not args.pairs:
        other = {}
        try:
            b1 = w.config1(10_000, seed=11)
            m1 = run_pairs(args, env, b1, 200, 3, 10_000, want_clocks=False)
            other["1"] = {"workload": WL[1], "kernel_gcups": m1["cells"] * 200 / (m1["dev_ms"] * 1e-3) / 1e9, "e2e_gcups": m1["cells"] * 200 / (m1["e2e_ms"] * 1e-3) / 1e9,
                          "ms_per_batch": m1["dev_ms"] / 200, "parity": m1["parity"], "cpu_gcups": m1["cpu"]["cells"] / m1["cpu"]["secs"] / 1e9}
            m3 = run_config3(args, env, 200, 2, 2, ref_regions=4)
            other["3"] = {"workload": WL[3], **{k: m3[k] for k in ("regions", "reads", "ssw_pairs", "reads_per_s", "e2e_gcups", "gpu_gcups_ssw_only", "s_per_pass")},
                          "parity": m3.get("parity"), "cpu_reads_per_s": (m3.get("cpu_baseline") or {}).get("reads_per_s")}
            for fl in (0, 1):
                b4 = make_pair_workload(w, 4, 1024, seed=14, flag=fl, threads=env["threads"])
                m4 = run_pairs(args, env, b4, 2, 1, 16, cap=8192, want_clocks=False)
                other[f"4_flag{fl}"] = {"workload": WL[4], "pairs": 1024, "flag": fl, "kernel_gcups": m4["cells"] * 2 / (m4["dev_ms"] * 1e-3) / 1e9, "e2e_gcups": m4["cells"] * 2 / (m4["e2e_ms"] * 1e-3) / 1e9,
                                        "phase_ms": m4["phase_ms"], "parity": m4["parity"], "cpu_gcups": m4["cpu"]["cells"] / m4["cpu"]["secs"] / 1e9}
            m5 = run_config5(args, env, 20_000, 1, 1, 0, sample_every=100)
            other["5_sample"] = {"workload": WL[5], "pairs": 20_000, "flag": 0, "kernel_gcups": m5["cells"] / (m5["dev_ms"] * 1e-3) / 1e9, "e2e_gcups": m5["cells"] / (m5["wall_ms"] * 1e-3) / 1e9,
                                 "parity": {"pairs": m5["parity_pairs"], "mismatches": m5["parity_bad"]},
                                 "cpu_gcups": m5["ref_cells"] / m5["ref_secs"] / 1e9 * m5["ref_threads"] if m5["ref_secs"] else None,
                                 "note": "20 k-pair sample of the length distribution; the full 10 M-pair stream is `bench.py --config 5`"}
        except Exception as ex:          # a failing side config must not lose the main line
            other["error"] = repr(ex)[:300]
        line["configs"] = other

    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
