"""Shared helpers of the realigner tests: golden regions (tests/golden/realign_golden.json.gz, made from the compiled reference
by tests/golden/make_golden_realign.py) and a clean-subprocess runner for the compiled reference realigner."""
import dataclasses
import gzip
import importlib
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
w = importlib.import_module("workloads")


def golden_regions():
    with gzip.open(os.path.join(HERE, "golden", "realign_golden.json.gz"), "rt") as f:
        doc = json.load(f)
    return [(w.RegionWorkload(**d["region"]), d["positions"], d["cigars"]) for d in doc]


def run_reference(regions, lib_path):
    """the reference object must run without numpy in the process (see ref_realigner_runner.py)"""
    p = subprocess.run([sys.executable, os.path.join(HERE, "ref_realigner_runner.py"), lib_path],
                       input=json.dumps([dataclasses.asdict(r) for r in regions]).encode(), capture_output=True, check=True)
    return [(a[0], a[1]) for a in json.loads(p.stdout)]


def mismatches(got, want):
    bad = []
    for k, ((gp, gc), (wp, wc)) in enumerate(zip(got, want)):
        for i in range(len(wp)):
            if gp[i] != wp[i] or gc[i] != wc[i]:
                bad.append((k, i, (gp[i], gc[i]), (wp[i], wc[i])))
    return bad
