"""The boundary, proven with the reference's OWN callers: the unmodified files of the reference --
bin/realignment/pyssw.py, fast_align_reads2ref.py, realign/ssw_cpp.cpp, realign/realigner.cpp -- copied as they are into the git-ignored
baseline/_ref/ by oracle.build_ref_callers(), are run against the PRODUCT library and must return what they return on the compiled
reference ssw.c.  Nothing in this file goes through the product's Python mirrors."""
import importlib.util
import json
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "megapath-nano_b200")
PRODUCT_LIBSSW = os.path.join(PKG, "realign", "libssw.so")


@pytest.fixture(scope="module")
def callers():
    from oracle import oracle
    oracle.require_ref()
    d = oracle.build_ref_callers()
    if d is None:
        pytest.fail("baseline/_ref (the reference's own callers) missing: run __graft_entry__.build() where /root/reference exists")
    return d


def queries_601(seed, n):
    """the ONT path's shape: ~400 bp queries (+-200 bp window with edits) against a 601 bp reference"""
    rng = np.random.default_rng(seed)
    ref = "".join("ACGT"[i] for i in rng.integers(0, 4, size=601))
    qs = []
    for k in range(n):
        s = int(rng.integers(0, 150)); q = list(ref[s:s + int(rng.integers(300, 440))])
        for _ in range(int(rng.integers(0, 9))):
            p = int(rng.integers(5, len(q) - 5)); q[p] = "ACGT"[int(rng.integers(0, 4))]
        if k % 2 == 0:
            p = int(rng.integers(10, len(q) - 10)); q[p:p] = list("ACCGTTAGGCAT"[:1 + k % 11])
        if k % 3 == 1:
            p = int(rng.integers(10, len(q) - 20)); del q[p:p + 1 + k % 9]
        if k % 5 == 0:
            q[len(q) // 3] = "N"
        qs.append("".join(q))
    qs.append("ACGTACGTACGTAC")            # <= 30 bases: the maskLen = 15 branch of pyssw.py:142
    return ref, qs


def test_reference_pyssw_module_on_the_product_library(callers):
    """pyssw.py (unmodified): SSW(lib_path=<product libssw.so>).align == the same module on the compiled reference ssw.c"""
    from oracle import oracle
    spec = importlib.util.spec_from_file_location("reference_pyssw", os.path.join(callers, "pyssw.py"))
    ref_pyssw = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_pyssw)
    ref, qs = queries_601(41, 40)
    out = {}
    for name, lib in (("product", PRODUCT_LIBSSW), ("reference", oracle.ref_path())):
        a = ref_pyssw.SSW(lib_path=lib)
        a.set_reference_sequence(ref)
        out[name] = [a.align(q) for q in qs]
    assert out["product"] == out["reference"]
    assert all(score > 0 and cigar for score, cigar, _ in out["product"][:-1])


DRIVER = r"""
import json, sys
sys.path.insert(0, sys.argv[1])
from fast_align_reads2ref import FastPassAligner
job = json.load(sys.stdin)
a = FastPassAligner(ctg_name="chr", consensus=job["consensus"], reference=job["reference"], reference_start=job["start"], read_name_list=job["names"])
json.dump(a.align_reads(), sys.stdout)
"""


def test_reference_fast_pass_aligner_zero_edit(callers, tmp_path):
    """fast_align_reads2ref.py + pyssw.py (unmodified) in a directory whose realign/libssw.so is the product: the ONT path's call
    sequence (local_realignment.py:321-327 -> FastPassAligner.align_reads) without a single edit on the reference side"""
    from oracle import oracle
    ref, qs = queries_601(43, 24)
    job = json.dumps({"consensus": qs[:-1], "reference": ref, "start": 1000, "names": [f"r{k}" for k in range(len(qs) - 1)]}).encode()
    got = {}
    for name, lib in (("product", PRODUCT_LIBSSW), ("reference", oracle.ref_path())):
        d = tmp_path / name
        (d / "realign").mkdir(parents=True)
        for f in ("pyssw.py", "fast_align_reads2ref.py"):
            shutil.copyfile(os.path.join(callers, f), d / f)
        shutil.copyfile(lib, d / "realign" / "libssw.so")
        p = subprocess.run([sys.executable, "-c", DRIVER, str(d)], input=job, capture_output=True, check=True)
        got[name] = json.loads(p.stdout)
    assert got["product"] == got["reference"] and len(got["product"]) == len(qs) - 1


def test_reference_cpp_sources_linked_to_the_product_library(callers):
    """the reference's ssw_cpp.cpp + realigner.cpp, compiled against include/ssw.h and linked to the product
    libssw.so instead of ssw.c, reproduce the golden regions of the all-CPU reference realigner"""
    from realign_util import golden_regions, mismatches
    import dataclasses
    gold = golden_regions()
    obj = os.path.join(callers, "realigner_refsrc_on_product")
    assert os.path.exists(obj)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ref_realigner_runner.py"), obj],
                       input=json.dumps([dataclasses.asdict(rg) for rg, _, _ in gold]).encode(), capture_output=True, check=True)
    got = [(a[0], a[1]) for a in json.loads(p.stdout)]
    assert not mismatches(got, [(pp, cc) for _, pp, cc in gold])


def test_aligner_class_equals_the_reference_class_field_by_field(callers, tmp_path):
    """StripedSmithWaterman::Aligner of the product against the reference's own ssw_cpp.cpp + ssw.c through the same probe program: every
    Alignment field including `mismatches` (CalculateNumberMismatch, ssw_cpp.cpp:123-207) and the raw cigar words"""
    exe = tmp_path / "probe_product"
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), "-o", str(exe), os.path.join(ROOT, "tests", "cpp", "ssw_cpp_probe_common.cpp"),
                    os.path.join(PKG, "libmpn_ssw.so"), "-Wl,-rpath," + PKG], check=True)
    rng = np.random.default_rng(19)
    ref = "".join("ACGT"[i] for i in rng.integers(0, 4, size=900))
    qs = []
    for k in range(60):
        s = int(rng.integers(0, 600)); q = list(ref[s:s + int(rng.integers(40, 290))])
        for _ in range(int(rng.integers(0, 6))):
            p = int(rng.integers(2, len(q) - 2)); q[p] = "ACGT"[int(rng.integers(0, 4))]
        if k % 3 == 0:
            p = int(rng.integers(5, len(q) - 5)); q[p:p] = list("GATTACA"[:1 + k % 7])
        if k % 4 == 1:
            p = int(rng.integers(5, len(q) - 12)); del q[p:p + 1 + k % 6]
        if k % 7 == 0:
            q = list("TTTTTTTTTT") + q + list("GGGGGGG")
        if k % 9 == 0:
            q[len(q) // 2] = "N"
        qs.append("".join(q))
    inp = tmp_path / "in.txt"
    inp.write_text(ref + "\n" + "\n".join(qs) + "\n")
    want = subprocess.run([os.path.join(callers, "ssw_cpp_probe_ref"), str(inp)], capture_output=True, text=True, check=True).stdout
    got = subprocess.run([str(exe), str(inp)], capture_output=True, text=True, check=True).stdout
    assert got == want and len(got.strip().split("\n")) == len(qs) + 4
