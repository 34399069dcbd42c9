"""GPU tier: the C++ front end (include/ssw_cpp.h, StripedSmithWaterman::Aligner) compiled into a small program the way a C++ caller
would use it.  Align() per query, the batched AlignBatch() and the reference's post-processing (soft clips, '=' / 'X' runs, mismatch
count: ssw_cpp.cpp:50-86,123-207) are checked against the compiled reference ssw.c + a Python restatement of that post-processing."""
import importlib
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "megapath-nano_b200")
w = importlib.import_module("workloads")


def expected(ref, q, res, cig):
    """ConvertAlignment + CalculateNumberMismatch on one s_align record"""
    score, score2, rb, re_, qb, qe, re2, n = [int(x) for x in res]
    out, mism, ti, qi, run, isx = [], 0, rb, qb, 0, False
    if qb > 0:
        out.append(f"{qb}S")

    def flush():
        nonlocal run
        if run:
            out.append(f"{run}{'X' if isx else '='}")
        run = 0

    for wd in cig[:n]:
        ln, op = int(wd) >> 4, int(wd) & 15
        if op == 0:
            for _ in range(ln):
                x = ref[ti] != q[qi]
                if run and x != isx:
                    flush()
                isx = x; run += 1; mism += x; ti += 1; qi += 1
        elif op == 1:
            flush(); qi += ln; mism += ln; out.append(f"{ln}I")
        else:
            flush(); ti += ln; mism += ln; out.append(f"{ln}D")
    flush()
    if len(q) - qe - 1 > 0:
        out.append(f"{len(q) - qe - 1}S")
    return [score, score2, rb, re_, qb, qe, re2, mism, "".join(out)]


def test_aligner_class_against_reference(tmp_path):
    from oracle import oracle
    exe = tmp_path / "probe"
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), "-o", str(exe), os.path.join(ROOT, "tests", "cpp", "ssw_cpp_probe.cpp"),
                    os.path.join(PKG, "libmpn_ssw.so"), "-Wl,-rpath," + PKG], check=True)
    rng = np.random.default_rng(17)
    ref = "".join("ACGT"[i] for i in rng.integers(0, 4, size=700))
    queries = []
    for k in range(30):
        s = int(rng.integers(0, 400)); q = list(ref[s:s + int(rng.integers(60, 280))])
        for _ in range(int(rng.integers(0, 5))):
            p = int(rng.integers(5, len(q) - 5)); q[p] = "ACGT"[int(rng.integers(0, 4))]
        if k % 3 == 0:
            p = int(rng.integers(10, len(q) - 10)); q[p:p] = list("GATTACA"[:1 + k % 5])
        if k % 4 == 1:
            p = int(rng.integers(10, len(q) - 12)); del q[p:p + 1 + k % 6]
        if k % 7 == 0:
            q = list("TTTTTTTT") + q + list("GGGGGG")
        if k % 9 == 0:
            q[len(q) // 2] = "N"
        queries.append("".join(q))
    inp = tmp_path / "in.txt"
    inp.write_text(ref + "\n" + "\n".join(queries) + "\n")
    lines = subprocess.run([str(exe), str(inp)], capture_output=True, text=True, check=True).stdout.strip().split("\n")
    assert len(lines) == len(queries) + 1
    # the reference's view of the same pairs: flag 0x0f, filters 0 / 32767, maskLen = query length (ssw_cpp.cpp:343-346), N row = -mismatch
    enc = lambda s: np.array(["ACGTN".index(c) for c in s], dtype=np.int8)
    reads = [enc(q) for q in queries]
    ro = np.concatenate([[0], np.cumsum([len(r) for r in reads])]).astype(np.int64)
    refs = np.tile(enc(ref), len(queries)); fo = np.arange(len(queries) + 1, dtype=np.int64) * len(ref)
    oracle.require_ref()
    impl = "ref"
    res, cig, _ = oracle.run_batch(np.concatenate(reads), ro, refs, fo, np.array([len(q) for q in queries], np.int32), w.dna_matrix(4, 6), 5, gapO=8, gapE=2,
                                   flag=0x0f, filters=0, filterd=32767, threads=4, impl=impl, cigar_cap=256)
    for k, q in enumerate(queries):
        f = lines[k].split()
        assert f[0] == "1", ("Align and AlignBatch disagree", k)
        want = expected(enc(ref), enc(q), res[k], cig[k])
        got = [int(x) for x in f[1:9]] + [f[9] if len(f) > 9 else ""]
        assert got == want, (k, got, want)
    assert lines[-1].split()[1] == lines[0].split()[1]
