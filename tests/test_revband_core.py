"""Banded reverse pass (megapath-nano_b200/csrc/sw_revband_core.h): the per-lane routine the kernel runs, compiled for the host with its
packed arithmetic emulated, against the compiled reference's begin positions (ssw.c:820-832).  Pins the band bound of the header comment
-- every alignment reaching score1 in the reversed sub-rectangle stays within (X - gapO + gapE) / gapE diagonals below and
(X - gapO + gapE) / (match + gapE) above the main diagonal, X = match * (read_end1 + 1) - score1 -- on degenerate alphabets and cheap gaps,
where co-optimal alignments abound."""
import ctypes as ct
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402
import workloads as w  # noqa: E402


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("revband") / "librevband_host.so")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-o", out, os.path.join(ROOT, "tests", "cpp", "revband_host.cpp")], check=True)
    L = ct.CDLL(out)
    L.revband_host.restype = ct.c_int
    L.revband_host.argtypes = [ct.c_void_p, ct.c_int64, ct.c_int64] + [ct.c_int] * 9 + [ct.POINTER(ct.c_int)] * 4
    return L


def run(lib, batch, n_is_mm=1, force_nw=0, cigar_cap=64):
    if not oracle.have_ref():
        pytest.skip("compiled reference not built (oracle/_ref)")
    out, _, _ = oracle.run_batch(batch.reads, batch.read_off, batch.refs, batch.ref_off, batch.masklen, batch.mat, batch.n, batch.gapO, batch.gapE, flag=1, threads=4, cigar_cap=cigar_cap)
    mat = batch.mat.reshape(5, 5)
    mt, mm = int(mat[0, 0]), int(mat[0, 1])
    pad = 64
    arena = np.zeros(pad + len(batch.reads) + len(batch.refs) + pad + 8, dtype=np.int8)
    a = arena[(-arena.ctypes.data) % 8:]
    a[pad:pad + len(batch.reads)] = batch.reads
    rbase = pad + len(batch.reads)
    a[rbase:rbase + len(batch.refs)] = batch.refs
    stats = {"done": 0, "bad": 0, "ineligible": 0, "bailed": 0}
    for p in range(batch.npairs):
        S, _, rb_, re_, qb, qe = [int(x) for x in out[p, :6]]
        if S <= 0:
            continue
        col, row, nw, h0 = ct.c_int(), ct.c_int(), ct.c_int(), ct.c_int()
        rc = lib.revband_host(a.ctypes.data, pad + int(batch.read_off[p]) + qe, rbase + int(batch.ref_off[p]) + re_, qe + 1, re_ + 1, S, mt, mm,
                              batch.gapO, batch.gapE, n_is_mm, force_nw, ct.byref(col), ct.byref(row), ct.byref(nw), ct.byref(h0))
        if rc == 2:
            stats["ineligible"] += 1
        elif rc == 1:
            stats["bailed"] += 1
        else:
            stats["done"] += 1
            stats["bad"] += (col.value, row.value) != (re_ - rb_, qe - qb)
    return stats


def test_short_reads_like_configs1(lib):
    st = run(lib, w.make_pairs(600, (150, 300), 1000, err=0.02, seed=12, flag=1))
    assert st["bad"] == 0 and st["done"] > 550 and st["bailed"] == 0


def test_wider_class_gives_the_same_cell(lib):
    st = run(lib, w.make_pairs(200, (150, 300), 1000, err=0.02, seed=13, flag=1), force_nw=20)
    assert st["bad"] == 0 and st["done"] > 180


def test_tiny_and_noisy(lib):
    assert run(lib, w.make_pairs(400, (1, 40), 60, err=0.05, seed=2, flag=1))["bad"] == 0
    st = run(lib, w.make_pairs(300, (50, 120), 300, err=0.10, seed=3, flag=1))
    assert st["bad"] == 0 and st["done"] > 100


def test_degenerate_alphabets_and_cheap_gaps(lib):
    b = w.make_pairs(500, (30, 80), 160, err=0.08, seed=8, flag=1)
    b.reads = (b.reads & 1).astype(np.int8); b.refs = (b.refs & 1).astype(np.int8); b.mat = w.dna_matrix(2, 3); b.gapO = 4; b.gapE = 1
    assert run(lib, b)["bad"] == 0
    b = w.make_pairs(500, (30, 80), 160, err=0.3, seed=9, flag=1)
    b.reads = (b.reads % 3 == 0).astype(np.int8); b.refs = (b.refs % 3 == 0).astype(np.int8); b.mat = w.dna_matrix(1, 1); b.gapO = 2; b.gapE = 1
    assert run(lib, b)["bad"] == 0
    b = w.make_pairs(400, (30, 80), 160, err=0.1, seed=10, flag=1); b.mat = w.dna_matrix(3, 1); b.gapO = 1; b.gapE = 1
    assert run(lib, b)["bad"] == 0


def test_n_bases(lib):
    st = run(lib, w.make_pairs(400, (60, 200), 400, err=0.03, seed=21, flag=1, n_frac=0.01), n_is_mm=1)
    assert st["bad"] == 0 and st["bailed"] == 0 and st["done"] > 300
    b = w.make_pairs(400, (60, 200), 400, err=0.03, seed=22, flag=1, n_frac=0.003); b.mat = w.dna_matrix(4, 6, n_zero=True)
    st = run(lib, b, n_is_mm=0)
    assert st["bad"] == 0 and st["bailed"] > 50 and st["done"] > 100


def test_random_scoring_and_shapes(lib):
    """random match / mismatch / gap scores (gapO > gapE >= 1), lengths, error rates and alphabet sizes (a longer run of this loop, 2.9 M banded
    pairs, found no difference)"""
    rng = np.random.default_rng(5)
    done = 0
    for _ in range(40):
        mt = int(rng.integers(1, 9)); mm = int(rng.integers(1, 9)); ge = int(rng.integers(1, 5)); go = ge + int(rng.integers(1, 10))
        lo = int(rng.integers(1, 150)); hi = lo + int(rng.integers(0, 200)); fl = int(hi * rng.uniform(1.0, 3.0)) + 8
        alpha = int(rng.choice([4, 4, 2, 3]))
        b = w.make_pairs(100, (lo, hi), fl, err=float(rng.choice([0.0, 0.01, 0.03, 0.08, 0.2])), seed=int(rng.integers(1 << 30)), flag=1, n_frac=float(rng.choice([0, 0, 0.01])))
        if alpha < 4:
            b.reads = np.where(b.reads < 4, b.reads % alpha, b.reads).astype(np.int8)
            b.refs = np.where(b.refs < 4, b.refs % alpha, b.refs).astype(np.int8)
        b.mat = w.dna_matrix(mt, mm); b.gapO = go; b.gapE = ge
        st = run(lib, b, cigar_cap=2048)
        assert st["bad"] == 0, (mt, mm, go, ge, lo, hi, fl, alpha)
        done += st["done"]
    assert done > 2000
