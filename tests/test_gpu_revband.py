"""Banded reverse pass on the device (csrc/sw_revband.cuh): begin positions and everything downstream of them (CIGAR) against the
compiled reference, on batches chosen to hit every route -- all four band classes, pairs the band refuses (wide bands -> the N variants
of the packed kernel), N bases with N = mismatch (handled in the band) and N = 0 (bail out), degenerate alphabets with cheap gaps."""
import importlib
import os

import numpy as np
import pytest

from golden_util import diff, CAP

pytestmark = pytest.mark.gpu

w = importlib.import_module("workloads")
B = importlib.import_module("megapath-nano_b200.batch")


@pytest.fixture(scope="module")
def eng():
    e = B.Engine(0)
    yield e
    e.close()


def check(eng, b, cap=CAP):
    from oracle import oracle
    oracle.require_ref()
    rec, cig = eng.align(b)
    g, gc = B.as_table(rec, cig, cap)
    r, c = oracle.run_batch(b.reads, b.read_off, b.refs, b.ref_off, b.masklen, b.mat, b.n, gapO=b.gapO, gapE=b.gapE, flag=b.flag, filters=b.filters,
                            filterd=b.filterd, score_size=b.score_size, threads=os.cpu_count() or 8, impl="ref", cigar_cap=cap)[:2]
    bad = diff(g, gc, r, c)
    assert len(bad) == 0, (b.name, len(bad), bad[:5], g[bad[:1]], r[bad[:1]])


def test_all_band_classes_and_refused_pairs(eng):
    check(eng, w.make_pairs(30_000, (150, 300), 1000, err=0.02, seed=31, flag=1, name="2 % errors"))
    check(eng, w.make_pairs(8_000, (50, 400), 600, err=0.10, seed=32, flag=1, name="10 % errors: many pairs beyond 64 diagonals"), cap=256)
    check(eng, w.make_pairs(6_000, (1, 60), 90, err=0.05, seed=33, flag=1, name="tiny reads"))


def test_n_bases(eng):
    check(eng, w.make_pairs(10_000, (60, 250), 500, err=0.03, seed=34, flag=1, n_frac=0.01, name="N = mismatch"))
    b = w.make_pairs(10_000, (60, 250), 500, err=0.03, seed=35, flag=1, n_frac=0.003, name="N = 0")
    b.mat = w.dna_matrix(4, 6, n_zero=True)
    check(eng, b)


def test_degenerate_alphabets_and_other_scores(eng):
    b = w.make_pairs(8_000, (30, 120), 200, err=0.08, seed=36, flag=1, name="binary alphabet")
    b.reads = (b.reads & 1).astype(np.int8); b.refs = (b.refs & 1).astype(np.int8); b.mat = w.dna_matrix(2, 3); b.gapO = 4; b.gapE = 1
    check(eng, b, cap=256)
    b = w.make_pairs(8_000, (30, 120), 200, err=0.3, seed=37, flag=1, name="skewed binary, match 1")
    b.reads = (b.reads % 3 == 0).astype(np.int8); b.refs = (b.refs % 3 == 0).astype(np.int8); b.mat = w.dna_matrix(1, 1); b.gapO = 2; b.gapE = 1
    check(eng, b, cap=256)
    b = w.make_pairs(8_000, (100, 300), 500, err=0.04, seed=38, flag=1, name="match 2 / mismatch 2, gaps 3 / 1")
    b.mat = w.dna_matrix(2, 2); b.gapO = 3; b.gapE = 1
    check(eng, b)


def test_mixed_with_long_reads(eng):
    """packed bins (banded) and the multi-strip bin (full reverse pass) in one batch"""
    a = w.make_pairs(3_000, (100, 300), 600, err=0.02, seed=39, flag=1)
    c = w.make_pairs(40, (2_000, 3_000), 1.2, err=0.05, seed=40, flag=1)
    b = w.concat_batches([a, c]) if hasattr(w, "concat_batches") else None
    if b is None:
        reads = np.concatenate([a.reads, c.reads]); refs = np.concatenate([a.refs, c.refs])
        ro = np.concatenate([a.read_off, c.read_off[1:] + a.read_off[-1]]); fo = np.concatenate([a.ref_off, c.ref_off[1:] + a.ref_off[-1]])
        b = w.PairBatch(reads, ro, refs, fo, np.concatenate([a.masklen, c.masklen]), flag=1, name="short + long")
    check(eng, b, cap=2048)
