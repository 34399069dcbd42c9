"""CPU tests of the checker itself: the scalar restatement (oracle/ssw_oracle.c) against the golden vectors produced by the
compiled reference, against known answers, and -- when the compiled reference is present -- a live fuzz against it."""
import importlib

import numpy as np
import pytest

from oracle import oracle
from golden_util import golden_cases, diff, CAP

w = importlib.import_module("workloads")


def run(b, impl, threads=4):
    return oracle.run_batch(b.reads, b.read_off, b.refs, b.ref_off, b.masklen, b.mat, b.n, gapO=b.gapO, gapE=b.gapE, flag=b.flag, filters=b.filters,
                            filterd=b.filterd, score_size=b.score_size, threads=threads, impl=impl, cigar_cap=CAP)[:2]


def test_port_matches_golden_vectors():
    total = 0
    for k, b, res, cig in golden_cases():
        if b.read_len.max() > 5000:          # the scalar port needs minutes on the ONT-scale cases; they are covered on the GPU side
            continue
        r, c = run(b, "port")
        bad = diff(r, c, res, cig)
        assert len(bad) == 0, (k, b.name, bad[:5], r[bad[:1]], res[bad[:1]])
        total += b.npairs
    assert total > 3000


def test_known_answers():
    # ACGTACGT vs TTACGTACGTTTACGTACGT, flag 1, mask 15 -> 32, ref [2, 9], read [0, 7], "8M"; earliest of two equal hits wins
    enc = lambda s: np.array(["ACGT".index(c) for c in s], dtype=np.int8)
    read, ref = enc("ACGTACGT"), enc("TTACGTACGTTTACGTACGT")
    b = w.PairBatch(read, np.array([0, 8]), ref, np.array([0, 20]), np.array([15], dtype=np.int32), flag=1)
    r, c = run(b, "port")
    assert list(r[0]) == [32, 0, 2, 9, 0, 7, 0, 1] and c[0, 0] == (8 << 4)
    # nothing aligns: score 0, ref_end1 -1 (byte mode), cigar "1M" when a cigar is requested
    b = w.PairBatch(enc("AAAA"), np.array([0, 4]), enc("CCCCCC"), np.array([0, 6]), np.array([15], dtype=np.int32), flag=1)
    r, c = run(b, "port")
    assert list(r[0]) == [0, 0, -1, -1, 0, 0, 0, 1] and c[0, 0] == (1 << 4)
    b.flag = 0
    r, c = run(b, "port")
    assert list(r[0]) == [0, 0, -1, -1, -1, 0, 0, 0]


@pytest.mark.skipif(not oracle.have_ref(), reason="compiled reference (oracle/_ref) not present")
def test_port_matches_compiled_reference_fuzz():
    for seed in range(200, 212):
        flag = [1, 0, 8, 0x0f, 2, 4][seed % 6]
        b = w.fuzz_pairs(120, seed, alphabet=2 if seed % 3 == 0 else 4, flag=flag, max_read=300, max_ref=350)
        b.filters = 80 if flag == 2 else 0
        b.filterd = 25 if flag == 4 else 32767
        r, c = run(b, "ref")
        p, cp = run(b, "port")
        bad = diff(r, c, p, cp)
        assert len(bad) == 0, (seed, bad[:5])


@pytest.mark.skipif(not oracle.have_ref(), reason="compiled reference (oracle/_ref) not present")
def test_golden_vectors_are_current():
    """the committed fixture must be what the compiled reference produces today"""
    for k, b, res, cig in golden_cases():
        if k % 5 or b.read_len.max() > 5000:
            continue
        r, c = run(b, "ref")
        assert len(diff(r, c, res, cig)) == 0, (k, b.name)
