"""Parity tests proper: the CUDA path, called through the C ABI, against (a) the golden vectors of the compiled reference,
(b) the oracle on seeded inputs, (c) size-independent properties at BASELINE sizes.  Bit-exact on every field and CIGAR word."""
import ctypes as ct
import importlib
import os

import numpy as np
import pytest

from golden_util import golden_cases, diff, CAP

pytestmark = pytest.mark.gpu

w = importlib.import_module("workloads")
B = importlib.import_module("megapath-nano_b200.batch")
pyssw = importlib.import_module("megapath-nano_b200.pyssw")


@pytest.fixture(scope="module")
def eng():
    e = B.Engine(0)           # raises if the CUDA library or the GPU is missing: no fallback
    yield e
    e.close()


def gpu_table(eng, b, cap=CAP):
    rec, cig = eng.align(b)
    return B.as_table(rec, cig, cap) + (rec, cig)


def oracle_table(b, cap=CAP, threads=8):
    from oracle import oracle
    oracle.require_ref()
    impl = "ref"
    return oracle.run_batch(b.reads, b.read_off, b.refs, b.ref_off, b.masklen, b.mat, b.n, gapO=b.gapO, gapE=b.gapE, flag=b.flag, filters=b.filters,
                            filterd=b.filterd, score_size=b.score_size, threads=threads, impl=impl, cigar_cap=cap)[:2]


def test_golden_vectors(eng):
    npairs = 0
    for k, b, res, cig in golden_cases():
        g, gc, _, _ = gpu_table(eng, b)
        bad = diff(g, gc, res, cig)
        assert len(bad) == 0, (k, b.name, bad[:5], g[bad[:1]], res[bad[:1]])
        npairs += b.npairs
    assert npairs > 3000


@pytest.mark.parametrize("seed", range(300, 316))
def test_fuzz_against_oracle(eng, seed):
    flag = [1, 0, 8, 0x0f, 2, 4, 3, 5][seed % 8]
    b = w.fuzz_pairs(250, seed, alphabet=2 if seed % 3 == 0 else 4, flag=flag)
    b.filters = 90 if flag in (2, 3) else 0
    b.filterd = 35 if flag in (4, 5) else 32767
    g, gc, _, _ = gpu_table(eng, b)
    r, c = oracle_table(b)
    bad = diff(g, gc, r, c)
    assert len(bad) == 0, (seed, bad[:5], g[bad[:1]], r[bad[:1]])


def test_config1_full_size(eng):
    """BASELINE configs[0] at its full size: 10 000 pairs, 250 bp x 500 bp, flag 0"""
    b = w.config1()
    g, gc, _, _ = gpu_table(eng, b)
    r, c = oracle_table(b, threads=os.cpu_count() or 8)
    assert len(diff(g, gc, r, c)) == 0


def test_config2_sample_against_oracle(eng):
    b = w.config2(60_000, seed=77)
    g, gc, _, _ = gpu_table(eng, b)
    r, c = oracle_table(b, threads=os.cpu_count() or 8)
    assert len(diff(g, gc, r, c)) == 0


def _rescore(b, rec, cig, i):
    """score of the reported alignment path: must equal score1 (a property that needs no oracle)"""
    mat = np.asarray(b.mat).reshape(b.n, b.n)
    read = b.reads[b.read_off[i]:b.read_off[i + 1]]; ref = b.refs[b.ref_off[i]:b.ref_off[i + 1]]
    qi, ti = int(rec["read_begin1"][i]), int(rec["ref_begin1"][i])
    o, l = int(rec["cigar_off"][i]), int(rec["cigar_len"][i])
    s = 0
    for wd in cig[o:o + l]:
        n, op = int(wd) >> 4, int(wd) & 15
        if op == 0:
            s += int(mat[ref[ti:ti + n], read[qi:qi + n]].sum()); qi += n; ti += n
        elif op == 1:
            s -= b.gapO + (n - 1) * b.gapE; qi += n
        else:
            s -= b.gapO + (n - 1) * b.gapE; ti += n
    return s, qi - 1, ti - 1


def test_config2_full_size_properties(eng):
    """BASELINE configs[1] at a size the oracle would need minutes for: CIGAR paths re-score to score1, spans are consistent,
    results do not depend on batch order, and a re-run is identical."""
    n = 300_000
    b = w.config2(n, seed=78)
    rec, cig = eng.align(b)
    assert (rec["status"] == 0).all()
    rng = np.random.default_rng(1)
    for i in rng.integers(0, n, size=400):
        s, qe, te = _rescore(b, rec, cig, int(i))
        assert s == int(rec["score1"][i]) and qe == int(rec["read_end1"][i]) and te == int(rec["ref_end1"][i]), i
    assert (rec["score1"] >= rec["score2"]).all()
    assert ((rec["ref_begin1"] >= 0) & (rec["ref_begin1"] <= rec["ref_end1"]) & (rec["ref_end1"] < 1000)).all()
    rec2, cig2 = eng.align(b)
    t1, c1 = B.as_table(rec, cig, 32); t2, c2 = B.as_table(rec2, cig2, 32)
    assert np.array_equal(t1, t2) and np.array_equal(c1, c2)
    perm = rng.permutation(20_000)
    sb = b.subset(perm)
    rec3, cig3 = eng.align(sb)
    t3, c3 = B.as_table(rec3, cig3, 32)
    assert np.array_equal(t3, t1[perm]) and np.array_equal(c3, c1[perm])


def test_edge_shapes(eng):
    """empty batch, 1-base sequences, reads longer than targets, all-N reads, reads needing the 32-bit kernel"""
    dna = w.dna_matrix()
    e = w.PairBatch(np.zeros(0, np.int8), np.zeros(1, np.int64), np.zeros(0, np.int8), np.zeros(1, np.int64), np.zeros(0, np.int32), mat=dna, flag=1)
    rec, cig = eng.align(e)
    assert len(rec) == 0
    rng = np.random.default_rng(3)
    reads, refs = [], []
    for rl, fl in ((1, 1), (1, 40), (40, 1), (16, 16), (17, 300), (300, 17), (129, 129), (1281, 1500), (1500, 200), (2000, 2100)):
        t = rng.integers(0, 4, size=fl).astype(np.int8)
        r = np.resize(t, rl).copy(); r[rng.random(rl) < 0.05] = 3
        reads.append(r); refs.append(t)
    reads.append(np.full(50, 4, np.int8)); refs.append(rng.integers(0, 4, size=80).astype(np.int8))       # all-N read
    n4 = rng.integers(0, 4, size=200).astype(np.int8); n4[::7] = 4
    reads.append(n4); refs.append(np.concatenate([rng.integers(0, 4, size=30).astype(np.int8), np.where(n4 == 4, 1, n4), rng.integers(0, 4, size=30).astype(np.int8)]))
    ro = np.concatenate([[0], np.cumsum([len(x) for x in reads])]).astype(np.int64)
    fo = np.concatenate([[0], np.cumsum([len(x) for x in refs])]).astype(np.int64)
    for flag in (0, 1, 0x0f):
        b = w.PairBatch(np.concatenate(reads), ro, np.concatenate(refs), fo, np.full(len(reads), 15, np.int32), mat=dna, flag=flag)
        g, gc, _, _ = gpu_table(eng, b)
        r, c = oracle_table(b)
        bad = diff(g, gc, r, c)
        assert len(bad) == 0, (flag, bad, g[bad[:1]], r[bad[:1]])


def test_legacy_per_pair_abi(eng):
    """ssw_init / ssw_align / align_destroy / init_destroy of realign/libssw.so, driven exactly like the reference's batch driver"""
    from oracle import oracle
    lib = ct.CDLL(pyssw.lib_path)
    b = w.fuzz_pairs(60, 41, flag=1, max_read=200, max_ref=250)
    g, gc, _ = oracle.run_batch(b.reads, b.read_off, b.refs, b.ref_off, b.masklen, b.mat, b.n, gapO=b.gapO, gapE=b.gapE, flag=1, threads=2, impl="lib", lib=lib, cigar_cap=CAP)
    r, c = oracle_table(b)
    assert len(diff(g, gc, r, c)) == 0


def test_pyssw_mirror(eng):
    """the Python operator interface of the reference (pyssw.SSW.align) and its new batched form agree with each other and with ssw.c"""
    rng = np.random.default_rng(9)
    ref = "".join("ACGT"[i] for i in rng.integers(0, 4, size=601))
    a = pyssw.SSW()
    a.set_reference_sequence(ref)
    queries = []
    for k in range(24):
        s = int(rng.integers(0, 200)); q = list(ref[s:s + 380])
        del q[100:100 + (k % 5)]
        q[200:200] = list("ACGTT"[:k % 4])
        if k % 6 == 0:
            q[50] = "N"
        queries.append("".join(q))
    queries.append("ACGTACGTAC")          # <= 30 bases: maskLen 15 branch of pyssw.py:142
    single = [a.align(q) for q in queries]
    batch = a.align_batch(queries)
    assert single == batch
    for (score, cigar, beg), q in zip(single, queries):
        assert score > 0 and beg >= 0 and cigar
    # cross-check one against the reference library when it is available
    from oracle import oracle
    r = pyssw.SSW(lib_path=oracle.require_ref())
    r.set_reference_sequence(ref)
    assert [r.align(q) for q in queries[:6]] == single[:6]


@pytest.mark.parametrize("flag,match,mismatch", [(0, 4, 6), (1, 4, 6), (1, 2, 3)])
def test_config4_ont_scale(eng, flag, match, mismatch):
    """BASELINE configs[3]: 10 kb reads at 8 % error vs 12 kb windows.  match=4 puts score1 on the int16 clamp of ssw.c:425
    (32-bit kernel with the clamp); match=2 stays below it (long un-saturated path).  flag 1 adds the reverse pass + wide-band traceback."""
    b = w.config4(12, seed=14 + flag + match, match=match, mismatch=mismatch, flag=flag)
    cap = 4096
    g, gc, rec, _ = gpu_table(eng, b, cap=cap)
    r, c = oracle_table(b, cap=cap, threads=os.cpu_count() or 8)
    bad = diff(g, gc, r, c)
    assert len(bad) == 0, (bad[:5], g[bad[:2]], r[bad[:2]])
    if match == 4:
        assert int(rec["score1"].max()) >= 32000
    if flag:
        assert int(rec["cigar_len"].min()) > 100


def test_config5_mixed_lengths_sample(eng):
    """BASELINE configs[4] on a seeded sample: read length log-uniform on [100, 20000], target = 1.2 x read, 5 % errors, flag 0 for the
    bulk and flag 1 on a subsample -- every length bin of the scheduler (packed strips, 32-bit kernel, clamp) in one batch."""
    b = w.config5(600, seed=15, flag=0)
    g, gc, _, _ = gpu_table(eng, b)
    r, c = oracle_table(b, threads=os.cpu_count() or 8)
    bad = diff(g, gc, r, c)
    assert len(bad) == 0, (bad[:5], g[bad[:2]], r[bad[:2]], b.read_len[bad[:5]])
    sub = b.subset(np.arange(0, 600, 5))
    sub.flag = 1
    cap = 8192
    g, gc, _, _ = gpu_table(eng, sub, cap=cap)
    r, c = oracle_table(sub, cap=cap, threads=os.cpu_count() or 8)
    bad = diff(g, gc, r, c)
    assert len(bad) == 0, (bad[:5], g[bad[:2]], r[bad[:2]], sub.read_len[bad[:5]])


def test_cigar_dense_pairs_exceed_the_typical_arena(eng):
    """reads that delete every third target base, cheap gaps: about one CIGAR word per read base, far beyond the typical-case
    arena budget (24 words per pair + 1 per 4 read bases) -> the engine must size for the worst case and re-run, not fail"""
    rng = np.random.default_rng(21)
    reads, refs = [], []
    for _ in range(96):
        t = rng.integers(0, 4, size=int(rng.integers(400, 700))).astype(np.int8)
        keep = np.ones(len(t), bool); keep[2::3] = False
        reads.append(t[keep][:300]); refs.append(t)
    ro = np.concatenate([[0], np.cumsum([len(x) for x in reads])]).astype(np.int64)
    fo = np.concatenate([[0], np.cumsum([len(x) for x in refs])]).astype(np.int64)
    mat = w.dna_matrix(5, 6)
    b = w.PairBatch(np.concatenate(reads), ro, np.concatenate(refs), fo, np.full(len(reads), 40, np.int32), mat=mat, gapO=2, gapE=1, flag=1)
    cap = 1024
    rec, cig = eng.align(b, cigar_cap=int(b.read_len.sum() + b.ref_len.sum()))
    g, gc = B.as_table(rec, cig, cap)
    r, c = oracle_table(b, cap=cap)
    assert len(diff(g, gc, r, c)) == 0
    assert int(rec["cigar_len"].sum()) > b.npairs * 24 + int(b.read_len.sum()) // 4 + 4096


def test_large_alphabet_goes_through_the_int32_kernels(eng):
    """the ABI is generic in n (ssw.h:77, any n x n matrix): a 20-letter protein-like matrix takes the 32-bit score kernel and the
    warp traceback (the packed kernels assume n <= 8)"""
    rng = np.random.default_rng(33)
    n = 20
    mat = rng.integers(-4, 3, size=(n, n)).astype(np.int8)
    mat = np.minimum(mat, mat.T)
    for i in range(n):
        mat[i, i] = int(rng.integers(4, 10))
    reads, refs = [], []
    for _ in range(150):
        t = rng.integers(0, n, size=int(rng.integers(30, 400))).astype(np.int8)
        s = int(rng.integers(0, max(1, len(t) - 20)))
        r = t[s:s + int(rng.integers(10, 250))].copy()
        r[rng.random(len(r)) < 0.1] = rng.integers(0, n)
        if len(r) > 40 and rng.random() < 0.5:
            r = np.delete(r, slice(20, 20 + int(rng.integers(1, 6))))
        reads.append(r); refs.append(t)
    ro = np.concatenate([[0], np.cumsum([len(x) for x in reads])]).astype(np.int64)
    fo = np.concatenate([[0], np.cumsum([len(x) for x in refs])]).astype(np.int64)
    for flag, gapO, gapE in ((1, 11, 1), (0, 5, 2)):
        b = w.PairBatch(np.concatenate(reads), ro, np.concatenate(refs), fo, np.maximum(15, np.diff(ro) // 2).astype(np.int32), mat=mat.reshape(-1), n=n,
                        gapO=gapO, gapE=gapE, flag=flag)
        g, gc, _, _ = gpu_table(eng, b)
        r, c = oracle_table(b)
        bad = diff(g, gc, r, c)
        assert len(bad) == 0, (flag, bad[:5], g[bad[:1]], r[bad[:1]])


@pytest.mark.parametrize("npairs", [6, 600])
def test_long_reads_with_n_bases(eng, npairs):
    """reads beyond the short-read strips that contain N (code 4): the packed multi-strip kernel flags them and the int32 kernel
    redoes them; with 6 pairs the multi-warp form of the long kernel is in use, with 600 the one-warp-per-pair form"""
    b = w.make_pairs(npairs, (1400, 3200) if npairs < 100 else (1300, 1700), 1.2, err=0.05, seed=77 + npairs, flag=1, chunk=64, n_frac=0.004)
    assert int((b.reads == 4).sum()) > 0
    cap = 2048
    g, gc, _, _ = gpu_table(eng, b, cap=cap)
    r, c = oracle_table(b, cap=cap, threads=os.cpu_count() or 8)
    bad = diff(g, gc, r, c)
    assert len(bad) == 0, (bad[:5], g[bad[:2]], r[bad[:2]])


@pytest.mark.parametrize("ncol", ["minus_mismatch", "zero", "plus_one", "varying"])
def test_short_reads_with_n_bases(eng, ncol):
    """reads of every short-read strip size that contain N (code 4): with a constant N column (both matrix builders of the reference,
    ssw_cpp.cpp:23-48 and pyssw.py:61-79) they stay in the packed kernel (second PRMT per row); with a varying N column they are flagged
    and redone by the int32 kernel.  Targets contain N too."""
    parts = [w.make_pairs(300, (lo, hi), 1.6, err=0.04, seed=900 + lo, flag=1, chunk=128, n_frac=0.006)
             for lo, hi in ((20, 64), (65, 330), (331, 700), (701, 1280))]
    reads = np.concatenate([p.reads for p in parts]); refs = np.concatenate([p.refs for p in parts])
    rl = np.concatenate([p.read_len for p in parts]); fl = np.concatenate([p.ref_len for p in parts])
    ro = np.zeros(len(rl) + 1, dtype=np.int64); np.cumsum(rl, out=ro[1:])
    fo = np.zeros(len(fl) + 1, dtype=np.int64); np.cumsum(fl, out=fo[1:])
    assert int((reads == 4).sum()) > 500 and int((refs == 4).sum()) > 500
    mat = w.dna_matrix(2, 3, n_zero=(ncol == "zero")).reshape(5, 5).copy()
    if ncol == "plus_one":
        mat[:, 4] = 1; mat[4, :] = 1
    elif ncol == "varying":
        mat[:, 4] = [0, -1, -2, -3, 1]; mat[4, :] = mat[:, 4]
    for flag in (1, 0):
        b = w.PairBatch(reads, ro, refs, fo, np.maximum(rl // 2, 15).astype(np.int32), mat=mat.reshape(-1), flag=flag)
        cap = 2048
        g, gc, _, _ = gpu_table(eng, b, cap=cap)
        r, c = oracle_table(b, cap=cap, threads=os.cpu_count() or 8)
        bad = diff(g, gc, r, c)
        assert len(bad) == 0, (ncol, flag, bad[:5], g[bad[:2]], r[bad[:2]])


def test_targets_longer_than_65535_columns(eng):
    """short reads against targets of 66 k - 90 k bases: column indices no longer fit the 16-bit keys of the second-best reduction
    (sw_finish.cuh takes its 64-bit path), the score kernel runs 70 k+ wavefront steps per pair; a second copy of the read far from the
    first gives a real score2 / ref_end2"""
    rng = np.random.default_rng(4242)
    reads, refs, masks = [], [], []
    for k in range(12):
        fl = int(rng.integers(66_000, 90_000))
        t = rng.integers(0, 4, size=fl).astype(np.int8)
        rl = int(rng.integers(60, 400))
        s1 = int(rng.integers(0, fl // 2 - rl)); s2 = int(rng.integers(fl // 2, fl - rl))
        r = t[s1:s1 + rl].copy()
        t[s2:s2 + rl] = r                                   # exact second copy
        r[rng.random(rl) < 0.03] = rng.integers(0, 4)
        if k % 3 == 0:
            r = np.delete(r, slice(rl // 2, rl // 2 + 3))
        reads.append(r); refs.append(t); masks.append(max(15, len(r) // 2))
    ro = np.concatenate([[0], np.cumsum([len(x) for x in reads])]).astype(np.int64)
    fo = np.concatenate([[0], np.cumsum([len(x) for x in refs])]).astype(np.int64)
    for flag in (1, 0):
        b = w.PairBatch(np.concatenate(reads), ro, np.concatenate(refs), fo, np.array(masks, dtype=np.int32), flag=flag)
        g, gc, _, _ = gpu_table(eng, b)
        r, c = oracle_table(b, threads=os.cpu_count() or 8)
        bad = diff(g, gc, r, c)
        assert len(bad) == 0, (flag, bad[:5], g[bad[:2]], r[bad[:2]])
        assert int((g[:, 1] > 100).sum()) >= 8              # score2 comes from the second copy


def test_nibble_packed_input_equals_int8_input(eng):
    """mpn_align_batch_packed4 (half the host->device bytes, expanded on the device) == mpn_align_batch, on odd lengths / odd offsets and on
    a batch large enough to be cut into pipeline ranges"""
    for b in (w.fuzz_pairs(400, 77, flag=1, random_matrix=False), w.config2(3000, seed=5), w.make_pairs(300_000, (21, 59), 83, err=0.03, seed=9, flag=1)):
        rec, cig = eng.align(b)
        want, wantc = B.as_table(rec, cig, 64)
        prec, pcig = eng.align_packed4(b, B.pack4(b.reads), B.pack4(b.refs))
        got, gotc = B.as_table(prec, pcig, 64)
        assert (got == want).all() and (gotc == wantc).all(), b.name


def test_two_bit_packed_input_equals_int8_input(eng):
    """mpn_align_batch_packed2 (a quarter of the host->device bytes; N codes travel as an exception list) == mpn_align_batch, on odd lengths /
    odd offsets, with and without N, and on a batch large enough to be cut into pipeline ranges"""
    for b in (w.fuzz_pairs(400, 78, flag=1, random_matrix=False), w.config2(3000, seed=6), w.make_pairs(5000, (20, 90), 130, err=0.04, seed=10, flag=1, n_frac=0.02),
              w.make_pairs(300_000, (21, 59), 83, err=0.03, seed=11, flag=1, n_frac=0.001)):
        rec, cig = eng.align(b)
        want, wantc = B.as_table(rec, cig, 64)
        r2, rx = B.pack2(b.reads)
        f2, fx = B.pack2(b.refs)
        prec, pcig = eng.align_packed2(b, r2, rx, f2, fx)
        got, gotc = B.as_table(prec, pcig, 64)
        assert (got == want).all() and (gotc == wantc).all(), b.name


def test_large_spans_batch_goes_through_the_chunk_pipeline(eng):
    """mpn_align_batch_spans with more pairs than one pipeline range: arena uploaded once, pairs cut into ranges -- same records as the CSR call"""
    b = w.make_pairs(320_000, (21, 59), 83, err=0.03, seed=12, flag=1)
    rec, cig = eng.align(b)
    want, wantc = B.as_table(rec, cig, 64)
    arena = np.concatenate([b.reads, b.refs])
    rd_start = b.read_off[:-1]; rf_start = len(b.reads) + b.ref_off[:-1]
    srec, scig = eng.align_spans(b, arena, rd_start, b.read_len.astype(np.int32), rf_start, b.ref_len.astype(np.int32), b.masklen)
    got, gotc = B.as_table(srec, scig, 64)
    assert (got == want).all() and (gotc == wantc).all()
