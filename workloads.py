"""Seeded synthetic inputs for the five BASELINE.json configs.  Test / bench infrastructure: not part of the product package.

Every generator returns a `PairBatch`: CSR-packed int8 code arrays (A0 C1 G2 T3 N4, as ssw_cpp.cpp:8-21 /
pyssw.py:86-98 translate them) plus per-pair maskLen and the scoring that the reference callers use
(ssw_cpp.cpp:23-48 builds +match/-mismatch with N = -mismatch; pyssw.py:61-79 builds N = 0).
Reads are sampled from their target window with substitutions / insertions / deletions in equal parts.
Generation is numpy-only and vectorised, so the 1 M-pair config builds in seconds.
"""
from dataclasses import dataclass, field
import numpy as np


def dna_matrix(match=4, mismatch=6, n_zero=False):
    """5x5 matrix as ssw_cpp.cpp:23-48 (N row/col = -mismatch) or pyssw.py:61-79 (N row/col = 0)."""
    m = np.full((5, 5), -mismatch, dtype=np.int8)
    for i in range(4):
        m[i, i] = match
    if n_zero:
        m[4, :] = 0
        m[:, 4] = 0
    return m.reshape(-1)


@dataclass
class PairBatch:
    reads: np.ndarray       # int8 codes, concatenated
    read_off: np.ndarray    # int64, n+1
    refs: np.ndarray        # int8 codes, concatenated
    ref_off: np.ndarray     # int64, n+1
    masklen: np.ndarray     # int32, n
    mat: np.ndarray = field(default_factory=dna_matrix)
    n: int = 5
    gapO: int = 8
    gapE: int = 2
    flag: int = 0
    filters: int = 0
    filterd: int = 32767
    score_size: int = 2
    name: str = ""

    @property
    def npairs(self):
        return len(self.read_off) - 1

    @property
    def read_len(self):
        return np.diff(self.read_off)

    @property
    def ref_len(self):
        return np.diff(self.ref_off)

    @property
    def cells(self):
        """forward-matrix cells, the GCUPS numerator."""
        return int((self.read_len.astype(np.int64) * self.ref_len.astype(np.int64)).sum())

    def subset(self, idx):
        idx = np.asarray(idx, dtype=np.int64)
        rl, fl = self.read_len[idx], self.ref_len[idx]
        ro = np.zeros(len(idx) + 1, dtype=np.int64); np.cumsum(rl, out=ro[1:])
        fo = np.zeros(len(idx) + 1, dtype=np.int64); np.cumsum(fl, out=fo[1:])
        reads = np.empty(ro[-1], dtype=np.int8); refs = np.empty(fo[-1], dtype=np.int8)
        for k, i in enumerate(idx):
            reads[ro[k]:ro[k + 1]] = self.reads[self.read_off[i]:self.read_off[i + 1]]
            refs[fo[k]:fo[k + 1]] = self.refs[self.ref_off[i]:self.ref_off[i + 1]]
        return PairBatch(reads, ro, refs, fo, self.masklen[idx].copy(), self.mat, self.n, self.gapO, self.gapE, self.flag,
                         self.filters, self.filterd, self.score_size, self.name + f"[subset {len(idx)}]")

    def shard(self, rank, world):
        """Contiguous shard of a length-sorted batch dealt round-robin: pair i goes to rank i % world."""
        if world == 1:
            return self
        return self.subset(np.arange(rank, self.npairs, world))


def _mutated_reads(rng, targets, tlen, rlen, offset, err):
    """targets: [n, tlen] int8; returns list-free CSR reads of per-pair length rlen[i], sampled from targets[i, offset[i]:]
    with error rate err split 1/3 substitution, 1/3 insertion, 1/3 deletion."""
    n = targets.shape[0]
    lmax = int(rlen.max())
    u = rng.random((n, lmax), dtype=np.float32)
    is_sub = u < err / 3
    is_ins = (u >= err / 3) & (u < 2 * err / 3)
    is_del = (u >= 2 * err / 3) & (u < err)
    step = np.ones((n, lmax), dtype=np.int32)
    step[is_ins] = 0
    step[is_del] = 2
    src = offset[:, None].astype(np.int64) + np.cumsum(step, axis=1) - step + (is_del.astype(np.int64))
    np.clip(src, 0, tlen - 1, out=src)
    base = np.take_along_axis(targets, src, axis=1)
    rnd = rng.integers(0, 4, size=(n, lmax), dtype=np.int8)
    base = np.where(is_ins, rnd, base)
    base = np.where(is_sub, (base + 1 + (rnd % 3)) % 4, base).astype(np.int8)
    keep = np.arange(lmax)[None, :] < rlen[:, None]
    reads = base[keep]
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(rlen, out=off[1:])
    return reads, off


def make_pairs(npairs, read_len, ref_len, err=0.02, seed=1, flag=0, mask="half", name="", chunk=65536, n_frac=0.0, **kw):
    """Generic generator.  read_len: int or (lo, hi) inclusive uniform; ref_len: int, or float multiple of the read length."""
    rng = np.random.default_rng(seed)
    reads_l, refs_l, rlens, flens = [], [], [], []
    for c0 in range(0, npairs, chunk):
        n = min(chunk, npairs - c0)
        rl = (np.full(n, read_len, dtype=np.int64) if np.isscalar(read_len)
              else rng.integers(read_len[0], read_len[1] + 1, size=n, dtype=np.int64))
        if isinstance(ref_len, float):
            fl = np.maximum((rl * ref_len).astype(np.int64), rl + 8)
        else:
            fl = np.full(n, ref_len, dtype=np.int64)
        tmax = int(fl.max())
        targets = rng.integers(0, 4, size=(n, tmax), dtype=np.int8)
        if n_frac > 0:
            targets[rng.random((n, tmax)) < n_frac] = 4
        slack = np.maximum(fl - rl - (rl * err).astype(np.int64) - 4, 0)
        offset = (rng.random(n) * (slack + 1)).astype(np.int64)
        r, ro = _mutated_reads(rng, targets, tmax, rl, offset, err)
        keep = np.arange(tmax)[None, :] < fl[:, None]
        refs_l.append(targets[keep]); reads_l.append(r); rlens.append(rl); flens.append(fl)
    rl = np.concatenate(rlens); fl = np.concatenate(flens)
    ro = np.zeros(npairs + 1, dtype=np.int64); np.cumsum(rl, out=ro[1:])
    fo = np.zeros(npairs + 1, dtype=np.int64); np.cumsum(fl, out=fo[1:])
    if mask == "half":
        ml = np.maximum(rl // 2, 15).astype(np.int32)
    elif mask == "qlen":
        ml = rl.astype(np.int32)
    else:
        ml = np.full(npairs, int(mask), dtype=np.int32)
    return PairBatch(np.concatenate(reads_l), ro, np.concatenate(refs_l), fo, ml, flag=flag, name=name, **kw)


def config1(npairs=10_000, seed=11):
    """BASELINE configs[0]: score + end positions only, 250 bp reads vs 500 bp targets, flag 0, maskLen 125."""
    return make_pairs(npairs, 250, 500, err=0.02, seed=seed, flag=0, mask=125, name="config1: 250bp x 500bp flag0")


def config2(npairs=1_000_000, seed=12):
    """BASELINE configs[1]: full ssw_align with traceback + CIGAR, reads U{150..300} vs 1 kb haplotypes, flag 1, maskLen readLen/2."""
    return make_pairs(npairs, (150, 300), 1000, err=0.02, seed=seed, flag=1, mask="half", name="config2: 150-300bp x 1kb flag1 (CIGAR)")


def config4(npairs=1024, seed=14, match=4, mismatch=6, flag=0):
    """BASELINE configs[3]: ONT-scale, 10 kb reads at 8 % error vs 12 kb windows (scores land on the 32767 clamp with match=4)."""
    b = make_pairs(npairs, 10_000, 12_000, err=0.08, seed=seed, flag=flag, mask="half", name=f"config4: 10kb x 12kb flag{flag}", chunk=64)
    b.mat = dna_matrix(match, mismatch)
    return b


def config5(npairs=10_000_000, seed=15, lo=100, hi=20_000, flag=0):
    """BASELINE configs[4]: mixed lengths, read length log-uniform on [lo, hi], target = 1.2 x read, 5 % errors.
    (The distribution is not fixed by BASELINE.json; log-uniform is this repo's stated choice)
    Built bin by bin so the padded generator never allocates npairs x hi."""
    rng = np.random.default_rng(seed)
    rl = np.exp(rng.uniform(np.log(lo), np.log(hi), size=npairs)).astype(np.int64)
    rl.sort()
    parts = []
    edges = np.unique(np.concatenate([[0], np.searchsorted(rl, np.geomspace(lo, hi, 40)[1:-1]), [npairs]]))
    for a, b in zip(edges[:-1], edges[1:]):
        if b <= a:
            continue
        parts.append((rl[a:b], seed * 1000 + int(a)))
    reads_l, refs_l, rls, fls = [], [], [], []
    for lens, s in parts:
        sub = make_pairs(len(lens), (int(lens.min()), int(lens.max())), 1.2, err=0.05, seed=s, flag=flag,
                         chunk=max(1, min(65536, 200_000_000 // int(lens.max() * 1.3))))
        reads_l.append(sub.reads); refs_l.append(sub.refs); rls.append(sub.read_len); fls.append(sub.ref_len)
    rl2 = np.concatenate(rls); fl2 = np.concatenate(fls)
    ro = np.zeros(len(rl2) + 1, dtype=np.int64); np.cumsum(rl2, out=ro[1:])
    fo = np.zeros(len(fl2) + 1, dtype=np.int64); np.cumsum(fl2, out=fo[1:])
    ml = np.maximum(rl2 // 2, 15).astype(np.int32)
    return PairBatch(np.concatenate(reads_l), ro, np.concatenate(refs_l), fo, ml, flag=flag, name=f"config5: mixed {lo}bp-{hi}bp x{len(rl2)}")


_seqgen = None


def seqgen():
    """tools/libseqgen.so (tools/seqgen.c): pairs as a pure function of (seed, global pair index), generated at memory speed on host threads"""
    global _seqgen
    if _seqgen is None:
        import ctypes as ct, os, subprocess
        here = os.path.dirname(os.path.abspath(__file__))
        so, src = os.path.join(here, "tools", "libseqgen.so"), os.path.join(here, "tools", "seqgen.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-pthread", "-o", so, src], check=True)
        L = ct.CDLL(so)
        L.seqgen_pairs.argtypes = [ct.c_uint64, ct.c_int64, ct.c_int64, ct.c_void_p, ct.c_void_p, ct.c_double, ct.c_void_p, ct.c_void_p, ct.c_int]
        L.seqgen_pairs.restype = None
        _seqgen = L
    return _seqgen


def gen_pairs_fast(rl, fl, seed, first_index=0, err=0.05, threads=4, flag=0, name="", reads_out=None, refs_out=None, **kw):
    """PairBatch for the given per-pair lengths through tools/seqgen.c.  reads_out / refs_out: optional preallocated int8 buffers
    (e.g. views of pinned torch tensors) that are filled in place."""
    rl = np.ascontiguousarray(rl, dtype=np.int64); fl = np.ascontiguousarray(fl, dtype=np.int64)
    n = len(rl)
    ro = np.zeros(n + 1, dtype=np.int64); np.cumsum(rl, out=ro[1:])
    fo = np.zeros(n + 1, dtype=np.int64); np.cumsum(fl, out=fo[1:])
    reads = np.empty(int(ro[-1]), dtype=np.int8) if reads_out is None else reads_out[:int(ro[-1])]
    refs = np.empty(int(fo[-1]), dtype=np.int8) if refs_out is None else refs_out[:int(fo[-1])]
    seqgen().seqgen_pairs(int(seed), int(first_index), n, ro.ctypes.data, fo.ctypes.data, float(err), reads.ctypes.data, refs.ctypes.data, int(threads))
    ml = np.maximum(rl // 2, 15).astype(np.int32)
    return PairBatch(reads, ro, refs, fo, ml, flag=flag, name=name, **kw)


class MixedStream:
    """BASELINE configs[4] -- 10 M mixed-length pairs, read length log-uniform on [lo, hi], target = 1.2 x read, 5 % errors -- as a stream
    of LENGTH-SORTED CHUNKS that never sits in RAM at once (~80 GB of bases at full size).  The pairs are sorted by read length once
    (the lengths alone are 80 MB); chunk k is a contiguous run of that order holding about `chunk_cost` cost units (forward cells x the
    relative cost of the score kernel the pairs take) and at most `max_pairs` pairs; its bases are a pure function of (seed, global
    index).  plan(world) deals the chunks to ranks, heaviest first, each to the least loaded rank (length-sorted
    chunks dealt by cells, no collective)."""

    def __init__(self, npairs=10_000_000, seed=15, lo=100, hi=20_000, flag=0, err=0.05, chunk_cost=None, max_pairs=400_000, match=4, world=1):
        rng = np.random.default_rng(seed)
        rl = np.exp(rng.uniform(np.log(lo), np.log(hi), size=npairs)).astype(np.int64)
        rl.sort()
        self.rl = rl
        self.fl = np.maximum((rl * 1.2).astype(np.int64), rl + 8)
        self.npairs, self.seed, self.flag, self.err = npairs, seed, flag, err
        cells = self.rl * self.fl
        # the engine's classifier (engine.cu pair_cost_per_cell): packed short-read kernel up to 1280 rows while no H can reach the int16 clamp
        short = ((np.minimum(self.rl, self.fl) + 1) * match <= 32767) & (self.rl <= 1280)
        cost = cells * np.where(short, 1.0, 1.55) + 4096.0
        self.total_cells = int(cells.sum())
        if chunk_cost is None:
            # a chunk must fill a GPU (thousands of long pairs: the multi-strip kernel runs one warp per pair) -> 2.5e12 cost units (~0.7 s);
            # smaller workloads still give every rank about a dozen chunks to balance on
            chunk_cost = min(2.5e12, max(2.0e11, float(cost.sum()) / (12.0 * world)))
        cum = np.cumsum(cost)
        bounds = [0]
        while bounds[-1] < npairs:
            a = bounds[-1]
            base = cum[a - 1] if a else 0.0
            b = int(np.searchsorted(cum, base + chunk_cost, side="left")) + 1
            bounds.append(min(npairs, max(a + 1, min(b, a + max_pairs))))
        self.bounds = np.array(bounds, dtype=np.int64)
        self.nchunks = len(bounds) - 1
        self.chunk_cost = np.add.reduceat(cost, self.bounds[:-1])
        self.chunk_cells = np.add.reduceat(cells, self.bounds[:-1])

    def plan(self, world):
        """chunk ids per rank: heaviest chunk first, each to the least loaded rank (deterministic: every rank derives the same plan)"""
        order = np.argsort(-self.chunk_cost, kind="stable")
        load = np.zeros(world)
        out = [[] for _ in range(world)]
        for k in order:
            r = int(np.argmin(load))
            out[r].append(int(k)); load[r] += self.chunk_cost[k]
        return out

    def chunk(self, k, threads=4, reads_out=None, refs_out=None):
        a, b = int(self.bounds[k]), int(self.bounds[k + 1])
        return gen_pairs_fast(self.rl[a:b], self.fl[a:b], self.seed, first_index=a, err=self.err, threads=threads, flag=self.flag,
                              name=f"config5 chunk {k}: pairs {a}..{b}", reads_out=reads_out, refs_out=refs_out)

    def chunk_bytes(self, k):
        a, b = int(self.bounds[k]), int(self.bounds[k + 1])
        return int(self.rl[a:b].sum()), int(self.fl[a:b].sum())


def fuzz_pairs(npairs, seed, max_read=700, max_ref=500, alphabet=4, flag=1, random_matrix=True, with_n=True):
    """The adversarial distribution of SURVEY.md section 8c: random 5x5 matrices, gapE 1..3, gapO > gapE, low-complexity
    alphabets (ties), reads derived from the target with jumps / random blocks / substitutions, occasional N, lengths from 1,
    random maskLen >= 15.  One scoring per batch (the batched ABI takes one matrix per batch)."""
    rng = np.random.default_rng(seed)
    if random_matrix:
        match = int(rng.integers(1, 6)); mism = int(rng.integers(1, 7))
        mat = np.full((5, 5), -mism, dtype=np.int8)
        for i in range(4):
            mat[i, i] = match
        if rng.random() < 0.5:
            mat[4, :] = 0; mat[:, 4] = 0
        elif rng.random() < 0.5:
            x = -int(rng.integers(0, 4)); mat[4, :] = x; mat[:, 4] = x
        if rng.random() < 0.3:     # fully random symmetric-free matrix, still with a positive diagonal
            mat = rng.integers(-6, 3, size=(5, 5)).astype(np.int8)
            for i in range(4):
                mat[i, i] = int(rng.integers(1, 6))
        gapE = int(rng.integers(1, 4)); gapO = gapE + int(rng.integers(1, 9))
    else:
        mat = dna_matrix().reshape(5, 5); gapO, gapE = 8, 2
    reads_l, refs_l, rl, fl = [], [], [], []
    for _ in range(npairs):
        flen = int(rng.integers(1, max_ref + 1)) if rng.random() < 0.9 else int(rng.integers(1, 20))
        ref = rng.integers(0, alphabet, size=flen, dtype=np.int8)
        if rng.random() < 0.15:
            rlen = int(rng.integers(1, 17))
        else:
            rlen = int(rng.integers(1, max_read + 1))
        out = []
        pos = int(rng.integers(0, flen))
        while len(out) < rlen:
            u = rng.random()
            if u < 0.03:
                pos += int(rng.integers(-30, 31))
            elif u < 0.06:
                out.extend(rng.integers(0, alphabet, size=int(rng.integers(1, 21))).tolist())
                continue
            pos = min(max(pos, 0), flen - 1)
            b = int(ref[pos])
            if rng.random() < 0.06:
                b = int(rng.integers(0, alphabet))
            if with_n and rng.random() < 0.004:
                b = 4
            out.append(b)
            pos += 1
            if pos >= flen:
                pos = int(rng.integers(0, flen))
        read = np.array(out[:rlen], dtype=np.int8)
        if with_n and rng.random() < 0.1 and flen > 3:
            ref = ref.copy(); ref[rng.integers(0, flen, size=2)] = 4
        reads_l.append(read); refs_l.append(ref); rl.append(rlen); fl.append(flen)
    rl = np.array(rl, dtype=np.int64); fl = np.array(fl, dtype=np.int64)
    ro = np.zeros(npairs + 1, dtype=np.int64); np.cumsum(rl, out=ro[1:])
    fo = np.zeros(npairs + 1, dtype=np.int64); np.cumsum(fl, out=fo[1:])
    ml = np.maximum(15, rng.integers(15, 40, size=npairs) if rng.random() < 0.3 else (rl // 2 + rng.integers(0, 6, size=npairs))).astype(np.int32)
    return PairBatch(np.concatenate(reads_l), ro, np.concatenate(refs_l), fo, ml, mat=mat.reshape(-1).astype(np.int8), gapO=gapO, gapE=gapE,
                     flag=flag, name=f"fuzz seed {seed}")


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE configs[2]: the realigner's amplicon workload (SURVEY.md section 8d "Config 3").  One region = what
# realign_illumina_reads.py:567-605 hands to realign_reads: a reference window with prefix/suffix flanks, the candidate
# haplotypes (flank + consensus + flank, space separated) and the reads overlapping the window with their current
# position and CIGAR.  The de Bruijn assembly that proposes haplotypes is out of scope, so haplotypes are the window with
# 1-3 planted SNVs / small indels; reads are sampled from a few "true" haplotypes with 0.5 % error.
import importlib as _importlib
RegionWorkload = _importlib.import_module("megapath-nano_b200.realigner").Region      # the product's input record of realign_reads / realign_regions


def _rand_dna(rng, n):
    return "".join("ACGT"[i] for i in rng.integers(0, 4, size=n))


def _plant_variants(rng, center, k):
    s = list(center)
    for _ in range(k):
        if len(s) < 40:
            break
        p = int(rng.integers(10, len(s) - 10))
        u = rng.random()
        if u < 0.5:
            s[p] = "ACGT"[("ACGT".index(s[p]) + 1 + int(rng.integers(0, 3))) % 4]
        elif u < 0.75:
            s[p:p] = list(_rand_dna(rng, int(rng.integers(1, 11))))
        else:
            del s[p:p + int(rng.integers(1, 11))]
    return "".join(s)


def _noisy_copy(rng, s, err):
    out = []
    for ch in s:
        u = rng.random()
        if u < err / 3:
            out.append("ACGT"[("ACGT".index(ch) + 1 + int(rng.integers(0, 3))) % 4])
        elif u < 2 * err / 3:
            out.append("ACGT"[int(rng.integers(0, 4))]); out.append(ch)
        elif u < err:
            continue
        else:
            out.append(ch)
    return "".join(out)


def config3(nregions=64, seed=13, max_reads=1000, max_haps=32, err=0.005, n_frac=0.0):
    """list of RegionWorkload (seeded)."""
    rng = np.random.default_rng(seed)
    regions = []
    for _ in range(nregions):
        wlen = int(rng.integers(160, 1001)); pre = int(rng.integers(20, 271)); suf = int(rng.integers(20, 271))
        prefix, center, suffix = _rand_dna(rng, pre), _rand_dna(rng, wlen), _rand_dna(rng, suf)
        reference = prefix + center + suffix
        nh = int(rng.integers(2, max_haps + 1))
        haps = [reference]
        while len(haps) < nh:
            h = prefix + _plant_variants(rng, center, int(rng.integers(1, 4))) + suffix
            if h not in haps:
                haps.append(h)
        order = rng.permutation(nh)
        haps = [haps[i] for i in order]
        truth = [haps[i] for i in rng.choice(nh, size=min(nh, int(rng.integers(1, 4))), replace=False)]
        nr = int(rng.integers(50, max_reads + 1))
        reads, positions, cigars = [], [], []
        ref_start = int(rng.integers(1000, 5_000_000))
        for _ in range(nr):
            h = truth[int(rng.integers(0, len(truth)))]
            rl = int(rng.integers(100, 251)); rl = min(rl, len(h) - 1)
            st = int(rng.integers(0, len(h) - rl + 1))
            r = _noisy_copy(rng, h[st:st + rl], err)
            if n_frac > 0:
                r = "".join("N" if rng.random() < n_frac else c for c in r)
            if not r:
                r = "A"
            reads.append(r); positions.append(ref_start + st); cigars.append(f"{len(r)}M")
        regions.append(RegionWorkload(reference, haps, reads, positions, cigars, ref_start, pre, suf))
    return regions


def dbg_windows(nwindows=32, seed=21, max_reads=400, err=0.005, repeat_frac=0.2, n_frac=0.0005, lowq_frac=0.01):
    """Assembler inputs shaped like realign_illumina_reads.py:532-553: a 160-1000 bp reference window, reads
    of 100-250 bp sampled from 1-3 truth haplotypes (reference + planted SNVs / indels) that overlap the window (so they may start before
    it or run past it), 0.5 % sequencing error, a few N and a few low-quality positions per read.  A fraction of the windows carries a
    tandem repeat, which pushes the smallest acyclic k up.  Returns a list of (ref, reads, low_quality_fields)."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(nwindows):
        wlen = int(rng.integers(160, 1001))
        ref = _rand_dna(rng, wlen)
        if rng.random() < repeat_frac:
            unit = _rand_dna(rng, int(rng.integers(2, 30)))
            at = int(rng.integers(20, wlen - 60))
            rep = (unit * 40)[:int(rng.integers(20, 60))]
            ref = ref[:at] + rep + ref[at + len(rep):]
        flank_l, flank_r = _rand_dna(rng, 120), _rand_dna(rng, 120)
        truth = [ref] + [_plant_variants(rng, ref, int(rng.integers(1, 4))) for _ in range(int(rng.integers(0, 3)))]
        reads, lowq = [], []
        for _ in range(int(rng.integers(0 if rng.random() < 0.05 else 20, max_reads + 1))):
            h = flank_l + truth[int(rng.integers(0, len(truth)))] + flank_r
            rl = int(rng.integers(100, 251))
            st = int(rng.integers(0, max(1, len(h) - rl)))
            r = _noisy_copy(rng, h[st:st + rl], err)
            if n_frac > 0:
                r = "".join("N" if rng.random() < n_frac else c for c in r)
            reads.append(r)
            lowq.append(" ".join(str(i) for i in np.nonzero(rng.random(len(r)) < lowq_frac)[0]))
        out.append((ref, reads, lowq))
    return out


def dbg_cap_windows(seed=41):
    """Assembler inputs around the 256-path cap of CandidatePaths (debruijn_graph.cpp:293-299): isolated multi-allelic sites, every allele
    combination read twice, so the number of source-to-sink paths is the product of the allele counts -- 243, 256 (two shapes), 288, 324,
    512 -- plus one window whose LAST branching sits right before the sink (the case in which the visiting order of successors decides
    whether the cap trips).  Returns a list of (ref, reads, low_quality_fields, expected number of paths)."""
    rng = np.random.default_rng(seed)
    out = []
    for alleles in ([3] * 5, [2] * 8, [4] * 4, [2] * 5 + [3] * 2, [3] * 4 + [4], [2] * 9, [2] * 7 + [3]):
        gap = 40
        ref = _rand_dna(rng, gap * (len(alleles) + 1))
        sites = [gap * (k + 1) for k in range(len(alleles))]
        if alleles == [2] * 7 + [3]:
            sites[-1] = len(ref) - 14                      # last branching close to the sink (k is 10 or a little more)
        reads = []
        # every read covers two neighbouring sites (and nothing else varies), all allele pairs, twice -> every edge has weight >= 2
        for k in range(len(sites)):
            lo = max(0, sites[k] - 30); hi = min(len(ref), (sites[k + 1] if k + 1 < len(sites) else sites[k]) + 30)
            for a in range(alleles[k]):
                for b in range(alleles[k + 1] if k + 1 < len(sites) else 1):
                    r = list(ref)
                    r[sites[k]] = "ACGT"[("ACGT".index(ref[sites[k]]) + a) % 4]
                    if k + 1 < len(sites):
                        r[sites[k + 1]] = "ACGT"[("ACGT".index(ref[sites[k + 1]]) + b) % 4]
                    reads += ["".join(r[lo:hi])] * 2
        n = 1
        for x in alleles:
            n *= x
        out.append((ref, reads, [""] * len(reads), n))
    return out


@dataclass
class WindowWorkload:
    """one candidate window before assembly (realign_illumina_reads.py:532-580): `chrom` is a piece of reference sequence whose first
    base sits at chromosome position chrom_start; the window is chrom[win_start:win_end]; reads carry chromosome positions."""
    chrom: str
    chrom_start: int
    win_start: int
    win_end: int
    reads: list
    positions: list
    cigars: list
    low_quality: list


def config3_windows(nwindows=32, seed=51, max_reads=400, err=0.005, lowq_frac=0.005):
    """BASELINE configs[2] one step earlier: windows whose haplotypes are NOT given but have to be assembled from the reads."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(nwindows):
        wlen = int(rng.integers(160, 1001)); pad = 300
        chrom = _rand_dna(rng, pad + wlen + pad)
        center = chrom[pad:pad + wlen]
        truth = [center] + [_plant_variants(rng, center, int(rng.integers(1, 4))) for _ in range(int(rng.integers(1, 3)))]
        chrom_start = int(rng.integers(1000, 5_000_000))
        reads, positions, cigars, lowq = [], [], [], []
        for _ in range(int(rng.integers(50, max_reads + 1))):
            h = chrom[:pad] + truth[int(rng.integers(0, len(truth)))] + chrom[pad + wlen:]
            rl = int(rng.integers(100, 251))
            st = int(rng.integers(pad - 80, max(pad - 79, pad + wlen - 20)))          # overlaps the window, overhang up to ~250 bp
            r = _noisy_copy(rng, h[st:st + rl], err) or "A"
            reads.append(r); positions.append(chrom_start + st); cigars.append(f"{len(r)}M")
            lowq.append(" ".join(str(i) for i in np.nonzero(rng.random(len(r)) < lowq_frac)[0]))
        out.append(WindowWorkload(chrom, chrom_start, pad, pad + wlen, reads, positions, cigars, lowq))
    return out


def fastpass_adversarial(nregions=12, seed=81):
    """Regions that stress the fast pass's order-dependent rules (realigner.cpp:170-253): tandem repeats and duplicated blocks (equal-score
    placements at several starts: the first evaluated wins), reads hanging over either end of the haplotype (start clamp at 0, skipped
    overhang), reads of 1..40 bases (not indexed up to 32), haplotypes shorter than k, N in reads and haplotypes, haplotypes that lose
    coverage inside the window (dropped), window flanks from 0 to longer than the haplotype."""
    rng = np.random.default_rng(seed)
    regions = []
    for g in range(nregions):
        wlen = int(rng.integers(40, 500))
        unit = _rand_dna(rng, int(rng.integers(1, 45)))
        core = _rand_dna(rng, wlen)
        if g % 3 != 2:
            at = int(rng.integers(0, max(1, wlen - 20)))
            rep = (unit * 200)[:int(rng.integers(40, 200))]
            core = core[:at] + rep + core[at:]
            if g % 3 == 1:
                core = core + _rand_dna(rng, 30) + rep[:90] + _rand_dna(rng, 40)
        pre, suf = int(rng.integers(0, 60)), int(rng.integers(0, 60))
        reference = _rand_dna(rng, pre) + core + _rand_dna(rng, suf)
        haps = [reference]
        for _ in range(int(rng.integers(1, 7))):
            h = reference[:pre] + _plant_variants(rng, core, int(rng.integers(1, 4))) + reference[len(reference) - suf:]
            if rng.random() < 0.2:
                h = "".join("N" if rng.random() < 0.01 else c for c in h)
            haps.append(h)
        if rng.random() < 0.3:
            haps.append(_rand_dna(rng, int(rng.integers(1, 40))))
        order = rng.permutation(len(haps)); haps = [haps[i] for i in order]
        reads, positions, cigars = [], [], []
        for _ in range(int(rng.integers(5, 120))):
            h = haps[int(rng.integers(0, len(haps)))]
            u = rng.random()
            rl = int(rng.integers(1, 41)) if u < 0.1 else int(rng.integers(33, 257))
            ext = _rand_dna(rng, 60) + h + _rand_dna(rng, 60)                           # lets reads hang over both ends
            st = int(rng.integers(0, max(1, len(ext) - rl)))
            r = ext[st:st + rl]
            nm = int(rng.integers(0, 5)) if rng.random() < 0.5 else 0
            r = list(r)
            for _ in range(nm):
                p = int(rng.integers(0, len(r))); r[p] = "ACGTN"[int(rng.integers(0, 5))]
            r = "".join(r) or "A"
            reads.append(r); positions.append(1000 + st); cigars.append(f"{len(r)}M")
        if g % 4 == 0:
            pre, suf = 0, 0
        elif g % 4 == 1:
            suf = len(reference) + 5
        regions.append(RegionWorkload(reference, haps, reads, positions, cigars, 1000, pre, suf))
    return regions
