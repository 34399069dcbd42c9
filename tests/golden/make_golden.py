#!/usr/bin/env python
"""Regenerates tests/golden/ssw_golden.npz by running the UNMODIFIED reference ssw.c (compiled by oracle/Makefile into
oracle/_ref/libssw_ref.so) on seeded inputs.  Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py

Each case = one batch (shared scoring / flag) of CSR-packed pairs plus the reference's outputs: the 7 scalar s_align fields,
cigarLen (-1 where ssw_align returned NULL) and the CIGAR words.  Cases cover: the known-answer vectors of SURVEY.md section 8c,
the adversarial fuzz distribution (random matrices / gaps, 2- and 4-letter alphabets, N, lengths from 1, every flag
combination, filters), byte-mode and word-mode pairs, score_size 0/1/2, and small samples of BASELINE configs 1 and 2."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402

w = importlib.import_module("workloads")
CAP = 192


def enc(s):
    return np.array(["ACGTN".index(c) for c in s], dtype=np.int8)


def manual_batch(pairs, mat, flag, masks, **kw):
    reads = [enc(r) for r, _ in pairs]
    refs = [enc(t) for _, t in pairs]
    ro = np.concatenate([[0], np.cumsum([len(r) for r in reads])]).astype(np.int64)
    fo = np.concatenate([[0], np.cumsum([len(r) for r in refs])]).astype(np.int64)
    return w.PairBatch(np.concatenate(reads), ro, np.concatenate(refs), fo, np.array(masks, dtype=np.int32), mat=mat, flag=flag, **kw)


def main():
    assert oracle.have_ref() or os.path.exists("/root/reference"), "needs the compiled reference"
    oracle.build()
    cases = []
    dna = w.dna_matrix()
    # known-answer vectors
    kv = [("AAAA", "CCCCCC"), ("ACGTACGT", "TTACGTACGTTTACGTACGT"), ("ACGTACGT", "TTACGTACGTTTACGTACGT"), ("A", "A"), ("A", "C"), ("ACGT", "A"),
          ("NNNN", "ACGTACGT"), ("ACGTNACGT", "GGACGTAACGTGG")]
    for flag, filters, filterd in ((1, 0, 32767), (0, 0, 32767), (2, 100, 32767), (4, 0, 3), (0x0f, 0, 32767), (8, 0, 0), (0x0f, 0, 0)):
        cases.append(manual_batch(kv, dna, flag, [15, 15, 3, 15, 15, 15, 15, 15], filters=filters, filterd=filterd, name=f"known flag{flag}"))
    for ss in (0, 1, 2):
        b = w.fuzz_pairs(60, 900 + ss, flag=1, max_read=260, max_ref=300)
        b.score_size = ss
        b.name = f"fuzz score_size {ss}"
        cases.append(b)
    for seed in range(24):
        flag = [1, 0, 8, 0x0f, 2, 4, 3, 6][seed % 8]
        b = w.fuzz_pairs(80, seed, alphabet=2 if seed % 3 == 0 else 4, flag=flag, max_read=400, max_ref=450)
        b.filters = 60 if flag in (2, 3, 6) else 0
        b.filterd = 40 if flag in (4, 6) else 32767
        cases.append(b)
    cases.append(w.config1(300, seed=21))
    b = w.config1(300, seed=22); b.flag = 8; cases.append(b)
    cases.append(w.config2(400, seed=23))
    b = w.config2(200, seed=24); b.flag = 0x0f; b.masklen[:] = b.read_len.astype(np.int32); b.name = "config3-style flag 0x0f maskLen=qlen"; cases.append(b)
    b = w.make_pairs(40, (900, 1400), 2000, err=0.05, seed=25, flag=1, name="mid-length 0.9-1.4kb x 2kb"); cases.append(b)
    b = w.make_pairs(6, 9000, 10500, err=0.08, seed=26, flag=1, name="ONT-scale 9kb x 10.5kb (int16 clamp edge)", chunk=8); cases.append(b)
    b = w.make_pairs(6, 9000, 10500, err=0.08, seed=27, flag=0, name="ONT-scale +2/-3", chunk=8); b.mat = w.dna_matrix(2, 3); cases.append(b)

    out = {"ncases": np.array(len(cases))}
    for k, b in enumerate(cases):
        r, c, _ = oracle.run_batch(b.reads, b.read_off, b.refs, b.ref_off, b.masklen, b.mat, b.n, gapO=b.gapO, gapE=b.gapE, flag=b.flag, filters=b.filters,
                                   filterd=b.filterd, score_size=b.score_size, threads=8, impl="ref", cigar_cap=CAP)
        lens = r[:, 7]
        assert lens.max() <= CAP or b.read_len.max() > 2000, (b.name, lens.max())
        pre = f"c{k}_"
        out[pre + "reads"] = b.reads; out[pre + "read_off"] = b.read_off; out[pre + "refs"] = b.refs; out[pre + "ref_off"] = b.ref_off
        out[pre + "masklen"] = b.masklen; out[pre + "mat"] = np.asarray(b.mat, dtype=np.int8)
        out[pre + "params"] = np.array([b.n, b.gapO, b.gapE, b.flag, b.filters, b.filterd, b.score_size], dtype=np.int32)
        out[pre + "res"] = r; out[pre + "cigar"] = c
        out[pre + "name"] = np.array(b.name)
        print(f"case {k:2d} {b.name:50s} pairs {b.npairs:4d} word-mode {int((r[:,0] + 6 >= 255).sum()):4d} null {int((lens < 0).sum())} max cigarLen {lens.max()}")
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ssw_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
