"""Window haplotype assembler (include/debruijn_graph.h): the Boost-free C++ implementation against
(a) the reference itself -- its unmodified debruijn_graph.cpp compiled over oracle/boost_shim (oracle/_ref/debruijn_graph_ref), live and
through the committed golden windows made from it -- (b) the string-keyed restatement oracle/dbg_oracle.py, and (c) properties that hold
for the reference by construction.  Host code only -> runs in the CPU tier."""
import ctypes
import importlib
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
w = importlib.import_module("workloads")
D = importlib.import_module("megapath-nano_b200.debruijn")
from oracle import dbg_oracle


@pytest.fixture(scope="module", autouse=True)
def built():
    if not os.path.exists(D.LIB_PATH):
        subprocess.run(["bash", os.path.join(ROOT, "build.sh"), "dbg"], check=True)
    D.load()


def oracle(ref, reads, lowq):
    return dbg_oracle.get_consensus(ref, ",".join(reads), ",".join(lowq))


def test_exports_and_struct_layout():
    L = D.load()
    for sym in ("get_consensus", "free_memory", "mpn_dbg_consensus_packed", "mpn_dbg_free"):
        assert hasattr(L, sym)
    assert ctypes.sizeof(D.DBGPointer) == 8 + 500 * 8            # int + padding, char*[500] (debruijn_graph.h:40-44)


@pytest.mark.parametrize("seed", [21, 22, 23])
def test_seeded_windows_against_restatement(seed):
    wins = w.dbg_windows(12, seed=seed, max_reads=150)
    got, ks = D.consensus_windows(wins, with_k=True)
    multi = 0
    for (ref, reads, lowq), g, k in zip(wins, got, ks):
        exp, ek = oracle(ref, reads, lowq)
        assert g == exp and k == ek
        assert D.get_consensus(ref, reads, lowq) == exp          # the reference's per-window call gives the same
        multi += len(exp) > 1
    assert multi >= 3                                             # planted variants are found


def test_properties():
    for ref, reads, lowq in w.dbg_windows(10, seed=31, max_reads=120):
        haps, k = D.consensus_windows([(ref, reads, lowq)], with_k=True)
        haps, k = haps[0], k[0]
        if k == 0:
            assert haps == []
            continue
        assert haps == sorted(haps)
        assert ref in haps                                        # reference edges are never pruned: its path always survives
        kmers = set()
        for s in [ref] + reads:
            kmers.update(s[i:i + k] for i in range(len(s) - k + 1))
        for h in haps:
            assert h.startswith(ref[:k])                          # every path starts at the source ...
            assert h.endswith(ref[-k:])                           # ... and, after pruning, ends at the sink
            assert all(h[i:i + k] in kmers for i in range(len(h) - k + 1))


def test_edge_cases():
    rng = np.random.default_rng(5)
    ref = "".join("ACGT"[i] for i in rng.integers(0, 4, size=300))
    cases = {
        "no reads": (ref, [""], [""]),
        "reference too short": ("ACGTACGTAC", ["ACGTACGTACGT"], [""]),
        "eleven bases": ("ACGTTGCATGA", ["ACGTTGCATGA"] * 3, [""] * 3),
        "reads shorter than k": (ref, ["ACGTA", "AC", ""], ["", "", ""]),
        "every k cyclic": ("A" * 150, ["A" * 100], [""]),
        "read makes a cycle": (ref, [ref[100:140] + ref[80:140]] * 2, [""] * 2),
        "tandem repeat in the reference": (ref[:100] + "CAG" * 15 + ref[100:], [ref[60:100] + "CAG" * 16 + ref[100:150]] * 3, [""] * 3),
        "SNV seen once is pruned": (ref, [ref[50:120] + "T" + ref[121:200]], [""]),
        "SNV seen twice": (ref, [ref[50:120] + ("T" if ref[120] != "T" else "G") + ref[121:200]] * 2, [""] * 2),
        "low quality hides the SNV": (ref, [ref[50:120] + ("T" if ref[120] != "T" else "G") + ref[121:200]] * 2, ["70", "70"]),
        "N splits the read": (ref, [ref[50:120] + "N" + ref[121:200]] * 2, [""] * 2),
        "bad base inside the first k-mer": (ref, ["ACGN" + ref[60:200]] * 2, ["", "1 2"]),
        "fewer quality fields than reads": (ref, [ref[20:150], ref[30:160]], [""]),
        "insertion and deletion": (ref, [ref[40:100] + "GATTACA" + ref[100:170], ref[150:200] + ref[212:280]] * 2, [""] * 4),
    }
    for name, (r, reads, lowq) in cases.items():
        exp, ek = oracle(r, reads, lowq)
        got = D.get_consensus(r, reads, lowq)
        assert got == exp, name
    assert D.get_consensus(*cases["every k cyclic"]) == []
    assert D.get_consensus(*cases["reference too short"]) == []
    assert len(D.get_consensus(*cases["SNV seen once is pruned"])) == 1
    assert len(D.get_consensus(*cases["SNV seen twice"])) == 2
    assert len(D.get_consensus(*cases["low quality hides the SNV"])) == 1
    assert len(D.get_consensus(*cases["insertion and deletion"])) == 4


def test_live_path_cap():
    """nine well-separated biallelic SNVs -> 512 paths: more than 256 alive -> no haplotypes at all (debruijn_graph.cpp:287-289);
    eight -> 256 paths, allowed"""
    rng = np.random.default_rng(9)
    ref = "".join("ACGT"[i] for i in rng.integers(0, 4, size=600))
    def alt(n):
        s = list(ref)
        for j in range(n):
            p = 40 + 60 * j
            s[p] = "ACGT"[("ACGT".index(s[p]) + 1) % 4]
        return "".join(s)
    for n, expect in ((8, 256), (9, 0)):
        a = alt(n)
        reads = [a[i:i + 50] for i in range(0, 550, 5)] * 2
        got = D.get_consensus(ref, reads, [""] * len(reads))
        exp, _ = oracle(ref, reads, [""] * len(reads))
        assert got == exp and len(got) == expect


def test_assembled_regions_through_the_realigner_host_logic():
    """the chain of realign_illumina_reads.py:551-612 on the CPU tier: windows -> assembler -> realigner inputs -> the product's
    realigner host logic (linked with the CPU checkers, oracle/_hosttest) == the compiled reference realigner on the same inputs"""
    from oracle import oracle
    from realign_util import run_reference, mismatches
    R = importlib.import_module("megapath-nano_b200.realigner")
    wins = w.config3_windows(5, seed=61, max_reads=90)
    regions, kept = D.regions_from_windows(wins)
    assert len(regions) >= 3 and all(len(rg.haplotypes) >= 2 and rg.reference in rg.haplotypes for rg in regions)
    for rg, k in zip(regions, kept):                              # planted variants come back as haplotypes
        assert all(h.startswith(rg.reference[:rg.ref_prefix + 10]) and h.endswith(rg.reference[-rg.ref_suffix - 10:]) for h in rg.haplotypes)
    if not oracle.have_ref():
        pytest.skip("compiled reference realigner not present")
    hostlib = oracle.build_hosttest()
    os.environ["MPN_SHIM_REF"] = oracle.ref_path()
    want = run_reference(regions, oracle.realigner_ref_path())
    got = R.realign_regions(regions, hostlib)
    bad = mismatches(got, want)
    assert not bad, bad[:5]
    moved = sum(c != f"{len(r)}M" for rg, (p, cs) in zip(regions, got) for r, c in zip(rg.reads, cs))
    assert moved > 0                                              # reads of the alternative haplotypes get new CIGARs


def test_live_against_the_compiled_reference():
    """the product and the reference object (unmodified debruijn_graph.cpp over the Boost stand-in), same ctypes call sequence, on seeded
    windows incl. repeats / N / low-quality positions and on windows around the 256-path cap"""
    from oracle import oracle as O
    ref_obj = O.dbg_ref_path()
    if ref_obj is None:
        pytest.skip("oracle/_ref/debruijn_graph_ref not built (needs /root/reference; the prebuilt object travels to the GPU box)")
    wins = w.dbg_windows(24, seed=7101, max_reads=120) + w.dbg_windows(8, seed=7102, max_reads=40, repeat_frac=1.0) + \
        w.dbg_windows(8, seed=7103, max_reads=200, n_frac=0.01, lowq_frac=0.05) + [(r, rd, lq) for r, rd, lq, _ in w.dbg_cap_windows(seed=43)]
    multi = 0
    for ref, reads, lowq in wins:
        want = D.get_consensus(ref, reads, lowq, lib_path=ref_obj)
        assert D.get_consensus(ref, reads, lowq) == want
        assert oracle(ref, reads, lowq)[0] == want                # the restatement agrees with the reference too
        multi += len(want) > 1
    assert multi >= 10
    caps = w.dbg_cap_windows(seed=43)
    assert [len(D.get_consensus(r, rd, lq)) for r, rd, lq, _ in caps] == [n if n <= 256 else 0 for _, _, _, n in caps]


def test_golden_windows():
    """committed fixtures (tests/golden/dbg_golden.json.gz, made by tests/golden/make_golden_dbg.py from the compiled reference): the product
    and the restatement still give them"""
    import gzip, json
    with gzip.open(os.path.join(ROOT, "tests", "golden", "dbg_golden.json.gz"), "rt") as f:
        doc = json.load(f)
    assert len(doc) >= 20
    got, ks = D.consensus_windows([(d["ref"], d["reads"], d["lowq"]) for d in doc], with_k=True)
    for d, g, k in zip(doc, got, ks):
        assert g == d["haplotypes"] and k == d["k"]
        assert oracle(d["ref"], d["reads"], d["lowq"]) == (d["haplotypes"], d["k"])
    assert sum(len(d["haplotypes"]) > 1 for d in doc) >= 5
