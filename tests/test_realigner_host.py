"""CPU tests of the region realigner's HOST logic (k-mer fast pass, pair planning, '='/'X' CIGAR post-processing, CIGAR
composition: csrc/realign_region.cpp + csrc/ssw_cpp_layer.cpp).  The product links these files with the CUDA engine; here the
test-only build oracle/_hosttest/realigner_hosttest.so links the same two sources with oracle/host_shim.cpp, which answers
mpn_align_batch from the CPU checkers, so the logic around the kernels can be pinned without a GPU:
  * against the golden regions produced by the compiled reference realigner,
  * live against the compiled reference (when oracle/_ref is present) on fresh seeded regions,
  * batched mpn_realign_regions == per-region realign_reads."""
import importlib
import os

import pytest

from oracle import oracle
from realign_util import golden_regions, run_reference, mismatches

w = importlib.import_module("workloads")
R = importlib.import_module("megapath-nano_b200.realigner")


@pytest.fixture(scope="module")
def hostlib():
    path = oracle.build_hosttest()
    os.environ["MPN_SHIM_REF"] = oracle.ref_path() if oracle.have_ref() else ""
    return path


def test_host_logic_matches_golden_regions(hostlib):
    gold = golden_regions()
    got = [R.realign_reads(rg, hostlib) for rg, _, _ in gold]
    bad = mismatches(got, [(p, c) for _, p, c in gold])
    assert not bad, bad[:5]
    assert sum(len(p) for _, p, _ in gold) > 1000


def test_batched_regions_equal_per_region_calls(hostlib):
    regions = [rg for rg, _, _ in golden_regions()]
    one = [R.realign_reads(rg, hostlib) for rg in regions]
    many = R.realign_regions(regions, hostlib)
    assert one == many
    assert R.realign_regions_packed(regions, hostlib) == many        # flat-buffer entry point
    st = R.last_stats(hostlib)
    assert st["pairs"] > 0 and st["cells"] > 0
    assert R.realign_regions_packed([], hostlib) == []
    assert R.realign_regions([], hostlib) == []


@pytest.mark.skipif(not oracle.have_ref(), reason="compiled reference realigner not present")
@pytest.mark.parametrize("seed,n_frac", [(501, 0.0), (502, 0.0), (503, 0.01)])
def test_host_logic_live_against_compiled_reference(hostlib, seed, n_frac):
    regions = w.config3(6, seed=seed, max_reads=150, max_haps=10, n_frac=n_frac)
    want = run_reference(regions, oracle.realigner_ref_path())
    got = R.realign_regions(regions, hostlib)
    bad = mismatches(got, want)
    assert not bad, bad[:5]


@pytest.mark.parametrize("seed", [81, 82, 83, 84])
def test_fast_pass_diagonal_run_formulation_equals_the_kmer_index(hostlib, seed):
    """the fast pass alone: the index-free "diagonal run" formulation that the GPU kernel implements (here its scalar stand-in in
    oracle/host_shim.cpp) against the product's host k-mer index path (the reference's algorithm, pinned end to end above), on regions built to
    hit the order-dependent rules: equal-score placements, clamped starts, dropped haplotypes, N, tiny reads and haplotypes"""
    regions = w.fastpass_adversarial(10, seed=seed) + w.config3(2, seed=seed, max_reads=80, max_haps=6, n_frac=0.002)
    s0, p0, _ = R.fastpass_only(regions, 0, hostlib)
    s1, p1, _ = R.fastpass_only(regions, 1, hostlib)
    assert s0 == s1
    assert p0 == p1
    assert sum(1 for x in s1 if x > 0) > 10 and sum(1 for x in p1[1::2] if x >= 0) > 100
