"""GPU parity of the region realigner (megapath-nano_b200/realign/realigner, the drop-in for the reference's `realigner` shared
object): realign_reads / mpn_realign_regions through the C ABI against the golden regions of the compiled reference, live
against the compiled reference on fresh BASELINE configs[2] regions, and batched == per-region."""
import importlib

import pytest

from realign_util import golden_regions, run_reference, mismatches

pytestmark = pytest.mark.gpu

w = importlib.import_module("workloads")
R = importlib.import_module("megapath-nano_b200.realigner")


def test_realign_reads_matches_golden_regions():
    gold = golden_regions()
    got = [R.realign_reads(rg) for rg, _, _ in gold]
    bad = mismatches(got, [(p, c) for _, p, c in gold])
    assert not bad, bad[:5]


def test_batched_regions_match_golden_and_single_calls():
    gold = golden_regions()
    regions = [rg for rg, _, _ in gold]
    many = R.realign_regions(regions)
    assert not mismatches(many, [(p, c) for _, p, c in gold])
    st = R.last_stats()
    assert st["pairs"] > 1000 and st["cells"] > 0


@pytest.mark.parametrize("seed,n_frac", [(601, 0.0), (602, 0.005)])
def test_config3_live_against_compiled_reference(seed, n_frac):
    """BASELINE configs[2]: per-region haplotypes x overlapping reads, variable small batches"""
    from oracle import oracle
    oracle.require_ref()
    regions = w.config3(24, seed=seed, max_reads=400, max_haps=16, n_frac=n_frac)
    want = run_reference(regions, oracle.realigner_ref_path())
    got = R.realign_regions(regions)
    bad = mismatches(got, want)
    assert not bad, bad[:5]
    single = [R.realign_reads(rg) for rg in regions[:4]]
    assert single == got[:4]


def test_assembled_haplotypes_through_the_gpu_realigner():
    """SURVEY.md section 8f N4 -> N3 -> hot path: haplotypes assembled by realign/debruijn_graph from the reads of each window, then
    realign_regions on the GPU == the compiled reference realigner on the same inputs"""
    from oracle import oracle
    oracle.require_ref()
    D = importlib.import_module("megapath-nano_b200.debruijn")
    regions, kept = D.regions_from_windows(w.config3_windows(12, seed=71, max_reads=150))
    assert len(regions) >= 8
    want = run_reference(regions, oracle.realigner_ref_path())
    got = R.realign_regions(regions)
    bad = mismatches(got, want)
    assert not bad, bad[:5]


@pytest.mark.parametrize("seed", [81, 82, 83, 84, 85, 86])
def test_fast_pass_kernel_equals_the_kmer_index(seed):
    """the fast pass alone, GPU kernel (mpn_fastpass: bit-parallel diagonals, no index) against the host k-mer
    index path (the reference's algorithm), exact on every haplotype score and every (score, position): adversarial regions (equal-score
    placements, clamped starts, dropped haplotypes, N, tiny reads / haplotypes) plus config-3 regions; a region with a lower-case base
    is flagged by the kernel and redone on the host"""
    regions = w.fastpass_adversarial(16, seed=seed) + w.config3(4, seed=seed, max_reads=300, max_haps=12, n_frac=0.002 if seed % 2 else 0.0)
    odd = w.config3(1, seed=seed + 100, max_reads=60, max_haps=4)[0]
    odd.reads[3] = odd.reads[3][:50] + "a" + odd.reads[3][51:]
    regions.append(odd)
    s0, p0, ms = R.fastpass_only(regions, 0)
    s1, p1, _ = R.fastpass_only(regions, 1)
    assert ms > 0
    assert s0 == s1
    assert p0 == p1
    assert sum(1 for x in s1 if x > 0) > 20 and sum(1 for x in p1[1::2] if x >= 0) > 500
