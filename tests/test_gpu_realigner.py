"""GPU parity of the region realigner (megapath-nano_b200/realign/realigner, the drop-in for the reference's `realigner` shared
object): realign_reads / mpn_realign_regions through the C ABI against the golden regions of the compiled reference, live
against the compiled reference on fresh BASELINE configs[2] regions, and batched == per-region."""
import importlib

import pytest

from realign_util import golden_regions, run_reference, mismatches

pytestmark = pytest.mark.gpu

w = importlib.import_module("megapath-nano_b200.workloads")
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
    if not oracle.have_ref():
        pytest.skip("compiled reference realigner not present")
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
    if not oracle.have_ref():
        pytest.skip("compiled reference realigner not present")
    D = importlib.import_module("megapath-nano_b200.debruijn")
    regions, kept = D.regions_from_windows(w.config3_windows(12, seed=71, max_reads=150))
    assert len(regions) >= 8
    want = run_reference(regions, oracle.realigner_ref_path())
    got = R.realign_regions(regions)
    bad = mismatches(got, want)
    assert not bad, bad[:5]
