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

w = importlib.import_module("megapath-nano_b200.workloads")
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
