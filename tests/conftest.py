import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")


import pytest


@pytest.fixture(autouse=True)
def _gpu_tier_needs_the_compiled_reference(request):
    """GPU tier: parity is claimed against the compiled, unmodified reference (oracle/_ref) only.  If it is absent every `-m gpu` test
    FAILS -- no silent fall-back to the scalar restatement and no skips (VERDICT r01 weak #1b)."""
    if request.node.get_closest_marker("gpu"):
        from oracle import oracle
        try:
            oracle.require_ref()
        except RuntimeError as ex:
            pytest.fail(str(ex))
