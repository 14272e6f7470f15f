import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    """A hung kernel (e.g. a barrier that never completes) must fail the test, not stall the whole run."""
    if not config.pluginmanager.hasplugin("timeout"):
        return
    for item in items:
        if item.get_closest_marker("timeout") is None:
            item.add_marker(pytest.mark.timeout(900))


def _build_native():
    """Build whatever is missing (the driver normally calls __graft_entry__.build() first)."""
    import __graft_entry__ as ge
    ge.build()


@pytest.fixture(scope="session", autouse=True)
def native_built():
    _build_native()


@pytest.fixture(scope="session")
def oracle():
    from oracle_lib import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reflib():
    from oracle_lib import RefLib
    if not RefLib.available():
        pytest.skip("oracle/_ref not built (no /root/reference at build time)")
    return RefLib()


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "conv_golden.npz"))


@pytest.fixture(scope="session")
def strip_golden():
    import numpy as np
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "strip_golden.npz"))
