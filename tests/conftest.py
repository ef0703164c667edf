import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    import __graft_entry__ as ge
    ge.build_harness()
    ge.build_oracle()
    return ge


@pytest.fixture(scope="session")
def port(built):
    from oracle import pyref
    return pyref.Port()


@pytest.fixture(scope="session")
def reference(built):
    from oracle import pyref
    try:
        return pyref.Reference("ser")
    except FileNotFoundError:
        pytest.skip("oracle/_ref not built (no /root/reference here and no prebuilt copy)")


@pytest.fixture(scope="session")
def hc_lib():
    """The C-ABI library on a CUDA device, tables uploaded."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from nyx_b200 import capi, synth
    from oracle import pyref
    hc = capi.NyxHC()
    hc.tables_upload(hc.tabulate_rates(pyref.TREECOOL, synth.mean_rhob()))
    return hc
