import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def g19():
    return importlib.import_module("2019global_b200")


@pytest.fixture(scope="session")
def abi():
    return importlib.import_module("2019global_b200.abi")


def _checker(which):
    from oracle import binding
    if not binding.available(which):
        pytest.skip("oracle library '%s' not built (run __graft_entry__.build())" % which)
    return binding.CheckerLib(which)


@pytest.fixture(scope="session")
def oracle():
    """This repo's plain-C restatement of the reference (oracle/ref_restate.c)."""
    return _checker("oracle")


@pytest.fixture(scope="session")
def reflib():
    """The unmodified reference compiled behind oracle/ref_harness (oracle/_ref)."""
    return _checker("ref")
