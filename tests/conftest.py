import importlib
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o
    o.lib()
    return o


@pytest.fixture(scope="session")
def s256():
    """The product package (directory name has a hyphen, so import by string)."""
    return importlib.import_module("secp256k1-voi_b200")


@pytest.fixture(scope="session")
def engine(s256):
    """A live engine on cuda:0 through the C ABI; fails loudly without a GPU."""
    eng = s256.Engine()
    yield eng
    eng.close()
