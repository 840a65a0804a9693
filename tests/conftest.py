import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def native_lib():
    """Path of the built C-ABI library (built on demand; nvcc cross-compiles without a GPU)."""
    from mahi_mpc_b200 import api, build
    if not os.path.exists(api.LIB_PATH):
        build.build_all()
    return api.LIB_PATH


@pytest.fixture(scope="session")
def coracle():
    """C restatement (oracle/mpc_oracle.c) -- the checker, never the thing under test."""
    from oracle import coracle as co
    co.build()
    return co.COracle("O3")


@pytest.fixture(scope="session")
def golden():
    import json
    gd = os.path.join(ROOT, "tests", "golden")
    with open(os.path.join(gd, "dp_nodes.json")) as fh:
        nodes = json.load(fh)
    with open(os.path.join(gd, "dp_solves.json")) as fh:
        solves = json.load(fh)
    return dict(nodes=nodes, solves=solves)
