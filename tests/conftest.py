"""pytest configuration: the `gpu` marker, repo-root imports, shared fixtures."""
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc
    orc.build()
    return orc


@pytest.fixture(scope="session")
def minidata(oracle):
    """(colnames, coldescs, columns) of the reference's 500-row table."""
    return oracle.load_tsv(os.path.join(GOLDEN, "minidata.tsv"))


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN, "phase3_golden.json")) as f:
        return json.load(f)["entries"]


@pytest.fixture(scope="session")
def ctx():
    """One libmbcol context on cuda:0 for the whole GPU session."""
    import mbcol
    c = mbcol.Context(0)
    yield c
    c.close()
