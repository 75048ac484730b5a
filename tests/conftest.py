import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure).  Builds oracle/liboracle.so on first use."""
    from oracle import oracle as orc
    orc.lib()
    return orc


@pytest.fixture(scope="session")
def golden_api():
    import json
    with open(os.path.join(GOLDEN, "api.json")) as fh:
        return json.load(fh)


@pytest.fixture()
def in_repo(monkeypatch):
    """CLI golden cases use repo-relative input paths."""
    monkeypatch.chdir(REPO)
