import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
TESTS = os.path.join(ROOT, "tests")
if TESTS not in sys.path:
    sys.path.insert(1, TESTS)          # `from fakes import fake_gallery` (test doubles)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture()
def store_dir(tmp_path, monkeypatch):
    d = tmp_path / "rbod_store"
    monkeypatch.setenv("RBOD_STORE_DIR", str(d))
    import qdrant_client

    yield str(d)
    qdrant_client._close_all()


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
