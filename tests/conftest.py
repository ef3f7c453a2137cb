import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from tests import oracle_ffi
    return oracle_ffi.Oracle()


@pytest.fixture(scope="session")
def refcnf():
    from tests import oracle_ffi
    lib = oracle_ffi.load_ref_cnf()
    if lib is None:
        pytest.skip("oracle/_ref/libref_cnf.so not built (needs /root/reference)")
    return lib


@pytest.fixture
def tc_env(monkeypatch):
    """Set PQH_TC_* knobs for one test.  The library reads them when a device context is
    created (pqh_knn_init), so every change re-creates the context, and so does the teardown."""
    from pq_huffman_b200 import knn

    def apply(**kv):
        for key, val in kv.items():
            if val is None:
                monkeypatch.delenv(key, raising=False)
            else:
                monkeypatch.setenv(key, str(val))
        knn.init([0])

    yield apply
    monkeypatch.undo()
    knn.init([0])
