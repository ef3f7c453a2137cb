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
