import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Make sure the native pieces exist (cheap no-op when they are up to date)."""
    import __graft_entry__ as g
    g.build()
    return True


@pytest.fixture(scope="session")
def oracle(built):
    from oracle.pyoracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference(built):
    from oracle.pyoracle import Reference, have_reference
    if not have_reference():
        pytest.skip("oracle/_ref/libsf_ref.so not available (built only where /root/reference exists)")
    return Reference()


@pytest.fixture(scope="session")
def ctx(built):
    from slowflow_b200 import Context
    c = Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def mt_checker(built):
    """(library, symbol prefix) of the CPU checker for the multi-frame path: the reference's own unmodified driver
    (oracle/_ref) where it was built, else the bit-identical restatement (oracle/sf_oracle_mt.cpp)."""
    from oracle.pyoracle import Oracle, Reference, have_reference
    if have_reference():
        return Reference().lib, "sf_ref_"
    return Oracle().lib, "sfo_"
