import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def built():
    import __graft_entry__ as g
    g.build()
    return True


@pytest.fixture(scope="session")
def engine_mod(built):
    from nextdenovo_b200 import engine
    return engine


@pytest.fixture(scope="session")
def reads_full(golden_dir):
    import nextdenovo_b200 as nd
    return nd.read_2bit(os.path.join(golden_dir, "reads_test.seed.2bit"))


@pytest.fixture(scope="session")
def reads_sub(golden_dir):
    import nextdenovo_b200 as nd
    return nd.read_2bit(os.path.join(golden_dir, "sub120.2bit"))
