import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent
for p in (str(REPO), str(REPO / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_table():
    from oracle import boys

    return boys.table()


@pytest.fixture(scope="session")
def engine(oracle_table):
    """Engine on cuda:0 using the SAME Boys table as the oracle (the reference's table depends on
    nmax at the 1e-9 level, so parity requires sharing it)."""
    import sympleints_b200

    eng = sympleints_b200.get_engine(0, 4, 5, boys_table=oracle_table)
    yield eng
    eng.close()
