import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (oracle/tk_oracle.py) -- the checker."""
    return entry.load_oracle()


@pytest.fixture(scope="session")
def tk():
    """The product package (ctypes over libtensorkrylov_b200.so); builds the library if it is missing."""
    if not os.path.exists(os.path.join(entry.PKG_DIR, "libtensorkrylov_b200.so")):
        entry.build()
    pkg = entry.load_package()
    # Parity tests compare with the oracle, whose spectral data (minors of A_1) comes from LAPACK like the reference's:
    # feed the same numbers (what the Julia wrapper does with Julia's own SpectralData).  The library's own
    # eigen-extremes (tk_schedule) differ from LAPACK's by eps * cond(minor) and have their own tests.
    pkg.api.DEFAULT_SPECTRAL = "lapack"
    return pkg


@pytest.fixture(scope="session")
def tables(orc, tk):
    return orc.ExpSumTables.from_packed(tk.TABLES_PATH)


@pytest.fixture(scope="session")
def gpu(tk):
    if tk.device_count() < 1:
        pytest.skip("no CUDA device visible")
    return 0


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))
