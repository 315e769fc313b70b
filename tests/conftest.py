import os
import sys
from pathlib import Path

# The C oracle and numpy/scipy use OpenMP / BLAS thread pools; a handful of threads is plenty for
# the test sizes and keeps many-core boxes from oversubscribing next to torch's own pools.
os.environ.setdefault("OMP_NUM_THREADS", "4")
os.environ.setdefault("OPENBLAS_NUM_THREADS", "4")
os.environ.setdefault("MKL_NUM_THREADS", "4")

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    return O


@pytest.fixture(scope="session")
def c_oracle(oracle):
    return oracle.COracle()


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    d = ROOT / "tests" / "golden"
    return {p.stem: np.load(p, allow_pickle=False) for p in sorted(d.glob("*.npz"))}
