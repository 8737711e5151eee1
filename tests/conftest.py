import glob
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
COV_PARAMS = [(10, 8), (32, 10), (1, 5), (3, 1), (10, 32), (2, 3)]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_inputs():
    names = []
    for pat in ("*.fa", "*.fq", "*.fa.gz"):
        names += [os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, pat))]
    return sorted(names)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
