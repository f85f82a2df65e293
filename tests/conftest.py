import gzip
import json
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def _load(name):
    with gzip.open(GOLDEN / name, "rt") as f:
        return json.load(f)


@pytest.fixture(scope="session")
def ref_vectors():
    """Outputs of the unmodified reference package (tests/golden/make_golden.py)."""
    return _load("reference_vectors.json.gz")


@pytest.fixture(scope="session")
def simple3_space():
    return _load("simple3_space.json.gz")


@pytest.fixture(scope="session")
def batched_vectors():
    """Reference methods driven in the batched (greenlets-blocked-in-get_reward) schedule with finite
    n_exhausted (tests/golden/make_golden.py section 9)."""
    return _load("batched_vectors.json.gz")
