import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the product library and the oracle once per session (no-ops when fresh)."""
    from regex_b200 import build as product_build
    product_build.build()
    from oracle import oracle as O
    O.build()


def load_vectors():
    import json
    with open(os.path.join(GOLDEN, "reference_tests.json")) as f:
        return json.load(f)
