import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    """tests/golden/reference_golden.npz -> {case: {key: ndarray}} (outputs of the REAL
    reference, produced by tests/golden/make_golden.py in the build container)."""
    cases = {}
    for fn in ("reference_golden.npz", "reference_golden_postfusion.npz", "reference_golden_staging.npz", "reference_golden_grid.npz"):
        z = np.load(os.path.join(ROOT, "tests", "golden", fn))
        for k in z.files:
            c, name = k.split("/", 1)
            cases.setdefault(c, {})[name] = z[k]
    return cases
