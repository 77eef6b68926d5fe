import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    return np.load(os.path.join(ROOT, "tests", "golden", "head_golden.npz"), allow_pickle=False)


def golden_dp(golden, kind, D=2304):
    import numpy as np

    if kind == "zero":
        return np.zeros(D, np.float32)
    w = golden["w_values"].astype(np.float64)
    return np.log(w / (1 - w)).astype(np.float32)


def case_fields(key):
    """'wvalues_eps0.1_hard' -> ('wvalues', 0.1, True)"""
    kind, eps, mode = key.split("_")
    return kind, float(eps[3:]), mode == "hard"
