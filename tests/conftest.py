import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def polyvt_golden():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "polyvt_golden.npz"))


@pytest.fixture(scope="session")
def scenario_fixture():
    import dlii_jssp_b200 as J
    return J.scenario.scenario_from_fixture(os.path.join(GOLDEN, "scenario_8_3_4_565.npz"))
