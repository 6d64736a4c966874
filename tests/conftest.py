import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # in-process communicator of the multi-rank GPU tests: a rank whose peer failed gives up after 30 s, not 120
    os.environ.setdefault("MCHRG_B200_COMM_TIMEOUT_S", "30")


def load_golden(name):
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        return json.load(f)


GOLDEN_CASES = ["validation_01_h3p", "validation_02_heada", "validation_03_etliclm", "kat_eeq_actinides"]
GEOM_CASES = ["geom_h2plus", "geom_znooh"]


@pytest.fixture(scope="session")
def oracle():
    import oracle as orc
    return orc


@pytest.fixture(scope="session")
def mc():
    """The product package with a live context (GPU tests only)."""
    import multicharge_b200 as m
    m.default_context(0)
    return m


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = max(np.abs(b).max(), 1e-2)
    return float(np.abs(a - b).max() / scale)
