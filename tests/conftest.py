import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def sandbag_cases():
    return np.load(os.path.join(GOLDEN, "sandbag_cases.npz"))


@pytest.fixture(scope="session")
def sandbag_minref():
    return json.load(open(os.path.join(GOLDEN, "sandbag_minref.json")))


@pytest.fixture(scope="session")
def sandbag_api():
    return json.load(open(os.path.join(GOLDEN, "sandbag_api.json")))


@pytest.fixture(scope="session")
def cyclone_cases():
    return (np.load(os.path.join(GOLDEN, "cyclone_cases.npz")),
            json.load(open(os.path.join(GOLDEN, "cyclone_cases.json"))))


@pytest.fixture(scope="session")
def mt_golden():
    return json.load(open(os.path.join(GOLDEN, "mt19937.json")))


SANDBAG_CASE_NAMES = ["c3_f65", "c3_f50", "c3_f40", "c2_f65", "c2_f50", "c1_f65", "c1_f40", "c10_f65",
                      "c10_f70", "c3_f100", "c3_tiny", "c4_skew"]
CYCLONE_CASE_NAMES = ["a_poisson", "b_lognormal", "c_minpairs", "d_miniter", "e_quirks"]
