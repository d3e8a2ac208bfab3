import json
import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


def load_golden(name):
    with open(os.path.join(GOLDEN_DIR, name)) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def jssp_golden():
    return load_golden("jssp_hamiltonians.json")


@pytest.fixture(scope="session")
def genome_golden():
    return load_golden("genomes.json")


@pytest.fixture(scope="session")
def cvar_golden():
    return load_golden("cvar.json")
