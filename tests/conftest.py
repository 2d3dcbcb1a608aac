import gzip
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ec10k_sfa():
    with gzip.open(os.path.join(ROOT, "tests", "golden", "ec10k_reads.sfa.gz"), "rb") as f:
        return f.read()


@pytest.fixture(scope="session")
def ec10k_golden():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "ec10k_survey.json")) as f:
        return json.load(f)
