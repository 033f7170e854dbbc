import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN, "reference_vectors.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def lexa():
    with open(os.path.join(GOLDEN, "lexa_sites.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def lexa_site_lists(lexa):
    return [mo["sites"] for mo in lexa["motifs"]]


@pytest.fixture(scope="session")
def lexa_weights(lexa_site_lists):
    counts = [len(s) for s in lexa_site_lists]
    return {"equal": [1.0 / len(counts)] * len(counts),
            "site_count": [c / float(sum(counts)) for c in counts]}
