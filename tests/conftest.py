import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def pkg():
    return importlib.import_module("26al-nbody_b200")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "enrich_golden.npz"))


def golden_case(g, tag):
    return {k[len(tag) + 1:]: g[k] for k in g.files if k.startswith(tag + "_")}


@pytest.fixture(scope="session")
def ctx(pkg):
    """One CUDA context for the whole GPU session (fails loudly without a B200)."""
    c = pkg.Context(0)
    yield c
    c.close()
