import importlib
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = "multiple-sequence-alignment-openmp-openmpi_b200"
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def seqalk():
    """The product package (loads libseqalk.so; fails loudly if it is not built)."""
    pkg = importlib.import_module(PKG)
    pkg.lib()
    return pkg


@pytest.fixture(scope="session")
def workloads():
    return importlib.import_module(PKG + ".workloads")


@pytest.fixture(scope="session")
def oracle():
    """The CPU checker (oracle/); only tests may use it."""
    from oracle import binding

    binding.lib()
    return binding


@pytest.fixture(scope="session")
def golden():
    return json.load(open(os.path.join(GOLDEN_DIR, "expected.json")))


@pytest.fixture(scope="session")
def load_fixture(workloads, golden):
    def _load(name):
        e = golden[name]
        pxy, pgap, genes = workloads.parse_input(open(os.path.join(GOLDEN_DIR, e["file"]), "rb").read())
        if "permutation" in e:
            genes = [genes[i] for i in e["permutation"]]
        return pxy, pgap, genes, e

    return _load


@pytest.fixture(scope="session")
def underscore_cases():
    """Inputs containing '_' with the unmodified reference's outputs (tests/golden/make_golden.py)."""
    return json.load(open(os.path.join(GOLDEN_DIR, "underscore.json")))


@pytest.fixture(scope="session")
def fullsize():
    """Oracle goldens of BASELINE configs 3-5 at full size (tests/golden/make_golden_fullsize.py)."""
    return json.load(open(os.path.join(GOLDEN_DIR, "fullsize.json")))
