import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    """The product package (directory name starts with a digit, hence importlib)."""
    return importlib.import_module("3d_sift_cuda_b200")


@pytest.fixture(scope="session")
def oracle():
    from oracle_bindings import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    from oracle_bindings import Reference
    if not Reference.available():
        pytest.skip("oracle/_ref/libref3dsift.so not built (needs /root/reference at build time)")
    return Reference()


@pytest.fixture(scope="session")
def engine(pkg):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    pkg.load_library()
    e = pkg.Engine(0)
    yield e
    e.close()
