import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
HOSTSIM = os.path.join(ROOT, "tests", "hostsim")
if HOSTSIM not in sys.path:
    sys.path.insert(0, HOSTSIM)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    path = os.path.join(ROOT, "tests", "golden", "elastic_golden.npz")
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def wb():
    """The product package with its CUDA library built (nvcc cross-compiles without a GPU)."""
    from wildboar_b200 import _build
    _build.build()
    import wildboar_b200
    return wildboar_b200
