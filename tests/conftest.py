import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def views():
    return {v: np.load(os.path.join(GOLDEN, "views", v + ".npz"))["xyz"] for v in ("cheff000", "cheff001", "cheff002")}


@pytest.fixture(scope="session")
def golden():
    return {v: np.load(os.path.join(GOLDEN, "golden_%s.npz" % v)) for v in ("cheff000", "cheff001", "cheff002")}


def forest_path(name="synthetic-T100-D15"):
    return os.path.join(GOLDEN, "forests", name + ".yaml.gz")


@pytest.fixture(scope="session")
def main_forest(oracle):
    return oracle.load_forest_yaml(forest_path())


@pytest.fixture(scope="session")
def kpl():
    """The product library through its ctypes binding; building it is part of the fixture."""
    from keypoint_learning_b200 import build as B
    B.build_lib()
    import keypoint_learning_b200 as K
    K.load_library()
    return K
