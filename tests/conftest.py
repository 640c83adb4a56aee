import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def kat(golden_dir):
    import json
    with open(os.path.join(golden_dir, "kat.json")) as fh:
        return json.load(fh)["entries"]


@pytest.fixture(scope="session")
def fixture_pixels(golden_dir):
    import numpy as np
    return dict(np.load(os.path.join(golden_dir, "fixture_pixels.npz")))


@pytest.fixture(scope="session")
def jg():
    """The product library through its ctypes plumbing; GPU tests fail (not skip) if it is missing."""
    import imagecodecs_b200 as m
    return m


@pytest.fixture(scope="session")
def gpu(jg):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device in this environment")
    assert os.path.exists(jg.LIB_PATH), "libjpeg_gpu.so must be built in-tree (python -m imagecodecs_b200.build)"
    assert jg.init([0]) >= 1
    return jg
