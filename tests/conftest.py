"""Test configuration: puts the repo root and the drop-in module directory on sys.path, registers
the `gpu` marker, builds libmdm_sm100.so if it is missing (nvcc cross-compiles without a GPU)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "masked-diffusion-model_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    from mdm_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        from mdm_b200 import build
        build.build()
    return _lib.LIB_PATH


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"))
    return load
