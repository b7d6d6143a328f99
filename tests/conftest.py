"""Shared fixtures: repo root on sys.path, `gpu` marker, golden loaders."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Golden:
    """npz with 'case/field' keys -> dict of cases."""

    def __init__(self, name):
        self._z = np.load(os.path.join(GOLDEN, name), allow_pickle=False)
        self.cases = sorted({k.split("/")[0] for k in self._z.files})

    def case(self, case):
        pre = case + "/"
        return {k[len(pre):]: self._z[k] for k in self._z.files if k.startswith(pre)}


@pytest.fixture(scope="session")
def pins():
    return Golden("reference_pins.npz")


@pytest.fixture(scope="session")
def runs():
    return Golden("reference_runs.npz")
