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
    """npz files with 'case/field' keys -> dict of cases."""

    def __init__(self, *names):
        self._z = {}
        for name in names:
            z = np.load(os.path.join(GOLDEN, name), allow_pickle=False)
            for k in z.files:
                assert k not in self._z, f"duplicate fixture key {k}"
                self._z[k] = z[k]
        self.cases = sorted({k.split("/")[0] for k in self._z})

    def case(self, case):
        pre = case + "/"
        return {k[len(pre):]: v for k, v in self._z.items() if k.startswith(pre)}


@pytest.fixture(scope="session")
def pins():
    return Golden("reference_pins.npz")


@pytest.fixture(scope="session")
def runs():
    return Golden("reference_runs.npz", "reference_runs2.npz")
