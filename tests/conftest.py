import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG_NAME = "indirect_learning_pose-shape_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def pkg():
    return importlib.import_module(PKG_NAME)


@pytest.fixture(scope="session")
def smpl_io(pkg):
    return pkg.smpl_io


@pytest.fixture(scope="session")
def host_model(smpl_io):
    """Seeded synthetic SMPL-shaped model on the real template geometry (the real pickle is not shipped)."""
    return smpl_io.make_synthetic_smpl(seed=0)


@pytest.fixture(scope="session")
def parts_by_vs(smpl_io):
    return {vs: smpl_io.golden_part_vertices(vs) for vs in (None, 2, 5)}


@pytest.fixture(scope="session")
def make_params(pkg):
    synth = importlib.import_module(PKG_NAME + ".synth")
    return synth.make_params
