import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def load(name):
        if name not in cache:
            cache[name] = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
        return cache[name]

    return load


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc
    orc.build()
    return orc


@pytest.fixture(scope="session")
def synth():
    from livecell_instance_segmentation_b200 import synth as s
    return s


@pytest.fixture
def tune():
    """Set liblcr tuning switches for one test (the library reads the environment only once, so tests go through
    lcr_set_tuning); everything set here is reset afterwards."""
    from livecell_instance_segmentation_b200 import _lib
    touched = set()

    def setter(**switches):
        for k, v in switches.items():
            touched.add(k)
            _lib.set_tuning(k, v)

    yield setter
    for k in touched:
        _lib.set_tuning(k, os.environ.get(k))


@pytest.fixture
def roi_cpu_coords():
    """RoIAlign sample coordinates rounded as torchvision's CPU op rounds them — the rule of the CPU-generated golden vectors
    and of the oracle's default — instead of the library default (torchvision's CUDA op, what the reference runs on a GPU)."""
    from livecell_instance_segmentation_b200 import ops
    prev = ops.set_roi_coord_rule("cpu")
    yield
    ops.set_roi_coord_rule(prev)


def pytest_collection_modifyitems(config, items):
    """GPU tests are skipped (not failed) where no CUDA device exists."""
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
