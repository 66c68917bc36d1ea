import os
import os.path as osp
import sys

import numpy as np
import pytest

ROOT = osp.dirname(osp.dirname(osp.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = osp.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")
    config.addinivalue_line("markers", "needs_reference: needs /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    have_ref = osp.isdir("/root/reference/models")
    for item in items:
        if "needs_reference" in item.keywords and not have_ref:
            item.add_marker(pytest.mark.skip(reason="/root/reference not present"))


@pytest.fixture(scope="session")
def golden_geometry():
    return np.load(osp.join(GOLDEN, "geometry.npz"))


@pytest.fixture(scope="session")
def golden_pointnet2():
    return np.load(osp.join(GOLDEN, "pointnet2.npz"))


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from pose2room_b200 import _lib
    _lib.load()  # fail loudly if the extension is missing on a GPU box
    torch.backends.cuda.matmul.allow_tf32 = False   # torch references in the tests must be true fp32
    torch.backends.cudnn.allow_tf32 = False
    return torch.device("cuda:0")
