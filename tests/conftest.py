import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    return load


@pytest.fixture(scope="session")
def oracle():
    from oracle import ref_oracle
    ref_oracle.build()
    return ref_oracle


@pytest.fixture(scope="session")
def slamfe():
    import slamfe as pkg
    return pkg


@pytest.fixture(params=["mma", "int"])
def matcher_kernel(request):
    """Runs a test once per matcher kernel: "mma" = tcgen05 tensor-core kernel (csrc/hamming_mma.cu, the
    default), "int" = INT-pipe carry-save kernel (csrc/hamming.cu).  Both must give identical keys."""
    from slamfe import ops
    old = ops.set_matcher_kernel(request.param)
    yield request.param
    ops.set_matcher_kernel(old)
