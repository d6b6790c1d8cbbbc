import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(params=[0, 1], ids=["ffma_fp32", "tcgen05_tf32"])
def backend(request):
    """Run a GPU test under both convolution back ends (include/dincae_b200.h:
    dincae_set_conv_backend).  Yields the back-end id; restores the default afterwards."""
    from dincae_b200 import _lib
    lib = _lib.load()
    assert lib.dincae_set_conv_backend(request.param) == 0
    yield request.param
    lib.dincae_set_conv_backend(1)
