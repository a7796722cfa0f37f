import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) device; run with -m gpu")


@pytest.fixture(scope="session")
def ctx():
    """Session-wide device context (GPU tests only).  No skip on failure: on a GPU
    box a missing library / device must fail loudly."""
    import torch
    assert torch.cuda.is_available(), "GPU test selected but no CUDA device is visible"
    from page_segmentation_b200 import runtime
    return runtime.get_context(0)
