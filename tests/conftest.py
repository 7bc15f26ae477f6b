import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def lib_built():
    """Build libkbotstep.so once per session if it is missing (nvcc cross-compiles without a GPU)."""
    import kbot_joystick_b200 as k

    if not k.LIB_PATH.exists():
        import __graft_entry__ as g

        g.build()
    return k.LIB_PATH


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
