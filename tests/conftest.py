import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    import torch
    # the CPU oracle regresses badly when torch grabs all 128 hardware threads of the GPU box's host
    torch.set_num_threads(max(1, min(os.cpu_count() or 1, 32)))
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")
    config.addinivalue_line("markers", "slow: long CPU test")


@pytest.fixture(scope="session")
def has_cuda():
    import torch
    return torch.cuda.is_available()
