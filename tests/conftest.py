import os
import sys

import pytest

# a learner that never reaches a P2P barrier must fail a test in a minute, not after the production default of 600 s
os.environ.setdefault("B200RL_P2P_TIMEOUT_S", "60")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from distrl_llm_b200 import _capi
    _capi.load_library()
    _capi.check(_capi.lib().b200rl_check_device(), "check_device")
    return torch.device("cuda:0")
