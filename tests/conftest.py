import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA (sm_100a) device; run on the B200 box with -m gpu")


@pytest.fixture(scope="session")
def sd_ed():
    from oracle.weights import make_state_dict
    return make_state_dict("ed", 0)


@pytest.fixture(scope="session")
def sd_vae():
    from oracle.weights import make_state_dict
    return make_state_dict("vae", 0)


@pytest.fixture(scope="session")
def golden():
    import torch

    def load(name):
        return torch.load(os.path.join(ROOT, "tests", "golden", f"{name}_golden.pt"), weights_only=False)
    return load
