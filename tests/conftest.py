import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def ref_ext(name):
    """The reference's own extension rebuilt for sm_100a (oracle/_ref/<name>.so) or None."""
    from oracle import build_ref
    if not build_ref.available(name):
        return None
    try:
        return build_ref.load_ref(name)
    except Exception:  # ABI mismatch etc. -> treat as unavailable, the oracle still checks
        return None
