import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "tests"))
sys.path.insert(0, os.path.join(REPO, "root-sim_b200"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_built():
    """Build the CPU oracle (and, where /root/reference exists, the real reference under oracle/_ref)."""
    subprocess.run(["bash", os.path.join(REPO, "oracle", "build_ref.sh")], check=True, capture_output=True)
    subprocess.run(["make", "-C", os.path.join(REPO, "oracle")], check=True, capture_output=True)
    return os.path.join(REPO, "oracle", "_build")


def have_reference_binary(model):
    return os.path.exists(os.path.join(REPO, "oracle", "_ref", model + "_trace"))


def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
