import ctypes
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def emu_backend():
    """Host compilation of csrc/mpc_b200.cu (tests/emu): the SAME kernel code run as loops, used to test the
    host logic in the GPU-less container.  Test infrastructure — the product never loads it."""
    import __graft_entry__ as g
    from python_mpc_b200 import _lib
    return _lib.Backend(ctypes.CDLL(g.build_emu()), "cpu")


@pytest.fixture(scope="session")
def cuda_backend():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from python_mpc_b200 import _lib
    return _lib.cuda_backend()


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    d = os.path.join(ROOT, "tests", "golden")
    return {f[:-4]: np.load(os.path.join(d, f)) for f in os.listdir(d) if f.endswith(".npz")}
