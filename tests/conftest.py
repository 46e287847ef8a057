import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")


@pytest.fixture(scope="session")
def tf_cuda():
    """The CUDA-enabled TensorFrost module initialised on cuda:0.  One backend per process (reference design), so every
    GPU test shares it.  No fallback: if the module, the runtime library or the device is missing this raises."""
    import tensorfrost_b200
    return tensorfrost_b200.load()


@pytest.fixture(scope="session")
def tfcuda_lib():
    """libtfcuda.so through ctypes, initialised on cuda:0 (the C-ABI the GPU parity tests call through)."""
    from tensorfrost_b200 import abi
    abi.init(-1)
    return abi.lib()
