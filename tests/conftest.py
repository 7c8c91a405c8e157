import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def native_lib():
    """Build (if stale) and load libcerebro_b200.so; CPU-safe (no device calls)."""
    from cerebro_b200 import build, _lib

    build.build()
    return _lib.load()


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        # a plain `pytest` on a CPU-only box skips the GPU tests; the GPU visit (tools/gpu_round.sh) sets CB_REQUIRE_GPU=1 so that
        # a box without a usable device fails loudly instead of reporting a green, empty run
        if os.environ.get("CB_REQUIRE_GPU") == "1":
            pytest.fail("GPU test selected but no CUDA device is visible")
        pytest.skip("no CUDA device visible")
    return 0
