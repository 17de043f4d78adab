import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


import pytest


@pytest.fixture(autouse=True)
def _default_precision():
    """Every test starts and ends in the library's default arithmetic mode (bf16x3 = the hand-written tensor-core path)."""
    yield
    mod = sys.modules.get("aerial_depth_completion_b200.dense")
    if mod is not None and mod.PRECISION != "bf16x3":
        mod.set_precision("bf16x3")
