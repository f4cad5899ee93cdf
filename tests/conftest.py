import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA GPU (run with `pytest -m gpu` on a B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def native_lib():
    """The built native library for CPU-side tests of host-only entry points (op lists, CLI).  On a fresh
    checkout it is compiled first (nvcc cross-compiles without a GPU); if no toolchain is present the
    dependent tests are skipped -- on a GPU box a missing library is an error, not a skip."""
    from distkeras_b200 import _native

    try:
        return _native.lib()
    except RuntimeError:
        try:
            import build_native

            build_native.build(verbose=False)
            return _native.lib()
        except Exception as exc:
            import torch

            if torch.cuda.is_available():
                raise
            pytest.skip(f"native library unavailable on this CPU box: {exc}")
