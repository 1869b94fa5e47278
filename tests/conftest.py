import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _poison_cuda_allocator(request):
    """ARIA_TEST_POISON=1: before every GPU test, fill a spread of caching-allocator block sizes with NaN and free them, so a
    kernel that reads memory nobody wrote fails deterministically instead of depending on the allocator's history (how the
    round-1 LoRA flake was root-caused).  Off by default (costs ~1 ms per test)."""
    if os.environ.get("ARIA_TEST_POISON") == "1" and "gpu" in request.keywords:
        import torch

        if torch.cuda.is_available():
            junk = [torch.full((n,), float("nan"), dtype=torch.bfloat16, device="cuda")
                    for n in (512, 3000, 4096, 24576, 1 << 16, 98304, 1 << 18, 1 << 20, 1 << 22, 1 << 26)]
            torch.cuda.synchronize()
            del junk
    yield
