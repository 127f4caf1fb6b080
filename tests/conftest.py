import os, sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "proof-of-burn_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: multi-GB witness; skipped unless POB_SLOW=1")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    skip_gpu = pytest.mark.skip(reason="no CUDA device")
    skip_slow = pytest.mark.skip(reason="multi-GB extra shape; set POB_SLOW=1 to run")
    for it in items:
        if "gpu" in it.keywords and not has_gpu:
            it.add_marker(skip_gpu)
        if "slow" in it.keywords and os.environ.get("POB_SLOW") != "1":
            it.add_marker(skip_slow)
