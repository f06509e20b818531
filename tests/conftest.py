import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "cs-304-speech-recognition-code_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_ready() -> bool:
    try:
        import torch
        return bool(torch.cuda.is_available())
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests are skipped (not failed) on a box without a CUDA device."""
    if _cuda_ready():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_hmm.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def golden_mfcc():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_mfcc.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def built_lib():
    sys.path.insert(0, PKG)
    import build as _b
    return _b.build()
