import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests are skipped on a box without a CUDA device (a plain `pytest` run here stays green) unless
    PPCSEQ_REQUIRE_GPU=1; on the B200 box they run and fail loudly if the library is missing."""
    if os.environ.get("PPCSEQ_REQUIRE_GPU") == "1":
        return
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device (set PPCSEQ_REQUIRE_GPU=1 to fail instead)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def built_lib():
    """The product library; GPU tests fail (not skip) if it is missing."""
    import ppcseq_b200
    return ppcseq_b200.lib()
