import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    # the CUDA library is a build artefact (git-ignored): on a fresh checkout build it before the first
    # test needs it -- nvcc cross-compiles sm_100a without a GPU.  On the GPU box the prebuilt .so
    # travels with the snapshot and this is a no-op (the build is skipped when nothing is stale).
    from ldsr_b200 import _lib
    if not os.path.exists(_lib.SO_PATH):
        import shutil
        if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
            from ldsr_b200 import build as b
            b.build()


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
