import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")
    config.addinivalue_line("markers", "requires_reference: needs the reference tree at /root/reference")


def pytest_collection_modifyitems(config, items):
    from oracle import ref_harness
    have_ref = ref_harness.available()
    skip_ref = pytest.mark.skip(reason="reference tree not present on this host")
    for item in items:
        if "requires_reference" in item.keywords and not have_ref:
            item.add_marker(skip_ref)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name))
    return load


@pytest.fixture(scope="session")
def manifests():
    import json
    with open(os.path.join(GOLDEN, "manifests.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def cuda_lib():
    """Build (if needed) and load the C-ABI library; GPU tests go through it, never through the oracle."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from tair_b200 import _lib, build
    if not os.path.exists(_lib.LIB_PATH):
        build.build()
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return _lib.lib()
