import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

REFERENCE_DIR = "/root/reference"   # exists only in the build container, never on the GPU box


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) GPU; run with -m gpu")


def have_reference():
    return os.path.exists(os.path.join(REFERENCE_DIR, "module", "unet.py"))


@pytest.fixture(scope="session")
def reference_model_cls():
    if not have_reference():
        pytest.skip("reference checkout not present on this machine")
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    from module.unet import Model
    return Model
