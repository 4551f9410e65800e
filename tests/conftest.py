import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "exp-trmf-nips16_b200")
for p in (ROOT, PKG, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _have_gpu():
    try:
        import trmf
        return trmf.trmf._clib.clib_float32.trmf_b200_device_count() > 0
    except Exception:
        return False


def pytest_collection_finish(session):
    # GPU tests selected on a box without a GPU must fail loudly, not skip: the product has no CPU path.
    if any(item.get_closest_marker("gpu") is not None for item in session.items) and not _have_gpu():
        pytest.exit("tests marked `gpu` were selected but no CUDA device is visible: this library is GPU-only "
                    "(run them with gpurun, or deselect with -m 'not gpu')", returncode=2)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


def load_golden(name):
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    with np.load(path) as z:
        return {k: z[k] for k in z.files}
