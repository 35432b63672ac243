import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _have_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu_ctx_factory():
    """Creates libqconsensus contexts on cuda:0; fails loudly (never skips) when marked gpu without a GPU."""
    from quoracle_b200 import native
    made = []

    def make(table=None, w=None):
        ctx = native.Context(0, table, w)
        made.append(ctx)
        return ctx

    yield make
    for c in made:
        c.close()


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a GPU must fail, not silently pass: nothing to do here; the tests
    # call into libqconsensus, which returns QC_ERR_NO_DEVICE.
    return
