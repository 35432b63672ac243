import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def pytest_sessionstart(session):
    """A checkout without the built library (the .so files are git-ignored) compiles it once, with nvcc, before the
    first test -- the same call __graft_entry__.build() makes.  Nothing here replaces the GPU path: without nvcc the
    build raises and every test that needs the library fails loudly."""
    from quoracle_b200 import build
    if not build.LIB.exists():
        build.build()


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


def pytest_terminal_summary(terminalreporter, exitstatus, config):
    """Margin-filter accounting of the GPU parity tests (tests/parity.py): printed, never dropped."""
    try:
        from tests import parity
    except Exception:
        return
    if not parity.LOG:
        return
    filt = [r for r in parity.LOG if r["filtered"]]
    unf = [r for r in parity.LOG if not r["filtered"]]
    tr = terminalreporter
    tr.write_sep("-", "index parity vs the f64 oracle (tests/parity.py)")
    tr.write_line(f"margin-safe inputs (|cos-thr| >= {parity.SAFE_MARGIN}, vote gap >= {parity.SAFE_MARGIN}): "
                  f"{sum(r['decisions'] for r in unf)} decisions in {len(unf)} batches, bit-identical indices on 100 %, no filter")
    tr.write_line(f"random inputs: {sum(r['decisions'] for r in filt)} decisions in {len(filt)} batches, bit-identical where the "
                  f"oracle margin exceeds the per-test margin; set aside by the filter: "
                  f"{sum(r['rejected_thr'] for r in filt)} (threshold rule) / {sum(r['rejected_vote'] for r in filt)} (vote)")
    worst = sorted(filt, key=lambda r: -(r["rejected_thr"] + r["rejected_vote"]))[:5]
    for r in worst:
        if r["rejected_thr"] + r["rejected_vote"]:
            tr.write_line(f"  {r['test']}: {r['rejected_thr']} / {r['rejected_vote']} of {r['decisions']} (margin {r['margin']:g})")
