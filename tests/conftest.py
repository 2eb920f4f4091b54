"""pytest configuration: registers the ``gpu`` marker and puts the repo root on sys.path."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import torch

    cache = {}

    def load(name):
        if name not in cache:
            cache[name] = torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)
        return cache[name]

    return load


def pytest_terminal_summary(terminalreporter):
    """element-wise relative errors recorded by tolerance.check_close (entries with |ref| >= 1e-3 of the tensor scale)"""
    try:
        import tolerance
    except ImportError:
        return
    rows = sorted(tolerance.ELEMENTWISE, reverse=True)[:12]
    if not rows:
        return
    terminalreporter.write_line("")
    terminalreporter.write_line("element-wise relative error of the %d checked tensors (entries >= 1e-3 of the tensor scale): worst 12"
                                % len(tolerance.ELEMENTWISE))
    terminalreporter.write_line("   max elementwise   99.9 %         norm-wise    tensor")
    for mx, q, nw, what in rows:
        terminalreporter.write_line("   %.2e          %.2e      %.2e     %s" % (mx, q, nw, what[:90]))
