import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """Make sure the product library (and the thin torch extension) are built and current -- a no-op when the built
    files that travel with the tree are up to date.  A failure here is not swallowed: the product has no fallback."""
    from oneshotdet_b200 import build as osd_build

    osd_build.build()
    try:   # the thin torch extension is a second binding of the same kernels; its own tests skip when it is absent
        osd_build.build_torch_extension()
    except Exception as exc:  # noqa: BLE001
        import warnings

        warnings.warn(f"oneshotdet_b200._C_torch could not be built: {exc}")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def ref_ops():
    """The reference's own compiled CPU ops (oracle/_ref/osd_ref_C.so), or None when it is not there."""
    from oracle import build_ref

    return build_ref.load_ref()
