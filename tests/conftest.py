import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def data_dir():
    return os.path.join(ROOT, "tests", "data")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture
def as_written_arithmetic():
    """The oracle's K1 float sums without the FMA contractions of the reference's GPU build (every
    multiply/add rounded on its own): what the numpy restatement and the g++ build of the
    reference's kernel text compute.  The default is restored afterwards."""
    import oracle

    oracle.set_device_arith(False)
    yield
    oracle.set_device_arith(True)


def pytest_collection_modifyitems(config, items):
    """GPU tests are skipped (not failed) on a box without a CUDA device, so that a plain
    `pytest tests` works there too; on a GPU box a missing library is still a hard failure."""
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if not gpu_items:
        return
    try:
        from unified_cvo_b200 import _abi

        n = int(_abi.load_library().cvo_b200_device_count())
    except Exception:  # library missing: let the tests fail loudly where a GPU is expected
        return
    if n <= 0:
        skip = pytest.mark.skip(reason="no CUDA device visible")
        for it in gpu_items:
            it.add_marker(skip)
