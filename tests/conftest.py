import os
import sys

import numpy as np
import pytest

# The thread-rank tests run several "ranks" on ONE GPU with spinning wait kernels (tests/_thread_group.py): a lazy
# module load in the middle of a step would wait for the whole device, i.e. for a kernel that waits for this thread.
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
# ... and every rank thread brings 3-4 streams: with the default of 8 hardware queues, one rank's signal kernel can end
# up queued behind another rank's wait kernel (which waits for that signal).  One queue per stream avoids it.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, GOLDEN):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def bits(a):
    """Bit pattern view for exact comparison (NaN-safe, -0.0-safe)."""
    a = np.ascontiguousarray(a)
    return a.view({1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}[a.dtype.itemsize])


def assert_bits_equal(got, want, what=""):
    got, want = np.asarray(got), np.asarray(want)
    assert got.dtype == want.dtype, f"{what}: dtype {got.dtype} != {want.dtype}"
    assert got.shape == want.shape, f"{what}: shape {got.shape} != {want.shape}"
    bad = np.flatnonzero(bits(got).ravel() != bits(want).ravel())
    assert bad.size == 0, (f"{what}: {bad.size} of {want.size} elements differ bitwise; first at {bad[:5]}: "
                           f"got {got.ravel()[bad[:5]]} want {want.ravel()[bad[:5]]}")


@pytest.fixture(scope="session")
def golden_kernel():
    return np.load(os.path.join(GOLDEN, "kernel_cases.npz"))


@pytest.fixture(scope="session")
def golden_api():
    return np.load(os.path.join(GOLDEN, "api_cases.npz"))


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def golden_next():
    return np.load(os.path.join(GOLDEN, "next_cases.npz"))
