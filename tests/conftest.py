import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    import pyoracle
    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="session")
def ffi():
    lib_path = os.path.join(ROOT, "gnss-sdr-rs_b200", "libgnss_b200.so")
    if not os.path.exists(lib_path):
        import __graft_entry__
        __graft_entry__.build()
    import gnss_sdr_rs_b200._ffi as f
    f.lib()
    return f


@pytest.fixture(scope="session")
def gpu(ffi):
    """One gb_handle on cuda:0.  The product has no CPU fallback: without a device this errors."""
    if ffi.lib().gb_device_count() < 1:
        pytest.fail("no CUDA device: -m gpu tests must run on the GPU box")
    hd = ffi.Handle(0)
    yield hd
    hd.close()
