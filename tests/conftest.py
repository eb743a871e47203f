import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle.bindings import OracleLib, build
    build()
    return OracleLib()


@pytest.fixture(scope="session")
def ref_lib():
    from oracle.bindings import RefLib, have_ref
    if not have_ref():
        pytest.skip("oracle/_ref/libnip_ref.so not built (needs /root/reference)")
    return RefLib()


@pytest.fixture(scope="session")
def gpu_lib():
    import nip_b200.api as api
    lib = api.load_library()           # raises if the extension is missing: no fallback
    if lib.nipgpu_device_check(0) != 0:
        pytest.fail("no usable B200 device: " + lib.nipgpu_last_error().decode())
    return api
