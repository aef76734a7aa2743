import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _ensure(path, make_dir, target=None):
    if not path.exists():
        cmd = ["make", "-C", str(make_dir)] + ([target] if target else [])
        subprocess.run(cmd, check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return path


@pytest.fixture(scope="session")
def oracle():
    import aadtest
    _ensure(ROOT / "oracle" / "liboracle.so", ROOT / "oracle", "liboracle.so")
    return aadtest.Oracle(ROOT / "oracle" / "liboracle.so")


def _ref(name):
    import aadtest
    from aad_b200.capi import AADCApi
    if Path("/root/reference/src").is_dir():
        _ensure(ROOT / "oracle" / "_ref" / name, ROOT / "oracle", "ref")
    p = ROOT / "oracle" / "_ref" / name
    if not p.exists():
        pytest.skip(f"{p} not built (the reference sources are not on this machine)")
    return AADCApi(p)


@pytest.fixture(scope="session")
def ref():
    """the unmodified reference codec compiled from /root/reference (2 channels max)"""
    return _ref("libaad_ref.so")


@pytest.fixture(scope="session")
def ref_wrapv():
    return _ref("libaad_ref_wrapv.so")


@pytest.fixture(scope="session")
def ref8():
    """reference with AAD_MAX_NUM_CHANNELS patched to 8"""
    return _ref("libaad_ref8.so")


@pytest.fixture(scope="session")
def product():
    """(AADCApi, GpuApi) over libaad_b200.so -- the thing under test"""
    import aad_b200
    _ensure(aad_b200.LIBRARY_PATH, ROOT / "aad_b200" / "csrc")
    return aad_b200.load()


@pytest.fixture(scope="session")
def gpu_ctx(product):
    api, gpu = product
    if gpu.device_count() < 1:
        pytest.fail("no CUDA device visible: GPU tests must run on the B200 box")
    h = gpu.create(0)
    yield h
    gpu.destroy(h)
