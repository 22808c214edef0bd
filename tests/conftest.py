"""pytest plumbing: the `gpu` marker, module loading (the package directory has a hyphen), and
session fixtures for the three native libraries (product, host-emulation test build, oracles)."""
import ctypes as C
import importlib.util
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _load_pkg():
    if "vbkkt" in sys.modules:
        return sys.modules["vbkkt"]
    spec = importlib.util.spec_from_file_location(
        "vbkkt", ROOT / "linear-programming-vanderbei_b200" / "__init__.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules["vbkkt"] = mod
    spec.loader.exec_module(mod)
    return mod


def _load_build():
    spec = importlib.util.spec_from_file_location(
        "vbkkt_build", ROOT / "linear-programming-vanderbei_b200" / "build.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def vbkkt():
    return _load_pkg()


@pytest.fixture(scope="session")
def product_lib(vbkkt):
    """The nvcc-built product library.  Built in-tree if missing (needs nvcc, no GPU)."""
    if not vbkkt.LIB_PATH.exists():
        _load_build().build_product()
    return vbkkt.load()


@pytest.fixture(scope="session")
def emu_lib(vbkkt):
    """TEST build of the same sources on the host thread emulator (tests/emu)."""
    path = _load_build().build_emu()
    return vbkkt.load(path)


@pytest.fixture(scope="session")
def oracle_lib():
    """Plain-C restatement of the reference path (oracle/kkt_oracle.c) -- checker only."""
    subprocess.run(["make", "-C", str(ROOT / "oracle"), "restatement"], check=True,
                   stdout=subprocess.DEVNULL)
    import harness as H
    lib = C.CDLL(str(ROOT / "oracle" / "libkkt_oracle.so"))
    H.declare_oracle(lib)
    return lib


@pytest.fixture(scope="session")
def gpu_lib(product_lib):
    """Product library on a box with a CUDA device; fails (not skips) when the library is there
    but no device is: GPU tests must never pass on a fallback."""
    assert product_lib.vbk_device_count() > 0, "no CUDA device visible to libvbkkt.so"
    return product_lib
