"""Shared fixtures.  `-m "not gpu"` tests run on a CPU-only box; `-m gpu` tests need a B200."""
from __future__ import annotations

import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "cuda-to-sycl-nbody_b200"), os.path.join(ROOT, "tools"), os.path.join(ROOT, "tests"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _make(target_dir, target, product):
    if not os.path.exists(product):
        subprocess.run(["make", "-C", target_dir, target], check=True, capture_output=True)
    return product


@pytest.fixture(scope="session")
def oracle():
    """CPU oracle (oracle/nbody_oracle.c) -- the checker, never the thing under test on GPU."""
    import oracle_lib
    _make(os.path.join(ROOT, "oracle"), "oracle", oracle_lib.ORACLE_LIB)
    return oracle_lib.Oracle()


@pytest.fixture(scope="session")
def nb():
    """ctypes binding of the product's C ABI; builds the library if it is missing."""
    import nbody_b200
    _make(os.path.join(ROOT, "cuda-to-sycl-nbody_b200"), "all", nbody_b200.LIB_PATH)
    nbody_b200.load_library()
    return nbody_b200


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference simulator (oracle/_ref), GPU only."""
    import refsim
    if not refsim.available():
        if os.path.isdir("/root/reference"):
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True, capture_output=True)
        else:
            pytest.skip("oracle/_ref/libnbody_ref.so not built and /root/reference absent")
    refsim.load()
    return refsim


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
