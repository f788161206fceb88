"""pytest configuration: the `gpu` marker, and session fixtures for the oracle
builds, the CPU kernel emulator (test infrastructure) and the GPU context."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_c():
    """ctypes handle of the C++ restatement (oracle/ref_cpu.cpp), built on demand."""
    from tests import helpers
    return helpers.load_oracle_c()


@pytest.fixture(scope="session")
def emu_lib_path():
    """TEST-ONLY build of the .cu sources on the CPU fiber emulator."""
    import halo2_pse_b200  # noqa: F401  (registers the package)
    from halo2_pse_b200 import build
    return build.build_emulator()


@pytest.fixture(scope="session")
def emu_ctx(emu_lib_path):
    import halo2_pse_b200 as h
    ctx = h.Context(0, lib_path=emu_lib_path)
    yield ctx
    ctx.close()


@pytest.fixture(scope="session")
def gpu_ctx():
    """The product library on cuda:0.  Fails loudly (no fallback) if it cannot load."""
    import halo2_pse_b200 as h
    ctx = h.Context(0)
    yield ctx
    ctx.close()
