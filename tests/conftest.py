"""pytest configuration.

Two backends run the same parity cases (tests/cases.py):
  * `-m gpu`     : the product, vorbispizza_b200/libvpz.so (sm_100a kernels) on a real B200.
  * `-m "not gpu"`: tests/emu/libvpz_emu.so -- the SAME host engine and the SAME kernel source compiled
                    against a CUDA execution-model emulator (test infrastructure, never shipped), so the
                    host logic and the kernel logic are exercised on a box without a GPU.
The checker in both cases is the CPU oracle (oracle/, test infrastructure).
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

DATA = os.path.join(ROOT, "tests", "data")
GOLDEN = os.path.join(ROOT, "tests", "golden")
FILES = ["1test", "2test", "3test", "issue6test"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


def load_file(name):
    with open(os.path.join(DATA, name + ".ogg"), "rb") as f:
        return f.read()


@pytest.fixture(scope="session")
def emu_lib_path():
    # VPZ_EMU_LIB: another build of the emulated library (e.g. one compiled with -fsanitize=address and run with
    # LD_PRELOAD=libasan.so: out-of-bounds accesses of the kernels show up as host heap errors)
    if os.environ.get("VPZ_EMU_LIB"):
        return os.environ["VPZ_EMU_LIB"]
    d = os.path.join(ROOT, "tests", "emu")
    subprocess.check_call(["make", "-C", d, "-s"])
    return os.path.join(d, "libvpz_emu.so")


@pytest.fixture(scope="session")
def emu_ctx(emu_lib_path):
    from vorbispizza_b200 import Context
    ctx = Context(0, lib_path=emu_lib_path)
    yield ctx
    ctx.close()


@pytest.fixture(scope="session")
def gpu_ctx():
    """The product library on cuda:0.  No skip, no fallback: a missing library or GPU is a failure."""
    from vorbispizza_b200 import Context
    ctx = Context(0)
    assert b"sm_100a" in ctx.lib.vpz_version()
    yield ctx
    ctx.close()
