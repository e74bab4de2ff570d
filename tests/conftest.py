"""Shared fixtures.  Tests marked `gpu` need a B200 and call the product library through
the C ABI; everything else runs on CPU (oracle vs golden vectors, host logic, ABI symbols)."""
import ctypes as C
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from gdtk_b200 import _abi  # noqa: E402

GAS_DATA = os.path.join(ROOT, "tests", "golden", "gas")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def build_oracle():
    so = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
    src = os.path.join(ROOT, "oracle", "oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    return so


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle bound through the same host classes as the product (prefix orc_)."""
    lib = _abi.load_library(build_oracle(), "orc_")
    # oracle-only entry points used by the known-answer tests
    lib.gas_update = lib.cdll.orc_gas_update
    lib.gas_update.restype = C.c_int
    lib.gas_update.argtypes = [C.c_int, C.c_int, _abi.DP]
    lib.cea_eval = lib.cdll.orc_cea_eval
    lib.cea_eval.restype = C.c_int
    lib.cea_eval.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, _abi.DP]
    lib.face_flux = lib.cdll.orc_face_flux
    lib.face_flux.restype = C.c_int
    lib.face_flux.argtypes = [C.c_int, _abi.DP, _abi.DP, _abi.DP, _abi.DP, _abi.DP]
    lib.debug_face_flux = lib.cdll.orc_debug_face_flux
    lib.debug_face_flux.restype = C.c_int
    lib.debug_face_flux.argtypes = [C.c_int, C.c_int, _abi.DP, _abi.DP, _abi.DP, _abi.DP, C.POINTER(C.c_int)]
    lib.set_option = lib.cdll.orc_set_option
    lib.set_option.restype = C.c_int
    lib.set_option.argtypes = [C.c_int, C.c_char_p, C.c_int]
    return lib


def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def product():
    """The CUDA library; fails loudly if it is not built."""
    return _abi.load_library()
