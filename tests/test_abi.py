"""The C-ABI shared library loads without a GPU and exports every symbol include/eb200.h
declares; without a CUDA device it refuses to run (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT, have_gpu
from gdtk_b200 import _abi, Config, cases

HEADER = os.path.join(ROOT, "include", "eb200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(eb200_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for must in ("eb200_init", "eb200_block_create", "eb200_block_set_geometry", "eb200_block_set_bc",
                 "eb200_commit", "eb200_upload_flow", "eb200_download_flow", "eb200_compute_dt",
                 "eb200_step", "eb200_finalize", "eb200_last_error", "eb200_set_exchange"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(_abi.DEFAULT_LIBRARY)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} is declared in include/eb200.h but not exported"


def test_python_binding_covers_the_header():
    bound = {"eb200_" + n for n in _abi.SIGNATURES}
    assert set(declared_symbols()) == bound


def test_oracle_exports_the_same_abi(oracle):
    for name in _abi.SIGNATURES:
        assert hasattr(oracle, name)


@pytest.mark.skipif(have_gpu(), reason="only meaningful on a machine without a GPU")
def test_no_cpu_fallback():
    lib = _abi.load_library()
    cfg = Config(dimensions=2, flux_calculator="ausmdv").to_struct(cases.ideal_air())
    h = lib.init(C.byref(cfg))
    assert h < 0
    assert "no CPU fallback" in lib.error() or "CUDA" in lib.error()


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _abi.Library(str(tmp_path / "libeb200.so"), "eb200_")


def test_config_rejects_options_outside_the_path():
    gm = cases.ideal_air()
    Config().to_struct(gm)              # the reference default (adaptive_hanel_ausmdv) is on the path
    Config(flux_calculator="adaptive").to_struct(gm)    # = adaptive_efm_ausmdv
    with pytest.raises(ValueError):
        Config(flux_calculator="adaptive_hlle_ausmdv").to_struct(gm)
    with pytest.raises(ValueError):
        Config(shock_detector_smoothing=2).to_struct(gm)
    with pytest.raises(ValueError):
        Config(flux_calculator="ausmdv", viscous=True).to_struct(gm)
    with pytest.raises(ValueError):
        Config(flux_calculator="ausmdv", interpolation_order=3).to_struct(gm)
    with pytest.raises(AttributeError):
        Config(no_such_option=1)
